// Device-side view of a Plan and the launcher interface (ps_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ps_plan.h"

namespace ps {

#ifndef PS_TPB
#define PS_TPB 128
#endif
constexpr int TPB = PS_TPB;         // threads per CTA == branches per tile
#ifndef PS_BUS_TPB
#define PS_BUS_TPB 128
#endif
constexpr int BTPB = PS_BUS_TPB;    // threads per bus-tile CTA (at most BTPB buses per tile)
constexpr int NSCAL = 4;            // PS_NSCALARS

struct DevPlan {                    // passed to kernels by value (constant bank)
  int form, src, flags;
  int nb, nr, ng;
  int o_a, o_b, o_Wd, o_Wr, o_Wi, o_pg, o_qg, o_f[4], o_cg, o_bss;
  int bus_rows;
  int ntiles, n_bus_tiles;
  int j_br0, h_br0;                                 // first J / H slot of the branch rows
  int pb_slots;                                     // PB* bus block: slot-parallel fast path available
  int cb_slots;                                     // CB* bus block: slot-parallel fast path available (same tables, see ps_capi.cu)
  const double *pb_jt, *pb_ht;                      // PB* slot tables (ps_capi.cu: upload): coefficients and
  const int32_t *pb_ji, *pb_hi;                     // indices, 4 shifted copies each:
  int pb_jstride, pb_hstride;                       // copy a at + a * stride + (4 - a) % 4 (aligned at entry a + 4q)
  int pbs_prefer;                                   // the plan's choice between the two PB* residual kernels
  int pbs_ok;                                       // PB* scenario-resident residual kernel available (ps_plan.h: pbs_code)
  const int32_t* pbs_code;                          // [ng | nr | nr] bus | rank << 24, -1 = no term
  const uint8_t* pbs_wmax;                          // highest rank per step (generator, Sij, Sji steps)
  int cost_order;
  const int32_t *f, *t;
  const double *Gf, *Bf, *Gt, *Bt, *gf, *bf, *gt, *bt, *Imax, *Gl, *Bl, *Pd0, *Qd0;
  const double *c0, *c1, *c2, *c3;                  // [2*nr] derived per-side constants
  const double *k0, *k1, *k2;                       // [ng] cost coefficients (may be null)
  const int32_t *gen_ptr, *gen_idx, *sh_ptr, *sh_idx;   // CSR lists per bus (Sij / Sji live in the half-edge tables)
  const int32_t *bus_jpp, *bus_hpp;                 // per bus: first J / H contribution (emission order)
  const int32_t *he_bus;                            // half-edge -> bus
  const int32_t *jsrc, *hsrc, *ms_dst, *ms_ptr, *ms_list;   // slot -> contribution maps of the bus block (ps_plan.h)
  const int32_t *he_ptr, *he_ks, *he_other;         // half-edges in bus order (Sij then Sji per bus)
  const double* he_c;                               // [4][nhe] PN* admittance entries staged per tile
  int nhe;
  const Tile* tiles;
  const int32_t* tile_map;                          // bus_kernel: CTA -> tile (nullptr: identity); ps_plan.h: hub_tiles
  const int32_t* hub_tiles; int n_hub_tiles;
  // warp tiles of bus_warp_kernel (ps_plan.h: WTile); n_wtiles == 0 -> the CTA-tile bus_kernel runs
  const WTile* wtiles;
  int n_wtiles, wt_jspace, wt_hspace, wt_gspace, wt_lists;
  const uint16_t *wt_dst, *wt_ownj, *wt_ownh, *wt_owng;
  const uint64_t* wt_multi;
  const int32_t *wt_ownj_ptr, *wt_ownh_ptr, *wt_owng_ptr;
};

struct EvalArgs {
  int what, S, chunk;
  const double* x;  long long ldx;
  const double* mu; long long ldmu;
  const double* Pd; const double* Qd; long long ldl;   // per-scenario loads (nullptr -> network loads)
  const uint8_t* status; long long ldst;               // per-scenario branch status (nullptr -> all in service)
  double* g; long long ldg;
  double* J; long long ldJ;
  double* H; long long ldH;
  double* scal;                                        // [S][NSCAL] (nullptr -> skip); must be zeroed
  double* diag; long long lddiag;                      // CB* fast path: scratch [S][4*nb] own-voltage Jacobian sums (nullptr -> generic path)
  double viol_tol;
  int force_any;                                       // development: bus_warp_kernel takes its run-time-flag instantiation (PS_WT_ANY)
};

// Launches one evaluation (bus kernel + branch kernel) on `stream`; *launches is incremented per kernel.
// kernel ids for the optional per-kernel timing: events ev[2*id], ev[2*id+1] are recorded around kernel `id`
enum KernelId : int { K_BUS = 0, K_BUS_JH = 1, K_OBJ = 2, K_BRANCH = 3, K_COUNT = 4 };
cudaError_t launch_eval(const DevPlan& P, const EvalArgs& A, int64_t smem_branch_bytes, int64_t smem_bus_bytes,
                        cudaStream_t stream, int64_t* launches, cudaEvent_t* ev = nullptr, unsigned* ev_mask = nullptr);
// opt the kernels of one formulation into the needed dynamic shared memory (once per device)
cudaError_t configure_eval(int form, int src, int64_t smem_branch_bytes, int64_t smem_bus_bytes, int64_t smem_wt_bytes = 0,
                           int64_t smem_pbs_bytes = 0);
// dynamic shared memory of pb_bus_g_scn_kernel (the bus accumulators of one scenario); 0 when it cannot run
int64_t pbs_smem_bytes(const DevPlan& P);
// dynamic shared memory of one bus_warp_kernel CTA (0 when the plan has no warp tiles)
int64_t wt_smem_bytes(const DevPlan& P);

// out[s][i] = in[s][perm[i]] (a caller-imposed Hessian slot order, ps_set_hessian_order)
cudaError_t launch_permute_rows(const int32_t* perm, int n, int64_t S, const double* in, long long ldin, double* out,
                                long long ldout, cudaStream_t stream);

// COO -> CSC assembly (replaces src/opt/algorithms/common.jl:12-20): nz[j] = sum_{q in seg j} dE[perm[q]]
cudaError_t launch_csc_assemble(const int32_t* segptr, const int32_t* perm, int nnz_csc, int64_t S,
                                const double* dE, long long ldE, double* nz, long long ldnz, cudaStream_t stream);

// affine + quadratic rows (replaces src/opt/MOI_wrapper.jl:864-891, 973-1002, 1030-1042)
struct RowsDev {
  int m, nnzj, nnzh;
  const int32_t *aptr, *avar, *qptr, *qv1, *qv2;   // CSR terms, 0-based variables
  const double *acoef, *qcoef, *cst;
  const int32_t *jx;  const double* jc;            // J entry e = jx[e] < 0 ? jc[e] : jc[e] * x[jx[e]]
  const int32_t *hrow; const double* hc;           // H entry e = hc[e] * lambda[hrow[e]]
};
cudaError_t launch_rows_eval(const RowsDev& R, int what, int64_t S, const double* x, long long ldx, const double* lam,
                             long long ldl, double* g, long long ldg, double* J, long long ldJ, double* H, long long ldH,
                             cudaStream_t stream);

// KKT / merit metrics (replaces src/opt/algorithms/common.jl:35-98, slp.jl:108-162 at alpha = 0)
constexpr int NKKT = 8;
struct KktDev {
  int m, n;
  const double *gL, *gU, *xL, *xU;
  const int32_t *colptr, *rowval;      // CSC pattern, 0-based
  const int32_t *rptr, *rperm;         // the same entries grouped by row: row i owns rperm[rptr[i] .. rptr[i+1])
};
struct KktArgs {
  const double *E, *x, *lam, *mU, *mL, *df, *nz, *nu, *p, *f;
  long long ldE, ldx, ldl, ldm, lddf, ldnz, ldnu, ldp;
  double* out;
};
cudaError_t launch_kkt(const KktDev& K, const KktArgs& A, int64_t S, cudaStream_t stream);

}  // namespace ps
