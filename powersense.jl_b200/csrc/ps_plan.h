// Host-side plan: variable layout, NL row order, JuMP-compatible Jacobian / Hessian structure and
// the slot tables the kernels consume.  Built once per (network, formulation, source, flags).
// Restates what JuMP derives from the @NL* calls of src/opf/formulations.jl:170-428
// (structure rules: SURVEY.md Appendix C.2 / C.3).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ps_patterns.h"

namespace ps {

struct HostNet {             // deep copy of ps_network (include/powersense_b200.h)
  int64_t nbus = 0, nbr = 0, ngen = 0, nbss = 0;
  std::vector<int32_t> f, t;                 // 0-based
  std::vector<double> Gf, Bf, Gt, Bt, gf, bf, gt, bt, Imax, Pd, Qd, Gl, Bl;
  std::vector<int32_t> gen_ptr, gen_idx, sij_ptr, sij_idx, sji_ptr, sji_idx, sh_ptr, sh_idx;  // 0-based ids
  int64_t n_tgh = 0;
  std::vector<double> c0, c1, c2;            // objective (formulations.jl:128-138); may be empty
  int cost_order = 0;
};

struct Tile {                // one CTA work item: a contiguous range of buses or branches
  int32_t kind;              // 0 = bus rows, 1 = branch rows, 2 = objective scalar
  int32_t u0, u1;            // unit range
  int32_t g0, gn, j0, jn, h0, hn;  // output slot ranges (per scenario)
  // bus tiles: slots with more than one contribution ("multi-slots"), sorted by contribution count (descending):
  // descriptors [m0, m0+mn) in Plan::ms_dst / ms_ptr, extra-contribution lists from l0 (ln entries) in Plan::ms_list
  int32_t mj0, mjn, mjl0, mjln, mh0, mhn, mhl0, mhln;
};

// One warp's work item of bus_warp_kernel: a contiguous range of buses with at most 32 half-edges.  The warp keeps a
// shared-memory *image* of the tile's J block, H block and residual rows (exactly the bytes that go to HBM), each
// followed by extra cells for the second and later contributions of multi-contribution slots ("J / H / G space").
// The extra cells of one slot are contiguous and in emission order.
struct WTile {
  int32_t u0, u1;            // buses
  int32_t j0, jn, h0, hn;    // output slot ranges (per scenario); the residual rows are 2*u0 .. 2*u1
  int32_t he0, nhe;          // half-edges (global, bus order)
  int32_t jx, hx, gx;        // extra cells of the J / H / G space
  int32_t dst0;              // Plan::wt_dst: [he_nj + he_nh + he_npg][32] destinations of lane's contributions (type-major)
  int32_t m0, mn;            // multi-slots of all three spaces, sorted by count (descending): Plan::wt_multi
};

#ifndef PS_PBS_E
#define PS_PBS_E 1
#endif
constexpr int PBS_THREADS = 1024;                 // CTA size of the scenario-resident PB* residual kernel
constexpr int PBS_WAVE = PBS_THREADS * PS_PBS_E;  // elements per step: PS_PBS_E per thread

struct Plan {
  int form = 0, src = 0, flags = 0;
  HostNet net;
  int64_t n_var = 0, m_nl = 0, nnzJ = 0, nnzH = 0;
  // 0-based variable offsets; -1 when the block does not exist (formulations.jl:20-67, :121)
  int64_t o_a = 0, o_b = 0, o_Wd = -1, o_Wr = -1, o_Wi = -1, o_pg = 0, o_qg = 0, o_f[4] = {-1, -1, -1, -1},
          o_cg = -1, o_bss = -1, o_tgh = -1;
  int64_t bus_rows = 0;      // 2*nbus or 0
  int rows_per_branch = 0, RJ = 0, RH = 0;          // per-branch block sizes
  std::vector<int64_t> jrow, jcol, hrow, hcol;      // 1-based structure (MOI tuples)
  std::vector<int32_t> jptr, hptr;                  // row -> first slot
  std::vector<double> lo, hi;                       // NLPBoundsPair per row
  // branch position tables: [swap][tag] -> slot inside the branch's block (swap = f > t)
  uint8_t br_jpos[2][MAX_BR_JTAGS] = {}, br_hpos[2][MAX_BR_HTAGS] = {};
  // bus position streams (emission order of the bus kernel); bit 15 = accumulate into the slot
  std::vector<uint16_t> bus_jpos, bus_hpos;
  std::vector<int32_t> bus_jpp, bus_hpp;            // per-bus start in the streams (size nbus+1)
  // half-edges in the bus kernel's order: for bus i, Sij(i) ascending then Sji(i) ascending
  std::vector<int32_t> he_ptr;                      // [nbus+1]
  std::vector<int32_t> he_ks;                       // branch * 2 + side (side 0: bus is the from end)
  std::vector<int32_t> he_other;                    // the bus at the other end
  std::vector<int32_t> he_bus;                      // the bus a half-edge belongs to
  // slot -> contributions of the bus block (contributions in the bus kernel's emission order, see bus_jpos):
  // slot e of J sums CJ[jperm[jsp[e] .. jsp[e+1])] left to right (emission order); same for H
  std::vector<int32_t> jsp, jperm, hsp, hperm;
  // per bus tile (see Tile): jsrc / hsrc[e] = first contribution of slot e (global contribution index);
  // ms_dst = tile-relative cell that receives the sum, ms_ptr = tile-relative [begin, end) into ms_list,
  // ms_list = tile-relative contribution indices added in emission order
  std::vector<int32_t> jsrc, hsrc, ms_dst, ms_ptr, ms_list;
  std::vector<double> he_c;                         // [4][nhe] staged constants of the PN* balances (g_s, b_s, c, d)
  // PB* balances without switched shunts: per-slot codes of the bus block for the slot-parallel fast path
  //   J: -1 -> 1.0 ; k >= 0 -> status of branch k ; -2 - (4*bus + w) -> voltage slot w (see pb_bus_jh_kernel)
  //   H: 2*bus + (0: P row, 1: Q row)
  std::vector<int32_t> pb_jcode, pb_hcode;
  bool pb_slots = false;
  // PB* balances, scenario-resident residual kernel (pb_bus_g_scn_kernel): the terms of a balance are visited in
  // three element-ordered passes (generators, Sij side of every branch, Sji side of every branch) of PBS_WAVE
  // consecutive elements per step, so that the injections / flow variables are read with coalesced loads.
  //   pbs_code [ngen | nbr | nbr]: bus | rank << 24 of the element's term, -1 when the element is in no list;
  //            rank = number of earlier elements of the same step that belong to the same bus (a bus's terms are
  //            added in list order: ranks are served in sub-rounds)
  //   pbs_wmax per step (generator steps, Sij steps, Sji steps): highest rank of the step
  // pbs_ok: every adjacency list is ascending and ranks fit (else the list-walking pb_bus_g_kernel runs)
  std::vector<int32_t> pbs_code;
  std::vector<uint8_t> pbs_wmax;
  bool pbs_ok = false;
  bool pbs_prefer = false;                          // ... and few enough ranks per step for it to be the faster kernel
  bool pb_implicit = false;                         // PB* bus rows: J slots follow [voltage][gens][Sij][Sji]
  int64_t smem_bytes = 0;                           // dynamic shared memory of the kernel (max over tile kinds)
  int64_t smem_bus_bytes = 0, smem_branch_bytes = 0; // bus kernel / staged branch kernel
  // derived per-side constants (side 0 = from, 1 = to), each [2*nbr]
  std::vector<double> c0, c1, c2, c3;
  std::vector<Tile> tiles;
  // warp tiles of bus_warp_kernel.  A bus with more than 32 half-edges does not fit a warp: the CTA tiles (`tiles`) that
  // contain such a bus are listed in hub_tiles and run through bus_kernel; the warp tiles cover all other buses
  std::vector<WTile> wtiles;
  std::vector<int32_t> hub_tiles;
  std::vector<uint16_t> wt_dst;
  // multi-slot descriptor: bits 0-13 destination slot, 14-15 space (0 J, 1 H, 2 G), 16-31 first extra cell, 32-47 count
  std::vector<uint64_t> wt_multi;
  // destinations of the buses' own contributions: J (generators, shunt), H (shunt), G (generator injections (P, Q) per
  // generator, then P shunt, Q shunt, Q switched-shunt terms, -Pd, -Qd)
  std::vector<uint16_t> wt_ownj, wt_ownh, wt_owng;
  std::vector<int32_t> wt_ownj_ptr, wt_ownh_ptr, wt_owng_ptr;    // per bus [nbus+1]
  int32_t wt_lists_max = 0;                         // uint16 entries of the per-warp staged lists (max over tiles)
  int32_t wt_jspace = 0, wt_hspace = 0, wt_gspace = 0;   // doubles of the spaces incl. alignment slack (max over tiles)
  int n_bus_tiles = 0;
  int tile_units_bus = 0, tile_units_br = 0;
  int64_t smem_doubles = 0;                         // max over tiles of staged outputs (padded)
  int NRP = 0, RJP = 0, RHP = 0;                    // per-thread staging strides of a branch tile (odd)
  int64_t nx_branch = 0;                            // variables the branch rows touch
  int64_t nx_nl = 0;                                // variables the NL rows touch (roofline accounting)
};

// returns "" on success, else an error message
std::string build_plan(const HostNet& net, int form, int src, int flags, int tpb, int bus_tpb, Plan& out);

}  // namespace ps
