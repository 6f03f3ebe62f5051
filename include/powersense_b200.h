/* powersense_b200.h -- C ABI of the B200-native evaluator for the Powersense.jl AC-OPF hot path.
 *
 * The reference (PowerSense/Powersense.jl v0.0.6) has no FFI: its hot path sits behind
 * MathOptInterface's abstract NLP evaluator, whose complete method set is enumerated by the
 * reference's own fake, `EmptyNLPEvaluator` (src/opt/MOI_wrapper.jl:56-77).  Every entry point
 * below replaces one of those methods (or the data / assembly step either side of them) and
 * cites the reference interface it stands in for.  A Julia `ccall` shim that binds them is in
 * INTEGRATION.md (julia/GpuNLPEvaluator.jl); powersense.jl_b200/capi.py is the ctypes binding
 * used by the tests.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in the signatures (`void* stream` is a cudaStream_t or NULL);
 *   - every function returns PS_OK (0) or a non-zero error code; ps_last_error() gives the text
 *     (the Julia shim turns non-zero into `error(...)`, the only error channel the MOI boundary
 *     has: src/opt/MOI_wrapper.jl:526);
 *   - all index arrays crossing the ABI are int64 and 1-based exactly like the MOI tuples
 *     (`Vector{Tuple{Int64,Int64}}`, src/opt/MOI_wrapper.jl:813,855);
 *   - the caller allocates every output; the evaluator only fills (same as MOI);
 *   - a handle is single-threaded (JuMP's evaluator is stateful and not thread-safe either);
 *   - there is NO CPU fallback: ps_create fails with PS_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef POWERSENSE_B200_H
#define POWERSENSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_ABI_VERSION 2
/* `device` argument of ps_create: CUDA ordinal, PS_DEVICE_CURRENT, or PS_DEVICE_NONE for a structure-only
 * handle (ps_dims / ps_constraint_bounds / ps_*_structure / ps_variable_offsets work; every evaluation entry
 * point returns PS_ERR_STATE -- there is no CPU evaluation path) */
#define PS_DEVICE_CURRENT (-1)
#define PS_DEVICE_NONE (-2)

/* error codes */
enum {
  PS_OK = 0,
  PS_ERR_ARG = 1,      /* bad argument (null pointer, size, code out of range) */
  PS_ERR_NETWORK = 2,  /* inconsistent network tables (endpoint out of range, self loop, adjacency) */
  PS_ERR_CUDA = 3,     /* CUDA runtime failure / no usable device (there is no CPU fallback) */
  PS_ERR_STATE = 4     /* call sequence error (e.g. Hessian requested but not initialised) */
};

/* formulation codes: the nine singletons of src/opf/type.jl:7-15, in that order */
enum {
  PS_PBRAPV = 0, PS_PBRARV = 1, PS_PBRAWV = 2, PS_CBRARV = 3, PS_CBRAWV = 4,
  PS_PNPAPV = 5, PS_PNRAPV = 6, PS_PNRARV = 7, PS_PNRAWV = 8
};
/* data.source_type selects the branch-limit family (src/opf/formulations.jl:299 vs :363) */
enum { PS_SRC_MATPOWER = 0, PS_SRC_ARPAE = 1 };
/* run_opf! keyword flags that change the variable layout (src/opf/formulations.jl:60-67,121) */
enum { PS_FLAG_FACTS_BSHUNT = 1, PS_FLAG_OBJ_LINEAR = 2 };
/* `what` bitmask of the batched evaluator */
enum { PS_EVAL_G = 1, PS_EVAL_J = 2, PS_EVAL_H = 4 };
/* per-scenario scalars written by ps_batch_eval (PS_NSCALARS doubles per scenario):
 *   [0] objective value (src/opf/formulations.jl:122-138),
 *   [1] max violation of the NL rows = norm_violations(E, g_L, g_U, ..., Inf) restricted to the NLP
 *       block (src/opt/algorithms/common.jl:76-98),
 *   [2] number of NL rows violated by more than `viol_tol`,
 *   [3] reserved (0). */
#define PS_NSCALARS 4

/* The PowersenseData tables the NL rows read, as struct-of-arrays host pointers
 * (replaces `mutable struct PowersenseData`, src/opf/PowersenseData.jl:4-105; the fields filled at
 * :186-211).  Copied at ps_create. */
typedef struct ps_network {
  int64_t nbus, nbr, ngen, nbss;
  int64_t n_tgh;                     /* sum(data.lps): number of tgh variables (formulations.jl:121) */
  const int64_t *br_f, *br_t;        /* [nbr] data.br[k] = (f, t), 1-based (PowersenseData.jl:210) */
  const double *Gf, *Bf, *Gt, *Bt;   /* [nbr] PowersenseData.jl:198-201 */
  const double *gf, *bf, *gt, *bt;   /* [nbr] PowersenseData.jl:202-205 */
  const double *Imax;                /* [nbr] PowersenseData.jl:206-207 */
  const double *Pd, *Qd, *Gl, *Bl;   /* [nbus] PowersenseData.jl:186-189 */
  /* data.node[string(i)] = {gen, Sij, Sji, shunt} (PowersenseData.jl:212) as CSR lists:
   * ptr[nbus+1] 0-based offsets, idx = 1-based element ids in ascending order */
  const int64_t *gen_ptr, *gen_idx, *sij_ptr, *sij_idx, *sji_ptr, *sji_idx, *sh_ptr, *sh_idx;
  /* objective data (formulations.jl:128-138); may be NULL when cost_order == 0 */
  const double *c0, *c1, *c2;        /* [ngen] */
  int32_t cost_order;                /* data.cost_order: 3 quadratic, 2 linear, else constant */
} ps_network;

typedef struct ps_handle ps_handle;  /* one (network, formulation, source_type, flags) model   */
typedef struct ps_batch ps_batch;    /* S scenarios of that model resident on one device       */
typedef struct ps_csc ps_csc;        /* a fixed COO -> CSC assembly pattern                    */

/* ---- lifetime ---------------------------------------------------------------------------- */
/* Replaces the JuMP model build of the NL block (create_opf_model!, src/opf/formulations.jl:170-428)
 * plus MOI.initialize (src/opt/MOI_wrapper.jl:1117): derives the variable layout, the NL row order,
 * both sparsity structures and the device tables.  `device` is the CUDA ordinal (see PS_DEVICE_*). */
int ps_create(const ps_network* net, int formulation, int source_type, int flags, int device, ps_handle** out);
void ps_destroy(ps_handle* h);
/* text of the last error on this handle (h == NULL: last ps_create failure) */
const char* ps_last_error(const ps_handle* h);
int ps_abi_version(void);

/* ---- structure (one-off) ------------------------------------------------------------------ */
/* n_var = all model variables (incl. cg / tgh the NL rows never read), m_nl = NL rows,
 * nnz of MOI.jacobian_structure / MOI.hessian_lagrangian_structure */
int ps_dims(const ps_handle* h, int64_t* n_var, int64_t* m_nl, int64_t* nnz_jac, int64_t* nnz_hess);
/* MOI.NLPBlockData.constraint_bounds (src/opt/MOI_wrapper.jl:1091-1094): (0,0) for ==, (-inf,0) for <= */
int ps_constraint_bounds(const ps_handle* h, double* lo, double* hi);
/* MOI.jacobian_structure (src/opt/MOI_wrapper.jl:813): rows in @NLconstraint order, columns ascending */
int ps_jacobian_structure(const ps_handle* h, int64_t* row, int64_t* col);
/* MOI.hessian_lagrangian_structure (src/opt/MOI_wrapper.jl:855): per-row blocks, lower-triangular tuples */
int ps_hessian_structure(const ps_handle* h, int64_t* row, int64_t* col);
/* ---- Hessian slot order -------------------------------------------------------------------------------------------
 * JuMP lays every NL row's Hessian block out as [diagonal slots of the row's variables, ascending][off-diagonal slots in
 * the order its acyclic-colouring recovery produces]; the second part depends on Julia's Set hash order and
 * cannot be derived without running Julia.  ps_hessian_structure therefore reports a documented canonical order (diagonal
 * slots ascending, then lower-triangular pairs (max, min) ascending) -- and these entry points let a caller that HAS the
 * JuMP order (julia/make_golden.jl dumps it; julia/GpuNLPEvaluator.jl can fetch it from a throw-away JuMP evaluator)
 * install it, so that structure and values cross the ABI in exactly JuMP's order.  The kernels are untouched: they keep
 * writing the canonical order, a gather pass (device) applies the permutation to H before it leaves the library.
 *
 * ps_hessian_row_blocks: ptr[r] .. ptr[r+1] (0-based, m_nl + 1 entries) = the slots of NL row r in the canonical structure.
 * ps_set_hessian_order: perm[i] (1-based, a permutation of 1..nnz_hess) = canonical position of the entry reported at
 *   position i from now on, by ps_hessian_structure and by every evaluation entry point; NULL restores the canonical order.
 * ps_match_hessian_structure: derive that permutation from a target structure (1-based (var, var) tuples, either triangle),
 *   assumed -- like JuMP's -- to be the concatenation of the same per-row blocks in row order: inside every row block the
 *   target must hold the same set of variable pairs; fails with PS_ERR_ARG naming the first row whose block differs.
 * ps_hessian_order: the permutation in force (identity when none). */
int ps_hessian_row_blocks(const ps_handle* h, int64_t* ptr);
int ps_set_hessian_order(ps_handle* h, const int64_t* perm);
int ps_match_hessian_structure(ps_handle* h, const int64_t* row, const int64_t* col);
int ps_hessian_order(const ps_handle* h, int64_t* perm);

/* 0-based offsets of the variable blocks (formulations.jl:20-67): out[0..15] =
 * a(theta|vi), b(V|vr), Wd, Wr, Wi, pg, qg, F0, F1, F2, F3, cg, bss, tgh, n_var, nx_nl ; -1 = absent */
int ps_variable_offsets(const ps_handle* h, int64_t out[16]);

/* Execution plan facts for reports and tests: out[0..7] = number of bus tiles, number of branch tiles, dynamic shared
 * memory of the bus kernel (bytes), of the staged branch kernel (bytes), 1 if the branch rows can take the direct
 * (register -> 32-byte sector) store path, 1 if the PB* bus rows use the implicit slot order, J / H slot where the
 * branch rows start (0-based; the direct path needs these on a 32-byte sector: see ps_preferred_layout). */
int ps_plan_info(const ps_handle* h, int64_t out[8]);
/* Leading dimension (doubles, multiple of 4) and leading pad (doubles, 0..3) for the batched g / J / H arrays
 * (index 0, 1, 2) such that, on a 32-byte aligned allocation, row s starts at base + pad + s*ld and the branch block
 * of every row starts on a 32-byte sector -- the condition for the direct store path of ps_batch_eval.  Other
 * layouts are accepted and take the staged path. */
int ps_preferred_layout(const ps_handle* h, int64_t ld[3], int64_t pad[3]);

/* ---- single-scenario callbacks: host pointers, synchronous --------------------------------
 * Like JuMP's evaluator (which keeps its forward pass while x is unchanged) the handle remembers the x of
 * the last call: a callback at a new x evaluates g AND J in one launch (when the caller's previous point asked for both),
 * keeps them on the device, and the sibling callback at the same x is served from there without a launch or an upload of
 * x.  eval_hessian needs the multipliers, which arrive with it: one launch of its own (x is not uploaded again).
 * ps_launch_count reports the kernel launches issued by these callbacks so far. */
/* MOI.eval_constraint(d, g, x) (call site src/opt/MOI_wrapper.jl:969).  `g` may point into a larger
 * array (the wrapper passes view(g, row:length(g)), :967). */
int ps_eval_constraint(ps_handle* h, const double* x, double* g);
/* MOI.eval_constraint_jacobian(d, J, x) (src/opt/MOI_wrapper.jl:1026; J = view(values, 1+offset:end), :1025) */
int ps_eval_jacobian(ps_handle* h, const double* x, double* J);
/* MOI.eval_hessian_lagrangian(d, H, x, sigma, mu) (src/opt/MOI_wrapper.jl:1061).  sigma multiplies an NL
 * objective only; Powersense OPF models have none (has_objective = false), so it is accepted and unused. */
int ps_eval_hessian(ps_handle* h, const double* x, double sigma, const double* mu, double* H);
/* all three in one launch (JuMP reuses its forward pass when x is unchanged; this is the fused equivalent);
 * any of g / J / H may be NULL */
int ps_eval_all(ps_handle* h, const double* x, const double* mu, double* g, double* J, double* H);
/* MOI.eval_objective / eval_objective_gradient exist on the evaluator but are never called for OPF models
 * (has_objective = false, src/opt/MOI_wrapper.jl:897,938); the objective of formulations.jl:122-138 is
 * provided for the batched driver's scalars. */
int ps_eval_objective(ps_handle* h, const double* x, double* f);
int64_t ps_launch_count(const ps_handle* h);

/* ---- batched, device-resident evaluation (the roofline path) ------------------------------ */
/* Per-scenario data lives on `device` (defaults: the network's own loads, every branch in service).
 * Lifetime: a ps_batch must be destroyed before its ps_handle, a ps_kkt may outlive its ps_csc (it copies what it needs).
 * Concurrency: calls on one batch may use different streams as long as their scenario ranges do not overlap (the CB*
 * models keep a per-scenario scratch between their two bus kernels); one in-flight evaluation per scenario range. */
int ps_batch_create(ps_handle* h, int64_t S, ps_batch** out);
void ps_batch_destroy(ps_batch* b);
/* Pd, Qd: host [S][nbus] row-major (scenario loads: the only per-scenario constants of the NL rows) */
int ps_batch_set_loads(ps_batch* b, const double* Pd, const double* Qd);
/* st: host [S][nbr] (1 in service, 0 out): contingency scenarios on the fixed sparsity pattern.  The
 * branch tables are multiplied by st exactly like PowersenseData.jl:198-207 (Imax: 0 -> 10000).  NULL resets. */
int ps_batch_set_branch_status(ps_batch* b, const uint8_t* st);
int ps_batch_set_viol_tol(ps_batch* b, double tol);
/* Evaluate scenarios [s_begin, s_begin + s_count) of the batch.  All data pointers are DEVICE pointers;
 * row s of an array starts at ptr + s*ld (ld in doubles, s counted from 0 at s_begin).  mu may be NULL
 * unless PS_EVAL_H; outputs not selected by `what` may be NULL; scalars [s_count][PS_NSCALARS] may be NULL.
 * Asynchronous on `stream`. */
int ps_batch_eval(ps_batch* b, int what, int64_t s_begin, int64_t s_count,
                  const double* x, int64_t ldx, const double* mu, int64_t ldmu,
                  double* g, int64_t ldg, double* J, int64_t ldJ, double* H, int64_t ldH,
                  double* scalars, void* stream);
/* Same, HOST buffers (ideally pinned: ps_host_alloc): scenarios are streamed through device staging
 * buffers in chunks, H2D copies / kernels / D2H copies overlapped on three streams.  Synchronous. */
int ps_batch_eval_host(ps_batch* b, int what, int64_t s_begin, int64_t s_count,
                       const double* x, int64_t ldx, const double* mu, int64_t ldmu,
                       double* g, int64_t ldg, double* J, int64_t ldJ, double* H, int64_t ldH,
                       double* scalars);
/* number of kernel launches issued by this batch so far (bench.py reports it as gpu_launches) */
int64_t ps_batch_launch_count(const ps_batch* b);
/* algorithmic HBM bytes of one scenario evaluation for `what` */
int64_t ps_algorithmic_bytes(const ps_handle* h, int what);
/* the share of it that belongs to the branch-row kernel (flow definitions, limits, cones): the x entries and
 * multipliers those rows read plus their g / J / H slots */
int64_t ps_algorithmic_bytes_branch(const ps_handle* h, int what);
/* Optional per-kernel timing: when enabled, ps_batch_eval records CUDA events around each kernel it launches;
 * ps_batch_kernel_times waits for the last evaluation and returns ms[0..3] = bus rows (g, or the whole bus
 * kernel), bus J/H slots (PB* fast path), objective scalar, branch rows (0 when a kernel did not run). */
int ps_batch_enable_timing(ps_batch* b, int on);
int ps_batch_kernel_times(ps_batch* b, float ms[4]);

/* pinned host memory for the *_host entry points */
int ps_host_alloc(void** p, int64_t bytes);
void ps_host_free(void* p);
/* Page-lock an array the caller already owns (a solver's g / J / H / x vectors, which it reuses for every iteration):
 * the single-scenario callbacks and ps_batch_eval_host then copy to / from it at the link's full rate instead of through
 * the driver's staging of pageable memory (case13659pegase, 12.6 MB of g + J + H per call: see DESIGN.md).  Optional;
 * unregister before the array is freed. */
int ps_host_register(void* p, int64_t bytes);
int ps_host_unregister(void* p);

/* ---- COO -> CSC assembly with a fixed precomputed pattern ---------------------------------- */
/* Replaces compute_jacobian_matrix(m, n, j_str, dE) (src/opt/algorithms/common.jl:12-20), which
 * inserts `J[r,c] += dE[i]` one element at a time into an empty CSC on every call.  The pattern
 * (colptr / rowval / permutation with duplicate segments) is built once from the 1-based COO tuples. */
int ps_csc_create(int64_t m, int64_t n, int64_t nnz_coo, const int64_t* row, const int64_t* col,
                  int device, ps_csc** out);
void ps_csc_destroy(ps_csc* p);
int ps_csc_dims(const ps_csc* p, int64_t* nnz_csc);
/* SparseMatrixCSC fields: colptr [n+1] and rowval [nnz_csc], 1-based */
int ps_csc_structure(const ps_csc* p, int64_t* colptr, int64_t* rowval);
/* nzval[j] = sum of dE over the COO entries mapped to CSC slot j (in COO order).  Host pointers. */
int ps_csc_assemble(ps_csc* p, const double* dE, double* nzval);
/* batched, DEVICE pointers: dE [S][ldE] -> nzval [S][ldnz]; asynchronous on `stream` */
int ps_csc_assemble_batch(ps_csc* p, int64_t S, const double* dE, int64_t ldE, double* nzval, int64_t ldnz,
                          void* stream);

/* ---- affine + quadratic rows of the solver-level model, on device ------------------------------ */
/* Replaces the row evaluation Powersense.Optimizer does itself for the `@constraint` rows that precede the NLP block
 * in g / J / lambda (src/opt/MOI_wrapper.jl:864-891 eval_function, :973-1002 fill_constraint_jacobian!, :1030-1042
 * fill_hessian_lagrangian!, structures :780-800 and :834-839), so that a batched driver keeps the whole g / J / H on the
 * device.  Row r (MOI.ScalarAffineFunction / ScalarQuadraticFunction) is given in CSR form, terms in MOI order:
 *   value  = constant[r] + sum coef*x[var] + sum (var1 == var2 ? 0.5 : 1) * qcoef * x[var1] * x[var2]
 *   J      : one entry per affine term (coef), then per quadratic term qcoef*x[var2] at (r,var1) and, when
 *            var1 != var2, qcoef*x[var1] at (r,var2)
 *   H      : one entry per quadratic term, lambda[r] * qcoef at (var1, var2)
 * The objective is a 1-row block (lambda[0] = obj_factor; its "J" entries are the sparse gradient terms).
 * Variable indices are 1-based; structures are 1-based with rows counted inside this block. */
typedef struct ps_rows ps_rows;
int ps_rows_create(int64_t n_var, int64_t m, const int64_t* aff_ptr, const int64_t* aff_var, const double* aff_coef,
                   const int64_t* quad_ptr, const int64_t* quad_var1, const int64_t* quad_var2, const double* quad_coef,
                   const double* constant, int device, ps_rows** out);
void ps_rows_destroy(ps_rows* r);
int ps_rows_dims(const ps_rows* r, int64_t* nnz_jac, int64_t* nnz_hess);
int ps_rows_jacobian_structure(const ps_rows* r, int64_t* row, int64_t* col);
int ps_rows_hessian_structure(const ps_rows* r, int64_t* row, int64_t* col);
/* host pointers, one x (what: PS_EVAL_G | PS_EVAL_J | PS_EVAL_H; lambda needed for H only) */
int ps_rows_eval(ps_rows* r, int what, const double* x, const double* lambda, double* g, double* J, double* H);
/* DEVICE pointers, S scenarios, row s of an array at ptr + s*ld; asynchronous on `stream` */
int ps_rows_eval_batch(ps_rows* r, int what, int64_t S, const double* x, int64_t ldx, const double* lambda, int64_t ldl,
                       double* g, int64_t ldg, double* J, int64_t ldJ, double* H, int64_t ldH, void* stream);

/* ---- KKT / merit metrics of the SLP iteration, per scenario, on device ------------------------------ */
/* Replaces, for a batch of scenarios, norm_violations (src/opt/algorithms/common.jl:76-98), norm_complementarity
 * (:51-69), KT_residuals (:35-44, which row-slices a CSC once per row), and the merit value / directional derivative of
 * the line search at alpha = 0 (src/opt/algorithms/slp.jl:108-131, 138-162; not the feasibility-restoration branch).
 * The Jacobian comes assembled on the fixed CSC pattern of ps_csc_create (nzval from ps_csc_assemble_batch).
 * out[s][0..PS_NKKT): 0 norm_violations p=1, 1 norm_violations p=Inf, 2 norm_complementarity p=Inf, 3 KT_residuals,
 * 4 phi = f + sum nu_i * viol_i, 5 D = df'p - sum nu_i * viol_i, 6 ||df||_2, 7 max_i |lambda_i| * ||J[i,:]||_2. */
#define PS_NKKT 8
typedef struct ps_kkt ps_kkt;
/* bounds: g_L, g_U [m], x_L, x_U [n] (host; +-Inf allowed); the pattern's index arrays are copied */
int ps_kkt_create(ps_csc* pattern, const double* g_L, const double* g_U, const double* x_L, const double* x_U, ps_kkt** out);
void ps_kkt_destroy(ps_kkt* k);
/* DEVICE pointers; row s of an array at ptr + s*ld; f [S]; nu / p may be NULL (out[4], out[5] are then f and 0);
 * asynchronous on `stream` */
int ps_kkt_eval_batch(ps_kkt* k, int64_t S, const double* E, int64_t ldE, const double* x, int64_t ldx,
                      const double* lambda, int64_t ldl, const double* mult_x_U, const double* mult_x_L, int64_t ldm,
                      const double* df, int64_t lddf, const double* nzval, int64_t ldnz, const double* nu, int64_t ldnu,
                      const double* p, int64_t ldp, const double* f, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POWERSENSE_B200_H */
