// C ABI of the B200-native evaluator (include/powersense_b200.h): handles, device tables, the
// single-scenario MOI-style callbacks, the batched device-resident driver, the host-streaming
// driver and the fixed-pattern COO -> CSC assembly.  No CPU fallback anywhere in this file: every
// numeric result is produced by the kernels in ps_kernels.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/powersense_b200.h"
#include "ps_kernels.cuh"

using namespace ps;

namespace {

thread_local std::string g_create_error;

struct DevMem {                      // owns device allocations of one object
  std::vector<void*> ptrs;
  cudaError_t err = cudaSuccess;
  template <class T> T* alloc(size_t n) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) { if (err == cudaSuccess) err = e; return nullptr; }
    ptrs.push_back(p);
    return (T*)p;
  }
  template <class T> T* upload(const std::vector<T>& v) {
    T* p = alloc<T>(v.size());
    if (p && !v.empty()) {
      cudaError_t e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
      if (e != cudaSuccess && err == cudaSuccess) err = e;
    }
    return p;
  }
  void release() { for (void* p : ptrs) cudaFree(p); ptrs.clear(); }
};

}  // namespace

struct ps_handle {
  Plan plan;
  DevPlan dev{};
  DevMem mem;
  int device = 0;
  std::string err;
  ps_batch* single = nullptr;        // S = 1 batch behind the MOI-style callbacks
  double *h_in = nullptr, *h_out = nullptr;   // pinned staging for the single-scenario calls
  // result cache of the single-scenario callbacks: h_in[0 .. n_var) / d_x hold the x of the last call,
  // cache_dev says which of d_g / d_J belong to it; want_prev / want_cur = outputs asked for at the previous / current x
  bool cache_x = false;
  int cache_dev = 0, want_prev = EV_G | EV_J, want_cur = 0;
  // Hessian slot order imposed by the caller (ps_set_hessian_order): reported slot i = canonical slot hperm[i]
  std::vector<int32_t> hperm;
  int32_t* d_hperm = nullptr;
  double *d_x = nullptr, *d_mu = nullptr, *d_g = nullptr, *d_J = nullptr, *d_H = nullptr, *d_scal = nullptr;
  cudaStream_t stream = nullptr;
};

struct ps_batch {
  ps_handle* h = nullptr;
  int device = 0;                    // copied from the handle: ps_batch_destroy does not touch it
  int64_t S = 0;
  bool status_active = false;        // a contingency mask is in force (ps_batch_set_branch_status(NULL) clears it)
  double* d_Hs = nullptr;            // canonical-order Hessian slab when the handle carries a slot permutation
  int64_t Hs_rows = 0;
  DevMem mem;
  double *d_Pd = nullptr, *d_Qd = nullptr;
  uint8_t* d_status = nullptr;
  double* d_diag = nullptr;          // CB* fast path scratch [S][4*nbus] (allocated on first use)
  double viol_tol = 1e-6;
  int64_t launches = 0;
  int chunk_override = 0;
  bool timing = false;               // per-kernel CUDA events (ps_batch_enable_timing)
  cudaEvent_t tev[2 * K_COUNT] = {};
  unsigned tmask = 0;
  // host-streaming state (ps_batch_eval_host)
  int64_t stage_S = 0;
  double *sx[2] = {nullptr, nullptr}, *smu[2] = {nullptr, nullptr}, *sg[2] = {nullptr, nullptr},
         *sJ[2] = {nullptr, nullptr}, *sH[2] = {nullptr, nullptr}, *sscal[2] = {nullptr, nullptr};
  cudaStream_t st_in = nullptr, st_k = nullptr, st_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

struct ps_csc {
  int64_t m = 0, n = 0, nnz_coo = 0, nnz = 0;
  std::vector<int64_t> colptr, rowval;     // 1-based (SparseMatrixCSC)
  std::vector<int32_t> segptr, perm;       // CSC slot j sums dE[perm[segptr[j] .. segptr[j+1])]
  int device = 0;
  DevMem mem;
  int32_t *d_segptr = nullptr, *d_perm = nullptr;
  double *d_in = nullptr, *d_out = nullptr;
  std::string err;
};

struct ps_rows {
  int64_t n_var = 0, m = 0, nnzj = 0, nnzh = 0;
  std::vector<int64_t> jrow, jcol, hrow, hcol;   // 1-based structures
  RowsDev dev{};
  DevMem mem;
  int device = 0;
  double *d_x = nullptr, *d_lam = nullptr, *d_g = nullptr, *d_J = nullptr, *d_H = nullptr;
};

struct ps_kkt {
  int64_t nnz = 0;                   // copied from the pattern (the ps_csc may be destroyed first)
  KktDev dev{};
  DevMem mem;
  int device = 0;
};

namespace {

int fail(ps_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
int cuda_fail(ps_handle* h, cudaError_t e, const char* where) {
  return fail(h, PS_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

template <class T> std::vector<T> vec(const T* p, int64_t n) { return p ? std::vector<T>(p, p + n) : std::vector<T>(); }

bool to_i32_csr(const int64_t* ptr, const int64_t* idx, int64_t nb, int64_t nmax, std::vector<int32_t>& optr,
                std::vector<int32_t>& oidx) {
  if (!ptr) return false;
  optr.resize(nb + 1);
  for (int64_t i = 0; i <= nb; i++) {
    if (ptr[i] < 0 || (i && ptr[i] < ptr[i - 1])) return false;
    optr[i] = (int32_t)ptr[i];
  }
  if (ptr[0] != 0) return false;
  const int64_t n = ptr[nb];
  if (n && !idx) return false;
  oidx.resize(n);
  for (int64_t q = 0; q < n; q++) {
    if (idx[q] < 1 || idx[q] > nmax) return false;
    oidx[q] = (int32_t)(idx[q] - 1);
  }
  return true;
}

int pick_chunk(const ps_batch* b, int64_t S, int ntiles) {
  if (b->chunk_override > 0) return b->chunk_override;
  if (const char* e = getenv("PS_CHUNK")) { const int c = atoi(e); if (c > 0) return c; }
  // enough CTAs for ~8 waves of 148 SMs x 3 resident CTAs, while amortising per-tile constants
  const int64_t want = 148 * 3 * 8;
  int64_t c = (S * ntiles) / want;
  return (int)std::max<int64_t>(1, std::min<int64_t>(16, c));
}

int do_eval(ps_batch* b, int what, int64_t s_begin, int64_t s_count, const double* x, int64_t ldx, const double* mu,
            int64_t ldmu, double* g, int64_t ldg, double* J, int64_t ldJ, double* H, int64_t ldH, double* scalars,
            cudaStream_t stream) {
  ps_handle* h = b->h;
  const Plan& P = h->plan;
  if (s_count == 0) return PS_OK;
  if (s_begin < 0 || s_count < 0 || s_begin + s_count > b->S) return fail(h, PS_ERR_ARG, "scenario range outside the batch");
  if (!x || ldx < P.n_var) return fail(h, PS_ERR_ARG, "x is null or ldx < n_var");
  if (P.m_nl == 0) what &= ~EV_G;                    // empty blocks (e.g. a W-relaxation without branches): nothing to write
  if (P.nnzJ == 0) what &= ~EV_J;
  if (P.nnzH == 0) what &= ~EV_H;
  if ((what & EV_G) && (!g || ldg < P.m_nl)) return fail(h, PS_ERR_ARG, "g is null or ldg < m_nl");
  if ((what & EV_J) && (!J || ldJ < P.nnzJ)) return fail(h, PS_ERR_ARG, "J is null or ldJ < nnz_jac");
  if ((what & EV_H) && (!H || ldH < P.nnzH)) return fail(h, PS_ERR_ARG, "H is null or ldH < nnz_hess");
  if ((what & EV_H) && (!mu || ldmu < P.m_nl)) return fail(h, PS_ERR_ARG, "mu is null or ldmu < m_nl (needed for the Hessian)");
  if (!(what & (EV_G | EV_J | EV_H)) && !scalars) return PS_OK;      // nothing to write
  if (s_count > 0x7fffffff) return fail(h, PS_ERR_ARG, "too many scenarios in one call");
  cudaError_t e;
  if (scalars) {
    e = cudaMemsetAsync(scalars, 0, (size_t)s_count * NSCAL * sizeof(double), stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMemsetAsync(scalars)");
  }
  EvalArgs A{};
  A.what = what; A.S = (int)s_count; A.chunk = pick_chunk(b, s_count, h->dev.ntiles);
  A.x = x; A.ldx = ldx; A.mu = mu; A.ldmu = ldmu;
  A.Pd = b->d_Pd ? b->d_Pd + s_begin * P.net.nbus : nullptr;
  A.Qd = b->d_Qd ? b->d_Qd + s_begin * P.net.nbus : nullptr;
  A.ldl = P.net.nbus;
  A.status = (b->d_status && b->status_active) ? b->d_status + s_begin * P.net.nbr : nullptr;
  A.ldst = P.net.nbr;
  A.g = g; A.ldg = ldg; A.J = J; A.ldJ = ldJ; A.H = H; A.ldH = ldH;
  A.scal = scalars; A.viol_tol = b->viol_tol;
  A.diag = nullptr; A.lddiag = 4 * P.net.nbus;
  A.force_any = getenv("PS_WT_ANY") ? 1 : 0;
  if (h->dev.cb_slots && !A.status && (what & EV_J)) {
    if (!b->d_diag) {
      b->d_diag = b->mem.alloc<double>((size_t)b->S * 4 * P.net.nbus);
      if (!b->d_diag) return cuda_fail(h, b->mem.err, "cudaMalloc(CB diagonal scratch)");
    }
    A.diag = b->d_diag + s_begin * 4 * P.net.nbus;
  }
  b->tmask = 0;
  if ((what & EV_H) && h->d_hperm) {
    // a caller-imposed Hessian slot order: the kernels write their canonical order into a slab (laid out like
    // ps_preferred_layout, so the fast store paths apply), one gather per slab moves it to the caller's rows
    const int64_t ldHs = (P.nnzH + 3) / 4 * 4, pad = (4 - P.hptr[P.bus_rows] % 4) % 4;
    const int64_t rows = std::max<int64_t>(1, std::min<int64_t>(s_count, (int64_t(256) << 20) / (ldHs * 8)));
    if (b->Hs_rows < rows) {
      // (the old slab, if any, stays owned by b->mem until the batch is destroyed; slabs only grow on the first calls)
      b->d_Hs = b->mem.alloc<double>((size_t)(rows * ldHs + 4));
      if (!b->d_Hs) return cuda_fail(h, b->mem.err, "cudaMalloc(Hessian slab)");
      b->Hs_rows = rows;
    }
    for (int64_t c0 = 0; c0 < s_count; c0 += rows) {
      const int64_t n = std::min<int64_t>(rows, s_count - c0);
      EvalArgs As = A;
      As.S = (int)n; As.chunk = pick_chunk(b, n, h->dev.ntiles);
      As.x = x + c0 * ldx; As.mu = mu + c0 * ldmu;
      if (As.Pd) { As.Pd += c0 * P.net.nbus; As.Qd += c0 * P.net.nbus; }
      if (As.status) As.status += c0 * P.net.nbr;
      if (As.g) As.g += c0 * ldg;
      if (As.J) As.J += c0 * ldJ;
      if (As.scal) As.scal += c0 * NSCAL;
      if (As.diag) As.diag += c0 * As.lddiag;
      As.H = b->d_Hs + pad; As.ldH = ldHs;
      e = launch_eval(h->dev, As, P.smem_branch_bytes, P.smem_bus_bytes, stream, &b->launches, b->timing ? b->tev : nullptr, &b->tmask);
      if (e == cudaSuccess) { e = launch_permute_rows(h->d_hperm, (int)P.nnzH, n, As.H, ldHs, H + c0 * ldH, ldH, stream); b->launches += 1; }
      if (e != cudaSuccess) return cuda_fail(h, e, "eval_kernel launch (permuted Hessian)");
    }
  } else {
    e = launch_eval(h->dev, A, P.smem_branch_bytes, P.smem_bus_bytes, stream, &b->launches, b->timing ? b->tev : nullptr, &b->tmask);
    if (e != cudaSuccess) return cuda_fail(h, e, "eval_kernel launch");
  }
  b->launches += (scalars ? 1 : 0);                  // the memset node
  return PS_OK;
}

}  // namespace

extern "C" {

int ps_abi_version(void) { return PS_ABI_VERSION; }

const char* ps_last_error(const ps_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ps_create(const ps_network* net, int formulation, int source_type, int flags, int device, ps_handle** out) {
  if (!net || !out) return fail(nullptr, PS_ERR_ARG, "null argument");
  *out = nullptr;
  if (net->nbus <= 0 || net->nbr < 0 || net->ngen < 0 || net->nbss < 0 || net->n_tgh < 0)
    return fail(nullptr, PS_ERR_ARG, "negative or zero table size");
  if (net->nbus > (1 << 28) || net->nbr > (1 << 28) || net->ngen > (1 << 28))
    return fail(nullptr, PS_ERR_ARG, "table too large for 32-bit device indices");
  const int64_t nb = net->nbus, nr = net->nbr, ng = net->ngen;
  if (nr && (!net->br_f || !net->br_t || !net->Gf || !net->Bf || !net->Gt || !net->Bt || !net->gf || !net->bf ||
             !net->gt || !net->bt || !net->Imax))
    return fail(nullptr, PS_ERR_ARG, "null branch table");
  if (!net->Pd || !net->Qd || !net->Gl || !net->Bl) return fail(nullptr, PS_ERR_ARG, "null bus table");
  HostNet hn;
  hn.nbus = nb; hn.nbr = nr; hn.ngen = ng; hn.nbss = net->nbss; hn.n_tgh = net->n_tgh;
  hn.f.resize(nr); hn.t.resize(nr);
  for (int64_t k = 0; k < nr; k++) {
    if (net->br_f[k] < 1 || net->br_f[k] > nb || net->br_t[k] < 1 || net->br_t[k] > nb)
      return fail(nullptr, PS_ERR_NETWORK, "branch endpoint out of range");
    hn.f[k] = (int32_t)(net->br_f[k] - 1);
    hn.t[k] = (int32_t)(net->br_t[k] - 1);
  }
  hn.Gf = vec(net->Gf, nr); hn.Bf = vec(net->Bf, nr); hn.Gt = vec(net->Gt, nr); hn.Bt = vec(net->Bt, nr);
  hn.gf = vec(net->gf, nr); hn.bf = vec(net->bf, nr); hn.gt = vec(net->gt, nr); hn.bt = vec(net->bt, nr);
  hn.Imax = vec(net->Imax, nr);
  hn.Pd = vec(net->Pd, nb); hn.Qd = vec(net->Qd, nb); hn.Gl = vec(net->Gl, nb); hn.Bl = vec(net->Bl, nb);
  if (!to_i32_csr(net->gen_ptr, net->gen_idx, nb, ng, hn.gen_ptr, hn.gen_idx)) return fail(nullptr, PS_ERR_NETWORK, "bad gen adjacency list");
  if (!to_i32_csr(net->sij_ptr, net->sij_idx, nb, nr, hn.sij_ptr, hn.sij_idx)) return fail(nullptr, PS_ERR_NETWORK, "bad Sij adjacency list");
  if (!to_i32_csr(net->sji_ptr, net->sji_idx, nb, nr, hn.sji_ptr, hn.sji_idx)) return fail(nullptr, PS_ERR_NETWORK, "bad Sji adjacency list");
  if (!to_i32_csr(net->sh_ptr, net->sh_idx, nb, std::max<int64_t>(net->nbss, 0), hn.sh_ptr, hn.sh_idx))
    return fail(nullptr, PS_ERR_NETWORK, "bad shunt adjacency list");
  hn.c0 = vec(net->c0, ng); hn.c1 = vec(net->c1, ng); hn.c2 = vec(net->c2, ng);
  hn.cost_order = net->cost_order;

  ps_handle* h = new ps_handle();
  const std::string msg = build_plan(hn, formulation, source_type, flags, TPB, BTPB, h->plan);
  if (!msg.empty()) {
    const bool arg = msg.rfind("unknown", 0) == 0;
    delete h;
    return fail(nullptr, arg ? PS_ERR_ARG : PS_ERR_NETWORK, msg);
  }
  const Plan& P = h->plan;
  if (P.n_var >= (int64_t(1) << 31)) { delete h; return fail(nullptr, PS_ERR_ARG, "n_var too large for 32-bit device offsets"); }

  // ---- device: there is no CPU fallback.  PS_DEVICE_NONE gives a structure-only handle (dims, bounds,
  // sparsity structures: host logic); every evaluation entry point fails on such a handle.
  if (device == PS_DEVICE_NONE) {
    h->device = PS_DEVICE_NONE;
    *out = h;
    return PS_OK;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    delete h;
    return fail(nullptr, PS_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                          (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  if (device < 0) { e = cudaGetDevice(&device); if (e != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaGetDevice"); } }
  if (device >= ndev) { delete h; return fail(nullptr, PS_ERR_ARG, "device ordinal out of range"); }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaSetDevice"); }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaGetDeviceProperties"); }
  if (prop.major != 10) {
    delete h;
    return fail(nullptr, PS_ERR_CUDA, "device is not sm_100 (B200): the kernels are built for sm_100a only");
  }
  h->device = device;
  DevMem& M = h->mem;
  DevPlan& D = h->dev;
  D.form = P.form; D.src = P.src; D.flags = P.flags;
  D.nb = (int)nb; D.nr = (int)nr; D.ng = (int)ng;
  D.o_a = (int)P.o_a; D.o_b = (int)P.o_b; D.o_Wd = (int)P.o_Wd; D.o_Wr = (int)P.o_Wr; D.o_Wi = (int)P.o_Wi;
  D.o_pg = (int)P.o_pg; D.o_qg = (int)P.o_qg;
  for (int q = 0; q < 4; q++) D.o_f[q] = (int)P.o_f[q];
  D.o_cg = (int)P.o_cg; D.o_bss = (int)P.o_bss;
  D.bus_rows = (int)P.bus_rows;
  D.ntiles = (int)P.tiles.size();
  D.n_bus_tiles = P.n_bus_tiles;
  D.j_br0 = P.jptr[P.bus_rows]; D.h_br0 = P.hptr[P.bus_rows];
  D.pb_slots = P.pb_slots ? 1 : 0;
  if (P.pb_slots) {
    // PB* slot tables: slot e of the bus block is coefficient[e] * m, with m = x[index[e]] (own-voltage slots),
    // 1 (generator slots: index -1) or the branch status (flow slots: index -2 - branch); the Hessian slots are
    // coefficient[e] * mu[index[e]].  Four copies of each table, copy a shifted so that entry a + 4q sits on a
    // 32-byte (values) / 16-byte (indices) boundary: the slot-parallel kernel reads the four entries of one output
    // sector with one vector load whatever the row alignment a is.
    const int64_t nj = (int64_t)P.pb_jcode.size(), nh = (int64_t)P.pb_hcode.size();
    std::vector<double> jt((size_t)nj), ht((size_t)nh);
    std::vector<int32_t> ji((size_t)nj), hi((size_t)nh);
    for (int64_t e = 0; e < nj; e++) {
      const int32_t c = P.pb_jcode[e];
      if (c == -1) { jt[e] = 1.0; ji[e] = -1; }
      else if (c >= 0) { jt[e] = 1.0; ji[e] = -2 - c; }
      else {
        const int32_t v = -2 - c, i = v >> 2, w = v & 3;
        jt[e] = (w < 2) ? -2.0 * P.net.Gl[i] : 2.0 * P.net.Bl[i];
        ji[e] = (int32_t)(((w & 1) ? P.o_b : P.o_a) + i);
      }
    }
    for (int64_t e = 0; e < nh; e++) {
      const int32_t c = P.pb_hcode[e], i = c >> 1;
      ht[e] = (c & 1) ? 2.0 * P.net.Bl[i] : -2.0 * P.net.Gl[i];
      hi[e] = c;
    }
    auto shifted = [&](const auto& v, int64_t& stride) {
      using T = typename std::decay<decltype(v)>::type::value_type;
      stride = ((int64_t)v.size() + 4 + 3) / 4 * 4;
      std::vector<T> all((size_t)(4 * stride), T(0));
      for (int a = 0; a < 4; a++) std::copy(v.begin(), v.end(), all.begin() + a * stride + (4 - a) % 4);
      return all;
    };
    int64_t sj = 0, sh = 0;
    D.pb_jt = M.upload(shifted(jt, sj)); D.pb_ji = M.upload(shifted(ji, sj));
    D.pb_ht = M.upload(shifted(ht, sh)); D.pb_hi = M.upload(shifted(hi, sh));
    D.pb_jstride = (int)sj; D.pb_hstride = (int)sh;
  }
  D.pbs_ok = P.pbs_ok ? 1 : 0;
  D.pbs_prefer = P.pbs_prefer ? 1 : 0;
  if (P.pbs_ok) { D.pbs_code = M.upload(P.pbs_code); D.pbs_wmax = M.upload(P.pbs_wmax); }
  D.cb_slots = 0;
  if (is_CB(P.form) && nb > 0 && P.bus_rows == 2 * nb) {
    // CB* slot tables (same layout and kernel as the PB* ones).  Every entry of the bus block is coefficient * m:
    //   J  own-voltage slots: m = the sum the residual kernel leaves in the scratch (index -2 - (4*bus + w), w = P-vr, P-vi,
    //      Q-vr, Q-vi); generator / W slots: constants (index -1); current slots: +-1 * x[own voltage] (formulations.jl:206-214)
    //   H  coefficient * mu[row]: +-1 for the (voltage, current) pairs, -2Gl / 2Bl for the shunt diagonal, 0 elsewhere
    bool ok = true;
    if (P.flags & FLAG_FACTS) for (int64_t i = 0; i < nb; i++) ok = ok && P.net.sh_ptr[i + 1] == P.net.sh_ptr[i];
    const int64_t nj = D.j_br0, nh = D.h_br0;
    std::vector<double> jt((size_t)nj), ht((size_t)nh);
    std::vector<int32_t> ji((size_t)nj), hi((size_t)nh);
    auto flow_q = [&](int64_t v) -> int { for (int q = 0; q < 4; q++) if (v >= P.o_f[q] && v < P.o_f[q] + nr) return q; return -1; };
    for (int64_t r = 0; r < 2 * nb && ok; r++) {
      const int64_t i = r >> 1; const int pq = (int)(r & 1);
      const int64_t vr = P.o_b + i, vi = P.o_a + i;
      for (int32_t e = P.jptr[r]; e < P.jptr[r + 1] && ok; e++) {
        const int64_t v = P.jcol[e] - 1;
        const int q = flow_q(v);
        if (v == vr) { jt[e] = 1.0; ji[e] = (int32_t)(-2 - (4 * i + (pq ? 2 : 0))); }
        else if (v == vi) { jt[e] = 1.0; ji[e] = (int32_t)(-2 - (4 * i + (pq ? 3 : 1))); }
        else if ((v >= P.o_pg && v < P.o_pg + ng) || (v >= P.o_qg && v < P.o_qg + ng)) { jt[e] = 1.0; ji[e] = -1; }
        else if (P.form == CBRAWV && v == P.o_Wd + i) { jt[e] = pq ? P.net.Bl[i] : -P.net.Gl[i]; ji[e] = -1; }
        else if (q >= 0) {
          const bool isIr = (q & 1) == 0;
          if (pq == 0) { jt[e] = 1.0; ji[e] = (int32_t)(isIr ? vr : vi); }
          else { jt[e] = isIr ? 1.0 : -1.0; ji[e] = (int32_t)(isIr ? vi : vr); }
        } else ok = false;
      }
      for (int32_t e = P.hptr[r]; e < P.hptr[r + 1] && ok; e++) {
        int64_t a = P.hrow[e] - 1, c = P.hcol[e] - 1;
        ht[e] = 0.0; hi[e] = -1;
        if (a == c) {
          if (P.form == CBRARV && (a == vr || a == vi)) { ht[e] = pq ? 2.0 * P.net.Bl[i] : -2.0 * P.net.Gl[i]; hi[e] = (int32_t)r; }
          continue;
        }
        if (a != vr && a != vi) std::swap(a, c);
        const int q = flow_q(c);
        if ((a != vr && a != vi) || q < 0) { ok = false; break; }
        const bool isIr = (q & 1) == 0, own_r = (a == vr);
        // P: (vr, Ir), (vi, Ii) -> +mu_P ; Q: (vi, Ir) -> +mu_Q, (vr, Ii) -> -mu_Q ; the other pairs do not occur
        if (pq == 0) { if (own_r != isIr) { ok = false; break; } ht[e] = 1.0; }
        else { if (own_r == isIr) { ok = false; break; } ht[e] = own_r ? -1.0 : 1.0; }
        hi[e] = (int32_t)r;
      }
    }
    if (ok) {
      auto shifted = [&](const auto& v, int64_t& stride) {
        using T = typename std::decay<decltype(v)>::type::value_type;
        stride = ((int64_t)v.size() + 4 + 3) / 4 * 4;
        std::vector<T> all((size_t)(4 * stride), T(0));
        for (int a = 0; a < 4; a++) std::copy(v.begin(), v.end(), all.begin() + a * stride + (4 - a) % 4);
        return all;
      };
      int64_t sj = 0, sh = 0;
      D.pb_jt = M.upload(shifted(jt, sj)); D.pb_ji = M.upload(shifted(ji, sj));
      D.pb_ht = M.upload(shifted(ht, sh)); D.pb_hi = M.upload(shifted(hi, sh));
      D.pb_jstride = (int)sj; D.pb_hstride = (int)sh;
      D.cb_slots = 1;
    }
  }
  D.cost_order = P.net.cost_order;
  D.f = M.upload(P.net.f); D.t = M.upload(P.net.t);
  D.Gf = M.upload(P.net.Gf); D.Bf = M.upload(P.net.Bf); D.Gt = M.upload(P.net.Gt); D.Bt = M.upload(P.net.Bt);
  D.gf = M.upload(P.net.gf); D.bf = M.upload(P.net.bf); D.gt = M.upload(P.net.gt); D.bt = M.upload(P.net.bt);
  D.Imax = M.upload(P.net.Imax); D.Gl = M.upload(P.net.Gl); D.Bl = M.upload(P.net.Bl);
  D.Pd0 = M.upload(P.net.Pd); D.Qd0 = M.upload(P.net.Qd);
  D.c0 = M.upload(P.c0); D.c1 = M.upload(P.c1); D.c2 = M.upload(P.c2); D.c3 = M.upload(P.c3);
  D.k0 = P.net.c0.empty() ? nullptr : M.upload(P.net.c0);
  D.k1 = P.net.c1.empty() ? nullptr : M.upload(P.net.c1);
  D.k2 = P.net.c2.empty() ? nullptr : M.upload(P.net.c2);
  D.gen_ptr = M.upload(P.net.gen_ptr); D.gen_idx = M.upload(P.net.gen_idx);
  D.sh_ptr = M.upload(P.net.sh_ptr); D.sh_idx = M.upload(P.net.sh_idx);
  D.bus_jpp = M.upload(P.bus_jpp); D.bus_hpp = M.upload(P.bus_hpp);
  D.he_ptr = M.upload(P.he_ptr); D.he_ks = M.upload(P.he_ks); D.he_other = M.upload(P.he_other);
  D.he_c = M.upload(P.he_c); D.nhe = (int)P.he_ks.size();
  D.he_bus = M.upload(P.he_bus); D.jsrc = M.upload(P.jsrc); D.hsrc = M.upload(P.hsrc);
  D.ms_dst = M.upload(P.ms_dst); D.ms_ptr = M.upload(P.ms_ptr); D.ms_list = M.upload(P.ms_list);
  D.tiles = M.upload(P.tiles);
  D.tile_map = nullptr; D.n_hub_tiles = (int)P.hub_tiles.size(); D.hub_tiles = M.upload(P.hub_tiles);
  D.n_wtiles = (int)P.wtiles.size();
  D.wtiles = M.upload(P.wtiles);
  D.wt_jspace = P.wt_jspace; D.wt_hspace = P.wt_hspace; D.wt_gspace = P.wt_gspace; D.wt_lists = P.wt_lists_max;
  D.wt_dst = M.upload(P.wt_dst); D.wt_multi = M.upload(P.wt_multi);
  D.wt_ownj = M.upload(P.wt_ownj); D.wt_ownh = M.upload(P.wt_ownh);
  D.wt_owng = M.upload(P.wt_owng);
  D.wt_ownj_ptr = M.upload(P.wt_ownj_ptr); D.wt_ownh_ptr = M.upload(P.wt_ownh_ptr); D.wt_owng_ptr = M.upload(P.wt_owng_ptr);
  // single-scenario staging
  const int64_t n_in = P.n_var + P.m_nl, n_out = NSCAL;
  h->d_x = M.alloc<double>(P.n_var); h->d_mu = M.alloc<double>(P.m_nl);
  h->d_g = M.alloc<double>(P.m_nl); h->d_J = M.alloc<double>(P.nnzJ); h->d_H = M.alloc<double>(P.nnzH);
  h->d_scal = M.alloc<double>(NSCAL);
  if (M.err != cudaSuccess) { e = M.err; M.release(); delete h; return cuda_fail(nullptr, e, "device table upload"); }
  if ((e = cudaHostAlloc((void**)&h->h_in, std::max<int64_t>(n_in, 1) * sizeof(double), cudaHostAllocDefault)) != cudaSuccess ||
      (e = cudaHostAlloc((void**)&h->h_out, std::max<int64_t>(n_out, 1) * sizeof(double), cudaHostAllocDefault)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = configure_eval(P.form, P.src, P.smem_branch_bytes, P.smem_bus_bytes, wt_smem_bytes(D), pbs_smem_bytes(D))) != cudaSuccess) {
    ps_destroy(h);
    return cuda_fail(nullptr, e, "handle setup");
  }
  int rc = ps_batch_create(h, 1, &h->single);
  if (rc != PS_OK) { g_create_error = h->err; ps_destroy(h); return rc; }
  *out = h;
  return PS_OK;
}

void ps_destroy(ps_handle* h) {
  if (!h) return;
  if (h->device == PS_DEVICE_NONE) { delete h; return; }
  cudaSetDevice(h->device);
  if (h->single) ps_batch_destroy(h->single);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->h_in) cudaFreeHost(h->h_in);
  if (h->h_out) cudaFreeHost(h->h_out);
  h->mem.release();
  delete h;
}

int ps_dims(const ps_handle* h, int64_t* n_var, int64_t* m_nl, int64_t* nnz_jac, int64_t* nnz_hess) {
  if (!h) return PS_ERR_ARG;
  if (n_var) *n_var = h->plan.n_var;
  if (m_nl) *m_nl = h->plan.m_nl;
  if (nnz_jac) *nnz_jac = h->plan.nnzJ;
  if (nnz_hess) *nnz_hess = h->plan.nnzH;
  return PS_OK;
}

int ps_constraint_bounds(const ps_handle* h, double* lo, double* hi) {
  if (!h || !lo || !hi) return PS_ERR_ARG;
  std::copy(h->plan.lo.begin(), h->plan.lo.end(), lo);
  std::copy(h->plan.hi.begin(), h->plan.hi.end(), hi);
  return PS_OK;
}

int ps_jacobian_structure(const ps_handle* h, int64_t* row, int64_t* col) {
  if (!h || !row || !col) return PS_ERR_ARG;
  std::copy(h->plan.jrow.begin(), h->plan.jrow.end(), row);
  std::copy(h->plan.jcol.begin(), h->plan.jcol.end(), col);
  return PS_OK;
}

int ps_hessian_structure(const ps_handle* h, int64_t* row, int64_t* col) {
  if (!h || !row || !col) return PS_ERR_ARG;
  const Plan& P = h->plan;
  if (h->hperm.empty()) {
    std::copy(P.hrow.begin(), P.hrow.end(), row);
    std::copy(P.hcol.begin(), P.hcol.end(), col);
  } else {
    for (int64_t i = 0; i < P.nnzH; i++) { row[i] = P.hrow[h->hperm[i]]; col[i] = P.hcol[h->hperm[i]]; }
  }
  return PS_OK;
}

int ps_hessian_row_blocks(const ps_handle* h, int64_t* ptr) {
  if (!h || !ptr) return PS_ERR_ARG;
  for (int64_t r = 0; r <= h->plan.m_nl; r++) ptr[r] = h->plan.hptr[r];
  return PS_OK;
}

int ps_hessian_order(const ps_handle* h, int64_t* perm) {
  if (!h || !perm) return PS_ERR_ARG;
  for (int64_t i = 0; i < h->plan.nnzH; i++) perm[i] = h->hperm.empty() ? i + 1 : (int64_t)h->hperm[i] + 1;
  return PS_OK;
}

int ps_set_hessian_order(ps_handle* hh, const int64_t* perm) {
  if (!hh) return PS_ERR_ARG;
  ps_handle* h = hh;
  const int64_t n = h->plan.nnzH;
  std::vector<int32_t> p;
  if (perm) {
    p.resize((size_t)n);
    std::vector<uint8_t> seen((size_t)n, 0);
    bool ident = true;
    for (int64_t i = 0; i < n; i++) {
      if (perm[i] < 1 || perm[i] > n || seen[perm[i] - 1]) return fail(h, PS_ERR_ARG, "Hessian order is not a permutation of 1..nnz_hess");
      seen[perm[i] - 1] = 1;
      p[i] = (int32_t)(perm[i] - 1);
      ident = ident && perm[i] == i + 1;
    }
    if (ident) p.clear();
  }
  if (h->device != PS_DEVICE_NONE) {
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    int32_t* d = nullptr;
    if (!p.empty()) {
      d = h->mem.upload(p);                          // (an earlier table stays owned by the handle until ps_destroy)
      if (!d || h->mem.err != cudaSuccess) return cuda_fail(h, h->mem.err, "cudaMalloc(Hessian order)");
    }
    h->d_hperm = d;
  }
  h->hperm.swap(p);
  return PS_OK;
}

int ps_match_hessian_structure(ps_handle* h, const int64_t* row, const int64_t* col) {
  if (!h) return PS_ERR_ARG;
  if (!row || !col) return fail(h, PS_ERR_ARG, "null structure");
  const Plan& P = h->plan;
  std::vector<int64_t> perm((size_t)P.nnzH);
  std::vector<std::pair<std::pair<int64_t, int64_t>, int64_t>> blk;
  for (int64_t r = 0; r < P.m_nl; r++) {
    const int64_t e0 = P.hptr[r], e1 = P.hptr[r + 1];
    blk.clear();
    for (int64_t e = e0; e < e1; e++) blk.push_back({{std::max(P.hrow[e], P.hcol[e]), std::min(P.hrow[e], P.hcol[e])}, e});
    std::sort(blk.begin(), blk.end());
    std::vector<uint8_t> used(blk.size(), 0);
    for (int64_t i = e0; i < e1; i++) {
      const std::pair<int64_t, int64_t> key{std::max(row[i], col[i]), std::min(row[i], col[i])};
      auto it = std::lower_bound(blk.begin(), blk.end(), std::make_pair(key, (int64_t)-1));
      while (it != blk.end() && it->first == key && used[it - blk.begin()]) ++it;
      if (it == blk.end() || it->first != key)
        return fail(h, PS_ERR_ARG, "Hessian structure differs in the block of NL row " + std::to_string(r + 1) + ": pair (" +
                                       std::to_string(row[i]) + ", " + std::to_string(col[i]) + ") at position " + std::to_string(i + 1));
      used[it - blk.begin()] = 1;
      perm[i] = it->second + 1;
    }
  }
  return ps_set_hessian_order(h, perm.data());
}

int ps_variable_offsets(const ps_handle* h, int64_t out[16]) {
  if (!h || !out) return PS_ERR_ARG;
  const Plan& P = h->plan;
  const int64_t v[16] = {P.o_a, P.o_b, P.o_Wd, P.o_Wr, P.o_Wi, P.o_pg, P.o_qg, P.o_f[0], P.o_f[1], P.o_f[2], P.o_f[3],
                         P.o_cg, P.o_bss, P.o_tgh, P.n_var, P.nx_nl};
  std::copy(v, v + 16, out);
  return PS_OK;
}

int ps_plan_info(const ps_handle* h, int64_t out[8]) {
  if (!h || !out) return PS_ERR_ARG;
  const Plan& P = h->plan;
  out[0] = P.n_bus_tiles;
  out[1] = (int64_t)P.tiles.size() - P.n_bus_tiles - 1;
  out[2] = P.smem_bus_bytes;
  out[3] = P.smem_branch_bytes;
  out[4] = direct_ok(P.form, P.src) ? 1 : 0;
  out[5] = (P.pb_implicit ? 1 : 0) | (P.pb_slots ? 2 : 0);
  out[6] = P.jptr[P.bus_rows];
  out[7] = P.hptr[P.bus_rows];
  return PS_OK;
}

int ps_preferred_layout(const ps_handle* h, int64_t ld[3], int64_t pad[3]) {
  if (!h || !ld || !pad) return PS_ERR_ARG;
  const Plan& P = h->plan;
  const int64_t n[3] = {P.m_nl, P.nnzJ, P.nnzH};
  const int64_t first[3] = {P.bus_rows, P.jptr[P.bus_rows], P.hptr[P.bus_rows]};
  for (int q = 0; q < 3; q++) {
    ld[q] = (n[q] + 3) / 4 * 4;
    pad[q] = (4 - first[q] % 4) % 4;
  }
  return PS_OK;
}

int64_t ps_algorithmic_bytes(const ps_handle* h, int what) {
  if (!h) return 0;
  const Plan& P = h->plan;
  // reads x restricted to the variables NL rows touch, mu (Hessian only), Pd/Qd (g only);
  // writes g, J, H.  Network constants are shared by all scenarios and excluded.
  int64_t n = P.nx_nl;
  if (what & EV_G) n += P.m_nl + (has_bus_rows(P.form) ? 2 * P.net.nbus : 0);
  if (what & EV_J) n += P.nnzJ;
  if (what & EV_H) n += P.nnzH + P.m_nl;
  return 8 * n;
}

int64_t ps_algorithmic_bytes_branch(const ps_handle* h, int what) {
  if (!h) return 0;
  const Plan& P = h->plan;
  const int64_t nr = P.net.nbr, NR = P.rows_per_branch;
  int64_t n = P.nx_branch;
  if (what & EV_G) n += NR * nr;
  if (what & EV_J) n += (int64_t)P.RJ * nr;
  if (what & EV_H) n += (int64_t)P.RH * nr + NR * nr;
  return 8 * n;
}

// ---- single-scenario callbacks ------------------------------------------------------------
// The x of the last call stays in pinned h_in (the comparison key) and on the device, together with the g / J computed
// from it: JuMP's evaluator reuses its forward pass while x is unchanged and solvers rely on that -- Ipopt
// calls eval_g, eval_jac_g and eval_h at the same point.  Outputs go from the device straight to the caller's buffers.
static int eval_single(ps_handle* h, int what, const double* x, const double* mu, double* g, double* J, double* H,
                       double* f) {
  if (!h) return PS_ERR_ARG;
  if (!x) return fail(h, PS_ERR_ARG, "x is null");
  if (h->device == PS_DEVICE_NONE) return fail(h, PS_ERR_STATE, "structure-only handle (PS_DEVICE_NONE): no CPU fallback for evaluation");
  const Plan& P = h->plan;
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
  if ((what & EV_H) && !mu) return fail(h, PS_ERR_ARG, "mu is null");
  const size_t xb = (size_t)P.n_var * sizeof(double);
  const bool same_x = h->cache_x && !getenv("PS_NO_CACHE") && std::memcmp(h->h_in, x, xb) == 0;
  if (!same_x) {
    std::memcpy(h->h_in, x, xb);
    e = cudaMemcpyAsync(h->d_x, h->h_in, xb, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) { h->cache_x = false; return cuda_fail(h, e, "H2D copy"); }
    h->cache_x = true; h->cache_dev = 0;
    if (h->want_cur) h->want_prev = h->want_cur;
    h->want_cur = 0;
  }
  h->want_cur |= what & (EV_G | EV_J);
  // what has to be computed: the requested g / J that the device does not hold for this x -- widened to g + J when the
  // caller's previous point asked for both -- and H, whose multipliers arrive with the call
  int need = what & (EV_G | EV_J) & ~h->cache_dev;
  if (need && (h->want_prev & (EV_G | EV_J)) == (EV_G | EV_J)) need = (EV_G | EV_J) & ~h->cache_dev;
  if (P.m_nl == 0) need &= ~EV_G;
  if (P.nnzJ == 0) need &= ~EV_J;
  const int lw = need | (what & EV_H);
  if (what & EV_H) {
    std::memcpy(h->h_in + P.n_var, mu, (size_t)P.m_nl * sizeof(double));
    e = cudaMemcpyAsync(h->d_mu, h->h_in + P.n_var, (size_t)P.m_nl * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "H2D copy");
  }
  if (lw || f) {
    int rc = do_eval(h->single, lw, 0, 1, h->d_x, P.n_var, h->d_mu, P.m_nl, h->d_g, P.m_nl, h->d_J, P.nnzJ, h->d_H, P.nnzH,
                     f ? h->d_scal : nullptr, h->stream);
    if (rc != PS_OK) { h->cache_x = false; return rc; }
    h->cache_dev |= need;
  }
  if ((what & EV_G) && P.m_nl) e = cudaMemcpyAsync(g, h->d_g, (size_t)P.m_nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && (what & EV_J) && P.nnzJ)
    e = cudaMemcpyAsync(J, h->d_J, (size_t)P.nnzJ * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && (what & EV_H) && P.nnzH)
    e = cudaMemcpyAsync(H, h->d_H, (size_t)P.nnzH * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && f)
    e = cudaMemcpyAsync(h->h_out, h->d_scal, NSCAL * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);   // callers read the output immediately
  if (e != cudaSuccess) { h->cache_x = false; return cuda_fail(h, e, "D2H copy / kernel"); }
  if (f) *f = h->h_out[0];
  return PS_OK;
}

int64_t ps_launch_count(const ps_handle* h) { return (h && h->single) ? h->single->launches : 0; }

int ps_eval_constraint(ps_handle* h, const double* x, double* g) {
  if (h && !g) return fail(h, PS_ERR_ARG, "g is null");
  return eval_single(h, EV_G, x, nullptr, g, nullptr, nullptr, nullptr);
}
int ps_eval_jacobian(ps_handle* h, const double* x, double* J) {
  if (h && !J) return fail(h, PS_ERR_ARG, "J is null");
  return eval_single(h, EV_J, x, nullptr, nullptr, J, nullptr, nullptr);
}
int ps_eval_hessian(ps_handle* h, const double* x, double sigma, const double* mu, double* H) {
  (void)sigma;  // no NL objective in any Powersense OPF model (has_objective = false)
  if (h && !H) return fail(h, PS_ERR_ARG, "H is null");
  return eval_single(h, EV_H, x, mu, nullptr, nullptr, H, nullptr);
}
int ps_eval_all(ps_handle* h, const double* x, const double* mu, double* g, double* J, double* H) {
  const int what = (g ? EV_G : 0) | (J ? EV_J : 0) | (H ? EV_H : 0);
  if (h && !what) return fail(h, PS_ERR_ARG, "all outputs are null");
  return eval_single(h, what, x, mu, g, J, H, nullptr);
}
int ps_eval_objective(ps_handle* h, const double* x, double* f) {
  if (h && !f) return fail(h, PS_ERR_ARG, "f is null");
  return eval_single(h, 0, x, nullptr, nullptr, nullptr, nullptr, f);
}

// ---- batched -----------------------------------------------------------------------------
int ps_batch_create(ps_handle* h, int64_t S, ps_batch** out) {
  if (!h || !out) return PS_ERR_ARG;
  *out = nullptr;
  if (S <= 0) return fail(h, PS_ERR_ARG, "S must be positive");
  if (h->device == PS_DEVICE_NONE) return fail(h, PS_ERR_STATE, "structure-only handle (PS_DEVICE_NONE): no CPU fallback for evaluation");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
  ps_batch* b = new ps_batch();
  b->h = h; b->S = S; b->device = h->device;
  *out = b;
  return PS_OK;
}

void ps_batch_destroy(ps_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  for (int q = 0; q < 2; q++) {
    if (b->ev_in[q]) cudaEventDestroy(b->ev_in[q]);
    if (b->ev_k[q]) cudaEventDestroy(b->ev_k[q]);
    if (b->ev_out[q]) cudaEventDestroy(b->ev_out[q]);
  }
  for (int q = 0; q < 2 * K_COUNT; q++) if (b->tev[q]) cudaEventDestroy(b->tev[q]);
  if (b->st_in) cudaStreamDestroy(b->st_in);
  if (b->st_k) cudaStreamDestroy(b->st_k);
  if (b->st_out) cudaStreamDestroy(b->st_out);
  b->mem.release();
  delete b;
}

int ps_batch_set_loads(ps_batch* b, const double* Pd, const double* Qd) {
  if (!b) return PS_ERR_ARG;
  ps_handle* h = b->h;
  if (!Pd || !Qd) return fail(h, PS_ERR_ARG, "Pd / Qd is null");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
  const size_t n = (size_t)b->S * h->plan.net.nbus;
  if (!b->d_Pd) { b->d_Pd = b->mem.alloc<double>(n); b->d_Qd = b->mem.alloc<double>(n); }
  if (!b->d_Pd || !b->d_Qd) return cuda_fail(h, b->mem.err, "cudaMalloc(loads)");
  if ((e = cudaMemcpy(b->d_Pd, Pd, n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(b->d_Qd, Qd, n * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
    return cuda_fail(h, e, "cudaMemcpy(loads)");
  return PS_OK;
}

int ps_batch_set_branch_status(ps_batch* b, const uint8_t* st) {
  if (!b) return PS_ERR_ARG;
  ps_handle* h = b->h;
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
  const size_t n = (size_t)b->S * h->plan.net.nbr;
  if (!st) {                       // reset: every branch in service (the buffer stays allocated; the kernels no longer read it)
    b->status_active = false;
    return PS_OK;
  }
  for (size_t q = 0; q < n; q++) if (st[q] > 1) return fail(h, PS_ERR_ARG, "branch status must be 0 or 1");
  if (!b->d_status) b->d_status = b->mem.alloc<uint8_t>(n);
  if (!b->d_status) return cuda_fail(h, b->mem.err, "cudaMalloc(status)");
  if ((e = cudaMemcpy(b->d_status, st, n, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(h, e, "cudaMemcpy(status)");
  b->status_active = true;
  return PS_OK;
}

int ps_batch_set_viol_tol(ps_batch* b, double tol) {
  if (!b) return PS_ERR_ARG;
  if (!(tol >= 0.0)) return fail(b->h, PS_ERR_ARG, "tolerance must be >= 0");
  b->viol_tol = tol;
  return PS_OK;
}

int64_t ps_batch_launch_count(const ps_batch* b) { return b ? b->launches : 0; }

int ps_batch_enable_timing(ps_batch* b, int on) {
  if (!b) return PS_ERR_ARG;
  cudaError_t e = cudaSetDevice(b->h->device);
  if (e != cudaSuccess) return cuda_fail(b->h, e, "cudaSetDevice");
  if (on && !b->tev[0])
    for (int q = 0; q < 2 * K_COUNT; q++)
      if ((e = cudaEventCreate(&b->tev[q])) != cudaSuccess) return cuda_fail(b->h, e, "cudaEventCreate");
  b->timing = on != 0;
  return PS_OK;
}

int ps_batch_kernel_times(ps_batch* b, float ms[4]) {
  if (!b || !ms) return PS_ERR_ARG;
  if (!b->timing) return fail(b->h, PS_ERR_STATE, "timing is not enabled on this batch");
  for (int id = 0; id < K_COUNT; id++) {
    ms[id] = 0.0f;
    if (b->tmask & (1u << id)) {
      cudaError_t e = cudaEventSynchronize(b->tev[2 * id + 1]);
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms[id], b->tev[2 * id], b->tev[2 * id + 1]);
      if (e != cudaSuccess) return cuda_fail(b->h, e, "cudaEventElapsedTime");
    }
  }
  return PS_OK;
}

int ps_batch_eval(ps_batch* b, int what, int64_t s_begin, int64_t s_count, const double* x, int64_t ldx, const double* mu,
                  int64_t ldmu, double* g, int64_t ldg, double* J, int64_t ldJ, double* H, int64_t ldH, double* scalars,
                  void* stream) {
  if (!b) return PS_ERR_ARG;
  cudaError_t e = cudaSetDevice(b->h->device);
  if (e != cudaSuccess) return cuda_fail(b->h, e, "cudaSetDevice");
  return do_eval(b, what, s_begin, s_count, x, ldx, mu, ldmu, g, ldg, J, ldJ, H, ldH, scalars, (cudaStream_t)stream);
}

// Host buffers: scenarios stream through two sets of device staging buffers.  Copy-in (st_in), kernel (st_k)
// and copy-out (st_out) of consecutive chunks overlap; events order the three stages of each chunk and the reuse
// of a staging set two chunks later.
int ps_batch_eval_host(ps_batch* b, int what, int64_t s_begin, int64_t s_count, const double* x, int64_t ldx,
                       const double* mu, int64_t ldmu, double* g, int64_t ldg, double* J, int64_t ldJ, double* H,
                       int64_t ldH, double* scalars) {
  if (!b) return PS_ERR_ARG;
  ps_handle* h = b->h;
  const Plan& P = h->plan;
  if (s_count == 0) return PS_OK;
  if (s_begin < 0 || s_count < 0 || s_begin + s_count > b->S) return fail(h, PS_ERR_ARG, "scenario range outside the batch");
  if (!x || ldx < P.n_var) return fail(h, PS_ERR_ARG, "x is null or ldx < n_var");
  if (P.m_nl == 0) what &= ~EV_G;                    // empty blocks (e.g. a W-relaxation without branches): nothing to write
  if (P.nnzJ == 0) what &= ~EV_J;
  if (P.nnzH == 0) what &= ~EV_H;
  if ((what & EV_G) && (!g || ldg < P.m_nl)) return fail(h, PS_ERR_ARG, "g is null or ldg < m_nl");
  if ((what & EV_J) && (!J || ldJ < P.nnzJ)) return fail(h, PS_ERR_ARG, "J is null or ldJ < nnz_jac");
  if ((what & EV_H) && (!H || ldH < P.nnzH)) return fail(h, PS_ERR_ARG, "H is null or ldH < nnz_hess");
  if ((what & EV_H) && (!mu || ldmu < P.m_nl)) return fail(h, PS_ERR_ARG, "mu is null or ldmu < m_nl (needed for the Hessian)");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
  if (!b->st_in) {
    // staging chunk: ~128 MB of outputs per set (short pipeline fill; the copies stay at PCIe rate: tools/pcie_probe.py)
    const int64_t per = (P.m_nl + P.nnzJ + P.nnzH + P.n_var) * 8;
    int64_t C = std::max<int64_t>(1, std::min<int64_t>(b->S, (int64_t(128) << 20) / std::max<int64_t>(per, 1)));
    if (const char* env = getenv("PS_STAGE")) { const int64_t c = atoll(env); if (c > 0) C = std::min<int64_t>(c, b->S); }
    b->stage_S = C;
    for (int q = 0; q < 2; q++) {
      b->sx[q] = b->mem.alloc<double>((size_t)C * P.n_var); b->smu[q] = b->mem.alloc<double>((size_t)C * P.m_nl);
      b->sg[q] = b->mem.alloc<double>((size_t)C * P.m_nl); b->sJ[q] = b->mem.alloc<double>((size_t)C * P.nnzJ);
      b->sH[q] = b->mem.alloc<double>((size_t)C * P.nnzH); b->sscal[q] = b->mem.alloc<double>((size_t)C * NSCAL);
    }
    if (b->mem.err != cudaSuccess) return cuda_fail(h, b->mem.err, "cudaMalloc(staging)");
    if ((e = cudaStreamCreateWithFlags(&b->st_in, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&b->st_k, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&b->st_out, cudaStreamNonBlocking)) != cudaSuccess)
      return cuda_fail(h, e, "cudaStreamCreate");
    for (int q = 0; q < 2; q++)
      if ((e = cudaEventCreateWithFlags(&b->ev_in[q], cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&b->ev_k[q], cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&b->ev_out[q], cudaEventDisableTiming)) != cudaSuccess)
        return cuda_fail(h, e, "cudaEventCreate");
  }
  const int64_t C = b->stage_S;
  int64_t it = 0;
  // rows that are contiguous on the host go as one linear copy (the driver splits pitched copies per row)
  auto copy_rows = [](void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width, int64_t rows, cudaMemcpyKind kind,
                      cudaStream_t st) {
    if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, (size_t)(width * rows), kind, st);
    return cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)rows, kind, st);
  };
  for (int64_t c0 = 0; c0 < s_count; c0 += C, it++) {
    const int q = (int)(it & 1);
    const int64_t n = std::min<int64_t>(C, s_count - c0);
    // the staging set is free once the copy-out of chunk it-2 is done (inputs: once its kernel is done)
    if (it >= 2) { cudaStreamWaitEvent(b->st_in, b->ev_k[q], 0); }
    e = copy_rows(b->sx[q], P.n_var * 8, x + c0 * ldx, ldx * 8, P.n_var * 8, n, cudaMemcpyHostToDevice, b->st_in);
    if (e == cudaSuccess && (what & EV_H))
      e = copy_rows(b->smu[q], P.m_nl * 8, mu + c0 * ldmu, ldmu * 8, P.m_nl * 8, n, cudaMemcpyHostToDevice, b->st_in);
    if (e != cudaSuccess) return cuda_fail(h, e, "H2D copy");
    cudaEventRecord(b->ev_in[q], b->st_in);
    cudaStreamWaitEvent(b->st_k, b->ev_in[q], 0);
    if (it >= 2) cudaStreamWaitEvent(b->st_k, b->ev_out[q], 0);
    int rc = do_eval(b, what, s_begin + c0, n, b->sx[q], P.n_var, b->smu[q], P.m_nl, b->sg[q], P.m_nl, b->sJ[q], P.nnzJ,
                     b->sH[q], P.nnzH, scalars ? b->sscal[q] : nullptr, b->st_k);
    if (rc != PS_OK) return rc;
    cudaEventRecord(b->ev_k[q], b->st_k);
    cudaStreamWaitEvent(b->st_out, b->ev_k[q], 0);
    if (what & EV_G) e = copy_rows(g + c0 * ldg, ldg * 8, b->sg[q], P.m_nl * 8, P.m_nl * 8, n, cudaMemcpyDeviceToHost, b->st_out);
    if (e == cudaSuccess && (what & EV_J))
      e = copy_rows(J + c0 * ldJ, ldJ * 8, b->sJ[q], P.nnzJ * 8, P.nnzJ * 8, n, cudaMemcpyDeviceToHost, b->st_out);
    if (e == cudaSuccess && (what & EV_H))
      e = copy_rows(H + c0 * ldH, ldH * 8, b->sH[q], P.nnzH * 8, P.nnzH * 8, n, cudaMemcpyDeviceToHost, b->st_out);
    if (e == cudaSuccess && scalars)
      e = cudaMemcpyAsync(scalars + c0 * NSCAL, b->sscal[q], (size_t)n * NSCAL * 8, cudaMemcpyDeviceToHost, b->st_out);
    if (e != cudaSuccess) return cuda_fail(h, e, "D2H copy");
    cudaEventRecord(b->ev_out[q], b->st_out);
  }
  if ((e = cudaStreamSynchronize(b->st_out)) != cudaSuccess || (e = cudaStreamSynchronize(b->st_k)) != cudaSuccess ||
      (e = cudaStreamSynchronize(b->st_in)) != cudaSuccess)
    return cuda_fail(h, e, "pipeline synchronize");
  return PS_OK;
}

int ps_host_alloc(void** p, int64_t bytes) {
  if (!p || bytes < 0) return PS_ERR_ARG;
  cudaError_t e = cudaHostAlloc(p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault);
  if (e != cudaSuccess) { g_create_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); return PS_ERR_CUDA; }
  return PS_OK;
}
void ps_host_free(void* p) { if (p) cudaFreeHost(p); }
int ps_host_register(void* p, int64_t bytes) {
  if (!p || bytes <= 0) return PS_ERR_ARG;
  cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return PS_ERR_CUDA; }
  return PS_OK;
}
int ps_host_unregister(void* p) {
  if (!p) return PS_ERR_ARG;
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); return PS_ERR_CUDA; }
  return PS_OK;
}

// ---- COO -> CSC ----------------------------------------------------------------------------
int ps_csc_create(int64_t m, int64_t n, int64_t nnz_coo, const int64_t* row, const int64_t* col, int device, ps_csc** out) {
  if (!out) return PS_ERR_ARG;
  *out = nullptr;
  if (m < 0 || n < 0 || nnz_coo < 0 || (nnz_coo && (!row || !col))) return fail(nullptr, PS_ERR_ARG, "bad COO arguments");
  if (nnz_coo >= (int64_t(1) << 31)) return fail(nullptr, PS_ERR_ARG, "COO too large");
  for (int64_t q = 0; q < nnz_coo; q++)
    if (row[q] < 1 || row[q] > m || col[q] < 1 || col[q] > n) return fail(nullptr, PS_ERR_ARG, "COO index out of range");
  ps_csc* p = new ps_csc();
  p->m = m; p->n = n; p->nnz_coo = nnz_coo;
  // stable sort by (col, row): duplicates keep COO order, which is the order common.jl:16-18 adds them in
  std::vector<int32_t> ord(nnz_coo);
  std::iota(ord.begin(), ord.end(), 0);
  std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) {
    return col[a] != col[b] ? col[a] < col[b] : row[a] < row[b];
  });
  p->perm = ord;
  p->colptr.assign(n + 1, 0);
  p->segptr.clear();
  for (int64_t q = 0; q < nnz_coo; q++) {
    const int32_t a = ord[q];
    if (q == 0 || col[a] != col[ord[q - 1]] || row[a] != row[ord[q - 1]]) {
      p->segptr.push_back((int32_t)q);
      p->rowval.push_back(row[a]);
    }
  }
  p->segptr.push_back((int32_t)nnz_coo);
  p->nnz = (int64_t)p->rowval.size();
  // SparseMatrixCSC.colptr (1-based): colptr[c] = 1 + number of stored entries in columns 1..c
  {
    std::vector<int64_t> cnt(n + 1, 0);
    for (int64_t j = 0; j < p->nnz; j++) cnt[col[ord[p->segptr[j]]]]++;
    p->colptr[0] = 1;
    for (int64_t c = 1; c <= n; c++) p->colptr[c] = p->colptr[c - 1] + cnt[c];
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { delete p; return fail(nullptr, PS_ERR_CUDA, "no CUDA device available (no CPU fallback)"); }
  if (device < 0) cudaGetDevice(&device);
  if ((e = cudaSetDevice(device)) != cudaSuccess) { delete p; return cuda_fail(nullptr, e, "cudaSetDevice"); }
  p->device = device;
  p->d_segptr = p->mem.upload(p->segptr);
  p->d_perm = p->mem.upload(p->perm);
  p->d_in = p->mem.alloc<double>(nnz_coo);
  p->d_out = p->mem.alloc<double>(p->nnz);
  if (p->mem.err != cudaSuccess) { e = p->mem.err; p->mem.release(); delete p; return cuda_fail(nullptr, e, "csc upload"); }
  *out = p;
  return PS_OK;
}

void ps_csc_destroy(ps_csc* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  p->mem.release();
  delete p;
}

int ps_csc_dims(const ps_csc* p, int64_t* nnz_csc) {
  if (!p || !nnz_csc) return PS_ERR_ARG;
  *nnz_csc = p->nnz;
  return PS_OK;
}

int ps_csc_structure(const ps_csc* p, int64_t* colptr, int64_t* rowval) {
  if (!p || !colptr || !rowval) return PS_ERR_ARG;
  std::copy(p->colptr.begin(), p->colptr.end(), colptr);
  std::copy(p->rowval.begin(), p->rowval.end(), rowval);
  return PS_OK;
}

int ps_csc_assemble(ps_csc* p, const double* dE, double* nzval) {
  if (!p || (p->nnz_coo && !dE) || (p->nnz && !nzval)) return PS_ERR_ARG;
  if (p->nnz == 0) return PS_OK;
  cudaError_t e = cudaSetDevice(p->device);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_in, dE, (size_t)p->nnz_coo * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_csc_assemble(p->d_segptr, p->d_perm, (int)p->nnz, 1, p->d_in, p->nnz_coo, p->d_out, p->nnz, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(nzval, p->d_out, (size_t)p->nnz * 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { p->err = cudaGetErrorString(e); g_create_error = p->err; return PS_ERR_CUDA; }
  return PS_OK;
}

int ps_csc_assemble_batch(ps_csc* p, int64_t S, const double* dE, int64_t ldE, double* nzval, int64_t ldnz, void* stream) {
  if (!p || S < 0 || (S && (!dE || !nzval)) || ldE < p->nnz_coo || ldnz < p->nnz) return PS_ERR_ARG;
  cudaError_t e = cudaSetDevice(p->device);
  if (e == cudaSuccess) e = launch_csc_assemble(p->d_segptr, p->d_perm, (int)p->nnz, S, dE, ldE, nzval, ldnz, (cudaStream_t)stream);
  if (e != cudaSuccess) { p->err = cudaGetErrorString(e); g_create_error = p->err; return PS_ERR_CUDA; }
  return PS_OK;
}

// ---- affine + quadratic rows ---------------------------------------------------------------
int ps_rows_create(int64_t n_var, int64_t m, const int64_t* aff_ptr, const int64_t* aff_var, const double* aff_coef,
                   const int64_t* quad_ptr, const int64_t* quad_var1, const int64_t* quad_var2, const double* quad_coef,
                   const double* constant, int device, ps_rows** out) {
  if (!out) return PS_ERR_ARG;
  *out = nullptr;
  if (n_var <= 0 || m < 0 || (m && (!aff_ptr || !quad_ptr))) return fail(nullptr, PS_ERR_ARG, "bad rows arguments");
  if (n_var >= (int64_t(1) << 31) || m >= (int64_t(1) << 31)) return fail(nullptr, PS_ERR_ARG, "rows block too large");
  const int64_t na = m ? aff_ptr[m] : 0, nq = m ? quad_ptr[m] : 0;
  if (m && (aff_ptr[0] != 0 || quad_ptr[0] != 0)) return fail(nullptr, PS_ERR_ARG, "row pointers must start at 0");
  if ((na && (!aff_var || !aff_coef)) || (nq && (!quad_var1 || !quad_var2 || !quad_coef))) return fail(nullptr, PS_ERR_ARG, "null term arrays");
  for (int64_t r = 0; r < m; r++)
    if (aff_ptr[r + 1] < aff_ptr[r] || quad_ptr[r + 1] < quad_ptr[r]) return fail(nullptr, PS_ERR_ARG, "row pointers must be non-decreasing");
  for (int64_t q = 0; q < na; q++) if (aff_var[q] < 1 || aff_var[q] > n_var) return fail(nullptr, PS_ERR_ARG, "affine variable index out of range");
  for (int64_t q = 0; q < nq; q++)
    if (quad_var1[q] < 1 || quad_var1[q] > n_var || quad_var2[q] < 1 || quad_var2[q] > n_var)
      return fail(nullptr, PS_ERR_ARG, "quadratic variable index out of range");
  ps_rows* R = new ps_rows();
  R->n_var = n_var; R->m = m;
  std::vector<int32_t> aptr(m + 1, 0), qptr(m + 1, 0), avar(na), qv1(nq), qv2(nq), jx, hrow;
  std::vector<double> acoef(aff_coef, aff_coef + na), qcoef(quad_coef, quad_coef + nq), cst(m, 0.0), jc, hc;
  for (int64_t r = 0; r <= m && m; r++) { aptr[r] = (int32_t)aff_ptr[r]; qptr[r] = (int32_t)quad_ptr[r]; }
  for (int64_t q = 0; q < na; q++) avar[q] = (int32_t)(aff_var[q] - 1);
  for (int64_t q = 0; q < nq; q++) { qv1[q] = (int32_t)(quad_var1[q] - 1); qv2[q] = (int32_t)(quad_var2[q] - 1); }
  if (constant) cst.assign(constant, constant + m);
  for (int64_t r = 0; r < m; r++) {
    for (int64_t q = aff_ptr[r]; q < aff_ptr[r + 1]; q++) {                 // MOI_wrapper.jl:780-784, 973-979
      R->jrow.push_back(r + 1); R->jcol.push_back(aff_var[q]);
      jx.push_back(-1); jc.push_back(aff_coef[q]);
    }
    for (int64_t q = quad_ptr[r]; q < quad_ptr[r + 1]; q++) {               // :786-799, 981-1002
      R->jrow.push_back(r + 1); R->jcol.push_back(quad_var1[q]);
      jx.push_back((int32_t)(quad_var2[q] - 1)); jc.push_back(quad_coef[q]);
      if (quad_var1[q] != quad_var2[q]) {
        R->jrow.push_back(r + 1); R->jcol.push_back(quad_var2[q]);
        jx.push_back((int32_t)(quad_var1[q] - 1)); jc.push_back(quad_coef[q]);
      }
      R->hrow.push_back(quad_var1[q]); R->hcol.push_back(quad_var2[q]);     // :834-839, 1030-1042
      hrow.push_back((int32_t)r); hc.push_back(quad_coef[q]);
    }
  }
  R->nnzj = (int64_t)jx.size(); R->nnzh = (int64_t)hc.size();
  if (device == PS_DEVICE_NONE) { R->device = PS_DEVICE_NONE; *out = R; return PS_OK; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { delete R; return fail(nullptr, PS_ERR_CUDA, "no CUDA device available (no CPU fallback)"); }
  if (device < 0) cudaGetDevice(&device);
  if ((e = cudaSetDevice(device)) != cudaSuccess) { delete R; return cuda_fail(nullptr, e, "cudaSetDevice"); }
  R->device = device;
  DevMem& M = R->mem;
  RowsDev& D = R->dev;
  D.m = (int)m; D.nnzj = (int)R->nnzj; D.nnzh = (int)R->nnzh;
  D.aptr = M.upload(aptr); D.avar = M.upload(avar); D.qptr = M.upload(qptr); D.qv1 = M.upload(qv1); D.qv2 = M.upload(qv2);
  D.acoef = M.upload(acoef); D.qcoef = M.upload(qcoef); D.cst = M.upload(cst);
  D.jx = M.upload(jx); D.jc = M.upload(jc); D.hrow = M.upload(hrow); D.hc = M.upload(hc);
  R->d_x = M.alloc<double>(n_var); R->d_lam = M.alloc<double>(m); R->d_g = M.alloc<double>(m);
  R->d_J = M.alloc<double>(R->nnzj); R->d_H = M.alloc<double>(R->nnzh);
  if (M.err != cudaSuccess) { e = M.err; M.release(); delete R; return cuda_fail(nullptr, e, "rows upload"); }
  *out = R;
  return PS_OK;
}

void ps_rows_destroy(ps_rows* r) {
  if (!r) return;
  if (r->device != PS_DEVICE_NONE) { cudaSetDevice(r->device); r->mem.release(); }
  delete r;
}

int ps_rows_dims(const ps_rows* r, int64_t* nnz_jac, int64_t* nnz_hess) {
  if (!r) return PS_ERR_ARG;
  if (nnz_jac) *nnz_jac = r->nnzj;
  if (nnz_hess) *nnz_hess = r->nnzh;
  return PS_OK;
}
int ps_rows_jacobian_structure(const ps_rows* r, int64_t* row, int64_t* col) {
  if (!r || !row || !col) return PS_ERR_ARG;
  std::copy(r->jrow.begin(), r->jrow.end(), row);
  std::copy(r->jcol.begin(), r->jcol.end(), col);
  return PS_OK;
}
int ps_rows_hessian_structure(const ps_rows* r, int64_t* row, int64_t* col) {
  if (!r || !row || !col) return PS_ERR_ARG;
  std::copy(r->hrow.begin(), r->hrow.end(), row);
  std::copy(r->hcol.begin(), r->hcol.end(), col);
  return PS_OK;
}

int ps_rows_eval_batch(ps_rows* r, int what, int64_t S, const double* x, int64_t ldx, const double* lambda, int64_t ldl,
                       double* g, int64_t ldg, double* J, int64_t ldJ, double* H, int64_t ldH, void* stream) {
  if (!r || S < 0) return PS_ERR_ARG;
  if (r->device == PS_DEVICE_NONE) { g_create_error = "structure-only rows block (PS_DEVICE_NONE): no CPU fallback"; return PS_ERR_STATE; }
  if (S == 0) return PS_OK;
  if (!x || ldx < r->n_var || ((what & EV_G) && (!g || ldg < r->m)) || ((what & EV_J) && (!J || ldJ < r->nnzj)) ||
      ((what & EV_H) && r->nnzh && (!H || ldH < r->nnzh || !lambda || ldl < r->m))) {
    g_create_error = "ps_rows_eval_batch: null pointer or leading dimension too small";
    return PS_ERR_ARG;
  }
  cudaError_t e = cudaSetDevice(r->device);
  if (e == cudaSuccess) e = launch_rows_eval(r->dev, what, S, x, ldx, lambda, ldl, g, ldg, J, ldJ, H, ldH, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = std::string("rows kernels: ") + cudaGetErrorString(e); return PS_ERR_CUDA; }
  return PS_OK;
}

int ps_rows_eval(ps_rows* r, int what, const double* x, const double* lambda, double* g, double* J, double* H) {
  if (!r || !x) return PS_ERR_ARG;
  if (r->device == PS_DEVICE_NONE) { g_create_error = "structure-only rows block (PS_DEVICE_NONE): no CPU fallback"; return PS_ERR_STATE; }
  if (((what & EV_G) && !g) || ((what & EV_J) && !J) || ((what & EV_H) && r->nnzh && (!H || !lambda))) return PS_ERR_ARG;
  cudaError_t e = cudaSetDevice(r->device);
  if (e == cudaSuccess) e = cudaMemcpy(r->d_x, x, (size_t)r->n_var * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && (what & EV_H) && r->nnzh) e = cudaMemcpy(r->d_lam, lambda, (size_t)r->m * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = launch_rows_eval(r->dev, what, 1, r->d_x, r->n_var, r->d_lam, r->m, r->d_g, r->m, r->d_J, r->nnzj, r->d_H, r->nnzh, nullptr);
  if (e == cudaSuccess && (what & EV_G) && r->m) e = cudaMemcpy(g, r->d_g, (size_t)r->m * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && (what & EV_J) && r->nnzj) e = cudaMemcpy(J, r->d_J, (size_t)r->nnzj * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && (what & EV_H) && r->nnzh) e = cudaMemcpy(H, r->d_H, (size_t)r->nnzh * 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { g_create_error = std::string("ps_rows_eval: ") + cudaGetErrorString(e); return PS_ERR_CUDA; }
  return PS_OK;
}

// ---- KKT / merit metrics -------------------------------------------------------------------
int ps_kkt_create(ps_csc* pattern, const double* g_L, const double* g_U, const double* x_L, const double* x_U, ps_kkt** out) {
  if (!out) return PS_ERR_ARG;
  *out = nullptr;
  if (!pattern || !g_L || !g_U || !x_L || !x_U) return fail(nullptr, PS_ERR_ARG, "ps_kkt_create: null argument");
  cudaError_t e = cudaSetDevice(pattern->device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
  ps_kkt* k = new ps_kkt();
  k->nnz = pattern->nnz; k->device = pattern->device;
  const int64_t m = pattern->m, n = pattern->n, nnz = pattern->nnz;
  std::vector<int32_t> colptr(n + 1), rowval(nnz), rptr(m + 1, 0), rperm(nnz);
  for (int64_t c = 0; c <= n; c++) colptr[c] = (int32_t)(pattern->colptr[c] - 1);
  for (int64_t q = 0; q < nnz; q++) { rowval[q] = (int32_t)(pattern->rowval[q] - 1); rptr[rowval[q] + 1]++; }
  for (int64_t i = 0; i < m; i++) rptr[i + 1] += rptr[i];
  { std::vector<int32_t> fill(rptr.begin(), rptr.end() - 1);
    for (int64_t q = 0; q < nnz; q++) rperm[fill[rowval[q]]++] = (int32_t)q; }     // ascending column inside a row
  KktDev& D = k->dev;
  D.m = (int)m; D.n = (int)n;
  D.gL = k->mem.upload(vec(g_L, m)); D.gU = k->mem.upload(vec(g_U, m));
  D.xL = k->mem.upload(vec(x_L, n)); D.xU = k->mem.upload(vec(x_U, n));
  D.colptr = k->mem.upload(colptr); D.rowval = k->mem.upload(rowval); D.rptr = k->mem.upload(rptr); D.rperm = k->mem.upload(rperm);
  if (k->mem.err != cudaSuccess) { e = k->mem.err; k->mem.release(); delete k; return cuda_fail(nullptr, e, "kkt upload"); }
  *out = k;
  return PS_OK;
}

void ps_kkt_destroy(ps_kkt* k) {
  if (!k) return;
  cudaSetDevice(k->device);
  k->mem.release();
  delete k;
}

int ps_kkt_eval_batch(ps_kkt* k, int64_t S, const double* E, int64_t ldE, const double* x, int64_t ldx, const double* lambda,
                      int64_t ldl, const double* mult_x_U, const double* mult_x_L, int64_t ldm, const double* df, int64_t lddf,
                      const double* nzval, int64_t ldnz, const double* nu, int64_t ldnu, const double* p, int64_t ldp,
                      const double* f, double* out, void* stream) {
  if (!k || S < 0) return PS_ERR_ARG;
  if (S == 0) return PS_OK;
  const int64_t m = k->dev.m, n = k->dev.n;
  if (!E || !x || !lambda || !mult_x_U || !mult_x_L || !df || !nzval || !out || ldE < m || ldx < n || ldl < m || ldm < n ||
      lddf < n || ldnz < k->nnz || (nu && ldnu < m) || (p && ldp < n)) {
    g_create_error = "ps_kkt_eval_batch: null pointer or leading dimension too small";
    return PS_ERR_ARG;
  }
  KktArgs A{};
  A.E = E; A.x = x; A.lam = lambda; A.mU = mult_x_U; A.mL = mult_x_L; A.df = df; A.nz = nzval; A.nu = nu; A.p = p; A.f = f;
  A.ldE = ldE; A.ldx = ldx; A.ldl = ldl; A.ldm = ldm; A.lddf = lddf; A.ldnz = ldnz; A.ldnu = ldnu; A.ldp = ldp; A.out = out;
  cudaError_t e = cudaSetDevice(k->device);
  if (e == cudaSuccess) e = launch_kkt(k->dev, A, S, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = std::string("kkt kernel: ") + cudaGetErrorString(e); return PS_ERR_CUDA; }
  return PS_OK;
}

}  // extern "C"
