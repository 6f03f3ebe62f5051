// Hand-written sm_100a kernels: evaluation of the NL constraint block g(x), its sparse Jacobian and the Hessian of
// the Lagrangian for every AC formulation of Powersense.jl (math: src/opf/formulations.jl:170-428; closed-form
// derivatives: SURVEY.md Appendix F), batched over scenarios; plus the affine / quadratic rows of the solver-level
// model, the fixed-pattern COO -> CSC assembly and the SLP iteration's KKT / merit metrics.
//
// The path is sparse FP64 and HBM-bound (>= 84 % of the bytes are output writes): no tensor cores.  Outputs are
// scenario-major [S][ld] in JuMP slot order, so what matters is that every store instruction writes whole 32-byte
// sectors of a row.  One evaluation = the bus rows + the branch rows:
//   branch rows   branch_direct_kernel: thread per branch, constants in registers across the scenario loop, next
//                 scenario's inputs prefetched, values kept by compile-time tag and emitted in slot order as 32-byte
//                 sector stores (the only topology dependence of the slot order, f < t, is a per-value select);
//                 branch_staged_kernel: any block size / alignment, through a conflict-free shared staging tile.
//   bus rows      bus_warp_kernel: a warp per tile of <= 32 half-edges; contributions are scattered straight into a
//                 shared-memory image of the tile's J / H / g block (destinations held in registers), multi-contribution
//                 slots and the residuals are ordered sums over contiguous extra cells, the image leaves through TMA;
//                 bus_kernel: its CTA-tile predecessor (cp.async staging, type-major cells, slot maps), kept for
//                 networks with a bus of more than 32 half-edges;
//                 pb_bus_g / pb_bus_jh and cb_bus_g / cb_bus_jh kernels: slot-parallel fast paths of the power- and
//                 current-branch-flow balances (coefficient / index tables, one 32-byte sector per thread).
// No atomics on data (only order-independent ones on the per-scenario scalars); results are bitwise reproducible.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "ps_kernels.cuh"

namespace ps {
namespace {

// one full 32-byte sector per thread: the store shape that keeps HBM writes at copy bandwidth when every
// thread owns a contiguous run of the output (8- and 16-byte stores of such runs are partial-sector writes and
// run ~6x slower: profiles/r01_store_patterns.txt).  Streaming (evict-first): outputs are never re-read.
__device__ __forceinline__ void st_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// V^3 and V^4 of the L4 current limits (formulations.jl:338-343).  JuMP evaluates `^` with a constant exponent other than
// 2 through pow(), i.e. to within an ulp of the exact power, whereas repeated products carry up to two
// roundings -- and on low-impedance branches those powers are multiplied by |y|^2 ~ 1e8.  These are the correctly rounded
// powers (double-double products, the FMA residuals are exact): equal to glibc's pow() in 99.92 % of 2e7 random arguments
// and to the exactly rounded power in all of them (tests/test_gpu_parity.py::test_correctly_rounded_powers).
__device__ __forceinline__ double pow3_cr(double v) {
  const double h = v * v, l = fma(v, v, -h);          // v^2 = h + l
  const double p = h * v, e = fma(h, v, -p);          // h v = p + e
  return p + (e + l * v);
}
__device__ __forceinline__ double pow4_cr(double v) {
  const double h = v * v, l = fma(v, v, -h);
  const double p = h * h, e = fma(h, h, -p);          // h^2 = p + e
  return p + (e + 2.0 * (h * l));
}

// ------------------------------------------------------------------------------- branch rows
// Thread lu of a branch tile stages its rows at sG[lu*NRP + r], sJ[lu*RJP + p], sH[lu*RHP + p];
// the strides are odd so that the 8-byte shared-memory stores of a warp are conflict free.
template <int FORM, int SRC>
struct Branch {
  static constexpr BranchSlots SL = branch_slots(FORM, SRC);
  static constexpr int NR = SL.nrows, RJ = SL.nj, RH = SL.nh;
  static constexpr int NRP = NR | 1, RJP = RJ | 1, RHP = RH | 1;
  static constexpr bool MP = (SRC == SRC_MATPOWER);
  static constexpr BranchPat PAT = branch_pattern(FORM, SRC);
  template <int R> struct Row { static constexpr int J0 = PAT.jtag0(R), H0 = PAT.htag0(R); };
  template <int T> struct JP { static constexpr int p0 = SL.jpos[0][T], p1 = SL.jpos[1][T]; };
  template <int T> struct HP { static constexpr int p0 = SL.hpos[0][T], p1 = SL.hpos[1][T]; };

  struct Out {
    double *g, *j, *h;     // this thread's staging rows
    bool sw, wg, wj, wh;
    double vmax, tol;
    int nviol;
    template <int R> __device__ __forceinline__ void G(double v, bool eq) {
      if (wg) g[R] = v;
      const double vv = eq ? fabs(v) : v;
      vmax = fmax(vmax, vv);
      nviol += (vv > tol) ? 1 : 0;
    }
    template <int T> __device__ __forceinline__ void J(double v) { j[sw ? JP<T>::p1 : JP<T>::p0] = v; }
    template <int T> __device__ __forceinline__ void H(double v) { h[sw ? HP<T>::p1 : HP<T>::p0] = v; }
  };

  // Register writer of the direct path: values are kept by *tag* in registers (all indices are compile-time
  // constants) and emitted in slot order as full 32-byte sectors, st.global.v4.f64, by the owning thread.
  template <int P> struct JI { static constexpr int t0 = SL.jinv[0][P], t1 = SL.jinv[1][P]; };
  template <int P> struct HI { static constexpr int t0 = SL.hinv[0][P], t1 = SL.hinv[1][P]; };
  template <bool WJ, bool WH, bool WG = WJ>
  struct OutR {                      // WJ / WH are compile-time: a J pass and an H pass keep register pressure low
    static constexpr bool wj = WJ, wh = WH;   // (WG alone: the residual-only pass of branch_g_kernel)
    double gt[NR > 0 ? NR : 1], jt[WJ && RJ > 0 ? RJ : 1], ht[WH && RH > 0 ? RH : 1];
    bool sw, track;
    double vmax, tol;
    int nviol;
    template <int R> __device__ __forceinline__ void G(double v, bool eq) {
      if constexpr (WG) {            // g and the violation scalars belong to the J pass
        gt[R] = v;
        const double vv = eq ? fabs(v) : v;
        vmax = fmax(vmax, vv);
        nviol += (vv > tol) ? 1 : 0;
      }
    }
    template <int T> __device__ __forceinline__ void J(double v) { if constexpr (WJ) jt[T] = v; }
    template <int T> __device__ __forceinline__ void H(double v) { if constexpr (WH) ht[T] = v; }
    template <int P> __device__ __forceinline__ double jsel() const {
      if constexpr (JI<P>::t0 == JI<P>::t1) return jt[JI<P>::t0];
      else return sw ? jt[JI<P>::t1] : jt[JI<P>::t0];
    }
    template <int P> __device__ __forceinline__ double hsel() const {
      if constexpr (HI<P>::t0 == HI<P>::t1) return ht[HI<P>::t0];
      else return sw ? ht[HI<P>::t1] : ht[HI<P>::t0];
    }
    template <int P0> __device__ __forceinline__ void storeJ4(double* p) const {
      st_v4(p + P0, jsel<P0>(), jsel<P0 + 1>(), jsel<P0 + 2>(), jsel<P0 + 3>());
    }
    template <int P0> __device__ __forceinline__ void storeH4(double* p) const {
      st_v4(p + P0, hsel<P0>(), hsel<P0 + 1>(), hsel<P0 + 2>(), hsel<P0 + 3>());
    }
    template <int... Q> __device__ __forceinline__ void storeJ(double* p, std::integer_sequence<int, Q...>) const { (storeJ4<4 * Q>(p), ...); }
    template <int... Q> __device__ __forceinline__ void storeH(double* p, std::integer_sequence<int, Q...>) const { (storeH4<4 * Q>(p), ...); }
    // the same values as 16-byte pairs into a (shared-memory) row, for the bulk-copy path
    template <int... Q> __device__ __forceinline__ void rowJ(double* r, std::integer_sequence<int, Q...>) const {
      ((*reinterpret_cast<double2*>(r + 2 * Q) = make_double2(jsel<2 * Q>(), jsel<2 * Q + 1>())), ...);
    }
    template <int... Q> __device__ __forceinline__ void rowH(double* r, std::integer_sequence<int, Q...>) const {
      ((*reinterpret_cast<double2*>(r + 2 * Q) = make_double2(hsel<2 * Q>(), hsel<2 * Q + 1>())), ...);
    }
    // Blocks of 4k+2 doubles per branch (CBRARV / CBRAWV MATPOWER limits, PBRAPV PSS/E limits): the branches of an
    // even/odd lane pair own 8k+4 doubles = whole sectors together.  The even lane writes its first k sectors and the
    // boundary sector (its last two values + the odd lane's first two, fetched with a shuffle); the odd lane writes
    // its remaining k sectors, which start on a sector boundary again.  All lanes of the warp must call this.
    template <bool ISJ, int P> __device__ __forceinline__ double psel() const {
      if constexpr (ISJ) return jsel<P>(); else return hsel<P>();
    }
    template <bool ISJ, int OFF, int... K> __device__ __forceinline__ void pair_sectors(double* p, std::integer_sequence<int, K...>) const {
      (st_v4(p + OFF + 4 * K, psel<ISJ, OFF + 4 * K>(), psel<ISJ, OFF + 4 * K + 1>(), psel<ISJ, OFF + 4 * K + 2>(),
             psel<ISJ, OFF + 4 * K + 3>()), ...);
    }
    template <bool ISJ, int R> __device__ __forceinline__ void store_pair(double* p, bool act, int lane, bool partner_act) const {
      static_assert(R % 4 == 2, "pair mode is for blocks of 4k+2 doubles");
      const double n0 = __shfl_down_sync(0xffffffffu, psel<ISJ, 0>(), 1), n1 = __shfl_down_sync(0xffffffffu, psel<ISJ, 1>(), 1);
      if (!act) return;
      if ((lane & 1) == 0) {
        pair_sectors<ISJ, 0>(p, std::make_integer_sequence<int, (R - 2) / 4>{});
        if (partner_act) st_v4(p + R - 2, psel<ISJ, R - 2>(), psel<ISJ, R - 1>(), n0, n1);
        else { p[R - 2] = psel<ISJ, R - 2>(); p[R - 1] = psel<ISJ, R - 1>(); }
      } else {
        pair_sectors<ISJ, 2>(p, std::make_integer_sequence<int, (R - 2) / 4>{});
      }
    }
  };
  static constexpr bool DIRECT_OK = (RJ % 2 == 0) && (RH % 2 == 0) && RJ > 0 && RH > 0;   // whole sectors per branch or per branch pair

  struct Cst { double gf, bf, Gf, Bf, gt, bt, Gt, Bt, Imax; double d[8]; };   // d: c0..c3 per side
  struct In { double af, at, bf, bt, wf, wt, wr, wi, F0, F1, F2, F3, st; double mu[NR > 0 ? NR : 1]; };

  static __device__ __forceinline__ Cst load_cst(const DevPlan& P, int u) {
    Cst c;
    c.gf = P.gf[u]; c.bf = P.bf[u]; c.Gf = P.Gf[u]; c.Bf = P.Bf[u];
    c.gt = P.gt[u]; c.bt = P.bt[u]; c.Gt = P.Gt[u]; c.Bt = P.Bt[u]; c.Imax = P.Imax[u];
    if constexpr (is_PN(FORM)) {
#pragma unroll
      for (int sd = 0; sd < 2; sd++) {
        c.d[0 + sd] = P.c0[sd * P.nr + u]; c.d[2 + sd] = P.c1[sd * P.nr + u];
        c.d[4 + sd] = P.c2[sd * P.nr + u]; c.d[6 + sd] = P.c3[sd * P.nr + u];
      }
    }
    return c;
  }

  static __device__ __forceinline__ In load_in(const DevPlan& P, const EvalArgs& A, int s, int k, int f, int t) {
    In in{};
    const double* __restrict__ x = A.x + (long long)s * A.ldx;
    in.st = A.status ? (double)A.status[(long long)s * A.ldst + k] : 1.0;
    constexpr bool NEED_V = FORM == PBRAPV || FORM == PBRARV || is_PN(FORM) || (FORM == CBRARV && MP);
    if constexpr (NEED_V) {
      in.af = x[P.o_a + f]; in.at = x[P.o_a + t];
      in.bf = x[P.o_b + f]; in.bt = x[P.o_b + t];
    }
    if constexpr (has_W(FORM)) { in.wf = x[P.o_Wd + f]; in.wt = x[P.o_Wd + t]; }
    if constexpr (has_WrWi(FORM)) { in.wr = x[P.o_Wr + k]; in.wi = x[P.o_Wi + k]; }
    if constexpr (is_PB(FORM) || is_CB(FORM)) {
      in.F0 = x[P.o_f[0] + k]; in.F1 = x[P.o_f[1] + k]; in.F2 = x[P.o_f[2] + k]; in.F3 = x[P.o_f[3] + k];
    }
    if (A.what & EV_H) {
      const double* __restrict__ mu = A.mu + (long long)s * A.ldmu + P.bus_rows + (long long)k * NR;
#pragma unroll
      for (int r = 0; r < NR; r++) in.mu[r] = mu[r];
    } else {
#pragma unroll
      for (int r = 0; r < NR; r++) in.mu[r] = 0.0;
    }
    return in;
  }

  // polar flow-definition row (PBRAPV :302-305 | :366-369): row = flow - T,  T from SURVEY F.1
  // SIDE 0: a = f ; SIDE 1: a = t.  PQ 0: P row, 1: Q row.  cs/sn = cos/sin(th_f - th_t)
  template <int R, int SIDE, int PQ, class O>
  static __device__ __forceinline__ void polar_flow(O& o, double flow, double Vf, double Vt, double cs, double sn,
                                                    double gs, double bs, double G, double B, double mu) {
    constexpr int J0 = Row<R>::J0, H0 = Row<R>::H0;
    const double Va = SIDE ? Vt : Vf, Vb = SIDE ? Vf : Vt;
    const double sd = SIDE ? -sn : sn;                       // sin(th_a - th_b)
    const double A = G * cs + B * sd, Bt = B * cs - G * sd;
    const double VV = Va * Vb;
    double T, dta, dVa, dVb, haa, haVa, haVb, hVaVa, hVaVb;
    if (PQ == 0) {
      T = (-gs) * (Va * Va) - A * Vf * Vt;                   // :302-303
      dta = -Bt * VV; dVa = -2.0 * gs * Va - A * Vb; dVb = -A * Va;
      haa = A * VV; haVa = -Bt * Vb; haVb = -Bt * Va; hVaVa = -2.0 * gs; hVaVb = -A;
    } else {
      T = bs * (Va * Va) + Bt * Vf * Vt;                     // :304-305
      dta = -A * VV; dVa = 2.0 * bs * Va + Bt * Vb; dVb = Bt * Va;
      haa = -Bt * VV; haVa = -A * Vb; haVb = -A * Va; hVaVa = 2.0 * bs; hVaVb = Bt;
    }
    o.template G<R>(flow - T, true);
    // J tags [flow, th_f, th_t, V_f, V_t]
    constexpr int tTA = SIDE ? 2 : 1, tTB = SIDE ? 1 : 2, tVA = SIDE ? 4 : 3, tVB = SIDE ? 3 : 4;
    if (o.wj) {
      o.template J<J0 + 0>(1.0);
      o.template J<J0 + tTA>(-dta);
      o.template J<J0 + tTB>(dta);
      o.template J<J0 + tVA>(-dVa);
      o.template J<J0 + tVB>(-dVb);
    }
    if (o.wh) {
      const double m = -mu;                                  // row = flow - T
      // clique4 tags: (thf,thf)0 (tht,tht)1 (Vf,Vf)2 (Vt,Vt)3 (thf,tht)4 (thf,Vf)5 (thf,Vt)6 (tht,Vf)7 (tht,Vt)8 (Vf,Vt)9
      constexpr int hAA = SIDE ? 1 : 0, hBB = SIDE ? 0 : 1, hVaVa_ = SIDE ? 3 : 2, hVbVb_ = SIDE ? 2 : 3;
      constexpr int hAVa = SIDE ? 8 : 5, hAVb = SIDE ? 7 : 6, hBVa = SIDE ? 6 : 7, hBVb = SIDE ? 5 : 8;
      o.template H<H0 + hAA>(m * haa);
      o.template H<H0 + hBB>(m * haa);
      o.template H<H0 + 4>(-m * haa);
      o.template H<H0 + hAVa>(m * haVa);
      o.template H<H0 + hAVb>(m * haVb);
      o.template H<H0 + hBVa>(-m * haVa);
      o.template H<H0 + hBVb>(-m * haVb);
      o.template H<H0 + hVaVa_>(m * hVaVa);
      o.template H<H0 + 9>(m * hVaVb);
      o.template H<H0 + hVbVb_>(0.0);
    }
  }

  // rectangular flow-definition row (PBRARV :309-312 | :373-376), T from SURVEY F.2
  template <int R, int SIDE, int PQ, class O>
  static __device__ __forceinline__ void rect_flow(O& o, double flow, double ra, double ia, double rb, double ib,
                                                   double gs, double bs, double G, double B, double mu) {
    constexpr int J0 = Row<R>::J0, H0 = Row<R>::H0;
    const double A2 = ra * ra + ia * ia, C = ra * rb + ia * ib, S = ra * ib - ia * rb;
    double T, dra, dia, drb, dib, hd, hC, hrib, hirb;
    if (PQ == 0) {
      T = ((-gs) * A2 - G * C) + B * S;                      // :309-310
      dra = -2.0 * gs * ra - G * rb + B * ib; dia = -2.0 * gs * ia - G * ib - B * rb;
      drb = -G * ra - B * ia; dib = -G * ia + B * ra;
      hd = -2.0 * gs; hC = -G; hrib = B; hirb = -B;
    } else {
      T = (bs * A2 + B * C) + G * S;                         // :311-312
      dra = 2.0 * bs * ra + B * rb + G * ib; dia = 2.0 * bs * ia + B * ib - G * rb;
      drb = B * ra - G * ia; dib = B * ia + G * ra;
      hd = 2.0 * bs; hC = B; hrib = G; hirb = -G;
    }
    o.template G<R>(flow - T, true);
    // J tags [flow, vi_f, vi_t, vr_f, vr_t]
    if (o.wj) {
      o.template J<J0 + 0>(1.0);
      o.template J<J0 + (SIDE ? 2 : 1)>(-dia);
      o.template J<J0 + (SIDE ? 1 : 2)>(-dib);
      o.template J<J0 + (SIDE ? 4 : 3)>(-dra);
      o.template J<J0 + (SIDE ? 3 : 4)>(-drb);
    }
    if (o.wh) {
      const double m = -mu;
      // H tags: (vi_f,vi_f)0 (vi_t,vi_t)1 (vr_f,vr_f)2 (vr_t,vr_t)3 (vr_f,vr_t)4 (vi_f,vi_t)5 (vr_f,vi_t)6 (vi_f,vr_t)7
      o.template H<H0 + (SIDE ? 1 : 0)>(m * hd);             // (ia,ia)
      o.template H<H0 + (SIDE ? 0 : 1)>(0.0);                // (ib,ib)
      o.template H<H0 + (SIDE ? 3 : 2)>(m * hd);             // (ra,ra)
      o.template H<H0 + (SIDE ? 2 : 3)>(0.0);                // (rb,rb)
      o.template H<H0 + 4>(m * hC);                          // (ra,rb)
      o.template H<H0 + 5>(m * hC);                          // (ia,ib)
      o.template H<H0 + (SIDE ? 7 : 6)>(m * hrib);           // (ra,ib)
      o.template H<H0 + (SIDE ? 6 : 7)>(m * hirb);           // (ia,rb)
    }
  }

  // L1: P^2 + Q^2 - Imax^2 [* |V_a|^2]   (:306-307,:313-314,:321-322 | :370-371,:377-378,:385-386)
  // va, vb: V_a (polar) | (vr_a, vi_a) (rect) | Wd_a (W)
  template <int R, class O>
  static __device__ __forceinline__ void limit_L1(O& o, double Pf, double Qf, double va, double vb, double Im2, double mu) {
    constexpr int J0 = Row<R>::J0, H0 = Row<R>::H0;
    const double lhs = Pf * Pf + Qf * Qf;
    double rhs = Im2;
    if constexpr (!MP) {
      if constexpr (FORM == PBRAPV) rhs = Im2 * (va * va);
      else if constexpr (FORM == PBRARV) rhs = Im2 * (va * va + vb * vb);
      else rhs = Im2 * va;
    }
    o.template G<R>(lhs - rhs, false);
    if (o.wj) {
      o.template J<J0 + 0>(2.0 * Pf);
      o.template J<J0 + 1>(2.0 * Qf);
      if constexpr (!MP) {
        if constexpr (FORM == PBRAPV) o.template J<J0 + 2>(-2.0 * Im2 * va);
        else if constexpr (FORM == PBRARV) { o.template J<J0 + 2>(-2.0 * Im2 * va); o.template J<J0 + 3>(-2.0 * Im2 * vb); }
        else o.template J<J0 + 2>(-Im2);
      }
    }
    if (o.wh) {
      o.template H<H0 + 0>(2.0 * mu);
      o.template H<H0 + 1>(2.0 * mu);
      if constexpr (!MP) {
        if constexpr (FORM == PBRAPV) o.template H<H0 + 2>(-2.0 * Im2 * mu);
        else if constexpr (FORM == PBRARV) { o.template H<H0 + 2>(-2.0 * Im2 * mu); o.template H<H0 + 3>(-2.0 * Im2 * mu); }
      }
    }
  }

  // L2 / L3 (CBRARV :328-329 | :392-393 ; CBRAWV :335-336 | :399-400).  a1 = vr_a | Wd_a, a0 = vi_a
  template <int R, class O>
  static __device__ __forceinline__ void limit_CB(O& o, double Ir, double Ii, double a1, double a0, double Im2, double mm) {
    constexpr int J0 = Row<R>::J0, H0 = Row<R>::H0;
    const double I2 = Ir * Ir + Ii * Ii;
    if constexpr (!MP) {
      o.template G<R>(I2 - Im2, false);
      if (o.wj) { o.template J<J0 + 0>(2.0 * Ir); o.template J<J0 + 1>(2.0 * Ii); }
      if (o.wh) { o.template H<H0 + 0>(2.0 * mm); o.template H<H0 + 1>(2.0 * mm); }
    } else if constexpr (FORM == CBRARV) {
      const double r = a1, im = a0, A2 = r * r + im * im;
      o.template G<R>(((Ir * Ir) * A2 + (Ii * Ii) * A2) - Im2, false);
      if (o.wj) { o.template J<J0 + 0>(2.0 * Ir * A2); o.template J<J0 + 1>(2.0 * Ii * A2);
                  o.template J<J0 + 2>(2.0 * r * I2); o.template J<J0 + 3>(2.0 * im * I2); }
      if (o.wh) { o.template H<H0 + 0>(mm * 2.0 * A2); o.template H<H0 + 1>(mm * 2.0 * A2);
                  o.template H<H0 + 2>(mm * 2.0 * I2); o.template H<H0 + 3>(mm * 2.0 * I2);
                  o.template H<H0 + 4>(mm * 4.0 * Ir * r); o.template H<H0 + 5>(mm * 4.0 * Ir * im);
                  o.template H<H0 + 6>(mm * 4.0 * Ii * r); o.template H<H0 + 7>(mm * 4.0 * Ii * im);
                  o.template H<H0 + 8>(0.0); }
    } else {
      const double W = a1;
      o.template G<R>(((Ir * Ir) * W + (Ii * Ii) * W) - Im2, false);
      if (o.wj) { o.template J<J0 + 0>(2.0 * Ir * W); o.template J<J0 + 1>(2.0 * Ii * W); o.template J<J0 + 2>(I2); }
      if (o.wh) { o.template H<H0 + 0>(mm * 2.0 * W); o.template H<H0 + 1>(mm * 2.0 * W); o.template H<H0 + 2>(0.0);
                  o.template H<H0 + 3>(mm * 2.0 * Ir); o.template H<H0 + 4>(mm * 2.0 * Ii); }
    }
  }

  // L4 (PNPAPV / PNRAPV :337-343 | :401-407), SURVEY F.5
  template <int SIDE, class O>
  static __device__ __forceinline__ void limit_L4(O& o, double thf, double tht, double Vf, double Vt, double gs, double bs,
                                                  double G, double B, double y3, double th1, double th2, double Im2, double mm) {
    constexpr int R = SIDE;
    constexpr int J0 = Row<R>::J0, H0 = Row<R>::H0;
    const double y1 = gs * gs + bs * bs, y2 = G * G + B * B;
    const double Va = SIDE ? Vt : Vf, Vb = SIDE ? Vf : Vt;
    const double psi = ((SIDE ? tht - thf : thf - tht) + th1) - th2;
    double s, c;
    sincos(psi, &s, &c);
    double val, dVa, dVb, dta, hVaVa, hVaVb, hVbVb, htt, htVa, htVb;
    if constexpr (MP) {
      const double Va2 = Va * Va, Va3 = pow3_cr(Va), Vb2 = Vb * Vb;    // pow(V, 3), pow(V, 4): see pow3_cr
      val = (y1 * pow4_cr(Va) + y2 * Vb2 * Va2) + y3 * Va3 * Vb * c;
      dVa = 4.0 * y1 * Va3 + 2.0 * y2 * Vb2 * Va + 3.0 * y3 * Va2 * Vb * c;
      dVb = 2.0 * y2 * Vb * Va2 + y3 * Va3 * c;
      dta = -y3 * Va3 * Vb * s;
      hVaVa = 12.0 * y1 * Va2 + 2.0 * y2 * Vb2 + 6.0 * y3 * Va * Vb * c;
      hVaVb = 4.0 * y2 * Va * Vb + 3.0 * y3 * Va2 * c;
      hVbVb = 2.0 * y2 * Va2;
      htt = -y3 * Va3 * Vb * c; htVa = -3.0 * y3 * Va2 * Vb * s; htVb = -y3 * Va3 * s;
    } else {
      val = (y1 * (Va * Va) + y2 * (Vb * Vb)) + y3 * Va * Vb * c;
      dVa = 2.0 * y1 * Va + y3 * Vb * c; dVb = 2.0 * y2 * Vb + y3 * Va * c; dta = -y3 * Va * Vb * s;
      hVaVa = 2.0 * y1; hVaVb = y3 * c; hVbVb = 2.0 * y2;
      htt = -y3 * Va * Vb * c; htVa = -y3 * Vb * s; htVb = -y3 * Va * s;
    }
    o.template G<R>(val - Im2, false);
    // J tags [th_f, th_t, V_f, V_t]
    if (o.wj) { o.template J<J0 + (SIDE ? 1 : 0)>(dta); o.template J<J0 + (SIDE ? 0 : 1)>(-dta);
                o.template J<J0 + (SIDE ? 3 : 2)>(dVa); o.template J<J0 + (SIDE ? 2 : 3)>(dVb); }
    if (o.wh) {
      o.template H<H0 + 0>(mm * htt); o.template H<H0 + 1>(mm * htt); o.template H<H0 + 4>(-mm * htt);
      o.template H<H0 + (SIDE ? 3 : 2)>(mm * hVaVa); o.template H<H0 + (SIDE ? 2 : 3)>(mm * hVbVb);
      o.template H<H0 + 9>(mm * hVaVb);
      o.template H<H0 + (SIDE ? 8 : 5)>(mm * htVa);          // (th_a, V_a)
      o.template H<H0 + (SIDE ? 7 : 6)>(mm * htVb);          // (th_a, V_b)
      o.template H<H0 + (SIDE ? 6 : 7)>(-mm * htVa);         // (th_b, V_a)
      o.template H<H0 + (SIDE ? 5 : 8)>(-mm * htVb);         // (th_b, V_b)
    }
  }

  // clique4 tag of a local pair; locals 0 ra, 1 ia, 2 rb, 3 ib -> roles (AF=0 vi_f, AT=1 vi_t, BF=2 vr_f, BT=3 vr_t)
  static constexpr int ctag(int side, int u, int v) {
    const int m0[4] = {2, 0, 3, 1}, m1[4] = {3, 1, 2, 0};
    int a = side ? m1[u] : m0[u], b = side ? m1[v] : m0[v];
    if (a > b) { const int q = a; a = b; b = q; }
    if (a == b) return a;
    return a == 0 ? 3 + b : (a == 1 ? 5 + b : 9);
  }

  // L5 (PNRARV :344-352 | :408-414), SURVEY F.5
  template <int SIDE, class O>
  static __device__ __forceinline__ void limit_L5(O& o, double ra, double ia, double rb, double ib, double y, double Y,
                                                  double y1, double y2, double Im2, double mm) {
    constexpr int R = SIDE;
    constexpr int J0 = Row<R>::J0, H0 = Row<R>::H0;
    const double A2 = ra * ra + ia * ia, B2 = rb * rb + ib * ib, C = ra * rb + ia * ib, S = ra * ib - ia * rb;
    const double Q = ((y1 * C + y2 * S) + y * A2) + Y * B2;
    const double q0 = y1 * rb + y2 * ib + 2.0 * y * ra, q1 = y1 * ib - y2 * rb + 2.0 * y * ia;
    const double q2 = y1 * ra - y2 * ia + 2.0 * Y * rb, q3 = y1 * ia + y2 * ra + 2.0 * Y * ib;
    // J tags [vi_f, vi_t, vr_f, vr_t]; locals (ra, ia, rb, ib)
    constexpr int t0 = SIDE ? 3 : 2, t1 = SIDE ? 1 : 0, t2 = SIDE ? 2 : 3, t3 = SIDE ? 0 : 1;
    if constexpr (MP) {
      o.template G<R>((((y1 * C * A2 + y2 * S * A2) + y * A2 * A2) + Y * B2 * A2) - Im2, false);
      if (o.wj) {
        o.template J<J0 + t0>(q0 * A2 + Q * (2.0 * ra)); o.template J<J0 + t1>(q1 * A2 + Q * (2.0 * ia));
        o.template J<J0 + t2>(q2 * A2); o.template J<J0 + t3>(q3 * A2);
      }
      if (o.wh) {
        const double a0 = 2.0 * ra, a1 = 2.0 * ia;
        o.template H<H0 + ctag(SIDE, 0, 0)>(mm * (2.0 * y * A2 + 2.0 * q0 * a0 + 2.0 * Q));
        o.template H<H0 + ctag(SIDE, 1, 1)>(mm * (2.0 * y * A2 + 2.0 * q1 * a1 + 2.0 * Q));
        o.template H<H0 + ctag(SIDE, 2, 2)>(mm * (2.0 * Y * A2));
        o.template H<H0 + ctag(SIDE, 3, 3)>(mm * (2.0 * Y * A2));
        o.template H<H0 + ctag(SIDE, 0, 1)>(mm * (q0 * a1 + q1 * a0));
        o.template H<H0 + ctag(SIDE, 0, 2)>(mm * (y1 * A2 + q2 * a0));
        o.template H<H0 + ctag(SIDE, 0, 3)>(mm * (y2 * A2 + q3 * a0));
        o.template H<H0 + ctag(SIDE, 1, 2)>(mm * (-y2 * A2 + q2 * a1));
        o.template H<H0 + ctag(SIDE, 1, 3)>(mm * (y1 * A2 + q3 * a1));
        o.template H<H0 + ctag(SIDE, 2, 3)>(0.0);
      }
    } else {
      o.template G<R>(Q - Im2, false);
      if (o.wj) { o.template J<J0 + t0>(q0); o.template J<J0 + t1>(q1); o.template J<J0 + t2>(q2); o.template J<J0 + t3>(q3); }
      if (o.wh) {
        // H tags: (vi_f,vi_f)0 (vi_t,vi_t)1 (vr_f,vr_f)2 (vr_t,vr_t)3 (vr_f,vr_t)4 (vi_f,vi_t)5 (vr_f,vi_t)6 (vi_f,vr_t)7
        o.template H<H0 + (SIDE ? 1 : 0)>(mm * 2.0 * y);     // (ia,ia)
        o.template H<H0 + (SIDE ? 0 : 1)>(mm * 2.0 * Y);     // (ib,ib)
        o.template H<H0 + (SIDE ? 3 : 2)>(mm * 2.0 * y);     // (ra,ra)
        o.template H<H0 + (SIDE ? 2 : 3)>(mm * 2.0 * Y);     // (rb,rb)
        o.template H<H0 + 4>(mm * y1);                       // (ra,rb)
        o.template H<H0 + 5>(mm * y1);                       // (ia,ib)
        o.template H<H0 + (SIDE ? 7 : 6)>(mm * y2);          // (ra,ib)
        o.template H<H0 + (SIDE ? 6 : 7)>(-mm * y2);         // (ia,rb)
      }
    }
  }

  template <class O>
  static __device__ __forceinline__ void run(O& o, const Cst& c, const In& in) {
    const double st = in.st;
    double Imax = c.Imax * st;
    if (Imax == 0.0) Imax = 10000.0;                          // PowersenseData.jl:206-207
    const double Im2 = Imax * Imax;
    const double gf = c.gf * st, bf = c.bf * st, Gf = c.Gf * st, Bf = c.Bf * st;
    const double gt = c.gt * st, bt = c.bt * st, Gt = c.Gt * st, Bt = c.Bt * st;
    const double* m = in.mu;
    if constexpr (FORM == PBRAPV) {
      const double thf = in.af, tht = in.at, Vf = in.bf, Vt = in.bt;
      double sn, cs;
      sincos(thf - tht, &sn, &cs);
      polar_flow<0, 0, 0>(o, in.F0, Vf, Vt, cs, sn, gf, bf, Gf, Bf, m[0]);
      polar_flow<1, 1, 0>(o, in.F2, Vf, Vt, cs, sn, gt, bt, Gt, Bt, m[1]);
      polar_flow<2, 0, 1>(o, in.F1, Vf, Vt, cs, sn, gf, bf, Gf, Bf, m[2]);
      polar_flow<3, 1, 1>(o, in.F3, Vf, Vt, cs, sn, gt, bt, Gt, Bt, m[3]);
      limit_L1<4>(o, in.F0, in.F1, Vf, 0.0, Im2, m[4]);
      limit_L1<5>(o, in.F2, in.F3, Vt, 0.0, Im2, m[5]);
    } else if constexpr (FORM == PBRARV) {
      const double i_f = in.af, i_t = in.at, r_f = in.bf, r_t = in.bt;
      rect_flow<0, 0, 0>(o, in.F0, r_f, i_f, r_t, i_t, gf, bf, Gf, Bf, m[0]);
      rect_flow<1, 1, 0>(o, in.F2, r_t, i_t, r_f, i_f, gt, bt, Gt, Bt, m[1]);
      rect_flow<2, 0, 1>(o, in.F1, r_f, i_f, r_t, i_t, gf, bf, Gf, Bf, m[2]);
      rect_flow<3, 1, 1>(o, in.F3, r_t, i_t, r_f, i_f, gt, bt, Gt, Bt, m[3]);
      limit_L1<4>(o, in.F0, in.F1, r_f, i_f, Im2, m[4]);
      limit_L1<5>(o, in.F2, in.F3, r_t, i_t, Im2, m[5]);
    } else if constexpr (FORM == PBRAWV || FORM == PNRAWV) {
      const double Wr = in.wr, Wi = in.wi, Wf = in.wf, Wt = in.wt;
      constexpr int J0 = Row<0>::J0, H0 = Row<0>::H0;
      o.template G<0>((Wr * Wr + Wi * Wi) - Wf * Wt, true);     // :316 / :354
      if (o.wj) { o.template J<J0 + 0>(2.0 * Wr); o.template J<J0 + 1>(2.0 * Wi);
                  o.template J<J0 + 2>(-Wt); o.template J<J0 + 3>(-Wf); }
      if (o.wh) { o.template H<H0 + 0>(2.0 * m[0]); o.template H<H0 + 1>(2.0 * m[0]);
                  o.template H<H0 + 2>(0.0); o.template H<H0 + 3>(0.0); o.template H<H0 + 4>(-m[0]); }
      if constexpr (FORM == PBRAWV) {
        limit_L1<1>(o, in.F0, in.F1, Wf, 0.0, Im2, m[1]);
        limit_L1<2>(o, in.F2, in.F3, Wt, 0.0, Im2, m[2]);
      }
    } else if constexpr (FORM == CBRARV) {
      limit_CB<0>(o, in.F0, in.F1, in.bf, in.af, Im2, m[0]);
      limit_CB<1>(o, in.F2, in.F3, in.bt, in.at, Im2, m[1]);
    } else if constexpr (FORM == CBRAWV) {
      limit_CB<0>(o, in.F0, in.F1, in.wf, 0.0, Im2, m[0]);
      limit_CB<1>(o, in.F2, in.F3, in.wt, 0.0, Im2, m[1]);
    } else if constexpr (FORM == PNPAPV || FORM == PNRAPV) {
      const double s2 = st * st;
      // d: [0,1] c0 = |G+jB| per side, [2,3] c1 = angle(G+jB), [4,5] c2 = y3, [6,7] c3 = angle(g_s+jb_s)
      limit_L4<0>(o, in.af, in.at, in.bf, in.bt, gf, bf, Gf, Bf, c.d[4] * s2, c.d[6], c.d[2], Im2, m[0]);
      limit_L4<1>(o, in.af, in.at, in.bf, in.bt, gt, bt, Gt, Bt, c.d[5] * s2, c.d[7], c.d[3], Im2, m[1]);
    } else if constexpr (FORM == PNRARV) {
      const double s2 = st * st;
      // d: [0,1] y, [2,3] Y, [4,5] y1, [6,7] y2
      limit_L5<0>(o, in.bf, in.af, in.bt, in.at, c.d[0] * s2, c.d[2] * s2, c.d[4] * s2, c.d[6] * s2, Im2, m[0]);
      limit_L5<1>(o, in.bt, in.at, in.bf, in.af, c.d[1] * s2, c.d[3] * s2, c.d[5] * s2, c.d[7] * s2, Im2, m[1]);
    }
  }
};

// --------------------------------------------------------------------------- bus rows
// Nodal balances (formulations.jl:170-297).  A bus tile (a contiguous range of buses, i.e. of rows and of g / J / H
// slots) stages, once per CTA, everything that does not depend on the scenario in shared memory: the tile's half-edge
// list (branch, side, other bus, owning bus), the admittance entries of the PN* balances, the generator / shunt lists and
// the slot -> contribution maps.  Per scenario the x entries the tile reads (own voltages, the flow / current variables
// or the neighbour voltages of every half-edge, generator injections), its multipliers and loads are gathered into a
// double-buffered shared array with cp.async one scenario ahead.
//
// Evaluation is CONTRIBUTION-MAJOR and free of divergence on the bus degree:
//   phase 1  half-edge-parallel: every half-edge writes its Jacobian (then Hessian) contributions and the pieces of
//            its flow term to shared memory, in the fixed emission order of the reference's term order
//            (generators, Sij ascending, Sji ascending, shunt terms);
//   phase 2  slot-parallel: every output slot sums its contributions left to right in emission order (duplicate
//            variables of a row: the diagonal block, parallel branches) and goes straight to HBM with warp-contiguous,
//            sector-aligned stores; every bus row folds its pieces in the reference's association.
// No atomics, bitwise reproducible, reference summation order.
struct SeqW {                      // writer of consecutive contributions (cells `stride` apart)
  double* p;
  int stride;
  bool on;
  __device__ __forceinline__ void put(double v) { if (on) *p = v; p += stride; }
  __device__ __forceinline__ void zero() { put(0.0); }          // a structural zero (ps_patterns.h: he_h_zero)
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// TMA bulk copy shared -> global (SASS UBLKCP): the copy engine writes whole lines, the LSU carries no global stores
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, int bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory view of one bus tile
struct BusTile {
  int nbt, nhe, ngt, nsh;            // buses, half-edges, generators, switched shunts in the tile
  int hp0, gp0, sp0, cj0, ch0;       // first global index of each list
  int xh, xg, xw, xs_, xm, xl, nx;   // offsets inside the gathered-input array
  double *cbj, *cbh, *gp, *xb0, *xb1, *hc;   // J / H contribution cells, g pieces [np][nhe], gathered inputs (x2), PN* constants
  int *heks, *hoth, *sgen, *ssh;
  uint16_t *hbus, *jsrc, *hsrc, *mdst, *mptr, *mlist;   // multi-slot arrays: J part then H part
  uint8_t* sst;
};

// per-bus invariants, loaded once per CTA and kept in registers across the scenario loop
struct BusInv {
  int g0, g1, h0, h1, f0, f1;        // tile-local ranges of the bus's generators / half-edges / switched shunts
  int oj, oh;                        // tile-local first cell of the bus's own (generator / shunt) J and H contributions
  double Gl, Bl;
};

// inputs of one half-edge: side, status, the two x entries of the far end (PN*: neighbour voltage) or of the branch
// (flow / current variables), the own voltage (theta_i, V_i) | (vi_i, vr_i), the row multipliers, PN* constants
struct HeIn {
  int side;
  double st, v0, v1, va0, vb0, muP, muQ, c0, c1, c2, c3;
};
struct PieceW {                    // residual pieces of the CTA-tile kernel: [np][nhe]
  double* p;
  int stride;
  __device__ __forceinline__ void operator()(int k, double v) const { p[k * stride] = v; }
};
// One half-edge: contributions in the emission order of ps_plan.cpp (which is the reference's source order) through the
// writers sj / sh, and the pieces of the flow term for the residual through piece(k, value).
template <int FORM, class WJ, class WH, class PC>
__device__ __forceinline__ void he_math(const HeIn& in, WJ& sj, WH& sh, const PC& piece, bool want_pieces) {
  const int side = in.side;
  const double st = in.st, v0 = in.v0, v1 = in.v1, va0 = in.va0, vb0 = in.vb0, muP = in.muP, muQ = in.muQ;
  if constexpr (FORM == PBRAPV || FORM == PBRARV) {            // :188-196
    sj.put(st);
    sj.put(st);
  } else if constexpr (FORM == CBRARV || FORM == CBRAWV) {     // :206-214  T4, v0/v1 = Ir / Ii
    const double Ir = v0, Ii = v1, ra = vb0, ia = va0;
    if (want_pieces) {
      piece(0, st * (ra * Ir)); piece(1, st * (ia * Ii));
      piece(2, st * (ia * Ir)); piece(3, st * (ra * Ii));
    }
    sj.put(st * Ir); sj.put(st * Ii); sj.put(st * ra); sj.put(st * ia);          // P: d/d(ra, ia, Ir, Ii)
    sj.put(-st * Ii); sj.put(st * Ir); sj.put(st * ia); sj.put(-st * ra);        // Q
    sh.zero(); sh.zero(); sh.put(st * muP); sh.zero(); sh.zero(); sh.put(st * muP);
    sh.zero(); sh.zero(); sh.put(st * muQ); sh.zero(); sh.zero(); sh.put(-st * muQ);
  } else if constexpr (FORM == PNPAPV || FORM == PNRAPV) {     // :215-234  T2 / T1, v0/v1 = theta_b / V_b
    const double gs = st * in.c0, bs = st * in.c1;
    const double thb = v0, Vb = v1, Va = vb0, dl = va0 - thb;
    double A, Bt;
    if constexpr (FORM == PNPAPV) {
      const double Y = st * in.c2, phi = in.c3;
      double s, c;
      sincos(dl - phi, &s, &c);
      A = Y * c; Bt = -(Y * s);
    } else {
      const double G = st * in.c2, B = st * in.c3;
      double s, c;
      sincos(dl, &s, &c);
      A = G * c + B * s; Bt = B * c - G * s;
    }
    const double Vf = side ? Vb : Va, Vt = side ? Va : Vb;    // source order V[f]*V[t]
    const double VV = Va * Vb;
    if (want_pieces) {                                         // acc = (acc -/+ p1) -/+ p2
      piece(0, gs * (Va * Va)); piece(1, A * Vf * Vt);
      piece(2, bs * (Va * Va)); piece(3, Bt * Vf * Vt);
    }
    // J: d/d(th_a, th_b, V_a, V_b)
    sj.put(-Bt * VV); sj.put(Bt * VV); sj.put(-2.0 * gs * Va - A * Vb); sj.put(-A * Va);
    sj.put(-A * VV); sj.put(A * VV); sj.put(2.0 * bs * Va + Bt * Vb); sj.put(Bt * Va);
    // H clique order over [th_a, th_b, V_a, V_b]: aa bb VaVa VbVb ab aVa aVb bVa bVb VaVb
    sh.put(muP * A * VV); sh.put(muP * A * VV); sh.put(muP * -2.0 * gs); sh.zero(); sh.put(-muP * A * VV);
    sh.put(-muP * Bt * Vb); sh.put(-muP * Bt * Va); sh.put(muP * Bt * Vb); sh.put(muP * Bt * Va); sh.put(-muP * A);
    sh.put(-muQ * Bt * VV); sh.put(-muQ * Bt * VV); sh.put(muQ * 2.0 * bs); sh.zero(); sh.put(muQ * Bt * VV);
    sh.put(-muQ * A * Vb); sh.put(-muQ * A * Va); sh.put(muQ * A * Vb); sh.put(muQ * A * Va); sh.put(muQ * Bt);
  } else if constexpr (FORM == PNRARV) {                       // :235-243  T3, v0/v1 = vi_b / vr_b
    const double gs = st * in.c0, bs = st * in.c1;
    const double G = st * in.c2, B = st * in.c3;
    const double ra = vb0, ia = va0, rb = v1, ib = v0;
    const double A2 = ra * ra + ia * ia, C = ra * rb + ia * ib, S = ra * ib - ia * rb;
    if (want_pieces) {                                         // P: ((acc - p0) - p1) + p2 ; Q: ((acc + p3) + p4) + p5
      piece(0, gs * A2); piece(1, G * C); piece(2, B * S);
      piece(3, bs * A2); piece(4, B * C); piece(5, G * S);
    }
    // J: d/d(ra, ia, rb, ib)
    sj.put(-2.0 * gs * ra - G * rb + B * ib); sj.put(-2.0 * gs * ia - G * ib - B * rb);
    sj.put(-G * ra - B * ia); sj.put(-G * ia + B * ra);
    sj.put(2.0 * bs * ra + B * rb + G * ib); sj.put(2.0 * bs * ia + B * ib - G * rb);
    sj.put(B * ra - G * ia); sj.put(B * ia + G * ra);
    // H: (ra,ra)(ia,ia)(rb,rb)(ib,ib)(ra,rb)(ia,ib)(ra,ib)(ia,rb)
    sh.put(muP * -2.0 * gs); sh.put(muP * -2.0 * gs); sh.zero(); sh.zero();
    sh.put(-muP * G); sh.put(-muP * G); sh.put(muP * B); sh.put(-muP * B);
    sh.put(muQ * 2.0 * bs); sh.put(muQ * 2.0 * bs); sh.zero(); sh.zero();
    sh.put(muQ * B); sh.put(muQ * B); sh.put(muQ * G); sh.put(-muQ * G);
  }
}

// the CTA-tile kernel's view: inputs from the gathered shared array, contributions to type-major cells
template <int FORM>
__device__ __forceinline__ void he_terms(const BusTile& T, const double* __restrict__ xs, int q, SeqW& sj, SeqW& sh,
                                         bool want_pieces, bool has_status) {
  HeIn in;
  in.side = T.heks[q] & 1;
  const int lb = T.hbus[q];
  in.st = has_status ? (double)T.sst[q] : 1.0;
  in.v0 = xs[T.xh + 2 * q]; in.v1 = xs[T.xh + 2 * q + 1];
  in.va0 = xs[lb]; in.vb0 = xs[T.nbt + lb];                    // own voltage: (theta_i, V_i) | (vi_i, vr_i)
  in.muP = sh.on ? xs[T.xm + 2 * lb] : 0.0; in.muQ = sh.on ? xs[T.xm + 2 * lb + 1] : 0.0;
  if constexpr (is_PN(FORM)) { in.c0 = T.hc[q]; in.c1 = T.hc[T.nhe + q]; in.c2 = T.hc[2 * T.nhe + q]; in.c3 = T.hc[3 * T.nhe + q]; }
  else { in.c0 = in.c1 = in.c2 = in.c3 = 0.0; }
  he_math<FORM>(in, sj, sh, PieceW{T.gp + q, T.nhe}, want_pieces);
}

// Generator and shunt contributions of one bus (everything of the bus that is not a half-edge).
// bss(q) = value of the q-th switched-shunt variable of the bus; sjg / sjs / shs = writers of the generator J entries,
// the shunt J entries and the shunt H entries (emission order of ps_plan.cpp)
struct OwnIn {
  double va0, vb0, W, muP, muQ, Gl, Bl;
  int ngen, nsh;
};
template <int FORM, class BSS, class WJ, class WH>
__device__ __forceinline__ void own_math(const OwnIn& o, const BSS& bssv, WJ& sjg, WJ& sjs, WH& shs) {
  for (int q = 0; q < o.ngen; q++) { sjg.put(1.0); sjg.put(1.0); }     // :177-180
  const double muP = o.muP, muQ = o.muQ;
  const double va0 = o.va0, vb0 = o.vb0;
  const double Gl = o.Gl, Bl = o.Bl;
  if constexpr (is_polar(FORM)) {                                       // :256-263
    const double V = vb0, V2 = V * V;
    sjs.put(-2.0 * Gl * V); sjs.put(2.0 * Bl * V);
    shs.put(-2.0 * Gl * muP); shs.put(2.0 * Bl * muQ);
    for (int q = 0; q < o.nsh; q++) {
      const double bss = bssv(q);
      sjs.put(V2); sjs.put(2.0 * bss * V);
      shs.put(0.0); shs.put(2.0 * bss * muQ); shs.put(2.0 * V * muQ);
    }
  } else if constexpr (FORM == PBRARV || FORM == CBRARV || FORM == PNRARV) {   // :264-271
    const double r = vb0, im = va0, A2 = r * r + im * im;
    sjs.put(-2.0 * Gl * r); sjs.put(-2.0 * Gl * im); sjs.put(2.0 * Bl * r); sjs.put(2.0 * Bl * im);
    shs.put(-2.0 * Gl * muP); shs.put(-2.0 * Gl * muP); shs.put(2.0 * Bl * muQ); shs.put(2.0 * Bl * muQ);
    for (int q = 0; q < o.nsh; q++) {
      const double bss = bssv(q);
      sjs.put(A2); sjs.put(2.0 * bss * r); sjs.put(2.0 * bss * im);
      shs.put(0.0); shs.put(2.0 * bss * muQ); shs.put(2.0 * bss * muQ);
      shs.put(2.0 * r * muQ); shs.put(2.0 * im * muQ); shs.put(0.0);
    }
  } else if constexpr (FORM == CBRAWV) {                                // :272-279
    const double W = o.W;
    sjs.put(-Gl); sjs.put(Bl);
    for (int q = 0; q < o.nsh; q++) {
      const double bss = bssv(q);
      sjs.put(W); sjs.put(bss);
      shs.put(0.0); shs.put(0.0); shs.put(muQ);
    }
  }
}
template <int FORM>
__device__ __forceinline__ void bus_own_terms(const DevPlan& P, const BusTile& T, const BusInv& B, const double* __restrict__ xs,
                                              int lu, SeqW sjg, SeqW sjs, SeqW shs) {
  OwnIn o;
  o.muP = shs.on ? xs[T.xm + 2 * lu] : 0.0; o.muQ = shs.on ? xs[T.xm + 2 * lu + 1] : 0.0;
  o.va0 = xs[lu]; o.vb0 = xs[T.nbt + lu];
  o.W = 0.0;
  if constexpr (FORM == CBRAWV) o.W = xs[T.xw + lu];
  o.Gl = B.Gl; o.Bl = B.Bl;
  o.ngen = B.g1 - B.g0; o.nsh = B.f1 - B.f0;
  own_math<FORM>(o, [&](int q) { return xs[T.xs_ + B.f0 + q]; }, sjg, sjs, shs);
}

// The two residuals of one bus from the staged pieces, in the reference's order and association (:171-292).
// gen(q, pg, qg): injections of the bus's q-th generator; flow(q, st, fp, fq): status and flow variables of its q-th
// half-edge (PB*); piece(q, k): k-th staged piece of its q-th half-edge; bssv(q): switched-shunt variables
template <int FORM, class GEN, class FLOW, class PIECE, class BSS>
__device__ __forceinline__ void residual_math(const OwnIn& o, int deg, double Pd, double Qd, const GEN& gen, const FLOW& flow,
                                              const PIECE& g, const BSS& bssv, double& gP, double& gQ) {
  double accP = 0.0, accQ = 0.0;                               // :171-172
  for (int q = 0; q < o.ngen; q++) {                           // :177-180
    double pg, qg;
    gen(q, pg, qg);
    accP = accP + pg;
    accQ = accQ + qg;
  }
  for (int q = 0; q < deg; q++) {                              // Sij ascending, then Sji ascending
    if constexpr (FORM == PBRAPV || FORM == PBRARV) {
      double st, fp, fq;
      flow(q, st, fp, fq);
      accP = accP + st * fp;
      accQ = accQ + st * fq;
    } else if constexpr (is_CB(FORM)) {
      accP = (accP + g(q, 0)) + g(q, 1);
      accQ = (accQ + g(q, 2)) - g(q, 3);
    } else if constexpr (FORM == PNPAPV || FORM == PNRAPV) {
      accP = (accP - g(q, 0)) - g(q, 1);
      accQ = (accQ + g(q, 2)) + g(q, 3);
    } else {
      accP = ((accP - g(q, 0)) - g(q, 1)) + g(q, 2);
      accQ = ((accQ + g(q, 3)) + g(q, 4)) + g(q, 5);
    }
  }
  const double va0 = o.va0, vb0 = o.vb0;
  if constexpr (is_polar(FORM)) {
    const double V2 = vb0 * vb0;
    accP = accP - o.Gl * V2;
    accQ = accQ + o.Bl * V2;
    for (int q = 0; q < o.nsh; q++) accQ = accQ + bssv(q) * V2;
  } else if constexpr (FORM == PBRARV || FORM == CBRARV || FORM == PNRARV) {
    const double A2 = vb0 * vb0 + va0 * va0;
    accP = accP - o.Gl * A2;
    accQ = accQ + o.Bl * A2;
    for (int q = 0; q < o.nsh; q++) accQ = accQ + bssv(q) * A2;
  } else {
    const double W = o.W;
    accP = accP - o.Gl * W;
    accQ = accQ + o.Bl * W;
    for (int q = 0; q < o.nsh; q++) accQ = accQ + bssv(q) * W;
  }
  gP = accP - Pd;                                              // :291
  gQ = accQ - Qd;                                              // :292
}
template <int FORM>
__device__ __forceinline__ void bus_residuals(const BusTile& T, const BusInv& B, const double* __restrict__ xs, int lu,
                                              bool has_status, double& gP, double& gQ) {
  OwnIn o;
  o.va0 = xs[lu]; o.vb0 = xs[T.nbt + lu];
  o.W = 0.0;
  if constexpr (FORM == CBRAWV) o.W = xs[T.xw + lu];
  o.muP = o.muQ = 0.0;
  o.Gl = B.Gl; o.Bl = B.Bl;
  o.ngen = B.g1 - B.g0; o.nsh = B.f1 - B.f0;
  const int nh = T.nhe;
  residual_math<FORM>(
      o, B.h1 - B.h0, xs[T.xl + lu], xs[T.xl + T.nbt + lu],
      [&](int q, double& pg, double& qg) { pg = xs[T.xg + 2 * (B.g0 + q)]; qg = xs[T.xg + 2 * (B.g0 + q) + 1]; },
      [&](int q, double& st, double& fp, double& fq) {
        st = has_status ? (double)T.sst[B.h0 + q] : 1.0;
        fp = xs[T.xh + 2 * (B.h0 + q)]; fq = xs[T.xh + 2 * (B.h0 + q) + 1];
      },
      [&](int q, int k) { return T.gp[k * nh + B.h0 + q]; }, [&](int q) { return xs[T.xs_ + B.f0 + q]; }, gP, gQ);
}

// issue the cp.async gather of one scenario's inputs of a bus tile
template <int FORM>
__device__ __forceinline__ void bus_gather(const DevPlan& P, const EvalArgs& A, const BusTile& T, const Tile& tl, int s, double* xs) {
  const double* __restrict__ x = A.x + (long long)s * A.ldx;
  const int tid = threadIdx.x;
  for (int q = tid; q < T.nbt; q += BTPB) {
    cp_async8(xs + q, x + P.o_a + tl.u0 + q);
    cp_async8(xs + T.nbt + q, x + P.o_b + tl.u0 + q);
    if constexpr (FORM == CBRAWV) cp_async8(xs + T.xw + q, x + P.o_Wd + tl.u0 + q);
    const double* pd = A.Pd ? A.Pd + (long long)s * A.ldl : P.Pd0;
    const double* qd = A.Qd ? A.Qd + (long long)s * A.ldl : P.Qd0;
    cp_async8(xs + T.xl + q, pd + tl.u0 + q);
    cp_async8(xs + T.xl + T.nbt + q, qd + tl.u0 + q);
  }
  if (A.what & EV_H) {
    const double* __restrict__ mu = A.mu + (long long)s * A.ldmu + 2 * tl.u0;
    for (int q = tid; q < 2 * T.nbt; q += BTPB) cp_async8(xs + T.xm + q, mu + q);
  }
  for (int q = tid; q < T.nhe; q += BTPB) {
    const int ks = T.heks[q], k = ks >> 1, side = ks & 1;
    if constexpr (is_PN(FORM)) {
      const int b = T.hoth[q];
      cp_async8(xs + T.xh + 2 * q, x + P.o_a + b);
      cp_async8(xs + T.xh + 2 * q + 1, x + P.o_b + b);
    } else {
      cp_async8(xs + T.xh + 2 * q, x + P.o_f[0 + 2 * side] + k);
      cp_async8(xs + T.xh + 2 * q + 1, x + P.o_f[1 + 2 * side] + k);
    }
  }
  for (int q = tid; q < T.ngt; q += BTPB) {
    const int g = T.sgen[q];
    cp_async8(xs + T.xg + 2 * q, x + P.o_pg + g);
    cp_async8(xs + T.xg + 2 * q + 1, x + P.o_qg + g);
  }
  for (int q = tid; q < T.nsh; q += BTPB) cp_async8(xs + T.xs_ + q, x + P.o_bss + T.ssh[q]);
}

// ------------------------------------------------------------------------------- helpers
// Stream a staged tile to HBM.  Consecutive threads write consecutive doubles, and the thread -> element map is
// shifted so that every warp-wide store covers whole 32-byte sectors of the destination (partial-sector writes
// are several times slower); only the few elements before the first sector boundary are written separately.
template <int NTHR, class F>
__device__ __forceinline__ void copy_aligned(double* __restrict__ dst, int n, int tid, F get) {
  const int a = min(n, (int)((4 - ((reinterpret_cast<unsigned long long>(dst) >> 3) & 3)) & 3));
  if (tid < a) dst[tid] = get(tid);
  for (int e = a + tid; e < n; e += NTHR) dst[e] = get(e);
}
// same, one whole 32-byte sector (4 doubles) per thread and store: 4x fewer store instructions
template <int NTHR, class F>
__device__ __forceinline__ void copy_sectors(double* __restrict__ dst, int n, int tid, F get) {
  const int a = min(n, (int)((4 - ((reinterpret_cast<unsigned long long>(dst) >> 3) & 3)) & 3));
  const int nfull = (n - a) >> 2, tail = a + 4 * nfull;
  for (int q = tid; q < nfull; q += NTHR) {
    const int e = a + 4 * q;
    st_v4(dst + e, get(e), get(e + 1), get(e + 2), get(e + 3));
  }
  if (tid < a) dst[tid] = get(tid);
  if (tid < n - tail) dst[tail + tid] = get(tail + tid);
}
template <int N, int NP>
__device__ __forceinline__ void copy_out_padded(double* __restrict__ dst, const double* tile, int n) {
  copy_sectors<TPB>(dst, n, threadIdx.x, [&](int e) { return tile[e + (e / N) * (NP - N)]; });   // N: compile-time
}
__device__ __forceinline__ void copy_out(double* __restrict__ dst, const double* tile, int n) {
  copy_aligned<TPB>(dst, n, threadIdx.x, [&](int e) { return tile[e]; });
}

// per-scenario violation scalars of one warp: max over the lanes through two 32-bit REDUX (non-negative doubles order like
// their bit patterns: high word first, then the low words of the lanes that hold the maximal high word) and the count
// through a third -- three warp-wide instructions where a shuffle tree takes fifteen (on the residual-only pass the
// shuffle version was a quarter of the branch kernel's time: profiles/r02_gonly_notes.txt)
__device__ __forceinline__ void reduce_scalars(double* scal, double vmax, int nviol) {
  if (!scal) return;
  if (!(vmax >= 0.0)) vmax = 0.0;                               // NaN is ignored, like fmax does
  const unsigned long long b = (unsigned long long)__double_as_longlong(vmax);
  const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(b >> 32));
  const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(b >> 32) == hi ? (unsigned)b : 0u);
  const int nv = __reduce_add_sync(0xffffffffu, nviol);
  if ((threadIdx.x & 31) == 0) {
    const unsigned long long m = ((unsigned long long)hi << 32) | lo;
    if (m) atomicMax((unsigned long long*)(scal + 1), m);
    if (nv) atomicAdd(scal + 2, (double)nv);                    // small integers: exact, order independent
  }
}

// objective of formulations.jl:122-138: sum(cg) (obj_type = :linear) or the polynomial cost in pg.
// The summation order is fixed independently of the CTA shape, so that every kernel that produces the scalar gives the same
// bits: 1024 virtual lanes, lane t sums the generators t, t + 1024, ... in order; each group of 32 consecutive lanes is
// folded by a xor-shuffle tree (16, 8, 4, 2, 1); the 32 group sums are added in order.
constexpr int OBJ_LANES = 1024;
__device__ __forceinline__ double objective_term(const DevPlan& P, const double* __restrict__ x, int i) {
  if (P.flags & FLAG_OBJ_LINEAR) return x[P.o_cg + i];
  const double p = x[P.o_pg + i];
  double v = P.k0 ? P.k0[i] : 0.0;
  if (P.cost_order >= 2 && P.k1) v += P.k1[i] * p;
  if (P.cost_order == 3 && P.k2) v += P.k2[i] * (p * p);
  return v;
}
__device__ __forceinline__ double objective_lane(const DevPlan& P, const double* __restrict__ x, int t) {
  double acc = 0.0;
  for (int i = t; i < P.ng; i += OBJ_LANES) acc += objective_term(P, x, i);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  return acc;                                                  // every lane of the group holds the group sum
}
__device__ __forceinline__ void objective_tile(const DevPlan& P, const EvalArgs& A, int s0, int s1, double* red) {
  static_assert(OBJ_LANES % TPB == 0, "virtual lanes are dealt to the CTA's threads in whole rounds");
  for (int s = s0; s < s1; s++) {
    const double* __restrict__ x = A.x + (long long)s * A.ldx;
#pragma unroll
    for (int j = 0; j < OBJ_LANES / TPB; j++) {                // thread tid plays the virtual lanes tid + TPB * j
      const int t = threadIdx.x + TPB * j;
      const double gsum = objective_lane(P, x, t);
      if ((threadIdx.x & 31) == 0) red[t >> 5] = gsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < OBJ_LANES / 32; w++) tot += red[w];
      A.scal[(long long)s * NSCAL + 0] = tot;
    }
    __syncthreads();
  }
}

// Staged path of the branch rows (any row size, any alignment): every thread evaluates its branch into a
// shared-memory tile laid out like the output (odd per-thread strides: conflict-free), then the CTA streams the
// tile to HBM with warp-contiguous, sector-aligned stores.
template <int FORM, int SRC>
__global__ void __launch_bounds__(TPB) branch_staged_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  extern __shared__ double smem[];
  using BR = Branch<FORM, SRC>;
  const Tile tl = P.tiles[P.n_bus_tiles + blockIdx.x];
  const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
  const bool wg = (A.what & EV_G) != 0, wj = (A.what & EV_J) != 0, wh = (A.what & EV_H) != 0;
  const int u = tl.u0 + threadIdx.x;
  const bool act = u < tl.u1;
  double* sG = smem;
  double* sJ = sG + TPB * BR::NRP;
  double* sH = sJ + TPB * BR::RJP;
  typename BR::Out o;
  o.g = sG + threadIdx.x * BR::NRP; o.j = sJ + threadIdx.x * BR::RJP; o.h = sH + threadIdx.x * BR::RHP;
  o.wg = wg; o.wj = wj; o.wh = wh; o.tol = A.viol_tol;
  typename BR::Cst cst{};
  typename BR::In cur{};
  int f = 0, t = 1;
  if (act) {
    f = P.f[u]; t = P.t[u];
    cst = BR::load_cst(P, u);
    cur = BR::load_in(P, A, s0, u, f, t);
  }
  o.sw = f > t;
  for (int s = s0; s < s1; s++) {
    typename BR::In nxt{};
    if (act && s + 1 < s1) nxt = BR::load_in(P, A, s + 1, u, f, t);   // prefetch the next scenario's inputs
    o.vmax = 0.0; o.nviol = 0;
    if (act) BR::run(o, cst, cur);
    __syncthreads();
    if (wg) copy_out_padded<BR::NR, BR::NRP>(A.g + (long long)s * A.ldg + tl.g0, sG, tl.gn);
    if (wj) copy_out_padded<BR::RJ, BR::RJP>(A.J + (long long)s * A.ldJ + tl.j0, sJ, tl.jn);
    if (wh) copy_out_padded<BR::RH, BR::RHP>(A.H + (long long)s * A.ldH + tl.h0, sH, tl.hn);
    reduce_scalars(A.scal ? A.scal + (long long)s * NSCAL : nullptr, o.vmax, o.nviol);
    __syncthreads();
    if (s + 1 < s1) cur = nxt;
  }
}

// Slots with several contributions (diagonal block of the nodal balances, parallel branches) are pre-reduced into the
// cell of their first contribution, left to right in emission order; the list is sorted by contribution count so the
// lanes of a warp run the same number of additions.  After that every slot has exactly one source cell.
__device__ __forceinline__ void reduce_multi(double* cb, const uint16_t* mdst, const uint16_t* mptr, const uint16_t* mlist, int mn) {
  for (int m = threadIdx.x; m < mn; m += BTPB) {
    const int d = mdst[m];
    double v = cb[d];
    for (int q = mptr[m], q1 = mptr[m + 1]; q < q1; q++) v = v + cb[mlist[q]];
    cb[d] = v;
  }
}
__device__ __forceinline__ void slots_out(double* __restrict__ dst, int n, const double* cb, const uint16_t* src) {
  copy_aligned<BTPB>(dst, n, threadIdx.x, [&](int e) { return cb[src[e]]; });
}

// Bus rows (nodal balances, formulations.jl:170-297) plus, in the last CTA column, the objective scalar.
#ifndef PS_BUS_MINB
#define PS_BUS_MINB 5
#endif
// threads per bus-tile CTA (the plan sizes the bus tiles for it: ps_capi.cu passes BTPB to build_plan)
// FULL = the whole evaluation (g + J + H) without a status mask: the output flags are compile-time constants there,
// which keeps the hot instantiation free of predicated stores and of the loop versions the compiler otherwise clones
template <int FORM, bool FULL>
__device__ __forceinline__ void bus_tile_body(const DevPlan& P, const EvalArgs& A, double* smem, int s0, int s1) {
  if constexpr (has_bus_rows(FORM)) {
    constexpr int NJ = he_nj(FORM), NH = he_nh(FORM), NP = he_np(FORM);
    const Tile tl = P.tiles[P.tile_map ? P.tile_map[blockIdx.x] : blockIdx.x];
    const bool wg = FULL || (A.what & EV_G) != 0, wj = FULL || (A.what & EV_J) != 0, wh = FULL || (A.what & EV_H) != 0;
    const int tid = threadIdx.x;
    // bus-owner threads are taken from the END of the CTA: the half-edge phase fills the first warps, so the per-bus
    // work (own terms, residuals) runs on warps that would otherwise idle at the barrier instead of lengthening warp 0
    const int lb = BTPB - 1 - tid;                  // tile-local bus of this thread
    const int u = tl.u0 + lb;
    const bool act = u < tl.u1;
    const bool facts = (P.flags & FLAG_FACTS) != 0;
    BusTile T;
    T.nbt = tl.u1 - tl.u0;
    T.hp0 = P.he_ptr[tl.u0]; T.nhe = P.he_ptr[tl.u1] - T.hp0;
    T.gp0 = P.gen_ptr[tl.u0]; T.ngt = P.gen_ptr[tl.u1] - T.gp0;
    T.sp0 = P.sh_ptr[tl.u0]; T.nsh = facts ? P.sh_ptr[tl.u1] - T.sp0 : 0;
    T.cj0 = P.bus_jpp[tl.u0]; T.ch0 = P.bus_hpp[tl.u0];
    const int ncj = P.bus_jpp[tl.u1] - T.cj0, nch = P.bus_hpp[tl.u1] - T.ch0;
    // gathered-input layout: own a | own b | half-edge pairs | generator pairs | Wd | bss | mu (P,Q) | Pd | Qd
    T.xh = 2 * T.nbt; T.xg = T.xh + 2 * T.nhe; T.xw = T.xg + 2 * T.ngt;
    T.xs_ = T.xw + (FORM == CBRAWV ? T.nbt : 0); T.xm = T.xs_ + T.nsh; T.xl = T.xm + 2 * T.nbt; T.nx = T.xl + 2 * T.nbt;
    T.cbj = smem;
    T.cbh = T.cbj + ncj;
    T.gp = T.cbh + nch;
    T.xb0 = T.gp + NP * T.nhe; T.xb1 = T.xb0 + T.nx;
    T.hc = T.xb1 + T.nx;
    T.heks = (int*)(T.hc + (is_PN(FORM) ? 4 * T.nhe : 0));
    T.hoth = T.heks + T.nhe;
    T.sgen = T.hoth + (is_PN(FORM) ? T.nhe : 0);
    T.ssh = T.sgen + T.ngt;
    T.hbus = (uint16_t*)(T.ssh + T.nsh);
    T.jsrc = T.hbus + T.nhe;
    T.hsrc = T.jsrc + tl.jn;
    T.mdst = T.hsrc + tl.hn;                      // [mjn | mhn]
    T.mptr = T.mdst + (tl.mjn + tl.mhn);          // [mjn + 1 | mhn + 1]
    T.mlist = T.mptr + (tl.mjn + tl.mhn + 2);     // [mjln | mhln]
    T.sst = (uint8_t*)(T.mlist + (tl.mjln + tl.mhln));
    // ---- scenario-independent staging (once per CTA)
    for (int q = tid; q < T.nhe; q += BTPB) {
      T.heks[q] = P.he_ks[T.hp0 + q];
      T.hbus[q] = (uint16_t)(P.he_bus[T.hp0 + q] - tl.u0);
      if constexpr (is_PN(FORM)) {
        T.hoth[q] = P.he_other[T.hp0 + q];
#pragma unroll
        for (int c = 0; c < 4; c++) T.hc[c * T.nhe + q] = P.he_c[(long long)c * P.nhe + T.hp0 + q];
      }
    }
    for (int q = tid; q < T.ngt; q += BTPB) T.sgen[q] = P.gen_idx[T.gp0 + q];
    for (int q = tid; q < T.nsh; q += BTPB) T.ssh[q] = P.sh_idx[T.sp0 + q];
    for (int q = tid; q < tl.jn; q += BTPB) T.jsrc[q] = (uint16_t)P.jsrc[tl.j0 + q];
    for (int q = tid; q < tl.hn; q += BTPB) T.hsrc[q] = (uint16_t)P.hsrc[tl.h0 + q];
    for (int q = tid; q < tl.mjn; q += BTPB) T.mdst[q] = (uint16_t)P.ms_dst[tl.mj0 + q];
    for (int q = tid; q < tl.mhn; q += BTPB) T.mdst[tl.mjn + q] = (uint16_t)P.ms_dst[tl.mh0 + q];
    for (int q = tid; q <= tl.mjn; q += BTPB) T.mptr[q] = (uint16_t)P.ms_ptr[tl.mj0 + q];
    for (int q = tid; q <= tl.mhn; q += BTPB) T.mptr[tl.mjn + 1 + q] = (uint16_t)P.ms_ptr[tl.mh0 + q];
    for (int q = tid; q < tl.mjln; q += BTPB) T.mlist[q] = (uint16_t)P.ms_list[tl.mjl0 + q];
    for (int q = tid; q < tl.mhln; q += BTPB) T.mlist[tl.mjln + q] = (uint16_t)P.ms_list[tl.mhl0 + q];
    BusInv B{};
    if (act) {
      B.g0 = P.gen_ptr[u] - T.gp0; B.g1 = P.gen_ptr[u + 1] - T.gp0;
      B.h0 = P.he_ptr[u] - T.hp0; B.h1 = P.he_ptr[u + 1] - T.hp0;
      B.f0 = P.sh_ptr[u] - T.sp0; B.f1 = facts ? P.sh_ptr[u + 1] - T.sp0 : B.f0;
      B.oj = NJ * T.nhe + (P.bus_jpp[u] - T.cj0) - NJ * B.h0;
      B.oh = NH * T.nhe + (P.bus_hpp[u] - T.ch0) - NH * B.h0;
      B.Gl = P.Gl[u]; B.Bl = P.Bl[u];
    }
    const bool has_status = !FULL && A.status != nullptr;
    const bool need_g = wg || A.scal;
    __syncthreads();
    bus_gather<FORM>(P, A, T, tl, s0, T.xb0);
    cp_async_commit();
    for (int s = s0; s < s1; s++) {
      const bool odd = ((s - s0) & 1) != 0;
      double* const xs = odd ? T.xb1 : T.xb0;
      if (s + 1 < s1) bus_gather<FORM>(P, A, T, tl, s + 1, odd ? T.xb0 : T.xb1);   // one scenario ahead
      cp_async_commit();
      if (has_status) {
        const uint8_t* __restrict__ stp = A.status + (long long)s * A.ldst;
        for (int q = tid; q < T.nhe; q += BTPB) T.sst[q] = stp[T.heks[q] >> 1];
      }
      cp_async_wait<1>();                                                          // this scenario's gather has landed
      __syncthreads();
      // ---- phase 1: half-edge-parallel contributions (J and H cells, residual pieces); the bus's own terms
      for (int q = tid; q < T.nhe; q += BTPB) {
        SeqW sj{T.cbj + q, T.nhe, wj}, sh{T.cbh + q, T.nhe, wh};
        he_terms<FORM>(T, xs, q, sj, sh, need_g, has_status);
      }
      if (act && (wj || wh)) {
        const int ng2 = 2 * (B.g1 - B.g0);
        bus_own_terms<FORM>(P, T, B, xs, lb, SeqW{T.cbj + B.oj, 1, wj}, SeqW{T.cbj + B.oj + ng2, 1, wj}, SeqW{T.cbh + B.oh, 1, wh});
      }
      __syncthreads();
      // ---- phase 1.5: slots with several contributions
      if (wj) reduce_multi(T.cbj, T.mdst, T.mptr, T.mlist, tl.mjn);
      if (wh) reduce_multi(T.cbh, T.mdst + tl.mjn, T.mptr + tl.mjn + 1, T.mlist + tl.mjln, tl.mhn);
      if (wj || wh) __syncthreads();
      // ---- phase 2: slot-parallel output, residual rows
      if (wj) slots_out(A.J + (long long)s * A.ldJ + tl.j0, tl.jn, T.cbj, T.jsrc);
      if (wh) slots_out(A.H + (long long)s * A.ldH + tl.h0, tl.hn, T.cbh, T.hsrc);
      double vmax = 0.0;
      int nviol = 0;
      if (need_g && act) {
        double gP, gQ;
        bus_residuals<FORM>(T, B, xs, lb, has_status, gP, gQ);
        if (wg) {
          double* gd = A.g + (long long)s * A.ldg + tl.g0 + 2 * lb;
          if ((reinterpret_cast<unsigned long long>(gd) & 15) == 0) *reinterpret_cast<double2*>(gd) = make_double2(gP, gQ);
          else { gd[0] = gP; gd[1] = gQ; }
        }
        vmax = fmax(fabs(gP), fabs(gQ));
        nviol = (fabs(gP) > A.viol_tol ? 1 : 0) + (fabs(gQ) > A.viol_tol ? 1 : 0);
      }
      reduce_scalars(A.scal ? A.scal + (long long)s * NSCAL : nullptr, vmax, nviol);
      __syncthreads();
    }
  }
}

template <int FORM>
__global__ void __launch_bounds__(BTPB, PS_BUS_MINB) bus_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  extern __shared__ double smem[];
  const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
  if ((A.what & (EV_G | EV_J | EV_H)) == (EV_G | EV_J | EV_H) && A.status == nullptr) bus_tile_body<FORM, true>(P, A, smem, s0, s1);
  else bus_tile_body<FORM, false>(P, A, smem, s0, s1);
}

// ---- bus rows, warp tiles ---------------------------------------------------------------------------------------
// The same nodal balances, restructured so that a WARP owns a small tile (<= 32 half-edges, ps_plan.h: WTile) and
// nothing is shared between warps: no CTA barrier, no cp.async staging, one shared-memory store per contribution.
// Everything is a slot that sums its contributions left to right in emission order -- the J and H entries and also
// the two residuals of a bus, written as flat sums of signed terms (x - y == x + (-y) exactly) in the reference's
// order: generator injections, the pieces of every half-edge, shunt terms, - load.  Per scenario
//   1. lane = half-edge: inputs straight from global memory (prefetched one scenario ahead in registers); he_math
//      writes every contribution to its DESTINATION: the slot's own position in a shared-memory image of the tile's
//      J / H / g block when it is the slot's first contribution, an extra cell behind the image otherwise.  The
//      destinations (uint16, fixed per lane) live in registers for the whole scenario loop;
//   2. lane = bus: generator and shunt contributions, the load terms;
//   3. lane = multi-contribution slot (lists sorted by count): slot += its extra cells, which are contiguous;
//   4. the image *is* the output: one cp.async.bulk (TMA) per J / H block hands it to the copy engine, which writes
//      whole lines to HBM while the warp evaluates the next scenario.  An odd first slot is matched by shifting the
//      image by one double, so that image and destination share their 16-byte phase (any other layout: a coalesced
//      copy loop).  The residual rows are stored, and folded into the violation scalars, by the lanes.
// one warp per CTA: the warps share nothing, and a CTA of several tiles holds its registers and shared memory until its
// slowest tile is done (4 warps per CTA: 3 % slower)
#ifndef PS_WT_WPC
#define PS_WT_WPC 1
#endif
constexpr int WPC = PS_WT_WPC;     // warps (= tiles) per CTA

struct RegW {                      // writer through register-held destinations (two uint16 per register)
  double* base;
  const uint32_t* d;
  int k;
  bool on;
  __device__ __forceinline__ void put(double v) {
    if (on) base[(d[k >> 1] >> ((k & 1) * 16)) & 0xffffu] = v;
    k++;
  }
  __device__ __forceinline__ void zero() { k++; }               // structural zeros are never written (the plan drops them)
};
struct NullW {                     // a block that is not asked for: its values are never formed (compile-time dead)
  static constexpr bool on = false;
  __device__ __forceinline__ void put(double) {}
  __device__ __forceinline__ void zero() {}
};
struct ListW {                     // writer through a destination list in shared memory
  double* base;
  const uint16_t* d;
  bool on;
  __device__ __forceinline__ void put(double v) {
    if (on) base[*d] = v;
    d++;
  }
};
// slot += its extra cells, left to right; lane per multi-contribution slot of any of the three spaces (one list,
// sorted by count, so the lanes of a round run about the same number of additions).  The descriptors are decoded ONCE per
// CTA (wt_predecode: 32-bit shared-memory addresses of the slot and of its first extra cell, count in the top bits; slots
// of a block that is not being computed get count 0), so a visit is one 8-byte load, the additions and one store --
// decoding the plan's packed form and rebuilding the shared-window addresses on every visit was 30 of the ~55
// instructions of a visit, and the visits 31 % of the kernel's instructions (profiles/r02_ncu_final_config3_summary.txt)
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ uint64_t wt_predecode(uint64_t w, const double* imgJ, const double* imgH, const double* imgG, bool wj, bool wh,
                                                 bool wgg) {
  const int kind = (int)(w >> 14) & 3;
  const bool on = kind == 0 ? wj : (kind == 1 ? wh : wgg);
  const unsigned base = (unsigned)__cvta_generic_to_shared(kind == 0 ? imgJ : (kind == 1 ? imgH : imgG));
  const unsigned dst = base + 8u * ((unsigned)w & 0x3fffu), src = base + 8u * (((unsigned)w >> 16) & 0xffffu);
  const unsigned n = on ? (unsigned)(w >> 32) & 0xfffu : 0u;
  return (uint64_t)dst | ((uint64_t)(src | (n << 20)) << 32);
}
__device__ __forceinline__ void reduce_multi_warp(const uint64_t* multi, int mn, int lane) {
  const unsigned mb = (unsigned)__cvta_generic_to_shared(multi);
  for (int m = lane; m < mn; m += 32) {
    unsigned dst, sn;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(dst), "=r"(sn) : "r"(mb + 8u * (unsigned)m) : "memory");
    const unsigned src = sn & 0xfffffu;
    const int n = (int)(sn >> 20);
    if (n == 0) continue;
    double v = lds_f64(dst);
    unsigned p = src;
    int q = n;
    for (; q >= 4; q -= 4, p += 32u) {                          // four loads in flight, then the additions in order
      const double a0 = lds_f64(p), a1 = lds_f64(p + 8u), a2 = lds_f64(p + 16u), a3 = lds_f64(p + 24u);
      v = (((v + a0) + a1) + a2) + a3;
    }
    if (q >= 2) { const double a0 = lds_f64(p), a1 = lds_f64(p + 8u); v = (v + a0) + a1; p += 16u; q -= 2; }
    if (q) v = v + lds_f64(p);
    sts_f64(dst, v);
  }
}
// hand a finished block image to the copy engine: [head][mid doubles, 16-byte aligned on both sides][tail]; when the
// 16-byte phases of image and destination differ (odd leading dimension) the warp copies it
struct WtBlock {
  unsigned simg;         // shared-memory address of the first bulk-copied double
  int head, mid, n;      // 0 / 1 leading doubles stored by a lane, bulk-copied doubles, block length
  bool bulk;
};
__device__ __forceinline__ WtBlock wt_block(const double* img, int n, bool bulk) {
  WtBlock b;
  b.head = (int)(((unsigned)__cvta_generic_to_shared(img) >> 3) & 1u);
  b.mid = (n - b.head) & ~1;
  b.n = n;
  b.simg = (unsigned)__cvta_generic_to_shared(img + b.head);
  b.bulk = bulk;
  return b;
}
__device__ __forceinline__ void wt_store_block(double* __restrict__ g, const double* img, const WtBlock& b, int lane) {
  if (b.bulk) {
    if (lane == 0 && b.mid > 0)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + b.head), "r"(b.simg), "r"(b.mid * 8) : "memory");
    if (lane == 1 && b.head && b.n > 0) g[0] = img[0];
    if (lane == 2 && b.head + b.mid < b.n) g[b.n - 1] = img[b.n - 1];
  } else {
    for (int e = lane; e < b.n; e += 32) g[e] = img[e];
  }
}
__device__ __forceinline__ void reduce_scalars_redux(double* scal, double vmax, int nviol) { reduce_scalars(scal, vmax, nviol); }

// MODE: which outputs are compile-time constants.  WT_ANY: read from A.what / A.status at run time (any combination);
// WT_FULL: g + J + H without a mask (the headline evaluation); WT_GJ: g + J -- the SLP's LP set-up: no Hessian term is formed.
// (Residuals alone -- eval_constraint, every trial point of the SLP line search -- go to bus_g_list_kernel.)
// WT_ST added to one of the three specialised kinds: the same with a branch-status mask (N-1 sweeps: config 5 evaluates,
// and its SLP line search re-evaluates, under a mask -- the run-time-flag instantiation costs those calls 10-30 %)
enum WtMode : int { WT_ANY = 0, WT_FULL = 1, WT_GJ = 3, WT_ST = 4 };
template <int FORM, int MODE>
__device__ __forceinline__ void bus_warp_body(const DevPlan& P, const EvalArgs& A, double* region, int tile, int s0, int s1) {
  if constexpr (has_bus_rows(FORM)) {
    constexpr int NJ = he_nj(FORM), NH = he_nh(FORM), NG = he_npg(FORM);
    constexpr int ND = (NJ + NH + NG + 1) / 2;
    constexpr bool PB = (FORM == PBRAPV || FORM == PBRARV);
    const WTile t = P.wtiles[tile];
    const int lane = threadIdx.x & 31;
    constexpr int KIND = MODE & 3;
    constexpr bool FULL = KIND == WT_FULL;
    const bool wg = KIND != WT_ANY || (A.what & EV_G) != 0;
    const bool wj = KIND == WT_FULL || KIND == WT_GJ || (KIND == WT_ANY && (A.what & EV_J) != 0);
    const bool wh = KIND == WT_FULL || (KIND == WT_ANY && (A.what & EV_H) != 0);
    const bool has_status = (MODE & WT_ST) != 0 || (KIND == WT_ANY && A.status != nullptr);
    const bool need_g = wg || A.scal;
    const bool facts = (P.flags & FLAG_FACTS) != 0;
    double* const spJ = region;
    double* const spH = spJ + P.wt_jspace;
    double* const spG = spH + P.wt_hspace;
    uint64_t* const multi = (uint64_t*)(spG + P.wt_gspace);     // multi-slot descriptors (ps_plan.h: wt_multi)
    const int oj0 = P.wt_ownj_ptr[t.u0], noj = P.wt_ownj_ptr[t.u1] - oj0;
    const int oh0 = P.wt_ownh_ptr[t.u0], noh = P.wt_ownh_ptr[t.u1] - oh0;
    const int og0 = P.wt_owng_ptr[t.u0], nog = P.wt_owng_ptr[t.u1] - og0;
    uint16_t* const ownj = (uint16_t*)(multi + t.mn);
    uint16_t* const ownh = ownj + noj;
    uint16_t* const owng = ownh + noh;
    // ---- scenario-independent set-up
    for (int q = lane; q < P.wt_hspace; q += 32) spH[q] = 0.0;    // slots whose contributions are all structural zeros stay 0
    for (int q = lane; q < noj; q += 32) ownj[q] = P.wt_ownj[oj0 + q];
    for (int q = lane; q < noh; q += 32) ownh[q] = P.wt_ownh[oh0 + q];
    for (int q = lane; q < nog; q += 32) owng[q] = P.wt_owng[og0 + q];
    // lane = half-edge
    const bool hact = lane < t.nhe;
    int ks = 0, hb = t.u0, oth = 0;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    uint32_t d[ND];
#pragma unroll
    for (int i = 0; i < ND; i++) d[i] = 0;
    if (hact) {
      const int q = t.he0 + lane;
      ks = P.he_ks[q]; hb = P.he_bus[q];
      if constexpr (is_PN(FORM)) {
        oth = P.he_other[q];
        c0 = P.he_c[q]; c1 = P.he_c[(long long)P.nhe + q]; c2 = P.he_c[2LL * P.nhe + q]; c3 = P.he_c[3LL * P.nhe + q];
      }
      const uint16_t* dt = P.wt_dst + t.dst0 + lane;
#pragma unroll
      for (int i = 0; i < ND; i++) {
        const unsigned lo = dt[(2 * i) * 32], hi = (2 * i + 1 < NJ + NH + NG) ? dt[(2 * i + 1) * 32] : 0u;
        d[i] = lo | (hi << 16);
      }
    }
    const int k = ks >> 1, side = ks & 1;
    const int xo0 = is_PN(FORM) ? P.o_a + oth : P.o_f[0 + 2 * side] + k;
    const int xo1 = is_PN(FORM) ? P.o_b + oth : P.o_f[1 + 2 * side] + k;
    const int xa = P.o_a + hb, xb = P.o_b + hb;
    // lane = bus
    const int nbt = t.u1 - t.u0, u = t.u0 + lane;
    const bool bact = lane < nbt;
    int g0 = 0, ngen = 0, h0l = 0, deg = 0, f0 = 0, nsh = 0, oj = 0, oh = 0, og = 0;
    double Gl = 0.0, Bl = 0.0;
    if (bact) {
      g0 = P.gen_ptr[u]; ngen = P.gen_ptr[u + 1] - g0;
      h0l = P.he_ptr[u] - t.he0; deg = P.he_ptr[u + 1] - P.he_ptr[u];
      if (facts) { f0 = P.sh_ptr[u]; nsh = P.sh_ptr[u + 1] - f0; }
      oj = P.wt_ownj_ptr[u] - oj0; oh = P.wt_ownh_ptr[u] - oh0; og = P.wt_owng_ptr[u] - og0;
      Gl = P.Gl[u]; Bl = P.Bl[u];
    }
    const int own_src = deg > 0 ? h0l : lane;                   // lane that holds this bus's own voltage / multipliers
    // row pointers of the current scenario (warp-uniform); the images share the 16-byte phase of their destination
    const double* __restrict__ xrow = A.x + (long long)s0 * A.ldx;
    const double* __restrict__ murow = wh ? A.mu + (long long)s0 * A.ldmu : nullptr;
    const double* __restrict__ pdrow = A.Pd ? A.Pd + (long long)s0 * A.ldl : P.Pd0;
    const double* __restrict__ qdrow = A.Qd ? A.Qd + (long long)s0 * A.ldl : P.Qd0;
    const long long ldload = A.Pd ? A.ldl : 0, ldloadq = A.Qd ? A.ldl : 0;
    const uint8_t* __restrict__ strow = has_status ? A.status + (long long)s0 * A.ldst : nullptr;
    double* jrow = wj ? A.J + (long long)s0 * A.ldJ + t.j0 : nullptr;
    double* hrow = wh ? A.H + (long long)s0 * A.ldH + t.h0 : nullptr;
    double* grow = wg ? A.g + (long long)s0 * A.ldg + 2 * t.u0 : nullptr;
    double* const imgJ = spJ + ((reinterpret_cast<unsigned long long>(jrow) >> 3) & 1);
    double* const imgH = spH + ((reinterpret_cast<unsigned long long>(hrow) >> 3) & 1);
    double* const imgG = spG;
    // image and destination share their 16-byte phase in every scenario iff the leading dimension is even
    const WtBlock blkJ = wt_block(imgJ, t.jn, (A.ldJ & 1) == 0), blkH = wt_block(imgH, t.hn, (A.ldH & 1) == 0);
    for (int q = lane; q < t.mn; q += 32) multi[q] = wt_predecode(P.wt_multi[t.m0 + q], imgJ, imgH, imgG, wj, wh, need_g);
    // the generator entries of J are the constant 1 (:177-180): written once, the image persists across scenarios
    if (bact && wj) for (int q = 0; q < 2 * ngen; q++) imgJ[P.wt_ownj[oj0 + oj + q]] = 1.0;
    struct Raw { double v0, v1, va0, vb0, muP, muQ, st, Pd, Qd; };
    auto load = [&](const double* __restrict__ x, const double* __restrict__ mu, const double* __restrict__ pd,
                    const double* __restrict__ qd, const uint8_t* __restrict__ stp) {
      Raw r;
      r.v0 = r.v1 = r.va0 = r.vb0 = r.muP = r.muQ = r.Pd = r.Qd = 0.0; r.st = 1.0;
      if (hact) {
        r.v0 = x[xo0]; r.v1 = x[xo1]; r.va0 = x[xa]; r.vb0 = x[xb];
        if (wh) { r.muP = mu[2 * hb]; r.muQ = mu[2 * hb + 1]; }
        if (has_status) r.st = (double)stp[k];
      }
      if (bact && need_g) { r.Pd = pd[u]; r.Qd = qd[u]; }
      return r;
    };
    Raw cur = load(xrow, murow, pdrow, qdrow, strow);
    __syncwarp();
    for (int s = s0; s < s1; s++) {
      Raw nxt = cur;
      if (s + 1 < s1)                                            // one scenario ahead
        nxt = load(xrow + A.ldx, wh ? murow + A.ldmu : nullptr, pdrow + ldload, qdrow + ldloadq, has_status ? strow + A.ldst : nullptr);
      const double* __restrict__ x = xrow;
      if (s > s0 && lane == 0) bulk_wait_read<0>();             // the copy engine has read the previous images
      __syncwarp();
      // ---- 1. half-edge contributions
      if (hact) {
        HeIn in;
        in.side = side; in.st = cur.st; in.v0 = cur.v0; in.v1 = cur.v1; in.va0 = cur.va0; in.vb0 = cur.vb0;
        in.muP = cur.muP; in.muQ = cur.muQ; in.c0 = c0; in.c1 = c1; in.c2 = c2; in.c3 = c3;
        auto gput = [&](int kk, double v) {
          imgG[(d[(NJ + NH + kk) >> 1] >> (((NJ + NH + kk) & 1) * 16)) & 0xffffu] = piece_neg(FORM, kk) ? -v : v;
        };
        if constexpr (KIND == WT_GJ) {
          RegW sj{imgJ, d, 0, true};
          NullW sh;
          he_math<FORM>(in, sj, sh, gput, need_g);
        } else {
          RegW sj{imgJ, d, 0, wj}, sh{imgH, d, NJ, wh};
          he_math<FORM>(in, sj, sh, gput, need_g);
        }
        if constexpr (PB) {
          if (need_g) { gput(0, cur.st * cur.v0); gput(1, cur.st * cur.v1); }
        }
      }
      // ---- 2. the bus's own terms (own voltage / multipliers come from the bus's first half-edge lane)
      OwnIn o;
      o.va0 = __shfl_sync(0xffffffffu, cur.va0, own_src); o.vb0 = __shfl_sync(0xffffffffu, cur.vb0, own_src);
      o.muP = __shfl_sync(0xffffffffu, cur.muP, own_src); o.muQ = __shfl_sync(0xffffffffu, cur.muQ, own_src);
      o.W = 0.0; o.Gl = Gl; o.Bl = Bl; o.ngen = 0; o.nsh = nsh;  // ngen = 0: the generator J entries are already in the image
      if (bact) {
        if (deg == 0) {                                         // isolated bus: nobody loaded its inputs
          o.va0 = x[P.o_a + u]; o.vb0 = x[P.o_b + u];
          o.muP = wh ? murow[2 * u] : 0.0; o.muQ = wh ? murow[2 * u + 1] : 0.0;
        }
        if constexpr (FORM == CBRAWV) o.W = x[P.o_Wd + u];
        auto bssv = [&](int q) { return x[P.o_bss + P.sh_idx[f0 + q]]; };
        if (wj || wh) {
          ListW lj{imgJ, ownj + oj + 2 * ngen, wj}, lh{imgH, ownh + oh, wh};
          own_math<FORM>(o, bssv, lj, lj, lh);
        }
        if (need_g) {                                           // residual terms: :177-180, :256-279, :291-292
          ListW lg{imgG, owng + og, true};
          for (int q = 0; q < ngen; q++) {
            const int gi = P.gen_idx[g0 + q];
            lg.put(x[P.o_pg + gi]); lg.put(x[P.o_qg + gi]);
          }
          double m2;                                            // |V|^2 of the shunt terms
          if constexpr (is_polar(FORM)) m2 = o.vb0 * o.vb0;
          else if constexpr (FORM == PBRARV || FORM == CBRARV || FORM == PNRARV) m2 = o.vb0 * o.vb0 + o.va0 * o.va0;
          else m2 = o.W;
          lg.put(-(Gl * m2)); lg.put(Bl * m2);
          for (int q = 0; q < nsh; q++) lg.put(bssv(q) * m2);
          lg.put(-cur.Pd); lg.put(-cur.Qd);
        }
      }
      __syncwarp();
      // ---- 3. slots with several contributions
      reduce_multi_warp(multi, t.mn, lane);
      // ---- 4. the images go out
      fence_async_smem();
      __syncwarp();
      if (wj) wt_store_block(jrow, imgJ, blkJ, lane);
      if (wh) wt_store_block(hrow, imgH, blkH, lane);
      if (lane == 0) bulk_commit();
      if (need_g) {
        double vmax = 0.0;
        int nviol = 0;
        for (int e = lane; e < 2 * nbt; e += 32) {
          const double v = imgG[e];
          if (wg) grow[e] = v;
          vmax = fmax(vmax, fabs(v));
          nviol += fabs(v) > A.viol_tol ? 1 : 0;
        }
        reduce_scalars_redux(A.scal ? A.scal + (long long)s * NSCAL : nullptr, vmax, nviol);
      }
      cur = nxt;
      xrow += A.ldx; pdrow += ldload; qdrow += ldloadq;
      if (wh) { murow += A.ldmu; hrow += A.ldH; }
      if (wj) jrow += A.ldJ;
      if (wg) grow += A.ldg;
      if (has_status) strow += A.ldst;
    }
    if (lane == 0) bulk_wait_read<0>();                         // shared memory must outlive the copies that read it
  }
}

#ifndef PS_WT_MINB
#define PS_WT_MINB 16               // 16 warps / SM: 128 registers per thread
#endif
__host__ __device__ inline int wt_region_doubles(const DevPlan& P) {   // per-warp shared memory, doubles (even)
  return P.wt_jspace + P.wt_hspace + P.wt_gspace + ((P.wt_lists + 7) / 8) * 2;
}
// one kernel per mode: each gets its own register allocation and code (the four bodies in ONE kernel cost the headline
// g + J + H instantiation 8 %: gpurun_out/r2c13_ab.log)
template <int FORM, int MODE>
__global__ void __launch_bounds__(32 * WPC, PS_WT_MINB) bus_warp_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int tile = blockIdx.x * WPC + (threadIdx.x >> 5);
  if (tile >= P.n_wtiles) return;
  const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
  double* r = smem + (threadIdx.x >> 5) * wt_region_doubles(P);
  bus_warp_body<FORM, MODE>(P, A, r, tile, s0, s1);
}
// which instantiation serves a call
__host__ inline int wt_mode(const EvalArgs& A) {
  const int w = A.what & (EV_G | EV_J | EV_H);
  const int st = A.status != nullptr ? WT_ST : 0;
  if (A.force_any) return WT_ANY;
  if (w == (EV_G | EV_J | EV_H)) return WT_FULL | st;
  if (w == (EV_G | EV_J)) return WT_GJ | st;
  return WT_ANY;
}

// ---- residuals only, thread per bus (any formulation with nodal balances) -------------------------------------------
// eval_constraint alone -- every trial point of the SLP line search -- writes 16 bytes per bus.  The warp-tile kernel's
// residual-only instantiation still runs its phases (half-edge lanes, bus lanes, ordered slot sums, copy-out) once per
// scenario: 245 SM cycles per 32 half-edges for one sincos and a dozen additions each.  Here a thread owns a bus and walks its
// generator / half-edge / shunt lists like pb_bus_g_kernel does for the PB* models, forming each half-edge's pieces with the
// same he_math and folding them with the same residual_math as the CTA-tile kernel (the reference's order and
// association): no shared memory, no phases, any degree (hub buses included).  Bitwise equal to the other bus kernels.
template <int FORM>
__global__ void __launch_bounds__(PS_TPB) bus_g_list_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  if constexpr (has_bus_rows(FORM)) {
    constexpr bool PB = (FORM == PBRAPV || FORM == PBRARV);
    const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
    const int i = blockIdx.x * PS_TPB + threadIdx.x;
    const bool act = i < P.nb;
    const bool facts = (P.flags & FLAG_FACTS) != 0;
    int g0 = 0, h0 = 0, f0 = 0;
    OwnIn o{};
    int deg = 0;
    if (act) {
      g0 = P.gen_ptr[i]; o.ngen = P.gen_ptr[i + 1] - g0;
      h0 = P.he_ptr[i]; deg = P.he_ptr[i + 1] - h0;
      if (facts) { f0 = P.sh_ptr[i]; o.nsh = P.sh_ptr[i + 1] - f0; }
      o.Gl = P.Gl[i]; o.Bl = P.Bl[i];
    }
    const bool wg = (A.what & EV_G) != 0;
    for (int s = s0; s < s1; s++) {
      double vmax = 0.0;
      int nviol = 0;
      if (act) {
        const double* __restrict__ x = A.x + (long long)s * A.ldx;
        const uint8_t* __restrict__ stp = A.status ? A.status + (long long)s * A.ldst : nullptr;
        const double* pd = A.Pd ? A.Pd + (long long)s * A.ldl : P.Pd0;
        const double* qd = A.Qd ? A.Qd + (long long)s * A.ldl : P.Qd0;
        o.va0 = x[P.o_a + i]; o.vb0 = x[P.o_b + i];
        if constexpr (FORM == CBRAWV) o.W = x[P.o_Wd + i];
        // the pieces of one half-edge, formed when the fold first asks for them
        double pc[6];
        int cq = -1;
        auto pieces = [&](int q) {
          const int e = h0 + q, ks = P.he_ks[e], k = ks >> 1;
          HeIn in;
          in.side = ks & 1;
          in.st = stp ? (double)stp[k] : 1.0;
          if constexpr (is_PN(FORM)) {
            const int oth = P.he_other[e];
            in.v0 = x[P.o_a + oth]; in.v1 = x[P.o_b + oth];
            in.c0 = P.he_c[e]; in.c1 = P.he_c[(long long)P.nhe + e]; in.c2 = P.he_c[2LL * P.nhe + e]; in.c3 = P.he_c[3LL * P.nhe + e];
          } else {
            in.v0 = x[P.o_f[0 + 2 * in.side] + k]; in.v1 = x[P.o_f[1 + 2 * in.side] + k];
            in.c0 = in.c1 = in.c2 = in.c3 = 0.0;
          }
          in.va0 = o.va0; in.vb0 = o.vb0; in.muP = in.muQ = 0.0;
          NullW sj, sh;
          he_math<FORM>(in, sj, sh, [&](int kk, double v) { pc[kk] = v; }, true);
          if constexpr (PB) { pc[0] = in.st; pc[1] = in.v0; pc[2] = in.v1; }
        };
        double gP, gQ;
        residual_math<FORM>(
            o, deg, pd[i], qd[i],
            [&](int q, double& pg, double& qg) { const int g = P.gen_idx[g0 + q]; pg = x[P.o_pg + g]; qg = x[P.o_qg + g]; },
            [&](int q, double& st, double& fp, double& fq) { if (q != cq) { cq = q; pieces(q); } st = pc[0]; fp = pc[1]; fq = pc[2]; },
            [&](int q, int k) { if (q != cq) { cq = q; pieces(q); } return pc[k]; },
            [&](int q) { return x[P.o_bss + P.sh_idx[f0 + q]]; }, gP, gQ);
        if (wg) {
          double* gd = A.g + (long long)s * A.ldg + 2 * i;
          if ((reinterpret_cast<unsigned long long>(gd) & 15) == 0) *reinterpret_cast<double2*>(gd) = make_double2(gP, gQ);
          else { gd[0] = gP; gd[1] = gQ; }
        }
        vmax = fmax(fabs(gP), fabs(gQ));
        nviol = (fabs(gP) > A.viol_tol ? 1 : 0) + (fabs(gQ) > A.viol_tol ? 1 : 0);
      }
      reduce_scalars(A.scal ? A.scal + (long long)s * NSCAL : nullptr, vmax, nviol);
    }
  }
}

// ---- PB* balances, fast path (PBRAPV / PBRARV without switched shunts) ------------------------------------------
// The nodal balance of a power-branch-flow model is linear in the generator and flow variables
// (formulations.jl:188-196): its Jacobian entries are the constant 1 (times the branch status) except the 1-2
// own-voltage slots, and its Hessian block is (-2 Gl mu_P | 2 Bl mu_Q).  So the bus block of J and H is produced
// slot-parallel from per-slot codes -- one 32-byte sector per thread, no staging, no barrier -- and the two
// residuals of a bus by one thread that walks the bus's generator / Sij / Sji lists in the reference's order.
constexpr int PBT = TPB;

// dst[0..n) = coef[e] * m(idx[e]): the few elements before the first 32-byte boundary and after the last one with
// scalar stores, everything else as whole sectors whose four table entries come from one vector load each
template <class F>
__device__ __forceinline__ void emit_sectors(double* __restrict__ dst, int n, int t, int nt, const double* __restrict__ coefs,
                                             const int32_t* __restrict__ idxs, int stride, F m) {
  const int a = min(n, (int)((4 - ((reinterpret_cast<unsigned long long>(dst) >> 3) & 3)) & 3));
  const long long off = (long long)(a & 3) * stride + ((4 - a) & 3);
  const double* __restrict__ coef = coefs + off;
  const int32_t* __restrict__ idx = idxs + off;
  const int nfull = (n - a) >> 2;
  for (int q = t; q < nfull; q += nt) {
    const int e = a + 4 * q;
    const int4 c = *reinterpret_cast<const int4*>(idx + e);
    const double4 k = *reinterpret_cast<const double4*>(coef + e);
    const double m0 = m(c.x), m1 = m(c.y), m2 = m(c.z), m3 = m(c.w);
    st_v4(dst + e, k.x * m0, k.y * m1, k.z * m2, k.w * m3);
  }
  if (t < a) dst[t] = coef[t] * m(idx[t]);
  const int tail = a + 4 * nfull;
  if (t < n - tail) dst[tail + t] = coef[tail + t] * m(idx[tail + t]);
}

template <int FORM>
__device__ __forceinline__ void pb_bus_jh_body(const DevPlan& P, const EvalArgs& A, int bx, int nbx) {
  const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
  const int t = bx * PBT + threadIdx.x, nt = nbx * PBT;
  for (int s = s0; s < s1; s++) {
    const double* __restrict__ x = A.x + (long long)s * A.ldx;
    if (A.what & EV_J) {
      double* Jd = A.J + (long long)s * A.ldJ;
      if (A.status) {
        const uint8_t* __restrict__ st = A.status + (long long)s * A.ldst;
        emit_sectors(Jd, P.j_br0, t, nt, P.pb_jt, P.pb_ji, P.pb_jstride,
                     [&](int c) { return c >= 0 ? x[c] : (c == -1 ? 1.0 : (double)st[-2 - c]); });
      } else {
        emit_sectors(Jd, P.j_br0, t, nt, P.pb_jt, P.pb_ji, P.pb_jstride, [&](int c) { return c >= 0 ? x[c] : 1.0; });
      }
    }
    if (A.what & EV_H) {
      const double* __restrict__ mu = A.mu + (long long)s * A.ldmu;
      emit_sectors(A.H + (long long)s * A.ldH, P.h_br0, t, nt, P.pb_ht, P.pb_hi, P.pb_hstride, [&](int c) { return mu[c]; });
    }
  }
}

template <int FORM>
__global__ void __launch_bounds__(PBT) pb_bus_jh_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  pb_bus_jh_body<FORM>(P, A, blockIdx.x, gridDim.x);
}

template <int FORM>
__device__ __forceinline__ void pb_bus_g_body(const DevPlan& P, const EvalArgs& A, int bx) {
  const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
  const int i = bx * PBT + threadIdx.x;
  const bool act = i < P.nb;
  int g0 = 0, g1 = 0, h0 = 0, h1 = 0;
  double Gl = 0.0, Bl = 0.0;
  if (act) { g0 = P.gen_ptr[i]; g1 = P.gen_ptr[i + 1]; h0 = P.he_ptr[i]; h1 = P.he_ptr[i + 1]; Gl = P.Gl[i]; Bl = P.Bl[i]; }
  const bool wg = (A.what & EV_G) != 0;
  for (int s = s0; s < s1; s++) {
    double vmax = 0.0;
    int nviol = 0;
    if (act) {
      const double* __restrict__ x = A.x + (long long)s * A.ldx;
      const uint8_t* __restrict__ stp = A.status ? A.status + (long long)s * A.ldst : nullptr;
      double accP = 0.0, accQ = 0.0;                           // formulations.jl:171-172
      for (int q = g0; q < g1; q++) {                          // :177-180
        const int g = P.gen_idx[q];
        accP = accP + x[P.o_pg + g];
        accQ = accQ + x[P.o_qg + g];
      }
      for (int q = h0; q < h1; q++) {                          // :188-196, Sij ascending then Sji ascending
        const int ks = P.he_ks[q], k = ks >> 1, side = ks & 1;
        const double stv = stp ? (double)stp[k] : 1.0;
        accP = accP + stv * x[P.o_f[0 + 2 * side] + k];
        accQ = accQ + stv * x[P.o_f[1 + 2 * side] + k];
      }
      if constexpr (FORM == PBRARV) {                          // :264-266
        const double im = x[P.o_a + i], r = x[P.o_b + i], A2 = r * r + im * im;
        accP = accP - Gl * A2;
        accQ = accQ + Bl * A2;
      } else {                                                 // :256-258
        const double V = x[P.o_b + i], V2 = V * V;
        accP = accP - Gl * V2;
        accQ = accQ + Bl * V2;
      }
      const double* pd = A.Pd ? A.Pd + (long long)s * A.ldl : P.Pd0;
      const double* qd = A.Qd ? A.Qd + (long long)s * A.ldl : P.Qd0;
      const double gP = accP - pd[i], gQ = accQ - qd[i];       // :291-292
      if (wg) {
        double* gd = A.g + (long long)s * A.ldg + 2 * i;
        if ((reinterpret_cast<unsigned long long>(gd) & 15) == 0) *reinterpret_cast<double2*>(gd) = make_double2(gP, gQ);
        else { gd[0] = gP; gd[1] = gQ; }
      }
      vmax = fmax(fabs(gP), fabs(gQ));
      nviol = (fabs(gP) > A.viol_tol ? 1 : 0) + (fabs(gQ) > A.viol_tol ? 1 : 0);
    }
    reduce_scalars(A.scal ? A.scal + (long long)s * NSCAL : nullptr, vmax, nviol);
  }
}

template <int FORM>
__global__ void __launch_bounds__(PBT) pb_bus_g_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  pb_bus_g_body<FORM>(P, A, blockIdx.x);
}

// ---- PB* residuals, scenario-resident -----------------------------------------------------------------------------
// pb_bus_g_kernel walks a bus's lists and gathers one 8-byte flow variable per 32-byte sector on a randomly numbered
// network: it is bound by the L1 -> L2 request rate at 4.6x the useful sectors (profiles/r02_ncu_pb_bus_g_summary.txt).
// Here the roles are swapped.  A CTA owns a whole scenario at a time and keeps the (P, Q) running sums of EVERY bus in
// shared memory (16 bytes per bus: 218 KB for case13659pegase); the terms are visited in element order -- generators,
// then the Sij side of every branch, then the Sji side -- PBS_WAVE consecutive elements per step, so pg / qg and the
// four flow arrays are read with perfectly coalesced loads, each sector once.  A bus still receives its terms in the
// reference's order (formulations.jl:171-196: generators, Sij ascending, Sji ascending): the adjacency lists are
// ascending (checked at plan time), passes and steps are separated by CTA barriers, and the elements of one step that
// share a bus are served in rank order (plan table pbs_code, one sub-round per rank; a random numbering needs one or two).
// The shunt term and the load are subtracted by the thread that owns the bus in those passes, which then writes the row.
// Bitwise identical to pb_bus_g_kernel (x - y == x + (-y); status 1.0 * v == v).
#ifndef PS_PBS_R
#define PS_PBS_R 4
#endif
constexpr int PBS_T = PBS_THREADS;
constexpr int PBS_E = PS_PBS_E;    // elements per thread and step (1: more per step -- fewer barriers, independent read-modify-writes --
                                   // measured slower at every depth that fits 64 registers: gpurun_out/r2c16_ab.log)
constexpr int PBS_MAX_STEPS = 1024;
constexpr int PBS_RB = 4;           // ring depth of the two bus-ordered passes (one element per thread and step)
constexpr int PBS_R = PS_PBS_R;    // steps whose inputs are in flight (a register ring: the CTA is alone on its SM -- the
                                   // accumulators fill its shared memory -- so memory latency is hidden by depth, not by occupancy)

// One pass over a term table: step j serves elements [j * PBS_WAVE, (j + 1) * PBS_WAVE) -- thread tid the PBS_E elements
// j * PBS_WAVE + e * PBS_T + tid -- with the inputs of the next PBS_R steps already in flight.
template <bool HAS_ST>
__device__ __forceinline__ void pbs_term_pass(double2* acc, const uint8_t* wmax, const int32_t* __restrict__ code, int n,
                                              const double* __restrict__ v0p, const double* __restrict__ v1p,
                                              const uint8_t* __restrict__ stp) {
  const int tid = threadIdx.x;
  // whole rings: the steps past the end have no terms (the loads are guarded) and cost one barrier each; unconditional
  // ring bodies keep every slot in a fixed register (a guarded body makes the compiler rotate the ring through moves,
  // which waits for a load one step after it was issued)
  const int nsteps = ((n + PBS_WAVE - 1) / PBS_WAVE + PBS_R - 1) / PBS_R * PBS_R;
  int c[PBS_R][PBS_E];
  double a0[PBS_R][PBS_E], a1[PBS_R][PBS_E];
  auto load = [&](int u, int j) {
#pragma unroll
    for (int q = 0; q < PBS_E; q++) {
      const int e = j * PBS_WAVE + q * PBS_T + tid;
      c[u][q] = -1; a0[u][q] = 0.0; a1[u][q] = 0.0;
      if (e < n) {
        c[u][q] = code[e]; a0[u][q] = v0p[e]; a1[u][q] = v1p[e];
        if constexpr (HAS_ST) { const double st = (double)stp[e]; a0[u][q] = st * a0[u][q]; a1[u][q] = st * a1[u][q]; }
      }
    }
  };
#pragma unroll
  for (int u = 0; u < PBS_R; u++) load(u, u);
  for (int j0 = 0; j0 < nsteps; j0 += PBS_R) {
#pragma unroll
    for (int u = 0; u < PBS_R; u++) {
      const int j = j0 + u;
      int cc[PBS_E];
      double b0[PBS_E], b1[PBS_E];
#pragma unroll
      for (int q = 0; q < PBS_E; q++) { cc[q] = c[u][q]; b0[q] = a0[u][q]; b1[q] = a1[u][q]; }
      load(u, j + PBS_R);                                      // refill: in flight while the next PBS_R - 1 steps are served
      for (int r = 0, mr = wmax[j]; r <= mr; r++) {            // elements of the step that share a bus: rank order
        // one rank: every read-modify-write of the sub-round goes to a different bus (ranks are per bus and step), so the
        // PBS_E of a thread are independent and overlap
#pragma unroll
        for (int q = 0; q < PBS_E; q++) {
          if ((cc[q] >> 24) == r) {                            // no term: code -1, rank -1
            double2* const p = acc + (cc[q] & 0xffffff);
            double2 v = *p;
            v.x = v.x + b0[q];
            v.y = v.y + b1[q];
            *p = v;
          }
        }
        __syncthreads();
      }
    }
  }
}

template <int FORM>
__global__ void __launch_bounds__(PBS_T, 1) pb_bus_g_scn_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  extern __shared__ __align__(16) double2 pbs_acc[];           // [nb] running (P, Q) sums of the scenario
  __shared__ double redv[PBS_T / 32], redo[PBS_T / 32];
  __shared__ int redn[PBS_T / 32];
  __shared__ uint8_t wmax[PBS_MAX_STEPS];                      // highest rank per term-table step
  const int tid = threadIdx.x;
  const bool wg = (A.what & EV_G) != 0;
  const int s_gen = (P.ng + PBS_WAVE - 1) / PBS_WAVE, s_br = (P.nr + PBS_WAVE - 1) / PBS_WAVE, s_bus = (P.nb + PBS_T - 1) / PBS_T;
  for (int j = tid; j < PBS_MAX_STEPS; j += PBS_T) wmax[j] = j < s_gen + 2 * s_br ? P.pbs_wmax[j] : 0;
  for (int s = blockIdx.x; s < A.S; s += gridDim.x) {
    const double* __restrict__ x = A.x + (long long)s * A.ldx;
    const uint8_t* __restrict__ stp = A.status ? A.status + (long long)s * A.ldst : nullptr;
    const double* __restrict__ pd = A.Pd ? A.Pd + (long long)s * A.ldl : P.Pd0;
    const double* __restrict__ qd = A.Qd ? A.Qd + (long long)s * A.ldl : P.Qd0;
    double* __restrict__ gd = wg ? A.g + (long long)s * A.ldg : nullptr;
    const bool al = (reinterpret_cast<unsigned long long>(gd) & 15) == 0;
    for (int i = tid; i < P.nb; i += PBS_T) pbs_acc[i] = make_double2(0.0, 0.0);     // formulations.jl:171-172
    __syncthreads();
    pbs_term_pass<false>(pbs_acc, wmax, P.pbs_code, P.ng, x + P.o_pg, x + P.o_qg, nullptr);                        // :177-180
    if (stp) {                                                                                                       // :188-196
      pbs_term_pass<true>(pbs_acc, wmax + s_gen, P.pbs_code + P.ng, P.nr, x + P.o_f[0], x + P.o_f[1], stp);        // Sij side
      pbs_term_pass<true>(pbs_acc, wmax + s_gen + s_br, P.pbs_code + P.ng + P.nr, P.nr, x + P.o_f[2], x + P.o_f[3], stp);   // Sji side
    } else {
      pbs_term_pass<false>(pbs_acc, wmax + s_gen, P.pbs_code + P.ng, P.nr, x + P.o_f[0], x + P.o_f[1], nullptr);
      pbs_term_pass<false>(pbs_acc, wmax + s_gen + s_br, P.pbs_code + P.ng + P.nr, P.nr, x + P.o_f[2], x + P.o_f[3], nullptr);
    }
    // the bus's own shunt term (:256-258 | :264-266), then minus the load (:291-292) and out; thread tid owns bus
    // j * PBS_T + tid in both passes: no barrier between them
    {
      double va[PBS_RB], vb[PBS_RB], gl[PBS_RB], bl[PBS_RB];
      auto load = [&](int u, int j) {
        const int e = j * PBS_T + tid;
        va[u] = vb[u] = gl[u] = bl[u] = 0.0;
        if (e < P.nb) {
          if constexpr (FORM == PBRARV) va[u] = x[P.o_a + e];
          vb[u] = x[P.o_b + e]; gl[u] = P.Gl[e]; bl[u] = P.Bl[e];
        }
      };
      constexpr int R2 = PBS_RB > 3 ? 3 : PBS_RB;               // four values per bus: a shallower ring
#pragma unroll
      for (int u = 0; u < R2; u++) load(u, u);
      for (int j0 = 0; j0 < s_bus; j0 += R2) {
#pragma unroll
        for (int u = 0; u < R2; u++) {
          const int j = j0 + u, e = j * PBS_T + tid;
          const double im = va[u], r = vb[u], Gl = gl[u], Bl = bl[u];
          load(u, j + R2);
          if (e < P.nb) {
            double m2;
            if constexpr (FORM == PBRARV) m2 = r * r + im * im;
            else m2 = r * r;
            double2 a = pbs_acc[e];
            a.x = a.x - Gl * m2;
            a.y = a.y + Bl * m2;
            pbs_acc[e] = a;
          }
        }
      }
    }
    double vmax = 0.0;
    int nviol = 0;
    {
      double lp[PBS_RB], lq[PBS_RB];
      auto load = [&](int u, int j) {
        const int e = j * PBS_T + tid;
        lp[u] = lq[u] = 0.0;
        if (e < P.nb) { lp[u] = pd[e]; lq[u] = qd[e]; }
      };
#pragma unroll
      for (int u = 0; u < PBS_RB; u++) load(u, u);
      for (int j0 = 0; j0 < s_bus; j0 += PBS_RB) {
#pragma unroll
        for (int u = 0; u < PBS_RB; u++) {
          const int j = j0 + u, e = j * PBS_T + tid;
          const double Pd = lp[u], Qd = lq[u];
          load(u, j + PBS_RB);
          if (e < P.nb) {
            const double2 a = pbs_acc[e];
            const double gP = a.x - Pd, gQ = a.y - Qd;
            if (wg) {
              if (al) *reinterpret_cast<double2*>(gd + 2 * e) = make_double2(gP, gQ);
              else { gd[2 * e] = gP; gd[2 * e + 1] = gQ; }
            }
            vmax = fmax(vmax, fmax(fabs(gP), fabs(gQ)));
            nviol += (fabs(gP) > A.viol_tol ? 1 : 0) + (fabs(gQ) > A.viol_tol ? 1 : 0);
          }
        }
      }
    }
    if (A.scal) {                                              // CTA-uniform
      static_assert(PBS_T == OBJ_LANES, "the CTA's threads are the objective's virtual lanes");
      const double gsum = objective_lane(P, x, tid);           // the objective scalar, in objective_tile's order
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
        nviol += __shfl_xor_sync(0xffffffffu, nviol, d);
      }
      if ((tid & 31) == 0) { redv[tid >> 5] = vmax; redn[tid >> 5] = nviol; redo[tid >> 5] = gsum; }
      __syncthreads();
      if (tid == 32) {
        double tot = 0.0;
        for (int w = 0; w < PBS_T / 32; w++) tot += redo[w];
        A.scal[(long long)s * NSCAL + 0] = tot;
      }
      if (tid < 32) {
        vmax = tid < PBS_T / 32 ? redv[tid] : 0.0; nviol = tid < PBS_T / 32 ? redn[tid] : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
          nviol += __shfl_xor_sync(0xffffffffu, nviol, d);
        }
        if (tid == 0) {
          double* scal = A.scal + (long long)s * NSCAL;
          if (vmax > 0.0) atomicMax((unsigned long long*)(scal + 1), (unsigned long long)__double_as_longlong(vmax));
          if (nviol) atomicAdd(scal + 2, (double)nviol);
        }
      }
    }
    __syncthreads();                                           // the accumulators are reused by the next scenario
  }
}

// objective scalar only (used with the PB* fast path, where no bus_kernel runs)
// ---- CB* balances, fast path (CBRARV / CBRAWV without switched shunts and without a status mask) --------------------
// The current-branch-flow balances are bilinear in (own voltage, branch current) (formulations.jl:206-214): every
// Jacobian entry of a current variable is +-1 * an own-voltage entry of x, every Hessian entry is +-1 (or the shunt's
// -2Gl / 2Bl) * the row multiplier or a structural 0, and the only sums are the four own-voltage Jacobian entries of a
// bus.  So: cb_bus_g_kernel, thread per bus, walks the bus's generator / Sij / Sji lists in the reference's order and
// leaves the two residuals and those four sums (scratch [S][4 nb]); cb_bus_jh_kernel then produces the whole bus block
// of J and H slot-parallel as coefficient[e] * m(index[e]) -- one 32-byte sector per thread, tables read with vector
// loads -- exactly like the PB* fast path.
template <int FORM>
__global__ void __launch_bounds__(PBT) cb_bus_g_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
  const int i = blockIdx.x * PBT + threadIdx.x;
  const bool act = i < P.nb;
  int g0 = 0, g1 = 0, h0 = 0, h1 = 0;
  double Gl = 0.0, Bl = 0.0;
  if (act) { g0 = P.gen_ptr[i]; g1 = P.gen_ptr[i + 1]; h0 = P.he_ptr[i]; h1 = P.he_ptr[i + 1]; Gl = P.Gl[i]; Bl = P.Bl[i]; }
  const bool wg = (A.what & EV_G) != 0, wd = (A.what & EV_J) != 0 && A.diag != nullptr;
  for (int s = s0; s < s1; s++) {
    double vmax = 0.0;
    int nviol = 0;
    if (act) {
      const double* __restrict__ x = A.x + (long long)s * A.ldx;
      const double ia = x[P.o_a + i], ra = x[P.o_b + i];
      double accP = 0.0, accQ = 0.0;                           // formulations.jl:171-172
      for (int q = g0; q < g1; q++) {                          // :177-180
        const int g = P.gen_idx[q];
        accP = accP + x[P.o_pg + g];
        accQ = accQ + x[P.o_qg + g];
      }
      double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;           // d/d(vr) P, d/d(vi) P, d/d(vr) Q, d/d(vi) Q
      bool first = true;
      for (int q = h0; q < h1; q++) {                          // :206-214, Sij ascending then Sji ascending
        const int ks = P.he_ks[q], k = ks >> 1, side = ks & 1;
        const double Ir = x[P.o_f[0 + 2 * side] + k], Ii = x[P.o_f[1 + 2 * side] + k];
        accP = (accP + ra * Ir) + ia * Ii;
        accQ = (accQ + ia * Ir) - ra * Ii;
        if (first) { d0 = Ir; d1 = Ii; d2 = -Ii; d3 = Ir; first = false; }
        else { d0 = d0 + Ir; d1 = d1 + Ii; d2 = d2 + (-Ii); d3 = d3 + Ir; }
      }
      if constexpr (FORM == CBRARV) {                          // :264-266
        const double A2 = ra * ra + ia * ia;
        accP = accP - Gl * A2;
        accQ = accQ + Bl * A2;
        const double t0 = -2.0 * Gl * ra, t1 = -2.0 * Gl * ia, t2 = 2.0 * Bl * ra, t3 = 2.0 * Bl * ia;
        if (first) { d0 = t0; d1 = t1; d2 = t2; d3 = t3; }
        else { d0 = d0 + t0; d1 = d1 + t1; d2 = d2 + t2; d3 = d3 + t3; }
      } else {                                                 // :272-274
        const double W = x[P.o_Wd + i];
        accP = accP - Gl * W;
        accQ = accQ + Bl * W;
      }
      if (wd) *reinterpret_cast<double4*>(A.diag + (long long)s * A.lddiag + 4 * i) = make_double4(d0, d1, d2, d3);
      const double* pd = A.Pd ? A.Pd + (long long)s * A.ldl : P.Pd0;
      const double* qd = A.Qd ? A.Qd + (long long)s * A.ldl : P.Qd0;
      const double gP = accP - pd[i], gQ = accQ - qd[i];       // :291-292
      if (wg) {
        double* gd = A.g + (long long)s * A.ldg + 2 * i;
        if ((reinterpret_cast<unsigned long long>(gd) & 15) == 0) *reinterpret_cast<double2*>(gd) = make_double2(gP, gQ);
        else { gd[0] = gP; gd[1] = gQ; }
      }
      vmax = fmax(fabs(gP), fabs(gQ));
      nviol = (fabs(gP) > A.viol_tol ? 1 : 0) + (fabs(gQ) > A.viol_tol ? 1 : 0);
    }
    reduce_scalars(A.scal ? A.scal + (long long)s * NSCAL : nullptr, vmax, nviol);
  }
}

__global__ void __launch_bounds__(PBT) cb_bus_jh_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
  const int t = blockIdx.x * PBT + threadIdx.x, nt = gridDim.x * PBT;
  for (int s = s0; s < s1; s++) {
    if (A.what & EV_J) {
      const double* __restrict__ x = A.x + (long long)s * A.ldx;
      const double* __restrict__ dg = A.diag + (long long)s * A.lddiag;
      emit_sectors(A.J + (long long)s * A.ldJ, P.j_br0, t, nt, P.pb_jt, P.pb_ji, P.pb_jstride,
                   [&](int c) { return c >= 0 ? x[c] : (c == -1 ? 1.0 : dg[-2 - c]); });
    }
    if (A.what & EV_H) {
      const double* __restrict__ mu = A.mu + (long long)s * A.ldmu;
      emit_sectors(A.H + (long long)s * A.ldH, P.h_br0, t, nt, P.pb_ht, P.pb_hi, P.pb_hstride,
                   [&](int c) { return c >= 0 ? mu[c] : 1.0; });
    }
  }
}

__global__ void __launch_bounds__(TPB) objective_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  __shared__ double red[OBJ_LANES / 32];
  const int s0 = blockIdx.x * A.chunk, s1 = min(A.S, s0 + A.chunk);
  objective_tile(P, A, s0, s1, red);
}

// Direct path of the branch rows: thread per branch, no shared-memory staging of J / H and no CTA barrier.
// Inputs of the next scenario are prefetched into registers while the current one is evaluated; the branch's
// J and H blocks (whole 32-byte sectors) go straight from registers to HBM; g goes through a per-warp staging
// row so that it is written with warp-contiguous stores.
// resident CTAs the direct kernel is compiled for: the short-block models (CB*, PN*: <= 178 registers unbounded) gain from
// a third CTA per SM (170 registers, no spills: PNPAPV branch rows 4.4 -> 4.9 TB/s), the PB* fallback (216 registers)
// would spill; four CTAs spill everywhere (3.5 TB/s)
#ifndef PS_DIRECT_MINB
#define PS_DIRECT_MINB 3
#endif
constexpr int direct_minb(int form) { return is_PB(form) ? 1 : PS_DIRECT_MINB; }
#ifndef PS_PREFETCH2
#define PS_PREFETCH2 1
#endif
template <int FORM, int SRC>
__device__ __forceinline__ void branch_direct_body(const DevPlan& P, const EvalArgs& A, int bx) {
  using BR = Branch<FORM, SRC>;
  if constexpr (BR::DIRECT_OK) {
    __shared__ __align__(16) double sW[TPB / 32][32 * BR::NR];     // per-warp staging row of g
    const int u = bx * TPB + threadIdx.x;
    const bool act = u < P.nr;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
    const bool wg = (A.what & EV_G) != 0, wj = (A.what & EV_J) != 0, wh = (A.what & EV_H) != 0;
    typename BR::template OutR<true, false> oj;
    typename BR::template OutR<false, true> oh;
    oj.tol = A.viol_tol; oh.tol = A.viol_tol;
    typename BR::Cst cst{};
    typename BR::In cur{};
    int f = 0, t = 1;
    if (act) {
      f = P.f[u]; t = P.t[u];
      cst = BR::load_cst(P, u);
      cur = BR::load_in(P, A, s0, u, f, t);
    }
    oj.sw = oh.sw = f > t;
    const long long jb = P.j_br0 + (long long)u * BR::RJ, hb = P.h_br0 + (long long)u * BR::RH;
    const int ubase = bx * TPB + w * 32;
    const int nact = max(0, min(32, P.nr - ubase));               // active lanes of this warp
    const int ng = nact * BR::NR;
    const long long gb = P.bus_rows + (long long)ubase * BR::NR;
#if PS_PREFETCH2
    typename BR::In nx1{};
    if (act && s0 + 1 < s1) nx1 = BR::load_in(P, A, s0 + 1, u, f, t);
#endif
    for (int s = s0; s < s1; s++) {
      typename BR::In nxt{};
#if PS_PREFETCH2
      if (act && s + 2 < s1) nxt = BR::load_in(P, A, s + 2, u, f, t);   // two scenarios ahead
#else
      if (act && s + 1 < s1) nxt = BR::load_in(P, A, s + 1, u, f, t);
#endif
      oj.vmax = 0.0; oj.nviol = 0;
      double* sw = sW[w];
      if (wj || wg || A.scal) {
        if (act) BR::run(oj, cst, cur);                            // pass 1: g + Jacobian tags
        if (wj) {
          double* jp = A.J + (long long)s * A.ldJ + jb;
          if constexpr (BR::RJ % 4 == 0) { if (act) oj.storeJ(jp, std::make_integer_sequence<int, BR::RJ / 4>{}); }
          else oj.template store_pair<true, BR::RJ>(jp, act, lane, u + 1 < P.nr);
        }
      }
      if (wh) {
        if (act) BR::run(oh, cst, cur);                            // pass 2: Hessian tags
        double* hp = A.H + (long long)s * A.ldH + hb;
        if constexpr (BR::RH % 4 == 0) { if (act) oh.storeH(hp, std::make_integer_sequence<int, BR::RH / 4>{}); }
        else oh.template store_pair<false, BR::RH>(hp, act, lane, u + 1 < P.nr);
      }
      if (wg) {
        double* sg = sw;
        if (act) {
#pragma unroll
          for (int r = 0; r < BR::NR; r++) sg[lane * BR::NR + r] = oj.gt[r];
        }
        __syncwarp();
        copy_aligned<32>(A.g + (long long)s * A.ldg + gb, ng, lane, [&](int e) { return sg[e]; });
        __syncwarp();
      }
      reduce_scalars(A.scal ? A.scal + (long long)s * NSCAL : nullptr, oj.vmax, oj.nviol);
#if PS_PREFETCH2
      cur = nx1; nx1 = nxt;
#else
      if (s + 1 < s1) cur = nxt;
#endif
    }
  }
}

// Bulk path of the branch rows: the same thread-per-branch evaluation, but every thread parks its J and H block in a
// private, padded shared-memory row (conflict-free 16-byte stores) and hands it to the copy engine with one
// cp.async.bulk per block (TMA, shared -> global).  The SM's load/store units then carry no global stores at all, the
// HBM writes are issued as whole lines, and the copy of scenario s overlaps the evaluation of scenario s + 1 (a row is
// only waited for right before it is overwritten).  No barrier: the rows and the bulk groups are per thread.
// Needs 16-byte aligned blocks (even block lengths, even leading dimensions); see launch_eval.
#ifndef PS_BULK_WARP
#define PS_BULK_WARP 1
#endif
#if PS_BULK_WARP
constexpr int bulk_row_stride(int r) { return r; }   // rows of a warp contiguous: one copy per warp and block
#else
constexpr int bulk_row_stride(int r) { return ((r / 2) & 1) ? r : r + 2; }   // doubles; (stride / 2) odd: conflict-free 16-byte stores
#endif

#ifndef PS_BULK_MINB
#define PS_BULK_MINB 2
#endif
// Rows handed to the copy engine together (one cp.async.bulk per group and block), and which lane evaluates which branch
// of the warp's 32: lane l takes branch (l % NG) * GR + l / NG (NG = 32 / GR groups), so that the eight lanes of a
// quarter-warp -- the unit a 16-byte shared-memory store is served in -- park their rows in eight DIFFERENT groups.  With
// a group stride of GR * row + 2 doubles (an odd number of 16-byte units) the eight 16-byte chunks fall into the eight
// bank groups: conflict-free for every row length when GR = 4; with GR = 8 (the default: half as many copies) a quarter-warp
// holds two rows of each of four groups -- conflict-free for the 12-unit J rows of PBRARV, two-way for its 18-unit H rows.
// (Lanes in branch order: row strides of 12 / 18 units put a quarter-warp on 2 / 4 bank groups -- 59 % of the kernel's
// shared-memory wavefronts were conflicts, profiles/r02_ncu_branch_bulk_summary.txt.)  Measured (gpurun_out/r2c37_ab.log):
// PBRAPV branch rows 4.87 -> 4.70 ms per 2048 scenarios (5.87 -> 6.09 TB/s), PBRARV 8.54 -> 8.51 per 4096 (not bound by
// these stores); GR = 4 and 8 equal.  Global loads and the g rows see the same 32 consecutive branches either way.
#ifndef PS_BULK_GROUP
#define PS_BULK_GROUP 8
#endif
constexpr int BULK_GR = PS_BULK_GROUP, BULK_NG = 32 / BULK_GR;
static_assert(BULK_GR == 4 || BULK_GR == 8, "eight lanes of a quarter-warp on eight (or four x two) groups");
#ifndef PS_BULK_PERMUTE
#define PS_BULK_PERMUTE 1
#endif
__device__ __forceinline__ int bulk_branch_of_lane(int l) { return PS_BULK_PERMUTE ? (l % BULK_NG) * BULK_GR + l / BULK_NG : l; }
template <int FORM, int SRC>
__global__ void __launch_bounds__(TPB, PS_BULK_MINB) branch_bulk_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  using BR = Branch<FORM, SRC>;
  if constexpr (BR::DIRECT_OK) {
    constexpr int SJ = bulk_row_stride(BR::RJ), SH = bulk_row_stride(BR::RH);
    extern __shared__ __align__(16) double rows[];                 // [TPB][SJ] | [TPB][SH]
    __shared__ __align__(16) double sW[TPB / 32][32 * BR::NR];     // per-warp staging row of g
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#if PS_BULK_WARP
    // rows of BULK_GR consecutive branches are contiguous (one copy per group); 16 bytes of padding between the groups
    const int bl = bulk_branch_of_lane(lane), tb = w * 32 + bl;    // this lane's branch within the warp / within the CTA
    double* const rj = rows + tb * SJ + (tb / BULK_GR) * 2;
    double* const rh = rows + TPB * SJ + (TPB / BULK_GR) * 2 + tb * SH + (tb / BULK_GR) * 2;
#else
    double* const rj = rows + threadIdx.x * SJ;
    double* const rh = rows + TPB * SJ + threadIdx.x * SH;
#endif
#if PS_BULK_WARP
    const int u = blockIdx.x * TPB + tb;
#else
    const int u = blockIdx.x * TPB + threadIdx.x;
    const int bl = lane;
#endif
    const bool act = u < P.nr;
    const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
    const bool wg = (A.what & EV_G) != 0, wj = (A.what & EV_J) != 0, wh = (A.what & EV_H) != 0;
    typename BR::template OutR<true, false> oj;
    typename BR::template OutR<false, true> oh;
    oj.tol = A.viol_tol; oh.tol = A.viol_tol;
    typename BR::Cst cst{};
    typename BR::In cur{};
    int f = 0, t = 1;
    if (act) {
      f = P.f[u]; t = P.t[u];
      cst = BR::load_cst(P, u);
      cur = BR::load_in(P, A, s0, u, f, t);
    }
    oj.sw = oh.sw = f > t;
    const long long jb = P.j_br0 + (long long)u * BR::RJ, hb = P.h_br0 + (long long)u * BR::RH;
    const int ubase = blockIdx.x * TPB + w * 32;
    const int nact = max(0, min(32, P.nr - ubase));
    const int ng = nact * BR::NR;
    const long long gb = P.bus_rows + (long long)ubase * BR::NR;
    const bool lead = (bl % BULK_GR) == 0;                         // issues (and waits for) the copies of its group's rows
    const int nrow8 = max(0, min(BULK_GR, P.nr - u));              // rows of this lane's group (meaningful on lead lanes)
#if PS_PREFETCH2
    typename BR::In nx1{};
    if (act && s0 + 1 < s1) nx1 = BR::load_in(P, A, s0 + 1, u, f, t);
#endif
    for (int s = s0; s < s1; s++) {
      typename BR::In nxt{};
#if PS_PREFETCH2
      if (act && s + 2 < s1) nxt = BR::load_in(P, A, s + 2, u, f, t);   // two scenarios ahead
#else
      if (act && s + 1 < s1) nxt = BR::load_in(P, A, s + 1, u, f, t);
#endif
      oj.vmax = 0.0; oj.nviol = 0;
      if (wj || wg || A.scal) {
        if (act) BR::run(oj, cst, cur);                            // pass 1: g + Jacobian tags
#if PS_BULK_WARP
        if (wj) {
          if (lead) bulk_wait_read<1>();                           // the copy that last read these rows (scenario s - 1) is done
          __syncwarp();
          if (act) oj.rowJ(rj, std::make_integer_sequence<int, BR::RJ / 2>{});
          fence_async_smem();
          __syncwarp();
          if (lead && nrow8 > 0) bulk_store(A.J + (long long)s * A.ldJ + jb, rj, nrow8 * BR::RJ * 8);
        }
      }
      if (lead) bulk_commit();
      if (wh) {
        if (act) BR::run(oh, cst, cur);                            // pass 2: Hessian tags
        if (lead) bulk_wait_read<1>();
        __syncwarp();
        if (act) oh.rowH(rh, std::make_integer_sequence<int, BR::RH / 2>{});
        fence_async_smem();
        __syncwarp();
        if (lead && nrow8 > 0) bulk_store(A.H + (long long)s * A.ldH + hb, rh, nrow8 * BR::RH * 8);
      }
      if (lead) bulk_commit();
#else
        if (wj && act) {
          bulk_wait_read<1>();                                     // the copy that last read this row (scenario s - 1) is done
          oj.rowJ(rj, std::make_integer_sequence<int, BR::RJ / 2>{});
          fence_async_smem();
          bulk_store(A.J + (long long)s * A.ldJ + jb, rj, BR::RJ * 8);
        }
      }
      bulk_commit();
      if (wh && act) {
        BR::run(oh, cst, cur);                                     // pass 2: Hessian tags
        bulk_wait_read<1>();
        oh.rowH(rh, std::make_integer_sequence<int, BR::RH / 2>{});
        fence_async_smem();
        bulk_store(A.H + (long long)s * A.ldH + hb, rh, BR::RH * 8);
      }
      bulk_commit();
#endif
      if (wg) {
        double* sg = sW[w];
        if (act) {
#pragma unroll
          for (int r = 0; r < BR::NR; r++) sg[bl * BR::NR + r] = oj.gt[r];
        }
        __syncwarp();
        copy_aligned<32>(A.g + (long long)s * A.ldg + gb, ng, lane, [&](int e) { return sg[e]; });
        __syncwarp();
      }
      reduce_scalars(A.scal ? A.scal + (long long)s * NSCAL : nullptr, oj.vmax, oj.nviol);
#if PS_PREFETCH2
      cur = nx1; nx1 = nxt;
#else
      if (s + 1 < s1) cur = nxt;
#endif
    }
    bulk_wait_read<0>();                                           // shared memory must outlive the copies that read it
  }
}

template <int FORM>
cudaError_t launch_bulk(const DevPlan& P, const EvalArgs& A, cudaStream_t st, dim3 grid) {
  if (P.src == SRC_MATPOWER) {
    using BR = Branch<FORM, SRC_MATPOWER>;
    const size_t smem = (size_t)(TPB * (bulk_row_stride(BR::RJ) + bulk_row_stride(BR::RH)) + (TPB / BULK_GR) * 4) * 8;
    branch_bulk_kernel<FORM, SRC_MATPOWER><<<grid, TPB, smem, st>>>(P, A);
  } else {
    using BR = Branch<FORM, SRC_ARPAE>;
    const size_t smem = (size_t)(TPB * (bulk_row_stride(BR::RJ) + bulk_row_stride(BR::RH)) + (TPB / BULK_GR) * 4) * 8;
    branch_bulk_kernel<FORM, SRC_ARPAE><<<grid, TPB, smem, st>>>(P, A);
  }
  return cudaGetLastError();
}

template <int FORM, int SRC>
__global__ void __launch_bounds__(TPB, direct_minb(FORM)) branch_direct_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  branch_direct_body<FORM, SRC>(P, A, blockIdx.x);
}

// Residual-only pass of the branch rows (eval_constraint alone: every trial point of the SLP line search, and the
// violation scalars): the derivative tags are compile-time dead, so this is the flow / limit expressions and nothing
// else; g leaves through a per-warp staging row as warp-contiguous, sector-aligned stores.  Any layout.
#ifndef PS_BG_MINB
#define PS_BG_MINB 1
#endif
template <int FORM, int SRC>
__global__ void __launch_bounds__(TPB, PS_BG_MINB) branch_g_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ EvalArgs A) {
  using BR = Branch<FORM, SRC>;
  if constexpr (BR::NR > 0) {
    __shared__ __align__(16) double sW[TPB / 32][32 * BR::NR];
    const int u = blockIdx.x * TPB + threadIdx.x;
    const bool act = u < P.nr;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int s0 = blockIdx.y * A.chunk, s1 = min(A.S, s0 + A.chunk);
    const bool wg = (A.what & EV_G) != 0;
    typename BR::template OutR<false, false, true> og;
    og.tol = A.viol_tol;
    typename BR::Cst cst{};
    typename BR::In cur{};
    int f = 0, t = 1;
    if (act) {
      f = P.f[u]; t = P.t[u];
      cst = BR::load_cst(P, u);
      cur = BR::load_in(P, A, s0, u, f, t);
    }
    og.sw = f > t;
    const int ubase = blockIdx.x * TPB + w * 32;
    const int nact = max(0, min(32, P.nr - ubase));
    const int ng = nact * BR::NR;
    const long long gb = P.bus_rows + (long long)ubase * BR::NR;
    // inputs two scenarios ahead (the first use of a scenario's inputs was 32 % of this kernel's stall samples at depth 1:
    // profiles/r02_ncu_gonly_summary.txt)
    typename BR::In nx1{};
    if (act && s0 + 1 < s1) nx1 = BR::load_in(P, A, s0 + 1, u, f, t);
    for (int s = s0; s < s1; s++) {
      typename BR::In nxt{};
      if (act && s + 2 < s1) nxt = BR::load_in(P, A, s + 2, u, f, t);
      og.vmax = 0.0; og.nviol = 0;
      if (act) BR::run(og, cst, cur);
      if (wg) {
        double* sg = sW[w];
        if (act) {
#pragma unroll
          for (int r = 0; r < BR::NR; r++) sg[lane * BR::NR + r] = og.gt[r];
        }
        __syncwarp();
        copy_aligned<32>(A.g + (long long)s * A.ldg + gb, ng, lane, [&](int e) { return sg[e]; });
        __syncwarp();
      }
      reduce_scalars(A.scal ? A.scal + (long long)s * NSCAL : nullptr, og.vmax, og.nviol);
      cur = nx1; nx1 = nxt;
    }
  }
}
template <int FORM>
cudaError_t launch_branch_g_form(const DevPlan& P, const EvalArgs& A, cudaStream_t st, dim3 grid) {
  if (P.src == SRC_MATPOWER) branch_g_kernel<FORM, SRC_MATPOWER><<<grid, TPB, 0, st>>>(P, A);
  else branch_g_kernel<FORM, SRC_ARPAE><<<grid, TPB, 0, st>>>(P, A);
  return cudaGetLastError();
}

template <int FORM>
cudaError_t launch_direct(const DevPlan& P, const EvalArgs& A, cudaStream_t st, dim3 grid) {
  if (P.src == SRC_MATPOWER) branch_direct_kernel<FORM, SRC_MATPOWER><<<grid, TPB, 0, st>>>(P, A);
  else branch_direct_kernel<FORM, SRC_ARPAE><<<grid, TPB, 0, st>>>(P, A);
  return cudaGetLastError();
}

template <int FORM>
cudaError_t launch_form(const DevPlan& P, const EvalArgs& A, size_t smem, cudaStream_t st, dim3 grid) {
  if (P.src == SRC_MATPOWER) branch_staged_kernel<FORM, SRC_MATPOWER><<<grid, TPB, smem, st>>>(P, A);
  else branch_staged_kernel<FORM, SRC_ARPAE><<<grid, TPB, smem, st>>>(P, A);
  return cudaGetLastError();
}

template <int FORM>
cudaError_t launch_bus_form(const DevPlan& P, const EvalArgs& A, size_t smem, cudaStream_t st, dim3 grid) {
  bus_kernel<FORM><<<grid, BTPB, smem, st>>>(P, A);
  return cudaGetLastError();
}

template <int FORM>
cudaError_t launch_bus_warp_form(const DevPlan& P, const EvalArgs& A, size_t smem, cudaStream_t st, dim3 grid) {
  switch (wt_mode(A)) {
    case WT_FULL: bus_warp_kernel<FORM, WT_FULL><<<grid, 32 * WPC, smem, st>>>(P, A); break;
    case WT_GJ: bus_warp_kernel<FORM, WT_GJ><<<grid, 32 * WPC, smem, st>>>(P, A); break;
    case WT_FULL | WT_ST: bus_warp_kernel<FORM, WT_FULL | WT_ST><<<grid, 32 * WPC, smem, st>>>(P, A); break;
    case WT_GJ | WT_ST: bus_warp_kernel<FORM, WT_GJ | WT_ST><<<grid, 32 * WPC, smem, st>>>(P, A); break;
    default: bus_warp_kernel<FORM, WT_ANY><<<grid, 32 * WPC, smem, st>>>(P, A); break;
  }
  return cudaGetLastError();
}

template <int FORM>
cudaError_t configure_form(int src, size_t smem_branch, size_t smem_bus, size_t smem_wt, size_t smem_pbs) {
  cudaError_t e = cudaFuncSetAttribute(bus_kernel<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bus);
  if (e != cudaSuccess) return e;
  if constexpr (FORM == PBRAPV || FORM == PBRARV) {
    if (smem_pbs > 0) {
      e = cudaFuncSetAttribute(pb_bus_g_scn_kernel<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pbs);
      if (e != cudaSuccess) return e;
    }
  }
  if (smem_wt > 0) {
    e = cudaFuncSetAttribute(bus_warp_kernel<FORM, WT_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wt);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bus_warp_kernel<FORM, WT_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wt);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bus_warp_kernel<FORM, WT_GJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wt);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bus_warp_kernel<FORM, WT_FULL | WT_ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wt);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bus_warp_kernel<FORM, WT_GJ | WT_ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wt);
    if (e != cudaSuccess) return e;
  }
  {
    using BM = Branch<FORM, SRC_MATPOWER>;
    using BA = Branch<FORM, SRC_ARPAE>;
    if (src == SRC_MATPOWER && BM::DIRECT_OK)
      e = cudaFuncSetAttribute(branch_bulk_kernel<FORM, SRC_MATPOWER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (TPB * (bulk_row_stride(BM::RJ) + bulk_row_stride(BM::RH)) + (TPB / BULK_GR) * 4) * 8);
    if (src != SRC_MATPOWER && BA::DIRECT_OK)
      e = cudaFuncSetAttribute(branch_bulk_kernel<FORM, SRC_ARPAE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (TPB * (bulk_row_stride(BA::RJ) + bulk_row_stride(BA::RH)) + (TPB / BULK_GR) * 4) * 8);
    if (e != cudaSuccess) return e;
  }
  if (src == SRC_MATPOWER)
    return cudaFuncSetAttribute(branch_staged_kernel<FORM, SRC_MATPOWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_branch);
  return cudaFuncSetAttribute(branch_staged_kernel<FORM, SRC_ARPAE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_branch);
}

__global__ void __launch_bounds__(256) csc_assemble_kernel(const int32_t* __restrict__ segptr, const int32_t* __restrict__ perm,
                                                           int nnz, const double* __restrict__ dE, long long ldE,
                                                           double* __restrict__ nz, long long ldnz) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= nnz) return;
  const int q0 = segptr[j], q1 = segptr[j + 1];
  const long long s = blockIdx.y + (long long)blockIdx.z * gridDim.y;
  const double* __restrict__ e = dE + s * ldE;
  double acc = 0.0;                                           // common.jl:16-18: J[r,c] += dE[i], in COO order
  for (int q = q0; q < q1; q++) acc += e[perm[q]];
  nz[s * ldnz + j] = acc;
}

__global__ void __launch_bounds__(256) permute_rows_kernel(const int32_t* __restrict__ perm, int n, const double* __restrict__ in,
                                                           long long ldin, double* __restrict__ out, long long ldout) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int src = perm[i];
  const long long s = blockIdx.y + (long long)blockIdx.z * gridDim.y;
  out[s * ldout + i] = in[s * ldin + src];
}

// ---- affine + quadratic rows (MOI_wrapper.jl:864-891, 973-1002, 1030-1042) -------------------------------------
__global__ void __launch_bounds__(256) rows_g_kernel(const RowsDev R, const double* __restrict__ x, long long ldx,
                                                     double* __restrict__ g, long long ldg) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= R.m) return;
  const double* __restrict__ xs = x + (long long)blockIdx.y * ldx;
  double v = R.cst[r];                                         // eval_function: constant, affine terms, quadratic terms
  for (int q = R.aptr[r]; q < R.aptr[r + 1]; q++) v += R.acoef[q] * xs[R.avar[q]];
  for (int q = R.qptr[r]; q < R.qptr[r + 1]; q++) {
    const int i = R.qv1[q], j = R.qv2[q];
    if (i == j) v += 0.5 * R.qcoef[q] * xs[i] * xs[j];
    else v += R.qcoef[q] * xs[i] * xs[j];
  }
  g[(long long)blockIdx.y * ldg + r] = v;
}
__global__ void __launch_bounds__(256) rows_j_kernel(const RowsDev R, const double* __restrict__ x, long long ldx,
                                                     double* __restrict__ J, long long ldJ) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= R.nnzj) return;
  const int ix = R.jx[e];
  const double c = R.jc[e];
  J[(long long)blockIdx.y * ldJ + e] = ix < 0 ? c : c * x[(long long)blockIdx.y * ldx + ix];
}
__global__ void __launch_bounds__(256) rows_h_kernel(const RowsDev R, const double* __restrict__ lam, long long ldl,
                                                     double* __restrict__ H, long long ldH) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= R.nnzh) return;
  H[(long long)blockIdx.y * ldH + e] = lam[(long long)blockIdx.y * ldl + R.hrow[e]] * R.hc[e];
}

// ---- KKT / merit metrics: one CTA per scenario, fixed-shape reductions (deterministic) ---------------------------
constexpr int KT = 256;
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < KT / 32; w++) t += red[w];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = red[0];
  for (int w = 1; w < KT / 32; w++) t = fmax(t, red[w]);
  return t;
}

__global__ void __launch_bounds__(KT) kkt_kernel(const KktDev K, const KktArgs A) {
  __shared__ double red[KT / 32];
  const long long s = blockIdx.x;
  const double* __restrict__ E = A.E + s * A.ldE;
  const double* __restrict__ x = A.x + s * A.ldx;
  const double* __restrict__ lam = A.lam + s * A.ldl;
  const double* __restrict__ nz = A.nz + s * A.ldnz;
  const double* __restrict__ df = A.df + s * A.lddf;
  const double* __restrict__ nu = A.nu ? A.nu + s * A.ldnu : nullptr;
  const double* __restrict__ pd = A.p ? A.p + s * A.ldp : nullptr;
  double v1 = 0.0, vinf = 0.0, cinf = 0.0, den = 0.0, pen = 0.0, scal = 0.0;
  for (int i = threadIdx.x; i < K.m; i += KT) {
    const double e = E[i], lo = K.gL[i], hi = K.gU[i], l = lam[i];
    double viol = 0.0;                                          // common.jl:84-90
    if (e > hi) viol = e - hi;
    else if (e < lo) viol = lo - e;
    v1 += viol; vinf = fmax(vinf, viol);
    if (nu) pen += nu[i] * viol;                                // slp.jl:124-128 / 155-159
    if (lo != hi) {                                             // common.jl:59-66
      double c = fmin(e - lo, hi - e) * l;
      if (c != c) c = 0.0;
      cinf = fmax(cinf, fabs(c));
      den += l * l;
    }
    double rs = 0.0;                                            // ||J[i,:]||^2 (common.jl:40-42)
    for (int q = K.rptr[i]; q < K.rptr[i + 1]; q++) { const double a = nz[K.rperm[q]]; rs += a * a; }
    scal = fmax(scal, fabs(l) * sqrt(rs));
  }
  double kt = 0.0, dn = 0.0, dp = 0.0;
  const double* __restrict__ mU = A.mU + s * A.ldm;
  const double* __restrict__ mL = A.mL + s * A.ldm;
  for (int j = threadIdx.x; j < K.n; j += KT) {
    const double xj = x[j];
    double viol = 0.0;                                          // common.jl:91-97
    if (xj > K.xU[j]) viol = xj - K.xU[j];
    else if (xj < K.xL[j]) viol = K.xL[j] - xj;
    v1 += viol; vinf = fmax(vinf, viol);
    double jt = 0.0;                                            // (J' lambda)_j over the column's stored entries
    for (int q = K.colptr[j]; q < K.colptr[j + 1]; q++) jt += nz[q] * lam[K.rowval[q]];
    const double r = df[j] - jt - mU[j] - mL[j];                // common.jl:38
    kt += r * r;
    dn += df[j] * df[j];
    if (pd) dp += df[j] * pd[j];                                // slp.jl:153
  }
  const double V1 = block_sum(v1, red), VI = block_max(vinf, red), CI = block_max(cinf, red), DEN = block_sum(den, red);
  const double PEN = block_sum(pen, red), SC = block_max(scal, red), KTS = block_sum(kt, red), DN = block_sum(dn, red);
  const double DP = block_sum(dp, red);
  if (threadIdx.x == 0) {
    double* o = A.out + s * NKKT;
    const double ndf = sqrt(DN);
    o[0] = V1; o[1] = VI;
    o[2] = CI / (1.0 + sqrt(DEN));                              // common.jl:68
    o[3] = sqrt(KTS) / fmax(fmax(1.0, ndf), SC);                // common.jl:38-43
    o[4] = (A.f ? A.f[s] : 0.0) + PEN;                          // slp.jl:121-129
    o[5] = DP - PEN;                                            // slp.jl:153-160
    o[6] = ndf; o[7] = SC;
  }
}

}  // namespace

cudaError_t launch_kkt(const KktDev& K, const KktArgs& A, int64_t S, cudaStream_t stream) {
  if (S <= 0) return cudaSuccess;
  kkt_kernel<<<(unsigned)S, KT, 0, stream>>>(K, A);
  return cudaGetLastError();
}

cudaError_t launch_rows_eval(const RowsDev& R, int what, int64_t S, const double* x, long long ldx, const double* lam,
                             long long ldl, double* g, long long ldg, double* J, long long ldJ, double* H, long long ldH,
                             cudaStream_t stream) {
  for (int64_t s0 = 0; s0 < S; s0 += 65535) {                  // gridDim.y limit
    const unsigned ny = (unsigned)std::min<int64_t>(65535, S - s0);
    if ((what & EV_G) && R.m > 0)
      rows_g_kernel<<<dim3((R.m + 255) / 256, ny), 256, 0, stream>>>(R, x + s0 * ldx, ldx, g + s0 * ldg, ldg);
    if ((what & EV_J) && R.nnzj > 0)
      rows_j_kernel<<<dim3((R.nnzj + 255) / 256, ny), 256, 0, stream>>>(R, x + s0 * ldx, ldx, J + s0 * ldJ, ldJ);
    if ((what & EV_H) && R.nnzh > 0)
      rows_h_kernel<<<dim3((R.nnzh + 255) / 256, ny), 256, 0, stream>>>(R, lam + s0 * ldl, ldl, H + s0 * ldH, ldH);
  }
  return cudaGetLastError();
}

int64_t wt_smem_bytes(const DevPlan& P) {
  if (P.n_wtiles <= 0) return 0;
  return (int64_t)WPC * wt_region_doubles(P) * 8;
}

// the scenario-resident PB* residual kernel needs the (P, Q) sums of every bus in one CTA's shared memory
int64_t pbs_smem_bytes(const DevPlan& P) {
  const int64_t b = (int64_t)P.nb * 16;
  const int64_t steps = (P.ng + PBS_WAVE - 1) / PBS_WAVE + 2 * ((P.nr + PBS_WAVE - 1) / PBS_WAVE);
  return (P.pbs_ok && b <= 222 * 1024 && steps + 16 <= PBS_MAX_STEPS) ? b : 0;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : 0;
  return v > 0 ? v : dflt;
}

cudaError_t configure_eval(int form, int src, int64_t smem_branch_bytes, int64_t smem_bus_bytes, int64_t smem_wt_bytes,
                           int64_t smem_pbs_bytes) {
  const size_t sb = (size_t)smem_branch_bytes, su = (size_t)smem_bus_bytes, sw = (size_t)smem_wt_bytes, sp = (size_t)smem_pbs_bytes;
  switch (form) {
    case PBRAPV: return configure_form<PBRAPV>(src, sb, su, sw, sp);
    case PBRARV: return configure_form<PBRARV>(src, sb, su, sw, sp);
    case PBRAWV: return configure_form<PBRAWV>(src, sb, su, sw, sp);
    case CBRARV: return configure_form<CBRARV>(src, sb, su, sw, sp);
    case CBRAWV: return configure_form<CBRAWV>(src, sb, su, sw, sp);
    case PNPAPV: return configure_form<PNPAPV>(src, sb, su, sw, sp);
    case PNRAPV: return configure_form<PNRAPV>(src, sb, su, sw, sp);
    case PNRARV: return configure_form<PNRARV>(src, sb, su, sw, sp);
    case PNRAWV: return configure_form<PNRAWV>(src, sb, su, sw, sp);
  }
  return cudaErrorInvalidValue;
}

#define PS_FORM_SWITCH(CALL)                                   \
  switch (P.form) {                                            \
    case PBRAPV: return CALL(PBRAPV);                          \
    case PBRARV: return CALL(PBRARV);                          \
    case PBRAWV: return CALL(PBRAWV);                          \
    case CBRARV: return CALL(CBRARV);                          \
    case CBRAWV: return CALL(CBRAWV);                          \
    case PNPAPV: return CALL(PNPAPV);                          \
    case PNRAPV: return CALL(PNRAPV);                          \
    case PNRARV: return CALL(PNRARV);                          \
    case PNRAWV: return CALL(PNRAWV);                          \
  }                                                            \
  return cudaErrorInvalidValue;

static cudaError_t launch_staged(const DevPlan& P, const EvalArgs& A, size_t smem, cudaStream_t stream, dim3 grid) {
#define PS_CALL(F) launch_form<F>(P, A, smem, stream, grid)
  PS_FORM_SWITCH(PS_CALL)
#undef PS_CALL
}
static cudaError_t launch_bus(const DevPlan& P, const EvalArgs& A, size_t smem, cudaStream_t stream, dim3 grid) {
#define PS_CALL(F) launch_bus_form<F>(P, A, smem, stream, grid)
  PS_FORM_SWITCH(PS_CALL)
#undef PS_CALL
}
template <int FORM>
cudaError_t launch_bus_g_list_form(const DevPlan& P, const EvalArgs& A, cudaStream_t st, dim3 grid) {
  bus_g_list_kernel<FORM><<<grid, PS_TPB, 0, st>>>(P, A);
  return cudaGetLastError();
}
static cudaError_t launch_bus_g_list(const DevPlan& P, const EvalArgs& A, cudaStream_t stream, dim3 grid) {
#define PS_CALL(F) launch_bus_g_list_form<F>(P, A, stream, grid)
  PS_FORM_SWITCH(PS_CALL)
#undef PS_CALL
}
static cudaError_t launch_bus_warp(const DevPlan& P, const EvalArgs& A, size_t smem, cudaStream_t stream, dim3 grid) {
#define PS_CALL(F) launch_bus_warp_form<F>(P, A, smem, stream, grid)
  PS_FORM_SWITCH(PS_CALL)
#undef PS_CALL
}
template <int FORM>
cudaError_t launch_pb_form(const DevPlan& P, const EvalArgs& A, cudaStream_t st, int nchunks, int64_t* launches,
                           cudaEvent_t* ev, unsigned* ev_mask) {
  auto rec = [&](int id, int end) { if (ev) { cudaEventRecord(ev[2 * id + end], st); if (end && ev_mask) *ev_mask |= 1u << id; } };
  if constexpr (FORM == PBRAPV || FORM == PBRARV) {
    int obj_from = 0;                                          // scenarios [0, obj_from) got their objective scalar from the bus kernel
    if ((A.what & EV_G) || A.scal) {
      rec(K_BUS, 0);
      // scenario-resident residuals when a batch fills the machine (a CTA works through whole scenarios one after the
      // other: a handful of scenarios is faster with the list-walking kernel); PS_PBSCN=1 / PS_NO_PBSCN force either
      const int64_t smem = pbs_smem_bytes(P);
      const bool scn = smem > 0 && !getenv("PS_NO_PBSCN") && ((A.S >= 32 && P.pbs_prefer) || getenv("PS_PBSCN"));
      if (scn) {
        static int n_sm = 0;
        if (n_sm == 0) { int dv = 0; cudaGetDevice(&dv); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dv); if (n_sm <= 0) n_sm = 148; }
        const int per_sm = (int)std::min<int64_t>(2, (220 * 1024) / smem);   // 1024-thread CTAs: at most 2 per SM
        const int slots = n_sm * per_sm;
        // a CTA works through whole scenarios: S scenarios take ceil(S / slots) rounds.  When the last round would be
        // mostly empty (512 scenarios on 148 SMs: 3.46 rounds) the remainder goes to the list-walking kernel instead --
        // same results bit for bit, disjoint scenario rows
        const int rem = A.S % slots;
        const int n_scn = (A.S > slots && rem > 0 && rem < (slots * 3) / 5) ? A.S - rem : A.S;
        EvalArgs As = A;
        As.S = n_scn;
        pb_bus_g_scn_kernel<FORM><<<dim3((unsigned)std::min(n_scn, slots)), PBS_T, (size_t)smem, st>>>(P, As);
        obj_from = n_scn;
        if (n_scn < A.S) {
          EvalArgs Ar = A;
          const long long o = n_scn;
          Ar.S = A.S - n_scn;
          Ar.x += o * A.ldx;
          if (Ar.Pd) Ar.Pd += o * A.ldl;
          if (Ar.Qd) Ar.Qd += o * A.ldl;
          if (Ar.status) Ar.status += o * A.ldst;
          if (Ar.g) Ar.g += o * A.ldg;
          if (Ar.scal) Ar.scal += o * NSCAL;
          const int nch = (Ar.S + Ar.chunk - 1) / Ar.chunk;
          pb_bus_g_kernel<FORM><<<dim3((unsigned)((P.nb + PBT - 1) / PBT), (unsigned)nch), PBT, 0, st>>>(P, Ar);
          if (launches) *launches += 1;
        }
      } else {
        pb_bus_g_kernel<FORM><<<dim3((unsigned)((P.nb + PBT - 1) / PBT), (unsigned)nchunks), PBT, 0, st>>>(P, A);
      }
      rec(K_BUS, 1);
      if (launches) *launches += 1;
    }
    if (A.what & (EV_J | EV_H)) {
      const int n = (P.j_br0 > P.h_br0 ? P.j_br0 : P.h_br0) / 4 + 1;
      rec(K_BUS_JH, 0);
      pb_bus_jh_kernel<FORM><<<dim3((unsigned)((n + PBT - 1) / PBT), (unsigned)nchunks), PBT, 0, st>>>(P, A);
      rec(K_BUS_JH, 1);
      if (launches) *launches += 1;
    }
    if (A.scal && obj_from < A.S) {
      EvalArgs Ao = A;
      Ao.S = A.S - obj_from;
      Ao.x += (long long)obj_from * A.ldx;
      Ao.scal += (long long)obj_from * NSCAL;
      Ao.chunk = Ao.S >= 65535 * 2 ? (Ao.S + 65534) / 65535 : (Ao.S > 1184 ? 2 : 1);   // one CTA per scenario while they fit a wave or two
      rec(K_OBJ, 0);
      objective_kernel<<<dim3((unsigned)((Ao.S + Ao.chunk - 1) / Ao.chunk)), TPB, 0, st>>>(P, Ao);
      rec(K_OBJ, 1);
      if (launches) *launches += 1;
    }
    return cudaGetLastError();
  } else {
    return cudaErrorInvalidValue;
  }
}
static cudaError_t launch_pb(const DevPlan& P, const EvalArgs& A, cudaStream_t stream, int nchunks, int64_t* launches,
                             cudaEvent_t* ev, unsigned* ev_mask) {
#define PS_CALL(F) launch_pb_form<F>(P, A, stream, nchunks, launches, ev, ev_mask)
  PS_FORM_SWITCH(PS_CALL)
#undef PS_CALL
}
static cudaError_t launch_cb(const DevPlan& P, const EvalArgs& A, cudaStream_t st, int nchunks, int64_t* launches,
                             cudaEvent_t* ev, unsigned* ev_mask) {
  auto rec = [&](int id, int end) { if (ev) { cudaEventRecord(ev[2 * id + end], st); if (end && ev_mask) *ev_mask |= 1u << id; } };
  if ((A.what & (EV_G | EV_J)) || A.scal) {
    const dim3 grid((unsigned)((P.nb + PBT - 1) / PBT), (unsigned)nchunks);
    rec(K_BUS, 0);
    if (P.form == CBRARV) cb_bus_g_kernel<CBRARV><<<grid, PBT, 0, st>>>(P, A);
    else cb_bus_g_kernel<CBRAWV><<<grid, PBT, 0, st>>>(P, A);
    rec(K_BUS, 1);
    if (launches) *launches += 1;
  }
  if (A.what & (EV_J | EV_H)) {
    const int n = (P.j_br0 > P.h_br0 ? P.j_br0 : P.h_br0) / 4 + 1;
    rec(K_BUS_JH, 0);
    cb_bus_jh_kernel<<<dim3((unsigned)((n + PBT - 1) / PBT), (unsigned)nchunks), PBT, 0, st>>>(P, A);
    rec(K_BUS_JH, 1);
    if (launches) *launches += 1;
  }
  if (A.scal) {
    rec(K_OBJ, 0);
    objective_kernel<<<dim3((unsigned)nchunks), TPB, 0, st>>>(P, A);
    rec(K_OBJ, 1);
    if (launches) *launches += 1;
  }
  return cudaGetLastError();
}
static cudaError_t launch_branch_bulk(const DevPlan& P, const EvalArgs& A, cudaStream_t stream, dim3 grid) {
#define PS_CALL(F) launch_bulk<F>(P, A, stream, grid)
  PS_FORM_SWITCH(PS_CALL)
#undef PS_CALL
}
static cudaError_t launch_branch_g(const DevPlan& P, const EvalArgs& A, cudaStream_t stream, dim3 grid) {
#define PS_CALL(F) launch_branch_g_form<F>(P, A, stream, grid)
  PS_FORM_SWITCH(PS_CALL)
#undef PS_CALL
}
static cudaError_t launch_branch_direct(const DevPlan& P, const EvalArgs& A, cudaStream_t stream, dim3 grid) {
#define PS_CALL(F) launch_direct<F>(P, A, stream, grid)
  PS_FORM_SWITCH(PS_CALL)
#undef PS_CALL
}

static bool sector_aligned(const double* p, long long first, long long ld) {
  return ((reinterpret_cast<unsigned long long>(p) >> 3) + (unsigned long long)first) % 4 == 0 && ld % 4 == 0;
}

struct KTimer {                        // records an event pair around one kernel when timing is enabled
  cudaEvent_t* ev; unsigned* mask; cudaStream_t st; int id;
  KTimer(cudaEvent_t* e, unsigned* m, cudaStream_t s, int i) : ev(e), mask(m), st(s), id(i) { if (ev) cudaEventRecord(ev[2 * id], st); }
  ~KTimer() { if (ev) { cudaEventRecord(ev[2 * id + 1], st); if (mask) *mask |= 1u << id; } }
};

// One evaluation = two launches on `stream`: the bus kernel (nodal balances + objective scalar) and the branch
// kernel (flow definitions / limits / cones) -- direct when the branch blocks of J and H start on 32-byte
// sectors in every scenario row, staged otherwise.
cudaError_t launch_eval(const DevPlan& P, const EvalArgs& A0, int64_t smem_branch_bytes, int64_t smem_bus_bytes,
                        cudaStream_t stream, int64_t* launches, cudaEvent_t* ev, unsigned* ev_mask) {
  if (A0.S <= 0) return cudaSuccess;
  EvalArgs A = A0;
  cudaError_t e = cudaSuccess;
  const bool pb_fast = P.pb_slots && !getenv("PS_NO_PBFAST");
  const bool cb_fast = P.cb_slots && !A.status && (A.diag || !(A.what & EV_J)) && !getenv("PS_NO_CBFAST");
  const int min_chunk = (A.S + 65534) / 65535;                 // gridDim.y <= 65535: very long batches take longer chunks
  if (cb_fast) {
    A.chunk = std::max(env_int("PS_CHUNK_PB", 4), min_chunk);
    const int nchunks = (A.S + A.chunk - 1) / A.chunk;
    e = launch_cb(P, A, stream, nchunks, launches, ev, ev_mask);
    if (e != cudaSuccess) return e;
  } else if (pb_fast) {
    A.chunk = std::max(env_int("PS_CHUNK_PB", 4), min_chunk);
    const int nchunks = (A.S + A.chunk - 1) / A.chunk;
    e = launch_pb(P, A, stream, nchunks, launches, ev, ev_mask);
    if (e != cudaSuccess) return e;
  } else if (P.n_bus_tiles > 0 || A.scal) {
    // bus tiles are few and pay a per-CTA staging cost: longer scenario chunks than the branch kernel
    const int64_t smem_wt = wt_smem_bytes(P);
    const bool warp_tiles = smem_wt > 0 && smem_wt <= 200 * 1024 && !getenv("PS_NO_WARPBUS");
    const long long want = 148LL * 4 * 6;
    long long c = ((long long)A.S * (P.n_bus_tiles + 1)) / want;
    const long long cw = ((long long)A.S * ((P.n_wtiles + WPC - 1) / WPC)) / (148LL * PS_WT_MINB * 8);
    A.chunk = warp_tiles ? env_int("PS_CHUNK_WT", (int)(cw < 1 ? 1 : (cw > 16 ? 16 : cw))) : env_int("PS_CHUNK_BUS", (int)(c < 1 ? 1 : (c > 32 ? 32 : c)));
    A.chunk = std::max(A.chunk, min_chunk);
    const int nchunks = (A.S + A.chunk - 1) / A.chunk;
    // residuals only: thread per bus (PS_NO_GLIST: the warp-tile kernel's residual-only instantiation)
    const bool g_list = !(A.what & (EV_J | EV_H)) && !A.force_any && !getenv("PS_NO_GLIST") && !getenv("PS_NO_WARPBUS") && P.n_bus_tiles > 0;
    if (g_list) {
      if (launches) *launches += 1;
      KTimer kt(ev, ev_mask, stream, K_BUS);
      EvalArgs Ag = A;
      Ag.chunk = std::max(env_int("PS_CHUNK_GL", 4), min_chunk);
      e = launch_bus_g_list(P, Ag, stream, dim3((unsigned)((P.nb + PS_TPB - 1) / PS_TPB), (unsigned)((A.S + Ag.chunk - 1) / Ag.chunk)));
    } else if (P.n_bus_tiles > 0) {
      if (launches) *launches += 1;
      KTimer kt(ev, ev_mask, stream, K_BUS);
      if (warp_tiles) {
        e = launch_bus_warp(P, A, (size_t)smem_wt, stream, dim3((unsigned)((P.n_wtiles + WPC - 1) / WPC), (unsigned)nchunks));
        if (e == cudaSuccess && P.n_hub_tiles > 0) {     // the CTA tiles with a bus of more than 32 half-edges
          DevPlan Ph = P;
          Ph.tile_map = P.hub_tiles;
          if (launches) *launches += 1;
          e = launch_bus(Ph, A, (size_t)smem_bus_bytes, stream, dim3((unsigned)P.n_hub_tiles, (unsigned)nchunks));
        }
      } else e = launch_bus(P, A, (size_t)smem_bus_bytes, stream, dim3((unsigned)P.n_bus_tiles, (unsigned)nchunks));
    }
    if (e != cudaSuccess) return e;
    if (A.scal) {                                  // objective scalar: its own small launch (fixed reduction shape)
      if (launches) *launches += 1;
      KTimer kt(ev, ev_mask, stream, K_OBJ);
      EvalArgs Ao = A;                             // one CTA per few scenarios, whatever the bus kernel's chunk is
      Ao.chunk = A.S >= 65535 * 2 ? (A.S + 65534) / 65535 : 2;
      objective_kernel<<<dim3((unsigned)((A.S + Ao.chunk - 1) / Ao.chunk)), TPB, 0, stream>>>(P, Ao);
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
  }
  if (P.nr == 0 || !(A.what & (EV_G | EV_J | EV_H)) ) {
    if (P.nr == 0 || !A.scal) return e;
  }
  A.chunk = std::max(A0.chunk, min_chunk);
  const int nchunks = (A.S + A.chunk - 1) / A.chunk;
  const unsigned nbt = (unsigned)((P.nr + TPB - 1) / TPB);
  bool direct = direct_ok(P.form, P.src) && !getenv("PS_NO_DIRECT");
  if (direct && (A.what & EV_J)) direct = sector_aligned(A.J, P.j_br0, A.ldJ);
  if (direct && (A.what & EV_H)) direct = sector_aligned(A.H, P.h_br0, A.ldH);
  // bulk (TMA store) path: 16-byte aligned blocks are enough
  auto pair_aligned = [](const double* p, long long first, long long ld) {
    return ((reinterpret_cast<unsigned long long>(p) >> 3) + (unsigned long long)first) % 2 == 0 && ld % 2 == 0;
  };
  // ... and it pays for long blocks only (PB*: 60-68 doubles per branch); for the short blocks of the CB* / PN* models the
  // per-warp copy set-up is not amortised and the direct path is a few percent faster (PS_BULK=1 forces it)
  const BranchPat bp = branch_pattern(P.form, P.src);
  bool bulk = direct_ok(P.form, P.src) && !getenv("PS_NO_BULK") && (bp.nj() + bp.nh() >= 48 || getenv("PS_BULK"));
  if (bulk && (A.what & EV_J)) bulk = pair_aligned(A.J, P.j_br0, A.ldJ);
  if (bulk && (A.what & EV_H)) bulk = pair_aligned(A.H, P.h_br0, A.ldH);
  if (launches) *launches += 1;
  KTimer kt(ev, ev_mask, stream, K_BRANCH);
  if (!(A.what & (EV_J | EV_H)) && !getenv("PS_NO_GONLY")) return launch_branch_g(P, A, stream, dim3(nbt, (unsigned)nchunks));
  if (bulk) return launch_branch_bulk(P, A, stream, dim3(nbt, (unsigned)nchunks));
  if (direct) return launch_branch_direct(P, A, stream, dim3(nbt, (unsigned)nchunks));
  return launch_staged(P, A, (size_t)smem_branch_bytes, stream, dim3(nbt, (unsigned)nchunks));
}

cudaError_t launch_permute_rows(const int32_t* perm, int n, int64_t S, const double* in, long long ldin, double* out,
                                long long ldout, cudaStream_t stream) {
  if (n <= 0 || S <= 0) return cudaSuccess;
  for (int64_t s0 = 0; s0 < S; s0 += 65535) {                  // gridDim.y limit
    const unsigned ny = (unsigned)std::min<int64_t>(65535, S - s0);
    permute_rows_kernel<<<dim3((n + 255) / 256, ny, 1), 256, 0, stream>>>(perm, n, in + s0 * ldin, ldin, out + s0 * ldout, ldout);
  }
  return cudaGetLastError();
}

cudaError_t launch_csc_assemble(const int32_t* segptr, const int32_t* perm, int nnz_csc, int64_t S, const double* dE,
                                long long ldE, double* nz, long long ldnz, cudaStream_t stream) {
  if (nnz_csc <= 0 || S <= 0) return cudaSuccess;
  const unsigned gy = (unsigned)(S < 65535 ? S : 65535);
  const unsigned gz = (unsigned)((S + gy - 1) / gy);
  if (gz > 65535) return cudaErrorInvalidConfiguration;
  // a partial last pass would run past S: require S to tile exactly, else launch the remainder separately
  const int64_t full = (int64_t)gy * (gz - (S % gy ? 1 : 0));
  if (full > 0) {
    const dim3 grid((nnz_csc + 255) / 256, gy, (unsigned)(full / gy));
    csc_assemble_kernel<<<grid, 256, 0, stream>>>(segptr, perm, nnz_csc, dE, ldE, nz, ldnz);
  }
  if (full < S) {
    const dim3 grid((nnz_csc + 255) / 256, (unsigned)(S - full), 1);
    csc_assemble_kernel<<<grid, 256, 0, stream>>>(segptr, perm, nnz_csc, dE + full * ldE, ldE, nz + full * ldnz, ldnz);
  }
  return cudaGetLastError();
}

}  // namespace ps
