// Row patterns of the NL constraint block, shared by the host plan builder (ps_plan.cpp) and the
// CUDA kernels (ps_kernels.cu).  One table per (formulation, source_type) describes, for the rows
// a *branch* produces, which variables ("roles") each row touches and which variable pairs carry
// a Hessian slot -- restated from src/opf/formulations.jl:299-428 with JuMP's structure rules
// (SURVEY.md Appendix C.2/C.3).  The order of entries inside a row below is the *tag* order the
// kernels emit values in; the builder turns tags into slot positions (sorted columns etc.).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define PS_HD __host__ __device__
#else
#define PS_HD
#endif

namespace ps {

enum Form : int { PBRAPV = 0, PBRARV, PBRAWV, CBRARV, CBRAWV, PNPAPV, PNRAPV, PNRARV, PNRAWV, NFORM };
enum Src : int { SRC_MATPOWER = 0, SRC_ARPAE = 1 };
enum Flags : int { FLAG_FACTS = 1, FLAG_OBJ_LINEAR = 2 };
enum What : int { EV_G = 1, EV_J = 2, EV_H = 4 };

// variables a branch row can touch
enum Role : int8_t {
  R_AF = 0,  // first voltage block at f  (theta_f | vi_f)   formulations.jl:23,26
  R_AT,      // first voltage block at t
  R_BF,      // second voltage block at f (V_f | vr_f)       formulations.jl:24,27
  R_BT,
  R_WDF, R_WDT, R_WR, R_WI,      // formulations.jl:30,33,34
  R_F0, R_F1, R_F2, R_F3,        // Pij,Qij,Pji,Qji (:47-50) or Irij,Iiij,Irji,Iiji (:53-56)
};

PS_HD constexpr bool is_polar(int f) { return f == PBRAPV || f == PNPAPV || f == PNRAPV; }
PS_HD constexpr bool has_W(int f) { return f == PBRAWV || f == CBRAWV || f == PNRAWV; }
PS_HD constexpr bool has_WrWi(int f) { return f == PBRAWV || f == PNRAWV; }
PS_HD constexpr bool is_PB(int f) { return f == PBRAPV || f == PBRARV || f == PBRAWV; }
PS_HD constexpr bool is_CB(int f) { return f == CBRARV || f == CBRAWV; }
PS_HD constexpr bool is_PN(int f) { return f == PNPAPV || f == PNRAPV || f == PNRARV; }
PS_HD constexpr bool has_bus_rows(int f) { return f != PBRAWV && f != PNRAWV; }  // :290-296

struct RowPat {
  int8_t eq;            // 1: '==' row (bounds 0,0); 0: '<=' row (bounds -inf,0)
  int8_t nj;
  int8_t j[5];
  int8_t nh;
  int8_t h[10][2];
};

struct BranchPat {
  int nrows;
  RowPat rows[6];
  PS_HD constexpr int jtag0(int r) const { int s = 0; for (int q = 0; q < r; q++) s += rows[q].nj; return s; }
  PS_HD constexpr int htag0(int r) const { int s = 0; for (int q = 0; q < r; q++) s += rows[q].nh; return s; }
  PS_HD constexpr int nj() const { return jtag0(nrows); }
  PS_HD constexpr int nh() const { return htag0(nrows); }
};

namespace detail {
PS_HD constexpr void set_j(RowPat& r, int8_t a, int8_t b = -1, int8_t c = -1, int8_t d = -1, int8_t e = -1) {
  int8_t v[5] = {a, b, c, d, e};
  r.nj = 0;
  for (int q = 0; q < 5; q++) if (v[q] >= 0) r.j[r.nj++] = v[q];
}
PS_HD constexpr void add_h(RowPat& r, int8_t a, int8_t b) { r.h[r.nh][0] = a; r.h[r.nh][1] = b; r.nh++; }
// full clique over (x0..x3): 4 diagonals then 6 pairs, in this fixed tag order:
// (0,0)(1,1)(2,2)(3,3)(0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
PS_HD constexpr void clique4(RowPat& r, int8_t x0, int8_t x1, int8_t x2, int8_t x3) {
  int8_t v[4] = {x0, x1, x2, x3};
  for (int q = 0; q < 4; q++) add_h(r, v[q], v[q]);
  for (int a = 0; a < 4; a++) for (int b = a + 1; b < 4; b++) add_h(r, v[a], v[b]);
}
}  // namespace detail

// Tag layout conventions used by the kernels:
//  polar rows over (th_a, th_b, V_a, V_b): J tags [.., AF, AT, BF, BT]; H = clique4(AF, AT, BF, BT)
//  rect rows over (vi, vr): J tags [.., AF, AT, BF, BT]; H = 4 diagonals (AF,AT,BF,BT) then the
//  cross pairs (BF,BT) (AF,AT) (BF,AT) (AF,BT)   [vr_f vr_t, vi_f vi_t, vr_f vi_t, vi_f vr_t]
PS_HD constexpr BranchPat branch_pattern(int form, int src) {
  using namespace detail;
  BranchPat p{};
  const bool mp = (src == SRC_MATPOWER);
  int n = 0;
  if (form == PBRAPV || form == PBRARV) {
    // Pij, Pji, Qij, Qji flow definitions (:302-305, :309-312 | :366-369, :373-376)
    const int8_t fl[4] = {R_F0, R_F2, R_F1, R_F3};
    for (int q = 0; q < 4; q++) {
      RowPat& r = p.rows[n++];
      r.eq = 1;
      set_j(r, fl[q], R_AF, R_AT, R_BF, R_BT);
      if (form == PBRAPV) clique4(r, R_AF, R_AT, R_BF, R_BT);
      else { add_h(r, R_AF, R_AF); add_h(r, R_AT, R_AT); add_h(r, R_BF, R_BF); add_h(r, R_BT, R_BT);
             add_h(r, R_BF, R_BT); add_h(r, R_AF, R_AT); add_h(r, R_BF, R_AT); add_h(r, R_AF, R_BT); }
    }
  }
  if (form == PBRAWV || form == PNRAWV) {                    // cone :316, :354 | :380, :416
    RowPat& r = p.rows[n++];
    r.eq = 1;
    set_j(r, R_WR, R_WI, R_WDF, R_WDT);
    add_h(r, R_WR, R_WR); add_h(r, R_WI, R_WI); add_h(r, R_WDF, R_WDF); add_h(r, R_WDT, R_WDT); add_h(r, R_WDF, R_WDT);
  }
  for (int side = 0; side < 2; side++) {                     // the two limit rows (from side, to side)
    if (form == PNRAWV) break;                                // its limits are @constraint rows (:357,:360)
    RowPat& r = p.rows[n++];
    r.eq = 0;
    const int8_t A = side ? R_AT : R_AF, B = side ? R_BT : R_BF, WD = side ? R_WDT : R_WDF;
    const int8_t FA = side ? R_F2 : R_F0, FB = side ? R_F3 : R_F1;
    if (is_PB(form)) {                                        // L1 :306-307,:313-314,:321-322 | :370-371,:377-378,:385-386
      if (mp) set_j(r, FA, FB);
      else if (form == PBRAPV) set_j(r, FA, FB, B);
      else if (form == PBRARV) set_j(r, FA, FB, B, A);
      else set_j(r, FA, FB, WD);
      add_h(r, FA, FA); add_h(r, FB, FB);
      if (!mp && form == PBRAPV) add_h(r, B, B);
      if (!mp && form == PBRARV) { add_h(r, B, B); add_h(r, A, A); }
    } else if (is_CB(form)) {                                 // L2/L3 :328-329,:335-336 | :392-393,:399-400
      if (!mp) { set_j(r, FA, FB); add_h(r, FA, FA); add_h(r, FB, FB); }
      else if (form == CBRARV) {
        set_j(r, FA, FB, B, A);                               // Ir, Ii, vr_a, vi_a
        add_h(r, FA, FA); add_h(r, FB, FB); add_h(r, B, B); add_h(r, A, A);
        add_h(r, FA, B); add_h(r, FA, A); add_h(r, FB, B); add_h(r, FB, A); add_h(r, B, A);
      } else {
        set_j(r, FA, FB, WD);
        add_h(r, FA, FA); add_h(r, FB, FB); add_h(r, WD, WD); add_h(r, FA, WD); add_h(r, FB, WD);
      }
    } else {                                                  // PN*: L4/L5 :337-352 | :401-414
      set_j(r, R_AF, R_AT, R_BF, R_BT);
      if (form == PNRARV && !mp) {
        add_h(r, R_AF, R_AF); add_h(r, R_AT, R_AT); add_h(r, R_BF, R_BF); add_h(r, R_BT, R_BT);
        add_h(r, R_BF, R_BT); add_h(r, R_AF, R_AT); add_h(r, R_BF, R_AT); add_h(r, R_AF, R_BT);
      } else clique4(r, R_AF, R_AT, R_BF, R_BT);
    }
  }
  p.nrows = n;
  return p;
}

constexpr int MAX_BR_JTAGS = 32;
constexpr int MAX_BR_HTAGS = 48;
constexpr int MAX_BR_ROWS = 6;

// ---------------------------------------------------------------------------------------------
// Slot positions of every tag inside its branch's block, as compile-time tables.
//
// A branch's rows touch the same roles for every branch; the only topology dependence of the
// JuMP column order (sorted by variable index) is whether f < t ("swap" = f > t),
// because the variable blocks are laid out in the fixed order of formulations.jl:20-57:
//   [theta|vi] [V|vr] [Wd] [Wr] [Wi] [pg] [qg] [F0] [F1] [F2] [F3].
// role_key() is an order-isomorphic image of the variable index of a role.
PS_HD constexpr int role_key(int role, int swap) {
  switch (role) {
    case R_AF:  return 0 + (swap ? 1 : 0);
    case R_AT:  return 0 + (swap ? 0 : 1);
    case R_BF:  return 2 + (swap ? 1 : 0);
    case R_BT:  return 2 + (swap ? 0 : 1);
    case R_WDF: return 4 + (swap ? 1 : 0);
    case R_WDT: return 4 + (swap ? 0 : 1);
    case R_WR:  return 6;
    case R_WI:  return 7;
    default:    return 8 + (role - R_F0);
  }
}

struct BranchSlots {
  uint8_t jpos[2][MAX_BR_JTAGS];   // [swap][tag] -> position inside the branch's J block
  uint8_t hpos[2][MAX_BR_HTAGS];   // [swap][tag] -> position inside the branch's H block
  uint8_t jinv[2][MAX_BR_JTAGS];   // [swap][position] -> tag (inverse of jpos)
  uint8_t hinv[2][MAX_BR_HTAGS];
  int nrows, nj, nh;
};

// Canonical layout: Jacobian columns ascending; Hessian block = vertices
// ascending (diagonal slots) then off-diagonal pairs (max,min) ascending.  Tags inside a row are
// unique (checked again at run time by the plan builder against the real variable indices).
PS_HD constexpr BranchSlots branch_slots(int form, int src) {
  BranchSlots s{};
  const BranchPat p = branch_pattern(form, src);
  s.nrows = p.nrows; s.nj = p.nj(); s.nh = p.nh();
  for (int swap = 0; swap < 2; swap++) {
    int jb = 0, hb = 0, jt = 0, ht = 0;
    for (int r = 0; r < p.nrows; r++) {
      const RowPat& rp = p.rows[r];
      for (int q = 0; q < rp.nj; q++) {
        int rank = 0;
        for (int o = 0; o < rp.nj; o++) if (role_key(rp.j[o], swap) < role_key(rp.j[q], swap)) rank++;
        s.jpos[swap][jt++] = (uint8_t)(jb + rank);
      }
      int ndiag = 0;
      for (int q = 0; q < rp.nh; q++) if (rp.h[q][0] == rp.h[q][1]) ndiag++;
      for (int q = 0; q < rp.nh; q++) {
        const int a = role_key(rp.h[q][0], swap), b = role_key(rp.h[q][1], swap);
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        int rank = 0;
        for (int o = 0; o < rp.nh; o++) {
          const int c = role_key(rp.h[o][0], swap), d = role_key(rp.h[o][1], swap);
          const int hi2 = c > d ? c : d, lo2 = c > d ? d : c;
          if ((hi == lo) != (hi2 == lo2)) continue;            // diagonals rank among diagonals
          if (hi2 < hi || (hi2 == hi && lo2 < lo)) rank++;
        }
        s.hpos[swap][ht++] = (uint8_t)(hb + (hi == lo ? rank : ndiag + rank));
      }
      jb += rp.nj;
      hb += rp.nh;
    }
    for (int t = 0; t < s.nj; t++) s.jinv[swap][s.jpos[swap][t]] = (uint8_t)t;
    for (int t = 0; t < s.nh; t++) s.hinv[swap][s.hpos[swap][t]] = (uint8_t)t;
  }
  return s;
}

// the direct (register -> 32-byte sector) store path needs whole sectors per branch or per pair of branches
PS_HD constexpr bool direct_ok(int form, int src) {
  const BranchSlots s = branch_slots(form, src);
  return s.nj > 0 && s.nh > 0 && s.nj % 2 == 0 && s.nh % 2 == 0;
}

// contributions of one half-edge to its bus's two balance rows (emission order of ps_plan.cpp / he_terms): Jacobian,
// Hessian, and the number of residual pieces staged for the bus thread
PS_HD constexpr int he_nj(int f) { return (f == PBRAPV || f == PBRARV) ? 2 : 8; }
PS_HD constexpr int he_nh(int f) { return (f == PBRAPV || f == PBRARV) ? 0 : (is_CB(f) ? 12 : (f == PNRARV ? 16 : 20)); }
PS_HD constexpr int he_np(int f) { return is_CB(f) ? 4 : (f == PNRARV ? 6 : (is_PN(f) ? 4 : 0)); }
// Hessian contributions of a half-edge that are structurally present but always 0.0 (he_math emits them with zero()):
// the warp-tile kernel never writes them (x + 0.0 == x; an all-zero slot is zeroed once)
PS_HD constexpr bool he_h_zero(int f, int ty) {
  return is_CB(f) ? (ty % 3 != 2) : (f == PNRARV ? ((ty & 7) == 2 || (ty & 7) == 3) : (is_PN(f) ? (ty % 10 == 3) : false));
}
// The residual of a nodal balance as a flat left-to-right sum of signed terms (formulations.jl:171-292): per half-edge
// he_npg terms -- the staged pieces above, or for the PB* models status * flow variable for the P and the Q row --
// term k belongs to row piece_row (0 = P, 1 = Q) and enters with the sign piece_neg (x - y == x + (-y) exactly)
PS_HD constexpr int he_npg(int f) { return (f == PBRAPV || f == PBRARV) ? 2 : he_np(f); }
PS_HD constexpr int piece_row(int f, int k) { return (f == PBRAPV || f == PBRARV) ? k : (f == PNRARV ? (k >= 3 ? 1 : 0) : (k >= 2 ? 1 : 0)); }
PS_HD constexpr bool piece_neg(int f, int k) { return is_PN(f) ? (k < 2) : (is_CB(f) ? (k == 3) : false); }

}  // namespace ps
