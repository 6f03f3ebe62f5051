/* TEST INFRASTRUCTURE ONLY -- CPU oracle #2 (plain C): restatement of the Powersense.jl NL
 * constraint block g(x), its sparse Jacobian and the Hessian of the Lagrangian.
 *
 * PARITY UNPINNED: the reference evaluates these rows with JuMP 0.21.10's NLPEvaluator
 * (Manifest.toml:187-191; not under /root/reference, no Julia in this image) and its tests pin no
 * g/J/H value or structure.  This file is pinned instead against
 *   - oracle/jump_oracle.py (generic expression-tree restatement of JuMP's published algorithm),
 *   - finite differences, and the 14-bus known answer 146.7 (test/opf/ACOPF_formulations.jl:9).
 *
 * Method: every textual term of src/opf/formulations.jl:170-428 is written here as the same
 * expression on small dense second-order jets (value, gradient, Hessian over the <= 6 variables
 * the term touches) -- i.e. automatic differentiation, like the reference, not hand-derived
 * formulas -- and is emitted into the row in the reference's order (gens, Sij ascending, Sji
 * ascending, shunt terms, then "- Pd": SURVEY.md C.1).  Structure follows SURVEY.md C.2/C.3:
 * Jacobian columns per row sorted ascending and unique; Hessian block per row = vertices
 * ascending as diagonal slots, then off-diagonal pairs (max,min) ascending (canonical order);
 * which pairs are structural is declared per term with the clique rule of JuMP's
 * compute_hessian_sparsity.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (powersense.jl_b200/csrc) never links or calls it. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { PBRAPV = 0, PBRARV, PBRAWV, CBRARV, CBRAWV, PNPAPV, PNRAPV, PNRARV, PNRAWV };
enum { SRC_MATPOWER = 0, SRC_ARPAE = 1 };
enum { FLAG_FACTS = 1, FLAG_OBJ_LINEAR = 2 };
enum { EV_G = 1, EV_J = 2, EV_H = 4, EV_ABS = 8 };   /* EV_ABS: accumulate |contribution| per slot (condition scale) */

typedef struct {
  int64_t nbus, nbr, ngen, nbss;
  const int64_t *br_f, *br_t;                 /* 1-based (data.br, PowersenseData.jl:210) */
  const double *Gf, *Bf, *Gt, *Bt, *gf, *bf, *gt, *bt, *Imax; /* PowersenseData.jl:198-207 */
  const double *Pd, *Qd, *Gl, *Bl;            /* PowersenseData.jl:186-189 */
  const int64_t *gen_ptr, *gen_idx, *sij_ptr, *sij_idx, *sji_ptr, *sji_idx, *sh_ptr, *sh_idx;
  const int64_t *lps;                         /* [ngen] points per generator (tgh count) */
} pso_net;

/* ------------------------------------------------------------------------------------ jets */
#define NV 6
typedef struct { double v, g[NV], h[NV][NV]; } jet;

static inline jet j_const(double c) { jet r; memset(&r, 0, sizeof r); r.v = c; return r; }
static inline jet j_var(double x, int i) { jet r = j_const(x); r.g[i] = 1.0; return r; }
static inline jet j_lin(jet a, double sa, jet b, double sb, int n) { /* sa*a + sb*b */
  jet r; memset(&r, 0, sizeof r);
  r.v = sa * a.v + sb * b.v;
  for (int i = 0; i < n; i++) { r.g[i] = sa * a.g[i] + sb * b.g[i];
    for (int j = 0; j < n; j++) r.h[i][j] = sa * a.h[i][j] + sb * b.h[i][j]; }
  return r;
}
static inline jet j_add(jet a, jet b, int n) { return j_lin(a, 1.0, b, 1.0, n); }
static inline jet j_sub(jet a, jet b, int n) { return j_lin(a, 1.0, b, -1.0, n); }
static inline jet j_scale(double c, jet a, int n) { /* c * a (constant first, as in the source) */
  jet r; memset(&r, 0, sizeof r);
  r.v = c * a.v;
  for (int i = 0; i < n; i++) { r.g[i] = c * a.g[i]; for (int j = 0; j < n; j++) r.h[i][j] = c * a.h[i][j]; }
  return r;
}
static inline jet j_mul(jet a, jet b, int n) {
  jet r; memset(&r, 0, sizeof r);
  r.v = a.v * b.v;
  for (int i = 0; i < n; i++) { r.g[i] = a.v * b.g[i] + b.v * a.g[i];
    for (int j = 0; j < n; j++)
      r.h[i][j] = a.v * b.h[i][j] + b.v * a.h[i][j] + a.g[i] * b.g[j] + a.g[j] * b.g[i]; }
  return r;
}
static inline jet j_fun(jet a, double f, double d1, double d2, int n) {
  jet r; memset(&r, 0, sizeof r);
  r.v = f;
  for (int i = 0; i < n; i++) { r.g[i] = d1 * a.g[i];
    for (int j = 0; j < n; j++) r.h[i][j] = d1 * a.h[i][j] + d2 * a.g[i] * a.g[j]; }
  return r;
}
static inline jet j_sq(jet a, int n) { return j_fun(a, a.v * a.v, 2.0 * a.v, 2.0, n); }  /* x^2 -> x*x (C.1) */
static inline jet j_pow(jet a, double p, int n) {
  return j_fun(a, pow(a.v, p), p * pow(a.v, p - 1.0), p * (p - 1.0) * pow(a.v, p - 2.0), n);
}
static inline jet j_cos(jet a, int n) { return j_fun(a, cos(a.v), -sin(a.v), -cos(a.v), n); }
static inline jet j_sin(jet a, int n) { return j_fun(a, sin(a.v), cos(a.v), -sin(a.v), n); }

/* ---------------------------------------------------------------------------------- model */
typedef struct {
  int form, src, flags;
  pso_net net;                      /* deep copies */
  int64_t n_var, m_nl, nnzJ, nnzH;
  int64_t o_a, o_b, o_Wd, o_Wr, o_Wi, o_pg, o_qg, o_f[4], o_cg, o_bss, o_tgh; /* 0-based offsets */
  int64_t *jrow, *jcol, *hrow, *hcol, *jptr, *hptr;
  double *lo, *hi;
  int32_t *jslot, *hslot;           /* contribution -> slot maps (emission order) */
  int64_t ncj, nch;
} pso_model;

/* emission context: STRUCT mode records keys, EVAL mode accumulates values */
typedef struct {
  const pso_model *m;
  int mode;                         /* 0 = struct, 1 = eval */
  int what;
  const double *x, *Pd, *Qd, *mu;
  double *g, *J, *H;
  int64_t row, cj, ch;
  double mu_row;
  /* struct mode buffers */
  int64_t *cjrow, *cjvar, capj;
  int64_t *chrow, *cha, *chb, caph;
  int bad;                          /* analytic non-zero outside the declared structure */
} ctx;

static void push_j(ctx *c, int64_t var) {
  if (c->cj == c->capj) { c->capj = c->capj ? 2 * c->capj : 1024;
    c->cjrow = realloc(c->cjrow, c->capj * 8); c->cjvar = realloc(c->cjvar, c->capj * 8); }
  c->cjrow[c->cj] = c->row; c->cjvar[c->cj] = var;
}
static void push_h(ctx *c, int64_t a, int64_t b) {
  if (c->ch == c->caph) { c->caph = c->caph ? 2 * c->caph : 1024;
    c->chrow = realloc(c->chrow, c->caph * 8); c->cha = realloc(c->cha, c->caph * 8); c->chb = realloc(c->chb, c->caph * 8); }
  c->chrow[c->ch] = c->row; c->cha[c->ch] = a > b ? a : b; c->chb[c->ch] = a > b ? b : a;
}
static inline void emit_j(ctx *c, int64_t var, double v) {
  if (c->mode == 0) push_j(c, var);
  else if (c->what & EV_J) c->J[c->m->jslot[c->cj]] += (c->what & EV_ABS) ? fabs(v) : v;
  c->cj++;
}
static inline void emit_h(ctx *c, int64_t a, int64_t b, double v) {
  if (c->mode == 0) push_h(c, a, b);
  else if (c->what & EV_H) c->H[c->m->hslot[c->ch]] += (c->what & EV_ABS) ? fabs(c->mu_row * v) : c->mu_row * v;
  c->ch++;
}
/* emit sign * term: gradient for every local variable, Hessian for the declared pairs */
static void emit_term(ctx *c, const jet *t, double sign, int n, const int64_t *gid,
                      const int8_t (*pairs)[2], int np) {
  for (int l = 0; l < n; l++) emit_j(c, gid[l], sign * t->g[l]);
  for (int p = 0; p < np; p++) emit_h(c, gid[pairs[p][0]], gid[pairs[p][1]], sign * t->h[pairs[p][0]][pairs[p][1]]);
  if (c->mode == 1) { /* every analytic second derivative must be covered by a declared pair */
    for (int a = 0; a < n; a++) for (int b = 0; b <= a; b++) if (t->h[a][b] != 0.0) {
      int ok = 0;
      for (int p = 0; p < np; p++)
        if ((pairs[p][0] == a && pairs[p][1] == b) || (pairs[p][0] == b && pairs[p][1] == a)) ok = 1;
      if (!ok) c->bad = 1;
    }
  }
}

/* structural pair tables (local indices), from JuMP's clique rule */
static const int8_t P_CLIQUE4[10][2] = {{0,0},{1,1},{2,2},{3,3},{0,1},{0,2},{0,3},{1,2},{1,3},{2,3}};
static const int8_t P_D2[1][2] = {{2,2}};                       /* c * V_a^2 with locals [th_a, th_b, V_a, V_b] */
static const int8_t P_D3[1][2] = {{3,3}};
static const int8_t P_A01[2][2] = {{0,0},{1,1}};                /* c * (r_a^2 + i_a^2), locals [r_a, i_a, ...] */
static const int8_t P_B23[2][2] = {{2,2},{3,3}};
static const int8_t P_C[6][2] = {{0,0},{2,2},{0,2},{1,1},{3,3},{1,3}};   /* c*(r_a r_b + i_a i_b) */
static const int8_t P_S[6][2] = {{0,0},{3,3},{0,3},{1,1},{2,2},{1,2}};   /* c*(r_a i_b - i_a r_b) */
static const int8_t P_VAVB[3][2] = {{2,2},{3,3},{2,3}};
static const int8_t P_CL2[3][2] = {{0,0},{1,1},{0,1}};
static const int8_t P_CL3[6][2] = {{0,0},{1,1},{2,2},{0,1},{0,2},{1,2}};
static const int8_t P_D0[1][2] = {{0,0}};
static const int8_t P_T4P[6][2] = {{0,0},{2,2},{0,2},{1,1},{3,3},{1,3}}; /* r_a*Ir + i_a*Ii, locals [r_a,i_a,Ir,Ii] */
static const int8_t P_T4Q[6][2] = {{1,1},{2,2},{1,2},{0,0},{3,3},{0,3}}; /* i_a*Ir - r_a*Ii */
static const int8_t P_L2[9][2] = {{0,0},{2,2},{3,3},{0,2},{0,3},{2,3},{1,1},{1,2},{1,3}}; /* [Ir,Ii,r,i] */
static const int8_t P_L3[5][2] = {{0,0},{2,2},{0,2},{1,1},{1,2}};                     /* [Ir,Ii,Wd] */
static const int8_t P_L6[5][2] = {{0,0},{1,1},{2,2},{3,3},{2,3}};                     /* [Wr,Wi,Wd_f,Wd_t] */

static inline int is_polar(int f) { return f == PBRAPV || f == PNPAPV || f == PNRAPV; }
static inline int has_W(int f) { return f == PBRAWV || f == CBRAWV || f == PNRAWV; }
static inline int is_PB(int f) { return f == PBRAPV || f == PBRARV || f == PBRAWV; }
static inline int is_CB(int f) { return f == CBRARV || f == CBRAWV; }

static void row_begin(ctx *c) { c->mu_row = (c->mode == 1 && c->mu) ? c->mu[c->row] : 1.0; }
static void row_end(ctx *c, double gval) {
  if (c->mode == 1 && (c->what & EV_G)) c->g[c->row] = gval;
  c->row++;
}

/* the whole NL block in reference order (formulations.jl:170-428) */
static void walk(ctx *c) {
  const pso_model *m = c->m; const pso_net *N = &m->net; const int F = m->form;
  const double *x = c->x; const int ev = c->mode == 1;
  const int facts = (m->flags & FLAG_FACTS) != 0;
#define XV(off, i) (ev ? x[(off) + (i)] : 1.0)
  c->row = 0; c->cj = 0; c->ch = 0;
  if (F != PBRAWV && F != PNRAWV) {
    for (int64_t i = 0; i < N->nbus; i++) {                       /* :170 */
      for (int pq = 0; pq < 2; pq++) {                            /* row P_i then row Q_i (:291-292) */
        row_begin(c);
        double acc = 0.0;                                         /* :171-172 */
        for (int64_t q = N->gen_ptr[i]; q < N->gen_ptr[i + 1]; q++) {   /* :177-180 */
          int64_t v = (pq ? m->o_qg : m->o_pg) + N->gen_idx[q] - 1;
          acc = acc + XV(0, v); emit_j(c, v, 1.0);
        }
        for (int side = 0; side < 2; side++) {
          const int64_t *ptr = side ? N->sji_ptr : N->sij_ptr, *idx = side ? N->sji_idx : N->sij_idx;
          for (int64_t q = ptr[i]; q < ptr[i + 1]; q++) {
            int64_t k = idx[q] - 1, f = N->br_f[k] - 1, t = N->br_t[k] - 1;
            int64_t a = side ? t : f, b = side ? f : t;
            double gs = side ? N->gt[k] : N->gf[k], bs = side ? N->bt[k] : N->bf[k];
            double G = side ? N->Gt[k] : N->Gf[k], B = side ? N->Bt[k] : N->Bf[k];
            if (F == PBRAPV || F == PBRARV) {                     /* :188-196 */
              int64_t v = m->o_f[(pq ? 1 : 0) + 2 * side] + k;
              acc = acc + XV(0, v); emit_j(c, v, 1.0);
            } else if (is_CB(F)) {                                /* :206-214, locals [vr_a, vi_a, Ir, Ii] */
              int64_t gid[4] = {m->o_b + a, m->o_a + a, m->o_f[0 + 2 * side] + k, m->o_f[1 + 2 * side] + k};
              jet ra = j_var(XV(0, gid[0]), 0), ia = j_var(XV(0, gid[1]), 1);
              jet Ir = j_var(XV(0, gid[2]), 2), Ii = j_var(XV(0, gid[3]), 3);
              if (!pq) { jet t1 = j_mul(ra, Ir, 4), t2 = j_mul(ia, Ii, 4); jet tt = j_add(t1, t2, 4);
                acc = acc + t1.v + t2.v; emit_term(c, &tt, 1.0, 4, gid, P_T4P, 6);
              } else { jet t1 = j_mul(ia, Ir, 4), t2 = j_mul(ra, Ii, 4); jet tt = j_sub(t1, t2, 4);
                acc = acc + t1.v - t2.v; emit_term(c, &tt, 1.0, 4, gid, P_T4Q, 6); }
            } else if (F == PNPAPV || F == PNRAPV) {              /* :215-234, locals [th_a, th_b, V_a, V_b] */
              int64_t gid[4] = {m->o_a + a, m->o_a + b, m->o_b + a, m->o_b + b};
              jet tha = j_var(XV(0, gid[0]), 0), thb = j_var(XV(0, gid[1]), 1);
              jet Va = j_var(XV(0, gid[2]), 2), Vb = j_var(XV(0, gid[3]), 3);
              jet Vf = side ? Vb : Va, Vt = side ? Va : Vb;       /* source writes V[f]*V[t] on both sides */
              jet p1 = j_scale(pq ? bs : gs, j_sq(Va, 4), 4), p2;
              if (F == PNPAPV) {
                double Y = hypot(G, B), thY = atan2(B, G);        /* :217, :222 abs/angle */
                jet ang = j_sub(j_sub(tha, thb, 4), j_const(thY), 4);
                jet tr = pq ? j_sin(ang, 4) : j_cos(ang, 4);
                p2 = j_mul(j_mul(j_scale(Y, tr, 4), Vf, 4), Vt, 4);
                acc = pq ? (acc + p1.v) - p2.v : (acc - p1.v) - p2.v;
                emit_term(c, &p1, pq ? 1.0 : -1.0, 4, gid, P_D2, 1);
                emit_term(c, &p2, -1.0, 4, gid, P_CLIQUE4, 10);
              } else {
                jet dl = j_sub(tha, thb, 4), cs = j_cos(dl, 4), sn = j_sin(dl, 4);
                jet co = pq ? j_sub(j_scale(B, cs, 4), j_scale(G, sn, 4), 4) : j_add(j_scale(G, cs, 4), j_scale(B, sn, 4), 4);
                p2 = j_mul(j_mul(co, Vf, 4), Vt, 4);
                acc = pq ? (acc + p1.v) + p2.v : (acc - p1.v) - p2.v;
                emit_term(c, &p1, pq ? 1.0 : -1.0, 4, gid, P_D2, 1);
                emit_term(c, &p2, pq ? 1.0 : -1.0, 4, gid, P_CLIQUE4, 10);
              }
            } else if (F == PNRARV) {                             /* :235-243, locals [vr_a, vi_a, vr_b, vi_b] */
              int64_t gid[4] = {m->o_b + a, m->o_a + a, m->o_b + b, m->o_a + b};
              jet ra = j_var(XV(0, gid[0]), 0), ia = j_var(XV(0, gid[1]), 1);
              jet rb = j_var(XV(0, gid[2]), 2), ib = j_var(XV(0, gid[3]), 3);
              jet A = j_add(j_sq(ra, 4), j_sq(ia, 4), 4);
              jet C = j_add(j_mul(ra, rb, 4), j_mul(ia, ib, 4), 4);
              jet S = j_sub(j_mul(ra, ib, 4), j_mul(ia, rb, 4), 4);
              if (!pq) { jet p1 = j_scale(gs, A, 4), p2 = j_scale(G, C, 4), p3 = j_scale(B, S, 4);
                acc = ((acc - p1.v) - p2.v) + p3.v;
                emit_term(c, &p1, -1.0, 4, gid, P_A01, 2); emit_term(c, &p2, -1.0, 4, gid, P_C, 6); emit_term(c, &p3, 1.0, 4, gid, P_S, 6);
              } else { jet p1 = j_scale(bs, A, 4), p2 = j_scale(B, C, 4), p3 = j_scale(G, S, 4);
                acc = acc + p1.v + p2.v + p3.v;
                emit_term(c, &p1, 1.0, 4, gid, P_A01, 2); emit_term(c, &p2, 1.0, 4, gid, P_C, 6); emit_term(c, &p3, 1.0, 4, gid, P_S, 6); }
            }
          }
        }
        /* shunt terms :256-279 */
        double cl = pq ? N->Bl[i] : N->Gl[i], sg = pq ? 1.0 : -1.0;
        if (is_polar(F)) {
          int64_t gid[2] = {m->o_b + i, 0};
          jet Vi = j_var(XV(0, gid[0]), 0); jet p = j_scale(cl, j_sq(Vi, 1), 1);
          acc = acc + sg * p.v; emit_term(c, &p, sg, 1, gid, P_D0, 1);
          if (pq && facts) for (int64_t q = N->sh_ptr[i]; q < N->sh_ptr[i + 1]; q++) {
            int64_t g2[2] = {m->o_bss + N->sh_idx[q] - 1, m->o_b + i};
            jet bs_ = j_var(XV(0, g2[0]), 0), V2 = j_var(XV(0, g2[1]), 1);
            jet pp = j_mul(bs_, j_sq(V2, 2), 2);
            acc = acc + pp.v; emit_term(c, &pp, 1.0, 2, g2, P_CL2, 3);
          }
        } else if (F == PBRARV || F == CBRARV || F == PNRARV) {
          int64_t gid[2] = {m->o_b + i, m->o_a + i};              /* [vr_i, vi_i] */
          jet r = j_var(XV(0, gid[0]), 0), im = j_var(XV(0, gid[1]), 1);
          jet p = j_scale(cl, j_add(j_sq(r, 2), j_sq(im, 2), 2), 2);
          acc = acc + sg * p.v; emit_term(c, &p, sg, 2, gid, P_A01, 2);
          if (pq && facts) for (int64_t q = N->sh_ptr[i]; q < N->sh_ptr[i + 1]; q++) {
            int64_t g3[3] = {m->o_bss + N->sh_idx[q] - 1, m->o_b + i, m->o_a + i};
            jet bs_ = j_var(XV(0, g3[0]), 0), r3 = j_var(XV(0, g3[1]), 1), i3 = j_var(XV(0, g3[2]), 2);
            jet pp = j_mul(bs_, j_add(j_sq(r3, 3), j_sq(i3, 3), 3), 3);
            acc = acc + pp.v; emit_term(c, &pp, 1.0, 3, g3, P_CL3, 6);
          }
        } else if (F == CBRAWV) {
          int64_t v = m->o_Wd + i;
          acc = acc + sg * (cl * XV(0, v)); emit_j(c, v, sg * cl);
          if (pq && facts) for (int64_t q = N->sh_ptr[i]; q < N->sh_ptr[i + 1]; q++) {
            int64_t g2[2] = {m->o_bss + N->sh_idx[q] - 1, v};
            jet bs_ = j_var(XV(0, g2[0]), 0), W = j_var(XV(0, g2[1]), 1);
            jet pp = j_mul(bs_, W, 2);
            acc = acc + pp.v; emit_term(c, &pp, 1.0, 2, g2, P_CL2, 3);
          }
        }
        double load = ev ? (pq ? c->Qd[i] : c->Pd[i]) : 0.0;
        row_end(c, acc - load);                                   /* :291-292: lhs - rhs */
      }
    }
  }
  const int mp = m->src == SRC_MATPOWER;                          /* :299 */
  for (int64_t k = 0; k < N->nbr; k++) {                          /* :300 / :364 */
    int64_t f = N->br_f[k] - 1, t = N->br_t[k] - 1;
    double Im2 = N->Imax[k] * N->Imax[k];
    if (F == PBRAPV || F == PBRARV) {
      /* four flow definitions: Pij, Pji, Qij, Qji (:302-305, :309-312) */
      for (int r4 = 0; r4 < 4; r4++) {
        int side = r4 & 1, pq = r4 >> 1;
        int64_t a = side ? t : f, b = side ? f : t;
        double gs = side ? N->gt[k] : N->gf[k], bs = side ? N->bt[k] : N->bf[k];
        double G = side ? N->Gt[k] : N->Gf[k], B = side ? N->Bt[k] : N->Bf[k];
        int64_t fv = m->o_f[pq + 2 * side] + k;
        row_begin(c);
        emit_j(c, fv, 1.0);
        double rhs;
        if (F == PBRAPV) {
          int64_t gid[4] = {m->o_a + a, m->o_a + b, m->o_b + a, m->o_b + b};
          jet tha = j_var(XV(0, gid[0]), 0), thb = j_var(XV(0, gid[1]), 1);
          jet Va = j_var(XV(0, gid[2]), 2), Vb = j_var(XV(0, gid[3]), 3);
          jet Vf = side ? Vb : Va, Vt = side ? Va : Vb;
          jet dl = j_sub(tha, thb, 4), cs = j_cos(dl, 4), sn = j_sin(dl, 4);
          jet p1 = j_scale(pq ? bs : -gs, j_sq(Va, 4), 4);
          jet co = pq ? j_sub(j_scale(B, cs, 4), j_scale(G, sn, 4), 4) : j_add(j_scale(G, cs, 4), j_scale(B, sn, 4), 4);
          jet p2 = j_mul(j_mul(co, Vf, 4), Vt, 4);
          rhs = pq ? p1.v + p2.v : p1.v - p2.v;
          emit_term(c, &p1, -1.0, 4, gid, P_D2, 1);
          emit_term(c, &p2, pq ? -1.0 : 1.0, 4, gid, P_CLIQUE4, 10);
        } else {
          int64_t gid[4] = {m->o_b + a, m->o_a + a, m->o_b + b, m->o_a + b};
          jet ra = j_var(XV(0, gid[0]), 0), ia = j_var(XV(0, gid[1]), 1);
          jet rb = j_var(XV(0, gid[2]), 2), ib = j_var(XV(0, gid[3]), 3);
          jet A = j_add(j_sq(ra, 4), j_sq(ia, 4), 4);
          jet C = j_add(j_mul(ra, rb, 4), j_mul(ia, ib, 4), 4);
          jet S = j_sub(j_mul(ra, ib, 4), j_mul(ia, rb, 4), 4);
          jet p1 = j_scale(pq ? bs : -gs, A, 4), p2 = j_scale(pq ? B : G, C, 4), p3 = j_scale(pq ? G : B, S, 4);
          rhs = pq ? p1.v + p2.v + p3.v : (p1.v - p2.v) + p3.v;
          emit_term(c, &p1, -1.0, 4, gid, P_A01, 2);
          emit_term(c, &p2, pq ? -1.0 : 1.0, 4, gid, P_C, 6);
          emit_term(c, &p3, -1.0, 4, gid, P_S, 6);
        }
        row_end(c, XV(0, fv) - rhs);
      }
    }
    if (F == PBRAWV || F == PNRAWV) {                             /* cone :316, :354 */
      int64_t gid[4] = {m->o_Wr + k, m->o_Wi + k, m->o_Wd + f, m->o_Wd + t};
      jet Wr = j_var(XV(0, gid[0]), 0), Wi = j_var(XV(0, gid[1]), 1), Wf = j_var(XV(0, gid[2]), 2), Wt = j_var(XV(0, gid[3]), 3);
      jet tt = j_sub(j_add(j_mul(Wr, Wr, 4), j_mul(Wi, Wi, 4), 4), j_mul(Wf, Wt, 4), 4);
      row_begin(c); emit_term(c, &tt, 1.0, 4, gid, P_L6, 5); row_end(c, tt.v);
    }
    if (is_PB(F)) {                                               /* L1 :306-307,:313-314,:321-322 / :370-371,:377-378,:385-386 */
      for (int side = 0; side < 2; side++) {
        int64_t a = side ? t : f;
        int64_t gid[2] = {m->o_f[0 + 2 * side] + k, m->o_f[1 + 2 * side] + k};
        jet P = j_var(XV(0, gid[0]), 0), Q = j_var(XV(0, gid[1]), 1);
        jet lhs = j_add(j_sq(P, 2), j_sq(Q, 2), 2);
        row_begin(c); emit_term(c, &lhs, 1.0, 2, gid, P_CL2, 2);
        double rhs = Im2;
        if (!mp) {
          if (F == PBRAPV) { int64_t g1[1] = {m->o_b + a}; jet Va = j_var(XV(0, g1[0]), 0);
            jet r = j_scale(Im2, j_sq(Va, 1), 1); rhs = r.v; emit_term(c, &r, -1.0, 1, g1, P_D0, 1);
          } else if (F == PBRARV) { int64_t g2[2] = {m->o_b + a, m->o_a + a};
            jet ra = j_var(XV(0, g2[0]), 0), ia = j_var(XV(0, g2[1]), 1);
            jet r = j_scale(Im2, j_add(j_sq(ra, 2), j_sq(ia, 2), 2), 2); rhs = r.v; emit_term(c, &r, -1.0, 2, g2, P_A01, 2);
          } else { int64_t v = m->o_Wd + a; rhs = Im2 * XV(0, v); emit_j(c, v, -Im2); }
        }
        row_end(c, lhs.v - rhs);
      }
    } else if (is_CB(F)) {                                        /* L2/L3 :328-329,:335-336 / :392-393,:399-400 */
      for (int side = 0; side < 2; side++) {
        int64_t a = side ? t : f;
        row_begin(c);
        double val;
        if (!mp) { int64_t gid[2] = {m->o_f[0 + 2 * side] + k, m->o_f[1 + 2 * side] + k};
          jet Ir = j_var(XV(0, gid[0]), 0), Ii = j_var(XV(0, gid[1]), 1);
          jet lhs = j_add(j_sq(Ir, 2), j_sq(Ii, 2), 2); val = lhs.v; emit_term(c, &lhs, 1.0, 2, gid, P_CL2, 2);
        } else if (F == CBRARV) { int64_t gid[4] = {m->o_f[0 + 2 * side] + k, m->o_f[1 + 2 * side] + k, m->o_b + a, m->o_a + a};
          jet Ir = j_var(XV(0, gid[0]), 0), Ii = j_var(XV(0, gid[1]), 1), r = j_var(XV(0, gid[2]), 2), im = j_var(XV(0, gid[3]), 3);
          jet A = j_add(j_sq(r, 4), j_sq(im, 4), 4);
          jet lhs = j_add(j_mul(j_sq(Ir, 4), A, 4), j_mul(j_sq(Ii, 4), A, 4), 4); val = lhs.v; emit_term(c, &lhs, 1.0, 4, gid, P_L2, 9);
        } else { int64_t gid[3] = {m->o_f[0 + 2 * side] + k, m->o_f[1 + 2 * side] + k, m->o_Wd + a};
          jet Ir = j_var(XV(0, gid[0]), 0), Ii = j_var(XV(0, gid[1]), 1), W = j_var(XV(0, gid[2]), 2);
          jet lhs = j_add(j_mul(j_sq(Ir, 3), W, 3), j_mul(j_sq(Ii, 3), W, 3), 3); val = lhs.v; emit_term(c, &lhs, 1.0, 3, gid, P_L3, 5); }
        row_end(c, val - Im2);
      }
    } else if (F == PNPAPV || F == PNRAPV) {                      /* L4 :337-343 / :401-407 */
      for (int side = 0; side < 2; side++) {
        int64_t a = side ? t : f, b = side ? f : t;
        double gs = side ? N->gt[k] : N->gf[k], bs = side ? N->bt[k] : N->bf[k];
        double G = side ? N->Gt[k] : N->Gf[k], B = side ? N->Bt[k] : N->Bf[k];
        double y1 = gs * gs + bs * bs, y2 = G * G + B * B, y3 = 2 * hypot(gs, bs) * hypot(G, B);
        double th1 = atan2(bs, gs), th2 = atan2(B, G);
        int64_t gid[4] = {m->o_a + a, m->o_a + b, m->o_b + a, m->o_b + b};
        jet tha = j_var(XV(0, gid[0]), 0), thb = j_var(XV(0, gid[1]), 1), Va = j_var(XV(0, gid[2]), 2), Vb = j_var(XV(0, gid[3]), 3);
        jet ang = j_sub(j_add(j_sub(tha, thb, 4), j_const(th1), 4), j_const(th2), 4), cs = j_cos(ang, 4);
        row_begin(c);
        double val;
        if (mp) {
          jet p1 = j_scale(y1, j_pow(Va, 4.0, 4), 4);
          jet p2 = j_mul(j_scale(y2, j_sq(Vb, 4), 4), j_sq(Va, 4), 4);
          jet p3 = j_mul(j_mul(j_scale(y3, j_pow(Va, 3.0, 4), 4), Vb, 4), cs, 4);
          val = p1.v + p2.v + p3.v;
          emit_term(c, &p1, 1.0, 4, gid, P_D2, 1); emit_term(c, &p2, 1.0, 4, gid, P_VAVB, 3); emit_term(c, &p3, 1.0, 4, gid, P_CLIQUE4, 10);
        } else {
          jet p1 = j_scale(y1, j_sq(Va, 4), 4), p2 = j_scale(y2, j_sq(Vb, 4), 4);
          jet p3 = j_mul(j_mul(j_scale(y3, Va, 4), Vb, 4), cs, 4);
          val = p1.v + p2.v + p3.v;
          emit_term(c, &p1, 1.0, 4, gid, P_D2, 1); emit_term(c, &p2, 1.0, 4, gid, P_D3, 1); emit_term(c, &p3, 1.0, 4, gid, P_CLIQUE4, 10);
        }
        row_end(c, val - Im2);
      }
    } else if (F == PNRARV) {                                     /* L5 :344-352 / :408-414 */
      for (int side = 0; side < 2; side++) {
        int64_t a = side ? t : f, b = side ? f : t;
        double gs = side ? N->gt[k] : N->gf[k], bs = side ? N->bt[k] : N->bf[k];
        double G = side ? N->Gt[k] : N->Gf[k], B = side ? N->Bt[k] : N->Bf[k];
        double y = gs * gs + bs * bs, Y = G * G + B * B, y1 = 2 * (gs * G + bs * B), y2 = 2 * (bs * G - gs * B);
        int64_t gid[4] = {m->o_b + a, m->o_a + a, m->o_b + b, m->o_a + b};
        jet ra = j_var(XV(0, gid[0]), 0), ia = j_var(XV(0, gid[1]), 1), rb = j_var(XV(0, gid[2]), 2), ib = j_var(XV(0, gid[3]), 3);
        jet A = j_add(j_sq(ra, 4), j_sq(ia, 4), 4), Bq = j_add(j_sq(rb, 4), j_sq(ib, 4), 4);
        jet C = j_add(j_mul(ra, rb, 4), j_mul(ia, ib, 4), 4), S = j_sub(j_mul(ra, ib, 4), j_mul(ia, rb, 4), 4);
        row_begin(c);
        double val;
        if (mp) {
          jet p1 = j_mul(j_scale(y1, C, 4), A, 4), p2 = j_mul(j_scale(y2, S, 4), A, 4);
          jet p3 = j_mul(j_scale(y, A, 4), A, 4), p4 = j_mul(j_scale(Y, Bq, 4), A, 4);
          val = p1.v + p2.v + p3.v + p4.v;
          /* one n-ary sum in the source; emitted term by term (same totals, same clique) so that the
           * EV_ABS pass sees the cancellation between the four products */
          emit_term(c, &p1, 1.0, 4, gid, P_CLIQUE4, 10); emit_term(c, &p2, 1.0, 4, gid, P_CLIQUE4, 10);
          emit_term(c, &p3, 1.0, 4, gid, P_CLIQUE4, 10); emit_term(c, &p4, 1.0, 4, gid, P_CLIQUE4, 10);
        } else {
          jet p1 = j_scale(y1, C, 4), p2 = j_scale(y2, S, 4), p3 = j_scale(y, A, 4), p4 = j_scale(Y, Bq, 4);
          val = p1.v + p2.v + p3.v + p4.v;
          emit_term(c, &p1, 1.0, 4, gid, P_C, 6); emit_term(c, &p2, 1.0, 4, gid, P_S, 6);
          emit_term(c, &p3, 1.0, 4, gid, P_A01, 2); emit_term(c, &p4, 1.0, 4, gid, P_B23, 2);
        }
        row_end(c, val - Im2);
      }
    }
  }
#undef XV
}

/* ------------------------------------------------------------------------ structure pass */
typedef struct { int64_t row, a, b, id; } key3;
static int cmp3(const void *p, const void *q) {
  const key3 *x = p, *y = q;
  if (x->row != y->row) return x->row < y->row ? -1 : 1;
  if (x->a != y->a) return x->a < y->a ? -1 : 1;
  if (x->b != y->b) return x->b < y->b ? -1 : 1;
  return x->id < y->id ? -1 : (x->id > y->id);
}

static void *dup_arr(const void *p, size_t n) { void *q = malloc(n ? n : 1); if (n) memcpy(q, p, n); return q; }

pso_model *pso_create(const pso_net *net, int form, int src, int flags) {
  pso_model *m = calloc(1, sizeof *m);
  m->form = form; m->src = src; m->flags = flags;
  pso_net *N = &m->net; *N = *net;
  int64_t nb = net->nbus, nr = net->nbr, ng = net->ngen;
#define DUPD(fld, n) N->fld = dup_arr(net->fld, (size_t)(n) * 8)
  DUPD(br_f, nr); DUPD(br_t, nr);
  DUPD(Gf, nr); DUPD(Bf, nr); DUPD(Gt, nr); DUPD(Bt, nr); DUPD(gf, nr); DUPD(bf, nr); DUPD(gt, nr); DUPD(bt, nr); DUPD(Imax, nr);
  DUPD(Pd, nb); DUPD(Qd, nb); DUPD(Gl, nb); DUPD(Bl, nb);
  DUPD(gen_ptr, nb + 1); DUPD(gen_idx, net->gen_ptr[nb]); DUPD(sij_ptr, nb + 1); DUPD(sij_idx, net->sij_ptr[nb]);
  DUPD(sji_ptr, nb + 1); DUPD(sji_idx, net->sji_ptr[nb]); DUPD(sh_ptr, nb + 1); DUPD(sh_idx, net->sh_ptr[nb]);
  DUPD(lps, ng);
#undef DUPD
  for (int64_t k = 0; k < nr; k++) if (net->br_f[k] == net->br_t[k]) { free(m); return NULL; }
  /* variable order: formulations.jl:20-67, :121 */
  int64_t o = 0;
  m->o_a = o; o += nb; m->o_b = o; o += nb;                      /* (theta,V) or (vi,vr) */
  m->o_Wd = m->o_Wr = m->o_Wi = -1;
  if (has_W(form)) { m->o_Wd = o; o += nb; }
  if (form == PBRAWV || form == PNRAWV) { m->o_Wr = o; o += nr; m->o_Wi = o; o += nr; }
  m->o_pg = o; o += ng; m->o_qg = o; o += ng;
  for (int q = 0; q < 4; q++) m->o_f[q] = -1;
  if (is_PB(form) || is_CB(form)) for (int q = 0; q < 4; q++) { m->o_f[q] = o; o += nr; }
  m->o_cg = -1; if (flags & FLAG_OBJ_LINEAR) { m->o_cg = o; o += ng; }
  m->o_bss = -1; if (flags & FLAG_FACTS) { m->o_bss = o; o += net->nbss; }
  m->o_tgh = -1; if (flags & FLAG_OBJ_LINEAR) { m->o_tgh = o; for (int64_t g = 0; g < ng; g++) o += net->lps[g]; }
  m->n_var = o;
  /* structure pass */
  ctx c; memset(&c, 0, sizeof c); c.m = m; c.mode = 0;
  walk(&c);
  m->m_nl = c.row; m->ncj = c.cj; m->nch = c.ch;
  m->jslot = malloc((size_t)(c.cj ? c.cj : 1) * 4); m->hslot = malloc((size_t)(c.ch ? c.ch : 1) * 4);
  m->jptr = calloc(m->m_nl + 1, 8); m->hptr = calloc(m->m_nl + 1, 8);
  /* Jacobian: per row sorted unique columns */
  key3 *kj = malloc((size_t)(c.cj ? c.cj : 1) * sizeof(key3));
  for (int64_t q = 0; q < c.cj; q++) { kj[q].row = c.cjrow[q]; kj[q].a = c.cjvar[q]; kj[q].b = 0; kj[q].id = q; }
  qsort(kj, c.cj, sizeof(key3), cmp3);
  m->jrow = malloc((size_t)(c.cj ? c.cj : 1) * 8); m->jcol = malloc((size_t)(c.cj ? c.cj : 1) * 8);
  int64_t ns = 0;
  for (int64_t q = 0; q < c.cj; q++) {
    if (q == 0 || kj[q].row != kj[q - 1].row || kj[q].a != kj[q - 1].a) { m->jrow[ns] = kj[q].row + 1; m->jcol[ns] = kj[q].a + 1; ns++; }
    m->jslot[kj[q].id] = (int32_t)(ns - 1);
  }
  m->nnzJ = ns;
  for (int64_t s = 0; s < ns; s++) m->jptr[m->jrow[s]]++;
  for (int64_t r = 0; r < m->m_nl; r++) m->jptr[r + 1] += m->jptr[r];
  free(kj);
  /* Hessian: per row, vertices ascending (diagonal slots) then off-diagonal pairs ascending */
  key3 *kv = malloc((size_t)(2 * c.ch + 1) * sizeof(key3));     /* vertices: every endpoint of every pair */
  for (int64_t q = 0; q < c.ch; q++) { kv[2 * q] = (key3){c.chrow[q], c.cha[q], 0, 0}; kv[2 * q + 1] = (key3){c.chrow[q], c.chb[q], 0, 0}; }
  qsort(kv, 2 * c.ch, sizeof(key3), cmp3);
  key3 *ko = malloc((size_t)(c.ch + 1) * sizeof(key3)); int64_t no = 0;
  for (int64_t q = 0; q < c.ch; q++) if (c.cha[q] != c.chb[q]) ko[no++] = (key3){c.chrow[q], c.cha[q], c.chb[q], q};
  qsort(ko, no, sizeof(key3), cmp3);
  /* count per row */
  int64_t *nd = calloc(m->m_nl + 1, 8), *nof = calloc(m->m_nl + 1, 8);
  for (int64_t q = 0; q < 2 * c.ch; q++) if (q == 0 || kv[q].row != kv[q - 1].row || kv[q].a != kv[q - 1].a) nd[kv[q].row]++;
  for (int64_t q = 0; q < no; q++) if (q == 0 || ko[q].row != ko[q - 1].row || ko[q].a != ko[q - 1].a || ko[q].b != ko[q - 1].b) nof[ko[q].row]++;
  for (int64_t r = 0; r < m->m_nl; r++) m->hptr[r + 1] = m->hptr[r] + nd[r] + nof[r];
  m->nnzH = m->hptr[m->m_nl];
  m->hrow = malloc((size_t)(m->nnzH ? m->nnzH : 1) * 8); m->hcol = malloc((size_t)(m->nnzH ? m->nnzH : 1) * 8);
  /* diagonal slots + lookup of (row,var) -> slot by binary search over the unique vertex list */
  int64_t nuv = 0; key3 *uv = malloc((size_t)(2 * c.ch + 1) * sizeof(key3));
  { int64_t pos = 0, currow = -1;
    for (int64_t q = 0; q < 2 * c.ch; q++) if (q == 0 || kv[q].row != kv[q - 1].row || kv[q].a != kv[q - 1].a) {
      if (kv[q].row != currow) { currow = kv[q].row; pos = m->hptr[currow]; }
      m->hrow[pos] = kv[q].a + 1; m->hcol[pos] = kv[q].a + 1;
      uv[nuv++] = (key3){kv[q].row, kv[q].a, 0, pos}; pos++; } }
  { int64_t pos = 0, currow = -1;
    for (int64_t q = 0; q < no; q++) {
      int isnew = (q == 0 || ko[q].row != ko[q - 1].row || ko[q].a != ko[q - 1].a || ko[q].b != ko[q - 1].b);
      if (ko[q].row != currow) { currow = ko[q].row; pos = m->hptr[currow] + nd[currow]; }
      if (isnew) { m->hrow[pos] = ko[q].a + 1; m->hcol[pos] = ko[q].b + 1; pos++; }
      m->hslot[ko[q].id] = (int32_t)(pos - 1); } }
  for (int64_t q = 0; q < c.ch; q++) if (c.cha[q] == c.chb[q]) {
    int64_t lo = 0, hi = nuv - 1, found = -1;
    while (lo <= hi) { int64_t mid = (lo + hi) / 2;
      if (uv[mid].row < c.chrow[q] || (uv[mid].row == c.chrow[q] && uv[mid].a < c.cha[q])) lo = mid + 1;
      else if (uv[mid].row == c.chrow[q] && uv[mid].a == c.cha[q]) { found = uv[mid].id; break; }
      else hi = mid - 1; }
    m->hslot[q] = (int32_t)found;
  }
  free(kv); free(ko); free(uv); free(nd); free(nof);
  free(c.cjrow); free(c.cjvar); free(c.chrow); free(c.cha); free(c.chb);
  /* bounds (C.1): '==' -> (0,0); '<=' -> (-inf,0) */
  m->lo = malloc((size_t)(m->m_nl ? m->m_nl : 1) * 8); m->hi = malloc((size_t)(m->m_nl ? m->m_nl : 1) * 8);
  int64_t r = 0;
  if (form != PBRAWV && form != PNRAWV) for (; r < 2 * nb; r++) { m->lo[r] = 0; m->hi[r] = 0; }
  for (int64_t k = 0; k < nr; k++) {
    if (form == PBRAPV || form == PBRARV) for (int q = 0; q < 4; q++, r++) { m->lo[r] = 0; m->hi[r] = 0; }
    if (form == PBRAWV || form == PNRAWV) { m->lo[r] = 0; m->hi[r] = 0; r++; }
    if (form != PNRAWV) for (int q = 0; q < 2; q++, r++) { m->lo[r] = -INFINITY; m->hi[r] = 0; }
  }
  return m;
}

void pso_destroy(pso_model *m) {
  if (!m) return;
  pso_net *N = &m->net;
  const void *ptrs[] = {N->br_f, N->br_t, N->Gf, N->Bf, N->Gt, N->Bt, N->gf, N->bf, N->gt, N->bt, N->Imax, N->Pd, N->Qd, N->Gl, N->Bl,
                        N->gen_ptr, N->gen_idx, N->sij_ptr, N->sij_idx, N->sji_ptr, N->sji_idx, N->sh_ptr, N->sh_idx, N->lps,
                        m->jrow, m->jcol, m->hrow, m->hcol, m->jptr, m->hptr, m->lo, m->hi, m->jslot, m->hslot};
  for (size_t q = 0; q < sizeof ptrs / sizeof *ptrs; q++) free((void *)ptrs[q]);
  free(m);
}

void pso_dims(const pso_model *m, int64_t *n_var, int64_t *m_nl, int64_t *nnzJ, int64_t *nnzH) {
  *n_var = m->n_var; *m_nl = m->m_nl; *nnzJ = m->nnzJ; *nnzH = m->nnzH;
}
void pso_structure(const pso_model *m, int64_t *jrow, int64_t *jcol, int64_t *hrow, int64_t *hcol) {
  memcpy(jrow, m->jrow, (size_t)m->nnzJ * 8); memcpy(jcol, m->jcol, (size_t)m->nnzJ * 8);
  memcpy(hrow, m->hrow, (size_t)m->nnzH * 8); memcpy(hcol, m->hcol, (size_t)m->nnzH * 8);
}
void pso_bounds(const pso_model *m, double *lo, double *hi) {
  memcpy(lo, m->lo, (size_t)m->m_nl * 8); memcpy(hi, m->hi, (size_t)m->m_nl * 8);
}

/* one scenario.  Pd/Qd NULL -> the network's own loads.  Returns 0, or 1 if an analytic second
 * derivative fell outside the declared structure. */
int pso_eval(const pso_model *m, int what, const double *x, const double *Pd, const double *Qd,
             const double *mu, double *g, double *J, double *H) {
  ctx c; memset(&c, 0, sizeof c);
  c.m = m; c.mode = 1; c.what = what; c.x = x; c.mu = mu; c.g = g; c.J = J; c.H = H;
  c.Pd = Pd ? Pd : m->net.Pd; c.Qd = Qd ? Qd : m->net.Qd;
  if (what & EV_J) memset(J, 0, (size_t)m->nnzJ * 8);
  if (what & EV_H) memset(H, 0, (size_t)m->nnzH * 8);
  walk(&c);
  return c.bad;
}

/* S scenarios, OpenMP over scenarios.  Strides in doubles; Pd/Qd/mu may be NULL. */
int pso_eval_batch(const pso_model *m, int what, int64_t S, const double *x, int64_t ldx,
                   const double *Pd, const double *Qd, int64_t ldl, const double *mu, int64_t ldmu,
                   double *g, int64_t ldg, double *J, int64_t ldJ, double *H, int64_t ldH, int nthreads) {
  int bad = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
  for (int64_t s = 0; s < S; s++)
    bad |= pso_eval(m, what, x + s * ldx, Pd ? Pd + s * ldl : NULL, Qd ? Qd + s * ldl : NULL,
                    mu ? mu + s * ldmu : NULL, g ? g + s * ldg : NULL, J ? J + s * ldJ : NULL, H ? H + s * ldH : NULL);
  return bad;
}

int pso_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
