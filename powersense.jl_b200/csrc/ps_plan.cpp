// Plan builder (host, C++): see ps_plan.h.  Everything here is structure only -- no values.
#include "ps_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <utility>

namespace ps {
namespace {

typedef std::pair<int64_t, int64_t> pr;

struct RowAcc {                       // contributions of one row in emission order
  std::vector<int64_t> jv;
  std::vector<pr> hp;                 // (max,min)
  void J(int64_t v) { jv.push_back(v); }
  void H(int64_t a, int64_t b) { hp.push_back(a >= b ? pr(a, b) : pr(b, a)); }
  void clear() { jv.clear(); hp.clear(); }
};

struct RowCanon {                     // canonical layout of one row
  std::vector<int64_t> cols;          // sorted unique (C.2)
  std::vector<int64_t> verts;         // sorted unique Hessian vertices -> diagonal slots (C.3)
  std::vector<pr> off;                // sorted unique off-diagonal (max,min)
  std::vector<int32_t> jslot, hslot;  // per contribution
};

void canon(const RowAcc& a, RowCanon& c) {
  c.cols = a.jv;
  std::sort(c.cols.begin(), c.cols.end());
  c.cols.erase(std::unique(c.cols.begin(), c.cols.end()), c.cols.end());
  c.jslot.resize(a.jv.size());
  for (size_t q = 0; q < a.jv.size(); q++)
    c.jslot[q] = (int32_t)(std::lower_bound(c.cols.begin(), c.cols.end(), a.jv[q]) - c.cols.begin());
  c.verts.clear();
  c.off.clear();
  for (const pr& p : a.hp) {
    c.verts.push_back(p.first);
    c.verts.push_back(p.second);
    if (p.first != p.second) c.off.push_back(p);
  }
  std::sort(c.verts.begin(), c.verts.end());
  c.verts.erase(std::unique(c.verts.begin(), c.verts.end()), c.verts.end());
  std::sort(c.off.begin(), c.off.end());
  c.off.erase(std::unique(c.off.begin(), c.off.end()), c.off.end());
  c.hslot.resize(a.hp.size());
  for (size_t q = 0; q < a.hp.size(); q++) {
    const pr& p = a.hp[q];
    if (p.first == p.second)
      c.hslot[q] = (int32_t)(std::lower_bound(c.verts.begin(), c.verts.end(), p.first) - c.verts.begin());
    else
      c.hslot[q] = (int32_t)(c.verts.size() + (std::lower_bound(c.off.begin(), c.off.end(), p) - c.off.begin()));
  }
}

void append_row(Plan& P, int64_t row, const RowCanon& c, bool eq) {
  P.jptr.push_back((int32_t)P.jrow.size());
  P.hptr.push_back((int32_t)P.hrow.size());
  for (int64_t v : c.cols) { P.jrow.push_back(row + 1); P.jcol.push_back(v + 1); }
  for (int64_t v : c.verts) { P.hrow.push_back(v + 1); P.hcol.push_back(v + 1); }
  for (const pr& p : c.off) { P.hrow.push_back(p.first + 1); P.hcol.push_back(p.second + 1); }
  P.lo.push_back(eq ? 0.0 : -INFINITY);
  P.hi.push_back(0.0);
}

}  // namespace

std::string build_plan(const HostNet& net, int form, int src, int flags, int tpb, int bus_tpb, Plan& P) {
  if (form < 0 || form >= NFORM) return "unknown formulation code";
  if (src != SRC_MATPOWER && src != SRC_ARPAE) return "unknown source_type code";
  P = Plan();
  P.form = form; P.src = src; P.flags = flags; P.net = net;
  const int64_t nb = net.nbus, nr = net.nbr, ng = net.ngen;
  for (int64_t k = 0; k < nr; k++) {
    if (net.f[k] < 0 || net.f[k] >= nb || net.t[k] < 0 || net.t[k] >= nb) return "branch endpoint out of range";
    if (net.f[k] == net.t[k]) return "self-loop branch (f == t) is not supported";
  }
  // ---- variable order (formulations.jl:20-67, :121)
  int64_t o = 0;
  P.o_a = o; o += nb; P.o_b = o; o += nb;
  if (has_W(form)) { P.o_Wd = o; o += nb; }
  if (has_WrWi(form)) { P.o_Wr = o; o += nr; P.o_Wi = o; o += nr; }
  P.o_pg = o; o += ng; P.o_qg = o; o += ng;
  if (is_PB(form) || is_CB(form)) for (int q = 0; q < 4; q++) { P.o_f[q] = o; o += nr; }
  if (flags & FLAG_OBJ_LINEAR) { P.o_cg = o; o += ng; }
  if (flags & FLAG_FACTS) { P.o_bss = o; o += net.nbss; }
  if (flags & FLAG_OBJ_LINEAR) { P.o_tgh = o; o += net.n_tgh; }
  P.n_var = o;
  const bool facts = (flags & FLAG_FACTS) != 0;
  const bool polar = is_polar(form);

  // ---- bus rows (formulations.jl:170-297): emission order == the bus kernel's order
  int64_t row = 0;
  P.bus_jpp.assign(1, 0);
  P.bus_hpp.assign(1, 0);
  if (has_bus_rows(form)) {
    RowAcc acc[2];
    RowCanon can[2];
    std::vector<std::pair<int, int>> sj, sh;     // streams: (row 0/1, contribution index)
    for (int64_t i = 0; i < nb; i++) {
      acc[0].clear(); acc[1].clear(); sj.clear(); sh.clear();
      auto J = [&](int r, int64_t v) { sj.push_back({r, (int)acc[r].jv.size()}); acc[r].J(v); };
      auto H = [&](int r, int64_t a, int64_t b) { sh.push_back({r, (int)acc[r].hp.size()}); acc[r].H(a, b); };
      for (int32_t q = net.gen_ptr[i]; q < net.gen_ptr[i + 1]; q++) {              // :177-180
        J(0, P.o_pg + net.gen_idx[q]);
        J(1, P.o_qg + net.gen_idx[q]);
      }
      for (int side = 0; side < 2; side++) {
        const std::vector<int32_t>& ptr = side ? net.sji_ptr : net.sij_ptr;
        const std::vector<int32_t>& idx = side ? net.sji_idx : net.sij_idx;
        for (int32_t q = ptr[i]; q < ptr[i + 1]; q++) {
          const int64_t k = idx[q];
          const int64_t a = side ? net.t[k] : net.f[k], b = side ? net.f[k] : net.t[k];
          if (a != i) return "adjacency list inconsistent with branch endpoints";
          if (form == PBRAPV || form == PBRARV) {                                  // :188-196
            J(0, P.o_f[0 + 2 * side] + k);
            J(1, P.o_f[1 + 2 * side] + k);
          } else if (is_CB(form)) {                                                // :206-214
            const int64_t ra = P.o_b + a, ia = P.o_a + a, Ir = P.o_f[0 + 2 * side] + k, Ii = P.o_f[1 + 2 * side] + k;
            for (int r = 0; r < 2; r++) { J(r, ra); J(r, ia); J(r, Ir); J(r, Ii); }
            H(0, ra, ra); H(0, Ir, Ir); H(0, ra, Ir); H(0, ia, ia); H(0, Ii, Ii); H(0, ia, Ii);
            H(1, ia, ia); H(1, Ir, Ir); H(1, ia, Ir); H(1, ra, ra); H(1, Ii, Ii); H(1, ra, Ii);
          } else if (form == PNPAPV || form == PNRAPV) {                           // :215-234
            const int64_t v[4] = {P.o_a + a, P.o_a + b, P.o_b + a, P.o_b + b};
            for (int r = 0; r < 2; r++) for (int q4 = 0; q4 < 4; q4++) J(r, v[q4]);
            for (int r = 0; r < 2; r++) {
              for (int q4 = 0; q4 < 4; q4++) H(r, v[q4], v[q4]);
              for (int x = 0; x < 4; x++) for (int y = x + 1; y < 4; y++) H(r, v[x], v[y]);
            }
          } else if (form == PNRARV) {                                             // :235-243
            const int64_t ra = P.o_b + a, ia = P.o_a + a, rb = P.o_b + b, ib = P.o_a + b;
            for (int r = 0; r < 2; r++) { J(r, ra); J(r, ia); J(r, rb); J(r, ib); }
            for (int r = 0; r < 2; r++) {
              H(r, ra, ra); H(r, ia, ia); H(r, rb, rb); H(r, ib, ib);
              H(r, ra, rb); H(r, ia, ib); H(r, ra, ib); H(r, ia, rb);
            }
          }
        }
      }
      // shunt terms :256-279 (always emitted, even when Gl = Bl = 0)
      if (polar) {
        const int64_t Vi = P.o_b + i;
        J(0, Vi); J(1, Vi); H(0, Vi, Vi); H(1, Vi, Vi);
      } else if (form == PBRARV || form == CBRARV || form == PNRARV) {
        const int64_t r_ = P.o_b + i, i_ = P.o_a + i;
        for (int r = 0; r < 2; r++) { J(r, r_); J(r, i_); }
        for (int r = 0; r < 2; r++) { H(r, r_, r_); H(r, i_, i_); }
      } else if (form == CBRAWV) {
        J(0, P.o_Wd + i); J(1, P.o_Wd + i);
      }
      if (facts) for (int32_t q = net.sh_ptr[i]; q < net.sh_ptr[i + 1]; q++) {     // :259-263,:267-271,:275-279
        const int64_t bs = P.o_bss + net.sh_idx[q];
        if (polar) {
          const int64_t Vi = P.o_b + i;
          J(1, bs); J(1, Vi); H(1, bs, bs); H(1, Vi, Vi); H(1, bs, Vi);
        } else if (form == CBRAWV) {
          const int64_t W = P.o_Wd + i;
          J(1, bs); J(1, W); H(1, bs, bs); H(1, W, W); H(1, bs, W);
        } else {
          const int64_t r_ = P.o_b + i, i_ = P.o_a + i;
          J(1, bs); J(1, r_); J(1, i_);
          H(1, bs, bs); H(1, r_, r_); H(1, i_, i_); H(1, bs, r_); H(1, bs, i_); H(1, r_, i_);
        }
      }
      canon(acc[0], can[0]);
      canon(acc[1], can[1]);
      const int64_t lenJ0 = (int64_t)can[0].cols.size(), lenH0 = (int64_t)(can[0].verts.size() + can[0].off.size());
      const int64_t lenJ = lenJ0 + (int64_t)can[1].cols.size();
      const int64_t lenH = lenH0 + (int64_t)(can[1].verts.size() + can[1].off.size());
      if (lenJ >= 32768 || lenH >= 32768) return "bus degree too large for 15-bit slot positions";
      std::vector<char> seenJ(lenJ, 0), seenH(lenH, 0);
      for (auto& e : sj) {
        int64_t p = (e.first ? lenJ0 : 0) + can[e.first].jslot[e.second];
        P.bus_jpos.push_back((uint16_t)(p | (seenJ[p] ? 0x8000 : 0)));
        seenJ[p] = 1;
      }
      for (auto& e : sh) {
        int64_t p = (e.first ? lenH0 : 0) + can[e.first].hslot[e.second];
        P.bus_hpos.push_back((uint16_t)(p | (seenH[p] ? 0x8000 : 0)));
        seenH[p] = 1;
      }
      for (int64_t p = 0; p < lenH; p++) if (!seenH[p]) return "internal: Hessian slot without contribution";
      P.bus_jpp.push_back((int32_t)P.bus_jpos.size());
      P.bus_hpp.push_back((int32_t)P.bus_hpos.size());
      append_row(P, row++, can[0], true);                                          // :291
      append_row(P, row++, can[1], true);                                          // :292
    }
  } else {
    P.bus_jpp.assign(nb + 1, 0);
    P.bus_hpp.assign(nb + 1, 0);
  }
  P.bus_rows = row;

  // ---- branch rows (formulations.jl:299-428)
  const BranchPat pat = branch_pattern(form, src);
  P.rows_per_branch = pat.nrows;
  bool have_tab[2] = {false, false};
  {
    RowAcc acc;
    RowCanon can;
    for (int64_t k = 0; k < nr; k++) {
      const int64_t f = net.f[k], t = net.t[k];
      const int swap = f > t;
      auto var = [&](int8_t role) -> int64_t {
        switch (role) {
          case R_AF: return P.o_a + f;
          case R_AT: return P.o_a + t;
          case R_BF: return P.o_b + f;
          case R_BT: return P.o_b + t;
          case R_WDF: return P.o_Wd + f;
          case R_WDT: return P.o_Wd + t;
          case R_WR: return P.o_Wr + k;
          case R_WI: return P.o_Wi + k;
          default: return P.o_f[role - R_F0] + k;
        }
      };
      uint8_t jp[MAX_BR_JTAGS] = {}, hp[MAX_BR_HTAGS] = {};
      int jbase = 0, hbase = 0, jt = 0, ht = 0;
      for (int r = 0; r < pat.nrows; r++) {
        const RowPat& rp = pat.rows[r];
        acc.clear();
        for (int q = 0; q < rp.nj; q++) acc.J(var(rp.j[q]));
        for (int q = 0; q < rp.nh; q++) acc.H(var(rp.h[q][0]), var(rp.h[q][1]));
        canon(acc, can);
        if ((int)can.cols.size() != rp.nj) return "internal: duplicate Jacobian tag in a branch row";
        if ((int)(can.verts.size() + can.off.size()) != rp.nh) return "internal: duplicate Hessian tag in a branch row";
        for (int q = 0; q < rp.nj; q++) jp[jt++] = (uint8_t)(jbase + can.jslot[q]);
        for (int q = 0; q < rp.nh; q++) hp[ht++] = (uint8_t)(hbase + can.hslot[q]);
        jbase += rp.nj;
        hbase += rp.nh;
        append_row(P, row++, can, rp.eq != 0);
      }
      P.RJ = jbase; P.RH = hbase;
      if (!have_tab[swap]) {
        std::copy(jp, jp + MAX_BR_JTAGS, P.br_jpos[swap]);
        std::copy(hp, hp + MAX_BR_HTAGS, P.br_hpos[swap]);
        have_tab[swap] = true;
      } else if (!std::equal(jp, jp + jt, P.br_jpos[swap]) || !std::equal(hp, hp + ht, P.br_hpos[swap])) {
        return "internal: branch slot pattern is not uniform";
      }
    }
    if (nr == 0) { P.RJ = pat.nj(); P.RH = pat.nh(); }
    // the kernels use the compile-time tables of branch_slots(); they must agree with the ones
    // derived above from the real variable indices
    const BranchSlots sl = branch_slots(form, src);
    if (sl.nj != P.RJ || sl.nh != P.RH || sl.nrows != pat.nrows) return "internal: branch_slots size mismatch";
    for (int sw = 0; sw < 2; sw++) {
      if (!have_tab[sw]) {
        std::copy(sl.jpos[sw], sl.jpos[sw] + MAX_BR_JTAGS, P.br_jpos[sw]);
        std::copy(sl.hpos[sw], sl.hpos[sw] + MAX_BR_HTAGS, P.br_hpos[sw]);
      } else if (!std::equal(sl.jpos[sw], sl.jpos[sw] + P.RJ, P.br_jpos[sw]) ||
                 !std::equal(sl.hpos[sw], sl.hpos[sw] + P.RH, P.br_hpos[sw])) {
        return "internal: compile-time branch slot table disagrees with the variable order";
      }
    }
  }
  P.m_nl = row;
  P.jptr.push_back((int32_t)P.jrow.size());
  P.hptr.push_back((int32_t)P.hrow.size());
  P.nnzJ = (int64_t)P.jrow.size();
  P.nnzH = (int64_t)P.hrow.size();
  if (P.nnzJ >= (int64_t(1) << 31) || P.nnzH >= (int64_t(1) << 31)) return "structure too large for 32-bit slot offsets";
  {
    std::vector<int64_t> u(P.jcol);
    std::sort(u.begin(), u.end());
    P.nx_nl = (int64_t)(std::unique(u.begin(), u.end()) - u.begin());
    std::vector<int64_t> ub(P.jcol.begin() + P.jptr[P.bus_rows], P.jcol.end());
    std::sort(ub.begin(), ub.end());
    P.nx_branch = (int64_t)(std::unique(ub.begin(), ub.end()) - ub.begin());
  }

  // ---- derived per-side constants (host-precomputed like formulations.jl:217,222,338-350)
  P.c0.assign(2 * nr, 0.0); P.c1.assign(2 * nr, 0.0); P.c2.assign(2 * nr, 0.0); P.c3.assign(2 * nr, 0.0);
  for (int side = 0; side < 2; side++)
    for (int64_t k = 0; k < nr; k++) {
      const double gs = side ? net.gt[k] : net.gf[k], bs = side ? net.bt[k] : net.bf[k];
      const double G = side ? net.Gt[k] : net.Gf[k], B = side ? net.Bt[k] : net.Bf[k];
      const int64_t q = side * nr + k;
      if (form == PNPAPV || form == PNRAPV) {
        // T2 (:217,:222): |G+jB|, angle(G+jB) in c0,c1 ; L4 (:338-342): y3 in c2, th1-th2 pieces in c3 (th1), c1 (th2)
        P.c0[q] = std::hypot(G, B);
        P.c1[q] = std::atan2(B, G);
        P.c2[q] = 2.0 * std::hypot(gs, bs) * std::hypot(G, B);
        P.c3[q] = std::atan2(bs, gs);
      } else if (form == PNRARV) {
        // L5 (:345-346,:349-350): y, Y, y1, y2
        P.c0[q] = gs * gs + bs * bs;
        P.c1[q] = G * G + B * B;
        P.c2[q] = 2.0 * (gs * G + bs * B);
        P.c3[q] = 2.0 * (bs * G - gs * B);
      }
    }

  // ---- tiles
  P.tile_units_br = tpb;
  // half-edge tables (bus kernel order) and the PN* constants staged in shared memory per tile
  P.he_ptr.assign(nb + 1, 0);
  for (int64_t i = 0; i < nb; i++) {
    for (int side = 0; side < 2; side++) {
      const std::vector<int32_t>& ptr = side ? net.sji_ptr : net.sij_ptr;
      const std::vector<int32_t>& idx = side ? net.sji_idx : net.sij_idx;
      for (int32_t q = ptr[i]; q < ptr[i + 1]; q++) {
        const int32_t k = idx[q];
        P.he_ks.push_back(2 * k + side);
        P.he_other.push_back(side ? net.f[k] : net.t[k]);
        P.he_bus.push_back((int32_t)i);
      }
    }
    P.he_ptr[i + 1] = (int32_t)P.he_ks.size();
  }
  const int64_t nhe = (int64_t)P.he_ks.size();
  if (is_PN(form)) {
    P.he_c.assign(4 * nhe, 0.0);
    for (int64_t q = 0; q < nhe; q++) {
      const int64_t k = P.he_ks[q] >> 1;
      const int side = P.he_ks[q] & 1;
      P.he_c[0 * nhe + q] = side ? net.gt[k] : net.gf[k];
      P.he_c[1 * nhe + q] = side ? net.bt[k] : net.bf[k];
      if (form == PNPAPV) {                       // |G+jB|, angle(G+jB)  (:217, :222)
        P.he_c[2 * nhe + q] = P.c0[side * nr + k];
        P.he_c[3 * nhe + q] = P.c1[side * nr + k];
      } else {
        P.he_c[2 * nhe + q] = side ? net.Gt[k] : net.Gf[k];
        P.he_c[3 * nhe + q] = side ? net.Bt[k] : net.Bf[k];
      }
    }
  }
  // slot -> contribution maps of the bus block (inverse of the position streams, stable in emission order)
  if (has_bus_rows(form)) {
    auto invert = [&](const std::vector<uint16_t>& pos, const std::vector<int32_t>& pp, const std::vector<int32_t>& ptr,
                      std::vector<int32_t>& sp, std::vector<int32_t>& perm) {
      const int64_t nslots = ptr[2 * nb];
      sp.assign(nslots + 1, 0);
      perm.assign(pos.size(), 0);
      for (int64_t i = 0; i < nb; i++)
        for (int32_t c = pp[i]; c < pp[i + 1]; c++) sp[ptr[2 * i] + (pos[c] & 0x7fff) + 1]++;
      for (int64_t e = 0; e < nslots; e++) sp[e + 1] += sp[e];
      std::vector<int32_t> fill(sp.begin(), sp.end() - 1);
      for (int64_t i = 0; i < nb; i++)
        for (int32_t c = pp[i]; c < pp[i + 1]; c++) perm[fill[ptr[2 * i] + (pos[c] & 0x7fff)]++] = c;   // ascending c per slot
    };
    invert(P.bus_jpos, P.bus_jpp, P.jptr, P.jsp, P.jperm);
    invert(P.bus_hpos, P.bus_hpp, P.hptr, P.hsp, P.hperm);
    P.jsrc.assign(P.jsp.size() - 1, 0);            // filled per tile below (tile-relative cells)
    P.hsrc.assign(P.hsp.size() - 1, 0);
  }
  // PB* balances: check that the canonical slot order equals the emission order the specialised bus rows
  // assume ([voltage slots][gens][Sij][Sji], no accumulation); true whenever the adjacency lists are ascending
  if (form == PBRAPV || form == PBRARV) {
    const int nv = form == PBRARV ? 2 : 1;
    bool ok = true;
    for (int64_t i = 0; i < nb && ok; i++) {
      const int64_t ngi = net.gen_ptr[i + 1] - net.gen_ptr[i], nhi = P.he_ptr[i + 1] - P.he_ptr[i];
      const int64_t nshi = facts ? net.sh_ptr[i + 1] - net.sh_ptr[i] : 0;
      if (nshi) continue;                          // such tiles take the generic path
      const int64_t len = nv + ngi + nhi;
      const uint16_t* jp = P.bus_jpos.data() + P.bus_jpp[i];
      if (P.bus_jpp[i + 1] - P.bus_jpp[i] != 2 * len) { ok = false; break; }
      // emission: (P,Q) per gen, (P,Q) per half-edge, then the voltage terms: rect J(0,vr) J(0,vi) J(1,vr) J(1,vi)
      int64_t e = 0;
      for (int64_t c = 0; c < ngi + nhi && ok; c++) {
        ok = jp[e] == (uint16_t)(nv + c) && jp[e + 1] == (uint16_t)(len + nv + c);
        e += 2;
      }
      if (!ok) break;
      if (nv == 2) ok = jp[e] == 1 && jp[e + 1] == 0 && jp[e + 2] == len + 1 && jp[e + 3] == len;
      else ok = jp[e] == 0 && jp[e + 1] == len;
      if (P.hptr[2 * i + 2] - P.hptr[2 * i] != 2 * nv) ok = false;
    }
    P.pb_implicit = ok;
    // slot codes of the whole bus block (only when no bus carries a switched shunt term)
    bool any_sh = false;
    if (facts) for (int64_t i = 0; i < nb; i++) any_sh = any_sh || net.sh_ptr[i + 1] > net.sh_ptr[i];
    if (ok && !any_sh && nb > 0) {
      P.pb_jcode.assign(P.jptr[P.bus_rows], -1);
      P.pb_hcode.assign(P.hptr[P.bus_rows], 0);
      for (int64_t i = 0; i < nb; i++) {
        const int64_t ngi = net.gen_ptr[i + 1] - net.gen_ptr[i], nhi = P.he_ptr[i + 1] - P.he_ptr[i];
        for (int pq = 0; pq < 2; pq++) {
          int32_t* jc = P.pb_jcode.data() + P.jptr[2 * i + pq];
          if (nv == 2) { jc[0] = -2 - (int32_t)(4 * i + 2 * pq + 0); jc[1] = -2 - (int32_t)(4 * i + 2 * pq + 1); }
          else jc[0] = -2 - (int32_t)(4 * i + 2 * pq + 1);
          for (int64_t c = 0; c < nhi; c++) jc[nv + ngi + c] = P.he_ks[P.he_ptr[i] + c] >> 1;
          int32_t* hc = P.pb_hcode.data() + P.hptr[2 * i + pq];
          for (int q = 0; q < nv; q++) hc[q] = (int32_t)(2 * i + pq);
        }
      }
      P.pb_slots = true;
      // element-ordered term tables of the scenario-resident residual kernel
      const int64_t ng = net.ngen;
      P.pbs_code.assign((size_t)(ng + 2 * nr), -1);
      P.pbs_ok = nb < (1 << 24);
      auto fill = [&](const std::vector<int32_t>& ptr, const std::vector<int32_t>& idx, int32_t* code) {
        for (int64_t i = 0; i < nb && P.pbs_ok; i++)
          for (int32_t q = ptr[i]; q < ptr[i + 1]; q++) {
            if ((q > ptr[i] && idx[q] <= idx[q - 1]) || code[idx[q]] != -1) { P.pbs_ok = false; break; }   // list order == element order
            code[idx[q]] = (int32_t)i;
          }
      };
      fill(net.gen_ptr, net.gen_idx, P.pbs_code.data());
      fill(net.sij_ptr, net.sij_idx, P.pbs_code.data() + ng);
      fill(net.sji_ptr, net.sji_idx, P.pbs_code.data() + ng + nr);
      std::vector<int32_t> seen((size_t)nb, -1), cnt((size_t)nb, 0);
      auto ranks = [&](int32_t* code, int64_t n) {
        for (int64_t w0 = 0; w0 < n && P.pbs_ok; w0 += PBS_WAVE) {
          int mx = 0;
          for (int64_t e = w0; e < std::min<int64_t>(n, w0 + PBS_WAVE); e++) {
            const int32_t b = code[e];
            if (b < 0) continue;
            if (seen[b] != (int32_t)(w0 / PBS_WAVE)) { seen[b] = (int32_t)(w0 / PBS_WAVE); cnt[b] = 0; }
            const int r = cnt[b]++;
            if (r > 127) { P.pbs_ok = false; break; }
            code[e] = b | (r << 24);
            mx = std::max(mx, r);
          }
          P.pbs_wmax.push_back((uint8_t)mx);
        }
        std::fill(seen.begin(), seen.end(), -1);     // step ids restart with every pass
      };
      ranks(P.pbs_code.data(), ng);
      ranks(P.pbs_code.data() + ng, nr);
      ranks(P.pbs_code.data() + ng + nr, nr);
      // a bus's terms inside one step are served in rank order, one CTA barrier per rank: worth it on a randomly numbered
      // branch list (1-3 ranks per step), not on a from-bus ordered one, where a step holds every outgoing branch of its
      // buses (and where the list-walking kernel's gathers hit the same sectors anyway)
      int64_t rounds = 0;
      for (uint8_t w : P.pbs_wmax) rounds += w + 1;
      P.pbs_prefer = P.pbs_ok && rounds <= 3 * (int64_t)P.pbs_wmax.size() + 8;
      if (!P.pbs_ok) { P.pbs_code.clear(); P.pbs_wmax.clear(); }
    }
  }
  // bus tiles: greedy, at most tpb buses and at most `budget` bytes of shared memory per tile
  const int npieces = he_np(form);                 // g pieces per half-edge
  // prefix counts (per row) of slots with more than one contribution
  std::vector<int64_t> nmJ(2 * nb + 1, 0), nmH(2 * nb + 1, 0);
  if (has_bus_rows(form))
    for (int64_t r = 0; r < 2 * nb; r++) {
      int64_t cj = 0, ch = 0;
      for (int32_t e = P.jptr[r]; e < P.jptr[r + 1]; e++) cj += (P.jsp[e + 1] - P.jsp[e] > 1);
      for (int32_t e = P.hptr[r]; e < P.hptr[r + 1]; e++) ch += (P.hsp[e + 1] - P.hsp[e] > 1);
      nmJ[r + 1] = nmJ[r] + cj;
      nmH[r + 1] = nmH[r] + ch;
    }
  auto bus_tile_bytes = [&](int64_t u0, int64_t u1) -> int64_t {
    const int64_t nbt = u1 - u0, nh = P.he_ptr[u1] - P.he_ptr[u0], ngt = net.gen_ptr[u1] - net.gen_ptr[u0];
    const int64_t nsh = facts ? net.sh_ptr[u1] - net.sh_ptr[u0] : 0;
    const int64_t jn = P.jptr[2 * u1] - P.jptr[2 * u0], hn = P.hptr[2 * u1] - P.hptr[2 * u0];
    const int64_t nx = 2 * nbt + 2 * nh + 2 * ngt + (form == CBRAWV ? nbt : 0) + nsh + 4 * nbt;
    const int64_t ncj = P.bus_jpp[u1] - P.bus_jpp[u0], nch = P.bus_hpp[u1] - P.bus_hpp[u0];
    int64_t b = 8 * (ncj + nch + npieces * nh + 2 * nx + (is_PN(form) ? 4 * nh : 0) + 2);
    b += 4 * (nh + (is_PN(form) ? nh : 0) + ngt + nsh + 2);
    const int64_t nm = (nmJ[2 * u1] - nmJ[2 * u0]) + (nmH[2 * u1] - nmH[2 * u0]);            // multi-slots of the tile
    b += 2 * (nh + jn + hn + (2 * nm + 2) + ((ncj - jn) + (nch - hn)) + 16);   // hbus, jsrc, hsrc, mdst + mptr, mlist
    b += nh + 64;                                  // status bytes + alignment slack
    return b;
  };
  int64_t budget = 44 * 1024;                              // 5 CTAs / SM (bus_kernel is built for 5 resident CTAs)
  if (const char* e = std::getenv("PS_BUS_SMEM_KB")) { const int v = std::atoi(e); if (v >= 8 && v <= 200) budget = (int64_t)v * 1024; }   // development knob
  const int64_t hard = 220 * 1024;
  int64_t mx = 0;
  P.tiles.clear();
  if (has_bus_rows(form)) {
    int64_t u0 = 0;
    while (u0 < nb) {
      int64_t u1 = u0 + 1;
      if (bus_tile_bytes(u0, u1) > hard) return "a bus does not fit in shared memory (bus degree too large)";
      while (u1 < nb && u1 - u0 < bus_tpb && bus_tile_bytes(u0, u1 + 1) <= budget &&
             P.bus_jpp[u1 + 1] - P.bus_jpp[u0] < 60000 && P.bus_hpp[u1 + 1] - P.bus_hpp[u0] < 60000) u1++;
      Tile t{};
      t.kind = 0; t.u0 = (int32_t)u0; t.u1 = (int32_t)u1;
      t.g0 = (int32_t)(2 * u0); t.gn = (int32_t)(2 * (u1 - u0));
      t.j0 = P.jptr[2 * u0]; t.jn = P.jptr[2 * u1] - t.j0;
      t.h0 = P.hptr[2 * u0]; t.hn = P.hptr[2 * u1] - t.h0;
      // multi-slot descriptors of this tile
      // Shared-memory cell of a contribution (tile-relative).  Half-edge contributions are stored type-major
      // (cell = type * nhe_tile + half-edge: conflict-free stores by consecutive half-edge threads); the bus's own
      // contributions (generators, shunt terms) follow in emission order.
      const int32_t nhe_t = P.he_ptr[u1] - P.he_ptr[u0];
      auto cell = [&](int32_t c, bool isJ) -> int32_t {
        const std::vector<int32_t>& pp = isJ ? P.bus_jpp : P.bus_hpp;
        const int64_t i = (std::upper_bound(pp.begin() + u0, pp.begin() + u1 + 1, c) - pp.begin()) - 1;   // owning bus
        const int32_t per = isJ ? he_nj(form) : he_nh(form);
        const int32_t ng2 = isJ ? 2 * (net.gen_ptr[i + 1] - net.gen_ptr[i]) : 0;
        const int32_t deg = P.he_ptr[i + 1] - P.he_ptr[i];
        const int32_t off = c - pp[i];
        const int32_t own0 = per * nhe_t + (pp[i] - pp[u0]) - per * (P.he_ptr[i] - P.he_ptr[u0]);
        if (off < ng2) return own0 + off;
        if (off < ng2 + per * deg) {
          const int32_t lh = (off - ng2) / per, ty = (off - ng2) % per;
          return ty * nhe_t + (P.he_ptr[i] - P.he_ptr[u0]) + lh;
        }
        return own0 + off - per * deg;
      };
      auto multi = [&](const std::vector<int32_t>& sp, const std::vector<int32_t>& perm, std::vector<int32_t>& src, bool isJ,
                       int32_t e0, int32_t en, int32_t& m0, int32_t& mn, int32_t& l0, int32_t& ln) {
        std::vector<int32_t> ms;
        for (int32_t e = e0; e < e0 + en; e++) {
          src[e] = cell(perm[sp[e]], isJ);
          if (sp[e + 1] - sp[e] > 1) ms.push_back(e);
        }
        std::stable_sort(ms.begin(), ms.end(), [&](int32_t a, int32_t b) { return sp[a + 1] - sp[a] > sp[b + 1] - sp[b]; });
        m0 = (int32_t)P.ms_dst.size(); mn = (int32_t)ms.size(); l0 = (int32_t)P.ms_list.size();
        int32_t run = 0;
        for (int32_t e : ms) {
          P.ms_dst.push_back(src[e]);
          P.ms_ptr.push_back(run);
          for (int32_t q = sp[e] + 1; q < sp[e + 1]; q++) { P.ms_list.push_back(cell(perm[q], isJ)); run++; }
        }
        P.ms_ptr.push_back(run);                   // one sentinel per tile and kind (mn + 1 entries)
        ln = run;
      };
      // ms_ptr holds mn+1 entries per call, so its offset differs from ms_dst's: keep them in step with a pad
      {
        int32_t m0, mn, l0, ln;
        while (P.ms_dst.size() < P.ms_ptr.size()) P.ms_dst.push_back(0);
        multi(P.jsp, P.jperm, P.jsrc, true, t.j0, t.jn, m0, mn, l0, ln);
        t.mj0 = m0; t.mjn = mn; t.mjl0 = l0; t.mjln = ln;
        while (P.ms_dst.size() < P.ms_ptr.size()) P.ms_dst.push_back(0);
        multi(P.hsp, P.hperm, P.hsrc, false, t.h0, t.hn, m0, mn, l0, ln);
        t.mh0 = m0; t.mhn = mn; t.mhl0 = l0; t.mhln = ln;
      }
      mx = std::max<int64_t>(mx, bus_tile_bytes(u0, u1));
      P.tiles.push_back(t);
      u0 = u1;
    }
  }
  // ---- warp tiles (bus_warp_kernel): contiguous bus ranges with at most 32 half-edges and 32 buses.  Every
  // contribution gets a destination in the tile's J (H, G) space: the first contribution of a slot goes to the slot's
  // own position in the output image, later ones (diagonal block, parallel branches, every further term of a residual)
  // to extra cells behind the image, contiguous per slot; the kernel adds them left to right in emission order.
  P.wtiles.clear();
  if (has_bus_rows(form) && nb > 0) {
    const int NJ = he_nj(form), NH = he_nh(form), NG = he_npg(form);
    bool ok = true;
    // CTA tiles holding a bus that does not fit a warp stay with bus_kernel; skip[i] = end of the hub tile that starts at i
    P.hub_tiles.clear();
    std::vector<int64_t> skip(nb + 1, -1);
    for (int32_t ti = 0; ti < (int32_t)P.tiles.size(); ti++) {
      const Tile& ct = P.tiles[ti];
      if (ct.kind != 0) continue;
      bool hub = false;
      for (int64_t i = ct.u0; i < ct.u1; i++) hub = hub || P.he_ptr[i + 1] - P.he_ptr[i] > 32;
      if (hub) { P.hub_tiles.push_back(ti); skip[ct.u0] = ct.u1; }
    }
    P.wt_ownj_ptr.assign(nb + 1, 0);
    P.wt_ownh_ptr.assign(nb + 1, 0);
    P.wt_owng_ptr.assign(nb + 1, 0);
    auto own_gap = [&](int64_t a, int64_t b) {       // buses a .. b belong to a hub tile: no own lists
      for (int64_t i = a; i < b; i++) {
        P.wt_ownj_ptr[i + 1] = (int32_t)P.wt_ownj.size();
        P.wt_ownh_ptr[i + 1] = (int32_t)P.wt_ownh.size();
        P.wt_owng_ptr[i + 1] = (int32_t)P.wt_owng.size();
      }
    };
    int64_t u0 = 0;
    while (ok && u0 < nb) {
      if (skip[u0] >= 0) { own_gap(u0, skip[u0]); u0 = skip[u0]; continue; }
      int64_t u1 = u0 + 1;
      while (u1 < nb && skip[u1] < 0 && u1 - u0 < 32 && P.he_ptr[u1 + 1] - P.he_ptr[u0] <= 32) u1++;
      WTile t{};
      t.u0 = (int32_t)u0; t.u1 = (int32_t)u1;
      t.j0 = P.jptr[2 * u0]; t.jn = P.jptr[2 * u1] - t.j0;
      t.h0 = P.hptr[2 * u0]; t.hn = P.hptr[2 * u1] - t.h0;
      t.he0 = P.he_ptr[u0]; t.nhe = P.he_ptr[u1] - t.he0;
      const int32_t cj0 = P.bus_jpp[u0], ncj = P.bus_jpp[u1] - cj0, ch0 = P.bus_hpp[u0], nch = P.bus_hpp[u1] - ch0;
      // residual terms of the tile in emission order, and the row (slot) each one belongs to
      std::vector<int32_t> gslot, gbus0(u1 - u0 + 1, 0);
      for (int64_t i = u0; i < u1; i++) {
        const int32_t r0 = (int32_t)(2 * (i - u0));
        const int32_t ngen = net.gen_ptr[i + 1] - net.gen_ptr[i], deg = P.he_ptr[i + 1] - P.he_ptr[i];
        const int32_t nshi = facts ? net.sh_ptr[i + 1] - net.sh_ptr[i] : 0;
        for (int32_t q = 0; q < ngen; q++) { gslot.push_back(r0); gslot.push_back(r0 + 1); }
        for (int32_t lh = 0; lh < deg; lh++) for (int k = 0; k < NG; k++) gslot.push_back(r0 + piece_row(form, k));
        gslot.push_back(r0); gslot.push_back(r0 + 1);                   // shunt terms
        for (int32_t q = 0; q < nshi; q++) gslot.push_back(r0 + 1);     // switched shunts (Q row)
        gslot.push_back(r0); gslot.push_back(r0 + 1);                   // - Pd, - Qd
        gbus0[i - u0 + 1] = (int32_t)gslot.size();
      }
      const int32_t ncg = (int32_t)gslot.size(), gn = (int32_t)(2 * (u1 - u0));
      if (ncj + 2 >= 65536 || nch + 2 >= 65536 || ncg + 2 >= 65536) { ok = false; break; }
      std::vector<int32_t> cdj(ncj, 0), cdh(nch, 0), cdg(ncg, 0);
      // slot -> contributions (tile-local ids, ascending = emission order)
      std::vector<std::vector<int32_t>> lj(t.jn), lh_(t.hn), lg(gn);
      for (int32_t e = 0; e < t.jn; e++) for (int32_t q = P.jsp[t.j0 + e]; q < P.jsp[t.j0 + e + 1]; q++) lj[e].push_back(P.jperm[q] - cj0);
      // structural zeros of the half-edge Hessian terms are dropped from the sums (x + 0.0 == x); a slot made of zeros
      // only keeps one of them: the kernel zeroes the H space once and never writes such a contribution
      std::vector<char> zh(nch, 0);
      for (int64_t i = u0; i < u1; i++) {
        const int32_t cH = P.bus_hpp[i] - ch0, deg = P.he_ptr[i + 1] - P.he_ptr[i];
        for (int32_t lh = 0; lh < deg; lh++) for (int ty = 0; ty < NH; ty++) zh[cH + lh * NH + ty] = he_h_zero(form, ty);
      }
      for (int32_t e = 0; e < t.hn; e++) {
        for (int32_t q = P.hsp[t.h0 + e]; q < P.hsp[t.h0 + e + 1]; q++) if (!zh[P.hperm[q] - ch0]) lh_[e].push_back(P.hperm[q] - ch0);
        if (lh_[e].empty()) lh_[e].push_back(P.hperm[P.hsp[t.h0 + e]] - ch0);
      }
      for (int32_t c = 0; c < ncg; c++) lg[gslot[c]].push_back(c);
      struct MS { int32_t kind, e, x0, cnt; };
      std::vector<MS> ms;
      auto spaces = [&](int kind, const std::vector<std::vector<int32_t>>& sl, std::vector<int32_t>& cd, int32_t& nx) {
        const int32_t en = (int32_t)sl.size();
        nx = 0;
        for (int32_t e = 0; e < en; e++) {
          cd[sl[e][0]] = e;
          if (sl[e].size() > 1) ms.push_back(MS{kind, e, en + nx, (int32_t)sl[e].size() - 1});
          for (size_t q = 1; q < sl[e].size(); q++) cd[sl[e][q]] = en + nx++;
        }
      };
      spaces(0, lj, cdj, t.jx);
      spaces(1, lh_, cdh, t.hx);
      spaces(2, lg, cdg, t.gx);
      if (ncj + 2 >= 16384 || nch + 2 >= 16384 || ncg + 2 >= 16384) { ok = false; break; }
      for (const MS& m : ms) if (m.cnt >= 4096) ok = false;           // the kernel's decoded descriptors keep 12 bits of count
      if (!ok) break;
      std::stable_sort(ms.begin(), ms.end(), [](const MS& a, const MS& b) { return a.cnt > b.cnt; });
      t.m0 = (int32_t)P.wt_multi.size(); t.mn = (int32_t)ms.size();
      for (const MS& m : ms)
        P.wt_multi.push_back((uint64_t)m.e | ((uint64_t)m.kind << 14) | ((uint64_t)m.x0 << 16) | ((uint64_t)m.cnt << 32));
      t.dst0 = (int32_t)P.wt_dst.size();
      P.wt_dst.resize(P.wt_dst.size() + (size_t)(NJ + NH + NG) * 32, 0);
      uint16_t* dst = P.wt_dst.data() + t.dst0;
      for (int64_t i = u0; i < u1; i++) {
        const int32_t ngen = net.gen_ptr[i + 1] - net.gen_ptr[i], ng2 = 2 * ngen, deg = P.he_ptr[i + 1] - P.he_ptr[i];
        const int32_t cJ = P.bus_jpp[i] - cj0, cH = P.bus_hpp[i] - ch0, cG = gbus0[i - u0], lane0 = P.he_ptr[i] - t.he0;
        for (int32_t lh = 0; lh < deg; lh++) {
          for (int ty = 0; ty < NJ; ty++) dst[ty * 32 + lane0 + lh] = (uint16_t)cdj[cJ + ng2 + lh * NJ + ty];
          for (int ty = 0; ty < NH; ty++) dst[(NJ + ty) * 32 + lane0 + lh] = (uint16_t)cdh[cH + lh * NH + ty];
          for (int ty = 0; ty < NG; ty++) dst[(NJ + NH + ty) * 32 + lane0 + lh] = (uint16_t)cdg[cG + ng2 + lh * NG + ty];
        }
        for (int32_t c = cJ; c < cJ + ng2; c++) P.wt_ownj.push_back((uint16_t)cdj[c]);
        for (int32_t c = cJ + ng2 + NJ * deg; c < P.bus_jpp[i + 1] - cj0; c++) P.wt_ownj.push_back((uint16_t)cdj[c]);
        for (int32_t c = cH + NH * deg; c < P.bus_hpp[i + 1] - ch0; c++) P.wt_ownh.push_back((uint16_t)cdh[c]);
        for (int32_t c = cG; c < cG + ng2; c++) P.wt_owng.push_back((uint16_t)cdg[c]);
        for (int32_t c = cG + ng2 + NG * deg; c < gbus0[i - u0 + 1]; c++) P.wt_owng.push_back((uint16_t)cdg[c]);
        P.wt_ownj_ptr[i + 1] = (int32_t)P.wt_ownj.size();
        P.wt_ownh_ptr[i + 1] = (int32_t)P.wt_ownh.size();
        P.wt_owng_ptr[i + 1] = (int32_t)P.wt_owng.size();
      }
      const int32_t lists = 4 * t.mn + (P.wt_ownj_ptr[u1] - P.wt_ownj_ptr[u0]) +
                            (P.wt_ownh_ptr[u1] - P.wt_ownh_ptr[u0]) + (P.wt_owng_ptr[u1] - P.wt_owng_ptr[u0]);
      P.wt_lists_max = std::max(P.wt_lists_max, lists);
      P.wt_jspace = std::max(P.wt_jspace, (ncj + 3) & ~1);          // + 1 for the 16-byte phase of the image, even
      P.wt_hspace = std::max(P.wt_hspace, (nch + 3) & ~1);
      P.wt_gspace = std::max(P.wt_gspace, (ncg + 3) & ~1);
      P.wtiles.push_back(t);
      u0 = u1;
    }
    if (!ok) { P.wtiles.clear(); P.wt_dst.clear(); P.wt_multi.clear(); P.hub_tiles.clear(); }
  }
  P.tile_units_bus = bus_tpb;
  P.n_bus_tiles = (int)P.tiles.size();
  P.NRP = pat.nrows | 1; P.RJP = P.RJ | 1; P.RHP = P.RH | 1;
  const int64_t mxb = nr > 0 ? (int64_t)8 * tpb * (P.NRP + P.RJP + P.RHP) : 0;
  for (int64_t u0 = 0; u0 < nr; u0 += tpb) {
    const int64_t u1 = std::min<int64_t>(nr, u0 + tpb);
    Tile t;
    t.kind = 1; t.u0 = (int32_t)u0; t.u1 = (int32_t)u1;
    t.g0 = (int32_t)(P.bus_rows + u0 * pat.nrows); t.gn = (int32_t)((u1 - u0) * pat.nrows);
    t.j0 = P.jptr[P.bus_rows] + (int32_t)(u0 * P.RJ); t.jn = (int32_t)((u1 - u0) * P.RJ);
    t.h0 = P.hptr[P.bus_rows] + (int32_t)(u0 * P.RH); t.hn = (int32_t)((u1 - u0) * P.RH);
    P.tiles.push_back(t);
  }
  {
    Tile t{};                                      // objective scalar (formulations.jl:122-138)
    t.kind = 2; t.u0 = 0; t.u1 = (int32_t)ng;
    P.tiles.push_back(t);
  }
  P.smem_bytes = std::max<int64_t>(std::max(mx, mxb), 1024);
  P.smem_bus_bytes = std::max<int64_t>(mx, 1024);
  P.smem_branch_bytes = std::max<int64_t>(mxb, 1024);
  P.smem_doubles = (P.smem_bytes + 7) / 8;
  return "";
}

}  // namespace ps
