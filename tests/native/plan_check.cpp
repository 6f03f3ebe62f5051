// CPU check of the warp-tile tables of ps_plan.cpp (test infrastructure): replays bus_warp_kernel's data movement --
// contribution -> destination (register table / own lists), extra cells, multi-slot lists -- with random contribution
// values and compares every output slot with the left-to-right sum over the slot's contributions (jsp / jperm).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "ps_plan.h"

using namespace ps;

static HostNet make_net(int nb, int nr, int ng, unsigned seed, bool hub) {
  std::mt19937 rng(seed);
  HostNet n;
  n.nbus = nb; n.nbr = nr; n.ngen = ng; n.nbss = 0;
  for (int k = 0; k < nr; k++) {
    int f, t;
    if (k < nb - 1) { f = k + 1; t = (int)(rng() % (k + 1)); }           // spanning tree
    else if (hub && k % 3 == 0) { f = 0; t = 1 + (int)(rng() % (nb - 1)); }
    else if (k % 17 == 0 && k > 0) { f = n.f[k - 1]; t = n.t[k - 1]; }    // parallel branch
    else { f = (int)(rng() % nb); t = (int)(rng() % nb); if (f == t) t = (f + 1) % nb; }
    if (rng() & 1) std::swap(f, t);
    n.f.push_back(f); n.t.push_back(t);
  }
  auto fill = [&](std::vector<double>& v, size_t m) { v.resize(m); for (auto& x : v) x = 0.1 + (rng() % 1000) / 500.0; };
  fill(n.Gf, nr); fill(n.Bf, nr); fill(n.Gt, nr); fill(n.Bt, nr); fill(n.gf, nr); fill(n.bf, nr); fill(n.gt, nr); fill(n.bt, nr);
  fill(n.Imax, nr); fill(n.Pd, nb); fill(n.Qd, nb); fill(n.Gl, nb); fill(n.Bl, nb);
  std::vector<std::vector<int>> gen(nb), sij(nb), sji(nb);
  for (int g = 0; g < ng; g++) gen[rng() % nb].push_back(g);
  for (int k = 0; k < nr; k++) { sij[n.f[k]].push_back(k); sji[n.t[k]].push_back(k); }
  auto csr = [&](const std::vector<std::vector<int>>& l, std::vector<int32_t>& ptr, std::vector<int32_t>& idx) {
    ptr.assign(1, 0);
    for (auto& v : l) { for (int e : v) idx.push_back(e); ptr.push_back((int32_t)idx.size()); }
  };
  csr(gen, n.gen_ptr, n.gen_idx); csr(sij, n.sij_ptr, n.sij_idx); csr(sji, n.sji_ptr, n.sji_idx);
  n.sh_ptr.assign(nb + 1, 0);
  return n;
}

// kind: 0 = J, 1 = H, 2 = G (residual rows as flat term sums)
static int check(const Plan& P, int kind) {
  const int NJ = he_nj(P.form), NH = he_nh(P.form), NG = he_npg(P.form);
  const bool facts = (P.flags & FLAG_FACTS) != 0;
  const std::vector<int32_t>& optr = kind == 0 ? P.wt_ownj_ptr : (kind == 1 ? P.wt_ownh_ptr : P.wt_owng_ptr);
  const std::vector<uint16_t>& own = kind == 0 ? P.wt_ownj : (kind == 1 ? P.wt_ownh : P.wt_owng);
  const int per = kind == 0 ? NJ : (kind == 1 ? NH : NG), ty0 = kind == 0 ? 0 : (kind == 1 ? NJ : NJ + NH);
  std::mt19937 rng(7 + kind);
  int64_t covered = 0;
  for (const WTile& t : P.wtiles) {
    if (t.nhe > 32 || t.u1 - t.u0 > 32) { printf("tile too large\n"); return 1; }
    const int en = kind == 0 ? t.jn : (kind == 1 ? t.hn : 2 * (t.u1 - t.u0)), nx = kind == 0 ? t.jx : (kind == 1 ? t.hx : t.gx);
    // the tile's contributions in emission order with their slot: from the plan's slot maps (J, H) or restated (G)
    std::vector<int> cslot;
    if (kind < 2) {
      const std::vector<int32_t>& pp = kind == 0 ? P.bus_jpp : P.bus_hpp;
      const std::vector<int32_t>& sp = kind == 0 ? P.jsp : P.hsp;
      const std::vector<int32_t>& perm = kind == 0 ? P.jperm : P.hperm;
      const int e0 = kind == 0 ? t.j0 : t.h0;
      cslot.assign(pp[t.u1] - pp[t.u0], -1);
      for (int e = e0; e < e0 + en; e++) for (int q = sp[e]; q < sp[e + 1]; q++) cslot[perm[q] - pp[t.u0]] = e - e0;
    } else {
      for (int i = t.u0; i < t.u1; i++) {
        const int r0 = 2 * (i - t.u0), ngen = P.net.gen_ptr[i + 1] - P.net.gen_ptr[i], deg = P.he_ptr[i + 1] - P.he_ptr[i];
        const int nsh = facts ? P.net.sh_ptr[i + 1] - P.net.sh_ptr[i] : 0;
        for (int q = 0; q < ngen; q++) { cslot.push_back(r0); cslot.push_back(r0 + 1); }
        for (int lh = 0; lh < deg; lh++) for (int k = 0; k < NG; k++) cslot.push_back(r0 + piece_row(P.form, k));
        cslot.push_back(r0); cslot.push_back(r0 + 1);
        for (int q = 0; q < nsh; q++) cslot.push_back(r0 + 1);
        cslot.push_back(r0); cslot.push_back(r0 + 1);
      }
    }
    const int nc = (int)cslot.size();
    // structural zeros of the half-edge Hessian terms: value 0.0, never written by the kernel
    std::vector<char> zero(nc, 0);
    if (kind == 1) {
      for (int i = t.u0; i < t.u1; i++) {
        const int cH = P.bus_hpp[i] - P.bus_hpp[t.u0], deg = P.he_ptr[i + 1] - P.he_ptr[i];
        for (int lh = 0; lh < deg; lh++) for (int ty = 0; ty < NH; ty++) zero[cH + lh * NH + ty] = he_h_zero(P.form, ty);
      }
    }
    if (en + nx > nc || (kind != 1 && en + nx != nc)) { printf("space size mismatch\n"); return 1; }
    if (en + nx + 2 > (kind == 0 ? P.wt_jspace : (kind == 1 ? P.wt_hspace : P.wt_gspace))) { printf("space max too small\n"); return 1; }
    std::vector<double> val(nc), sp_(en + nx, kind == 1 ? 0.0 : -1e300), want(en, 0.0);
    std::vector<char> hit(en + nx, 0), seen(en, 0);
    for (int c = 0; c < nc; c++) {
      val[c] = zero[c] ? 0.0 : (double)(rng() % 100000) / 7.0 - 5000.0;
      if (cslot[c] < 0 || cslot[c] >= en) { printf("contribution without slot\n"); return 1; }
      want[cslot[c]] = seen[cslot[c]] ? want[cslot[c]] + val[c] : val[c];
      seen[cslot[c]] = 1;
    }
    int cc = 0;                                           // running contribution index of the replay below
    auto put = [&](int dst, double v) {
      if (zero[cc]) return;                               // RegW::zero(): skipped
      if (dst < 0 || dst >= en + nx || hit[dst]) { printf("bad destination %d\n", dst); exit(1); }
      hit[dst] = 1; sp_[dst] = v;
    };
    int c = 0;
    for (int i = t.u0; i < t.u1; i++) {
      const int ng2 = kind == 1 ? 0 : 2 * (P.net.gen_ptr[i + 1] - P.net.gen_ptr[i]), deg = P.he_ptr[i + 1] - P.he_ptr[i];
      int o = optr[i];
      for (int q = 0; q < ng2; q++) { cc = c; put(own[o++], val[c++]); }
      for (int lh = 0; lh < deg; lh++)
        for (int ty = 0; ty < per; ty++) { cc = c; put(P.wt_dst[t.dst0 + (ty0 + ty) * 32 + (P.he_ptr[i] - t.he0) + lh], val[c++]); }
      while (o < optr[i + 1]) { cc = c; put(own[o++], val[c++]); }
    }
    if (c != nc) { printf("contribution count mismatch %d != %d (kind %d)\n", c, nc, kind); return 1; }
    if (kind != 1) for (char h : hit) if (!h) { printf("unwritten cell\n"); return 1; }
    for (int m = 0; m < t.mn; m++) {
      const uint64_t w = P.wt_multi[t.m0 + m];
      if (m > 0 && ((w >> 32) & 0xffff) > ((P.wt_multi[t.m0 + m - 1] >> 32) & 0xffff)) { printf("multi-slots not sorted by count\n"); return 1; }
      if ((int)((w >> 14) & 3) != kind) continue;
      const int d = (int)(w & 0x3fff), x0 = (int)((w >> 16) & 0xffff), cnt = (int)((w >> 32) & 0xffff);
      double v = sp_[d];
      for (int q = 0; q < cnt; q++) v = v + sp_[x0 + q];
      sp_[d] = v;
    }
    for (int e = 0; e < en; e++) {
      if (!seen[e] || want[e] != sp_[e]) { printf("slot %d differs (kind %d)\n", e, kind); return 1; }
      covered++;
    }
  }
  int64_t want_n = kind == 0 ? P.jptr[P.bus_rows] : (kind == 1 ? P.hptr[P.bus_rows] : P.bus_rows);
  for (int32_t ti : P.hub_tiles) {                       // CTA tiles that stay with bus_kernel
    const Tile& ct = P.tiles[ti];
    want_n -= kind == 0 ? ct.jn : (kind == 1 ? ct.hn : 2 * (ct.u1 - ct.u0));
    for (const WTile& t : P.wtiles) if (t.u0 < ct.u1 && ct.u0 < t.u1) { printf("warp tile overlaps a hub tile\n"); return 1; }
  }
  if (!P.wtiles.empty() && covered != want_n) { printf("coverage %lld != %lld\n", (long long)covered, (long long)want_n); return 1; }
  return 0;
}

int main() {
  int bad = 0, with_tiles = 0;
  for (int form = 0; form < NFORM; form++) {
    if (!has_bus_rows(form)) continue;
    for (int cfg = 0; cfg < 4; cfg++) {
      const int nb = cfg == 0 ? 5 : (cfg == 1 ? 200 : (cfg == 2 ? 1500 : 60));
      const int nr = cfg == 0 ? 6 : (cfg == 1 ? 330 : (cfg == 2 ? 2300 : 200));
      HostNet n = make_net(nb, nr, nb / 3 + 1, 11 + cfg, cfg == 3);
      Plan P;
      const std::string msg = build_plan(n, form, SRC_MATPOWER, 0, 128, 128, P);
      if (!msg.empty()) { printf("form %d cfg %d: %s\n", form, cfg, msg.c_str()); bad++; continue; }
      if (cfg == 3 && P.hub_tiles.empty()) {            // a hub bus with more than 32 half-edges stays with the CTA-tile kernel
        printf("form %d: hub network without hub tiles\n", form); bad++; continue;
      }
      if (cfg != 3 && !P.hub_tiles.empty()) { printf("form %d cfg %d: unexpected hub tiles\n", form, cfg); bad++; continue; }
      if (P.wtiles.empty()) {                            // fine only when the hub tiles hold every bus
        int64_t held = 0;
        for (int32_t ti : P.hub_tiles) held += P.tiles[ti].u1 - P.tiles[ti].u0;
        if (held != nb) { printf("form %d cfg %d: no warp tiles\n", form, cfg); bad++; }
        continue;
      }
      with_tiles++;
      for (int kind = 0; kind < 3; kind++) bad += check(P, kind);
    }
  }
  printf("%s (%d plans with warp tiles)\n", bad ? "FAILED" : "ok", with_tiles);
  return bad ? 1 : 0;
}
