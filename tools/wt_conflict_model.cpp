// Development tool: bank-conflict model of bus_warp_kernel's half-edge stores, computed from the real plan tables
// (8-byte shared-memory stores: 16 bank pairs, one wavefront per half-warp when conflict-free).
//   g++ -O2 -std=c++17 -Ipowersense.jl_b200/csrc -Iinclude tools/wt_conflict_model.cpp powersense.jl_b200/csrc/ps_plan.cpp -o /tmp/wtc && /tmp/wtc 5
// On a case2869pegase-sized random network (PNPAPV = form 5) it gives 5.57 wavefronts per store -- ncu measures 5.5
// (profiles/r02_ncu_config3_summary.txt) -- 4.25 if the extra cells of multi-contribution slots were laid out conflict-free,
// 3.52 with per-bus padding of the J / H images on top (one bulk copy per bus instead of one per tile); ideal 2.  See DESIGN.md section 7.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <algorithm>
#include <map>
#include "ps_plan.h"
using namespace ps;
static HostNet make_net(int nb, int nr, int ng, unsigned seed) {
  std::mt19937 rng(seed);
  HostNet n;
  n.nbus = nb; n.nbr = nr; n.ngen = ng; n.nbss = 0;
  for (int k = 0; k < nr; k++) {
    int f, t;
    if (k < nb - 1) { f = k + 1; t = (int)(rng() % (k + 1)); }
    else { f = (int)(rng() % nb); t = (int)(rng() % nb); if (f == t) t = (f + 1) % nb; }
    if (rng() & 1) std::swap(f, t);
    n.f.push_back(f); n.t.push_back(t);
  }
  // random permutation of the branch list (synth.py default)
  std::vector<int> perm(nr); for (int i = 0; i < nr; i++) perm[i] = i; std::shuffle(perm.begin(), perm.end(), rng);
  std::vector<int32_t> f2(nr), t2(nr); for (int i = 0; i < nr; i++) { f2[i] = n.f[perm[i]]; t2[i] = n.t[perm[i]]; } n.f = f2; n.t = t2;
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
static int wavefronts(const std::vector<int>& addr) {   // addr[lane] or -1
  int w = 0;
  for (int half = 0; half < 2; half++) {
    std::map<int, std::vector<int>> bank;
    for (int l = 16 * half; l < 16 * half + 16; l++) if (addr[l] >= 0) { auto& v = bank[addr[l] & 15]; if (std::find(v.begin(), v.end(), addr[l]) == v.end()) v.push_back(addr[l]); }
    int mx = 0; bool any = false;
    for (auto& kv : bank) { mx = std::max<int>(mx, kv.second.size()); any = true; }
    w += any ? mx : 0;
  }
  return w;
}
int main(int argc, char** argv) {
  const int form = argc > 1 ? atoi(argv[1]) : PNPAPV;
  HostNet n = make_net(2869, 4582, 510, 5);
  Plan P;
  std::string err = build_plan(n, form, 0, 0, 128, 128, P);
  if (!err.empty()) { printf("plan: %s\n", err.c_str()); return 1; }
  const int NJ = he_nj(form), NH = he_nh(form), NG = he_npg(form), NT = NJ + NH + NG;
  long long stores = 0, wf = 0, ideal = 0, wf_fixed = 0, st_fixed = 0, wf_extra = 0;
  std::vector<long long> per_ty_w(NT, 0), per_ty_n(NT, 0);
  for (const WTile& t : P.wtiles) {
    const int enJ = t.jn, enH = t.hn, enG = 2 * (t.u1 - t.u0);
    for (int ty = 0; ty < NT; ty++) {
      if (ty >= NJ && ty < NJ + NH && he_h_zero(form, ty - NJ)) continue;
      std::vector<int> addr(32, -1), addr_img(32, -1);
      const int en = ty < NJ ? enJ : (ty < NJ + NH ? enH : enG);
      int nact = 0;
      for (int l = 0; l < t.nhe; l++) { const int d = P.wt_dst[t.dst0 + ty * 32 + l]; addr[l] = d; nact++; if (d < en) addr_img[l] = d; }
      const int w = wavefronts(addr);
      stores++; wf += w; ideal += (t.nhe > 16 ? 2 : 1);
      per_ty_w[ty] += w; per_ty_n[ty]++;
      // model: extras made conflict-free (each extra lane in its own bank pair, not colliding with anything): only the image-bound lanes conflict
      wf_fixed += std::max(wavefronts(addr_img), t.nhe > 16 ? 2 : 1);
    }
  }
  // model 2: per-bus padding of the J and H images (even pads, greedy per tile), extras conflict-free
  long long wf_pad = 0; long long pad_cells = 0;
  for (const WTile& t : P.wtiles) {
    for (int space = 0; space < 2; space++) {
      const int ty0 = space == 0 ? 0 : NJ, ty1 = space == 0 ? NJ : NJ + NH;
      const int en = space == 0 ? t.jn : t.hn;
      const std::vector<int32_t>& ptr = space == 0 ? P.jptr : P.hptr;
      const int e0 = space == 0 ? t.j0 : t.h0;
      const int nbt = t.u1 - t.u0;
      std::vector<int> shift(nbt, 0);
      // lanes of bus i: he_ptr
      auto cost = [&](int upto) {   // total wavefronts over the space's store instructions counting buses [0, upto]
        long long c = 0;
        for (int ty = ty0; ty < ty1; ty++) {
          if (space == 1 && he_h_zero(form, ty - NJ)) continue;
          std::vector<int> addr(32, -1);
          for (int i = 0; i <= upto; i++)
            for (int l = P.he_ptr[t.u0 + i] - t.he0; l < P.he_ptr[t.u0 + i + 1] - t.he0; l++) {
              const int d = P.wt_dst[t.dst0 + ty * 32 + l];
              if (d < en) addr[l] = d + shift[i];
            }
          c += wavefronts(addr);
        }
        return c;
      };
      int cum = 0;
      for (int i = 0; i < nbt; i++) {
        long long best = -1; int bp = 0;
        for (int pad = 0; pad < 16; pad += 2) { shift[i] = cum + pad; const long long c = cost(i); if (best < 0 || c < best) { best = c; bp = pad; } }
        cum += bp; shift[i] = cum; pad_cells += bp;
      }
      // final cost, at least the ideal per instruction
      for (int ty = ty0; ty < ty1; ty++) {
        if (space == 1 && he_h_zero(form, ty - NJ)) continue;
        std::vector<int> addr(32, -1);
        for (int i = 0; i < nbt; i++)
          for (int l = P.he_ptr[t.u0 + i] - t.he0; l < P.he_ptr[t.u0 + i + 1] - t.he0; l++) {
            const int d = P.wt_dst[t.dst0 + ty * 32 + l];
            if (d < en) addr[l] = d + shift[i];
          }
        wf_pad += std::max(wavefronts(addr), t.nhe > 16 ? 2 : 1);
      }
    }
  }
  long long nJH = 0; for (int ty = 0; ty < NJ + NH; ty++) nJH += per_ty_n[ty];
  printf("per-bus padded images + conflict-free extras: %.2f wavefronts per J/H store; pad cells per tile %.1f\n", (double)wf_pad / nJH, (double)pad_cells / P.wtiles.size());
  printf("form %d: %zu tiles, %lld store instructions, %.2f wavefronts per store (ideal %.2f); with conflict-free extras %.2f\n", form, P.wtiles.size(), stores,
         (double)wf / stores, (double)ideal / stores, (double)wf_fixed / stores);
  for (int ty = 0; ty < NT; ty++) if (per_ty_n[ty]) printf("  ty %2d (%s): %.2f\n", ty, ty < NJ ? "J" : (ty < NJ + NH ? "H" : "G"), (double)per_ty_w[ty] / per_ty_n[ty]);
  return 0;
}
