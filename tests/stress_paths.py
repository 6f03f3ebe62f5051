"""Randomised cross-check of the kernel paths (test infrastructure, needs a GPU; test_gpu_stress.py runs a short pass,
`python tests/stress_paths.py 200` the long one): many small seeded networks -- parallel
branches, out-of-service elements, switched shunts, both source types, hubs, isolated buses -- every formulation, random
layouts; the default path, the warp-tile path (PS_NO_CBFAST / PS_NO_PBFAST) and the CTA-tile path (PS_NO_WARPBUS) must
agree bitwise, with and without a status mask, the residual-only and g + J calls must give the bits of the full evaluation,
and the default path must match the CPU oracle.
usage: python tests/stress_paths.py [n_networks]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import powersense_b200 as ps  # noqa: E402
from powersense_b200 import capi, synth  # noqa: E402
from oracle import c_oracle  # noqa: E402
from cases import tol_err_scaled  # noqa: E402


def main(n):
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(int(os.environ.get("PS_STRESS_SEED", "2024")))
    bad = 0
    for it in range(n):
        nb = int(rng.integers(3, 260))
        nr = int(nb - 1 + rng.integers(0, 2 * nb))
        ng = int(rng.integers(1, max(2, nb // 3)))
        src = "arpae" if rng.random() < 0.4 else "matpower"
        nbss = int(rng.integers(0, 6)) if src == "arpae" else 0
        net = synth.synth_network(nb, nr, ng, seed=1000 + it, source_type=src, nbss=nbss, frac_out=float(rng.choice([0.0, 0.1])),
                                  frac_parallel=float(rng.choice([0.0, 0.05, 0.2])))
        if rng.random() < 0.25 and nr > 40:                      # a hub bus
            k = int(rng.integers(33, min(nr, 70)))
            net.f[:k] = 1
            net.t[:k] = 2 + (np.arange(k) % (nb - 1))
        d = ps.create_PowersenseData(net)
        S = int(rng.integers(1, 12))
        for F in ps.ALL_FORMULATIONS:
            x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=it)
            h = capi.Handle(d, F.code, device=0)
            mu = rng.normal(size=(S, h.m_nl))
            tx, tmu = torch.from_numpy(x).to(dev), torch.from_numpy(mu).to(dev)
            st = (rng.random((S, max(d.nbr, 1))) > 0.1).astype(np.uint8)[:, :d.nbr]
            for masked in (False, True):
                b = capi.Batch(h, S)
                b.set_loads(Pd, Qd)
                if masked and d.nbr:
                    b.set_branch_status(st)
                pads = [int(v) for v in rng.integers(0, 4, 3)]
                odd = bool(rng.random() < 0.3)

                def run(what=7):
                    def rows(width, pad):
                        ld = width + pad + ((width + pad + 1) % 2 if odd else (width + pad) % 2)
                        flat = torch.full((pad + S * max(ld, 1),), np.nan, dtype=torch.float64, device=dev)
                        return flat[pad:].view(S, max(ld, 1))[:, :width]
                    g, J, H = rows(h.m_nl, pads[0]), rows(h.nnzJ, pads[1]), rows(h.nnzH, pads[2])
                    sc = torch.empty(S, capi.PS_NSCALARS, dtype=torch.float64, device=dev)
                    b.eval_torch(what, tx, tmu, g, J, H, sc)
                    torch.cuda.synchronize()
                    return [a.cpu().numpy() for a in (g, J, H, sc)]
                ref = run()
                # the callbacks a solver issues separately -- residuals alone (thread-per-bus kernel, residual-only branch
                # kernel), g + J (their own instantiations) -- give the same bits as the full evaluation
                for what, names in ((1, "gs"), (3, "gJs")):
                    out = run(what)
                    for name, a, c in zip("gJHs", ref, out):
                        if name in names and not np.array_equal(a, c):
                            bad += 1
                            print(f"MISMATCH net {it} (nb {nb} nr {nr} {src} nbss {nbss}) {F.name} masked={masked} what={what} {name}", flush=True)
                for env in ({"PS_NO_CBFAST": "1", "PS_NO_PBFAST": "1"}, {"PS_NO_CBFAST": "1", "PS_NO_PBFAST": "1", "PS_NO_WARPBUS": "1"},
                            {"PS_NO_BULK": "1", "PS_NO_DIRECT": "1"}):
                    os.environ.update(env)
                    try:
                        out = run()
                    finally:
                        for k in env:
                            del os.environ[k]
                    for name, a, c in zip("gJHs", ref, out):
                        if not np.array_equal(a, c):
                            bad += 1
                            print(f"MISMATCH net {it} (nb {nb} nr {nr} {src} nbss {nbss}) {F.name} masked={masked} {env} {name}", flush=True)
                if not masked:
                    o = c_oracle.COracle(d, F.code)
                    og, oJ, oH = o.eval_batch(x, mu, Pd, Qd)
                    for s in range(S):
                        gs, Js, Hs = o.condition_scale(x[s], mu[s], Pd[s], Qd[s])
                        for name, a, r, scl in (("g", ref[0][s], og[s], gs), ("J", ref[1][s], oJ[s], Js), ("H", ref[2][s], oH[s], Hs)):
                            if tol_err_scaled(a, r, scl) > 1.0:
                                bad += 1
                                print(f"ORACLE net {it} {F.name} s={s} {name} err {tol_err_scaled(a, r, scl):.3g}", flush=True)
                b.close()
            h.close()
        if it % 5 == 4:
            print(f"{it + 1} networks, {bad} problems", flush=True)
    print("stress:", "FAILED" if bad else "ok", f"({n} networks x {len(ps.ALL_FORMULATIONS)} formulations)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 20))
