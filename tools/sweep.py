"""Development sweep (not part of the bench contract): times ps_batch_eval for several cases /
formulations / scenario-chunk sizes on one GPU and prints achieved algorithmic GB/s."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import powersense_b200 as ps  # noqa: E402
from powersense_b200 import capi, synth  # noqa: E402


def run(case, Fname, S, chunks, what=7, steps=5, scal=True):
    d = synth.named_case(case)
    F = ps.BY_NAME[Fname]
    h = capi.Handle(d, F.code, device=0)
    b = capi.Batch(h, S)
    dev = torch.device("cuda:0")
    xb, Pd, Qd, _ = synth.scenario_batch(d, F, 8, seed=1)
    x = torch.from_numpy(xb).to(dev).repeat(-(-S // 8), 1)[:S].contiguous()
    x *= 1.0 + 1e-3 * (torch.rand_like(x) - 0.5)
    mu = torch.randn(S, h.m_nl, dtype=torch.float64, device=dev)
    (ldg, ldJ, ldH), (pg, pJ, pH) = h.preferred_layout()

    def rows(width, ld, pad):
        if os.environ.get("PS_PLAIN_LAYOUT"):
            return torch.empty(S, width, dtype=torch.float64, device=dev)
        flat = torch.empty(pad + S * ld, dtype=torch.float64, device=dev)
        return flat[pad:].view(S, ld)[:, :width]
    g, J, H = rows(h.m_nl, ldg, pg), rows(h.nnzJ, ldJ, pJ), rows(h.nnzH, ldH, pH)
    sc = torch.empty(S, 4, dtype=torch.float64, device=dev) if scal else None
    ab = h.algorithmic_bytes(what)
    for c in chunks:
        os.environ["PS_CHUNK"] = str(c)
        for _ in range(2):
            b.eval_torch(what, x, mu, g, J, H, sc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            b.eval_torch(what, x, mu, g, J, H, sc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"{case} {Fname} S={S} what={what} scal={scal} chunk={c}: {ms:.3f} ms  {S / ms * 1e3:.0f} evals/s  {ab * S / ms / 1e6:.0f} GB/s", flush=True)
    b.enable_timing(True)
    b.eval_torch(what, x, mu, g, J, H, sc)
    kt = b.kernel_times()
    abr = h.algorithmic_bytes_branch(what)
    print(f"   kernels (ms): {kt}; branch rows {abr * S / max(kt['branch'], 1e-9) / 1e6:.0f} GB/s, "
          f"bus rows {(ab - abr) * S / max(kt['bus'] + kt['bus_jh'], 1e-9) / 1e6:.0f} GB/s", flush=True)
    b.enable_timing(False)
    b.close()
    h.close()


if __name__ == "__main__":
    sel = sys.argv[1] if len(sys.argv) > 1 else "all"
    if sel in ("all", "pbrarv"):
        run("case13659pegase", "PBRARVmodel", 1024, [1, 2, 4, 8, 16, 32])
        run("case13659pegase", "PBRARVmodel", 1024, [8], what=1)
        run("case13659pegase", "PBRARVmodel", 1024, [8], what=2)
        run("case13659pegase", "PBRARVmodel", 1024, [8], what=4)
    if sel == "noscal":
        run("case13659pegase", "PBRARVmodel", 1024, [2, 4, 8], scal=False)
        run("case13659pegase", "PBRARVmodel", 1024, [4], what=1, scal=False)
        run("case13659pegase", "PBRARVmodel", 1024, [4], what=2, scal=False)
        run("case13659pegase", "PBRARVmodel", 1024, [4], what=4, scal=False)
    if sel == "tune":
        for cpb in (2, 4, 8, 16):
            os.environ["PS_CHUNK_PB"] = str(cpb)
            run("case13659pegase", "PBRARVmodel", 2048, [4, 8, 16])
    if sel == "all9":
        for F in ps.ALL_FORMULATIONS:
            run("case13659pegase", F.name, 512, [16])
    if sel == "pn":
        run("case9241pegase", "PNPAPVmodel", 512, [16])
        run("case13659pegase", "CBRARVmodel", 512, [16])
    if sel == "ncu_pn":
        run("case9241pegase", "PNPAPVmodel", 256, [4], steps=1)
    if sel == "pbs":
        run("case13659pegase", "PBRARVmodel", 2048, [8])
        run("case13659pegase", "PBRARVmodel", 2048, [8], what=1)
        os.environ["PS_NO_PBSCN"] = "1"
        run("case13659pegase", "PBRARVmodel", 2048, [8])
        del os.environ["PS_NO_PBSCN"]
    if sel == "ncu_g":
        run("case13659pegase", "PBRARVmodel", 256, [4], what=1, steps=1)
    if sel == "ncu_c3":
        run("case2869pegase", "PNPAPVmodel", 512, [4], steps=1)
    if sel == "ncu_cb":
        run("case13659pegase", "CBRARVmodel", 256, [4], steps=1)
    if sel == "latency":
        for case, Fname in (("case14", "PNPAPVmodel"), ("case300", "PNPAPVmodel"), ("case13659pegase", "PBRARVmodel")):
            d = synth.named_case(case)
            F = ps.BY_NAME[Fname]
            h = capi.Handle(d, F.code, device=0)
            x, _, _, _ = synth.scenario_batch(d, F, 1, seed=1)
            x = x[0]
            mu = np.random.default_rng(0).normal(size=h.m_nl)
            g, J, H = np.zeros(h.m_nl), np.zeros(h.nnzJ), np.zeros(h.nnzH)
            x2 = x * (1.0 + 1e-9)

            def triple():                       # what Ipopt does at a new point: eval_g, eval_jac_g, eval_h at the same x
                triple.k ^= 1
                xx = x2 if triple.k else x
                h.eval_constraint(xx, g); h.eval_jacobian(xx, J); h.eval_hessian(xx, 1.0, mu, H)
            triple.k = 0

            def fresh(fn):                      # defeat the x-keyed cache: alternate between two points
                def run():
                    fresh.k ^= 1
                    fn(x2 if fresh.k else x)
                return run
            fresh.k = 0
            for name, fn in (("eval_constraint (new x)", fresh(lambda xx: h.eval_constraint(xx, g))),
                             ("eval_jacobian (new x)", fresh(lambda xx: h.eval_jacobian(xx, J))),
                             ("eval_jacobian (same x as the last eval_constraint: cached)", lambda: h.eval_jacobian(x, J)),
                             ("eval_hessian (new x)", fresh(lambda xx: h.eval_hessian(xx, 1.0, mu, H))),
                             ("eval_constraint + eval_jacobian + eval_hessian at one new x", triple),
                             ("eval_all (new x)", fresh(lambda xx: h.eval_all(xx, mu, g, J, H)))):
                for _ in range(5):
                    fn()
                t0 = time.perf_counter()
                for _ in range(50):
                    fn()
                dt = (time.perf_counter() - t0) / 50
                print(f"{case} {Fname} {name}: {dt * 1e6:.0f} us per call (host pointers, synchronous)", flush=True)
            h.close()
    if sel == "ncu":
        run("case13659pegase", "PBRARVmodel", 256, [4], steps=1)
    if sel in ("all", "others"):
        run("case13659pegase", "CBRARVmodel", 1024, [4, 16])
        run("case2869pegase", "PNPAPVmodel", 1024, [4, 16])
        run("case9241pegase", "PNPAPVmodel", 512, [4, 16])
        run("case13659pegase", "PNRARVmodel", 512, [4, 16])
        run("case13659pegase", "PBRAPVmodel", 512, [4, 16])
