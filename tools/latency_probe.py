"""Where does a single-scenario ps_eval_all spend its time?  (development probe: case13659pegase PBRARV through the host-pointer
callbacks, output subsets and kernel-path knobs; prints microseconds per call at a new x)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import powersense_b200 as ps  # noqa: E402
from powersense_b200 import capi, synth  # noqa: E402

case, Fname = (sys.argv[1:3] + ["case13659pegase", "PBRARVmodel"])[:2] if len(sys.argv) >= 3 else ("case13659pegase", "PBRARVmodel")
d = synth.named_case(case)
F = ps.BY_NAME[Fname]
h = capi.Handle(d, F.code, device=0)
x, _, _, _ = synth.scenario_batch(d, F, 1, seed=1)
x = x[0]
x2 = x * (1.0 + 1e-9)
mu = np.random.default_rng(0).normal(size=h.m_nl)
g, J, H = np.zeros(h.m_nl), np.zeros(h.nnzJ), np.zeros(h.nnzH)
g[:] = 1; J[:] = 1; H[:] = 1
lib = capi.lib()
pd = capi._pd


def timeit(name, fn, n=20):
    k = [0]

    def run():
        k[0] ^= 1
        fn(x2 if k[0] else x)
    for _ in range(5):
        run()
    t0 = time.perf_counter()
    for _ in range(n):
        run()
    print(f"{case} {Fname} {name}: {(time.perf_counter() - t0) / n * 1e6:.0f} us", flush=True)


N = None
for label, env in (("default", {}), ("PS_NO_BULK", {"PS_NO_BULK": "1"})):
    for k_, v_ in env.items():
        os.environ[k_] = v_
    timeit(f"[{label}] eval_all g+J+H", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), pd(g), pd(J), pd(H)))
    timeit(f"[{label}] eval_all g+J", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), pd(g), pd(J), N))
    timeit(f"[{label}] eval_all J+H", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), N, pd(J), pd(H)))
    timeit(f"[{label}] eval_all g+H", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), pd(g), N, pd(H)))
    timeit(f"[{label}] eval_all H", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), N, N, pd(H)))
    timeit(f"[{label}] eval_all g", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), pd(g), N, N))
    for k_ in env:
        del os.environ[k_]

with capi.host_registered(g, J, H):
    timeit("[caller's arrays page-locked: ps_host_register] eval_all g+J+H", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), pd(g), pd(J), pd(H)))
    timeit("[page-locked] eval_all g", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), pd(g), N, N))
    timeit("[page-locked] eval_all g+J", lambda xx: lib.ps_eval_all(h.h, pd(xx), pd(mu), pd(g), pd(J), N))
