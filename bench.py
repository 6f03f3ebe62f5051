#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (BASELINE.json metric, config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one pass of the hot path -- the fused g + Jacobian + Hessian-of-the-Lagrangian evaluation of every
NL row -- over one batch of synthetic scenarios of case13659pegase in the rectangular branch-flow formulation
(PBRARVmodel), through the C ABI (ps_batch_eval).  Weak scaling: every GPU owns `--scenarios` scenarios
(4096 by default); scenarios are independent, so the only collective is the NCCL all_gather of the
per-scenario scalars, which is inside the timed region.

`value`   whole-job evaluations/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`     the same metric through ps_batch_eval_host: pinned HOST buffers in, HOST buffers out, H2D/D2H copies
          inside the timed region.
`roofline` achieved algorithmic HBM GB/s of the evaluation kernel against MEASURED_PEAKS.json.
`cpu_baseline` the CPU oracle (a C port of the reference's formulas; the reference itself is Julia/JuMP and
          cannot run here) on the box's host cores, on a bounded sample of the same workload.
`--impl reference` times that CPU port alone (all host threads) and prints the same line shape."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "scenario Jacobian+Hessian evals/sec & achieved HBM GB/s (case13659pegase)"
UNIT = "evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--case", default="case13659pegase")
    ap.add_argument("--formulation", default="PBRARVmodel")
    ap.add_argument("--scenarios", type=int, default=4096, help="scenarios per GPU (weak scaling)")
    ap.add_argument("--e2e-scenarios", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--what", type=int, default=7, help="bitmask G=1 J=2 H=4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--callbacks", action="store_true", help="also time the g-only and g+J passes (outside the timed region)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_config(args, h=None):
    cfg = {"workload": f"{args.case} (synthetic topology, published dimensions) {args.formulation} g+J+H batched evaluation",
           "case": args.case, "formulation": args.formulation, "what": args.what,
           "scenarios_per_gpu": args.scenarios, "l2": "inputs+outputs per step >> 126 MB L2 (no flush needed)",
           "synthetic_branch_order": os.environ.get("PS_SYNTH_ORDER", "random") + " (random = permuted branch list, the adversarial order "
                                     "for the gathers; file = ascending from-bus like real case files)"}
    return cfg


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._thr = index, [], threading.Event(), None
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nv = None

    def _poll(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.rows.append((time.time(), sm, int(reasons)))
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self, t0, t1):
        if self._nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self._nv
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}
        seen = set()
        for _, _, r in rows:
            for bit, nm in names.items():
                if r & bit:
                    seen.add(nm)
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": mx,
                "reasons": sorted(seen), "samples": len(rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_inputs(d, F, S, seed):
    """Host-side seeded base scenarios; expanded on the device by the caller."""
    from powersense_b200 import synth
    B = min(S, 16)
    x, Pd, Qd, lay = synth.scenario_batch(d, F, B, seed=seed)
    return x, Pd, Qd, lay


def cpu_baseline(d, F, args, seconds, nthreads=0):
    """Times oracle/ps_oracle.c (the CPU port of the reference's formulas) on a bounded sample."""
    from oracle import c_oracle
    from powersense_b200 import synth
    o = c_oracle.COracle(d, F.code)
    cores = nthreads or os.cpu_count() or c_oracle.lib().pso_max_threads()    # torchrun sets OMP_NUM_THREADS=1: ask the OS
    n0 = max(cores, 4)
    x, Pd, Qd, _ = synth.scenario_batch(d, F, n0, seed=77)
    mu = np.random.default_rng(5).normal(size=(n0, o.m_nl))
    out = (np.zeros((n0, o.m_nl)), np.zeros((n0, o.nnzJ)), np.zeros((n0, o.nnzH)))
    t = time.time(); o.eval_batch(x, mu, Pd, Qd, what=args.what, nthreads=cores, out=out); dt0 = time.time() - t
    reps = max(1, int(seconds / max(dt0, 1e-3)))
    reps = min(reps, 50)
    t = time.time()
    for _ in range(reps):
        o.eval_batch(x, mu, Pd, Qd, what=args.what, nthreads=cores, out=out)
    dt = time.time() - t
    return {"value": n0 * reps / dt, "unit": UNIT, "cores": int(cores), "kind": "port",
            "sample": f"{reps} x {n0} scenarios of {args.case} {F.name} (what={args.what}), oracle/ps_oracle.c with OpenMP over scenarios"}


def run_reference(args, rank):
    """Reference arm: the reference's CPU evaluator is JuMP (Julia), which cannot run here; the committed C port
    of its formulas (oracle/ps_oracle.c) stands in, on all host threads, one bounded sample per step."""
    if rank != 0:
        return
    import __graft_entry__ as ge
    ge.build()
    import powersense_b200 as ps
    from powersense_b200 import synth
    from oracle import c_oracle
    d = synth.named_case(args.case)
    F = ps.BY_NAME[args.formulation]
    o = c_oracle.COracle(d, F.code)
    cores = os.cpu_count() or c_oracle.lib().pso_max_threads()     # torchrun sets OMP_NUM_THREADS=1: ask the OS
    n = max(cores, 4)
    x, Pd, Qd, _ = synth.scenario_batch(d, F, n, seed=77)
    mu = np.random.default_rng(5).normal(size=(n, o.m_nl))
    out = (np.zeros((n, o.m_nl)), np.zeros((n, o.nnzJ)), np.zeros((n, o.nnzH)))
    for _ in range(args.warmup):
        o.eval_batch(x, mu, Pd, Qd, what=args.what, nthreads=cores, out=out)
    t = time.time()
    for _ in range(args.steps):
        o.eval_batch(x, mu, Pd, Qd, what=args.what, nthreads=cores, out=out)
    dt = time.time() - t
    val = n * args.steps / dt
    cfg = workload_config(args)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(cores), "kind": "port",
                             "sample": f"each step = {n} scenarios of {args.case} {F.name} (what={args.what}) on {cores} host threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the evaluator has no CPU fallback")
    import __graft_entry__ as ge
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()                      # no-op when the in-tree .so files are up to date
    if world > 1:
        dist.barrier()
    import powersense_b200 as ps
    from powersense_b200 import capi, synth
    from powersense_b200.driver import ScenarioDriver

    d = synth.named_case(args.case)
    F = ps.BY_NAME[args.formulation]
    S = args.scenarios
    drv = ScenarioDriver(d, F, S * world, rank, world, device_index=local, what=args.what)
    h = drv.handle
    dev = drv.device
    # ---- synthetic inputs: 16 seeded base scenarios, expanded and perturbed on the device per global scenario
    xb, Pdb, Qdb, lay = make_inputs(d, F, S, seed=1000 + rank)
    B = xb.shape[0]
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242 + rank)
    xbd = torch.from_numpy(xb).to(dev)
    for s0 in range(0, S, 256):
        n = min(256, S - s0)
        idx = (torch.arange(s0, s0 + n, device=dev) % B)
        noise = 1.0 + 1e-3 * (torch.rand(n, h.n_var, generator=gen, device=dev, dtype=torch.float64) - 0.5)
        drv.x[s0:s0 + n] = xbd[idx] * noise
        del noise
    if drv.mu is not None:
        for s0 in range(0, S, 256):
            n = min(256, S - s0)
            drv.mu[s0:s0 + n] = torch.randn(n, h.m_nl, generator=gen, device=dev, dtype=torch.float64)
    rng = np.random.default_rng(31 + rank)
    reps = -(-S // B)
    Pd = np.tile(Pdb, (reps, 1))[:S] * (1.0 + 0.01 * rng.standard_normal((S, 1)))
    Qd = np.tile(Qdb, (reps, 1))[:S] * (1.0 + 0.01 * rng.standard_normal((S, 1)))
    drv.set_loads(Pd, Qd)
    del Pd, Qd
    torch.cuda.synchronize()

    def step():
        drv.evaluate()
        if world > 1:
            drv.gather_scalars()

    sampler = ClockSampler(local)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = drv.batch.launch_count
    drv.batch.enable_timing(True)          # CUDA events around every kernel, recorded on the launching stream
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ktimes = []
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for k in range(args.steps):
        ev[k][0].record()
        drv.evaluate()
        ev[k][1].record()
        if world > 1:
            drv.gather_scalars()
        ktimes.append(drv.batch.kernel_times())      # waits for this step's kernels (event sync, no device idle of note)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    kavg = {k: float(np.mean([t[k] for t in ktimes])) for k in ktimes[0]}
    launches = drv.batch.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.summary(t0, t1)
    ms_per_step = ms_total / args.steps
    value = S * world / (ms_per_step * 1e-3)
    alg_bytes = h.algorithmic_bytes(args.what)
    alg_branch = h.algorithmic_bytes_branch(args.what)
    peak, peak_src = measured_peak()
    achieved_step = alg_bytes * S / (kernel_ms * 1e-3) / 1e9
    achieved = alg_branch * S / (kavg["branch"] * 1e-3) / 1e9      # the dominant kernel: branch rows
    info = h.plan_info()
    # ScenarioDriver allocates the preferred (sector-aligned) layout: the bulk (TMA store) branch kernel runs unless disabled
    if info["direct_ok"]:
        nr_ = max(len(d.br), 1)
        long_blocks = (h.nnzJ - info["j_br0"]) // nr_ + (h.nnzH - info["h_br0"]) // nr_ >= 48     # launch_eval's rule
        use_bulk = (long_blocks or os.environ.get("PS_BULK")) and not os.environ.get("PS_NO_BULK")
        kname = "branch_bulk_kernel" if use_bulk else "branch_direct_kernel"
        if os.environ.get("PS_NO_BULK") and os.environ.get("PS_NO_DIRECT"):
            kname = "branch_staged_kernel"
    else:
        kname = "branch_staged_kernel"
    kname += f"<{args.formulation[:6]}>"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")             # dram bytes/launch from the committed ncu capture
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            key = f"{args.case}:{args.formulation}:{S}:{args.what}"
            traffic = tj.get(key, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # one scenario's scalars make a quick sanity line for the log (stderr)
    sc = drv.scalars[:2].cpu().numpy()
    print(f"[rank {rank}] S={S} step {kernel_ms:.3f} ms ({achieved_step:.0f} GB/s algorithmic); kernels {kavg}; branch kernel "
          f"{achieved:.0f} GB/s; scalars[0]={sc[0].tolist()}",
          file=sys.stderr, flush=True)

    # ---- the other callbacks of the boundary: the SLP line search asks for g only, its LP set-up for g + J.
    # Not part of `value` (opt-in, --callbacks, so that the default command launches the g + J + H pass only); per GPU,
    # device-resident, CUDA events, 5 passes each after one warm-up pass
    callbacks = {}
    drv.batch.enable_timing(False)
    for name, w in (("g", 1), ("g+J", 3)):
        if not args.callbacks or (args.what & w) != w:
            continue
        drv.evaluate(w)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            drv.evaluate(w)
        c1.record()
        torch.cuda.synchronize()
        ms = c0.elapsed_time(c1) / 5
        ab = h.algorithmic_bytes(w)
        callbacks[name] = {"evals_per_s_per_gpu": S / (ms * 1e-3), "ms_per_pass": ms, "algorithmic_gbs": ab * S / (ms * 1e-3) / 1e9}
    # ---- e2e: host buffers through ps_batch_eval_host
    e2e = None
    if not args.no_e2e:
        Se = min(args.e2e_scenarios, S)
        hx = capi.pinned_empty((Se, h.n_var)); hx[:] = drv.x[:Se].cpu().numpy()
        hmu = capi.pinned_empty((Se, h.m_nl)); hmu[:] = drv.mu[:Se].cpu().numpy() if drv.mu is not None else 0.0
        hg = capi.pinned_empty((Se, h.m_nl)) if args.what & 1 else None
        hJ = capi.pinned_empty((Se, h.nnzJ)) if args.what & 2 else None
        hH = capi.pinned_empty((Se, h.nnzH)) if args.what & 4 else None
        hs = capi.pinned_empty((Se, capi.PS_NSCALARS))
        l0 = drv.batch.launch_count
        drv.batch.eval_host(args.what, hx, hmu, hg, hJ, hH, hs)          # warm-up (allocates the staging buffers)
        if world > 1:
            dist.barrier()
        ta = time.time()
        for _ in range(args.e2e_steps):
            drv.batch.eval_host(args.what, hx, hmu, hg, hJ, hH, hs)
        tb = time.time() - ta
        if world > 1:
            t = torch.tensor([tb], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tb = float(t.item())
        if args.what & 2:
            assert np.array_equal(hJ[0], drv.J[0].cpu().numpy()), "e2e path disagrees with the device path"
        h2d = Se * (h.n_var + (h.m_nl if args.what & 4 else 0)) * 8
        d2h = Se * ((h.m_nl if args.what & 1 else 0) + (h.nnzJ if args.what & 2 else 0) + (h.nnzH if args.what & 4 else 0) +
                    capi.PS_NSCALARS) * 8
        e2e = {"value": Se * world * args.e2e_steps / tb, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "scenarios_per_step_per_gpu": Se, "steps": args.e2e_steps,
               "note": "ps_batch_eval_host: pinned host x/mu in, host g/J/H/scalars out, chunked 3-stream pipeline; PCIe-bound"}
        del hx, hmu, hg, hJ, hH, hs
    sampler.stop()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(d, F, args, args.cpu_seconds)

    if rank == 0:
        cfg = workload_config(args)
        cfg.update({"n_var": h.n_var, "m_nl": h.m_nl, "nnz_jac": h.nnzJ, "nnz_hess": h.nnzH,
                    "algorithmic_bytes_per_eval": alg_bytes, "scenarios_total": S * world,
                    "parallelism": f"scenario-sharded x{world}, all_gather of {capi.PS_NSCALARS} scalars/scenario" if world > 1 else "1 GPU"})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "callbacks": callbacks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "kernel": kname,
                             "kernel_ms_per_launch": kavg["branch"], "units_per_launch": S,
                             "algorithmic_bytes_per_unit": alg_branch,
                             "step": {"achieved": achieved_step, "frac": achieved_step / peak, "ms": kernel_ms,
                                      "algorithmic_bytes_per_unit": alg_bytes, "kernels_ms": kavg,
                                      "note": "all kernels of one evaluation (bus rows, bus J/H slots, objective, branch rows)"}},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    drv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
