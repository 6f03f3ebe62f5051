#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (BASELINE.json metric, config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one pass of the hot path -- the fused g + Jacobian + Hessian-of-the-Lagrangian evaluation of every
NL row -- over one batch of synthetic scenarios of case13659pegase in the rectangular branch-flow formulation
(PBRARVmodel), through the C ABI (ps_batch_eval).  BASELINE config 4 as written: `--scenarios` scenarios (4096) IN
TOTAL, sharded over the N GPUs ("scaling": "strong"; N = 1 is the whole batch on one GPU); scenarios are independent, so
the only collective is the NCCL all_gather of the per-scenario scalars, inside the timed region (issued on its own
stream so that it overlaps the next evaluation).  For N > 1 the weak-scaled figure (4096 scenarios per GPU) is reported
beside it under "weak".

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
    ap.add_argument("--scenarios", type=int, default=4096, help="scenarios in total (config 4: sharded over the GPUs)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra weak-scaled pass (--scenarios per GPU)")
    ap.add_argument("--e2e-scenarios", type=int, default=1024)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs block (config 3 and config 5's evaluator)")
    ap.add_argument("--what", type=int, default=7, help="bitmask G=1 J=2 H=4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-callbacks", action="store_true", help="skip the g-only and g+J passes (timed outside the timed region)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_config(args, dims, world):
    """The same dictionary for the native and the reference arm (dims = n_var, m_nl, nnz_jac, nnz_hess)."""
    n_var, m_nl, nnzJ, nnzH = dims
    cfg = {"workload": f"{args.case} (synthetic topology, published dimensions) {args.formulation} g+J+H batched evaluation",
           "case": args.case, "formulation": args.formulation, "what": args.what,
           "scenarios_total": args.scenarios, "l2": "inputs+outputs per step >> 126 MB L2 (no flush needed)",
           "synthetic_branch_order": os.environ.get("PS_SYNTH_ORDER", "random") + " (random = permuted branch list, the adversarial order "
                                     "for the gathers; file = ascending from-bus like real case files)",
           "n_var": int(n_var), "m_nl": int(m_nl), "nnz_jac": int(nnzJ), "nnz_hess": int(nnzH),
           "parallelism": f"{args.scenarios} scenarios sharded over {world} GPU(s), all_gather of 4 scalars/scenario" if world > 1 else "1 GPU"}
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
    cfg = workload_config(args, (o.n_var, o.m_nl, o.nnzJ, o.nnzH), int(os.environ.get("WORLD_SIZE", "1")))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
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
    peak, peak_src = measured_peak()

    def fill_inputs(drv, seed_off=0):
        """16 seeded base scenarios, expanded and perturbed on the device per global scenario"""
        h, S, dev = drv.handle, drv.S, drv.device
        xb, Pdb, Qdb, _ = make_inputs(d, F, S, seed=1000 + rank + seed_off)
        B = xb.shape[0]
        gen = torch.Generator(device=dev)
        gen.manual_seed(4242 + rank + seed_off)
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
        rng = np.random.default_rng(31 + rank + seed_off)
        reps = -(-S // B)
        Pd = np.tile(Pdb, (reps, 1))[:S] * (1.0 + 0.01 * rng.standard_normal((S, 1)))
        Qd = np.tile(Qdb, (reps, 1))[:S] * (1.0 + 0.01 * rng.standard_normal((S, 1)))
        drv.set_loads(Pd, Qd)
        torch.cuda.synchronize()

    def timed_steps(drv, steps, warmup, sampler=None):
        """W warm-up steps, then K steps between a barrier + synchronize on both sides: one evaluation of the rank's shard +
        the gather of its scalars per step; CUDA events; max over ranks.  No host synchronisation inside the loop."""
        for _ in range(max(warmup, 3)):
            drv.evaluate()
            drv.gather_scalars_async()
        drv.wait_gathers()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if sampler is not None:
            sampler.start()
        l0 = drv.batch.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            drv.evaluate()
            drv.gather_scalars_async()
        drv.wait_gathers()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=drv.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, drv.batch.launch_count - l0, (t0, t1)

    def kernel_pass(drv, n=3):
        """per-kernel durations (CUDA events recorded around every kernel on the launching stream, inside the library) and
        the whole evaluation, sampled in a separate pass so that the timed loop carries no host synchronisation"""
        drv.batch.enable_timing(True)
        kt, ev = [], []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); drv.evaluate(); b.record()
            kt.append(drv.batch.kernel_times())
            torch.cuda.synchronize()
            ev.append(a.elapsed_time(b))
        drv.batch.enable_timing(False)
        return {k: float(np.mean([t[k] for t in kt])) for k in kt[0]}, float(np.mean(ev))

    S_total = args.scenarios
    drv = ScenarioDriver(d, F, S_total, rank, world, device_index=local, what=args.what)
    h, S, dev = drv.handle, drv.S, drv.device
    fill_inputs(drv)
    sampler = ClockSampler(local)
    ms_per_step, launches, (t0, t1) = timed_steps(drv, args.steps, args.warmup, sampler)
    clocks = sampler.summary(t0, t1)
    value = S_total / (ms_per_step * 1e-3)
    kavg, kernel_ms = kernel_pass(drv)
    alg_bytes = h.algorithmic_bytes(args.what)
    alg_branch = h.algorithmic_bytes_branch(args.what)
    # the whole step, from the timed loop itself (what `value` measures; at N > 1 it includes the exposed part of the gather)
    achieved_step = alg_bytes * S / (ms_per_step * 1e-3) / 1e9
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
            e = tj.get(f"{args.case}:{args.formulation}:{args.what}", {})
            if e.get("dram_bytes_per_scenario") is not None:
                traffic = e["dram_bytes_per_scenario"] * S          # per launch, like `achieved`
        except Exception:
            traffic = None
    sc = drv.scalars[:2].cpu().numpy()
    print(f"[rank {rank}] S={S} of {S_total}: step {ms_per_step:.3f} ms; evaluation {kernel_ms:.3f} ms ({achieved_step:.0f} GB/s "
          f"algorithmic); kernels {kavg}; branch kernel {achieved:.0f} GB/s; scalars[0]={sc[0].tolist()}", file=sys.stderr, flush=True)

    def passes(drv_, w, n=5):
        drv_.evaluate(w)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(n):
            drv_.evaluate(w)
        c1.record()
        torch.cuda.synchronize()
        return c0.elapsed_time(c1) / n

    # ---- the other callbacks of the boundary: the SLP line search asks for g only, its LP set-up for g + J.
    # Outside the timed region; per GPU, device-resident, CUDA events, 5 passes each after one warm-up pass
    callbacks = {}
    for name, w in (("g", 1), ("g+J", 3)):
        if args.no_callbacks or (args.what & w) != w:
            continue
        ms = passes(drv, w)
        ab = h.algorithmic_bytes(w)
        callbacks[name] = {"evals_per_s_per_gpu": S / (ms * 1e-3), "ms_per_pass": ms, "algorithmic_gbs": ab * S / (ms * 1e-3) / 1e9,
                           "frac_of_peak": ab * S / (ms * 1e-3) / 1e9 / peak}
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

        def host_steps(w, n):
            drv.batch.eval_host(w, hx, hmu, hg if w & 1 else None, hJ if w & 2 else None, hH if w & 4 else None, hs)   # warm-up
            if world > 1:
                dist.barrier()
            ta = time.time()
            for _ in range(n):
                drv.batch.eval_host(w, hx, hmu, hg if w & 1 else None, hJ if w & 2 else None, hH if w & 4 else None, hs)
            tb = time.time() - ta
            if world > 1:
                t = torch.tensor([tb], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                tb = float(t.item())
            return tb
        tb = host_steps(args.what, args.e2e_steps)
        if args.what & 2:
            assert np.array_equal(hJ[0], drv.J[0].cpu().numpy()), "e2e path disagrees with the device path"
        h2d = Se * (h.n_var + (h.m_nl if args.what & 4 else 0)) * 8
        d2h = Se * ((h.m_nl if args.what & 1 else 0) + (h.nnzJ if args.what & 2 else 0) + (h.nnzH if args.what & 4 else 0) +
                    capi.PS_NSCALARS) * 8
        e2e = {"value": Se * world * args.e2e_steps / tb, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "scenarios_per_step_per_gpu": Se, "steps": args.e2e_steps,
               "d2h_gbs_per_gpu": d2h * args.e2e_steps / tb / 1e9,
               "note": "ps_batch_eval_host: pinned host x/mu in, host g/J/H/scalars out, chunked 3-stream pipeline; bound by the "
                       "host link (PCIe 5 x16: ~56 GB/s per direction measured alone, tools/pcie_probe.py), at N > 1 by the host's "
                       "aggregate copy bandwidth (profiles/r02_host_ceiling.txt)"}
        if (args.what & 3) == 3 and not args.no_callbacks:
            # what the SLP consumes per iteration (g + J: no Hessian crosses the link)
            tgj = host_steps(3, max(3, args.e2e_steps // 2))
            e2e["g+J"] = {"value": Se * world * max(3, args.e2e_steps // 2) / tgj, "unit": UNIT,
                          "d2h_bytes_per_step": int(Se * (h.m_nl + h.nnzJ + capi.PS_NSCALARS) * 8)}
        del hx, hmu, hg, hJ, hH, hs
    sampler.stop()
    n_var, m_nl, nnzJ, nnzH = h.n_var, h.m_nl, h.nnzJ, h.nnzH
    drv.close()
    del drv
    torch.cuda.empty_cache()

    # ---- N > 1: the weak-scaled figure (every GPU owns `--scenarios` scenarios) beside the strong-scaled headline
    weak = None
    if world > 1 and not args.no_weak:
        drvw = ScenarioDriver(d, F, S_total * world, rank, world, device_index=local, what=args.what)
        fill_inputs(drvw, seed_off=7)
        msw, _, _ = timed_steps(drvw, max(5, args.steps // 2), 3)
        weak = {"value": S_total * world / (msw * 1e-3), "unit": UNIT, "ms_per_step": msw, "scenarios_per_gpu": S_total,
                "scaling": "weak"}
        drvw.close()
        del drvw
        torch.cuda.empty_cache()

    # ---- the BASELINE configs that use the nodal formulations, so that the driver's record carries them: 5 passes each,
    # device-resident, outside the timed region, one GPU's worth (rank 0)
    extra = {}
    if rank == 0 and not args.no_extra:
        for label, case, fname, Sx, masked in (("config3", "case2869pegase", "PNPAPVmodel", 1024, False),
                                                ("config5_evaluator", "case9241pegase", "PNPAPVmodel", 512, True)):
            dx = synth.named_case(case)
            Fx = ps.BY_NAME[fname]
            dr = ScenarioDriver(dx, Fx, Sx, 0, 1, device_index=local, what=7)
            xb, Pdb, Qdb, _ = synth.scenario_batch(dx, Fx, 16, seed=5)
            reps = -(-Sx // 16)
            dr.x[:] = torch.from_numpy(np.tile(xb, (reps, 1))[:Sx]).to(dev)
            dr.x *= 1.0 + 1e-3 * (torch.rand_like(dr.x) - 0.5)
            dr.mu[:] = torch.randn(Sx, dr.handle.m_nl, dtype=torch.float64, device=dev)
            dr.set_loads(np.tile(Pdb, (reps, 1))[:Sx], np.tile(Qdb, (reps, 1))[:Sx])
            if masked:
                from powersense_b200.driver import contingency_status
                dr.set_branch_status(contingency_status(dx.nbr, 0, Sx))     # N-1 sweep: scenario s opens branch s
            res = {"case": case, "formulation": fname, "scenarios": Sx, "contingency_mask": masked}
            for name, w in (("g+J+H", 7), ("g", 1), ("g+J", 3)):
                ms = passes(dr, w)
                ab = dr.handle.algorithmic_bytes(w)
                res[name] = {"evals_per_s": Sx / (ms * 1e-3), "ms_per_pass": ms, "algorithmic_gbs": ab * Sx / (ms * 1e-3) / 1e9,
                             "frac_of_peak": ab * Sx / (ms * 1e-3) / 1e9 / peak}
            kx, _ = kernel_pass(dr, 2)
            res["kernels_ms"] = kx
            extra[label] = res
            dr.close()
            del dr
            torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(d, F, args, args.cpu_seconds)

    if rank == 0:
        cfg = workload_config(args, (n_var, m_nl, nnzJ, nnzH), world)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "callbacks": callbacks, "weak": weak, "extra_configs": extra,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "kernel": kname,
                             "kernel_ms_per_launch": kavg["branch"], "units_per_launch": S,
                             "algorithmic_bytes_per_unit": alg_branch,
                             "note": "`frac` is the dominant kernel's (branch rows); the whole evaluation -- what `value` measures -- is under `step`",
                             "step": {"achieved": achieved_step, "frac": achieved_step / peak, "ms": ms_per_step,
                                      "algorithmic_bytes_per_unit": alg_bytes, "units_per_gpu": S, "kernels_ms": kavg,
                                      "evaluation_ms_with_kernel_events": kernel_ms,
                                      "note": "per GPU: algorithmic bytes of the rank's shard / ms_per_step of the timed loop; kernels_ms = the "
                                              "kernels of one evaluation (bus rows, bus J/H slots, objective, branch rows), sampled in a "
                                              "separate pass with CUDA events around every kernel (the events add gaps)"}},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
