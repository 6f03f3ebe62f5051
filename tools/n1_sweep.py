"""BASELINE config 5 at fixture scale: the N-1 branch-contingency sweep of the reference's 14-bus case solved with the
lock-step batched SLP line search (powersense_b200.slp.BatchedSlpLS): one scenario per branch outage plus the base case,
every evaluation on the GPU, the LP subproblems on the host.  Prints status / objective / iterations per scenario and
where the wall time goes.  usage: python tools/n1_sweep.py [formulation [case [n_outages [lp_workers]]]] -- case: case14 (the
reference's fixture, default) or a synthetic named case (powersense_b200.synth), e.g. case118"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import powersense_b200 as ps  # noqa: E402
from powersense_b200 import slp  # noqa: E402
from cases import load_fixture  # noqa: E402


def main():
    F = ps.BY_NAME[sys.argv[1]] if len(sys.argv) > 1 else ps.PNPAPVmodel
    case = sys.argv[2] if len(sys.argv) > 2 else "case14"
    from powersense_b200 import synth
    fixture = case in ("case14", "case300", "case3")
    feasible = case.endswith(":feasible")               # a synthetic named case with loads from a solved power flow
    M = copy.deepcopy(load_fixture(case)) if fixture else synth.named_case(case.split(":")[0], feasible=feasible)
    obj = os.environ.get("PS_OBJ", "linear" if case == "case14" else "quadratic")
    model = ps.create_opf_model(M, formulation=F, obj_type=obj)
    nout = min(int(sys.argv[3]), M.nbr) if len(sys.argv) > 3 else M.nbr
    S = nout + 1
    st = np.ones((S, M.nbr), dtype=np.uint8)
    for k in range(nout):
        st[k + 1, k] = 0
    Pd, Qd = np.tile(M.Pd, (S, 1)), np.tile(M.Qd, (S, 1))
    workers = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    drv = slp.BatchedSlpLS(model, S, Pd, Qd, status=st, options=slp.SlpOptions(max_iter=int(os.environ.get("PS_MAX_ITER", "100")), lp_workers=workers))
    res = drv.run()
    names = {0: "LOCALLY_SOLVED", 2: "LOCALLY_INFEASIBLE", 6: "ALMOST_LOCALLY_SOLVED", -1: "ITERATION_LIMIT", -3: "NUMERICAL_ERROR",
             -5: "OTHER"}
    g = res["g"]
    viol = np.maximum(np.maximum(g - drv.g_U, drv.g_L - g), 0.0).max(axis=1)
    print(f"{F.name} {case}: {S} scenarios (base + {nout} single-branch outages)")
    import collections
    print("  status counts:", dict(collections.Counter(names.get(int(v), int(v)) for v in res["status"])))
    ok = res["status"] == 0
    if ok.any():
        print(f"  LOCALLY_SOLVED: objective {res['obj'][ok].min():.3f} .. {res['obj'][ok].max():.3f} (base case {res['obj'][0]:.3f}), "
              f"iterations {int(res['iter'][ok].min())} .. {int(res['iter'][ok].max())}, max violation {viol[ok].max():.2e}")
    for s in range(S if S <= 40 else 0):
        what = "base case" if s == 0 else f"branch {s} ({M.br[s - 1][0]} - {M.br[s - 1][1]}) out"
        print(f"  {what:28s} {names.get(int(res['status'][s]), res['status'][s]):22s} objective {res['obj'][s]:9.3f}  "
              f"iterations {int(res['iter'][s]):3d}  max violation {viol[s]:.2e}")
    tm = res["timing"]
    print(f"wall {tm['total_s']:.2f} s: host LPs {tm['host_lp_s']:.2f} s ({tm['lp_calls']} calls, {f"{workers} worker processes" if workers else "serial"}), GPU evaluations + metrics "
          f"{tm['gpu_eval_s']:.2f} s ({len(res['history'])} lock-step iterations, {tm['trial_batches']} batched trial points)")
    drv.close()
    model.close()


if __name__ == "__main__":
    main()
