"""The N-1 sweep of tools/n1_sweep.py with the TEST-ONLY CPU evaluation backend (tests/oracle_backend.py: the oracles instead
of the GPU) -- for studying the SLP's convergence where no GPU is at hand.  Not a product path: the product's backend is
GpuBackend.  usage: python tools/n1_sweep_cpu.py case300 [n_outages [lp_workers [max_iter [first_outage]]]]"""
import copy
import os
import sys
import json
import collections

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import powersense_b200 as ps  # noqa: E402
from powersense_b200 import slp, synth  # noqa: E402
from cases import load_fixture  # noqa: E402
from oracle_backend import OracleBackend  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "case300"
    fixture = case in ("case14", "case300", "case3")
    M = copy.deepcopy(load_fixture(case)) if fixture else synth.named_case(case.split(":")[0], feasible=True)
    nout = min(int(sys.argv[2]), M.nbr) if len(sys.argv) > 2 else M.nbr
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 600
    k0 = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    obj = os.environ.get("PS_OBJ", "linear" if case == "case14" else "quadratic")
    model = ps.create_opf_model(M, formulation=ps.PNPAPVmodel, obj_type=obj, device=-2)
    S = nout + 1
    st = np.ones((S, M.nbr), dtype=np.uint8)
    for k in range(nout):
        st[k + 1, k0 + k] = 0
    be = OracleBackend(model, S, status=st)
    drv = slp.BatchedSlpLS(None, S, backend=be, options=slp.SlpOptions(max_iter=max_iter, lp_workers=workers))
    res = drv.run()
    names = {0: "LOCALLY_SOLVED", 2: "LOCALLY_INFEASIBLE", 6: "ALMOST_LOCALLY_SOLVED", -1: "ITERATION_LIMIT", -3: "NUMERICAL_ERROR", -5: "OTHER"}
    g = res["g"]
    viol = np.maximum(np.maximum(g - drv.g_U, drv.g_L - g), 0.0).max(axis=1)
    print(f"{case} PNPAPVmodel obj={obj}: base + outages {k0 + 1}..{k0 + nout}, max_iter {max_iter}")
    print("  status counts:", dict(collections.Counter(names.get(int(v), int(v)) for v in res["status"])))
    ok = res["status"] == 0
    if ok.any():
        print(f"  LOCALLY_SOLVED: objective {res['obj'][ok].min():.3f} .. {res['obj'][ok].max():.3f} (base {res['obj'][0]:.3f}), iterations "
              f"{int(res['iter'][ok].min())} .. {int(res['iter'][ok].max())}, max violation {viol[ok].max():.2e}")
    print("  timing:", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in res["timing"].items()})
    out = {"status": res["status"].tolist(), "obj": res["obj"].tolist(), "iter": res["iter"].tolist(), "viol": viol.tolist(), "k0": k0}
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"n1_cpu_{case.replace(':', '_')}_{k0}_{nout}.json"), "w"))
    drv.close()


if __name__ == "__main__":
    main()
