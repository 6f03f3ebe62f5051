"""Development probe: case300 (MATPOWER) PNPAPV with the quadratic objective; the reference's example quotes 7.1973e+05
(examples/opf/ACOPF_formulations_example.jl:27)."""
import copy, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import powersense_b200 as ps
from powersense_b200 import slp
from cases import load_fixture

which = sys.argv[1] if len(sys.argv) > 1 else "slp"
M = copy.deepcopy(load_fixture("case300"))
t0 = time.time()
if which == "slsqp":
    res = ps.run_opf(M, formulation=ps.PNPAPVmodel, obj_type="quadratic", maxiter=300, verbose=1)
    print("SLSQP cost", M.cost, "time", time.time() - t0, getattr(res, "message", None), getattr(res, "nit", None))
else:
    ps.create_opf_model(M, formulation=ps.PNPAPVmodel, obj_type="quadratic")
    drv = slp.BatchedSlpLS(M.model, 1, options=slp.SlpOptions(max_iter=int(sys.argv[2]) if len(sys.argv) > 2 else 100))
    out = drv.run()
    print("SLP status", out["status"], "obj", out["obj"], "iters", out["iter"], "time", time.time() - t0)
    for h in out["history"][-5:]:
        print({k: (float(v[0]) if hasattr(v, "__len__") else v) for k, v in h.items()})
