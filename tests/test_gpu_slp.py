"""N4: the lock-step batched SLP line search (the reference's `OPT` algorithm, slp_line_search.jl:80-278) with every
evaluation on the GPU -- NL rows, affine/quadratic rows, CSC assembly, KKT metrics, backtracking trial points -- and the LP
subproblems on the host (scipy/HiGHS standing in for GLPK).  The reference's own SLP test asserts the 14-bus objective
146.7 at rtol 0.1 with PNPAPVmodel (examples/opf/ACOPF_formulations_example.jl:7, test/opf/ACOPF_formulations.jl:9)."""
import copy

import numpy as np
import pytest

import powersense_b200 as ps
from powersense_b200 import slp
from powersense_b200.driver import contingency_status

from cases import load_fixture

pytestmark = pytest.mark.gpu


def test_case14_slp_load_scenarios_and_contingencies():
    M = copy.deepcopy(load_fixture("case14"))
    model = ps.create_opf_model(M, formulation=ps.PNPAPVmodel, obj_type="linear")
    S = 5
    scale = np.array([1.0, 1.03, 0.97, 1.0, 1.0])
    Pd, Qd = M.Pd[None, :] * scale[:, None], M.Qd[None, :] * scale[:, None]
    st = np.ones((S, M.nbr), dtype=np.uint8)
    st[3, 2] = 0                                         # N-1: branch 3 (2 -> 3) out
    st[4, 9] = 0                                         # N-1: branch 10 out
    drv = slp.BatchedSlpLS(model, S, Pd, Qd, status=st, options=slp.SlpOptions(max_iter=80))
    res = drv.run()
    assert np.all(res["status"] == 0), res["status"]     # LOCALLY_SOLVED for every scenario
    assert res["obj"][0] == pytest.approx(146.7, rel=0.1)                 # the reference's assertion
    assert res["obj"][0] == pytest.approx(146.77, rel=5e-3)               # and close to the NLP optimum found by SLSQP
    assert res["obj"][1] > res["obj"][0] > res["obj"][2]                  # more load costs more
    assert res["obj"][3] >= res["obj"][0] - 0.05 and res["obj"][4] >= res["obj"][0] - 0.05   # an outage cannot make it cheaper (SLP tolerance)
    # primal feasibility at the reference's tolerance (tol_infeas = 0.01, parameters.jl:18) for every scenario
    g = res["g"]
    viol = np.maximum(np.maximum(g - drv.g_U, drv.g_L - g), 0.0).max(axis=1)
    assert np.all(viol <= 0.01 + 1e-9), viol
    assert np.all(res["iter"] < 80)
    drv.close()
    model.close()


def test_case14_n1_branch_contingency_sweep():
    """BASELINE config 5 at fixture scale: base case + every single-branch outage of the 14-bus case in one lock-step
    batch (tools/n1_sweep.py prints the table).  Every outage that leaves the network connected is LOCALLY_SOLVED with an
    objective at or above the base case's (to the SLP tolerance); opening the radial branch 7 - 8 islands the condenser
    bus, which the SLP reports as not solved to optimality."""
    M = copy.deepcopy(load_fixture("case14"))
    model = ps.create_opf_model(M, formulation=ps.PNPAPVmodel, obj_type="linear")
    S = M.nbr + 1
    st = np.ones((S, M.nbr), dtype=np.uint8)
    for k in range(M.nbr):
        st[k + 1, k] = 0
    drv = slp.BatchedSlpLS(model, S, np.tile(M.Pd, (S, 1)), np.tile(M.Qd, (S, 1)), status=st, options=slp.SlpOptions(max_iter=100))
    res = drv.run()
    radial = 1 + [tuple(b) for b in M.br.tolist()].index((7, 8))
    connected = np.array([s != radial for s in range(S)])
    assert np.all(res["status"][connected] == 0), res["status"]
    assert res["status"][radial] != 0
    assert res["obj"][0] == pytest.approx(146.7, rel=0.1)
    assert np.all(res["obj"][connected] >= res["obj"][0] - 0.05) and np.all(res["obj"][connected] < 148.0)
    g = res["g"]
    viol = np.maximum(np.maximum(g - drv.g_U, drv.g_L - g), 0.0).max(axis=1)
    assert np.all(viol[connected] <= 0.01 + 1e-9)
    tm = res["timing"]
    assert tm["lp_calls"] > 0 and tm["gpu_eval_s"] > 0.0
    drv.close()
    model.close()


def test_case300_gpu_backend_reproduces_the_cpu_checker_iterate_by_iterate():
    """BASELINE config 5 on the reference's own examples/opf/case300.m: base case + three single-branch outages, 12
    lock-step iterations of the line-search SLP.  The GPU backend (NL rows, affine / quadratic rows, objective, CSC assembly,
    KKT / merit metrics: all on the device) and the test-only CPU backend built on the oracles drive the same host
    algorithm; objective, step length, infeasibility and merit derivative agree iteration by iteration."""
    from oracle_backend import OracleBackend
    M = copy.deepcopy(load_fixture("case300"))
    S = 4
    st = np.ones((S, M.nbr), dtype=np.uint8)
    for s, k in enumerate((5, 170, 300), start=1):
        st[s, k] = 0
    opts = slp.SlpOptions(max_iter=12)
    model = ps.create_opf_model(M, formulation=ps.PNPAPVmodel, obj_type="quadratic")
    gpu = slp.BatchedSlpLS(model, S, np.tile(M.Pd, (S, 1)), np.tile(M.Qd, (S, 1)), status=st, options=opts)
    a = gpu.run()
    cpu = slp.BatchedSlpLS(None, S, backend=OracleBackend(model, S, status=st), options=opts)
    b = cpu.run()
    assert len(a["history"]) == len(b["history"]) >= 12        # lock-step iterations (a scenario's own count stops at max_iter)
    for ha, hb in zip(a["history"], b["history"]):
        np.testing.assert_allclose(ha["f"], hb["f"], rtol=1e-7, atol=1e-6)
        np.testing.assert_allclose(ha["alpha"], hb["alpha"], rtol=1e-9)
        np.testing.assert_allclose(ha["prim_infeas"], hb["prim_infeas"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(ha["D"], hb["D"], rtol=1e-5, atol=1e-6)
    # (x itself is not compared: the common angle shift is a null direction of the model -- no reference bus is fixed -- and
    # the LP's choice along it depends on the last bits of its input)
    np.testing.assert_allclose(a["obj"], b["obj"], rtol=1e-7)
    np.testing.assert_allclose(a["g"], b["g"], rtol=1e-4, atol=1e-4)      # 23 LP steps amplify the last-bit input differences
    tm = a["timing"]
    assert tm["lp_calls"] >= 12 * S and tm["trial_batches"] > 12
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["iter"], b["iter"])
    gpu.close(); cpu.close(); model.close()
