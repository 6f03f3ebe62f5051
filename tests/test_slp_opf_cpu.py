"""The SLP driver's host logic on real OPF models, in the CPU suite: powersense_b200.slp.BatchedSlpLS (algorithm, LP
subproblems, ConvexFeasibleInitialization, contingency handling) with the test-only CPU evaluation backend of
tests/oracle_backend.py.  The product's backend is GpuBackend; tests/test_gpu_slp.py checks that the two produce the same
iterates."""
import copy

import numpy as np
import pytest

import powersense_b200 as ps
from powersense_b200 import slp, synth

from cases import load_fixture
from oracle_backend import OracleBackend


def test_case14_known_answer_and_contingencies_with_the_cpu_backend():
    """The reference's own SLP assertion (test/opf/ACOPF_formulations.jl:9: PNPAPVmodel, obj_type = :linear, 146.7 at
    rtol 0.1) plus two N-1 outages on the fixed pattern."""
    M = copy.deepcopy(load_fixture("case14"))
    model = ps.create_opf_model(M, formulation=ps.PNPAPVmodel, obj_type="linear", device=-2)     # structure-only: no GPU here
    S = 3
    st = np.ones((S, M.nbr), dtype=np.uint8)
    st[1, 2] = 0
    st[2, 9] = 0
    drv = slp.BatchedSlpLS(None, S, backend=OracleBackend(model, S, status=st), options=slp.SlpOptions(max_iter=80))
    res = drv.run()
    assert np.all(res["status"] == 0), res["status"]
    assert res["obj"][0] == pytest.approx(146.7, rel=0.1) and res["obj"][0] == pytest.approx(146.77, rel=5e-3)
    assert np.all(res["obj"][1:] >= res["obj"][0] - 0.05)
    g = res["g"]
    assert np.all(np.maximum(np.maximum(g - drv.g_U, drv.g_L - g), 0.0).max(axis=1) <= 0.01 + 1e-9)
    drv.close()


def test_convex_feasible_initialization_is_the_l1_projection_onto_the_affine_rows():
    """convex_initialization! (slp.jl:23-48, model of MOI_wrapper.jl:1121-1145): from a start that violates generator
    bounds the LP returns the nearest point (L1) inside every affine row, touching only the offending entries."""
    M = copy.deepcopy(load_fixture("case14"))
    model = ps.create_opf_model(M, formulation=ps.PNPAPVmodel, obj_type="quadratic", device=-2)
    be = OracleBackend(model, 2)
    drv = slp.BatchedSlpLS(None, 2, backend=be, options=slp.SlpOptions())
    lay = model.lay
    x0 = np.tile(model.x0, (2, 1))
    x0[:, lay["V"]:lay["V"] + M.nbus] = 1.0
    x0 = drv._convex_initialization(x0)                      # a start inside every affine row ...
    assert np.array_equal(drv._convex_initialization(x0.copy()), x0)      # ... is left alone
    x0[0, lay["pg"]] = M.Pmax[0] + 3.0                       # above its upper bound
    x0[1, lay["V"] + 4] = M.Vmin[4] - 0.2                    # below the voltage bound
    x1 = drv._convex_initialization(x0.copy())
    A, cst, lo, hi = be.linear_rows()
    for s in range(2):
        v = A @ x1[s] + cst
        assert np.all(v <= hi + 1e-9) and np.all(v >= lo - 1e-9)
    assert x1[0, lay["pg"]] == pytest.approx(M.Pmax[0]) and x1[1, lay["V"] + 4] == pytest.approx(M.Vmin[4])
    changed = np.abs(x1 - x0) > 1e-12
    assert changed[0].sum() == 1 and changed[1].sum() == 1
    drv.close()


def test_feasible_synthetic_case_is_feasible_and_solvable():
    """synth.named_case(..., feasible=True): the operating point the case is built around satisfies every NL row and every
    affine row, and the SLP solves the OPF from it (BASELINE config 5 needs a solvable large case; the default synthetic
    loads are random and in general not a solvable OPF)."""
    d = synth.named_case("case118", feasible=True)
    model = ps.create_opf_model(d, formulation=ps.PNPAPVmodel, obj_type="quadratic", device=-2)
    be = OracleBackend(model, 1)
    import torch
    X = torch.from_numpy(model.x0[None, :].copy())
    E, f = torch.zeros(1, be.m, dtype=torch.float64), torch.zeros(1, 1, dtype=torch.float64)
    be.eval_g(X, E, f)
    e = E.numpy()[0]
    assert np.maximum(np.maximum(e - be.g_U, be.g_L - e), 0.0).max() <= 1e-9          # the start is feasible
    drv = slp.BatchedSlpLS(None, 1, backend=be, options=slp.SlpOptions(max_iter=150))
    res = drv.run()
    assert res["status"][0] in (0, 6), res["status"]
    assert res["obj"][0] <= float(f[0, 0]) + 1e-6                                      # the optimiser did not make it dearer
    drv.close()
