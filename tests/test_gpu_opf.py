"""End-to-end known answers (the numeric answers the reference's tests and examples hold on this path): the 14-bus
PSS/E case solved with obj_type = :linear gives cost ~ 146.7 at rtol 0.1 for PNPAPV, PBRARV + box and CBRARV +- box
(test/opf/ACOPF_formulations.jl:7-9, examples/opf/ACOPF_formulations_example.jl:5-19), and MATPOWER case300 with the
quadratic objective gives 7.1973e+05 (example :23-27).  The NL block is served by the
GPU evaluator through the MOI-style callbacks; the NLP solver is scipy's SLSQP (no Ipopt / GLPK here)."""
import copy

import numpy as np
import pytest

import powersense_b200 as ps

from cases import load_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F,box", [(ps.PNPAPVmodel, False), (ps.PBRARVmodel, True), (ps.CBRARVmodel, True), (ps.CBRARVmodel, False)],
                         ids=["PNPAPV", "PBRARV+box", "CBRARV+box", "CBRARV"])
def test_case14_known_objective(F, box):
    M = copy.deepcopy(load_fixture("case14"))
    res = ps.run_opf(M, formulation=F, box_constraints=box, obj_type="linear")
    model = M.model
    x = res.x
    g = model.nl_value(x)
    viol = np.maximum(np.maximum(g - model.nl_hi, model.nl_lo - g), 0.0).max()
    rv = model.rows_value(x)
    viol = max(viol, np.maximum(np.maximum(rv - model.rows.hi, model.rows.lo - rv), 0.0).max())
    assert viol < 1e-5, viol
    assert M.cost == pytest.approx(146.7, rel=0.1), M.cost          # the reference's own assertion
    # solution is written back like run_opf! does (formulations.jl:608-651)
    assert M.Pg.shape == (5,) and np.all(M.Pg >= M.Pmin - 1e-6) and np.all(M.Pg <= M.Pmax + 1e-6)
    if F is ps.PNPAPVmodel:
        assert np.all(M.V >= 0.9 - 1e-6) and np.all(M.V <= 1.1 + 1e-6)
    model.close()


def test_case300_known_objective():
    """The reference's second example (examples/opf/ACOPF_formulations_example.jl:23-27): MATPOWER case300, PNPAPV,
    obj_type = :quadratic, quoted there as 7.1973e+05 (MATPOWER's published AC-OPF optimum of the case is 719725.1).
    Dense SLSQP on 738 variables: about a minute of host time."""
    M = copy.deepcopy(load_fixture("case300"))
    res = ps.run_opf(M, formulation=ps.PNPAPVmodel, obj_type="quadratic", maxiter=300)
    model = M.model
    g = model.nl_value(res.x)
    viol = np.maximum(np.maximum(g - model.nl_hi, model.nl_lo - g), 0.0).max()
    rv = model.rows_value(res.x)
    viol = max(viol, np.maximum(np.maximum(rv - model.rows.hi, model.rows.lo - rv), 0.0).max())
    assert viol < 1e-5, viol
    assert M.cost == pytest.approx(7.1973e5, rel=1e-4), M.cost
    model.close()


def test_moi_method_set():
    """The evaluator implements exactly the methods of EmptyNLPEvaluator (src/opt/MOI_wrapper.jl:56-77)."""
    d = load_fixture("case14")
    ev = ps.GpuNLPEvaluator(d, ps.PNPAPVmodel)
    assert set(ev.features_available()) == {"Grad", "Jac", "Hess"}
    with pytest.raises(RuntimeError):
        ev.hessian_lagrangian_structure()                   # before initialize([:Hess])
    ev.initialize(["Grad", "Jac", "Hess"])
    js, hs = ev.jacobian_structure(), ev.hessian_lagrangian_structure()
    assert len(js) == 386 and len(hs) == 1044 and ev.num_constraints == 68 and ev.num_variables == 93
    assert all(isinstance(t, tuple) and len(t) == 2 and t[0] >= 1 and t[1] >= 1 for t in js[:5])
    x = np.random.default_rng(0).uniform(0.9, 1.1, ev.num_variables)
    g, J, H = np.empty(68), np.empty(386), np.empty(1044)
    ev.eval_constraint(g, x)
    ev.eval_constraint_jacobian(J, x)
    ev.eval_hessian_lagrangian(H, x, 1.0, np.ones(68))
    assert np.isfinite(g).all() and np.isfinite(J).all() and np.isfinite(H).all()
    with pytest.raises(RuntimeError):
        ev.eval_objective(x)
    lo, hi = ev.constraint_bounds()
    assert np.all(lo[:28] == 0) and np.all(hi == 0) and np.all(np.isinf(lo[28:]))
    ev.close()
