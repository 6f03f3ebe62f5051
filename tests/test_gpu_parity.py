"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(include/powersense_b200.h via powersense.jl_b200/capi.py), against the CPU oracle and the committed
golden vectors.  Tolerance is BASELINE.json's: values within 1e-12 relative / 1e-10 absolute in FP64;
sparsity structures and index arrays bit-exact."""
import numpy as np
import pytest

import powersense_b200 as ps
from powersense_b200 import capi, synth
from oracle import c_oracle

from cases import FIXTURE_CASES, load_fixture, load_golden, synthetic, tol_err, tol_err_scaled

pytestmark = pytest.mark.gpu

FORMS = ps.ALL_FORMULATIONS


def _torch():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a GPU"
    return torch


def _case(name):
    return load_fixture(name) if name in FIXTURE_CASES else synthetic(name)


def _batch_eval(h, x, mu, Pd=None, Qd=None, status=None, what=7, scal=True):
    torch = _torch()
    dev = torch.device("cuda:0")
    S = x.shape[0]
    b = capi.Batch(h, S)
    if Pd is not None:
        b.set_loads(Pd, Qd)
    if status is not None:
        b.set_branch_status(status)
    tx, tmu = torch.from_numpy(x).to(dev), torch.from_numpy(mu).to(dev)
    g = torch.full((S, h.m_nl), np.nan, dtype=torch.float64, device=dev)
    J = torch.full((S, h.nnzJ), np.nan, dtype=torch.float64, device=dev)
    H = torch.full((S, h.nnzH), np.nan, dtype=torch.float64, device=dev)
    sc = torch.full((S, capi.PS_NSCALARS), np.nan, dtype=torch.float64, device=dev) if scal else None
    b.eval_torch(what, tx, tmu, g, J, H, sc)
    torch.cuda.synchronize()
    b.close()
    return g.cpu().numpy(), J.cpu().numpy(), H.cpu().numpy(), (sc.cpu().numpy() if scal else None)


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
@pytest.mark.parametrize("case", FIXTURE_CASES)
def test_golden_vectors_single_scenario(case, F):
    """The MOI-style host-pointer callbacks against the committed golden vectors (jump_oracle)."""
    d = load_fixture(case)
    z = load_golden(case, F)
    h = capi.Handle(d, F.code, device=0)
    assert h.n_var == int(z["n_var"])
    jr, jc = h.jacobian_structure()
    hr, hc = h.hessian_structure()
    assert np.array_equal(jr, z["jrow"]) and np.array_equal(jc, z["jcol"])
    assert np.array_equal(hr, z["hrow"]) and np.array_equal(hc, z["hcol"])
    lo, hi = h.constraint_bounds()
    assert np.array_equal(lo, z["lo"]) and np.array_equal(hi, z["hi"])
    g = h.eval_constraint(z["x"])
    J = h.eval_jacobian(z["x"])
    H = h.eval_hessian(z["x"], 1.0, z["mu"])
    gs, Js, Hs = c_oracle.COracle(d, F.code).condition_scale(z["x"], z["mu"])
    assert tol_err_scaled(g, z["g"], gs) <= 1.0
    assert tol_err_scaled(J, z["J"], Js) <= 1.0
    assert tol_err_scaled(H, z["H"], Hs) <= 1.0
    g2, J2, H2 = h.eval_all(z["x"], z["mu"])
    assert np.array_equal(g, g2) and np.array_equal(J, J2) and np.array_equal(H, H2)
    h.close()


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
@pytest.mark.parametrize("case", ["case14", "case300", "case118", "arpae30", "mp40_out", "hub"])
def test_batch_vs_oracle(case, F):
    d = _case(case)
    S = 5
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=7)
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(3).normal(size=(S, h.m_nl))
    g, J, H, sc = _batch_eval(h, x, mu, Pd, Qd)
    o = c_oracle.COracle(d, F.code)
    og, oJ, oH = o.eval_batch(x, mu, Pd, Qd)
    assert not np.isnan(g).any() and not np.isnan(J).any() and not np.isnan(H).any()
    gsmax = np.zeros(S)
    for s in range(S):
        gs, Js, Hs = o.condition_scale(x[s], mu[s], Pd[s], Qd[s])
        gsmax[s] = gs.max()
        assert tol_err_scaled(g[s], og[s], gs) <= 1.0
        assert tol_err_scaled(J[s], oJ[s], Js) <= 1.0
        assert tol_err_scaled(H[s], oH[s], Hs) <= 1.0
    # scalars: max violation of the NL rows (common.jl:76-98 with p = Inf) and count above tolerance
    lo, hi = o.bounds()
    viol = np.maximum(np.maximum(og - hi, lo - og), 0.0)
    assert tol_err_scaled(sc[:, 1], viol.max(axis=1), gsmax) <= 1.0
    safe = np.abs(viol - 1e-6) > 1e-12 * gsmax[:, None] + 1e-9        # rows not sitting on the tolerance
    assert np.array_equal(sc[:, 2], (viol > 1e-6).sum(axis=1)) or not safe.all()
    # objective (formulations.jl:122-138, obj_type = :linear -> sum(cg))
    off = h.variable_offsets()
    assert tol_err(sc[:, 0], x[:, off["cg"]:off["cg"] + d.ngen].sum(axis=1)) <= 1.0
    # batched == looped single-scenario, bitwise (network loads for the single-scenario path)
    g1, J1, H1, _ = _batch_eval(h, x, mu)
    for s in range(2):
        a, b, c = h.eval_all(x[s], mu[s])
        assert np.array_equal(a, g1[s]) and np.array_equal(b, J1[s]) and np.array_equal(c, H1[s])
    h.close()


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
@pytest.mark.parametrize("case", ["wellcond", "wellcond_arpae", "case3"])
def test_strict_tolerance_on_well_conditioned_networks(case, F):
    """BASELINE.json's tolerance as stated -- 1e-12 relative / 1e-10 absolute against the value itself --
    where the reference formulas are well conditioned (branch |y| of order 10, not 1e4)."""
    d = _case(case)
    S = 6
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=23)
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(4).normal(size=(S, h.m_nl))
    g, J, H, _ = _batch_eval(h, x, mu, Pd, Qd)
    og, oJ, oH = c_oracle.COracle(d, F.code).eval_batch(x, mu, Pd, Qd)
    assert tol_err(g, og) <= 1.0 and tol_err(J, oJ) <= 1.0 and tol_err(H, oH) <= 1.0
    h.close()


@pytest.mark.parametrize("F", [ps.PBRARVmodel, ps.PNPAPVmodel, ps.CBRARVmodel, ps.PNRARVmodel], ids=lambda f: f.name)
def test_quadratic_objective_and_flags(F):
    d = load_fixture("case300")
    for facts, lin in ((True, False), (False, True), (False, False)):
        h = capi.Handle(d, F.code, facts_bshunt=facts, obj_linear=lin, device=0)
        o = c_oracle.COracle(d, F.code, facts_bshunt=facts, obj_linear=lin)
        assert (h.n_var, h.m_nl, h.nnzJ, h.nnzH) == (o.n_var, o.m_nl, o.nnzJ, o.nnzH)
        x, Pd, Qd, lay = synth.scenario_batch(d, F, 3, seed=2, facts_bshunt=facts, obj_linear=lin)
        mu = np.random.default_rng(5).normal(size=(3, h.m_nl))
        g, J, H, sc = _batch_eval(h, x, mu, Pd, Qd)
        og, oJ, oH = o.eval_batch(x, mu, Pd, Qd)
        for s in range(3):
            gs, Js, Hs = o.condition_scale(x[s], mu[s], Pd[s], Qd[s])
            assert max(tol_err_scaled(g[s], og[s], gs), tol_err_scaled(J[s], oJ[s], Js), tol_err_scaled(H[s], oH[s], Hs)) <= 1.0
        pg = x[:, lay["pg"]:lay["pg"] + d.ngen]
        if lin:
            obj = x[:, lay["cg"]:lay["cg"] + d.ngen].sum(axis=1)
        else:                                   # formulations.jl:128-133 (cost_order == 3 for case300)
            obj = (d.c2 * pg * pg + d.c1 * pg + d.c0).sum(axis=1)
        assert np.allclose(sc[:, 0], obj, rtol=1e-12, atol=1e-9)
        h.close()


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
def test_contingency_status_mask(F):
    """Per-scenario branch outages on the fixed pattern equal the oracle run on the network rebuilt with
    that branch out of service (PowersenseData.jl:198-207: tables scaled by br_status, Imax 0 -> 10000; the
    branch leaves Sij/Sji, GO_read2.jl:92-93), after dropping the structurally absent slots."""
    net = synth.synth_network(40, 70, 9, seed=21, frac_parallel=0.1)
    d = ps.create_PowersenseData(net)
    S = 4
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=9)
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(1).normal(size=(S, h.m_nl))
    out = [3, 17, 40, 69]
    st = np.ones((S, d.nbr), dtype=np.uint8)
    for s, k in enumerate(out):
        st[s, k] = 0
    g, J, H, _ = _batch_eval(h, x, mu, Pd, Qd, status=st)
    jr, jc = h.jacobian_structure()
    hr, hc = h.hessian_structure()
    import copy
    for s, k in enumerate(out):
        net2 = copy.deepcopy(net)
        net2.br_status = net.br_status.copy()
        net2.br_status[k] = 0
        d2 = ps.create_PowersenseData(net2)
        o2 = c_oracle.COracle(d2, F.code)
        g2, J2, H2 = o2.eval(x[s], mu[s], Pd[s], Qd[s])
        gs2, _, _ = o2.condition_scale(x[s], mu[s], Pd[s], Qd[s])
        assert tol_err_scaled(g[s], g2, gs2) <= 1.0
        # compare as dense-by-key maps: slots absent from the outage structure must be zero on the GPU
        r2, c2, hr2, hc2 = o2.structure()
        ref = {}
        _, Js2, Hs2 = o2.condition_scale(x[s], mu[s], Pd[s], Qd[s])
        jscale = {(r, c): v for r, c, v in zip(r2, c2, Js2)}
        for r, c, v in zip(r2, c2, J2):
            ref[(r, c)] = v
        for r, c, v in zip(jr, jc, J[s]):
            assert abs(v - ref.get((r, c), 0.0)) <= 1e-10 + 1e-12 * max(abs(v), jscale.get((r, c), 0.0)), (F.name, s, r, c, v, ref.get((r, c)))
        # Hessian of the Lagrangian is linear in mu: compare the weighted sum per variable pair
        # (duplicates across row blocks summed, which is what a solver does with the COO triplets)
        ref, acc, hscale = {}, {}, {}
        for a, b, v, sc_ in zip(hr2, hc2, H2, Hs2):
            ref[(a, b)] = ref.get((a, b), 0.0) + v
            hscale[(a, b)] = hscale.get((a, b), 0.0) + sc_
        for a, b, v in zip(hr, hc, H[s]):
            acc[(a, b)] = acc.get((a, b), 0.0) + v
        for key, v in acc.items():
            assert abs(v - ref.get(key, 0.0)) <= 1e-9 + 1e-12 * max(abs(v), hscale.get(key, 0.0)), (F.name, s, key, v, ref.get(key))
        for key, v in ref.items():
            assert key in acc or abs(v) <= 1e-12
    h.close()


def test_view_offset_semantics():
    """The MOI wrapper hands the evaluator views into larger arrays (MOI_wrapper.jl:967,1025,1059-1060):
    outputs are written at the given pointer and nothing else is touched."""
    d = load_fixture("case14")
    F = ps.PNPAPVmodel
    z = load_golden("case14", F)
    h = capi.Handle(d, F.code, device=0)
    off = 37
    big = np.full(off + h.m_nl + 5, -7.0)
    h.eval_constraint(z["x"], big[off:])
    assert np.all(big[:off] == -7.0) and np.all(big[off + h.m_nl:] == -7.0)
    assert tol_err(big[off:off + h.m_nl], z["g"]) <= 1.0
    bigJ = np.full(off + h.nnzJ + 5, -7.0)
    h.eval_jacobian(z["x"], bigJ[off:])
    assert np.all(bigJ[:off] == -7.0) and np.all(bigJ[off + h.nnzJ:] == -7.0)
    lam = np.concatenate([np.full(off, 3.0), z["mu"]])
    bigH = np.full(off + h.nnzH + 5, -7.0)
    h.eval_hessian(z["x"], 1.0, lam[off:], bigH[off:])
    assert np.all(bigH[:off] == -7.0) and np.all(bigH[off + h.nnzH:] == -7.0)
    assert tol_err(bigH[off:off + h.nnzH], z["H"]) <= 1.0
    h.close()


def test_errors_and_partial_what():
    d = load_fixture("case14")
    F = ps.PBRARVmodel
    h = capi.Handle(d, F.code, device=0)
    torch = _torch()
    dev = torch.device("cuda:0")
    x, Pd, Qd, _ = synth.scenario_batch(d, F, 3, seed=1)
    tx = torch.from_numpy(x).to(dev)
    b = capi.Batch(h, 3)
    J = torch.full((3, h.nnzJ + 3), np.nan, dtype=torch.float64, device=dev)    # padded leading dimension
    b.eval_torch(capi.EVAL_J, tx, None, None, J, None, None)
    torch.cuda.synchronize()
    o = c_oracle.COracle(d, F.code)
    _, oJ, _ = o.eval_batch(x, None, None, None, what=2)
    Jh = J.cpu().numpy()
    assert tol_err(Jh[:, :h.nnzJ], oJ) <= 1.0 and np.isnan(Jh[:, h.nnzJ:]).all()
    with pytest.raises(capi.PowersenseError) as e:                               # Hessian without mu
        b.eval_torch(capi.EVAL_H, tx, None, None, None, torch.empty(3, h.nnzH, dtype=torch.float64, device=dev), None)
    assert e.value.code == capi.PS_ERR_ARG
    with pytest.raises(capi.PowersenseError):                                    # range outside the batch
        b.eval_torch(capi.EVAL_J, tx, None, None, J, None, None, s_begin=2)
    # sub-range of the batch uses that range's loads
    b.set_loads(Pd, Qd)
    g = torch.empty(2, h.m_nl, dtype=torch.float64, device=dev)
    b.eval_torch(capi.EVAL_G, tx[1:], None, g, None, None, None, s_begin=1)
    torch.cuda.synchronize()
    og, _, _ = o.eval_batch(x[1:], None, Pd[1:], Qd[1:], what=1)
    assert tol_err(g.cpu().numpy(), og) <= 1.0
    b.close()
    h.close()


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
def test_partial_what_matches_full_evaluation(F):
    """eval_constraint / eval_constraint_jacobian / eval_hessian_lagrangian are separate callbacks (the SLP line search
    only asks for g, the LP set-up for g + J): every subset of {g, J, H}, with and without the scalars, with and without a
    status mask and on a sub-range of the batch, writes exactly the requested blocks of the full evaluation -- bitwise
    -- and leaves the others untouched.  Covers the warp-tile bus kernel and the slot-parallel PB* / CB* kernels (whose
    Jacobian depends on the scratch the residual kernel leaves)."""
    torch = _torch()
    dev = torch.device("cuda:0")
    d = synthetic("case118")
    S = 6
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=5)
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(6).normal(size=(S, h.m_nl))
    tx, tmu = torch.from_numpy(x).to(dev), torch.from_numpy(mu).to(dev)
    nr = len(d.br)
    for masked in (False, True):
        b = capi.Batch(h, S)
        b.set_loads(Pd, Qd)
        if masked:
            st = np.ones((S, nr), dtype=np.uint8)
            for s in range(S):
                st[s, (7 * s + 3) % nr] = 0
            b.set_branch_status(st)

        def run(what, scal, lo=0):
            n = S - lo
            g = torch.full((n, h.m_nl), np.nan, dtype=torch.float64, device=dev)
            J = torch.full((n, h.nnzJ), np.nan, dtype=torch.float64, device=dev)
            H = torch.full((n, h.nnzH), np.nan, dtype=torch.float64, device=dev)
            sc = torch.full((n, capi.PS_NSCALARS), np.nan, dtype=torch.float64, device=dev) if scal else None
            b.eval_torch(what, tx[lo:], tmu[lo:] if what & 4 else None, g if what & 1 else None, J if what & 2 else None,
                         H if what & 4 else None, sc, s_begin=lo)
            torch.cuda.synchronize()
            return g.cpu().numpy(), J.cpu().numpy(), H.cpu().numpy(), None if sc is None else sc.cpu().numpy()
        full = run(7, True)
        assert not any(np.isnan(a).any() for a in full)
        for what in (1, 2, 3, 4, 5, 6, 7):
            for scal in (False, True):
                for lo in (0, 2):
                    out = run(what, scal, lo)
                    for bit, a, r in ((1, out[0], full[0]), (2, out[1], full[1]), (4, out[2], full[2])):
                        if what & bit:
                            assert np.array_equal(a, r[lo:]), (F.name, masked, what, scal, lo, bit)
                        else:
                            assert np.isnan(a).all(), (F.name, masked, what, scal, lo, bit)
                    if scal:
                        assert np.array_equal(out[3], full[3][lo:]), (F.name, masked, what, lo, "scalars")
        out = run(0, True)                                   # scalars only
        assert np.array_equal(out[3], full[3]) and all(np.isnan(a).all() for a in out[:3])
        b.close()
    h.close()


def test_host_streaming_matches_device_path():
    d = synthetic("case118")
    F = ps.PBRARVmodel
    S = 37
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=4)
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(8).normal(size=(S, h.m_nl))
    g0, J0, H0, sc0 = _batch_eval(h, x, mu, Pd, Qd)
    import os
    os.environ["PS_STAGE"] = "8"                    # force several pipeline chunks, incl. a ragged last one
    try:
        b = capi.Batch(h, S)
        b.set_loads(Pd, Qd)
        px = capi.pinned_empty((S, h.n_var)); px[:] = x
        pmu = capi.pinned_empty((S, h.m_nl)); pmu[:] = mu
        g, J, H = capi.pinned_empty((S, h.m_nl)), capi.pinned_empty((S, h.nnzJ)), capi.pinned_empty((S, h.nnzH))
        sc = capi.pinned_empty((S, capi.PS_NSCALARS))
        b.eval_host(7, px, pmu, g, J, H, sc)
        assert np.array_equal(g, g0) and np.array_equal(J, J0) and np.array_equal(H, H0) and np.array_equal(sc, sc0)
        # plain (pageable) numpy buffers work too
        g2 = np.zeros((S, h.m_nl))
        b.eval_host(1, x, None, g2, None, None, None)
        assert np.array_equal(g2, g0)
        b.close()
    finally:
        del os.environ["PS_STAGE"]
    h.close()


@pytest.mark.parametrize("F,case,S", [(ps.PBRARVmodel, "case13659pegase", 24), (ps.PNPAPVmodel, "case2869pegase", 48),
                                      (ps.CBRARVmodel, "case13659pegase", 8), (ps.PNPAPVmodel, "case9241pegase", 8)],
                         ids=lambda v: getattr(v, "name", str(v)))
def test_full_size_cases(F, case, S):
    """BASELINE.json's full network sizes: a few scenarios against the oracle plus size-independent
    properties (H linear in mu; J of the flow-definition rows has +1 on the flow variable; identical
    scenarios give identical rows)."""
    d = synthetic(case)
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=13)
    x[S - 1] = x[0]; Pd[S - 1] = Pd[0]; Qd[S - 1] = Qd[0]
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(6).normal(size=(S, h.m_nl))
    mu[S - 1] = mu[0]
    g, J, H, sc = _batch_eval(h, x, mu, Pd, Qd)
    o = c_oracle.COracle(d, F.code)
    n_chk = 3
    og, oJ, oH = o.eval_batch(x[:n_chk], mu[:n_chk], Pd[:n_chk], Qd[:n_chk])
    for s in range(n_chk):
        gs, Js, Hs = o.condition_scale(x[s], mu[s], Pd[s], Qd[s])
        assert tol_err_scaled(g[s], og[s], gs) <= 1.0 and tol_err_scaled(J[s], oJ[s], Js) <= 1.0 and tol_err_scaled(H[s], oH[s], Hs) <= 1.0
    assert np.array_equal(g[0], g[S - 1]) and np.array_equal(J[0], J[S - 1]) and np.array_equal(H[0], H[S - 1])
    _, _, H2, _ = _batch_eval(h, x, 2.0 * mu, Pd, Qd, what=4, scal=False)
    assert np.array_equal(H2, 2.0 * H)             # linearity in mu (scaling by 2 is exact in FP64)
    h.close()


def test_config5_n1_contingency_sweep_full_size():
    """BASELINE config 5 (evaluator part): case9241pegase PNPAPV, one scenario per outaged branch on the fixed
    pattern.  A few scenarios are checked against the oracle run on the network rebuilt with that branch out of service
    (g exactly; J / H through the rows / pairs the two structures share); size-independent properties for the rest:
    the outaged branch's limit rows equal -Imax0^2 with Imax0 = 10000 (PowersenseData.jl:206-207), its Jacobian
    and Hessian slots vanish, and every other branch row is bitwise equal to the no-outage evaluation."""
    import copy
    from powersense_b200.driver import contingency_status
    from powersense_b200.synth import NAMED_CASES, synth_network
    nb, nr, ng = NAMED_CASES["case9241pegase"]
    net = synth_network(nb, nr, ng, seed=nb)
    d = ps.create_PowersenseData(net)
    F = ps.PNPAPVmodel
    S = 12
    x, Pd, Qd, _ = synth.scenario_batch(d, F, 1, seed=3)
    x, Pd, Qd = np.repeat(x, S, 0), np.repeat(Pd, S, 0), np.repeat(Qd, S, 0)       # same point, different outages
    h = capi.Handle(d, F.code, device=0)
    mu = np.repeat(np.random.default_rng(2).normal(size=(1, h.m_nl)), S, 0)
    st = contingency_status(d.nbr, s_begin=4000, s_count=S)
    out = [int(np.argmin(r)) for r in st]
    g0, J0, H0, _ = _batch_eval(h, x[:1], mu[:1], Pd[:1], Qd[:1])
    g, J, H, _ = _batch_eval(h, x, mu, Pd, Qd, status=st)
    jr, jc = h.jacobian_structure()
    info = h.plan_info()
    br_rows = 2 * d.nbus
    for s, k in enumerate(out):
        rows = np.array([br_rows + 2 * k, br_rows + 2 * k + 1])
        assert np.all(g[s, rows] == -1.0e8)                                        # 0 - 10000^2
        jsl = slice(info["j_br0"] + 8 * k, info["j_br0"] + 8 * k + 8)
        hsl = slice(info["h_br0"] + 20 * k, info["h_br0"] + 20 * k + 20)
        assert np.all(J[s, jsl] == 0.0) and np.all(H[s, hsl] == 0.0)
        keep = np.ones(d.nbr, bool); keep[k] = False
        assert np.array_equal(g[s, br_rows:].reshape(-1, 2)[keep], g0[0, br_rows:].reshape(-1, 2)[keep])
        assert np.array_equal(J[s, info["j_br0"]:].reshape(-1, 8)[keep], J0[0, info["j_br0"]:].reshape(-1, 8)[keep])
        # only the two end buses of the outaged branch see different balance rows
        f, t = int(d.br[k][0]) - 1, int(d.br[k][1]) - 1
        same = np.ones(d.nbus, bool); same[[f, t]] = False
        assert np.array_equal(g[s, :br_rows].reshape(-1, 2)[same], g0[0, :br_rows].reshape(-1, 2)[same])
    for s in (0, 5):                                                               # oracle on the rebuilt network
        net2 = copy.deepcopy(net)
        net2.br_status = net.br_status.copy(); net2.br_status[out[s]] = 0
        d2 = ps.create_PowersenseData(net2)
        o2 = c_oracle.COracle(d2, F.code)
        g2, J2, H2 = o2.eval(x[s], mu[s], Pd[s], Qd[s])
        gs2, _, _ = o2.condition_scale(x[s], mu[s], Pd[s], Qd[s])
        assert tol_err_scaled(g[s], g2, gs2) <= 1.0
    h.close()


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
def test_edge_networks(F):
    """Degenerate inputs: a network without branches, an isolated bus without generator or load, a bus with two
    generators, a single scenario, and a scenario count that does not divide the scenario chunk."""
    from powersense_b200.data import Network
    # (a) three buses, no branch at all
    net = synth.synth_network(6, 5, 3, seed=1)
    net0 = Network(**{**net.__dict__})
    for k in ("f", "t", "br_r", "br_x", "g_fr", "b_fr", "g_to", "b_to", "tap", "shift", "br_status", "rate_a", "gfM", "bfM"):
        setattr(net0, k, getattr(net, k)[:0])
    d0 = ps.create_PowersenseData(net0)
    # (b) bus 6 isolated (its only branches removed), two generators on bus 1
    net1 = synth.synth_network(6, 7, 3, seed=2)
    keep = (net1.f != 6) & (net1.t != 6)
    for k in ("f", "t", "br_r", "br_x", "g_fr", "b_fr", "g_to", "b_to", "tap", "shift", "br_status", "rate_a", "gfM", "bfM"):
        setattr(net1, k, getattr(net1, k)[keep])
    net1.gen_bus[:] = [1, 1, 3]
    d1 = ps.create_PowersenseData(net1)
    for d, S in ((d0, 1), (d1, 1), (d1, 19)):
        h = capi.Handle(d, F.code, device=0)
        o = c_oracle.COracle(d, F.code)
        assert (h.n_var, h.m_nl, h.nnzJ, h.nnzH) == (o.n_var, o.m_nl, o.nnzJ, o.nnzH)
        x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=3)
        mu = np.random.default_rng(1).normal(size=(S, max(h.m_nl, 1)))[:, :h.m_nl]
        if h.m_nl == 0:                                    # PNRAWV / PBRAWV without branches: nothing to evaluate
            h.close()
            continue
        g, J, H, sc = _batch_eval(h, x, mu, Pd, Qd)
        og, oJ, oH = o.eval_batch(x, mu, Pd, Qd)
        assert tol_err(g, og) <= 1.0 and tol_err(J, oJ) <= 1.0 and tol_err(H, oH) <= 1.0
        h.close()


@pytest.mark.parametrize("F,case", [(ps.PBRARVmodel, "case13659pegase"), (ps.PBRAPVmodel, "case2869pegase"),
                                    (ps.PNPAPVmodel, "case2869pegase"), (ps.CBRARVmodel, "case2869pegase"),
                                    (ps.PNRARVmodel, "case2869pegase"), (ps.PNRAPVmodel, "case118"),
                                    (ps.CBRAWVmodel, "case118")],
                         ids=lambda v: getattr(v, "name", str(v)))
def test_all_code_paths_agree_bitwise(F, case):
    """The bulk (TMA store), direct (register -> sector) and staged branch kernels, the slot-parallel PB* bus kernels, the
    warp-tile bus kernel and the CTA-tile contribution-major bus kernel evaluate the same expressions in the same order: their outputs are bitwise
    equal, on the preferred (sector-aligned) layout, on a 16-byte aligned one and on a plain one, and two runs are
    bitwise reproducible."""
    import os
    torch = _torch()
    dev = torch.device("cuda:0")
    d = synthetic(case)
    S = 9
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=31)
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(9).normal(size=(S, h.m_nl))
    b = capi.Batch(h, S)
    b.set_loads(Pd, Qd)
    tx, tmu = torch.from_numpy(x).to(dev), torch.from_numpy(mu).to(dev)
    (ldg, ldJ, ldH), (pg, pJ, pH) = h.preferred_layout()

    def run(layout):
        def rows(width, ld, pad):
            if layout == "plain":      # odd leading dimension: rows are only 8-byte aligned -> staged branch kernel
                return torch.full((S, width + 1 + width % 2), np.nan, dtype=torch.float64, device=dev)[:, :width]
            if layout == "pair":       # ld = 2 mod 4: 16-byte aligned blocks, sectors are not -> bulk (TMA) branch kernel only
                ld2 = ld + 2
                flat = torch.full((pad + S * ld2,), np.nan, dtype=torch.float64, device=dev)
                return flat[pad:].view(S, ld2)[:, :width]
            flat = torch.full((pad + S * ld,), np.nan, dtype=torch.float64, device=dev)
            return flat[pad:].view(S, ld)[:, :width]
        g, J, H = rows(h.m_nl, ldg, pg), rows(h.nnzJ, ldJ, pJ), rows(h.nnzH, ldH, pH)
        sc = torch.empty(S, capi.PS_NSCALARS, dtype=torch.float64, device=dev)
        b.eval_torch(7, tx, tmu, g, J, H, sc)
        torch.cuda.synchronize()
        return g.cpu().numpy(), J.cpu().numpy(), H.cpu().numpy(), sc.cpu().numpy()
    ref = run("sector")
    assert all(np.array_equal(a, c) for a, c in zip(ref, run("sector")))              # reproducible
    variants = [({}, "plain"), ({}, "pair"), ({"PS_BULK": "1"}, "sector"), ({"PS_BULK": "1"}, "pair"),
                ({"PS_NO_BULK": "1"}, "sector"), ({"PS_NO_BULK": "1"}, "pair"),
                ({"PS_NO_BULK": "1", "PS_NO_DIRECT": "1"}, "sector"), ({"PS_NO_PBFAST": "1"}, "sector"),
                ({"PS_NO_BULK": "1", "PS_NO_DIRECT": "1", "PS_NO_PBFAST": "1"}, "plain"), ({"PS_CHUNK": "1"}, "sector"),
                ({"PS_CHUNK": "5"}, "sector"),
                # bus rows: the slot-parallel CB* kernels (default for CB*) against the warp-tile kernel (default otherwise)
                # against the CTA-tile kernel, and the warp-tile kernel's scenario chunking
                ({"PS_NO_CBFAST": "1"}, "sector"), ({"PS_NO_CBFAST": "1"}, "plain"),
                ({"PS_NO_CBFAST": "1", "PS_NO_WARPBUS": "1"}, "sector"), ({"PS_NO_CBFAST": "1", "PS_NO_WARPBUS": "1"}, "plain"),
                ({"PS_NO_CBFAST": "1", "PS_NO_WARPBUS": "1", "PS_NO_PBFAST": "1"}, "pair"),
                ({"PS_NO_CBFAST": "1", "PS_CHUNK_WT": "1"}, "sector"),
                ({"PS_NO_CBFAST": "1", "PS_CHUNK_WT": "5", "PS_NO_PBFAST": "1"}, "plain"), ({"PS_CHUNK_PB": "3"}, "pair")]
    for env, layout in variants:
        os.environ.update(env)
        try:
            out = run(layout)
        finally:
            for k in env:
                del os.environ[k]
        for name, a, c in zip("gJHs", ref, out):
            assert np.array_equal(a, c), (F.name, env, layout, name)
    b.close()
    h.close()


@pytest.mark.parametrize("F", [ps.PBRARVmodel, ps.PBRAPVmodel], ids=lambda v: v.name)
@pytest.mark.parametrize("order", ["random", "file"])
def test_scenario_resident_pb_residuals(F, order):
    """pb_bus_g_scn_kernel (bus sums of a whole scenario in shared memory, terms visited in element order with coalesced
    loads) is bitwise equal to the list-walking pb_bus_g_kernel: on a randomly numbered network (ranks 0-1), on a from-bus
    ordered one (every step has chains of equal buses: one sub-round per rank), with a contingency mask, for g alone and
    for g + J + H, with more scenarios than CTAs and with fewer."""
    import os
    torch = _torch()
    dev = torch.device("cuda:0")
    d = synth.named_case("case2869pegase", branch_order=order)
    h = capi.Handle(d, F.code, device=0)
    for S, masked in ((333, False), (40, True)):
        x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=5)
        mu = np.random.default_rng(3).normal(size=(S, h.m_nl))
        b = capi.Batch(h, S)
        b.set_loads(Pd, Qd)
        if masked:
            st = np.ones((S, d.nbr), dtype=np.uint8)
            st[np.arange(S), (np.arange(S) * 37) % d.nbr] = 0
            b.set_branch_status(st)
        tx, tmu = torch.from_numpy(x).to(dev), torch.from_numpy(mu).to(dev)

        def run(what):
            g = torch.full((S, h.m_nl), np.nan, dtype=torch.float64, device=dev)
            J = torch.full((S, h.nnzJ), np.nan, dtype=torch.float64, device=dev)
            H = torch.full((S, h.nnzH), np.nan, dtype=torch.float64, device=dev)
            sc = torch.empty(S, capi.PS_NSCALARS, dtype=torch.float64, device=dev)
            b.eval_torch(what, tx, tmu, g, J, H, sc)
            torch.cuda.synchronize()
            return g.cpu().numpy(), J.cpu().numpy(), H.cpu().numpy(), sc.cpu().numpy()
        for what in (1, 7):
            os.environ["PS_PBSCN"] = "1"
            try:
                new = run(what)                               # S >= 32: the scenario-resident kernel (the default where the plan offers it)
            finally:
                del os.environ["PS_PBSCN"]
            os.environ["PS_NO_PBSCN"] = "1"
            try:
                ref = run(what)
            finally:
                del os.environ["PS_NO_PBSCN"]
            for name, a, c in zip("gJHs", ref, new):
                assert np.array_equal(a, c, equal_nan=True), (F.name, order, S, what, name)
            assert not np.isnan(new[0]).any()
        b.close()
    h.close()


@pytest.mark.parametrize("F", [ps.PNPAPVmodel, ps.PBRARVmodel, ps.CBRARVmodel, ps.PNRAWVmodel], ids=lambda v: v.name)
def test_hessian_order_hook_round_trips_bitwise(F):
    """With a caller-imposed slot order (ps_match_hessian_structure) every entry point -- the single-scenario callback, the
    batched device path (several slabs), the host-streaming path -- returns exactly canonical[perm]; the kernels are not
    involved (they keep writing the canonical order).  Resetting restores the canonical output."""
    import os
    from test_capi_cpu import _jump_like_order
    torch = _torch()
    dev = torch.device("cuda:0")
    d = synthetic("case118")
    S = 7
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=3)
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(1).normal(size=(S, h.m_nl))
    g0, J0, H0, _ = _batch_eval(h, x, mu, Pd, Qd)
    perm, tr, tc = _jump_like_order(h, 11)
    h.match_hessian_structure(tr, tc)
    g1, J1, H1, _ = _batch_eval(h, x, mu, Pd, Qd)
    assert np.array_equal(g1, g0) and np.array_equal(J1, J0) and np.array_equal(H1, H0[:, perm])
    b = capi.Batch(h, S)
    b.set_loads(Pd, Qd)
    hx, hmu = capi.pinned_empty((S, h.n_var)), capi.pinned_empty((S, h.m_nl))
    hx[:], hmu[:] = x, mu
    hH = capi.pinned_empty((S, h.nnzH))
    os.environ["PS_STAGE"] = "3"
    try:
        b.eval_host(4, hx, hmu, None, None, hH, None)
    finally:
        del os.environ["PS_STAGE"]
    assert np.array_equal(np.asarray(hH), H0[:, perm])
    b.close()
    h2 = capi.Handle(d, F.code, device=0)          # network loads: the single-scenario callbacks use them
    Hc = h2.eval_hessian(x[0], 1.0, mu[0]).copy()
    h2.set_hessian_order(perm + 1)
    assert np.array_equal(h2.eval_hessian(x[0], 1.0, mu[0]), Hc[perm])
    g, J, Hp = h2.eval_all(x[0], mu[0])
    assert np.array_equal(Hp, Hc[perm])
    h2.set_hessian_order(None)
    assert np.array_equal(h2.eval_hessian(x[0], 1.0, mu[0]), Hc)
    h2.close()
    h.set_hessian_order(None)
    assert np.array_equal(_batch_eval(h, x, mu, Pd, Qd)[2], H0)
    h.close()


def test_single_scenario_cache_keyed_on_x():
    """JuMP's evaluator keeps its forward pass while x is unchanged.  Here: eval_constraint at a new x runs
    ONE fused g + J launch, eval_constraint_jacobian at the same x is served from the device without a launch,
    eval_hessian_lagrangian (whose multipliers arrive with the call) costs one more; a g-only caller (the SLP line search)
    stops paying for J after its first point; results equal the uncached path bitwise."""
    import os
    d = synthetic("case118")
    F = ps.PNPAPVmodel
    h = capi.Handle(d, F.code, device=0)
    x, _, _, _ = synth.scenario_batch(d, F, 3, seed=8)
    mu = np.random.default_rng(2).normal(size=h.m_nl)
    os.environ["PS_NO_CACHE"] = "1"
    try:
        g_ref = [h.eval_constraint(x[k]).copy() for k in range(3)]
        J_ref = [h.eval_jacobian(x[k]).copy() for k in range(3)]
        H_ref = h.eval_hessian(x[0], 1.0, mu).copy()
    finally:
        del os.environ["PS_NO_CACHE"]
    per_eval = 2                                    # PNPAPV: bus kernel + branch kernel per evaluation
    h.eval_constraint(x[0]); h.eval_jacobian(x[0])  # a caller that asks for g and J at a point (Ipopt, the SLP's LP set-up) ...
    n0 = h.launch_count()
    g = h.eval_constraint(x[1])                     # ... gets both from ONE fused evaluation at the next new x
    n1 = h.launch_count()
    J = h.eval_jacobian(x[1])                       # same x: no launch
    n2 = h.launch_count()
    H = h.eval_hessian(x[1], 1.0, mu)               # multipliers arrive now: one evaluation (H only), no x upload
    n3 = h.launch_count()
    g2 = h.eval_constraint(x[1])                    # still cached
    n4 = h.launch_count()
    assert n1 - n0 == per_eval and n2 == n1 and n3 - n2 == per_eval and n4 == n3
    assert np.array_equal(g, g_ref[1]) and np.array_equal(J, J_ref[1]) and np.array_equal(g2, g_ref[1])
    assert np.array_equal(h.eval_hessian(x[0], 1.0, mu), H_ref)
    xx = x[2].copy()                                # a caller that mutates its x in place is seen (the key is the content)
    assert np.array_equal(h.eval_constraint(xx), g_ref[2])
    xx[:] = x[0]
    assert np.array_equal(h.eval_constraint(xx), g_ref[0]) and np.array_equal(h.eval_jacobian(xx), J_ref[0])
    # g-only caller: after one point at which J was not asked for, a new x evaluates g alone
    h.eval_constraint(x[1]); h.eval_constraint(x[2])
    a = h.launch_count()
    gg = h.eval_constraint(x[0])
    Jl = h.eval_jacobian(x[0])                      # asked after all: a second launch, J only
    assert np.array_equal(gg, g_ref[0]) and np.array_equal(Jl, J_ref[0]) and h.launch_count() - a == 2 * per_eval
    h.close()


def test_callbacks_into_page_locked_caller_arrays():
    """ps_host_register: the single-scenario callbacks write into arrays the caller page-locked (a solver's own g / J / H
    vectors, views into larger arrays included) -- same values as into pageable ones; unregistering twice is an error."""
    d = synthetic("case118")
    F = ps.PBRARVmodel
    h = capi.Handle(d, F.code, device=0)
    x, _, _, _ = synth.scenario_batch(d, F, 2, seed=4)
    mu = np.random.default_rng(5).normal(size=h.m_nl)
    g0, J0, H0 = (a.copy() for a in h.eval_all(x[0], mu))
    big = np.full(h.m_nl + h.nnzJ + h.nnzH + 7, np.nan)
    g, J, H = big[3:3 + h.m_nl], big[3 + h.m_nl:3 + h.m_nl + h.nnzJ], big[5 + h.m_nl + h.nnzJ:5 + h.m_nl + h.nnzJ + h.nnzH]
    with capi.host_registered(big):
        h.eval_all(x[1], mu, g, J, H)               # another point first: nothing comes from the cache
        h.eval_all(x[0], mu, g, J, H)
        assert np.array_equal(g, g0) and np.array_equal(J, J0) and np.array_equal(H, H0)
        assert np.isnan(big[:3]).all() and np.isnan(big[-2:]).all()
    assert capi.lib().ps_host_unregister(big.ctypes.data_as(capi._VP)) != capi.PS_OK
    h.eval_all(x[0], mu, g, J, H)                   # pageable again
    assert np.array_equal(H, H0)
    h.close()


def test_correctly_rounded_powers():
    """The L4 limit rows hold V^4 and V^3 (formulations.jl:338-343); the reference evaluates them through pow() (JuMP's `^`
    with a constant exponent other than 2).  The kernel's powers are the correctly rounded ones: on a network whose L4 rows
    reduce to V_a^4 - 1 (g_s = 1, everything else 0, Imax = 1) g + 1 and J / 4 must equal the exactly rounded rational
    powers, bit for bit (repeated products miss that in 25-50 % of the arguments; glibc's own pow() in 0.08 %)."""
    import copy
    from fractions import Fraction
    d = copy.deepcopy(load_fixture("case3"))
    for name in ("gf", "gt"):
        getattr(d, name)[:] = 1.0
    for name in ("bf", "bt", "Gf", "Bf", "Gt", "Bt"):
        getattr(d, name)[:] = 0.0
    d.Imax[:] = 1.0
    F = ps.PNPAPVmodel
    h = capi.Handle(d, F.code, device=0)
    off = h.variable_offsets()
    jr, jc = h.jacobian_structure()
    rng = np.random.default_rng(12)
    S = 2000
    x, _, _, _ = synth.scenario_batch(d, F, S, seed=1)
    V = rng.uniform(0.85, 1.18, size=(S, d.nbus))
    x[:, off["b"]:off["b"] + d.nbus] = V
    mu = np.zeros((S, h.m_nl))
    g, J, _, _ = _batch_eval(h, x, mu, what=3)
    bad4 = bad3 = 0
    for k in range(d.nbr):
        for side in range(2):
            a = int(d.br[k][side]) - 1
            row = 2 * d.nbus + 2 * k + side
            e = int(np.flatnonzero((jr == row + 1) & (jc == off["b"] + a + 1))[0])
            for s in range(0, S, 7):
                v = Fraction(float(V[s, a]))
                bad4 += (g[s, row] + 1.0) != float(v ** 4)
                bad3 += (J[s, e] / 4.0) != float(v ** 3)
    assert bad4 == 0 and bad3 == 0, (bad4, bad3)
    h.close()


@pytest.mark.parametrize("F", [ps.PNPAPVmodel, ps.PNRAPVmodel, ps.PNRARVmodel, ps.CBRARVmodel, ps.PBRARVmodel], ids=lambda v: v.name)
def test_specialised_callback_kernels_match_the_general_one(F):
    """bus_warp_kernel is instantiated for the callbacks a solver actually issues -- residuals only (eval_constraint: every
    trial point of the SLP line search; no derivative is formed) and g + J (the LP set-up; no Hessian term is formed) -- next
    to the run-time-flag version that serves odd combinations, each kind also under a branch-status mask.  Same expressions,
    same order: bitwise equal."""
    import os
    torch = _torch()
    dev = torch.device("cuda:0")
    d = synthetic("case2869pegase")
    S = 37
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=21)
    h = capi.Handle(d, F.code, device=0)
    mu = np.random.default_rng(2).normal(size=(S, h.m_nl))
    base = {"PS_NO_CBFAST": "1", "PS_NO_PBFAST": "1"}           # the warp-tile kernel for every formulation
    st = np.ones((S, d.nbr), dtype=np.uint8)                    # an N-1 mask: the three kinds are also instantiated with one
    st[np.arange(S), (np.arange(S) * 53) % d.nbr] = 0
    os.environ.update(base)
    try:
        for status in (None, st):
            for what in (1, 3, 7, 2, 5):
                new = _batch_eval(h, x, mu, Pd, Qd, status=status, what=what)
                os.environ["PS_WT_ANY"] = "1"
                try:
                    ref = _batch_eval(h, x, mu, Pd, Qd, status=status, what=what)
                finally:
                    del os.environ["PS_WT_ANY"]
                for name, a, c in zip("gJHs", new, ref):
                    assert np.array_equal(a, c, equal_nan=True), (F.name, what, name, status is not None)
    finally:
        for k in base:
            del os.environ[k]
    h.close()
