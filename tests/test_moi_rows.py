"""N1: affine + quadratic rows of the solver-level model on the device (ps_rows_*) against the CPU restatement of
src/opt/MOI_wrapper.jl:780-800, 834-839, 864-891, 973-1002, 1030-1042."""
import numpy as np
import pytest

from powersense_b200 import capi
from oracle import moi_rows_oracle as mo


def _random_rows(rng, n, m):
    rows = []
    for _ in range(m):
        na, nq = int(rng.integers(0, 6)), int(rng.integers(0, 4))
        aff = [(float(rng.normal()), int(rng.integers(1, n + 1))) for _ in range(na)]
        quad = []
        for _ in range(nq):
            i, j = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
            if rng.random() < 0.4:
                j = i
            quad.append((float(rng.normal()), i, j))
        rows.append((float(rng.normal()), aff, quad))
    rows.append((0.0, [], []))                                   # an empty row
    rows.append((1.5, [(2.0, 3), (4.0, 3)], [(3.0, 2, 2), (3.0, 2, 2)]))   # duplicate terms are kept, not merged
    return rows


def test_structure_matches_reference_rules():
    rng = np.random.default_rng(0)
    rows = _random_rows(rng, 9, 12)
    R = capi.RowsBlock(9, *mo.to_csr(rows)[:7], constant=mo.to_csr(rows)[7], device=-2)
    jr, jc = R.jacobian_structure()
    hr, hc = R.hessian_structure()
    assert list(zip(jr.tolist(), jc.tolist())) == mo.jacobian_structure(rows)
    assert list(zip(hr.tolist(), hc.tolist())) == mo.hessian_structure(rows)
    with pytest.raises(capi.PowersenseError) as e:
        R.eval(np.zeros(9))
    assert e.value.code == capi.PS_ERR_STATE
    with pytest.raises(capi.PowersenseError):
        capi.RowsBlock(3, [0, 1], [7], [1.0], [0, 0], [], [], [], device=-2)       # variable index out of range


@pytest.mark.gpu
def test_values_single_and_batched():
    import torch
    rng = np.random.default_rng(1)
    n = 40
    rows = _random_rows(rng, n, 200)
    csr = mo.to_csr(rows)
    R = capi.RowsBlock(n, *csr[:7], constant=csr[7], device=0)
    S = 5
    X = rng.normal(size=(S, n))
    L = rng.normal(size=(S, R.m))
    g, J, H = R.eval(X[0], L[0])
    assert np.array_equal(g, np.array([mo.eval_function(r, X[0]) for r in rows]))
    assert np.array_equal(J, mo.fill_constraint_jacobian(rows, X[0]))
    assert np.array_equal(H, mo.fill_hessian_lagrangian(rows, L[0]))
    dev = torch.device("cuda:0")
    tx, tl = torch.from_numpy(X).to(dev), torch.from_numpy(L).to(dev)
    tg = torch.full((S, R.m + 1), np.nan, dtype=torch.float64, device=dev)
    tJ = torch.full((S, R.nnzJ + 2), np.nan, dtype=torch.float64, device=dev)
    tH = torch.full((S, R.nnzH), np.nan, dtype=torch.float64, device=dev)
    R.eval_torch(7, tx, tl, tg, tJ, tH)
    torch.cuda.synchronize()
    for s in range(S):
        assert np.array_equal(tg[s, :R.m].cpu().numpy(), np.array([mo.eval_function(r, X[s]) for r in rows]))
        assert np.array_equal(tJ[s, :R.nnzJ].cpu().numpy(), mo.fill_constraint_jacobian(rows, X[s]))
        assert np.array_equal(tH[s].cpu().numpy(), mo.fill_hessian_lagrangian(rows, L[s]))
    assert torch.isnan(tg[:, R.m:]).all() and torch.isnan(tJ[:, R.nnzJ:]).all()
    R.close()


@pytest.mark.gpu
def test_opf_model_rows_of_case14_in_solver_order():
    """The `@constraint` rows of the 14-bus CBRARV model (formulations.jl:123-168, 324-334) in Powersense.Optimizer's
    order [lin_le|lin_ge|lin_eq|quad_le|quad_ge|quad_eq] (MOI_wrapper.jl:767-773), evaluated on the device."""
    import copy
    import powersense_b200 as ps
    from cases import load_fixture
    M = copy.deepcopy(load_fixture("case14"))
    model = ps.create_opf_model(M, formulation=ps.CBRARVmodel, obj_type="linear")
    rows, lo, hi, g_order = model.moi_rows()
    assert sum(g_order) == len(rows) and g_order[3] == g_order[4] == 14 and g_order[5] == 0   # Vmin^2 <= vr^2+vi^2 <= Vmax^2 split
    assert g_order[2] == 3 * 5 + 4 * 20                                                       # cost rows + current definitions
    csr = mo.to_csr(rows)
    R = capi.RowsBlock(model.n, *csr[:7], constant=csr[7], device=0)
    x = np.random.default_rng(3).uniform(0.5, 1.5, model.n)
    lam = np.random.default_rng(4).normal(size=R.m)
    g, J, H = R.eval(x, lam)
    assert np.allclose(g, [mo.eval_function(r, x) for r in rows], rtol=0, atol=0)
    assert np.array_equal(J, mo.fill_constraint_jacobian(rows, x)) and np.array_equal(H, mo.fill_hessian_lagrangian(rows, lam))
    # the same rows in Interval form on the host (what run_opf hands to scipy) give the same values
    v = model.rows_value(x)
    k = g_order[0] + g_order[1] + g_order[2]
    assert np.allclose(np.sort(g[k:k + 14]), np.sort(v[[i for i, q in enumerate(model.rows.quad) if q is not None]]), rtol=1e-14)
    R.close()
    model.close()


@pytest.mark.parametrize("case", ["case14", "case300", "arpae30"])
def test_w_relaxation_rows_agree_with_the_rectangular_models(case):
    """PBRAWV / PNRAWV through create_opf_model (formulations.jl:162-167, 173-296, 317-320 | 381-384, 355-360 | 419-425):
    at a point where the W variables are consistent with the voltages (Wd = vr^2 + vi^2, Wr + j Wi = V_f conj(V_t)... as the
    reference defines them) the W-space rows must reproduce the rectangular nodal model the relaxation came from -- the
    affine balances equal PNRARV's NL balance residuals, PNRAWV's limit rows equal PNRARV's L5 rows (quadratic for
    MATPOWER data, affine for PSS/E data), PBRAWV's flow definitions vanish when the flows are their W-expressions, and the
    three copies of Wd == vr^2 + vi^2 are present.  Structure-only handles: no GPU needed."""
    import copy
    import powersense_b200 as ps
    from powersense_b200 import synth
    from oracle import c_oracle
    from cases import load_fixture, synthetic
    d = copy.deepcopy(load_fixture(case)) if case in ("case14", "case300") else synthetic(case)
    nb, nr = d.nbus, d.nbr
    rng = np.random.default_rng(5)
    xr, Pd, Qd, layr = synth.scenario_batch(d, ps.PNRARVmodel, 1, seed=3, obj_linear=False)
    vi, vr = xr[0, layr["vi"]:layr["vi"] + nb], xr[0, layr["vr"]:layr["vr"] + nb]
    f, t = d.br[:, 0] - 1, d.br[:, 1] - 1
    Wd = vr * vr + vi * vi
    Wr, Wi = vr[f] * vr[t] + vi[f] * vi[t], vr[f] * vi[t] - vi[f] * vr[t]
    o = c_oracle.COracle(d, ps.PNRARVmodel.code, True, False)
    g_ref, _, _ = o.eval(xr[0], None, d.Pd, d.Qd, what=1)            # PNRARV: [P_i, Q_i ...] residuals, then L5 rows
    for F in (ps.PNRAWVmodel, ps.PBRAWVmodel):
        model = ps.create_opf_model(d, formulation=F, obj_type="quadratic", device=-2)
        lay = model.lay
        x = rng.normal(size=model.n)
        x[lay["vi"]:lay["vi"] + nb], x[lay["vr"]:lay["vr"] + nb], x[lay["Wd"]:lay["Wd"] + nb] = vi, vr, Wd
        x[lay["Wr"]:lay["Wr"] + nr], x[lay["Wi"]:lay["Wi"] + nr] = Wr, Wi
        for key in ("pg", "qg"):
            x[lay[key]:lay[key] + d.ngen] = xr[0, layr[key]:layr[key] + d.ngen]
        if d.nbss:
            x[lay["bss"]:lay["bss"] + d.nbss] = xr[0, layr["bss"]:layr["bss"] + d.nbss]
        if F is ps.PBRAWVmodel:
            (Pij, Qij, Pji, Qji), _ = synth.flows_from_voltages(d, vr, vi)
            for key, v in (("Pij", Pij), ("Qij", Qij), ("Pji", Pji), ("Qji", Qji)):
                x[lay[key]:lay[key] + nr] = v
        v = model.rows_value(x)
        lo, hi = model.rows.lo, model.rows.hi
        eq = np.flatnonzero(lo == hi)
        # the rows are in creation order: ... [Wd bounds][Wd == vr^2+vi^2 x 3][per bus: P, Q][per branch rows]
        ng = d.ngen
        n_head = 2 * ng + (d.nbss if True else 0) + nb + nb       # pg, qg, bss bounds, V bounds (quadratic), Wd bounds
        wrows = v[n_head:n_head + 3 * nb]
        assert np.allclose(wrows, 0.0, atol=1e-12) and np.all(lo[n_head:n_head + 3 * nb] == 0.0)          # 3 copies of Wd == vr^2 + vi^2
        bal = slice(n_head + 3 * nb, n_head + 3 * nb + 2 * nb)
        scale = 1.0 + np.abs(g_ref[:2 * nb])
        assert np.allclose((v[bal] - lo[bal]) / scale, g_ref[:2 * nb] / scale, atol=1e-9), (case, F.name)      # balances == PNRARV's NL rows
        br = v[n_head + 5 * nb:]
        if F is ps.PBRAWVmodel:
            assert np.allclose(br[:4 * nr], 0.0, atol=1e-9)                                                # flow definitions
        else:
            lim = br[:2 * nr] - hi[n_head + 5 * nb:n_head + 5 * nb + 2 * nr]
            sc = 1.0 + np.abs(g_ref[2 * nb:])
            assert np.allclose(lim / sc, g_ref[2 * nb:] / sc, atol=1e-7), (case, F.name)                    # limits == L5 rows
            quad = [q is not None for q in model.rows.quad[n_head + 5 * nb:n_head + 5 * nb + 2 * nr]]
            assert all(quad) if d.source_type == "matpower" else not any(quad)
        rows, lo2, hi2, g_order = model.moi_rows()
        assert sum(g_order) == len(rows) and g_order[5] >= 3 * nb
        model.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fname", ["PBRAWVmodel", "PNRAWVmodel"])
def test_w_relaxation_rows_on_the_device(fname):
    """The `@constraint` rows of the two W-relaxation models (affine balances, W-space flow definitions, quadratic
    limits, the tripled Wd rows) in Powersense.Optimizer's order, evaluated by ps_rows_* on the device, bit-exact against
    the statement-by-statement restatement; and the model's NL block (cones, L1 limits) evaluates through the same handle."""
    import copy
    import powersense_b200 as ps
    from cases import load_fixture
    M = copy.deepcopy(load_fixture("case300"))
    model = ps.create_opf_model(M, formulation=ps.BY_NAME[fname], obj_type="linear", box_constraints=True)
    rows, lo, hi, g_order = model.moi_rows()
    csr = mo.to_csr(rows)
    R = capi.RowsBlock(model.n, *csr[:7], constant=csr[7], device=0)
    x = np.random.default_rng(3).uniform(0.5, 1.5, model.n)
    lam = np.random.default_rng(4).normal(size=R.m)
    g, J, H = R.eval(x, lam)
    assert np.array_equal(g, np.array([mo.eval_function(r, x) for r in rows]))
    assert np.array_equal(J, mo.fill_constraint_jacobian(rows, x)) and np.array_equal(H, mo.fill_hessian_lagrangian(rows, lam))
    assert g_order[5] >= 3 * M.nbus
    gnl = model.nl_value(x)
    assert gnl.shape == (M.nbr * (3 if fname == "PBRAWVmodel" else 1),) and np.isfinite(gnl).all()
    R.close()
    model.close()
