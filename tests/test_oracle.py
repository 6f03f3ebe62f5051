"""CPU tests of the oracle (test infrastructure): the C restatement against the committed golden vectors
(expression-tree restatement of JuMP's algorithm), against finite differences, against the closed-form
structure counts of SURVEY.md 8(d), and the cross-formulation identities.  PARITY UNPINNED against the
reference's own evaluator (JuMP 0.21.10 / Julia cannot run here): these tests pin the oracle to two
independent restatements and to first principles instead."""
import numpy as np
import pytest

import powersense_b200 as ps
from powersense_b200 import synth
from oracle import c_oracle, jump_oracle

from cases import FIXTURE_CASES, load_fixture, load_golden, synthetic, tol_err, tol_err_scaled

FORMS = ps.ALL_FORMULATIONS


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
@pytest.mark.parametrize("case", FIXTURE_CASES)
def test_c_oracle_matches_golden(case, F):
    d = load_fixture(case)
    z = load_golden(case, F)
    o = c_oracle.COracle(d, F.code)
    assert o.n_var == int(z["n_var"])
    jr, jc, hr, hc = o.structure()
    assert np.array_equal(jr, z["jrow"]) and np.array_equal(jc, z["jcol"])
    assert np.array_equal(hr, z["hrow"]) and np.array_equal(hc, z["hcol"])
    lo, hi = o.bounds()
    assert np.array_equal(lo, z["lo"]) and np.array_equal(hi, z["hi"])
    g, J, H = o.eval(z["x"], z["mu"])
    gs, Js, Hs = o.condition_scale(z["x"], z["mu"])
    assert tol_err_scaled(g, z["g"], gs) <= 1.0
    assert tol_err_scaled(J, z["J"], Js) <= 1.0
    assert tol_err_scaled(H, z["H"], Hs) <= 1.0


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
@pytest.mark.parametrize("case", ["arpae30", "mp40_out"])
def test_c_oracle_matches_tree_oracle_live(case, F):
    """Synthetic networks with out-of-service elements, parallel branches, switched shunts (bss variables)."""
    d = synthetic(case)
    m = jump_oracle.build_model(d, F.name)
    o = c_oracle.COracle(d, F.code)
    x, _, _, _ = synth.scenario_batch(d, F, 1, seed=3)
    mu = np.random.default_rng(0).normal(size=o.m_nl)
    jr, jc, _ = m.jacobian_structure()
    hr, hc, _ = m.hessian_structure()
    a, b, c, e = o.structure()
    assert np.array_equal(jr, a) and np.array_equal(jc, b) and np.array_equal(hr, c) and np.array_equal(hc, e)
    g, J, H = m.evaluate(x[0], mu)
    g2, J2, H2 = o.eval(x[0], mu)
    gs, Js, Hs = o.condition_scale(x[0], mu)
    assert max(tol_err_scaled(g2, g, gs), tol_err_scaled(J2, J, Js), tol_err_scaled(H2, H, Hs)) <= 1.0


@pytest.mark.parametrize("F", FORMS, ids=lambda f: f.name)
def test_finite_differences(F):
    """J against central differences of g, and the Lagrangian Hessian against central differences of J^T mu."""
    d = synthetic("wellcond")
    o = c_oracle.COracle(d, F.code)
    x, Pd, Qd, _ = synth.scenario_batch(d, F, 1, seed=5)
    x = x[0]
    mu = np.random.default_rng(1).normal(size=o.m_nl)
    g, J, H = o.eval(x, mu)
    jr, jc, hr, hc = o.structure()
    n = o.n_var
    touched = np.unique(jc) - 1
    rng = np.random.default_rng(2)
    cols = rng.choice(touched, size=min(25, len(touched)), replace=False)

    def jt_mu(xx):
        _, JJ, _ = o.eval(xx, mu, what=2)
        out = np.zeros(n)
        np.add.at(out, jc - 1, JJ * mu[jr - 1])
        return out
    Hd = {}
    for a, b, v in zip(hr, hc, H):
        Hd[(a - 1, b - 1)] = Hd.get((a - 1, b - 1), 0.0) + v
    for j in cols:
        h = 1e-6 * max(1.0, abs(x[j]))
        xp, xm = x.copy(), x.copy()
        xp[j] += h
        xm[j] -= h
        gp, _, _ = o.eval(xp, what=1)
        gm, _, _ = o.eval(xm, what=1)
        fd = (gp - gm) / (2 * h)
        col = np.zeros(o.m_nl)
        sel = jc - 1 == j
        col[jr[sel] - 1] = J[sel]
        assert np.allclose(fd, col, rtol=1e-5, atol=1e-5 * (1 + np.abs(col).max()))
        fdh = (jt_mu(xp) - jt_mu(xm)) / (2 * h)
        hcol = np.zeros(n)
        for (a, b), v in Hd.items():           # lower-triangular tuples: mirror
            if b == j:
                hcol[a] += v
            if a == j and a != b:
                hcol[b] += v
        assert np.allclose(fdh, hcol, rtol=1e-4, atol=1e-4 * (1 + np.abs(hcol).max())), F.name


def test_structure_counts_closed_form():
    """m_nl / nnzJ / nnzH equal the closed-form counts of SURVEY.md 8(d) (MATPOWER limit family, nbss = 0)."""
    d = synthetic("case118")
    nb, nr = d.nbus, d.nbr
    Ga = len(d.gen_idx)
    pairs = {(min(a, b), max(a, b)) for (a, b), st in zip(d.br, np.ones(nr)) if True}
    # E' = distinct connected bus pairs among in-service branches (all in service in this case)
    E = len(pairs)
    want = {
        "PBRAPVmodel": (2 * nb + 6 * nr, 2 * Ga + 2 * nb + 28 * nr, 2 * nb + 44 * nr),
        "PBRARVmodel": (2 * nb + 6 * nr, 2 * Ga + 4 * nb + 28 * nr, 4 * nb + 36 * nr),
        "CBRARVmodel": (2 * nb + 2 * nr, 2 * Ga + 4 * nb + 16 * nr, 4 * nb + 34 * nr),
        "CBRAWVmodel": (2 * nb + 2 * nr, 2 * Ga + 6 * nb + 14 * nr, 4 * nb + 26 * nr),
        "PNPAPVmodel": (2 * nb + 2 * nr, 2 * Ga + 4 * nb + 8 * E + 8 * nr, 6 * nb + 28 * E + 20 * nr),
        "PNRAPVmodel": (2 * nb + 2 * nr, 2 * Ga + 4 * nb + 8 * E + 8 * nr, 6 * nb + 28 * E + 20 * nr),
        "PNRARVmodel": (2 * nb + 2 * nr, 2 * Ga + 4 * nb + 8 * E + 8 * nr, 4 * nb + 24 * E + 20 * nr),
        "PBRAWVmodel": (3 * nr, 8 * nr, 9 * nr),
        "PNRAWVmodel": (nr, 4 * nr, 5 * nr),
    }
    for F in FORMS:
        o = c_oracle.COracle(d, F.code)
        assert (o.m_nl, o.nnzJ, o.nnzH) == want[F.name], F.name
        jr, jc, hr, hc = o.structure()
        # rows in order, columns strictly ascending inside a row; Hessian tuples lower triangular
        assert np.all(np.diff(jr) >= 0)
        same = np.diff(jr) == 0
        assert np.all(np.diff(jc)[same] > 0)
        assert np.all(hr >= hc)


def test_case14_dimensions_and_known_parameters():
    """the 14-bus PSS/E fixture through create_PowersenseData."""
    d = load_fixture("case14")
    assert (d.nbus, d.nbr, d.ngen, d.nbss, d.source_type) == (14, 20, 5, 0, "arpae")
    assert np.isclose(d.Pd.sum(), 2.34528)
    assert np.isclose(d.Bl[8], 0.19)
    k = 0                                       # branch 1 (1 -> 2)
    assert np.allclose([d.Gf[k], d.Bf[k], d.gf[k], d.bf[k], d.Imax[k]],
                       [-4.9991316008, 15.2630865232, 4.9991316008, -15.2366865232, 1.416])
    k = 17                                      # branch 18 (4 -> 7, tap 0.978)
    assert np.allclose([d.Gf[k], d.Bf[k], d.bf[k], d.bt[k], d.Imax[k]],
                       [0.0, 4.8895126603, -4.9995016977, -4.7819433818, 0.41], atol=1e-9)
    o = c_oracle.COracle(d, ps.PNPAPVmodel.code)
    assert (o.n_var - int(np.sum(d.lps)) - d.ngen, o.m_nl, o.nnzJ, o.nnzH) == (14 + 14 + 5 + 5, 68, 386, 1044)
    assert o.n_var == 93


def test_cross_formulation_identities():
    """Same physical state in polar and rectangular coordinates: PNPAPV == PNRAPV == PNRARV bus residuals, and the
    PB* flow-definition residuals vanish when the flow variables equal the flow expressions."""
    d = synthetic("wellcond")
    rng = np.random.default_rng(4)
    nb = d.nbus
    V, th = rng.uniform(0.95, 1.05, nb), rng.normal(0, 0.05, nb)
    vr, vi = V * np.cos(th), V * np.sin(th)
    pg, qg = rng.uniform(0, 1, d.ngen), rng.uniform(-0.5, 0.5, d.ngen)
    res = {}
    for F in (ps.PNPAPVmodel, ps.PNRAPVmodel, ps.PNRARVmodel, ps.PBRARVmodel, ps.PBRAPVmodel):
        lay = synth.variable_layout(d, F)
        x = np.zeros(lay["n_var"])
        if "theta" in lay:
            x[lay["theta"]:lay["theta"] + nb], x[lay["V"]:lay["V"] + nb] = th, V
        else:
            x[lay["vi"]:lay["vi"] + nb], x[lay["vr"]:lay["vr"] + nb] = vi, vr
        x[lay["pg"]:lay["pg"] + d.ngen], x[lay["qg"]:lay["qg"] + d.ngen] = pg, qg
        if "Pij" in lay:
            pf, _ = synth.flows_from_voltages(d, vr, vi)
            for key, val in zip(("Pij", "Qij", "Pji", "Qji"), pf):
                x[lay[key]:lay[key] + d.nbr] = val
        g, _, _ = c_oracle.COracle(d, F.code).eval(x, what=1)
        res[F.name] = g
    assert np.allclose(res["PNPAPVmodel"][:2 * nb], res["PNRAPVmodel"][:2 * nb], atol=1e-11)
    assert np.allclose(res["PNPAPVmodel"][:2 * nb], res["PNRARVmodel"][:2 * nb], atol=1e-11)
    assert np.allclose(res["PBRARVmodel"][:2 * nb], res["PNRARVmodel"][:2 * nb], atol=1e-11)   # balances agree
    for name in ("PBRARVmodel", "PBRAPVmodel"):
        flowdef = res[name][2 * nb:].reshape(d.nbr, 6)[:, :4]
        assert np.abs(flowdef).max() < 1e-11


def test_condition_scale_flags_ill_conditioned_limit_rows():
    """The L4 current-limit rows (formulations.jl:338-343) subtract O(|y|^2) products: the oracle's condition
    scale must dwarf the value there, and be comparable to the value on well conditioned rows."""
    d = synthetic("case118")                    # x down to 1e-4 p.u. -> |y|^2 up to 1e8
    F = ps.PNPAPVmodel
    o = c_oracle.COracle(d, F.code)
    x, Pd, Qd, _ = synth.scenario_batch(d, F, 1, seed=1)
    g, J, H = o.eval(x[0], np.ones(o.m_nl), Pd[0], Qd[0])
    gs, Js, Hs = o.condition_scale(x[0], np.ones(o.m_nl), Pd[0], Qd[0])
    lim = slice(2 * d.nbus, None)
    assert (gs[lim] / np.maximum(np.abs(g[lim]), 1e-300)).max() > 1e3


def test_synthetic_branch_orders_are_the_same_network():
    """synth_network(branch_order="file") lists the same branches as the default random permutation, in (mostly)
    ascending from-bus order like the reference's shipped case files."""
    from powersense_b200 import synth
    a = synth.synth_network(300, 411, 69, seed=7, branch_order="random")
    b = synth.synth_network(300, 411, 69, seed=7, branch_order="file")
    assert sorted(zip(a.f.tolist(), a.t.tolist())) == sorted(zip(b.f.tolist(), b.t.tolist()))
    descents = int(np.sum(np.diff(b.f) < 0))
    assert descents <= 0.1 * len(b.f) < int(np.sum(np.diff(a.f) < 0))
