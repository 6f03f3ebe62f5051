"""N3: per-scenario KKT / merit metrics on the device against the CPU restatement of common.jl:35-98 and
slp.jl:108-162.  Floating-point sums are reduced in a fixed tree on the device: tolerance 1e-12 relative."""
import numpy as np
import pytest

import powersense_b200 as ps
from powersense_b200 import capi, synth
from oracle import csc_oracle, kkt_oracle

from cases import load_fixture

pytestmark = pytest.mark.gpu


def test_kkt_metrics_on_the_case300_jacobian():
    import torch
    d = load_fixture("case300")
    F = ps.PNRARVmodel
    h = capi.Handle(d, F.code, device=0)
    jr, jc = h.jacobian_structure()
    m, n = h.m_nl, h.n_var
    pat = capi.CscPattern(m, n, jr, jc, device=0)
    lo, hi = h.constraint_bounds()
    rng = np.random.default_rng(0)
    x_L, x_U = np.full(n, -np.inf), np.full(n, np.inf)
    x_L[-50:] = 0.0
    x_U[:40] = 1.0
    km = capi.KktMetrics(pat, lo, hi, x_L, x_U)
    S = 4
    x, Pd, Qd, _ = synth.scenario_batch(d, F, S, seed=6)
    x[:, -50:] = rng.normal(size=(S, 50))                       # some bound violations
    E = np.stack([h.eval_constraint(x[s]) for s in range(S)])
    Jv = np.stack([h.eval_jacobian(x[s]) for s in range(S)])
    lam, nu = rng.normal(size=(S, m)), rng.uniform(1, 3, size=(S, m))
    mU, mL = rng.normal(size=(S, n)) * 0.1, rng.normal(size=(S, n)) * 0.1
    df, pdir, f = rng.normal(size=(S, n)), rng.normal(size=(S, n)), rng.normal(size=S)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nz = torch.empty(S, pat.nnz, dtype=torch.float64, device=dev)
    pat.assemble_torch(T(Jv), nz)
    out = km.eval_torch(T(E), T(x), T(lam), T(mU), T(mL), T(df), nz, T(nu), T(pdir), T(f)).cpu().numpy()
    for s in range(S):
        Jac = csc_oracle.compute_jacobian_matrix(m, n, jr, jc, Jv[s])
        want = [kkt_oracle.norm_violations(E[s], lo, hi, x[s], x_L, x_U, 1),
                kkt_oracle.norm_violations(E[s], lo, hi, x[s], x_L, x_U, np.inf),
                kkt_oracle.norm_complementarity(E[s], lo, hi, lam[s]),
                kkt_oracle.KT_residuals(df[s], lam[s], mU[s], mL[s], Jac),
                kkt_oracle.compute_phi0(f[s], E[s], lo, hi, nu[s]),
                kkt_oracle.compute_derivative(df[s], pdir[s], E[s], lo, hi, nu[s]),
                np.linalg.norm(df[s])]
        assert np.allclose(out[s, :7], want, rtol=1e-12, atol=1e-12), (s, out[s, :7], want)
    assert np.isfinite(out).all()
    # without nu / p: phi = f, D = 0
    out2 = km.eval_torch(T(E), T(x), T(lam), T(mU), T(mL), T(df), nz, None, None, T(f)).cpu().numpy()
    assert np.array_equal(out2[:, 4], f) and np.all(out2[:, 5] == 0.0) and np.array_equal(out2[:, :4], out[:, :4])
    km.close(); pat.close(); h.close()
