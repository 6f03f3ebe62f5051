"""N4 pinned on the reference's own SLP test: the toy NLP of test/opt/ext_solver.jl (and examples/opt/opt_example.jl),
which test/runtests.jl:19-25 solves with Powersense.Optimizer + GLPK and checks as X = Y = -1 (rtol 1e-4),
LOCALLY_SOLVED.  The product's lock-step SLP driver (powersense_b200.slp.BatchedSlpLS: the algorithm and the LP
subproblem) runs here unchanged; only its evaluation backend is replaced by a test-only CPU one built on the oracle
(the product has no CPU evaluation path: the default backend needs the CUDA library and a GPU)."""
import numpy as np
import pytest
import torch

from oracle import kkt_oracle
from powersense_b200 import capi, slp


class HostToyBackend:
    """min X^2 + X  s.t.  X^2 - X == 2,  X*Y == 1,  X*Y >= 0,  X >= -2   (ext_solver.jl:31-38).
    Row order as Powersense.Optimizer sees it (MOI_wrapper.jl:780-800): the affine row, then the NL rows."""

    def __init__(self, S):
        self.S, self.dev, self.n, self.m = S, torch.device("cpu"), 2, 4
        inf = np.inf
        self.g_L = np.array([-2.0, 2.0, 1.0, 0.0])
        self.g_U = np.array([inf, 2.0, 1.0, inf])
        self.xl, self.xu, self.x0 = np.full(2, -inf), np.full(2, inf), np.zeros(2)
        # dense 4x2 Jacobian in CSC order, 1-based like ps_csc_structure
        self.colptr = np.array([1, 5, 9], dtype=np.int64)
        self.rowval = np.array([1, 2, 3, 4, 1, 2, 3, 4], dtype=np.int64)
        self.nnz = 8

    @staticmethod
    def _g(x, y):
        return np.array([x, x * x - x, x * y, x * y])

    def eval_g(self, X, E, f):
        for s in range(self.S):
            x, y = X[s].tolist()
            E[s] = torch.from_numpy(self._g(x, y))
            f[s, 0] = x * x + x

    def eval_all(self, X, E, f, df, nz):
        self.eval_g(X, E, f)
        for s in range(self.S):
            x, y = X[s].tolist()
            df[s] = torch.tensor([2 * x + 1, 0.0], dtype=torch.float64)
            nz[s] = torch.tensor([1.0, 2 * x - 1, y, y, 0.0, 0.0, x, x], dtype=torch.float64)

    def metrics(self, E, X, lam, mU, mL, df, nz, nu, P, f, out):
        for s in range(self.S):
            Jac = nz[s].numpy().reshape(2, 4).T
            e, x = E[s].numpy(), X[s].numpy()
            out[s, 0] = kkt_oracle.norm_violations(e, self.g_L, self.g_U, x, self.xl, self.xu, 1)
            out[s, 1] = kkt_oracle.norm_violations(e, self.g_L, self.g_U, x, self.xl, self.xu, np.inf)
            out[s, 2] = kkt_oracle.norm_complementarity(e, self.g_L, self.g_U, lam[s].numpy())
            out[s, 3] = kkt_oracle.KT_residuals(df[s].numpy(), lam[s].numpy(), mU[s].numpy(), mL[s].numpy(), Jac)
            fs = float(f[s])
            out[s, 4] = kkt_oracle.compute_phi0(fs, e, self.g_L, self.g_U, nu[s].numpy()) if nu is not None else fs
            out[s, 5] = (kkt_oracle.compute_derivative(df[s].numpy(), P[s].numpy(), e, self.g_L, self.g_U, nu[s].numpy())
                         if nu is not None and P is not None else 0.0)
        return out

    def close(self):
        pass


@pytest.mark.parametrize("S", [1, 3])
def test_reference_toy_problem_known_answer(S):
    be = HostToyBackend(S)
    drv = slp.BatchedSlpLS(None, S, backend=be, options=slp.SlpOptions())
    x0 = np.zeros((S, 2))
    if S > 1:                                  # scenarios started elsewhere advance in lock-step to the same point
        x0[1] = [-0.5, -0.5]
        x0[2] = [-2.0, -3.0]
    out = drv.run(x0)
    assert (out["status"] == 0).all(), out["status"]                     # LOCALLY_SOLVED
    np.testing.assert_allclose(out["x"][0], -1.0, rtol=1e-4)             # test/runtests.jl:20-21 (the reference's start, X = Y = 0)
    np.testing.assert_allclose(out["x"], -1.0, rtol=1e-3)                # the other starts stop inside tol_infeas = 0.01 of it
    np.testing.assert_allclose(out["obj"], 0.0, atol=1e-4)
    drv.close()


def test_first_iterations_follow_the_reference_statement_by_statement():
    """A hand trace of slp_line_search.jl:127-249 on the toy problem from the reference's start X = Y = 0.
    Iteration 1: the LP is infeasible (row X*Y == 1 has a zero gradient and residual 1) -> feasibility restoration, no
    step, iter stays 1 (:163-176).  Iteration 2, restoration LP in the reference's parametrisation (subproblem.jl:404-500):
    rows violated from below get b' = b - |viol|, i.e. right-hand sides 4 (X^2 - X == 2) and 2 (X*Y == 1) and slack
    bounds -2 / -1; with X >= -2 the optimum is slack_sum = 4 (degenerate in p); phi = sum of violations = 3 (slp.jl:112-119),
    D = slack_sum - sum of violations = 1 (:140-149).  compute_nu!(::SlpLS) (slp_line_search.jl:285-295): iter == 1 ->
    nu = |lambda|; afterwards nu = max(nu, |lambda|) -- never the row-norm rule of the generic method."""
    drv = slp.BatchedSlpLS(None, 1, backend=HostToyBackend(1), options=slp.SlpOptions())
    out = drv.run(np.zeros((1, 2)))
    h = out["history"]
    assert h[0]["restoration"][0] and not h[0]["lp_solved"][0] and h[0]["iters"][0] == 1 and h[1]["iters"][0] == 1
    assert h[1]["restoration"][0] and h[1]["lp_solved"][0]
    assert h[1]["slack_sum"][0] == pytest.approx(4.0) and h[1]["phi"][0] == pytest.approx(3.0) and h[1]["D"][0] == pytest.approx(1.0)
    np.testing.assert_array_equal(h[1]["nu"][0], np.abs(h[1]["lambda"][0]))          # iter == 1: nu = |lambda|
    for a, b in zip(h[1:], h[2:]):
        if b["lp_solved"][0]:
            np.testing.assert_array_equal(b["nu"][0], np.maximum(a["nu"][0], np.abs(b["lambda"][0])))
    # Armijo on the merit function with the reference's eta / tau (parameters.jl): alpha = tau^k, first k that passes
    for r in h[1:]:
        k = np.log(r["alpha"][0]) / np.log(0.9)
        assert abs(k - round(k)) < 1e-9
    assert out["status"][0] == 0
    drv.close()
    # the textbook elastic LP (slacks >= 0) gives slack_sum = 1 and D = -2 at the same point: the reference's is different
    drv = slp.BatchedSlpLS(None, 1, backend=HostToyBackend(1), options=slp.SlpOptions(restoration_lp="standard"))
    h = drv.run(np.zeros((1, 2)))["history"]
    assert h[1]["slack_sum"][0] == pytest.approx(1.0) and h[1]["D"][0] == pytest.approx(-2.0)
    drv.close()


def test_pooled_lps_reproduce_the_serial_run():
    """The LP subproblems of one lock-step iteration are independent; solved by worker processes (SlpOptions.lp_workers)
    they give exactly the iterates of the serial run."""
    S = 3
    x0 = np.array([[0.0, 0.0], [-0.5, -0.5], [-2.0, -3.0]])
    runs = []
    for workers in (0, 2):
        drv = slp.BatchedSlpLS(None, S, backend=HostToyBackend(S), options=slp.SlpOptions(lp_workers=workers))
        runs.append(drv.run(x0))
        drv.close()
    a, b = runs
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["iter"], b["iter"])
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["lambda"], b["lambda"])
    assert b["timing"]["lp_calls"] == a["timing"]["lp_calls"] > 0
