"""Host side of the batched SLP: the LP subproblem of one scenario (the reference's slack-augmented LP,
src/opt/algorithms/subproblem.jl:55-334, and its dual extraction :657-700), solved with scipy's HiGHS where the
reference's tests use GLPK.  The LPs of the scenarios of one lock-step iteration are independent: `LpPool` spreads them
over worker processes (numpy / scipy only -- the workers never touch CUDA)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


@dataclass
class LpStatic:                        # everything of the LP that does not change between iterations and scenarios
    n: int
    m: int
    colptr: np.ndarray                 # CSC pattern of the constraint Jacobian, 1-based (SparseMatrixCSC)
    rowval: np.ndarray
    i_le: np.ndarray
    i_ge: np.ndarray
    i_eq: np.ndarray
    g_L: np.ndarray
    g_U: np.ndarray
    xl: np.ndarray
    xu: np.ndarray
    delta: float
    tol_error: float
    restoration: str = "reference"     # "reference": the parametrisation of subproblem.jl:404-500; "standard": slacks >= 0


def solve_lp(st: LpStatic, nzval, E, df, x, restoration):
    """Normal phase: min df'p  s.t.  g_L <= E + J p <= g_U, max(-Delta, x_L - x) <= p <= min(Delta, x_U - x) -- the
    reference fixes every slack at zero there (subproblem.jl:520-540).

    Feasibility restoration (subproblem.jl:369-500): min sum(slacks) over the elastic rows
        le  J p - u     <= c_ub - b'      ge  J p + u >= c_lb - b'      eq  J p + u - v = c - b'
    in the reference's own parametrisation (st.restoration == "reference"): with viol_i the signed violation of row i at
    the current point (:405-411), b' = b - |viol| (:412) and the slacks bounded below by -|viol| instead of 0 -- u for the
    one-sided rows (:478-495), v when an equality is violated from above, u when from below (:415-476).  For rows violated
    from ABOVE this is the standard elastic LP with its slack shifted by |viol|; for rows violated from BELOW the shift
    has the wrong sign in the reference (b - |viol| moves b further away), so the LP asks for three times the
    linearised correction there.  Replicated as is -- the iterates are meant to be comparable with
    slp_line_search.jl:80-249; "standard" gives the textbook LP (slacks >= 0, b unshifted).  The slack sum returned is the
    reference's (evaluate_total_slack!, slp.jl:209-214: shifted slacks, it can be negative), which is what
    compute_derivative (slp.jl:140-143) and the exit test slack_sum <= tol_infeas (slp_line_search.jl:197) consume.
    Returns (p, lambda, mu_upper, mu_lower, slack_sum, status)."""
    n, m = st.n, st.m

    J = sp.csc_matrix((nzval, st.rowval - 1, st.colptr - 1), shape=(m, n)).tocsr()
    le, ge, eq = st.i_le, st.i_ge, st.i_eq
    nle, nge, neq = len(le), len(ge), len(eq)
    ns = (nle + nge + 2 * neq) if restoration else 0
    I = sp.identity
    Zs = lambda r, c: sp.csr_matrix((r, c))
    ub_blocks, eq_block = [], None
    if restoration:        # columns: [p | s_le | s_ge | s1_eq | s2_eq]
        if nle:
            ub_blocks.append(sp.hstack([J[le], -I(nle), Zs(nle, nge + 2 * neq)]))              # A p - s <= c_ub - b
        if nge:
            ub_blocks.append(sp.hstack([-J[ge], Zs(nge, nle), -I(nge), Zs(nge, 2 * neq)]))     # A p + s >= c_lb - b
        if neq:
            eq_block = sp.hstack([J[eq], Zs(neq, nle + nge), I(neq), -I(neq)], format="csr")   # A p + s1 - s2 = c_lb - b
    else:
        if nle:
            ub_blocks.append(J[le])
        if nge:
            ub_blocks.append(-J[ge])
        if neq:
            eq_block = J[eq]
    A_ub = sp.vstack(ub_blocks, format="csr") if ub_blocks else None
    b_ub = np.concatenate([st.g_U[le] - E[le], -(st.g_L[ge] - E[ge])]) if ub_blocks else None
    b_eq = st.g_L[eq] - E[eq] if neq else None
    slack_lb = np.zeros(ns)
    if restoration and st.restoration == "reference":
        # signed violation (:405-411) and the shift b' = b - |viol| (:412): every right-hand side c - b' grows by |viol|
        v_le = np.maximum(E[le] - st.g_U[le], 0.0)                 # |viol| of the rows violated from above
        v_ge = np.maximum(st.g_L[ge] - E[ge], 0.0)                 # ... from below
        d_eq = E[eq] - st.g_L[eq]                                  # > 0: above, < 0: below
        a_eq = np.abs(d_eq)
        if ub_blocks:
            b_ub = b_ub + np.concatenate([v_le, -v_ge])            # (ge rows are stored negated)
        if neq:
            b_eq = b_eq + a_eq
        slack_lb = np.concatenate([-v_le, -v_ge, np.where(d_eq > 0, 0.0, -a_eq), np.where(d_eq > 0, -a_eq, 0.0)])
    if st.tol_error > 0:
        if b_ub is not None:
            b_ub[np.abs(b_ub) <= st.tol_error] = 0.0
        if b_eq is not None:
            b_eq[np.abs(b_eq) <= st.tol_error] = 0.0
    c = np.concatenate([np.zeros(n) if restoration else df, np.ones(ns)])
    ub = np.minimum(st.delta, st.xu - x)                      # subproblem.jl:126-129
    lb = np.maximum(-st.delta, st.xl - x)
    bounds = np.column_stack([np.concatenate([lb, slack_lb]), np.concatenate([ub, np.full(ns, np.inf)])])
    res = linprog(c, A_ub=A_ub, b_ub=b_ub, A_eq=eq_block, b_eq=b_eq, bounds=bounds, method="highs")
    if res.status == 0:
        p = res.x[:n]
        lam = np.zeros(m)                                       # subproblem.jl:673-679 (MOI dual sign convention)
        if ub_blocks:
            lam[le] = res.ineqlin.marginals[:nle]
            lam[ge] = -res.ineqlin.marginals[nle:]
        if neq:
            lam[eq] = res.eqlin.marginals
        mU, mL = res.upper.marginals[:n].copy(), res.lower.marginals[:n].copy()
        mU[p < st.xu - x] = 0.0                               # :686-693: the trust region is not a variable bound
        mL[p > st.xl - x] = 0.0
        return p, lam, mU, mL, float(res.x[n:].sum()), "OPTIMAL"
    if res.status == 2:
        return np.zeros(n), np.zeros(m), np.zeros(n), np.zeros(n), 0.0, "INFEASIBLE"
    return np.zeros(n), np.zeros(m), np.zeros(n), np.zeros(n), 0.0, "OTHER"



_STATIC = None


def _init_worker(static: LpStatic):
    global _STATIC
    _STATIC = static


def _solve_in_worker(args):
    return solve_lp(_STATIC, *args)


class LpPool:
    """Worker processes ("spawn": the parent holds a CUDA context) that each keep the static part of the LP."""

    def __init__(self, static: LpStatic, workers: int):
        import multiprocessing as mp
        self.pool = mp.get_context("spawn").Pool(workers, initializer=_init_worker, initargs=(static,))

    def solve_many(self, jobs):
        """jobs: list of (nzval, E, df, x, restoration); results in the same order."""
        return self.pool.map(_solve_in_worker, jobs, chunksize=max(1, len(jobs) // (4 * self.pool._processes) or 1))

    def close(self):
        self.pool.close()
        self.pool.join()
