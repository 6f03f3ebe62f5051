"""Lock-step batched SLP line search: S scenarios (load cases / N-1 contingencies) of one OPF model solved together
with the Powersense `OPT` algorithm, every evaluation on the GPU.

Restates the reference's line-search SLP (src/opt/algorithms/slp_line_search.jl:80-278 `run!` / `compute_alpha`,
src/opt/algorithms/slp.jl:83-162 `compute_nu!` / `compute_phi` / `compute_derivative`, the slack-augmented LP of
src/opt/algorithms/subproblem.jl:55-334 and its dual extraction :657-700) with the per-iteration work split as:

  device, one launch per batch   NL rows g / J (ps_batch_eval), affine + quadratic rows (ps_rows_eval_batch), objective
                                 value and gradient terms (a 1-row ps_rows block), COO -> CSC assembly
                                 (ps_csc_assemble_batch), primal / dual infeasibility, complementarity, merit value and
                                 directional derivative (ps_kkt_eval_batch), and every backtracking trial point
  host, per scenario             the LP subproblem (scipy `linprog` / HiGHS here, GLPK in the reference's tests): an
                                 external serial simplex, out of scope of the GPU path

What follows the reference statement by statement: the start-point clamp (:100-107), ConvexFeasibleInitialization
(level 1, the default of parameters.jl:15: the L1-nearest point to the start that satisfies the affine rows,
slp.jl:23-48 with the model of MOI_wrapper.jl:1121-1145; a failure keeps the clamped start like the try/catch of
:116-122), the SlpLS penalty rule nu_i = |lambda_i| at a scenario's first iteration and max(nu_i, |lambda_i|) afterwards
(slp_line_search.jl:285-295 -- not the generic rule of slp.jl:83-95, which belongs to the trust-region variant), the
restoration LP in the reference's own parametrisation (lp_host.solve_lp), and the exit logic of :156-240.

Differences from the reference, stated: scenarios advance in lock-step, so a finished scenario simply stops updating and
`iter` is kept per scenario; an unexpected LP status leaves the status at the initial -5 unless the point is feasible (the
reference's `slp.ret == -3` at :158 is a comparison, not an assignment -- mirrored, not "fixed"); the final E is
re-evaluated at the returned x (the reference returns the E of its last eval_functions! call, which is the same point).
The `@constraint` rows are shared by all scenarios, which holds for the formulations whose affine rows do not involve
branch parameters (every formulation except CBRARV / CBRAWV when branch status masks are used)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import scipy.sparse as sp

from . import capi, lp_host
from .opf import OPFModel


@dataclass
class SlpOptions:                      # src/opt/parameters.jl:16-30
    tol_direction: float = 1e-6
    tol_residual: float = 0.01
    tol_infeas: float = 0.01
    tol_error: float = 0.0
    max_iter: int = 200
    eta: float = 0.4
    tau: float = 0.9
    min_alpha: float = 1e-8
    delta: float = 1000.0              # sub_optimize!(slp, Δ = 1000.0), slp.jl:50
    lp_workers: int = 0                # > 0: the scenarios' LPs of one iteration are solved by that many worker processes
    convex_init: int = 1               # ConvexFeasibleInitialization (parameters.jl:15); 0 = off
    restoration_lp: str = "reference"  # lp_host.solve_lp: "reference" (subproblem.jl:404-500 as written) | "standard"

class GpuBackend:
    """Device side of the iteration for an OPF model: every evaluation is one batched launch over the S scenarios
    (ps_batch_eval, ps_rows_eval_batch, ps_csc_assemble_batch, ps_kkt_eval_batch).  BatchedSlpLS only needs the small
    interface below (attributes dev, n, m, nnz, g_L, g_U, xl, xu, x0, colptr, rowval; eval_g, eval_all, metrics), which
    the tests also implement on a CPU checker to run the same algorithm on the reference's own toy problem."""

    def __init__(self, model: OPFModel, S: int, Pd=None, Qd=None, status=None):
        import torch
        self.torch = torch
        self.model, self.S = model, int(S)
        h = model.evaluator.handle
        self.h = h
        self.dev = torch.device("cuda", torch.cuda.current_device())
        n = self.n = model.n
        rows, lo, hi, g_order = model.moi_rows()
        self.m_rows, self.m = len(rows), len(rows) + h.m_nl
        self.g_order = g_order + [h.m_nl]
        self.g_L = np.concatenate([lo, model.nl_lo])
        self.g_U = np.concatenate([hi, model.nl_hi])
        self.xl, self.xu, self.x0 = model.xl, model.xu, model.x0
        # device blocks: rows (N1), objective (1-row block), NL batch, CSC pattern (a11), metrics (N3)
        from_rows = _rows_to_csr(rows)
        self.rows = capi.RowsBlock(n, *from_rows[:7], constant=from_rows[7], device=self.dev.index)
        obj_aff = [(float(c), int(j) + 1) for j, c in enumerate(model.c) if c != 0.0]
        obj_quad = [(float(q), int(j) + 1, int(j) + 1) for j, q in enumerate(model.qdiag) if q != 0.0]
        oc = _rows_to_csr([(model.c0, obj_aff, obj_quad)])
        self.obj = capi.RowsBlock(n, *oc[:7], constant=oc[7], device=self.dev.index)
        self.obj_jcol = torch.from_numpy(self.obj.jacobian_structure()[1] - 1).to(self.dev)
        self.batch = capi.Batch(h, self.S)
        if Pd is not None:
            self.batch.set_loads(Pd, Qd)
        if status is not None:
            self.batch.set_branch_status(status)
        rj_r, rj_c = self.rows.jacobian_structure()
        nj_r, nj_c = h.jacobian_structure()
        self.nnz_rows = len(rj_r)
        jrow = np.concatenate([rj_r, nj_r + self.m_rows])        # MOI_wrapper.jl:810-830
        jcol = np.concatenate([rj_c, nj_c])
        self.pattern = capi.CscPattern(self.m, n, jrow, jcol, device=self.dev.index)
        self.colptr, self.rowval = self.pattern.structure()
        self.nnz = self.pattern.nnz
        self.kkt = capi.KktMetrics(self.pattern, self.g_L, self.g_U, model.xl, model.xu)
        Z = lambda *shape: torch.zeros(*shape, dtype=torch.float64, device=self.dev)
        self.dE = Z(self.S, self.nnz_rows + h.nnzJ)
        self.gobj = Z(self.S, max(self.obj.nnzJ, 1))

    def linear_rows(self):
        """(A csr [m_lin, n], constant, lo, hi) of the model's affine rows: what ConvexFeasibleInitialization level 1 puts
        into its LP (linear_le / linear_ge / linear_eq constraints, MOI_wrapper.jl:1136-1144)."""
        rows, lo, hi, _ = self.model.moi_rows()
        keep = [r for r, (c, aff, quad) in enumerate(rows) if not quad]
        data, ri, ci, cst = [], [], [], []
        for q, r in enumerate(keep):
            c, aff, _ = rows[r]
            cst.append(c)
            for coef, var in aff:
                data.append(coef); ri.append(q); ci.append(var - 1)
        A = sp.csr_matrix((data, (ri, ci)), shape=(len(keep), self.n))
        return A, np.array(cst), np.asarray(lo)[keep], np.asarray(hi)[keep]

    def eval_g(self, X, E, f):
        """E = [affine/quadratic rows | NL rows](X), f = objective(X) -- eval_g / eval_f of the closures of
        MOI_wrapper.jl:1168-1180."""
        self.rows.eval_torch(capi.EVAL_G, X, None, E, None, None)
        self.batch.eval_torch(capi.EVAL_G, X, None, E[:, self.m_rows:], None, None, None)
        self.obj.eval_torch(capi.EVAL_G, X, None, f, None, None)

    def eval_all(self, X, E, f, df, nz):
        self.rows.eval_torch(capi.EVAL_G | capi.EVAL_J, X, None, E, self.dE, None)
        self.batch.eval_torch(capi.EVAL_G | capi.EVAL_J, X, None, E[:, self.m_rows:], self.dE[:, self.nnz_rows:], None, None)
        self.obj.eval_torch(capi.EVAL_G | capi.EVAL_J, X, None, f, self.gobj, None)
        df.zero_()                                                  # fill_gradient! (MOI_wrapper.jl:906-930): dense, duplicates added
        if self.obj.nnzJ:
            df.index_add_(1, self.obj_jcol, self.gobj)
        self.pattern.assemble_torch(self.dE, nz)                    # compute_jacobian_matrix (common.jl:12-20)

    def metrics(self, E, X, lam, mU, mL, df, nz, nu, P, f, out):
        """[S, PS_NKKT]: viol1, viol_inf, complementarity, KKT residual, phi(0), D(0) (nu / P may be None)."""
        return self.kkt.eval_torch(E, X, lam, mU, mL, df, nz, nu, P, f, out=out)

    def close(self):
        for o in (self.kkt, self.pattern, self.batch, self.rows, self.obj):
            o.close()


class BatchedSlpLS:
    def __init__(self, model: Optional[OPFModel], S: int, Pd=None, Qd=None, status=None, options: Optional[SlpOptions] = None,
                 backend=None):
        import torch
        self.torch = torch
        self.S, self.opt = int(S), options or SlpOptions()
        be = self.be = backend if backend is not None else GpuBackend(model, S, Pd, Qd, status)
        self.dev = be.dev
        n, m = self.n, self.m = be.n, be.m
        self.g_L, self.g_U, self.xl, self.xu, self.x0 = be.g_L, be.g_U, be.xl, be.xu, be.x0
        self.colptr, self.rowval = be.colptr, be.rowval
        f64 = torch.float64
        Z = lambda *shape: torch.zeros(*shape, dtype=f64, device=self.dev)
        S = self.S
        self.X, self.Xt = Z(S, n), Z(S, n)
        self.E, self.Et = Z(S, m), Z(S, m)
        self.nz = Z(S, be.nnz)
        self.lam, self.nu = Z(S, m), Z(S, m)
        self.mU, self.mL, self.df, self.P = Z(S, n), Z(S, n), Z(S, n), Z(S, n)
        self.f, self.ft = Z(S, 1), Z(S, 1)
        self.metrics, self.metrics_t = Z(S, capi.PS_NKKT), Z(S, capi.PS_NKKT)
        # row classes of the LP (subproblem.jl:141-318): le / ge / eq
        le = np.isinf(self.g_L) & np.isfinite(self.g_U)
        ge = np.isfinite(self.g_L) & np.isinf(self.g_U)
        eq = self.g_L == self.g_U
        if not np.all(le | ge | eq):
            raise NotImplementedError("two-sided rows reach Powersense.Optimizer split by the SplitInterval bridge")
        self.i_le, self.i_ge, self.i_eq = np.nonzero(le)[0], np.nonzero(ge)[0], np.nonzero(eq)[0]
        self._lp_static = lp_host.LpStatic(n=n, m=m, colptr=np.asarray(self.colptr), rowval=np.asarray(self.rowval), i_le=self.i_le,
                                           i_ge=self.i_ge, i_eq=self.i_eq, g_L=np.asarray(self.g_L), g_U=np.asarray(self.g_U),
                                           xl=np.asarray(self.xl), xu=np.asarray(self.xu), delta=self.opt.delta,
                                           tol_error=self.opt.tol_error, restoration=self.opt.restoration_lp)
        self._lp_pool = lp_host.LpPool(self._lp_static, self.opt.lp_workers) if self.opt.lp_workers > 0 else None

    # ---- evaluations (device) ----------------------------------------------------------------------------------------
    def _eval_g(self, X, E, f):
        self.be.eval_g(X, E, f)

    def _eval_all(self):
        self.be.eval_all(self.X, self.E, self.f, self.df, self.nz)

    def _metrics(self, with_dir: bool):
        self.be.metrics(self.E, self.X, self.lam, self.mU, self.mL, self.df, self.nz, self.nu if with_dir else None,
                        self.P if with_dir else None, self.f.view(-1), self.metrics)
        return self.metrics.cpu().numpy()

    # ---- host: the LP subproblem of one scenario (subproblem.jl:55-334, 349-700) ---------------------------------
    def _lp(self, nzval, E, df, x, restoration):
        """The LP subproblem of one scenario (lp_host.solve_lp; subproblem.jl:55-334, 657-700)."""
        return lp_host.solve_lp(self._lp_static, nzval, E, df, x, restoration)

    def _convex_initialization(self, x0):
        """convex_initialization! (slp.jl:23-48) on the model of MOI_wrapper.jl:1121-1145: variables (x, s+, s-), minimise
        sum(s+ + s-) subject to the model's affine rows on x, s+- >= 0 and x + s+ - s- = x0 -- the L1-nearest point to the
        start that satisfies the affine rows (x itself is free there: the column bounds are not part of that model).  One LP
        per distinct start point; a failure keeps the clamped start (the reference's try/catch)."""
        from scipy.optimize import linprog
        A, cst, lo, hi = self.be.linear_rows()
        if A.shape[0] == 0:
            return x0
        n = self.n
        I = sp.identity(n, format="csr")
        fin_u, fin_l = np.isfinite(hi), np.isfinite(lo)
        eq = fin_u & fin_l & (lo == hi)
        ub_rows = [A[fin_u & ~eq], -A[fin_l & ~eq]]
        b_ub = np.concatenate([(hi - cst)[fin_u & ~eq], -(lo - cst)[fin_l & ~eq]])
        Z = sp.csr_matrix((ub_rows[0].shape[0] + ub_rows[1].shape[0], 2 * n))
        A_ub = sp.hstack([sp.vstack(ub_rows), Z], format="csr")
        c = np.concatenate([np.zeros(n), np.ones(2 * n)])
        bounds = [(None, None)] * n + [(0.0, None)] * (2 * n)
        out = x0.copy()
        uniq, inv = np.unique(x0, axis=0, return_inverse=True)
        for q, xs in enumerate(uniq):
            A_eq = sp.vstack([sp.hstack([A[eq], sp.csr_matrix((int(eq.sum()), 2 * n))]), sp.hstack([I, I, -I])], format="csr")
            b_eq = np.concatenate([(lo - cst)[eq], xs])
            try:
                res = linprog(c, A_ub=A_ub if A_ub.shape[0] else None, b_ub=b_ub if A_ub.shape[0] else None, A_eq=A_eq, b_eq=b_eq,
                              bounds=bounds, method="highs")
            except Exception:
                continue
            if res.status == 0:
                out[np.asarray(inv).ravel() == q] = res.x[:n]
        return out

    # ---- the algorithm (slp_line_search.jl:80-249) -----------------------------------------------------------------
    def run(self, x0: Optional[np.ndarray] = None):
        t, o, S, n, m = self.torch, self.opt, self.S, self.n, self.m
        x0 = self.x0 if x0 is None else x0
        x0 = np.broadcast_to(np.asarray(x0, dtype=np.float64), (S, n)).copy()
        xl, xu = self.xl, self.xu
        x0 = np.where(xl > -np.inf, np.maximum(x0, xl), x0)        # :100-107 (clamp to the column bounds)
        x0 = np.where(xu < np.inf, np.minimum(x0, xu), x0)
        if o.convex_init > 0 and hasattr(self.be, "linear_rows"):  # :116-122
            x0 = self._convex_initialization(x0)
        self.X.copy_(t.from_numpy(x0))
        restoration = np.zeros(S, dtype=bool)
        ret = np.full(S, -5)
        done = np.zeros(S, dtype=bool)
        iters = np.ones(S, dtype=np.int64)
        slack_sum = np.zeros(S)
        nu = np.zeros((S, m))
        hist = []
        it = 1
        import time
        tm = {"gpu_eval_s": 0.0, "host_lp_s": 0.0, "lp_calls": 0, "trial_batches": 0}   # where the wall time goes
        t_run = time.perf_counter()
        while not done.all():
            t0 = time.perf_counter()
            self._eval_all()                                        # eval_functions!, slp.jl:201-206
            mets = self._metrics(with_dir=False)
            tm["gpu_eval_s"] += time.perf_counter() - t0
            prim, dual, compl = mets[:, 1].copy(), mets[:, 3].copy(), mets[:, 2].copy()   # :141-143
            nz, E, df, X = self.nz.cpu().numpy(), self.E.cpu().numpy(), self.df.cpu().numpy(), self.X.cpu().numpy()
            P, lam, mU, mL = np.zeros((S, n)), np.zeros((S, m)), np.zeros((S, n)), np.zeros((S, n))
            skip_step = np.zeros(S, dtype=bool)
            todo = np.nonzero(~done)[0]
            t0 = time.perf_counter()
            jobs = [(nz[s], E[s], df[s], X[s], bool(restoration[s])) for s in todo]
            sols = self._lp_pool.solve_many(jobs) if self._lp_pool is not None else [self._lp(*j) for j in jobs]   # :147-148
            tm["host_lp_s"] += time.perf_counter() - t0
            tm["lp_calls"] += len(jobs)
            for s, (p, l, a, b, ssum, status) in zip(todo, sols):
                if status == "INFEASIBLE":                          # :160-175
                    if restoration[s]:
                        ret[s] = 6 if prim[s] <= o.tol_infeas else 2
                        done[s] = True
                    else:
                        restoration[s] = True
                        skip_step[s] = True
                    continue
                if status != "OPTIMAL":                             # :156-162 (`slp.ret == -3` is a no-op there: ret stays -5)
                    ret[s] = 6 if prim[s] <= o.tol_infeas else ret[s]
                    done[s] = True
                    continue
                P[s], lam[s], mU[s], mL[s], slack_sum[s] = p, l, a, b, ssum
                if iters[s] == 1:                                   # compute_nu!(::SlpLS), slp_line_search.jl:285-295
                    nu[s] = np.abs(l)
                else:
                    nu[s] = np.maximum(nu[s], np.abs(l))
            active = ~done & ~skip_step
            self.P.copy_(t.from_numpy(P)); self.lam.copy_(t.from_numpy(lam))
            self.mU.copy_(t.from_numpy(mU)); self.mL.copy_(t.from_numpy(mL)); self.nu.copy_(t.from_numpy(nu))
            mets = self._metrics(with_dir=True)                     # phi (slp.jl:108-131) and D (:138-162) at alpha = 0
            viol1 = mets[:, 0]
            phi0 = np.where(restoration, _viol_sum(E, self.g_L, self.g_U), mets[:, 4])
            D = np.where(restoration, slack_sum - _viol_sum(E, self.g_L, self.g_U), mets[:, 5])
            # ---- compute_alpha (:256-278): backtracking, every trial point of every scenario in one batch
            alpha = np.ones(S)
            valid = np.ones(S, dtype=bool)
            searching = active.copy()
            while searching.any():
                t0 = time.perf_counter()
                tm["trial_batches"] += 1
                self.Xt.copy_(self.X + t.from_numpy(alpha[:, None] * P).to(self.dev))
                self._eval_g(self.Xt, self.Et, self.ft)
                # merit at the trial point on the device (compute_phi, slp.jl:108-131): out[4] = f + sum nu*viol, and
                # out[0] = sum of violations (the step respects the column bounds, so only the rows contribute)
                self.be.metrics(self.Et, self.Xt, self.lam, self.mU, self.mL, self.df, self.nz, self.nu, self.P,
                                self.ft.view(-1), self.metrics_t)
                mt = self.metrics_t.cpu().numpy()
                tm["gpu_eval_s"] += time.perf_counter() - t0
                phi_t = np.where(restoration, mt[:, 0], mt[:, 4])
                ok = phi_t <= phi0 + o.eta * alpha * D
                for s in np.nonzero(searching)[0]:
                    if ok[s]:
                        searching[s] = False
                    elif alpha[s] < o.min_alpha:
                        if restoration[s]:
                            ret[s] = -3
                        valid[s] = False
                        searching[s] = False
                    else:
                        alpha[s] *= o.tau
            hist.append({"iter": it, "prim_infeas": prim.copy(), "dual_infeas": dual.copy(), "compl": compl.copy(),
                         "alpha": alpha.copy(), "f": self.f.cpu().numpy().ravel().copy(), "viol1": viol1.copy(),
                         "phi": phi0.copy(), "D": D.copy(), "slack_sum": slack_sum.copy(), "restoration": restoration.copy(),
                         "p_inf": np.abs(P).max(axis=1), "nu": nu.copy(), "lambda": lam.copy(), "iters": iters.copy(),
                         "lp_solved": (~done & ~skip_step).copy()})
            step = np.zeros(S, dtype=bool)
            for s in np.nonzero(active)[0]:
                if slack_sum[s] <= o.tol_infeas:                    # :197-199
                    restoration[s] = False
                if iters[s] >= o.max_iter:                          # :202-208
                    ret[s] = 6 if prim[s] <= o.tol_infeas else -1
                    done[s] = True
                    continue
                pinf = np.abs(P[s]).max() if n else 0.0
                if (prim[s] <= o.tol_infeas and compl[s] <= o.tol_residual) or pinf <= o.tol_direction:   # :210-222
                    if restoration[s]:
                        restoration[s] = False
                        iters[s] += 1
                        continue
                    elif dual[s] <= o.tol_residual:
                        ret[s] = 0
                        done[s] = True
                        continue
                if not valid[s]:                                    # :225-240
                    if ret[s] == -3:
                        ret[s] = 6 if prim[s] <= o.tol_infeas else 2
                        done[s] = True
                    else:
                        restoration[s] = True
                        iters[s] += 1
                    continue
                step[s] = True
                iters[s] += 1
            iters[skip_step & ~done] += 0                           # `continue` without iter += 1 (:171-173)
            upd = t.from_numpy((alpha * step)[:, None] * P).to(self.dev)
            self.X += upd                                           # :243
            it += 1
            if it > 10 * o.max_iter:
                break
        self._eval_g(self.X, self.E, self.f)
        tm["total_s"] = time.perf_counter() - t_run
        return {"x": self.X.cpu().numpy(), "obj": self.f.cpu().numpy().ravel(), "status": ret, "iter": iters, "history": hist,
                "g": self.E.cpu().numpy(), "lambda": self.lam.cpu().numpy(), "timing": tm}

    def close(self):
        if self._lp_pool is not None:
            self._lp_pool.close()
            self._lp_pool = None
        self.be.close()


def _viol_rows(E, g_L, g_U):
    return np.where(E > g_U, E - g_U, np.where(E < g_L, g_L - E, 0.0))


def _viol_sum(E, g_L, g_U):
    return _viol_rows(E, g_L, g_U).sum(axis=1)


def _rows_to_csr(rows):
    ap, av, ac, qp, q1, q2, qc, cst = [0], [], [], [0], [], [], [], []
    for c, aff, quad in rows:
        cst.append(c)
        for coef, var in aff:
            av.append(var); ac.append(coef)
        for coef, i, j in quad:
            q1.append(i); q2.append(j); qc.append(coef)
        ap.append(len(av)); qp.append(len(q1))
    return (np.array(ap, np.int64), np.array(av, np.int64), np.array(ac, np.float64), np.array(qp, np.int64),
            np.array(q1, np.int64), np.array(q2, np.int64), np.array(qc, np.float64), np.array(cst, np.float64))
