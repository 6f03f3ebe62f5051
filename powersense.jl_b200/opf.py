"""create_opf_model! / run_opf! for the AC formulations (src/opf/formulations.jl:3-472, 584-653), with the NL block
served by GpuNLPEvaluator instead of JuMP's evaluator.

What stays on the host here is what the reference leaves to the *solver* side of the MOI boundary: the affine and
quadratic `@constraint` rows (formulations.jl:123-127, 141-168, 324-334, 430-471), the variable bounds (tgh >= 0,
:121) and the objective (:122-138).  The NL rows -- balances, flow definitions, limits -- are evaluated on the
GPU through the C ABI.  The reference hands the model to an MOI solver (Ipopt, or its own SLP `OPT` with GLPK); neither
exists in this image, so `run_opf` drives scipy's SLSQP (default; dense, small cases) or trust-constr (sparse, exact
Hessians) with the same callbacks."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import scipy.sparse as sp

from .data import PowersenseData
from .evaluator import GpuNLPEvaluator
from .synth import variable_layout
from .types import (ACOPF_fromulation, CBRARVmodel, CBRAWVmodel, PBRAPVmodel, PBRARVmodel, PBRAWVmodel, PNPAPVmodel,
                    PNRAPVmodel, PNRARVmodel, PNRAWVmodel)

POLAR = (PBRAPVmodel, PNPAPVmodel, PNRAPVmodel)
INF = float("inf")


@dataclass
class Rows:
    """lo <= A x + 0.5 x'Q_i x <= hi  (MOI's ScalarAffine / ScalarQuadratic rows)."""
    A: sp.csr_matrix
    quad: List[Optional[sp.csr_matrix]]       # per row: symmetric Q_i or None
    lo: np.ndarray
    hi: np.ndarray


class OPFModel:
    def __init__(self, data: PowersenseData, formulation: ACOPF_fromulation = PBRAPVmodel, initialize: bool = True,
                 FACTS_bShunt: bool = True, box_constraints: bool = False, obj_type: str = "linear", device: int = 0):
        # (PBRAWV / PNRAWV are built as the reference builds them, although they are incomplete relaxations there: Wr, Wi
        # are never tied to vr, vi -- the two rows that would do it are commented out at formulations.jl:417-418)
        if obj_type not in ("linear", "quadratic", "feasible"):
            raise ValueError("obj_type must be 'linear', 'quadratic' or 'feasible'")
        self.data, self.formulation, self.obj_type = data, formulation, obj_type
        d, F = data, formulation
        lin = obj_type == "linear"
        self.lay = lay = variable_layout(d, F, FACTS_bShunt, lin)
        n = self.n = lay["n_var"]
        nb, nr, ng = d.nbus, d.nbr, d.ngen
        # ---- start values (formulations.jl:20-67; the struct's own defaults when `initialize`)
        x0 = np.zeros(n)
        if initialize:
            if F in POLAR:
                x0[lay["theta"]:lay["theta"] + nb], x0[lay["V"]:lay["V"] + nb] = d.theta, d.V
            else:
                x0[lay["vi"]:lay["vi"] + nb], x0[lay["vr"]:lay["vr"] + nb] = d.vi, d.vr
            if "Wd" in lay:
                x0[lay["Wd"]:lay["Wd"] + nb] = d.Wd
            x0[lay["pg"]:lay["pg"] + ng], x0[lay["qg"]:lay["qg"] + ng] = d.Pg, d.Qg
            if lin:
                x0[lay["cg"]:lay["cg"] + ng] = d.cg
        self.x0 = x0
        # ---- variable bounds: tgh >= 0 (:121)
        self.xl, self.xu = np.full(n, -INF), np.full(n, INF)
        if lin:
            self.xl[lay["tgh"]:] = 0.0
        # ---- objective (:122-138)
        self.c = np.zeros(n)
        self.qdiag = np.zeros(n)                  # 0.5 * x' diag(qdiag) x
        self.c0 = 0.0
        if lin:
            self.c[lay["cg"]:lay["cg"] + ng] = 1.0
        elif obj_type == "quadratic":
            self.c0 = float(np.sum(d.c0))
            if d.cost_order >= 2:
                self.c[lay["pg"]:lay["pg"] + ng] = d.c1
            if d.cost_order == 3:
                self.qdiag[lay["pg"]:lay["pg"] + ng] = 2.0 * d.c2
        # ---- affine / quadratic rows
        ar, ac, av, lo, hi, quad = [], [], [], [], [], []

        def row(cols, vals, l, h, Q=None):
            r = len(lo)
            ar.extend([r] * len(cols)); ac.extend(cols); av.extend(vals)
            lo.append(l); hi.append(h); quad.append(Q)

        def diagq(i, j, v):                       # 0.5 x'Qx with Q = v on (i,i) and (j,j)  ->  0.5 v (xi^2 + xj^2)
            return sp.csr_matrix(([v, v], ([i, j], [i, j])), shape=(n, n))
        if lin:                                   # :123-127
            t = lay["tgh"]
            for i in range(ng):
                k = int(d.lps[i])
                cols = list(range(t, t + k))
                row([lay["cg"] + i] + cols, [1.0] + list(-np.asarray(d.cgh[i][:k])), 0.0, 0.0)
                row([lay["pg"] + i] + cols, [1.0] + list(-np.asarray(d.pgh[i][:k])), 0.0, 0.0)
                row(cols, [1.0] * k, 1.0, 1.0)
                t += k
        for i in range(ng):                       # :145-146
            row([lay["pg"] + i], [1.0], d.Pmin[i], d.Pmax[i])
        for i in range(ng):
            row([lay["qg"] + i], [1.0], d.Qmin[i], d.Qmax[i])
        if FACTS_bShunt:                          # :151
            for i in range(d.nbss):
                row([lay["bss"] + i], [1.0], d.bmin[i], d.bmax[i])
        if F in POLAR:                            # :156
            for i in range(nb):
                row([lay["V"] + i], [1.0], d.Vmin[i], d.Vmax[i])
        else:                                     # :158  Vmin^2 <= vr^2 + vi^2 <= Vmax^2
            for i in range(nb):
                row([], [], d.Vmin[i] ** 2, d.Vmax[i] ** 2, diagq(lay["vr"] + i, lay["vi"] + i, 2.0))
        if "Wd" in lay:                           # :162-163
            for i in range(nb):
                row([lay["Wd"] + i], [1.0], d.Vmin[i] ** 2, d.Vmax[i] ** 2)
            for i in range(nb):
                row([lay["Wd"] + i], [1.0], 0.0, 0.0, diagq(lay["vr"] + i, lay["vi"] + i, -2.0))
        if F in (PBRAWVmodel, PNRAWVmodel):       # :165-168: the same quadratic rows, emitted twice more
            for _ in range(2):
                for i in range(nb):
                    row([lay["Wd"] + i], [1.0], 0.0, 0.0, diagq(lay["vr"] + i, lay["vi"] + i, -2.0))
            # nodal balances in W-space: `@constraint` rows, not NL rows (:173-296); terms on the same variable merge
            # like JuMP's AffExpr does
            for i in range(nb):
                pc, pv, qc_, qv_ = [], [], [], []
                for g in d.gen_idx[d.gen_ptr[i]:d.gen_ptr[i + 1]]:                         # :181-185
                    pc.append(lay["pg"] + int(g) - 1); pv.append(1.0)
                    qc_.append(lay["qg"] + int(g) - 1); qv_.append(1.0)
                for side, (ptr, idx) in enumerate(((d.sij_ptr, d.sij_idx), (d.sji_ptr, d.sji_idx))):
                    for k in idx[ptr[i]:ptr[i + 1]]:
                        k = int(k) - 1
                        if F is PBRAWVmodel:                                               # :197-205
                            pc.append(lay["Pji" if side else "Pij"] + k); pv.append(1.0)
                            qc_.append(lay["Qji" if side else "Qij"] + k); qv_.append(1.0)
                        else:                                                              # :244-253
                            a = int(d.br[k][side]) - 1
                            gs, bs = (d.gt[k], d.bt[k]) if side else (d.gf[k], d.bf[k])
                            G, B = (d.Gt[k], d.Bt[k]) if side else (d.Gf[k], d.Bf[k])
                            sgn = -1.0 if side else 1.0
                            pc += [lay["Wd"] + a, lay["Wr"] + k, lay["Wi"] + k]; pv += [-gs, -G, sgn * B]
                            qc_ += [lay["Wd"] + a, lay["Wr"] + k, lay["Wi"] + k]; qv_ += [bs, B, sgn * G]
                pc.append(lay["Wd"] + i); pv.append(-d.Gl[i])                             # :280-281
                qc_.append(lay["Wd"] + i); qv_.append(d.Bl[i])
                Q = None
                sh = d.sh_idx[d.sh_ptr[i]:d.sh_ptr[i + 1]] if FACTS_bShunt else []
                if len(sh):                                                                # :282-286: bss[sh] * Wd[i]
                    ii = [lay["bss"] + int(q) - 1 for q in sh]
                    w = [lay["Wd"] + i] * len(ii)
                    Q = sp.csr_matrix(([1.0] * (2 * len(ii)), (ii + w, w + ii)), shape=(n, n))
                row(pc, pv, d.Pd[i], d.Pd[i])                                             # :294
                row(qc_, qv_, d.Qd[i], d.Qd[i], Q)                                        # :295
            mp_ = d.source_type == "matpower"
            for k in range(nr):
                f, t = int(d.br[k][0]) - 1, int(d.br[k][1]) - 1
                Wd, Wr, Wi = lay["Wd"], lay["Wr"] + k, lay["Wi"] + k
                if F is PBRAWVmodel:                                                       # :317-320 | :381-384
                    row([lay["Pij"] + k, Wd + f, Wr, Wi], [1.0, d.gf[k], d.Gf[k], -d.Bf[k]], 0.0, 0.0)
                    row([lay["Pji"] + k, Wd + t, Wr, Wi], [1.0, d.gt[k], d.Gt[k], d.Bt[k]], 0.0, 0.0)
                    row([lay["Qij"] + k, Wd + f, Wr, Wi], [1.0, -d.bf[k], -d.Bf[k], -d.Gf[k]], 0.0, 0.0)
                    row([lay["Qji"] + k, Wd + t, Wr, Wi], [1.0, -d.bt[k], -d.Bt[k], d.Gt[k]], 0.0, 0.0)
                else:                                                                      # :355-360 | :419-425
                    for side, a, b in ((0, f, t), (1, t, f)):
                        gs, bs = (d.gt[k], d.bt[k]) if side else (d.gf[k], d.bf[k])
                        G, B = (d.Gt[k], d.Bt[k]) if side else (d.Gf[k], d.Bf[k])
                        y, Y = gs * gs + bs * bs, G * G + B * B
                        y1, y2 = 2.0 * (gs * G + bs * B), 2.0 * (bs * G - gs * B)
                        if side:
                            y2 = -y2
                        im2 = d.Imax[k] ** 2
                        if mp_:       # y1 Wr Wd_a + y2 Wi Wd_a + y Wd_a Wd_a + Y Wd_b Wd_a <= Imax^2  (quadratic)
                            ii = [Wr, Wi, Wd + a, Wd + b]
                            cc = [y1, y2, 2.0 * y, Y]                     # 0.5 x'Qx: the square term carries 2 y on the diagonal
                            r_, c_, v_ = [], [], []
                            for q, cf in zip(ii, cc):
                                if q == Wd + a:
                                    r_.append(q); c_.append(q); v_.append(cf)
                                else:
                                    r_ += [q, Wd + a]; c_ += [Wd + a, q]; v_ += [cf, cf]
                            row([], [], -INF, im2, sp.csr_matrix((v_, (r_, c_)), shape=(n, n)))
                        else:
                            row([Wr, Wi, Wd + a, Wd + b], [y1, y2, y, Y], -INF, im2)
        if F in (CBRARVmodel, CBRAWVmodel):       # current definitions :324-327 / :331-334 (same in :388-398)
            vr, vi = lay["vr"], lay["vi"]
            for k in range(nr):
                f, t = int(d.br[k][0]) - 1, int(d.br[k][1]) - 1
                row([lay["Irij"] + k, vr + f, vi + f, vr + t, vi + t], [1.0, d.gf[k], -d.bf[k], d.Gf[k], -d.Bf[k]], 0.0, 0.0)
                row([lay["Irji"] + k, vr + t, vi + t, vr + f, vi + f], [1.0, d.gt[k], -d.bt[k], d.Gt[k], -d.Bt[k]], 0.0, 0.0)
                row([lay["Iiij"] + k, vi + f, vr + f, vi + t, vr + t], [1.0, d.gf[k], d.bf[k], d.Gf[k], d.Bf[k]], 0.0, 0.0)
                row([lay["Iiji"] + k, vi + t, vr + t, vi + f, vr + f], [1.0, d.gt[k], d.bt[k], d.Gt[k], d.Bt[k]], 0.0, 0.0)
        if box_constraints:                       # :430-471, in the reference's row order
            if F not in POLAR:
                for key in ("vi", "vr"):
                    for i in range(nb):
                        row([lay[key] + i], [1.0], -d.Vmax[i], d.Vmax[i])
            if F in (PBRAWVmodel, PNRAWVmodel):
                for k in range(nr):
                    f, t = int(d.br[k][0]) - 1, int(d.br[k][1]) - 1
                    b = d.Vmax[f] * d.Vmax[t]
                    row([lay["Wr"] + k], [1.0], -b, b)
                    row([lay["Wi"] + k], [1.0], -b, b)
            mp = d.source_type == "matpower"
            if "Pij" in lay:
                if mp:                             # vectorised: all Pij, all Pji, all Qij, all Qji
                    for key in ("Pij", "Pji", "Qij", "Qji"):
                        for k in range(nr):
                            row([lay[key] + k], [1.0], -d.Imax[k], d.Imax[k])
                else:
                    for k in range(nr):
                        f, t = int(d.br[k][0]) - 1, int(d.br[k][1]) - 1
                        bf, bt = d.Imax[k] * d.Vmax[f], d.Imax[k] * d.Vmax[t]
                        for key, b in (("Pij", bf), ("Pji", bt), ("Qij", bf), ("Qji", bt)):
                            row([lay[key] + k], [1.0], -b, b)
            elif "Irij" in lay:
                if mp:
                    for k in range(nr):
                        f, t = int(d.br[k][0]) - 1, int(d.br[k][1]) - 1
                        for key, e in (("Irij", f), ("Iiij", f), ("Irji", t), ("Iiji", t)):
                            row([lay[key] + k], [d.Vmin[e]], -d.Imax[k], d.Imax[k])
                else:
                    for key in ("Irij", "Iiij", "Irji", "Iiji"):
                        for k in range(nr):
                            row([lay[key] + k], [1.0], -d.Imax[k], d.Imax[k])
        m = len(lo)
        self.rows = Rows(sp.csr_matrix((av, (ar, ac)), shape=(m, n)), quad, np.array(lo, float), np.array(hi, float))
        # ---- NL block on the GPU
        self.evaluator = GpuNLPEvaluator(d, F, FACTS_bShunt, lin, device=device)
        self.evaluator.initialize(["Grad", "Jac", "Hess"])
        assert self.evaluator.num_variables == n
        self.nl_lo, self.nl_hi = self.evaluator.constraint_bounds()
        js = np.array(self.evaluator.jacobian_structure(), dtype=np.int64).reshape(-1, 2)
        hs = np.array(self.evaluator.hessian_lagrangian_structure(), dtype=np.int64).reshape(-1, 2)
        self.jrow, self.jcol = js[:, 0] - 1, js[:, 1] - 1
        self.hrow, self.hcol = hs[:, 0] - 1, hs[:, 1] - 1

    # objective
    def f(self, x):
        return float(self.c @ x + 0.5 * np.dot(self.qdiag * x, x) + self.c0)

    def grad_f(self, x):
        return self.c + self.qdiag * x

    # affine + quadratic rows (host: the solver side of the MOI boundary)
    def rows_value(self, x):
        v = self.rows.A @ x
        for i, Q in enumerate(self.rows.quad):
            if Q is not None:
                v[i] += 0.5 * x @ (Q @ x)
        return v

    # NL block (GPU)
    def nl_value(self, x):
        g = np.empty(self.evaluator.num_constraints)
        self.evaluator.eval_constraint(g, x)
        return g

    def nl_jac(self, x):
        J = np.empty(len(self.jrow))
        self.evaluator.eval_constraint_jacobian(J, x)
        return sp.coo_matrix((J, (self.jrow, self.jcol)), shape=(self.evaluator.num_constraints, self.n)).tocsr()

    def nl_hess(self, x, mu):
        H = np.empty(len(self.hrow))
        self.evaluator.eval_hessian_lagrangian(H, x, 1.0, mu)
        L = sp.coo_matrix((H, (self.hrow, self.hcol)), shape=(self.n, self.n)).tocsr()      # lower triangle, duplicates summed
        return L + sp.tril(L, -1).T

    def moi_rows(self):
        """The affine / quadratic rows as Powersense.Optimizer stores them: every `Interval` row reaches it through JuMP's
        SplitInterval bridge as one >= and one <= row, and rows are ordered [lin_le | lin_ge | lin_eq | quad_le | quad_ge |
        quad_eq] in front of the NLP block (src/opt/MOI_wrapper.jl:767-773).  Returns (rows, lower, upper, g_order) with
        rows = [(constant, [(coef, var)], [(coef, var1, var2)])], 1-based variables, MOI's quadratic convention (diagonal
        coefficient doubled in storage, :884-888)."""
        A, quad, lo, hi = self.rows.A, self.rows.quad, self.rows.lo, self.rows.hi
        buckets = {k: [] for k in ("lin_le", "lin_ge", "lin_eq", "quad_le", "quad_ge", "quad_eq")}
        for i in range(A.shape[0]):
            a = A.getrow(i)
            aff = [(float(v), int(c) + 1) for c, v in zip(a.indices, a.data)]
            q = []
            if quad[i] is not None:
                Q = quad[i].tocoo()
                q = [(float(v), int(r) + 1, int(c) + 1) for r, c, v in zip(Q.row, Q.col, Q.data) if r >= c]
            kind = "quad" if q else "lin"
            row = (0.0, aff, q)
            if lo[i] == hi[i]:
                buckets[kind + "_eq"].append((row, lo[i], hi[i]))
            else:
                if np.isfinite(hi[i]):
                    buckets[kind + "_le"].append((row, -INF, hi[i]))
                if np.isfinite(lo[i]):
                    buckets[kind + "_ge"].append((row, lo[i], INF))
        order = ("lin_le", "lin_ge", "lin_eq", "quad_le", "quad_ge", "quad_eq")
        rows = [r for k in order for (r, _, _) in buckets[k]]
        lower = np.array([l for k in order for (_, l, _) in buckets[k]])
        upper = np.array([u for k in order for (_, _, u) in buckets[k]])
        return rows, lower, upper, [len(buckets[k]) for k in order]

    def close(self):
        self.evaluator.close()


def create_opf_model(data: PowersenseData, **kw) -> OPFModel:
    """create_opf_model!(data; formulation, initialize, FACTS_bShunt, box_constraints, obj_type) (formulations.jl:3-10)."""
    data.model = OPFModel(data, **kw)
    return data.model


def run_opf(data: PowersenseData, formulation: ACOPF_fromulation = PBRAPVmodel, create_model: bool = True,
            solver: str = "SLSQP", initialize: bool = True, FACTS_bShunt: bool = True, box_constraints: bool = False,
            solve: bool = True, obj_type: str = "linear", device: int = 0, maxiter: int = 1000, verbose: int = 0):
    """run_opf!(data; formulation, create_model, solver, initialize, FACTS_bShunt, box_constraints, solve, obj_type)
    (formulations.jl:584-653).  Writes the solution back into `data` exactly like the reference (:608-651)."""
    from scipy.optimize import Bounds, LinearConstraint, NonlinearConstraint, minimize
    if create_model or data.model is None:
        create_opf_model(data, formulation=formulation, initialize=initialize, FACTS_bShunt=FACTS_bShunt,
                         box_constraints=box_constraints, obj_type=obj_type, device=device)
    M: OPFModel = data.model
    if not solve:
        return None
    lin_rows = np.array([q is None for q in M.rows.quad])
    cons = []
    if lin_rows.any():
        cons.append(LinearConstraint(M.rows.A[lin_rows], M.rows.lo[lin_rows], M.rows.hi[lin_rows]))
    if (~lin_rows).any():
        idx = np.nonzero(~lin_rows)[0]
        Qs = [M.rows.quad[i] for i in idx]
        Aq = M.rows.A[idx]

        def qv(x):
            return Aq @ x + np.array([0.5 * x @ (Q @ x) for Q in Qs])

        def qj(x):
            return sp.vstack([sp.csr_matrix(Q @ x) for Q in Qs]) + Aq

        def qh(x, v):
            H = sp.csr_matrix((M.n, M.n))
            for vi_, Q in zip(v, Qs):
                H = H + vi_ * Q
            return H
        cons.append(NonlinearConstraint(qv, M.rows.lo[idx], M.rows.hi[idx], jac=qj, hess=qh))
    cons.append(NonlinearConstraint(M.nl_value, M.nl_lo, M.nl_hi, jac=M.nl_jac, hess=M.nl_hess))
    if solver.upper() == "SLSQP":
        # dense SQP (small cases): same callbacks, Jacobians densified, no Hessian callbacks (BFGS inside SLSQP)
        dense = []
        for c in cons:
            if isinstance(c, LinearConstraint):
                dense.append(LinearConstraint(c.A.toarray(), c.lb, c.ub))
            else:
                dense.append(NonlinearConstraint(c.fun, c.lb, c.ub, jac=(lambda x, j=c.jac: j(x).toarray())))
        res = minimize(M.f, M.x0, jac=M.grad_f, method="SLSQP", bounds=Bounds(M.xl, M.xu), constraints=dense,
                       options={"maxiter": maxiter, "ftol": 1e-10, "disp": bool(verbose)})
    else:
        res = minimize(M.f, M.x0, jac=M.grad_f, hess=lambda x: sp.diags(M.qdiag), method=solver, bounds=Bounds(M.xl, M.xu),
                       constraints=cons, options={"maxiter": maxiter, "verbose": verbose, "gtol": 1e-6, "xtol": 1e-9})
    x, lay, d = res.x, M.lay, data
    nb, nr, ng = d.nbus, d.nbr, d.ngen
    d.cost = float(res.fun)                                        # :609
    if M.formulation in POLAR:                                     # :610-616
        d.theta, d.V = x[lay["theta"]:lay["theta"] + nb].copy(), x[lay["V"]:lay["V"] + nb].copy()
    else:
        d.vi, d.vr = x[lay["vi"]:lay["vi"] + nb].copy(), x[lay["vr"]:lay["vr"] + nb].copy()
    if "Wd" in lay:
        d.Wd = x[lay["Wd"]:lay["Wd"] + nb].copy()
    d.Pg, d.Qg = x[lay["pg"]:lay["pg"] + ng].copy(), x[lay["qg"]:lay["qg"] + ng].copy()   # :629-630
    for key in ("Pij", "Qij", "Pji", "Qji", "Irij", "Iiij", "Irji", "Iiji"):          # :634-645
        if key in lay:
            setattr(d, key, x[lay[key]:lay[key] + nr].copy())
    if FACTS_bShunt and d.nbss:
        d.Bss = x[lay["bss"]:lay["bss"] + d.nbss].copy()
    return res
