"""TEST INFRASTRUCTURE ONLY -- a CPU evaluation backend for powersense_b200.slp.BatchedSlpLS built on the oracles
(oracle/ps_oracle.c for the NL rows, numpy for the affine / quadratic rows and the KKT / merit metrics restating
oracle/kkt_oracle.py vectorised over scenarios).  The product has no CPU evaluation path; this exists so that the SLP
driver's host logic (the algorithm, the LP subproblems, the contingency handling) can be exercised on real OPF models in
the CPU test-suite, and so that the GPU backend can be checked against it iteration by iteration."""
import copy

import numpy as np
import scipy.sparse as sp
import torch

from oracle import c_oracle


def masked_data(d, k):
    """The network with branch k out of service ON THE FIXED PATTERN, i.e. what ps_batch_set_branch_status does: the
    branch tables are multiplied by the status like PowersenseData.jl:198-207 (Imax: 0 -> 10000); adjacency unchanged."""
    m = copy.copy(d)
    for name in ("Gf", "Bf", "Gt", "Bt", "gf", "bf", "gt", "bt"):
        a = np.array(getattr(d, name), dtype=np.float64, copy=True)
        a[k] = 0.0
        setattr(m, name, a)
    im = np.array(d.Imax, dtype=np.float64, copy=True)
    im[k] = 10000.0
    m.Imax = im
    return m


class OracleBackend:
    def __init__(self, model, S, Pd=None, Qd=None, status=None, facts_bshunt=True):
        self.model, self.S = model, int(S)
        self.dev = torch.device("cpu")
        d, F = model.data, model.formulation
        lin = model.obj_type == "linear"
        n = self.n = model.n
        rows, lo, hi, g_order = model.moi_rows()
        self.rows = rows
        base = c_oracle.COracle(d, F.code, facts_bshunt, lin)
        self.m_rows, self.m_nl = len(rows), base.m_nl
        self.m = self.m_rows + self.m_nl
        self.g_L = np.concatenate([lo, model.nl_lo])
        self.g_U = np.concatenate([hi, model.nl_hi])
        self.xl, self.xu, self.x0 = model.xl, model.xu, model.x0
        self.Pd = np.tile(d.Pd, (S, 1)) if Pd is None else np.asarray(Pd, dtype=np.float64)
        self.Qd = np.tile(d.Qd, (S, 1)) if Qd is None else np.asarray(Qd, dtype=np.float64)
        # one oracle per distinct outage pattern
        self.oracles, self.which = [base], np.zeros(S, dtype=np.int64)
        if status is not None:
            seen = {}
            for s in range(S):
                out = tuple(np.flatnonzero(np.asarray(status[s]) == 0).tolist())
                if not out:
                    continue
                if out not in seen:
                    dm = d
                    for k in out:
                        dm = masked_data(dm, k)
                    seen[out] = len(self.oracles)
                    self.oracles.append(c_oracle.COracle(dm, F.code, facts_bshunt, lin))
                self.which[s] = seen[out]
        # affine / quadratic rows as sparse operators: value = cst + A x + sum_q coef (0.5 if i == j) x_i x_j
        ar, ac, av, cst = [], [], [], []
        q_r, q_i, q_j, q_c = [], [], [], []
        jr, jc = [], []                                     # Jacobian structure of the rows block (MOI_wrapper.jl:780-799)
        for r, (c, aff, quad) in enumerate(rows):
            cst.append(c)
            for coef, var in aff:
                ar.append(r); ac.append(var - 1); av.append(coef)
                jr.append(r); jc.append(var - 1)
            for coef, i, j in quad:
                q_r.append(r); q_i.append(i - 1); q_j.append(j - 1); q_c.append(coef)
                jr.append(r); jc.append(i - 1)
                if i != j:
                    jr.append(r); jc.append(j - 1)
        self.A = sp.csr_matrix((av, (ar, ac)), shape=(self.m_rows, n))
        self.cst = np.array(cst)
        self.q = (np.array(q_r, int), np.array(q_i, int), np.array(q_j, int), np.array(q_c))
        nj_r, nj_c, _, _ = base.structure()
        self.nnz_rows = len(jr)
        row = np.concatenate([np.array(jr, int), nj_r - 1 + self.m_rows])
        col = np.concatenate([np.array(jc, int), nj_c - 1])
        # CSC pattern with duplicates merged, sorted by (col, row) like ps_csc_create
        order = np.lexsort((row, col))
        key = col[order] * (self.m + 1) + row[order]
        first = np.concatenate([[True], key[1:] != key[:-1]])
        self.slot_of = np.empty(len(row), dtype=np.int64)
        self.slot_of[order] = np.cumsum(first) - 1
        self.nnz = int(first.sum())
        self.rowval = (row[order][first] + 1).astype(np.int64)
        cnt = np.bincount(col[order][first], minlength=n)
        self.colptr = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int64)
        self._row0, self._col0 = row[order][first], col[order][first]
        self.c, self.qdiag, self.c0 = model.c, model.qdiag, model.c0

    def linear_rows(self):
        keep = [r for r, (c, aff, quad) in enumerate(self.rows) if not quad]
        lo, hi = self.g_L[:self.m_rows], self.g_U[:self.m_rows]
        return self.A[keep], self.cst[keep], lo[keep], hi[keep]

    # ---- evaluations
    def _rows_g(self, X):
        v = (self.A @ X.T).T + self.cst
        qr, qi, qj, qc = self.q
        if len(qr):
            w = np.where(qi == qj, 0.5, 1.0) * qc
            np.add.at(v, (slice(None), qr), w * X[:, qi] * X[:, qj])
        return v

    def _nl(self, X, what):
        S = X.shape[0]
        g = np.zeros((S, self.m_nl))
        J = np.zeros((S, len(self.slot_of) - self.nnz_rows)) if what & 2 else None
        for q, o in enumerate(self.oracles):
            idx = np.flatnonzero(self.which == q)
            if not len(idx):
                continue
            og, oJ, _ = o.eval_batch(X[idx], None, self.Pd[idx], self.Qd[idx], what=what)
            g[idx] = og
            if what & 2:
                J[idx] = oJ
        return g, J

    def eval_g(self, X, E, f):
        x = X.numpy()
        g, _ = self._nl(x, 1)
        E[:] = torch.from_numpy(np.concatenate([self._rows_g(x), g], axis=1))
        f[:, 0] = torch.from_numpy(x @ self.c + 0.5 * np.einsum("sj,j,sj->s", x, self.qdiag, x) + self.c0)

    def eval_all(self, X, E, f, df, nz):
        x = X.numpy()
        g, J = self._nl(x, 3)
        E[:] = torch.from_numpy(np.concatenate([self._rows_g(x), g], axis=1))
        f[:, 0] = torch.from_numpy(x @ self.c + 0.5 * np.einsum("sj,j,sj->s", x, self.qdiag, x) + self.c0)
        df[:] = torch.from_numpy(self.c[None, :] + self.qdiag[None, :] * x)
        # rows-block Jacobian values in structure order (fill_constraint_jacobian!, MOI_wrapper.jl:973-1002)
        S = x.shape[0]
        vals = np.zeros((S, self.nnz_rows))
        p = 0
        for c, aff, quad in self.rows:
            for coef, var in aff:
                vals[:, p] = coef; p += 1
            for coef, i, j in quad:
                vals[:, p] = coef * x[:, j - 1]; p += 1
                if i != j:
                    vals[:, p] = coef * x[:, i - 1]; p += 1
        dE = np.concatenate([vals, J], axis=1)
        out = np.zeros((S, self.nnz))
        np.add.at(out, (slice(None), self.slot_of), dE)              # compute_jacobian_matrix: J[r, c] += dE[i]
        nz[:] = torch.from_numpy(out)

    def metrics(self, E, X, lam, mU, mL, df, nz, nu, P, f, out):
        """oracle/kkt_oracle.py, vectorised over the scenarios (common.jl:35-98, slp.jl:108-162)."""
        E, X, lam, mU, mL, df, nz = (a.numpy() for a in (E, X, lam, mU, mL, df, nz))
        gL, gU = self.g_L, self.g_U
        viol = np.where(E > gU, E - gU, np.where(E < gL, gL - E, 0.0))
        violx = np.where(X > self.xu, X - self.xu, np.where(X < self.xl, self.xl - X, 0.0))
        o = out.numpy()
        o[:, 0] = viol.sum(axis=1) + violx.sum(axis=1)
        o[:, 1] = np.maximum(viol.max(axis=1, initial=0.0), violx.max(axis=1, initial=0.0))
        ineq = gL != gU
        with np.errstate(invalid="ignore"):
            c = np.minimum(E - gL, gU - E) * lam
        c = np.where(np.isnan(c) | ~ineq, 0.0, c)
        o[:, 2] = np.abs(c).max(axis=1, initial=0.0) / (1.0 + np.sqrt((lam * lam * ineq).sum(axis=1)))
        S = E.shape[0]
        JtL = np.zeros((S, self.n))
        np.add.at(JtL, (slice(None), self._col0), nz * lam[:, self._row0])
        rs = np.zeros((S, self.m))
        np.add.at(rs, (slice(None), self._row0), nz * nz)
        ndf = np.linalg.norm(df, axis=1)
        scal = np.maximum(np.maximum(1.0, ndf), (np.abs(lam) * np.sqrt(rs)).max(axis=1, initial=0.0))
        o[:, 3] = np.linalg.norm(df - JtL - mU - mL, axis=1) / scal
        fv = f.numpy() if f is not None else np.zeros(S)
        if nu is not None:
            pen = (nu.numpy() * viol).sum(axis=1)
            o[:, 4] = fv + pen
            o[:, 5] = ((df * P.numpy()).sum(axis=1) if P is not None else 0.0) - pen
        else:
            o[:, 4], o[:, 5] = fv, 0.0
        o[:, 6], o[:, 7] = ndf, scal
        return out

    def close(self):
        pass
