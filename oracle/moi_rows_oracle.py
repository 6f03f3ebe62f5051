"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the affine / quadratic row evaluation of Powersense.Optimizer
(src/opt/MOI_wrapper.jl): eval_function (:864-891), fill_constraint_jacobian! (:973-1002), fill_hessian_lagrangian!
(:1030-1042) and the two sparsity builders (:780-800, :834-839).  Plain Python loops in the reference's term order.
Parity unpinned against a Julia run (no Julia here); the functions are a few lines each and are followed literally.

A row is (constant, affine_terms=[(coef, var)], quadratic_terms=[(coef, var1, var2)]) with 1-based variables."""
import numpy as np


def eval_function(row, x):
    cst, aff, quad = row
    v = cst
    for coef, var in aff:                               # :872-878
        v += coef * x[var - 1]
    for coef, i, j in quad:                             # :880-889
        if i == j:
            v += 0.5 * coef * x[i - 1] * x[j - 1]
        else:
            v += coef * x[i - 1] * x[j - 1]
    return v


def jacobian_structure(rows):
    out = []
    for r, (cst, aff, quad) in enumerate(rows, start=1):
        for coef, var in aff:                           # :780-784
            out.append((r, var))
        for coef, i, j in quad:                         # :786-799
            out.append((r, i))
            if i != j:
                out.append((r, j))
    return out


def hessian_structure(rows):
    return [(i, j) for (_, _, quad) in rows for (_, i, j) in quad]      # :834-839


def fill_constraint_jacobian(rows, x):
    vals = []
    for cst, aff, quad in rows:
        for coef, var in aff:                           # :973-979
            vals.append(coef)
        for coef, i, j in quad:                         # :981-1002
            vals.append(coef * x[j - 1])
            if i != j:
                vals.append(coef * x[i - 1])
    return np.array(vals)


def fill_hessian_lagrangian(rows, lam):
    return np.array([lam[r] * coef for r, (_, _, quad) in enumerate(rows) for (coef, _, _) in quad])   # :1036-1041


def to_csr(rows):
    """Row list -> the CSR arrays of ps_rows_create."""
    ap, av, ac, qp, q1, q2, qc, cst = [0], [], [], [0], [], [], [], []
    for c, aff, quad in rows:
        cst.append(c)
        for coef, var in aff:
            av.append(var); ac.append(coef)
        for coef, i, j in quad:
            q1.append(i); q2.append(j); qc.append(coef)
        ap.append(len(av)); qp.append(len(q1))
    return (np.array(ap, np.int64), np.array(av, np.int64), np.array(ac), np.array(qp, np.int64), np.array(q1, np.int64),
            np.array(q2, np.int64), np.array(qc), np.array(cst))
