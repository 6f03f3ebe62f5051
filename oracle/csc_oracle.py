"""TEST INFRASTRUCTURE ONLY -- CPU restatement of compute_jacobian_matrix (src/opt/algorithms/common.jl:12-20):

    J = spzeros(m, n); for i = 1:length(j_str)  J[j_str[i][1], j_str[i][2]] += dE[i]  end

as a dense accumulation in COO order (small cases), returned in SparseMatrixCSC form (1-based colptr / rowval,
nzval).  Julia's `setindex!` on a SparseMatrixCSC does not store an entry whose first assigned value is zero; the
product keeps a *fixed* pattern (every structural COO position, explicit zeros included), so the comparison is on the
dense matrix and on the pattern of structural positions.  Parity unpinned against a Julia run (no Julia here)."""
import numpy as np


def compute_jacobian_matrix(m, n, rows, cols, dE):
    J = np.zeros((m, n))
    for r, c, v in zip(rows, cols, dE):            # common.jl:16-18, in COO order
        J[r - 1, c - 1] += v
    return J


def structural_csc(m, n, rows, cols):
    """colptr / rowval (1-based) of the fixed pattern: every (row, col) that appears in the COO list."""
    S = np.zeros((m, n), dtype=bool)
    S[np.asarray(rows) - 1, np.asarray(cols) - 1] = True
    colptr = np.ones(n + 1, dtype=np.int64)
    rowval = []
    for c in range(n):
        rr = np.nonzero(S[:, c])[0]
        rowval.extend((rr + 1).tolist())
        colptr[c + 1] = colptr[c] + len(rr)
    return colptr, np.array(rowval, dtype=np.int64)
