"""GPU tests of the fixed-pattern COO -> CSC assembly (replaces compute_jacobian_matrix, common.jl:12-20)."""
import numpy as np
import pytest

import powersense_b200 as ps
from powersense_b200 import capi, synth
from oracle import csc_oracle

from cases import load_fixture

pytestmark = pytest.mark.gpu


def test_small_with_duplicates_and_empty_columns():
    rows = np.array([2, 1, 2, 3, 1, 2, 3], dtype=np.int64)
    cols = np.array([1, 1, 1, 4, 4, 1, 2], dtype=np.int64)        # (2,1) three times; column 3 empty
    dE = np.array([1.0, 2.0, 4.0, 8.0, 16.0, 32.0, 0.0])
    p = capi.CscPattern(3, 4, rows, cols, device=0)
    colptr, rowval = p.structure()
    cp, rv = csc_oracle.structural_csc(3, 4, rows, cols)
    assert np.array_equal(colptr, cp) and np.array_equal(rowval, rv)
    nz = p.assemble(dE)
    D = np.zeros((3, 4))
    for c in range(4):
        for q in range(colptr[c] - 1, colptr[c + 1] - 1):
            D[rowval[q] - 1, c] = nz[q]
    assert np.array_equal(D, csc_oracle.compute_jacobian_matrix(3, 4, rows, cols, dE))
    p.close()
    e = capi.CscPattern(2, 2, np.zeros(0, np.int64), np.zeros(0, np.int64), device=0)      # empty COO
    assert e.nnz == 0 and np.array_equal(e.structure()[0], [1, 1, 1])
    e.close()
    with pytest.raises(capi.PowersenseError):
        capi.CscPattern(2, 2, [3], [1], device=0)                                           # row out of range


@pytest.mark.parametrize("F", [ps.PBRARVmodel, ps.PNPAPVmodel], ids=lambda f: f.name)
def test_jacobian_of_case300_single_and_batched(F):
    """The NL Jacobian of case300 with a solver-side prefix of affine rows in front (MOI_wrapper.jl:810-830 puts the
    NL rows after them), assembled for one vector and for a device batch."""
    import torch
    d = load_fixture("case300")
    h = capi.Handle(d, F.code, device=0)
    jr, jc = h.jacobian_structure()
    m_pre = 7
    pre_r = np.array([1, 1, 3, 7, 7], dtype=np.int64)
    pre_c = np.array([5, 9, 5, 2, 2], dtype=np.int64)              # duplicate (7,2)
    rows = np.concatenate([pre_r, jr + m_pre])
    cols = np.concatenate([pre_c, jc])
    m, n = m_pre + h.m_nl, h.n_var
    p = capi.CscPattern(m, n, rows, cols, device=0)
    x, _, _, _ = synth.scenario_batch(d, F, 3, seed=2)
    S = x.shape[0]
    dE = np.zeros((S, len(rows)))
    rng = np.random.default_rng(0)
    for s in range(S):
        dE[s, :5] = rng.normal(size=5)
        dE[s, 5:] = h.eval_jacobian(x[s])
    colptr, rowval = p.structure()
    cp, rv = csc_oracle.structural_csc(m, n, rows, cols)
    assert np.array_equal(colptr, cp) and np.array_equal(rowval, rv)
    nz0 = p.assemble(dE[0])
    import scipy.sparse as sp
    A = sp.csc_matrix((nz0, rowval - 1, colptr - 1), shape=(m, n)).toarray()
    assert np.array_equal(A, csc_oracle.compute_jacobian_matrix(m, n, rows, cols, dE[0]))
    dev = torch.device("cuda:0")
    tE = torch.from_numpy(dE).to(dev)
    tz = torch.full((S, p.nnz + 2), np.nan, dtype=torch.float64, device=dev)
    p.assemble_torch(tE, tz)
    torch.cuda.synchronize()
    z = tz.cpu().numpy()
    assert np.array_equal(z[0, :p.nnz], nz0) and np.isnan(z[:, p.nnz:]).all()
    for s in range(S):
        A = sp.csc_matrix((z[s, :p.nnz], rowval - 1, colptr - 1), shape=(m, n)).toarray()
        assert np.array_equal(A, csc_oracle.compute_jacobian_matrix(m, n, rows, cols, dE[s]))
    p.close()
    h.close()
