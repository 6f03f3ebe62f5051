"""CPU tests of the C-ABI boundary: the shared library loads and exports every symbol the header declares, the
host logic (plan builder: layout, row order, both sparsity structures, bounds) matches the oracle, and error
handling behaves -- no compute calls (there is no GPU here and no CPU fallback in the library)."""
import ctypes as C
import re

import numpy as np
import pytest

import powersense_b200 as ps
from powersense_b200 import capi, synth
from oracle import c_oracle

from cases import FIXTURE_CASES, load_fixture, load_golden, synthetic


def _declared_symbols():
    txt = open(capi.HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/powersense_b200.h but not exported"
    assert set(names) == set(capi.EXPORTS), set(names) ^ set(capi.EXPORTS)
    assert L.ps_abi_version() == 2


@pytest.mark.parametrize("F", ps.ALL_FORMULATIONS, ids=lambda f: f.name)
@pytest.mark.parametrize("case", list(FIXTURE_CASES) + ["arpae30", "mp40_out", "hub"])
def test_structure_only_handle_matches_oracle(case, F):
    d = load_fixture(case) if case in FIXTURE_CASES else synthetic(case)
    h = capi.Handle(d, F.code, device=-2)                       # PS_DEVICE_NONE
    o = c_oracle.COracle(d, F.code)
    assert (h.n_var, h.m_nl, h.nnzJ, h.nnzH) == (o.n_var, o.m_nl, o.nnzJ, o.nnzH)
    jr, jc, hr, hc = o.structure()
    a, b = h.jacobian_structure()
    assert np.array_equal(a, jr) and np.array_equal(b, jc)
    a, b = h.hessian_structure()
    assert np.array_equal(a, hr) and np.array_equal(b, hc)
    lo, hi = o.bounds()
    a, b = h.constraint_bounds()
    assert np.array_equal(a, lo) and np.array_equal(b, hi)
    if case in FIXTURE_CASES:
        z = load_golden(case, F)
        assert np.array_equal(h.jacobian_structure()[1], z["jcol"]) and np.array_equal(h.hessian_structure()[0], z["hrow"])
    off = h.variable_offsets()
    lay = synth.variable_layout(d, F)
    names = {"theta": "a", "vi": "a", "V": "b", "vr": "b", "Wd": "Wd", "Wr": "Wr", "Wi": "Wi", "pg": "pg", "qg": "qg",
             "Pij": "F0", "Qij": "F1", "Pji": "F2", "Qji": "F3", "Irij": "F0", "Iiij": "F1", "Irji": "F2", "Iiji": "F3",
             "cg": "cg", "bss": "bss", "tgh": "tgh", "n_var": "n_var"}
    for k, v in lay.items():
        assert off[names[k]] == v, (k, off[names[k]], v)
    (ldg, ldJ, ldH), (pg, pJ, pH) = h.preferred_layout()
    info = h.plan_info()
    assert ldg % 4 == 0 and ldJ % 4 == 0 and ldH % 4 == 0 and ldJ >= h.nnzJ
    assert (info["j_br0"] + pJ) % 4 == 0 and (info["h_br0"] + pH) % 4 == 0
    # evaluation on a structure-only handle must fail loudly: there is no CPU path
    with pytest.raises(capi.PowersenseError) as e:
        h.eval_constraint(np.zeros(h.n_var))
    assert e.value.code == capi.PS_ERR_STATE
    with pytest.raises(capi.PowersenseError):
        capi.Batch(h, 4)
    h.close()


def _jump_like_order(h, seed):
    """A structure in the shape JuMP's has: the same per-row blocks, the off-diagonal part of every block in
    another order and with the other triangle for some pairs."""
    rng = np.random.default_rng(seed)
    hr, hc = h.hessian_structure()
    ptr = h.hessian_row_blocks()
    perm = np.arange(h.nnzH)
    for r in range(h.m_nl):
        e0, e1 = int(ptr[r]), int(ptr[r + 1])
        off = np.flatnonzero(hr[e0:e1] != hc[e0:e1]) + e0
        perm[off] = rng.permutation(off)
    tr, tc = hr[perm].copy(), hc[perm].copy()
    flip = (rng.random(h.nnzH) < 0.5) & (tr != tc)
    tr[flip], tc[flip] = tc[flip], tr[flip]
    return perm, tr, tc


@pytest.mark.parametrize("F", ps.ALL_FORMULATIONS, ids=lambda f: f.name)
def test_hessian_order_hook_on_the_structure(F):
    """ps_set_hessian_order / ps_match_hessian_structure (the permutation-import hook for JuMP's intra-block order, SURVEY.md
    section 7 "hard parts"): a JuMP-shaped target structure is matched block by block, reported back exactly, and reset."""
    d = load_fixture("case14")
    h = capi.Handle(d, F.code, device=-2)
    hr0, hc0 = h.hessian_structure()
    assert np.array_equal(h.hessian_order(), np.arange(1, h.nnzH + 1))
    ptr = h.hessian_row_blocks()
    assert ptr[0] == 0 and ptr[-1] == h.nnzH and np.all(np.diff(ptr) >= 0)
    perm, tr, tc = _jump_like_order(h, 7 + F.code)
    h.match_hessian_structure(tr, tc)
    assert np.array_equal(h.hessian_order(), perm + 1)
    hr, hc = h.hessian_structure()
    assert np.array_equal(hr, hr0[perm]) and np.array_equal(hc, hc0[perm])      # our triangle convention, the target's order
    assert np.array_equal(np.maximum(hr, hc), np.maximum(tr, tc)) and np.array_equal(np.minimum(hr, hc), np.minimum(tr, tc))
    h.set_hessian_order(None)
    assert np.array_equal(h.hessian_structure()[0], hr0)
    h.set_hessian_order(perm + 1)
    assert np.array_equal(h.hessian_structure()[1], hc0[perm])
    # errors: not a permutation; a pair that does not belong to the row's block
    bad = perm + 1
    if h.nnzH > 1:
        bad = bad.copy(); bad[0] = bad[1]
        with pytest.raises(capi.PowersenseError):
            h.set_hessian_order(bad)
        tr2 = tr.copy(); tr2[0] = h.n_var
        with pytest.raises(capi.PowersenseError, match="NL row 1"):
            h.match_hessian_structure(tr2, tc)
    h.close()


def test_flags_change_the_layout_like_the_reference():
    d = synthetic("arpae30")
    for F in (ps.PBRARVmodel, ps.PNPAPVmodel, ps.CBRAWVmodel):
        for facts in (True, False):
            for lin in (True, False):
                h = capi.Handle(d, F.code, facts_bshunt=facts, obj_linear=lin, device=-2)
                o = c_oracle.COracle(d, F.code, facts_bshunt=facts, obj_linear=lin)
                assert (h.n_var, h.m_nl, h.nnzJ, h.nnzH) == (o.n_var, o.m_nl, o.nnzJ, o.nnzH)
                jr, jc, hr, hc = o.structure()
                assert np.array_equal(h.jacobian_structure()[1], jc) and np.array_equal(h.hessian_structure()[1], hc)
                h.close()


def test_create_errors():
    L = capi.lib()
    d = synthetic("arpae30")
    net, keep = capi.make_network(d)
    h = capi._VP()
    assert L.ps_create(C.byref(net), 99, 0, 3, -2, C.byref(h)) == capi.PS_ERR_ARG            # unknown formulation
    assert b"formulation" in L.ps_last_error(None)
    assert L.ps_create(C.byref(net), 1, 7, 3, -2, C.byref(h)) == capi.PS_ERR_ARG             # unknown source type
    assert L.ps_create(None, 1, 0, 3, -2, C.byref(h)) == capi.PS_ERR_ARG
    bad = keep["br_t"].copy()
    bad[3] = keep["br_f"][3]                                                               # self loop
    net2, keep2 = capi.make_network(d)
    net2.br_t = bad.ctypes.data_as(C.POINTER(C.c_int64))
    assert L.ps_create(C.byref(net2), 5, 1, 3, -2, C.byref(h)) == capi.PS_ERR_NETWORK
    bad2 = keep["br_f"].copy()
    bad2[0] = d.nbus + 5                                                                   # endpoint out of range
    net3, keep3 = capi.make_network(d)
    net3.br_f = bad2.ctypes.data_as(C.POINTER(C.c_int64))
    assert L.ps_create(C.byref(net3), 5, 1, 3, -2, C.byref(h)) == capi.PS_ERR_NETWORK


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = synthetic("arpae30")
    with pytest.raises(capi.PowersenseError) as e:
        capi.Handle(d, ps.PBRARVmodel.code, device=0)
    assert e.value.code == capi.PS_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.PowersenseError):
        capi.CscPattern(3, 3, [1, 2], [1, 2])


def test_product_never_imports_the_oracle():
    """The package must not reference oracle/ anywhere (a product path through the oracle would void parity)."""
    import os
    pkg = os.path.dirname(capi.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"
