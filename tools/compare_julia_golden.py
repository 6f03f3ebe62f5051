#!/usr/bin/env python
"""Pin libpowersense_b200 against the REFERENCE evaluator (JuMP's NLPEvaluator behind Powersense.jl).

The reference cannot run in the authoring container (no Julia); julia/make_golden.jl dumps its outputs wherever it can.
This tool is the other half:

    python tools/compare_julia_golden.py export  DIR      # the evaluation points of tests/golden/*.npz as raw Float64
    julia --project=<Powersense.jl> julia/make_golden.jl <Powersense.jl> DIR
    python tools/compare_julia_golden.py compare DIR      # needs a B200: values come from the library, not from a CPU path

`compare` feeds the library the reference's own PowersenseData tables (DIR/<case>_data.*.bin), then checks, per case and
formulation:  dimensions and bounds; jacobian_structure tuple for tuple; hessian_lagrangian_structure per NL-row block as
a sorted multiset of variable pairs (JuMP's intra-block order and triangle are implementation-defined);
g / J / H values, Hessian entries matched pair by pair inside each row block, against BASELINE's tolerance
|a - ref| <= 1e-10 + 1e-12 |ref|, reporting the unscaled max abs / max rel error.  It writes DIR/<case>_<form>.hperm.bin
(int64, 1-based): the permutation that makes ps_hessian_structure / ps_eval_hessian report JuMP's order
(ps_set_hessian_order), DIR/report.json, and checks the repo's readers against the reference's tables.
Committing DIR as tests/golden/julia/ turns tests/test_julia_golden.py from "skipped: parity unpinned" into the pin."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASES = ("case14", "case300", "case3")
F64_FIELDS = "Vmax Vmin Pmax Pmin Qmax Qmin bmax bmin Imax Gf Bf Gt Bt gf bf gt bt Pd Qd Gl Bl c0 c1 c2".split()
I64_FIELDS = "gen_ptr gen_idx sij_ptr sij_idx sji_ptr sji_idx sh_ptr sh_idx lps".split()
ATOL, RTOL = 1e-10, 1e-12


def _bin(path, dtype):
    return np.fromfile(path, dtype=dtype)


def export_points(out_dir: str, golden_dir: str | None = None) -> int:
    """x / mu of every committed golden vector as raw little-endian Float64 (what julia/make_golden.jl reads)."""
    golden_dir = golden_dir or os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    n = 0
    for f in sorted(os.listdir(golden_dir)):
        if not f.endswith("model.npz"):
            continue
        z = np.load(os.path.join(golden_dir, f))
        stem = os.path.join(out_dir, f[:-4])
        np.ascontiguousarray(z["x"], dtype="<f8").tofile(stem + ".x.bin")
        np.ascontiguousarray(z["mu"], dtype="<f8").tofile(stem + ".mu.bin")
        n += 1
    return n


def load_julia_data(d: str, case: str):
    """PowersenseData from the tables julia/make_golden.jl dumped (None when absent)."""
    from powersense_b200.data import PowersenseData
    stem = os.path.join(d, f"{case}_data")
    if not os.path.exists(stem + ".meta.bin"):
        return None
    nbus, ngen, nbr, nbss, cost_order, slack = (int(v) for v in _bin(stem + ".meta.bin", "<i8"))
    kw = {k: _bin(f"{stem}.{k}.bin", "<f8") for k in F64_FIELDS}
    kw.update({k: _bin(f"{stem}.{k}.bin", "<i8") for k in I64_FIELDS})
    kw["br"] = _bin(stem + ".br.bin", "<i8").reshape(nbr, 2)
    lens = _bin(stem + ".pgh_len.bin", "<i8")
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(int)
    pf, cf = _bin(stem + ".pgh_flat.bin", "<f8"), _bin(stem + ".cgh_flat.bin", "<f8")
    pgh = [pf[offs[i]:offs[i + 1]] for i in range(len(lens))]
    cgh = [cf[offs[i]:offs[i + 1]] for i in range(len(lens))]
    data = PowersenseData(nbus=nbus, ngen=ngen, nbr=nbr, nbss=nbss, pgh=pgh, cgh=cgh, cost_order=cost_order,
                          source_type=open(stem + ".source_type.txt").read().strip(), slack=slack, **kw)
    return data


def load_julia_outputs(d: str, case: str, form: str):
    stem = os.path.join(d, f"{case}_{form}")
    if not os.path.exists(stem + ".g.bin"):
        return None
    out = {k: _bin(f"{stem}.{k}.bin", "<f8") for k in ("x", "mu", "g", "J", "H", "lo", "hi")}
    out.update({k: _bin(f"{stem}.{k}.bin", "<i8") for k in ("jrow", "jcol", "hrow", "hcol")})
    return out


def match_hessian_blocks(ptr, hr, hc, tr, tc):
    """Canonicalise every NL-row block as a sorted multiset of (max, min) pairs.  Returns (perm, problems): perm[i] = 0-based
    position, in OUR structure (hr, hc), of the target's entry i; problems = [(row, ours-only pairs, theirs-only pairs)]."""
    perm = np.full(len(tr), -1, dtype=np.int64)
    problems = []
    for r in range(len(ptr) - 1):
        e0, e1 = int(ptr[r]), int(ptr[r + 1])
        ours = {}
        for e in range(e0, e1):
            ours.setdefault((max(hr[e], hc[e]), min(hr[e], hc[e])), []).append(e)
        theirs_only = []
        for i in range(e0, min(e1, len(tr))):
            key = (max(tr[i], tc[i]), min(tr[i], tc[i]))
            lst = ours.get(key)
            if lst:
                perm[i] = lst.pop(0)
            else:
                theirs_only.append(key)
        ours_only = [k for k, v in ours.items() for _ in v]
        if ours_only or theirs_only:
            problems.append((r + 1, ours_only, theirs_only))
    return perm, problems


def value_report(a, ref):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if a.size == 0:
        return {"n": 0, "max_abs": 0.0, "max_rel": 0.0, "n_outside_tol": 0, "worst": []}
    err = np.abs(a - ref)
    rel = err / np.maximum(np.abs(ref), 1e-300)
    bad = np.flatnonzero(err > ATOL + RTOL * np.abs(ref))
    worst = bad[np.argsort(-err[bad])][:10]
    return {"n": int(a.size), "max_abs": float(err.max()), "max_rel": float(rel[np.abs(ref) > 0].max()) if np.any(np.abs(ref) > 0) else 0.0,
            "n_outside_tol": int(bad.size), "worst": [(int(i), float(a[i]), float(ref[i])) for i in worst]}


def compare_case(data, jl, form_code, evaluate):
    """One (case, formulation).  evaluate(data, code, x, mu) -> (handle-like with structure methods, g, J, H in the
    library's canonical order).  Returns the report dict and the 1-based Hessian permutation (None if blocks differ)."""
    h, g, J, H = evaluate(data, form_code, jl["x"], jl["mu"])
    rep = {"dims": {"ours": [h.n_var, h.m_nl, h.nnzJ, h.nnzH], "julia": [len(jl["x"]), len(jl["g"]), len(jl["J"]), len(jl["H"])]}}
    rep["dims"]["equal"] = rep["dims"]["ours"] == rep["dims"]["julia"]
    if not rep["dims"]["equal"]:
        return rep, None
    lo, hi = h.constraint_bounds()
    rep["bounds_equal"] = bool(np.array_equal(lo, jl["lo"]) and np.array_equal(hi, jl["hi"]))
    jr, jc = h.jacobian_structure()
    rep["jacobian_structure_equal"] = bool(np.array_equal(jr, jl["jrow"]) and np.array_equal(jc, jl["jcol"]))
    hr, hc = h.hessian_structure()
    ptr = h.hessian_row_blocks()
    perm, problems = match_hessian_blocks(ptr, hr, hc, jl["hrow"], jl["hcol"])
    rep["hessian_blocks_equal"] = not problems
    rep["hessian_block_problems"] = [(r, [list(map(int, p)) for p in a[:5]], [list(map(int, p)) for p in b[:5]]) for r, a, b in problems[:10]]
    rep["hessian_order_identical"] = bool(not problems and np.array_equal(perm, np.arange(len(perm))))
    rep["hessian_triangle_identical"] = bool(not problems and np.array_equal(hr[perm], jl["hrow"]) and np.array_equal(hc[perm], jl["hcol"]))
    rep["g"] = value_report(g, jl["g"])
    rep["J"] = value_report(J, jl["J"]) if rep["jacobian_structure_equal"] else None
    rep["H"] = value_report(H[perm], jl["H"]) if not problems else None
    rep["within_tolerance"] = bool(rep["g"]["n_outside_tol"] == 0 and rep["J"] is not None and rep["J"]["n_outside_tol"] == 0 and
                                   rep["H"] is not None and rep["H"]["n_outside_tol"] == 0)
    return rep, (perm + 1 if not problems else None)


def compare_tables(ours, ref):
    """The repo's readers / create_PowersenseData against the reference's tables."""
    out = {}
    for k in ("nbus", "ngen", "nbr", "nbss", "cost_order", "source_type"):
        out[k] = getattr(ours, k) == getattr(ref, k)
    for k in F64_FIELDS + I64_FIELDS + ["br"]:
        a, b = np.asarray(getattr(ours, k)), np.asarray(getattr(ref, k))
        out[k] = bool(a.shape == b.shape and np.array_equal(a, b))
        if not out[k] and a.shape == b.shape and a.dtype.kind == "f":
            out[k] = f"max abs diff {float(np.max(np.abs(a - b))):.3g}"
    return out


def gpu_evaluate(data, code, x, mu):
    from powersense_b200 import capi
    h = capi.Handle(data, code, device=0)
    g, J, H = h.eval_all(x, mu)
    return h, g, J, H


def compare_dir(d: str, evaluate=gpu_evaluate, forms=None, fixture_loader=None) -> dict:
    import powersense_b200 as ps
    report = {"tolerance": f"|a - ref| <= {ATOL} + {RTOL} |ref|", "cases": {}, "tables": {}}
    for case in CASES:
        data = load_julia_data(d, case)
        if data is None:
            continue
        if fixture_loader is not None:
            report["tables"][case] = compare_tables(fixture_loader(case), data)
        for F in (forms or ps.ALL_FORMULATIONS):
            jl = load_julia_outputs(d, case, F.name)
            if jl is None:
                continue
            rep, perm = compare_case(data, jl, F.code, evaluate)
            report["cases"][f"{case}:{F.name}"] = rep
            if perm is not None:
                perm.astype("<i8").tofile(os.path.join(d, f"{case}_{F.name}.hperm.bin"))
    ok = [r.get("within_tolerance", False) and r.get("jacobian_structure_equal", False) and r.get("hessian_blocks_equal", False)
          for r in report["cases"].values()]
    report["summary"] = {"compared": len(ok), "pinned": int(sum(ok)),
                         "verdict": "parity pinned" if ok and all(ok) else ("nothing to compare" if not ok else "differences: see cases")}
    return report


def main(argv):
    if len(argv) < 3 or argv[1] not in ("export", "compare"):
        print(__doc__)
        return 2
    if argv[1] == "export":
        print(f"wrote {export_points(argv[2])} evaluation points to {argv[2]}")
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import load_fixture            # the committed tables of the repo's own readers
    rep = compare_dir(argv[2], fixture_loader=load_fixture)
    with open(os.path.join(argv[2], "report.json"), "w") as f:
        json.dump(rep, f, indent=1)
    for k, r in rep["cases"].items():
        print(k, "OK" if r.get("within_tolerance") and r.get("jacobian_structure_equal") and r.get("hessian_blocks_equal") else "DIFFERS",
              {n: (r[n]["max_abs"], r[n]["max_rel"]) for n in "gJH" if r.get(n)})
    print(rep["summary"])
    return 0 if rep["summary"]["verdict"] == "parity pinned" else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv))
