"""Third, independent restatement of the NL rows: every row of every AC formulation written
down ONCE MORE, symbolically, straight from src/opf/formulations.jl:170-428 (sympy), differentiated by sympy and evaluated
in 30-digit arithmetic (mpmath) at the committed golden points.  Both oracles -- the expression-tree one that produced the
goldens (oracle/jump_oracle.py) and the C one (oracle/ps_oracle.c) -- must agree with it to 1e-13 (relative to the entry's
magnitude, floor 1), the Jacobian pattern must be exactly the row's free symbols in ascending variable order (JuMP's rule,
SURVEY.md C.2), and every non-zero second derivative must own a slot in the row's Hessian block.
Nothing here shares code with either oracle or with the kernels."""
import mpmath
import numpy as np
import pytest
import sympy as sp

import powersense_b200 as ps
from powersense_b200 import capi, synth
from oracle import c_oracle

from cases import load_fixture, load_golden

mpmath.mp.dps = 30
ZERO = sp.Symbol("ZERO", real=True)


def build_rows(d, F, X):
    """The NL rows `lhs - rhs` in creation order.  X(name, i) -> the sympy symbol of variable `name`, 1-based element i."""
    name = F.name
    mp_ = d.source_type == "matpower"
    cos, sin = sp.cos, sp.sin
    fl = lambda v: 0.0 if v == ZERO else float(v)
    # data parameters: exact binary64 values; a parameter that is exactly zero stays in the expression as the symbol ZERO
    # (JuMP's patterns are structural: a variable multiplied by a zero parameter keeps its slots)
    R = lambda v: sp.Float(float(v), 40) if float(v) != 0.0 else ZERO
    rows = []
    polar = name in ("PBRAPVmodel", "PNPAPVmodel", "PNRAPVmodel")
    if name not in ("PBRAWVmodel", "PNRAWVmodel"):
        for i in range(1, d.nbus + 1):                          # :170-297
            P, Q = sp.Integer(0), sp.Integer(0)
            for g in d.gen_idx[d.gen_ptr[i - 1]:d.gen_ptr[i]]:
                P += X("pg", g); Q += X("qg", g)
            for side, (ptr, idx) in enumerate(((d.sij_ptr, d.sij_idx), (d.sji_ptr, d.sji_idx))):
                for k in idx[ptr[i - 1]:ptr[i]]:
                    f, t = int(d.br[k - 1][0]), int(d.br[k - 1][1])
                    a, b = (t, f) if side else (f, t)
                    gs, bs = (d.gt, d.bt)[0 if side else 0][k - 1] if False else ((d.gt[k - 1], d.bt[k - 1]) if side else (d.gf[k - 1], d.bf[k - 1]))
                    G, B = (d.Gt[k - 1], d.Bt[k - 1]) if side else (d.Gf[k - 1], d.Bf[k - 1])
                    gs, bs, G, B = R(gs), R(bs), R(G), R(B)
                    if name in ("PBRAPVmodel", "PBRARVmodel"):                 # :188-196
                        P += X("Pji" if side else "Pij", k); Q += X("Qji" if side else "Qij", k)
                    elif name in ("CBRARVmodel", "CBRAWVmodel"):               # :206-214
                        Ir, Ii = X("Irji" if side else "Irij", k), X("Iiji" if side else "Iiij", k)
                        P += X("vr", a) * Ir + X("vi", a) * Ii
                        Q += X("vi", a) * Ir - X("vr", a) * Ii
                    elif name == "PNPAPVmodel":                                # :215-224
                        Y = R(abs(complex(fl(G), fl(B)))); thY = sp.Float(float(np.angle(complex(fl(G), fl(B)))), 40)
                        dl = X("theta", a) - X("theta", b)
                        P += -gs * X("V", a) ** 2 - Y * cos(dl - thY) * X("V", f) * X("V", t)
                        Q += bs * X("V", a) ** 2 - Y * sin(dl - thY) * X("V", f) * X("V", t)
                    elif name == "PNRAPVmodel":                                # :225-234
                        dl = X("theta", a) - X("theta", b)
                        P += -gs * X("V", a) ** 2 - (G * cos(dl) + B * sin(dl)) * X("V", f) * X("V", t)
                        Q += bs * X("V", a) ** 2 + (B * cos(dl) - G * sin(dl)) * X("V", f) * X("V", t)
                    elif name == "PNRARVmodel":                                # :235-243
                        ra, ia, rb, ib = X("vr", a), X("vi", a), X("vr", b), X("vi", b)
                        P += -gs * (ra ** 2 + ia ** 2) - G * (ra * rb + ia * ib) + B * (ra * ib - ia * rb)
                        Q += bs * (ra ** 2 + ia ** 2) + B * (ra * rb + ia * ib) + G * (ra * ib - ia * rb)
            if polar:
                m2 = X("V", i) ** 2
            elif name == "CBRAWVmodel":
                m2 = X("Wd", i)
            else:
                m2 = X("vr", i) ** 2 + X("vi", i) ** 2
            P += -R(d.Gl[i - 1]) * m2                                          # :256-279
            Q += R(d.Bl[i - 1]) * m2
            for sh in d.sh_idx[d.sh_ptr[i - 1]:d.sh_ptr[i]]:
                Q += X("bss", sh) * m2
            rows.append(P - R(d.Pd[i - 1])); rows.append(Q - R(d.Qd[i - 1]))   # :291-292
    for k in range(1, d.nbr + 1):                                              # :299-428
        f, t = int(d.br[k - 1][0]), int(d.br[k - 1][1])
        gf, bf, Gf, Bf = R(d.gf[k - 1]), R(d.bf[k - 1]), R(d.Gf[k - 1]), R(d.Bf[k - 1])
        gt, bt, Gt, Bt = R(d.gt[k - 1]), R(d.bt[k - 1]), R(d.Gt[k - 1]), R(d.Bt[k - 1])
        Im2 = R(d.Imax[k - 1]) ** 2
        if name == "PBRAPVmodel":
            thf, tht, Vf, Vt = X("theta", f), X("theta", t), X("V", f), X("V", t)
            rows.append(X("Pij", k) - (-gf * Vf ** 2 - (Gf * cos(thf - tht) + Bf * sin(thf - tht)) * Vf * Vt))
            rows.append(X("Pji", k) - (-gt * Vt ** 2 - (Gt * cos(tht - thf) + Bt * sin(tht - thf)) * Vf * Vt))
            rows.append(X("Qij", k) - (bf * Vf ** 2 + (Bf * cos(thf - tht) - Gf * sin(thf - tht)) * Vf * Vt))
            rows.append(X("Qji", k) - (bt * Vt ** 2 + (Bt * cos(tht - thf) - Gt * sin(tht - thf)) * Vf * Vt))
            rows.append(X("Pij", k) ** 2 + X("Qij", k) ** 2 - (Im2 if mp_ else Im2 * Vf ** 2))
            rows.append(X("Pji", k) ** 2 + X("Qji", k) ** 2 - (Im2 if mp_ else Im2 * Vt ** 2))
        elif name == "PBRARVmodel":
            rf, if_, rt, it = X("vr", f), X("vi", f), X("vr", t), X("vi", t)
            rows.append(X("Pij", k) - (-gf * (rf ** 2 + if_ ** 2) - Gf * (rf * rt + if_ * it) + Bf * (rf * it - if_ * rt)))
            rows.append(X("Pji", k) - (-gt * (rt ** 2 + it ** 2) - Gt * (rt * rf + it * if_) + Bt * (rt * if_ - it * rf)))
            rows.append(X("Qij", k) - (bf * (rf ** 2 + if_ ** 2) + Bf * (rf * rt + if_ * it) + Gf * (rf * it - if_ * rt)))
            rows.append(X("Qji", k) - (bt * (rt ** 2 + it ** 2) + Bt * (rt * rf + it * if_) + Gt * (rt * if_ - it * rf)))
            rows.append(X("Pij", k) ** 2 + X("Qij", k) ** 2 - (Im2 if mp_ else Im2 * (rf ** 2 + if_ ** 2)))
            rows.append(X("Pji", k) ** 2 + X("Qji", k) ** 2 - (Im2 if mp_ else Im2 * (rt ** 2 + it ** 2)))
        elif name in ("PBRAWVmodel", "PNRAWVmodel"):
            rows.append(X("Wr", k) * X("Wr", k) + X("Wi", k) * X("Wi", k) - X("Wd", f) * X("Wd", t))
            if name == "PBRAWVmodel":
                rows.append(X("Pij", k) ** 2 + X("Qij", k) ** 2 - (Im2 if mp_ else Im2 * X("Wd", f)))
                rows.append(X("Pji", k) ** 2 + X("Qji", k) ** 2 - (Im2 if mp_ else Im2 * X("Wd", t)))
        elif name in ("CBRARVmodel", "CBRAWVmodel"):
            for a, Ir, Ii in ((f, X("Irij", k), X("Iiij", k)), (t, X("Irji", k), X("Iiji", k))):
                if not mp_:
                    rows.append(Ir ** 2 + Ii ** 2 - Im2)
                else:
                    m2 = X("vr", a) ** 2 + X("vi", a) ** 2 if name == "CBRARVmodel" else X("Wd", a)
                    rows.append(Ir ** 2 * m2 + Ii ** 2 * m2 - Im2)
        elif name in ("PNPAPVmodel", "PNRAPVmodel"):
            for a, b, gs, bs, G, B in ((f, t, gf, bf, Gf, Bf), (t, f, gt, bt, Gt, Bt)):
                y1 = gs ** 2 + bs ** 2; y2 = G ** 2 + B ** 2
                y3 = R(2.0 * abs(complex(fl(gs), fl(bs))) * abs(complex(fl(G), fl(B))))
                th1 = sp.Float(float(np.angle(complex(fl(gs), fl(bs)))), 40)
                th2 = sp.Float(float(np.angle(complex(fl(G), fl(B)))), 40)
                Va, Vb = X("V", a), X("V", b)
                c = cos(X("theta", a) - X("theta", b) + th1 - th2)
                if mp_:
                    rows.append(y1 * Va ** 4 + y2 * Vb ** 2 * Va ** 2 + y3 * Va ** 3 * Vb * c - Im2)
                else:
                    rows.append(y1 * Va ** 2 + y2 * Vb ** 2 + y3 * Va * Vb * c - Im2)
        elif name == "PNRARVmodel":
            for a, b, gs, bs, G, B in ((f, t, gf, bf, Gf, Bf), (t, f, gt, bt, Gt, Bt)):
                y, Y = gs ** 2 + bs ** 2, G ** 2 + B ** 2
                y1, y2 = 2 * (gs * G + bs * B), 2 * (bs * G - gs * B)
                ra, ia, rb, ib = X("vr", a), X("vi", a), X("vr", b), X("vi", b)
                A2, B2, C, S = ra ** 2 + ia ** 2, rb ** 2 + ib ** 2, ra * rb + ia * ib, ra * ib - ia * rb
                if mp_:
                    rows.append(y1 * C * A2 + y2 * S * A2 + y * A2 * A2 + Y * B2 * A2 - Im2)
                else:
                    rows.append(y1 * C + y2 * S + y * A2 + Y * B2 - Im2)
    return rows


@pytest.mark.parametrize("F", ps.ALL_FORMULATIONS, ids=lambda f: f.name)
@pytest.mark.parametrize("case", ["case3", "case14"])
def test_sympy_restatement_agrees_with_both_oracles(case, F):
    d = load_fixture(case)
    z = load_golden(case, F)
    lay = synth.variable_layout(d, F)
    n = int(z["n_var"])
    syms = sp.symbols(f"x1:{n + 1}", real=True)
    X = lambda name, i: syms[lay[name] + int(i) - 1]
    rows = build_rows(d, F, X)
    h = capi.Handle(d, F.code, device=-2)
    assert len(rows) == h.m_nl == len(z["g"])
    xs, mu = z["x"], z["mu"]
    point = {s: mpmath.mpf(float(v)) for s, v in zip(syms, xs)}
    sub = {s: sp.Float(float(v), 40) for s, v in zip(syms, xs)}
    sub[ZERO] = sp.Integer(0)
    ev = lambda e: float(sp.N(e.xreplace(sub), 30)) if e.free_symbols else float(e)
    jr, jc = h.jacobian_structure()
    hr, hc = h.hessian_structure()
    hptr = h.hessian_row_blocks()
    index = {s: i + 1 for i, s in enumerate(syms)}
    g_s, J_s, H_s = np.zeros(h.m_nl), np.zeros(h.nnzJ), np.zeros(h.nnzH)
    e = 0
    for r, row in enumerate(rows):
        g_s[r] = ev(row)
        cols = sorted(index[s] for s in row.free_symbols if s != ZERO)
        assert jc[e:e + len(cols)].tolist() == cols and np.all(jr[e:e + len(cols)] == r + 1), (r, cols)   # JuMP's pattern rule
        grads = {c: sp.diff(row, syms[c - 1]) for c in cols}
        for c in cols:
            J_s[e] = ev(grads[c]); e += 1
        blk = {(int(max(a, b)), int(min(a, b))): q for q, (a, b) in enumerate(zip(hr[hptr[r]:hptr[r + 1]], hc[hptr[r]:hptr[r + 1]]), start=int(hptr[r]))}
        for ia, a in enumerate(cols):
            for b in cols[:ia + 1]:
                d2 = sp.diff(grads[a], syms[b - 1])
                if d2 != 0:
                    assert (a, b) in blk, (r, a, b)              # every non-zero second derivative owns a slot
                    H_s[blk[(a, b)]] = float(mu[r]) * ev(d2)
    assert e == h.nnzJ
    o = c_oracle.COracle(d, F.code)
    og, oJ, oH = o.eval(xs, mu, d.Pd, d.Qd)
    gs, Js, Hs = o.condition_scale(xs, mu, d.Pd, d.Qd)
    for name, ref, a, b, sc in (("g", g_s, z["g"], og, gs), ("J", J_s, z["J"], oJ, Js), ("H", H_s, z["H"], oH, Hs)):
        tol = 1e-13 * np.maximum(1.0, np.maximum(np.abs(ref), sc))
        assert np.all(np.abs(a - ref) <= tol), (name, "jump_oracle", float(np.max(np.abs(a - ref) / tol)))
        assert np.all(np.abs(b - ref) <= tol), (name, "ps_oracle.c", float(np.max(np.abs(b - ref) / tol)))
    h.close()
