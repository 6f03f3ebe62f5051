"""TEST INFRASTRUCTURE ONLY -- CPU oracle #1: a miniature JuMP-style NLP evaluator.

PARITY UNPINNED: the reference's evaluator is JuMP 0.21.10's NLPEvaluator (Manifest.toml:187-191),
which is not under /root/reference and cannot run here (no Julia).  The reference's own tests pin
nothing at derivative level; the only known answer on this path is the
14-bus objective 146.7 +- 10 % (test/opf/ACOPF_formulations.jl:9), checked in tests/.

What this file does: it restates `create_opf_model!` (src/opf/formulations.jl:3-472) line by line
as expression *trees* -- one tree per @NLexpression / @NLconstraint, parsed the way Julia parses
the source text -- and then derives everything generically from the trees with JuMP's published
algorithms:
  * Jacobian structure: per row, sorted unique variables of the tape and of every subexpression
    it depends on (JuMP _Derivatives compute_gradient_sparsity!).
  * Hessian structure: per row, the edge set of compute_hessian_sparsity (children of a node that
    is non-linear w.r.t. the output form a clique, diagonals included; linearly entering
    subexpressions contribute their own edge sets); block layout = vertices ascending as diagonal
    slots, then off-diagonal edges.  JuMP's intra-block off-diagonal order depends on Set hash
    order + acyclic colouring and cannot be reproduced without Julia; we use the documented
    canonical order (lower-triangular tuples (max,min), ascending) -- SURVEY.md C.3.
  * Values: second-order forward-mode AD on the trees (x^2 evaluated as x*x, n-ary sums and
    products folded left to right, a-b-c as (a-b)-c: C.1).
No closed-form derivative appears in this file, which is what makes it an independent check of
oracle/ps_oracle.c and of the CUDA kernels.  Pure-Python loops: small cases only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

# ---------------------------------------------------------------------------------- trees
# node := ('v', idx) | ('c', value) | ('s', subexpr_id) | ('+', [n..]) | ('-', [a, b]) |
#         ('neg', a) | ('pos', a) | ('*', [n..]) | ('^', a, p) | ('cos', a) | ('sin', a)


def V(i):
    return ('v', int(i))


def C(x):
    return ('c', float(x))


def add(*a):
    return ('+', list(a))


def sub(a, b):
    return ('-', [a, b])


def neg(a):
    return ('neg', a)


def mul(*a):
    return ('*', list(a))


def pw(a, p):
    return ('^', a, float(p))


def children(n):
    k = n[0]
    if k in ('+', '-', '*'):
        return n[1]
    if k in ('neg', 'pos', 'cos', 'sin'):
        return [n[1]]
    if k == '^':
        return [n[1], C(n[2])]
    return []


class Model:
    """Holds subexpressions and NL constraints in creation order (JuMP: @NLexpression,
    @NLconstraint)."""

    def __init__(self, n_var: int):
        self.n_var = n_var
        self.subs: List[tuple] = []
        self.rows: List[Tuple[tuple, float, float]] = []
        self._sub_vars: Dict[int, frozenset] = {}
        self._sub_edges: Dict[int, frozenset] = {}
        self._const_cache: Dict[int, bool] = {}

    def nlexpr(self, node) -> tuple:
        self.subs.append(node)
        return ('s', len(self.subs) - 1)

    def nlcon(self, lhs, sense: str, rhs):
        # JuMP moves everything to the left: lhs - rhs with bounds (0,0) / (-inf,0) (C.1)
        lo, hi = (0.0, 0.0) if sense == '==' else (-math.inf, 0.0)
        self.rows.append((sub(lhs, rhs), lo, hi))

    # ---- variables / constness
    def vars_of(self, n) -> frozenset:
        k = n[0]
        if k == 'v':
            return frozenset([n[1]])
        if k == 'c':
            return frozenset()
        if k == 's':
            if n[1] not in self._sub_vars:
                self._sub_vars[n[1]] = self.vars_of(self.subs[n[1]])
            return self._sub_vars[n[1]]
        out = frozenset()
        for c in children(n):
            out |= self.vars_of(c)
        return out

    def is_const(self, n) -> bool:
        return len(self.vars_of(n)) == 0

    # ---- Hessian edge set (compute_hessian_sparsity)
    def edges_of(self, n) -> frozenset:
        out = set()
        self._visit(n, out)
        return frozenset(out)

    def _sub_edge(self, sid) -> frozenset:
        if sid not in self._sub_edges:
            self._sub_edges[sid] = self.edges_of(self.subs[sid])
        return self._sub_edges[sid]

    def _visit(self, n, out: set):
        k = n[0]
        if k == 's':                       # enters linearly: append its edge list
            out |= self._sub_edge(n[1])
            return
        if k in ('v', 'c'):
            return
        ch = children(n)
        nonlin = False
        for c in ch:
            if self.is_const(c):
                continue                   # "definitely not nonlinear"
            if k in ('+', '-', 'neg', 'pos'):
                pass
            elif k == '*':
                if any((o is not c) and not self.is_const(o) for o in ch):
                    nonlin = True
            else:                          # ^, cos, sin
                nonlin = True
            if nonlin:
                break
        if nonlin:
            grp = sorted(set().union(*[self.vars_of(c) for c in ch]))
            for a in grp:
                for b in grp:
                    if b <= a:
                        out.add((a, b))
            return
        for c in ch:
            if not self.is_const(c):
                self._visit(c, out)

    # ---- structure
    def jacobian_structure(self):
        rows, cols, ptr = [], [], [0]
        for r, (n, _, _) in enumerate(self.rows):
            vs = sorted(self.vars_of(n))
            rows += [r + 1] * len(vs)
            cols += [v + 1 for v in vs]
            ptr.append(len(cols))
        return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64), np.array(ptr, dtype=np.int64)

    def hessian_blocks(self):
        """Per row: (vertices ascending, off-diagonal edges (i>j) ascending)."""
        blocks = []
        for (n, _, _) in self.rows:
            e = self.edges_of(n)
            verts = sorted({a for a, _ in e} | {b for _, b in e})
            off = sorted((a, b) for (a, b) in e if a != b)
            blocks.append((verts, off))
        return blocks

    def hessian_structure(self):
        rows, cols, ptr = [], [], [0]
        for verts, off in self.hessian_blocks():
            for v in verts:
                rows.append(v + 1)
                cols.append(v + 1)
            for a, b in off:
                rows.append(a + 1)
                cols.append(b + 1)
            ptr.append(len(rows))
        return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64), np.array(ptr, dtype=np.int64)

    def bounds(self):
        return (np.array([r[1] for r in self.rows]), np.array([r[2] for r in self.rows]))

    # ---- values: second-order forward AD
    def _jet(self, n, x, memo):
        k = n[0]
        if k == 'v':
            return (x[n[1]], {n[1]: 1.0}, {})
        if k == 'c':
            return (n[1], {}, {})
        if k == 's':
            if n[1] not in memo:
                memo[n[1]] = self._jet(self.subs[n[1]], x, memo)
            return memo[n[1]]
        if k == '+':
            acc = self._jet(n[1][0], x, memo)
            for c in n[1][1:]:
                acc = _jadd(acc, self._jet(c, x, memo), 1.0)
            return acc
        if k == '-':
            return _jadd(self._jet(n[1][0], x, memo), self._jet(n[1][1], x, memo), -1.0)
        if k == 'neg':
            a = self._jet(n[1], x, memo)
            return (-a[0], {i: -v for i, v in a[1].items()}, {i: -v for i, v in a[2].items()})
        if k == 'pos':
            return self._jet(n[1], x, memo)
        if k == '*':
            acc = self._jet(n[1][0], x, memo)
            for c in n[1][1:]:
                acc = _jmul(acc, self._jet(c, x, memo))
            return acc
        if k == '^':
            a = self._jet(n[1], x, memo)
            p = n[2]
            if p == 2.0:
                return _junary(a, a[0] * a[0], 2.0 * a[0], 2.0)
            return _junary(a, a[0] ** p, p * a[0] ** (p - 1), p * (p - 1) * a[0] ** (p - 2))
        if k == 'cos':
            a = self._jet(n[1], x, memo)
            return _junary(a, math.cos(a[0]), -math.sin(a[0]), -math.cos(a[0]))
        if k == 'sin':
            a = self._jet(n[1], x, memo)
            return _junary(a, math.sin(a[0]), math.cos(a[0]), -math.sin(a[0]))
        raise ValueError(k)

    def evaluate(self, x, mu=None, want_h=True):
        """Returns g [m], J [nnzJ], H [nnzH] (H[slot] = mu_row * d2 g_row / dxi dxj, single mixed
        partial for off-diagonals: SURVEY.md C.3)."""
        m = len(self.rows)
        g = np.zeros(m)
        Jv: List[float] = []
        Hv: List[float] = []
        blocks = self.hessian_blocks() if want_h else None
        memo: Dict[int, tuple] = {}
        for r, (n, _, _) in enumerate(self.rows):
            val, grad, hess = self._jet(n, x, memo)
            g[r] = val
            for v in sorted(self.vars_of(n)):
                Jv.append(grad.get(v, 0.0))
            if want_h:
                verts, off = blocks[r]
                w = 1.0 if mu is None else mu[r]
                for v in verts:
                    Hv.append(w * hess.get((v, v), 0.0))
                for a, b in off:
                    Hv.append(w * hess.get((a, b), 0.0))
                for key, hv in hess.items():      # every analytic non-zero must have a slot
                    if hv != 0.0:
                        a, b = key
                        assert (a in verts and b in verts and (a == b or (a, b) in set(off))), (r, key)
        return g, np.array(Jv), np.array(Hv)


def _jadd(a, b, s):
    g = dict(a[1])
    for i, v in b[1].items():
        g[i] = g.get(i, 0.0) + s * v
    h = dict(a[2])
    for i, v in b[2].items():
        h[i] = h.get(i, 0.0) + s * v
    return (a[0] + s * b[0], g, h)


def _jmul(a, b):
    g = {}
    for i, v in a[1].items():
        g[i] = g.get(i, 0.0) + b[0] * v
    for i, v in b[1].items():
        g[i] = g.get(i, 0.0) + a[0] * v
    h = {}
    for i, v in a[2].items():
        h[i] = h.get(i, 0.0) + b[0] * v
    for i, v in b[2].items():
        h[i] = h.get(i, 0.0) + a[0] * v
    for i, ga in a[1].items():
        for j, gb in b[1].items():
            key = (i, j) if i >= j else (j, i)
            h[key] = h.get(key, 0.0) + (2.0 if i == j else 1.0) * ga * gb
    return (a[0] * b[0], g, h)


def _junary(a, f, d1, d2):
    g = {i: d1 * v for i, v in a[1].items()}
    h = {i: d1 * v for i, v in a[2].items()}
    items = list(a[1].items())
    for p, (i, gi) in enumerate(items):
        for (j, gj) in items[p:]:
            key = (i, j) if i >= j else (j, i)
            h[key] = h.get(key, 0.0) + d2 * gi * gj
    return (f, g, h)


# ---------------------------------------------------------------------------------- the model
POLAR = ("PBRAPVmodel", "PNPAPVmodel", "PNRAPVmodel")
RECT = ("PBRARVmodel", "CBRARVmodel", "PNRARVmodel", "PBRAWVmodel", "CBRAWVmodel", "PNRAWVmodel")
WFORMS = ("PBRAWVmodel", "CBRAWVmodel", "PNRAWVmodel")


def variable_offsets(d, formulation: str, facts_bshunt=True, obj_linear=True):
    """formulations.jl:20-67 and :121 -- JuMP index = creation order (0-based offsets here)."""
    off, o = {}, 0
    nb, nr, ng = d.nbus, d.nbr, d.ngen
    blocks = []
    if formulation in POLAR:
        blocks += [("theta", nb), ("V", nb)]                       # :22-24
    else:
        blocks += [("vi", nb), ("vr", nb)]                         # :25-27 (imaginary first)
    if formulation in WFORMS:
        blocks += [("Wd", nb)]                                     # :29-31
    if formulation in ("PBRAWVmodel", "PNRAWVmodel"):
        blocks += [("Wr", nr), ("Wi", nr)]                         # :32-35
    blocks += [("pg", ng), ("qg", ng)]                             # :38-43
    if formulation in ("PBRAPVmodel", "PBRARVmodel", "PBRAWVmodel"):
        blocks += [("Pij", nr), ("Qij", nr), ("Pji", nr), ("Qji", nr)]   # :46-51
    if formulation in ("CBRARVmodel", "CBRAWVmodel"):
        blocks += [("Irij", nr), ("Iiij", nr), ("Irji", nr), ("Iiji", nr)]  # :52-57
    if obj_linear:
        blocks += [("cg", ng)]                                     # :60-62
    if facts_bshunt:
        blocks += [("bss", d.nbss)]                                # :65-67
    if obj_linear:
        blocks += [("tgh", int(np.sum(d.lps)))]                    # :121
    for name, n in blocks:
        off[name] = o
        o += n
    off["n_var"] = o
    return off


def build_model(d, formulation: str, facts_bshunt=True, obj_linear=True) -> Model:
    """NL part of create_opf_model! (formulations.jl:170-428).  `d` is any object with the
    PowersenseData fields (Gf.., br (nbr,2) 1-based, gen/sij/sji/sh CSR lists, Pd.., source_type)."""
    F = formulation
    off = variable_offsets(d, F, facts_bshunt, obj_linear)
    m = Model(off["n_var"])
    X = lambda name, i: V(off[name] + i - 1)                       # 1-based element -> variable node
    lst = lambda ptr, idx, i: [int(e) for e in idx[ptr[i]:ptr[i + 1]]]
    has_nl_bus = F not in ("PBRAWVmodel", "PNRAWVmodel")           # :171-172, :290-296
    for i in range(1, d.nbus + 1):                                 # :170
        if not has_nl_bus:
            break
        P = m.nlexpr(C(0.0))                                       # :171
        Q = m.nlexpr(C(0.0))                                       # :172
        gens = lst(d.gen_ptr, d.gen_idx, i - 1)
        Sij = lst(d.sij_ptr, d.sij_idx, i - 1)
        Sji = lst(d.sji_ptr, d.sji_idx, i - 1)
        shs = lst(d.sh_ptr, d.sh_idx, i - 1)
        for g in gens:                                             # :177-180
            P = m.nlexpr(add(P, X("pg", g)))
            Q = m.nlexpr(add(Q, X("qg", g)))
        for side, ks in ((0, Sij), (1, Sji)):
            for k in ks:
                f, t = int(d.br[k - 1][0]), int(d.br[k - 1][1])
                a, b = (f, t) if side == 0 else (t, f)
                gs, bs = (d.gf[k - 1], d.bf[k - 1]) if side == 0 else (d.gt[k - 1], d.bt[k - 1])
                G, B = (d.Gf[k - 1], d.Bf[k - 1]) if side == 0 else (d.Gt[k - 1], d.Bt[k - 1])
                if F in ("PBRAPVmodel", "PBRARVmodel"):            # :188-196
                    P = m.nlexpr(add(P, X("Pij" if side == 0 else "Pji", k)))
                    Q = m.nlexpr(add(Q, X("Qij" if side == 0 else "Qji", k)))
                elif F in ("CBRARVmodel", "CBRAWVmodel"):          # :206-214
                    Ir = X("Irij" if side == 0 else "Irji", k)
                    Ii = X("Iiij" if side == 0 else "Iiji", k)
                    P = m.nlexpr(add(P, mul(X("vr", a), Ir), mul(X("vi", a), Ii)))
                    Q = m.nlexpr(sub(add(Q, mul(X("vi", a), Ir)), mul(X("vr", a), Ii)))
                elif F == "PNPAPVmodel":                           # :215-225
                    Y, thY = abs(complex(G, B)), math.atan2(B, G)
                    ang = sub(sub(X("theta", a), X("theta", b)), C(thY))
                    # note :223-224 writes V[f]*V[t] on both sides (a commutative swap for side 1)
                    VfVt = (X("V", f), X("V", t))
                    P = m.nlexpr(sub(sub(P, mul(C(gs), pw(X("V", a), 2))), mul(C(Y), ('cos', ang), *VfVt)))
                    Q = m.nlexpr(sub(add(Q, mul(C(bs), pw(X("V", a), 2))), mul(C(Y), ('sin', ang), *VfVt)))
                elif F == "PNRAPVmodel":                           # :226-234
                    dl = sub(X("theta", a), X("theta", b))
                    VfVt = (X("V", f), X("V", t))
                    P = m.nlexpr(sub(sub(P, mul(C(gs), pw(X("V", a), 2))),
                                     mul(add(mul(C(G), ('cos', dl)), mul(C(B), ('sin', dl))), *VfVt)))
                    Q = m.nlexpr(add(add(Q, mul(C(bs), pw(X("V", a), 2))),
                                     mul(sub(mul(C(B), ('cos', dl)), mul(C(G), ('sin', dl))), *VfVt)))
                elif F == "PNRARVmodel":                           # :235-243
                    A_ = add(pw(X("vr", a), 2), pw(X("vi", a), 2))
                    C_ = add(mul(X("vr", a), X("vr", b)), mul(X("vi", a), X("vi", b)))
                    S_ = sub(mul(X("vr", a), X("vi", b)), mul(X("vi", a), X("vr", b)))
                    P = m.nlexpr(add(sub(sub(P, mul(C(gs), A_)), mul(C(G), C_)), mul(C(B), S_)))
                    Q = m.nlexpr(add(Q, mul(C(bs), A_), mul(C(B), C_), mul(C(G), S_)))
        if F in POLAR:                                             # :256-263
            P = m.nlexpr(sub(P, mul(C(d.Gl[i - 1]), pw(X("V", i), 2))))
            Q = m.nlexpr(add(Q, mul(C(d.Bl[i - 1]), pw(X("V", i), 2))))
            if facts_bshunt:
                for sh in shs:
                    Q = m.nlexpr(add(Q, mul(X("bss", sh), pw(X("V", i), 2))))
        elif F in ("PBRARVmodel", "CBRARVmodel", "PNRARVmodel"):   # :264-271
            A_ = add(pw(X("vr", i), 2), pw(X("vi", i), 2))
            P = m.nlexpr(sub(P, mul(C(d.Gl[i - 1]), A_)))
            Q = m.nlexpr(add(Q, mul(C(d.Bl[i - 1]), A_)))
            if facts_bshunt:
                for sh in shs:
                    Q = m.nlexpr(add(Q, mul(X("bss", sh), A_)))
        elif F == "CBRAWVmodel":                                   # :272-279
            P = m.nlexpr(sub(P, mul(C(d.Gl[i - 1]), X("Wd", i))))
            Q = m.nlexpr(add(Q, mul(C(d.Bl[i - 1]), X("Wd", i))))
            if facts_bshunt:
                for sh in shs:
                    Q = m.nlexpr(add(Q, mul(X("bss", sh), X("Wd", i))))
        m.nlcon(P, '==', C(d.Pd[i - 1]))                           # :291
        m.nlcon(Q, '==', C(d.Qd[i - 1]))                           # :292
    matpower = d.source_type == "matpower"                         # :299
    for k in range(1, d.nbr + 1):                                  # :300 / :364
        f, t = int(d.br[k - 1][0]), int(d.br[k - 1][1])
        Im2 = pw(C(d.Imax[k - 1]), 2)
        sides = ((f, t, d.gf[k - 1], d.bf[k - 1], d.Gf[k - 1], d.Bf[k - 1]),
                 (t, f, d.gt[k - 1], d.bt[k - 1], d.Gt[k - 1], d.Bt[k - 1]))
        if F == "PBRAPVmodel":                                     # :302-307 / :366-371
            VfVt = (X("V", f), X("V", t))
            rhsP, rhsQ = [], []
            for (a, b, gs, bs, G, B) in sides:
                dl = sub(X("theta", a), X("theta", b))
                rhsP.append(sub(mul(neg(C(gs)), pw(X("V", a), 2)),
                                mul(add(mul(C(G), ('cos', dl)), mul(C(B), ('sin', dl))), *VfVt)))
                rhsQ.append(add(mul(C(bs), pw(X("V", a), 2)),
                                mul(sub(mul(C(B), ('cos', dl)), mul(C(G), ('sin', dl))), *VfVt)))
            m.nlcon(X("Pij", k), '==', rhsP[0])
            m.nlcon(X("Pji", k), '==', rhsP[1])
            m.nlcon(X("Qij", k), '==', rhsQ[0])
            m.nlcon(X("Qji", k), '==', rhsQ[1])
            for (P_, Q_, a) in (("Pij", "Qij", f), ("Pji", "Qji", t)):
                lhs = add(pw(X(P_, k), 2), pw(X(Q_, k), 2))
                m.nlcon(lhs, '<=', Im2 if matpower else mul(Im2, pw(X("V", a), 2)))
        elif F == "PBRARVmodel":                                   # :309-314 / :373-378
            rhsP, rhsQ = [], []
            for (a, b, gs, bs, G, B) in sides:
                A_ = add(pw(X("vr", a), 2), pw(X("vi", a), 2))
                C_ = add(mul(X("vr", a), X("vr", b)), mul(X("vi", a), X("vi", b)))
                S_ = sub(mul(X("vr", a), X("vi", b)), mul(X("vi", a), X("vr", b)))
                rhsP.append(add(sub(mul(neg(C(gs)), A_), mul(C(G), C_)), mul(C(B), S_)))
                rhsQ.append(add(mul(('pos', C(bs)), A_), mul(C(B), C_), mul(C(G), S_)))
            m.nlcon(X("Pij", k), '==', rhsP[0])
            m.nlcon(X("Pji", k), '==', rhsP[1])
            m.nlcon(X("Qij", k), '==', rhsQ[0])
            m.nlcon(X("Qji", k), '==', rhsQ[1])
            for (P_, Q_, a) in (("Pij", "Qij", f), ("Pji", "Qji", t)):
                lhs = add(pw(X(P_, k), 2), pw(X(Q_, k), 2))
                m.nlcon(lhs, '<=', Im2 if matpower else mul(Im2, add(pw(X("vr", a), 2), pw(X("vi", a), 2))))
        elif F == "PBRAWVmodel":                                   # :316,321-322 / :380,385-386
            m.nlcon(add(mul(X("Wr", k), X("Wr", k)), mul(X("Wi", k), X("Wi", k))), '==', mul(X("Wd", f), X("Wd", t)))
            for (P_, Q_, a) in (("Pij", "Qij", f), ("Pji", "Qji", t)):
                lhs = add(pw(X(P_, k), 2), pw(X(Q_, k), 2))
                m.nlcon(lhs, '<=', Im2 if matpower else mul(Im2, X("Wd", a)))
        elif F in ("CBRARVmodel", "CBRAWVmodel"):                  # :328-329,335-336 / :392-393,399-400
            for (Ir_, Ii_, a) in (("Irij", "Iiij", f), ("Irji", "Iiji", t)):
                if not matpower:
                    lhs = add(pw(X(Ir_, k), 2), pw(X(Ii_, k), 2))
                elif F == "CBRARVmodel":
                    A1 = add(pw(X("vr", a), 2), pw(X("vi", a), 2))
                    A2 = add(pw(X("vr", a), 2), pw(X("vi", a), 2))
                    lhs = add(mul(pw(X(Ir_, k), 2), A1), mul(pw(X(Ii_, k), 2), A2))
                else:
                    lhs = add(mul(pw(X(Ir_, k), 2), X("Wd", a)), mul(pw(X(Ii_, k), 2), X("Wd", a)))
                m.nlcon(lhs, '<=', Im2)
        elif F in ("PNPAPVmodel", "PNRAPVmodel"):                  # :337-343 / :401-407
            for (a, b, gs, bs, G, B) in sides:
                y1 = gs * gs + bs * bs                                  # abs2
                y2 = G * G + B * B
                y3 = 2 * abs(complex(gs, bs)) * abs(complex(G, B))
                th1, th2 = math.atan2(bs, gs), math.atan2(B, G)
                ang = sub(add(sub(X("theta", a), X("theta", b)), C(th1)), C(th2))
                if matpower:
                    lhs = add(mul(C(y1), pw(X("V", a), 4)),
                              mul(C(y2), pw(X("V", b), 2), pw(X("V", a), 2)),
                              mul(C(y3), pw(X("V", a), 3), X("V", b), ('cos', ang)))
                else:
                    lhs = add(mul(C(y1), pw(X("V", a), 2)), mul(C(y2), pw(X("V", b), 2)),
                              mul(C(y3), X("V", a), X("V", b), ('cos', ang)))
                m.nlcon(lhs, '<=', Im2)
        elif F == "PNRARVmodel":                                   # :344-352 / :408-414
            for (a, b, gs, bs, G, B) in sides:
                y, Y = gs * gs + bs * bs, G * G + B * B
                y1, y2 = 2 * (gs * G + bs * B), 2 * (bs * G - gs * B)
                A_ = lambda: add(pw(X("vr", a), 2), pw(X("vi", a), 2))
                B_ = add(pw(X("vr", b), 2), pw(X("vi", b), 2))
                C_ = add(mul(X("vr", a), X("vr", b)), mul(X("vi", a), X("vi", b)))
                S_ = sub(mul(X("vr", a), X("vi", b)), mul(X("vi", a), X("vr", b)))
                if matpower:
                    lhs = add(mul(C(y1), C_, A_()), mul(C(y2), S_, A_()), mul(C(y), A_(), A_()), mul(C(Y), B_, A_()))
                else:
                    lhs = add(mul(C(y1), C_), mul(C(y2), S_), mul(C(y), A_()), mul(C(Y), B_))
                m.nlcon(lhs, '<=', Im2)
        elif F == "PNRAWVmodel":                                   # :354 / :416
            m.nlcon(add(mul(X("Wr", k), X("Wr", k)), mul(X("Wi", k), X("Wi", k))), '==', mul(X("Wd", f), X("Wd", t)))
    return m
