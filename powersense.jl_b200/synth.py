"""Seeded synthetic networks and scenario batches.

The MATPOWER pegase / case118 files are not shipped with the reference (only the 14-bus PSS/E
case, case300.m and case3.m are), so the named BASELINE cases are synthesised to their published
dimensions: a connected graph with exactly nb buses / nr branches (random spanning tree with a
locality bias, extra edges, ~1.5 % parallel branches, ~12 % transformers, ~10 % of those phase
shifting).  Everything is reproducible from the seed."""
from __future__ import annotations

import numpy as np

from .data import Network, PowersenseData, create_PowersenseData
from .types import ACOPF_fromulation

NAMED_CASES = {            # published MATPOWER dimensions: (nbus, nbranch, ngen)
    "case14": (14, 20, 5),
    "case118": (118, 186, 54),
    "case300": (300, 411, 69),
    "case2869pegase": (2869, 4582, 510),
    "case9241pegase": (9241, 16049, 1445),
    "case13659pegase": (13659, 20467, 4092),
}


def synth_network(nb: int, nr: int, ng: int, seed: int = 0, source_type: str = "matpower",
                  frac_out: float = 0.0, nbss: int = 0, frac_parallel: float = 0.015, x_min: float = 1e-4,
                  branch_order: str | None = None) -> Network:
    """branch_order: "random" (default) permutes the branch list -- the adversarial order for every kernel that gathers by
    branch or by end bus, and the one all reported numbers use; "file" lists the branches the way MATPOWER / PSS-E case
    files do -- ascending from-bus, with a few percent of the rows out of place (the reference's examples/opf/case300.m:
    17 of its 411 rows break the ascending order; its 14-bus case: none).  PS_SYNTH_ORDER overrides the default."""
    import os
    branch_order = branch_order or os.environ.get("PS_SYNTH_ORDER", "random")
    assert branch_order in ("file", "random")
    rng = np.random.default_rng(seed)
    assert nr >= nb - 1
    # spanning tree, parent close in index with a heavy (log-uniform) tail
    f = np.zeros(nr, dtype=np.int64)
    t = np.zeros(nr, dtype=np.int64)
    for i in range(1, nb):
        d = int(np.exp(rng.uniform(0, np.log(max(i, 1) + 1e-9)))) if i > 1 else 1
        d = min(max(d, 1), i)
        f[i - 1], t[i - 1] = i - d, i
    k = nb - 1
    npar = int(round(frac_parallel * nr)) if nr - k > 4 else 0
    npar = min(npar, nr - k)
    while k < nr - npar:
        a = int(rng.integers(0, nb))
        d = int(np.exp(rng.uniform(0, np.log(nb))))
        b = a + d if rng.random() < 0.5 else a - d
        if 0 <= b < nb and b != a:
            f[k], t[k] = a, b
            k += 1
    for _ in range(npar):
        src = int(rng.integers(0, k))
        f[k], t[k] = f[src], t[src]
        k += 1
    flip = rng.random(nr) < 0.3
    f, t = np.where(flip, t, f), np.where(flip, f, t)
    perm = rng.permutation(nr)
    f, t = f[perm] + 1, t[perm] + 1
    if branch_order == "file":
        key = f.astype(np.float64) * (nb + 1) + t
        moved = rng.random(nr) < 0.04                        # rows that break the ascending order
        key[moved] = rng.uniform(0.0, float(nb + 1) * (nb + 1), int(moved.sum()))
        order = np.argsort(key, kind="stable")
        f, t = f[order], t[order]
    x = np.exp(rng.uniform(np.log(x_min), np.log(0.5), nr))
    r = x * rng.uniform(0.05, 0.5, nr)
    is_tr = rng.random(nr) < 0.12
    tap = np.where(is_tr, rng.uniform(0.9, 1.1, nr), 1.0)
    shift = np.where(is_tr & (rng.random(nr) < 0.10), rng.uniform(-0.2, 0.2, nr), 0.0)
    bch = np.where(is_tr, 0.0, rng.uniform(0.0, 0.3, nr))
    rate = rng.uniform(1.0, 20.0, nr)
    status = (rng.random(nr) >= frac_out).astype(np.int64)
    gen_bus = rng.integers(1, nb + 1, ng).astype(np.int64)
    gen_status = (rng.random(ng) >= min(frac_out, 0.5)).astype(np.int64)
    pmax = rng.uniform(0.5, 5.0, ng)
    pmin = pmax * rng.uniform(0.0, 0.3, ng)
    qmax = rng.uniform(0.2, 2.0, ng)
    loaded = rng.random(nb) < 0.6
    Pd = np.where(loaded, rng.uniform(0.0, 2.0, nb), 0.0)
    Qd = Pd * rng.uniform(-0.2, 0.5, nb)
    BL = np.where(rng.random(nb) < 0.05, rng.uniform(-0.5, 0.5, nb), 0.0)
    GL = np.where(rng.random(nb) < 0.02, rng.uniform(0.0, 0.1, nb), 0.0)
    costs = [[float(rng.uniform(0.01, 0.2)), float(rng.uniform(1, 40)), float(rng.uniform(0, 100))] for _ in range(ng)]
    lin = None
    if source_type != "matpower":
        lin = []
        for g in range(ng):
            p = np.linspace(pmin[g], pmax[g], 10)
            lin.append(np.stack([p, costs[g][0] * p * p + costs[g][1] * p + costs[g][2]], axis=1))
    ss_bus = rng.integers(1, nb + 1, nbss).astype(np.int64)
    return Network(
        source_type=source_type, baseMVA=100.0, slack=int(gen_bus[0]) if ng else 1,
        bus_id=np.arange(1, nb + 1, dtype=np.int64), vmax=np.full(nb, 1.1), vmin=np.full(nb, 0.9),
        Pd=Pd, Qd=Qd, GL=GL, BL=BL, vm=np.ones(nb), va=np.zeros(nb),
        gen_bus=gen_bus, gen_status=gen_status, pg=np.zeros(ng), qg=np.zeros(ng),
        pmin=pmin, pmax=pmax, qmin=-qmax, qmax=qmax, gen_cost=costs, gen_lin=lin if lin is not None else [None] * ng,
        f=f, t=t, br_r=r, br_x=x, g_fr=np.zeros(nr), b_fr=bch / 2, g_to=np.zeros(nr), b_to=bch / 2,
        tap=tap, shift=shift, br_status=status, rate_a=rate, gfM=np.zeros(nr), bfM=np.zeros(nr),
        ss_bus=ss_bus, ss_status=np.ones(nbss, dtype=np.int64),
        ss_max=rng.uniform(0.0, 0.5, nbss), ss_min=-rng.uniform(0.0, 0.5, nbss),
    )


def named_case(name: str, seed: int | None = None, **kw) -> PowersenseData:
    nb, nr, ng = NAMED_CASES[name]
    return create_PowersenseData(synth_network(nb, nr, ng, seed=nb if seed is None else seed, **kw))


def bfs_depth(d: PowersenseData) -> np.ndarray:
    nb = d.nbus
    adj = [[] for _ in range(nb)]
    for a, b in d.br:
        adj[a - 1].append(b - 1)
        adj[b - 1].append(a - 1)
    depth = np.full(nb, -1, dtype=np.int64)
    for root in range(nb):
        if depth[root] >= 0:
            continue
        depth[root] = 0
        q = [root]
        while q:
            nq = []
            for u in q:
                for v in adj[u]:
                    if depth[v] < 0:
                        depth[v] = depth[u] + 1
                        nq.append(v)
            q = nq
    return depth


def flows_from_voltages(d: PowersenseData, vr: np.ndarray, vi: np.ndarray):
    """Defining expressions of the flow / current variables (formulations.jl:309-312, 324-327),
    vectorised over the leading scenario axis.  Returns (Pij,Qij,Pji,Qji), (Irij,Iiij,Irji,Iiji)."""
    f, t = d.br[:, 0] - 1, d.br[:, 1] - 1
    rf, if_, rt, it = vr[..., f], vi[..., f], vr[..., t], vi[..., t]
    Af, At = rf * rf + if_ * if_, rt * rt + it * it
    C = rf * rt + if_ * it
    Sf = rf * it - if_ * rt
    St = -Sf
    Pij = -d.gf * Af - d.Gf * C + d.Bf * Sf
    Qij = d.bf * Af + d.Bf * C + d.Gf * Sf
    Pji = -d.gt * At - d.Gt * C + d.Bt * St
    Qji = d.bt * At + d.Bt * C + d.Gt * St
    Irij = -d.gf * rf + d.bf * if_ - d.Gf * rt + d.Bf * it
    Irji = -d.gt * rt + d.bt * it - d.Gt * rf + d.Bt * if_
    Iiij = -d.gf * if_ - d.bf * rf - d.Gf * it - d.Bf * rt
    Iiji = -d.gt * it - d.bt * rt - d.Gt * if_ - d.Bt * rf
    return (Pij, Qij, Pji, Qji), (Irij, Iiij, Irji, Iiji)


def variable_layout(d: PowersenseData, formulation: ACOPF_fromulation, facts_bshunt: bool = True,
                    obj_linear: bool = True):
    """JuMP variable order = creation order (formulations.jl:20-67, then :121); 0-based offsets."""
    name = formulation.name
    nb, nr, ng = d.nbus, d.nbr, d.ngen
    off, o = {}, 0

    def add(key, n):
        nonlocal o
        off[key] = o
        o += n
    if name in ("PBRAPVmodel", "PNPAPVmodel", "PNRAPVmodel"):
        add("theta", nb); add("V", nb)
    else:
        add("vi", nb); add("vr", nb)
    if name in ("PBRAWVmodel", "CBRAWVmodel", "PNRAWVmodel"):
        add("Wd", nb)
    if name in ("PBRAWVmodel", "PNRAWVmodel"):
        add("Wr", nr); add("Wi", nr)
    add("pg", ng); add("qg", ng)
    if name in ("PBRAPVmodel", "PBRARVmodel", "PBRAWVmodel"):
        add("Pij", nr); add("Qij", nr); add("Pji", nr); add("Qji", nr)
    if name in ("CBRARVmodel", "CBRAWVmodel"):
        add("Irij", nr); add("Iiij", nr); add("Irji", nr); add("Iiji", nr)
    if obj_linear:
        add("cg", ng)
    if facts_bshunt:
        add("bss", d.nbss)
    if obj_linear:
        add("tgh", int(np.sum(d.lps)))
    off["n_var"] = o
    return off


def scenario_batch(d: PowersenseData, formulation: ACOPF_fromulation, S: int, seed: int = 0,
                   facts_bshunt: bool = True, obj_linear: bool = True, load_sigma: float = 0.1):
    """Synthetic evaluation points for S scenarios: x [S, n_var], Pd/Qd [S, nb].
    Flow/current variables are their defining expressions times (1 + 1e-2 N(0,1))."""
    rng = np.random.default_rng(seed)
    lay = variable_layout(d, formulation, facts_bshunt, obj_linear)
    nb, nr, ng = d.nbus, d.nbr, d.ngen
    depth = bfs_depth(d)
    V = rng.uniform(0.94, 1.06, (S, nb))
    th = depth[None, :] * rng.normal(0.0, 0.01, (S, 1)) + rng.normal(0.0, 0.01, (S, nb))
    vr, vi = V * np.cos(th), V * np.sin(th)
    x = rng.normal(0.0, 1.0, (S, lay["n_var"]))              # cg / tgh etc.: anything (NL rows ignore them)
    if "theta" in lay:
        x[:, lay["theta"]:lay["theta"] + nb] = th
        x[:, lay["V"]:lay["V"] + nb] = V
    else:
        x[:, lay["vi"]:lay["vi"] + nb] = vi
        x[:, lay["vr"]:lay["vr"] + nb] = vr
    f, t = d.br[:, 0] - 1, d.br[:, 1] - 1
    if "Wd" in lay:
        x[:, lay["Wd"]:lay["Wd"] + nb] = V * V * (1 + 1e-2 * rng.normal(size=(S, nb)))
    if "Wr" in lay:
        x[:, lay["Wr"]:lay["Wr"] + nr] = (vr[:, f] * vr[:, t] + vi[:, f] * vi[:, t]) * (1 + 1e-2 * rng.normal(size=(S, nr)))
        x[:, lay["Wi"]:lay["Wi"] + nr] = (vr[:, f] * vi[:, t] - vi[:, f] * vr[:, t]) * (1 + 1e-2 * rng.normal(size=(S, nr)))
    x[:, lay["pg"]:lay["pg"] + ng] = rng.uniform(d.Pmin, np.maximum(d.Pmax, d.Pmin + 1e-9), (S, ng))
    x[:, lay["qg"]:lay["qg"] + ng] = rng.uniform(d.Qmin, np.maximum(d.Qmax, d.Qmin + 1e-9), (S, ng))
    pf, cf = flows_from_voltages(d, vr, vi)
    if "Pij" in lay:
        for key, val in zip(("Pij", "Qij", "Pji", "Qji"), pf):
            x[:, lay[key]:lay[key] + nr] = val * (1 + 1e-2 * rng.normal(size=(S, nr)))
    if "Irij" in lay:
        for key, val in zip(("Irij", "Iiij", "Irji", "Iiji"), cf):
            x[:, lay[key]:lay[key] + nr] = val * (1 + 1e-2 * rng.normal(size=(S, nr)))
    if "bss" in lay and d.nbss:
        x[:, lay["bss"]:lay["bss"] + d.nbss] = rng.uniform(d.bmin, np.maximum(d.bmax, d.bmin + 1e-9), (S, d.nbss))
    Pd = np.clip(d.Pd[None, :] * (1 + load_sigma * rng.normal(size=(S, nb))), 0.0, None)
    Qd = d.Qd[None, :] * (1 + load_sigma * rng.normal(size=(S, nb)))
    return x, Pd, Qd, lay
