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
                  branch_order: str | None = None, nominal_taps: bool = False) -> Network:
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
    if nominal_taps:           # feasible cases: off-nominal taps / phase shifters around low-impedance loops force
        tap, shift = np.ones(nr), np.zeros(nr)     # circulating flows of hundreds of p.u. that no dispatch can serve
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


def named_case(name: str, seed: int | None = None, feasible: bool = False, **kw) -> PowersenseData:
    """feasible=True: loads, generator limits and branch ratings are re-derived from a power-flow solution so that the case
    is a feasible AC-OPF (make_feasible); the default keeps the random loads of SURVEY.md 8(d), which is all the evaluator
    benchmarks need but is in general NOT a solvable OPF."""
    nb, nr, ng = NAMED_CASES[name]
    if feasible:
        kw.setdefault("x_min", 2e-3)
        kw.setdefault("nominal_taps", True)
    net = synth_network(nb, nr, ng, seed=nb if seed is None else seed, **kw)
    if feasible:
        net = make_feasible(net, seed=nb if seed is None else seed)
    d = create_PowersenseData(net)
    if feasible:                                                 # like a case file that carries its solved power flow:
        fp = d.feasible_point = net.feasible_point               # the start values of create_opf_model! (formulations.jl:20-67)
        d.theta, d.V, d.Pg, d.Qg = fp["theta"].copy(), fp["V"].copy(), fp["pg"].copy(), fp["qg"].copy()
        d.vr, d.vi = d.V * np.cos(d.theta), d.V * np.sin(d.theta)
    return d


def make_feasible(net: Network, seed: int = 0, max_angle: float = 0.25) -> Network:
    """A feasible AC-OPF on a synthetic network: choose a dispatch (every in-service generator at 40 % of its range, pmin
    = 0), scale the random loads to it, solve the linearised (DC) power flow for the bus angles, then evaluate the AC flow
    equations of the reference (formulations.jl:309-312 at |V| = 1) at that point and DEFINE the loads as what balances
    them exactly (formulations.jl:171-292 with the residual set to zero).  Branch ratings are raised to 1.3 x the flow at
    that point and the reactive limits widened to hold the reactive dispatch, so the point is strictly inside every
    inequality: the OPF and every N-1 case that keeps the network connected start from a feasible neighbourhood."""
    import copy
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    net = copy.deepcopy(net)
    nb, ng = net.nbus, net.ngen
    net.pmin = np.zeros(ng)
    on = net.gen_status > 0
    pg = np.where(on, 0.4 * net.pmax, 0.0)
    scale = pg.sum() / max(net.Pd.sum(), 1e-9)
    net.Pd, net.Qd = net.Pd * scale, net.Qd * scale
    d = create_PowersenseData(net)
    f, t = d.br[:, 0] - 1, d.br[:, 1] - 1
    # linearisation at |V| = 1, theta = 0: Pij ~ -gf - Gf + Bf (th_t - th_f), Pji ~ -gt - Gt + Bt (th_f - th_t)
    sl = net.slack - 1
    rows = np.concatenate([f, f, t, t]); cols = np.concatenate([t, f, f, t]); vals = np.concatenate([d.Bf, -d.Bf, d.Bt, -d.Bt])
    keep = rows != sl
    L = sp.coo_matrix((np.concatenate([vals[keep], [1.0]]), (np.concatenate([rows[keep], [sl]]), np.concatenate([cols[keep], [sl]]))),
                      shape=(nb, nb)).tocsc()
    lu = spla.splu(L)
    inj0 = np.zeros(nb)
    np.add.at(inj0, f, -d.gf - d.Gf)
    np.add.at(inj0, t, -d.gt - d.Gt)
    for _ in range(8):
        gen_at = np.zeros(nb)
        np.add.at(gen_at, net.gen_bus - 1, pg)
        rhs = -(gen_at + inj0 - d.Gl - net.Pd)                   # balance(theta) = gen + inj0 + L theta - Gl - Pd = 0
        rhs[sl] = 0.0
        th = lu.solve(rhs)
        spread = np.abs(th[f] - th[t]).max() if len(f) else 0.0
        if spread <= max_angle:
            break
        k = 0.8 * max_angle / spread                             # lighter loading: smaller angle differences
        pg, net.Pd, net.Qd = pg * k, net.Pd * k, net.Qd * k
    vr, vi = np.cos(th), np.sin(th)
    (Pij, Qij, Pji, Qji), _ = flows_from_voltages(d, vr, vi)
    accP, accQ = np.zeros(nb), np.zeros(nb)
    np.add.at(accP, net.gen_bus - 1, pg)
    live = np.asarray(net.br_status) > 0                         # out-of-service branches are in no adjacency list
    np.add.at(accP, f[live], Pij[live]); np.add.at(accP, t[live], Pji[live])
    np.add.at(accQ, f[live], Qij[live]); np.add.at(accQ, t[live], Qji[live])
    net.Pd = accP - d.Gl                                         # |V|^2 = 1
    # reactive side: keep the random reactive loads, let the generators of a bus cover the rest where there are any
    need = accQ + d.Bl - net.Qd                                  # = -sum(qg) needed at the bus
    ngen_at = np.zeros(nb)
    np.add.at(ngen_at, (net.gen_bus - 1)[on], 1.0)
    qg = np.where(on, -need[net.gen_bus - 1] / np.maximum(ngen_at[net.gen_bus - 1], 1.0), 0.0)
    has = ngen_at > 0
    net.Qd = np.where(has, net.Qd, accQ + d.Bl)
    net.qmax = np.maximum(net.qmax, 1.3 * np.abs(qg) + 0.1)
    net.qmin = -net.qmax
    net.pg, net.qg = pg, qg
    net.vm, net.va = np.ones(nb), th
    S = np.maximum(np.hypot(Pij, Qij), np.hypot(Pji, Qji))
    net.rate_a = np.maximum(net.rate_a, 1.3 * S + 0.05)
    net.feasible_point = {"theta": th, "V": np.ones(nb), "pg": pg, "qg": qg}
    return net


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
