"""PowersenseData tables as struct-of-arrays FP64 buffers.

Host-side mirror of the reference data layer (src/opf/PowersenseData.jl:4-214): the same field
names, the same branch-parameter arithmetic (PowersenseData.jl:191-211, quirks included, see
SURVEY.md Appendix D) and the same adjacency rules (GO_read2.jl:85-94 / GO_read.jl:303-319), but
every table is a flat numpy array and the per-bus adjacency Dict is four CSR lists so the
whole thing can be handed to the C ABI (include/powersense_b200.h, ps_network) without copies.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

F64 = np.float64
I64 = np.int64


@dataclass
class Network:
    """Parsed case, per-unit, internal 1-based indices (what process_data() returns upstream).

    Arrays are element-indexed in file order; `f`, `t`, `gen_bus`, `ss_bus` hold 1-based
    *internal* bus indices (reference "f_bus_index" etc.)."""
    source_type: str = "matpower"          # "matpower" | "arpae" (GO_read.jl:303)
    baseMVA: float = 100.0
    slack: int = 1
    # buses
    bus_id: np.ndarray = None
    vmax: np.ndarray = None
    vmin: np.ndarray = None
    Pd: np.ndarray = None
    Qd: np.ndarray = None
    GL: np.ndarray = None
    BL: np.ndarray = None
    vm: np.ndarray = None
    va: np.ndarray = None
    # generators
    gen_bus: np.ndarray = None
    gen_status: np.ndarray = None
    pg: np.ndarray = None
    qg: np.ndarray = None
    pmin: np.ndarray = None
    pmax: np.ndarray = None
    qmin: np.ndarray = None
    qmax: np.ndarray = None
    gen_cost: List[List[float]] = field(default_factory=list)       # matpower polynomial, highest order first
    gen_lin: List[Optional[np.ndarray]] = field(default_factory=list)  # arpae piecewise points [[p, c], ...]
    # branches
    f: np.ndarray = None
    t: np.ndarray = None
    br_r: np.ndarray = None
    br_x: np.ndarray = None
    g_fr: np.ndarray = None
    b_fr: np.ndarray = None
    g_to: np.ndarray = None
    b_to: np.ndarray = None
    tap: np.ndarray = None
    shift: np.ndarray = None
    br_status: np.ndarray = None
    rate_a: np.ndarray = None              # NaN == key absent (PowersenseData.jl:206 haskey)
    gfM: np.ndarray = None
    bfM: np.ndarray = None
    # switched shunts ("sshunt")
    ss_bus: np.ndarray = None
    ss_status: np.ndarray = None
    ss_max: np.ndarray = None
    ss_min: np.ndarray = None

    @property
    def nbus(self) -> int:
        return len(self.vmax)

    @property
    def ngen(self) -> int:
        return len(self.gen_bus)

    @property
    def nbr(self) -> int:
        return len(self.f)

    @property
    def nbss(self) -> int:
        return 0 if self.ss_bus is None else len(self.ss_bus)


def _csr(lists: Sequence[Sequence[int]]):
    ptr = np.zeros(len(lists) + 1, dtype=I64)
    for i, l in enumerate(lists):
        ptr[i + 1] = ptr[i] + len(l)
    idx = np.fromiter((e for l in lists for e in l), dtype=I64, count=int(ptr[-1]))
    return ptr, idx


@dataclass
class PowersenseData:
    """Field-for-field counterpart of `mutable struct PowersenseData` (PowersenseData.jl:4-105)
    restricted to what the OPF model builder reads.  `br` is (nbr, 2) int64, 1-based.
    `node` (PowersenseData.jl:40) is stored as CSR lists of 1-based element ids."""
    nbus: int
    ngen: int
    nbr: int
    nbss: int
    Vmax: np.ndarray
    Vmin: np.ndarray
    Pmax: np.ndarray
    Pmin: np.ndarray
    Qmax: np.ndarray
    Qmin: np.ndarray
    bmax: np.ndarray
    bmin: np.ndarray
    Imax: np.ndarray
    Gf: np.ndarray
    Bf: np.ndarray
    Gt: np.ndarray
    Bt: np.ndarray
    gf: np.ndarray
    bf: np.ndarray
    gt: np.ndarray
    bt: np.ndarray
    br: np.ndarray
    Pd: np.ndarray
    Qd: np.ndarray
    Gl: np.ndarray
    Bl: np.ndarray
    gen_ptr: np.ndarray
    gen_idx: np.ndarray
    sij_ptr: np.ndarray
    sij_idx: np.ndarray
    sji_ptr: np.ndarray
    sji_idx: np.ndarray
    sh_ptr: np.ndarray
    sh_idx: np.ndarray
    lps: np.ndarray
    pgh: List[np.ndarray]
    cgh: List[np.ndarray]
    c0: np.ndarray
    c1: np.ndarray
    c2: np.ndarray
    cost_order: int
    source_type: str
    slack: int
    # start / solution vectors (PowersenseData.jl:19-23,34-38,47-54)
    Pg: np.ndarray = None
    Qg: np.ndarray = None
    Bss: np.ndarray = None
    V: np.ndarray = None
    theta: np.ndarray = None
    Wd: np.ndarray = None
    Wr: np.ndarray = None
    Wi: np.ndarray = None
    vr: np.ndarray = None
    vi: np.ndarray = None
    cg: np.ndarray = None
    cost: float = 0.0
    model: object = None

    @property
    def node(self):
        """Same shape as the reference's `data.node[string(i)]` dict (1-based keys as str)."""
        out = {}
        for i in range(self.nbus):
            out[str(i + 1)] = {
                "gen": self.gen_idx[self.gen_ptr[i]:self.gen_ptr[i + 1]].tolist(),
                "Sij": self.sij_idx[self.sij_ptr[i]:self.sij_ptr[i + 1]].tolist(),
                "Sji": self.sji_idx[self.sji_ptr[i]:self.sji_ptr[i + 1]].tolist(),
                "shunt": self.sh_idx[self.sh_ptr[i]:self.sh_ptr[i + 1]].tolist(),
            }
        return out


def branch_parameters(net: Network):
    """PowersenseData.jl:191-211.  Returns Gf,Bf,Gt,Bt,gf,bf,gt,bt,Imax (each [nbr])."""
    st = net.br_status.astype(F64)
    y = 1.0 / (net.br_r + 1j * net.br_x)                       # :193-194
    g, b = y.real, y.imag
    alpha = 1.0 / net.tap * np.cos(-net.shift)                 # :195-197
    beta = 1.0 / net.tap * np.sin(-net.shift)
    Gf = st * (-g * alpha - b * beta)                          # :198
    Bf = st * (g * beta - b * alpha)                           # :199
    Gt = st * (-g * alpha + b * beta)                          # :200
    Bt = st * (-g * beta - b * alpha)                          # :201
    gf = st * (net.gfM + (g + net.g_fr) / (net.tap * net.tap))  # :202
    bf = st * (net.bfM + (b + net.b_fr) / (net.tap * net.tap))  # :203
    gt = st * (g + net.g_fr)                                   # :204 (uses g_fr, not g_to: replicated)
    bt = st * (b + net.b_fr)                                   # :205 (uses b_fr, not b_to: replicated)
    Imax = np.where(np.isnan(net.rate_a), 10000.0, st * np.nan_to_num(net.rate_a))  # :206
    Imax = np.where(Imax == 0, 10000.0, Imax)                  # :207
    return Gf, Bf, Gt, Bt, gf, bf, gt, bt, Imax


def adjacency(net: Network):
    """GO_read2.jl:80-94 / GO_read.jl:303-319: per-bus gen/Sij/Sji/shunt lists, ascending element
    id, only elements with status > 0."""
    nb = net.nbus
    gen = [[] for _ in range(nb)]
    sij = [[] for _ in range(nb)]
    sji = [[] for _ in range(nb)]
    sh = [[] for _ in range(nb)]
    for g in range(net.ngen):
        if net.gen_bus[g] * net.gen_status[g] > 0:
            gen[int(net.gen_bus[g]) - 1].append(g + 1)
    for k in range(net.nbr):
        if net.f[k] * net.br_status[k] > 0:
            sij[int(net.f[k]) - 1].append(k + 1)
        if net.t[k] * net.br_status[k] > 0:
            sji[int(net.t[k]) - 1].append(k + 1)
    if net.source_type != "matpower":                          # GO_read2.jl:75-76: matpower has none
        for s in range(net.nbss):
            if net.ss_bus[s] * net.ss_status[s] > 0:
                sh[int(net.ss_bus[s]) - 1].append(s + 1)
    return gen, sij, sji, sh


def create_PowersenseData(path) -> PowersenseData:
    """create_PowersenseData(path) (PowersenseData.jl:107-214).

    `path` is a MATPOWER `.m` path (str), a `[raw, rop]` pair of PSS/E-33 paths, or an already
    parsed `Network` (what the reference's process_data returns)."""
    if isinstance(path, Network):
        net = path
    else:
        from .readers import process_data
        net = process_data(path)
    nb, ng, nr = net.nbus, net.ngen, net.nbr
    nbss = net.nbss if net.source_type != "matpower" else 0     # PowersenseData.jl:109
    cost_segments = 30                                          # :110
    gst = net.gen_status.astype(F64)
    Pmin, Pmax = gst * net.pmin, gst * net.pmax                 # :125-128
    Qmin, Qmax = gst * net.qmin, gst * net.qmax
    c0, c1, c2 = np.zeros(ng), np.zeros(ng), np.zeros(ng)
    lps = np.zeros(ng, dtype=I64)
    pgh: List[np.ndarray] = []
    cgh: List[np.ndarray] = []
    cost_order = 0
    for i in range(ng):
        if net.source_type == "matpower":                       # :130-156
            cost = net.gen_cost[i]
            cost_order = max(cost_order, len(cost))             # GO_read2.jl:59
            if len(cost) == 3:
                c2[i], c1[i], c0[i] = cost
                lseg = (Pmax[i] - Pmin[i]) / (cost_segments - 1)
                p = Pmin[i] + np.arange(cost_segments) * lseg
                pts = np.stack([p, c2[i] * p ** 2 + c1[i] * p + c0[i]], axis=1)
            elif len(cost) == 2:
                c1[i], c0[i] = cost
                pts = np.array([[Pmin[i], c0[i]], [Pmax[i], c1[i]]])   # sic (:146-147)
            elif len(cost) == 1:
                c0[i] = cost[0]
                pts = np.array([[Pmin[i], c0[i]]])
            else:
                pts = np.array([[Pmin[i], c0[i]]])
        else:
            pts = np.array(net.gen_lin[i], dtype=F64).reshape(-1, 2)
        lps[i] = len(pts)
        if Pmin[i] <= Pmax[i] and not (Pmax[i] == 0 and Pmin[i] == 0):      # :162-175
            cc = np.where(pts[:, 1] < 10e-6, 10e-6, pts[:, 1])
            pp = pts[:, 0].copy()
            if Pmin[i] < pp[0]:
                pp[0] = Pmin[i]
            if Pmax[i] > pp[-1]:
                pp[-1] = Pmax[i]
            cgh.append(cc)
            pgh.append(pp)
        elif Pmax[i] == 0 and Pmin[i] == 0:                                   # :176-180
            lps[i] = 1
            cgh.append(np.zeros(1))
            pgh.append(np.zeros(1))
        else:                                                   # Pmin > Pmax: reference leaves them empty
            cgh.append(np.zeros(0))
            pgh.append(np.zeros(0))
    Gf, Bf, Gt, Bt, gf, bf, gt, bt, Imax = branch_parameters(net)
    gen, sij, sji, sh = adjacency(net)
    gp, gi = _csr(gen)
    ijp, iji = _csr(sij)
    jip, jii = _csr(sji)
    shp, shi = _csr(sh)
    if nbss:
        sst = net.ss_status.astype(F64)
        bmax, bmin = sst * net.ss_max, sst * net.ss_min         # :112-116
    else:
        bmax, bmin = np.zeros(0), np.zeros(0)
    vm = net.vm if net.vm is not None else np.ones(nb)
    va = net.va if net.va is not None else np.zeros(nb)
    d = PowersenseData(
        nbus=nb, ngen=ng, nbr=nr, nbss=nbss,
        Vmax=net.vmax.astype(F64), Vmin=net.vmin.astype(F64),
        Pmax=Pmax, Pmin=Pmin, Qmax=Qmax, Qmin=Qmin, bmax=bmax, bmin=bmin, Imax=Imax,
        Gf=Gf, Bf=Bf, Gt=Gt, Bt=Bt, gf=gf, bf=bf, gt=gt, bt=bt,
        br=np.stack([net.f.astype(I64), net.t.astype(I64)], axis=1),
        Pd=net.Pd.astype(F64), Qd=net.Qd.astype(F64), Gl=net.GL.astype(F64), Bl=net.BL.astype(F64),
        gen_ptr=gp, gen_idx=gi, sij_ptr=ijp, sij_idx=iji, sji_ptr=jip, sji_idx=jii,
        sh_ptr=shp, sh_idx=shi, lps=lps, pgh=pgh, cgh=cgh, c0=c0, c1=c1, c2=c2,
        cost_order=cost_order if net.source_type == "matpower" else 0,
        source_type=net.source_type, slack=int(net.slack),
    )
    # start values exactly as the struct constructor leaves them (PowersenseData.jl:96-104)
    d.Pg, d.Qg, d.Bss = np.zeros(ng), np.zeros(ng), np.zeros(nbss)
    d.V, d.theta = np.ones(nb), np.zeros(nb)
    d.Wd, d.Wr, d.Wi = np.ones(nb), np.ones(nr), np.zeros(nr)
    d.vr, d.vi = np.ones(nb), np.zeros(nb)
    d.cg = np.zeros(ng)
    return d
