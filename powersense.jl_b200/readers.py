"""Case readers (host-only, one-off): PSS/E-33 `.raw` + `.rop` and MATPOWER v2 `.m`.

Counterparts of src/opf/GO_read.jl:253-404 (process_raw / process_rop) and
src/opf/GO_read2.jl:1-97 (PowerModels.parse_file + update_format).  The reference locates PSS/E
sections by comma count; we use the standard "0 / END OF ..." section terminators, which yields
the same records for well-formed files, and read the same fields.
The MATPOWER branch reproduces the PowerModels 0.18 per-unit conventions Powersense relies on
(E.3); bus indices follow file order (the reference follows Dict hash order, GO_read2.jl:18-20,
which is upstream of the evaluator boundary)."""
from __future__ import annotations

import math
import re
from typing import List

import numpy as np

from .data import Network


def process_data(path) -> Network:
    """process_data(path) (GO_read2.jl:1-13)."""
    if isinstance(path, str):
        return read_matpower(path)
    return read_rop(path[1], read_raw(path[0]))


# ----------------------------------------------------------------------------- PSS/E-33
def _sections(lines: List[str], start: int):
    secs, cur = [], []
    for ln in lines[start:]:
        s = ln.strip()
        if s.upper().startswith("Q") and len(s) <= 2:
            break
        if re.match(r"^0\s*(/|$)", s):
            secs.append(cur)
            cur = []
        else:
            cur.append(ln)
    secs.append(cur)
    return secs


def _flds(ln: str):
    return [x.strip() for x in ln.split("/")[0].split(",")] if "'" not in ln else \
        [x.strip() for x in re.split(r",(?=(?:[^']*'[^']*')*[^']*$)", ln)]


def read_raw(path_raw: str) -> Network:
    lines = open(path_raw).read().splitlines()
    baseMVA = float(lines[0].split(",")[1])
    secs = _sections(lines, 3)
    bus_s, load_s, fsh_s, gen_s, br_s, tr_s = secs[0], secs[1], secs[2], secs[3], secs[4], secs[5]
    ssh_s = secs[16] if len(secs) > 16 else []
    ids, vm, va, vmax, vmin = [], [], [], [], []
    for ln in bus_s:                                            # add_buses GO_read.jl:25-58
        f = _flds(ln)
        ids.append(int(f[0])); vm.append(float(f[7])); va.append(math.radians(float(f[8])))
        vmax.append(float(f[9])); vmin.append(float(f[10]))
    nb = len(ids)
    index = {b: i + 1 for i, b in enumerate(ids)}
    Pd, Qd, GL, BL = np.zeros(nb), np.zeros(nb), np.zeros(nb), np.zeros(nb)
    for ln in load_s:                                           # add_loads :59-71
        f = _flds(ln)
        i = index[int(f[0])] - 1
        st = int(f[2])
        Pd[i] += st * float(f[5]) / baseMVA
        Qd[i] += st * float(f[6]) / baseMVA
    for ln in fsh_s:                                            # add_shunts :73-85
        f = _flds(ln)
        i = index[int(f[0])] - 1
        st = int(f[2])
        GL[i] += st * float(f[3]) / baseMVA
        BL[i] += st * float(f[4]) / baseMVA
    g = dict(bus=[], st=[], pg=[], qg=[], qt=[], qb=[], pt=[], pb=[], id=[])
    for ln in gen_s:                                            # add_gens :86-122
        f = _flds(ln)
        g["bus"].append(index[int(f[0])]); g["id"].append(f[1].strip("' "))
        g["pg"].append(float(f[2]) / baseMVA); g["qg"].append(float(f[3]) / baseMVA)
        g["qt"].append(float(f[4]) / baseMVA); g["qb"].append(float(f[5]) / baseMVA)
        g["st"].append(int(f[14])); g["pt"].append(float(f[16]) / baseMVA); g["pb"].append(float(f[17]) / baseMVA)
    slack_g = 0
    for i in range(len(g["bus"])):
        if g["st"][i] * g["pt"][i] > g["pt"][slack_g]:
            slack_g = i
    br = dict(f=[], t=[], r=[], x=[], bfr=[], tap=[], sh=[], ra=[], gfM=[], bfM=[])
    for ln in br_s:                                             # add_Lbranches :124-167 (status==1 only)
        f = _flds(ln)
        if int(f[13]) != 1:
            continue
        br["f"].append(index[int(f[0])]); br["t"].append(index[int(f[1])])
        br["r"].append(float(f[3])); br["x"].append(float(f[4])); br["bfr"].append(float(f[5]) / 2)
        br["ra"].append(float(f[6]) / baseMVA); br["tap"].append(1.0); br["sh"].append(0.0)
        br["gfM"].append(0.0); br["bfM"].append(0.0)
    for q in range(0, len(tr_s) - 3, 4):                        # add_Tbranches :168-224
        t1, t2, t3, t4 = (_flds(tr_s[q + j]) for j in range(4))
        if int(t1[11]) != 1:
            continue
        br["f"].append(index[int(t1[0])]); br["t"].append(index[int(t1[1])])
        br["r"].append(float(t2[0])); br["x"].append(float(t2[1])); br["bfr"].append(0.0)
        br["tap"].append(float(t3[0]) / float(t4[0])); br["sh"].append(float(t3[2]) * math.pi / 180)
        br["ra"].append(float(t3[3]) / baseMVA)
        br["gfM"].append(float(t1[7])); br["bfM"].append(float(t1[8]))
    ss = dict(bus=[], st=[], mx=[], mn=[])
    for ln in ssh_s:                                            # add_switch_shunts :226-251
        f = _flds(ln)
        st = int(f[3])
        mx = mn = 0.0
        for i in range(1, 9):
            n = int(float(f[2 * (4 + i)])) if len(f) > 2 * (4 + i) else 0
            b = float(f[2 * (5 + i) - 1]) if len(f) > 2 * (5 + i) - 1 else 0.0
            bl = n * b / baseMVA
            if bl > 0:
                mx += bl * st
            elif bl < 0:
                mn += bl * st
        ss["bus"].append(index[int(f[0])]); ss["st"].append(st); ss["mx"].append(mx); ss["mn"].append(mn)
    nr = len(br["f"])
    A = lambda v, dt=np.float64: np.array(v, dtype=dt)
    net = Network(
        source_type="arpae", baseMVA=baseMVA, slack=ids[g["bus"][slack_g] - 1] if g["bus"] else 1,
        bus_id=A(ids, np.int64), vmax=A(vmax), vmin=A(vmin), Pd=Pd, Qd=Qd, GL=GL, BL=BL, vm=A(vm), va=A(va),
        gen_bus=A(g["bus"], np.int64), gen_status=A(g["st"], np.int64), pg=A(g["pg"]), qg=A(g["qg"]),
        pmin=A(g["pb"]), pmax=A(g["pt"]), qmin=A(g["qb"]), qmax=A(g["qt"]),
        gen_cost=[[] for _ in g["bus"]], gen_lin=[None for _ in g["bus"]],
        f=A(br["f"], np.int64), t=A(br["t"], np.int64), br_r=A(br["r"]), br_x=A(br["x"]),
        g_fr=np.zeros(nr), b_fr=A(br["bfr"]), g_to=np.zeros(nr), b_to=A(br["bfr"]),
        tap=A(br["tap"]), shift=A(br["sh"]), br_status=np.ones(nr, dtype=np.int64), rate_a=A(br["ra"]),
        gfM=A(br["gfM"]), bfM=A(br["bfM"]),
        ss_bus=A(ss["bus"], np.int64), ss_status=A(ss["st"], np.int64), ss_max=A(ss["mx"]), ss_min=A(ss["mn"]),
    )
    net._gen_ids = g["id"]
    return net


def read_rop(path_rop: str, net: Network) -> Network:
    """process_rop GO_read.jl:340-404: dispatch units -> dispatch tables -> piecewise cost tables."""
    lines = open(path_rop).read().splitlines()
    secs = _sections(lines, 1)
    disp = [_flds(l) for l in secs[4]] if len(secs) > 4 else []     # Generator Dispatch data
    apdt = [_flds(l) for l in secs[5]] if len(secs) > 5 else []     # Active Power Dispatch Tables
    plct = secs[9] if len(secs) > 9 else []                         # Piece-wise Linear Cost Tables
    tables, q = {}, 0
    while q < len(plct):
        hdr = _flds(plct[q])
        n = int(hdr[2])
        pts = np.array([[float(v) for v in _flds(plct[q + 1 + c])[:2]] for c in range(n)])
        tables[int(hdr[0])] = pts / net.baseMVA                     # :367,370
        q += n + 1
    tbl_to_ctbl = {int(r[0]): int(r[6]) for r in apdt}
    ids = {int(b): i for i, b in enumerate(net.bus_id)}
    for r in disp:
        bus_no, gid, dsptbl = int(r[0]), r[1].strip("' "), int(r[3])
        pts = tables[tbl_to_ctbl[dsptbl]]
        for gi in range(net.ngen):
            if net.bus_id[net.gen_bus[gi] - 1] == bus_no and net._gen_ids[gi] == gid:
                net.gen_lin[gi] = pts
    _ = ids
    return net


# ----------------------------------------------------------------------------- MATPOWER v2
def _mat(text: str, name: str):
    m = re.search(r"mpc\." + name + r"\s*=\s*\[(.*?)\];", text, re.S)
    if not m:
        return np.zeros((0, 0))
    rows = []
    for ln in m.group(1).splitlines():
        ln = ln.split("%")[0].strip().rstrip(";").strip()
        if ln:
            rows.append([float(v) for v in re.split(r"[\s,]+", ln) if v])
    return np.array(rows, dtype=np.float64)


def read_matpower(path: str) -> Network:
    """MATPOWER case -> per-unit Network with PowerModels 0.18 conventions:
    loads/shunts/limits divided by baseMVA, angles in radians, tap 0 -> 1, b split b/2 per end,
    polynomial costs rescaled to per-unit with leading zeros trimmed, rate_a == 0 replaced by the
    PowerModels thermal-limit estimate |y| * vmax * sqrt(vf^2+vt^2-2 vf vt cos(theta_max))."""
    text = open(path).read()
    base = float(re.search(r"mpc\.baseMVA\s*=\s*([0-9.eE+-]+)", text).group(1))
    bus, gen, brn, gc = _mat(text, "bus"), _mat(text, "gen"), _mat(text, "branch"), _mat(text, "gencost")
    nb, ng, nr = len(bus), len(gen), len(brn)
    ids = bus[:, 0].astype(np.int64)
    index = {int(b): i + 1 for i, b in enumerate(ids)}
    slack = int(ids[np.where(bus[:, 1] == 3)[0][0]]) if np.any(bus[:, 1] == 3) else int(ids[0])
    costs: List[List[float]] = []
    for i in range(ng):
        if len(gc) > i and gc[i, 0] == 2:
            n = int(gc[i, 3])
            c = list(gc[i, 4:4 + n])
            c = [c[j] * base ** (n - 1 - j) for j in range(n)]
            while c and c[0] == 0.0:
                c.pop(0)
            if len(c) > 3:
                c = c[-3:]
            costs.append(c)
        else:
            costs.append([])
    f = np.array([index[int(v)] for v in brn[:, 0]], dtype=np.int64)
    t = np.array([index[int(v)] for v in brn[:, 1]], dtype=np.int64)
    r, x, b = brn[:, 2], brn[:, 3], brn[:, 4]
    tap = np.where(brn[:, 8] == 0, 1.0, brn[:, 8])
    shift = np.radians(brn[:, 9])
    status = brn[:, 10].astype(np.int64)
    rate_a = brn[:, 5] / base
    angmin = np.radians(brn[:, 11]) if brn.shape[1] > 12 else np.full(nr, -math.pi / 3)
    angmax = np.radians(brn[:, 12]) if brn.shape[1] > 12 else np.full(nr, math.pi / 3)
    pad = 1.0472
    angmin = np.where(angmin <= -math.pi / 2, -pad, angmin)
    angmax = np.where(angmax >= math.pi / 2, pad, angmax)
    both0 = (angmin == 0) & (angmax == 0)
    angmin = np.where(both0, -pad, angmin)
    angmax = np.where(both0, pad, angmax)
    vmaxb = bus[:, 11]
    for k in range(nr):
        if rate_a[k] <= 0.0:
            th = max(abs(angmin[k]), abs(angmax[k]))
            z = complex(r[k], x[k])
            ymag = abs(1.0 / z) if z != 0 else 0.0
            vf, vt = vmaxb[f[k] - 1], vmaxb[t[k] - 1]
            cmax = math.sqrt(vf * vf + vt * vt - 2 * vf * vt * math.cos(th))
            rate_a[k] = ymag * max(vf, vt) * cmax
    return Network(
        source_type="matpower", baseMVA=base, slack=slack,
        bus_id=ids, vmax=bus[:, 11].copy(), vmin=bus[:, 12].copy(),
        Pd=bus[:, 2] / base, Qd=bus[:, 3] / base, GL=bus[:, 4] / base, BL=bus[:, 5] / base,
        vm=bus[:, 7].copy(), va=np.radians(bus[:, 8]),
        gen_bus=np.array([index[int(v)] for v in gen[:, 0]], dtype=np.int64),
        gen_status=(gen[:, 7] > 0).astype(np.int64),
        pg=gen[:, 1] / base, qg=gen[:, 2] / base, pmin=gen[:, 9] / base, pmax=gen[:, 8] / base,
        qmin=gen[:, 4] / base, qmax=gen[:, 3] / base, gen_cost=costs, gen_lin=[None] * ng,
        f=f, t=t, br_r=r.copy(), br_x=x.copy(), g_fr=np.zeros(nr), b_fr=b / 2, g_to=np.zeros(nr), b_to=b / 2,
        tap=tap, shift=shift, br_status=status, rate_a=rate_a, gfM=np.zeros(nr), bfM=np.zeros(nr),
        ss_bus=np.zeros(0, dtype=np.int64), ss_status=np.zeros(0, dtype=np.int64),
        ss_max=np.zeros(0), ss_min=np.zeros(0),
    )
