"""Shared test inputs: the reference's shipped cases (from the committed tests/golden/*.npz tables, so that
nothing reads /root/reference at run time) and seeded synthetic networks."""
import functools
import os

import numpy as np

import powersense_b200 as ps
from powersense_b200 import synth
from powersense_b200.data import PowersenseData

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURE_CASES = ("case14", "case3", "case300")


@functools.lru_cache(maxsize=None)
def load_fixture(name: str) -> PowersenseData:
    z = np.load(os.path.join(GOLDEN, f"{name}_data.npz"))
    nbus, ngen, nbr, nbss, cost_order, slack = (int(v) for v in z["meta"])
    lens = z["pgh_len"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    pgh = [z["pgh_flat"][offs[i]:offs[i + 1]] for i in range(len(lens))]
    cgh = [z["cgh_flat"][offs[i]:offs[i + 1]] for i in range(len(lens))]
    kw = {k: z[k] for k in ("Vmax Vmin Pmax Pmin Qmax Qmin bmax bmin Imax Gf Bf Gt Bt gf bf gt bt br Pd Qd Gl Bl gen_ptr "
                            "gen_idx sij_ptr sij_idx sji_ptr sji_idx sh_ptr sh_idx lps c0 c1 c2").split()}
    d = PowersenseData(nbus=nbus, ngen=ngen, nbr=nbr, nbss=nbss, pgh=pgh, cgh=cgh, cost_order=cost_order,
                       source_type=str(z["source_type"]), slack=slack, **kw)
    d.Pg, d.Qg, d.Bss = np.zeros(ngen), np.zeros(ngen), np.zeros(nbss)
    d.V, d.theta = np.ones(nbus), np.zeros(nbus)
    d.Wd, d.Wr, d.Wi = np.ones(nbus), np.ones(nbr), np.zeros(nbr)
    d.vr, d.vi = np.ones(nbus), np.zeros(nbus)
    d.cg = np.zeros(ngen)
    return d


def load_golden(name: str, formulation) -> dict:
    return dict(np.load(os.path.join(GOLDEN, f"{name}_{formulation.name}.npz")))


@functools.lru_cache(maxsize=None)
def synthetic(name: str) -> PowersenseData:
    """Seeded synthetic stand-ins; 'arpae30' exercises the PSS/E limit family, switched
    shunts (bss variables) and out-of-service elements."""
    if name == "arpae30":
        return ps.create_PowersenseData(synth.synth_network(30, 45, 8, seed=3, source_type="arpae", nbss=4, frac_out=0.1))
    if name == "mp40_out":
        return ps.create_PowersenseData(synth.synth_network(40, 70, 9, seed=11, frac_out=0.15, frac_parallel=0.1))
    if name == "wellcond":                 # |y| <= ~50: no catastrophic cancellation in the L4 / L5 rows
        return ps.create_PowersenseData(synth.synth_network(60, 95, 12, seed=17, x_min=0.02))
    if name == "wellcond_arpae":
        return ps.create_PowersenseData(synth.synth_network(60, 95, 12, seed=18, x_min=0.02, source_type="arpae", nbss=3))
    if name == "hub":                      # one bus of degree ~60: uneven bus tiles
        net = synth.synth_network(200, 330, 20, seed=5)
        net.f[:60] = 1
        net.t[:60] = np.arange(2, 62)
        return ps.create_PowersenseData(net)
    return synth.named_case(name)


def tol_err(a, ref):
    """max |a - ref| / (1e-10 + 1e-12 |ref|): <= 1 means within BASELINE's 1e-12 relative / 1e-10 absolute."""
    a, ref = np.asarray(a), np.asarray(ref)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - ref) / (1e-10 + 1e-12 * np.abs(ref))))


def tol_err_scaled(a, ref, scale):
    """Same, with the relative part taken against max(|ref|, scale): scale is the oracle's condition scale
    (COracle.condition_scale), the magnitude of the terms the reference formula sums for that output."""
    a, ref, scale = np.asarray(a), np.asarray(ref), np.asarray(scale)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - ref) / (1e-10 + 1e-12 * np.maximum(np.abs(ref), scale))))
