"""ctypes front-end of oracle/ps_oracle.c (TEST INFRASTRUCTURE ONLY; parity unpinned, see the
header of ps_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_P64 = C.POINTER(C.c_int64)
_PD = C.POINTER(C.c_double)


class _Net(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("nbus", "nbr", "ngen", "nbss")] + \
               [(n, _P64) for n in ("br_f", "br_t")] + \
               [(n, _PD) for n in ("Gf", "Bf", "Gt", "Bt", "gf", "bf", "gt", "bt", "Imax", "Pd", "Qd", "Gl", "Bl")] + \
               [(n, _P64) for n in ("gen_ptr", "gen_idx", "sij_ptr", "sij_idx", "sji_ptr", "sji_idx", "sh_ptr", "sh_idx", "lps")]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libps_oracle.so")
    src = os.path.join(_HERE, "ps_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libps_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.pso_create.restype = C.c_void_p
        _LIB.pso_create.argtypes = [C.POINTER(_Net), C.c_int, C.c_int, C.c_int]
        _LIB.pso_destroy.argtypes = [C.c_void_p]
        _LIB.pso_dims.argtypes = [C.c_void_p] + [_P64] * 4
        _LIB.pso_structure.argtypes = [C.c_void_p] + [_P64] * 4
        _LIB.pso_bounds.argtypes = [C.c_void_p, _PD, _PD]
        _LIB.pso_eval.argtypes = [C.c_void_p, C.c_int] + [_PD] * 7
        _LIB.pso_eval_batch.argtypes = [C.c_void_p, C.c_int, C.c_int64, _PD, C.c_int64, _PD, _PD, C.c_int64,
                                         _PD, C.c_int64, _PD, C.c_int64, _PD, C.c_int64, _PD, C.c_int64, C.c_int]
    return _LIB


def _pd(a):
    return None if a is None else a.ctypes.data_as(_PD)


def _p64(a):
    return a.ctypes.data_as(_P64)


class COracle:
    """Restatement of the NL block for one (network, formulation, source_type, flags)."""

    def __init__(self, d, formulation_code: int, facts_bshunt=True, obj_linear=True):
        L = lib()
        self._keep = {}
        net = _Net()
        net.nbus, net.nbr, net.ngen, net.nbss = d.nbus, d.nbr, d.ngen, d.nbss
        k = self._keep
        k["br_f"] = np.ascontiguousarray(d.br[:, 0], dtype=np.int64)
        k["br_t"] = np.ascontiguousarray(d.br[:, 1], dtype=np.int64)
        for n in ("Gf", "Bf", "Gt", "Bt", "gf", "bf", "gt", "bt", "Imax", "Pd", "Qd", "Gl", "Bl"):
            k[n] = np.ascontiguousarray(getattr(d, n), dtype=np.float64)
        for n in ("gen_ptr", "gen_idx", "sij_ptr", "sij_idx", "sji_ptr", "sji_idx", "sh_ptr", "sh_idx", "lps"):
            k[n] = np.ascontiguousarray(getattr(d, n), dtype=np.int64)
        for n, a in k.items():
            setattr(net, n, _p64(a) if a.dtype == np.int64 else _pd(a))
        flags = (1 if facts_bshunt else 0) | (2 if obj_linear else 0)
        src = 0 if d.source_type == "matpower" else 1
        self.h = L.pso_create(C.byref(net), formulation_code, src, flags)
        if not self.h:
            raise ValueError("pso_create failed (self-loop branch?)")
        dims = [C.c_int64() for _ in range(4)]
        L.pso_dims(self.h, *[C.byref(v) for v in dims])
        self.n_var, self.m_nl, self.nnzJ, self.nnzH = (v.value for v in dims)
        self._has_bus_rows = formulation_code not in (2, 8)      # PBRAWV / PNRAWV have no NL bus rows

    def __del__(self):
        if lib is None:
            return
        if getattr(self, "h", None):
            lib().pso_destroy(self.h)
            self.h = None

    def structure(self):
        jr, jc = np.zeros(self.nnzJ, np.int64), np.zeros(self.nnzJ, np.int64)
        hr, hc = np.zeros(self.nnzH, np.int64), np.zeros(self.nnzH, np.int64)
        lib().pso_structure(self.h, _p64(jr), _p64(jc), _p64(hr), _p64(hc))
        return jr, jc, hr, hc

    def bounds(self):
        lo, hi = np.zeros(self.m_nl), np.zeros(self.m_nl)
        lib().pso_bounds(self.h, _pd(lo), _pd(hi))
        return lo, hi

    def eval(self, x, mu=None, Pd=None, Qd=None, what=7):
        g, J, H = np.zeros(self.m_nl), np.zeros(self.nnzJ), np.zeros(self.nnzH)
        x = np.ascontiguousarray(x, dtype=np.float64)
        bad = lib().pso_eval(self.h, what, _pd(x), _pd(Pd), _pd(Qd), _pd(mu), _pd(g), _pd(J), _pd(H))
        if bad:
            raise AssertionError("analytic second derivative outside the declared Hessian structure")
        return g, J, H

    def eval_batch(self, x, mu=None, Pd=None, Qd=None, what=7, nthreads=0, out=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        S = x.shape[0]
        if out is None:
            out = (np.zeros((S, self.m_nl)), np.zeros((S, self.nnzJ)), np.zeros((S, self.nnzH)))
        g, J, H = out
        bad = lib().pso_eval_batch(self.h, what, S, _pd(x), x.shape[1], _pd(Pd), _pd(Qd),
                                   0 if Pd is None else Pd.shape[1], _pd(mu), 0 if mu is None else mu.shape[1],
                                   _pd(g), g.shape[1], _pd(J), J.shape[1], _pd(H), H.shape[1], nthreads)
        if bad:
            raise AssertionError("analytic second derivative outside the declared Hessian structure")
        return g, J, H

    def condition_scale(self, x, mu=None, Pd=None, Qd=None):
        """Per-output magnitude of the summed terms (sum of |contribution| per slot; for g the first-order
        proxy sum_j |dg/dx_j|_abs |x_j| + |load|).  Rows such as the L4 / L5 current limits
        (formulations.jl:338-352) subtract O(|y|^2) products to get an O(1) result: two FP64 evaluations that
        differ only in rounding agree to ~1e-16 of this scale, not of the result.  Tests use it to state the
        tolerance as 1e-12 relative to max(|value|, scale)."""
        _, Ja, Ha = self.eval(x, mu, Pd, Qd, what=2 | 4 | 8)
        jr, jc, _, _ = self.structure()
        gs = np.zeros(self.m_nl)
        np.add.at(gs, jr - 1, Ja * np.abs(np.asarray(x)[jc - 1]))
        nb = len(self._keep["Pd"])
        if True:
            pd = np.abs(self._keep["Pd"] if Pd is None else Pd)
            qd = np.abs(self._keep["Qd"] if Qd is None else Qd)
            if self._has_bus_rows:
                gs[0:2 * nb:2] += pd
                gs[1:2 * nb:2] += qd
        return gs, Ja, Ha
