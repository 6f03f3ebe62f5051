"""ctypes binding of the C ABI (include/powersense_b200.h -> powersense.jl_b200/libpowersense_b200.so).

This is the same binding a Julia maintainer writes with `ccall`; the tests call the
CUDA path through it.  There is no CPU fallback: if the shared library is missing, `lib()` raises,
and if no sm_100 device is usable `ps_create` fails with PS_ERR_CUDA."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PS_LIB") or os.path.join(_HERE, "libpowersense_b200.so")   # PS_LIB: development variants
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "powersense_b200.h")

PS_OK, PS_ERR_ARG, PS_ERR_NETWORK, PS_ERR_CUDA, PS_ERR_STATE = 0, 1, 2, 3, 4
PS_NSCALARS = 4
EVAL_G, EVAL_J, EVAL_H = 1, 2, 4

_P64 = C.POINTER(C.c_int64)
_PD = C.POINTER(C.c_double)
_PU8 = C.POINTER(C.c_uint8)
_VP = C.c_void_p


class PsNetwork(C.Structure):
    """struct ps_network (include/powersense_b200.h)."""
    _fields_ = ([(n, C.c_int64) for n in ("nbus", "nbr", "ngen", "nbss", "n_tgh")] +
                [(n, _P64) for n in ("br_f", "br_t")] +
                [(n, _PD) for n in ("Gf", "Bf", "Gt", "Bt", "gf", "bf", "gt", "bt", "Imax", "Pd", "Qd", "Gl", "Bl")] +
                [(n, _P64) for n in ("gen_ptr", "gen_idx", "sij_ptr", "sij_idx", "sji_ptr", "sji_idx", "sh_ptr", "sh_idx")] +
                [(n, _PD) for n in ("c0", "c1", "c2")] + [("cost_order", C.c_int32)])


class PowersenseError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"powersense_b200 error {code}: {msg}")
        self.code = code


_SIGS = {
    "ps_abi_version": (C.c_int, []),
    "ps_create": (C.c_int, [C.POINTER(PsNetwork), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_VP)]),
    "ps_destroy": (None, [_VP]),
    "ps_last_error": (C.c_char_p, [_VP]),
    "ps_dims": (C.c_int, [_VP, _P64, _P64, _P64, _P64]),
    "ps_constraint_bounds": (C.c_int, [_VP, _PD, _PD]),
    "ps_jacobian_structure": (C.c_int, [_VP, _P64, _P64]),
    "ps_hessian_structure": (C.c_int, [_VP, _P64, _P64]),
    "ps_hessian_row_blocks": (C.c_int, [_VP, _P64]),
    "ps_set_hessian_order": (C.c_int, [_VP, _P64]),
    "ps_match_hessian_structure": (C.c_int, [_VP, _P64, _P64]),
    "ps_hessian_order": (C.c_int, [_VP, _P64]),
    "ps_variable_offsets": (C.c_int, [_VP, _P64]),
    "ps_plan_info": (C.c_int, [_VP, _P64]),
    "ps_preferred_layout": (C.c_int, [_VP, _P64, _P64]),
    "ps_eval_constraint": (C.c_int, [_VP, _PD, _PD]),
    "ps_eval_jacobian": (C.c_int, [_VP, _PD, _PD]),
    "ps_eval_hessian": (C.c_int, [_VP, _PD, C.c_double, _PD, _PD]),
    "ps_eval_all": (C.c_int, [_VP, _PD, _PD, _PD, _PD, _PD]),
    "ps_eval_objective": (C.c_int, [_VP, _PD, _PD]),
    "ps_launch_count": (C.c_int64, [_VP]),
    "ps_batch_create": (C.c_int, [_VP, C.c_int64, C.POINTER(_VP)]),
    "ps_batch_destroy": (None, [_VP]),
    "ps_batch_set_loads": (C.c_int, [_VP, _PD, _PD]),
    "ps_batch_set_branch_status": (C.c_int, [_VP, _PU8]),
    "ps_batch_set_viol_tol": (C.c_int, [_VP, C.c_double]),
    "ps_batch_eval": (C.c_int, [_VP, C.c_int, C.c_int64, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64,
                                _VP, C.c_int64, _VP, C.c_int64, _VP, _VP]),
    "ps_batch_eval_host": (C.c_int, [_VP, C.c_int, C.c_int64, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64,
                                     _VP, C.c_int64, _VP, C.c_int64, _VP]),
    "ps_batch_launch_count": (C.c_int64, [_VP]),
    "ps_algorithmic_bytes": (C.c_int64, [_VP, C.c_int]),
    "ps_algorithmic_bytes_branch": (C.c_int64, [_VP, C.c_int]),
    "ps_batch_enable_timing": (C.c_int, [_VP, C.c_int]),
    "ps_batch_kernel_times": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "ps_host_alloc": (C.c_int, [C.POINTER(_VP), C.c_int64]),
    "ps_host_free": (None, [_VP]),
    "ps_host_register": (C.c_int, [_VP, C.c_int64]),
    "ps_host_unregister": (C.c_int, [_VP]),
    "ps_csc_create": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _P64, _P64, C.c_int, C.POINTER(_VP)]),
    "ps_csc_destroy": (None, [_VP]),
    "ps_csc_dims": (C.c_int, [_VP, _P64]),
    "ps_csc_structure": (C.c_int, [_VP, _P64, _P64]),
    "ps_csc_assemble": (C.c_int, [_VP, _PD, _PD]),
    "ps_csc_assemble_batch": (C.c_int, [_VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP]),
    "ps_rows_create": (C.c_int, [C.c_int64, C.c_int64, _P64, _P64, _PD, _P64, _P64, _P64, _PD, _PD, C.c_int, C.POINTER(_VP)]),
    "ps_rows_destroy": (None, [_VP]),
    "ps_rows_dims": (C.c_int, [_VP, _P64, _P64]),
    "ps_rows_jacobian_structure": (C.c_int, [_VP, _P64, _P64]),
    "ps_rows_hessian_structure": (C.c_int, [_VP, _P64, _P64]),
    "ps_rows_eval": (C.c_int, [_VP, C.c_int, _PD, _PD, _PD, _PD, _PD]),
    "ps_rows_eval_batch": (C.c_int, [_VP, C.c_int, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64,
                                     _VP, C.c_int64, _VP]),
    "ps_kkt_create": (C.c_int, [_VP, _PD, _PD, _PD, _PD, C.POINTER(_VP)]),
    "ps_kkt_destroy": (None, [_VP]),
    "ps_kkt_eval_batch": (C.c_int, [_VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP, _VP, C.c_int64,
                                    _VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP, C.c_int64, _VP, _VP, _VP]),
}
EXPORTS = tuple(_SIGS)

_LIB = None


def lib():
    """Load libpowersense_b200.so (raises if it has not been built: there is no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C powersense.jl_b200/csrc).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _pd(a):
    return None if a is None else a.ctypes.data_as(_PD)


def _p64(a):
    return a.ctypes.data_as(_P64)


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


def make_network(d):
    """PowersenseData (powersense.jl_b200/data.py) -> (PsNetwork, keep-alive dict)."""
    keep = {}
    net = PsNetwork()
    net.nbus, net.nbr, net.ngen, net.nbss = d.nbus, d.nbr, d.ngen, d.nbss
    net.n_tgh = int(np.sum(d.lps))
    keep["br_f"] = np.ascontiguousarray(d.br[:, 0], dtype=np.int64)
    keep["br_t"] = np.ascontiguousarray(d.br[:, 1], dtype=np.int64)
    for n in ("Gf", "Bf", "Gt", "Bt", "gf", "bf", "gt", "bt", "Imax", "Pd", "Qd", "Gl", "Bl", "c0", "c1", "c2"):
        keep[n] = np.ascontiguousarray(getattr(d, n), dtype=np.float64)
    for n in ("gen_ptr", "gen_idx", "sij_ptr", "sij_idx", "sji_ptr", "sji_idx", "sh_ptr", "sh_idx"):
        keep[n] = np.ascontiguousarray(getattr(d, n), dtype=np.int64)
    for n, a in keep.items():
        setattr(net, n, _p64(a) if a.dtype == np.int64 else _pd(a))
    net.cost_order = int(d.cost_order)
    return net, keep


class Handle:
    """ps_handle: one (network, formulation, source_type, flags) model on one device."""

    def __init__(self, d, formulation_code: int, facts_bshunt: bool = True, obj_linear: bool = True, device: int = -1):
        L = lib()
        net, self._keep = make_network(d)
        flags = (1 if facts_bshunt else 0) | (2 if obj_linear else 0)
        src = 0 if d.source_type == "matpower" else 1
        h = _VP()
        rc = L.ps_create(C.byref(net), int(formulation_code), src, flags, int(device), C.byref(h))
        if rc != PS_OK:
            raise PowersenseError(rc, (L.ps_last_error(None) or b"").decode())
        self.h = h
        dims = [C.c_int64() for _ in range(4)]
        self._check(L.ps_dims(self.h, *[C.byref(v) for v in dims]))
        self.n_var, self.m_nl, self.nnzJ, self.nnzH = (v.value for v in dims)

    def _check(self, rc):
        if rc != PS_OK:
            raise PowersenseError(rc, (lib().ps_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            lib().ps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # structure
    def constraint_bounds(self):
        lo, hi = np.zeros(self.m_nl), np.zeros(self.m_nl)
        self._check(lib().ps_constraint_bounds(self.h, _pd(lo), _pd(hi)))
        return lo, hi

    def jacobian_structure(self):
        r, c = np.zeros(self.nnzJ, np.int64), np.zeros(self.nnzJ, np.int64)
        self._check(lib().ps_jacobian_structure(self.h, _p64(r), _p64(c)))
        return r, c

    def hessian_structure(self):
        r, c = np.zeros(self.nnzH, np.int64), np.zeros(self.nnzH, np.int64)
        self._check(lib().ps_hessian_structure(self.h, _p64(r), _p64(c)))
        return r, c

    def hessian_row_blocks(self):
        """0-based offsets [m_nl + 1] of the per-row blocks of the Hessian structure."""
        ptr = np.zeros(self.m_nl + 1, np.int64)
        self._check(lib().ps_hessian_row_blocks(self.h, _p64(ptr)))
        return ptr

    def set_hessian_order(self, perm):
        """perm[i] = 1-based canonical position of the entry reported at position i (None: canonical order)."""
        if perm is None:
            self._check(lib().ps_set_hessian_order(self.h, None))
            return
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        assert perm.size == self.nnzH
        self._check(lib().ps_set_hessian_order(self.h, _p64(perm)))

    def match_hessian_structure(self, row, col):
        """Install the slot order of a target structure (e.g. JuMP's own, dumped by julia/make_golden.jl)."""
        row, col = np.ascontiguousarray(row, dtype=np.int64), np.ascontiguousarray(col, dtype=np.int64)
        assert row.size == self.nnzH and col.size == self.nnzH
        self._check(lib().ps_match_hessian_structure(self.h, _p64(row), _p64(col)))

    def hessian_order(self):
        perm = np.zeros(self.nnzH, np.int64)
        self._check(lib().ps_hessian_order(self.h, _p64(perm)))
        return perm

    def launch_count(self):
        """kernel launches issued by the single-scenario callbacks so far"""
        return int(lib().ps_launch_count(self.h))

    def variable_offsets(self):
        out = np.zeros(16, np.int64)
        self._check(lib().ps_variable_offsets(self.h, _p64(out)))
        names = ("a", "b", "Wd", "Wr", "Wi", "pg", "qg", "F0", "F1", "F2", "F3", "cg", "bss", "tgh", "n_var", "nx_nl")
        return dict(zip(names, out.tolist()))

    def plan_info(self):
        out = np.zeros(8, np.int64)
        self._check(lib().ps_plan_info(self.h, _p64(out)))
        names = ("bus_tiles", "branch_tiles", "smem_bus_bytes", "smem_branch_bytes", "direct_ok", "pb_implicit", "j_br0", "h_br0")
        return dict(zip(names, out.tolist()))

    def preferred_layout(self):
        """((ld_g, ld_J, ld_H), (pad_g, pad_J, pad_H)) in doubles: see ps_preferred_layout."""
        ld, pad = np.zeros(3, np.int64), np.zeros(3, np.int64)
        self._check(lib().ps_preferred_layout(self.h, _p64(ld), _p64(pad)))
        return tuple(ld.tolist()), tuple(pad.tolist())

    def algorithmic_bytes(self, what=7):
        return int(lib().ps_algorithmic_bytes(self.h, what))

    def algorithmic_bytes_branch(self, what=7):
        return int(lib().ps_algorithmic_bytes_branch(self.h, what))

    # single-scenario callbacks (outputs may be views into larger arrays, like the MOI wrapper's)
    def eval_constraint(self, x, g=None):
        x = _f64(x, self.n_var)
        g = np.zeros(self.m_nl) if g is None else g
        assert g.flags.c_contiguous and g.dtype == np.float64 and g.size >= self.m_nl
        self._check(lib().ps_eval_constraint(self.h, _pd(x), _pd(g)))
        return g

    def eval_jacobian(self, x, J=None):
        x = _f64(x, self.n_var)
        J = np.zeros(self.nnzJ) if J is None else J
        assert J.flags.c_contiguous and J.dtype == np.float64 and J.size >= self.nnzJ
        self._check(lib().ps_eval_jacobian(self.h, _pd(x), _pd(J)))
        return J

    def eval_hessian(self, x, sigma, mu, H=None):
        x, mu = _f64(x, self.n_var), _f64(mu, self.m_nl)
        H = np.zeros(self.nnzH) if H is None else H
        assert H.flags.c_contiguous and H.dtype == np.float64 and H.size >= self.nnzH
        self._check(lib().ps_eval_hessian(self.h, _pd(x), float(sigma), _pd(mu), _pd(H)))
        return H

    def eval_all(self, x, mu, g=None, J=None, H=None):
        """g, J, H: optional preallocated outputs (a solver reuses its arrays; fresh ones cost their page faults)"""
        x, mu = _f64(x, self.n_var), _f64(mu, self.m_nl)
        g = np.zeros(self.m_nl) if g is None else g
        J = np.zeros(self.nnzJ) if J is None else J
        H = np.zeros(self.nnzH) if H is None else H
        self._check(lib().ps_eval_all(self.h, _pd(x), _pd(mu), _pd(g), _pd(J), _pd(H)))
        return g, J, H

    def eval_objective(self, x):
        x = _f64(x, self.n_var)
        f = C.c_double()
        self._check(lib().ps_eval_objective(self.h, _pd(x), C.byref(f)))
        return f.value


class Batch:
    """ps_batch: S scenarios of a Handle's model resident on its device."""

    def __init__(self, handle: Handle, S: int):
        self.handle = handle
        self.S = int(S)
        b = _VP()
        handle._check(lib().ps_batch_create(handle.h, self.S, C.byref(b)))
        self.b = b

    def close(self):
        if getattr(self, "b", None) and getattr(self.handle, "h", None):
            lib().ps_batch_destroy(self.b)
        self.b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_loads(self, Pd, Qd):
        nb = self.handle._keep["Pd"].size
        Pd, Qd = _f64(Pd, self.S * nb), _f64(Qd, self.S * nb)
        self.handle._check(lib().ps_batch_set_loads(self.b, _pd(Pd), _pd(Qd)))

    def set_branch_status(self, st):
        if st is None:
            self.handle._check(lib().ps_batch_set_branch_status(self.b, None))
            return
        st = np.ascontiguousarray(st, dtype=np.uint8)
        assert st.size == self.S * self.handle._keep["br_f"].size
        self.handle._check(lib().ps_batch_set_branch_status(self.b, st.ctypes.data_as(_PU8)))

    def set_viol_tol(self, tol):
        self.handle._check(lib().ps_batch_set_viol_tol(self.b, float(tol)))

    def enable_timing(self, on=True):
        self.handle._check(lib().ps_batch_enable_timing(self.b, 1 if on else 0))

    def kernel_times(self):
        """ms of the last evaluation per kernel: {'bus', 'bus_jh', 'objective', 'branch'} (waits for it)."""
        ms = (C.c_float * 4)()
        self.handle._check(lib().ps_batch_kernel_times(self.b, ms))
        return dict(zip(("bus", "bus_jh", "objective", "branch"), [float(v) for v in ms]))

    @property
    def launch_count(self):
        return int(lib().ps_batch_launch_count(self.b))

    def eval_ptrs(self, what, s_begin, s_count, x, ldx, mu, ldmu, g, ldg, J, ldJ, H, ldH, scalars, stream=None):
        """Raw device-pointer call (ints or None)."""
        self.handle._check(lib().ps_batch_eval(self.b, what, s_begin, s_count, x, ldx, mu, ldmu, g, ldg, J, ldJ, H, ldH,
                                               scalars, stream))

    def eval_torch(self, what, x, mu=None, g=None, J=None, H=None, scalars=None, s_begin=0, stream=None):
        """Device tensors (torch, float64, row-major [s_count, ld]); torch is only the allocator here."""
        def p(t):
            return None if t is None else t.data_ptr()

        def ld(t):
            return 0 if t is None else t.stride(0)
        for t in (x, mu, g, J, H, scalars):
            assert t is None or (t.is_cuda and t.dtype.is_floating_point and t.element_size() == 8 and t.stride(-1) == 1)
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(x.device).cuda_stream
        self.eval_ptrs(what, s_begin, x.shape[0], p(x), ld(x), p(mu), ld(mu), p(g), ld(g), p(J), ld(J), p(H), ld(H),
                       p(scalars), stream)

    def eval_host(self, what, x, mu=None, g=None, J=None, H=None, scalars=None, s_begin=0):
        """Host numpy arrays [s_count, ld] (ideally views of pinned memory)."""
        def p(a):
            return None if a is None else a.ctypes.data

        def ld(a):
            return 0 if a is None else a.strides[0] // 8
        for a in (x, mu, g, J, H, scalars):
            assert a is None or (a.dtype == np.float64 and a.strides[-1] == 8)
        self.handle._check(lib().ps_batch_eval_host(self.b, what, s_begin, x.shape[0], p(x), ld(x), p(mu), ld(mu), p(g), ld(g),
                                                    p(J), ld(J), p(H), ld(H), p(scalars)))


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by cudaHostAlloc memory (ps_host_alloc); freed when the array is collected."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = _VP()
    rc = lib().ps_host_alloc(C.byref(p), n)
    if rc != PS_OK:
        raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    import weakref
    weakref.finalize(buf, lib().ps_host_free, p)
    return arr


class host_registered:
    """Context manager: page-lock numpy arrays the caller owns for the duration of the block (ps_host_register), e.g. the
    g / J / H vectors a solver hands to the single-scenario callbacks at every iteration."""

    def __init__(self, *arrays):
        self.arrays = [a for a in arrays if a is not None and a.size]
        self.done = []

    def __enter__(self):
        for a in self.arrays:
            assert a.flags.c_contiguous
            rc = lib().ps_host_register(a.ctypes.data_as(_VP), a.nbytes)
            if rc != PS_OK:
                self.__exit__(None, None, None)
                raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())
            self.done.append(a)
        return self

    def __exit__(self, *exc):
        for a in self.done:
            lib().ps_host_unregister(a.ctypes.data_as(_VP))
        self.done = []
        return False


class CscPattern:
    """ps_csc: fixed COO -> CSC assembly pattern (replaces compute_jacobian_matrix, common.jl:12-20)."""

    def __init__(self, m, n, rows, cols, device=-1):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        p = _VP()
        rc = lib().ps_csc_create(m, n, rows.size, _p64(rows), _p64(cols), device, C.byref(p))
        if rc != PS_OK:
            raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())
        self.p, self.m, self.n, self.nnz_coo = p, m, n, rows.size
        k = C.c_int64()
        lib().ps_csc_dims(self.p, C.byref(k))
        self.nnz = k.value

    def close(self):
        if getattr(self, "p", None):
            lib().ps_csc_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def structure(self):
        colptr, rowval = np.zeros(self.n + 1, np.int64), np.zeros(self.nnz, np.int64)
        rc = lib().ps_csc_structure(self.p, _p64(colptr), _p64(rowval))
        assert rc == PS_OK
        return colptr, rowval

    def assemble(self, dE):
        dE = _f64(dE, self.nnz_coo)
        nz = np.zeros(self.nnz)
        rc = lib().ps_csc_assemble(self.p, _pd(dE), _pd(nz))
        if rc != PS_OK:
            raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())
        return nz

    def assemble_torch(self, dE, nz, stream=None):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(dE.device).cuda_stream
        rc = lib().ps_csc_assemble_batch(self.p, dE.shape[0], dE.data_ptr(), dE.stride(0), nz.data_ptr(), nz.stride(0), stream)
        if rc != PS_OK:
            raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())


class RowsBlock:
    """ps_rows: affine + quadratic rows of the solver-level model on the device (replaces the row evaluation of
    src/opt/MOI_wrapper.jl:864-1042).  Terms in CSR form, 1-based variable indices, MOI term order."""

    def __init__(self, n_var, aff_ptr, aff_var, aff_coef, quad_ptr, quad_var1, quad_var2, quad_coef, constant=None, device=-1):
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._keep = [i64(aff_ptr), i64(aff_var), f64(aff_coef), i64(quad_ptr), i64(quad_var1), i64(quad_var2), f64(quad_coef)]
        ap, av, ac, qp, q1, q2, qc = self._keep
        self.m = len(ap) - 1
        self.n_var = int(n_var)
        cst = f64(np.zeros(self.m) if constant is None else constant)
        p = _VP()
        rc = lib().ps_rows_create(self.n_var, self.m, _p64(ap), _p64(av), _pd(ac), _p64(qp), _p64(q1), _p64(q2), _pd(qc), _pd(cst),
                                  int(device), C.byref(p))
        if rc != PS_OK:
            raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())
        self.p = p
        a, b = C.c_int64(), C.c_int64()
        lib().ps_rows_dims(self.p, C.byref(a), C.byref(b))
        self.nnzJ, self.nnzH = a.value, b.value

    def close(self):
        if getattr(self, "p", None):
            lib().ps_rows_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != PS_OK:
            raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())

    def jacobian_structure(self):
        r, c = np.zeros(self.nnzJ, np.int64), np.zeros(self.nnzJ, np.int64)
        self._check(lib().ps_rows_jacobian_structure(self.p, _p64(r), _p64(c)))
        return r, c

    def hessian_structure(self):
        r, c = np.zeros(self.nnzH, np.int64), np.zeros(self.nnzH, np.int64)
        self._check(lib().ps_rows_hessian_structure(self.p, _p64(r), _p64(c)))
        return r, c

    def eval(self, x, lam=None, what=7):
        x = _f64(x, self.n_var)
        lam = _f64(np.zeros(self.m) if lam is None else lam, self.m)
        g, J, H = np.zeros(self.m), np.zeros(self.nnzJ), np.zeros(self.nnzH)
        self._check(lib().ps_rows_eval(self.p, what, _pd(x), _pd(lam), _pd(g), _pd(J), _pd(H)))
        return g, J, H

    def eval_torch(self, what, x, lam, g, J, H, stream=None):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        p = lambda t: None if t is None else t.data_ptr()
        ld = lambda t: 0 if t is None else t.stride(0)
        self._check(lib().ps_rows_eval_batch(self.p, what, x.shape[0], p(x), ld(x), p(lam), ld(lam), p(g), ld(g), p(J), ld(J),
                                             p(H), ld(H), stream))


PS_NKKT = 8


class KktMetrics:
    """ps_kkt: per-scenario KKT / merit metrics on the device (replaces common.jl:35-98, slp.jl:108-162 at alpha = 0)."""

    def __init__(self, pattern: CscPattern, g_L, g_U, x_L, x_U):
        self.pattern = pattern
        gl, gu = _f64(g_L, pattern.m), _f64(g_U, pattern.m)
        xl, xu = _f64(x_L, pattern.n), _f64(x_U, pattern.n)
        k = _VP()
        rc = lib().ps_kkt_create(pattern.p, _pd(gl), _pd(gu), _pd(xl), _pd(xu), C.byref(k))
        if rc != PS_OK:
            raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())
        self.k = k

    def close(self):
        if getattr(self, "k", None):
            lib().ps_kkt_destroy(self.k)
            self.k = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval_torch(self, E, x, lam, mU, mL, df, nz, nu=None, p=None, f=None, out=None, stream=None):
        import torch
        S = E.shape[0]
        if out is None:
            out = torch.empty(S, PS_NKKT, dtype=torch.float64, device=E.device)
        if stream is None:
            stream = torch.cuda.current_stream(E.device).cuda_stream
        P = lambda t: None if t is None else t.data_ptr()
        ld = lambda t: 0 if t is None else t.stride(0)
        assert mU.stride(0) == mL.stride(0)
        rc = lib().ps_kkt_eval_batch(self.k, S, P(E), ld(E), P(x), ld(x), P(lam), ld(lam), P(mU), P(mL), ld(mU), P(df), ld(df),
                                     P(nz), ld(nz), P(nu), ld(nu), P(p), ld(p), P(f), P(out), stream)
        if rc != PS_OK:
            raise PowersenseError(rc, (lib().ps_last_error(None) or b"").decode())
        return out
