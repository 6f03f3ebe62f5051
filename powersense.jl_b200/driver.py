"""Batched scenario driver: S scenarios of one network, sharded across the ranks of one box.

The reference has no scenario batching (one PowersenseData = one network + one load vector, one solve at a
time: src/opf/formulations.jl:584-653).  Scenarios are independent units, so each rank (one process per GPU)
owns a contiguous block and evaluates it with ps_batch_eval; there is no data-path collective.  The only
exchange is the gather of the per-scenario scalars (objective, max NL violation, violated-row count) over
NCCL (NVLink 5 / NVSwitch) -- `gather_scalars`.  torch is used for device memory, streams and
torch.distributed only."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import capi
from .types import ACOPF_fromulation, EVAL_G, EVAL_H, EVAL_J


def shard_range(S_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, begin + count) of scenarios owned by `rank` (blocks differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(S_total), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def contingency_status(nbr: int, s_begin: int, s_count: int) -> np.ndarray:
    """N-1 branch sweep (BASELINE config 5): global scenario s opens branch s mod nbr."""
    st = np.ones((s_count, nbr), dtype=np.uint8)
    k = (np.arange(s_begin, s_begin + s_count) % nbr)
    st[np.arange(s_count), k] = 0
    return st


class ScenarioDriver:
    """Owns one rank's shard: the model handle, the batch and the device buffers (row-major [S_local, ld])."""

    def __init__(self, data, formulation: ACOPF_fromulation, S_total: int, rank: int = 0, world_size: int = 1,
                 device_index: Optional[int] = None, facts_bshunt: bool = True, obj_linear: bool = True,
                 what: int = EVAL_G | EVAL_J | EVAL_H):
        import torch
        self.torch = torch
        self.rank, self.world_size = rank, world_size
        self.S_total = int(S_total)
        self.s_begin, self.S = shard_range(S_total, world_size, rank)
        self.device_index = rank if device_index is None else device_index
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        self.handle = capi.Handle(data, formulation.code, facts_bshunt, obj_linear, device=self.device_index)
        self.batch = capi.Batch(self.handle, max(self.S, 1))
        self.what = what
        h = self.handle
        f64 = torch.float64
        n = max(self.S, 1)
        self.x = torch.zeros(n, h.n_var, dtype=f64, device=self.device)
        self.mu = torch.zeros(n, h.m_nl, dtype=f64, device=self.device) if what & EVAL_H else None
        # outputs in the layout ps_preferred_layout asks for: rows of ld doubles starting `pad` doubles into a
        # 32-byte aligned allocation, so that the branch block of every row starts on a 32-byte sector
        (ldg, ldJ, ldH), (pg, pJ, pH) = h.preferred_layout()

        def rows(width, ld, pad):
            flat = torch.empty(pad + n * ld, dtype=f64, device=self.device)
            return flat[pad:].view(n, ld)[:, :width]
        self.g = rows(h.m_nl, ldg, pg) if what & EVAL_G else None
        self.J = rows(h.nnzJ, ldJ, pJ) if what & EVAL_J else None
        self.H = rows(h.nnzH, ldH, pH) if what & EVAL_H else None
        # two sets of per-scenario scalars: the gather of evaluation k (NCCL, its own stream) overlaps evaluation k + 1
        self._scal = [torch.zeros(n, capi.PS_NSCALARS, dtype=f64, device=self.device) for _ in range(2)]
        self._flip = 0
        self.scalars = self._scal[0]
        self._comm_stream = None
        self._gather_ev = [None, None]        # gather of buffer q finished (the next evaluation into q waits for it)
        self._gather_out = None

    def set_loads(self, Pd: np.ndarray, Qd: np.ndarray):
        self.batch.set_loads(Pd, Qd)

    def set_branch_status(self, st):
        self.batch.set_branch_status(st)

    def evaluate(self, what: Optional[int] = None):
        """One pass of the hot path over this rank's shard (asynchronous on torch's current stream)."""
        if self.S == 0:
            return
        ev = self._gather_ev[self._flip]
        if ev is not None:                                # the gather that last read this scalar buffer
            self.torch.cuda.current_stream().wait_event(ev)
        self.scalars = self._scal[self._flip]
        self.batch.eval_torch(self.what if what is None else what, self.x, self.mu, self.g, self.J, self.H, self.scalars)

    def gather_scalars(self):
        """[S_total, PS_NSCALARS] on every rank: NCCL all_gather over NVLink when world_size > 1."""
        if self.world_size == 1:
            return self.scalars[:self.S]
        return all_gather_rows(self.scalars[:self.S], self.S_total, self.world_size)

    def gather_scalars_async(self):
        """The same exchange, off the evaluation's critical path: issued on a communication stream behind an event, into a
        preallocated buffer, so that it overlaps the next evaluation (which writes the other scalar buffer).  Call
        wait_gathers() before reading gathered_scalars.  Equal shards take one all_gather_into_tensor and nothing else."""
        torch = self.torch
        if self.world_size == 1:
            self.gathered_scalars = self.scalars[:self.S]
            return
        import torch.distributed as dist
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
            smax = -(-self.S_total // self.world_size)
            self._gather_out = torch.empty(self.world_size * smax, capi.PS_NSCALARS, dtype=torch.float64, device=self.device)
            self._gather_pad = torch.zeros(smax, capi.PS_NSCALARS, dtype=torch.float64, device=self.device)
        done = torch.cuda.Event()
        done.record()                                      # evaluation k has written its scalars
        src = self.scalars
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(done)
            if self.S * self.world_size == self.S_total:
                dist.all_gather_into_tensor(self._gather_out, src[:self.S])
            else:
                self._gather_pad[:self.S] = src[:self.S]
                dist.all_gather_into_tensor(self._gather_out, self._gather_pad)
            ev = torch.cuda.Event()
            ev.record()
        self._gather_ev[self._flip] = ev
        self._flip ^= 1
        self.gathered_scalars = None

    def wait_gathers(self):
        """Make torch's current stream wait for every outstanding gather; returns [S_total, PS_NSCALARS]."""
        if self.world_size == 1:
            return self.scalars[:self.S]
        for ev in self._gather_ev:
            if ev is not None:
                self.torch.cuda.current_stream().wait_event(ev)
        even = self.S * self.world_size == self.S_total
        self.gathered_scalars = self._gather_out if even else gather_unpad(self._gather_out, self.S_total, self.world_size)
        return self.gathered_scalars

    def close(self):
        self.batch.close()
        self.handle.close()


def all_gather_rows(local_rows, S_total: int, world_size: int):
    """all_gather of per-rank row blocks whose sizes follow shard_range (they differ by at most one row, so the
    collective is padded to the largest shard).  Works on any torch.distributed backend (NCCL on the GPUs, gloo in
    the CPU tests); the only collective of the whole path."""
    import torch
    import torch.distributed as dist
    smax = -(-S_total // world_size)
    k = local_rows.shape[1]
    pad = torch.zeros(smax, k, dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    out = torch.empty(world_size * smax, k, dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, pad)
    return gather_unpad(out, S_total, world_size)


def gather_unpad(padded, S_total: int, world_size: int):
    """Drop the per-rank padding of an all_gather of ceil(S/world) rows per rank (works on torch or numpy)."""
    smax = -(-S_total // world_size)
    parts = []
    for r in range(world_size):
        _, cnt = shard_range(S_total, world_size, r)
        parts.append(padded[r * smax:r * smax + cnt])
    if hasattr(padded, "new_empty"):
        import torch
        return torch.cat(parts, dim=0)
    return np.concatenate(parts, axis=0)
