"""World-size-2 (and 3) CPU tests of the multi-GPU host logic over gloo: scenario sharding, the padded all_gather of
the per-scenario scalars (the only collective of the path) and the contingency-status generator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from powersense_b200 import driver


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, n = driver.shard_range(S_total, world, rank)
        # every scenario's "scalars" encode its global index, so the gathered table is checkable on every rank
        local = torch.stack([torch.arange(b, b + n, dtype=torch.float64) * (k + 1) for k in range(4)], dim=1)
        full = driver.all_gather_rows(local, S_total, world)
        want = torch.stack([torch.arange(S_total, dtype=torch.float64) * (k + 1) for k in range(4)], dim=1)
        ok = full.shape == want.shape and bool(torch.equal(full, want))
        q.put((rank, ok, b, n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,S_total", [(2, 10), (2, 7), (3, 8)])
def test_all_gather_rows_over_gloo(world, S_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, S_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    cover = sorted((b, n) for _, _, b, n in res)
    assert cover[0][0] == 0 and sum(n for _, n in cover) == S_total
    for (b0, n0), (b1, _) in zip(cover, cover[1:]):
        assert b0 + n0 == b1


def test_shard_range_properties():
    for S in (0, 1, 5, 4096, 4099):
        for w in (1, 2, 3, 8):
            parts = [driver.shard_range(S, w, r) for r in range(w)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == S
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
            for (b0, n0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + n0 == b1
    with pytest.raises(ValueError):
        driver.shard_range(4, 2, 2)


def test_gather_unpad_numpy():
    S, w = 7, 3
    smax = 3
    padded = np.zeros((w * smax, 2))
    k = 0
    for r in range(w):
        _, n = driver.shard_range(S, w, r)
        padded[r * smax:r * smax + n, 0] = np.arange(k, k + n)
        k += n
    out = driver.gather_unpad(padded, S, w)
    assert out.shape == (S, 2) and np.array_equal(out[:, 0], np.arange(S))


def test_contingency_status():
    st = driver.contingency_status(5, s_begin=3, s_count=6)
    assert st.shape == (6, 5) and st.dtype == np.uint8
    assert [int(np.argmin(r)) for r in st] == [3, 4, 0, 1, 2, 3] and np.all(st.sum(axis=1) == 4)
