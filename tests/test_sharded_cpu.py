"""Host side of the mega-request sharding, no GPU: the slicing rule (Python statement vs the
library's mr_group_slice), and — with world_size-2 gloo processes — the handle exchange that connects the
members, followed by the gather-by-slices + ordering logic on a stand-in scorer.  The CUDA path itself
(peer stores from the scoring kernel, device-side wait, ordering) is covered by tests/test_group_gpu.py."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metarank_b200 import _capi, sharded
from oracle import oracle


@pytest.mark.parametrize("n,world", [(10, 2), (10_000, 8), (7, 8), (0, 4), (1, 1), (9, 4), (1280, 8), (1281, 8),
                                     (4096, 3), (100_000, 8)])
def test_shard_ranges_partition_the_items(n, world):
    rs = [sharded.shard_range(n, world, r) for r in range(world)]
    assert rs[0][0] == 0 and rs[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    per = -(-(-(-n // world)) // 128) * 128 if n else 0
    assert all(hi - lo <= per for lo, hi in rs)
    assert all(lo % 128 == 0 for lo, hi in rs if hi > lo)  # slices start on scorer tiles
    # the library's rule is the same one
    lib = _capi.lib()
    for r in range(world):
        lo, hi = C.c_int32(-1), C.c_int32(-1)
        lib.mr_group_slice(C.c_int32(n), C.c_int32(world), C.c_int32(r), C.byref(lo), C.byref(hi))
        assert (lo.value, hi.value) == rs[r]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. the members' 64-byte handles reach every rank in rank order
    mine = bytes([rank + 1]) * sharded.HANDLE_BYTES
    table = sharded.exchange_handles(mine)
    handles_ok = table == b"".join(bytes([r + 1]) * sharded.HANDLE_BYTES for r in range(world))
    # 2. every rank scores its slice only; the slices tile the request; every rank derives the same order
    rng = np.random.Generator(np.random.PCG64(123))
    truth = rng.standard_normal(n_items)
    truth[::17] = truth[min(3, n_items - 1)]  # ties -> stability matters
    lo, hi = sharded.shard_range(n_items, world, rank)
    full = torch.zeros(n_items, dtype=torch.float64)
    full[lo:hi] = torch.from_numpy(truth[lo:hi].copy())  # what the peer stores do on the GPU
    dist.all_reduce(full)                                   # slices are disjoint: the sum is the gather
    scores = full.numpy()
    ok = handles_ok and np.array_equal(scores, truth)
    q.put((rank, bool(ok), (lo, hi), oracle.rank_order(scores).tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [1001, 3, 300])
def test_handle_exchange_and_slices_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(30) for p in ps]
    assert all(r[1] for r in res), res
    assert [r[2] for r in res] == [sharded.shard_range(n_items, 2, 0), sharded.shard_range(n_items, 2, 1)]
    assert res[0][3] == res[1][3]
