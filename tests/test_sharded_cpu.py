"""world_size-2 gloo test of the mega-request sharding logic (host side only; the per-slice
scorer is a stand-in callable so no GPU is needed — the CUDA scorer itself is covered by the
-m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metarank_b200 import sharded
from oracle import oracle


@pytest.mark.parametrize("n,world", [(10, 2), (10_000, 8), (7, 8), (0, 4), (1, 1), (9, 4)])
def test_shard_ranges_partition_the_items(n, world):
    rs = [sharded.shard_range(n, world, r) for r in range(world)]
    assert rs[0][0] == 0 and rs[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    per = -(-n // world) if n else 0
    assert all(hi - lo <= per for lo, hi in rs)


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
    rng = np.random.Generator(np.random.PCG64(123))
    truth = rng.standard_normal(n_items)
    truth[::17] = truth[min(3, n_items - 1)]  # ties -> stability matters
    calls = []

    def score_slice(lo, hi):
        calls.append((lo, hi))
        return torch.from_numpy(truth[lo:hi].copy())

    sc = sharded.ShardedScorer(score_slice, oracle.rank_order)
    scores, order = sc.rerank(n_items)
    ok = np.array_equal(scores, truth) and np.array_equal(order, oracle.rank_order(truth))
    q.put((rank, bool(ok), calls))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [1001, 3])
def test_gather_and_order_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(30) for p in ps]
    assert all(r[1] for r in res), res
    # every rank scored exactly its own contiguous slice, once
    assert [r[2] for r in res] == [[sharded.shard_range(n_items, 2, 0)], [sharded.shard_range(n_items, 2, 1)]]
