"""Item-sharded scoring of one mega-request across the GPUs of a box.

One process per GPU (torchrun); model and state are replicated.  Ordinary traffic is sharded by
REQUEST with no collective at all (bench.py).  Only a request too large for one GPU's latency
budget (BASELINE config #5: 10 000 items x 2000 trees) is split:

  1. every rank assembles the full request's feature matrix (cheap; per-request aggregates —
     diversity top-N, min-max / position normalisation, interacted_with histograms — need the
     whole item list, S/feature/DiversityFeature.scala:67-130, S/ml/onnx/Normalize.scala:13-46),
  2. rank g scores rows [g*ceil(N/G), (g+1)*ceil(N/G)) with the GBDT kernel,
  3. one all_gather of ceil(N/G) f64 scores per rank (NCCL over NVLink; <= 10 kB per rank, so
     it is latency-bound and a fused compute+collective kernel would buy nothing),
  4. every rank derives the same stable descending order (Ranker.rerank's sortBy(-score)).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous range of rank `rank`: [rank*ceil(N/G), min(N, (rank+1)*ceil(N/G)))."""
    per = -(-n_items // world) if world > 0 else n_items
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def gather_scores(local_scores, n_items: int, group=None):
    """all_gather of equally padded per-rank score slices -> the full score vector on every rank.
    `local_scores` is a 1-D float64 torch tensor (CUDA with NCCL, CPU with gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_scores[:n_items]
    per = -(-n_items // world)
    padded = torch.zeros(per, dtype=torch.float64, device=local_scores.device)
    padded[: local_scores.numel()] = local_scores
    out = torch.empty(per * world, dtype=torch.float64, device=local_scores.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_items]


class ShardedScorer:
    """Scores one request across the process group.

    score_slice(lo, hi) -> 1-D float64 tensor with the scores of items [lo, hi) of the request.
    The product passes a closure over the CUDA path (`cuda_slice_scorer`); CPU tests pass any
    callable, which is how the sharding/gather/order logic is covered with gloo.
    """

    def __init__(self, score_slice, order_fn, group=None):
        self.score_slice = score_slice
        self.order_fn = order_fn
        self.group = group

    def rerank(self, n_items: int):
        import torch.distributed as dist

        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        lo, hi = shard_range(n_items, world, rank)
        local = self.score_slice(lo, hi)
        scores = gather_scores(local, n_items, self.group)
        scores_h = scores.detach().cpu().numpy()
        return scores_h, self.order_fn(scores_h)


def cuda_slice_scorer(ranker, booster, arrays):
    """score_slice for the CUDA path, everything device-resident: assemble the whole request on this
    GPU (mr_rank_device with model = NULL -> dense matrix in HBM), score only rows [lo, hi)."""
    import torch

    from . import features as F

    n, dim = arrays["total_items"], ranker.mapping.dim
    dev = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int64) if v.dtype == np.uint64 else np.ascontiguousarray(v)).cuda()
           for k, v in arrays.items() if isinstance(v, np.ndarray)}
    d_feat = torch.empty(max(n * dim, 1), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    F.rank_device(ranker.state, None, arrays["n_requests"], n, dev["offsets"].data_ptr(), dev["ids"].data_ptr(), 0, 0,
                  d_feat.data_ptr(), stream, dev["users"].data_ptr(), dev["sessions"].data_ptr())

    def score_slice(lo, hi):
        out = torch.empty(max(hi - lo, 0), dtype=torch.float64, device="cuda")
        if hi > lo:
            booster.predict_device(d_feat.data_ptr() + lo * dim * 8, hi - lo, dim, out.data_ptr(), stream)
        return out

    score_slice.keepalive = (dev, d_feat)
    return score_slice, n
