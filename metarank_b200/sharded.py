"""Item-sharded ranking of one mega-request across the GPUs of a box — the ctypes mirror
of the mr_group_* entry points of include/mr_b200.h.

Ordinary traffic is sharded by REQUEST with no exchange at all (bench.py).  Only a request too large for one
GPU's latency budget (BASELINE config #5: 10 000 items x 2000 trees) is split by item: every member of the
group assembles and scores a contiguous item range, the scoring kernel's final store writes each score into
the exchange buffer of EVERY member over NVLink (peer memory: CUDA IPC between processes, direct peer access
inside one process) and raises the member's flag there; each member then waits for the other flags on the
device and orders the full score vector.  There is no host round trip and no collective call on the data
path — `torch.distributed` (or any other transport) is only used ONCE, to exchange the 64-byte IPC handles.

`shard_range` is the host-side statement of the slicing rule (the library's `mr_group_slice`); the gloo tests
hold the two to each other and cover the handle exchange with world_size 2 on CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

HANDLE_BYTES = 64
TILE = 128  # slices are whole scorer tiles


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous range of member `rank`: ceil(N / G) rounded up to whole 128-item tiles."""
    if world <= 0 or n_items <= 0:
        return 0, 0
    per = -(-(-(-n_items // world)) // TILE) * TILE
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def exchange_handles(local: bytes, group=None) -> bytes:
    """All-gather of the members' IPC handles over torch.distributed (NCCL needs CUDA tensors, gloo takes CPU
    ones): returns world x 64 bytes in rank order."""
    import torch
    import torch.distributed as dist

    assert len(local) == HANDLE_BYTES
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = "cuda" if backend == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(dev)
    out = torch.empty(world * HANDLE_BYTES, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return bytes(out.cpu().numpy().tobytes())


class Group:
    """One member (rank of world) of an mr_group on `ctx`'s device."""

    def __init__(self, ctx, rank: int, world: int, max_items: int):
        from ._capi import check, lib

        self.ctx, self.rank, self.world, self.max_items = ctx, rank, world, max_items
        self._h = C.c_void_p()
        check(lib().mr_group_create(ctx.handle, C.c_int32(rank), C.c_int32(world), C.c_int32(max_items), C.byref(self._h)))

    def export(self) -> bytes:
        from ._capi import check, lib

        buf = (C.c_uint8 * HANDLE_BYTES)()
        check(lib().mr_group_export(self._h, buf))
        return bytes(buf)

    def connect(self, handles: bytes) -> None:
        """Multi-process: `handles` = exchange_handles(self.export())."""
        from ._capi import check, lib

        assert len(handles) == self.world * HANDLE_BYTES
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        check(lib().mr_group_connect(self._h, buf))

    def connect_distributed(self, group=None) -> None:
        self.connect(exchange_handles(self.export(), group) if self.world > 1 else self.export())

    @staticmethod
    def connect_local(members: list["Group"]) -> None:
        """Single process driving every member (one per GPU, or several on one GPU in tests)."""
        from ._capi import check, lib

        arr = (C.c_void_p * len(members))(*[m._h for m in members])
        check(lib().mr_group_connect_local(arr, C.c_int32(len(members))))

    def slice(self, n_items: int) -> tuple[int, int]:
        from ._capi import lib

        lo, hi = C.c_int32(), C.c_int32()
        lib().mr_group_slice(C.c_int32(n_items), C.c_int32(self.world), C.c_int32(self.rank), C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def rank_arrays(self, state, model, arrays: dict, want_order: bool = True):
        """mr_group_rank (collective: every member calls it with the same single-request arrays)."""
        from ._capi import check, lib
        from .features import RankBatch

        N = arrays["total_items"]
        b = RankBatch(arrays["n_requests"], arrays["offsets"].ctypes.data, arrays["ids"].ctypes.data,
                      arrays["users"].ctypes.data, arrays["sessions"].ctypes.data, arrays["req_f64"].ctypes.data,
                      arrays["req_u64"].ctypes.data, arrays["req_vec"].ctypes.data, arrays["req_vp"].ctypes.data,
                      arrays["item_f64"].ctypes.data if arrays.get("item_f64") is not None else None,
                      arrays["tok_off"].ctypes.data if "tok_off" in arrays else None,
                      arrays["tok_hash"].ctypes.data if "tok_hash" in arrays else None,
                      arrays["tok_w"].ctypes.data if "tok_w" in arrays else None)
        scores = np.empty(max(N, 1), dtype=np.float64)
        order = np.empty(max(N, 1), dtype=np.int32) if want_order else None
        check(lib().mr_group_rank(self._h, state._h, model._h, C.byref(b), C.c_void_p(scores.ctypes.data),
                                  C.c_void_p(order.ctypes.data) if order is not None else None))
        return scores[:N], (order[:N] if order is not None else None)

    def rank_device(self, state, model, n_items: int, d_offsets: int, d_item_ids: int, d_scores: int, d_order: int,
                    stream: int = 0) -> None:
        """mr_group_rank_device: device pointers, enqueued on `stream` without synchronising."""
        from ._capi import check, lib
        from .features import RankBatch

        b = RankBatch(1, d_offsets, d_item_ids, None, None, None, None, None, None, None)
        check(lib().mr_group_rank_device(self._h, state._h, model._h, C.byref(b), C.c_int32(n_items),
                                         C.c_void_p(d_scores or None), C.c_void_p(d_order or None), C.c_void_p(stream)))

    def free(self) -> None:
        from ._capi import lib

        if self._h:
            lib().mr_group_free(self._h)
            self._h = C.c_void_p()
