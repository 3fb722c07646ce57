"""Host-side mirror of ltrlib's `Booster` as Metarank uses it
(reference S/ml/rank/LambdaMARTRanker.scala:229-230,348,362,365,373,392):

    LightGBMBooster(bytes) / XGBoostBooster(bytes)
    booster.predictMat(values, rows, cols) -> Array[Double]
    booster.save() / weights() / close() / isClosed()

Every call goes through the C ABI (include/mr_b200.h).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, lib


class Context:
    """mr_ctx: one per process per GPU."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().mr_init(C.c_int32(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            check(lib().mr_shutdown(self._h))
            self._h = C.c_void_p()

    @property
    def handle(self):
        return self._h

    def rank_order(self, scores, offsets=None) -> np.ndarray:
        """Ranker.rerank's sortBy(-score) permutation(s) (S/ml/Ranker.scala:52-67)."""
        scores = np.ascontiguousarray(scores, dtype=np.float64)
        if offsets is None:
            offsets = np.array([0, scores.size], dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        order = np.empty(scores.size, dtype=np.int32)
        check(lib().mr_rank_order(self._h, C.c_void_p(scores.ctypes.data), C.c_void_p(offsets.ctypes.data),
                                  C.c_int32(offsets.size - 1), C.c_void_p(order.ctypes.data)))
        return order


class B200Booster:
    """Drop-in for ltrlib `Booster[_]` backed by the sm_100a scoring kernel."""

    KIND = None

    def __init__(self, ctx: Context, blob: bytes, kind: int | None = None, n_features: int = 0):
        kind = self.KIND if kind is None else kind
        self._ctx = ctx
        self._h = C.c_void_p()
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(lib().mr_model_load(ctx.handle, C.c_int32(kind), buf, C.c_size_t(len(blob)),
                                  C.c_int32(n_features), C.byref(self._h)))

    @classmethod
    def from_metarank_blob(cls, ctx: Context, blob: bytes, feature_names: list[str] | None):
        """LambdaMARTPredictor.load (S/ml/rank/LambdaMARTRanker.scala:192-236)."""
        self = cls.__new__(cls)
        self._ctx = ctx
        self._h = C.c_void_p()
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        if feature_names is None:
            names, n = None, -1
        else:
            enc = [s.encode("utf-8") for s in feature_names]
            names = (C.c_char_p * max(len(enc), 1))(*enc)
            n = len(enc)
        check(lib().mr_model_load_metarank(ctx.handle, buf, C.c_size_t(len(blob)), names, C.c_int32(n),
                                           C.byref(self._h)))
        return self

    # -- Booster API -----------------------------------------------------------------
    def predictMat(self, values, rows: int, cols: int) -> np.ndarray:
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        if values.size != rows * cols:
            raise ValueError(f"values has {values.size} cells, rows*cols = {rows * cols}")
        out = np.empty(rows, dtype=np.float64)
        check(lib().mr_model_predict_mat(self._h, C.c_void_p(values.ctypes.data), C.c_int32(rows),
                                         C.c_int32(cols), C.c_void_p(out.ctypes.data)))
        return out

    def predict_device(self, d_values_ptr: int, rows: int, cols: int, d_out_ptr: int, stream: int = 0) -> None:
        """Device-resident predictMat on a caller-provided CUDA stream (no sync)."""
        check(lib().mr_model_predict_mat_device(self._h, C.c_void_p(d_values_ptr), C.c_int32(rows),
                                                C.c_int32(cols), C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def codes_bytes(self, rows: int) -> int:
        lib().mr_model_codes_bytes.restype = C.c_size_t
        return int(lib().mr_model_codes_bytes(self._h, C.c_int32(rows)))

    def bin_device(self, d_values_ptr: int, rows: int, cols: int, d_codes_ptr: int, stream: int = 0) -> None:
        check(lib().mr_model_bin_device(self._h, C.c_void_p(d_values_ptr), C.c_int32(rows), C.c_int32(cols),
                                        C.c_void_p(d_codes_ptr), C.c_void_p(stream)))

    def score_codes_device(self, d_codes_ptr: int, rows: int, d_out_ptr: int, stream: int = 0) -> None:
        check(lib().mr_model_score_codes_device(self._h, C.c_void_p(d_codes_ptr), C.c_int32(rows),
                                                C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def save(self) -> bytes:
        p = C.POINTER(C.c_uint8)()
        n = C.c_size_t()
        check(lib().mr_model_save(self._h, C.byref(p), C.byref(n)))
        return bytes(C.string_at(p, n.value))

    def weights(self) -> np.ndarray:
        n = self.info().n_features
        out = np.zeros(n, dtype=np.float64)
        check(lib().mr_model_weights(self._h, C.c_void_p(out.ctypes.data), C.c_int32(n)))
        return out

    def close(self) -> None:
        if self._h:
            check(lib().mr_model_close(self._h))

    def isClosed(self) -> bool:
        return bool(lib().mr_model_is_closed(self._h))

    # -- extras ----------------------------------------------------------------------
    def info(self) -> _capi.ModelInfo:
        inf = _capi.ModelInfo()
        check(lib().mr_model_get_info(self._h, C.byref(inf)))
        return inf

    def mean_path(self, values, rows: int, cols: int) -> float:
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        out = C.c_double()
        check(lib().mr_model_count_path(self._h, C.c_void_p(values.ctypes.data), C.c_int32(rows), C.c_int32(cols),
                                        C.byref(out)))
        return out.value

    def set_option(self, key: str, value: int) -> None:
        check(lib().mr_model_set_option(self._h, key.encode(), C.c_int32(value)))

    def free(self) -> None:
        if self._h:
            lib().mr_model_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class LightGBMBooster(B200Booster):
    KIND = 0


class XGBoostBooster(B200Booster):
    KIND = 1


def inspect_model(kind: int, blob: bytes, chunk_kb: int = 0) -> _capi.ModelInfo:
    """Host-only parse + pack of a booster blob (no GPU needed); raises MrError on bad input."""
    inf = _capi.ModelInfo()
    buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b"\0")
    check(lib().mr_model_inspect(C.c_int32(kind), buf, C.c_size_t(len(blob)), C.c_int32(chunk_kb), C.byref(inf)))
    return inf


def selfcheck_model(kind: int, blob: bytes, samples: int = 64, max_tile: int = 0) -> tuple[int, int]:
    """Host-only layout check of the throughput scorer's packing (mr_model_selfcheck): (form bits | tile size << 8, leaf
    mismatches)."""
    form, bad = C.c_int32(), C.c_int64()
    buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b"\0")
    check(lib().mr_model_selfcheck(C.c_int32(kind), buf, C.c_size_t(len(blob)), C.c_int32(samples), C.c_int32(max_tile), C.byref(form), C.byref(bad)))
    return form.value, bad.value
