"""ORACLE ctypes front-end (test infrastructure only).

Wraps oracle/liboracle.so (built by `make -C oracle` from gbdt_oracle.c and
assemble_oracle.c).  See the headers of those files for what they restate.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import model_parse

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


class _LgbModel(C.Structure):
    _fields_ = [("n_trees", C.c_int32), ("n_features", C.c_int32),
                ("node_off", C.c_void_p), ("leaf_off", C.c_void_p), ("split_feature", C.c_void_p),
                ("threshold", C.c_void_p), ("decision_type", C.c_void_p), ("left_child", C.c_void_p),
                ("right_child", C.c_void_p), ("leaf_value", C.c_void_p), ("cat_b_off", C.c_void_p),
                ("cat_boundaries", C.c_void_p), ("cat_t_off", C.c_void_p), ("cat_threshold", C.c_void_p)]


class _XgbModel(C.Structure):
    _fields_ = [("n_trees", C.c_int32), ("n_features", C.c_int32), ("base_score", C.c_float),
                ("node_off", C.c_void_p), ("left", C.c_void_p), ("right", C.c_void_p),
                ("split_index", C.c_void_p), ("split_cond", C.c_void_p), ("default_left", C.c_void_p)]


def _cat(arrs, dtype):
    arrs = [np.asarray(a, dtype=dtype) for a in arrs]
    return np.ascontiguousarray(np.concatenate(arrs) if arrs else np.zeros(0, dtype=dtype))


def _offsets(lens):
    return np.ascontiguousarray(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32))


class OracleBooster:
    """CPU restatement of ltrlib `Booster.predictMat` for a parsed model."""

    def __init__(self, kind: int, blob: bytes):
        self.kind = kind
        if kind == 0:
            m = model_parse.parse_lightgbm_text(blob)
            tr = m["trees"]
            k = dict(
                node_off=_offsets([t["num_leaves"] - 1 for t in tr]),
                leaf_off=_offsets([t["num_leaves"] for t in tr]),
                split_feature=_cat([t["split_feature"] for t in tr], np.int32),
                threshold=_cat([t["threshold"] for t in tr], np.float64),
                decision_type=_cat([t["decision_type"] for t in tr], np.int32),
                left_child=_cat([t["left_child"] for t in tr], np.int32),
                right_child=_cat([t["right_child"] for t in tr], np.int32),
                leaf_value=_cat([t["leaf_value"] for t in tr], np.float64),
                cat_b_off=_offsets([len(t["cat_boundaries"]) for t in tr]),
                cat_boundaries=_cat([t["cat_boundaries"] for t in tr], np.int32),
                cat_t_off=_offsets([len(t["cat_threshold"]) for t in tr]),
                cat_threshold=_cat([t["cat_threshold"] for t in tr], np.uint32),
            )
            self._keep = k
            self._m = _LgbModel(len(tr), m["n_features"], *[k[f].ctypes.data for f, _ in _LgbModel._fields_[2:]])
        elif kind == 1:
            m = model_parse.parse_xgboost(blob)
            tr = m["trees"]
            k = dict(
                node_off=_offsets([len(t["left"]) for t in tr]),
                left=_cat([t["left"] for t in tr], np.int32),
                right=_cat([t["right"] for t in tr], np.int32),
                split_index=_cat([t["split_index"] for t in tr], np.int32),
                split_cond=_cat([t["split_cond"] for t in tr], np.float32),
                default_left=_cat([t["default_left"] for t in tr], np.uint8),
            )
            self._keep = k
            self._m = _XgbModel(len(tr), m["n_features"], float(m["base_score"]),
                                *[k[f].ctypes.data for f, _ in _XgbModel._fields_[3:]])
        else:
            raise ValueError(f"unsupported booster tag {kind}")
        self.model = m
        self.n_features = m["n_features"]
        self.n_trees = len(m["trees"])
        self.visited = 0

    def predictMat(self, values, rows: int, cols: int, threads: int = 1) -> np.ndarray:
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        assert values.size == rows * cols
        if cols < self.n_features:
            raise ValueError(f"matrix has {cols} columns, model needs {self.n_features}")
        out = np.empty(rows, dtype=np.float64)
        vis = C.c_int64(0)
        fn = lib().oracle_lgb_predict if self.kind == 0 else lib().oracle_xgb_predict
        fn(C.byref(self._m), C.c_void_p(values.ctypes.data), C.c_int32(rows), C.c_int32(cols),
           C.c_void_p(out.ctypes.data), C.c_int32(threads), C.byref(vis))
        self.visited = vis.value
        return out


def rank_order(scores) -> np.ndarray:
    """Permutation of Ranker.rerank's `sortBy(-_.score)` (stable)."""
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    n = scores.size
    order = np.empty(n, dtype=np.int32)
    scratch = np.empty(max(n, 1), dtype=np.int32)
    lib().oracle_rank_order(C.c_void_p(scores.ctypes.data), C.c_int32(n), C.c_void_p(order.ctypes.data),
                            C.c_void_p(scratch.ctypes.data))
    return order


def gather_rows(cat: np.ndarray, idx: np.ndarray, threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    """out[r] = cat[idx[r]] with the rows spread over `threads` (bench.py's CPU arm)."""
    cat = np.ascontiguousarray(cat, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    rows, cols = idx.size, cat.shape[1]
    if out is None:
        out = np.empty((rows, cols), dtype=np.float64)
    lib().oracle_gather_rows(C.c_void_p(cat.ctypes.data), C.c_void_p(idx.ctypes.data), C.c_int32(rows), C.c_int32(cols),
                             C.c_void_p(out.ctypes.data), C.c_int32(threads))
    return out


def rank_order_batch(scores, offsets, threads: int = 1) -> np.ndarray:
    """rank_order per request of a batch (offsets as in mr_rank_batch), requests in parallel."""
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    order = np.empty(scores.size, dtype=np.int32)
    scratch = np.empty(max(scores.size, 1), dtype=np.int32)
    lib().oracle_rank_order_batch(C.c_void_p(scores.ctypes.data), C.c_void_p(offsets.ctypes.data), C.c_int32(offsets.size - 1),
                                  C.c_void_p(order.ctypes.data), C.c_void_p(scratch.ctypes.data), C.c_int32(threads))
    return order
