"""ORACLE (test infrastructure only — never imported by the product path).

Independent CPU parsers for the two booster blobs Metarank stores in a model
(reference S/ml/rank/LambdaMARTRanker.scala:192-236, `boosterType` 0 = LightGBM
text, 1 = XGBoost bytes).  The prediction arithmetic itself is NOT in the
reference tree: it lives in io.github.metarank:ltrlib:0.2.6 ->
lightgbm4j:4.6.0-1 (LightGBM 4.6.0) / metarank's xgboost4j fork (build.sbt:57-58),
none of which exists in this container.  The parsers below restate the public
model formats of those libraries (LightGBM `gbdt_model_text.cpp` / `tree.cpp`
key=value blocks; XGBoost JSON/UBJSON model schema).  PARITY UNPINNED: the
reference's tests hold no golden score for this boundary.
"""
from __future__ import annotations

import io
import json
import struct

import numpy as np


# --------------------------------------------------------------------------- metarank blob

def parse_metarank_blob(blob: bytes):
    """Framing of LambdaMARTPredictor.load (S/ml/rank/LambdaMARTRanker.scala:192-236)."""
    s = io.BytesIO(blob)
    version = struct.unpack(">b", s.read(1))[0]
    if version not in (2, 3):
        raise ValueError(f"unsupported bitstream version {version}")
    n = struct.unpack(">i", s.read(4))[0]
    names = []
    for _ in range(n):
        ln = struct.unpack(">H", s.read(2))[0]
        names.append(s.read(ln).decode("utf-8"))
    kind = struct.unpack(">b", s.read(1))[0]
    size = struct.unpack(">i", s.read(4))[0]
    booster = s.read(size)
    if len(booster) != size:
        raise ValueError("truncated booster blob")
    if kind not in (0, 1):
        raise ValueError(f"unsupported booster tag {kind}")
    return version, names, kind, booster


# --------------------------------------------------------------------------- LightGBM text

def parse_lightgbm_text(blob: bytes) -> dict:
    """LightGBM model text -> flat arrays.

    Returns dict(kind='lightgbm', n_features, trees=[dict(num_leaves, split_feature,
    threshold, decision_type, left_child, right_child, leaf_value, cat_boundaries,
    cat_threshold)]).  Thresholds / leaf values are parsed with Python float()
    (correctly rounded, like LightGBM's fast_double_parser path).
    """
    text = blob.decode("utf-8", errors="replace")
    lines = text.split("\n")
    header = {}
    i = 0
    while i < len(lines) and not lines[i].startswith("Tree="):
        if "=" in lines[i]:
            k, v = lines[i].split("=", 1)
            header[k.strip()] = v.strip()
        i += 1
    if "max_feature_idx" not in header:
        raise ValueError("not a LightGBM model: max_feature_idx missing")
    if int(header.get("num_class", "1")) != 1 or int(header.get("num_tree_per_iteration", "1")) != 1:
        raise ValueError("multiclass models are not supported")
    trees = []
    cur = None
    for ln in lines[i:]:
        ln = ln.strip()
        if ln.startswith("Tree="):
            cur = {}
            trees.append(cur)
        elif ln.startswith("end of trees"):
            break
        elif "=" in ln and cur is not None:
            k, v = ln.split("=", 1)
            cur[k] = v
    out = []
    for t in trees:
        nl = int(t["num_leaves"])
        if int(t.get("is_linear", "0")) != 0:
            raise ValueError("linear trees are not supported")

        def arr(key, conv, n):
            if n == 0:
                return []
            vals = t[key].split()
            if len(vals) != n:
                raise ValueError(f"{key}: expected {n} values, got {len(vals)}")
            return [conv(x) for x in vals]

        d = dict(num_leaves=nl)
        d["leaf_value"] = np.array(arr("leaf_value", float, nl), dtype=np.float64)
        ni = nl - 1
        d["split_feature"] = np.array(arr("split_feature", int, ni), dtype=np.int32)
        d["threshold"] = np.array(arr("threshold", float, ni), dtype=np.float64)
        d["decision_type"] = np.array(arr("decision_type", int, ni), dtype=np.int32)
        d["left_child"] = np.array(arr("left_child", int, ni), dtype=np.int32)
        d["right_child"] = np.array(arr("right_child", int, ni), dtype=np.int32)
        ncat = int(t.get("num_cat", "0"))
        if ncat > 0:
            cb = [int(x) for x in t["cat_boundaries"].split()]
            ct = [int(x) for x in t["cat_threshold"].split()]
            if len(cb) != ncat + 1:
                raise ValueError("cat_boundaries size mismatch")
        else:
            cb, ct = [0], []
        d["cat_boundaries"] = np.array(cb, dtype=np.int32)
        d["cat_threshold"] = np.array(ct, dtype=np.uint32)
        out.append(d)
    return dict(kind="lightgbm", n_features=int(header["max_feature_idx"]) + 1, trees=out,
                objective=header.get("objective", ""))


# --------------------------------------------------------------------------- XGBoost JSON / UBJSON

class _F32(float):
    """marks a UBJSON float32 so that no double-rounding happens on re-parse"""


def _ubj_read(s: io.BytesIO):
    def rd(fmt, n):
        b = s.read(n)
        if len(b) != n:
            raise ValueError("truncated UBJSON")
        return struct.unpack(fmt, b)[0]

    def scalar(tag):
        if tag == b"Z": return None
        if tag == b"T": return True
        if tag == b"F": return False
        if tag == b"i": return rd(">b", 1)
        if tag == b"U": return rd(">B", 1)
        if tag == b"I": return rd(">h", 2)
        if tag == b"l": return rd(">i", 4)
        if tag == b"L": return rd(">q", 8)
        if tag == b"d": return float(np.float32(rd(">f", 4)))
        if tag == b"D": return rd(">d", 8)
        if tag == b"C": return s.read(1).decode("latin1")
        if tag == b"S":
            n = scalar(s.read(1))
            return s.read(n).decode("utf-8")
        raise ValueError(f"bad UBJSON tag {tag!r}")

    def value(tag=None):
        tag = tag or s.read(1)
        if tag == b"{":
            out = {}
            while True:
                t = s.read(1)
                if t == b"}":
                    return out
                n = scalar(t)
                k = s.read(n).decode("utf-8")
                out[k] = value()
        if tag == b"[":
            t = s.read(1)
            if t == b"$":
                ety = s.read(1)
                if s.read(1) != b"#":
                    raise ValueError("typed array without count")
                n = scalar(s.read(1))
                return [scalar(ety) for _ in range(n)]
            if t == b"#":
                n = scalar(s.read(1))
                return [value() for _ in range(n)]
            out = []
            while t != b"]":
                out.append(value(t))
                t = s.read(1)
            return out
        return scalar(tag)

    return value()


def parse_xgboost(blob: bytes) -> dict:
    """XGBoost model bytes (JSON text or UBJSON) -> flat arrays per tree.

    float fields are rounded to f32 exactly once (np.float32 of the decimal token /
    the stored f32), matching XGBoost's own `float` model storage.
    """
    head = blob.lstrip()[:1]
    if head == b"{" and blob.lstrip()[1:2] in (b'"', b"}", b" ", b"\n"):
        doc = json.loads(blob.decode("utf-8"), parse_float=lambda tok: float(np.float32(tok)))
    elif head == b"{":
        doc = _ubj_read(io.BytesIO(blob))
    elif blob[:4] == b"bs64":
        raise ValueError("base64-wrapped XGBoost binary model is not supported")
    else:
        return _parse_xgboost_binary(blob)
    learner = doc["learner"]
    gb = learner["gradient_booster"]
    if gb["name"] != "gbtree":
        raise ValueError(f"booster {gb['name']} not supported")
    lmp = learner["learner_model_param"]
    bs = lmp["base_score"]
    if isinstance(bs, str):
        bs = bs.strip("[]")
    base_score = np.float32(bs)
    n_features = int(lmp["num_feature"])
    trees = []
    for t in gb["model"]["trees"]:
        if any(int(x) != 0 for x in t.get("split_type", [])):
            raise ValueError("categorical XGBoost splits are not supported")
        trees.append(dict(
            left=np.array(t["left_children"], dtype=np.int32),
            right=np.array(t["right_children"], dtype=np.int32),
            split_index=np.array(t["split_indices"], dtype=np.int32),
            split_cond=np.array(t["split_conditions"], dtype=np.float32),
            default_left=np.array(t["default_left"], dtype=np.uint8),
        ))
    return dict(kind="xgboost", n_features=n_features, base_score=base_score, trees=trees,
                objective=learner.get("objective", {}).get("name", ""))


def _parse_xgboost_binary(blob: bytes) -> dict:
    """XGBoost's deprecated binary encoding (the default of Booster.toByteArray() / save_raw() up to 2.0; restated from
    xgboost's public headers: src/learner.cc LearnerModelParamLegacy (136 B), src/gbm/gbtree_model.h GBTreeModelParam
    (160 B), include/xgboost/tree_model.h TreeParam (148 B), RegTree::Node (20 B: parent, cleft, cright, sindex with
    the default-left flag in bit 31, leaf value | split condition) and RTreeNodeStat (16 B)).  Nodes keep their file
    order; pruned nodes stay in the arrays, unreachable."""
    import struct

    p = 4 if blob[:4] == b"binf" else 0
    if len(blob) < p + 136:
        raise ValueError("unsupported XGBoost model encoding")
    base_score, n_features, num_class = struct.unpack_from("<fIi", blob, p)
    major, _, num_target = struct.unpack_from("<III", blob, p + 20)
    p += 136
    if not (base_score == base_score) or n_features == 0 or n_features > 1 << 24 or major > 10:
        raise ValueError("unsupported XGBoost model encoding")
    if num_class > 1 or num_target > 1:
        raise ValueError("multi-output XGBoost models are not supported")
    names = []
    for _ in range(2):
        (n,) = struct.unpack_from("<Q", blob, p)
        if n > 256:
            raise ValueError("XGBoost binary model: name too long")
        names.append(blob[p + 8:p + 8 + n].decode())
        p += 8 + n
    if names[1] != "gbtree":
        raise ValueError(f"booster {names[1]} not supported")
    (n_trees,) = struct.unpack_from("<i", blob, p)
    p += 160
    trees = []
    for _ in range(n_trees):
        _, n, _, _, _, leaf_vec = struct.unpack_from("<iiiiIi", blob, p)
        p += 148
        if n < 1 or p + n * 36 > len(blob):
            raise ValueError("XGBoost binary model: truncated tree")
        nodes = np.frombuffer(blob, dtype=np.dtype([("parent", "<i4"), ("l", "<i4"), ("r", "<i4"), ("s", "<u4"), ("v", "<f4")]),
                              count=n, offset=p)
        p += n * 36
        if leaf_vec != 0:
            (lv,) = struct.unpack_from("<Q", blob, p)
            p += 8 + 4 * lv
        leaf = nodes["l"] == -1
        trees.append(dict(
            left=nodes["l"].astype(np.int32), right=np.where(leaf, -1, nodes["r"]).astype(np.int32),
            split_index=np.where(leaf, 0, nodes["s"] & 0x7FFFFFFF).astype(np.int32),
            split_cond=nodes["v"].astype(np.float32),
            default_left=np.where(leaf, 0, nodes["s"] >> 31).astype(np.uint8),
        ))
    if p + 4 * n_trees > len(blob):
        raise ValueError("XGBoost binary model: truncated tree_info")
    return dict(kind="xgboost", n_features=int(n_features), base_score=np.float32(base_score), trees=trees, objective=names[0])


# --------------------------------------------------------------------------- second opinion

def predict_python(model: dict, X) -> np.ndarray:
    """Straight-line pure-Python evaluation of a parsed model, written independently of
    oracle/gbdt_oracle.c (different language, different data structures) so that the two can be
    cross-checked on random models: the GBDT boundary has no golden vector in the reference."""
    X = np.asarray(X, dtype=np.float64)
    out = np.empty(X.shape[0], dtype=np.float64)
    if model["kind"] == "lightgbm":
        k_zero = float(np.float32(1e-35))
        for r in range(X.shape[0]):
            s = 0.0
            for t in model["trees"]:
                if t["num_leaves"] <= 1:
                    s += float(t["leaf_value"][0])
                    continue
                node = 0
                while node >= 0:
                    x = float(X[r, t["split_feature"][node]])
                    dt = int(t["decision_type"][node])
                    if dt & 1:  # categorical
                        left = False
                        if x == x and -2147483649.0 < x < 2147483648.0 and int(x) >= 0:
                            iv = int(x)
                            ci = int(t["threshold"][node])
                            b, e = int(t["cat_boundaries"][ci]), int(t["cat_boundaries"][ci + 1])
                            if iv // 32 < e - b:
                                left = bool((int(t["cat_threshold"][b + iv // 32]) >> (iv % 32)) & 1)
                    else:
                        mt = (dt >> 2) & 3
                        if x != x and mt != 2:
                            x = 0.0
                        if (mt == 1 and -k_zero <= x <= k_zero) or (mt == 2 and x != x):
                            left = bool(dt & 2)
                        else:
                            left = x <= float(t["threshold"][node])
                    node = int(t["left_child"][node] if left else t["right_child"][node])
                s += float(t["leaf_value"][~node])
            out[r] = s
        return out
    for r in range(X.shape[0]):
        s = np.float32(model["base_score"])
        for t in model["trees"]:
            nid = 0
            while t["left"][nid] != -1:
                fv = np.float32(X[r, t["split_index"][nid]])
                if fv != fv:
                    nid = int(t["left"][nid] if t["default_left"][nid] else t["right"][nid])
                else:
                    nid = int(t["left"][nid] if fv < t["split_cond"][nid] else t["right"][nid])
            s = np.float32(s + t["split_cond"][nid])
        out[r] = float(s)
    return out
