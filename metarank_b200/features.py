"""Host-side mirror of the reference's feature plumbing on the /rank path, over the C ABI:

    FeatureMapping   S/FeatureMapping.scala:56-99       -> mr_schema (columns, inputs)
    DeviceState      Persistence.values (KVStore.put)   -> mr_state_upsert / mr_state_flush
    Ranker.rerank    S/ml/Ranker.scala:27-83            -> mr_rank

This is what the JVM shim does in Scala; here it is Python so the
parity tests can drive the CUDA path with the same literal events/requests as the oracle.
Plain-data conventions (shared with the tests, no shared code with oracle/):
  Key   = (scope, feature_name); scope = ("global",) | ("item", id) | ("user", id) | ("session", id) |
          ("field", name, value) | ("irf", name, value, item) | ("ranking", id)
  Value = ("scalar", float|str|list[str]|list[float]) | ("counter", int) | ("pcounter", [int]) | ("blist", [ids])
"""
from __future__ import annotations

import ctypes as C
import json
import struct

import numpy as np

from . import _capi
from ._capi import check, lib

MR_IN_REQ_F64, MR_IN_REQ_U64, MR_IN_REQ_VEC, MR_IN_ITEM_F64, MR_IN_REQ_TOKENS = 0, 1, 2, 3, 4
_SCOPE_TAG = {"global": 0, "item": 1, "user": 2, "session": 3, "field": 4, "irf": 5, "ranking": 6}


def hash64(s) -> int:
    b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    lib().mr_hash64.restype = C.c_uint64
    return int(lib().mr_hash64(b, C.c_size_t(len(b))))


def token_count(s: str) -> int:
    b = s.encode("utf-8")
    lib().mr_token_count.restype = C.c_int32
    return int(lib().mr_token_count(b, C.c_size_t(len(b))))


def _parse_iso(s: str):
    """ZonedDateTime.parse(value, ISO_DATE_TIME) for the local_time extractor (offset required)."""
    import datetime as dt
    import re

    s2 = re.sub(r"\[.*\]$", "", s)
    if s2.endswith("Z"):
        s2 = s2[:-1] + "+00:00"
    try:
        d = dt.datetime.fromisoformat(s2)
    except ValueError:
        return None
    return d if d.tzinfo is not None else None


def _map_datetime(parse: str, d) -> float:
    """LocalDateTimeFeature's DateTimeMapper (S/feature/LocalDateTimeFeature.scala:44-80)."""
    if parse == "time_of_day":
        return (d.hour * 3600 + d.minute * 60 + d.second) / 3600.0
    if parse == "day_of_week":
        return float(d.isoweekday())
    if parse == "month_of_year":
        return float(d.month)
    if parse == "year":
        return float(d.year)
    return float(int(d.timestamp()))  # "second"


def match_tokens(conf: dict, text: str) -> list[str]:
    """FieldMatcher.tokenize for the matchers whose analyzer is first-party enough to mirror here: language
    "whitespace" (Lucene WhitespaceTokenizer: split on whitespace, nothing else).  NgramMatcher.tokenize
    (S/feature/matcher/NgramMatcher.scala:9-30): every n-gram of every term; TermMatcher.tokenize
    (matcher/TermMatcher.scala:7-12): the terms; both sorted and unique (FieldMatcher.unique).  Any other
    language is a Lucene analyzer and has to be run by the caller (request["tokens"][feature])."""
    m = conf["method"]
    if m.get("language") != "whitespace":
        raise ValueError(f"feature {conf['name']}: language {m.get('language')!r} needs caller-side tokens")
    terms = text.split()
    if m["type"] == "ngram":
        n = int(m["n"])
        terms = [t[j:j + n] for t in terms for j in range(0, len(t) - n + 1)]
    return sorted(set(terms))


def bm25_idf(method: dict, token: str) -> float:
    """BM25Matcher.score's termIDF (S/feature/matcher/BM25Matcher.scala:26-27)."""
    import math

    gtf = method.get("termfreq", {}).get(token, 0)
    return math.log(1.0 + (method["docs"] - gtf + 0.5) / (gtf + 0.5))


def _field_name(s: str) -> tuple[str, str]:
    ev, fld = s.split(".", 1)
    return ("item" if ev == "metadata" else ev), fld


class RankBatch(C.Structure):
    _fields_ = [("n_requests", C.c_int32), ("item_offsets", C.c_void_p), ("item_ids", C.c_void_p),
                ("user_ids", C.c_void_p), ("session_ids", C.c_void_p), ("req_f64", C.c_void_p),
                ("req_u64", C.c_void_p), ("req_vec", C.c_void_p), ("req_vec_present", C.c_void_p),
                ("item_f64", C.c_void_p), ("req_tok_offsets", C.c_void_p), ("req_tok_hashes", C.c_void_p),
                ("req_tok_weights", C.c_void_p), ("max_items_per_request", C.c_int32)]


class StateInfo(C.Structure):
    _fields_ = [("rows", C.c_int64 * 6), ("device_bytes", C.c_int64), ("item_row_bytes", C.c_int64)]


class FeatureMapping:
    """mr_schema: features config (list of Metarank feature dicts) + the model's feature list."""

    def __init__(self, ctx, features: list[dict], model_features: list[str]):
        self.ctx = ctx
        self.features = [dict(f) for f in features]
        self.model_features = list(model_features)
        doc = json.dumps({"features": self.features, "model_features": self.model_features}).encode()
        self._h = C.c_void_p()
        check(lib().mr_schema_create(ctx.handle if ctx is not None else None, doc, C.c_size_t(len(doc)),
                                     C.byref(self._h)))
        self.dim = int(lib().mr_schema_dim(self._h))
        self.by_name = {f["name"]: f for f in self.features}
        self.n_req_f64 = self._n(MR_IN_REQ_F64)
        self.n_req_u64 = self._n(MR_IN_REQ_U64)
        self.n_req_vec = self._n(MR_IN_REQ_VEC)
        self.n_item_f64 = self._n(MR_IN_ITEM_F64)
        self.n_req_tok = self._n(MR_IN_REQ_TOKENS)
        self.vec_stride = int(lib().mr_schema_vec_stride(self._h))

    def _n(self, kind):
        n = C.c_int32(0)
        lib().mr_schema_input_slot(self._h, C.c_int32(kind), None, C.byref(n))
        return n.value

    def offset(self, feature: str):
        d = C.c_int32(0)
        o = int(lib().mr_schema_feature_offset(self._h, feature.encode(), C.byref(d)))
        return (o, d.value) if o >= 0 else None

    def input_slot(self, kind: int, feature: str) -> int:
        return int(lib().mr_schema_input_slot(self._h, C.c_int32(kind), feature.encode(), None))

    def vec_offset(self, slot: int):
        d = C.c_int32(0)
        o = int(lib().mr_schema_vec_offset(self._h, C.c_int32(slot), C.byref(d)))
        return o, d.value

    def free(self):
        if self._h:
            lib().mr_schema_free(self._h)
            self._h = C.c_void_p()

    # ---- request -> inputs (the JVM shim's job) -------------------------------------
    def _encode_string(self, conf, values):
        possible = conf["values"]
        if conf.get("encode", "onehot") == "index":
            return [float(possible.index(values[0]) + 1) if values and values[0] in possible else 0.0]
        out = [0.0] * len(possible)
        for v in values:
            if v in possible:
                out[possible.index(v)] = 1.0
        return out

    def pack_requests(self, requests: list[dict]):
        """RankingEvents -> the flat arrays of mr_rank_batch (numpy, kept alive by the caller)."""
        R = len(requests)
        offs = np.zeros(R + 1, dtype=np.int32)
        for r, q in enumerate(requests):
            offs[r + 1] = offs[r] + len(q["items"])
        N = int(offs[R])
        ids = np.zeros(max(N, 1), dtype=np.uint64)
        users = np.zeros(max(R, 1), dtype=np.uint64)
        sessions = np.zeros(max(R, 1), dtype=np.uint64)
        req_f64 = np.full((max(R, 1), max(self.n_req_f64, 1)), np.nan)
        req_u64 = np.zeros((max(R, 1), max(self.n_req_u64, 1)), dtype=np.uint64)
        req_vec = np.zeros((max(R, 1), max(self.vec_stride, 1)), dtype=np.float32)
        req_vp = np.zeros((max(R, 1), max(self.n_req_vec, 1)), dtype=np.uint8)
        item_f64 = np.full((max(N, 1), max(self.n_item_f64, 1)), np.nan)
        tok_lists = [[[] for _ in range(self.n_req_tok)] for _ in range(R)]  # [request][slot] -> [(hash, weight)]
        num = lambda v: isinstance(v, (int, float)) and not isinstance(v, bool)  # noqa: E731
        strl = lambda v: isinstance(v, list) and all(isinstance(x, str) for x in v)  # noqa: E731
        for r, q in enumerate(requests):
            if q.get("user") is not None:
                users[r] = hash64(q["user"])
            if q.get("session") is not None:
                sessions[r] = hash64(q["session"])
            rf = {n: v for n, v in q.get("fields", [])}  # fieldsMap: last duplicate wins
            rfirst = {}
            for n, v in q.get("fields", []):
                rfirst.setdefault(n, v)
            for name in self.model_features:
                conf = self.by_name.get(name)
                if conf is None:
                    continue
                t = conf["type"]
                if t in ("number", "word_count") and conf.get("scope") == "ranking":
                    slot = self.input_slot(MR_IN_REQ_F64, name)
                    v = rf.get(_field_name(conf.get("source", conf.get("field")))[1])
                    if t == "number" and num(v):
                        req_f64[r, slot] = float(v)
                    if t == "word_count" and isinstance(v, str):
                        req_f64[r, slot] = float(token_count(v))
                elif t == "string" and _field_name(conf.get("source", conf.get("field")))[0] == "ranking":
                    slot = self.input_slot(MR_IN_REQ_F64, name)
                    v = rfirst.get(_field_name(conf.get("source", conf.get("field")))[1])
                    enc = self._encode_string(conf, [v] if isinstance(v, str) else v if strl(v) else [])
                    req_f64[r, slot:slot + len(enc)] = enc
                elif t == "rate" and str(conf.get("scope", "item")).startswith("ranking."):
                    slot = self.input_slot(MR_IN_REQ_U64, name)
                    v = rf.get(conf["scope"].split(".", 1)[1])
                    if isinstance(v, str):
                        req_u64[r, slot] = hash64(v)
                elif t == "item_age":
                    req_u64[r, self.input_slot(MR_IN_REQ_U64, name)] = np.int64(int(q.get("timestamp", 0))).astype(np.uint64)
                elif t == "local_time":
                    import datetime as dt

                    fld = _field_name(conf["source"])[1]
                    d = None
                    if fld == "timestamp":
                        d = dt.datetime.fromtimestamp(int(q.get("timestamp", 0)) // 1000, tz=dt.timezone.utc)
                    elif isinstance(rf.get(fld), str):
                        d = _parse_iso(rf[fld])
                    if d is not None:
                        req_f64[r, self.input_slot(MR_IN_REQ_F64, name)] = _map_datetime(conf["parse"], d)
                elif t == "field_match" and conf["method"]["type"] in ("ngram", "term", "bm25"):
                    slot = self.input_slot(MR_IN_REQ_TOKENS, name)
                    qf = rf.get(_field_name(conf["rankingField"])[1])
                    if isinstance(qf, str):  # only a StringField is tokenized (FieldMatchFeature.scala:62-68)
                        toks = (q.get("tokens") or {}).get(name)
                        if toks is None:
                            toks = match_tokens(conf, qf)
                        w = [bm25_idf(conf["method"], tk) for tk in toks] if conf["method"]["type"] == "bm25" else [0.0] * len(toks)
                        tok_lists[r][slot] = [(hash64(tk), wk) for tk, wk in zip(toks, w)]
                elif t == "field_match":
                    slot = self.input_slot(MR_IN_REQ_VEC, name)
                    q_emb = (q.get("embeddings") or {}).get(name)
                    qf = rf.get(_field_name(conf["rankingField"])[1])
                    if q_emb is not None and (isinstance(qf, str) or strl(qf)):  # no query field -> feature missing
                        o, d = self.vec_offset(slot)
                        req_vec[r, o:o + d] = np.asarray(q_emb, dtype=np.float32)
                        req_vp[r, slot] = 1
            for j, it in enumerate(q["items"]):
                i = int(offs[r]) + j
                ids[i] = hash64(it["id"])
                flds = it.get("fields", [])
                for name in self.model_features:
                    conf = self.by_name.get(name)
                    if conf is None:
                        continue
                    t = conf["type"]
                    if t == "relevancy":
                        first = next(((n, v) for n, v in flds if n == "relevancy"), None)
                        if first is not None and num(first[1]):
                            item_f64[i, self.input_slot(MR_IN_ITEM_F64, name)] = float(first[1])
                    elif t == "number" and conf.get("scope") != "ranking":
                        fld = _field_name(conf.get("source", conf.get("field")))[1]
                        ov = next((v for n, v in flds if n == fld and num(v)), None)
                        if ov is not None:
                            item_f64[i, self.input_slot(MR_IN_ITEM_F64, name)] = float(ov)
                    elif t == "string" and _field_name(conf.get("source", conf.get("field")))[0] != "ranking":
                        fld = _field_name(conf.get("source", conf.get("field")))[1]
                        ov = next(([v] if isinstance(v, str) else list(v) for n, v in flds
                                   if n == fld and (isinstance(v, str) or strl(v))), None)
                        if ov is not None:
                            slot = self.input_slot(MR_IN_ITEM_F64, name)
                            enc = self._encode_string(conf, ov)
                            item_f64[i, slot:slot + len(enc)] = enc
        if self.n_item_f64 == 0 or np.isnan(item_f64).all():
            item_f64 = None  # mr_rank_batch.item_f64 == NULL: no per-item inputs at all
        return dict(offsets=offs, ids=ids, users=users, sessions=sessions, req_f64=req_f64, req_u64=req_u64,
                    req_vec=req_vec, req_vp=req_vp, item_f64=item_f64, n_requests=R, total_items=N,
                    **self._pack_tokens(tok_lists))

    def _pack_tokens(self, tok_lists):
        if self.n_req_tok == 0:
            return {}
        offs, hs, ws = [0], [], []
        for per_req in tok_lists:
            for lst in per_req:
                hs += [h for h, _ in lst]
                ws += [w for _, w in lst]
                offs.append(len(hs))
        return dict(tok_off=np.asarray(offs, dtype=np.int32), tok_hash=np.asarray(hs + [0], dtype=np.uint64),
                    tok_w=np.asarray(ws + [0.0], dtype=np.float64))


def pack_feature_values(values: dict) -> bytes:
    """Map[Key, FeatureValue] -> mr_state_upsert wire format (include/mr_b200.h)."""
    out = []
    for (scope, name), (kind, v) in values.items():
        nb = name.encode("utf-8")
        out.append(struct.pack("<H", len(nb)) + nb)
        tag = _SCOPE_TAG[scope[0]]
        out.append(struct.pack("<B", tag))
        if tag in (1, 2, 3, 6):
            out.append(struct.pack("<Q", hash64(scope[1])))
        elif tag == 4:
            out.append(struct.pack("<Q", hash64(scope[2])))
        elif tag == 5:
            out.append(struct.pack("<QQ", hash64(scope[2]), hash64(scope[3])))
        if kind == "scalar":
            if isinstance(v, bool):
                out.append(struct.pack("<BB", 7, 1 if v else 0))
            elif isinstance(v, (int, float)):
                out.append(struct.pack("<Bd", 0, float(v)))
            elif isinstance(v, str):
                out.append(struct.pack("<BQ", 1, hash64(v)))
            elif isinstance(v, (list, tuple, np.ndarray)) and len(v) and not isinstance(v[0], str):
                a = np.asarray(v, dtype=np.float64)
                out.append(struct.pack("<BI", 3, a.size) + a.tobytes())
            else:
                hs = np.array([hash64(x) for x in v], dtype=np.uint64)
                out.append(struct.pack("<BI", 2, hs.size) + hs.tobytes())
        elif kind == "counter":
            out.append(struct.pack("<Bq", 4, int(v)))
        elif kind == "pcounter":
            a = np.asarray(v, dtype=np.int64)
            out.append(struct.pack("<BI", 5, a.size) + a.tobytes())
        elif kind == "blist":
            hs = np.array([hash64(x) for x in v], dtype=np.uint64)
            out.append(struct.pack("<BI", 6, hs.size) + hs.tobytes())
        else:
            raise ValueError(kind)
    return b"".join(out)


class DecodedRequests:
    """mr_requests_decode: RankingEvent JSON (one object or an array) -> the packed mr_rank_batch, natively."""

    def __init__(self, mapping: "FeatureMapping", body):
        data = body.encode("utf-8") if isinstance(body, str) else bytes(body)
        self._h = C.c_void_p()
        check(lib().mr_requests_decode(mapping._h, data, C.c_size_t(len(data)), C.byref(self._h)))
        n = C.c_int32(0)
        lib().mr_requests_batch.restype = C.POINTER(RankBatch)
        self.batch = lib().mr_requests_batch(self._h, C.byref(n)).contents
        self.total_items = n.value
        self.n_requests = self.batch.n_requests
        self.mapping = mapping

    def arrays(self) -> dict:
        """Copies of the packed arrays, shaped like FeatureMapping.pack_requests returns them."""
        m, b, R, N = self.mapping, self.batch, self.n_requests, self.total_items

        def arr(ptr, ctype, count, dtype, shape=None):
            if not ptr or count == 0:
                return None
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).astype(dtype).copy()
            return a.reshape(shape) if shape else a

        out = dict(n_requests=R, total_items=N, offsets=arr(b.item_offsets, C.c_int32, R + 1, np.int32),
                   ids=arr(b.item_ids, C.c_uint64, max(N, 1), np.uint64),
                   users=arr(b.user_ids, C.c_uint64, max(R, 1), np.uint64),
                   sessions=arr(b.session_ids, C.c_uint64, max(R, 1), np.uint64),
                   req_f64=arr(b.req_f64, C.c_double, max(R, 1) * max(m.n_req_f64, 1), np.float64, (max(R, 1), max(m.n_req_f64, 1))),
                   req_u64=arr(b.req_u64, C.c_uint64, max(R, 1) * max(m.n_req_u64, 1), np.uint64, (max(R, 1), max(m.n_req_u64, 1))),
                   req_vec=arr(b.req_vec, C.c_float, max(R, 1) * max(m.vec_stride, 1), np.float32, (max(R, 1), max(m.vec_stride, 1))),
                   req_vp=arr(b.req_vec_present, C.c_uint8, max(R, 1) * max(m.n_req_vec, 1), np.uint8, (max(R, 1), max(m.n_req_vec, 1))),
                   item_f64=arr(b.item_f64, C.c_double, max(N, 1) * max(m.n_item_f64, 1), np.float64, (max(N, 1), max(m.n_item_f64, 1))))
        if m.n_req_tok:
            n_off = R * m.n_req_tok + 1
            out["tok_off"] = arr(b.req_tok_offsets, C.c_int32, n_off, np.int32)
            n_tok = int(out["tok_off"][-1])
            out["tok_hash"] = arr(b.req_tok_hashes, C.c_uint64, n_tok + 1, np.uint64)
            out["tok_w"] = arr(b.req_tok_weights, C.c_double, n_tok + 1, np.float64)
        return out

    def item_id(self, index: int) -> str:
        n = C.c_size_t(0)
        lib().mr_requests_item_id.restype = C.c_void_p
        p = lib().mr_requests_item_id(self._h, C.c_int32(index), C.byref(n))
        return C.string_at(p, n.value).decode("utf-8")

    def timestamp(self, request: int) -> int:
        lib().mr_requests_timestamp.restype = C.c_int64
        return int(lib().mr_requests_timestamp(self._h, C.c_int32(request)))

    def free(self):
        if self._h:
            lib().mr_requests_free(self._h)
            self._h = C.c_void_p()


def transcode_feature_values(blob: bytes):
    """mr_feature_values_transcode: reference binary FeatureValue stream -> mr_state_upsert records (host
    only).  Returns (records bytes, n decoded, n unsupported, consumed bytes)."""
    n, nr, nu, c = C.c_size_t(0), C.c_int64(0), C.c_int64(0), C.c_size_t(0)
    src = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b"\0")
    check(lib().mr_feature_values_transcode(src, C.c_size_t(len(blob)), None, C.c_size_t(0), C.byref(n), C.byref(nr),
                                            C.byref(nu), C.byref(c)))
    out = (C.c_uint8 * max(n.value, 1))()
    check(lib().mr_feature_values_transcode(src, C.c_size_t(len(blob)), out, C.c_size_t(n.value), C.byref(n), C.byref(nr),
                                            C.byref(nu), C.byref(c)))
    return bytes(out[:n.value]), nr.value, nu.value, c.value


def _pack_scope(scope) -> bytes:
    tag = _SCOPE_TAG[scope[0]]
    out = struct.pack("<B", tag)
    if tag in (1, 2, 3, 6):
        out += struct.pack("<Q", hash64(scope[1]))
    elif tag == 4:
        out += struct.pack("<Q", hash64(scope[2]))
    elif tag == 5:
        out += struct.pack("<QQ", hash64(scope[2]), hash64(scope[3]))
    return out


def pack_writes(writes: list) -> bytes:
    """The extractors' raw writes -> mr_state_apply_writes wire format.  A write is
    ("put", key, ts, scalar) | ("inc", key, ts, n) | ("pinc", key, ts, n) | ("append", key, ts, item_id)
    with key = (scope, feature_name) — Write.{Put, Increment, PeriodicIncrement, Append} (S/model/Write.scala)."""
    out = []
    for kind, (scope, name), ts, v in writes:
        nb = name.encode("utf-8")
        head = struct.pack("<H", len(nb)) + nb + _pack_scope(scope)
        if kind == "put":
            val = pack_feature_values({(scope, name): ("scalar", v)})
            out.append(head + struct.pack("<Bq", 0, int(ts)) + val[len(head):])
        elif kind == "inc":
            out.append(head + struct.pack("<Bqq", 1, int(ts), int(v)))
        elif kind == "pinc":
            out.append(head + struct.pack("<Bqq", 2, int(ts), int(v)))
        elif kind == "append":
            items = v if isinstance(v, list) else [v]
            for it in reversed(items):  # a list is prepended as a block, keeping its order (MemBoundedList.put)
                out.append(head + struct.pack("<BqQ", 3, int(ts), hash64(it)))
        else:
            raise ValueError(kind)
    return b"".join(out)


class DeviceState:
    """mr_state: device-resident Persistence.values for one FeatureMapping."""

    def __init__(self, ctx, mapping: FeatureMapping):
        self.mapping = mapping
        self._h = C.c_void_p()
        check(lib().mr_state_create(ctx.handle, mapping._h, C.byref(self._h)))

    def put_packed(self, blob: bytes):
        a, s = C.c_int64(0), C.c_int64(0)
        buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b"\0")
        check(lib().mr_state_upsert(self._h, buf, C.c_size_t(len(blob)), C.byref(a), C.byref(s)))
        return a.value, s.value

    def put(self, values: dict):
        """KVStore.put(Map[Key, FeatureValue])"""
        return self.put_packed(pack_feature_values(values))

    def load_feature_values(self, blob: bytes):
        """Bulk load from the reference's binary store format: a stream of delimited FeatureValues
        (BinaryVCodec.encodeDelimited over FeatureValueCodec).  Returns (applied, skipped, consumed bytes)."""
        a, s, c = C.c_int64(0), C.c_int64(0), C.c_size_t(0)
        buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b"\0")
        check(lib().mr_state_load_feature_values(self._h, buf, C.c_size_t(len(blob)), C.byref(a), C.byref(s), C.byref(c)))
        return a.value, s.value, c.value

    def apply_writes(self, writes: list):
        """FeatureValueFlow.commitWrite for a batch of raw writes (see pack_writes)."""
        blob = pack_writes(writes)
        a, s = C.c_int64(0), C.c_int64(0)
        buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b"\0")
        check(lib().mr_state_apply_writes(self._h, buf, C.c_size_t(len(blob)), C.byref(a), C.byref(s)))
        return a.value, s.value

    def flush(self):
        check(lib().mr_state_flush(self._h))

    def info(self) -> StateInfo:
        inf = StateInfo()
        check(lib().mr_state_get_info(self._h, C.byref(inf)))
        return inf

    def free(self):
        if self._h:
            lib().mr_state_free(self._h)
            self._h = C.c_void_p()


class Ranker:
    """Ranker(mapping, store).rerank (S/ml/Ranker.scala:27-83), batched."""

    def __init__(self, mapping: FeatureMapping, state: DeviceState):
        self.mapping, self.state = mapping, state

    def rank_arrays(self, arrays: dict, model=None, want_order=True, want_features=False, out_scores=None,
                    out_order=None):
        """mr_rank on already-packed arrays; returns (scores, order, features).  out_scores / out_order may be
        caller-provided (e.g. page-locked) numpy arrays; page-locked buffers are DMA'd in place."""
        N = arrays["total_items"]
        b = RankBatch(arrays["n_requests"], arrays["offsets"].ctypes.data, arrays["ids"].ctypes.data,
                      arrays["users"].ctypes.data, arrays["sessions"].ctypes.data, arrays["req_f64"].ctypes.data,
                      arrays["req_u64"].ctypes.data, arrays["req_vec"].ctypes.data, arrays["req_vp"].ctypes.data,
                      arrays["item_f64"].ctypes.data if arrays["item_f64"] is not None else None,
                      arrays["tok_off"].ctypes.data if "tok_off" in arrays else None,
                      arrays["tok_hash"].ctypes.data if "tok_hash" in arrays else None,
                      arrays["tok_w"].ctypes.data if "tok_w" in arrays else None)
        scores = out_scores if out_scores is not None else np.empty(max(N, 1), dtype=np.float64)
        order = (out_order if out_order is not None else np.empty(max(N, 1), dtype=np.int32)) \
            if want_order and model is not None else None
        feats = np.empty((max(N, 1), max(self.mapping.dim, 1)), dtype=np.float64) if want_features else None
        check(lib().mr_rank(self.state._h, model._h if model is not None else None, C.byref(b),
                            C.c_void_p(scores.ctypes.data),
                            C.c_void_p(order.ctypes.data) if order is not None else None,
                            C.c_void_p(feats.ctypes.data) if feats is not None else None))
        return (scores[:N] if model is not None else None, order[:N] if order is not None else None,
                feats[:N, :self.mapping.dim] if feats is not None else None)

    def rank_decoded(self, dec: "DecodedRequests", model=None, want_order=True, want_features=False):
        """mr_rank straight on the batch mr_requests_decode packed (no Python-side packing)."""
        N = dec.total_items
        scores = np.empty(max(N, 1), dtype=np.float64)
        order = np.empty(max(N, 1), dtype=np.int32) if want_order and model is not None else None
        feats = np.empty((max(N, 1), max(self.mapping.dim, 1)), dtype=np.float64) if want_features else None
        check(lib().mr_rank(self.state._h, model._h if model is not None else None, C.byref(dec.batch),
                            C.c_void_p(scores.ctypes.data),
                            C.c_void_p(order.ctypes.data) if order is not None else None,
                            C.c_void_p(feats.ctypes.data) if feats is not None else None))
        return (scores[:N] if model is not None else None, order[:N] if order is not None else None,
                feats[:N, :self.mapping.dim] if feats is not None else None)

    def make_query(self, requests: list[dict]) -> list[np.ndarray]:
        """Ranker.makeQuery: the dense f64[N x dim] matrix per request (ltrlib Query.values)."""
        arrays = self.mapping.pack_requests(requests)
        _, _, feats = self.rank_arrays(arrays, None, want_order=False, want_features=True)
        offs = arrays["offsets"]
        return [feats[offs[r]:offs[r + 1]] for r in range(len(requests))]

    def rerank(self, requests: list[dict], model, explain: bool = False) -> list[dict]:
        """One RankResponse-shaped dict per request: items sorted by -score (stable)."""
        arrays = self.mapping.pack_requests(requests)
        scores, order, feats = self.rank_arrays(arrays, model, want_order=True, want_features=explain)
        offs = arrays["offsets"]
        out = []
        for r, q in enumerate(requests):
            b = int(offs[r])
            items = []
            for k in order[b:int(offs[r + 1])]:
                e = {"item": q["items"][k]["id"], "score": float(scores[b + k])}
                if explain:
                    e["features"] = feats[b + k].tolist()
                items.append(e)
            out.append({"items": items})
        return out


def pack_number_columns(names: list[str], ids_u64: np.ndarray, columns: np.ndarray) -> bytes:
    """Vectorised mr_state_upsert packing of item-scoped SDouble scalars: one record per
    (item, feature) for a dense [n_items x n_features] matrix (NaN entries are skipped = missing)."""
    out = []
    ids_u64 = np.ascontiguousarray(ids_u64, dtype=np.uint64)
    for j, name in enumerate(names):
        nb = name.encode("utf-8")
        dt = np.dtype([("nl", "<u2"), ("name", f"S{len(nb)}"), ("scope", "u1"), ("id", "<u8"), ("kind", "u1"),
                       ("val", "<f8")])
        col = columns[:, j]
        keep = ~np.isnan(col)
        rec = np.zeros(int(keep.sum()), dtype=dt)
        rec["nl"] = len(nb)
        rec["name"] = nb
        rec["scope"] = 1
        rec["id"] = ids_u64[keep]
        rec["kind"] = 0
        rec["val"] = col[keep]
        out.append(rec.tobytes())
    return b"".join(out)


def rank_device(state: DeviceState, model, n_requests: int, total_items: int, d_offsets: int, d_item_ids: int,
                d_scores: int, d_order: int = 0, d_features: int = 0, stream: int = 0, d_user_ids: int = 0,
                d_session_ids: int = 0, max_items: int = 0) -> None:
    """mr_rank_device: every pointer is a device address; enqueued on `stream`, no sync.
    max_items: the largest request's item count when known (mr_rank_batch.max_items_per_request)."""
    b = RankBatch(n_requests, d_offsets, d_item_ids, d_user_ids or None, d_session_ids or None, None, None, None,
                  None, None)
    b.max_items_per_request = int(max_items)
    check(lib().mr_rank_device(state._h, model._h if model is not None else None, C.byref(b),
                               C.c_int32(total_items), C.c_void_p(d_scores), C.c_void_p(d_order or None),
                               C.c_void_p(d_features or None), C.c_void_p(stream)))


def rank_device_status(state: DeviceState, stream: int = 0) -> None:
    check(lib().mr_rank_device_status(state._h, C.c_void_p(stream)))
