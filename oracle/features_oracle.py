"""ORACLE (test infrastructure only) — CPU restatement of Metarank's feature-vector
assembly on the /rank path, plus the slice of the WRITE path that the reference's own
tests use to produce state, so the golden vectors of src/test/scala/ai/metarank/feature/*
can be replayed here.

Read path restated (S/ = src/main/scala/ai/metarank/):
  FeatureValueLoader.fromStateBackend  S/fstore/FeatureValueLoader.scala:11-25  (state is a dict, so
                                        the two batched reads collapse to dict lookups)
  ItemValue.fromState                  S/model/ItemValue.scala:25-72
  ClickthroughQuery.collectFeatureValues S/flow/ClickthroughQuery.scala:50-74
  extractors' value()/values()         file:line cited on each class below
Write path restated (only to replay reference tests):
  FeatureValueFlow.process             S/flow/FeatureValueFlow.scala:24-92
  Mem{Scalar,Counter,PeriodicCounter,BoundedList}  S/fstore/memory/*.scala
  PeriodicCounterFeature.fromMap       S/model/Feature.scala:140-162

Data model (plain Python, shared by tests when they feed the CUDA path):
  Key   = (scope, feature_name); scope = ("global",) | ("item", id) | ("user", id) |
          ("session", id) | ("field", name, value) | ("irf", name, value, item) | ("ranking", id)
  Value = ("scalar", float | str | list[str] | list[float] | bool) | ("counter", int) |
          ("pcounter", [int, ...]) | ("blist", [item ids, newest first])
  Event = dict(event="item"|"user"|"interaction"|"ranking", id, timestamp(ms), fields=[(name, value)],
          item, user, session, ranking, type, items=[dict(id, fields=[...])])
Pinned by tests/test_features_golden.py against the reference's test vectors.
"""
from __future__ import annotations

import math
import re

import numpy as np

NAN = float("nan")
DAY = 24 * 3600 * 1000


# --------------------------------------------------------------------------- helpers

def parse_duration_ms(s) -> int:
    """DurationJson: '24h', '7d', '60s', '10m', '100ms'."""
    if isinstance(s, (int, float)):
        return int(s)
    m = re.fullmatch(r"\s*(\d+)\s*(ms|s|m|h|d)\s*", s)
    if not m:
        raise ValueError(f"cannot parse duration {s!r}")
    return int(m.group(1)) * {"ms": 1, "s": 1000, "m": 60_000, "h": 3_600_000, "d": DAY}[m.group(2)]


def parse_scope_type(s: str):
    """ScopeType decoder (S/model/ScopeType.scala)."""
    if s in ("global", "item", "user", "session", "ranking"):
        return (s,)
    m = re.fullmatch(r"item\.([a-zA-Z0-9\-_]+)", s)
    if m:
        return ("item_field", m.group(1))
    m = re.fullmatch(r"ranking\.([a-zA-Z0-9\-_]+)", s)
    if m:
        return ("ranking_field", m.group(1))
    raise ValueError(f"scope type {s} not supported")


def parse_field_name(s: str):
    """FieldName decoder (S/model/FieldName.scala): returns (event, field)."""
    m = re.fullmatch(r"interaction:([a-zA-Z0-9_]+)\.([a-zA-Z0-9_]+)", s)
    if m:
        return ("interaction:" + m.group(1), m.group(2))
    m = re.fullmatch(r"([a-z\*]+)\.([a-zA-Z0-9_]+)", s)
    if not m:
        raise ValueError(f"cannot decode source field '{s}'")
    src = {"metadata": "item", "item": "item", "user": "user", "ranking": "ranking", "*": "*"}.get(m.group(1))
    if src is None:
        raise ValueError(f"cannot decode source field {m.group(1)}")
    return (src, m.group(2))


def fields_map(fields):
    """Event.fieldsMap: fields.map(f => f.name -> f).toMap — the LAST duplicate wins."""
    return {n: v for n, v in fields}


def is_num(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def is_strlist(v):
    return isinstance(v, list) and all(isinstance(x, str) for x in v)


def token_count(s: str) -> int:
    """WordCountFeature.tokenCount: "\\s+".r.split(s).length with java.util.regex semantics
    (S/feature/WordCountFeature.scala:73-76).  Java's \\s = [ \\t\\n\\x0B\\f\\r]; a leading match
    yields an empty first token, trailing empty tokens are removed, "" -> [""]."""
    if s == "":
        return 1
    parts = re.split(r"[ \t\n\x0b\f\r]+", s)
    while parts and parts[-1] == "":
        parts.pop()
    return len(parts)


def to_start_of_period(ts: int, period: int) -> int:
    """Timestamp.toStartOfPeriod (S/model/Timestamp.scala:18-21): floor(ts.toDouble / period)."""
    return int(math.floor(float(ts) / period)) * period


def java_long_div(a: int, b: int) -> int:
    """Long / Long: truncation toward zero; / 0 raises ArithmeticException."""
    if b == 0:
        raise ZeroDivisionError("java.lang.ArithmeticException: / by zero")
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def fdiv(a: float, b: float) -> float:
    """IEEE double division (Python raises on /0.0, the JVM does not)."""
    if b == 0 or not math.isfinite(b) or not math.isfinite(a):
        with np.errstate(all="ignore"):
            return float(np.float64(a) / np.float64(b))
    return a / b


# --------------------------------------------------------------------------- state primitives (write path)

class MemState:
    """The four Mem* features (S/fstore/memory/Mem{ScalarFeature,Counter,PeriodicCounter,BoundedList}.scala)."""

    def __init__(self):
        self.scalars, self.counters, self.pcounters, self.lists = {}, {}, {}, {}

    def put(self, w, configs):
        kind, key = w[0], w[1]
        if kind == "put":
            self.scalars[key] = w[3]
        elif kind == "inc":
            self.counters[key] = self.counters.get(key, 0) + w[3]
        elif kind == "pinc":
            conf = configs[key[1]]
            bucket = to_start_of_period(w[2], conf["period"])
            m = self.pcounters.setdefault(key, {})
            m[bucket] = m.get(bucket, 0) + w[3]
        elif kind == "append":
            conf = configs[key[1]]
            ts, value = w[2], w[3]
            new = [(ts, v) for v in value] if isinstance(value, list) else [(ts, value)]
            if key not in self.lists:
                self.lists[key] = new  # first write is NOT trimmed (MemBoundedList.put :22-29)
            else:
                res = new + self.lists[key]
                res = [tv for tv in res if tv[0] >= ts - conf["duration"]][: conf["count"]]
                self.lists[key] = res
        else:
            raise ValueError(kind)

    def compute_value(self, w, configs):
        kind, key = w[0], w[1]
        if kind == "put":
            return ("scalar", self.scalars[key]) if key in self.scalars else None
        if kind == "inc":
            return ("counter", self.counters[key]) if key in self.counters else None
        if kind == "pinc":
            if key not in self.pcounters:
                return None
            return ("pcounter", periodic_from_map(self.pcounters[key], configs[key[1]]))
        if kind == "append":
            return ("blist", [v for _, v in self.lists[key]]) if key in self.lists else None
        raise ValueError(kind)


def periodic_from_map(buckets: dict, conf) -> list:
    """PeriodicCounterFeature.fromMap (S/model/Feature.scala:140-162): windows are anchored
    at the LAST bucket, [last - period*start, last - period*end + period], both inclusive."""
    ts = sorted(buckets)
    if not ts:
        return []
    last = ts[-1]
    out = []
    for start_off, end_off in conf["ranges"]:
        start = last - conf["period"] * start_off
        end = last - conf["period"] * end_off + conf["period"]
        out.append(sum(buckets[t] for t in ts if start <= t <= end))
    return out


# --------------------------------------------------------------------------- extractors

class BaseFeature:
    dim = 1
    name = ""
    scope = ("item",)

    def states(self):  # name -> config
        return {}

    def writes(self, ev, store):
        return []

    # BaseFeature.writeKey (S/feature/BaseFeature.scala:18-26)
    def write_key(self, ev, scope, name):
        e = ev["event"]
        if scope == ("global",):
            return (("global",), name)
        if scope == ("user",) and e == "interaction":
            return (("user", ev["user"]), name) if ev.get("user") is not None else None
        if scope == ("user",) and e == "user":
            return (("user", ev["user"]), name)
        if scope == ("session",) and e == "interaction":
            return (("session", ev["session"]), name) if ev.get("session") is not None else None
        if scope == ("item",) and e in ("interaction", "item"):
            return (("item", ev["item"]), name)
        return None

    # BaseFeature.readKey (S/feature/BaseFeature.scala:28-36)
    def read_key(self, req, scope, name, item_id):
        if scope == ("global",):
            return (("global",), name)
        if scope == ("item",):
            return (("item", item_id), name)
        if scope == ("user",):
            return (("user", req["user"]), name) if req.get("user") is not None else None
        if scope == ("session",):
            return (("session", req["session"]), name) if req.get("session") is not None else None
        if scope == ("ranking",):
            return (("ranking", req["id"]), name)
        return None

    def value(self, req, state, item):
        raise NotImplementedError

    def values(self, req, state, mode="online"):
        return [self.value(req, state, it) for it in req["items"]]


class NumberFeature(BaseFeature):
    """S/feature/NumberFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        self.scope = parse_scope_type(c["scope"])
        self.field = parse_field_name(c.get("source", c.get("field")))[1]
        self.refresh = parse_duration_ms(c.get("refresh", 0))

    def states(self):
        return {self.name: dict(kind="scalar", scope=self.scope, refresh=self.refresh)}

    def writes(self, ev, store):  # :44-56
        key = self.write_key(ev, self.scope, self.name)
        if key is None:
            return []
        for n, v in ev.get("fields", []):
            if n == self.field:
                return [("put", key, ev["timestamp"], float(v))] if is_num(v) else []
        return []

    def value(self, req, state, item):  # :58-69
        key = self.read_key(req, self.scope, self.name, item["id"])
        fv = state.get(key) if key is not None else None
        if fv is not None and fv[0] == "scalar" and is_num(fv[1]):
            return [float(fv[1])]
        return [NAN]

    def values(self, req, state, mode="online"):  # :71-97
        if self.scope == ("ranking",):
            v = fields_map(req.get("fields", [])).get(self.field)
            return [[float(v)] if is_num(v) else [NAN] for _ in req["items"]]
        out = []
        for it in req["items"]:
            ov = next((v for n, v in it.get("fields", []) if n == self.field and is_num(v)), None)
            out.append([float(ov)] if ov is not None else self.value(req, state, it))
        return out


class WordCountFeature(BaseFeature):
    """S/feature/WordCountFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        self.scope = parse_scope_type(c["scope"])
        self.field = parse_field_name(c["source"])[1]

    def states(self):
        return {self.name: dict(kind="scalar", scope=self.scope, refresh=0)}

    def writes(self, ev, store):  # :36-48 (fields.find: the FIRST field of that name)
        key = self.write_key(ev, self.scope, self.name)
        if key is None:
            return []
        for n, v in ev.get("fields", []):
            if n == self.field:
                return [("put", key, ev["timestamp"], float(token_count(v)))] if isinstance(v, str) else []
        return []

    def value(self, req, state, item):  # :53-70
        if self.scope == ("ranking",):
            v = fields_map(req.get("fields", [])).get(self.field)
            return [float(token_count(v))] if isinstance(v, str) else [NAN]
        key = self.read_key(req, self.scope, self.name, item["id"])
        fv = state.get(key) if key is not None else None
        if fv is not None and fv[0] == "scalar" and is_num(fv[1]):
            return [float(fv[1])]
        return [NAN]


class RefererFeature(BaseFeature):
    """S/feature/RefererFeature.scala:39-110.  The URL is parsed at WRITE time (snowplow referers.json, a third-party
    database that stays with the JVM): writeField :70-90 stores SString(medium) under the user or the session.  value()
    :94-109 reads it back and maps it through `possibleValues` :47-54; absent / not a string / not in the table ->
    CategoryValue("unknown", 0)."""

    MEDIUMS = {"unknown": 0, "search": 1, "internal": 2, "social": 3, "email": 4, "paid": 5}

    def __init__(self, c):
        self.name = c["name"]
        self.scope = parse_scope_type(c["scope"])
        self.src_event, self.field = parse_field_name(c["source"])
        self.dim = 1
        self.categorical = True

    def states(self):
        return {self.name: dict(kind="scalar", scope=self.scope, refresh=0)}

    def writes(self, ev, store):
        return []  # the medium is written by the JVM's parser; tests put SString(medium) into the state directly

    def value(self, req, state, item):
        if self.scope[0] not in ("user", "session"):
            return [0.0]
        key = self.read_key(req, self.scope, self.name, item["id"])
        fv = state.get(key) if key is not None else None
        if fv is not None and fv[0] == "scalar" and isinstance(fv[1], str):
            return [float(self.MEDIUMS.get(fv[1], 0))]
        return [0.0]


class StringFeature(BaseFeature):
    """S/feature/StringFeature.scala (index + onehot encoders)"""

    def __init__(self, c):
        self.name = c["name"]
        self.scope = parse_scope_type(c["scope"])
        self.src_event, self.field = parse_field_name(c.get("source", c.get("field")))
        self.encode_kind = c.get("encode", "onehot")
        self.possible = list(c["values"])
        self.dim = 1 if self.encode_kind == "index" else len(self.possible)
        self.categorical = self.encode_kind == "index"

    def states(self):
        return {self.name: dict(kind="scalar", scope=self.scope, refresh=0)}

    def encode(self, values):
        if self.encode_kind == "index":  # :124-137 — CategoryValue(index + 1), 0 = nil
            if values and values[0] in self.possible:
                return [float(self.possible.index(values[0]) + 1)]
            return [0.0]
        out = [0.0] * self.dim  # OneHotEncoder.fromValues
        for v in values:
            if v in self.possible:
                out[self.possible.index(v)] = 1.0
        return out

    def writes(self, ev, store):
        key = self.write_key(ev, self.scope, self.name)
        if key is None:
            return []
        for n, v in ev.get("fields", []):
            if n == self.field:
                if isinstance(v, str):
                    return [("put", key, ev["timestamp"], [v])]
                if is_strlist(v):
                    return [("put", key, ev["timestamp"], list(v))]
                return []
        return []

    def value(self, req, state, item):  # :70-79
        key = self.read_key(req, self.scope, self.name, item["id"])
        fv = state.get(key) if key is not None else None
        if fv is not None and fv[0] == "scalar" and is_strlist(fv[1]):
            return self.encode(fv[1])
        return self.encode([])

    def values(self, req, state, mode="online"):  # :81-107
        if self.src_event == "ranking":
            v = next((v for n, v in req.get("fields", []) if n == self.field), None)
            const = self.encode([v]) if isinstance(v, str) else self.encode(v) if is_strlist(v) else self.encode([])
            return [const for _ in req["items"]]
        out = []
        for it in req["items"]:
            ov = None
            for n, v in it.get("fields", []):
                if n == self.field and isinstance(v, str):
                    ov = [v]
                    break
                if n == self.field and is_strlist(v):
                    ov = list(v)
                    break
            out.append(self.encode(ov) if ov is not None else self.value(req, state, it))
        return out


class InteractionCountFeature(BaseFeature):
    """S/feature/InteractionCountFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        self.scope = parse_scope_type(c["scope"])
        self.interaction = c["interaction"]

    def states(self):
        return {self.name: dict(kind="counter", scope=self.scope, refresh=0)}

    def writes(self, ev, store):
        if ev["event"] == "interaction" and ev["type"] == self.interaction:
            key = self.write_key(ev, self.scope, self.name)
            return [("inc", key, ev["timestamp"], 1)] if key is not None else []
        return []

    def value(self, req, state, item):  # :44-59 — missing is 0.0
        key = self.read_key(req, self.scope, self.name, item["id"])
        fv = state.get(key) if key is not None else None
        return [float(fv[1])] if fv is not None and fv[0] == "counter" else [0.0]


class WindowInteractionCountFeature(BaseFeature):
    """S/feature/WindowInteractionCountFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        self.scope = parse_scope_type(c["scope"])
        self.interaction = c["interaction"]
        self.period = parse_duration_ms(c["bucket"])
        self.periods = list(c["periods"])
        self.dim = len(self.periods)

    def states(self):
        return {self.name: dict(kind="pcounter", scope=self.scope, refresh=0, period=self.period,
                                ranges=[(p, 0) for p in self.periods])}

    def writes(self, ev, store):
        key = self.write_key(ev, self.scope, self.name)
        if key is not None and ev["event"] == "interaction" and ev["type"] == self.interaction:
            return [("pinc", key, ev["timestamp"], 1)]
        return []

    def value(self, req, state, item):  # :50-63
        key = self.read_key(req, self.scope, self.name, item["id"])
        fv = state.get(key) if key is not None else None
        if fv is not None and fv[0] == "pcounter" and len(fv[1]) == self.dim:
            return [float(v) for v in fv[1]]
        return [NAN] * self.dim


class RateFeature(BaseFeature):
    """S/feature/RateFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        self.top, self.bottom = c["top"], c["bottom"]
        self.scope = parse_scope_type(c["scope"]) if c.get("scope") else ("item",)
        self.period = parse_duration_ms(c["bucket"])
        self.periods = list(c["periods"])
        self.dim = len(self.periods)
        self.weight = c["normalize"]["weight"] if c.get("normalize") else None
        self.refresh_opt = parse_duration_ms(c["refresh"]) if c.get("refresh") is not None else None
        self.refresh = self.refresh_opt if self.refresh_opt is not None else 3_600_000  # getOrElse(1.hour)
        n = self.name
        self.top_target, self.bottom_target = f"{n}_{self.top}", f"{n}_{self.bottom}"
        self.top_global, self.bottom_global = f"{n}_{self.top}_norm", f"{n}_{self.bottom}_norm"
        self.item_field, self.ranking_field = f"{n}_field", f"{n}_rfield"

    def states(self):
        pc = dict(kind="pcounter", refresh=self.refresh, period=self.period, ranges=[(p, 0) for p in self.periods])
        f_refresh = self.refresh_opt if self.refresh_opt is not None else 0  # getOrElse(0.hour)
        return {
            self.top_target: dict(pc, scope=self.scope), self.bottom_target: dict(pc, scope=self.scope),
            self.top_global: dict(pc, scope=("global",)), self.bottom_global: dict(pc, scope=("global",)),
            self.item_field: dict(kind="scalar", scope=("item",), refresh=f_refresh),
            self.ranking_field: dict(kind="scalar", scope=("ranking",), refresh=f_refresh),
        }

    def _make_write(self, scope, ev, counter, global_counter):  # :220-234
        w = [("pinc", (scope, counter), ev["timestamp"], 1)]
        if self.weight is not None:
            w.append(("pinc", (("global",), global_counter), ev["timestamp"], 1))
        return w

    def writes(self, ev, store):  # :103-218
        e = ev["event"]
        if e == "ranking" and self.scope[0] == "ranking_field":
            v = fields_map(ev.get("fields", [])).get(self.scope[1])
            if is_strlist(v) and v:
                v = v[0]
            if isinstance(v, str):
                return [("put", (("ranking", ev["id"]), self.ranking_field), ev["timestamp"], v)]
            return []
        if e == "item" and self.scope[0] == "item_field":
            v = fields_map(ev.get("fields", [])).get(self.scope[1])
            if is_strlist(v) and v:
                v = v[0]
            if isinstance(v, str):
                return [("put", (("item", ev["item"]), self.item_field), ev["timestamp"], v)]
            return []
        if e == "interaction":
            if self.scope == ("item",):
                if ev["type"] == self.top:
                    return self._make_write(("item", ev["item"]), ev, self.top_target, self.top_global)
                if ev["type"] == self.bottom:
                    return self._make_write(("item", ev["item"]), ev, self.bottom_target, self.bottom_global)
                return []
            if self.scope[0] == "ranking_field":
                if ev.get("ranking") is None:
                    raise RuntimeError("got interaction event grouped by ranking field, but without ranking id")
                fv = store.scalars.get((("ranking", ev["ranking"]), self.ranking_field))
                if not isinstance(fv, str):
                    return []
                sc = ("irf", self.scope[1], fv, ev["item"])
            else:
                fv = store.scalars.get((("item", ev["item"]), self.item_field))
                if not isinstance(fv, str):
                    return []
                sc = ("field", self.scope[1], fv)
            if ev["type"] == self.top:
                return self._make_write(sc, ev, self.top_target, self.top_global)
            if ev["type"] == self.bottom:
                return self._make_write(sc, ev, self.bottom_target, self.bottom_global)
        return []

    def value(self, req, state, item):  # :290-356
        target = None
        if self.scope == ("item",):
            target = ("item", item["id"])
        elif self.scope[0] == "item_field":
            fv = state.get((("item", item["id"]), self.item_field))
            if fv is not None and fv[0] == "scalar" and isinstance(fv[1], str):
                target = ("field", self.scope[1], fv[1])
        elif self.scope[0] == "ranking_field":
            v = fields_map(req.get("fields", [])).get(self.scope[1])
            if isinstance(v, str):
                target = ("irf", self.scope[1], v, item["id"])
        missing = [NAN] * self.dim
        if target is None:
            return missing

        def pc(key):
            fv = state.get(key)
            return fv[1] if fv is not None and fv[0] == "pcounter" and len(fv[1]) == self.dim else None

        top, bottom = pc((target, self.top_target)), pc((target, self.bottom_target))
        if top is None or bottom is None:
            return missing
        if self.weight is None:
            return [fdiv(float(t), float(b)) for t, b in zip(top, bottom)]  # Long / Double
        tg, bg = pc((("global",), self.top_global)), pc((("global",), self.bottom_global))
        if tg is None or bg is None:
            return missing
        w = float(self.weight)
        # (w + top) / (w * (bottomGlobal / topGlobal) + bottom) with Long / Long INTEGER division (:343-350)
        return [fdiv(w + float(t), w * float(java_long_div(g_b, g_t)) + float(b))
                for t, b, g_t, g_b in zip(top, bottom, tg, bg)]


class InteractedWithFeature(BaseFeature):
    """S/feature/InteractedWithFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        self.interaction = c["interaction"]
        f = c["field"]
        fl = [f] if isinstance(f, str) else list(f)
        self.fields = []
        for x in fl:
            ev, fld = parse_field_name(x)
            if ev != "item":
                raise ValueError("can only be applied to item fields")
            self.fields.append(fld)
        # `fields` is a Scala immutable Map built with .toMap: insertion order up to 4 entries
        if len(self.fields) > 4:
            raise NotImplementedError("Scala HashMap iteration order for > 4 fields is not restated")
        self.scope = parse_scope_type(c["scope"])
        if self.scope not in (("user",), ("session",)):
            raise ValueError("can only be scoped to user/session")
        self.count = c.get("count") or 100
        self.duration = parse_duration_ms(c.get("duration") or "24h")
        self.dim = len(fl)
        self.interactions = self.name + "_interactions"

    def states(self):
        st = {self.interactions: dict(kind="blist", scope=self.scope, refresh=0, count=self.count, duration=self.duration)}
        for f in self.fields:
            st[f"{self.name}_{f}"] = dict(kind="scalar", scope=("item",), refresh=0)
        return st

    def _visitor_key(self, user, session):  # :125-129
        if self.scope == ("session",):
            return (("session", session), self.interactions) if session is not None else None
        return (("user", user), self.interactions) if user is not None else None

    def writes(self, ev, store):  # :68-97
        if ev["event"] == "item":
            out = []
            for n, v in ev.get("fields", []):
                if n in self.fields:
                    vals = [v] if isinstance(v, str) else list(v) if is_strlist(v) else []
                    out.append(("put", (("item", ev["item"]), f"{self.name}_{n}"), ev["timestamp"], vals))
            return out
        if ev["event"] == "interaction" and ev["type"] == self.interaction:
            key = self._visitor_key(ev.get("user"), ev.get("session"))
            return [("append", key, ev["timestamp"], ev["item"])] if key is not None else []
        return []

    def values(self, req, state, mode="online"):  # :133-164
        visitor_fields = {}
        vk = self._visitor_key(req.get("user"), req.get("session"))
        fv = state.get(vk) if vk is not None else None
        if fv is not None and fv[0] == "blist":
            for f in self.fields:
                hist = {}
                for it in fv[1]:
                    if not isinstance(it, str):
                        continue
                    sv = state.get((("item", it), f"{self.name}_{f}"))
                    if sv is not None and sv[0] == "scalar" and is_strlist(sv[1]):
                        for tag in sv[1]:
                            hist[tag] = hist.get(tag, 0) + 1
                visitor_fields[f] = hist
        out = []
        for it in req["items"]:
            row = []
            for f in self.fields:
                hist = visitor_fields.get(f, {})
                sv = state.get((("item", it["id"]), f"{self.name}_{f}"))
                tags = sv[1] if sv is not None and sv[0] == "scalar" and is_strlist(sv[1]) else []
                cnt = 0.0
                for t in tags:
                    cnt = cnt + hist.get(t, 0)
                row.append(cnt)
            out.append(row)
        return out


class RelevancyFeature(BaseFeature):
    """S/feature/RelevancyFeature.scala:36-51"""

    def __init__(self, c):
        self.name = c["name"]

    def values(self, req, state, mode="online"):
        out = []
        for it in req["items"]:
            first = next(((n, v) for n, v in it.get("fields", []) if n == "relevancy"), None)
            out.append([float(first[1])] if first is not None and is_num(first[1]) else [NAN])
        return out


class PositionFeature(BaseFeature):
    """S/feature/PositionFeature.scala:30-35"""

    def __init__(self, c):
        self.name = c["name"]
        self.position = int(c["position"])

    def values(self, req, state, mode="online"):
        if mode == "online":
            return [[float(self.position)] for _ in req["items"]]
        return [[float(i)] for i, _ in enumerate(req["items"])]


def percentile50_legacy(data):
    """org.apache.commons.math3.stat.descriptive.rank.Percentile.evaluate(50.0) with the default
    LEGACY estimation and NaNStrategy.REMOVED (used at S/feature/DiversityFeature.scala:118-129)."""
    n = len(data)
    if n == 0:
        return NAN
    if n == 1:
        return data[0]
    work = sorted(v for v in data if v == v)
    n = len(work)
    if n == 0:
        return NAN
    pos = 0.5 * (n + 1)
    fpos = math.floor(pos)
    ip = int(fpos)
    dif = pos - fpos
    if pos < 1:
        return work[0]
    if pos >= n:
        return work[n - 1]
    lower, upper = work[ip - 1], work[ip]
    return lower + dif * (upper - lower)


class DiversityFeature(BaseFeature):
    """S/feature/DiversityFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        ev, self.field = parse_field_name(c["source"])
        if ev != "item":
            raise ValueError("diversity feature can only accept item fields")
        self.top = c.get("top", 20) if c.get("top") is not None else 20  # YAML default (:164)

    def states(self):
        return {self.name: dict(kind="scalar", scope=("item",), refresh=0)}

    def writes(self, ev, store):  # :42-62
        if ev["event"] != "item":
            return []
        for n, v in ev.get("fields", []):
            if n == self.field:
                if is_num(v):
                    return [("put", (("item", ev["item"]), self.name), ev["timestamp"], float(v))]
                if isinstance(v, str) or is_strlist(v):
                    return [("put", (("item", ev["item"]), self.name), ev["timestamp"], v)]
                return []
        return []

    def values(self, req, state, mode="online"):  # :67-130
        fvs = []
        for it in req["items"]:
            fv = state.get((("item", it["id"]), self.name))
            if fv is not None and fv[0] == "scalar":
                fvs.append((it["id"], fv[1]))
        empty = [[0.0] for _ in req["items"]]
        if not fvs:
            return empty
        head = fvs[0][1]
        if isinstance(head, str) or is_strlist(head):
            sv = [(i, [v] if isinstance(v, str) else list(v)) for i, v in fvs if isinstance(v, str) or is_strlist(v)]
            fmap = dict(sv)
            counts = {}
            for _, tags in sv[: self.top]:
                for t in tags:
                    counts[t] = counts.get(t, 0) + 1
            total = 0.0
            for c in counts.values():
                total = total + c
            out = []
            for it in req["items"]:
                if it["id"] not in fmap:
                    out.append([NAN])
                else:
                    s = 0.0
                    for t in fmap[it["id"]]:
                        s = s + counts.get(t, 0)
                    out.append([fdiv(s, total)])
            return out
        if is_num(head):
            dv = [(i, float(v)) for i, v in fvs if is_num(v)]
            fmap = dict(dv)
            median = percentile50_legacy([v for _, v in dv][: self.top])
            return [[fmap[it["id"]] - median] if it["id"] in fmap else [NAN] for it in req["items"]]
        return empty


def cosine_dist(query_f32, item_f64) -> float:
    """CosineDistance.dist (S/ml/onnx/distance/DistanceFunction.scala:13-27): sequential double
    sums; query(i)*query(i) is a FLOAT product."""
    q = np.asarray(query_f32, dtype=np.float32)
    e = np.asarray(item_f64, dtype=np.float64)
    top = np.float64(0.0)
    a = np.float64(0.0)
    b = np.float64(0.0)
    for i in range(len(q)):
        top = top + np.float64(q[i]) * e[i]
        a = a + np.float64(np.float32(q[i] * q[i]))
        b = b + e[i] * e[i]
    with np.errstate(all="ignore"):
        return float(top / (np.sqrt(a) * np.sqrt(b)))


def normalize_scale(kind: str, values: list) -> list:
    """Normalize.scale (S/ml/onnx/Normalize.scala:13-46)."""
    if kind == "noop":
        return list(values)
    if kind == "linear":
        sc = [v for v in values if v == v]
        if not sc:
            return list(values)
        mn, mx = min(sc), max(sc)
        return [fdiv(v - mn, mx - mn) for v in values]
    if kind == "position":
        size = float(len(values))

        def key(v):  # java.lang.Double.compare: NaN greatest
            return (1, 0.0) if v != v else (0, v)

        order = sorted(range(len(values)), key=lambda i: key(values[i]))  # stable
        out = list(values)
        for si, oi in enumerate(order):
            if values[oi] == values[oi]:
                out[oi] = si / size
        return out
    raise ValueError(kind)


class FieldMatchBiencoderFeature(BaseFeature):
    """S/feature/FieldMatchBiencoderFeature.scala:80-109 with the query embedding supplied
    (rankingCache hit); the ONNX forward is out of scope."""

    def __init__(self, c):
        self.name = c["name"]
        self.ranking_field = parse_field_name(c["rankingField"])[1]
        self.item_field = parse_field_name(c["itemField"])[1]
        self.vdim = int(c["method"]["dim"])
        self.norm = c.get("norm", "noop") or "noop"

    def states(self):
        return {self.name: dict(kind="scalar", scope=("item",), refresh=0)}

    def values(self, req, state, mode="online", query_embedding=None):
        # the query STRING comes from the ranking field (StringField, or StringListField joined by " ");
        # without it the feature is missing whatever the caches hold (:84-88)
        qf = fields_map(req.get("fields", [])).get(self.ranking_field)
        has_query = isinstance(qf, str) or is_strlist(qf)
        q = (req.get("embeddings") or {}).get(self.name) if query_embedding is None else query_embedding
        if q is None or not has_query:
            return [[NAN] for _ in req["items"]]
        raw = []
        for it in req["items"]:
            fv = state.get((("item", it["id"]), self.name))
            if fv is not None and fv[0] == "scalar" and isinstance(fv[1], (list, np.ndarray)) and not is_strlist(list(fv[1])[:1] or [0]):
                raw.append(cosine_dist(q, fv[1]))
            else:
                raw.append(NAN)
        return [[v] for v in normalize_scale(self.norm, raw)]


class FieldMatchTokensFeature(BaseFeature):
    """S/feature/FieldMatchFeature.scala:30-93 with method ngram / term / bm25.

    Tokenisation is Lucene's (S/util/TextAnalyzer.scala) and therefore NOT restated, except for the
    `whitespace` analyzer (split on whitespace only): NgramMatcher.tokenize / TermMatcher.tokenize
    (S/feature/matcher/NgramMatcher.scala:9-30, TermMatcher.scala:7-12) then are pure string code.  For any
    other language the tokens come with the event / request: event["tokens"][feature], request["tokens"][feature]
    (what matcher.tokenize(field) returned: sorted, unique).
    Scoring: FieldMatcher.score (matcher/FieldMatcher.scala:15-49, Jaccard over two sorted unique arrays) and
    BM25Matcher.score (matcher/BM25Matcher.scala:19-33)."""

    K1, B = 1.2, 0.75

    def __init__(self, c):
        self.name = c["name"]
        self.ranking_field = parse_field_name(c["rankingField"])[1]
        self.item_field = parse_field_name(c["itemField"])[1]
        self.method = dict(c["method"])
        self.state_name = self.name + "_" + self.item_field  # conf.name :32

    def states(self):
        return {self.state_name: dict(kind="scalar", scope=("item",), refresh=0)}

    def tokenize(self, text, given=None):
        if given is not None:
            return list(given)
        if self.method.get("language") != "whitespace":
            raise NotImplementedError(f"{self.name}: analyzer {self.method.get('language')!r} is Lucene's; pass tokens")
        terms = text.split()
        if self.method["type"] == "ngram":  # NgramMatcher.tokenize
            n = int(self.method["n"])
            terms = [t[j:j + n] for t in terms for j in range(0, len(t) - n + 1)]
        return sorted(set(terms))  # FieldMatcher.unique: sort + dedupe

    def writes(self, ev, store):  # :39-54
        if ev.get("event") != "item":
            return []
        key = (("item", ev["item"]), self.state_name)
        for n, v in ev.get("fields", []):
            if n == self.item_field:
                given = (ev.get("tokens") or {}).get(self.name)
                if isinstance(v, str):
                    return [("put", key, ev["timestamp"], self.tokenize(v, given))]
                if is_strlist(v):
                    return [("put", key, ev["timestamp"], self.tokenize(" ".join(v), given))]
                return []
        return []

    def score(self, query, doc):
        if self.method["type"] == "bm25":  # BM25Matcher.score: no early exit, tf of a unique-token doc is 0/1
            m = self.method
            total = 0.0
            freq = {}
            for d in doc:
                freq[d] = freq.get(d, 0) + 1
            for term in query:
                tf = freq.get(term, 0)
                gtf = m.get("termfreq", {}).get(term, 0)
                idf = math.log(1.0 + (m["docs"] - gtf + 0.5) / (gtf + 0.5))
                total += idf * (tf * (self.K1 + 1.0)) / (tf + self.K1 * (1.0 - self.B + self.B * (len(doc) / m["avgdl"])))
            return total
        if not query or not doc:
            return 0.0
        inter = len(set(query) & set(doc))
        return float(inter) / float(len(set(query) | set(doc)))

    def values(self, req, state, mode="online"):  # :60-93
        qf = fields_map(req.get("fields", [])).get(self.ranking_field)
        if not isinstance(qf, str):
            return [[0.0] for _ in req["items"]]
        q = self.tokenize(qf, (req.get("tokens") or {}).get(self.name))
        out = []
        for it in req["items"]:
            fv = state.get((("item", it["id"]), self.state_name))
            if fv is not None and fv[0] == "scalar" and (is_strlist(fv[1]) or (isinstance(fv[1], (list, tuple)) and len(fv[1]) == 0)):
                out.append([self.score(q, list(fv[1]))])
            else:
                out.append([0.0])
        return out


class BooleanFeature(BaseFeature):
    """S/feature/BooleanFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        self.scope = parse_scope_type(c["scope"])
        self.field = parse_field_name(c.get("field", c.get("source")))[1]

    def states(self):
        return {self.name: dict(kind="scalar", scope=self.scope, refresh=0)}

    def writes(self, ev, store):
        key = self.write_key(ev, self.scope, self.name)
        if key is None:
            return []
        for n, v in ev.get("fields", []):
            if n == self.field:
                return [("put", key, ev["timestamp"], bool(v))] if isinstance(v, bool) else []
        return []

    def value(self, req, state, item):  # :47-62
        key = self.read_key(req, self.scope, self.name, item["id"])
        fv = state.get(key) if key is not None else None
        if fv is not None and fv[0] == "scalar" and isinstance(fv[1], bool):
            return [1.0 if fv[1] else 0.0]
        return [NAN]


class NumVectorFeature(BaseFeature):
    """S/feature/NumVectorFeature.scala (reducers are applied on the WRITE side; `random` is not restated)"""

    def __init__(self, c):
        self.name = c["name"]
        self.scope = parse_scope_type(c["scope"])
        self.field = parse_field_name(c["source"])[1]
        self.reducers = list(c["reduce"]) if c.get("reduce") else ["min", "max", "size", "avg"]
        self.dim = sum(int(r[6:]) if re.fullmatch(r"vector[0-9]+", r) else 1 for r in self.reducers)

    def states(self):
        return {self.name: dict(kind="scalar", scope=self.scope, refresh=0)}

    @staticmethod
    def reduce(name, v):  # :77-168
        if re.fullmatch(r"vector[0-9]+", name):
            n = int(name[6:])
            return (list(v) + [0.0] * n)[:n]  # Arrays.copyOfRange pads with zeros
        if name == "size":
            return [float(len(v))]
        if name == "sum":
            s = 0.0
            for x in v:
                s += x
            return [s]
        if name == "euclidean_distance":
            s = 0.0
            for x in v:
                s += x * x
            return [math.sqrt(s)]
        if not v:
            return [0.0]
        if name == "first":
            return [v[0]]
        if name == "last":
            return [v[-1]]
        if name == "min":
            m = 1.7976931348623157e308  # Double.MaxValue
            for x in v:
                if x < m:
                    m = x
            return [m]
        if name == "max":
            m = -1.7976931348623157e308  # scala.Double.MinValue = -Double.MaxValue
            for x in v:
                if x > m:
                    m = x
            return [m]
        if name == "avg":
            s = 0.0
            for x in v:
                s += x
            return [s / len(v)]
        raise NotImplementedError(name)

    def writes(self, ev, store):
        key = self.write_key(ev, self.scope, self.name)
        if key is None:
            return []
        for n, v in ev.get("fields", []):
            if n == self.field:
                vals = [float(v)] if is_num(v) else [float(x) for x in v] if isinstance(v, list) and all(is_num(x) for x in v) else None
                if vals is None:
                    return []
                out = []
                for r in self.reducers:
                    out.extend(self.reduce(r, vals))
                return [("put", key, ev["timestamp"], out)]
        return []

    def value(self, req, state, item):  # :55-70
        key = self.read_key(req, self.scope, self.name, item["id"])
        fv = state.get(key) if key is not None else None
        if fv is not None and fv[0] == "scalar" and isinstance(fv[1], (list, np.ndarray)) and (len(fv[1]) == 0 or is_num(fv[1][0])):
            return [float(x) for x in fv[1]]
        return [NAN] * self.dim


def parse_iso_datetime(s: str):
    """ZonedDateTime.parse(value, ISO_DATE_TIME): offset date-time with an optional [zone] suffix."""
    import datetime as _dt

    s2 = re.sub(r"\[.*\]$", "", s)
    if s2.endswith("Z"):
        s2 = s2[:-1] + "+00:00"
    try:
        d = _dt.datetime.fromisoformat(s2)
    except ValueError:
        return None
    return d if d.tzinfo is not None else None


class ItemAgeFeature(BaseFeature):
    """S/feature/ItemAgeFeature.scala"""

    def __init__(self, c):
        self.name = c["name"]
        self.field = parse_field_name(c["source"])[1]
        self.scope = ("item",)

    def states(self):
        return {self.name: dict(kind="scalar", scope=("item",), refresh=0)}

    def writes(self, ev, store):  # :33-66
        key = self.write_key(ev, ("item",), self.name)
        if key is None:
            return []
        if self.field == "timestamp":
            v = ev["timestamp"] / 1000.0
        else:
            v = next((v for n, v in ev.get("fields", []) if n == self.field), None)
        if is_num(v):
            out = float(v)
        elif isinstance(v, str):
            d = parse_iso_datetime(v)
            if d is not None:
                out = float(int(d.timestamp()))
            else:
                try:
                    out = float(v)
                except ValueError:
                    return []
        else:
            return []
        return [("put", key, ev["timestamp"], out)]

    def value(self, req, state, item):  # :74-86
        fv = state.get((("item", item["id"]), self.name))
        if fv is not None and fv[0] == "scalar" and is_num(fv[1]):
            ms = float(fv[1]) * 1000.0
            upd = 0 if ms != ms else int(math.floor(ms + 0.0)) + (1 if (ms - math.floor(ms)) >= 0.5 else 0)  # Math.round
            return [float(abs(int(req["timestamp"]) - upd) // 1000)]
        return [NAN]


class LocalDateTimeFeature(BaseFeature):
    """S/feature/LocalDateTimeFeature.scala (a RankingFeature: one value per request)"""

    def __init__(self, c):
        self.name = c["name"]
        ev, self.field = parse_field_name(c["source"])
        if ev != "ranking":
            raise ValueError("can only work with ranking event fields")
        self.parse = c["parse"]

    @staticmethod
    def map_datetime(parse, d) -> float:  # :44-80
        if parse == "time_of_day":
            return (d.hour * 3600 + d.minute * 60 + d.second) / 3600.0
        if parse == "day_of_week":
            return float(d.isoweekday())
        if parse == "month_of_year":
            return float(d.month)
        if parse == "year":
            return float(d.year)
        if parse == "second":
            return float(int(d.timestamp()))
        raise ValueError(f"parsing method {parse} is not supported")

    def request_value(self, req) -> float:
        import datetime as _dt

        if self.field == "timestamp":
            d = _dt.datetime.fromtimestamp(int(req["timestamp"]) // 1000, tz=_dt.timezone.utc)
            return self.map_datetime(self.parse, d)
        v = fields_map(req.get("fields", [])).get(self.field)
        if isinstance(v, str):
            d = parse_iso_datetime(v)
            if d is not None:
                return self.map_datetime(self.parse, d)
        return NAN

    def values(self, req, state, mode="online"):
        v = self.request_value(req)
        return [[v] for _ in req["items"]]


FEATURE_TYPES = {
    "number": NumberFeature, "word_count": WordCountFeature, "string": StringFeature, "referer": RefererFeature,
    "interaction_count": InteractionCountFeature, "window_count": WindowInteractionCountFeature,
    "rate": RateFeature, "interacted_with": InteractedWithFeature, "relevancy": RelevancyFeature,
    "position": PositionFeature, "diversity": DiversityFeature, "boolean": BooleanFeature,
    "vector": NumVectorFeature, "item_age": ItemAgeFeature, "local_time": LocalDateTimeFeature,
}


def make_feature(conf: dict) -> BaseFeature:
    t = conf["type"]
    if t == "field_match":
        if conf["method"]["type"] in ("ngram", "term", "bm25"):
            return FieldMatchTokensFeature(conf)
        if conf["method"]["type"] != "bi-encoder":
            raise NotImplementedError(conf["method"]["type"])
        return FieldMatchBiencoderFeature(conf)
    if t not in FEATURE_TYPES:
        raise NotImplementedError(f"feature type {t} is not supported")
    return FEATURE_TYPES[t](conf)


# --------------------------------------------------------------------------- mapping / flow / query

class FeatureMapping:
    """FeatureMapping.fromFeatureSchema + makeDatasetDescriptor (S/FeatureMapping.scala:56-99)."""

    def __init__(self, feature_confs: list, model_features: list):
        self.features = [make_feature(c) for c in feature_confs]
        self.model_features = list(model_features)
        self.configs = {}
        for f in self.features:
            self.configs.update(f.states())
        # DatasetDescriptor: model feature order, offsets = running sum of dims
        self.offsets, self.dim = {}, 0
        for name in self.model_features:
            f = next((x for x in self.features if x.name == name), None)
            if f is None:
                continue
            self.offsets[name] = (self.dim, f.dim)
            self.dim += f.dim


class FeatureValueFlow:
    """FeatureValueFlow.process (S/flow/FeatureValueFlow.scala:24-92) over MemPersistence."""

    def __init__(self, mapping: FeatureMapping, always_refresh: bool = False):
        # always_refresh mirrors the reference test harness, whose `updated` cache is
        # Scaffeine().maximumSize(0) and therefore never remembers a key
        # (src/test/scala/ai/metarank/feature/FeatureTest.scala:27)
        self.mapping = mapping
        self.store = MemState()
        self.updated = {}
        self.always_refresh = always_refresh
        self.write_log = []  # every committed write, in order (tests replay them through mr_state_apply_writes)

    def process(self, events) -> dict:
        out = {}
        for ev in events:
            writes = []
            for f in self.mapping.features:
                writes.extend(f.writes(ev, self.store))
            for w in writes:
                self.store.put(w, self.mapping.configs)
            self.write_log.extend(writes)
            for w in writes:
                key, ts = w[1], w[2]
                last = None if self.always_refresh else self.updated.get(key)
                if last is None:
                    self.updated[key] = ts
                    refresh = True
                else:
                    refresh = abs(ts - last) >= self.mapping.configs[key[1]]["refresh"]
                if refresh:
                    v = self.store.compute_value(w, self.mapping.configs)
                    if v is not None:
                        out[key] = v
        return out


def item_values(mapping: FeatureMapping, req: dict, state: dict, mode="online") -> list:
    """ItemValue.fromState restricted to the model's features: per item, {feature: [values]}."""
    feats = [f for f in mapping.features if f.name in mapping.model_features]
    per_feature = {}
    for f in feats:
        vals = f.values(req, state, mode)
        if len(vals) != len(req["items"]):
            raise RuntimeError(f"for {f.name} dim mismatch: there should be {len(req['items'])} per-document values")
        for v in vals:
            if len(v) != f.dim:
                raise RuntimeError(f"for {f.name} dim mismatch: {f.dim} != {len(v)}")
        per_feature[f.name] = vals
    return [{n: per_feature[n][i] for n in per_feature} for i in range(len(req["items"]))]


def dense_matrix(mapping: FeatureMapping, req: dict, state: dict, mode="online") -> np.ndarray:
    """ClickthroughQuery.collectFeatureValues: zero-initialised row, values scattered at
    DatasetDescriptor.offsets -> row-major f64[N x dim] (ltrlib Query.values)."""
    ivs = item_values(mapping, req, state, mode)
    out = np.zeros((len(ivs), mapping.dim), dtype=np.float64)
    for i, iv in enumerate(ivs):
        for name, vals in iv.items():
            if name in mapping.offsets:
                o, d = mapping.offsets[name]
                out[i, o:o + d] = vals
    return out
