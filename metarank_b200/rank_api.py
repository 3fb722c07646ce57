"""POST /rank/<model>?explain= — the reference's HTTP surface for the hot path, as a host-side
mirror over the C ABI (reference S/api/routes/RankApi.scala:25-41,58-84).

In a deployment this layer stays on the JVM (http4s + circe); it exists here so the whole
request -> response path can be exercised and tested the way T/main/api/RankApiTest.scala does
(in-process `routes.apply(request)`, no socket).  `serve()` adds a thread-per-connection socket
front end from the standard library for manual poking; it is not a production server.

JSON shapes follow the reference's codecs:
  RankingEvent   S/model/Event.scala:44-99  (id, timestamp as long | numeric string | ISO date-time,
                 user?, session?, fields?[{name, value}], items[{id, relevancy?, fields?, label?}];
                 `relevancy` is sugar for a leading NumberField("relevancy", r), :84-93)
  Field          S/model/Field.scala:36-58  (string | bool | number | string[] | number[])
  RankResponse   S/api/routes/RankApi.scala:58-84, printed by JsonChunk (keys sorted, nulls dropped,
                 S/api/JsonChunk.scala:8-17); features = {name: double | [double] | "cat@index"}, NaN -> null
                 (S/model/MValue.scala:70-76)
"""
from __future__ import annotations

import datetime as _dt
import json
import math
import re
import time
from urllib.parse import parse_qs, urlparse

from .features import FeatureMapping, Ranker


class DecodingFailure(ValueError):
    pass


class ModelError(Exception):
    """RankApi.ModelError: unknown / untrained model (S/ml/Ranker.scala:36-43,88-94)."""


def decode_timestamp(v) -> int:
    if isinstance(v, bool):
        raise DecodingFailure("cannot decode timestamp")
    if isinstance(v, int):
        return v
    if isinstance(v, float) and v == int(v):
        return int(v)
    if isinstance(v, str):
        s = v.strip()
        if re.fullmatch(r"-?\d+", s):
            return int(s)
        s2 = re.sub(r"\[.*\]$", "", s)
        if s2.endswith("Z"):
            s2 = s2[:-1] + "+00:00"
        try:
            d = _dt.datetime.fromisoformat(s2)
        except ValueError:
            d = None
        if d is not None and d.tzinfo is not None:  # Instant.toEpochMilli: exact, truncated to the millisecond
            return (d - _dt.datetime(1970, 1, 1, tzinfo=_dt.timezone.utc)) // _dt.timedelta(milliseconds=1)
    raise DecodingFailure(f"cannot decode timestamp {v!r}")


def decode_field(o) -> tuple:
    if not isinstance(o, dict) or not isinstance(o.get("name"), str):
        raise DecodingFailure("field needs a string 'name'")
    name = o["name"]
    if "value" not in o:
        raise DecodingFailure("field value not found")
    v = o["value"]
    if v is None:
        raise DecodingFailure(f"null value in field {name}")
    if isinstance(v, (bool, str)):
        return (name, v)
    if isinstance(v, (int, float)):
        return (name, float(v))
    if isinstance(v, list):
        if all(isinstance(x, str) for x in v):
            return (name, list(v))
        if all(isinstance(x, (int, float)) and not isinstance(x, bool) for x in v):
            return (name, [float(x) for x in v])
        raise DecodingFailure(f"cannot decode field {name}: got list of {v}")
    raise DecodingFailure(f"cannot decode field {name}: got object {v}")


def decode_ranking_event(text: str) -> dict:
    """circe decode[RankingEvent] (the `event` discriminator is optional on this endpoint)."""
    def _no_constants(name):  # NaN / Infinity are not JSON: circe's parser rejects them (T/model/FieldTest.scala:15-17)
        raise DecodingFailure(f"{name} is not valid JSON")

    try:
        o = json.loads(text, parse_constant=_no_constants)
    except json.JSONDecodeError as e:
        raise DecodingFailure(str(e)) from e
    if not isinstance(o, dict):
        raise DecodingFailure("ranking event must be a JSON object")
    for k in ("id", "timestamp", "items"):
        if k not in o:
            raise DecodingFailure(f"required field '{k}' missing in JSON")
    if not isinstance(o["items"], list) or not o["items"]:
        raise DecodingFailure("items must be a non-empty list")  # NonEmptyList
    items = []
    for it in o["items"]:
        if not isinstance(it, dict) or not isinstance(it.get("id"), str):
            raise DecodingFailure("item needs a string 'id'")
        fields = [decode_field(f) for f in (it.get("fields") or [])]
        rel = it.get("relevancy")
        if rel is not None:
            if isinstance(rel, bool) or not isinstance(rel, (int, float)):
                raise DecodingFailure("relevancy must be a number")
            fields = [("relevancy", float(rel))] + fields
        items.append(dict(id=it["id"], fields=fields, label=it.get("label")))
    for k in ("user", "session"):
        if o.get(k) is not None and not isinstance(o[k], str):
            raise DecodingFailure(f"'{k}' must be a string")
    out = dict(event="ranking", id=str(o["id"]), timestamp=decode_timestamp(o["timestamp"]),
               user=o.get("user"), session=o.get("session"),
               fields=[decode_field(f) for f in (o.get("fields") or [])], items=items)
    for k in ("embeddings", "tokens"):  # extensions: what the caller's models / analyzers produced for this request
        if isinstance(o.get(k), dict):
            out[k] = o[k]
    return out


def _num(x: float):
    return None if (isinstance(x, float) and (math.isnan(x) or math.isinf(x))) else x  # Json.fromDoubleOrNull


def encode_response(resp: dict) -> str:
    """JsonChunk: sorted keys, two-space indent, null values dropped."""
    def drop(o):
        if isinstance(o, dict):
            return {k: drop(v) for k, v in o.items() if v is not None}
        if isinstance(o, list):
            return [drop(v) for v in o]  # nulls inside arrays are kept (vector values)
        return o
    return json.dumps(drop(resp), sort_keys=True, indent=2)


class RankApi:
    """RankApi(ranker).routes — `models` maps a model name to (FeatureMapping, DeviceState, booster)."""

    def __init__(self, models: dict):
        self.models = models
        self.requests = {}  # metarank_rank_requests{model} counter (S/util/analytics/Metrics.scala:5-21)

    def _mvalues(self, mapping: FeatureMapping, row) -> dict:
        out = {}
        for name in mapping.model_features:
            od = mapping.offset(name)
            conf = mapping.by_name.get(name)
            if od is None or conf is None:
                continue
            o, d = od
            vals = [float(x) for x in row[o:o + d]]
            if conf["type"] == "referer":
                idx = int(vals[0])
                out[name] = f"{('unknown', 'search', 'internal', 'social', 'email', 'paid')[idx] if 0 <= idx <= 5 else 'unknown'}@{idx}"
            elif conf["type"] == "string" and conf.get("encode") == "index":
                idx = int(vals[0])
                cat = conf["values"][idx - 1] if 1 <= idx <= len(conf["values"]) else "nil"
                out[name] = f"{cat}@{idx}"  # CategoryValue
            elif d == 1 and conf["type"] not in ("rate", "window_count", "interacted_with", "vector") \
                    and not (conf["type"] == "string" and conf.get("encode", "onehot") == "onehot"):
                out[name] = _num(vals[0])  # SingleValue
            else:
                out[name] = [_num(v) for v in vals]  # VectorValue
        return out

    def rerank(self, request: dict, model: str, explain: bool) -> dict:
        """Ranker.rerank (S/ml/Ranker.scala:27-83) -> RankResponse as a plain dict."""
        start = time.time()
        if model not in self.models:
            raise ModelError(f"model {model} is not configured")
        mapping, state, booster = self.models[model]
        resp = Ranker(mapping, state).rerank([request], booster, explain=explain)[0]
        items = []
        for e in resp["items"]:
            it = {"item": e["item"], "score": e["score"], "features": None}
            if explain:
                it["features"] = self._mvalues(mapping, e["features"])
            items.append(it)
        # explain=true also dumps every loaded FeatureValue grouped by scope in the reference; the state
        # lives in HBM here, so the groups are returned empty (SURVEY.md appendix D allows dropping it)
        st = {"session": [], "user": [], "global": [], "item": []} if explain else None
        return {"state": st, "items": items, "took": int((time.time() - start) * 1000)}

    def routes(self, method: str, url: str, body: str):
        """(status, content_type, body) for one request; anything but POST /rank/<model> is 404."""
        u = urlparse(url)
        m = re.fullmatch(r"/rank/([^/]+)", u.path)
        if method != "POST" or not m:
            return 404, "text/plain", "Not found"
        model = m.group(1)
        q = parse_qs(u.query).get("explain", [None])[0]
        if q not in (None, "true", "false"):
            return 400, "text/plain", "explain must be a boolean"
        self.requests[model] = self.requests.get(model, 0) + 1
        try:
            req = decode_ranking_event(body)
            return 200, "application/json", encode_response(self.rerank(req, model, q == "true"))
        except Exception as e:  # ErrorAction.httpRoutes logs and answers 500 (S/main/command/Serve.scala:101-103)
            return 500, "text/plain", f"{type(e).__name__}: {e}"


def serve(api: RankApi, host: str = "127.0.0.1", port: int = 8080):
    """Blocking stdlib HTTP front end for manual testing."""
    from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

    class H(BaseHTTPRequestHandler):
        def do_POST(self):  # noqa: N802
            n = int(self.headers.get("Content-Length", "0"))
            status, ctype, body = api.routes("POST", self.path, self.rfile.read(n).decode("utf-8"))
            data = body.encode("utf-8")
            self.send_response(status)
            self.send_header("Content-Type", ctype)
            self.send_header("Content-Length", str(len(data)))
            self.end_headers()
            self.wfile.write(data)

        def log_message(self, *a):
            pass

    ThreadingHTTPServer((host, port), H).serve_forever()
