"""Native request decoder (metarank_b200/csrc/request_codec.cpp, host only) against the Python
shim, which the device tests hold to the oracle: JSON body -> decode_ranking_event -> FeatureMapping.pack_requests
must give the same arrays as mr_requests_decode, value for value, on configs that use every request-side input."""
import json

import numpy as np
import pytest

from metarank_b200 import _capi, features as F
from metarank_b200.rank_api import DecodingFailure, decode_ranking_event

FEATS = [
    dict(name="price", type="number", scope="item", source="item.price"),
    dict(name="qlen", type="word_count", scope="ranking", source="ranking.query"),
    dict(name="budget", type="number", scope="ranking", source="ranking.budget"),
    dict(name="platform", type="string", scope="ranking", source="ranking.platform", encode="onehot", values=["ios", "android", "web"]),
    dict(name="country", type="string", scope="ranking", source="ranking.country", encode="index", values=["de", "fr", "us"]),
    dict(name="color", type="string", scope="item", source="item.color", encode="index", values=["red", "green", "blue"]),
    dict(name="size", type="string", scope="item", source="item.size", encode="onehot", values=["s", "m", "l", "xl"]),
    dict(name="ctr", type="rate", top="click", bottom="impression", scope="ranking.query", bucket="24h", periods=[7, 30]),
    dict(name="rel", type="relevancy"),
    dict(name="age", type="item_age", source="item.updated_at"),
    dict(name="tod", type="local_time", source="ranking.timestamp", parse="time_of_day"),
    dict(name="dow", type="local_time", source="ranking.local_ts", parse="day_of_week"),
    dict(name="moy", type="local_time", source="ranking.local_ts", parse="month_of_year"),
    dict(name="yr", type="local_time", source="ranking.local_ts", parse="year"),
    dict(name="sec", type="local_time", source="ranking.local_ts", parse="second"),
    dict(name="sim", type="field_match", rankingField="ranking.query", itemField="item.title", distance="cos",
         method=dict(type="bi-encoder", dim=4)),
    dict(name="ng", type="field_match", rankingField="ranking.query", itemField="item.title",
         method=dict(type="ngram", n=3, language="whitespace")),
    dict(name="tm", type="field_match", rankingField="ranking.query", itemField="item.title",
         method=dict(type="term", language="en")),
    dict(name="bm", type="field_match", rankingField="ranking.query", itemField="item.title",
         method=dict(type="bm25", language="whitespace", docs=40, avgdl=3.5, termfreq={"red": 7, "shoes": 3, "кеды": 1})),
]
MODEL = [f["name"] for f in FEATS]
WORDS = ["red", "shoes", "кеды", "x", "läuft", "😀ok", "blue-ish", "a"]


def _random_body(rng, n_events):
    def field_value(kind):
        if kind == "str":
            return " ".join(str(w) for w in rng.choice(WORDS, int(rng.integers(0, 4))))
        if kind == "num":
            return float(rng.integers(-5, 50)) if rng.random() < 0.5 else int(rng.integers(0, 9))
        if kind == "bool":
            return bool(rng.random() < 0.5)
        if kind == "strlist":
            return [str(w) for w in rng.choice(["ios", "web", "red", "blue", "xl", "m"], int(rng.integers(0, 3)))]
        return [float(x) for x in rng.integers(0, 5, int(rng.integers(1, 3)))]

    events = []
    for e in range(n_events):
        fields = []
        for name, kinds in [("query", ["str", "str", "strlist", "num"]), ("budget", ["num", "num", "str", "bool"]),
                            ("platform", ["str", "strlist", "num"]), ("country", ["str", "strlist"]),
                            ("local_ts", ["iso", "iso", "str"])]:
            for _ in range(int(rng.integers(0, 3))):  # absent, once, or duplicated (first/last-wins rules)
                k = str(rng.choice(kinds))
                if name in ("platform", "country") and k == "str":
                    v = str(rng.choice(["ios", "web", "de", "us", "zz"]))
                elif k == "iso":
                    v = str(rng.choice(["2024-03-17T23:59:58+01:00", "2023-12-31T22:00:00.250-05:30[America/Nowhere]",
                                        "1999-01-01T00:00Z", "2024-02-29T12:34:56,7Z", "2024-03-17T10:00:00", "17.03.2024"]))
                else:
                    v = field_value(k)
                fields.append(dict(name=name, value=v))
        items = []
        for j in range(int(rng.integers(1, 6))):
            it = dict(id=str(rng.choice(["p1", "p2", "товар", "p😀", f"i{j}"])))
            fl = []
            if rng.random() < 0.4:
                it["relevancy"] = float(rng.integers(0, 4))
            for name, kinds in [("price", ["num", "str"]), ("color", ["str", "strlist", "num"]), ("size", ["strlist", "str"]),
                                ("relevancy", ["num", "str"])]:
                if rng.random() < 0.35:
                    k = str(rng.choice(kinds))
                    v = str(rng.choice(["red", "blue", "xl", "zz"])) if k == "str" else field_value(k)
                    fl.append(dict(name=name, value=v))
            if fl or rng.random() < 0.3:
                it["fields"] = fl
            if rng.random() < 0.2:
                it["label"] = "x"
            items.append(it)
        ts = rng.choice(["int", "str", "iso", "float"])
        ev = dict(id=f"r{e}", items=items, fields=fields,
                  timestamp={"int": 1710716400123, "str": " 1710716400123 ", "iso": "2024-03-17T23:00:00.123+00:00",
                             "float": 1710716400000.0}[str(ts)])
        if rng.random() < 0.7:
            ev["user"] = str(rng.choice(["alice", "боб"]))
        if rng.random() < 0.7:
            ev["session"] = "s1"
        if rng.random() < 0.3:
            ev["user"] = None
        if rng.random() < 0.7:
            ev["embeddings"] = {"sim": [float(x) for x in rng.normal(size=4)]}
        ev["tokens"] = {"tm": sorted({str(w) for w in rng.choice(WORDS, int(rng.integers(0, 4)))})}
        events.append(ev)
    return events


def _python_arrays(fm, events):
    reqs = [decode_ranking_event(json.dumps(e)) for e in events]
    return fm.pack_requests(reqs), reqs


def _same(a, b, what):
    if a is None or b is None:
        assert a is None and b is None, what
        return
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype.kind == "f":
        assert np.array_equal(a, b, equal_nan=True), (what, a, b)
    else:
        assert np.array_equal(a, b), (what, a, b)


@pytest.mark.parametrize("seed", range(8))
def test_native_decoder_packs_what_the_python_shim_packs(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    fm = F.FeatureMapping(None, FEATS, MODEL)
    events = _random_body(rng, int(rng.integers(1, 7)))
    want, reqs = _python_arrays(fm, events)
    body = json.dumps(events if len(events) > 1 or seed % 2 else events[0], ensure_ascii=bool(seed % 3))  # \\uXXXX escapes too
    dec = F.DecodedRequests(fm, body)
    got = dec.arrays()
    assert got["n_requests"] == want["n_requests"] and got["total_items"] == want["total_items"]
    for k in ["offsets", "ids", "users", "sessions", "req_f64", "req_u64", "req_vec", "req_vp", "item_f64", "tok_off"]:
        _same(got[k], want[k], k)
    n_tok = int(want["tok_off"][-1])
    _same(got["tok_hash"][:n_tok], want["tok_hash"][:n_tok], "tok_hash")
    _same(got["tok_w"][:n_tok], want["tok_w"][:n_tok], "tok_w")
    i = 0
    for r, q in enumerate(reqs):
        assert dec.timestamp(r) == q["timestamp"]
        for it in q["items"]:
            assert dec.item_id(i) == it["id"]
            i += 1
    dec.free(); fm.free()


def test_native_decoder_rejects_what_the_reference_decoder_rejects():
    fm = F.FeatureMapping(None, FEATS, MODEL)
    ok = dict(id="r", timestamp=1, items=[dict(id="p1")])
    bad_bodies = [
        "[1, 2]", "{}", json.dumps(dict(ok, items=[])), json.dumps(dict(id="r", items=[dict(id="p")])),
        json.dumps(dict(ok, timestamp=True)), json.dumps(dict(ok, timestamp="yesterday")), json.dumps(dict(ok, timestamp=1.5)),
        json.dumps(dict(ok, items=[dict(id=5)])), json.dumps(dict(ok, items=[dict(id="p", relevancy="high")])),
        json.dumps(dict(ok, fields=[dict(name="q", value=None)])), json.dumps(dict(ok, fields=[dict(name="q", value={"a": 1})])),
        json.dumps(dict(ok, fields=[dict(name="q", value=["a", 1])])), json.dumps(dict(ok, fields=[dict(value="x")])),
        json.dumps(dict(ok, user=7)), "{not json",
        # T/model/FieldTest.scala:15-17,31-33,40-42: NaN literals and a list of booleans are decoding failures
        '{"id": "r", "timestamp": 1, "items": [{"id": "p"}], "fields": [{"name": "f", "value": NaN}]}',
        '{"id": "r", "timestamp": 1, "items": [{"id": "p"}], "fields": [{"name": "f", "value": [1, 2, 3, NaN]}]}',
        json.dumps(dict(ok, fields=[dict(name="t", value=[True, False])])),
    ]
    for body in bad_bodies:
        with pytest.raises(_capi.MrError):
            F.DecodedRequests(fm, body)
        with pytest.raises((DecodingFailure, Exception)):
            fm.pack_requests([decode_ranking_event(body)])
    # a Lucene-analyzed field_match without caller-side tokens cannot be served natively: loud, not silent
    with pytest.raises(_capi.MrError) as e:
        F.DecodedRequests(fm, json.dumps(dict(ok, fields=[dict(name="query", value="red shoes")])))
    assert "tokens" in str(e.value)
    fm.free()


def test_native_decoder_on_the_reference_event_json_cases():
    """T/model/EventJsonTest.scala:86-121 ("decode ranking": string timestamp, user / session, two string fields,
    three items with the relevancy sugar) and :155-160 (timestamp as long / numeric string / ISO instant)."""
    body = """{
      "event": "ranking",
      "id": "81f46c34-a4bb-469c-8708-f8127cd67d27",
      "timestamp": "1599391467000",
      "user": "user1",
      "session": "session1",
      "fields": [
          {"name": "query", "value": "jeans"},
          {"name": "source", "value": "search"}
      ],
      "items": [
        {"id": "product3", "relevancy":  2.0},
        {"id": "product1", "relevancy":  1.0},
        {"id": "product2", "relevancy":  0.5}
      ]
    }"""
    feats = [dict(name="rel", type="relevancy"),
             dict(name="qwords", type="word_count", scope="ranking", source="ranking.query"),
             dict(name="src", type="string", scope="ranking", source="ranking.source", encode="index", values=["ads", "search"])]
    fm = F.FeatureMapping(None, feats, ["rel", "qwords", "src"])
    dec = F.DecodedRequests(fm, body)
    a = dec.arrays()
    assert dec.timestamp(0) == 1599391467000
    assert a["users"].tolist() == [F.hash64("user1")] and a["sessions"].tolist() == [F.hash64("session1")]
    assert a["ids"].tolist() == [F.hash64(x) for x in ("product3", "product1", "product2")]
    assert [dec.item_id(k) for k in range(3)] == ["product3", "product1", "product2"]
    assert a["item_f64"][:, fm.input_slot(F.MR_IN_ITEM_F64, "rel")].tolist() == [2.0, 1.0, 0.5]
    assert a["req_f64"][0, fm.input_slot(F.MR_IN_REQ_F64, "qwords")] == 1.0
    assert a["req_f64"][0, fm.input_slot(F.MR_IN_REQ_F64, "src")] == 2.0   # "search" is the 2nd value: index + 1
    dec.free()
    for ts, want in [("123", 123), ('"123"', 123), ('"2022-06-22T11:21:39Z"', 1655896899000)]:
        d = F.DecodedRequests(fm, '{"id": "r", "timestamp": %s, "items": [{"id": "p"}]}' % ts)
        assert d.timestamp(0) == want
        d.free()
    fm.free()


def test_native_decoder_is_insensitive_to_key_order_whitespace_and_unknown_keys():
    """`value` before `name`, items before fields, pretty-printing, nested unknown keys: same arrays."""
    from metarank_b200 import synth

    feats, model = synth.ranklens_config()
    state, item_ids, sessions = synth.ranklens_state(n_items=200, n_sessions=10, seed=21)
    reqs = synth.ranklens_requests(item_ids, sessions, 12, 30, seed=22)
    fm = F.FeatureMapping(None, feats, model)
    bodies = [dict(items=[dict(fields=[dict(value=v, name=n) for n, v in it.get("fields", [])], id=it["id"], label=1)
                          for it in q["items"]],
                   fields=[dict(value=v, extra={"a": [1, {"b": None}]}, name=n) for n, v in q.get("fields", [])],
                   session=q.get("session"), user=q.get("user"), timestamp=str(q["timestamp"]), id=q["id"],
                   event="ranking", unknown=[1, 2, {"x": 'y"z\\'}]) for q in reqs]
    want = fm.pack_requests(reqs)
    for text in (json.dumps(bodies, indent=3), json.dumps(bodies, separators=(",", ":"))):
        got = F.DecodedRequests(fm, text).arrays()
        for k in ["offsets", "ids", "users", "sessions", "req_f64", "req_u64", "item_f64"]:
            _same(got[k], want[k], k)
    fm.free()


def test_native_decoder_survives_corrupted_bodies():
    """Byte flips, truncations and hostile escapes end in MR_ERR_PARSE / MR_ERR_INVALID_ARG or a valid decode,
    never in a crash or an out-of-bounds read."""
    rng = np.random.Generator(np.random.PCG64(77))
    fm = F.FeatureMapping(None, FEATS, MODEL)
    base = json.dumps(_random_body(rng, 3), ensure_ascii=True).encode()
    ok = err = 0
    for it in range(3000):
        b = bytearray(base)
        for _ in range(int(rng.integers(1, 5))):
            k = int(rng.integers(0, len(b)))
            b[k] = int(rng.choice([rng.integers(0, 256), ord('"'), ord("\\"), ord("{"), ord("["), ord("u"), ord(",")]))
        if rng.random() < 0.3:
            b = b[:int(rng.integers(0, len(b)))]
        try:
            d = F.DecodedRequests(fm, bytes(b))
            a = d.arrays()
            assert a["offsets"][-1] == a["total_items"]
            d.free()
            ok += 1
        except _capi.MrError:
            err += 1
    assert ok > 0 and err > 0
    for tail in ['"\\u12', '"\\ud83d\\u', '"\\', '"abc', "[[[[[[[[", '{"id": "r", "timestamp": 1, "items": [{"id": "\\ud83d']:
        with pytest.raises(_capi.MrError):
            F.DecodedRequests(fm, '{"id": "r", "timestamp": 1, "items": [{"id": "p"}], "fields": [{"name": "q", "value": ' + tail)
    with pytest.raises(_capi.MrError):  # a megabyte of '[' must not blow the stack
        F.DecodedRequests(fm, '{"id": "r", "timestamp": 1, "items": [{"id": "p"}], "junk": ' + "[" * 1_000_000)
    fm.free()


def test_native_decoder_is_as_strict_as_jawn_and_takes_the_last_duplicate():
    """duplicate keys resolve to the LAST occurrence (circe's JsonObject, json.loads), bytes after the
    top-level value are a failure, and '+1', '01', '1.', '.5', \\uZZZZ are not JSON."""
    fm = F.FeatureMapping(None, FEATS, MODEL)
    base = '"items": [{"id": "p1"}, {"id": "p2", "relevancy": 1, "relevancy": 3}]'
    body = '{"id": "a", "timestamp": 1000, "timestamp": 2000, "user": "u1", "user": null, "id": "b", ' + base + "}"
    dec = F.DecodedRequests(fm, body)
    want = fm.pack_requests([decode_ranking_event(body)])
    got = dec.arrays()
    assert dec.timestamp(0) == 2000 == decode_ranking_event(body)["timestamp"]
    for k in ["offsets", "ids", "users", "sessions", "req_f64", "req_u64", "item_f64"]:
        _same(got[k], want[k], k)
    assert got["users"][0] == 0  # the later `null` wins
    dec.free()
    two_item_lists = '{"id": "a", "timestamp": 1, "items": [{"id": "x"}], "items": [{"id": "p1"}, {"id": "p2"}]}'
    dec = F.DecodedRequests(fm, two_item_lists)
    assert dec.total_items == 2 and [dec.item_id(i) for i in range(2)] == ["p1", "p2"]
    dec.free()
    ok = '{"id": "r", "timestamp": 1, "items": [{"id": "p"}]}'
    for bad in [ok + " garbage", ok + "{}", ok + ",", "[" + ok + "] 1",
                ok.replace('"timestamp": 1', '"timestamp": +1'), ok.replace('"timestamp": 1', '"timestamp": 01'),
                ok.replace('"timestamp": 1', '"timestamp": 1.'), ok.replace('{"id": "p"}', '{"id": "p", "relevancy": .5}'),
                ok.replace('{"id": "p"}', '{"id": "p", "relevancy": 1e}'), ok.replace('"id": "r"', '"id": "\\\\uZZZZ"'.replace("\\\\", "\\")),
                ok.replace('{"id": "p"}', '{"id": "p", "relevancy": -}')]:
        with pytest.raises(_capi.MrError):
            F.DecodedRequests(fm, bad)
        with pytest.raises(Exception):
            decode_ranking_event(bad)
    for good in [ok + "  \n\t ", ok.replace('{"id": "p"}', '{"id": "p", "relevancy": -0.5e+2}'), ok.replace('"timestamp": 1', '"timestamp": 0')]:
        F.DecodedRequests(fm, good).free()
    fm.free()


def _decode_all(fm, body, threads, monkeypatch):
    monkeypatch.setenv("MR_DECODE_THREADS", str(threads))
    dec = F.DecodedRequests(fm, body)
    out = dec.arrays()
    out["_ids"] = [dec.item_id(i) for i in range(dec.total_items)]
    out["_ts"] = [dec.timestamp(r) for r in range(dec.n_requests)]
    dec.free()
    return out


def test_large_batches_decode_the_same_on_several_threads(monkeypatch):
    """A body of hundreds of events is split at its top-level element boundaries and parsed / packed by several threads
    (request_codec.cpp decode_requests); the batch must be the one the sequential walk builds, array for array — with
    brackets, commas and escaped quotes inside strings, nested arrays, duplicate fields, tokens and embeddings in play."""
    rng = np.random.Generator(np.random.PCG64(77))
    fm = F.FeatureMapping(None, FEATS, MODEL)
    events = _random_body(rng, 900)
    for k, tricky in enumerate(['a "quoted, [bracket]" \\ back', "]}],[{", 'x\\"y,z', "{\"k\": [1, 2]}"]):
        events[50 * k + 3]["fields"].append(dict(name="query", value=tricky))
        events[50 * k + 4]["items"][0]["id"] = tricky
    for indent, ascii_ in ((None, True), (1, False)):
        body = json.dumps(events, ensure_ascii=ascii_, indent=indent)
        assert len(body) > 256 << 10
        one = _decode_all(fm, body, 1, monkeypatch)
        for threads in (2, 8):
            many = _decode_all(fm, body, threads, monkeypatch)
            assert many["n_requests"] == one["n_requests"] == 900 and many["total_items"] == one["total_items"]
            for key in ["offsets", "ids", "users", "sessions", "req_f64", "req_u64", "req_vec", "req_vp", "item_f64", "tok_off",
                        "tok_hash", "tok_w"]:
                _same(many[key], one[key], key)
            assert many["_ids"] == one["_ids"] and many["_ts"] == one["_ts"]
    # no item carries fields: the override matrix is not built at all (mr_rank_batch.item_f64 == NULL) on either path
    bare = [dict(id=f"r{r}", timestamp=1710716400123, items=[dict(id=f"i{r}_{j}") for j in range(40)]) for r in range(800)]
    body = json.dumps(bare)
    assert len(body) > 256 << 10
    for threads in (1, 8):
        got = _decode_all(fm, body, threads, monkeypatch)
        assert got["item_f64"] is None and got["total_items"] == 32000
    fm.free()


def test_large_batches_fail_like_the_sequential_walk(monkeypatch):
    """What the sequential decoder rejects, the threaded one rejects with the same status and message: it drops back to
    the sequential walk on anything unusual, and a request-level error is the one of the FIRST offending request."""
    rng = np.random.Generator(np.random.PCG64(78))
    fm = F.FeatureMapping(None, FEATS, MODEL)
    events = _random_body(rng, 700)
    good = json.dumps(events)
    assert len(good) > 256 << 10

    def outcome(body, threads):
        monkeypatch.setenv("MR_DECODE_THREADS", str(threads))
        try:
            F.DecodedRequests(fm, body).free()
            return None
        except _capi.MrError as ex:
            return ex.status, ex.message

    cut = good.index('{"id": "r350"')
    bad_bodies = [
        good[:cut] + '{"id": "r350", "items": []},' + good[cut:],                    # an event the decoder rejects, mid-array
        good[:cut] + '{"id": "r350" "items": [{"id": "x"}], "timestamp": 1},' + good[cut:],   # malformed JSON inside one element
        good[:-1] + ', 17]',                                                            # an element that is not an object
        good + ' trailing',                                                             # bytes after the array
        good[:-1],                                                                      # unterminated array
        good[:cut] + ',' + good[cut:],                                                  # empty element
    ]
    import copy
    wrong = copy.deepcopy(events)
    for r, emb in ((10, [1.0, 2.0]), (600, [1.0])):   # wrong embedding dims in request 10 AND in request 600: the first decides
        wrong[r]["embeddings"] = {"sim": emb}
        wrong[r]["fields"].append(dict(name="query", value="red shoes"))
    bad_bodies.append(json.dumps(wrong))
    for body in bad_bodies:
        want = outcome(body, 1)
        assert want is not None, body[:80]
        for threads in (2, 8):
            assert outcome(body, threads) == want
    assert outcome(good, 8) is None
    fm.free()
