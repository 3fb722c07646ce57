"""JSON edge of POST /rank (no GPU): RankingEvent decoding and RankResponse encoding follow the
reference's circe codecs (golden JSON from T/model/EventJsonTest.scala)."""
import json

import pytest

from metarank_b200 import rank_api as ra


def test_decode_ranking_event_golden():
    # T/model/EventJsonTest.scala:86-124
    text = """{
      "event": "ranking", "id": "81f46c34-a4bb-469c-8708-f8127cd67d27", "timestamp": "1599391467000",
      "user": "user1", "session": "session1",
      "fields": [{"name": "query", "value": "jeans"}, {"name": "source", "value": "search"}],
      "items": [{"id": "product3", "relevancy": 2.0}, {"id": "product1", "relevancy": 1.0}, {"id": "product2", "relevancy": 0.5}]
    }"""
    ev = ra.decode_ranking_event(text)
    assert ev["id"] == "81f46c34-a4bb-469c-8708-f8127cd67d27" and ev["timestamp"] == 1599391467000
    assert (ev["user"], ev["session"]) == ("user1", "session1")
    assert ev["fields"] == [("query", "jeans"), ("source", "search")]
    assert [(i["id"], i["fields"]) for i in ev["items"]] == [
        ("product3", [("relevancy", 2.0)]), ("product1", [("relevancy", 1.0)]), ("product2", [("relevancy", 0.5)])]


def test_decode_timestamp_long_string_iso():
    # T/model/EventJsonTest.scala:155-160
    assert ra.decode_timestamp(123) == 123
    assert ra.decode_timestamp("123") == 123
    assert ra.decode_timestamp("2022-06-22T11:21:39Z") == 1655896899000


def test_decode_fields_all_types_and_errors():
    ev = ra.decode_ranking_event(json.dumps({"id": "r", "timestamp": 1, "items": [
        {"id": "a", "relevancy": 1, "fields": [{"name": "price", "value": 3}, {"name": "tags", "value": ["x", "y"]},
                                               {"name": "ok", "value": True}, {"name": "v", "value": [1, 2.5]}]}]}))
    assert ev["items"][0]["fields"] == [("relevancy", 1.0), ("price", 3.0), ("tags", ["x", "y"]), ("ok", True), ("v", [1.0, 2.5])]
    assert ev["user"] is None and ev["session"] is None and ev["fields"] == []
    for bad in ('{"id":"r","timestamp":1,"items":[]}', '{"id":"r","items":[{"id":"a"}]}', "not json",
                '{"id":"r","timestamp":1,"items":[{"id":"a","fields":[{"name":"x","value":null}]}]}',
                '{"id":"r","timestamp":1,"items":[{"id":"a","fields":[{"name":"x","value":{"a":1}}]}]}',
                '{"id":"r","timestamp":1,"items":[{"id":"a","fields":[{"name":"x","value":["a",1]}]}]}'):
        with pytest.raises(ra.DecodingFailure):
            ra.decode_ranking_event(bad)


def test_response_encoding_drops_nulls_sorts_keys_and_maps_nan():
    # JsonChunk (dropNullValues, sortKeys); MValue NaN -> null (T/model/MValueJsonTest.scala)
    resp = {"state": None, "took": 1, "items": [{"item": "p1", "score": 0.5, "features": {"b": None, "a": [1.0, None]}}]}
    out = ra.encode_response(resp)
    d = json.loads(out)
    assert "state" not in d and list(d) == ["items", "took"]
    assert d["items"][0] == {"features": {"a": [1.0, None]}, "item": "p1", "score": 0.5}
    assert ra._num(float("nan")) is None and ra._num(1.5) == 1.5


def test_routes_reject_everything_but_post_rank():
    api = ra.RankApi({})
    assert api.routes("GET", "/rank/m", "")[0] == 404
    assert api.routes("POST", "/feedback", "")[0] == 404
    status, _, body = api.routes("POST", "/rank/nope", '{"id":"r","timestamp":1,"items":[{"id":"a"}]}')
    assert status == 500 and "model nope is not configured" in body  # ModelError -> 500 like the reference
