"""Pins the feature-assembly oracle on the reference's own golden vectors (no GPU)."""
import math

import numpy as np
import pytest

import golden_cases as G
from oracle import features_oracle as fo


def _same(a, b):
    return (a != a and b != b) or a == b


@pytest.mark.parametrize("case", G.CASES + G.ORACLE_ONLY_CASES, ids=[c["name"] for c in G.CASES + G.ORACLE_ONLY_CASES])
def test_reference_golden_vector(case):
    mapping = fo.FeatureMapping(case["features"], case["model_features"])
    state = fo.FeatureValueFlow(mapping, always_refresh=True).process(case["events"])
    ivs = fo.item_values(mapping, case["request"], state, mode="online")
    for name, want in case["expected"].items():
        got = [iv[name] for iv in ivs]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert len(g) == len(w) and all(_same(x, y) for x, y in zip(g, w)), (case["ref"], name, got, want)


def test_position_offline_mode():
    # T/feature/PositionFeatureTest.scala:17-22
    f = fo.PositionFeature(dict(name="pos", position=5))
    assert f.values(G.ranking(["p1", "p2", "p3"]), {}, mode="offline") == [[0.0], [1.0], [2.0]]


def test_dense_layout_matches_clickthrough_query_test():
    # T/flow/ClickthroughQueryTest.scala:152-159: columns=5, rows=3 and the 15 row-major values
    mapping = fo.FeatureMapping(G.LAYOUT_FEATURES, [f["name"] for f in G.LAYOUT_FEATURES])
    assert mapping.dim == 5
    out = np.zeros((3, mapping.dim))
    for i, iv in enumerate(G.LAYOUT_ITEM_VALUES):
        for name, vals in iv.items():
            o, d = mapping.offsets[name]
            out[i, o:o + d] = vals
    assert out.reshape(-1).tolist() == G.LAYOUT_EXPECTED


@pytest.mark.parametrize("offsets,expected", [
    ([0], [1, 1]),                                   # increment once
    ([-h * 3600_000 for h in range(0, 10)], [1, 10]),  # intra-day burst (now-9h .. now)
    ([-d * 86400_000 for d in range(0, 10)], [1, 8]),  # once a day
    ([-7 * d * 86400_000 for d in range(0, 10)], [1, 2]),  # once a week
])
def test_periodic_counter_windows(offsets, expected):
    # T/fstore/PeriodicCounterSuite.scala:22-144, ranges (0,0) and (7,0), period 1 day
    conf = dict(period=86400_000, ranges=[(0, 0), (7, 0)])
    st = fo.MemState()
    key = (("item", "p"), "f1")
    for o in sorted(offsets):
        st.put(("pinc", key, G.NOW + o, 1), {"f1": conf})
    assert st.compute_value(("pinc", key, G.NOW, 1), {"f1": conf}) == ("pcounter", expected)


def test_normalize_golden():
    # T/ml/onnx/NormalizeTest.scala:10-59
    nan = float("nan")
    assert fo.normalize_scale("linear", [1.0, 2.0, 3.0]) == [0.0, 0.5, 1.0]
    r = fo.normalize_scale("linear", [1.0, 2.0, nan])
    assert r[:2] == [0.0, 1.0] and math.isnan(r[2])
    assert fo.normalize_scale("position", [1.0, 4.0, 3.0, 2.0, 5.0]) == [0.0, 0.6, 0.4, 0.2, 0.8]
    r = fo.normalize_scale("position", [nan, 1.0, 4.0, 3.0, 2.0])
    assert math.isnan(r[0]) and r[1:] == [0.0, 0.6, 0.4, 0.2]


def test_token_count_java_split_semantics():
    # WordCountFeature.tokenCount = "\\s+".split(s).length (S/feature/WordCountFeature.scala:73-76)
    assert [fo.token_count(s) for s in ["foo, bar, baz!", "foo bar", "", " lead", "trail  ", "   ", "a\tb\nc"]] == \
        [3, 2, 1, 2, 1, 0, 3]


def test_percentile_legacy():
    # commons-math3 Percentile(50), LEGACY: pos = 0.5 (n + 1)
    assert fo.percentile50_legacy([10.0, 20.0, 40.0, 15.0, 5.0]) == 15.0
    assert fo.percentile50_legacy([1.0, 2.0]) == 1.5
    assert fo.percentile50_legacy([7.0]) == 7.0
    assert math.isnan(fo.percentile50_legacy([]))
    assert fo.percentile50_legacy([float("nan"), 3.0, 1.0]) == 2.0


def test_normalized_rate_zero_global_top_raises():
    # Long / Long with topGlobal == 0 is an ArithmeticException in the reference (:343-350)
    f = fo.RateFeature(dict(G.RATE, normalize={"weight": 10}))
    st = {(("item", "p1"), "ctr_click"): ("pcounter", [0, 0]), (("item", "p1"), "ctr_impression"): ("pcounter", [3, 3]),
          (("global",), "ctr_click_norm"): ("pcounter", [0, 0]), (("global",), "ctr_impression_norm"): ("pcounter", [3, 3])}
    with pytest.raises(ZeroDivisionError):
        f.value(G.ranking(["p1"]), st, {"id": "p1"})


def test_rate_plain_division_edge_cases():
    f = fo.RateFeature(G.RATE)
    st = {(("item", "p1"), "ctr_click"): ("pcounter", [0, 2]), (("item", "p1"), "ctr_impression"): ("pcounter", [0, 0])}
    v = f.value(G.ranking(["p1"]), st, {"id": "p1"})
    assert math.isnan(v[0]) and v[1] == math.inf
    st[(("item", "p1"), "ctr_click")] = ("pcounter", [1])  # wrong length -> missing
    assert all(math.isnan(x) for x in f.value(G.ranking(["p1"]), st, {"id": "p1"}))


def test_bm25_matches_the_reference_test_values():
    """T/feature/matcher/BM25MatcherTest.scala:21-27: 0.15 +- 0.01 and 1.34 +- 0.01 (the reference's own tolerance:
    math.log is not bit-reproducible across JVMs either)."""
    bm = fo.FieldMatchTokensFeature(dict(name="m", rankingField="ranking.q", itemField="item.t",
                                         method=dict(type="bm25", language="en", docs=3, avgdl=3.0,
                                                     termfreq={"foo": 1, "bar": 2, "baz": 3})))
    assert abs(bm.score(["baz"], ["bar", "baz"]) - 0.15) <= 0.01
    assert abs(bm.score(["foo"], ["foo"]) - 1.34) <= 0.01
    ng = fo.FieldMatchTokensFeature(dict(name="m", rankingField="ranking.q", itemField="item.t",
                                         method=dict(type="ngram", n=3, language="whitespace")))
    assert ng.tokenize("fooba foo") == ["foo", "oba", "oob"]          # NgramMatcherTest.scala:11-14
    assert ng.tokenize("foobar") == ["bar", "foo", "oba", "oob"]      # :16-19
    assert (ng.score(list("abc"), list("abc")), ng.score(["a"], ["a", "b"]), ng.score(["c", "d"], ["a", "b"])) == (1.0, 0.5, 0.0)


def test_token_match_schema_and_request_packing_on_the_host():
    """mr_schema_create accepts field_match ngram / term / bm25 and exposes one MR_IN_REQ_TOKENS slot per feature;
    the Python shim packs the request's token lists the way mr_rank_batch documents."""
    from metarank_b200 import _capi, features as F

    feats = [dict(name="tm", type="field_match", rankingField="ranking.query", itemField="item.title",
                  method=dict(type="ngram", n=3, language="whitespace")),
             dict(name="price", type="number", scope="item", source="item.price"),
             dict(name="bm", type="field_match", rankingField="ranking.query", itemField="item.title",
                  method=dict(type="bm25", language="en", docs=3, avgdl=3.0, termfreq={"foo": 1}))]
    fm = F.FeatureMapping(None, feats, ["tm", "price", "bm"])
    assert fm.dim == 3 and fm.n_req_tok == 2
    assert (fm.input_slot(F.MR_IN_REQ_TOKENS, "tm"), fm.input_slot(F.MR_IN_REQ_TOKENS, "bm"),
            fm.input_slot(F.MR_IN_REQ_TOKENS, "price")) == (0, 1, -1)
    r1 = G.ranking(["p1", "p2"], [("query", "foobar")])
    r1["tokens"] = {"bm": ["foo", "zzz"]}
    r2 = G.ranking(["p3"])  # no query field: empty lists
    a = fm.pack_requests([r1, r2])
    assert a["tok_off"].tolist() == [0, 4, 6, 6, 6]
    assert a["tok_hash"][:6].tolist() == [F.hash64(t) for t in ["bar", "foo", "oba", "oob", "foo", "zzz"]]
    assert a["tok_w"][:4].tolist() == [0.0] * 4 and a["tok_w"][4] == F.bm25_idf(feats[2]["method"], "foo")
    fm.free()
    with pytest.raises(_capi.MrError) as e:
        F.FeatureMapping(None, [dict(feats[2], method=dict(type="bm25", language="en"))], ["bm"])
    assert "avgdl" in str(e.value)


def test_referer_reads_the_stored_medium():
    """T/feature/RefererFeatureTest.scala:37-51: the write path (JVM: snowplow's referer database) stores
    SString("search") under the user for http://www.google.com; the ranking then yields CategoryValue("search", 1).
    The table S/feature/RefererFeature.scala:47-54 for the other mediums; absent / unknown strings -> 0."""
    from oracle import features_oracle as fo

    feats = [dict(name="ref_medium", type="referer", source="ranking.ref", scope="user")]
    mapping = fo.FeatureMapping(feats, ["ref_medium"])
    req = G.ranking(["p1", "p2"], user="u1")
    state = {(("user", "u1"), "ref_medium"): ("scalar", "search")}
    assert fo.dense_matrix(mapping, req, state).tolist() == [[1.0], [1.0]]
    for medium, idx in (("unknown", 0), ("internal", 2), ("social", 3), ("email", 4), ("paid", 5), ("carrier pigeon", 0)):
        state = {(("user", "u1"), "ref_medium"): ("scalar", medium)}
        assert fo.dense_matrix(mapping, req, state).tolist() == [[float(idx)]] * 2
    assert fo.dense_matrix(mapping, req, {}).tolist() == [[0.0], [0.0]]
    assert fo.dense_matrix(mapping, G.ranking(["p1"], user=None), state).tolist() == [[0.0]]
