"""Golden vectors transcribed from the reference's own tests (src/test/scala/ai/metarank/...).

Each case = (feature config dicts, events replayed through the write path, ranking request,
expected per-feature values) with the reference file:line it was taken from.  They pin the
CPU oracle (tests/test_features_golden.py) and, through the oracle AND directly, the CUDA
path (tests/test_features_gpu.py).
"""
import itertools

NOW = 1622505601000  # Timestamp.date(2021, 6, 1, 0, 0, 1), T/fstore/FeatureSuite.scala:14
_ids = itertools.count()


def item_event(item, fields=()):
    """T/util/TestItemEvent.scala"""
    return dict(event="item", id=f"e{next(_ids)}", item=item, timestamp=NOW, fields=list(fields))


def interaction(item, ranking, typ="click", user="u1", session="s1", ts=NOW):
    """T/util/TestInteractionEvent.scala"""
    return dict(event="interaction", id=f"e{next(_ids)}", item=item, timestamp=ts, ranking=ranking, user=user,
                session=session, type=typ, fields=[])


def ranking(items, fields=(), rid=None, user="u1", session="s1", item_fields=None):
    """T/util/TestRankingEvent.scala: every item carries relevancy 1.0"""
    return dict(event="ranking", id=rid or f"r{next(_ids)}", timestamp=NOW, user=user, session=session,
                fields=list(fields),
                items=[dict(id=i, fields=[("relevancy", 1.0)] + list((item_fields or {}).get(i, []))) for i in items])


CASES = []


def case(name, ref, features, events, request, expected, model_features=None):
    CASES.append(dict(name=name, ref=ref, features=features, events=events, request=request, expected=expected,
                      model_features=model_features or [f["name"] for f in features]))


RATE = dict(name="ctr", type="rate", top="click", bottom="impression", bucket="24h", periods=[7, 14], refresh="0s")

case("rate_plain", "T/feature/RateFeatureTest.scala:61-74", [RATE],
     [interaction("p1", "x", "impression")] * 4 + [interaction("p1", "x", "click")],
     ranking(["p1"]), {"ctr": [[0.25, 0.25]]})

case("rate_normalized_integer_division", "T/feature/NormRateFeatureTest.scala:64-81",
     [dict(RATE, normalize={"weight": 10})],
     [interaction("p1", "x", "impression")] * 3 + [interaction("p1", "x", "click")]
     + [interaction("p2", "x", "impression")] * 90 + [interaction("p2", "x", "click")] * 9,
     ranking(["p1"]), {"ctr": [[0.11827956989247312, 0.11827956989247312]]})

_scoped_events = [
    item_event("p1", [("color", "red")]), item_event("p2", [("color", "red")]), item_event("p3", [("color", "red")]),
    item_event("p4", [("color", "green")]), item_event("p5", [("size", "xl")]),
    interaction("p1", "x", "impression"), interaction("p2", "x", "impression"), interaction("p3", "x", "impression"),
    interaction("p2", "x", "impression"), interaction("p4", "x", "impression"), interaction("p5", "x", "impression"),
    interaction("p1", "x", "click"), interaction("p4", "x", "click"), interaction("p5", "x", "click"),
]
case("rate_item_field_scope", "T/feature/ScopedRateFeatureTest.scala:60-83", [dict(RATE, scope="item.color")],
     _scoped_events, ranking(["p1"]), {"ctr": [[0.25, 0.25]]})
case("rate_item_field_scope_string_lists", "T/feature/ScopedRateFeatureTest.scala:85-108", [dict(RATE, scope="item.color")],
     [item_event("p1", [("color", ["red"])]), item_event("p2", [("color", ["red"])]),
      item_event("p3", [("color", ["red"])])] + _scoped_events[3:],
     ranking(["p1"]), {"ctr": [[0.25, 0.25]]})

case("rate_ranking_field_scope", "T/feature/RankFieldScopedRateFeatureTest.scala:48-65",
     [dict(RATE, scope="ranking.query")],
     [ranking(["p1", "p2"], [("query", "test")], rid="r1"),
      interaction("p1", "r1", "impression"), interaction("p2", "r1", "impression"), interaction("p1", "r1", "click"),
      ranking(["p1", "p2"], [("query", "test")], rid="r2"),
      interaction("p1", "r2", "impression"), interaction("p2", "r2", "impression"), interaction("p2", "r2", "click")],
     ranking(["p1"], [("query", "test")]), {"ctr": [[0.5, 0.5]]})

case("window_count", "T/feature/WindowInteractionCountFeatureTest.scala:46-56",
     [dict(name="cnt", type="window_count", interaction="click", bucket="24h", periods=[1], scope="item")],
     [interaction("p1", "e0")] * 3, ranking(["p1"]), {"cnt": [[3.0]]})

case("interaction_count", "T/feature/InteractionCountTest.scala:50-57",
     [dict(name="cnt", type="interaction_count", interaction="click", scope="item")],
     [interaction("p1", "x")] * 3, ranking(["p1"]), {"cnt": [[3.0]]})

_seen = dict(name="seen", type="interacted_with", interaction="impression", field="item.color", scope="session",
             count=10, duration="24h")
_seen_events = [item_event("p1", [("color", "red")]), item_event("p2", [("color", "green")]),
                interaction("p1", "i1", "impression"), interaction("p2", "i1", "impression")]
case("interacted_with_one_field", "T/feature/InteractedWithFeatureTest.scala:105-119", [_seen], _seen_events,
     ranking(["p1", "p2", "p3"]), {"seen": [[1.0], [1.0], [0.0]]})
case("interacted_with_two_fields", "T/feature/InteractedWithFeatureTest.scala:121-144",
     [dict(_seen, field=["item.color", "item.tags"])], _seen_events,
     ranking(["p1", "p2", "p3"]), {"seen": [[1.0, 0.0], [1.0, 0.0], [0.0, 0.0]]})

case("word_count_item", "T/feature/WordCountFeatureItemTest.scala:46-53",
     [dict(name="title_words", type="word_count", scope="item", source="metadata.title")],
     [item_event("p1", [("title", "foo, bar, baz!")])], ranking(["p1"]), {"title_words": [[3.0]]})
case("word_count_ranking", "T/feature/WordCountFeatureRankingTest.scala:36-43",
     [dict(name="query_words", type="word_count", scope="ranking", source="ranking.query")],
     [], ranking(["p1"], [("query", "foo bar")]), {"query_words": [[2.0]]})

_rel_req = ranking(["p1", "p2"])
_rel_req["items"][0]["fields"] = [("relevancy", 1.0)]
_rel_req["items"][1]["fields"] = [("relevancy", 2.0)]
case("relevancy", "T/feature/RelevancyTest.scala:16-25", [dict(name="rel", type="relevancy")], [], _rel_req,
     {"rel": [[1.0], [2.0]]})

case("position_online", "T/feature/PositionFeatureTest.scala:24-28", [dict(name="pos", type="position", position=5)],
     [], ranking(["p1", "p2", "p3"]), {"pos": [[5.0], [5.0], [5.0]]})

_prices = lambda ps: [item_event(f"p{i + 1}", [("price", p)]) for i, p in enumerate(ps)]  # noqa: E731
case("diversity_numbers", "T/feature/DiversityFeatureTest.scala:12-32",
     [dict(name="divnum", type="diversity", source="item.price", top=2147483647)], _prices([10.0, 20.0, 40.0, 15.0, 5.0]),
     ranking(["p1", "p2", "p3", "p4", "p5"]), {"divnum": [[-5.0], [5.0], [25.0], [0.0], [-10.0]]})
case("diversity_top3_numbers", "T/feature/DiversityFeatureTest.scala:34-55",
     [dict(name="divnum", type="diversity", source="item.price", top=3)], _prices([10.0, 20.0, 30.0, 5.0, 1.0]),
     ranking(["p1", "p2", "p3", "p4", "p5"]), {"divnum": [[-10.0], [0.0], [10.0], [-15.0], [-19.0]]})
case("diversity_strings", "T/feature/DiversityFeatureTest.scala:57-78",
     [dict(name="divstr", type="diversity", source="item.cat", top=2147483647)],
     [item_event(f"p{i + 1}", [("cat", c)]) for i, c in enumerate("abcab")],
     ranking(["p1", "p2", "p3", "p4", "p5"]), {"divstr": [[0.4], [0.4], [0.2], [0.4], [0.4]]})
case("diversity_string_lists", "T/feature/DiversityFeatureTest.scala:79-99",
     [dict(name="divstrl", type="diversity", source="item.cat", top=2147483647)],
     [item_event("p1", [("cat", ["a"])]), item_event("p2", [("cat", ["b", "c"])]),
      item_event("p3", [("cat", ["a", "b", "c"])]), item_event("p4", [("cat", ["a", "b", "c", "d"])])],
     ranking(["p1", "p2", "p3", "p4"]), {"divstrl": [[0.3], [0.6], [0.9], [1.0]]})

case("vector_default_reducers", "T/feature/NumVectorFeatureTest.scala:62-70",
     [dict(name="vec", type="vector", source="item.vec", scope="item")],
     [item_event("p1", [("vec", [1.0, 2.0, 3.0])])], ranking(["p1", "p2"]),
     {"vec": [[1.0, 3.0, 3.0, 2.0], [float("nan")] * 4]})

_UPD = 1646085600  # ZonedDateTime.of(2022, 3, 1, 0, 0, 0, 0, UTC+2).toEpochSecond
_NOWMS = 1648418400000  # ZonedDateTime.of(2022, 3, 28, 0, 0, 0, 0, UTC+2) in epoch millis
_age_req = ranking(["p1"])
_age_req["timestamp"] = _NOWMS
_age_ev = item_event("p1", [("updated_at", "2022-03-01T00:00:00+02:00[UTC+02:00]")])
_age_ev["timestamp"] = _UPD * 1000
case("item_age_iso_string", "T/feature/ItemAgeFeatureTest.scala:90-101",
     [dict(name="itemage", type="item_age", source="item.updated_at")], [_age_ev], _age_req, {"itemage": [[2332800.0]]})

_lt = lambda parse, src="ranking.localts": [dict(name="x", type="local_time", source=src, parse=parse)]  # noqa: E731
_lt_req = ranking(["p1"], [("localts", "2022-03-28T12:00:00+02:00[UTC+02:00]")])
for _p, _v in (("time_of_day", 12.0), ("day_of_week", 1.0), ("month_of_year", 3.0), ("year", 2022.0), ("second", 1648461600.0)):
    case(f"local_time_{_p}", "T/feature/LocalDateTimeFeatureTest.scala:44-62", _lt(_p), [], _lt_req, {"x": [[_v]]})
_lt_native = ranking(["p1"])
_lt_native["timestamp"] = 1648461600000
case("local_time_native_timestamp", "T/feature/LocalDateTimeFeatureTest.scala:64-73", _lt("year", "ranking.timestamp"), [],
     _lt_native, {"x": [[2022.0]]})
case("local_time_bad_format_is_missing", "T/feature/LocalDateTimeFeatureTest.scala:33-42", _lt("time_of_day"), [],
     ranking(["p1"], [("localts", "now")]), {"x": [[float("nan")]]})

# Known-answer cases derived from the reference source where its tests hold no value
case("boolean_item", "S/feature/BooleanFeature.scala:47-62 (derived)",
     [dict(name="avail", type="boolean", scope="item", source="item.availability")],
     [item_event("p1", [("availability", True)]), item_event("p2", [("availability", False)])],
     ranking(["p1", "p2", "p3"]), {"avail": [[1.0], [0.0], [float("nan")]]})
case("number_item_and_override", "S/feature/NumberFeature.scala:58-97 (derived)",
     [dict(name="price", type="number", scope="item", source="metadata.price")],
     [item_event("p1", [("price", 10.0)]), item_event("p2", [("price", 20.0)])],
     ranking(["p1", "p2", "p3"], item_fields={"p2": [("price", 99.0)]}),
     {"price": [[10.0], [99.0], [float("nan")]]})
case("string_index_unknown_is_zero", "S/feature/StringFeature.scala:124-137 (derived)",
     [dict(name="color", type="string", scope="item", source="metadata.color", encode="index", values=["red", "green", "blue"])],
     [item_event("p1", [("color", "green")]), item_event("p2", [("color", ["pink", "red"])])],
     ranking(["p1", "p2", "p3"]), {"color": [[2.0], [0.0], [0.0]]})
case("string_onehot", "S/util/OneHotEncoder.scala:12-24 (derived)",
     [dict(name="color", type="string", scope="item", source="metadata.color", values=["red", "green", "blue"])],
     [item_event("p1", [("color", ["blue", "red", "pink"])])],
     ranking(["p1", "p2"]), {"color": [[1.0, 0.0, 1.0], [0.0, 0.0, 0.0]]})

# More of the reference's own vectors for extractors whose cases above are derived from the source.  They pin the ORACLE
# only (tests/test_features_golden.py): the device is held to the oracle on these extractors by the cases above and by the
# randomised configs of tests/test_features_gpu.py.
ORACLE_ONLY_CASES = []


def oracle_case(name, ref, features, events, request, expected):
    ORACLE_ONLY_CASES.append(dict(name=name, ref=ref, features=features, events=events, request=request, expected=expected,
                                  model_features=[f["name"] for f in features]))


_color = dict(name="color", type="string", scope="item", source="metadata.color", values=["red", "green", "blue"])
oracle_case("string_item_onehot_default", "T/feature/StringFeatureTest.scala:101-108", [_color],
            [item_event("p1", [("color", "green")])], ranking(["p1"]), {"color": [[0.0, 1.0, 0.0]]})
oracle_case("string_from_ranking_field", "T/feature/StringFeatureTest.scala:110-125",
            [dict(_color, source="ranking.color")], [], ranking(["p1"], [("color", "red")]), {"color": [[1.0, 0.0, 0.0]]})
_click = dict(interaction("p1", "p0"), fields=[("country", "EU")])
oracle_case("string_session_scope", "T/feature/StringFeatureTest.scala:127-144",
            [dict(name="country", type="string", scope="session", source="interaction:click.country", values=["US", "EU"])],
            [_click], ranking(["p1"]), {"country": [[0.0, 1.0]]})
oracle_case("string_onehot_explicit", "T/feature/StringFeatureTest.scala:146-164",
            [dict(name="country", type="string", scope="session", source="interaction:click.country", values=["us", "eu"],
                  encode="onehot")],
            [dict(interaction("p1", "p0"), fields=[("country", "eu")])], ranking(["p1"]), {"country": [[0.0, 1.0]]})
oracle_case("string_override_from_rank_item", "T/feature/StringFeatureTest.scala:166-172", [_color], [],
            ranking(["p1"], item_fields={"p1": [("color", "red")]}), {"color": [[1.0, 0.0, 0.0]]})
_pop = dict(name="popularity", type="number", scope="item", source="metadata.popularity")
oracle_case("number_item", "T/feature/NumberFeatureTest.scala:92-99", [_pop],
            [item_event("p1", [("popularity", 100)])], ranking(["p1"]), {"popularity": [[100.0]]})
oracle_case("number_override_from_rank_item", "T/feature/NumberFeatureTest.scala:101-109", [_pop], [],
            ranking(["p1"], item_fields={"p1": [("popularity", 100)]}), {"popularity": [[100.0]]})
oracle_case("number_ranking_scope", "T/feature/NumberFeatureTest.scala:111-126",
            [dict(name="weather_temp", type="number", scope="ranking", source="ranking.temp")], [],
            ranking(["p1"], [("temp", 10)]), {"weather_temp": [[10.0]]})

# ClickthroughQuery dense layout: T/flow/ClickthroughQueryTest.scala:152-159
LAYOUT_FEATURES = [
    dict(name="price", type="number", scope="item", source="metadata.price"),
    dict(name="category", type="string", scope="item", source="metadata.category", encode="index", values=["socks", "shirts"]),
    dict(RATE, name="ctr"),
    dict(name="clicked_category", type="interacted_with", interaction="click", field="metadata.category",
         scope="session", count=10, duration="24h"),
]
LAYOUT_ITEM_VALUES = [
    {"category": [1.0], "ctr": [0.2, 0.1], "price": [10.0], "clicked_category": [1.0]},
    {"price": [5.0], "ctr": [0.1, 0.05], "category": [2.0], "clicked_category": [0.0]},
    {"ctr": [0.2, 0.2], "clicked_category": [1.0], "price": [3.0], "category": [1.0]},
]
LAYOUT_EXPECTED = [10.0, 1.0, 0.2, 0.1, 1.0, 5.0, 2.0, 0.1, 0.05, 0.0, 3.0, 1.0, 0.2, 0.2, 1.0]

# ---- field_match ngram / term / bm25 (S/feature/FieldMatchFeature.scala).  The reference's tests tokenize with
# Lucene's English analyzer; tokenisation stays with the caller here, so the cases either use the `whitespace`
# analyzer on strings the English analyzer leaves untouched, or carry the token lists the reference's tests state.
_TITLE_MATCH = dict(name="title_match", type="field_match", rankingField="ranking.query", itemField="item.title",
                    method=dict(type="ngram", n=3, language="whitespace"))
case("field_match_ngram_score", "T/feature/FieldMatchFeatureTest.scala:61-68 (puts :49-59: [bar, foo, oba, oob])",
     [_TITLE_MATCH], [item_event("p1", [("title", "foobar")])],
     ranking(["p1", "p2"], [("query", "foo")]), {"title_match": [[0.25], [0.0]]})
case("field_match_ngram_no_query_field", "S/feature/FieldMatchFeature.scala:91 (None -> 0 for every item)",
     [_TITLE_MATCH], [item_event("p1", [("title", "foobar")])], ranking(["p1"]), {"title_match": [[0.0]]})
case("field_match_ngram_duplicates_and_string_list",
     "T/feature/matcher/NgramMatcherTest.scala:11-14 ('fooba foo' -> foo, oba, oob); StringListField joined by ' ' (S :46)",
     [_TITLE_MATCH], [item_event("p1", [("title", ["fooba", "foo"])]), item_event("p2", [("title", 7.0)])],
     ranking(["p1", "p2"], [("query", "oob zzz")]), {"title_match": [[0.25], [0.0]]})
_TERM = dict(name="tm", type="field_match", rankingField="ranking.query", itemField="item.title",
             method=dict(type="term", language="en"))
_term_item = item_event("p1", [("title", "greetings to hamsters!")])
_term_item["tokens"] = {"tm": ["greet", "hamster"]}            # TermMatcherTest.scala:11-14
_term_req = ranking(["p1"], [("query", "greet")])
_term_req["tokens"] = {"tm": ["greet"]}
case("field_match_term_half_match", "T/feature/matcher/TermMatcherTest.scala:11-23 (tokens as stated there; {a} vs {a,b} = 0.5)",
     [_TERM], [_term_item], _term_req, {"tm": [[0.5]]})
_BM25 = dict(name="bm", type="field_match", rankingField="ranking.query", itemField="item.title",
             method=dict(type="bm25", language="en", docs=3, avgdl=3.0, termfreq={"foo": 1, "bar": 2, "baz": 3}))
_bm_items = [item_event("p1", [("title", "bar baz")]), item_event("p2", [("title", "foo")])]
_bm_items[0]["tokens"] = {"bm": ["bar", "baz"]}
_bm_items[1]["tokens"] = {"bm": ["foo"]}
_bm_req = ranking(["p1", "p2", "p3"], [("query", "baz")])
_bm_req["tokens"] = {"bm": ["baz"]}
case("field_match_bm25_high_freq_query", "T/feature/matcher/BM25MatcherTest.scala:11-23 (0.15 +- 0.01)",
     [_BM25], _bm_items, _bm_req, {})
_bm_req2 = ranking(["p2", "p1"], [("query", "foo")])
_bm_req2["tokens"] = {"bm": ["foo"]}
case("field_match_bm25_low_freq_query", "T/feature/matcher/BM25MatcherTest.scala:25-27 (1.34 +- 0.01)",
     [_BM25], _bm_items, _bm_req2, {})
