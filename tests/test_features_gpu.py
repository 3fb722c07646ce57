"""GPU parity of feature assembly + the full rerank path (through the C ABI) against the
CPU oracle and the reference's golden vectors.  Bit-exact: the assembled matrix is compared
with array_equal (NaN == NaN), scores likewise, ordering identical."""
import numpy as np
import pytest

import golden_cases as G
from metarank_b200 import synth
from oracle import features_oracle as fo
from oracle import oracle

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), equal_nan=True)


def _device(ctx, features, model_features, state):
    from metarank_b200 import features as F

    fm = F.FeatureMapping(ctx, features, model_features)
    ds = F.DeviceState(ctx, fm)
    applied, skipped = ds.put(state)
    ds.flush()
    return fm, ds, F.Ranker(fm, ds), applied, skipped


@pytest.mark.parametrize("case", G.CASES, ids=[c["name"] for c in G.CASES])
def test_golden_vectors_on_device(ctx, case):
    mapping = fo.FeatureMapping(case["features"], case["model_features"])
    state = fo.FeatureValueFlow(mapping, always_refresh=True).process(case["events"])
    want = fo.dense_matrix(mapping, case["request"], state)
    fm, ds, rk, applied, _ = _device(ctx, case["features"], case["model_features"], state)
    try:
        assert fm.dim == mapping.dim
        got = rk.make_query([case["request"]])[0]
        assert _eq(got, want), (case["ref"], got, want)
        for name, exp in case["expected"].items():  # and directly against the reference's numbers
            o, d = fm.offset(name)
            assert _eq(got[:, o:o + d], np.array(exp)), (case["ref"], name)
    finally:
        ds.free(); fm.free()


def test_dense_layout_columns(ctx):
    from metarank_b200 import features as F

    fm = F.FeatureMapping(ctx, G.LAYOUT_FEATURES, [f["name"] for f in G.LAYOUT_FEATURES])
    assert fm.dim == 5  # T/flow/ClickthroughQueryTest.scala:152
    assert [fm.offset(n) for n in ("price", "category", "ctr", "clicked_category")] == [(0, 1), (1, 1), (2, 2), (4, 1)]
    fm.free()


@pytest.mark.parametrize("n_items,per_request", [(300, 24), (2000, 1000)])
def test_ranklens_shaped_assembly_and_rerank(ctx, n_items, per_request):
    """BASELINE config #3 shape (1000-item requests over the ranklens feature set, 24 columns)."""
    import metarank_b200 as mb

    feats, model = synth.ranklens_config()
    state, item_ids, sessions = synth.ranklens_state(n_items=n_items, n_sessions=40, seed=45)
    reqs = synth.ranklens_requests(item_ids, sessions, 6, per_request, seed=46)
    reqs.append(dict(reqs[0], session=None, user=None))  # no visitor
    reqs.append(dict(reqs[1], items=reqs[1]["items"][:1]))  # single item
    mapping = fo.FeatureMapping(feats, model)
    assert mapping.dim == 24
    fm, ds, rk, applied, skipped = _device(ctx, feats, model, state)
    blob = synth.lightgbm_model_text(60, 24, seed=9, cat_features={7: 16})
    booster = mb.LightGBMBooster(ctx, blob, n_features=24)
    ob = oracle.OracleBooster(0, blob)
    try:
        assert fm.dim == 24 and applied > 0
        got = rk.make_query(reqs)
        resp = rk.rerank(reqs, booster, explain=True)
        for r, q in enumerate(reqs):
            want = fo.dense_matrix(mapping, q, state)
            assert _eq(got[r], want), f"request {r}: first diff col {np.argwhere(~((got[r] == want) | ((got[r] != got[r]) & (want != want))))[:5]}"
            scores = ob.predictMat(want, *want.shape)
            order = oracle.rank_order(scores)
            assert [e["item"] for e in resp[r]["items"]] == [q["items"][k]["id"] for k in order]
            assert _eq([e["score"] for e in resp[r]["items"]], scores[order])
            assert _eq([e["features"] for e in resp[r]["items"]], want[order])
    finally:
        booster.free(); ds.free(); fm.free()


def test_cosine_biencoder_c4(ctx):
    """BASELINE config #4 shape: cosine(384-d) + scalar columns, 256 items, XGBoost scorer."""
    import metarank_b200 as mb

    rng = np.random.Generator(np.random.PCG64(77))
    dim = 384
    feats = [dict(name="sim", type="field_match", rankingField="ranking.query", itemField="item.title",
                  method=dict(type="bi-encoder", dim=dim), distance="cos")]
    feats += [dict(name=f"n{k}", type="number", scope="item", source=f"metadata.n{k}") for k in range(15)]
    model = [f["name"] for f in feats]
    state, ids = {}, [f"i{k}" for k in range(300)]
    for it in ids:
        e = rng.standard_normal(dim).astype(np.float32)
        e /= np.linalg.norm(e)
        if rng.random() > 0.05:
            state[(("item", it), "sim")] = ("scalar", e.astype(np.float64).tolist())
        for k in range(15):
            state[(("item", it), f"n{k}")] = ("scalar", float(rng.standard_normal()))
    for norm in ("noop", "linear", "position"):
        feats[0] = dict(feats[0], norm=norm)
        reqs = []
        for r in range(3):
            pick = [ids[int(j)] for j in rng.choice(300, 256, replace=False)]
            q = rng.standard_normal(dim).astype(np.float32)
            reqs.append(dict(event="ranking", id=f"r{r}", timestamp=0, user=None, session=None,
                             fields=[("query", "some text")], embeddings={"sim": q.tolist()},
                             items=[dict(id=i, fields=[]) for i in pick]))
        reqs.append(dict(reqs[0], embeddings={}))  # no query embedding -> NaN column
        mapping = fo.FeatureMapping(feats, model)
        fm, ds, rk, _, _ = _device(ctx, feats, model, state)
        blob = synth.xgboost_model_json(200, 16, depth=6, seed=1234 + 4)
        bx = mb.XGBoostBooster(ctx, blob, n_features=16)
        try:
            got = rk.make_query(reqs)
            resp = rk.rerank(reqs, bx)
            for r, q in enumerate(reqs):
                want = fo.dense_matrix(mapping, q, state)
                assert _eq(got[r], want), (norm, r)
                scores = oracle.OracleBooster(1, blob).predictMat(want, *want.shape)
                order = oracle.rank_order(scores)
                assert [e["item"] for e in resp[r]["items"]] == [q["items"][k]["id"] for k in order]
        finally:
            bx.free(); ds.free(); fm.free()


def test_scopes_user_session_global_and_request_inputs(ctx):
    feats = [
        dict(name="u_clicks", type="interaction_count", interaction="click", scope="user"),
        dict(name="s_win", type="window_count", interaction="click", scope="session", bucket="24h", periods=[1, 7]),
        dict(name="g_num", type="number", scope="global", source="metadata.x"),
        dict(name="q_len", type="word_count", scope="ranking", source="ranking.query"),
        dict(name="q_num", type="number", scope="ranking", source="ranking.boost"),
        dict(name="q_cat", type="string", scope="item", source="ranking.kind", encode="index", values=["web", "app"]),
        dict(name="q_hot", type="string", scope="item", source="ranking.kind", values=["web", "app", "tv"]),
        dict(name="color", type="string", scope="item", source="item.color", encode="index", values=["red", "green"]),
        dict(name="rel", type="relevancy"),
        dict(G.RATE, name="rq", scope="ranking.query"),
    ]
    model = [f["name"] for f in feats]
    state = {
        (("user", "u1"), "u_clicks"): ("counter", 7), (("session", "s1"), "s_win"): ("pcounter", [2, 9]),
        (("session", "s2"), "s_win"): ("pcounter", [2]),  # wrong length -> NaN
        (("global",), "g_num"): ("scalar", 3.5), (("item", "a"), "color"): ("scalar", ["green"]),
        (("irf", "query", "shoes", "a"), "rq_click"): ("pcounter", [1, 2]),
        (("irf", "query", "shoes", "a"), "rq_impression"): ("pcounter", [4, 4]),
    }
    reqs = [
        dict(event="ranking", id="r1", timestamp=0, user="u1", session="s1",
             fields=[("query", "shoes"), ("boost", 2.0), ("kind", "app")],
             items=[dict(id="a", fields=[("relevancy", 0.5)]), dict(id="b", fields=[("color", "red")]), dict(id="c", fields=[])]),
        dict(event="ranking", id="r2", timestamp=0, user="nobody", session="s2", fields=[("query", "  red   shoes ")],
             items=[dict(id="a", fields=[])]),
        dict(event="ranking", id="r3", timestamp=0, user=None, session=None, fields=[("kind", ["tv", "web"])],
             items=[dict(id="a", fields=[]), dict(id="zz", fields=[])]),
    ]
    mapping = fo.FeatureMapping(feats, model)
    fm, ds, rk, _, _ = _device(ctx, feats, model, state)
    try:
        got = rk.make_query(reqs)
        for r, q in enumerate(reqs):
            assert _eq(got[r], fo.dense_matrix(mapping, q, state)), r
    finally:
        ds.free(); fm.free()


def test_normalized_rate_zero_global_is_arithmetic_error(ctx):
    import metarank_b200 as mb

    feats = [dict(G.RATE, normalize={"weight": 10})]
    state = {(("item", "p1"), "ctr_click"): ("pcounter", [0, 0]), (("item", "p1"), "ctr_impression"): ("pcounter", [3, 3]),
             (("global",), "ctr_click_norm"): ("pcounter", [0, 0]), (("global",), "ctr_impression_norm"): ("pcounter", [3, 3])}
    fm, ds, rk, _, _ = _device(ctx, feats, ["ctr"], state)
    try:
        with pytest.raises(mb.MrError) as e:  # java.lang.ArithmeticException -> HTTP 500 in the reference
            rk.make_query([G.ranking(["p1"])])
        assert e.value.status == 7
        assert _eq(rk.make_query([G.ranking(["p2"])])[0], [[np.nan, np.nan]])  # item without state never divides
    finally:
        ds.free(); fm.free()


def test_state_updates_are_visible_after_flush(ctx):
    feats = [dict(name="price", type="number", scope="item", source="metadata.price")]
    fm, ds, rk, _, _ = _device(ctx, feats, ["price"], {(("item", "p1"), "price"): ("scalar", 1.0)})
    try:
        assert _eq(rk.make_query([G.ranking(["p1", "p2"])])[0], [[1.0], [np.nan]])
        ds.put({(("item", "p2"), "price"): ("scalar", 2.0), (("item", "p1"), "price"): ("scalar", 5.0),
                (("item", "p1"), "unrelated"): ("scalar", 9.0)})
        ds.flush()
        assert _eq(rk.make_query([G.ranking(["p1", "p2"])])[0], [[5.0], [2.0]])
        many = {(("item", f"x{i}"), "price"): ("scalar", float(i)) for i in range(5000)}  # forces table growth
        ds.put(many)
        ds.flush()
        got = rk.make_query([G.ranking([f"x{i}" for i in range(0, 5000, 7)] + ["p1"])])[0]
        assert _eq(got[:, 0], [float(i) for i in range(0, 5000, 7)] + [5.0])
        assert ds.info().rows[1] == 5002
        # a few scattered rows of a large table: the scatter-kernel path of mr_state_flush
        ds.put({(("item", "x3"), "price"): ("scalar", -3.0), (("item", "x4000"), "price"): ("scalar", -4000.0)})
        ds.flush()
        got = rk.make_query([G.ranking(["x3", "x4", "x4000", "x4999", "p2"])])[0]
        assert _eq(got[:, 0], [-3.0, 4.0, -4000.0, 4999.0, 2.0])
    finally:
        ds.free(); fm.free()


def test_sparse_flush_of_embedding_rows(ctx):
    feats = [dict(name="sim", type="field_match", rankingField="ranking.query", itemField="item.title",
                  method=dict(type="bi-encoder", dim=16), distance="cos")]
    rng = np.random.Generator(np.random.PCG64(3))
    state = {(("item", f"i{k}"), "sim"): ("scalar", rng.standard_normal(16)) for k in range(400)}
    fm, ds, rk, _, _ = _device(ctx, feats, ["sim"], state)
    mapping = fo.FeatureMapping(feats, ["sim"])
    try:
        upd = {(("item", "i7"), "sim"): ("scalar", rng.standard_normal(16)), (("item", "i390"), "sim"): ("scalar", rng.standard_normal(16))}
        state.update(upd)
        ds.put(upd)
        ds.flush()
        req = dict(event="ranking", id="r", timestamp=0, user=None, session=None, fields=[("query", "q")],
                   embeddings={"sim": rng.standard_normal(16).astype(np.float32)},
                   items=[dict(id=f"i{k}", fields=[]) for k in (6, 7, 8, 389, 390, 391)])
        assert _eq(rk.make_query([req])[0], fo.dense_matrix(mapping, req, state))
    finally:
        ds.free(); fm.free()


def test_unsupported_feature_types_fail_loudly(ctx):
    import metarank_b200 as mb
    from metarank_b200 import features as F

    with pytest.raises(mb.MrError) as e:
        F.FeatureMapping(ctx, [dict(name="ua", type="ua", field="platform", source="ranking.ua")], ["ua"])
    assert e.value.status == 5 and "ua" in e.value.message


def test_rank_device_api_matches_host_api(ctx):
    """mr_rank_device (device pointers, caller's stream) == mr_rank (host buffers), with and without
    the explain matrix, for the binned (LightGBM) and the f32 (XGBoost) scorer."""
    import torch

    import metarank_b200 as mb
    from metarank_b200 import features as F

    feats, model = synth.ranklens_config()
    state, item_ids, sessions = synth.ranklens_state(n_items=500, n_sessions=30, seed=5)
    reqs = synth.ranklens_requests(item_ids, sessions, 40, 100, seed=6)
    fm, ds, rk, _, _ = _device(ctx, feats, model, state)
    arrays = fm.pack_requests(reqs)
    N, R = arrays["total_items"], arrays["n_requests"]
    d = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int64) if v.dtype == np.uint64 else np.ascontiguousarray(v)).cuda()
         for k, v in arrays.items() if isinstance(v, np.ndarray)}
    st = torch.cuda.current_stream().cuda_stream
    for booster in (mb.LightGBMBooster(ctx, synth.lightgbm_model_text(80, 24, seed=3, cat_features={7: 16})),
                    mb.XGBoostBooster(ctx, synth.xgboost_model_json(60, 24, depth=5, seed=4))):
        want_s, want_o, want_f = rk.rank_arrays(arrays, booster, want_order=True, want_features=True)
        for explain in (False, True):
            d_s = torch.zeros(N, dtype=torch.float64, device="cuda")
            d_o = torch.zeros(N, dtype=torch.int32, device="cuda")
            d_f = torch.zeros(N * fm.dim, dtype=torch.float64, device="cuda")
            F.rank_device(ds, booster, R, N, d["offsets"].data_ptr(), d["ids"].data_ptr(), d_s.data_ptr(), d_o.data_ptr(),
                          d_f.data_ptr() if explain else 0, st, d["users"].data_ptr(), d["sessions"].data_ptr())
            F.rank_device_status(ds, st)
            assert _eq(d_s.cpu().numpy(), want_s) and np.array_equal(d_o.cpu().numpy(), want_o)
            if explain:
                assert _eq(d_f.cpu().numpy().reshape(N, fm.dim), want_f)
        booster.free()
    ds.free(); fm.free()


def test_large_batch_is_sliced_and_pipelined(ctx):
    """A batch beyond the 256 K-item slice size goes through several lanes; results must not depend on it."""
    import metarank_b200 as mb
    from metarank_b200 import features as F

    names = [f"f{j}" for j in range(8)]
    feats = [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names]
    fm = F.FeatureMapping(ctx, feats, names)
    ds = F.DeviceState(ctx, fm)
    cat = synth.feature_matrix(5000, 8, seed=1)
    ids = np.arange(1, 5001, dtype=np.uint64) * np.uint64(2654435761)
    ds.put_packed(F.pack_number_columns(names, ids, cat))
    ds.flush()
    rng = np.random.Generator(np.random.PCG64(2))
    sizes = rng.integers(1, 400, 2000)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    pick = rng.integers(0, 5000, int(offs[-1]))
    R = len(sizes)
    arrays = dict(offsets=offs, ids=ids[pick], users=np.zeros(R, dtype=np.uint64), sessions=np.zeros(R, dtype=np.uint64),
                  req_f64=np.zeros((R, 1)), req_u64=np.zeros((R, 1), dtype=np.uint64), req_vec=np.zeros((R, 1), dtype=np.float32),
                  req_vp=np.zeros((R, 1), dtype=np.uint8), item_f64=None, n_requests=R, total_items=int(offs[-1]))
    blob = synth.lightgbm_model_text(40, 8, seed=3)
    booster = mb.LightGBMBooster(ctx, blob)
    scores, order, feats_out = F.Ranker(fm, ds).rank_arrays(arrays, booster, want_order=True, want_features=True)
    want = oracle.OracleBooster(0, blob).predictMat(cat[pick], len(pick), 8, threads=0)
    assert _eq(feats_out, cat[pick]) and _eq(scores, want)
    for r in rng.integers(0, R, 50):
        assert np.array_equal(order[offs[r]:offs[r + 1]], oracle.rank_order(want[offs[r]:offs[r + 1]]))
    booster.free(); ds.free(); fm.free()


@pytest.mark.parametrize("n_feat", [63, 64, 100, 127, 140])
def test_wide_item_rows_take_the_coalesced_gather(ctx, n_feat):
    """Item rows of 65..128 words (64+ scalar features, BASELINE config #5) are gathered by the warp-per-row
    kernels with up to four coalesced loads (code rows on the scored path, dense rows with explain); wider rows
    fall back to the generic per-item kernel.  All three must equal the stored scalars / the oracle's scores."""
    import metarank_b200 as mb
    from metarank_b200 import features as F

    names = [f"f{j}" for j in range(n_feat)]
    fm = F.FeatureMapping(ctx, [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names], names)
    ds = F.DeviceState(ctx, fm)
    cat = synth.feature_matrix(700, n_feat, seed=n_feat)
    ids = np.arange(1, 701, dtype=np.uint64) * np.uint64(40503)
    for j0 in range(0, n_feat, 16):  # several upserts: rows are merged column group by column group
        ds.put_packed(F.pack_number_columns(names[j0:j0 + 16], ids, cat[:, j0:j0 + 16]))
    ds.flush()
    blob = synth.lightgbm_model_text(60, n_feat, seed=11)
    booster = mb.LightGBMBooster(ctx, blob)
    rng = np.random.Generator(np.random.PCG64(n_feat))
    pick = rng.integers(0, 700, 5000)
    offs = np.arange(0, 5001, 250, dtype=np.int32)
    R = len(offs) - 1
    arrays = dict(offsets=offs, ids=ids[pick], users=np.zeros(R, dtype=np.uint64), sessions=np.zeros(R, dtype=np.uint64),
                  req_f64=np.zeros((R, 1)), req_u64=np.zeros((R, 1), dtype=np.uint64), req_vec=np.zeros((R, 1), dtype=np.float32),
                  req_vp=np.zeros((R, 1), dtype=np.uint8), item_f64=None, n_requests=R, total_items=5000)
    rk = F.Ranker(fm, ds)
    want = oracle.OracleBooster(0, blob).predictMat(cat[pick], 5000, n_feat)
    scores, order, _ = rk.rank_arrays(arrays, booster)                               # code rows -> code gather
    assert _eq(scores, want)
    scores2, _, feats = rk.rank_arrays(arrays, booster, want_features=True)          # explain: dense rows
    assert _eq(feats, cat[pick]) and _eq(scores2, want)
    for r in range(R):
        assert np.array_equal(order[offs[r]:offs[r + 1]], oracle.rank_order(want[offs[r]:offs[r + 1]]))
    booster.free(); ds.free(); fm.free()


def test_rank_api_post_rank_like_reference_rank_api_test(ctx):
    """POST /rank/<model>?explain= in-process, as T/main/api/RankApiTest.scala:36-66 does."""
    import json

    import metarank_b200 as mb
    from metarank_b200 import rank_api as ra

    feats, model = synth.ranklens_config()
    state, item_ids, sessions = synth.ranklens_state(n_items=200, n_sessions=10, seed=5)
    fm, ds, rk, _, _ = _device(ctx, feats, model, state)
    blob = synth.lightgbm_model_text(40, 24, seed=3, cat_features={7: 16})
    booster = mb.LightGBMBooster(ctx, blob, n_features=24)
    api = ra.RankApi({"xgboost": (fm, ds, booster)})
    payload = json.dumps({"event": "ranking", "id": "r1", "timestamp": "1599391467000", "user": sessions[0],
                          "session": sessions[0], "items": [{"id": i, "relevancy": 0} for i in item_ids[:24]]})
    req = ra.decode_ranking_event(payload)
    mapping = fo.FeatureMapping(feats, model)
    want = fo.dense_matrix(mapping, req, state)
    scores = oracle.OracleBooster(0, blob).predictMat(want, *want.shape)
    order = oracle.rank_order(scores)
    for explain in ("true", "false"):
        status, ctype, body = api.routes("POST", f"/rank/xgboost?explain={explain}", payload)
        assert status == 200 and ctype == "application/json"
        resp = json.loads(body)
        assert [e["item"] for e in resp["items"]] == [item_ids[k] for k in order]
        assert [e["score"] for e in resp["items"]] == [float(scores[k]) for k in order]
        assert ("state" in resp) == (explain == "true")
        assert all(("features" in e) == (explain == "true") for e in resp["items"])
    # explain payload: names re-attached from the DatasetDescriptor, NaN dropped, category as "cat@index"
    e0 = json.loads(api.routes("POST", "/rank/xgboost?explain=true", payload)[2])["items"][0]
    k0 = int(order[0])
    assert e0["features"]["position"] == 5.0 and len(e0["features"]["profile"]) == 4
    assert e0["features"]["genre"].endswith("@%d" % int(want[k0, 7]))
    assert api.routes("POST", "/rank/missing", payload)[0] == 500
    booster.free(); ds.free(); fm.free()


@pytest.mark.parametrize("case", [c for c in G.CASES if c["events"]], ids=[c["name"] for c in G.CASES if c["events"]])
def test_native_write_path_replays_reference_tests(ctx, case):
    """the extractors' raw writes (Put / Increment / PeriodicIncrement / Append) go to
    mr_state_apply_writes instead of refreshed FeatureValues; the library keeps MemCounter /
    MemPeriodicCounter+fromMap / MemBoundedList semantics itself.  Same golden vectors as the read path."""
    from metarank_b200 import features as F

    mapping = fo.FeatureMapping(case["features"], case["model_features"])
    flow = fo.FeatureValueFlow(mapping, always_refresh=True)
    state = flow.process(case["events"])
    fm = F.FeatureMapping(ctx, case["features"], case["model_features"])
    ds = F.DeviceState(ctx, fm)
    try:
        applied, skipped = ds.apply_writes(flow.write_log)
        ds.flush()
        got = F.Ranker(fm, ds).make_query([case["request"]])[0]
        assert _eq(got, fo.dense_matrix(mapping, case["request"], state)), case["ref"]
        for name, exp in case["expected"].items():
            o, d = fm.offset(name)
            assert _eq(got[:, o:o + d], np.array(exp)), (case["ref"], name)
    finally:
        ds.free(); fm.free()


def test_native_write_path_random_event_stream(ctx):
    """Periodic counters over many days (window anchoring at the last bucket, out-of-order events),
    bounded lists beyond count/duration, counters in user/session scope — incremental visibility too."""
    from metarank_b200 import features as F

    rng = np.random.Generator(np.random.PCG64(11))
    feats = [
        dict(G.RATE, name="ctr", periods=[1, 7, 30], normalize={"weight": 10}),
        dict(name="wc", type="window_count", interaction="click", scope="item", bucket="24h", periods=[1, 3]),
        dict(name="uclicks", type="interaction_count", interaction="click", scope="user"),
        dict(name="seen", type="interacted_with", interaction="click", field=["item.color", "item.tags"], scope="session",
             count=5, duration="48h"),
        dict(name="price", type="number", scope="item", source="metadata.price"),
    ]
    model = [f["name"] for f in feats]
    items = [f"p{i}" for i in range(12)]
    events = [G.item_event(it, [("color", ["red", "green", "blue"][i % 3]), ("tags", [f"t{i % 4}", f"t{(i * 7) % 5}"]),
                                ("price", float(i))]) for i, it in enumerate(items)]
    t = G.NOW - 40 * 86400_000
    for _ in range(600):
        t += int(rng.integers(0, 6 * 3600_000))
        ts = t - (int(rng.integers(0, 3 * 86400_000)) if rng.random() < 0.1 else 0)  # some late events
        typ = "click" if rng.random() < 0.3 else "impression"
        events.append(G.interaction(items[int(rng.integers(0, 12))], "r", typ, user=f"u{int(rng.integers(0, 3))}",
                                    session=f"s{int(rng.integers(0, 4))}", ts=ts))
    mapping = fo.FeatureMapping(feats, model)
    flow = fo.FeatureValueFlow(mapping, always_refresh=True)
    fm = F.FeatureMapping(ctx, feats, model)
    ds = F.DeviceState(ctx, fm)
    try:
        state, done = {}, 0
        for cut in (len(items), 100, 350, len(events)):
            state.update(flow.process(events[done:cut]))
            ds.apply_writes(flow.write_log)
            flow.write_log.clear()
            ds.flush()
            done = cut
            for sess, user in (("s0", "u0"), ("s3", "u2"), (None, None)):
                req = G.ranking(items + ["unknown"], user=user, session=sess)
                got = F.Ranker(fm, ds).make_query([req])[0]
                assert _eq(got, fo.dense_matrix(mapping, req, state)), (cut, sess)
    finally:
        ds.free(); fm.free()


def _random_config(rng):
    """A random feature set over every supported extractor kind and scope, plus state and requests
    salted with the awkward cases: missing keys, wrong value types, wrong lengths, unknown items."""
    scopes = ["item", "user", "session", "global"]
    feats, k = [], 0

    def nm(p):
        nonlocal k
        k += 1
        return f"{p}{k}"

    for _ in range(int(rng.integers(1, 4))):
        feats.append(dict(name=nm("num"), type="number", scope=str(rng.choice(scopes + ["ranking"])), source="item.price"))
    feats.append(dict(name=nm("wc"), type="word_count", scope=str(rng.choice(["item", "ranking"])), source="item.title"))
    feats.append(dict(name=nm("cat"), type="string", scope="item", source="item.color", encode="index", values=["red", "green", "blue"]))
    feats.append(dict(name=nm("hot"), type="string", scope=str(rng.choice(["item", "user"])), source="item.size", values=["s", "m", "l", "xl"]))
    feats.append(dict(name=nm("cnt"), type="interaction_count", interaction="click", scope=str(rng.choice(scopes))))
    feats.append(dict(name=nm("win"), type="window_count", interaction="click", scope=str(rng.choice(scopes)), bucket="24h",
                      periods=[1, 7, 30][: int(rng.integers(1, 4))]))
    feats.append(dict(name=nm("rate"), type="rate", top="click", bottom="imp", bucket="24h", periods=[7, 30],
                      **({"normalize": {"weight": float(rng.integers(1, 20))}} if rng.random() < 0.5 else {}),
                      **({"scope": str(rng.choice(["item", "item.color", "ranking.query"]))} if rng.random() < 0.8 else {})))
    feats.append(dict(name=nm("seen"), type="interacted_with", interaction="click", scope=str(rng.choice(["user", "session"])),
                      field=["item.tags", "item.color"][: int(rng.integers(1, 3))], count=20, duration="24h"))
    feats.append(dict(name=nm("rel"), type="relevancy"))
    feats.append(dict(name=nm("pos"), type="position", position=int(rng.integers(0, 9))))
    feats.append(dict(name=nm("divn"), type="diversity", source="item.price", top=int(rng.choice([1, 3, 20]))))
    feats.append(dict(name=nm("divs"), type="diversity", source="item.tags"))
    feats.append(dict(name=nm("flag"), type="boolean", scope=str(rng.choice(["item", "session"])), source="item.ok"))
    feats.append(dict(name=nm("vec"), type="vector", scope="item", source="item.vec", reduce=["min", "vector2", "avg"]))
    feats.append(dict(name=nm("age"), type="item_age", source="item.updated"))
    feats.append(dict(name=nm("tod"), type="local_time", source="ranking.timestamp", parse=str(rng.choice(["time_of_day", "day_of_week", "year"]))))
    feats.append(dict(name=nm("sim"), type="field_match", rankingField="ranking.query", itemField="item.title",
                      method=dict(type="bi-encoder", dim=8), norm=str(rng.choice(["noop", "linear", "position"]))))
    order = rng.permutation(len(feats))
    model = [feats[i]["name"] for i in order if rng.random() < 0.9]
    return feats, model


def _random_state(rng, feats, items, users, sessions):
    st = {}
    tags = [f"t{i}" for i in range(6)]

    def maybe(p=0.85):
        return rng.random() < p

    def pc(n):
        return ("pcounter", [int(x) for x in rng.integers(0, 50, n)])

    for f in feats:
        t, n = f["type"], f["name"]
        scope_ids = {"item": [("item", i) for i in items], "user": [("user", u) for u in users],
                     "session": [("session", s) for s in sessions], "global": [("global",)]}
        sc = f.get("scope", "item")
        if t == "number" and sc != "ranking":
            for s in scope_ids[sc]:
                if maybe():
                    st[(s, n)] = ("scalar", float(rng.standard_normal())) if maybe(0.9) else ("scalar", "oops")
        elif t == "word_count" and sc == "item":
            for s in scope_ids["item"]:
                if maybe():
                    st[(s, n)] = ("scalar", float(rng.integers(0, 9)))
        elif t == "string":
            for s in scope_ids[sc]:
                if maybe():
                    vals = list(rng.choice(f["values"] + ["zzz"], int(rng.integers(0, 3))))
                    st[(s, n)] = ("scalar", [str(v) for v in vals]) if maybe(0.9) else ("scalar", 3.0)
        elif t == "interaction_count":
            for s in scope_ids[sc]:
                if maybe():
                    st[(s, n)] = ("counter", int(rng.integers(0, 100)))
        elif t == "window_count":
            P = len(f["periods"])
            for s in scope_ids[sc]:
                if maybe():
                    st[(s, n)] = pc(P) if maybe(0.9) else pc(P + 1)
        elif t == "rate":
            scope = f.get("scope", "item")
            for name in (f"{n}_click_norm", f"{n}_imp_norm"):
                if maybe(0.95):
                    st[(("global",), name)] = ("pcounter", [int(x) for x in rng.integers(1, 500, 2)])
            for i in items:
                if scope == "item":
                    tgt = ("item", i)
                elif scope == "item.color":
                    col = str(rng.choice(["red", "green", "blue"]))
                    if maybe():
                        st[(("item", i), f"{n}_field")] = ("scalar", col)
                    tgt = ("field", "color", col)
                else:
                    tgt = ("irf", "query", str(rng.choice(["shoes", "hats"])), i)
                if maybe():
                    st[(tgt, f"{n}_click")] = pc(2)
                if maybe():
                    st[(tgt, f"{n}_imp")] = pc(2) if maybe(0.95) else pc(3)
        elif t == "interacted_with":
            for s in scope_ids[sc]:
                if maybe(0.7):
                    st[(s, f"{n}_interactions")] = ("blist", [str(x) for x in rng.choice(items + ["ghost"], int(rng.integers(0, 15)))])
            for fld in ([f["field"]] if isinstance(f["field"], str) else f["field"]):
                for i in items:
                    if maybe():
                        st[(("item", i), f"{n}_{fld.split('.')[1]}")] = ("scalar", [str(x) for x in rng.choice(tags, int(rng.integers(0, 4)))])
        elif t == "diversity":
            numeric = f["source"] == "item.price"
            for i in items:
                if maybe(0.8):
                    st[(("item", i), n)] = ("scalar", float(rng.integers(0, 9))) if numeric else \
                        ("scalar", [str(x) for x in rng.choice(tags, int(rng.integers(0, 4)))])
        elif t == "boolean":
            for s in scope_ids[sc]:
                if maybe():
                    st[(s, n)] = ("scalar", bool(rng.random() < 0.5))
        elif t == "vector":
            for i in items:
                if maybe():
                    st[(("item", i), n)] = ("scalar", [float(x) for x in rng.standard_normal(4)])
        elif t == "item_age":
            for i in items:
                if maybe():
                    st[(("item", i), n)] = ("scalar", float(1_600_000_000 + int(rng.integers(0, 10**7)) + rng.random()))
        elif t == "field_match":
            for i in items:
                if maybe():
                    st[(("item", i), n)] = ("scalar", rng.standard_normal(8).astype(np.float32).astype(np.float64))
    return st


@pytest.mark.parametrize("seed", range(12))
def test_random_feature_configs_match_oracle(ctx, seed):
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    items = [f"i{k}" for k in range(25)]
    users, sessions = ["u0", "u1"], ["s0", "s1", "s2"]
    feats, model = _random_config(rng)
    state = _random_state(rng, feats, items, users, sessions)
    for key in list(state):  # keep normalized rates away from a zero global top counter (-> ArithmeticException)
        if key[1].endswith("_click_norm"):
            state[key] = ("pcounter", [max(1, v) for v in state[key][1]])
    reqs = []
    for r in range(6):
        n = int(rng.integers(1, 40))
        picks = [str(x) for x in rng.choice(items + ["nope1", "nope2"], n)]  # duplicates and unknown items
        fields = [("query", str(rng.choice(["shoes", "hats", "socks"])))] if rng.random() < 0.8 else []
        if rng.random() < 0.5:
            fields.append(("price", float(rng.integers(0, 5))))
        if rng.random() < 0.5:
            fields.append(("title", "  some  words here "))
        sim = next(f["name"] for f in feats if f["type"] == "field_match")
        emb = {sim: rng.standard_normal(8).astype(np.float32)} if rng.random() < 0.8 else {}
        its = []
        for p in picks:
            fl = []
            if rng.random() < 0.7:
                fl.append(("relevancy", float(rng.integers(0, 4))))
            if rng.random() < 0.2:
                fl.append(("price", float(rng.integers(10, 20))))
            if rng.random() < 0.2:
                fl.append(("color", str(rng.choice(["red", "zzz"]))))
            its.append(dict(id=p, fields=fl))
        reqs.append(dict(event="ranking", id=f"r{r}", timestamp=1_650_000_000_000 + int(rng.integers(0, 10**9)),
                         user=str(rng.choice(users + ["ux"])) if rng.random() < 0.8 else None,
                         session=str(rng.choice(sessions + ["sx"])) if rng.random() < 0.8 else None,
                         fields=fields, embeddings=emb, items=its))
    mapping = fo.FeatureMapping(feats, model)
    fm, ds, rk, _, _ = _device(ctx, feats, model, state)
    try:
        assert fm.dim == mapping.dim
        got = rk.make_query(reqs)
        for r, q in enumerate(reqs):
            want = fo.dense_matrix(mapping, q, state)
            if not _eq(got[r], want):
                bad = np.argwhere(~((got[r] == want) | ((got[r] != got[r]) & (want != want))))
                cols = {n: mapping.offsets[n] for n in mapping.offsets}
                raise AssertionError(f"seed {seed} request {r}: mismatch at {bad[:6].tolist()} cols={cols}\n"
                                     f"got={got[r][bad[0][0]].tolist()}\nwant={want[bad[0][0]].tolist()}")
    finally:
        ds.free(); fm.free()


def test_two_contexts_in_one_process_on_two_gpus():
    """A JVM drives all GPUs of the box from one process: one mr_ctx per device, requests round-robin.
    Skipped on single-GPU boxes."""
    import torch

    import metarank_b200 as mb

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    blob = synth.lightgbm_model_text(120, 30, seed=5)
    X = synth.feature_matrix(5000, 30, seed=6)
    want = oracle.OracleBooster(0, blob).predictMat(X, *X.shape)
    ctxs = [mb.Context(0), mb.Context(1)]
    boosters = [mb.LightGBMBooster(c, blob) for c in ctxs]
    import threading

    out = [None, None]

    def work(k):
        for _ in range(5):
            out[k] = boosters[k].predictMat(X, *X.shape)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert np.array_equal(out[0], want) and np.array_equal(out[1], want)
    for b in boosters:
        b.free()
    for c in ctxs:
        c.close()


def test_ordering_edge_sizes_ties_and_specials(ctx):
    """Per-request ordering across the four device code paths (warp bitonic network <= 128 items, CTA bitonic
    sort <= 4096, chip-wide rank by counting <= 20 000, chunk sort + merge beyond), with heavy ties and NaN scores; through mr_rank
    (host offsets -> size hint) and mr_rank_device (no hint).  Oracle: stable sortBy(-score),
    S/ranking/Ranker.scala:58-60."""
    import torch

    import metarank_b200 as mb
    from metarank_b200 import features as F

    names = ["a", "b"]
    fm = F.FeatureMapping(ctx, [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names], names)
    ds = F.DeviceState(ctx, fm)
    n_cat = 6000
    rng = np.random.Generator(np.random.PCG64(11))
    cat = np.stack([rng.integers(0, 4, n_cat).astype(np.float64), rng.integers(0, 3, n_cat).astype(np.float64)], axis=1)
    ids = np.arange(1, n_cat + 1, dtype=np.uint64) * np.uint64(11400714819323198485)
    ds.put_packed(F.pack_number_columns(names, ids, cat)); ds.flush()
    # two depth-1 trees with leaves that produce ties, a NaN score, and both zeros
    model = "\n".join([
        "tree", "version=v3", "num_class=1", "num_tree_per_iteration=1", "label_index=0", "max_feature_idx=1",
        "objective=lambdarank", "feature_names=a b", "feature_infos=[0:3] [0:2]", "tree_sizes=0 0", "",
        "Tree=0", "num_leaves=3", "num_cat=0", "split_feature=0 0", "split_gain=1 1", "threshold=0.5 1.5",
        "decision_type=2 2", "left_child=-1 -2", "right_child=1 -3", "leaf_value=0.0 -0.0 1.0", "leaf_weight=1 1 1",
        "leaf_count=1 1 1", "internal_value=0 0", "internal_weight=0 0", "internal_count=3 2", "is_linear=0", "shrinkage=1", "",
        "Tree=1", "num_leaves=3", "num_cat=0", "split_feature=1 1", "split_gain=1 1", "threshold=0.5 1.5",
        "decision_type=2 2", "left_child=-1 -2", "right_child=1 -3", "leaf_value=0.0 -0.0 nan", "leaf_weight=1 1 1",
        "leaf_count=1 1 1", "internal_value=0 0", "internal_weight=0 0", "internal_count=3 2", "is_linear=0", "shrinkage=1", "",
        "end of trees", ""]).encode()
    booster = mb.LightGBMBooster(ctx, model)
    sizes = np.array([0, 1, 2, 3, 31, 32, 33, 100, 127, 128, 129, 130, 500, 1024, 4095, 4096, 4097, 5000, 0, 7, 9001, 4100, 1,
                      20000, 20001, 6, 23000], dtype=np.int64)  # (4096, 20000]: rank by counting; beyond: chunk sort + merge
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    N, R = int(offs[-1]), len(sizes)
    pick = rng.integers(0, n_cat, N)
    arrays = dict(offsets=offs, ids=ids[pick], users=np.zeros(R, dtype=np.uint64), sessions=np.zeros(R, dtype=np.uint64),
                  req_f64=np.zeros((R, 1)), req_u64=np.zeros((R, 1), dtype=np.uint64), req_vec=np.zeros((R, 1), dtype=np.float32),
                  req_vp=np.zeros((R, 1), dtype=np.uint8), item_f64=None, n_requests=R, total_items=N)
    want_s = oracle.OracleBooster(0, model).predictMat(cat[pick], N, 2, threads=0)
    assert np.isnan(want_s).any() and (want_s == 0).any()  # (a sum that starts at +0.0 never yields -0.0)
    want_o = np.concatenate([oracle.rank_order(want_s[offs[r]:offs[r + 1]]) for r in range(R)]).astype(np.int32)
    scores, order, _ = F.Ranker(fm, ds).rank_arrays(arrays, booster, want_order=True)
    assert _eq(scores, want_s) and np.array_equal(order, want_o)
    # small requests only: the hint lets mr_rank skip the CTA-wide sort
    small = sizes <= 128
    so = np.concatenate([[0], np.cumsum(sizes[small])]).astype(np.int32)
    sp = np.concatenate([pick[offs[r]:offs[r + 1]] for r in range(R) if small[r]])
    a2 = dict(arrays, offsets=so, ids=ids[sp], users=arrays["users"][:small.sum()], sessions=arrays["sessions"][:small.sum()],
              req_f64=np.zeros((small.sum(), 1)), req_u64=np.zeros((small.sum(), 1), dtype=np.uint64),
              req_vec=np.zeros((small.sum(), 1), dtype=np.float32), req_vp=np.zeros((small.sum(), 1), dtype=np.uint8),
              n_requests=int(small.sum()), total_items=int(so[-1]))
    s2, o2, _ = F.Ranker(fm, ds).rank_arrays(a2, booster, want_order=True)
    w2 = oracle.OracleBooster(0, model).predictMat(cat[sp], len(sp), 2, threads=0)
    assert np.array_equal(o2, np.concatenate([oracle.rank_order(w2[so[r]:so[r + 1]]) for r in range(len(so) - 1)]))
    # device entry point
    d_off = torch.from_numpy(offs).cuda(); d_ids = torch.from_numpy(ids[pick].view(np.int64)).cuda()
    d_s = torch.zeros(N, dtype=torch.float64, device="cuda"); d_o = torch.full((N,), -1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    F.rank_device(ds, booster, R, N, d_off.data_ptr(), d_ids.data_ptr(), d_s.data_ptr(), d_o.data_ptr(), 0, st)
    F.rank_device_status(ds, st)
    assert _eq(d_s.cpu().numpy(), want_s) and np.array_equal(d_o.cpu().numpy(), want_o)
    booster.free(); ds.free(); fm.free()


def test_code_rows_follow_state_updates(ctx):
    """The per-model code rows (materialised u16 codes of the item table, rank_api.cu CodeCache) must track
    the state: sparse updates (incremental refresh from the flush log), table growth (rebuild), more
    flushes than the log holds, two models sharing one state, a repacked model, missing values being
    filled in, and unknown items."""
    import metarank_b200 as mb
    from metarank_b200 import features as F

    names = [f"f{j}" for j in range(9)]
    feats = [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names]
    fm = F.FeatureMapping(ctx, feats, names)
    ds = F.DeviceState(ctx, fm)
    rng = np.random.Generator(np.random.PCG64(77))
    n0 = 3000
    cat = synth.feature_matrix(n0 + 4000, 9, seed=5)  # rows beyond n0 are added later
    ids = np.arange(1, len(cat) + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    ds.put_packed(F.pack_number_columns(names, ids[:n0], cat[:n0])); ds.flush()
    known = n0
    lgb_blob = synth.lightgbm_model_text(60, 9, seed=8)
    xgb_blob = synth.xgboost_model_json(40, 9, depth=5, seed=9)
    lgb, xgb = mb.LightGBMBooster(ctx, lgb_blob), mb.XGBoostBooster(ctx, xgb_blob)
    o_lgb, o_xgb = oracle.OracleBooster(0, lgb_blob), oracle.OracleBooster(1, xgb_blob)
    rk = F.Ranker(fm, ds)

    def check(booster, orc, n_req=40):
        sizes = rng.integers(1, 60, n_req)
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        N = int(offs[-1])
        pick = rng.integers(0, min(known + 50, len(ids)), N)  # some ids are not in the table yet: all-missing rows
        X = np.where((pick < known)[:, None], cat[pick], np.nan)
        R = n_req
        arrays = dict(offsets=offs, ids=ids[pick],
                      users=np.zeros(R, dtype=np.uint64), sessions=np.zeros(R, dtype=np.uint64), req_f64=np.zeros((R, 1)),
                      req_u64=np.zeros((R, 1), dtype=np.uint64), req_vec=np.zeros((R, 1), dtype=np.float32),
                      req_vp=np.zeros((R, 1), dtype=np.uint8), item_f64=None, n_requests=R, total_items=N)
        # ids beyond `known` but inside `ids` are simply not stored yet -> unknown items
        scores, order, _ = rk.rank_arrays(arrays, booster, want_order=True)
        want = orc.predictMat(np.ascontiguousarray(X), N, 9, threads=0)
        assert _eq(scores, want)
        for r in range(R):
            assert np.array_equal(order[offs[r]:offs[r + 1]], oracle.rank_order(want[offs[r]:offs[r + 1]]))

    check(lgb, o_lgb); check(xgb, o_xgb)
    # sparse update, NaN appearing / disappearing
    def put(rows, new):  # a NaN entry is "no write": the stored value stays (pack_number_columns skips it)
        ds.put_packed(F.pack_number_columns(names, ids[rows], new)); ds.flush()
        cat[rows] = np.where(np.isnan(new), cat[rows], new)

    rows = rng.choice(known, 25, replace=False)
    put(rows, synth.feature_matrix(25, 9, seed=6))   # also fills values that were missing
    check(lgb, o_lgb); check(xgb, o_xgb)
    # many small flushes between two ranks (more than the change log keeps)
    for k in range(70):
        r = int(rng.integers(0, known))
        put(np.array([r]), synth.feature_matrix(1, 9, seed=100 + k))
    check(lgb, o_lgb)
    # growth: new items (device rows reallocate)
    ds.put_packed(F.pack_number_columns(names, ids[known:known + 4000], cat[known:known + 4000])); ds.flush()
    known += 4000
    check(lgb, o_lgb); check(xgb, o_xgb)
    # a repacked model gets fresh rows
    lgb.set_option("chunk_kb", 8)
    check(lgb, o_lgb)
    lgb.set_option("variant", 2)   # generic binned scorer: identity tile mapping
    check(lgb, o_lgb)
    lgb.set_option("variant", -1)
    check(lgb, o_lgb)
    lgb.free(); xgb.free(); ds.free(); fm.free()


def test_request_item_fields_override_stored_values_in_the_scored_path(ctx):
    """NumberFeature.values (S/feature/NumberFeature.scala:84-93): a field on the ranking item wins over the
    stored value.  With a model attached the codes normally come from the per-model code rows; a batch that
    carries item fields for an overridable column must bypass them.  Scores with and without the explain
    matrix, against the oracle's feature rows."""
    import metarank_b200 as mb
    from metarank_b200 import features as F

    feats = [dict(name="price", type="number", scope="item", source="item.price"),
             dict(name="pop", type="number", scope="item", source="item.pop")]
    mapping = fo.FeatureMapping(feats, ["price", "pop"])
    events = [dict(event="item", id=f"e{i}", timestamp=G.NOW, item=f"p{i}", fields=[("price", float(i)), ("pop", float(10 - i))])
              for i in range(1, 9)]
    state = fo.FeatureValueFlow(mapping, always_refresh=True).process(events)
    fm, ds, rk, _, _ = _device(ctx, feats, ["price", "pop"], state)
    blob = synth.lightgbm_model_text(30, 2, seed=21)
    booster = mb.LightGBMBooster(ctx, blob)
    orc = oracle.OracleBooster(0, blob)
    plain = G.ranking([f"p{i}" for i in range(1, 9)] + ["unknown"])
    over = G.ranking([f"p{i}" for i in range(1, 9)] + ["unknown"],
                     item_fields={"p2": [("price", 100.0)], "p5": [("price", -3.5)], "unknown": [("price", 7.0)]})
    for req in (plain, over, plain):
        want_x = fo.dense_matrix(mapping, req, state)
        want_s = orc.predictMat(np.ascontiguousarray(want_x), len(want_x), 2, threads=0)
        arrays = fm.pack_requests([req])
        for explain in (False, True):
            scores, order, fx = rk.rank_arrays(arrays, booster, want_order=True, want_features=explain)
            assert _eq(scores, want_s), (req is over, explain)
            assert np.array_equal(order, oracle.rank_order(want_s))
            if explain:
                assert _eq(fx, want_x)
    x_over = fo.dense_matrix(mapping, over, state)
    assert x_over[1, 0] == 100.0 and x_over[4, 0] == -3.5 and x_over[8, 0] == 7.0 and np.isnan(x_over[8, 1])
    booster.free(); ds.free(); fm.free()


def _state_to_feature_values(state, legacy=False):
    """Oracle state {(scope, name): (kind, value)} -> FeatureValue dicts for oracle/codec_oracle.py."""
    out = []
    for n, ((scope, name), (kind, v)) in enumerate(state.items()):
        key = (tuple(scope), name)
        if kind == "scalar":
            if isinstance(v, (list, tuple, np.ndarray)) and len(v) == 0:
                v = ("strings", [])
            elif isinstance(v, np.ndarray):
                v = [float(x) for x in v]
            out.append(dict(type="scalar", key=key, ts=1000 + n, value=v, expire_ms=86400000))
        elif kind == "counter":
            out.append(dict(type="counter", key=key, ts=1000 + n, value=int(v), expire_ms=86400000))
        elif kind == "pcounter":
            out.append(dict(type="pcounter", key=key, ts=1000 + n, expire_ms=86400000,
                            values=[dict(start=0, end=86400000, periods=j + 1, value=int(x)) for j, x in enumerate(v)]))
        elif kind == "blist":
            out.append(dict(type="blist", key=key, ts=1000 + n, expire_ms=86400000,
                            values=[(5000 - j, x) for j, x in enumerate(v)]))
        else:
            raise ValueError(kind)
    return out


@pytest.mark.parametrize("legacy", [False, True])
def test_state_loaded_from_the_reference_binary_format(ctx, legacy):
    """a state exported in the reference's binary store format (delimited FeatureValueCodec
    records, S/fstore/codec/impl/FeatureValueCodec.scala:41-237) loads natively and ranks exactly like the same
    state put through mr_state_upsert — ranklens feature set: scalars, string lists, counters, periodic
    counters, bounded lists, every scope."""
    from metarank_b200 import features as F
    from oracle import codec_oracle as co

    feats, model = synth.ranklens_config()
    state, item_ids, sessions = synth.ranklens_state(n_items=300, n_sessions=20, seed=15)
    reqs = synth.ranklens_requests(item_ids, sessions, 12, 40, seed=16)
    blob = co.encode_delimited(_state_to_feature_values(state, legacy), legacy)
    mapping = fo.FeatureMapping(feats, model)
    fm = F.FeatureMapping(ctx, feats, model)
    ds = F.DeviceState(ctx, fm)
    applied, skipped, consumed = ds.load_feature_values(blob)
    ds.flush()
    fm2, ds2, rk2, applied2, skipped2 = _device(ctx, feats, model, state)
    assert (applied, skipped) == (applied2, skipped2) and consumed == len(blob) and applied > 0
    rk = F.Ranker(fm, ds)
    got, ref = rk.make_query(reqs), rk2.make_query(reqs)
    for r, q in enumerate(reqs):
        want = fo.dense_matrix(mapping, q, state)
        assert _eq(got[r], want) and _eq(ref[r], want)
    ds.free(); fm.free(); ds2.free(); fm2.free()


def test_concurrent_ranks_and_flushes_keep_code_rows_coherent(ctx):
    """Host threads rank through five models (more than the code-row cache holds -> evictions) while another
    thread keeps updating and flushing OTHER items; every result must equal the oracle, and after the threads
    join the updated items must rank with their new values."""
    import threading

    import metarank_b200 as mb
    from metarank_b200 import features as F

    names = [f"f{j}" for j in range(6)]
    fm = F.FeatureMapping(ctx, [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names], names)
    ds = F.DeviceState(ctx, fm)
    n_items = 4000
    cat = synth.feature_matrix(n_items, 6, seed=31)
    ids = np.arange(1, n_items + 1, dtype=np.uint64) * np.uint64(0xD6E8FEB86659FD93)
    ds.put_packed(F.pack_number_columns(names, ids, cat)); ds.flush()
    blobs = [synth.lightgbm_model_text(40, 6, seed=50 + k) for k in range(5)]
    boosters = [mb.LightGBMBooster(ctx, b) for b in blobs]
    oracles = [oracle.OracleBooster(0, b) for b in blobs]
    stable = np.arange(0, 2000)        # never updated
    hot = np.arange(2000, n_items)     # updated concurrently
    errors = []

    def ranker(tid):
        rng = np.random.Generator(np.random.PCG64(100 + tid))
        rk = F.Ranker(fm, ds)
        try:
            for it in range(25):
                m = int(rng.integers(0, 5))
                sizes = rng.integers(1, 80, 20)
                offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
                pick = rng.choice(stable, int(offs[-1]))
                R = len(sizes)
                arrays = dict(offsets=offs, ids=ids[pick], users=np.zeros(R, dtype=np.uint64), sessions=np.zeros(R, dtype=np.uint64),
                              req_f64=np.zeros((R, 1)), req_u64=np.zeros((R, 1), dtype=np.uint64),
                              req_vec=np.zeros((R, 1), dtype=np.float32), req_vp=np.zeros((R, 1), dtype=np.uint8),
                              item_f64=None, n_requests=R, total_items=int(offs[-1]))
                scores, order, _ = rk.rank_arrays(arrays, boosters[m], want_order=True)
                want = oracles[m].predictMat(np.ascontiguousarray(cat[pick]), len(pick), 6, threads=1)
                if not _eq(scores, want):
                    errors.append((tid, it, "scores"))
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    new_hot = synth.feature_matrix(len(hot), 6, seed=77)
    new_hot = np.where(np.isnan(new_hot), 0.25, new_hot)

    def updater():
        try:
            for k in range(0, len(hot), 100):
                rows = hot[k:k + 100]
                ds.put_packed(F.pack_number_columns(names, ids[rows], new_hot[k:k + 100]))
                ds.flush()
        except Exception as e:  # noqa: BLE001
            errors.append(("updater", repr(e)))

    threads = [threading.Thread(target=ranker, args=(t,)) for t in range(3)] + [threading.Thread(target=updater)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    cat[hot] = new_hot
    rk = F.Ranker(fm, ds)
    arrays = dict(offsets=np.array([0, len(hot)], dtype=np.int32), ids=ids[hot], users=np.zeros(1, dtype=np.uint64),
                  sessions=np.zeros(1, dtype=np.uint64), req_f64=np.zeros((1, 1)), req_u64=np.zeros((1, 1), dtype=np.uint64),
                  req_vec=np.zeros((1, 1), dtype=np.float32), req_vp=np.zeros((1, 1), dtype=np.uint8), item_f64=None,
                  n_requests=1, total_items=len(hot))
    for m in range(5):
        scores, _, _ = rk.rank_arrays(arrays, boosters[m], want_order=True)
        assert _eq(scores, oracles[m].predictMat(np.ascontiguousarray(cat[hot]), len(hot), 6, threads=1))
    for b in boosters:
        b.free()
    ds.free(); fm.free()


def test_token_field_match_random_batches_and_slices(ctx):
    """field_match ngram / term / bm25 on the device against the oracle: random vocabularies, items without a
    title or with a non-string one, requests without a query, a batch large enough to be cut into several
    pipelined slices (the slice's token range is rebased in the kernel), with and without a model."""
    import metarank_b200 as mb
    from metarank_b200 import features as F

    rng = np.random.Generator(np.random.PCG64(404))
    vocab = ["".join(rng.choice(list("abcdefgh"), int(rng.integers(2, 7)))) for _ in range(60)]
    tf = {w: int(rng.integers(1, 50)) for w in vocab[:40]}
    feats = [dict(name="ng", type="field_match", rankingField="ranking.query", itemField="item.title",
                  method=dict(type="ngram", n=2, language="whitespace")),
             dict(name="tm", type="field_match", rankingField="ranking.query", itemField="item.title",
                  method=dict(type="term", language="whitespace")),
             dict(name="bm", type="field_match", rankingField="ranking.q2", itemField="item.title",
                  method=dict(type="bm25", language="en", docs=50, avgdl=3.5, termfreq=tf)),
             dict(name="price", type="number", scope="item", source="item.price")]
    model = ["ng", "price", "bm", "tm"]
    mapping = fo.FeatureMapping(feats, model)
    n_items = 1500
    events = []
    for i in range(n_items):
        u = rng.random()
        words = [str(w) for w in rng.choice(vocab, int(rng.integers(0, 7)))]
        fields = [("price", float(i % 17))]
        if u < 0.8:
            fields.append(("title", " ".join(words)))
        elif u < 0.9:
            fields.append(("title", words))          # StringListField: joined by " "
        elif u < 0.95:
            fields.append(("title", 3.0))            # wrong type: no write
        ev = dict(event="item", id=f"e{i}", timestamp=G.NOW, item=f"p{i}", fields=fields)
        ev["tokens"] = {"bm": sorted(set(words))}    # the English analyzer is the caller's: stand-in tokens
        events.append(ev)
    state = fo.FeatureValueFlow(mapping, always_refresh=True).process(events)
    fm, ds, rk, _, _ = _device(ctx, feats, model, state)
    reqs = []
    for r in range(760):
        n = int(rng.integers(300, 420))
        items = [f"p{int(k)}" for k in rng.integers(0, n_items + 20, n)]
        fields = []
        qwords = [str(w) for w in rng.choice(vocab, int(rng.integers(0, 5)))]
        if rng.random() < 0.85:
            fields.append(("query", " ".join(qwords)))
        q2 = [str(w) for w in rng.choice(vocab, int(rng.integers(0, 4)))]
        if rng.random() < 0.85:
            fields.append(("q2", " ".join(q2)))
        q = G.ranking(items, fields, rid=f"r{r}")
        q["tokens"] = {"bm": sorted(set(q2))}
        reqs.append(q)
    arrays = fm.pack_requests(reqs)
    assert arrays["total_items"] > (1 << 18)  # more than one slice
    want = np.concatenate([fo.dense_matrix(mapping, q, state) for q in reqs])
    assert (want[:, 0] > 0).any() and (want[:, 2] != 0).any() and (want[:, 3] > 0).any()
    _, _, got = rk.rank_arrays(arrays, None, want_features=True)
    assert _eq(got, want)
    blob = synth.lightgbm_model_text(60, 4, seed=12)
    booster = mb.LightGBMBooster(ctx, blob)
    scores, order, _ = rk.rank_arrays(arrays, booster, want_order=True)
    ws = oracle.OracleBooster(0, blob).predictMat(np.ascontiguousarray(want), len(want), 4, threads=0)
    assert _eq(scores, ws)
    offs = arrays["offsets"]
    for r in rng.integers(0, len(reqs), 25):
        assert np.array_equal(order[offs[r]:offs[r + 1]], oracle.rank_order(ws[offs[r]:offs[r + 1]]))
    booster.free(); ds.free(); fm.free()


def test_rank_from_natively_decoded_json(ctx):
    """POST /rank bodies -> mr_requests_decode -> mr_rank, no Python packing in between: same features, scores
    and order as the Python shim's path, which the other tests hold to the oracle (ranklens feature set)."""
    import json

    import metarank_b200 as mb
    from metarank_b200 import features as F

    feats, model = synth.ranklens_config()
    state, item_ids, sessions = synth.ranklens_state(n_items=400, n_sessions=25, seed=21)
    reqs = synth.ranklens_requests(item_ids, sessions, 30, 60, seed=22)
    fm, ds, rk, _, _ = _device(ctx, feats, model, state)
    bodies = []
    for q in reqs:  # the request dicts of the oracle -> the JSON the reference's API takes
        o = dict(id=q["id"], timestamp=q["timestamp"], user=q.get("user"), session=q.get("session"),
                 fields=[dict(name=n, value=v) for n, v in q.get("fields", [])],
                 items=[dict(id=it["id"], fields=[dict(name=n, value=v) for n, v in it.get("fields", [])]) for it in q["items"]])
        bodies.append(o)
    dec = F.DecodedRequests(fm, json.dumps(bodies))
    arrays = fm.pack_requests(reqs)
    booster = mb.LightGBMBooster(ctx, synth.lightgbm_model_text(80, fm.dim, seed=3, cat_features={7: 16}))
    want = rk.rank_arrays(arrays, booster, want_order=True, want_features=True)
    got = rk.rank_decoded(dec, booster, want_order=True, want_features=True)
    assert _eq(got[2], want[2]) and _eq(got[0], want[0]) and np.array_equal(got[1], want[1])
    for r, q in enumerate(reqs[:5]):
        assert _eq(got[2][arrays["offsets"][r]:arrays["offsets"][r + 1]], fo.dense_matrix(fo.FeatureMapping(feats, model), q, state))
    dec.free(); booster.free(); ds.free(); fm.free()


def _json_body(q):
    """An oracle request dict -> the JSON body the reference's API takes (plus the embeddings / tokens keys)."""
    import json

    def val(v):
        if isinstance(v, np.ndarray):
            return [float(x) for x in v]
        if isinstance(v, (np.floating, np.integer)):
            return float(v)
        return v

    o = dict(id=q["id"], timestamp=q["timestamp"], fields=[dict(name=n, value=val(v)) for n, v in q.get("fields", [])],
             items=[dict(id=it["id"], fields=[dict(name=n, value=val(v)) for n, v in it.get("fields", [])]) for it in q["items"]])
    for k in ("user", "session"):
        if q.get(k) is not None:
            o[k] = q[k]
    if q.get("embeddings"):
        o["embeddings"] = {k: [float(x) for x in v] for k, v in q["embeddings"].items()}
    if q.get("tokens"):
        o["tokens"] = q["tokens"]
    return json.dumps(o)


@pytest.mark.parametrize("case", G.CASES, ids=[c["name"] for c in G.CASES])
def test_golden_vectors_through_the_native_request_decoder(ctx, case):
    """The reference's golden vectors again, but the request travels as a JSON body through mr_requests_decode
    (no Python packing): the assembled matrix must still be the oracle's, bit for bit."""
    from metarank_b200 import features as F

    mapping = fo.FeatureMapping(case["features"], case["model_features"])
    state = fo.FeatureValueFlow(mapping, always_refresh=True).process(case["events"])
    want = fo.dense_matrix(mapping, case["request"], state)
    fm, ds, rk, _, _ = _device(ctx, case["features"], case["model_features"], state)
    try:
        dec = F.DecodedRequests(fm, _json_body(case["request"]))
        _, _, got = rk.rank_decoded(dec, None, want_features=True)
        assert _eq(got, want), (case["ref"], got, want)
        dec.free()
    finally:
        ds.free(); fm.free()


def test_reupserted_lists_reuse_their_pool_region(ctx):
    """FeatureValueSink re-emits the same keys continuously; list payloads (interacted_with bounded lists, tag
    lists, token sets) must be rewritten in place instead of growing the pool with every upsert,
    and what the kernels read must be the latest value, also when a list grows or shrinks."""
    from metarank_b200 import features as F

    feats = [dict(name="seen", type="interacted_with", interaction="click", field=["item.tags"], scope="session", count=50,
                  duration="24h"),
             dict(name="d", type="diversity", source="metadata.tags")]
    model = ["seen", "d"]
    mapping = fo.FeatureMapping(feats, model)
    fm = F.FeatureMapping(ctx, feats, model)
    ds = F.DeviceState(ctx, fm)
    rng = np.random.Generator(np.random.PCG64(5))
    items = [f"i{k}" for k in range(40)]
    try:
        state = {}
        for it in items:
            tags = [f"t{int(x)}" for x in rng.integers(0, 9, int(rng.integers(1, 5)))]
            state[(("item", it), "seen_tags")] = ("scalar", tags)
            state[(("item", it), "d")] = ("scalar", tags)
        sizes = []
        for rnd in range(60):
            n = int(rng.integers(1, 40)) if rnd % 7 else 50  # grows and shrinks
            state[(("session", "s"), "seen_interactions")] = ("blist", [items[int(j)] for j in rng.integers(0, 40, n)])
            it = items[int(rng.integers(0, 40))]
            tags = [f"t{int(x)}" for x in rng.integers(0, 9, int(rng.integers(1, 6)))]
            state[(("item", it), "seen_tags")] = ("scalar", tags)
            state[(("item", it), "d")] = ("scalar", tags)
            ds.put(state)  # the WHOLE state again, like a sink that re-emits every key
            ds.flush()
            sizes.append(int(ds.info().device_bytes))
            req = G.ranking(items[:25], session="s")
            assert _eq(F.Ranker(fm, ds).make_query([req])[0], fo.dense_matrix(mapping, req, state)), rnd
        # the pool settles: a list that outgrows its region moves to one of twice the size, so 60 re-emissions of the
        # whole state cost a bounded factor — appending every time would have multiplied the pool 60-fold
        assert sizes[-1] <= 1.5 * sizes[1], sizes
    finally:
        ds.free(); fm.free()


def test_writes_cannot_silently_continue_loaded_values(ctx):
    """A PeriodicCounterValue / BoundedListValue that arrived as a refreshed value does not carry the raw state
    behind it (day buckets, entry timestamps): a PeriodicIncrement / Append on such a key must fail loudly instead
    of resetting the windows / the history; the reverse — a refreshed value over written state —
    simply supersedes it."""
    import metarank_b200 as mb
    from metarank_b200 import features as F

    feats = [dict(name="wc", type="window_count", interaction="click", scope="item", bucket="24h", periods=[1, 7]),
             dict(name="seen", type="interacted_with", interaction="click", field=["item.tags"], scope="session", count=5,
                  duration="48h")]
    model = ["wc", "seen"]
    mapping = fo.FeatureMapping(feats, model)
    flow = fo.FeatureValueFlow(mapping, always_refresh=True)
    fm = F.FeatureMapping(ctx, feats, model)
    ds = F.DeviceState(ctx, fm)
    try:
        ev = [G.item_event("a", [("tags", ["x", "y"])]), G.item_event("b", [("tags", ["y"])]),
              G.interaction("a", "r", "click", session="s", ts=G.NOW - 3600_000),
              G.interaction("b", "r", "click", session="s", ts=G.NOW)]
        state = dict(flow.process(ev))
        ds.apply_writes(flow.write_log)
        ds.flush()
        req = G.ranking(["a", "b"], session="s")
        assert _eq(F.Ranker(fm, ds).make_query([req])[0], fo.dense_matrix(mapping, req, state))
        # refreshed values over written state: they win, and from then on raw writes for those keys are refused
        state[(("item", "a"), "wc")] = ("pcounter", [5, 9])
        state[(("session", "s"), "seen_interactions")] = ("blist", ["b"])
        ds.put({k: state[k] for k in [(("item", "a"), "wc"), (("session", "s"), "seen_interactions")]})
        ds.flush()
        assert _eq(F.Ranker(fm, ds).make_query([req])[0], fo.dense_matrix(mapping, req, state))
        flow.write_log.clear()
        flow.process([G.interaction("a", "r", "click", session="s", ts=G.NOW + 1000)])
        hit = 0
        for w in flow.write_log:
            kind, (scope, name), ts, v = w
            touches_loaded = (name == "wc" and scope == ("item", "a")) or (name == "seen_interactions" and scope == ("session", "s"))
            if touches_loaded:
                hit += 1
                with pytest.raises(mb.MrError) as e:
                    ds.apply_writes([w])
                assert e.value.status == 5 and "refreshed value" in e.value.message
        assert hit == 2
        # a key that only ever saw raw writes keeps working
        flow.write_log.clear()
        flow.process([G.interaction("b", "r", "click", session="s2", ts=G.NOW + 2000)])
        a, s = ds.apply_writes(flow.write_log)
        assert a > 0
    finally:
        ds.free(); fm.free()


@pytest.mark.parametrize("dim", [384, 20, 18, 132])
def test_embedding_storage_switches_from_binary32_to_f64(ctx, dim):
    """Embeddings whose every element round-trips through binary32 (anything an ONNX encoder produced) live on the
    device as f32 and are scored by cosine_f32_kernel; the first element that needs all 53 bits switches the whole
    array to f64 (sparse update -> full re-upload) and the old kernel.  Either way the column equals the reference's
    sequential sums bit for bit.  Requests of 100 items put request boundaries inside the kernel's 128-item CTAs
    (the per-CTA widened query vs the per-lane fallback); dim 18 is not a multiple of 4 (always f64), 132 has a
    4-float tail chunk."""
    feats = [dict(name="sim", type="field_match", rankingField="ranking.query", itemField="item.title",
                  method=dict(type="bi-encoder", dim=dim), distance="cos")]
    rng = np.random.Generator(np.random.PCG64(dim))
    n = 700
    state = {}
    for k in range(n):
        if rng.random() > 0.04:
            state[(("item", f"i{k}"), "sim")] = ("scalar", rng.standard_normal(dim).astype(np.float32).astype(np.float64).tolist())
    fm, ds, rk, _, _ = _device(ctx, feats, ["sim"], state)
    mapping = fo.FeatureMapping(feats, ["sim"])

    def check():
        reqs = []
        for r in range(7):
            pick = rng.choice(n + 5, 100, replace=False)  # a few unknown items too
            reqs.append(dict(event="ranking", id=f"r{r}", timestamp=0, user=None, session=None, fields=[("query", "q")],
                             embeddings={} if r == 3 else {"sim": rng.standard_normal(dim).astype(np.float32)},
                             items=[dict(id=f"i{int(j)}", fields=[]) for j in pick]))
        got = rk.make_query(reqs)
        for r, q in enumerate(reqs):
            assert _eq(got[r], fo.dense_matrix(mapping, q, state)), r

    try:
        check()
        upd = {(("item", "i9"), "sim"): ("scalar", rng.standard_normal(dim).astype(np.float32).astype(np.float64).tolist())}
        state.update(upd); ds.put(upd); ds.flush()   # sparse update, still binary32
        check()
        upd = {(("item", "i11"), "sim"): ("scalar", (rng.standard_normal(dim) * (1 + 2.0**-40)).tolist())}
        state.update(upd); ds.put(upd); ds.flush()   # needs 53 bits: the array becomes f64
        check()
        upd = {(("item", f"i{k}"), "sim"): ("scalar", rng.standard_normal(dim).tolist()) for k in (1, 2, n - 1)}
        state.update(upd); ds.put(upd); ds.flush()
        check()
    finally:
        ds.free(); fm.free()


def test_upsert_batch_is_validated_before_anything_is_applied(ctx):
    """A packed buffer whose LATER record is malformed (a vector of the wrong dim, a cut record) must not leave its earlier
    records applied: the whole buffer is checked first."""
    import metarank_b200 as mb
    from metarank_b200 import features as F

    feats = [dict(name="price", type="number", scope="item", source="metadata.price"),
             dict(name="vec", type="vector", scope="item", source="metadata.vec", reduce=["vector3"])]
    model = ["price", "vec"]
    mapping = fo.FeatureMapping(feats, model)
    fm = F.FeatureMapping(ctx, feats, model)
    ds = F.DeviceState(ctx, fm)
    try:
        good = {(("item", "a"), "price"): ("scalar", 3.5), (("item", "a"), "vec"): ("scalar", [1.0, 2.0, 3.0])}
        ds.put(good); ds.flush()
        req = G.ranking(["a", "b"])
        before = F.Ranker(fm, ds).make_query([req])[0]
        assert _eq(before, fo.dense_matrix(mapping, req, good))
        ok_rec = F.pack_feature_values({(("item", "b"), "price"): ("scalar", 9.0)})
        bad_vec = F.pack_feature_values({(("item", "b"), "vec"): ("scalar", [1.0, 2.0])})      # dim 2, the feature says 3
        for blob in (ok_rec + bad_vec, ok_rec + bad_vec[:-3], ok_rec + b"\x05\x00pri"):
            with pytest.raises(mb.MrError):
                ds.put_packed(blob)
            ds.flush()
            assert _eq(F.Ranker(fm, ds).make_query([req])[0], before)   # item b's price was NOT applied
        ds.put_packed(ok_rec); ds.flush()
        good[(("item", "b"), "price")] = ("scalar", 9.0)
        assert _eq(F.Ranker(fm, ds).make_query([req])[0], fo.dense_matrix(mapping, req, good))
    finally:
        ds.free(); fm.free()


def test_referer_medium_from_user_and_session_state(ctx):
    """`referer` (S/feature/RefererFeature.scala:94-109): the medium the JVM's parser stored under the user / the session,
    through the fixed medium -> index table; the golden of T/feature/RefererFeatureTest.scala:46-50 (search -> 1) and every
    other branch against the oracle."""
    from metarank_b200 import features as F

    feats = [dict(name="ref_medium", type="referer", source="ranking.ref", scope="user"),
             dict(name="ref_s", type="referer", source="interaction:click.ref", scope="session"),
             dict(name="price", type="number", scope="item", source="metadata.price")]
    model = ["ref_medium", "price", "ref_s"]
    mapping = fo.FeatureMapping(feats, model)
    fm = F.FeatureMapping(ctx, feats, model)
    ds = F.DeviceState(ctx, fm)
    try:
        state = {(("user", "u1"), "ref_medium"): ("scalar", "search"), (("item", "p1"), "price"): ("scalar", 2.5)}
        for k, m in enumerate(("unknown", "search", "internal", "social", "email", "paid", "smoke signals")):
            state[(("user", f"v{k}"), "ref_medium")] = ("scalar", m)
            state[(("session", f"s{k}"), "ref_s")] = ("scalar", m)
        ds.put(state); ds.flush()
        reqs = [G.ranking(["p1", "p2"], user="u1", session="s3")]
        reqs += [G.ranking(["p1"], user=f"v{k}", session=f"s{6 - k}") for k in range(7)]
        reqs += [G.ranking(["p2", "p1"], user="nobody", session="nothing"), G.ranking(["p1"], user=None, session=None)]
        got = F.Ranker(fm, ds).make_query(reqs)
        assert got[0][:, 0].tolist() == [1.0, 1.0]   # RefererFeatureTest: CategoryValue("search", 1)
        for q, g in zip(reqs, got):
            assert _eq(g, fo.dense_matrix(mapping, q, state))
    finally:
        ds.free(); fm.free()
