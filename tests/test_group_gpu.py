"""Mega-request sharding on the GPU (BASELINE configs[4]): mr_group_rank must return exactly
what mr_rank returns on one GPU — scores bit for bit, the same order — however the request is cut.  The members
of a group may share a device, so the whole exchange (peer stores from the scoring kernel, flags, device-side
wait, ordering of the gathered vector) runs on a single-GPU box: one thread per member, like a JVM would."""
import threading

import numpy as np
import pytest

import metarank_b200 as mb
from metarank_b200 import features as F, sharded, synth
from oracle import features_oracle as fo, oracle

pytestmark = pytest.mark.gpu


def _number_setup(ctx, n_feat, n_cat, seed):
    names = [f"f{j}" for j in range(n_feat)]
    feats = [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names]
    fm = F.FeatureMapping(ctx, feats, names)
    st = F.DeviceState(ctx, fm)
    cat = synth.feature_matrix(n_cat, n_feat, seed=seed)
    ids = (np.arange(1, n_cat + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    st.put_packed(F.pack_number_columns(names, ids, cat))
    st.flush()
    return fm, st, cat, ids


def _arrays(ids):
    n = len(ids)
    return dict(offsets=np.array([0, n], dtype=np.int32), ids=np.ascontiguousarray(ids), users=np.zeros(1, dtype=np.uint64),
                sessions=np.zeros(1, dtype=np.uint64), req_f64=np.zeros((1, 1)), req_u64=np.zeros((1, 1), dtype=np.uint64),
                req_vec=np.zeros((1, 1), dtype=np.float32), req_vp=np.zeros((1, 1), dtype=np.uint8), item_f64=None,
                n_requests=1, total_items=n)


def _run_members(members, st, booster, arrays):
    out = [None] * len(members)
    err = []

    def go(k):
        try:
            out[k] = members[k].rank_arrays(st, booster, arrays)
        except Exception as e:  # noqa: BLE001
            err.append(e)

    ts = [threading.Thread(target=go, args=(k,)) for k in range(len(members))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not err, err
    return out


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_group_rank_equals_single_gpu_rank(ctx, world):
    n_feat, n_trees = 24, 300
    fm, st, cat, ids = _number_setup(ctx, n_feat, 30_000, seed=61)
    blob = synth.lightgbm_model_text(n_trees, n_feat, seed=62)
    booster = mb.LightGBMBooster(ctx, blob, n_features=n_feat)
    ob = oracle.OracleBooster(0, blob)
    rk = F.Ranker(fm, st)
    members = [sharded.Group(ctx, r, world, 12_000) for r in range(world)]
    sharded.Group.connect_local(members)
    rng = np.random.Generator(np.random.PCG64(63))
    try:
        for n in [10_000, 12_000, 5000, 4097, 1280, 129, 128, 100, 1, 7777]:  # sizes alternate the exchange parity too
            pick = rng.choice(30_000, n, replace=False)
            req_ids = ids[pick].copy()
            if n > 10:
                req_ids[3] = np.uint64(12345)  # an unknown item: NaN row
            arrays = _arrays(req_ids)
            want_sc, want_od, _ = rk.rank_arrays(arrays, booster)
            X = cat[pick].copy()
            if n > 10:
                X[3] = np.nan
            osc = ob.predictMat(X, n, n_feat)
            assert np.array_equal(want_sc, osc)
            assert np.array_equal(want_od, oracle.rank_order(osc))
            for sc, od in _run_members(members, st, booster, arrays):
                assert np.array_equal(sc, osc), n
                assert np.array_equal(od, want_od), n
    finally:
        [m.free() for m in members]
        booster.free(); st.free(); fm.free()


@pytest.mark.timeout(300)
def test_group_ties_nan_scores_and_throughput_scorer(ctx):
    """Duplicate items (tied scores: stability), and a slice large enough for the throughput scorer
    (latency_rows forced low) so that the compact kernel's own peer stores are exercised."""
    n_feat = 16
    fm, st, cat, ids = _number_setup(ctx, n_feat, 5000, seed=71)
    blob = synth.lightgbm_model_text(120, n_feat, seed=72)
    booster = mb.LightGBMBooster(ctx, blob, n_features=n_feat)
    booster.set_option("latency_rows", 256)
    ob = oracle.OracleBooster(0, blob)
    members = [sharded.Group(ctx, r, 2, 9000) for r in range(2)]
    sharded.Group.connect_local(members)
    try:
        rng = np.random.Generator(np.random.PCG64(73))
        pick = rng.integers(0, 200, 9000)  # heavy duplication -> many exact ties
        arrays = _arrays(ids[pick])
        osc = ob.predictMat(cat[pick], 9000, n_feat)
        for sc, od in _run_members(members, st, booster, arrays):
            assert np.array_equal(sc, osc)
            assert np.array_equal(od, oracle.rank_order(osc))
    finally:
        [m.free() for m in members]
        booster.free(); st.free(); fm.free()


@pytest.mark.timeout(300)
def test_group_with_per_request_aggregates_and_exact_kernel(ctx):
    """A schema whose features read the whole item list (diversity, interacted_with: every member assembles
    the full request, only the scoring is split) and a model the binned scorers refuse (zero-as-missing: the
    exact f64 kernel + the publishing fallback)."""
    feats, model = synth.ranklens_config()
    state, item_ids, sessions = synth.ranklens_state(n_items=600, n_sessions=20, seed=81)
    reqs = synth.ranklens_requests(item_ids, sessions, 1, 500, seed=82)
    fm = F.FeatureMapping(ctx, feats, model)
    ds = F.DeviceState(ctx, fm)
    ds.put(state); ds.flush()
    mapping = fo.FeatureMapping(feats, model)
    want = fo.dense_matrix(mapping, reqs[0], state)
    arrays = fm.pack_requests(reqs)
    members = [sharded.Group(ctx, r, 3, 1000) for r in range(3)]
    sharded.Group.connect_local(members)
    try:
        for zero_missing in (False, True):
            blob = synth.lightgbm_model_text(80, fm.dim, seed=83, cat_features={7: 16}, zero_missing=zero_missing)
            booster = mb.LightGBMBooster(ctx, blob, n_features=fm.dim)
            osc = oracle.OracleBooster(0, blob).predictMat(want, *want.shape)
            for sc, od in _run_members(members, ds, booster, arrays):
                assert np.array_equal(sc, osc)
                assert np.array_equal(od, oracle.rank_order(osc))
            booster.free()
    finally:
        [m.free() for m in members]
        ds.free(); fm.free()


@pytest.mark.timeout(120)
def test_group_errors(ctx):
    fm, st, cat, ids = _number_setup(ctx, 8, 500, seed=91)
    blob = synth.lightgbm_model_text(10, 8, seed=92)
    booster = mb.LightGBMBooster(ctx, blob, n_features=8)
    g = sharded.Group(ctx, 0, 2, 256)
    try:
        with pytest.raises(mb.MrError) as e:  # not connected
            g.rank_arrays(st, booster, _arrays(ids[:10]))
        assert e.value.status == 1
        one = sharded.Group(ctx, 0, 1, 256)
        with pytest.raises(mb.MrError) as e:  # larger than the group's capacity
            one.rank_arrays(st, booster, _arrays(ids[:300]))
        assert e.value.status == 1
        two = dict(_arrays(ids[:10]), offsets=np.array([0, 5, 10], dtype=np.int32), n_requests=2,
                   users=np.zeros(2, dtype=np.uint64), sessions=np.zeros(2, dtype=np.uint64))
        with pytest.raises(mb.MrError) as e:  # batches are sharded by request, not by item
            one.rank_arrays(st, booster, two)
        assert e.value.status == 5
        sc, od = one.rank_arrays(st, booster, _arrays(ids[:0]))
        assert sc.size == 0 and od.size == 0
        one.free()
        with pytest.raises(mb.MrError):
            sharded.Group(ctx, 2, 2, 10)
    finally:
        g.free()
        booster.free(); st.free(); fm.free()
