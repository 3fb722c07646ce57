"""GPU parity of Booster.predictMat (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): ordering bit-exact, |Δscore| <= 1e-5.  The kernel
accumulates per item in tree order in the library's own precision, so we assert the
stronger property: scores are BIT-IDENTICAL to the oracle.
"""
import numpy as np
import pytest

from metarank_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu

TOL = 1e-5  # the stated tolerance; the tests below additionally require exact equality


def _check(ctx, kind, blob, X, **opts):
    import metarank_b200 as mb

    ob = oracle.OracleBooster(kind, blob)
    want = ob.predictMat(X, *X.shape)
    b = mb.B200Booster(ctx, blob, kind=kind, n_features=X.shape[1])
    try:
        for k, v in opts.items():
            b.set_option(k, v)
        got = b.predictMat(X, *X.shape)
    finally:
        b.free()
    assert got.shape == want.shape
    assert np.all(np.abs(got - want) <= TOL)
    assert np.array_equal(got, want), f"max |Δ| = {np.max(np.abs(got - want))}"
    assert np.array_equal(ctx.rank_order(got), oracle.rank_order(want))
    return got


@pytest.mark.parametrize("variant,ilp", [(0, 2), (0, 4), (2, 2), (4, 1), (5, 1)])
@pytest.mark.parametrize("threads", [32, 128, 256])
def test_c2_lightgbm_100x30x500(ctx, variant, ilp, threads):
    blob = synth.lightgbm_model_text(500, 30, seed=1234 + 2)
    X = synth.feature_matrix(100, 30, seed=42 + 2)
    _check(ctx, 0, blob, X, variant=variant, ilp=ilp, threads=threads)


@pytest.mark.parametrize("variant", [0, 2, 4, 5])
@pytest.mark.parametrize("chunk_kb", [4, 32, 200])
def test_chunking_is_invisible(ctx, chunk_kb, variant):
    blob = synth.lightgbm_model_text(500, 30, seed=7, stump_every=11)
    X = synth.feature_matrix(3000, 30, seed=8)
    _check(ctx, 0, blob, X, chunk_kb=chunk_kb, variant=variant)


@pytest.mark.parametrize("rows", [1, 2, 31, 32, 33, 255, 257, 1000, 4097])
def test_ragged_row_counts(ctx, rows):
    blob = synth.lightgbm_model_text(64, 13, seed=3)  # odd column count: scalar tile loads
    X = synth.feature_matrix(rows, 13, seed=rows)
    _check(ctx, 0, blob, X)


def test_categorical_zero_missing_and_stumps(ctx):
    cat = {2: 40, 7: 100}
    blob = synth.lightgbm_model_text(120, 10, seed=11, cat_features=cat, zero_missing=True, stump_every=7)
    rng = np.random.Generator(np.random.PCG64(5))
    X = synth.feature_matrix(2000, 10, seed=6)
    X[:, 2] = rng.integers(-2, 45, 2000)  # negative and out-of-bitset categories
    X[:, 7] = rng.integers(0, 130, 2000)
    X[rng.random(2000) < 0.1, 2] = np.nan
    X[rng.random(2000) < 0.05, 7] = 3e10  # beyond int range -> right child
    X[rng.random(2000) < 0.2, 0] = 0.0
    X[rng.random(2000) < 0.1, 1] = np.nan
    X[rng.random(2000) < 0.05, 3] = np.inf
    X[rng.random(2000) < 0.05, 4] = -np.inf
    X[rng.random(2000) < 0.05, 5] = 1e-36  # inside LightGBM's zero band
    for variant in (0, 2):  # zero-band models cannot be binned: variant 2 silently uses the f64 kernel
        _check(ctx, 0, blob, X, variant=variant)
    blob = synth.lightgbm_model_text(120, 10, seed=12, cat_features=cat, zero_missing=False, stump_every=7)
    for variant in (0, 2):  # categorical bitsets + NaN routing through the binned kernel
        for threads in (0, 64):
            _check(ctx, 0, blob, X, variant=variant, threads=threads)


def test_binned_codes_at_threshold_boundaries(ctx):
    """Values exactly on / one ulp around every threshold must route like the f64 compare."""
    from oracle import model_parse

    blob = synth.lightgbm_model_text(40, 6, seed=33)
    m = model_parse.parse_lightgbm_text(blob)
    rows = []
    for t in m["trees"]:
        for f, thr in zip(t["split_feature"], t["threshold"]):
            for v in (thr, np.nextafter(thr, np.inf), np.nextafter(thr, -np.inf)):
                r = np.zeros(6)
                r[f] = v
                rows.append(r)
    X = np.array(rows)
    for v in (2, 4, 5):
        _check(ctx, 0, blob, X, variant=v)
    xb = synth.xgboost_model_json(30, 5, depth=5, seed=34)
    mx = model_parse.parse_xgboost(xb)
    rows = []
    for t in mx["trees"]:
        for f, thr, l in zip(t["split_index"], t["split_cond"], t["left"]):
            if l == -1:
                continue
            for v in (float(thr), float(np.nextafter(np.float32(thr), np.float32(np.inf))),
                      float(np.nextafter(np.float32(thr), np.float32(-np.inf))), float(thr) + 1e-12, float(thr) - 1e-12):
                r = np.zeros(5)
                r[f] = v
                rows.append(r)
    for v in (2, 4, 5):
        _check(ctx, 1, xb, np.array(rows), variant=v)


def test_deep_unbalanced_lightgbm(ctx):
    blob = synth.lightgbm_model_text(50, 20, num_leaves=255, max_depth=0, seed=21)
    X = synth.feature_matrix(777, 20, seed=22)
    for v in (-1, 0, 2, 4, 5):
        _check(ctx, 0, blob, X, variant=v)
    # one tree of 20 000 leaves: too large for the compact layout's 16-bit offsets -> generic binned kernel
    big = synth.lightgbm_model_text(2, 20, num_leaves=20000, max_depth=0, seed=23)
    _check(ctx, 0, big, X, chunk_kb=220)


@pytest.mark.parametrize("fmt", ["json", "ubj"])
@pytest.mark.parametrize("depth,full", [(6, True), (8, False)])
def test_c4_xgboost(ctx, fmt, depth, full):
    gen = synth.xgboost_model_json if fmt == "json" else synth.xgboost_model_ubj
    blob = gen(200, 16, depth=depth, seed=1234 + 4, full=full)
    X = synth.feature_matrix(256, 16, seed=42 + 4)
    for variant in (-1, 0, 2, 4, 5):  # -1 on 256 rows = the low-latency tree-parallel path
        _check(ctx, 1, blob, X, variant=variant)


def test_low_latency_path_small_batches(ctx):
    """<= 2048 rows in auto mode: per-tree leaf values over (chunk x group) CTAs + in-order sum."""
    blob = synth.lightgbm_model_text(500, 30, seed=41, stump_every=13)
    for rows in (1, 100, 129, 2048, 2049):
        X = synth.feature_matrix(rows, 30, seed=rows)
        _check(ctx, 0, blob, X)
    xb = synth.xgboost_model_ubj(120, 9, depth=7, seed=42, full=False)
    for rows in (3, 100, 1000):
        _check(ctx, 1, xb, synth.feature_matrix(rows, 9, seed=rows))


def test_xgboost_f32_rounding_of_inputs(ctx):
    # values that differ only below binary32 precision must route identically to the oracle
    blob = synth.xgboost_model_json(50, 4, depth=5, seed=9)
    rng = np.random.Generator(np.random.PCG64(10))
    X = rng.standard_normal((500, 4))
    X += rng.standard_normal((500, 4)) * 1e-9
    _check(ctx, 1, blob, X)


def test_wide_rows_fall_back_to_global_reads(ctx):
    blob = synth.lightgbm_model_text(20, 1200, seed=31)
    X = synth.feature_matrix(300, 1200, seed=32)
    _check(ctx, 0, blob, X, variant=0)
    _check(ctx, 0, blob, X)  # auto: too wide for the binning kernel -> exact f64 kernel


def test_c5_mega_request_properties(ctx):
    """BASELINE config #5 shape at full size: 10 000 x 64, 2000 trees."""
    import metarank_b200 as mb

    blob = synth.lightgbm_model_text(2000, 64, seed=1234 + 5)
    X = synth.feature_matrix(10_000, 64, seed=42 + 5)
    got = _check(ctx, 0, blob, X)
    # size-independent properties: row permutation equivariance, duplication
    b = mb.B200Booster(ctx, blob, kind=0)
    perm = np.random.Generator(np.random.PCG64(1)).permutation(X.shape[0])
    got_p = b.predictMat(X[perm], *X.shape)
    assert np.array_equal(got_p, got[perm])
    twice = b.predictMat(np.concatenate([X[:100], X[:100]]), 200, 64)
    assert np.array_equal(twice[:100], twice[100:])
    b.free()


def test_empty_and_errors(ctx):
    import metarank_b200 as mb

    blob = synth.lightgbm_model_text(5, 4, seed=1)
    b = mb.B200Booster(ctx, blob, kind=0)
    assert b.predictMat(np.zeros((0, 4)), 0, 4).shape == (0,)
    with pytest.raises(mb.MrError) as e:
        b.predictMat(np.zeros((3, 5)), 3, 5)
    assert e.value.status == 1
    assert b.save() == blob
    assert b.weights().sum() == b.info().n_internal_nodes
    assert not b.isClosed()
    b.close()
    b.close()  # idempotent
    assert b.isClosed()
    with pytest.raises(mb.MrError) as e:
        b.predictMat(np.zeros((1, 4)), 1, 4)
    assert e.value.status == 4
    with pytest.raises(mb.MrError):
        mb.B200Booster(ctx, b"not a model", kind=0)
    with pytest.raises(mb.MrError) as e:
        mb.B200Booster(ctx, blob, kind=0, n_features=9)
    assert e.value.status == 6


def test_metarank_blob_framing(ctx):
    import metarank_b200 as mb

    names = [f"f{i}" for i in range(6)]
    lgb = synth.lightgbm_model_text(10, 6, seed=2)
    X = synth.feature_matrix(50, 6, seed=3)
    want = oracle.OracleBooster(0, lgb).predictMat(X, 50, 6)
    for version in (2, 3):
        blob = synth.metarank_model_blob(names, 0, lgb, version=version)
        b = mb.B200Booster.from_metarank_blob(ctx, blob, names)
        assert np.array_equal(b.predictMat(X, 50, 6), want)
        b.free()
    with pytest.raises(mb.MrError) as e:
        mb.B200Booster.from_metarank_blob(ctx, blob, names[:-1] + ["other"])
    assert e.value.status == 6 and "booster trained with" in e.value.message


def test_mean_path_matches_oracle(ctx):
    import metarank_b200 as mb

    blob = synth.lightgbm_model_text(100, 30, seed=4)
    X = synth.feature_matrix(500, 30, seed=5)
    ob = oracle.OracleBooster(0, blob)
    ob.predictMat(X, 500, 30)
    b = mb.B200Booster(ctx, blob, kind=0)
    # single-leaf trees cost one dummy node on the device; none here
    assert b.mean_path(X, 500, 30) == pytest.approx(ob.visited / (500 * 100), rel=1e-12)
    b.free()


def test_concurrent_predicts(ctx):
    import threading

    import metarank_b200 as mb

    blob = synth.lightgbm_model_text(200, 30, seed=6)
    b = mb.B200Booster(ctx, blob, kind=0)
    ob = oracle.OracleBooster(0, blob)
    errs = []

    def work(seed):
        try:
            X = synth.feature_matrix(100 + seed, 30, seed=seed)
            want = ob.predictMat(X, *X.shape)
            for _ in range(20):
                assert np.array_equal(b.predictMat(X, *X.shape), want)
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    ts = [threading.Thread(target=work, args=(s,)) for s in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    b.free()
    assert not errs, errs


def test_latency_path_beyond_2048_rows_and_pipelined_sum(ctx):
    """The tree-parallel path takes mega-request sized batches too (one 10 000-item request): leaf values over
    (chunk x 128-item group) CTAs, then the in-order sum with its 4-stage cp.async pipeline; tree counts around
    the stage / batch boundaries (24, 72, 96, 97) and row counts around the CTA shapes (32 / 64 / 128 threads)."""
    for n_trees in (1, 23, 24, 25, 72, 96, 97, 500):
        blob = synth.lightgbm_model_text(n_trees, 12, seed=300 + n_trees, stump_every=7)
        for rows in (1, 33, 2049, 4736, 4737, 9473):
            _check(ctx, 0, blob, synth.feature_matrix(rows, 12, seed=rows), latency_rows=16384)
    xb = synth.xgboost_model_json(130, 9, depth=6, seed=301)
    _check(ctx, 1, xb, synth.feature_matrix(5000, 9, seed=5), latency_rows=16384)


def test_removed_variants_are_rejected(ctx):
    import metarank_b200 as mb

    b = mb.B200Booster(ctx, synth.lightgbm_model_text(3, 4, seed=1), kind=0)
    for v in (1, 3, 6):
        with pytest.raises(mb.MrError) as e:
            b.set_option("variant", v)
        assert e.value.status == 1
    b.free()


def test_slim_scorer_layouts_and_fallbacks(ctx):
    """The 4-byte-node scorer (variant 5 / auto on large batches): every CTA-tile size it has (512 items up to 62 tile
    columns, 256 up to 126, 128 up to 254), batches around the tile boundaries, XGBoost f32, single-leaf trees, NaN in
    both directions on one feature (duplicated tile columns), categorical splits, models it must refuse (trees or bitsets
    too large for a block -> the 8-byte compact kernel) — all bit-identical to the oracle, like every other scorer."""
    import metarank_b200 as mb

    for n_feat, trees in ((30, 120), (70, 60), (120, 40), (200, 20), (300, 12)):
        blob = synth.lightgbm_model_text(trees, n_feat, seed=500 + n_feat, stump_every=9)
        for rows in (1, 127, 513, 5000, 40_000):
            X = synth.feature_matrix(rows, n_feat, seed=rows + n_feat)
            _check(ctx, 0, blob, X, variant=5)
        _check(ctx, 0, blob, synth.feature_matrix(40_000, n_feat, seed=3))  # auto: 40 000 rows is the throughput path
    for fmt in (synth.xgboost_model_json, synth.xgboost_model_ubj):
        xb = fmt(90, 16, depth=6, seed=77)
        for rows in (33, 4096, 40_000):
            _check(ctx, 1, xb, synth.feature_matrix(rows, 16, seed=rows), variant=5)
    deep = synth.lightgbm_model_text(40, 12, num_leaves=120, max_depth=0, seed=78)   # 2 KB tree blocks
    _check(ctx, 0, deep, synth.feature_matrix(3000, 12, seed=4), variant=5)
    huge = synth.lightgbm_model_text(6, 12, num_leaves=400, max_depth=0, seed=79)   # no slim form: silently the compact kernel
    _check(ctx, 0, huge, synth.feature_matrix(3000, 12, seed=5), variant=5)
    # categorical splits: entries flagged in bit 0 leave the level loop and are resolved against the block's bitsets
    for cats in ({3: 20}, {0: 7, 5: 300, 9: 33}):
        cat = synth.lightgbm_model_text(50, 10, seed=80, cat_features=cats, stump_every=11)
        Xc = synth.feature_matrix(6000, 10, seed=6)
        for f, n in cats.items():  # integer category values, some NaN / negative / beyond the bitset
            Xc[:, f] = np.random.Generator(np.random.PCG64(f)).integers(-2, n + 40, 6000)
            Xc[::17, f] = np.nan
        _check(ctx, 0, cat, Xc, variant=5)
        _check(ctx, 0, cat, Xc, variant=4)
    wide_cat = synth.lightgbm_model_text(20, 10, seed=81, cat_features={2: 40000})   # 5 KB bitsets: no slim form -> compact kernel
    _check(ctx, 0, wide_cat, synth.feature_matrix(3000, 10, seed=7), variant=5)


def test_small_categorical_in_loop_form_and_parameter_root_table(ctx, monkeypatch):
    """Models whose bitsets all live in categories 0..15 carry `0xFFF0 | category` codes (gbdt_model.h kMetaCat16): the slim
    scorer resolves their categorical nodes inside its level loop, every other scorer decodes the code back.  Level 0 of
    the slim walk comes from a root table in the kernel's parameter space (<= 512 / <= 1920 trees; beyond that, or with
    MR_NO_ROOT_TAB, from the chunk's own table).  Categories: in range, fractional (truncated like static_cast<int>),
    negative, >= 16, beyond int range, NaN."""
    rng = np.random.Generator(np.random.PCG64(91))
    for cats, trees in (({3: 16}, 60), ({0: 7, 5: 12, 9: 16}, 40), ({1: 2}, 700), ({4: 16}, 2100)):
        blob = synth.lightgbm_model_text(trees, 10, seed=90 + trees, cat_features=cats, stump_every=13)
        X = synth.feature_matrix(6000, 10, seed=trees)
        for f, n in cats.items():
            X[:, f] = rng.integers(-2, n + 6, 6000)
            X[::17, f] = np.nan
            X[1::29, f] += 0.75
            X[5::131, f] = 3e10
            X[7::131, f] = -0.5   # static_cast<int>(-0.5) == 0: category 0
        for variant in (5, 4, 2, 0):
            _check(ctx, 0, blob, X, variant=variant)
        _check(ctx, 0, blob, X[:300])                   # auto on a small batch: the tree-parallel latency path
        _check(ctx, 0, blob, np.tile(X, (7, 1)))        # auto on 42 000 rows: the throughput path
        monkeypatch.setenv("MR_NO_ROOT_TAB", "1")       # the same models through the chunk-resident root tables
        _check(ctx, 0, blob, X, variant=5)
        monkeypatch.delenv("MR_NO_ROOT_TAB")
        monkeypatch.setenv("MR_NO_CAT16", "1")          # ... and through the wide-bitset form (the loop leaves on a categorical node)
        _check(ctx, 0, blob, X, variant=5)
        monkeypatch.delenv("MR_NO_CAT16")
    # numeric models: leaves hanging directly off the root, single-leaf trees, both root-table sizes, XGBoost f32
    for trees, leaves in ((30, 2), (300, 3), (513, 31), (1920, 8), (1921, 8)):
        blob = synth.lightgbm_model_text(trees, 12, num_leaves=leaves, seed=trees, stump_every=5)
        X = synth.feature_matrix(5000, 12, seed=trees + 1)
        _check(ctx, 0, blob, X, variant=5)
    xb = synth.xgboost_model_json(600, 16, depth=3, seed=92)
    _check(ctx, 1, xb, synth.feature_matrix(5000, 16, seed=93), variant=5)


def test_batches_of_several_hundred_thousand_rows(ctx):
    """Batches around the chip's one-wave capacity (148 SMs x 1536 rows): partial last rounds of the persistent CTAs, the
    small-categorical form, XGBoost f32 — bit-identical to the oracle at every size."""
    import metarank_b200 as mb

    models = [(0, synth.lightgbm_model_text(40, 24, seed=70, cat_features={7: 16}), 24),
              (0, synth.lightgbm_model_text(40, 30, seed=71, stump_every=9), 30),
              (1, synth.xgboost_model_json(30, 16, depth=6, seed=72), 16)]
    for kind, blob, nf in models:
        ob = oracle.OracleBooster(kind, blob)
        X = synth.feature_matrix(300_000, nf, seed=nf)
        if kind == 0 and nf == 24:
            X[:, 7] = np.random.Generator(np.random.PCG64(1)).integers(-1, 18, len(X))
            X[::37, 7] = np.nan
        want = ob.predictMat(X, *X.shape, threads=0)
        b = mb.B200Booster(ctx, blob, kind=kind, n_features=nf)
        try:
            for rows in (76_000, 131_072, 227_328, 256_000, 300_000):
                got = b.predictMat(X[:rows], rows, nf)
                assert np.array_equal(got, want[:rows]), (kind, nf, rows)
        finally:
            b.free()


def test_xgboost_deprecated_binary_model_scores_like_its_json_twin(ctx):
    """The encoding xgboost4j's toByteArray() writes up to 2.0: the same synthetic ensemble as JSON and as binary (with the
    `binf` prefix, with pruned nodes left in the arrays) gives the same bits through every scorer, and the oracle's own
    reader of the format agrees."""
    for depth, full, rows in ((6, True, 700), (8, False, 5000), (4, False, 40_000)):
        j = synth.xgboost_model_json(80, 16, depth=depth, seed=60 + depth, full=full)
        X = synth.feature_matrix(rows, 16, seed=depth)
        X[::13, 3] = np.nan
        want = _check(ctx, 1, j, X)
        for kw in ({}, {"magic": True, "deleted": 2}):
            b = synth.xgboost_model_binary(80, 16, depth=depth, seed=60 + depth, full=full, **kw)
            for variant in (-1, 0, 5):
                assert np.array_equal(_check(ctx, 1, b, X, variant=variant), want)
