"""Pins the GBDT oracle on hand-computed known answers.

The reference holds NO golden score for Booster.predictMat ("parity
unpinned"), so these cases are derived by hand from the published LightGBM / XGBoost
decision rules — each expected value below can be checked with pencil and paper
against the model text in the test.
"""
import json
import math

import numpy as np
import pytest

from metarank_b200 import synth
from oracle import model_parse, oracle

LGB_TEMPLATE = """tree
version=v4
num_class=1
num_tree_per_iteration=1
label_index=0
max_feature_idx=2
objective=lambdarank
feature_names=a b c
feature_infos=[0:1] [0:1] 0:1:2:3:40
tree_sizes=0

{trees}
end of trees

feature_importances:
a=1

parameters:
end of parameters

pandas_categorical:null
"""


def lgb(*trees):
    return LGB_TEMPLATE.format(trees="\n\n".join(trees)).encode()


# node 0: a <= 0.5 ? node 1 : leaf 2 ; node 1: b <= 1.5 ? leaf 0 : leaf 1
T_NUM = lambda dt0, dt1: f"""Tree=0
num_leaves=3
num_cat=0
split_feature=0 1
split_gain=1 1
threshold=0.5 1.5
decision_type={dt0} {dt1}
left_child=1 -1
right_child=-3 -2
leaf_value=10 20 30
leaf_weight=1 1 1
leaf_count=1 1 1
internal_value=0 0
internal_weight=1 1
internal_count=2 2
is_linear=0
shrinkage=1
"""

NAN = float("nan")


def predict(blob, rows):
    X = np.array(rows, dtype=np.float64)
    return oracle.OracleBooster(0, blob).predictMat(X, *X.shape).tolist()


def test_lightgbm_numerical_le_threshold():
    # decision_type 0: missing None, default right
    got = predict(lgb(T_NUM(0, 0)), [[0.5, 1.5, 0], [0.5000001, 0, 0], [0.4, 1.6, 0], [-1e300, -1e300, 0]])
    assert got == [10.0, 30.0, 20.0, 10.0]  # x <= thr goes LEFT (inclusive)


def test_lightgbm_nan_with_missing_none_becomes_zero():
    # missing None: NaN -> 0.0, then 0.0 <= 0.5 -> left ; 0.0 <= 1.5 -> left
    assert predict(lgb(T_NUM(0, 0)), [[NAN, NAN, 0]]) == [10.0]
    # threshold below zero: NaN -> 0.0 > -1 -> right
    t = T_NUM(0, 0).replace("threshold=0.5 1.5", "threshold=-1 1.5")
    assert predict(lgb(t), [[NAN, 0, 0]]) == [30.0]


def test_lightgbm_missing_nan_uses_default_direction():
    # decision_type 8 = missing NaN, default right ; 10 = missing NaN, default left
    assert predict(lgb(T_NUM(8, 8)), [[NAN, 0, 0], [0.1, NAN, 0]]) == [30.0, 20.0]
    assert predict(lgb(T_NUM(10, 10)), [[NAN, 99, 0], [0.1, NAN, 0]]) == [20.0, 10.0]


def test_lightgbm_missing_zero_band():
    # decision_type 4 = missing Zero, default right ; 6 = default left.  |x| <= 1e-35f is "zero"
    assert predict(lgb(T_NUM(4, 4)), [[0.0, 0.0, 0], [1e-36, 1.0, 0], [-1e-36, 1.0, 0], [1e-34, 1.0, 0]]) == \
        [30.0, 30.0, 30.0, 10.0]
    assert predict(lgb(T_NUM(6, 6)), [[0.0, 5.0, 0], [0.0, 0.0, 0]]) == [20.0, 10.0]
    # NaN with missing Zero: NaN -> 0.0 -> zero band -> default side
    assert predict(lgb(T_NUM(4, 4)), [[NAN, 0, 0]]) == [30.0]
    assert predict(lgb(T_NUM(6, 6)), [[NAN, NAN, 0]]) == [10.0]


T_CAT = """Tree=0
num_leaves=2
num_cat=1
split_feature=2
split_gain=1
threshold=0
decision_type=1
left_child=-1
right_child=-2
leaf_value=1 2
leaf_weight=1 1
leaf_count=1 1
internal_value=0
internal_weight=1
internal_count=2
cat_boundaries=0 2
cat_threshold=10 1
is_linear=0
shrinkage=1
"""


def test_lightgbm_categorical_bitset():
    # bitset words [10 (bits 1,3), 1 (bit 32)] -> categories {1, 3, 32} go left
    rows = [[0, 0, c] for c in (0, 1, 2, 3, 3.9, 31, 32, 33, 64, 1000, -1, NAN, 3e10, -0.5)]
    # 3.9 truncates to 3 (left); -0.5 truncates to 0 (not in set); -1 < 0 -> right
    want = [2, 1, 2, 1, 1, 2, 1, 2, 2, 2, 2, 2, 2, 2]
    assert predict(lgb(T_CAT), rows) == [float(w) for w in want]


def test_lightgbm_sum_is_f64_in_tree_order():
    # 0.1 + 0.2 + 0.3 in f64, left to right: (0.1 + 0.2) + 0.3 != 0.1 + (0.2 + 0.3)
    def stump(i, v):
        return f"Tree={i}\nnum_leaves=1\nnum_cat=0\nleaf_value={v!r}\nis_linear=0\nshrinkage=1\n"

    got = predict(lgb(stump(0, 0.1), stump(1, 0.2), stump(2, 0.3)), [[0, 0, 0]])
    assert got == [(0.1 + 0.2) + 0.3]
    assert got != [0.1 + (0.2 + 0.3)]


def test_lightgbm_threshold_parse_is_correctly_rounded():
    t = T_NUM(0, 0).replace("threshold=0.5 1.5", "threshold=0.1000000000000000055511151231257827 1.5")
    assert predict(lgb(t), [[0.1, 0, 0], [np.nextafter(0.1, 1), 0, 0]]) == [10.0, 30.0]


def _xgb(trees, base="5E-1", nf=2):
    return json.dumps({"learner": {
        "gradient_booster": {"name": "gbtree", "model": {"trees": trees, "tree_info": [0] * len(trees)}},
        "learner_model_param": {"base_score": base, "num_feature": str(nf), "num_class": "0"},
        "objective": {"name": "rank:ndcg"}}, "version": [2, 1, 4]}).encode()


def _xtree(split_cond, default_left):
    return {"left_children": [1, -1, -1], "right_children": [2, -1, -1], "split_indices": [0, 0, 0],
            "split_conditions": split_cond, "default_left": default_left, "split_type": [0, 0, 0]}


def test_xgboost_strict_less_default_and_f32():
    blob = _xgb([_xtree([0.5, 1.0, 2.0], [1, 0, 0])])
    X = np.array([[0.5, 0], [0.4999999, 0], [NAN, 0], [0.5 - 1e-12, 0]])
    got = oracle.OracleBooster(1, blob).predictMat(X, 4, 2).tolist()
    # 0.5 < 0.5 false -> right (2.0); 0.4999999 -> left; NaN -> default left;
    # 0.5-1e-12 rounds to 0.5f as binary32 -> NOT less -> right
    assert got == [2.5, 1.5, 1.5, 2.5]
    blob = _xgb([_xtree([0.5, 1.0, 2.0], [0, 0, 0])])
    assert oracle.OracleBooster(1, blob).predictMat(np.array([[NAN, 0.0]]), 1, 2).tolist() == [2.5]


def test_xgboost_f32_sequential_sum_from_base_score():
    leaves = [0.1, 0.2, 0.3, 1e-8, 1e-8]
    trees = [{"left_children": [-1], "right_children": [-1], "split_indices": [0], "split_conditions": [v],
              "default_left": [0], "split_type": [0]} for v in leaves]
    got = oracle.OracleBooster(1, _xgb(trees)).predictMat(np.zeros((1, 2)), 1, 2)[0]
    s = np.float32(0.5)
    for v in leaves:
        s = np.float32(s + np.float32(v))
    assert got == float(s)
    assert got != 0.5 + sum(leaves)  # f64 arithmetic would differ


def test_parsers_agree_json_ubj_and_reject_garbage():
    a = model_parse.parse_xgboost(synth.xgboost_model_json(5, 6, depth=4, seed=1))
    b = model_parse.parse_xgboost(synth.xgboost_model_ubj(5, 6, depth=4, seed=1))
    for ta, tb in zip(a["trees"], b["trees"]):
        for k in ta:
            assert np.array_equal(ta[k], tb[k])
    with pytest.raises(ValueError):
        model_parse.parse_xgboost(b"binf\x00\x00")
    with pytest.raises(ValueError):
        model_parse.parse_lightgbm_text(b"hello")


def test_metarank_blob_roundtrip():
    names = ["popularity", "ctr", "genre"]
    inner = synth.lightgbm_model_text(2, 3, seed=1)
    for v in (2, 3):
        ver, got_names, kind, booster = model_parse.parse_metarank_blob(synth.metarank_model_blob(names, 0, inner, v))
        assert (ver, got_names, kind, booster) == (v, names, 0, inner)
    with pytest.raises(ValueError):
        model_parse.parse_metarank_blob(b"\x07" + b"\0" * 20)


def test_rank_order_is_stable_descending_total_order():
    s = [0.1, 0.5, 0.5, NAN, -0.0, 0.0, 2.0, -math.inf, math.inf]
    assert oracle.rank_order(s).tolist() == [8, 6, 1, 2, 0, 5, 4, 7, 3]
    assert oracle.rank_order([]).tolist() == []
    assert oracle.rank_order([1.0] * 5).tolist() == [0, 1, 2, 3, 4]


def test_openmp_path_equals_scalar_path():
    blob = synth.lightgbm_model_text(50, 12, seed=3, cat_features={4: 20}, zero_missing=True)
    X = synth.feature_matrix(3000, 12, seed=4)
    ob = oracle.OracleBooster(0, blob)
    assert np.array_equal(ob.predictMat(X, *X.shape, threads=1), ob.predictMat(X, *X.shape, threads=0))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c_oracle_agrees_with_independent_python_evaluator(seed):
    """Two independent restatements (C over flat arrays, pure Python over the parsed dict) must agree
    bit for bit on random models with every node kind — the only cross-check available for a boundary
    the reference never asserts a value for."""
    rng = np.random.Generator(np.random.PCG64(seed))
    blob = synth.lightgbm_model_text(25, 9, seed=seed, cat_features={3: 50}, zero_missing=True, stump_every=6)
    X = synth.feature_matrix(120, 9, seed=seed + 10)
    X[:, 3] = rng.integers(-3, 60, 120)
    X[rng.random(120) < 0.1, 3] = np.nan
    X[rng.random(120) < 0.2, 2] = 0.0
    want = model_parse.predict_python(model_parse.parse_lightgbm_text(blob), X)
    assert np.array_equal(oracle.OracleBooster(0, blob).predictMat(X, *X.shape), want)
    xb = synth.xgboost_model_json(20, 7, depth=5, seed=seed, full=False)
    X7 = synth.feature_matrix(100, 7, seed=seed + 20)
    want = model_parse.predict_python(model_parse.parse_xgboost(xb), X7)
    assert np.array_equal(oracle.OracleBooster(1, xb).predictMat(X7, *X7.shape), want)


def test_oracle_traversal_and_sum_against_an_independent_gbdt_library():
    """The reference's scorer (LightGBM behind ltrlib) is not available here, so the GBDT oracle stays "parity
    unpinned" against it.  What CAN be pinned is the part every GBDT predictor shares — `x <= threshold` goes
    left, leaf values added in tree order in f64 — against an independent, widely used implementation that IS
    in the image: scikit-learn's GradientBoostingRegressor.  Its trees are written out as a LightGBM model text
    (first tree = the constant init prediction, leaf = learning_rate * value, same product sklearn forms) and the
    oracle's predictions must equal `predict`'s bit for bit on float32-representable inputs (sklearn traverses
    in float32).  LightGBM's own missing-value / categorical rules are NOT covered by this (no NaN support in
    GradientBoostingRegressor)."""
    sklearn_ensemble = pytest.importorskip("sklearn.ensemble")
    from metarank_b200 import synth

    rng = np.random.Generator(np.random.PCG64(12))
    n_feat = 7
    X = rng.normal(size=(900, n_feat)).astype(np.float32).astype(np.float64)
    y = 2 * X[:, 0] + np.sin(3 * X[:, 1]) + 1.5 * (X[:, 2] > 0.3) - X[:, 3] * X[:, 4] + 0.1 * rng.normal(size=len(X))
    gb = sklearn_ensemble.GradientBoostingRegressor(n_estimators=60, max_depth=5, learning_rate=0.1, subsample=0.8,
                                                    random_state=0).fit(X, y)
    lr = gb.learning_rate
    trees = [synth._Tree([], [], [], [], [], [float(gb.init_.constant_[0, 0])], [0], [])]
    for est in gb.estimators_[:, 0]:
        t = est.tree_
        internal = [i for i in range(t.node_count) if t.children_left[i] != -1]
        leaves = [i for i in range(t.node_count) if t.children_left[i] == -1]
        iid = {n: k for k, n in enumerate(internal)}
        lid = {n: k for k, n in enumerate(leaves)}
        ref = lambda c: iid[c] if c in iid else ~lid[c]  # noqa: E731
        if not internal:
            trees.append(synth._Tree([], [], [], [], [], [lr * float(t.value[0, 0, 0])], [0], []))
            continue
        trees.append(synth._Tree(
            split_feature=[int(t.feature[n]) for n in internal], threshold=[float(t.threshold[n]) for n in internal],
            decision_type=[2] * len(internal),  # numerical, default left, missing type None
            left_child=[ref(int(t.children_left[n])) for n in internal],
            right_child=[ref(int(t.children_right[n])) for n in internal],
            leaf_value=[lr * float(t.value[n, 0, 0]) for n in leaves], cat_boundaries=[0], cat_threshold=[]))
    blob = synth.lightgbm_text_from_trees(trees, n_feat, shrinkage=lr)
    Xt = rng.normal(size=(4000, n_feat)).astype(np.float32).astype(np.float64)
    Xt[:50] = X[:50]
    thr = np.concatenate([np.asarray(t.threshold, dtype=np.float64) for t in trees if t.threshold])
    Xt[50:50 + min(200, len(thr)), 0] = thr[:200].astype(np.float32)  # values at / next to thresholds
    got = oracle.OracleBooster(0, blob).predictMat(np.ascontiguousarray(Xt), len(Xt), n_feat, threads=0)
    want = gb.predict(Xt)
    assert np.array_equal(got, want), float(np.max(np.abs(got - want)))
    # and the independent Python evaluator of the same blob agrees as well
    from oracle import model_parse as mp
    assert np.array_equal(mp.predict_python(mp.parse_lightgbm_text(blob), Xt[:300]), want[:300])


def test_oracle_nan_routing_against_hist_gradient_boosting():
    """Same idea for the missing-value rule: scikit-learn's HistGradientBoostingRegressor predicts on raw f64 values
    with `x <= threshold` and a per-node `missing_go_to_left` for NaN — the semantics of a LightGBM node with
    missing type NaN and the default-left bit.  Its trees, exported as LightGBM text (decision_type = NaN-missing
    | default-left), must score bit-identically through the oracle on inputs with NaNs."""
    sklearn_ensemble = pytest.importorskip("sklearn.ensemble")
    from metarank_b200 import synth

    rng = np.random.Generator(np.random.PCG64(21))
    n_feat = 6
    X = rng.normal(size=(3000, n_feat))
    X[rng.random(X.shape) < 0.15] = np.nan
    y = np.nan_to_num(X[:, 0]) * 2 + np.where(np.isnan(X[:, 1]), 1.0, np.sin(3 * np.nan_to_num(X[:, 1]))) + \
        (np.nan_to_num(X[:, 2]) > 0.3) * 1.5 + 0.1 * rng.normal(size=len(X))
    hgb = sklearn_ensemble.HistGradientBoostingRegressor(max_iter=50, max_leaf_nodes=15, learning_rate=0.1,
                                                         early_stopping=False, random_state=0).fit(X, y)
    trees = [synth._Tree([], [], [], [], [], [float(np.ravel(hgb._baseline_prediction)[0])], [0], [])]
    n_left = n_right = 0
    for (pred,) in hgb._predictors:
        nodes = pred.nodes
        internal = [i for i in range(len(nodes)) if not nodes[i]["is_leaf"]]
        leaves = [i for i in range(len(nodes)) if nodes[i]["is_leaf"]]
        iid = {n: k for k, n in enumerate(internal)}
        lid = {n: k for k, n in enumerate(leaves)}
        ref = lambda c: iid[c] if c in iid else ~lid[c]  # noqa: E731
        assert not any(nodes[n]["is_categorical"] for n in internal)
        mgl = [bool(nodes[n]["missing_go_to_left"]) for n in internal]
        n_left += sum(mgl); n_right += len(mgl) - sum(mgl)
        trees.append(synth._Tree(
            split_feature=[int(nodes[n]["feature_idx"]) for n in internal],
            threshold=[float(nodes[n]["num_threshold"]) for n in internal],
            decision_type=[(2 << 2) | (2 if m else 0) for m in mgl],  # missing type NaN (+ default left)
            left_child=[ref(int(nodes[n]["left"])) for n in internal],
            right_child=[ref(int(nodes[n]["right"])) for n in internal],
            leaf_value=[float(nodes[n]["value"]) for n in leaves], cat_boundaries=[0], cat_threshold=[]))
    assert n_left > 0 and n_right > 0  # both NaN directions occur
    blob = synth.lightgbm_text_from_trees(trees, n_feat, shrinkage=0.1)
    Xt = rng.normal(size=(5000, n_feat))
    Xt[rng.random(Xt.shape) < 0.2] = np.nan
    thr = np.concatenate([np.asarray(t.threshold) for t in trees if t.threshold])
    Xt[:min(300, len(thr)), 1] = thr[:300]  # exactly on thresholds
    got = oracle.OracleBooster(0, blob).predictMat(np.ascontiguousarray(Xt), len(Xt), n_feat, threads=0)
    want = hgb.predict(Xt)
    assert np.array_equal(got, want), float(np.nanmax(np.abs(got - want)))


def test_xgboost_binary_encoding_reads_like_json():
    """The oracle's reader of XGBoost's deprecated binary encoding (xgboost4j's toByteArray() up to 2.0): the same synthetic
    ensemble as JSON, UBJSON and binary — with the `binf` prefix, with pruned nodes left in the arrays, without the
    attribute block — parses to the same arrays and scores to the same bits, through the C port and the pure-Python walk."""
    j = synth.xgboost_model_json(25, 12, depth=5, seed=44, full=False)
    pj = model_parse.parse_xgboost(j)
    X = synth.feature_matrix(300, 12, seed=45)
    X[::9, 2] = np.nan
    want = oracle.OracleBooster(1, j).predictMat(X, 300, 12)
    assert np.array_equal(oracle.OracleBooster(1, synth.xgboost_model_ubj(25, 12, depth=5, seed=44, full=False)).predictMat(X, 300, 12), want)
    for kw in ({}, {"magic": True}, {"deleted": 3}, {"with_attributes": False}, {"objective": "rank:pairwise"}):
        b = synth.xgboost_model_binary(25, 12, depth=5, seed=44, full=False, **kw)
        pb = model_parse.parse_xgboost(b)
        assert pb["n_features"] == pj["n_features"] and pb["base_score"] == pj["base_score"] and len(pb["trees"]) == len(pj["trees"])
        nd = kw.get("deleted", 0)
        for tb, tj in zip(pb["trees"], pj["trees"]):
            n = len(tj["left"])
            assert len(tb["left"]) == n + nd
            for k in ("left", "right", "split_index", "split_cond", "default_left"):
                assert np.array_equal(tb[k][:n], tj[k]), k
        assert np.array_equal(oracle.OracleBooster(1, b).predictMat(X, 300, 12), want)
        assert np.array_equal(model_parse.predict_python(pb, X[:40]), want[:40])
    with pytest.raises(ValueError):
        model_parse.parse_xgboost(b"bs64\tAAAA")
    with pytest.raises(ValueError):
        model_parse.parse_xgboost(b"not a model at all" * 20)
