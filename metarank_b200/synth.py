"""Seeded synthetic models and inputs for the /rank hot path.

Nothing here is an oracle or a kernel: it only *writes* LightGBM model text /
XGBoost model JSON in the public formats those libraries save (the blobs the
reference hands to ``LightGBMBooster(bytes)`` / ``XGBoostBooster(bytes)``,
reference S/ml/rank/LambdaMARTRanker.scala:228-232) and draws feature matrices
of the shapes BASELINE.json names.  All randomness is
``numpy.random.Generator(PCG64(seed))``.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass

import numpy as np

__all__ = [
    "feature_matrix", "column_kinds", "lightgbm_model_text", "xgboost_model_json",
    "xgboost_model_ubj", "metarank_model_blob", "CONFIGS",
]

# BASELINE.json configs (rows × features × trees); see SURVEY.md §8 shorthand.
CONFIGS = {
    "C2": dict(items=100, features=30, trees=500, leaves=16, max_depth=8, kind="lightgbm"),
    "C4": dict(items=256, features=16, trees=200, depth=6, kind="xgboost"),
    "C5": dict(items=10_000, features=64, trees=2000, leaves=16, max_depth=8, kind="lightgbm"),
}

_KINDS = ("normal", "lognormal", "count", "ctr", "nan5")


def column_kinds(n_features: int) -> list[str]:
    """Columns cycle through {N(0,1), logN(0,1), U{0..1000}, Beta(2,50), N(0,1)+5% NaN}."""
    return [_KINDS[j % len(_KINDS)] for j in range(n_features)]


def _draw_column(rng: np.random.Generator, kind: str, n: int) -> np.ndarray:
    if kind == "normal":
        return rng.standard_normal(n)
    if kind == "lognormal":
        return rng.lognormal(0.0, 1.0, n)
    if kind == "count":
        return rng.integers(0, 1001, n).astype(np.float64)
    if kind == "ctr":
        return rng.beta(2.0, 50.0, n)
    if kind == "nan5":
        x = rng.standard_normal(n)
        x[rng.random(n) < 0.05] = np.nan
        return x
    raise ValueError(kind)


def feature_matrix(rows: int, n_features: int, seed: int = 42) -> np.ndarray:
    """Row-major f64[rows × n_features], the layout of ltrlib ``Query.values``
    (reference S/flow/ClickthroughQuery.scala:50-74)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    kinds = column_kinds(n_features)
    out = np.empty((rows, n_features), dtype=np.float64)
    for j, k in enumerate(kinds):
        out[:, j] = _draw_column(rng, k, rows)
    return out


def _quantile_threshold(rng: np.random.Generator, kind: str) -> float:
    """A split threshold = a uniformly drawn quantile of the column's distribution."""
    sample = _draw_column(rng, "normal" if kind == "nan5" else kind, 64)
    v = float(sample[int(rng.integers(0, 64))])
    if kind == "count":
        v = float(np.floor(v)) + 0.5  # LightGBM puts thresholds between observed values
    return v


def _fmt(x: float) -> str:
    return repr(float(x)) if np.isfinite(x) else ("nan" if np.isnan(x) else ("inf" if x > 0 else "-inf"))


@dataclass
class _Tree:
    split_feature: list
    threshold: list
    decision_type: list
    left_child: list
    right_child: list
    leaf_value: list
    cat_boundaries: list
    cat_threshold: list


def _grow_lightgbm_tree(rng, n_features, kinds, num_leaves, max_depth, cat_features, shrinkage,
                        zero_missing=False):
    """Random leaf-wise growth with LightGBM's node numbering: split #s turns leaf `l`
    into internal node s whose left child keeps leaf id l and right child is leaf s+1."""
    leaf_depth = {0: 0}
    leaf_parent = {0: (-1, 0)}  # leaf -> (parent node, is_right)
    sf, thr, dt, lc, rc = [], [], [], [], []
    cat_b, cat_t = [0], []
    n_leaves = 1
    while n_leaves < num_leaves:
        cands = [l for l, d in leaf_depth.items() if max_depth <= 0 or d < max_depth]
        if not cands:
            break
        leaf = cands[int(rng.integers(0, len(cands)))]
        node = n_leaves - 1
        f = int(rng.integers(0, n_features))
        default_left = int(rng.random() < 0.5)
        if f in cat_features:
            n_cat = cat_features[f]
            n_words = (n_cat + 31) // 32
            words = [int(rng.integers(0, 2**32)) for _ in range(n_words)]
            if n_cat % 32:  # LightGBM's bitset ends at the largest category of the set: nothing beyond the feature's range
                words[-1] &= (1 << (n_cat % 32)) - 1
            cat_idx = len(cat_b) - 1
            cat_t.extend(words)
            cat_b.append(len(cat_t))
            sf.append(f); thr.append(float(cat_idx)); dt.append(1)  # categorical, missing None
        else:
            kind = kinds[f]
            t = _quantile_threshold(rng, kind)
            if kind == "nan5":
                missing = 2
            elif zero_missing and kind == "count":
                missing = 1
            else:
                missing = 0
            sf.append(f); thr.append(t); dt.append((default_left << 1) | (missing << 2))
        lc.append(~leaf); rc.append(~n_leaves)
        p, is_right = leaf_parent[leaf]
        if p >= 0:
            if is_right:
                rc[p] = node
            else:
                lc[p] = node
        d = leaf_depth[leaf] + 1
        leaf_depth[leaf] = d
        leaf_depth[n_leaves] = d
        leaf_parent[leaf] = (node, 0)
        leaf_parent[n_leaves] = (node, 1)
        n_leaves += 1
    lv = (rng.standard_normal(n_leaves) * 0.1 * shrinkage).tolist()
    return _Tree(sf, thr, dt, lc, rc, lv, cat_b, cat_t)


def lightgbm_model_text(n_trees: int, n_features: int, num_leaves: int = 16, max_depth: int = 8,
                        seed: int = 1234, cat_features: dict | None = None,
                        shrinkage: float = 0.1, zero_missing: bool = False,
                        stump_every: int = 0) -> bytes:
    """LightGBM ``SaveModelToString`` text (v4 header) for a lambdarank model.

    cat_features: {column -> n_categories} columns split with categorical bitsets.
    stump_every: every k-th tree is a single-leaf tree (num_leaves=1), which real
    models contain after early convergence.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    kinds = column_kinds(n_features)
    cat_features = cat_features or {}
    trees = []
    for t in range(n_trees):
        if stump_every and t % stump_every == stump_every - 1:
            trees.append(_Tree([], [], [], [], [], [float(rng.standard_normal() * 0.01)], [0], []))
        else:
            trees.append(_grow_lightgbm_tree(rng, n_features, kinds, num_leaves, max_depth, cat_features,
                                             shrinkage, zero_missing))
    return lightgbm_text_from_trees(trees, n_features, shrinkage)


def lightgbm_text_from_trees(trees: list, n_features: int, shrinkage: float = 0.1) -> bytes:
    """LightGBM model text for explicit trees (`_Tree`: split_feature, threshold, decision_type, left_child,
    right_child with ~leaf for leaves, leaf_value, cat_boundaries, cat_threshold)."""
    blocks = []
    for t, tr in enumerate(trees):
        nl = len(tr.leaf_value)
        n_cat = len(tr.cat_boundaries) - 1
        lines = [f"Tree={t}", f"num_leaves={nl}", f"num_cat={n_cat}"]
        if nl > 1:
            ni = nl - 1
            lines += [
                "split_feature=" + " ".join(map(str, tr.split_feature)),
                "split_gain=" + " ".join("1" for _ in range(ni)),
                "threshold=" + " ".join(_fmt(x) for x in tr.threshold),
                "decision_type=" + " ".join(map(str, tr.decision_type)),
                "left_child=" + " ".join(map(str, tr.left_child)),
                "right_child=" + " ".join(map(str, tr.right_child)),
                "leaf_value=" + " ".join(_fmt(x) for x in tr.leaf_value),
                "leaf_weight=" + " ".join("1" for _ in range(nl)),
                "leaf_count=" + " ".join("10" for _ in range(nl)),
                "internal_value=" + " ".join("0" for _ in range(ni)),
                "internal_weight=" + " ".join("1" for _ in range(ni)),
                "internal_count=" + " ".join("20" for _ in range(ni)),
            ]
            if n_cat > 0:
                lines += [
                    "cat_boundaries=" + " ".join(map(str, tr.cat_boundaries)),
                    "cat_threshold=" + " ".join(map(str, tr.cat_threshold)),
                ]
        else:
            lines += ["leaf_value=" + _fmt(tr.leaf_value[0])]
        lines += ["is_linear=0", f"shrinkage={_fmt(shrinkage)}", "", ""]
        blocks.append("\n".join(lines))
    header = [
        "tree", "version=v4", "num_class=1", "num_tree_per_iteration=1", "label_index=0",
        f"max_feature_idx={n_features - 1}", "objective=lambdarank",
        "feature_names=" + " ".join(f"Column_{j}" for j in range(n_features)),
        "feature_infos=" + " ".join("[-10:10]" for _ in range(n_features)),
        "tree_sizes=" + " ".join(str(len(b) + 1) for b in blocks),
        "", "",
    ]
    tail = ["end of trees", "", "feature_importances:", "", "parameters:", "[boosting: gbdt]",
            "[objective: lambdarank]", "end of parameters", "", "pandas_categorical:null", ""]
    return ("\n".join(header) + "\n".join(blocks) + "\n" + "\n".join(tail)).encode("utf-8")


# --------------------------------------------------------------------------- XGBoost

def _grow_xgb_tree(rng, n_features, kinds, depth, full, shrinkage):
    """XGBoost RegTree arrays (BFS ids, leaf <=> left_children == -1)."""
    left, right, sidx, scond, dleft = [], [], [], [], []

    def new_node():
        left.append(-1); right.append(-1); sidx.append(0); scond.append(0.0); dleft.append(0)
        return len(left) - 1

    root = new_node()
    frontier = [(root, 0)]
    while frontier:
        nid, d = frontier.pop(0)
        split = d < depth and (full or d < 2 or rng.random() < 0.75)
        if split:
            f = int(rng.integers(0, n_features))
            sidx[nid] = f
            scond[nid] = float(np.float32(_quantile_threshold(rng, kinds[f])))
            dleft[nid] = int(rng.random() < 0.5)
            l = new_node(); r = new_node()
            left[nid] = l; right[nid] = r
            frontier.append((l, d + 1)); frontier.append((r, d + 1))
        else:
            scond[nid] = float(np.float32(rng.standard_normal() * 0.1 * shrinkage))
    return left, right, sidx, scond, dleft


def _xgb_model_dict(n_trees, n_features, depth, seed, full, base_score, shrinkage):
    rng = np.random.Generator(np.random.PCG64(seed))
    kinds = column_kinds(n_features)
    trees = []
    for t in range(n_trees):
        l, r, si, sc, dl = _grow_xgb_tree(rng, n_features, kinds, depth, full, shrinkage)
        n = len(l)
        parents = [2147483647] * n
        for i in range(n):
            if l[i] >= 0:
                parents[l[i]] = i; parents[r[i]] = i
        trees.append({
            "base_weights": [0.0] * n, "categories": [], "categories_nodes": [],
            "categories_segments": [], "categories_sizes": [],
            "default_left": dl, "id": t, "left_children": l, "loss_changes": [0.0] * n,
            "parents": parents, "right_children": r, "split_conditions": sc,
            "split_indices": si, "split_type": [0] * n, "sum_hessian": [1.0] * n,
            "tree_param": {"num_deleted": "0", "num_feature": str(n_features),
                           "num_nodes": str(n), "size_leaf_vector": "1"},
        })
    bs = np.float32(base_score)
    return {
        "learner": {
            "attributes": {}, "feature_names": [], "feature_types": [],
            "gradient_booster": {
                "model": {
                    "gbtree_model_param": {"num_parallel_tree": "1", "num_trees": str(n_trees)},
                    "iteration_indptr": list(range(n_trees + 1)),
                    "tree_info": [0] * n_trees, "trees": trees,
                },
                "name": "gbtree",
            },
            "learner_model_param": {
                "base_score": np.format_float_scientific(bs, unique=True, exp_digits=1).upper(),
                "boost_from_average": "1", "num_class": "0", "num_feature": str(n_features),
                "num_target": "1",
            },
            "objective": {"name": "rank:ndcg", "lambdarank_param": {}},
        },
        "version": [2, 1, 4],
    }


def xgboost_model_json(n_trees: int, n_features: int, depth: int = 6, seed: int = 1234,
                       full: bool = True, base_score: float = 0.5, shrinkage: float = 0.1) -> bytes:
    """XGBoost ≥1.x JSON model (``save_raw("json")``) for a rank:ndcg gbtree."""
    d = _xgb_model_dict(n_trees, n_features, depth, seed, full, base_score, shrinkage)
    return json.dumps(d, separators=(",", ":")).encode("utf-8")


def _ubj(obj, float_ctx: str | None = None) -> bytes:
    """Minimal UBJSON writer in the dialect XGBoost emits (typed arrays for numeric lists)."""
    if isinstance(obj, dict):
        out = [b"{"]
        for k, v in obj.items():
            kb = k.encode()
            out.append(b"L" + struct.pack(">q", len(kb)) + kb)
            out.append(_ubj(v, k))
        out.append(b"}")
        return b"".join(out)
    if isinstance(obj, list):
        if obj and all(isinstance(x, float) for x in obj):
            return b"[$d#L" + struct.pack(">q", len(obj)) + b"".join(struct.pack(">f", x) for x in obj)
        if obj and all(isinstance(x, int) and not isinstance(x, bool) for x in obj):
            if float_ctx == "default_left" or float_ctx == "split_type":
                return b"[$U#L" + struct.pack(">q", len(obj)) + bytes(obj)
            if float_ctx in ("iteration_indptr",) or max(abs(x) for x in obj) >= 2**31:
                return b"[$L#L" + struct.pack(">q", len(obj)) + b"".join(struct.pack(">q", x) for x in obj)
            return b"[$l#L" + struct.pack(">q", len(obj)) + b"".join(struct.pack(">i", x) for x in obj)
        return b"[" + b"".join(_ubj(x) for x in obj) + b"]"
    if isinstance(obj, str):
        b = obj.encode()
        return b"SL" + struct.pack(">q", len(b)) + b
    if isinstance(obj, bool):
        return b"T" if obj else b"F"
    if isinstance(obj, int):
        return b"l" + struct.pack(">i", obj) if abs(obj) < 2**31 else b"L" + struct.pack(">q", obj)
    if isinstance(obj, float):
        return b"d" + struct.pack(">f", obj)
    if obj is None:
        return b"Z"
    raise TypeError(type(obj))


def xgboost_model_binary(n_trees: int, n_features: int, depth: int = 6, seed: int = 1234, full: bool = True,
                         base_score: float = 0.5, shrinkage: float = 0.1, magic: bool = False, deleted: int = 0,
                         objective: str = "rank:ndcg", with_attributes: bool = True) -> bytes:
    """Same model as :func:`xgboost_model_json` in XGBoost's deprecated BINARY encoding — what ``Booster.toByteArray()`` /
    ``save_raw()`` write by default up to XGBoost 2.0, i.e. what an xgboost4j-trained Metarank model holds
    (LearnerModelParamLegacy 136 B, two length-prefixed names, GBTreeModelParam 160 B, per tree TreeParam 148 B + 20-byte
    nodes + 16-byte stats, tree_info; little-endian).  magic: the 0.x ``binf`` prefix; deleted: pruned (unreachable) nodes
    appended to every tree, as RegTree keeps them after pruning; with_attributes: the attribute block that follows."""
    d = _xgb_model_dict(n_trees, n_features, depth, seed, full, base_score, shrinkage)
    out = [b"binf"] if magic else []
    out.append(struct.pack("<fIiiiIIIi", float(np.float32(base_score)), n_features, 0, 1 if with_attributes else 0, 0, 1, 7, 1, 1)
               + b"\0" * (25 * 4))
    for name in (objective, "gbtree"):
        out.append(struct.pack("<Q", len(name)) + name.encode())
    out.append(struct.pack("<iiiiqii", n_trees, 1, n_features, 0, 0, 1, 0) + b"\0" * (32 * 4))
    for t in d["learner"]["gradient_booster"]["model"]["trees"]:
        l, r, si, sc, dl, par = (t[k] for k in ("left_children", "right_children", "split_indices", "split_conditions",
                                               "default_left", "parents"))
        n = len(l)
        out.append(struct.pack("<iiiiIi", 1, n + deleted, deleted, 0, n_features, 0) + b"\0" * (31 * 4))
        for i in range(n):
            parent = -1 if par[i] == 2147483647 else (par[i] | (0x80000000 if l[par[i]] == i else 0))
            sindex = (si[i] | (0x80000000 if dl[i] else 0)) if l[i] >= 0 else 0
            out.append(struct.pack("<IiiIf", parent & 0xFFFFFFFF, l[i], r[i] if l[i] >= 0 else 0, sindex, sc[i]))
        for _ in range(deleted):
            out.append(struct.pack("<IiiIf", 0xFFFFFFFF, -1, 0, 0xFFFFFFFF, 0.0))
        out.append(struct.pack("<fffi", 0.0, 1.0, 0.0, 0) * (n + deleted))
    out.append(struct.pack(f"<{n_trees}i", *([0] * n_trees)))
    if with_attributes:
        k, v = b"best_iteration", str(n_trees - 1).encode()
        out.append(struct.pack("<Q", 1) + struct.pack("<Q", len(k)) + k + struct.pack("<Q", len(v)) + v)
    return b"".join(out)


def xgboost_model_ubj(n_trees: int, n_features: int, depth: int = 6, seed: int = 1234,
                      full: bool = True, base_score: float = 0.5, shrinkage: float = 0.1) -> bytes:
    """Same model as :func:`xgboost_model_json`, UBJSON-encoded (``save_raw("ubj")``)."""
    return _ubj(_xgb_model_dict(n_trees, n_features, depth, seed, full, base_score, shrinkage))


def metarank_model_blob(feature_names: list[str], booster_kind: int, booster_bytes: bytes,
                        version: int = 3) -> bytes:
    """Metarank's model framing (reference S/ml/rank/LambdaMARTRanker.scala:367-389):
    byte version, int nFeatures, writeUTF names, byte boosterType, int size, bytes,
    (v3) int nWarmup — big-endian DataOutputStream."""
    out = [struct.pack(">b", version), struct.pack(">i", len(feature_names))]
    for n in feature_names:
        b = n.encode("utf-8")  # modified UTF-8 == UTF-8 for BMP text without NUL
        out.append(struct.pack(">H", len(b)) + b)
    out.append(struct.pack(">b", booster_kind))
    out.append(struct.pack(">i", len(booster_bytes)))
    out.append(booster_bytes)
    if version >= 3:
        out.append(struct.pack(">i", 0))
    return b"".join(out)


# --------------------------------------------------------------------------- ranklens-shaped state (config #3)

RANKLENS_GENRES = ["drama", "comedy", "thriller", "action", "adventure", "romance", "crime", "science fiction",
                   "fantasy", "family", "horror", "mystery", "animation", "history", "music"]


def ranklens_config():
    """The ranklens feature set of the reference (src/test/resources/ranklens/config.yml:27-58 and the
    `features:` section), as the JSON-able dicts Metarank's config decoders accept.  24 columns."""
    num = lambda n, f: dict(name=n, type="number", scope="item", source=f"metadata.{f}")  # noqa: E731
    rate = dict(type="rate", top="click", bottom="impression", bucket="24h", periods=[7, 30])
    feats = [
        dict(rate, name="ctr_tag", scope="item.tag"),
        dict(rate, name="ctr_genre", scope="item.genre"),
        dict(name="position", type="position", position=5),
        num("popularity", "popularity"), num("vote_avg", "vote_avg"), num("vote_cnt", "vote_cnt"),
        num("budget", "budget"), num("release_date", "release_date"), num("runtime", "runtime"),
        dict(name="title_length", type="word_count", source="metadata.title", scope="item"),
        dict(name="genre", type="string", scope="item", source="metadata.genres", encode="index",
             values=RANKLENS_GENRES),
        dict(rate, name="ctr", normalize={"weight": 10}),
        dict(name="profile", type="interacted_with", interaction="click",
             field=["item.genres", "item.actors", "item.tags", "item.director"], scope="session", count=100,
             duration="24h"),
        dict(name="divers_genres", type="diversity", source="metadata.genres"),
        dict(name="divers_actors", type="diversity", source="metadata.actors"),
        dict(name="divers_tags", type="diversity", source="metadata.tags"),
        dict(name="divers_year", type="diversity", source="metadata.release_date"),
        dict(name="divers_popularity", type="diversity", source="metadata.popularity"),
        dict(name="visitor_click_count", type="interaction_count", interaction="click", scope="session"),
        dict(name="global_item_click_count", type="interaction_count", interaction="click", scope="item"),
        dict(name="day_item_click_count", type="window_count", interaction="click", scope="item", bucket="24h",
             periods=[1]),
    ]
    model = ["popularity", "vote_avg", "vote_cnt", "budget", "release_date", "runtime", "title_length", "genre",
             "ctr", "profile", "position", "divers_genres", "divers_actors", "divers_tags", "divers_year",
             "divers_popularity", "ctr_tag", "ctr_genre"]
    return feats, model


def _zipf_draw(rng, n_vocab, k):
    p = 1.0 / np.arange(1, n_vocab + 1) ** 1.2
    p /= p.sum()
    return rng.choice(n_vocab, size=k, replace=False, p=p)


def ranklens_state(n_items: int = 1000, n_sessions: int = 100, seed: int = 45, missing: float = 0.03,
                   with_field_scalars: bool = True):
    """FeatureValues (plain Key -> Value dict, see metarank_b200/features.py) of a ranklens-shaped
    catalogue, SURVEY.md §8d: per item 6 numeric scalars, title word count, 1-3 genres / 3-8 actors /
    2-10 tags / 1 director drawn Zipf(1.2) from vocabularies of 20/5000/1000/2000, click/impression
    windows for periods [7, 30] per item, per tag/genre value and globally; sessions with 0-100
    clicked items.  `missing` = fraction of state entries left out (exercises NaN paths)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    st = {}
    drop = lambda: rng.random() < missing  # noqa: E731
    genres_v = RANKLENS_GENRES + [f"g{i}" for i in range(5)]
    item_ids = [f"m{i}" for i in range(n_items)]
    tag_first, genre_first = {}, {}
    for i, it in enumerate(item_ids):
        sc = ("item", it)
        nums = dict(popularity=rng.lognormal(2, 1), vote_avg=rng.uniform(0, 10), vote_cnt=float(rng.integers(0, 20000)),
                    budget=float(rng.integers(0, 3 * 10**8)), release_date=float(rng.integers(0, 1_600_000_000)),
                    runtime=float(rng.integers(40, 240)))
        for k, v in nums.items():
            if not drop():
                st[(sc, k)] = ("scalar", float(v))
        if not drop():
            st[(sc, "title_length")] = ("scalar", float(rng.integers(1, 9)))
        genres = [genres_v[j] for j in _zipf_draw(rng, len(genres_v), int(rng.integers(1, 4)))]
        actors = [f"a{j}" for j in _zipf_draw(rng, 5000, int(rng.integers(3, 9)))]
        tags = [f"t{j}" for j in _zipf_draw(rng, 1000, int(rng.integers(2, 11)))]
        director = [f"d{j}" for j in _zipf_draw(rng, 2000, 1)]
        if not drop():
            st[(sc, "genre")] = ("scalar", genres)
        for fld, vals in (("genres", genres), ("actors", actors), ("tags", tags), ("director", director)):
            if not drop():
                st[(sc, f"profile_{fld}")] = ("scalar", vals)
        for fld, vals in (("genres", genres), ("actors", actors), ("tags", tags)):
            if not drop():
                st[(sc, f"divers_{fld}")] = ("scalar", vals)
        if (sc, "release_date") in st and not drop():
            st[(sc, "divers_year")] = st[(sc, "release_date")]
        if (sc, "popularity") in st and not drop():
            st[(sc, "divers_popularity")] = st[(sc, "popularity")]
        imp = rng.poisson([40.0, 160.0])
        clk = np.minimum(rng.poisson([3.0, 12.0]), imp)
        if not drop():
            st[(sc, "ctr_click")] = ("pcounter", [int(x) for x in clk])
        if not drop():
            st[(sc, "ctr_impression")] = ("pcounter", [int(x) for x in imp])
        if with_field_scalars and not drop():
            # the real ranklens events carry no singular `tag`/`genre` field, so on the real data
            # these two scalars never exist and ctr_tag/ctr_genre are NaN;
            # synthetic data exercises the populated path too
            st[(sc, "ctr_tag_field")] = ("scalar", tags[0])
            st[(sc, "ctr_genre_field")] = ("scalar", genres[0])
            tag_first[tags[0]] = 1
            genre_first[genres[0]] = 1
    for name, vocab in (("tag", tag_first), ("genre", genre_first)):
        for v in vocab:
            imp = rng.poisson([400.0, 1600.0])
            clk = np.minimum(rng.poisson([30.0, 120.0]), imp)
            sc = ("field", name, v)
            if not drop():
                st[(sc, f"ctr_{name}_click")] = ("pcounter", [int(x) for x in clk])
            if not drop():
                st[(sc, f"ctr_{name}_impression")] = ("pcounter", [int(x) for x in imp])
    st[(("global",), "ctr_click_norm")] = ("pcounter", [int(n_items * 3), int(n_items * 12)])
    st[(("global",), "ctr_impression_norm")] = ("pcounter", [int(n_items * 40), int(n_items * 160)])
    sessions = [f"s{i}" for i in range(n_sessions)]
    for s in sessions:
        n = int(min(100, rng.geometric(0.08) - 1))
        if n > 0:
            hist = [item_ids[int(j)] for j in rng.integers(0, n_items, n)]
            if rng.random() < 0.1:
                hist.append("unknown-item")  # an interacted item that has no state
            st[(("session", s), "profile_interactions")] = ("blist", hist)
    return st, item_ids, sessions


def ranklens_requests(item_ids, sessions, n_requests: int, items_per_request: int, seed: int = 46,
                      unknown: float = 0.02):
    """LatencyBenchmark.RandomRequest-shaped RankingEvents (T/util/benchmark/LatencyBenchmark.scala:44-58):
    distinct random items + a random session."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for r in range(n_requests):
        k = min(items_per_request, len(item_ids))
        ids = [item_ids[int(j)] for j in rng.choice(len(item_ids), size=k, replace=False)]
        ids = [f"nope{r}_{j}" if rng.random() < unknown else x for j, x in enumerate(ids)]
        sess = sessions[int(rng.integers(0, len(sessions)))] if rng.random() > 0.05 else None
        out.append(dict(event="ranking", id=f"req{r}", timestamp=0, user=sess, session=sess, fields=[],
                        items=[dict(id=i, fields=[("relevancy", 0.0)]) for i in ids]))
    return out
