// Parsers for the booster blobs + packing into TMA-stageable chunks.  See gbdt_model.h.
//
// Format notes (public formats of the third-party boosters; the arithmetic is not in
// the reference tree — see oracle/gbdt_oracle.c for the citation chain):
//  * LightGBM model text: "key=value" header, then "Tree=<i>" blocks with num_leaves,
//    num_cat, split_feature, threshold, decision_type, left_child, right_child,
//    leaf_value, [cat_boundaries, cat_threshold], terminated by "end of trees".
//  * XGBoost JSON/UBJSON: learner.gradient_booster.model.trees[] with left_children,
//    right_children, split_indices, split_conditions, default_left; leaf <=> left == -1,
//    leaf value = split_conditions[i]; learner_model_param.base_score / num_feature.
#include "gbdt_model.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <unordered_map>

#include "json.h"

namespace mr {

int HostTree::depth() const {
  if (feat.empty()) return 0;
  std::vector<std::pair<int, int>> st{{0, 1}};
  int best = 0;
  while (!st.empty()) {
    auto [n, d] = st.back();
    st.pop_back();
    best = std::max(best, d);
    if (left[n] >= 0) st.push_back({left[n], d + 1});
    if (right[n] >= 0) st.push_back({right[n], d + 1});
  }
  return best;
}

namespace {

struct Line {
  const char *p;
  size_t n;
};

bool starts(const Line &l, const char *s) {
  size_t n = strlen(s);
  return l.n >= n && memcmp(l.p, s, n) == 0;
}

// whitespace-separated list -> numbers.  strtod/strtol are correctly rounded, like the
// fast_double_parser path LightGBM >= 3.2 uses when it reads its own model text.
template <class T, class F>
std::vector<T> split_nums(const std::string &s, size_t expect, const char *key, F conv) {
  std::vector<T> out;
  out.reserve(expect);
  const char *p = s.c_str();
  while (*p) {
    while (*p == ' ' || *p == '\t' || *p == '\r') p++;
    if (!*p) break;
    char *end = nullptr;
    T v = conv(p, &end);
    if (end == p) fail(MR_ERR_PARSE, "lightgbm: bad number in '%s'", key);
    out.push_back(v);
    p = end;
  }
  if (out.size() != expect)
    fail(MR_ERR_PARSE, "lightgbm: '%s' has %zu values, expected %zu", key, out.size(), expect);
  return out;
}

void finalize(HostModel &m) {
  m.n_internal = 0;
  m.max_leaves = 0;
  m.max_depth = 0;
  m.has_cat = m.has_zero_missing = false;
  for (auto &t : m.trees) {
    // structure first: everything below (depth, packing, the kernels) walks child pointers without looking back
    {
      const int ni = (int)t.feat.size(), nl = (int)t.leaf.size();
      if (t.left.size() != t.feat.size() || t.right.size() != t.feat.size() || t.flags.size() != t.feat.size() ||
          t.thr.size() != t.feat.size())
        fail(MR_ERR_PARSE, "tree arrays disagree in length");
      if (nl < 1) fail(MR_ERR_PARSE, "a tree needs at least one leaf");
      std::vector<uint8_t> seen_node(ni, 0), seen_leaf(nl, 0);
      std::vector<int> st;
      if (ni > 0) { st.push_back(0); seen_node[0] = 1; }
      while (!st.empty()) {
        const int n = st.back();
        st.pop_back();
        for (int c : {t.left[n], t.right[n]}) {
          if (c >= 0) {
            if (c >= ni) fail(MR_ERR_PARSE, "child index out of range");
            if (seen_node[c]) fail(MR_ERR_PARSE, "tree node %d is reachable twice (cycle or shared subtree)", c);
            seen_node[c] = 1;
            st.push_back(c);
          } else {
            if (~c >= nl) fail(MR_ERR_PARSE, "child index out of range");
            seen_leaf[~c] = 1;
          }
        }
      }
      for (int n = 0; n < ni; n++)  // unreachable nodes are never walked, but their pointers must still be sane
        for (int c : {t.left[n], t.right[n]})
          if (c >= 0 ? c >= ni : ~c >= nl) fail(MR_ERR_PARSE, "child index out of range");
    }
    m.n_internal += (int64_t)t.feat.size();
    m.max_leaves = std::max<int>(m.max_leaves, (int)t.leaf.size());
    m.max_depth = std::max(m.max_depth, t.depth());
    for (size_t i = 0; i < t.feat.size(); i++) {
      if (t.flags[i] & NF_CATEGORICAL) m.has_cat = true;
      else if (((t.flags[i] >> NF_MISSING_SHIFT) & 3) == 1) m.has_zero_missing = true;
      if (t.feat[i] < 0 || t.feat[i] >= m.n_features)
        fail(MR_ERR_PARSE, "split feature %d outside [0,%d)", t.feat[i], m.n_features);
      int nl = (int)t.leaf.size(), ni = (int)t.feat.size();
      for (int c : {t.left[i], t.right[i]}) {
        if (c >= 0 ? c >= ni : ~c >= nl) fail(MR_ERR_PARSE, "child index out of range");
      }
    }
    if (t.leaf.size() > 32768) fail(MR_ERR_UNSUPPORTED, "trees with more than 32768 leaves are not supported");
  }
}

}  // namespace

HostModel parse_lightgbm_text(const uint8_t *blob, size_t len) {
  HostModel m;
  m.kind = MR_BOOSTER_LIGHTGBM;
  m.blob.assign(blob, blob + len);
  // split into lines
  std::vector<Line> lines;
  {
    const char *p = (const char *)blob, *e = p + len;
    while (p < e) {
      const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
      if (!q) q = e;
      size_t n = (size_t)(q - p);
      while (n && (p[n - 1] == '\r' || p[n - 1] == ' ')) n--;
      lines.push_back({p, n});
      p = q + 1;
    }
  }
  size_t i = 0;
  int max_feature_idx = -1, num_class = 1, per_iter = 1;
  long n_tree_sizes = -1, n_feature_names = -1;
  std::string objective;
  bool average_output = false;
  auto count_tokens = [](const std::string &v) {
    long n = 0;
    bool in = false;
    for (char ch : v) {
      const bool sp = ch == ' ' || ch == '\t';
      if (!sp && !in) n++;
      in = !sp;
    }
    return n;
  };
  for (; i < lines.size() && !starts(lines[i], "Tree="); i++) {
    std::string s(lines[i].p, lines[i].n);
    if (starts(lines[i], "max_feature_idx=")) max_feature_idx = atoi(s.c_str() + 16);
    else if (starts(lines[i], "num_class=")) num_class = atoi(s.c_str() + 10);
    else if (starts(lines[i], "num_tree_per_iteration=")) per_iter = atoi(s.c_str() + 23);
    else if (starts(lines[i], "objective=")) objective = s.substr(10);
    else if (s == "average_output") average_output = true;  // GBDT::SaveModelToString writes the bare word for rf boosting
    else if (starts(lines[i], "tree_sizes=")) n_tree_sizes = count_tokens(s.substr(11));
    else if (starts(lines[i], "feature_names=")) n_feature_names = count_tokens(s.substr(14));
  }
  if (max_feature_idx < 0) fail(MR_ERR_PARSE, "lightgbm: max_feature_idx missing (not a LightGBM model text)");
  if (num_class != 1 || per_iter != 1) fail(MR_ERR_UNSUPPORTED, "lightgbm: multiclass models are not supported");
  m.n_features = max_feature_idx + 1;
  if (n_feature_names >= 0 && n_feature_names != m.n_features)
    fail(MR_ERR_PARSE, "lightgbm: feature_names lists %ld names, max_feature_idx says %d features", n_feature_names, m.n_features);
  if (average_output)
    fail(MR_ERR_UNSUPPORTED, "lightgbm: average_output (random-forest boosting) models are not supported: the prediction is the "
                             "MEAN of the trees, Metarank trains gbdt LambdaMART (S/ml/rank/LambdaMARTRanker.scala:96-118)");
  {
    // Booster.predictMat is a NORMAL prediction (LGBM_BoosterPredictForMat): the objective's ConvertOutput applies.  It is the
    // identity for the ranking and plain regression objectives; everything else (sigmoid, exp, softmax, sqrt) is refused
    // rather than silently returned as raw scores.  The line is "<name> [key:value ...]".
    const std::string name = objective.substr(0, objective.find(' '));
    static const char *identity[] = {"", "lambdarank", "rank_xendcg", "regression", "regression_l1", "huber", "fair", "quantile",
                                     "mape", "custom"};
    bool ok = false;
    for (const char *k : identity) ok |= name == k;
    if (ok && name == "regression" && objective.find("sqrt") != std::string::npos) ok = false;
    if (!ok)
      fail(MR_ERR_UNSUPPORTED, "lightgbm: objective '%s' transforms the raw score on prediction; only identity objectives "
                               "(lambdarank, rank_xendcg, regression, ...) are supported", objective.c_str());
  }

  std::unordered_map<std::string, std::string> kv;
  auto flush = [&]() {
    if (kv.empty()) return;
    auto need = [&](const char *k) -> const std::string & {
      auto it = kv.find(k);
      if (it == kv.end()) fail(MR_ERR_PARSE, "lightgbm: tree block lacks '%s'", k);
      return it->second;
    };
    int nl = atoi(need("num_leaves").c_str());
    if (nl < 1) fail(MR_ERR_PARSE, "lightgbm: num_leaves < 1");
    if (kv.count("is_linear") && atoi(kv["is_linear"].c_str()) != 0)
      fail(MR_ERR_UNSUPPORTED, "lightgbm: linear trees are not supported");
    HostTree t;
    auto d = [](const char *p, char **e) { return strtod(p, e); };
    auto l = [](const char *p, char **e) { return (int32_t)strtol(p, e, 10); };
    auto u = [](const char *p, char **e) { return (uint32_t)strtoul(p, e, 10); };
    t.leaf = split_nums<double>(need("leaf_value"), (size_t)nl, "leaf_value", d);
    size_t ni = (size_t)nl - 1;
    if (ni > 0) {
      t.feat = split_nums<int32_t>(need("split_feature"), ni, "split_feature", l);
      t.thr = split_nums<double>(need("threshold"), ni, "threshold", d);
      auto dt = split_nums<int32_t>(need("decision_type"), ni, "decision_type", l);
      t.left = split_nums<int32_t>(need("left_child"), ni, "left_child", l);
      t.right = split_nums<int32_t>(need("right_child"), ni, "right_child", l);
      int ncat = kv.count("num_cat") ? atoi(kv["num_cat"].c_str()) : 0;
      std::vector<int32_t> cb;
      std::vector<uint32_t> ct;
      if (ncat > 0) {
        cb = split_nums<int32_t>(need("cat_boundaries"), (size_t)ncat + 1, "cat_boundaries", l);
        if (cb.back() < 0) fail(MR_ERR_PARSE, "lightgbm: bad cat_boundaries");
        ct = split_nums<uint32_t>(need("cat_threshold"), (size_t)cb.back(), "cat_threshold", u);
      }
      t.flags.resize(ni);
      t.cat_begin.assign(ni, 0);
      t.cat_n.assign(ni, 0);
      for (size_t k = 0; k < ni; k++) {
        uint32_t f = (uint32_t)dt[k] & 0xF;
        if (f & NF_CATEGORICAL) {
          int ci = (int)t.thr[k];
          if (ci < 0 || ci >= ncat) fail(MR_ERR_PARSE, "lightgbm: categorical threshold index out of range");
          if (cb[ci] < 0 || cb[ci + 1] < cb[ci] || cb[ci + 1] > (int)ct.size())
            fail(MR_ERR_PARSE, "lightgbm: bad cat_boundaries");
          t.cat_begin[k] = (int32_t)t.cat_words.size();
          t.cat_n[k] = cb[ci + 1] - cb[ci];
          t.cat_words.insert(t.cat_words.end(), ct.begin() + cb[ci], ct.begin() + cb[ci + 1]);
        } else {
          // where does NaN go?  Tree::NumericalDecision: NaN -> 0.0 unless missing_type == NaN;
          // then Zero/NaN missing types take the default side, otherwise 0.0 <= threshold.
          uint32_t mt = (f >> NF_MISSING_SHIFT) & 3;
          bool nan_left = (mt == 0) ? (0.0 <= t.thr[k]) : ((f & NF_DEFAULT_LEFT) != 0);
          if (nan_left) f |= NF_NAN_LEFT;
        }
        t.flags[k] = (uint8_t)f;
      }
    }
    m.trees.push_back(std::move(t));
    kv.clear();
  };
  bool in_tree = false, saw_end = false;
  long next_tree = 0;
  for (; i < lines.size(); i++) {
    const Line &ln = lines[i];
    if (starts(ln, "Tree=")) {
      flush();
      in_tree = true;
      kv["Tree"] = std::string(ln.p + 5, ln.n - 5);
      if (atol(kv["Tree"].c_str()) != next_tree) fail(MR_ERR_PARSE, "lightgbm: tree blocks out of order: 'Tree=%s' where Tree=%ld was expected", kv["Tree"].c_str(), next_tree);
      next_tree++;
      continue;
    }
    if (starts(ln, "end of trees")) { saw_end = true; break; }
    if (!in_tree) continue;
    const char *eq = (const char *)memchr(ln.p, '=', ln.n);
    if (!eq) continue;
    kv[std::string(ln.p, (size_t)(eq - ln.p))] = std::string(eq + 1, ln.n - (size_t)(eq - ln.p) - 1);
  }
  flush();
  // SaveModelToString always closes the tree section; its absence means the blob was cut (a model file that lost
  // its tail would otherwise load as a shorter, silently different ensemble)
  if (!saw_end) fail(MR_ERR_PARSE, "lightgbm: 'end of trees' missing: the model text is truncated");
  if (n_tree_sizes >= 0 && n_tree_sizes != (long)m.trees.size())
    fail(MR_ERR_PARSE, "lightgbm: tree_sizes announces %ld trees, the text holds %zu", n_tree_sizes, m.trees.size());
  finalize(m);
  return m;
}

// XGBoost's "deprecated" binary model: what Booster.toByteArray() / save_raw() / save_model("x.model") write by default
// up to XGBoost 2.0 (JSON / UBJSON are opt-in there and the default from 2.1) — and therefore what a Metarank model trained
// with ltrlib's xgboost4j holds.  Little-endian, fixed-size C structs (xgboost src/learner.cc LearnerModelParamLegacy,
// src/gbm/gbtree_model.h GBTreeModelParam, include/xgboost/tree_model.h TreeParam / RegTree::Node / RTreeNodeStat):
//   ["binf"]                                   optional 4-byte magic (0.4 - 0.9x save_raw)
//   LearnerModelParamLegacy   136 B            f32 base_score, u32 num_feature, i32 num_class, i32 contain_extra_attrs,
//                                              i32 contain_eval_metrics, u32 major, u32 minor, u32 num_target, i32 x 26
//   u64 n + bytes             name_obj         ("rank:pairwise", "rank:ndcg", ...)
//   u64 n + bytes             name_gbm         ("gbtree")
//   GBTreeModelParam          160 B            i32 num_trees, i32 num_roots / num_parallel_tree, i32 num_feature, i32 pad,
//                                              i64 num_pbuffer, i32 num_output_group, i32 size_leaf_vector, i32 x 32
//   per tree: TreeParam       148 B            i32 num_roots, i32 num_nodes, i32 num_deleted, i32 max_depth, u32 num_feature,
//                                              i32 size_leaf_vector, i32 x 31
//             Node x num_nodes    20 B each    i32 parent, i32 cleft (-1: leaf), i32 cright, u32 sindex (bit 31 = default
//                                              left, low 31 = feature; 0xFFFFFFFF = deleted), f32 leaf value | split condition
//             Stat x num_nodes    16 B each    f32 loss_chg, f32 sum_hess, f32 base_weight, i32 leaf_child_cnt
//             [u64 n + f32 x n]                leaf vector, only when size_leaf_vector != 0 (pre-1.0 multi-output; refused)
//   i32 x num_trees           tree_info        output group of each tree
//   (attributes, metrics, and from 1.0 a JSON configuration follow; prediction does not need them)
// Unpinned like the rest of the scorer side: no file XGBoost wrote exists here; restated from the public
// headers.  A node's test is the same as in the JSON form: missing -> default side, else left iff (f32)x < split condition.
static HostModel parse_xgboost_binary(const uint8_t *blob, size_t len) {
  HostModel m;
  m.kind = MR_BOOSTER_XGBOOST;
  m.blob.assign(blob, blob + len);
  size_t p = 0;
  auto need = [&](size_t n, const char *what) {
    if (len - p < n) fail(MR_ERR_PARSE, "xgboost binary model: truncated in %s (offset %zu, %zu more bytes needed, %zu left)", what, p, n, len - p);
  };
  auto rd32 = [&](size_t off) { uint32_t v; memcpy(&v, blob + off, 4); return v; };
  auto rdf = [&](size_t off) { float v; memcpy(&v, blob + off, 4); return v; };
  auto rd64 = [&](size_t off) { uint64_t v; memcpy(&v, blob + off, 8); return v; };
  if (len >= 4 && memcmp(blob, "binf", 4) == 0) p = 4;
  need(136, "LearnerModelParam");
  m.base_score = rdf(p);
  const uint32_t num_feature = rd32(p + 4);
  const int32_t num_class = (int32_t)rd32(p + 8);
  const uint32_t major = rd32(p + 20), num_target = rd32(p + 28);
  p += 136;
  if (!(m.base_score == m.base_score) || num_feature == 0 || num_feature > (1u << 24) || major > 10)
    fail(MR_ERR_UNSUPPORTED, "xgboost: unsupported model encoding (neither JSON / UBJSON nor a plausible binary header: num_feature %u, "
                             "version %u)", num_feature, major);
  if (num_class > 1) fail(MR_ERR_UNSUPPORTED, "xgboost: multiclass models are not supported");
  if (num_target > 1) fail(MR_ERR_UNSUPPORTED, "xgboost: multi-target models are not supported");
  auto rd_str = [&](const char *what) {
    need(8, what);
    const uint64_t n = rd64(p);
    p += 8;
    if (n > 256) fail(MR_ERR_PARSE, "xgboost binary model: %s is %llu bytes long", what, (unsigned long long)n);
    need((size_t)n, what);
    std::string v((const char *)blob + p, (size_t)n);
    p += (size_t)n;
    return v;
  };
  const std::string obj = rd_str("the objective name"), gbm = rd_str("the booster name");
  {
    static const char *identity[] = {"rank:ndcg", "rank:pairwise", "rank:map", "reg:squarederror", "reg:linear",
                                     "reg:absoluteerror", "reg:pseudohubererror", "reg:quantileerror"};
    bool ok = false;
    for (const char *k : identity) ok |= obj == k;
    if (!ok) fail(MR_ERR_UNSUPPORTED, "xgboost: objective '%s' transforms the margin on prediction; only rank:* and plain "
                                     "regression objectives are supported", obj.c_str());
  }
  if (gbm != "gbtree") fail(MR_ERR_UNSUPPORTED, "xgboost: booster '%s' not supported", gbm.c_str());
  m.n_features = (int)num_feature;
  need(160, "GBTreeModelParam");
  const int32_t num_trees = (int32_t)rd32(p), out_groups = (int32_t)rd32(p + 24), gb_leaf_vec = (int32_t)rd32(p + 28);
  p += 160;
  if (num_trees < 0 || (size_t)num_trees > len / 168) fail(MR_ERR_PARSE, "xgboost binary model: %d trees cannot fit %zu bytes", num_trees, len);
  if (out_groups > 1 || gb_leaf_vec > 1) fail(MR_ERR_UNSUPPORTED, "xgboost: multi-output models are not supported");
  for (int32_t ti = 0; ti < num_trees; ti++) {
    need(148, "TreeParam");
    const int32_t n = (int32_t)rd32(p + 4), leaf_vec = (int32_t)rd32(p + 20);
    p += 148;
    if (n < 1 || (size_t)n > (len - p) / 36) fail(MR_ERR_PARSE, "xgboost binary model: tree %d announces %d nodes", ti, n);
    if (leaf_vec > 1) fail(MR_ERR_UNSUPPORTED, "xgboost: vector leaves are not supported");
    need((size_t)n * 36, "a tree's nodes");
    const size_t nodes = p;
    p += (size_t)n * 36;
    if (leaf_vec != 0) {  // pre-1.0 files carry an (empty or scalar) leaf vector here
      need(8, "a tree's leaf vector");
      const uint64_t lv = rd64(p);
      p += 8;
      need((size_t)lv * 4, "a tree's leaf vector");
      p += (size_t)lv * 4;
    }
    // renumber from the root: deleted (pruned) nodes stay in the array but nothing points at them
    std::vector<int32_t> id(n, INT32_MIN);
    std::vector<int32_t> order;
    std::vector<int32_t> st{0};
    int ni = 0, nl = 0;
    while (!st.empty()) {
      const int32_t i = st.back();
      st.pop_back();
      if (i < 0 || i >= n) fail(MR_ERR_PARSE, "xgboost: child out of range");
      if (id[i] != INT32_MIN) fail(MR_ERR_PARSE, "xgboost binary model: node %d of tree %d is reachable twice", i, ti);
      const size_t o = nodes + (size_t)i * 20;
      const int32_t cl = (int32_t)rd32(o + 4), cr = (int32_t)rd32(o + 8);
      if (rd32(o + 12) == 0xFFFFFFFFu && cl != -1) fail(MR_ERR_PARSE, "xgboost binary model: tree %d reaches a deleted node", ti);
      order.push_back(i);
      if (cl == -1) id[i] = ~(nl++);
      else { id[i] = ni++; st.push_back(cr); st.push_back(cl); }
    }
    HostTree t;
    t.feat.resize(ni); t.thr.resize(ni); t.flags.resize(ni); t.left.resize(ni); t.right.resize(ni);
    t.cat_begin.assign(ni, 0); t.cat_n.assign(ni, 0);
    t.leaf.resize(nl);
    for (int32_t i : order) {
      const size_t o = nodes + (size_t)i * 20;
      if (id[i] < 0) { t.leaf[~id[i]] = (double)rdf(o + 16); continue; }
      const uint32_t sindex = rd32(o + 12);
      const int k2 = id[i];
      t.feat[k2] = (int32_t)(sindex & 0x7FFFFFFFu);
      t.thr[k2] = (double)rdf(o + 16);
      const bool dl = (sindex >> 31) != 0;
      t.flags[k2] = (uint8_t)((dl ? NF_DEFAULT_LEFT | NF_NAN_LEFT : 0) | (2u << NF_MISSING_SHIFT));
      t.left[k2] = id[(int32_t)rd32(o + 4)];
      t.right[k2] = id[(int32_t)rd32(o + 8)];
    }
    m.trees.push_back(std::move(t));
  }
  need((size_t)num_trees * 4, "tree_info");
  for (int32_t ti = 0; ti < num_trees; ti++)
    if (rd32(p + (size_t)ti * 4) != 0) fail(MR_ERR_UNSUPPORTED, "xgboost: multi-group models are not supported");
  finalize(m);
  return m;
}

HostModel parse_xgboost_model(const uint8_t *blob, size_t len) {
  HostModel m;
  m.kind = MR_BOOSTER_XGBOOST;
  m.blob.assign(blob, blob + len);
  size_t s = 0;
  while (s < len && (blob[s] == ' ' || blob[s] == '\n' || blob[s] == '\r' || blob[s] == '\t')) s++;
  if (len - s >= 4 && memcmp(blob + s, "bs64", 4) == 0)
    fail(MR_ERR_UNSUPPORTED, "xgboost: base64-wrapped binary model ('bs64' header of old Python pickles); decode it, or re-save the "
                             "booster (booster.toByteArray() / save_raw())");
  if (s >= len || blob[s] != '{') return parse_xgboost_binary(blob, len);
  JValue doc;
  size_t k = s + 1;
  while (k < len && (blob[k] == ' ' || blob[k] == '\n' || blob[k] == '\r' || blob[k] == '\t')) k++;
  if (k < len && (blob[k] == '"' || blob[k] == '}')) doc = JsonParser(blob + s, len - s).parse();
  else doc = UbjParser(blob + s, len - s).parse();

  const JValue &learner = doc.at("learner");
  const JValue &gb = learner.at("gradient_booster");
  if (gb.at("name").str != "gbtree") fail(MR_ERR_UNSUPPORTED, "xgboost: booster '%s' not supported", gb.at("name").str.c_str());
  const JValue &lmp = learner.at("learner_model_param");
  {
    // base_score is written as a string: "5E-1" up to XGBoost 2.x, "[5E-1]" (a vector intercept, one entry per target)
    // from 3.1 on; a number in hand-written files
    const JValue &bs = lmp.at("base_score");
    if (bs.kind == JValue::Str) {
      std::string t = bs.str;
      size_t a = 0, b = t.size();
      while (a < b && (t[a] == ' ' || t[a] == '[')) a++;
      while (b > a && (t[b - 1] == ' ' || t[b - 1] == ']')) b--;
      t = t.substr(a, b - a);
      if (t.find(',') != std::string::npos) fail(MR_ERR_UNSUPPORTED, "xgboost: multi-target base_score '%s' is not supported", bs.str.c_str());
      char *end = nullptr;
      m.base_score = strtof(t.c_str(), &end);
      if (t.empty() || end == t.c_str() || *end != 0) fail(MR_ERR_PARSE, "xgboost: base_score '%s' is not a number", bs.str.c_str());
    } else {
      m.base_score = bs.as_f32();
    }
  }
  m.n_features = (int)lmp.at("num_feature").as_int();
  if (const JValue *nc = lmp.get("num_class"))
    if (nc->as_int() > 1) fail(MR_ERR_UNSUPPORTED, "xgboost: multiclass models are not supported");
  if (const JValue *nt = lmp.get("num_target"))
    if (nt->as_int() > 1) fail(MR_ERR_UNSUPPORTED, "xgboost: multi-target models are not supported");
  if (const JValue *obj = learner.get("objective"))
    if (const JValue *on = obj->get("name")) {
      // Booster.predict applies the objective's PredTransform; it is the identity for rank:* and the plain regressions
      // (and base_score needs no ProbToMargin for them).  Anything else is refused rather than returned as margins.
      static const char *identity[] = {"rank:ndcg", "rank:pairwise", "rank:map", "reg:squarederror", "reg:linear",
                                       "reg:absoluteerror", "reg:pseudohubererror", "reg:quantileerror"};
      bool ok = false;
      for (const char *k : identity) ok |= on->str == k;
      if (!ok) fail(MR_ERR_UNSUPPORTED, "xgboost: objective '%s' transforms the margin on prediction; only rank:* and plain "
                                       "regression objectives are supported", on->str.c_str());
    }
  const JValue &trees = gb.at("model").at("trees");
  if (trees.kind != JValue::Arr) fail(MR_ERR_PARSE, "xgboost: trees is not an array");
  for (const JValue &jt : trees.arr) {
    const auto &L = jt.at("left_children").arr, &R = jt.at("right_children").arr;
    const auto &SI = jt.at("split_indices").arr, &SC = jt.at("split_conditions").arr, &DL = jt.at("default_left").arr;
    size_t n = L.size();
    if (R.size() != n || SI.size() != n || SC.size() != n || DL.size() != n || n == 0)
      fail(MR_ERR_PARSE, "xgboost: tree arrays have inconsistent sizes");
    if (const JValue *st = jt.get("split_type"))
      for (auto &v : st->arr)
        if (v.as_int() != 0)
          fail(MR_ERR_UNSUPPORTED, "xgboost: categorical splits are not supported (they need enable_categorical and typed "
                                   "features; Metarank hands XGBoost a plain float matrix, S/ml/rank/LambdaMARTRanker.scala:119-140)");
    if (const JValue *tp = jt.get("tree_param"))
      if (const JValue *nn = tp->get("num_nodes"))
        if ((size_t)nn->as_int() != L.size())
          fail(MR_ERR_PARSE, "xgboost: tree_param.num_nodes = %lld but the tree's arrays hold %zu nodes", (long long)nn->as_int(), L.size());
    // renumber: internal nodes and leaves get separate index spaces
    std::vector<int32_t> id(n);
    int ni = 0, nl = 0;
    for (size_t i = 0; i < n; i++) id[i] = (L[i].as_int() == -1) ? ~(nl++) : ni++;
    HostTree t;
    t.feat.resize(ni); t.thr.resize(ni); t.flags.resize(ni); t.left.resize(ni); t.right.resize(ni);
    t.cat_begin.assign(ni, 0); t.cat_n.assign(ni, 0);
    t.leaf.resize(nl);
    for (size_t i = 0; i < n; i++) {
      if (id[i] < 0) {
        t.leaf[~id[i]] = (double)SC[i].as_f32();
      } else {
        int64_t l = L[i].as_int(), r = R[i].as_int();
        if (l < 0 || r < 0 || (size_t)l >= n || (size_t)r >= n) fail(MR_ERR_PARSE, "xgboost: child out of range");
        int k2 = id[i];
        t.feat[k2] = (int32_t)SI[i].as_int();
        t.thr[k2] = (double)SC[i].as_f32();
        bool dl = DL[i].as_int() != 0;
        t.flags[k2] = (uint8_t)((dl ? NF_DEFAULT_LEFT | NF_NAN_LEFT : 0) | (2u << NF_MISSING_SHIFT));
        t.left[k2] = id[l];
        t.right[k2] = id[r];
      }
    }
    m.trees.push_back(std::move(t));
  }
  finalize(m);
  return m;
}

void parse_metarank_frame(const uint8_t *blob, size_t len, std::vector<std::string> &names, int &kind,
                          size_t &begin, size_t &end) {
  // java.io.DataInputStream: big-endian; readUTF = u16 length + modified UTF-8
  size_t p = 0;
  auto need = [&](size_t n) {
    if (len - p < n) fail(MR_ERR_PARSE, "metarank model blob truncated");
  };
  auto i32 = [&]() {
    need(4);
    int32_t v = (int32_t)((uint32_t)blob[p] << 24 | (uint32_t)blob[p + 1] << 16 | (uint32_t)blob[p + 2] << 8 | blob[p + 3]);
    p += 4;
    return v;
  };
  need(1);
  int version = (int8_t)blob[p++];
  if (version != 2 && version != 3) fail(MR_ERR_PARSE, "unsupported bitstream version %d", version);
  int32_t nf = i32();
  if (nf < 0) fail(MR_ERR_PARSE, "negative feature count");
  names.clear();
  for (int i = 0; i < nf; i++) {
    need(2);
    size_t n = (size_t)blob[p] << 8 | blob[p + 1];
    p += 2;
    need(n);
    names.emplace_back((const char *)blob + p, n);
    p += n;
  }
  need(1);
  kind = (int8_t)blob[p++];
  int32_t size = i32();
  if (size < 0) fail(MR_ERR_PARSE, "negative booster size");
  need((size_t)size);
  begin = p;
  end = p + (size_t)size;
  if (kind != 0 && kind != 1) fail(MR_ERR_UNSUPPORTED, "unsupported booster tag %d", kind);
}

// ------------------------------------------------------------------ packing

namespace {
inline size_t al16(size_t x) { return (x + 15) & ~size_t(15); }

struct DNodeHost {
  union {
    double thr64;
    struct { float thr32; uint32_t pad; } f;
    struct { uint32_t cat_off, cat_n; } c;
  };
  uint32_t ff;
  int16_t left, right;
};
static_assert(sizeof(DNodeHost) == 16, "DNode must be 16 bytes");
}  // namespace

PackedModel pack_model(const HostModel &m, size_t chunk_budget) {
  PackedModel pk;
  const bool f32 = m.kind == MR_BOOSTER_XGBOOST;
  const size_t leaf_sz = f32 ? 4 : 8;
  // a single-leaf tree becomes one dummy node whose two children are leaf 0
  auto n_nodes = [](const HostTree &t) { return t.feat.empty() ? size_t(1) : t.feat.size(); };
  auto tree_bytes = [&](const HostTree &t) {
    return n_nodes(t) * 16 + al16(t.leaf.size() * leaf_sz) + al16(t.cat_words.size() * 4);
  };
  size_t i = 0, nt = m.trees.size();
  if (nt == 0) {
    // empty ensemble: one empty chunk so the kernel still writes base scores
    ChunkDesc cd{0, 16, 0, 0};
    pk.bytes.assign(16, 0);
    pk.chunks.push_back(cd);
    pk.max_chunk_bytes = 16;
    return pk;
  }
  while (i < nt) {
    size_t j = i, body = 0;
    while (j < nt) {
      size_t nb = body + tree_bytes(m.trees[j]);
      size_t hdr = 16 + al16((j - i + 1) * 8);
      if (j > i && hdr + nb > chunk_budget) break;
      body = nb;
      j++;
    }
    size_t n = j - i;
    size_t hdr = 16 + al16(n * 8);
    size_t total = hdr + body;
    if (total >= (1u << 20)) fail(MR_ERR_UNSUPPORTED, "a single tree needs %zu bytes of shared memory", total);
    size_t base = pk.bytes.size();
    pk.bytes.resize(base + total, 0);
    uint8_t *c = pk.bytes.data() + base;
    uint32_t hn = (uint32_t)n;
    memcpy(c, &hn, 4);
    uint32_t *tab = (uint32_t *)(c + 16);
    size_t off = hdr;
    for (size_t k = 0; k < n; k++) {
      const HostTree &t = m.trees[i + k];
      size_t nn = n_nodes(t);
      size_t node_off = off, leaf_off = off + nn * 16;
      size_t cat_off = leaf_off + al16(t.leaf.size() * leaf_sz);
      tab[2 * k] = (uint32_t)node_off;
      tab[2 * k + 1] = (uint32_t)leaf_off;
      DNodeHost *nodes = (DNodeHost *)(c + node_off);
      if (t.feat.empty()) {
        nodes[0].thr64 = 0.0;
        nodes[0].ff = 0;
        nodes[0].left = nodes[0].right = (int16_t)~0;
      }
      for (size_t q = 0; q < t.feat.size(); q++) {
        DNodeHost &d = nodes[q];
        if (t.flags[q] & NF_CATEGORICAL) {
          d.c.cat_off = (uint32_t)((cat_off / 4) + (size_t)t.cat_begin[q]);
          d.c.cat_n = (uint32_t)t.cat_n[q];
        } else if (f32) {
          d.f.thr32 = (float)t.thr[q];
          d.f.pad = 0;
        } else {
          d.thr64 = t.thr[q];
        }
        d.ff = (uint32_t)t.feat[q] | ((uint32_t)t.flags[q] << 24);
        d.left = (int16_t)t.left[q];
        d.right = (int16_t)t.right[q];
      }
      if (f32) {
        float *lv = (float *)(c + leaf_off);
        for (size_t q = 0; q < t.leaf.size(); q++) lv[q] = (float)t.leaf[q];
      } else {
        memcpy(c + leaf_off, t.leaf.data(), t.leaf.size() * 8);
      }
      if (!t.cat_words.empty()) memcpy(c + cat_off, t.cat_words.data(), t.cat_words.size() * 4);
      off = cat_off + al16(t.cat_words.size() * 4);
    }
    ChunkDesc cd{(uint32_t)base, (uint32_t)total, (uint32_t)n, (uint32_t)i};
    pk.chunks.push_back(cd);
    pk.max_chunk_bytes = std::max<uint32_t>(pk.max_chunk_bytes, (uint32_t)total);
    i = j;
  }
  return pk;
}

// ------------------------------------------------------------------ binned packing

BinnedModel pack_binned(const HostModel &m, size_t chunk_budget) {
  BinnedModel B;
  const int F = m.n_features;
  if (F > 4095 || m.has_zero_missing) return B;
  // per-feature use: numerical thresholds / categorical
  std::vector<std::vector<double>> thr(F);
  B.is_cat.assign(F, 0);
  std::vector<uint8_t> is_num(F, 0);
  bool any_cat = false, small_cat = true;  // small_cat: every bitset lives in categories 0..15 (kMetaCat16)
  for (auto &t : m.trees)
    for (size_t q = 0; q < t.feat.size(); q++) {
      const int f = t.feat[q];
      if (t.flags[q] & NF_CATEGORICAL) {
        B.is_cat[f] = 1;
        any_cat = true;
        if ((size_t)t.cat_n[q] * 32 > 65000) return B;  // category ids must fit the u16 code
        for (int32_t wd = 0; wd < t.cat_n[q]; wd++)
          if (t.cat_words[(size_t)t.cat_begin[q] + wd] & (wd == 0 ? 0xFFFF0000u : 0xFFFFFFFFu)) small_cat = false;
      } else {
        is_num[f] = 1;
        thr[f].push_back(t.thr[q]);
      }
    }
  B.cat16 = any_cat && small_cat && getenv("MR_NO_CAT16") == nullptr;
  B.thr_off.assign(F + 1, 0);
  for (int f = 0; f < F; f++) {
    if (B.is_cat[f] && is_num[f]) return B;  // a column split both ways: not representable by one code
    auto &v = thr[f];
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    for (double x : v)
      if (x != x) return B;  // NaN threshold: comparisons are never true; keep the exact kernel
    if (v.size() > 65000) return B;
    B.thr_off[f + 1] = B.thr_off[f] + (uint32_t)v.size();
    B.thr.insert(B.thr.end(), v.begin(), v.end());
  }
  // bucket index per column (see BinMeta)
  B.meta.resize(F);
  for (int f = 0; f < F; f++) {
    const double *t = B.thr.data() + B.thr_off[f];
    const uint32_t mth = B.thr_off[f + 1] - B.thr_off[f];
    BinMeta &M = B.meta[f];
    M.mn = 0.0; M.scale = 0.0; M.g = 1; M.idx_off = (uint32_t)B.bucket_range.size();
    M.thr_off = B.thr_off[f]; M.flags = (B.is_cat[f] ? (kMetaCat | (B.cat16 ? kMetaCat16 : 0u)) : 0u) | (kMetaNoDup << 16);
    if (mth >= 2 && std::isfinite(t[0]) && std::isfinite(t[mth - 1])) {
      const double span = t[mth - 1] - t[0];
      const uint32_t g = std::min<uint32_t>(4 * mth, 32768);
      const double scale = (double)g / span;
      if (std::isfinite(span) && span > 0 && std::isfinite(scale) && std::isfinite(span * scale)) {
        M.mn = t[0]; M.scale = scale; M.g = g;
      }
    }
    auto bucket = [&](double x) -> uint32_t {
      if (M.g == 1 || x <= M.mn) return 0;
      const double v = (x - M.mn) * M.scale;
      return v >= (double)(M.g - 1) ? M.g - 1 : (uint32_t)v;
    };
    // prefix counts: cnt[b] = #{j : bucket(t_j) < b}; bucket b owns thresholds [cnt[b], cnt[b+1])
    std::vector<uint32_t> cnt(M.g + 1, 0);
    for (uint32_t j = 0; j < mth; j++) cnt[bucket(t[j]) + 1]++;
    for (uint32_t b = 0; b < M.g; b++) cnt[b + 1] += cnt[b];
    for (uint32_t b = 0; b < M.g; b++) B.bucket_range.push_back(cnt[b] | (cnt[b + 1] << 16));
  }
  B.tile_cols = F;
  const bool f32 = m.kind == MR_BOOSTER_XGBOOST;
  const size_t leaf_sz = f32 ? 4 : 8;
  struct BNodeHost { uint16_t k, ff; int16_t left, right; };
  static_assert(sizeof(BNodeHost) == 8, "BNode must be 8 bytes");
  auto n_nodes = [](const HostTree &t) { return t.feat.empty() ? size_t(1) : t.feat.size(); };
  auto n_cat_nodes = [](const HostTree &t) { size_t c = 0; for (auto fl : t.flags) c += (fl & NF_CATEGORICAL) != 0; return c; };
  auto tree_bytes = [&](const HostTree &t) {
    return al16(n_nodes(t) * 8) + al16(t.leaf.size() * leaf_sz) + al16(n_cat_nodes(t) * 8) + al16(t.cat_words.size() * 4);
  };
  PackedModel &pk = B.packed;
  size_t i = 0, nt = m.trees.size();
  if (nt == 0) {
    pk.bytes.assign(16, 0);
    pk.chunks.push_back(ChunkDesc{0, 16, 0, 0});
    pk.max_chunk_bytes = 16;
    B.ok = true;
    return B;
  }
  while (i < nt) {
    size_t j = i, body = 0;
    while (j < nt) {
      size_t nb = body + tree_bytes(m.trees[j]);
      size_t hdr = 16 + al16((j - i + 1) * 8);
      if (j > i && hdr + nb > chunk_budget) break;
      body = nb;
      j++;
    }
    const size_t n = j - i, hdr = 16 + al16(n * 8), total = hdr + body;
    if (total >= (1u << 19)) return B;  // cat-table offsets are u16 in units of 8 bytes
    const size_t base = pk.bytes.size();
    pk.bytes.resize(base + total, 0);
    uint8_t *c = pk.bytes.data() + base;
    uint32_t hn = (uint32_t)n;
    memcpy(c, &hn, 4);
    uint32_t *tab = (uint32_t *)(c + 16);
    size_t off = hdr;
    for (size_t k = 0; k < n; k++) {
      const HostTree &t = m.trees[i + k];
      const size_t nn = n_nodes(t);
      const size_t node_off = off, leaf_off = node_off + al16(nn * 8);
      const size_t ctab_off = leaf_off + al16(t.leaf.size() * leaf_sz);
      const size_t cw_off = ctab_off + al16(n_cat_nodes(t) * 8);
      tab[2 * k] = (uint32_t)node_off;
      tab[2 * k + 1] = (uint32_t)leaf_off;
      BNodeHost *nodes = (BNodeHost *)(c + node_off);
      uint32_t *ctab = (uint32_t *)(c + ctab_off);  // {word offset in chunk (u32 units), n words}
      if (t.feat.empty()) nodes[0] = BNodeHost{0, 0, (int16_t)~0, (int16_t)~0};
      size_t ci = 0;
      for (size_t q = 0; q < t.feat.size(); q++) {
        BNodeHost &d = nodes[q];
        const int f = t.feat[q];
        uint32_t fl = (t.flags[q] & NF_NAN_LEFT) ? (uint32_t)BF_NAN_LEFT : 0u;
        if (t.flags[q] & NF_CATEGORICAL) {
          fl = BF_CATEGORICAL;  // NaN always goes right at categorical nodes
          ctab[2 * ci] = (uint32_t)(cw_off / 4 + (size_t)t.cat_begin[q]);
          ctab[2 * ci + 1] = (uint32_t)t.cat_n[q];
          d.k = (uint16_t)((ctab_off + ci * 8) / 8);  // u16 index of the {off, n} pair, in 8-byte units
          ci++;
        } else {
          const double *b = B.thr.data() + B.thr_off[f], *e = B.thr.data() + B.thr_off[f + 1];
          d.k = (uint16_t)(std::lower_bound(b, e, t.thr[q]) - b);  // exact match exists by construction
        }
        d.ff = (uint16_t)((uint32_t)f | (fl << 12));
        d.left = (int16_t)t.left[q];
        d.right = (int16_t)t.right[q];
      }
      if (f32) {
        float *lv = (float *)(c + leaf_off);
        for (size_t q = 0; q < t.leaf.size(); q++) lv[q] = (float)t.leaf[q];
      } else {
        memcpy(c + leaf_off, t.leaf.data(), t.leaf.size() * 8);
      }
      if (!t.cat_words.empty()) memcpy(c + cw_off, t.cat_words.data(), t.cat_words.size() * 4);
      off = cw_off + al16(t.cat_words.size() * 4);
    }
    pk.chunks.push_back(ChunkDesc{(uint32_t)base, (uint32_t)total, (uint32_t)n, (uint32_t)i});
    pk.max_chunk_bytes = std::max<uint32_t>(pk.max_chunk_bytes, (uint32_t)total);
    i = j;
  }
  B.ok = true;
  return B;
}

// ------------------------------------------------------------------ compact packing

BinnedModel pack_compact(const HostModel &m, const BinnedModel &bn, size_t chunk_budget) {
  BinnedModel C;
  if (!bn.ok || m.n_features > 1023 || m.trees.empty()) return C;
  C.thr_off = bn.thr_off;
  C.thr = bn.thr;
  C.is_cat = bn.is_cat;
  C.cat16 = bn.cat16;
  C.meta = bn.meta;
  C.bucket_range = bn.bucket_range;
  // NaN direction per tile column (BinMeta::flags): a feature whose numerical nodes all send NaN the same
  // way needs one column; a feature with both kinds gets a second column for its NaN-left nodes
  const int F = m.n_features;
  std::vector<uint32_t> n_left(F, 0), n_right(F, 0), col_left(F, 0);
  for (auto &t : m.trees)
    for (size_t q = 0; q < t.feat.size(); q++)
      if (!(t.flags[q] & NF_CATEGORICAL)) ((t.flags[q] & NF_NAN_LEFT) ? n_left : n_right)[t.feat[q]]++;
  C.tile_cols = F;
  for (int f = 0; f < F; f++) {
    uint32_t fl = C.is_cat[f] ? (kMetaCat | (C.cat16 ? kMetaCat16 : 0u)) : 0u, dup = kMetaNoDup;
    col_left[f] = (uint32_t)f;
    if (n_left[f] && !n_right[f]) fl |= kMetaNanLow;
    else if (n_left[f] && n_right[f]) { dup = (uint32_t)C.tile_cols++; col_left[f] = dup; }
    C.meta[f].flags = fl | (dup << 16);
  }
  if (C.tile_cols > 1023) return C;  // the node's column field is 10 bits
  const bool f32 = m.kind == MR_BOOSTER_XGBOOST;
  if (chunk_budget == 0) {
    // auto: the scorer wants ~48 resident warps per SM (2 CTAs x 768 threads); give the double-buffered
    // chunks what the code tile of those warps leaves of the 227 KB (profiles/sweep_r1.md)
    const long long left = (225ll * 1024 - 1536ll * C.tile_cols * 2) / 4;
    chunk_budget = left >= 24 * 1024 ? 24 * 1024 : left >= 16 * 1024 ? 16 * 1024 : 8 * 1024;
  }
  chunk_budget = std::min<size_t>(chunk_budget, 65536 - 16);
  auto n_cat_nodes = [](const HostTree &t) { size_t c = 0; for (auto fl : t.flags) c += (fl & NF_CATEGORICAL) != 0; return c; };
  auto tree_bytes = [&](const HostTree &t) {
    // a single-leaf tree is stored as one dummy split with both children on its leaf: every walk starts at a node
    return (std::max<size_t>(t.feat.size(), 1) + t.leaf.size() + n_cat_nodes(t)) * 8 + ((t.cat_words.size() * 4 + 7) & ~size_t(7));
  };
  PackedModel &pk = C.packed;
  size_t i = 0, nt = m.trees.size();
  while (i < nt) {
    size_t j = i, body = 0;
    while (j < nt) {
      const size_t nb = body + tree_bytes(m.trees[j]);
      if (j > i && 16 + al16((j - i + 1) * 4) + nb > chunk_budget) break;
      body = nb;
      j++;
    }
    const size_t n = j - i, hdr = 16 + al16(n * 4), total = al16(hdr + body);
    if (total > 65536) return C;  // a single tree too large for 16-bit byte offsets
    const size_t base = pk.bytes.size();
    pk.bytes.resize(base + total, 0);
    uint8_t *c = pk.bytes.data() + base;
    const uint32_t hn = (uint32_t)n;
    memcpy(c, &hn, 4);
    uint32_t *roots = (uint32_t *)(c + 16);
    size_t off = hdr;
    for (size_t k = 0; k < n; k++) {
      const HostTree &t = m.trees[i + k];
      const size_t ni = t.feat.size(), node_off = off, leaf_off = off + std::max<size_t>(ni, 1) * 8;
      const size_t ctab_off = leaf_off + t.leaf.size() * 8, cw_off = ctab_off + n_cat_nodes(t) * 8;
      // a pointer to a node says what it points at: bit 0 = leaf value, bit 1 = categorical node — the numeric
      // level loop then leaves on (n & 3) != 0 with the test it needs anyway, at no cost per level
      auto child = [&](int cidx) -> uint32_t {
        if (cidx < 0) return (uint32_t)((leaf_off + (size_t)(~cidx) * 8) | 1u);
        return (uint32_t)(node_off + (size_t)cidx * 8) | ((t.flags[cidx] & NF_CATEGORICAL) ? 2u : 0u);
      };
      roots[k] = (uint32_t)node_off | ((ni && (t.flags[0] & NF_CATEGORICAL)) ? 2u : 0u);
      uint32_t *w = (uint32_t *)(c + node_off);
      if (ni == 0) { w[0] = 0xFFFFu << 16; w[1] = (uint32_t)(leaf_off | 1u) * 0x10001u; }
      uint32_t *ctab = (uint32_t *)(c + ctab_off);
      size_t ci = 0;
      for (size_t q = 0; q < ni; q++) {
        const int f = t.feat[q];
        if (t.flags[q] & NF_CATEGORICAL) {
          // bit 1 = categorical; the k field is the 8-byte index of the node's {bitset word offset, n words}
          ctab[2 * ci] = (uint32_t)(cw_off / 4 + (size_t)t.cat_begin[q]);
          ctab[2 * ci + 1] = (uint32_t)t.cat_n[q];
          w[2 * q] = ((uint32_t)f * 64u | 2u) | ((uint32_t)((ctab_off + ci * 8) / 8) << 16);
          ci++;
        } else {
          const double *b = C.thr.data() + C.thr_off[f], *e = C.thr.data() + C.thr_off[f + 1];
          const uint32_t kk = (uint32_t)(std::lower_bound(b, e, t.thr[q]) - b);
          const uint32_t col = (t.flags[q] & NF_NAN_LEFT) ? col_left[f] : (uint32_t)f;
          w[2 * q] = (col * 64u) | (kk << 16);
        }
        w[2 * q + 1] = child(t.left[q]) | (child(t.right[q]) << 16);
      }
      uint8_t *lv = c + leaf_off;
      for (size_t q = 0; q < t.leaf.size(); q++) {
        if (f32) { const float v = (float)t.leaf[q]; memcpy(lv + q * 8, &v, 4); }
        else memcpy(lv + q * 8, &t.leaf[q], 8);
      }
      if (!t.cat_words.empty()) memcpy(c + cw_off, t.cat_words.data(), t.cat_words.size() * 4);
      off = cw_off + ((t.cat_words.size() * 4 + 7) & ~size_t(7));
    }
    pk.chunks.push_back(ChunkDesc{(uint32_t)base, (uint32_t)total, (uint32_t)n, (uint32_t)i});
    pk.max_chunk_bytes = std::max<uint32_t>(pk.max_chunk_bytes, (uint32_t)total);
    i = j;
  }
  C.ok = true;
  return C;
}

SlimModel pack_slim(const HostModel &m, const BinnedModel &C, size_t chunk_budget, int max_T, int warps_per_sm, uint32_t min_tile_addr) {
  SlimModel S;
  if (!C.ok || m.trees.empty()) return S;
  const int F = m.n_features;
  for (int f = 0; f < F; f++)
    if (C.thr_off[f + 1] - C.thr_off[f] > 0x7C00u) return S;  // codes must stay below the binary16 NaN patterns
  S.n_pairs = (C.tile_cols + 1) / 2;
  S.cat16 = C.cat16;
  const bool cat16 = C.cat16;
  // root table for the kernel's parameter space (SlimModel::root_tab).  With it, level 0 of a walk ends on the root's
  // child entry and the level loop is entered without a leaf test: a leaf hanging directly off the root is reached
  // through a dummy split (both children = the leaf), like the root of a single-leaf tree.
  const bool want_tab = m.trees.size() <= (size_t)kSlimRootTabMax && (!m.has_cat || cat16) && getenv("MR_NO_ROOT_TAB") == nullptr;
  auto n_dummy = [&](const HostTree &t) -> size_t {
    if (!want_tab) return 0;
    return t.feat.empty() ? 2 : (size_t)(t.left[0] < 0) + (size_t)(t.right[0] < 0);
  };
  size_t max_block = 16;
  auto n_cat_nodes = [](const HostTree &t) { size_t c = 0; for (auto fl : t.flags) c += (fl & NF_CATEGORICAL) != 0; return c; };
  auto block_bytes = [&](const HostTree &t) {
    // entries (root pair + one pair per node) + leaf slots + per categorical node {bitset byte offset, n words} + the bitsets
    // (cat16: the bitset is inside the entry)
    const size_t need = (2 + 2 * (std::max<size_t>(t.feat.size(), 1) + n_dummy(t))) * 4 + t.leaf.size() * 8 +
                        (cat16 ? 0 : n_cat_nodes(t) * 8 + ((t.cat_words.size() * 4 + 7) & ~size_t(7)));
    size_t b = 16;
    while (b < need) b <<= 1;
    return b;
  };
  for (auto &t : m.trees) max_block = std::max(max_block, block_bytes(t));
  for (int T : {512, 256, 128}) {
    if (T > max_T) continue;
    const int shift = T == 512 ? 11 : T == 256 ? 10 : 9;
    const int cbase = std::max(1, (int)((min_tile_addr + (uint32_t)T * 4u - 1u) / ((uint32_t)T * 4u)));
    if (S.n_pairs + cbase <= (1 << (16 - shift)) && max_block <= (size_t)T * 4) { S.tile_T = T; S.col_base = cbase; break; }
  }
  if (!S.tile_T) return S;
  const int shift = S.tile_T == 512 ? 11 : S.tile_T == 256 ? 10 : 9;
  const bool f32 = m.kind == MR_BOOSTER_XGBOOST;
  if (chunk_budget == 0) {
    // ~48 resident warps per SM (or what the caller plans for): what their code tiles leave of the 227 KB, split over
    // the CTAs' two chunk buffers
    const int ctas = std::max(1, warps_per_sm * 32 / S.tile_T);
    const long long tile = (long long)(S.n_pairs + S.col_base) * S.tile_T * 4;
    const long long left = (225ll * 1024 / ctas - tile - 2048) / 2;
    chunk_budget = (size_t)std::max<long long>(2048, std::min<long long>(24 * 1024, (left / 2048) * 2048));
    // a chunk is a TMA copy, an mbarrier wait and a CTA-wide barrier: at least ~8 trees between two of them
    chunk_budget = std::max(chunk_budget, std::min<size_t>(8 * max_block + 2047, 16 * 1024) & ~size_t(2047));
  }
  chunk_budget = std::max<size_t>(chunk_budget, 2048);
  PackedModel &pk = S.packed;
  size_t i = 0, nt = m.trees.size();
  if (want_tab) S.root_tab.reserve(nt * 4);
  while (i < nt) {
    // greedy: header, then blocks, each aligned to its own size
    size_t j = i, end = 0;
    auto layout_end = [&](size_t from, size_t to) {
      size_t off = 16 + al16((to - from) * 8);
      for (size_t k = from; k < to; k++) {
        const size_t b = block_bytes(m.trees[k]);
        off = (off + b - 1) & ~(b - 1);
        off += b;
      }
      return off;
    };
    while (j < nt) {
      const size_t e = layout_end(i, j + 1);
      if (j > i && e > chunk_budget) break;
      end = e;
      j++;
    }
    const size_t n = j - i, total = al16(end);
    if (total > 65536) return S;
    const size_t base = pk.bytes.size();
    pk.bytes.resize(base + total, 0);
    uint8_t *c = pk.bytes.data() + base;
    const uint32_t hn[2] = {(uint32_t)n, (uint32_t)i};  // trees in the chunk, index of its first tree
    memcpy(c, hn, 8);
    uint32_t *roots = (uint32_t *)(c + 16);  // per tree: {block offset, copy of the root entry}
    size_t off = 16 + al16(n * 8);
    for (size_t k = 0; k < n; k++) {
      const HostTree &t = m.trees[i + k];
      const size_t b = block_bytes(t), ni = t.feat.size();
      off = (off + b - 1) & ~(b - 1);
      roots[2 * k] = (uint32_t)off;
      uint32_t *e = (uint32_t *)(c + off);
      const size_t leaf_base = (2 + 2 * (std::max<size_t>(ni, 1) + n_dummy(t))) * 4;  // a multiple of 8
      // breadth-first numbering of the child pairs: node q's children live in pair slot pair_of[q]
      std::vector<uint32_t> entry_of_node(ni, 0);  // entry index of internal node q
      uint32_t next_pair = 1;
      std::vector<int> order;
      if (ni) { order.push_back(0); entry_of_node[0] = 0; }
      auto leaf_entry = [&](int cidx) { return 0x80000000u | (uint32_t)(leaf_base + (size_t)(~cidx) * 8); };
      const uint32_t cbase = (uint32_t)S.col_base;
      const uint32_t dummy = cbase << shift;  // k = 0 on tile column 0; | child pair * 8
      if (!ni) {
        // a single-leaf tree is a dummy split whose children both are its leaf: every walk starts at an internal entry
        e[0] = dummy | (1u * 8u);
        if (want_tab) {
          e[2] = dummy | (2u * 8u); e[3] = dummy | (3u * 8u);
          e[4] = e[5] = e[6] = e[7] = leaf_entry(~0);
        } else {
          e[2] = e[3] = leaf_entry(~0);
        }
      }
      const size_t ctab_base = leaf_base + t.leaf.size() * 8, cw_base = ctab_base + n_cat_nodes(t) * 8;
      uint32_t *ctab = (uint32_t *)(c + off + ctab_base);
      size_t ci = 0;
      for (size_t h = 0; h < order.size(); h++) {
        const int q = order[h];
        const uint32_t pair = next_pair++;
        const int f = t.feat[q];
        bool swapped = false;
        if ((t.flags[q] & NF_CATEGORICAL) && cat16) {
          // bit 0 = categorical, high half = the bitset itself (categories 0..15): resolved inside the level loop.  Bit 31 is
          // the leaf flag, so a set holding category 15 is stored as its complement with the children exchanged — and with
          // bit 2 set, the bit the missing code (kCat16Missing) tests, so that a missing value still reaches the right child
          uint32_t bits = t.cat_n[q] > 0 ? (t.cat_words[(size_t)t.cat_begin[q]] & 0xFFFFu) : 0u;
          swapped = (bits & 0x8000u) != 0;
          if (swapped) bits = ~bits & 0xFFFFu;
          e[entry_of_node[q]] = (bits << 16) | ((((uint32_t)f >> 1) + cbase) << shift) | (((uint32_t)f & 1u) << 1) | (pair * 8u) |
                                (swapped ? 4u : 0u) | 1u;
        } else if (t.flags[q] & NF_CATEGORICAL) {
          // bit 0 = categorical: the level loop leaves on it (the same test that finds a leaf); the k field is the 8-byte
          // index, inside the block, of the node's {bitset byte offset, n words}
          ctab[2 * ci] = (uint32_t)(cw_base + (size_t)t.cat_begin[q] * 4);
          ctab[2 * ci + 1] = (uint32_t)t.cat_n[q];
          e[entry_of_node[q]] = (uint32_t)(((ctab_base + ci * 8) / 8) << 16) | ((((uint32_t)f >> 1) + cbase) << shift) |
                                (((uint32_t)f & 1u) << 1) | (pair * 8u) | 1u;
          ci++;
        } else {
          const double *tb = C.thr.data() + C.thr_off[f], *te = C.thr.data() + C.thr_off[f + 1];
          const uint32_t kk = (uint32_t)(std::lower_bound(tb, te, t.thr[q]) - tb);
          const uint32_t dup = C.meta[f].flags >> 16;
          const uint32_t col = ((t.flags[q] & NF_NAN_LEFT) && dup != kMetaNoDup) ? dup : (uint32_t)f;
          e[entry_of_node[q]] = (kk << 16) | (((col >> 1) + cbase) << shift) | ((col & 1u) << 1) | (pair * 8u);
        }
        const int ch[2] = {swapped ? t.right[q] : t.left[q], swapped ? t.left[q] : t.right[q]};
        for (int sd = 0; sd < 2; sd++) {
          if (ch[sd] < 0 && h == 0 && want_tab) {
            const uint32_t dp = next_pair++;
            e[2 * pair + sd] = dummy | (dp * 8u);
            e[2 * dp] = e[2 * dp + 1] = leaf_entry(ch[sd]);
          } else if (ch[sd] < 0) {
            e[2 * pair + sd] = leaf_entry(ch[sd]);
          } else { entry_of_node[ch[sd]] = 2 * pair + sd; order.push_back(ch[sd]); }
        }
      }
      roots[2 * k + 1] = e[0];
      if (want_tab) { S.root_tab.push_back((uint32_t)off); S.root_tab.push_back(e[0]); S.root_tab.push_back(e[2]); S.root_tab.push_back(e[3]); }
      if (!cat16 && !t.cat_words.empty()) memcpy(c + off + cw_base, t.cat_words.data(), t.cat_words.size() * 4);
      uint8_t *lv = c + off + leaf_base;
      for (size_t q = 0; q < t.leaf.size(); q++) {
        if (f32) { const float v = (float)t.leaf[q]; memcpy(lv + q * 8, &v, 4); }
        else memcpy(lv + q * 8, &t.leaf[q], 8);
      }
      off += b;
    }
    pk.chunks.push_back(ChunkDesc{(uint32_t)base, (uint32_t)total, (uint32_t)n, (uint32_t)i});
    pk.max_chunk_bytes = std::max<uint32_t>(pk.max_chunk_bytes, (uint32_t)total);
    i = j;
  }
  S.ok = true;
  return S;
}

}  // namespace mr

namespace mr {

// Host-side consistency check of the slim packing (mr_model_selfcheck; CPU-only CI): random code vectors are walked
// through (a) the parsed trees with the binned decision rule — `code <= k` on the node's tile column, bitset test on the
// category — and (b) the packed 4-byte entries exactly as gbdt_score_slim_kernel reads them (root table or the chunk's own
// table, masks of the tile size, dummy splits, in-entry bitsets).  Returns the number of (sample, tree) pairs whose leaf
// value differs.  It validates LAYOUT, not arithmetic: nothing is scored here.
size_t slim_pack_selfcheck(const HostModel &m, const BinnedModel &C, const SlimModel &S, int n_samples, uint64_t seed) {
  if (!S.ok) return 0;
  const int shift = S.tile_T == 512 ? 11 : S.tile_T == 256 ? 10 : 9;
  const uint32_t col_mask = ((0xFFFFu << shift) & 0xFFFFu) | 2u, child_mask = ((1u << shift) - 1u) & ~7u;
  auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(seed >> 33); };
  auto f16_le = [](uint16_t a, uint16_t b) {  // binary16 `a <= b`, false if either is NaN (the kernel's HSETP2)
    auto nan = [](uint16_t h) { return (h & 0x7C00u) == 0x7C00u && (h & 0x3FFu); };
    if (nan(a) || nan(b)) return false;
    auto key = [](uint16_t h) { return (h & 0x8000u) ? -(int)(h & 0x7FFFu) : (int)(h & 0x7FFFu); };
    return key(a) <= key(b);
  };
  size_t bad = 0;
  std::vector<uint16_t> tile((size_t)S.n_pairs * 2, 0);
  std::vector<int> cat_val(m.n_features, -1);
  std::vector<uint32_t> num_code(m.n_features, 0);
  std::vector<uint8_t> is_nan(m.n_features, 0);
  for (int s = 0; s < n_samples; s++) {
    for (int f = 0; f < m.n_features; f++) {
      const BinMeta &M = C.meta[f];
      const uint32_t nthr = C.thr_off[f + 1] - C.thr_off[f];
      is_nan[f] = rnd() % 8 == 0;
      uint16_t code;
      if (M.flags & kMetaCat) {
        cat_val[f] = is_nan[f] ? -1 : (int)(rnd() % 24) - 2;
        if (M.flags & kMetaCat16) code = (cat_val[f] >= 0 && cat_val[f] < 16) ? (uint16_t)(kCat16Base | cat_val[f]) : kCat16Missing;
        else code = cat_val[f] >= 0 ? (uint16_t)cat_val[f] : kBinNaN;
        tile[f] = code;
      } else {
        num_code[f] = rnd() % (nthr + 1);
        code = is_nan[f] ? kBinNaN : (uint16_t)num_code[f];
        tile[f] = (code == kBinNaN && (M.flags & kMetaNanLow)) ? 0 : code;
        const uint32_t dup = M.flags >> 16;
        if (dup != kMetaNoDup) tile[dup] = code == kBinNaN ? 0 : code;
      }
    }
    for (size_t ci = 0; ci < S.packed.chunks.size(); ci++) {
      const ChunkDesc &cd = S.packed.chunks[ci];
      const uint8_t *cb = S.packed.bytes.data() + cd.byte_off;
      const uint32_t *roots = (const uint32_t *)(cb + 16);
      for (uint32_t k = 0; k < cd.n_trees; k++) {
        const HostTree &t = m.trees[cd.first_tree + k];
        // (a) the parsed tree
        double want;
        if (t.feat.empty()) want = t.leaf[0];
        else {
          int n = 0;
          while (n >= 0) {
            const int f = t.feat[n];
            bool left;
            if (t.flags[n] & NF_CATEGORICAL) {
              const int cv = cat_val[f];
              left = cv >= 0 && (cv >> 5) < t.cat_n[n] && ((t.cat_words[(size_t)t.cat_begin[n] + (cv >> 5)] >> (cv & 31)) & 1u);
            } else if (is_nan[f]) {
              left = (t.flags[n] & NF_NAN_LEFT) != 0;
            } else {
              const double *tb = C.thr.data() + C.thr_off[f], *te = C.thr.data() + C.thr_off[f + 1];
              left = num_code[f] <= (uint32_t)(std::lower_bound(tb, te, t.thr[n]) - tb);
            }
            n = left ? t.left[n] : t.right[n];
          }
          want = t.leaf[~n];
        }
        // (b) the packed entries, the way the kernel walks them
        auto code_at = [&](uint32_t w) { const uint32_t pair = ((w & col_mask) >> shift) - (uint32_t)S.col_base; return tile[2 * pair + ((w >> 1) & 1u)]; };
        auto goes_left = [&](uint32_t w, uint16_t code, const uint8_t *blk) {
          if (S.cat16) {
            const bool pc = ((w >> (code & 31u)) & w & 1u) != 0;
            return pc || f16_le(code, (uint16_t)(w >> 16));
          }
          if (w & 1u) {  // wide categorical entry: {bitset byte offset, n words} at 8-byte index (w >> 16) & 0x7FFF
            if (code == kBinNaN) return false;
            const uint32_t *ct = (const uint32_t *)(blk + ((w >> 16) & 0x7FFFu) * 8u);
            const uint32_t wd = code >> 5;
            return wd < ct[1] && ((*(const uint32_t *)(blk + ct[0] + wd * 4u) >> (code & 31u)) & 1u) != 0;
          }
          return f16_le(code, (uint16_t)(w >> 16));
        };
        uint32_t block_off, w;
        int guard = 0;
        if (!S.root_tab.empty()) {
          const uint32_t *r = S.root_tab.data() + 4 * (size_t)(cd.first_tree + k);
          block_off = r[0];
          w = goes_left(r[1], code_at(r[1]), cb + block_off) ? r[2] : r[3];
          if ((int32_t)w < 0) { bad++; continue; }  // the root's children must be internal entries
        } else {
          block_off = roots[2 * k];
          w = roots[2 * k + 1];
        }
        if (block_off != roots[2 * k] || ((const uint32_t *)(cb + 4))[0] != cd.first_tree) { bad++; continue; }
        const uint8_t *blk = cb + block_off;
        while ((int32_t)w >= 0 && guard++ < 100000) {
          const uint32_t n = (w & child_mask) + (goes_left(w, code_at(w), blk) ? 0u : 4u);
          w = *(const uint32_t *)(blk + n);
        }
        double got;
        if (m.kind == MR_BOOSTER_XGBOOST) { float v; memcpy(&v, blk + (w & 0xFFFFu), 4); got = (double)v; want = (double)(float)want; }
        else memcpy(&got, blk + (w & 0xFFFFu), 8);
        if ((int32_t)w >= 0 || memcmp(&got, &want, 8) != 0) bad++;
      }
    }
  }
  return bad;
}

}  // namespace mr
