/* ORACLE — test infrastructure only.  Never linked into, imported by or called
 * from the product path (metarank_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline legs may use it.
 *
 * CPU restatement of the booster prediction Metarank reaches through
 *   booster.predictMat(values, rows, cols)   (reference
 *   src/main/scala/ai/metarank/ml/rank/LambdaMARTRanker.scala:348)
 * The arithmetic is NOT in /root/reference: it is io.github.metarank:ltrlib:0.2.6
 * -> lightgbm4j:4.6.0-1 (LightGBM 4.6.0, LGBM_BoosterPredictForMat) and metarank's
 * xgboost4j fork (XGBoosterPredict) (build.sbt:57-58), absent from this container.
 * The functions below restate the published algorithms:
 *   LightGBM  include/LightGBM/tree.h  Tree::NumericalDecision / CategoricalDecision /
 *             GetLeaf, src/boosting/gbdt_prediction.cpp GBDT::PredictRaw
 *   XGBoost   src/predictor/cpu_predictor.cc (GetLeafIndex / PredictByAllTrees),
 *             include/xgboost/tree_model.h RegTree::Node
 * PARITY UNPINNED: the reference's own tests assert no score at this boundary
 * (SURVEY.md §8c), so this oracle is anchored on the libraries' documented
 * semantics and on hand-computed known-answer trees (tests/test_oracle_gbdt.py).
 * The part all GBDT predictors share (x <= threshold goes left, leaves added in tree
 * order in f64) is additionally held bit for bit to an independent implementation,
 * scikit-learn's GradientBoostingRegressor, and the NaN rule of a node with missing
 * type NaN (default_left decides) to HistGradientBoostingRegressor, in the same test
 * file; that does not pin LightGBM's zero-as-missing / categorical rules or XGBoost's
 * binary32 path.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* LightGBM decision_type bits (tree.h): bit0 categorical, bit1 default-left,
 * bits 2-3 missing type (0 None, 1 Zero, 2 NaN). kZeroThreshold = 1e-35f. */
#define LGB_CAT_MASK 1
#define LGB_DEFAULT_LEFT_MASK 2
static const double kZeroThreshold = 1e-35f;

typedef struct {
  int32_t n_trees;
  int32_t n_features;
  const int32_t *node_off;    /* [n_trees+1] first internal node of tree t */
  const int32_t *leaf_off;    /* [n_trees+1] first leaf of tree t */
  const int32_t *split_feature;
  const double *threshold;
  const int32_t *decision_type;
  const int32_t *left_child;
  const int32_t *right_child;
  const double *leaf_value;
  const int32_t *cat_b_off;   /* [n_trees+1] offset into cat_boundaries per tree */
  const int32_t *cat_boundaries;
  const int32_t *cat_t_off;   /* [n_trees+1] offset into cat_threshold per tree */
  const uint32_t *cat_threshold;
} lgb_model_t;

static inline int lgb_find_in_bitset(const uint32_t *bits, int n, int pos) {
  int i1 = pos / 32;
  if (i1 >= n) return 0;
  return (bits[i1] >> (pos % 32)) & 1;
}

static inline double lgb_tree_predict(const lgb_model_t *m, int t, const double *row, int32_t *depth_acc) {
  int nl = m->leaf_off[t + 1] - m->leaf_off[t];
  const double *lv = m->leaf_value + m->leaf_off[t];
  if (nl <= 1) return lv[0];
  int o = m->node_off[t];
  const int32_t *sf = m->split_feature + o, *dt = m->decision_type + o;
  const int32_t *lc = m->left_child + o, *rc = m->right_child + o;
  const double *thr = m->threshold + o;
  int node = 0;
  while (node >= 0) {
    double fval = row[sf[node]];
    int d = dt[node];
    if (depth_acc) (*depth_acc)++;
    if (d & LGB_CAT_MASK) {
      /* Tree::CategoricalDecision */
      int go_left = 0;
      if (!isnan(fval)) {
        /* static_cast<int>(fval): out-of-range is cvttsd2si's INT_MIN on x86 -> "< 0" */
        int int_fval = (fval >= 2147483648.0 || fval <= -2147483649.0) ? INT32_MIN : (int)fval;
        if (int_fval >= 0) {
          int cat_idx = (int)thr[node];
          const int32_t *cb = m->cat_boundaries + m->cat_b_off[t];
          const uint32_t *ct = m->cat_threshold + m->cat_t_off[t];
          go_left = lgb_find_in_bitset(ct + cb[cat_idx], cb[cat_idx + 1] - cb[cat_idx], int_fval);
        }
      }
      node = go_left ? lc[node] : rc[node];
    } else {
      /* Tree::NumericalDecision */
      int missing_type = (d >> 2) & 3;
      if (isnan(fval) && missing_type != 2) fval = 0.0;
      if ((missing_type == 1 && fval >= -kZeroThreshold && fval <= kZeroThreshold) ||
          (missing_type == 2 && isnan(fval))) {
        node = (d & LGB_DEFAULT_LEFT_MASK) ? lc[node] : rc[node];
      } else {
        node = (fval <= thr[node]) ? lc[node] : rc[node];
      }
    }
  }
  return lv[~node];
}

/* GBDT::PredictRaw: f64 accumulation in tree order; lambdarank => no transform.
 * visited (optional) returns the total number of internal nodes evaluated. */
void oracle_lgb_predict(const lgb_model_t *m, const double *values, int32_t rows, int32_t cols,
                        double *out, int32_t threads, int64_t *visited) {
  int64_t total = 0;
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : total)
#endif
  for (int32_t r = 0; r < rows; r++) {
    const double *row = values + (int64_t)r * cols;
    double s = 0.0;
    int32_t d = 0;
    for (int t = 0; t < m->n_trees; t++) s += lgb_tree_predict(m, t, row, visited ? &d : 0);
    out[r] = s;
    total += d;
  }
  if (visited) *visited = total;
}

typedef struct {
  int32_t n_trees;
  int32_t n_features;
  float base_score;
  const int32_t *node_off; /* [n_trees+1] */
  const int32_t *left;
  const int32_t *right;
  const int32_t *split_index;
  const float *split_cond; /* leaf value when left == -1 */
  const uint8_t *default_left;
} xgb_model_t;

/* cpu_predictor.cc: features are float (DMatrix built from the double matrix with
 * missing = NaN), `fvalue < split_cond` goes left, NaN takes the default child, the
 * prediction buffer starts at base_score and each tree's leaf is added in tree order,
 * all in binary32; rank:* objectives apply no transform. */
void oracle_xgb_predict(const xgb_model_t *m, const double *values, int32_t rows, int32_t cols,
                        double *out, int32_t threads, int64_t *visited) {
  int64_t total = 0;
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : total)
#endif
  for (int32_t r = 0; r < rows; r++) {
    const double *row = values + (int64_t)r * cols;
    volatile float s = m->base_score; /* volatile: forbid wider intermediate precision */
    int64_t d = 0;
    for (int t = 0; t < m->n_trees; t++) {
      int o = m->node_off[t];
      int nid = 0;
      while (m->left[o + nid] != -1) {
        float fv = (float)row[m->split_index[o + nid]];
        if (isnan(fv))
          nid = m->default_left[o + nid] ? m->left[o + nid] : m->right[o + nid];
        else
          nid = (fv < m->split_cond[o + nid]) ? m->left[o + nid] : m->right[o + nid];
        d++;
      }
      s = s + m->split_cond[o + nid];
    }
    out[r] = (double)s;
    total += d;
  }
  if (visited) *visited = total;
}

/* Ranker.rerank's final ordering (reference S/ml/Ranker.scala:52-67):
 * items.sortBy(-_.score) — a stable sort under java.lang.Double.compare on the
 * negated score (NaN sorts last, -0.0 < 0.0).  Writes the permutation. */
static inline int64_t total_order_key(double x) {
  /* java.lang.Double.compare == compare of doubleToLongBits with the sign fix-up,
   * all NaNs canonicalised (doubleToLongBits collapses NaN payloads) */
  if (isnan(x)) return INT64_MAX;
  int64_t b;
  memcpy(&b, &x, 8);
  return b < 0 ? (int64_t)(b ^ INT64_MAX) : b;
}

void oracle_rank_order(const double *scores, int32_t n, int32_t *order, int32_t *scratch) {
  /* bottom-up stable merge sort on key(-score); ties keep request order */
  for (int i = 0; i < n; i++) order[i] = i;
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int i = lo, j = mid, k = lo;
      while (i < mid && j < hi) {
        int64_t a = total_order_key(-scores[order[i]]), b = total_order_key(-scores[order[j]]);
        if (b < a) scratch[k++] = order[j++]; else scratch[k++] = order[i++];
      }
      while (i < mid) scratch[k++] = order[i++];
      while (j < hi) scratch[k++] = order[j++];
    }
    memcpy(order, scratch, (unsigned long)n * 4);
  }
}

/* ---- CPU-arm helpers for bench.py (not restatements of reference code: the reference's feature assembly of
 * stored scalars is a hash-map read per (item, feature), FeatureValueLoader.scala:11-25; a dense row gather is the
 * CHEAPEST possible stand-in, so the CPU baseline is an upper bound on what the JVM path could do). */

/* out[r, :] = cat[idx[r], :], rows in parallel (so that the baseline's assembly is not a single-threaded
 * numpy fancy-index beside a multi-threaded scorer). */
void oracle_gather_rows(const double *cat, const int64_t *idx, int32_t rows, int32_t cols, double *out, int32_t threads) {
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int32_t r = 0; r < rows; r++) memcpy(out + (int64_t)r * cols, cat + idx[r] * (int64_t)cols, (size_t)cols * sizeof(double));
}

/* oracle_rank_order for n_requests back-to-back requests, requests in parallel. */
void oracle_rank_order_batch(const double *scores, const int32_t *offsets, int32_t n_requests, int32_t *order,
                             int32_t *scratch, int32_t threads) {
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
#endif
  for (int32_t r = 0; r < n_requests; r++)
    oracle_rank_order(scores + offsets[r], offsets[r + 1] - offsets[r], order + offsets[r], scratch + offsets[r]);
}
