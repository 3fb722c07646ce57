// C ABI for feature assembly + the full rerank path: mr_schema_*, mr_state_*, mr_rank*.
// Host mirror of Ranker.rerank (reference S/ml/Ranker.scala:27-83):
//   makeQuery  = lookup/cosine/prepass/assemble kernels over the device-resident state
//   predict    = gbdt_score kernel on the assembled matrix (never leaves HBM)
//   sortBy     = order kernel
#include <chrono>
#include <condition_variable>

#include "assemble_kernels.cuh"
#include "fv_codec.h"
#include "request_codec.h"
#include "internal.h"
#include "schema.h"
#include "sinks.cuh"
#include "state.h"

using namespace mr;

struct mr_schema {
  std::shared_ptr<const RequestPlan> req_plan;  // for mr_requests_decode
  Schema s;
  mr_ctx *ctx = nullptr;
  DFeature *d_plan = nullptr;
  FastCol *d_fast_cols = nullptr;
};

struct RankScratch {  // per-lane device scratch for one in-flight mr_rank
  uint8_t *buf = nullptr;
  size_t cap = 0;
};

// Per-model code rows of the item table (see code_rows_kernel): built on first use, refreshed from the
// store's change log after a flush.
struct CodeCache {
  uint64_t model_gen = 0, epoch = 0, last_use = 0;
  uint32_t *d = nullptr;
  size_t cap_rows = 0;
  int row_words = 0;
};

struct mr_state {
  std::mutex cache_mu;
  std::vector<CodeCache> caches;
  uint64_t cache_clock = 0;
  std::atomic<int> ranks_inflight{0};   // mr_rank calls between entry and return
  std::vector<uint32_t *> graveyard;    // evicted code rows another in-flight rank may still read
  mr_ctx *ctx = nullptr;
  mr_schema *schema = nullptr;
  std::unique_ptr<StateStore> store;
  bool dirty = false;
  uint32_t hist_pool_per_hist = 256;  // average tag-multiset entries reserved per (request, histogram)
  // device-batch API scratch (single stream use)
  uint8_t *d_scratch = nullptr;
  size_t d_scratch_cap = 0;
  int32_t *d_error = nullptr;
};

namespace {

inline size_t al(size_t x) { return (x + 255) & ~size_t(255); }

struct ScratchPlan {
  size_t item_req, item_row, visitor_row, cos, qnorm, reqagg, hist_desc, hist_pool, hist_cursor, error_flag, features, codes, leafvals, local_scores, total;
  uint32_t hist_pool_cap;
};

ScratchPlan plan_scratch(const Schema &S, int n_requests, int total_items, uint32_t per_hist, bool own_features,
                         int code_cols = 0, size_t leaf_bytes = 0, size_t local_score_bytes = 0) {
  ScratchPlan p{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += al(bytes); return r; };
  int n_cos = 0;
  for (auto &d : S.plan) n_cos += d.kind == FK_COSINE;
  p.item_req = take((size_t)total_items * 4);
  p.item_row = take((size_t)total_items * 4);
  p.visitor_row = take((size_t)n_requests * 8);
  p.cos = take((size_t)2 * n_cos * total_items * 8);
  p.qnorm = take((size_t)n_requests * std::max(n_cos, 1) * 8);
  p.reqagg = take((size_t)n_requests * std::max(S.n_reqagg, 1) * 32);
  p.hist_desc = take((size_t)n_requests * std::max(S.n_hist, 1) * 8);
  uint64_t cap = (uint64_t)n_requests * S.n_hist * per_hist;
  cap = std::min<uint64_t>(std::max<uint64_t>(cap, S.n_hist ? 4096 : 0), 0x7FFFFFF0ull);
  p.hist_pool_cap = (uint32_t)cap;
  p.hist_pool = take((size_t)cap * 8);
  p.hist_cursor = take(4);
  p.error_flag = take(4);
  p.features = own_features ? take((size_t)total_items * std::max(S.dim, 1) * 8) : 0;
  p.codes = code_cols > 0 ? take(binned_scratch_bytes(total_items, code_cols)) : 0;
  p.leafvals = leaf_bytes ? take(leaf_bytes) : 0;
  p.local_scores = local_score_bytes ? take(local_score_bytes) : 0;
  p.total = o;
  return p;
}

void fill_args(RankArgs &a, mr_state *st, uint8_t *scratch, const ScratchPlan &sp) {
  const Schema &S = st->store->schema;
  a.st = st->store->view();
  a.plan = st->schema->d_plan;
  a.n_plan = (int)S.plan.size();
  a.fast_cols = st->schema->d_fast_cols;
  a.n_fast = (int)S.fast_cols.size();
  a.dim = S.dim;
  a.n_req_f64 = (int)S.in_req_f64.size();
  a.n_req_u64 = (int)S.in_req_u64.size();
  a.n_req_vec = (int)S.in_req_vec.size();
  a.vec_stride = S.vec_stride;
  a.n_item_f64 = (int)S.in_item_f64.size();
  a.n_req_tok = (int)S.in_req_tok.size();
  a.item_req = (int32_t *)(scratch + sp.item_req);
  a.item_row = (uint32_t *)(scratch + sp.item_row);
  a.visitor_row = (uint32_t *)(scratch + sp.visitor_row);
  a.cos = (double *)(scratch + sp.cos);
  a.qnorm = (double *)(scratch + sp.qnorm);
  a.reqagg = (double *)(scratch + sp.reqagg);
  a.hist_desc = (uint2 *)(scratch + sp.hist_desc);
  a.hist_pool = (uint64_t *)(scratch + sp.hist_pool);
  a.hist_pool_cap = sp.hist_pool_cap;
  a.hist_cursor = (uint32_t *)(scratch + sp.hist_cursor);
  a.error_flag = (int32_t *)(scratch + sp.error_flag);
  a.n_hist = S.n_hist;
  a.n_reqagg = S.n_reqagg;
  int n_cos = 0;
  for (auto &d : S.plan) n_cos += d.kind == FK_COSINE;
  a.n_cos = n_cos;
}

// The binned scorer consumes u16 rank codes; when it is the scorer, the assemble kernel emits the
// codes itself and the f64 matrix is only materialised if the caller asked for it (explain).
bool fused_codes(const mr_model *model) { return model && model->use_binned(); }

const uint32_t *code_rows_for(mr_state *st, const mr_model *model, int *row_words_out);

// legacy_layout: groups-of-32 codes whatever the batch size (mega-request slices are cut at 128-item boundaries of it)
void set_codes(RankArgs &a, mr_state *st, const mr_model *model, uint8_t *scratch, const ScratchPlan &sp, bool legacy_layout = false) {
  a.codes = (uint16_t *)(scratch + sp.codes);
  a.bin = BinParams{};
  a.bin.thr_off = model->d_thr_off;
  a.bin.thr = model->d_thr;
  const BinnedLaunch B = model->binned_desc();  // the active code-based scorer's tile mapping
  a.bin.is_cat = model->d_is_cat;
  a.bin.meta = B.d_meta;
  a.bin.bucket_range = model->d_bucket_range;
  a.bin.n_features = model->host.n_features;
  a.bin.tile_cols = B.tile_cols;
  a.bin.xgb = model->host.kind == MR_BOOSTER_XGBOOST;
  a.bin.tile_T = legacy_layout ? 0 : model->code_layout(a.total_items);
  a.code_rows = a.out_features ? nullptr : code_rows_for(st, model, &a.code_row_words);
}

struct RankInFlight {  // counts a rank call in st->ranks_inflight for its duration (see code_rows_for)
  mr_state *s;
  explicit RankInFlight(mr_state *s_) : s(s_) { s->ranks_inflight.fetch_add(1); }
  ~RankInFlight() { s->ranks_inflight.fetch_sub(1); }
};

// Code rows for (state, model), current as of the last flush; nullptr when the path does not apply.
// Called with the store's shared lock held (no flush can run); builds are synchronous and serialised.
const uint32_t *code_rows_for(mr_state *st, const mr_model *model, int *row_words_out) {
  static const bool disabled = [] { const char *e = getenv("MR_CODE_ROWS"); return e && e[0] == '0'; }();
  const Schema &S = st->store->schema;
  if (disabled || !model || !model->use_binned() || S.fast_cols.empty()) return nullptr;
  StateStore &store = *st->store;
  const size_t n_rows = store.tables[SC_ITEM].d_n_rows;  // rows on the device (pending upserts are not visible yet)
  const BinnedLaunch B = model->binned_desc();
  const int crw = (B.tile_cols + 1) / 2;
  std::lock_guard<std::mutex> g(st->cache_mu);
  const bool alone = st->ranks_inflight.load() <= 1;  // a rank that fetched a pointer earlier counted itself first
  if (alone && !st->graveyard.empty()) {
    MR_CUDA_CHECK(cudaDeviceSynchronize());
    for (uint32_t *p : st->graveyard) cudaFree(p);
    st->graveyard.clear();
  }
  CodeCache *cc = nullptr;
  for (auto &c : st->caches)
    if (c.model_gen == model->code_gen) cc = &c;
  if (!cc) {
    if (st->caches.size() >= 4) {  // evict the least recently used
      size_t lru = 0;
      for (size_t k = 1; k < st->caches.size(); k++)
        if (st->caches[k].last_use < st->caches[lru].last_use) lru = k;
      if (alone) {
        MR_CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(st->caches[lru].d);
      } else if (st->caches[lru].d) {
        st->graveyard.push_back(st->caches[lru].d);  // freed once no other rank is in flight
      }
      st->caches.erase(st->caches.begin() + lru);
    }
    st->caches.emplace_back();
    cc = &st->caches.back();
    cc->model_gen = model->code_gen;
  }
  cc->last_use = ++st->cache_clock;
  RankArgs a{};
  a.st = store.view();
  a.fast_cols = st->schema->d_fast_cols;
  a.n_fast = (int)S.fast_cols.size();
  a.bin.thr_off = model->d_thr_off;
  a.bin.thr = model->d_thr;
  a.bin.is_cat = model->d_is_cat;
  a.bin.meta = B.d_meta;
  a.bin.bucket_range = model->d_bucket_range;
  a.bin.n_features = model->host.n_features;
  a.bin.tile_cols = B.tile_cols;
  a.bin.xgb = model->host.kind == MR_BOOSTER_XGBOOST;
  bool full = cc->d == nullptr || cc->cap_rows < n_rows + 1 || cc->row_words != crw;
  std::vector<uint32_t> rows;
  if (!full && cc->epoch != store.item_epoch) {
    if (store.item_log.empty() || store.item_log.front().epoch > cc->epoch + 1) full = true;  // log no longer reaches back
    for (auto &ch : store.item_log) {
      if (full) break;
      if (ch.epoch <= cc->epoch) continue;
      if (ch.all) full = true;
      else rows.insert(rows.end(), ch.rows.begin(), ch.rows.end());
    }
    if (rows.size() * 4 > n_rows) full = true;
  }
  if (full) {
    if (cc->d == nullptr || cc->cap_rows < n_rows + 1 || cc->row_words != crw) {
      if (cc->d) { MR_CUDA_CHECK(cudaDeviceSynchronize()); cudaFree(cc->d); cc->d = nullptr; }
      cc->cap_rows = n_rows + 1 + n_rows / 2;
      cc->row_words = crw;
      MR_CUDA_CHECK(cudaMalloc((void **)&cc->d, cc->cap_rows * (size_t)crw * 4));
      MR_CUDA_CHECK(cudaMemset(cc->d, 0, cc->cap_rows * (size_t)crw * 4));
    }
    launch_code_rows(a, cc->d, crw, (uint32_t)n_rows, nullptr, 0, 0);
    MR_CUDA_CHECK(cudaStreamSynchronize(0));
  } else if (!rows.empty()) {
    uint32_t *d_idx = nullptr;
    MR_CUDA_CHECK(cudaMalloc((void **)&d_idx, rows.size() * 4));
    MR_CUDA_CHECK(cudaMemcpy(d_idx, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice));
    launch_code_rows(a, cc->d, crw, (uint32_t)n_rows, d_idx, (uint32_t)rows.size(), 0);
    MR_CUDA_CHECK(cudaStreamSynchronize(0));
    cudaFree(d_idx);
  }
  cc->epoch = store.item_epoch;
  *row_words_out = crw;
  return cc->d;
}

void check_scored_dim(mr_state *st, mr_model *model) {
  if (model && model->host.n_features != st->store->schema.dim)
    fail(MR_ERR_FEATURE_MISMATCH, "booster reads %d features, the schema's dataset descriptor has %d columns",
         model->host.n_features, st->store->schema.dim);
}

}  // namespace

extern "C" {

uint64_t mr_hash64(const void *bytes, size_t len) { return hash64(bytes, len); }
int32_t mr_token_count(const char *utf8, size_t len) { return token_count(utf8, len); }

mr_status mr_schema_create(mr_ctx *ctx, const char *json, size_t len, mr_schema **out) {
  return guard([&] {
    if (!json || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    auto s = std::make_unique<mr_schema>();
    s->s = parse_schema_json(json, len);
    s->req_plan = make_request_plan(s->s);
    s->ctx = ctx;
    if (ctx) {  // ctx may be NULL for host-only validation of a config
      MR_CUDA_CHECK(cudaSetDevice(ctx->device));
      const size_t bytes = std::max<size_t>(s->s.plan.size(), 1) * sizeof(DFeature);
      MR_CUDA_CHECK(cudaMalloc((void **)&s->d_plan, bytes));
      if (!s->s.plan.empty())
        MR_CUDA_CHECK(cudaMemcpy(s->d_plan, s->s.plan.data(), s->s.plan.size() * sizeof(DFeature), cudaMemcpyHostToDevice));
      if (!s->s.fast_cols.empty()) {
        MR_CUDA_CHECK(cudaMalloc((void **)&s->d_fast_cols, s->s.fast_cols.size() * sizeof(FastCol)));
        MR_CUDA_CHECK(cudaMemcpy(s->d_fast_cols, s->s.fast_cols.data(), s->s.fast_cols.size() * sizeof(FastCol), cudaMemcpyHostToDevice));
      }
    }
    *out = s.release();
  });
}

mr_status mr_schema_free(mr_schema *s) {
  if (!s) return MR_OK;
  if (s->d_plan) cudaFree(s->d_plan);
  if (s->d_fast_cols) cudaFree(s->d_fast_cols);
  delete s;
  return MR_OK;
}

int32_t mr_schema_dim(const mr_schema *s) { return s ? s->s.dim : -1; }

int32_t mr_schema_feature_offset(const mr_schema *s, const char *feature, int32_t *dim_out) {
  if (!s || !feature) return -1;
  auto it = s->s.col_of.find(feature);
  if (it == s->s.col_of.end()) return -1;
  if (dim_out) *dim_out = it->second.second;
  return it->second.first;
}

int32_t mr_schema_input_slot(const mr_schema *s, int32_t kind, const char *feature, int32_t *n_out) {
  if (!s) return -1;
  const std::vector<std::string> *v = nullptr;
  std::vector<std::string> vec_names;
  switch (kind) {
    case MR_IN_REQ_F64: v = &s->s.in_req_f64; break;
    case MR_IN_REQ_U64: v = &s->s.in_req_u64; break;
    case MR_IN_ITEM_F64: v = &s->s.in_item_f64; break;
    case MR_IN_REQ_TOKENS: v = &s->s.in_req_tok; break;
    case MR_IN_REQ_VEC:
      for (auto &x : s->s.in_req_vec) vec_names.push_back(x.feature);
      v = &vec_names;
      break;
    default: return -1;
  }
  if (n_out) *n_out = (int32_t)v->size();
  if (!feature) return -1;
  for (size_t i = 0; i < v->size(); i++)
    if ((*v)[i] == feature) return (int32_t)i;
  return -1;
}

int32_t mr_schema_vec_stride(const mr_schema *s) { return s ? s->s.vec_stride : -1; }
int32_t mr_schema_vec_offset(const mr_schema *s, int32_t slot, int32_t *dim_out) {
  if (!s || slot < 0 || slot >= (int)s->s.in_req_vec.size()) return -1;
  if (dim_out) *dim_out = s->s.in_req_vec[slot].dim;
  return s->s.in_req_vec[slot].offset;
}

mr_status mr_state_create(mr_ctx *ctx, mr_schema *schema, mr_state **out) {
  return guard([&] {
    if (!ctx || !schema || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (!schema->d_plan) fail(MR_ERR_INVALID_ARG, "schema was created without a context");
    auto st = std::make_unique<mr_state>();
    st->ctx = ctx;
    st->schema = schema;
    st->store = std::make_unique<StateStore>(schema->s);
    st->store->device = ctx->device;
    st->dirty = true;
    MR_CUDA_CHECK(cudaSetDevice(ctx->device));
    MR_CUDA_CHECK(cudaMalloc((void **)&st->d_error, 4));
    *out = st.release();
  });
}

mr_status mr_state_free(mr_state *st) {
  if (!st) return MR_OK;
  cudaSetDevice(st->ctx->device);
  cudaDeviceSynchronize();
  if (st->d_scratch) cudaFree(st->d_scratch);
  if (st->d_error) cudaFree(st->d_error);
  for (auto &c : st->caches) cudaFree(c.d);
  for (uint32_t *p : st->graveyard) cudaFree(p);
  delete st;
  return MR_OK;
}

mr_status mr_state_upsert(mr_state *st, const uint8_t *packed, size_t len, int64_t *applied, int64_t *skipped) {
  return guard([&] {
    if (!st || (!packed && len)) fail(MR_ERR_INVALID_ARG, "null argument");
    st->store->upsert(packed, len, applied, skipped);
    st->dirty = true;
  });
}

mr_status mr_state_apply_writes(mr_state *st, const uint8_t *packed, size_t len, int64_t *applied, int64_t *skipped) {
  return guard([&] {
    if (!st || (!packed && len)) fail(MR_ERR_INVALID_ARG, "null argument");
    st->store->apply_writes(packed, len, applied, skipped);
    st->dirty = true;
  });
}

struct mr_requests {
  PackedRequests p;
};

mr_status mr_requests_decode(const mr_schema *schema, const char *json, size_t len, mr_requests **out) {
  return guard([&] {
    if (!schema || !json || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    auto r = std::make_unique<mr_requests>();
    decode_requests(schema->s, *schema->req_plan, json, len, r->p);
    *out = r.release();
  });
}

const mr_rank_batch *mr_requests_batch(const mr_requests *r, int32_t *total_items) {
  if (!r) return nullptr;
  if (total_items) *total_items = r->p.total_items;
  return &r->p.batch;
}

const char *mr_requests_item_id(const mr_requests *r, int32_t index, size_t *len) {
  if (!r || index < 0 || index >= (int32_t)r->p.item_ids.size()) return nullptr;
  if (len) *len = r->p.item_ids[index].size();
  return r->p.item_ids[index].c_str();
}

int64_t mr_requests_timestamp(const mr_requests *r, int32_t request) {
  if (!r || request < 0 || request >= (int32_t)r->p.timestamps.size()) return 0;
  return r->p.timestamps[request];
}

mr_status mr_requests_free(mr_requests *r) {
  delete r;
  return MR_OK;
}

mr_status mr_state_load_feature_values(mr_state *st, const uint8_t *bytes, size_t len, int64_t *applied,
                                       int64_t *skipped, size_t *consumed) {
  return guard([&] {
    if (!st || (!bytes && len)) fail(MR_ERR_INVALID_ARG, "null argument");
    std::vector<uint8_t> recs;
    FvStats fs;
    transcode_feature_values(bytes, len, recs, fs);  // throws before anything is applied
    int64_t ok = 0, skip = 0;
    st->store->upsert(recs.data(), recs.size(), &ok, &skip);
    st->dirty = true;
    if (applied) *applied = ok;
    if (skipped) *skipped = skip + fs.unsupported;
    if (consumed) *consumed = fs.consumed;
  });
}

mr_status mr_feature_values_transcode(const uint8_t *bytes, size_t len, uint8_t *out, size_t out_cap, size_t *out_len,
                                      int64_t *n_records, int64_t *n_unsupported, size_t *consumed) {
  return guard([&] {
    if ((!bytes && len) || !out_len) fail(MR_ERR_INVALID_ARG, "null argument");
    std::vector<uint8_t> recs;
    FvStats fs;
    transcode_feature_values(bytes, len, recs, fs);
    *out_len = recs.size();
    if (n_records) *n_records = fs.records;
    if (n_unsupported) *n_unsupported = fs.unsupported;
    if (consumed) *consumed = fs.consumed;
    if (out) {
      if (out_cap < recs.size()) fail(MR_ERR_INVALID_ARG, "output buffer too small: %zu bytes needed", recs.size());
      if (!recs.empty()) memcpy(out, recs.data(), recs.size());
    }
  });
}

mr_status mr_state_flush(mr_state *st) {
  return guard([&] {
    if (!st) fail(MR_ERR_INVALID_ARG, "null argument");
    st->store->flush();
    st->dirty = false;
  });
}

mr_status mr_state_get_info(mr_state *st, mr_state_info *out) {
  return guard([&] {
    if (!st || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    for (int t = 0; t < SC_N_TABLES; t++) out->rows[t] = (int64_t)st->store->tables[t].n_rows;
    out->device_bytes = st->store->device_bytes;
    {
      std::lock_guard<std::mutex> g(st->cache_mu);
      for (auto &c : st->caches) out->device_bytes += (int64_t)(c.cap_rows * (size_t)c.row_words * 4);  // per-model code rows
    }
    out->item_row_bytes = (int64_t)st->store->tables[SC_ITEM].row_words * 8;
  });
}

}  // extern "C"

namespace {

bool host_pinned(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Which caller buffers are page-locked (cudaHostAlloc / cudaHostRegister, e.g. a registered direct
// ByteBuffer): those are DMA'd in place, everything else is staged through the lane's pinned buffer.
struct PinInfo {
  bool ids = false, scores = false, order = false, features = false;
};

// One slice of requests in flight on one lane (stream + pinned staging + device scratch).
struct RankPending {
  bool active = false;
  int i0 = 0, n = 0;  // first item / item count of the slice
  size_t in_bytes = 0, ho_scores = 0, ho_order = 0, ho_feat = 0, ho_err = 0;
  PinInfo pin;
};

void rank_enqueue(mr_state *st, mr_model *model, const mr_rank_batch *b, int r0, int r1, Lane *lane, RankPending &pd,
                  bool want_order, bool want_features, const PinInfo &pin, double *out_scores, int32_t *out_order,
                  double *out_features) {
  const Schema &S = st->store->schema;
  const int R = r1 - r0, i0 = b->item_offsets[r0], N = b->item_offsets[r1] - i0;
  pd = RankPending{};
  pd.active = true;
  pd.i0 = i0;
  pd.n = N;
  pd.pin = pin;
  // ---- pack the slice into one pinned blob -> one H2D copy
  struct Seg { const void *src; size_t bytes, off; };
  Seg segs[12];
  size_t in_bytes = 0;
  int ns = 0;
  auto seg = [&](const void *p, size_t bytes) { segs[ns] = Seg{p, p ? bytes : 0, in_bytes}; in_bytes += al(segs[ns].bytes); return ns++; };
  const int s_off = seg(b->item_offsets + r0, (size_t)(R + 1) * 4);
  const int s_ids = seg(b->item_ids + i0, (size_t)N * 8);
  const int s_usr = seg(b->user_ids ? b->user_ids + r0 : nullptr, (size_t)R * 8);
  const int s_ses = seg(b->session_ids ? b->session_ids + r0 : nullptr, (size_t)R * 8);
  const size_t nrf = S.in_req_f64.size(), nru = S.in_req_u64.size(), nrv = S.in_req_vec.size(), nif = S.in_item_f64.size();
  const int s_rf = seg(b->req_f64 ? b->req_f64 + (size_t)r0 * nrf : nullptr, (size_t)R * nrf * 8);
  const int s_ru = seg(b->req_u64 ? b->req_u64 + (size_t)r0 * nru : nullptr, (size_t)R * nru * 8);
  const int s_rv = seg(b->req_vec ? b->req_vec + (size_t)r0 * S.vec_stride : nullptr, (size_t)R * S.vec_stride * 4);
  const int s_rp = seg(b->req_vec_present ? b->req_vec_present + (size_t)r0 * nrv : nullptr, (size_t)R * nrv);
  const int s_if = seg(b->item_f64 ? b->item_f64 + (size_t)i0 * nif : nullptr, (size_t)N * nif * 8);
  // token lists of the slice: its offsets verbatim (the kernel subtracts the first), its range of hashes / weights
  const size_t ntk = S.in_req_tok.size();
  const bool has_tok = ntk > 0 && b->req_tok_offsets && b->req_tok_hashes;
  const int32_t t0 = has_tok ? b->req_tok_offsets[(size_t)r0 * ntk] : 0, t1 = has_tok ? b->req_tok_offsets[(size_t)r1 * ntk] : 0;
  const int s_to = seg(has_tok ? b->req_tok_offsets + (size_t)r0 * ntk : nullptr, ((size_t)R * ntk + 1) * 4);
  const int s_th = seg(has_tok ? b->req_tok_hashes + t0 : nullptr, (size_t)(t1 - t0) * 8);
  const int s_tw = seg(has_tok && b->req_tok_weights ? b->req_tok_weights + t0 : nullptr, (size_t)(t1 - t0) * 8);

  const bool fused = fused_codes(model);
  const size_t leaf_bytes = (fused && model->use_latency(N)) ? latency_scratch_bytes(N, (int)model->host.trees.size()) : 0;
  ScratchPlan sp = plan_scratch(S, R, N, st->hist_pool_per_hist, want_features || !fused, fused ? model->code_cols() : 0, leaf_bytes);
  // device outputs are laid out [scores | order | error flag] so that the common (unpinned) case is ONE D2H copy
  const size_t scores_off = al(in_bytes) + sp.total, order_off = scores_off + al((size_t)N * 8);
  const size_t err_off = order_off + al((size_t)N * 4);
  const size_t d_total = err_off + 256;
  const size_t out_bytes = al((size_t)N * 8) + al((size_t)N * 4) + 256 + (want_features ? al((size_t)N * S.dim * 8) : 0) + 16;
  lane->ensure(in_bytes + out_bytes, d_total);
  for (int k = 0; k < ns; k++)
    if (segs[k].bytes && !(k == s_ids && pin.ids)) memcpy(lane->h_pinned + segs[k].off, segs[k].src, segs[k].bytes);
  if (i0 != 0) {  // rebase the slice's offsets to start at 0
    int32_t *o = (int32_t *)(lane->h_pinned + segs[s_off].off);
    for (int r = 0; r <= R; r++) o[r] -= i0;
  }
  uint8_t *d_in = lane->d_buf, *scratch = lane->d_buf + al(in_bytes);
  if (pin.ids) {
    // ids are the bulk of the request: DMA them straight from the caller's page-locked buffer
    if (segs[s_ids].off) MR_CUDA_CHECK(cudaMemcpyAsync(d_in, lane->h_pinned, segs[s_ids].off, cudaMemcpyHostToDevice, lane->stream));
    MR_CUDA_CHECK(cudaMemcpyAsync(d_in + segs[s_ids].off, segs[s_ids].src, segs[s_ids].bytes, cudaMemcpyHostToDevice, lane->stream));
    const size_t after = segs[s_ids].off + al(segs[s_ids].bytes);
    if (in_bytes > after) MR_CUDA_CHECK(cudaMemcpyAsync(d_in + after, lane->h_pinned + after, in_bytes - after, cudaMemcpyHostToDevice, lane->stream));
  } else {
    MR_CUDA_CHECK(cudaMemcpyAsync(d_in, lane->h_pinned, in_bytes, cudaMemcpyHostToDevice, lane->stream));
  }
  auto dp = [&](int si) -> const void * { return segs[si].bytes ? d_in + segs[si].off : nullptr; };
  RankArgs a{};
  fill_args(a, st, scratch, sp);
  a.n_requests = R;
  a.total_items = N;
  a.item_offsets = (const int32_t *)dp(s_off);
  a.item_ids = (const uint64_t *)dp(s_ids);
  a.user_ids = (const uint64_t *)dp(s_usr);
  a.session_ids = (const uint64_t *)dp(s_ses);
  a.req_f64 = (const double *)dp(s_rf);
  a.req_u64 = (const uint64_t *)dp(s_ru);
  a.req_vec = (const float *)dp(s_rv);
  a.req_vec_present = (const uint8_t *)dp(s_rp);
  a.item_f64 = (const double *)dp(s_if);
  a.req_tok_off = has_tok ? (const int32_t *)(d_in + segs[s_to].off) : nullptr;
  a.req_tok_hash = (const uint64_t *)dp(s_th);
  a.req_tok_w = (const double *)dp(s_tw);
  a.req_tok_base = t0;
  a.out_features = (want_features || !fused) ? (double *)(scratch + sp.features) : nullptr;
  a.error_flag = (int32_t *)(lane->d_buf + err_off);  // zeroed by lookup_kernel
  if (fused) set_codes(a, st, model, scratch, sp);
  launch_assemble(a, S, lane->stream);
  double *d_scores = (double *)(lane->d_buf + scores_off);
  int32_t *d_order = (int32_t *)(lane->d_buf + order_off);
  uint8_t *h_out = lane->h_pinned + in_bytes;
  pd.in_bytes = in_bytes;
  if (model) {
    if (fused) model->score_codes(a.codes, N, d_scores, lane->stream, leaf_bytes ? scratch + sp.leafvals : nullptr);
    else model->score(a.out_features, N, S.dim, d_scores, lane->stream);
    if (want_order) {
      int max_n = 1;  // host offsets are at hand: the CTA-wide sort is only launched when a request needs it
      for (int r = r0; r < r1; r++) max_n = std::max(max_n, b->item_offsets[r + 1] - b->item_offsets[r]);
      launch_rank_order(d_scores, a.item_offsets, R, N, d_order, lane->stream, max_n);
    }
  }
  // host staging mirrors the device layout: [scores | order | err | features]
  pd.ho_scores = 0;
  pd.ho_order = al((size_t)N * 8);
  pd.ho_err = pd.ho_order + al((size_t)N * 4);
  pd.ho_feat = pd.ho_err + 256;
  const bool any_pinned = pin.scores || pin.order;
  if (!any_pinned) {
    MR_CUDA_CHECK(cudaMemcpyAsync(h_out, d_scores, pd.ho_err + 4, cudaMemcpyDeviceToHost, lane->stream));
  } else {
    if (model)
      MR_CUDA_CHECK(cudaMemcpyAsync(pin.scores ? (void *)(out_scores + i0) : (void *)(h_out + pd.ho_scores), d_scores,
                                    (size_t)N * 8, cudaMemcpyDeviceToHost, lane->stream));
    if (model && want_order)
      MR_CUDA_CHECK(cudaMemcpyAsync(pin.order ? (void *)(out_order + i0) : (void *)(h_out + pd.ho_order), d_order,
                                    (size_t)N * 4, cudaMemcpyDeviceToHost, lane->stream));
    MR_CUDA_CHECK(cudaMemcpyAsync(h_out + pd.ho_err, a.error_flag, 4, cudaMemcpyDeviceToHost, lane->stream));
  }
  if (want_features)
    MR_CUDA_CHECK(cudaMemcpyAsync(pin.features ? (void *)(out_features + (size_t)i0 * S.dim) : (void *)(h_out + pd.ho_feat),
                                  a.out_features, (size_t)N * S.dim * 8, cudaMemcpyDeviceToHost, lane->stream));
}

// Waits for the slice and copies its outputs to the caller's buffers; returns the device error flag.
int32_t rank_finish(mr_state *st, mr_model *model, Lane *lane, RankPending &pd, double *out_scores, int32_t *out_order,
                    double *out_features) {
  if (!pd.active) return 0;
  pd.active = false;
  MR_CUDA_CHECK(cudaStreamSynchronize(lane->stream));
  const Schema &S = st->store->schema;
  const uint8_t *h_out = lane->h_pinned + pd.in_bytes;
  int32_t err;
  memcpy(&err, h_out + pd.ho_err, 4);
  if (err != 0) return err;
  if (model && out_scores && !pd.pin.scores) memcpy(out_scores + pd.i0, h_out + pd.ho_scores, (size_t)pd.n * 8);
  if (model && out_order && !pd.pin.order) memcpy(out_order + pd.i0, h_out + pd.ho_order, (size_t)pd.n * 4);
  if (out_features && !pd.pin.features) memcpy(out_features + (size_t)pd.i0 * S.dim, h_out + pd.ho_feat, (size_t)pd.n * S.dim * 8);
  return 0;
}

}  // namespace

extern "C" {

mr_status mr_rank(mr_state *st, mr_model *model, const mr_rank_batch *b, double *out_scores, int32_t *out_order,
                  double *out_features) {
  return guard([&] {
    if (!st || !b) fail(MR_ERR_INVALID_ARG, "null argument");
    if (model) check_model(model);
    if (b->n_requests < 0) fail(MR_ERR_INVALID_ARG, "negative request count");
    if (b->n_requests == 0) return;
    if (!b->item_offsets || (!b->item_ids && b->item_offsets[b->n_requests] > 0)) fail(MR_ERR_INVALID_ARG, "null item arrays");
    const int R = b->n_requests;
    if (b->item_offsets[0] != 0) fail(MR_ERR_INVALID_ARG, "item_offsets[0] must be 0");
    for (int r = 0; r < R; r++)
      if (b->item_offsets[r + 1] < b->item_offsets[r]) fail(MR_ERR_INVALID_ARG, "item_offsets must be non-decreasing");
    const int N = b->item_offsets[R];
    if (model && !out_scores && N > 0) fail(MR_ERR_INVALID_ARG, "out_scores is null");
    check_scored_dim(st, model);
    const Schema &S = st->store->schema;
    if (!S.in_req_f64.empty() && !b->req_f64) fail(MR_ERR_INVALID_ARG, "schema needs req_f64 inputs");
    if (!S.in_req_u64.empty() && !b->req_u64) fail(MR_ERR_INVALID_ARG, "schema needs req_u64 inputs");
    if (!S.in_req_vec.empty() && (!b->req_vec || !b->req_vec_present)) fail(MR_ERR_INVALID_ARG, "schema needs req_vec inputs");
    if (!S.in_req_tok.empty()) {
      if (!b->req_tok_offsets || !b->req_tok_hashes) fail(MR_ERR_INVALID_ARG, "schema needs req_tok inputs");
      const size_t n_off = (size_t)R * S.in_req_tok.size();
      if (b->req_tok_offsets[0] != 0) fail(MR_ERR_INVALID_ARG, "req_tok_offsets[0] must be 0");
      for (size_t k = 0; k < n_off; k++)
        if (b->req_tok_offsets[k + 1] < b->req_tok_offsets[k]) fail(MR_ERR_INVALID_ARG, "req_tok_offsets must be non-decreasing");
      bool bm25 = false;
      for (auto &d : S.plan) bm25 |= d.kind == FK_TOKEN_MATCH && d.aux0 == 1;
      if (bm25 && !b->req_tok_weights) fail(MR_ERR_INVALID_ARG, "a bm25 field_match needs req_tok_weights (the tokens' IDF)");
    }
    if (N == 0) return;
    std::unique_ptr<InflightGuard> ig;
    if (model) ig = std::make_unique<InflightGuard>(model);
    MR_CUDA_CHECK(cudaSetDevice(st->ctx->device));
    RankInFlight inflight(st);
    std::shared_lock<std::shared_mutex> read_guard(st->store->mu);  // no flush while kernels read the tables

    // Large batches are cut at request boundaries into slices of ~256 K items that alternate
    // between up to four lanes, so the H2D copy / kernels / D2H copy of neighbouring slices overlap.
    static const int kSliceItems = [] {
      const char *e = getenv("MR_RANK_SLICE_ITEMS");  // tuning knob; default measured in profiles/
      const int v = e ? atoi(e) : 0;
      return v > 0 ? v : (1 << 18);  // 256 K: e2e 405 M items/s vs 306 M at 128 K (tools/e2e_slices.py)
    }();
    std::vector<int> cuts{0};
    for (int r = 1; r <= R; r++)
      if (r == R || b->item_offsets[r + 1] - b->item_offsets[cuts.back()] > kSliceItems) cuts.push_back(r);
    const int n_slices = (int)cuts.size() - 1;
    PinInfo pin;
    if (N >= 4096) {  // only worth four driver queries for sizeable batches
      pin.ids = host_pinned(b->item_ids);
      pin.scores = model && host_pinned(out_scores);
      pin.order = model && out_order && host_pinned(out_order);
      pin.features = out_features && host_pinned(out_features);
    }
    for (int attempt = 0;; attempt++) {
      // up to kLanes slices in flight: while the GPU works on some, the host packs / unpacks others
      constexpr int kLanes = 4;
      const int n_lanes = std::min(kLanes, n_slices);
      std::vector<std::unique_ptr<LaneGuard>> guards;
      for (int l = 0; l < n_lanes; l++) guards.push_back(std::make_unique<LaneGuard>(st->ctx));
      RankPending pend[kLanes];
      int32_t err = 0;
      for (int sidx = 0; sidx < n_slices && err == 0; sidx++) {
        const int l = sidx % n_lanes;
        Lane *ln = guards[l]->lane.get();
        err = rank_finish(st, model, ln, pend[l], out_scores, out_order, out_features);
        if (err) break;
        rank_enqueue(st, model, b, cuts[sidx], cuts[sidx + 1], ln, pend[l], out_order != nullptr, out_features != nullptr,
                     pin, out_scores, out_order, out_features);
      }
      for (int l = 0; l < n_lanes; l++) {
        const int32_t e = rank_finish(st, model, guards[l]->lane.get(), pend[l], out_scores, out_order, out_features);
        if (!err) err = e;
      }
      if (err == -1 && attempt < 6) {  // tag-multiset pool too small: grow and redo the batch
        st->hist_pool_per_hist *= 4;
        continue;
      }
      if (err == -1) fail(MR_ERR_UNSUPPORTED, "per-request tag histograms exceed the scratch pool");
      if (err == MR_ERR_ARITHMETIC)
        fail(MR_ERR_ARITHMETIC, "/ by zero: normalized rate with a global top counter of 0 (java.lang.ArithmeticException in RateFeature.value)");
      break;
    }
  });
}

mr_status mr_rank_device(mr_state *st, mr_model *model, const mr_rank_batch *b, int32_t total_items,
                         double *d_out_scores, int32_t *d_out_order, double *d_out_features, void *cuda_stream) {
  return guard([&] {
    if (!st || !b) fail(MR_ERR_INVALID_ARG, "null argument");
    if (model) check_model(model);
    check_scored_dim(st, model);
    const int R = b->n_requests, N = total_items;
    if (R <= 0 || N <= 0) return;
    const Schema &S = st->store->schema;
    MR_CUDA_CHECK(cudaSetDevice(st->ctx->device));
    RankInFlight inflight(st);
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const bool fused = fused_codes(model);
    ScratchPlan sp = plan_scratch(S, R, N, st->hist_pool_per_hist, d_out_features == nullptr && !fused, fused ? model->code_cols() : 0);
    if (sp.total > st->d_scratch_cap) {
      MR_CUDA_CHECK(cudaDeviceSynchronize());
      if (st->d_scratch) cudaFree(st->d_scratch);
      st->d_scratch = nullptr;
      MR_CUDA_CHECK(cudaMalloc((void **)&st->d_scratch, sp.total));
      st->d_scratch_cap = sp.total;
    }
    RankArgs a{};
    fill_args(a, st, st->d_scratch, sp);
    a.error_flag = st->d_error;
    a.n_requests = R;
    a.total_items = N;
    a.item_offsets = b->item_offsets;
    a.item_ids = b->item_ids;
    a.user_ids = b->user_ids;
    a.session_ids = b->session_ids;
    a.req_f64 = b->req_f64;
    a.req_u64 = b->req_u64;
    a.req_vec = b->req_vec;
    a.req_vec_present = b->req_vec_present;
    a.item_f64 = b->item_f64;
    a.req_tok_off = (S.in_req_tok.empty() || !b->req_tok_hashes) ? nullptr : b->req_tok_offsets;
    a.req_tok_hash = b->req_tok_hashes;
    a.req_tok_w = b->req_tok_weights;
    a.req_tok_base = 0;
    a.out_features = d_out_features ? d_out_features : (fused ? nullptr : (double *)(st->d_scratch + sp.features));
    if (fused) set_codes(a, st, model, st->d_scratch, sp);
    launch_assemble(a, S, stream);
    if (model) {
      if (!d_out_scores) fail(MR_ERR_INVALID_ARG, "d_out_scores is null");
      if (fused) model->score_codes(a.codes, N, d_out_scores, stream);
      else model->score(a.out_features, N, S.dim, d_out_scores, stream);
      if (d_out_order) launch_rank_order(d_out_scores, a.item_offsets, R, N, d_out_order, stream, b->max_items_per_request);
    }
  });
}

mr_status mr_rank_device_status(mr_state *st, void *cuda_stream) {
  return guard([&] {
    if (!st) fail(MR_ERR_INVALID_ARG, "null argument");
    MR_CUDA_CHECK(cudaSetDevice(st->ctx->device));
    MR_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    int32_t err = 0;
    MR_CUDA_CHECK(cudaMemcpy(&err, st->d_error, 4, cudaMemcpyDeviceToHost));
    if (err == -1) {
      st->hist_pool_per_hist *= 4;
      fail(MR_ERR_UNSUPPORTED, "per-request tag histograms exceeded the scratch pool; pool grown, resubmit the batch");
    }
    if (err == MR_ERR_ARITHMETIC) fail(MR_ERR_ARITHMETIC, "/ by zero in normalized rate (global top counter is 0)");
    if (err == -2) fail(MR_ERR_CUDA, "a member of the group did not publish its slice within 2 s (peer process gone?)");
  });
}

}  // extern "C"

// =====================================================================================================
// Mega-request sharding (BASELINE configs[4]: one 10 000-item request, 2000 trees, 8 GPUs).
//
// Items are independent given the request, so a request too large for one GPU's latency budget is split into
// contiguous item ranges, one per member of an mr_group.  Every member assembles and scores ITS range only;
// the scorer's final store writes each score into the exchange buffer of EVERY member through peer memory
// (NVLink / NVSwitch; CUDA IPC between processes, direct peer access inside one process) and the last CTA
// raises the member's flag on every peer — the all-gather is the scorer's epilogue.  A one-warp kernel then
// waits for the other members' flags and the ordering kernels rank the full score vector.  No host round
// trip, no NCCL call, no allocation per request.
//
// Exchange buffer of a member (cudaMalloc, exported by IPC handle):
//   +0     u32 flag[2][8]      flag[parity][member] = sequence number of the last request that member published
//   +256   f64 score[2][cap]   double-buffered by request parity: a member may start request s+1 while a slow
//                              peer still orders request s (it cannot reach s+2 before that peer published s+1,
//                              which it does after it finished reading s)
// Members of one process that SHARE a device (tests put a whole group on one GPU) meet here between publishing and
// waiting: a host thread that allocates or frees device memory synchronises the device, so it would sit behind a
// peer's wait kernel that in turn waits for this thread's publish.  With one member per GPU — the deployment — there
// is nothing to meet about (a synchronising call only ever waits for its own device) and the pointer stays null.
struct LocalRendezvous {
  std::mutex m;
  std::condition_variable cv;
  int n = 0, arrived = 0;
  uint64_t generation = 0;
  void arrive_and_wait() {
    std::unique_lock<std::mutex> lk(m);
    const uint64_t gen = generation;
    if (++arrived == n) { arrived = 0; generation++; cv.notify_all(); return; }
    cv.wait_for(lk, std::chrono::seconds(3), [&] { return generation != gen; });  // bounded like the device-side wait
  }
};

struct mr_group {
  std::shared_ptr<LocalRendezvous> rendezvous;  // non-null only when local members share a device
  mr_ctx *ctx = nullptr;
  int rank = 0, world = 1, cap = 0;  // cap: item capacity (multiple of 128)
  uint8_t *xbuf = nullptr;
  size_t xbytes = 0;
  uint8_t *peer[8] = {};
  bool ipc_open[8] = {};
  bool connected = false;
  uint32_t seq = 0;
  uint32_t *d_done = nullptr;   // CTA arrival counter of the publishing kernel
  int32_t *d_offs = nullptr;    // {0, n_assembled, 0, N}
  int32_t *d_order = nullptr;   // [cap]
  int32_t *d_rank_tmp = nullptr;  // [cap] scratch of the mega-request ordering
  int32_t *d_err = nullptr;
  std::mutex mu;                // one collective at a time per member

  double *scores(int member, int parity) const { return (double *)(peer[member] + 256) + (size_t)parity * cap; }
  uint32_t *flag(int member, int parity, int who) const { return (uint32_t *)peer[member] + parity * 8 + who; }
};

namespace {

// Contiguous slice of member `rank`: ceil(N / G) rounded up to whole 128-item scorer tiles.
inline void group_slice(int n_items, int world, int rank, int *lo, int *hi) {
  const int per = ((((n_items + world - 1) / world) + 127) / 128) * 128;
  *lo = std::min(n_items, rank * per);
  *hi = std::min(n_items, *lo + per);
}

__global__ void group_prologue_kernel(int32_t *offs, int n_assembled, int n_total) {
  offs[0] = 0; offs[1] = n_assembled; offs[2] = 0; offs[3] = n_total;
}

// Fallback publisher: scorers that cannot store to the sinks themselves (the exact f64/f32 kernel, the generic
// binned kernel) and empty slices.  local == nullptr publishes nothing but still raises the flags.
__global__ void __launch_bounds__(256) group_publish_kernel(const double *local, int n, const ScoreSinks s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (local && i < n) store_score(nullptr, s, i, local[i]);
  publish_when_last(s);
}

// One lane per member: spin (acquire, system scope) until its flag carries this request's sequence number.
// Bounded: a peer that never arrives (crashed process) raises the error flag after `timeout_ns` instead of
// hanging the GPU.
__global__ void group_wait_kernel(const uint32_t *flags, int world, uint32_t seq, int32_t *err, unsigned long long timeout_ns) {
  const int g = threadIdx.x;
  if (g >= world) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + g) : "memory");
    if (v == seq) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) { atomicExch(err, -2); return; }
    __nanosleep(64);
  }
}

struct GroupCall {
  int N = 0, lo = 0, hi = 0;
  bool sliced = false;  // only [lo, hi) was assembled (else the whole request)
};

// Enqueues on `stream`: [assemble] -> score this member's slice (the stores go to every member) -> wait for the
// other members -> order the full score vector.  `a` carries the device pointers of the assembled item range.
void group_enqueue(mr_group *g, mr_state *st, mr_model *model, RankArgs &a, const ScratchPlan &sp, uint8_t *scratch,
                   const GroupCall &gc, cudaStream_t stream) {
  const Schema &S = st->store->schema;
  const bool fused = fused_codes(model);
  const uint32_t seq = g->seq;
  const int parity = (int)(seq & 1u);
  const int n_sl = gc.hi - gc.lo, first = gc.sliced ? 0 : gc.lo;
  { ProfScope _ps("group_prologue_kernel", stream); group_prologue_kernel<<<1, 1, 0, stream>>>(g->d_offs, a.total_items, gc.N); }
  MR_CUDA_CHECK(cudaGetLastError());
  a.item_offsets = g->d_offs;
  a.out_features = fused ? nullptr : (double *)(scratch + sp.features);
  if (fused) set_codes(a, st, model, scratch, sp, /*legacy_layout=*/true);
  if (a.total_items > 0) launch_assemble(a, S, stream);
  else MR_CUDA_CHECK(cudaMemsetAsync(a.error_flag, 0, 4, stream));
  ScoreSinks sk;
  sk.n_peer = g->world;
  for (int m = 0; m < g->world; m++) {
    sk.peer[m] = g->scores(m, parity);
    sk.flag[m] = g->flag(m, parity, g->rank);
  }
  sk.done = g->d_done;
  sk.seq = seq;
  sk.item_base = gc.lo;
  bool published = false;
  if (n_sl > 0) {
    if (fused && model->fuses_sinks(n_sl)) {
      uint16_t *codes = a.codes + (size_t)(first / 32) * model->code_cols() * 32;
      model->score_codes(codes, n_sl, nullptr, stream, sp.leafvals ? scratch + sp.leafvals : nullptr, &sk, /*layout=*/0);
      published = true;
    } else {
      double *local = (double *)(scratch + sp.local_scores);
      if (fused) model->score_codes(a.codes + (size_t)(first / 32) * model->code_cols() * 32, n_sl, local, stream, nullptr, nullptr, /*layout=*/0);
      else model->score(a.out_features + (size_t)first * S.dim, n_sl, S.dim, local, stream);
      { ProfScope _ps("group_publish_kernel", stream); group_publish_kernel<<<(n_sl + 255) / 256, 256, 0, stream>>>(local, n_sl, sk); }
      MR_CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
      published = true;
    }
  }
  if (!published) {
    { ProfScope _ps("group_publish_kernel", stream); group_publish_kernel<<<1, 256, 0, stream>>>(nullptr, 0, sk); }
    MR_CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
  if (g->rendezvous) g->rendezvous->arrive_and_wait();  // members sharing a device: everybody has published (see LocalRendezvous)
  { ProfScope _ps("group_wait_kernel", stream); group_wait_kernel<<<1, 32, 0, stream>>>(g->flag(g->rank, parity, 0), g->world, seq, a.error_flag, 2000000000ull); }
  MR_CUDA_CHECK(cudaGetLastError());
  launch_rank_order(g->scores(g->rank, parity), g->d_offs + 2, 1, gc.N, g->d_order, stream, gc.N, g->d_rank_tmp);
  g_kernel_launches += 2;
}

void group_check(mr_group *g, mr_state *st, mr_model *model, int n_requests, int N) {
  if (!g || !st || !model) fail(MR_ERR_INVALID_ARG, "null argument");
  check_model(model);
  if (!g->connected) fail(MR_ERR_INVALID_ARG, "group is not connected (mr_group_connect / mr_group_connect_local)");
  if (st->ctx != g->ctx || model->ctx != g->ctx) fail(MR_ERR_INVALID_ARG, "state / model belong to another context than the group member");
  if (n_requests != 1) fail(MR_ERR_UNSUPPORTED, "mr_group_rank splits ONE request; batches of ordinary requests are sharded by request (mr_rank per GPU)");
  if (N < 0 || N > g->cap) fail(MR_ERR_INVALID_ARG, "request has %d items, the group was created for %d", N, g->cap);
  check_scored_dim(st, model);
}

}  // namespace

extern "C" {

mr_status mr_group_create(mr_ctx *ctx, int32_t rank, int32_t world, int32_t max_items, mr_group **out) {
  return guard([&] {
    if (!ctx || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (world < 1 || world > 8 || rank < 0 || rank >= world) fail(MR_ERR_INVALID_ARG, "rank %d / world %d (1..8 members)", rank, world);
    if (max_items < 1) fail(MR_ERR_INVALID_ARG, "max_items must be positive");
    auto g = std::make_unique<mr_group>();
    g->ctx = ctx;
    g->rank = rank;
    g->world = world;
    g->cap = ((max_items + 127) / 128) * 128;
    g->xbytes = 256 + (size_t)2 * g->cap * 8;
    MR_CUDA_CHECK(cudaSetDevice(ctx->device));
    MR_CUDA_CHECK(cudaMalloc((void **)&g->xbuf, g->xbytes));
    MR_CUDA_CHECK(cudaMemset(g->xbuf, 0, g->xbytes));
    MR_CUDA_CHECK(cudaMalloc((void **)&g->d_done, 256));
    MR_CUDA_CHECK(cudaMemset(g->d_done, 0, 256));
    g->d_offs = (int32_t *)((uint8_t *)g->d_done + 64);
    g->d_err = (int32_t *)((uint8_t *)g->d_done + 128);
    MR_CUDA_CHECK(cudaMalloc((void **)&g->d_order, (size_t)g->cap * 4 * 4));
    g->d_rank_tmp = g->d_order + g->cap;  // 3 ints per item, 8-byte aligned (cap is a multiple of 128)
    g->peer[rank] = g->xbuf;
    g->connected = world == 1;
    *out = g.release();
  });
}

mr_status mr_group_export(mr_group *g, uint8_t *handle) {
  return guard([&] {
    if (!g || !handle) fail(MR_ERR_INVALID_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MR_GROUP_HANDLE_BYTES, "handle size");
    MR_CUDA_CHECK(cudaSetDevice(g->ctx->device));
    cudaIpcMemHandle_t h;
    MR_CUDA_CHECK(cudaIpcGetMemHandle(&h, g->xbuf));
    memcpy(handle, &h, sizeof h);
  });
}

mr_status mr_group_connect(mr_group *g, const uint8_t *handles) {
  return guard([&] {
    if (!g || !handles) fail(MR_ERR_INVALID_ARG, "null argument");
    MR_CUDA_CHECK(cudaSetDevice(g->ctx->device));
    for (int m = 0; m < g->world; m++) {
      if (m == g->rank || g->peer[m]) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + (size_t)m * MR_GROUP_HANDLE_BYTES, sizeof h);
      void *p = nullptr;
      MR_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      g->peer[m] = (uint8_t *)p;
      g->ipc_open[m] = true;
    }
    g->connected = true;
  });
}

mr_status mr_group_connect_local(mr_group *const *members, int32_t world) {
  return guard([&] {
    if (!members || world < 1 || world > 8) fail(MR_ERR_INVALID_ARG, "bad member list");
    for (int m = 0; m < world; m++)
      if (!members[m] || members[m]->world != world || members[m]->rank != m)
        fail(MR_ERR_INVALID_ARG, "members[%d] must be the group member of rank %d in a world of %d", m, m, world);
    for (int a = 0; a < world; a++) {
      mr_group *g = members[a];
      MR_CUDA_CHECK(cudaSetDevice(g->ctx->device));
      for (int b = 0; b < world; b++) {
        const int db = members[b]->ctx->device;
        if (db != g->ctx->device) {
          int can = 0;
          MR_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, g->ctx->device, db));
          if (!can) fail(MR_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", g->ctx->device, db);
          cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
          else MR_CUDA_CHECK(e);
        }
        g->peer[b] = members[b]->xbuf;
      }
      g->connected = true;
    }
    bool shared = false;
    for (int a = 0; a < world; a++)
      for (int b = a + 1; b < world; b++) shared |= members[a]->ctx->device == members[b]->ctx->device;
    if (shared) {
      auto rv = std::make_shared<LocalRendezvous>();
      rv->n = world;
      for (int a = 0; a < world; a++) members[a]->rendezvous = rv;
    }
  });
}

mr_status mr_group_free(mr_group *g) {
  if (!g) return MR_OK;
  cudaSetDevice(g->ctx->device);
  cudaDeviceSynchronize();
  for (int m = 0; m < g->world; m++)
    if (g->ipc_open[m]) cudaIpcCloseMemHandle(g->peer[m]);
  if (g->xbuf) cudaFree(g->xbuf);
  if (g->d_done) cudaFree(g->d_done);
  if (g->d_order) cudaFree(g->d_order);
  delete g;
  return MR_OK;
}

void mr_group_slice(int32_t n_items, int32_t world, int32_t rank, int32_t *lo, int32_t *hi) {
  int a = 0, b = 0;
  if (world > 0 && rank >= 0 && rank < world && n_items >= 0) group_slice(n_items, world, rank, &a, &b);
  if (lo) *lo = a;
  if (hi) *hi = b;
}

mr_status mr_group_rank(mr_group *g, mr_state *st, mr_model *model, const mr_rank_batch *b, double *out_scores,
                        int32_t *out_order) {
  return guard([&] {
    if (!b || !b->item_offsets) fail(MR_ERR_INVALID_ARG, "null argument");
    const int N = b->n_requests == 1 ? b->item_offsets[1] - b->item_offsets[0] : 0;
    group_check(g, st, model, b->n_requests, N);
    if (b->item_offsets[0] != 0) fail(MR_ERR_INVALID_ARG, "item_offsets[0] must be 0");
    if (N > 0 && (!b->item_ids || !out_scores)) fail(MR_ERR_INVALID_ARG, "null item / score arrays");
    const Schema &S = st->store->schema;
    if (!S.in_req_f64.empty() && !b->req_f64) fail(MR_ERR_INVALID_ARG, "schema needs req_f64 inputs");
    if (!S.in_req_u64.empty() && !b->req_u64) fail(MR_ERR_INVALID_ARG, "schema needs req_u64 inputs");
    if (!S.in_req_vec.empty() && (!b->req_vec || !b->req_vec_present)) fail(MR_ERR_INVALID_ARG, "schema needs req_vec inputs");
    if (!S.in_req_tok.empty() && (!b->req_tok_offsets || !b->req_tok_hashes)) fail(MR_ERR_INVALID_ARG, "schema needs req_tok inputs");
    std::lock_guard<std::mutex> one(g->mu);
    InflightGuard ig(model);
    MR_CUDA_CHECK(cudaSetDevice(g->ctx->device));
    RankInFlight inflight(st);
    std::shared_lock<std::shared_mutex> read_guard(st->store->mu);
    GroupCall gc;
    gc.N = N;
    group_slice(N, g->world, g->rank, &gc.lo, &gc.hi);
    // per-request aggregates (diversity, interacted_with histograms, cosine normalisation) read the WHOLE item list:
    // such schemas assemble the full request on every member and only the scoring is split
    gc.sliced = !S.needs_prepass;
    const int a0 = gc.sliced ? gc.lo : 0, a1 = gc.sliced ? gc.hi : N, n_asm = a1 - a0;
    LaneGuard lg(g->ctx);
    Lane *lane = lg.lane.get();
    struct Seg { const void *src; size_t bytes, off; };
    Seg segs[12];
    size_t in_bytes = 0;
    int ns = 0;
    auto seg = [&](const void *p, size_t bytes) { segs[ns] = Seg{p, p ? bytes : 0, in_bytes}; in_bytes += al(segs[ns].bytes); return ns++; };
    const size_t nrf = S.in_req_f64.size(), nru = S.in_req_u64.size(), nrv = S.in_req_vec.size(), nif = S.in_item_f64.size(), ntk = S.in_req_tok.size();
    const int s_ids = seg(b->item_ids ? b->item_ids + a0 : nullptr, (size_t)n_asm * 8);
    const int s_usr = seg(b->user_ids, 8);
    const int s_ses = seg(b->session_ids, 8);
    const int s_rf = seg(b->req_f64, nrf * 8);
    const int s_ru = seg(b->req_u64, nru * 8);
    const int s_rv = seg(b->req_vec, (size_t)S.vec_stride * 4);
    const int s_rp = seg(b->req_vec_present, nrv);
    const int s_if = seg(b->item_f64 ? b->item_f64 + (size_t)a0 * nif : nullptr, (size_t)n_asm * nif * 8);
    const bool has_tok = ntk > 0 && b->req_tok_offsets && b->req_tok_hashes;
    const int32_t t1 = has_tok ? b->req_tok_offsets[ntk] : 0;
    const int s_to = seg(has_tok ? b->req_tok_offsets : nullptr, (ntk + 1) * 4);
    const int s_th = seg(has_tok ? b->req_tok_hashes : nullptr, (size_t)t1 * 8);
    const int s_tw = seg(has_tok && b->req_tok_weights ? b->req_tok_weights : nullptr, (size_t)t1 * 8);
    const bool fused = fused_codes(model);
    const int n_sl = gc.hi - gc.lo;
    const size_t leaf_bytes = (fused && n_sl > 0 && model->use_latency(n_sl)) ? latency_scratch_bytes(n_sl, (int)model->host.trees.size()) : 0;
    ScratchPlan sp = plan_scratch(S, 1, std::max(n_asm, 1), st->hist_pool_per_hist, !fused, fused ? model->code_cols() : 0, leaf_bytes,
                                  (size_t)std::max(n_sl, 1) * 8);
    const size_t d_total = al(in_bytes) + sp.total + 256;
    const size_t out_bytes = al((size_t)N * 8) + al((size_t)N * 4) + 256;
    lane->ensure(in_bytes + out_bytes, d_total);
    for (int k = 0; k < ns; k++)
      if (segs[k].bytes) memcpy(lane->h_pinned + segs[k].off, segs[k].src, segs[k].bytes);
    uint8_t *d_in = lane->d_buf, *scratch = lane->d_buf + al(in_bytes);
    if (in_bytes) MR_CUDA_CHECK(cudaMemcpyAsync(d_in, lane->h_pinned, in_bytes, cudaMemcpyHostToDevice, lane->stream));
    auto dp = [&](int si) -> const void * { return segs[si].bytes ? d_in + segs[si].off : nullptr; };
    RankArgs a{};
    fill_args(a, st, scratch, sp);
    a.n_requests = 1;
    a.total_items = n_asm;
    a.item_ids = (const uint64_t *)dp(s_ids);
    a.user_ids = (const uint64_t *)dp(s_usr);
    a.session_ids = (const uint64_t *)dp(s_ses);
    a.req_f64 = (const double *)dp(s_rf);
    a.req_u64 = (const uint64_t *)dp(s_ru);
    a.req_vec = (const float *)dp(s_rv);
    a.req_vec_present = (const uint8_t *)dp(s_rp);
    a.item_f64 = (const double *)dp(s_if);
    a.req_tok_off = has_tok ? (const int32_t *)(d_in + segs[s_to].off) : nullptr;
    a.req_tok_hash = (const uint64_t *)dp(s_th);
    a.req_tok_w = (const double *)dp(s_tw);
    a.req_tok_base = 0;
    a.error_flag = (int32_t *)(lane->d_buf + al(in_bytes) + sp.total);
    g->seq++;
    group_enqueue(g, st, model, a, sp, scratch, gc, lane->stream);
    uint8_t *h_out = lane->h_pinned + in_bytes;
    const size_t ho_order = al((size_t)N * 8), ho_err = ho_order + al((size_t)N * 4);
    const int parity = (int)(g->seq & 1u);
    if (N > 0) MR_CUDA_CHECK(cudaMemcpyAsync(h_out, g->scores(g->rank, parity), (size_t)N * 8, cudaMemcpyDeviceToHost, lane->stream));
    if (N > 0 && out_order) MR_CUDA_CHECK(cudaMemcpyAsync(h_out + ho_order, g->d_order, (size_t)N * 4, cudaMemcpyDeviceToHost, lane->stream));
    MR_CUDA_CHECK(cudaMemcpyAsync(h_out + ho_err, a.error_flag, 4, cudaMemcpyDeviceToHost, lane->stream));
    MR_CUDA_CHECK(cudaStreamSynchronize(lane->stream));
    int32_t err;
    memcpy(&err, h_out + ho_err, 4);
    if (err == -2) fail(MR_ERR_CUDA, "a member of the group did not publish its slice within 2 s (peer process gone?)");
    if (err == -1) { st->hist_pool_per_hist *= 4; fail(MR_ERR_UNSUPPORTED, "per-request tag histograms exceeded the scratch pool; pool grown, resubmit on every member"); }
    if (err == MR_ERR_ARITHMETIC) fail(MR_ERR_ARITHMETIC, "/ by zero in normalized rate (global top counter is 0)");
    if (N > 0) memcpy(out_scores, h_out, (size_t)N * 8);
    if (N > 0 && out_order) memcpy(out_order, h_out + ho_order, (size_t)N * 4);
  });
}

mr_status mr_group_rank_device(mr_group *g, mr_state *st, mr_model *model, const mr_rank_batch *b, int32_t total_items,
                               double *d_out_scores, int32_t *d_out_order, void *cuda_stream) {
  return guard([&] {
    if (!b) fail(MR_ERR_INVALID_ARG, "null argument");
    const int N = total_items;
    group_check(g, st, model, b->n_requests, N);
    const Schema &S = st->store->schema;
    std::lock_guard<std::mutex> one(g->mu);
    MR_CUDA_CHECK(cudaSetDevice(g->ctx->device));
    RankInFlight inflight(st);
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    GroupCall gc;
    gc.N = N;
    group_slice(N, g->world, g->rank, &gc.lo, &gc.hi);
    gc.sliced = !S.needs_prepass;
    const int a0 = gc.sliced ? gc.lo : 0, a1 = gc.sliced ? gc.hi : N, n_asm = a1 - a0, n_sl = gc.hi - gc.lo;
    const bool fused = fused_codes(model);
    const size_t leaf_bytes = (fused && n_sl > 0 && model->use_latency(n_sl)) ? latency_scratch_bytes(n_sl, (int)model->host.trees.size()) : 0;
    ScratchPlan sp = plan_scratch(S, 1, std::max(n_asm, 1), st->hist_pool_per_hist, !fused, fused ? model->code_cols() : 0, leaf_bytes,
                                  (size_t)std::max(n_sl, 1) * 8);
    if (sp.total > st->d_scratch_cap) {
      MR_CUDA_CHECK(cudaDeviceSynchronize());
      if (st->d_scratch) cudaFree(st->d_scratch);
      st->d_scratch = nullptr;
      MR_CUDA_CHECK(cudaMalloc((void **)&st->d_scratch, sp.total));
      st->d_scratch_cap = sp.total;
    }
    const size_t nif = S.in_item_f64.size();
    RankArgs a{};
    fill_args(a, st, st->d_scratch, sp);
    a.error_flag = st->d_error;
    a.n_requests = 1;
    a.total_items = n_asm;
    a.item_ids = b->item_ids ? b->item_ids + a0 : nullptr;
    a.user_ids = b->user_ids;
    a.session_ids = b->session_ids;
    a.req_f64 = b->req_f64;
    a.req_u64 = b->req_u64;
    a.req_vec = b->req_vec;
    a.req_vec_present = b->req_vec_present;
    a.item_f64 = b->item_f64 ? b->item_f64 + (size_t)a0 * nif : nullptr;
    a.req_tok_off = (S.in_req_tok.empty() || !b->req_tok_hashes) ? nullptr : b->req_tok_offsets;
    a.req_tok_hash = b->req_tok_hashes;
    a.req_tok_w = b->req_tok_weights;
    a.req_tok_base = 0;
    g->seq++;
    group_enqueue(g, st, model, a, sp, st->d_scratch, gc, stream);
    const int parity = (int)(g->seq & 1u);
    if (d_out_scores && N > 0) MR_CUDA_CHECK(cudaMemcpyAsync(d_out_scores, g->scores(g->rank, parity), (size_t)N * 8, cudaMemcpyDeviceToDevice, stream));
    if (d_out_order && N > 0) MR_CUDA_CHECK(cudaMemcpyAsync(d_out_order, g->d_order, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
  });
}

}  // extern "C"
