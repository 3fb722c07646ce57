// C ABI of libmrgpu.so (include/mr_b200.h): context, booster handles, predictMat.
// Everything here is host code around the kernels; exceptions are converted to
// mr_status at the boundary and never escape.
#include "internal.h"

using namespace mr;

namespace mr {
thread_local std::string t_last_error;
void check_model(mr_model *m) {
  if (!m) fail(MR_ERR_INVALID_ARG, "model handle is null");
  if (m->closed.load()) fail(MR_ERR_CLOSED, "booster is closed");
}
}  // namespace mr

namespace {

__global__ void dyn_smem_base_probe(int *out) {
  extern __shared__ uint8_t probe_smem[];
  *out = (int)__cvta_generic_to_shared(probe_smem);
}

void check_matrix(mr_model *m, const void *values, int rows, int cols, const void *out) {
  if (rows < 0 || cols < 0) fail(MR_ERR_INVALID_ARG, "negative matrix dimension");
  if (rows > 0 && (!values || !out)) fail(MR_ERR_INVALID_ARG, "null matrix or output pointer");
  if (rows > 0 && cols != m->host.n_features)
    fail(MR_ERR_INVALID_ARG, "matrix has %d columns, the booster was trained with %d features", cols,
         m->host.n_features);
  if ((int64_t)rows * cols >= (int64_t)INT32_MAX)
    fail(MR_ERR_INVALID_ARG, "matrix of %d x %d exceeds the 2^31 cell limit", rows, cols);
}

mr_model *make_model(mr_ctx *ctx, HostModel &&hm, int n_features) {
  if (n_features > 0 && hm.n_features != n_features)
    fail(MR_ERR_FEATURE_MISMATCH, "booster reads %d features, dataset descriptor has %d", hm.n_features, n_features);
  auto m = std::make_unique<mr_model>();
  m->ctx = ctx;
  m->host = std::move(hm);
  MR_CUDA_CHECK(cudaSetDevice(ctx->device));
  m->repack();
  ctx->live_models++;
  return m.release();
}

}  // namespace

extern "C" {

const char *mr_last_error(void) { return t_last_error.c_str(); }
const char *mr_version(void) { return "libmrgpu 0.1.0 sm_100a"; }
int64_t mr_kernel_launches(void) { return (int64_t)g_kernel_launches; }

mr_status mr_profile_begin(void) {
  return guard([&] { profile_begin(); });
}
mr_status mr_profile_end(char *json, size_t cap, size_t *len) {
  return guard([&] {
    std::string s;
    profile_end(s);
    if (len) *len = s.size();
    if (json) {
      if (cap < s.size() + 1) fail(MR_ERR_INVALID_ARG, "profile buffer too small: %zu bytes needed", s.size() + 1);
      memcpy(json, s.c_str(), s.size() + 1);
    }
  });
}

mr_status mr_init(int32_t device, mr_ctx **out) {
  return guard([&] {
    if (!out) fail(MR_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      fail(MR_ERR_NO_DEVICE, "no CUDA device available (%s); libmrgpu has no CPU fallback",
           e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) fail(MR_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    MR_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      fail(MR_ERR_NO_DEVICE, "device %d is sm_%d%d; libmrgpu is built for sm_100a only", device, prop.major, prop.minor);
    MR_CUDA_CHECK(cudaSetDevice(device));
    {  // stream-ordered scratch (cudaMallocAsync) must never give memory back to the OS mid-serving:
       // a trim at a synchronisation point shows up as a multi-millisecond latency spike
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    }
    auto ctx = std::make_unique<mr_ctx>();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    {  // where a kernel's dynamic shared window starts in the shared address space (the slim scorer wants to know)
      int *d = nullptr, h = -1;
      if (cudaMalloc((void **)&d, 4) == cudaSuccess) {
        dyn_smem_base_probe<<<1, 32, 1024>>>(d);
        if (cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost) != cudaSuccess) h = -1;
        cudaFree(d);
      }
      cudaGetLastError();
      ctx->dyn_smem_base = h;
    }
    *out = ctx.release();
  });
}

mr_status mr_shutdown(mr_ctx *ctx) {
  return guard([&] {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto &l : ctx->free_lanes) {
      if (l->h_pinned) cudaFreeHost(l->h_pinned);
      if (l->d_buf) cudaFree(l->d_buf);
      if (l->stream) cudaStreamDestroy(l->stream);
    }
    delete ctx;
  });
}

mr_status mr_model_load(mr_ctx *ctx, int32_t kind, const uint8_t *blob, size_t len, int32_t n_features,
                        mr_model **out) {
  return guard([&] {
    if (!ctx || !blob || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    HostModel hm;
    if (kind == MR_BOOSTER_LIGHTGBM) hm = parse_lightgbm_text(blob, len);
    else if (kind == MR_BOOSTER_XGBOOST) hm = parse_xgboost_model(blob, len);
    else fail(MR_ERR_UNSUPPORTED, "unsupported booster tag %d", kind);
    *out = make_model(ctx, std::move(hm), n_features);
  });
}

mr_status mr_model_load_metarank(mr_ctx *ctx, const uint8_t *blob, size_t len, const char *const *feature_names,
                                 int32_t n_names, mr_model **out) {
  return guard([&] {
    if (!ctx || !blob || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    std::vector<std::string> names;
    int kind = 0;
    size_t b = 0, e = 0;
    parse_metarank_frame(blob, len, names, kind, b, e);
    if (n_names >= 0) {
      bool same = (size_t)n_names == names.size();
      for (int i = 0; same && i < n_names; i++) same = feature_names && feature_names[i] && names[i] == feature_names[i];
      if (!same) {
        std::string exp, act;
        for (auto &s : names) exp += (exp.empty() ? "" : ", ") + s;
        for (int i = 0; i < n_names; i++) act += (act.empty() ? "" : ", ") + std::string(feature_names && feature_names[i] ? feature_names[i] : "<null>");
        fail(MR_ERR_FEATURE_MISMATCH,
             "booster trained with List(%s) features, but config defines List(%s)\nYou may need to retrain the model with the newer config",
             exp.c_str(), act.c_str());
      }
    }
    HostModel hm = kind == 0 ? parse_lightgbm_text(blob + b, e - b) : parse_xgboost_model(blob + b, e - b);
    *out = make_model(ctx, std::move(hm), 0);
  });
}

mr_status mr_model_predict_mat_device(mr_model *m, const double *d_values, int32_t rows, int32_t cols,
                                      double *d_out_scores, void *cuda_stream) {
  return guard([&] {
    check_model(m);
    InflightGuard ig(m);
    check_matrix(m, d_values, rows, cols, d_out_scores);
    if (rows == 0) return;
    m->score(d_values, rows, cols, d_out_scores, (cudaStream_t)cuda_stream);
  });
}

// Host-buffer predictMat.  Large batches are cut into slices that alternate between two
// lanes (stream + buffers) so the H2D copy of slice i+1 overlaps the kernel of slice i and
// the D2H of slice i-1.  Page-locked caller buffers are DMA'd directly; pageable ones are
// staged through the lane's pinned buffer.
static bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

mr_status mr_model_predict_mat(mr_model *m, const double *values, int32_t rows, int32_t cols, double *out_scores) {
  return guard([&] {
    check_model(m);
    InflightGuard ig(m);
    check_matrix(m, values, rows, cols, out_scores);
    if (rows == 0) return;
    MR_CUDA_CHECK(cudaSetDevice(m->ctx->device));
    const int kSlice = 1 << 17;  // rows per slice
    const int n_slices = (rows + kSlice - 1) / kSlice;
    const bool direct_in = is_pinned(values), direct_out = is_pinned(out_scores);
    LaneGuard lane0(m->ctx);
    std::unique_ptr<LaneGuard> lane1;
    if (n_slices > 1) lane1 = std::make_unique<LaneGuard>(m->ctx);
    Lane *lanes[2] = {lane0.lane.get(), lane1 ? lane1->lane.get() : lane0.lane.get()};
    const int slice_rows = std::min(rows, kSlice);
    const size_t in_cap = (size_t)slice_rows * cols * sizeof(double), out_cap = (size_t)slice_rows * sizeof(double);
    const size_t d_out_off = (in_cap + 255) & ~size_t(255);
    for (int l = 0; l < (n_slices > 1 ? 2 : 1); l++)
      lanes[l]->ensure((direct_in ? 0 : in_cap) + (direct_out ? 0 : out_cap) + 16, d_out_off + out_cap);
    struct Pending { int r0, n; };
    Pending pend[2] = {{0, 0}, {0, 0}};
    auto drain = [&](int l) {
      if (pend[l].n == 0) return;
      MR_CUDA_CHECK(cudaStreamSynchronize(lanes[l]->stream));
      if (!direct_out)
        memcpy(out_scores + pend[l].r0, lanes[l]->h_pinned + (direct_in ? 0 : in_cap), (size_t)pend[l].n * sizeof(double));
      pend[l].n = 0;
    };
    for (int i = 0; i < n_slices; i++) {
      const int l = i & 1;
      Lane *ln = lanes[l];
      drain(l);  // buffers of this lane are free again
      const int r0 = i * kSlice, n = std::min(kSlice, rows - r0);
      const size_t in_bytes = (size_t)n * cols * sizeof(double), out_bytes = (size_t)n * sizeof(double);
      const double *src = values + (size_t)r0 * cols;
      if (!direct_in) {
        memcpy(ln->h_pinned, src, in_bytes);
        src = (const double *)ln->h_pinned;
      }
      double *d_in = (double *)ln->d_buf, *d_out = (double *)(ln->d_buf + d_out_off);
      MR_CUDA_CHECK(cudaMemcpyAsync(d_in, src, in_bytes, cudaMemcpyHostToDevice, ln->stream));
      m->score(d_in, n, cols, d_out, ln->stream);
      double *dst = direct_out ? out_scores + r0 : (double *)(ln->h_pinned + (direct_in ? 0 : in_cap));
      MR_CUDA_CHECK(cudaMemcpyAsync(dst, d_out, out_bytes, cudaMemcpyDeviceToHost, ln->stream));
      pend[l] = {r0, n};
    }
    drain(0);
    drain(1);
  });
}

mr_status mr_model_count_path(mr_model *m, const double *values, int32_t rows, int32_t cols, double *mean_path) {
  return guard([&] {
    check_model(m);
    InflightGuard ig(m);
    if (!mean_path) fail(MR_ERR_INVALID_ARG, "mean_path is null");
    std::vector<double> scratch((size_t)std::max(rows, 1));
    check_matrix(m, values, rows, cols, scratch.data());
    *mean_path = 0.0;
    if (rows == 0 || m->host.trees.empty()) return;
    MR_CUDA_CHECK(cudaSetDevice(m->ctx->device));
    LaneGuard lane(m->ctx);
    const size_t in_bytes = (size_t)rows * cols * sizeof(double), out_bytes = (size_t)rows * sizeof(double);
    const size_t d_out_off = (in_bytes + 255) & ~size_t(255), d_cnt_off = (d_out_off + out_bytes + 255) & ~size_t(255);
    lane->ensure(in_bytes + 8, d_cnt_off + 8);
    memcpy(lane->h_pinned, values, in_bytes);
    MR_CUDA_CHECK(cudaMemcpyAsync(lane->d_buf, lane->h_pinned, in_bytes, cudaMemcpyHostToDevice, lane->stream));
    MR_CUDA_CHECK(cudaMemsetAsync(lane->d_buf + d_cnt_off, 0, 8, lane->stream));
    ScoreLaunch L = m->launch_desc((double *)lane->d_buf, rows, cols, (double *)(lane->d_buf + d_out_off));
    L.d_visited = (unsigned long long *)(lane->d_buf + d_cnt_off);
    launch_gbdt_score(L, m->ctx->num_sms, lane->stream);
    unsigned long long cnt = 0;
    MR_CUDA_CHECK(cudaMemcpyAsync(lane->h_pinned, lane->d_buf + d_cnt_off, 8, cudaMemcpyDeviceToHost, lane->stream));
    MR_CUDA_CHECK(cudaStreamSynchronize(lane->stream));
    memcpy(&cnt, lane->h_pinned, 8);
    *mean_path = (double)cnt / ((double)rows * (double)m->host.trees.size());
  });
}

size_t mr_model_codes_bytes(mr_model *m, int32_t rows) {
  if (!m || m->closed.load() || rows <= 0 || !m->use_binned()) return 0;
  return binned_scratch_bytes(rows, m->code_cols());
}

mr_status mr_model_bin_device(mr_model *m, const double *d_values, int32_t rows, int32_t cols, void *d_codes,
                              void *cuda_stream) {
  return guard([&] {
    check_model(m);
    InflightGuard ig(m);
    check_matrix(m, d_values, rows, cols, d_codes);
    if (!m->use_binned()) fail(MR_ERR_UNSUPPORTED, "this model is scored by the f64/f32 kernel, it has no code form");
    if (rows == 0) return;
    BinnedLaunch B = m->binned_desc();
    B.d_values = d_values; B.rows = rows; B.cols = cols; B.d_bins = (uint16_t *)d_codes; B.codes_only = true;
    B.tile_T = m->code_layout(rows);  // what mr_model_score_codes_device will read for the same row count
    launch_gbdt_binned(B, m->ctx->num_sms, (cudaStream_t)cuda_stream);
  });
}

mr_status mr_model_score_codes_device(mr_model *m, const void *d_codes, int32_t rows, double *d_out_scores,
                                      void *cuda_stream) {
  return guard([&] {
    check_model(m);
    InflightGuard ig(m);
    if (rows < 0) fail(MR_ERR_INVALID_ARG, "negative row count");
    if (rows > 0 && (!d_codes || !d_out_scores)) fail(MR_ERR_INVALID_ARG, "null pointer");
    if (!m->use_binned()) fail(MR_ERR_UNSUPPORTED, "this model is scored by the f64/f32 kernel, it has no code form");
    if (rows == 0) return;
    m->score_codes((uint16_t *)d_codes, rows, d_out_scores, (cudaStream_t)cuda_stream);
  });
}

mr_status mr_model_save(mr_model *m, const uint8_t **blob, size_t *len) {
  return guard([&] {
    check_model(m);
    if (!blob || !len) fail(MR_ERR_INVALID_ARG, "null argument");
    *blob = m->host.blob.data();
    *len = m->host.blob.size();
  });
}

mr_status mr_model_weights(mr_model *m, double *out, int32_t n) {
  return guard([&] {
    check_model(m);
    if (!out || n < m->host.n_features) fail(MR_ERR_INVALID_ARG, "weights buffer too small");
    std::fill(out, out + n, 0.0);
    for (auto &t : m->host.trees)
      for (int f : t.feat) out[f] += 1.0;
  });
}

mr_status mr_model_get_info(mr_model *m, mr_model_info *out) {
  return guard([&] {
    if (!m || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    out->kind = m->host.kind;
    out->n_features = m->host.n_features;
    out->n_trees = (int32_t)m->host.trees.size();
    out->max_leaves = m->host.max_leaves;
    out->n_chunks = (int32_t)m->packed.chunks.size();
    out->has_categorical = m->host.has_cat;
    out->n_internal_nodes = m->host.n_internal;
    out->device_bytes = (int64_t)m->packed.bytes.size();
  });
}

mr_status mr_model_inspect(int32_t kind, const uint8_t *blob, size_t len, int32_t chunk_kb, mr_model_info *out) {
  return guard([&] {
    if (!blob || !out) fail(MR_ERR_INVALID_ARG, "null argument");
    HostModel hm;
    if (kind == MR_BOOSTER_LIGHTGBM) hm = parse_lightgbm_text(blob, len);
    else if (kind == MR_BOOSTER_XGBOOST) hm = parse_xgboost_model(blob, len);
    else fail(MR_ERR_UNSUPPORTED, "unsupported booster tag %d", kind);
    PackedModel pk = pack_model(hm, (size_t)(chunk_kb > 0 ? chunk_kb : 16) * 1024);
    out->kind = hm.kind;
    out->n_features = hm.n_features;
    out->n_trees = (int32_t)hm.trees.size();
    out->max_leaves = hm.max_leaves;
    out->n_chunks = (int32_t)pk.chunks.size();
    out->has_categorical = hm.has_cat;
    out->n_internal_nodes = hm.n_internal;
    out->device_bytes = (int64_t)pk.bytes.size();
  });
}

mr_status mr_model_selfcheck(int32_t kind, const uint8_t *blob, size_t len, int32_t samples, int32_t max_tile, int32_t *form, int64_t *mismatches) {
  return guard([&] {
    if (!blob || !form || !mismatches) fail(MR_ERR_INVALID_ARG, "null argument");
    HostModel hm;
    if (kind == MR_BOOSTER_LIGHTGBM) hm = parse_lightgbm_text(blob, len);
    else if (kind == MR_BOOSTER_XGBOOST) hm = parse_xgboost_model(blob, len);
    else fail(MR_ERR_UNSUPPORTED, "unsupported booster tag %d", kind);
    *form = 0; *mismatches = 0;
    const BinnedModel bn = pack_binned(hm, (size_t)(hm.n_features <= 32 ? 60 : 16) * 1024);
    if (!bn.ok) return;
    const BinnedModel cm = pack_compact(hm, bn, 0);
    if (!cm.ok) return;
    const SlimModel sl = pack_slim(hm, cm, 0, max_tile > 0 ? max_tile : 512, max_tile > 0 && max_tile < 512 ? 64 : 48);
    if (!sl.ok) return;
    *form = 1 | (sl.root_tab.empty() ? 0 : 2) | (sl.cat16 ? 4 : 0) | (sl.tile_T << 8);
    *mismatches = (int64_t)slim_pack_selfcheck(hm, cm, sl, std::max(1, samples), 0x9E3779B97F4A7C15ull);
  });
}

mr_status mr_model_walk_stats(mr_model *m, const double *d_values, int32_t rows, int32_t cols, double *lane_levels,
                              double *warp_levels, double *warp_trees, void *cuda_stream) {
  return guard([&] {
    check_model(m);
    check_matrix(m, d_values, rows, cols, d_values);
    if (!m->use_binned() || !m->use_compact()) fail(MR_ERR_UNSUPPORTED, "walk statistics exist for the compact binned scorers only");
    InflightGuard g(m);
    MR_CUDA_CHECK(cudaSetDevice(m->ctx->device));
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    BinnedLaunch B = m->binned_desc();  // the 8-byte compact form: same trees, same tile mapping as the slim one
    B.d_values = d_values; B.rows = rows; B.cols = cols;
    void *bins = nullptr;
    MR_CUDA_CHECK(cudaMalloc(&bins, std::max<size_t>(binned_scratch_bytes(rows, m->code_cols()), 16)));
    B.d_bins = (uint16_t *)bins;
    B.codes_only = true;  // groups-of-32 layout (tile_T = 0)
    unsigned long long out[3] = {0, 0, 0};
    try {
      launch_gbdt_binned(B, m->ctx->num_sms, stream);
      compact_walk_stats(B, out, stream);
    } catch (...) {
      cudaFree(bins);
      throw;
    }
    cudaFree(bins);
    if (lane_levels) *lane_levels = (double)out[0];
    if (warp_levels) *warp_levels = (double)out[1];
    if (warp_trees) *warp_trees = (double)out[2];
  });
}

mr_status mr_model_set_option(mr_model *m, const char *key, int32_t value) {
  return guard([&] {
    check_model(m);
    if (!key) fail(MR_ERR_INVALID_ARG, "key is null");
    std::lock_guard<std::mutex> g(m->mu);
    std::string k(key);
    if (k == "threads") m->opt_threads = value;
    else if (k == "variant") {
      if (value != -1 && value != 0 && value != 2 && value != 4 && value != 5)
        fail(MR_ERR_INVALID_ARG, "variant %d does not exist (-1 auto, 0 exact, 2 binned, 4 compact, 5 slim)", value);
      m->opt_variant = value;
      m->code_gen = mr_model::next_code_gen();
    }
    else if (k == "ilp") m->opt_ilp = value;
    else if (k == "latency_rows") m->opt_latency_rows = value;
    else if (k == "chunk_kb") {
      m->opt_chunk_kb = value;
      MR_CUDA_CHECK(cudaSetDevice(m->ctx->device));
      MR_CUDA_CHECK(cudaDeviceSynchronize());
      m->repack();
    } else fail(MR_ERR_NOT_FOUND, "unknown option '%s'", key);
  });
}

mr_status mr_model_close(mr_model *m) {
  return guard([&] {
    if (!m) fail(MR_ERR_INVALID_ARG, "model handle is null");
    bool was = m->closed.exchange(true);
    if (was) return;  // idempotent
    // in-flight predicts hold `inflight`; wait for them, then drop device memory
    while (m->inflight.load() > 0) std::this_thread::yield();
    cudaSetDevice(m->ctx->device);
    cudaDeviceSynchronize();
    m->release_device();
    m->ctx->live_models--;
  });
}

int32_t mr_model_is_closed(mr_model *m) { return (!m || m->closed.load()) ? 1 : 0; }

mr_status mr_model_free(mr_model *m) {
  if (!m) return MR_OK;
  mr_status s = mr_model_close(m);
  delete m;
  return s;
}

mr_status mr_rank_order(mr_ctx *ctx, const double *scores, const int32_t *offsets, int32_t n_requests,
                        int32_t *order) {
  return guard([&] {
    (void)ctx;
    if (n_requests < 0) fail(MR_ERR_INVALID_ARG, "negative request count");
    if (n_requests > 0 && (!scores || !offsets || !order)) fail(MR_ERR_INVALID_ARG, "null argument");
    for (int r = 0; r < n_requests; r++) {
      const int b = offsets[r], e = offsets[r + 1];
      if (e < b) fail(MR_ERR_INVALID_ARG, "offsets must be non-decreasing");
      int32_t *o = order + b;
      for (int i = 0; i < e - b; i++) o[i] = i;
      const double *s = scores + b;
      // java.lang.Double.compare on -score: total order with NaN last, -0.0 < 0.0
      auto key = [](double x) -> int64_t {
        x = -x;
        if (x != x) return INT64_MAX;
        int64_t bits;
        memcpy(&bits, &x, 8);
        return bits < 0 ? (bits ^ INT64_MAX) : bits;
      };
      std::stable_sort(o, o + (e - b), [&](int32_t a, int32_t c) { return key(s[a]) < key(s[c]); });
    }
  });
}

}  // extern "C"
