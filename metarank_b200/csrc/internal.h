// Internal definitions shared by the C-ABI translation units (api.cu, rank_api.cu).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.h"
#include "gbdt_kernels.cuh"
#include "gbdt_model.h"

namespace mr {
extern thread_local std::string t_last_error;

template <class F> mr_status guard(F &&f) {
  try {
    f();
    return MR_OK;
  } catch (const Error &e) {
    t_last_error = e.what();
    return e.code;
  } catch (const std::bad_alloc &) {
    t_last_error = "out of host memory";
    return MR_ERR_INVALID_ARG;
  } catch (const std::exception &e) {
    t_last_error = e.what();
    return MR_ERR_INVALID_ARG;
  } catch (...) {
    t_last_error = "unknown error";
    return MR_ERR_INVALID_ARG;
  }
}
}  // namespace mr

using namespace mr;

// A Lane is what one in-flight host call needs: a stream, pinned staging and device
// scratch.  Lanes are checked out from the context's pool, so concurrent JVM threads
// never share buffers and never allocate on the hot path once warmed up.
struct Lane {
  cudaStream_t stream = nullptr;
  uint8_t *h_pinned = nullptr;
  size_t h_bytes = 0;
  uint8_t *d_buf = nullptr;
  size_t d_bytes = 0;
  void ensure(size_t hb, size_t db) {
    if (hb > h_bytes) {
      if (h_pinned) cudaFreeHost(h_pinned);
      h_pinned = nullptr;
      h_bytes = 0;
      size_t n = std::max(hb, h_bytes * 2);
      MR_CUDA_CHECK(cudaMallocHost((void **)&h_pinned, n));
      h_bytes = n;
    }
    if (db > d_bytes) {
      if (d_buf) cudaFree(d_buf);
      d_buf = nullptr;
      d_bytes = 0;
      size_t n = std::max(db, d_bytes * 2);
      MR_CUDA_CHECK(cudaMalloc((void **)&d_buf, n));
      d_bytes = n;
    }
  }
};

struct mr_ctx {
  int device = 0;
  int num_sms = 0;
  int dyn_smem_base = -1;  // absolute shared-memory address of a kernel's dynamic window (probed by mr_init); the slim scorer
                           // places its code tile at an absolute address and needs the window to start below it
  std::mutex mu;
  std::vector<std::unique_ptr<Lane>> free_lanes;
  std::atomic<int> live_models{0};

  std::unique_ptr<Lane> checkout() {
    {
      std::lock_guard<std::mutex> g(mu);
      if (!free_lanes.empty()) {
        auto l = std::move(free_lanes.back());
        free_lanes.pop_back();
        return l;
      }
    }
    auto l = std::make_unique<Lane>();
    MR_CUDA_CHECK(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
    return l;
  }
  void checkin(std::unique_ptr<Lane> l) {
    std::lock_guard<std::mutex> g(mu);
    free_lanes.push_back(std::move(l));
  }
};

struct LaneGuard {
  mr_ctx *ctx;
  std::unique_ptr<Lane> lane;
  explicit LaneGuard(mr_ctx *c) : ctx(c), lane(c->checkout()) {}
  ~LaneGuard() {
    if (lane) ctx->checkin(std::move(lane));
  }
  Lane *operator->() { return lane.get(); }
};

struct mr_model {
  mr_ctx *ctx = nullptr;
  HostModel host;
  PackedModel packed;
  uint8_t *d_model = nullptr;
  ChunkDesc *d_chunks = nullptr;
  // binned form (exact integer traversal); binned.ok == false -> always the f64/f32 kernel
  BinnedModel binned, compact, lat;  // lat: compact layout in 4 KB chunks for the low-latency path
  SlimModel slim;                    // 4-byte-node form of `compact` (same tile mapping) for the throughput scorer's fast path
  uint8_t *d_smodel = nullptr;
  ChunkDesc *d_schunks = nullptr;
  uint8_t *d_bmodel = nullptr, *d_cmodel = nullptr, *d_lmodel = nullptr;
  ChunkDesc *d_bchunks = nullptr, *d_cchunks = nullptr, *d_lchunks = nullptr;
  uint32_t *d_thr_off = nullptr;
  double *d_thr = nullptr;
  uint8_t *d_is_cat = nullptr;
  BinMeta *d_meta = nullptr, *d_cmeta = nullptr;  // identity tile mapping (binned/threaded) / compact + lat mapping
  uint32_t *d_bucket_range = nullptr;
  uint32_t *d_ltree_off = nullptr;  // sum-kernel staging plan of the latency path: group_rel table
  uint4 *d_lgroups = nullptr;
  SumPlan sum_plan;
  std::atomic<bool> closed{false};
  std::atomic<int> inflight{0};
  std::mutex mu;  // guards repacking / device buffers
  int opt_threads = 0, opt_variant = -1, opt_ilp = 0, opt_chunk_kb = 0, opt_latency_rows = 0;

  void upload() {
    if (d_model) cudaFree(d_model);
    if (d_chunks) cudaFree(d_chunks);
    d_model = nullptr;
    d_chunks = nullptr;
    MR_CUDA_CHECK(cudaMalloc((void **)&d_model, packed.bytes.size()));
    MR_CUDA_CHECK(cudaMemcpy(d_model, packed.bytes.data(), packed.bytes.size(), cudaMemcpyHostToDevice));
    MR_CUDA_CHECK(cudaMalloc((void **)&d_chunks, packed.chunks.size() * sizeof(ChunkDesc)));
    MR_CUDA_CHECK(cudaMemcpy(d_chunks, packed.chunks.data(), packed.chunks.size() * sizeof(ChunkDesc),
                             cudaMemcpyHostToDevice));
  }
  template <class T> static T *to_device(const std::vector<T> &v) {
    T *d = nullptr;
    MR_CUDA_CHECK(cudaMalloc((void **)&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) MR_CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
  }
  void free_binned() {
    for (void *p : {(void *)d_bmodel, (void *)d_bchunks, (void *)d_thr_off, (void *)d_thr, (void *)d_is_cat,
                    (void *)d_cmodel, (void *)d_cchunks, (void *)d_lmodel,
                    (void *)d_lchunks, (void *)d_ltree_off, (void *)d_lgroups, (void *)d_smodel, (void *)d_schunks, (void *)d_meta, (void *)d_cmeta, (void *)d_bucket_range})
      if (p) cudaFree(p);
    d_bmodel = nullptr; d_bchunks = nullptr; d_thr_off = nullptr; d_thr = nullptr; d_is_cat = nullptr;
    d_cmodel = nullptr; d_cchunks = nullptr; d_lmodel = nullptr; d_lchunks = nullptr; d_ltree_off = nullptr; d_lgroups = nullptr; d_smodel = nullptr; d_schunks = nullptr; d_meta = nullptr; d_cmeta = nullptr; d_bucket_range = nullptr;
  }
  // identity of the current code mapping (thresholds + tile columns): consumers that cache codes key on it
  uint64_t code_gen = 0;
  static uint64_t next_code_gen() { static std::atomic<uint64_t> g{1}; return g.fetch_add(1); }
  void repack() {
    code_gen = next_code_gen();
    // Default policy: stream the ensemble through two small shared-memory buffers (TMA bulk
    // copies overlap the traversal).  Small chunks leave shared memory for the feature tile,
    // i.e. for resident warps, which is what bounds this kernel (profiles/ round-1 notes).
    size_t budget;
    if (opt_chunk_kb > 0) {
      budget = (size_t)opt_chunk_kb * 1024;
    } else {
      // measured (profiles/sweep_r1.md): narrow rows leave room for large chunks (fewer CTA-wide
      // barriers); wide rows need the shared memory for the code tile (occupancy)
      budget = (host.n_features <= 32 ? 60 : 16) * 1024;
    }
    packed = pack_model(host, budget);
    upload();
    free_binned();
    binned = pack_binned(host, budget);
    if (binned.ok) {
      d_bmodel = to_device(binned.packed.bytes);
      d_bchunks = to_device(binned.packed.chunks);
      d_thr_off = to_device(binned.thr_off);
      d_thr = to_device(binned.thr);
      d_is_cat = to_device(binned.is_cat);
      d_meta = to_device(binned.meta);
      d_bucket_range = to_device(binned.bucket_range);
      compact = pack_compact(host, binned, opt_chunk_kb > 0 ? budget : 0);  // 0: sized from the code tile
      if (compact.ok) {
        d_cmodel = to_device(compact.packed.bytes);
        d_cchunks = to_device(compact.packed.chunks);
        d_cmeta = to_device(compact.meta);
      }
      slim = SlimModel{};
      if (compact.ok) {
        const uint32_t min_tile = (uint32_t)std::max(0, ctx->dyn_smem_base) + 64u;  // (no slim scoring at all without a probed base)
        slim = pack_slim(host, compact, opt_chunk_kb > 0 ? budget : 0, 512, 48, min_tile);
        if (slim.ok) {
          d_smodel = to_device(slim.packed.bytes);
          d_schunks = to_device(slim.packed.chunks);
        }
      }
      sum_plan = SumPlan{};
      lat = pack_compact(host, binned, 4 * 1024);
      if (lat.ok && !compact.ok) lat.ok = false;  // same tile mapping as `compact` (d_cmeta) by construction
      if (lat.ok) {
        d_lmodel = to_device(lat.packed.bytes);
        d_lchunks = to_device(lat.packed.chunks);
        // groups of consecutive chunks (one contiguous byte range each) for the in-order sum's staging
        std::vector<uint4> groups;
        std::vector<uint32_t> rel;
        const auto &ch = lat.packed.chunks;
        for (size_t c = 0; c < ch.size();) {
          size_t e = c + 1;
          uint32_t bytes = ch[c].bytes, trees = ch[c].n_trees;
          while (e < ch.size() && bytes + ch[e].bytes <= kSumGroupBytes && trees + ch[e].n_trees <= (uint32_t)kSumGroupTrees &&
                 ch[e].byte_off == ch[e - 1].byte_off + ch[e - 1].bytes) {
            bytes += ch[e].bytes; trees += ch[e].n_trees; e++;
          }
          if (trees > (uint32_t)kSumGroupTrees) { lat.ok = false; break; }  // a chunk of > 64 single-leaf trees: not worth a path
          groups.push_back(make_uint4(ch[c].first_tree, trees, ch[c].byte_off, bytes));
          rel.resize(groups.size() * kSumGroupTrees, 0);
          uint32_t k = 0;
          for (size_t q = c; q < e; q++)
            for (uint32_t t = 0; t < ch[q].n_trees; t++) rel[(groups.size() - 1) * kSumGroupTrees + k++] = ch[q].byte_off - ch[c].byte_off;
          sum_plan.max_group_bytes = std::max(sum_plan.max_group_bytes, bytes);
          c = e;
        }
        if (lat.ok) {
          d_lgroups = to_device(groups);
          d_ltree_off = to_device(rel);
          sum_plan.d_groups = d_lgroups;
          sum_plan.d_group_rel = d_ltree_off;
          sum_plan.n_groups = (int)groups.size();
        }
      }
    }
  }
  bool use_compact() const { return compact.ok && (opt_variant == 4 || opt_variant == 5 || opt_variant < 0); }
  // Layout of the code buffer a batch of `rows` rows is scored from (BinParams::tile_T): the slim scorer's CTA-tile layout
  // when the model has a slim form, the batch is not one for the tree-parallel latency path, and the device's dynamic
  // shared window starts below the tile's absolute address; else 0 (groups of 32 rows: compact kernel / latency path).
  // variant 4 pins the 8-byte compact kernel, variant 5 the slim one (tests, A/B).
  // Layout of the code buffer a batch of `rows` rows is scored from (BinParams::tile_T): the slim scorer's CTA-tile layout
  // when the model has a slim form, the batch is not one for the tree-parallel latency path, and the device's dynamic
  // shared window starts below the tile's absolute address; else 0 (groups of 32 rows: compact kernel / latency path).
  // variant 4 pins the 8-byte compact kernel, variant 5 the slim one (tests, A/B).
  // (Smaller-tile forms of the same model for batches that fit the chip in one wave were measured and dropped: a
  // 256 000-row batch as 1000 x 256-item tiles, 7 CTAs per SM, took the 319 us the 500 x 512-item tiles take in two
  // rounds — profiles/ncu_r2_summary.md section 1b.)
  int code_layout(int rows) const {
    if (!slim.ok || !use_binned() || !use_compact() || opt_variant == 4 || opt_threads != 0) return 0;
    if (ctx->dyn_smem_base < 0 || ctx->dyn_smem_base + 64 > slim.col_base * slim.tile_T * 4) return 0;
    if (opt_variant < 0 && use_latency(rows)) return 0;
    return slim.tile_T;
  }
  // the code-based scorer that is active and the width of its code tile
  const BinnedModel &active_binned() const { return use_compact() ? compact : binned; }
  int code_cols() const { return active_binned().tile_cols; }
  bool use_binned() const {
    if ((size_t)2 * host.n_features * (4 * 32 + 2) * sizeof(uint16_t) > 200 * 1024) return false;  // bin_kernel tile
    if (binned.ok && 2 * (size_t)binned.packed.max_chunk_bytes + 64 * (size_t)2 * host.n_features * 2 > 200 * 1024)
      return false;  // an enormous tree: the exact kernel's HBM-resident slow path handles it
    // auto (-1): the compact binned kernel whenever the model allows it (no categorical splits, <= 1023
    // features); otherwise generic binned for LightGBM (f64 -> u16 quarters the tile) and the plain f32
    // kernel for XGBoost, whose features are already binary32 (profiles/sweep_r1.md)
    if (opt_variant < 0) return binned.ok && (compact.ok || host.kind == MR_BOOSTER_LIGHTGBM);
    return binned.ok && (opt_variant == 2 || opt_variant == 4);
  }
  BinnedLaunch binned_desc() const {
    BinnedLaunch B;
    const bool cmp = use_compact();
    const BinnedModel &M = cmp ? compact : binned;
    B.compact = cmp;
    B.d_model = cmp ? d_cmodel : d_bmodel;
    B.d_chunks = cmp ? d_cchunks : d_bchunks;
    B.n_chunks = (int)M.packed.chunks.size();
    B.max_chunk_bytes = M.packed.max_chunk_bytes;
    B.d_thr_off = d_thr_off; B.d_thr = d_thr; B.d_is_cat = d_is_cat;
    B.d_meta = cmp ? d_cmeta : d_meta; B.d_bucket_range = d_bucket_range;
    B.tile_cols = M.tile_cols;
    B.kind = host.kind; B.has_cat = host.has_cat; B.cat16 = M.cat16; B.base_score = host.base_score;
    B.n_features = host.n_features;
    B.threads = opt_threads; B.ilp = opt_ilp;
    return B;
  }
  // Scores rows whose u16 codes were already written to d_codes (fused assemble path).
  // The tree-parallel path wins while the batch cannot fill the chip with one thread per item: a thread's walk
  // through T trees is a T * depth long dependent chain whatever the batch size, the latency path's cost is the
  // leaf-value round trip (T * rows * 16 B).  Measured crossover: profiles/latency_path_r2.md.
  int latency_max_rows() const {
    if (opt_latency_rows > 0) return opt_latency_rows;
    const long long t = std::max<long long>(1, (long long)host.trees.size());
    return (int)std::max<long long>(kLatencyMaxRows, std::min<long long>(32768, 48000000ll / t));
  }
  bool use_latency(int rows) const {
    return lat.ok && rows <= latency_max_rows() && opt_variant < 0 && opt_threads == 0 &&
           use_compact() && 128 + 2 * ((size_t)lat.packed.max_chunk_bytes + 128) + (size_t)4 * compact.tile_cols * 64 <= 200 * 1024 &&
           128 + 3 * ((size_t)sum_plan.max_group_bytes + 128 + (size_t)kSumGroupTrees * 128 * 2 + kSumGroupTrees * 4) <= 220 * 1024;
  }
  // true: score_codes() can store to peer sinks from inside the scoring kernel
  bool fuses_sinks(int rows) const { return use_latency(rows) || (use_binned() && use_compact()); }
  // layout: the code buffer's layout — -1 = code_layout(rows) (what the producers were told for a batch of this size),
  // 0 = groups of 32 rows whatever the size (mega-request slices)
  void score_codes(uint16_t *d_codes, int rows, double *d_out, cudaStream_t stream, void *d_leaf_scratch = nullptr,
                   const ScoreSinks *sinks = nullptr, int layout = -1) const {
    BinnedLaunch B = binned_desc();
    B.rows = rows; B.cols = host.n_features; B.d_out = d_out; B.d_bins = d_codes; B.codes_ready = true;
    if (sinks) B.sinks = *sinks;
    if (layout < 0) layout = code_layout(rows);
    if (layout) {
      B.tile_T = slim.tile_T;
      B.slim_col_base = slim.col_base;
      B.d_model = d_smodel; B.d_chunks = d_schunks;
      B.n_chunks = (int)slim.packed.chunks.size();
      B.max_chunk_bytes = slim.packed.max_chunk_bytes;
      if (!slim.root_tab.empty()) { B.h_root_tab = slim.root_tab.data(); B.n_root_tab = (int)(slim.root_tab.size() / 4); }
      launch_gbdt_binned(B, ctx->num_sms, stream);
      return;
    }
    if (use_latency(rows)) {
      B.d_model = d_lmodel; B.d_chunks = d_lchunks;
      B.n_chunks = (int)lat.packed.chunks.size();
      B.max_chunk_bytes = lat.packed.max_chunk_bytes;
      if (d_leaf_scratch) {
        launch_gbdt_latency(B, (int)host.trees.size(), sum_plan, d_leaf_scratch, stream);
      } else {
        void *lv = nullptr;
        MR_CUDA_CHECK(cudaMallocAsync(&lv, latency_scratch_bytes(rows, (int)host.trees.size()), stream));
        launch_gbdt_latency(B, (int)host.trees.size(), sum_plan, lv, stream);
        MR_CUDA_CHECK(cudaFreeAsync(lv, stream));
      }
      return;
    }
    launch_gbdt_binned(B, ctx->num_sms, stream);
  }
  // Enqueue scoring of a device-resident matrix on `stream` with whichever kernel applies.
  void score(const double *d_values, int rows, int cols, double *d_out, cudaStream_t stream) const {
    if (use_binned()) {
      BinnedLaunch B = binned_desc();
      B.d_values = d_values; B.rows = rows; B.cols = cols; B.d_out = d_out;
      void *bins = nullptr;
      MR_CUDA_CHECK(cudaMallocAsync(&bins, std::max<size_t>(binned_scratch_bytes(rows, code_cols()), 16), stream));
      B.d_bins = (uint16_t *)bins;
      B.tile_T = code_layout(rows);
      if (use_latency(rows) || B.tile_T) {
        B.codes_only = true;
        launch_gbdt_binned(B, ctx->num_sms, stream);
        score_codes(B.d_bins, rows, d_out, stream);
      } else {
        launch_gbdt_binned(B, ctx->num_sms, stream);
      }
      MR_CUDA_CHECK(cudaFreeAsync(bins, stream));
    } else {
      ScoreLaunch L = launch_desc(d_values, rows, cols, d_out);
      launch_gbdt_score(L, ctx->num_sms, stream);
    }
  }
  void release_device() {
    free_binned();
    if (d_model) cudaFree(d_model);
    if (d_chunks) cudaFree(d_chunks);
    d_model = nullptr;
    d_chunks = nullptr;
  }
  ScoreLaunch launch_desc(const double *d_values, int rows, int cols, double *d_out) const {
    ScoreLaunch L;
    L.d_model = d_model;
    L.d_chunks = d_chunks;
    L.n_chunks = (int)packed.chunks.size();
    L.max_chunk_bytes = packed.max_chunk_bytes;
    L.kind = host.kind;
    L.has_cat = host.has_cat;
    L.has_zero = host.has_zero_missing;
    L.base_score = host.base_score;
    L.n_features = host.n_features;
    L.d_values = d_values;
    L.rows = rows;
    L.cols = cols;
    L.d_out = d_out;
    L.threads = opt_threads;
    L.ilp = opt_ilp;
    return L;
  }
};


namespace mr {
struct InflightGuard {
  mr_model *m;
  explicit InflightGuard(mr_model *mm) : m(mm) { m->inflight++; }
  ~InflightGuard() { m->inflight--; }
};
void check_model(mr_model *m);
}  // namespace mr
