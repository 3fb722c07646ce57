// Launch interface of the GBDT scoring kernels (gbdt_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "gbdt_model.h"

namespace mr {

struct ScoreLaunch {
  // model
  const uint8_t *d_model = nullptr;       // packed chunks in HBM
  const ChunkDesc *d_chunks = nullptr;    // chunk table in HBM
  int n_chunks = 0;
  uint32_t max_chunk_bytes = 0;
  int kind = 0;                           // MR_BOOSTER_*
  bool has_cat = false, has_zero = false;
  float base_score = 0.f;
  int n_features = 0;
  // batch
  const double *d_values = nullptr;       // row-major rows x cols
  int rows = 0, cols = 0;
  double *d_out = nullptr;
  unsigned long long *d_visited = nullptr;  // non-null: also count evaluated internal nodes
  // tuning (0 = choose)
  int threads = 0;
  int ilp = 0;
};

// Enqueues the scoring kernel on `stream`.  Throws mr::Error on failure.
void launch_gbdt_score(const ScoreLaunch &L, int num_sms, cudaStream_t stream);

// Where the finished scores go besides d_out.  The mega-request path (rank_api.cu, mr_group_rank) points `peer`
// at the score buffers of every GPU of the group (peer memory over NVLink): the scorer's final store IS the
// all-gather.  The CTA that finishes last stores `seq` to flag[g] on every peer (release, system scope).
struct ScoreSinks {
  double *peer[8] = {};
  uint32_t *flag[8] = {};
  uint32_t *done = nullptr;  // device counter of finished CTAs (zero between launches)
  uint32_t seq = 0;
  int n_peer = 0;            // 0: only d_out
  int item_base = 0;         // index of this slice's first item in the peers' buffers
};

// Binned path (gbdt_binned.cu): feature values -> u16 rank codes, then integer traversal.
struct BinnedLaunch {
  const uint8_t *d_model = nullptr;     // chunks of BNodes
  const ChunkDesc *d_chunks = nullptr;
  int n_chunks = 0;
  uint32_t max_chunk_bytes = 0;
  const uint32_t *d_thr_off = nullptr;  // [F + 1]
  const double *d_thr = nullptr;
  const uint8_t *d_is_cat = nullptr;    // [F]
  const BinMeta *d_meta = nullptr;      // [F]
  const uint32_t *d_bucket_range = nullptr;
  int kind = 0;
  bool has_cat = false;
  bool cat16 = false;                   // categorical columns carry the small-categorical code form (kMetaCat16)
  float base_score = 0.f;
  int n_features = 0;
  int tile_cols = 0;                    // columns of the code tile (BinnedModel::tile_cols)
  const double *d_values = nullptr;
  int rows = 0, cols = 0;
  double *d_out = nullptr;
  uint16_t *d_bins = nullptr;           // scratch: ceil(rows/32) * tile_cols * 32 codes
  int threads = 0, ilp = 0;
  bool codes_only = false;              // run bin_kernel only (no traversal)
  bool codes_ready = false;             // d_bins already holds the codes (fused assemble): skip bin_kernel
  bool compact = false;                 // model bytes are pack_compact() chunks -> fast lock-step kernel
  int tile_T = 0;                       // != 0: d_bins is in the slim layout and d_model / d_chunks are pack_slim() chunks
  const uint32_t *h_root_tab = nullptr;  // slim scorer: HOST copy of SlimModel::root_tab (4 words per tree), passed as a kernel parameter
  int n_root_tab = 0;                   // trees in it (0: the chunks' own root tables)
  int slim_col_base = 1;                // slim scorer: SlimModel::col_base of the form in d_model
  ScoreSinks sinks;                     // compact + latency kernels only
};
// bytes of a code buffer for `rows` rows in either layout (BinParams::tile_T): whole groups of 32 / whole CTA tiles
inline size_t binned_scratch_bytes(int rows, int tile_cols) {
  const size_t legacy = (size_t)((rows + 31) / 32) * (size_t)tile_cols * 32 * sizeof(uint16_t);
  const size_t slim = (size_t)((rows + 511) / 512) * 512 * (size_t)((tile_cols + 1) / 2) * 4;
  return legacy > slim ? legacy : slim;
}
void launch_gbdt_binned(const BinnedLaunch &L, int num_sms, cudaStream_t stream);
// Low-latency path for small batches: the leaf every row reaches in every tree, spread over (chunk x item-group)
// CTAs, then an in-order per-item sum of the values behind them.  d_leafslots: latency_scratch_bytes() of scratch.
// SumPlan: the small-chunk packing cut into groups of consecutive chunks for the sum kernel's staging
// (groups[g] = {first tree, n trees, byte offset, bytes}; group_rel[g * kSumGroupTrees + k] = offset of tree k's chunk in its group).
constexpr int kSumGroupTrees = 64;
constexpr uint32_t kSumGroupBytes = 16 * 1024;
struct SumPlan {
  const uint4 *d_groups = nullptr;
  const uint32_t *d_group_rel = nullptr;
  int n_groups = 0;
  uint32_t max_group_bytes = 0;
};
void launch_gbdt_latency(const BinnedLaunch &L, int n_trees, const SumPlan &sum, void *d_leafslots, cudaStream_t stream);
// Level steps the lock-step scorer executes on the codes in L.d_bins (compact layout only): out = {sum of the lanes'
// path lengths, sum over (warp, tree) of the deepest lane's path, number of (warp, tree) pairs}.  Synchronous.
void compact_walk_stats(const BinnedLaunch &L, unsigned long long out[3], cudaStream_t stream);
inline size_t latency_scratch_bytes(int rows, int n_trees) { return (size_t)n_trees * (size_t)((rows + 127) & ~127) * 2; }
constexpr int kLatencyMaxRows = 2048;

extern long long g_kernel_launches;

// Per-kernel timing for bench.py's roofline block (mr_profile_begin / mr_profile_end): while a profile is open,
// every launch site brackets its kernel with two CUDA events on the launching stream.  Off (one relaxed load per
// launch) on the serving path.
struct ProfScope {
  const char *name;
  cudaStream_t stream;
  cudaEvent_t e0 = nullptr;
  ProfScope(const char *n, cudaStream_t s);
  ~ProfScope();
};
void profile_begin();
bool profile_active();  // launch sites that replay CUDA graphs fall back to plain launches while a profile is open
// Waits for the device, then appends one JSON object per kernel name to `out`: {"kernel", "launches", "ms"}.
void profile_end(std::string &out);

}  // namespace mr
