// Device-resident feature state: the replacement for Persistence.values
// (KVStore[Key, FeatureValue], reference S/fstore/Persistence.scala:39,85-89) on the read
// path.  One table per scope kind; a table = open-addressing hash map (scope hash -> row)
// + fixed-width rows of u64 words (presence bits, then every slot of that scope) + a pool
// for variable-length payloads (tag hashes, bounded lists) + dense side arrays for
// embeddings.  The host keeps a shadow copy and uploads dirty ranges on flush.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <deque>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <unordered_map>
#include <vector>

#include "schema.h"

namespace mr {

constexpr int kMaxSides = 4;

struct DTable {  // device view, passed to kernels by value
  const uint64_t *keys;
  const uint32_t *vals;
  uint32_t mask;       // capacity - 1 (capacity is a power of two); 0 when empty
  uint32_t n_rows;
  const uint64_t *rows;
  const uint64_t *pool;
  int32_t row_words;
  int32_t pad;
};

struct DState {
  DTable t[SC_N_TABLES];
  // Embedding side arrays (one row of `dim` values per table row).  An array whose every element survives the
  // round trip through binary32 — true for anything an ONNX encoder produced: the reference only ever widens f32
  // outputs, S/model/Scalar.scala:16-33 — lives on the device as f32 (`side_f32`, half the bytes the cosine kernel
  // streams) and is widened in registers, bit for bit the stored double; otherwise as f64 (`side`).  Exactly one of
  // the two pointers is non-null.  side_bs[row] = sum of e*e in index order (CosineDistance's bSum, a function of the
  // item alone, computed once at upsert time).
  const double *side[kMaxSides];
  const float *side_f32[kMaxSides];
  const double *side_bs[kMaxSides];
  int32_t side_dim[kMaxSides];
};

__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
  return x;
}

struct HostTable {
  std::vector<uint64_t> keys;   // 0 = empty
  std::vector<uint32_t> vals;
  size_t n_rows = 0;
  int row_words = 1;
  std::vector<uint64_t> rows;
  std::vector<uint64_t> pool;
  std::vector<std::vector<double>> sides;  // one per side array of this table
  std::vector<int> side_ids;
  // device mirrors
  uint64_t *d_keys = nullptr; uint32_t *d_vals = nullptr; size_t d_cap = 0;
  uint64_t *d_rows = nullptr; size_t d_rows_cap = 0;
  uint64_t *d_pool = nullptr; size_t d_pool_cap = 0, pool_uploaded = 0;
  std::vector<double *> d_sides; std::vector<size_t> d_sides_cap;
  std::vector<std::vector<double>> side_bs;      // per row: sum of squares of the embedding, in index order
  std::vector<uint8_t> side_is_f32;              // every element stored so far round-trips through binary32
  std::vector<uint8_t> side_dev_mode;            // what the device holds: 0 nothing yet, 1 f32, 2 f64
  std::vector<float *> d_sides_f32; std::vector<size_t> d_sides_f32_cap;
  std::vector<double *> d_side_bs; std::vector<size_t> d_side_bs_cap;
  size_t d_keys_n = 0, d_n_rows = 0;  // what the device holds as of the last flush (ranking reads this snapshot)
  bool map_dirty = false;
  size_t dirty_lo = SIZE_MAX, dirty_hi = 0;  // row range touched since the last flush
  size_t pool_dirty_lo = SIZE_MAX, pool_dirty_hi = 0;  // pool entries rewritten in place (bounded lists)

  uint32_t find(uint64_t key) const;
  uint32_t find_or_insert(uint64_t key);
  void grow_map();
  std::vector<uint8_t> row_dirty;     // per row: touched since the last flush
  std::vector<uint32_t> dirty_rows;   // the touched rows (each once), for the scatter upload
  void touch(size_t row) {
    dirty_lo = std::min(dirty_lo, row);
    dirty_hi = std::max(dirty_hi, row + 1);
    if (row_dirty.size() <= row) row_dirty.resize(std::max(row + 1, row_dirty.size() * 2), 0);
    if (!row_dirty[row]) { row_dirty[row] = 1; dirty_rows.push_back((uint32_t)row); }
  }
};

struct StateStore {
  // resolves (slot, scope ids) to the row that holds it, inserting the row when new
  uint32_t row_for(const Slot &sl, int scope, uint64_t id0, uint64_t id1);
  Schema schema;
  HostTable tables[SC_N_TABLES];
  std::shared_mutex mu;  // shared: kernels reading the tables; exclusive: upsert / flush
  int device = 0;
  int64_t device_bytes = 0;

  explicit StateStore(const Schema &s);
  ~StateStore();
  void upsert(const uint8_t *p, size_t len, int64_t *applied, int64_t *skipped);
  // Write path (FeatureValueFlow.commitWrite + computeValue with refresh = always):
  // Put / Increment / PeriodicIncrement / Append with the reference's Mem* state semantics.
  void apply_writes(const uint8_t *p, size_t len, int64_t *applied, int64_t *skipped);
  // raw state behind the refreshed values, keyed by (table, row, slot)
  struct RawKey {
    uint64_t v;
    bool operator==(const RawKey &o) const { return v == o.v; }
  };
  struct RawKeyHash { size_t operator()(const RawKey &k) const { return (size_t)mix64(k.v); } };
  std::unordered_map<RawKey, std::map<int64_t, int64_t>, RawKeyHash> buckets;                // MemPeriodicCounter
  std::unordered_map<RawKey, std::deque<std::pair<int64_t, uint64_t>>, RawKeyHash> lists;     // MemBoundedList
  std::unordered_map<RawKey, uint32_t, RawKeyHash> list_region;  // pool offset of a list's fixed-capacity region
  std::unordered_map<RawKey, std::pair<uint32_t, uint32_t>, RawKeyHash> upsert_region;  // upserted lists: {pool offset, capacity}
  void flush();
  DState view() const;
  // Item-table change log for consumers that keep derived per-row data on the device (the per-model
  // code rows, rank_api.cu): one entry per flush that touched item rows.
  struct ItemChange {
    uint64_t epoch;
    bool all;                    // too many rows (or rows moved): rebuild everything
    std::vector<uint32_t> rows;  // otherwise the rows that changed
  };
  uint64_t item_epoch = 0;
  std::deque<ItemChange> item_log;  // last kItemLogMax flushes
  static constexpr size_t kItemLogMax = 64;
};

}  // namespace mr
