// Exact value -> u16 rank-code conversion shared by bin_kernel (gbdt_binned.cu) and the fused
// assemble kernel (assemble_kernels.cu).  See gbdt_model.h "binned layout".
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "gbdt_model.h"

namespace mr {

struct BinParams {
  const double *values;
  const uint32_t *thr_off;
  const double *thr;
  const uint8_t *is_cat;
  const BinMeta *meta;           // [n_features] monotone bucket index (see gbdt_model.h); may point to shared memory
  const uint32_t *bucket_range;
  uint16_t *bins;
  int rows, cols, n_features;
  int tile_cols;  // columns of the code tile (>= n_features, see BinMeta::flags)
  int xgb;  // 1: round to binary32 first, strict less (upper_bound)
  // Layout of `bins`.  0: [group of 32 items][tile column][lane] u16 (8-byte compact scorer, latency path).
  // 512 | 256 | 128 = items per CTA of the slim scorer: [CTA tile][column pair][item of the tile] u32, the pair's even
  // column in the low half — a CTA's tile is still ONE contiguous range, and a lane's read of any column hits its own bank.
  int tile_T;
};

// index (in u16 units) of (item, tile column) in a code buffer of the given layout
__device__ __forceinline__ size_t code_index(int tile_T, int tile_cols, int item, int col) {
  if (tile_T == 0) return ((size_t)(item >> 5) * tile_cols + col) * 32 + (item & 31);
  const int n_pairs = (tile_cols + 1) >> 1;
  return (((size_t)(item / tile_T) * n_pairs + (col >> 1)) * tile_T + (item % tile_T)) * 2 + (col & 1);
}

// code of x in a column whose header (M, categorical flag) the caller already holds in registers
template <bool XGB>
__device__ __forceinline__ uint16_t code_of_col_t(const BinParams &p, const BinMeta &M, bool cat, double x) {
  if (XGB) x = (double)__double2float_rn(x);  // XGBoost compares binary32 values, strict less (upper_bound)
  if (cat) {
    // LightGBM CategoricalDecision: static_cast<int>(x), NaN / negative / out-of-int-range -> right
    const bool in_range = (x < 2147483648.0) && (x > -2147483649.0);  // false for NaN
    const int iv = in_range ? __double2int_rz(x) : -1;
    if (M.flags & kMetaCat16) return (iv >= 0 && iv < 16) ? (uint16_t)(kCat16Base | (uint32_t)iv) : kCat16Missing;
    return (iv >= 0 && iv < 65000) ? (uint16_t)iv : kBinNaN;
  }
  if (x != x) return kBinNaN;
  // bucket(x) is monotone, so only the thresholds in x's own bucket need comparing
  uint32_t bk = 0;
  if (M.g > 1 && x > M.mn) {
    const double v = __dmul_rn(__dadd_rn(x, -M.mn), M.scale);
    bk = v >= (double)(M.g - 1) ? M.g - 1 : (uint32_t)v;
  }
  const uint32_t range = __ldg(p.bucket_range + M.idx_off + bk);
  uint32_t lo = range & 0xFFFFu, hi = range >> 16;
  const double *t = p.thr + M.thr_off;
  const uint32_t n = hi - lo;
  if (n <= 4) {
    // the usual case: a handful of thresholds share the bucket.  Fetch them side by side (one memory
    // latency instead of a chain of dependent probes) and count: the thresholds are sorted, so
    // #{t < x} (LightGBM) or #{t <= x} (XGBoost) among them is x's rank inside the bucket.
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    const double t0 = n > 0 ? __ldg(t + lo) : inf, t1 = n > 1 ? __ldg(t + lo + 1) : inf;
    const double t2 = n > 2 ? __ldg(t + lo + 2) : inf, t3 = n > 3 ? __ldg(t + lo + 3) : inf;
    if (XGB) lo += (uint32_t)(n > 0 && t0 <= x) + (uint32_t)(n > 1 && t1 <= x) + (uint32_t)(n > 2 && t2 <= x) + (uint32_t)(n > 3 && t3 <= x);
    else lo += (uint32_t)(n > 0 && t0 < x) + (uint32_t)(n > 1 && t1 < x) + (uint32_t)(n > 2 && t2 < x) + (uint32_t)(n > 3 && t3 < x);
    return (uint16_t)lo;
  }
  if (XGB) {
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (__ldg(t + m) <= x) lo = m + 1; else hi = m; }  // #{t <= x}
  } else {
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (__ldg(t + m) < x) lo = m + 1; else hi = m; }   // #{t < x}
  }
  return (uint16_t)lo;
}

__device__ __forceinline__ uint16_t code_of_col(const BinParams &p, const BinMeta &M, bool cat, double x) {
  return p.xgb ? code_of_col_t<true>(p, M, cat, x) : code_of_col_t<false>(p, M, cat, x);
}

__device__ __forceinline__ uint16_t code_of(const BinParams &p, int f, double x) {
  const BinMeta M = p.meta[f];
  return code_of_col(p, M, (M.flags & kMetaCat) != 0, x);
}

// Where a code lands in the scorer's tile (BinMeta::flags): the base column takes the NaN code the
// feature's nodes agree on; a feature with both NaN directions also fills its duplicate column.
__device__ __forceinline__ uint16_t base_code(const BinMeta &M, uint16_t c) {
  return (c == kBinNaN && (M.flags & kMetaNanLow)) ? (uint16_t)0 : c;
}
// the category behind a categorical column's code (kBinNaN: goes right), whichever form the model's codes are in
__device__ __forceinline__ uint32_t cat_of_code(uint32_t code, bool cat16) {
  if (!cat16) return code;
  return (code & 0xFFF0u) == kCat16Base ? (code & 15u) : (uint32_t)kBinNaN;
}
__device__ __forceinline__ uint32_t dup_col(const BinMeta &M) { return M.flags >> 16; }  // kMetaNoDup: none
__device__ __forceinline__ uint16_t dup_code(uint16_t c) { return c == kBinNaN ? (uint16_t)0 : c; }

}  // namespace mr
