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
};

// code of x in a column whose header (M, categorical flag) the caller already holds in registers
__device__ __forceinline__ uint16_t code_of_col(const BinParams &p, const BinMeta &M, bool cat, double x) {
  if (p.xgb) x = (double)__double2float_rn(x);
  if (x != x) return kBinNaN;
  if (cat) {
    // LightGBM CategoricalDecision: static_cast<int>(x), negative / out-of-int-range -> right
    const bool in_range = (x < 2147483648.0) && (x > -2147483649.0);
    const int iv = in_range ? __double2int_rz(x) : -1;
    return (iv >= 0 && iv < 65000) ? (uint16_t)iv : kBinNaN;
  }
  // bucket(x) is monotone, so only the thresholds in x's own bucket need comparing
  uint32_t bk = 0;
  if (M.g > 1 && x > M.mn) {
    const double v = __dmul_rn(__dadd_rn(x, -M.mn), M.scale);
    bk = v >= (double)(M.g - 1) ? M.g - 1 : (uint32_t)v;
  }
  const uint32_t range = __ldg(p.bucket_range + M.idx_off + bk);
  uint32_t lo = range & 0xFFFFu, hi = range >> 16;
  const double *t = p.thr + M.thr_off;
  if (p.xgb) {
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (__ldg(t + m) <= x) lo = m + 1; else hi = m; }  // #{t <= x}
  } else {
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (__ldg(t + m) < x) lo = m + 1; else hi = m; }   // #{t < x}
  }
  return (uint16_t)lo;
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
__device__ __forceinline__ uint32_t dup_col(const BinMeta &M) { return M.flags >> 16; }  // kMetaNoDup: none
__device__ __forceinline__ uint16_t dup_code(uint16_t c) { return c == kBinNaN ? (uint16_t)0 : c; }

}  // namespace mr
