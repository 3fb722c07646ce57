// Binned GBDT scoring on sm_100a (see gbdt_model.h "binned layout" for why it is exact).
//
//   bin_kernel      X (row-major f64) -> u16 rank codes, laid out [group of 32 items][feature][lane]
//                   so that a CTA's tile is one contiguous byte range;
//   score kernel    thread per item, trees in tree order (bit-identical sums, like the f64 kernel),
//                   per node one 8-byte LDS (node) + one 2-byte LDS (code) + integer compare.
//                   The tile of codes and the tree chunks are both staged by TMA bulk copies
//                   (cp.async.bulk + mbarrier complete_tx); chunks are double-buffered.
// Compared with the f64 kernel this halves the node bytes, quarters the tile bytes (4x more
// items resident per SM -> more warps to hide the LDS->LDS->compare chain) and takes the
// FP64 compare pipe out of the loop.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "binning.cuh"
#include "gbdt_kernels.cuh"
#include "sinks.cuh"
#include "tma.cuh"

namespace mr {
namespace {


// ------------------------------------------------------------------ binning
constexpr int kBinGroups = 4;  // groups of 32 items per CTA

__global__ void __launch_bounds__(256) bin_kernel(const BinParams p) {
  extern __shared__ uint16_t s_codes[];  // [tile_cols][kBinGroups*32 + 2]
  const int F = p.n_features, T = p.tile_cols;
  const int items_per_cta = kBinGroups * 32, pitch = items_per_cta + 2;
  const long long item0 = (long long)blockIdx.x * items_per_cta;
  const int n_here = (int)min((long long)items_per_cta, (long long)p.rows - item0);
  // phase 1: coalesced read of the tile's rows, one element per thread per step
  const int n_elem = items_per_cta * F;
  for (int e = threadIdx.x; e < n_elem; e += blockDim.x) {
    const int it = e / F, f = e - it * F;
    const BinMeta M = p.meta[f];
    uint16_t c = 0;
    if (it < n_here) c = code_of_col(p, M, (M.flags & kMetaCat) != 0, __ldg(p.values + (size_t)(item0 + it) * p.cols + f));
    s_codes[f * pitch + it] = base_code(M, c);
    if (dup_col(M) != kMetaNoDup) s_codes[dup_col(M) * pitch + it] = dup_code(c);
  }
  __syncthreads();
  // phase 2: coalesced write in [group][column][lane] order
  if (p.tile_T == 0) {
    uint16_t *out = p.bins + (size_t)blockIdx.x * kBinGroups * T * 32;
    const long long n_groups_total = (p.rows + 31) / 32;
    for (int o = threadIdx.x; o < items_per_cta * T; o += blockDim.x) {
      const int g = o / (T * 32), r = o - g * (T * 32), f = r >> 5, lane = r & 31;
      if ((long long)blockIdx.x * kBinGroups + g < n_groups_total) out[o] = s_codes[f * pitch + g * 32 + lane];
    }
  } else {
    // slim layout: [CTA tile][column pair][item] u32 — consecutive threads write consecutive items of one pair
    const int n_pairs = (T + 1) >> 1;
    uint32_t *out = reinterpret_cast<uint32_t *>(p.bins);
    for (int o = threadIdx.x; o < items_per_cta * n_pairs; o += blockDim.x) {
      const int pr = o / items_per_cta, it = o - pr * items_per_cta;
      const long long item = item0 + it;
      if (item < (((long long)p.rows + 31) & ~31ll)) {
        const uint32_t lo = s_codes[(2 * pr) * pitch + it], hi = (2 * pr + 1 < T) ? s_codes[(2 * pr + 1) * pitch + it] : 0u;
        out[((size_t)(item / p.tile_T) * n_pairs + pr) * p.tile_T + (size_t)(item % p.tile_T)] = lo | (hi << 16);
      }
    }
  }
}

// ------------------------------------------------------------------ traversal
struct BParams {
  const uint8_t *model;
  const ChunkDesc *chunks;
  const uint16_t *bins;
  double *out;
  int n_chunks;
  uint32_t chunk_stride;
  int rows, n_features;
  float base_score;
  int cat16;  // categorical columns carry the small-categorical code form (gbdt_model.h kMetaCat16)
  int col_base;  // slim scorer: pair p of the code tile lives at absolute shared address (p + col_base) * 4T
  ScoreSinks sinks;
};

template <bool HAS_CAT>
__device__ __forceinline__ int bstep(const uint2 nd, uint32_t code, const uint8_t *chunk, bool cat16) {
  const uint32_t k = nd.x & 0xFFFFu, fl = nd.x >> 28;
  bool left;
  if (HAS_CAT && (fl & BF_CATEGORICAL)) {
    left = false;
    code = cat_of_code(code, cat16);
    if (code != kBinNaN) {
      const uint2 ct = reinterpret_cast<const uint2 *>(chunk)[k];  // {bitset word offset, n words}
      const uint32_t w = code >> 5;
      if (w < ct.y) left = (reinterpret_cast<const uint32_t *>(chunk)[ct.x + w] >> (code & 31u)) & 1u;
    }
  } else {
    // k <= 65000 < kBinNaN, so `code <= k` is already false for NaN; NaN then follows NF_NAN_LEFT
    left = (code <= k) || (code == kBinNaN && (fl & BF_NAN_LEFT));
  }
  return left ? (int)(short)(nd.y & 0xFFFFu) : (int)(short)(nd.y >> 16);
}

template <typename Real, bool HAS_CAT, int ILP>
__global__ void __launch_bounds__(1024) gbdt_score_binned_kernel(const BParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  using AccT = Real;
  const int W = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = p.n_features;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);  // [0],[1] chunk buffers, [2] tile
  const bool resident = p.n_chunks == 1;
  uint8_t *cbuf0 = smem + 128;
  uint8_t *cbuf1 = cbuf0 + (resident ? 0u : p.chunk_stride);
  uint16_t *xs = reinterpret_cast<uint16_t *>(cbuf1 + p.chunk_stride);
  const uint16_t *xw = xs + (size_t)warp * F * 32 + lane;  // this lane's column of codes

  const int n_tiles = (p.rows + W - 1) / W;
  const int groups_per_tile = W >> 5;
  const int n_groups = (p.rows + 31) >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < n_tiles) {
    const ChunkDesc cd = p.chunks[0];
    mbar_arrive_expect_tx(&bars[0], cd.bytes);
    tma_bulk_g2s(cbuf0, p.model + cd.byte_off, cd.bytes, &bars[0]);
  }

  uint32_t it = 0, tile_it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
    // the tile's codes are one contiguous range: one TMA bulk copy
    if (tid == 0) {
      const int g0 = tile * groups_per_tile, g1 = min(n_groups, g0 + groups_per_tile);
      const uint32_t bytes = (uint32_t)(g1 - g0) * (uint32_t)F * 64u;
      fence_proxy_async();
      mbar_arrive_expect_tx(&bars[2], bytes);
      tma_bulk_g2s(xs, p.bins + (size_t)g0 * F * 32, bytes, &bars[2]);
    }
    mbar_wait(&bars[2], tile_it & 1);
    const int item = tile * W + tid;

    AccT acc = (sizeof(Real) == 4) ? (AccT)p.base_score : (AccT)0;
    for (int c = 0; c < p.n_chunks; ++c, ++it) {
      if (!resident && tid == 0) {
        const bool more = (c + 1 < p.n_chunks) || (tile + (int)gridDim.x < n_tiles);
        if (more) {
          const int nc = (c + 1 < p.n_chunks) ? c + 1 : 0;
          const ChunkDesc cd = p.chunks[nc];
          uint64_t *bar = &bars[(it + 1) & 1];
          fence_proxy_async();
          mbar_arrive_expect_tx(bar, cd.bytes);
          tma_bulk_g2s(((it + 1) & 1) ? cbuf1 : cbuf0, p.model + cd.byte_off, cd.bytes, bar);
        }
      }
      if (!resident || it == 0) mbar_wait(&bars[it & 1], (it >> 1) & 1);
      const uint8_t *cb = (!resident && (it & 1)) ? cbuf1 : cbuf0;
      const int ntree = (int)*reinterpret_cast<const uint32_t *>(cb);
      const uint2 *tab = reinterpret_cast<const uint2 *>(cb + 16);

      int t = 0;
      for (; t + ILP <= ntree; t += ILP) {
        const uint2 *nodes[ILP];
        const Real *leaves[ILP];
        int n[ILP];
#pragma unroll
        for (int k = 0; k < ILP; k++) {
          const uint2 to = tab[t + k];
          nodes[k] = reinterpret_cast<const uint2 *>(cb + to.x);
          leaves[k] = reinterpret_cast<const Real *>(cb + to.y);
          n[k] = 0;
        }
        bool any = true;
        while (any) {
          uint2 nd[ILP];
          uint32_t code[ILP];
#pragma unroll
          for (int k = 0; k < ILP; k++) nd[k] = nodes[k][n[k] < 0 ? 0 : n[k]];
#pragma unroll
          for (int k = 0; k < ILP; k++) code[k] = xw[((nd[k].x >> 16) & 0xFFFu) << 5];
          any = false;
#pragma unroll
          for (int k = 0; k < ILP; k++) {
            const int nx = bstep<HAS_CAT>(nd[k], code[k], cb, p.cat16 != 0);
            n[k] = n[k] >= 0 ? nx : n[k];
            any |= n[k] >= 0;
          }
        }
#pragma unroll
        for (int k = 0; k < ILP; k++) acc += (AccT)leaves[k][~n[k]];
      }
      for (; t < ntree; t++) {
        const uint2 to = tab[t];
        const uint2 *nodes = reinterpret_cast<const uint2 *>(cb + to.x);
        const Real *leaves = reinterpret_cast<const Real *>(cb + to.y);
        int n = 0;
        do {
          const uint2 nd = nodes[n];
          n = bstep<HAS_CAT>(nd, xw[((nd.x >> 16) & 0xFFFu) << 5], cb, p.cat16 != 0);
        } while (n >= 0);
        acc += (AccT)leaves[~n];
      }
      __syncthreads();  // chunk buffer (and, after the last chunk, the tile) may be overwritten next
    }
    if (item < p.rows) p.out[item] = (double)acc;
  }
}

// ------------------------------------------------------------------ compact lock-step traversal
// Same mapping as gbdt_score_binned_kernel, on the compact layout (gbdt_model.h): children are byte
// offsets (bit 0 = leaf), the feature's byte offset inside the warp's code tile is stored in the
// node, so one tree level is: LDS.64 node, shift+mask/or, LDS.U16 code, compare, select, test.
// One tree: from the root's byte offset to the leaf's (offset | 1).  One level = 9 SASS instructions
// (ALIGNED) or 10; spelled in PTX because nvcc otherwise routes the predicates through integer registers
// (18 instructions).  With HAS_CAT the PTX loop additionally leaves on a categorical node (bit 1 of
// word0), which is resolved in C++, then re-enters.  Every tree starts at an internal node (single-leaf
// trees are packed as a dummy split).
template <bool HAS_CAT, bool ALIGNED>
__device__ __forceinline__ uint32_t walk_tree(uint32_t n, const uint8_t *cb, uint32_t cb_addr, const uint8_t *xwarp,
                                              uint32_t xwarp_addr, uint32_t lane2, bool cat16 = false) {
  if (HAS_CAT) {
    // pointers carry the kind of their target (bit 0 leaf, bit 1 categorical node): the numeric loop runs until
    // either bit shows up, categorical nodes are resolved here and the walk re-enters
    do {
      if (!(n & 2u)) {
        asm volatile(
            "{\n"
            ".reg .pred pl, pq;\n"
            ".reg .b32 w0, w1, off, code, kk, sel, tmp;\n"
            "LVLC:\n"
            "add.u32 tmp, %1, %0;\n"
            "ld.shared.v2.u32 {w0, w1}, [tmp];\n"
            "lop3.b32 off, w0, 0xFFC0, %3, 0xEA;\n"
            "add.u32 off, off, %2;\n"
            "ld.shared.u16 code, [off];\n"
            "shr.u32 kk, w0, 16;\n"
            "setp.le.u32 pl, code, kk;\n"
            "selp.b32 sel, 0x4410, 0x4432, pl;\n"
            "prmt.b32 %0, w1, 0, sel;\n"
            "and.b32 tmp, %0, 3;\n"
            "setp.eq.u32 pq, tmp, 0;\n"
            "@pq bra LVLC;\n"
            "}\n"
            : "+r"(n)
            : "r"(cb_addr), "r"(xwarp_addr), "r"(lane2)
            : "memory");
        if (n & 1u) break;
      }
      // categorical node: NaN / negative / out-of-bitset go right (LightGBM CategoricalDecision)
      const uint2 nd = *reinterpret_cast<const uint2 *>(cb + (n & ~3u));
      const uint32_t code = cat_of_code(*reinterpret_cast<const uint16_t *>(xwarp + ((nd.x & 0xFFC0u) | lane2)), cat16);
      bool left = false;
      if (code != kBinNaN) {
        const uint2 ct = reinterpret_cast<const uint2 *>(cb)[nd.x >> 16];
        const uint32_t w = code >> 5;
        if (w < ct.y) left = (reinterpret_cast<const uint32_t *>(cb)[ct.x + w] >> (code & 31u)) & 1u;
      }
      n = __byte_perm(nd.y, 0u, left ? 0x4410u : 0x4432u);
    } while (!(n & 1u));
  } else if (ALIGNED) {
    asm volatile(
        "{\n"
        ".reg .pred pl, pq;\n"
        ".reg .b32 w0, w1, off, code, kk, sel, tmp;\n"
        "LVLA:\n"
        "add.u32 tmp, %1, %0;\n"
        "ld.shared.v2.u32 {w0, w1}, [tmp];\n"
        "lop3.b32 off, w0, 0xFFC0, %2, 0xEA;\n"   // (w0 & 0xFFC0) | (warp tile base | lane*2)
        "ld.shared.u16 code, [off];\n"
        "shr.u32 kk, w0, 16;\n"
        "setp.le.u32 pl, code, kk;\n"               // the NaN direction is baked into the column (BinMeta::flags)
        "selp.b32 sel, 0x4410, 0x4432, pl;\n"
        "prmt.b32 %0, w1, 0, sel;\n"
        "and.b32 tmp, %0, 1;\n"
        "setp.eq.u32 pq, tmp, 0;\n"
        "@pq bra LVLA;\n"
        "}\n"
        : "+r"(n)
        : "r"(cb_addr), "r"(xwarp_addr | lane2)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred pl, pq;\n"
        ".reg .b32 w0, w1, off, code, kk, sel, tmp;\n"
        "LVL:\n"
        "add.u32 tmp, %1, %0;\n"
        "ld.shared.v2.u32 {w0, w1}, [tmp];\n"
        "lop3.b32 off, w0, 0xFFC0, %3, 0xEA;\n"   // (w0 & 0xFFC0) | lane*2
        "add.u32 off, off, %2;\n"
        "ld.shared.u16 code, [off];\n"
        "shr.u32 kk, w0, 16;\n"
        "setp.le.u32 pl, code, kk;\n"
        "selp.b32 sel, 0x4410, 0x4432, pl;\n"
        "prmt.b32 %0, w1, 0, sel;\n"
        "and.b32 tmp, %0, 1;\n"
        "setp.eq.u32 pq, tmp, 0;\n"
        "@pq bra LVL;\n"
        "}\n"
        : "+r"(n)
        : "r"(cb_addr), "r"(xwarp_addr), "r"(lane2)
        : "memory");
  }
  return n;
}

// ALIGNED (tile_cols a power of two): every warp's tile starts on a multiple of its own size, so the code's
// address is (column offset | thread base) — one LOP3, no add.
template <typename Real, bool HAS_CAT, bool ALIGNED>
__global__ void __launch_bounds__(1024) gbdt_score_compact_kernel(const BParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int W = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = p.n_features;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  const bool resident = p.n_chunks == 1;
  uint8_t *cbuf0 = smem + 128;
  uint8_t *cbuf1 = cbuf0 + (resident ? 0u : p.chunk_stride);
  uint8_t *xs = cbuf1 + p.chunk_stride;
  if (ALIGNED) {
    const uint32_t wt = (uint32_t)F * 64u, a0 = smem_u32(xs);
    xs += ((a0 + wt - 1u) & ~(wt - 1u)) - a0;
  }
  const uint8_t *xwarp = xs + (size_t)warp * F * 64;  // this warp's [feature][lane] u16 codes
  const uint32_t lane2 = (uint32_t)lane * 2u;

  const int n_tiles = (p.rows + W - 1) / W;
  const int groups_per_tile = W >> 5;
  const int n_groups = (p.rows + 31) >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < n_tiles) {
    const ChunkDesc cd = p.chunks[0];
    mbar_arrive_expect_tx(&bars[0], cd.bytes);
    tma_bulk_g2s(cbuf0, p.model + cd.byte_off, cd.bytes, &bars[0]);
  }
  uint32_t it = 0, tile_it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
    if (tid == 0) {
      const int g0 = tile * groups_per_tile, g1 = min(n_groups, g0 + groups_per_tile);
      const uint32_t bytes = (uint32_t)(g1 - g0) * (uint32_t)F * 64u;
      fence_proxy_async();
      mbar_arrive_expect_tx(&bars[2], bytes);
      tma_bulk_g2s(xs, p.bins + (size_t)g0 * F * 32, bytes, &bars[2]);
    }
    mbar_wait(&bars[2], tile_it & 1);
    const int item = tile * W + tid;
    Real acc = (sizeof(Real) == 4) ? (Real)p.base_score : (Real)0;
    for (int c = 0; c < p.n_chunks; ++c, ++it) {
      if (!resident && tid == 0) {
        const bool more = (c + 1 < p.n_chunks) || (tile + (int)gridDim.x < n_tiles);
        if (more) {
          const int nc = (c + 1 < p.n_chunks) ? c + 1 : 0;
          const ChunkDesc cd = p.chunks[nc];
          uint64_t *bar = &bars[(it + 1) & 1];
          fence_proxy_async();
          mbar_arrive_expect_tx(bar, cd.bytes);
          tma_bulk_g2s(((it + 1) & 1) ? cbuf1 : cbuf0, p.model + cd.byte_off, cd.bytes, bar);
        }
      }
      if (!resident || it == 0) mbar_wait(&bars[it & 1], (it >> 1) & 1);
      const uint8_t *cb = (!resident && (it & 1)) ? cbuf1 : cbuf0;
      const int ntree = (int)*reinterpret_cast<const uint32_t *>(cb);
      const uint32_t *roots = reinterpret_cast<const uint32_t *>(cb + 16);
      const uint32_t cb_addr = smem_u32(cb), xwarp_addr = smem_u32(xwarp);
      auto walk = [&](uint32_t root) -> Real {
        const uint32_t n = walk_tree<HAS_CAT, ALIGNED>(root, cb, cb_addr, xwarp, xwarp_addr, lane2, p.cat16 != 0);
        return *reinterpret_cast<const Real *>(cb + (n - 1u));
      };
      // four roots per (warp-uniform) LDS.128; the leaf values are still added one by one, in tree order
      int t = 0;
#pragma unroll 1
      for (; t + 4 <= ntree; t += 4) {
        const uint4 r = *reinterpret_cast<const uint4 *>(roots + t);
        acc += walk(r.x);
        acc += walk(r.y);
        acc += walk(r.z);
        acc += walk(r.w);
      }
#pragma unroll 1
      for (; t < ntree; t++) acc += walk(roots[t]);
      __syncthreads();
    }
    if (item < p.rows) store_score(p.out, p.sinks, item, (double)acc);
  }
  if (p.sinks.n_peer) publish_when_last(p.sinks);
}

// ------------------------------------------------------------------ slim lock-step traversal (4-byte nodes)
// The throughput scorer's fast path (gbdt_model.h SlimModel).  Same mapping as the compact kernel — thread per item,
// trees in tree order, chunks double-buffered by TMA — with everything that cost a level an instruction or a
// shared-memory wavefront taken out:
//   * entries are 4 bytes: lanes on different nodes of a tree are served by ONE wavefront (an LDS.64 is two half-warp passes);
//   * the code tile is [column pair][item of the CTA] u32 at absolute shared address (pair + 1) * 4T: the code's address is
//     (entry & mask) | 4 * tid — one LOP3, no add, and lane l always reads bank l (no conflicts whatever the columns);
//   * k is compared as a binary16 pattern straight out of the entry's high half: one HSETP2, no shift;
//   * a child is (entry & mask) | block base, + 4 for the right one; a leaf is an entry with the sign bit set.
// One level: LOP3, LDS.U16, HSETP2, LOP3, IADD (predicated), LDS.32, ISETP, BRA.
template <int T> struct SlimConst {
  static constexpr uint32_t kShift = T == 512 ? 11u : T == 256 ? 10u : 9u;
  static constexpr uint32_t kColMask = ((0xFFFFu << kShift) & 0xFFFFu) | 2u;
  static constexpr uint32_t kChildMask = ((1u << kShift) - 1u) & ~7u;
};

// Root table as a kernel parameter (gbdt_model.h SlimModel::root_tab): NR entries of {block offset, root entry, left
// child entry, right child entry}, read with warp-uniform indices from the constant bank.  NR == 0: the chunk's own
// {block offset, root entry} table in shared memory.
template <int NR> struct SlimRoots { uint4 e[NR > 0 ? NR : 1]; };

// The numeric level of the walk, and the same level for models whose categorical nodes are in the in-loop form
// (gbdt_model.h kMetaCat16): `entry >> (code & 31)` puts the node's bitset bit for this category at bit 0, `& entry & 1`
// keeps it only at a categorical entry, and the compare takes it as its OR operand — a categorical code is a binary16 NaN,
// so the numeric half of the OR is false there, and a numeric entry has bit 0 clear, so the categorical half is false here.
#define MR_SLIM_TEST_NUM                                                                     \
  "mov.b32 {wlo, whi}, %0;\n"                                                                \
  "mov.b32 {clo, chi}, code;\n"                                                              \
  "setp.le.f16 pl, clo, whi;\n"
#define MR_SLIM_TEST_CAT16                                                                   \
  "shf.r.wrap.b32 cb, %0, 0, code;\n"                                                        \
  "and.b32 cb, cb, %0;\n"                                                                    \
  "and.b32 cb, cb, 1;\n"                                                                     \
  "setp.ne.u32 pc, cb, 0;\n"                                                                 \
  "mov.b32 {wlo, whi}, %0;\n"                                                                \
  "mov.b32 {clo, chi}, code;\n"                                                              \
  "setp.le.or.f16 pl, clo, whi, pc;\n"
// one level from the entry in %0 to the entry of the child taken; loops until a leaf entry (sign bit) arrives
#define MR_SLIM_LOOP(LABEL, TEST)                                                            \
  LABEL ":\n"                                                                                \
  "lop3.b32 off, %0, %1, %2, 0xEA;\n" /* (entry & column mask) | 4 * tid */                  \
  "ld.shared.u16 code, [off];\n"                                                             \
  TEST                                                                                       \
  "lop3.b32 n, %0, %3, %4, 0xEA;\n"   /* (entry & child mask) | block base */                \
  "@!pl add.u32 n, n, 4;\n"                                                                  \
  "ld.shared.u32 %0, [n];\n"                                                                 \
  "setp.ge.s32 pq, %0, 0;\n"                                                                 \
  "@pq bra " LABEL ";\n"
// level 0 with the root entry (%5, warp-uniform) and both child entries (%6, %7) already in registers: no node load,
// and no leaf test either — in a model with a root table the root's children are internal entries (pack_slim)
#define MR_SLIM_ROOT(TEST0)                                                                  \
  "lop3.b32 off, %5, %1, %2, 0xEA;\n"                                                        \
  "ld.shared.u16 code, [off];\n"                                                             \
  TEST0                                                                                      \
  "selp.b32 %0, %6, %7, pl;\n"
#define MR_SLIM_TEST0_NUM                                                                    \
  "mov.b32 {wlo, whi}, %5;\n"                                                                \
  "mov.b32 {clo, chi}, code;\n"                                                              \
  "setp.le.f16 pl, clo, whi;\n"
#define MR_SLIM_TEST0_CAT16                                                                  \
  "shf.r.wrap.b32 cb, %5, 0, code;\n"                                                        \
  "and.b32 cb, cb, %5;\n"                                                                    \
  "and.b32 cb, cb, 1;\n"                                                                     \
  "setp.ne.u32 pc, cb, 0;\n"                                                                 \
  "mov.b32 {wlo, whi}, %5;\n"                                                                \
  "mov.b32 {clo, chi}, code;\n"                                                              \
  "setp.le.or.f16 pl, clo, whi, pc;\n"
#define MR_SLIM_REGS                                                                         \
  ".reg .pred pl, pq, pc;\n"                                                                 \
  ".reg .b32 off, code, n, cb;\n"                                                            \
  ".reg .b16 wlo, whi, clo, chi;\n"

// CAT: 0 = no categorical node; 1 = categorical nodes with bitsets of any width (the loop leaves on them, they are resolved
// in C++ and the loop is re-entered); 2 = small-categorical form, resolved inside the loop.
template <typename Real, int T, int CAT, int NR>
__global__ void __launch_bounds__(T) gbdt_score_slim_kernel(const __grid_constant__ BParams p, const __grid_constant__ SlimRoots<NR> rt) {
  extern __shared__ __align__(128) uint8_t smem[];
  using K = SlimConst<T>;
  static_assert(NR == 0 || CAT != 1, "the parameter root table serves the in-loop forms only");
  const int tid = threadIdx.x;
  const uint32_t base = smem_u32(smem);         // absolute shared address of the dynamic window (small: asserted on the host)
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  const uint32_t tile_abs = (uint32_t)T * 4u * (uint32_t)p.col_base;   // pair p of the tile lives at (p + col_base) * 4T
  const uint32_t n_pairs = (uint32_t)(p.n_features + 1) >> 1;
  const uint32_t cb0_abs = ((uint32_t)T * 4u * (n_pairs + (uint32_t)p.col_base) + 2047u) & ~2047u, cb1_abs = cb0_abs + p.chunk_stride;
  uint8_t *tile_ptr = smem + (tile_abs - base);
  const bool resident = p.n_chunks == 1;
  const uint32_t tid4 = (uint32_t)tid * 4u;
  uint32_t lane_zero;
  asm volatile("shr.u32 %0, %1, 31;" : "=r"(lane_zero) : "r"(tid4));

  const int n_tiles = (p.rows + T - 1) / T;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < n_tiles) {
    const ChunkDesc cd = p.chunks[0];
    mbar_arrive_expect_tx(&bars[0], cd.bytes);
    tma_bulk_g2s(smem + (cb0_abs - base), p.model + cd.byte_off, cd.bytes, &bars[0]);
  }
  uint32_t it = 0, tile_it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
    if (tid == 0) {
      const uint32_t bytes = n_pairs * (uint32_t)T * 4u;  // the whole CTA tile is one contiguous range (buffers are padded to whole tiles)
      fence_proxy_async();
      mbar_arrive_expect_tx(&bars[2], bytes);
      tma_bulk_g2s(tile_ptr, reinterpret_cast<const uint8_t *>(p.bins) + (size_t)tile * bytes, bytes, &bars[2]);
    }
    mbar_wait(&bars[2], tile_it & 1);
    const int item = tile * T + tid;
    Real acc = (sizeof(Real) == 4) ? (Real)p.base_score : (Real)0;
    for (int c = 0; c < p.n_chunks; ++c, ++it) {
      if (!resident && tid == 0) {
        const bool more = (c + 1 < p.n_chunks) || (tile + (int)gridDim.x < n_tiles);
        if (more) {
          const int nc = (c + 1 < p.n_chunks) ? c + 1 : 0;
          const ChunkDesc cd = p.chunks[nc];
          uint64_t *bar = &bars[(it + 1) & 1];
          fence_proxy_async();
          mbar_arrive_expect_tx(bar, cd.bytes);
          tma_bulk_g2s(smem + ((((it + 1) & 1) ? cb1_abs : cb0_abs) - base), p.model + cd.byte_off, cd.bytes, bar);
        }
      }
      if (!resident || it == 0) mbar_wait(&bars[it & 1], (it >> 1) & 1);
      const uint32_t cb_abs = (!resident && (it & 1)) ? cb1_abs : cb0_abs;
      const uint8_t *cb = smem + (cb_abs - base);
      const int ntree = (int)*reinterpret_cast<const uint32_t *>(cb);
      auto leaf_value = [&](uint32_t tb, uint32_t w) -> Real {
        return *reinterpret_cast<const Real *>(smem + (((w & 0xFFFFu) | tb) - base));
      };
      if (NR > 0) {
        // level 0 from the parameter table: {block offset, root entry, left entry, right entry} of tree `first + t`
        const uint4 *tab = rt.e + *reinterpret_cast<const uint32_t *>(cb + 4);
        auto walk = [&](const uint4 r) -> Real {
          // `lane_zero` is 0, but only at run time: ptxas would otherwise prove the block base warp-uniform, keep it in a
          // uniform register — which a LOP3 cannot read — and re-materialise it with an extra move on every level
          const uint32_t tb = cb_abs + r.x + lane_zero;
          uint32_t w;
          if (CAT == 0) {
            asm volatile("{\n" MR_SLIM_REGS MR_SLIM_ROOT(MR_SLIM_TEST0_NUM) MR_SLIM_LOOP("SLVL", MR_SLIM_TEST_NUM) "}\n"
                         : "=r"(w)
                         : "n"(K::kColMask), "r"(tid4), "n"(K::kChildMask), "r"(tb), "r"(r.y), "r"(r.z), "r"(r.w)
                         : "memory");
          } else {
            asm volatile("{\n" MR_SLIM_REGS MR_SLIM_ROOT(MR_SLIM_TEST0_CAT16) MR_SLIM_LOOP("SLVL", MR_SLIM_TEST_CAT16) "}\n"
                         : "=r"(w)
                         : "n"(K::kColMask), "r"(tid4), "n"(K::kChildMask), "r"(tb), "r"(r.y), "r"(r.z), "r"(r.w)
                         : "memory");
          }
          return leaf_value(tb, w);
        };
        int t = 0;
#pragma unroll 1
        for (; t + 4 <= ntree; t += 4) {
          const uint4 r0 = tab[t], r1 = tab[t + 1], r2 = tab[t + 2], r3 = tab[t + 3];
          acc += walk(r0);
          acc += walk(r1);
          acc += walk(r2);
          acc += walk(r3);
        }
#pragma unroll 1
        for (; t < ntree; t++) acc += walk(tab[t]);
      } else {
        const uint32_t *roots = reinterpret_cast<const uint32_t *>(cb + 16);
        auto walk = [&](uint32_t block_off, uint32_t w) -> Real {
          // absolute address of the tree's block, aligned to the block's size (`lane_zero`: see above).  `w` is the root
          // entry (always an internal one), delivered with the root table.
          const uint32_t tb = cb_abs + block_off + lane_zero;
          if (CAT == 0) {
            asm volatile("{\n" MR_SLIM_REGS MR_SLIM_LOOP("SLVL", MR_SLIM_TEST_NUM) "}\n"
                         : "+r"(w)
                         : "n"(K::kColMask), "r"(tid4), "n"(K::kChildMask), "r"(tb)
                         : "memory");
          } else if (CAT == 2) {
            asm volatile("{\n" MR_SLIM_REGS MR_SLIM_LOOP("SLVL", MR_SLIM_TEST_CAT16) "}\n"
                         : "+r"(w)
                         : "n"(K::kColMask), "r"(tid4), "n"(K::kChildMask), "r"(tb)
                         : "memory");
          } else {
            // entries carry their kind: sign bit = leaf, bit 0 = categorical node.  The numeric loop runs until either shows up
            // (one LOP3 with a predicate result instead of the sign test), categorical nodes are resolved here, then it re-enters
            for (;;) {
              if (!(w & 1u)) {
                asm volatile(
                    "{\n"
                    ".reg .pred pl, pq;\n"
                    ".reg .b32 off, code, n, tmp;\n"
                    ".reg .b16 wlo, whi, clo, chi;\n"
                    "SLVC:\n"
                    "lop3.b32 off, %0, %1, %2, 0xEA;\n"
                    "ld.shared.u16 code, [off];\n"
                    "mov.b32 {wlo, whi}, %0;\n"
                    "mov.b32 {clo, chi}, code;\n"
                    "setp.le.f16 pl, clo, whi;\n"
                    "lop3.b32 n, %0, %3, %4, 0xEA;\n"
                    "@!pl add.u32 n, n, 4;\n"
                    "ld.shared.u32 %0, [n];\n"
                    "and.b32 tmp, %0, 0x80000001;\n"
                    "setp.eq.u32 pq, tmp, 0;\n"
                    "@pq bra SLVC;\n"
                    "}\n"
                    : "+r"(w)
                    : "n"(K::kColMask), "r"(tid4), "n"(K::kChildMask), "r"(tb)
                    : "memory");
                if ((int)w < 0) break;
              }
              // categorical node (LightGBM CategoricalDecision): NaN / negative / beyond the bitset go right
              const uint32_t code = *reinterpret_cast<const uint16_t *>(smem + (((w & K::kColMask) | tid4) - base));
              bool left = false;
              if (code != kBinNaN) {
                const uint2 ct = *reinterpret_cast<const uint2 *>(smem + ((tb + ((w >> 16) & 0x7FFFu) * 8u) - base));
                const uint32_t wd = code >> 5;
                if (wd < ct.y) left = (*reinterpret_cast<const uint32_t *>(smem + ((tb + ct.x + wd * 4u) - base)) >> (code & 31u)) & 1u;
              }
              const uint32_t n = ((w & K::kChildMask) | tb) + (left ? 0u : 4u);
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(n));
              if ((int)w < 0) break;
            }
          }
          return leaf_value(tb, w);
        };
        // two trees per (warp-uniform) LDS.128: {block, root entry} x 2; the leaf values are still added one by one, in tree order
        int t = 0;
#pragma unroll 1
        for (; t + 4 <= ntree; t += 4) {
          const uint4 r0 = *reinterpret_cast<const uint4 *>(roots + 2 * t), r1 = *reinterpret_cast<const uint4 *>(roots + 2 * t + 4);
          acc += walk(r0.x, r0.y);
          acc += walk(r0.z, r0.w);
          acc += walk(r1.x, r1.y);
          acc += walk(r1.z, r1.w);
        }
#pragma unroll 1
        for (; t < ntree; t++) acc += walk(roots[2 * t], roots[2 * t + 1]);
      }
      __syncthreads();
    }
    if (item < p.rows) store_score(p.out, p.sinks, item, (double)acc);
  }
  if (p.sinks.n_peer) publish_when_last(p.sinks);
}

// ------------------------------------------------------------------ low-latency path (small batches)
// A single /rank request is 100 items: one thread per item walking 500 trees is a 100+ us dependent
// chain on 4 warps of one SM.  For small batches the trees are spread over the chip instead: CTA
// (chunk c, item group g) walks only chunk c's trees for 128 items and stores every tree's leaf value;
// a second kernel adds the values per item IN TREE ORDER, so the result is still bit-identical to the
// sequential sum (a tree-parallel reduction would change the f64 rounding).
struct LParams {
  const uint8_t *model;
  const ChunkDesc *chunks;
  const uint16_t *bins;
  uint16_t *leafslots;       // [n_trees][rows_padded]: the leaf each row reached, as its 8-byte slot index in the tree's chunk
  const uint32_t *tree_off;  // [n_trees] byte offset of the tree's chunk in `model`
  double *out;
  int rows, rows_padded, n_features, n_trees;
  int n_chunks, chunks_per_cta, has_cat;  // has_cat: 0 none, 1 categorical nodes, 2 with small-categorical codes (kMetaCat16)
  uint32_t chunk_stride;  // bytes per chunk buffer
  float base_score;
};

template <typename Real>
__global__ void __launch_bounds__(128) gbdt_leaves_kernel(const LParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = p.n_features;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);  // [0], [1]: chunk buffers; [2]: the code tile
  const uint32_t stride = p.chunk_stride;
  uint8_t *cbuf = smem + 128;
  uint8_t *xs = cbuf + 2 * (size_t)stride;
  const int n_groups = (p.rows + 31) >> 5;
  const int g0 = blockIdx.y * 4, g1 = min(n_groups, g0 + 4);
  // this CTA walks chunks [c_lo, c_hi) for its 128 items: the tile is staged once, the chunks stream through two buffers
  const int c_lo = blockIdx.x * p.chunks_per_cta, c_hi = min(p.n_chunks, c_lo + p.chunks_per_cta);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
    const uint32_t tile_bytes = (uint32_t)(g1 - g0) * (uint32_t)F * 64u;
    mbar_arrive_expect_tx(&bars[2], tile_bytes);
    tma_bulk_g2s(xs, p.bins + (size_t)g0 * F * 32, tile_bytes, &bars[2]);
    const ChunkDesc cd = p.chunks[c_lo];
    mbar_arrive_expect_tx(&bars[0], cd.bytes);
    tma_bulk_g2s(cbuf, p.model + cd.byte_off, cd.bytes, &bars[0]);
  }
  __syncthreads();
  mbar_wait(&bars[2], 0);
  const bool has_items = g0 + warp < g1;
  const int item = (g0 + warp) * 32 + lane;
  const uint8_t *xwarp = xs + (size_t)warp * F * 64;
  const uint32_t lane2 = (uint32_t)lane * 2u;
  for (int c = c_lo, it = 0; c < c_hi; c++, it++) {
    if (tid == 0 && c + 1 < c_hi) {  // prefetch the next chunk into the other buffer (its readers passed the barrier below)
      const ChunkDesc nd = p.chunks[c + 1];
      uint64_t *bar = &bars[(it + 1) & 1];
      fence_proxy_async();
      mbar_arrive_expect_tx(bar, nd.bytes);
      tma_bulk_g2s(cbuf + (size_t)((it + 1) & 1) * stride, p.model + nd.byte_off, nd.bytes, bar);
    }
    mbar_wait(&bars[it & 1], (it >> 1) & 1);
    const uint8_t *cb = cbuf + (size_t)(it & 1) * stride;
    const uint32_t first_tree = p.chunks[c].first_tree;
    if (has_items) {
      const int ntree = (int)*reinterpret_cast<const uint32_t *>(cb);
      const uint32_t *roots = reinterpret_cast<const uint32_t *>(cb + 16);
      const uint32_t cb_addr = smem_u32(cb), xwarp_addr = smem_u32(xwarp);
      uint16_t *dst = p.leafslots + (size_t)first_tree * p.rows_padded + item;
      // the compact scorer's hand-scheduled level loop (10 SASS instructions; nvcc's own schedule of the same walk is 18)
      if (p.has_cat) {
        for (int t = 0; t < ntree; t++, dst += p.rows_padded)
          *dst = (uint16_t)((walk_tree<true, false>(roots[t], cb, cb_addr, xwarp, xwarp_addr, lane2, p.has_cat == 2) - 1u) >> 3);
      } else {
        for (int t = 0; t < ntree; t++, dst += p.rows_padded)
          *dst = (uint16_t)((walk_tree<false, false>(roots[t], cb, cb_addr, xwarp, xwarp_addr, lane2) - 1u) >> 3);
      }
    }
    __syncthreads();  // this buffer is refilled two chunks from now
  }
}

// In-order sum of the per-tree leaf values.  The adds are a dependent chain (tree order = the sequential reference's
// rounding: T dependent DADDs, 8.2 cycles each on B200 — tools/micro/fp64_latency.cu — i.e. 8 us for the 2000 trees of
// BASELINE config #5) and everything else is arranged so that the chain is all that remains.  The model's small-chunk
// packing is cut into GROUPS of consecutive chunks (<= 16 KB, <= 64 trees: one contiguous byte range).  A CTA is one
// producer warp and 1 or 4 consumer warps of 32 rows each, around a ring of kSumRing group buffers:
//   producer   per group ONE TMA bulk copy of the group's model bytes (leaf values included) plus 16-byte cp.async copies
//              of the rows' leaf slots (2 bytes per row and tree) and of the trees' offsets inside the group; all of them
//              complete on the buffer's `full` mbarrier, so the producer never waits for data — only for a free buffer;
//   consumer   waits for `full`; per tree: slot (LDS.U16) -> value at group + tree offset + 8 * slot (LDS.64) -> acc += value;
//              raises `empty`.  No per-lane gathers from global memory anywhere (8-byte cp.async gathers cost ~20 wavefronts
//              an instruction and saturated the L1 pipe at 76 %: profiles/ncu_r2_summary.md).
constexpr int kSumRing = 3;     // ring depth when the grid is several CTAs per SM
constexpr int kSumRingMax = 8;  // ... and the most a lone CTA per SM takes (mega-request slices: 79 CTAs on 148 SMs) — the ring
                                // then hides the whole fetch latency of a group behind the groups in flight ahead of it

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct SumParams {
  const uint8_t *model;       // small-chunk compact packing
  const uint16_t *leafslots;  // [n_trees][rows_padded]
  const uint4 *groups;        // per group {first tree, n trees, byte offset in model, bytes}
  const uint32_t *group_rel;  // [n_groups][kSumGroupTrees] byte offset of each tree's chunk inside its group
  double *out;
  int rows, rows_padded, n_groups;
  uint32_t group_stride;      // bytes reserved per ring buffer for the model bytes (multiple of 128)
  float base_score;
  int ring;                   // ring depth (kSumRing .. kSumRingMax)
};

template <typename Real>
__global__ void __launch_bounds__(160) gbdt_sum_kernel(const SumParams p, const ScoreSinks sinks) {
  extern __shared__ __align__(128) uint8_t s_sum_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(s_sum_raw), *empty = full + kSumRingMax;
  const int ring = p.ring;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = (blockDim.x >> 5) - 1;  // consumer warps = 32-row slabs of this CTA
  const int cta_rows = 32 * C;
  const int row0 = blockIdx.x * cta_rows;
  // ring buffer b: [model bytes: group_stride][slots: kSumGroupTrees x cta_rows u16][rel: kSumGroupTrees u32]
  const size_t buf_bytes = (size_t)p.group_stride + (size_t)kSumGroupTrees * cta_rows * 2 + kSumGroupTrees * 4;
  uint8_t *bufs = s_sum_raw + 128;
  if (threadIdx.x == 0) {
    for (int k = 0; k < ring; k++) { mbar_init(&full[k], 33); mbar_init(&empty[k], C); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {
    // ---- producer
    for (int g = 0; g < p.n_groups; g++) {
      const int b = g % ring, use = g / ring;
      if (use > 0) mbar_wait_spin(&empty[b], (use - 1) & 1);
      const uint4 gd = __ldg(p.groups + g);
      uint8_t *mb = bufs + (size_t)b * buf_bytes;
      uint16_t *slots = reinterpret_cast<uint16_t *>(mb + p.group_stride);
      uint32_t *rel = reinterpret_cast<uint32_t *>(slots + (size_t)kSumGroupTrees * cta_rows);
      if (lane == 0) {
        fence_proxy_async();
        mbar_arrive_expect_tx(&full[b], gd.w);
        tma_bulk_g2s(mb, p.model + gd.z, gd.w, &full[b]);
      }
      // 16-byte pieces of 8 rows: 4 (one consumer) or 16 (four) per tree — a power of two, and piece k of the group lands
      // at byte 16 k of the slot tile, so the loop is a shift, a mask and two adds per copy (this warp feeds four others)
      const int pshift = C == 1 ? 2 : 4;
      const uint32_t dst = smem_u32(slots);
      const uint16_t *src0 = p.leafslots + (size_t)gd.x * p.rows_padded + row0;
      for (int k = lane; k < ((int)gd.y << pshift); k += 32) {
        const uint16_t *src = src0 + (size_t)(k >> pshift) * p.rows_padded + (k & ((1 << pshift) - 1)) * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)k * 16u), "l"(src) : "memory");
      }
      if (lane * 4 < (int)gd.y)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(rel) + (uint32_t)lane * 16u),
                     "l"(p.group_rel + (size_t)g * kSumGroupTrees + lane * 4)
                     : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[b])) : "memory");
    }
  } else {
    // ---- consumer of rows [row0 + 32 (warp - 1), + 32)
    const int item = row0 + (warp - 1) * 32 + lane;
    const bool live = item < p.rows;
    Real acc = (sizeof(Real) == 4) ? (Real)p.base_score : (Real)0;
    for (int g = 0; g < p.n_groups; g++) {
      const int b = g % ring, use = g / ring;
      mbar_wait_spin(&full[b], use & 1);
      const int nt = (int)__ldg(&p.groups[g].y);
      if (live) {
        const uint8_t *mb = bufs + (size_t)b * buf_bytes;
        const uint16_t *sl = reinterpret_cast<const uint16_t *>(mb + p.group_stride) + (warp - 1) * 32 + lane;
        const uint32_t *rel = reinterpret_cast<const uint32_t *>(mb + p.group_stride + (size_t)kSumGroupTrees * cta_rows * 2);
        // Two dependent shared-memory reads per tree (slot, then the value behind it) — batches of 16 trees so that their
        // latencies overlap: 16 slots, then 16 values, then the 16 adds in tree order.  The value slot holds the model's
        // Real in its first bytes (f64, or f32 for XGBoost).
        int k0 = 0;
        for (; k0 + 16 <= nt; k0 += 16) {
          uint32_t off[16];
          Real v[16];
#pragma unroll
          for (int j = 0; j < 16; j++) off[j] = rel[k0 + j] + (uint32_t)sl[(size_t)(k0 + j) * cta_rows] * 8u;
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] = *reinterpret_cast<const Real *>(mb + off[j]);
#pragma unroll
          for (int j = 0; j < 16; j++) acc += v[j];
        }
        for (; k0 < nt; k0++) acc += *reinterpret_cast<const Real *>(mb + rel[k0] + (uint32_t)sl[(size_t)k0 * cta_rows] * 8u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
    }
    if (live) store_score(p.out, sinks, item, (double)acc);
  }
  if (sinks.n_peer) publish_when_last(sinks);
}

// ------------------------------------------------------------------ walk statistics (bench.py's roofline block)
// What the lock-step scorer executes on a batch, counted rather than assumed: a warp walks every tree to the
// depth of its deepest lane, so per (warp, tree) it issues max-depth level steps while its lanes need sum-depth/32
// of them.  One thread per item straight from HBM/L2 (no staging): a measuring aid, not a scoring path.
struct WalkStats { unsigned long long lane_levels, warp_levels, warp_trees; };

__global__ void __launch_bounds__(128) compact_walk_stats_kernel(const bool cat16, const uint8_t *model, const ChunkDesc *chunks, int n_chunks,
                                                                 const uint16_t *bins, int rows, int F, WalkStats *out) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = item < rows;
  const uint16_t *xw = bins + ((size_t)(item >> 5) * F) * 32 + lane;
  unsigned long long lane_levels = 0, warp_levels = 0, warp_trees = 0;
  for (int c = 0; c < n_chunks; c++) {
    const uint8_t *cb = model + chunks[c].byte_off;
    const int ntree = (int)*reinterpret_cast<const uint32_t *>(cb);
    const uint32_t *roots = reinterpret_cast<const uint32_t *>(cb + 16);
    for (int t = 0; t < ntree; t++) {
      uint32_t n = roots[t];
      int depth = 0;
      while (live && !(n & 1u)) {
        const uint2 nd = *reinterpret_cast<const uint2 *>(cb + (n & ~2u));
        uint32_t code = xw[(nd.x & 0xFFC0u) >> 1];
        bool left;
        if (nd.x & 2u) {
          left = false;
          code = cat_of_code(code, cat16);
          if (code != kBinNaN) {
            const uint2 ct = reinterpret_cast<const uint2 *>(cb)[nd.x >> 16];
            const uint32_t w = code >> 5;
            if (w < ct.y) left = (reinterpret_cast<const uint32_t *>(cb)[ct.x + w] >> (code & 31u)) & 1u;
          }
        } else {
          left = code <= (nd.x >> 16);
        }
        n = __byte_perm(nd.y, 0u, left ? 0x4410u : 0x4432u);
        depth++;
      }
      lane_levels += depth;
      warp_levels += __reduce_max_sync(0xFFFFFFFFu, depth);
      warp_trees++;
    }
  }
  const unsigned long long lanes = lane_levels;
  unsigned long long tot = lanes;
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
  if (lane == 0) {
    atomicAdd(&out->lane_levels, tot);
    atomicAdd(&out->warp_levels, warp_levels);
    atomicAdd(&out->warp_trees, warp_trees);
  }
}

template <typename Real, bool HAS_CAT, int ILP>
void launch_b(const BParams &p, int threads, size_t smem, int num_sms, cudaStream_t stream) {
  auto kern = gbdt_score_binned_kernel<Real, HAS_CAT, ILP>;
  MR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) fail(MR_ERR_CUDA, "binned gbdt kernel does not fit on an SM (smem %zu, threads %d)", smem, threads);
  const int n_tiles = (p.rows + threads - 1) / threads;
  const int grid = std::max(1, std::min(n_tiles, num_sms * per_sm));
  { ProfScope _ps("gbdt_score_binned_kernel", stream); kern<<<grid, threads, smem, stream>>>(p); }
  MR_CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}


}  // namespace

void compact_walk_stats(const BinnedLaunch &L, unsigned long long out[3], cudaStream_t stream) {
  WalkStats *d = nullptr;
  MR_CUDA_CHECK(cudaMalloc((void **)&d, sizeof(WalkStats)));
  MR_CUDA_CHECK(cudaMemsetAsync(d, 0, sizeof(WalkStats), stream));
  if (L.rows > 0)
    compact_walk_stats_kernel<<<(L.rows + 127) / 128, 128, 0, stream>>>(L.cat16, L.d_model, L.d_chunks, L.n_chunks, L.d_bins, L.rows, L.tile_cols, d);
  cudaError_t e = cudaGetLastError();
  WalkStats h{};
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(d);
  MR_CUDA_CHECK(e);
  out[0] = h.lane_levels; out[1] = h.warp_levels; out[2] = h.warp_trees;
}

void launch_gbdt_latency(const BinnedLaunch &L, int n_trees, const SumPlan &sum, void *d_leafslots, cudaStream_t stream) {
  // L.d_model / d_chunks = the small-chunk compact packing; L.d_bins already holds the codes
  if (L.rows <= 0) return;
  LParams p;
  p.model = L.d_model; p.chunks = L.d_chunks; p.bins = L.d_bins; p.leafslots = (uint16_t *)d_leafslots; p.tree_off = nullptr;
  p.out = L.d_out;
  p.rows = L.rows; p.rows_padded = (L.rows + 127) & ~127; p.n_features = L.tile_cols; p.n_trees = n_trees;
  p.base_score = L.base_score;
  p.n_chunks = L.n_chunks;
  p.has_cat = L.has_cat ? (L.cat16 ? 2 : 1) : 0;
  p.chunk_stride = (L.max_chunk_bytes + 127u) & ~127u;
  // (chunk, 128-item group) pairs are the unit of parallelism; once there are more of them than ~16 per SM a CTA takes
  // several chunks in a row for its group, so the code tile is staged once per CTA instead of once per 4 KB of trees
  const int n_item_groups = (L.rows + 127) / 128;
  p.chunks_per_cta = (int)std::max<long long>(1, std::min<long long>(16, ((long long)L.n_chunks * n_item_groups) / (148 * 16)));
  const size_t smem = 128 + 2 * (size_t)p.chunk_stride + (size_t)4 * L.tile_cols * 64;
  dim3 grid((unsigned)((L.n_chunks + p.chunks_per_cta - 1) / p.chunks_per_cta), (unsigned)n_item_groups);
  // the in-order sum: 32 rows per CTA while that keeps the chip busy (a mega-request slice of ~1 000 rows spreads over 32+
  // SMs), 128 rows per CTA for larger batches so that the model bytes are staged once per four slabs
  SumParams sp;
  sp.model = L.d_model; sp.leafslots = (const uint16_t *)d_leafslots; sp.groups = sum.d_groups; sp.group_rel = sum.d_group_rel;
  sp.out = L.d_out; sp.rows = L.rows; sp.rows_padded = p.rows_padded; sp.n_groups = sum.n_groups;
  sp.group_stride = (sum.max_group_bytes + 127u) & ~127u;
  sp.base_score = L.base_score;
  const int consumers = L.rows <= 148 * 32 ? 1 : 4;
  const int cta_rows = 32 * consumers;
  const size_t stage = (size_t)sp.group_stride + (size_t)kSumGroupTrees * cta_rows * 2 + kSumGroupTrees * 4;
  const int n_sum_ctas = (L.rows + cta_rows - 1) / cta_rows;
  sp.ring = n_sum_ctas <= 148 ? (int)std::max<size_t>(kSumRing, std::min<size_t>(kSumRingMax, (220 * 1024 - 128) / stage)) : kSumRing;
  const size_t sum_smem = 128 + sp.ring * stage;
  auto go = [&](auto leaves, auto sumk) {
    MR_CUDA_CHECK(cudaFuncSetAttribute(leaves, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { ProfScope _ps("gbdt_leaves_kernel", stream); leaves<<<grid, 128, smem, stream>>>(p); }
    MR_CUDA_CHECK(cudaGetLastError());
    MR_CUDA_CHECK(cudaFuncSetAttribute(sumk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sum_smem));
    { ProfScope _ps("gbdt_sum_kernel", stream); sumk<<<(L.rows + cta_rows - 1) / cta_rows, 32 * (consumers + 1), sum_smem, stream>>>(sp, L.sinks); }
    MR_CUDA_CHECK(cudaGetLastError());
    g_kernel_launches += 2;
  };
  if (L.kind == MR_BOOSTER_XGBOOST) go(gbdt_leaves_kernel<float>, gbdt_sum_kernel<float>);
  else go(gbdt_leaves_kernel<double>, gbdt_sum_kernel<double>);
}

void launch_gbdt_binned(const BinnedLaunch &L, int num_sms, cudaStream_t stream) {
  if (L.rows <= 0) return;
  const int F = L.tile_cols;  // the traversal only sees tile columns
  // ---- pass 1: codes (skipped when the caller — the fused assemble kernel — already wrote them)
  if (!L.codes_ready) {
  BinParams bp;
  bp.values = L.d_values; bp.thr_off = L.d_thr_off; bp.thr = L.d_thr; bp.is_cat = L.d_is_cat;
  bp.meta = L.d_meta; bp.bucket_range = L.d_bucket_range;
  bp.bins = L.d_bins; bp.rows = L.rows; bp.cols = L.cols; bp.n_features = L.n_features; bp.tile_cols = F;
  bp.xgb = L.kind == MR_BOOSTER_XGBOOST;
  bp.tile_T = L.tile_T;
  const int items_per_cta = kBinGroups * 32;
  const size_t bin_smem = (size_t)F * (items_per_cta + 2) * sizeof(uint16_t);
  if (bin_smem > 200 * 1024) fail(MR_ERR_UNSUPPORTED, "too many features for the binning kernel");
  MR_CUDA_CHECK(cudaFuncSetAttribute(bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem));
  { ProfScope _ps("bin_kernel", stream); bin_kernel<<<(L.rows + items_per_cta - 1) / items_per_cta, 256, bin_smem, stream>>>(bp); }
  MR_CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
  }

  if (L.codes_only) return;
  if (L.tile_T) {
    // ---- slim scorer: CTA = tile_T threads, tile at absolute shared address 4 * tile_T, two 2 KB-aligned chunk buffers
    BParams p;
    p.model = L.d_model; p.chunks = L.d_chunks; p.bins = L.d_bins; p.out = L.d_out;
    p.n_chunks = L.n_chunks; p.chunk_stride = (L.max_chunk_bytes + 2047u) & ~2047u;
    p.rows = L.rows; p.n_features = F; p.base_score = L.base_score; p.cat16 = L.cat16 ? 1 : 0;
    p.sinks = L.sinks;
    const int T = L.tile_T;
    const size_t n_pairs = (size_t)(F + 1) / 2;
    p.col_base = L.slim_col_base;
    const size_t cb0 = (((size_t)T * 4 * (n_pairs + (size_t)L.slim_col_base)) + 2047) & ~size_t(2047);
    const size_t smem = cb0 + (size_t)p.chunk_stride * (L.n_chunks == 1 ? 1 : 2);  // the window starts at (or just above) address 0
    auto go = [&](auto kern, const auto &roots) {
      MR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      MR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
      if (per_sm < 1) fail(MR_ERR_CUDA, "slim gbdt kernel does not fit on an SM (%zu B smem, %d threads)", smem, T);
      per_sm = std::min(per_sm, std::max(1, 1536 / T));  // ~48 warps saturate the shared-memory pipe (profiles/sweep_r1.md)
      const int n_tiles = (p.rows + T - 1) / T;
      static const bool debug = getenv("MR_DEBUG_LAUNCH") != nullptr;
      if (debug)
        fprintf(stderr, "[mr] slim scorer: rows %d tile_cols %d chunks %d x %u B, %d threads x %d CTAs (%d/SM), %zu B smem, cat %d, root table %zu B\n",
                p.rows, F, p.n_chunks, p.chunk_stride, T, std::max(1, std::min(n_tiles, num_sms * per_sm)), per_sm, smem,
                L.has_cat ? (L.cat16 ? 2 : 1) : 0, sizeof(roots) > 16 ? sizeof(roots) : (size_t)0);
      { ProfScope _ps("gbdt_score_slim_kernel", stream); kern<<<std::max(1, std::min(n_tiles, num_sms * per_sm)), T, smem, stream>>>(p, roots); }
      MR_CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    };
    // CAT: 0 none, 1 bitsets of any width (the loop leaves on them), 2 small-categorical form resolved inside the loop
    const int cat = !L.has_cat ? 0 : L.cat16 ? 2 : 1;
    const bool f32 = L.kind == MR_BOOSTER_XGBOOST;  // (XGBoost categorical splits are refused at load time)
    auto with_T = [&](auto tt) {
      constexpr int TT = decltype(tt)::value;
      auto with_roots = [&](auto nr) {
        constexpr int NR = decltype(nr)::value;
        SlimRoots<NR> roots;
        if (NR > 0) memcpy(roots.e, L.h_root_tab, (size_t)L.n_root_tab * 16);
        else roots.e[0] = make_uint4(0, 0, 0, 0);
        if (f32) go(gbdt_score_slim_kernel<float, TT, 0, NR>, roots);
        else if (cat == 2) go(gbdt_score_slim_kernel<double, TT, 2, NR>, roots);
        else if (cat == 0) go(gbdt_score_slim_kernel<double, TT, 0, NR>, roots);
        else if constexpr (NR == 0) go(gbdt_score_slim_kernel<double, TT, 1, 0>, roots);
      };
      // level 0 from a parameter-space root table when the model has one (<= kSlimRootTabMax trees, no wide bitsets)
      const int n_tab = (L.h_root_tab && cat != 1) ? L.n_root_tab : 0;
      if (n_tab > 0 && n_tab <= 512) with_roots(std::integral_constant<int, 512>{});
      else if (n_tab > 0 && n_tab <= kSlimRootTabMax) with_roots(std::integral_constant<int, kSlimRootTabMax>{});
      else with_roots(std::integral_constant<int, 0>{});
    };
    if (T == 512) with_T(std::integral_constant<int, 512>{});
    else if (T == 256) with_T(std::integral_constant<int, 256>{});
    else with_T(std::integral_constant<int, 128>{});
    return;
  }
  // ---- pass 2: traversal
  BParams p;
  p.model = L.d_model; p.chunks = L.d_chunks; p.bins = L.d_bins; p.out = L.d_out;
  p.n_chunks = L.n_chunks; p.chunk_stride = (L.max_chunk_bytes + 127u) & ~127u;
  p.rows = L.rows; p.n_features = F; p.base_score = L.base_score; p.cat16 = L.cat16 ? 1 : 0; p.col_base = 1;
  p.sinks = L.sinks;
  const size_t kMaxSmem = 227 * 1024;
  const bool aligned_tile = L.compact && !L.has_cat && (F & (F - 1)) == 0;  // + slack to align the tile
  const size_t fixed = 128 + (size_t)p.chunk_stride * (L.n_chunks == 1 ? 1 : 2) + (aligned_tile ? (size_t)F * 64 : 0);
  const size_t per_item = (size_t)F * sizeof(uint16_t);
  auto fit_threads = [&](int n_cta) -> int {
    const size_t per_cta = kMaxSmem / (size_t)n_cta, reserve = 1024;
    if (per_cta < fixed + reserve + 32 * per_item) return 0;
    size_t t = (per_cta - fixed - reserve) / std::max<size_t>(per_item, 1);
    t = std::min<size_t>(t, 1024) & ~size_t(31);
    return (int)t;
  };
  int threads = L.threads;
  int want_per_sm = 0;
  if (threads <= 0 && L.compact) {
    // Measured (profiles/sweep_r1.md, fine sweep): time falls with resident warps up to ~48 per SM
    // (24: 2.81 ms, 32: 2.29, 40: 2.17, 48: 2.05 on C2), then the shared-memory pipe is saturated and
    // more warps only queue on it (56: 2.07, 64: 2.22).  Take the CTA shape that gets closest to 48 warps,
    // in CTAs of whole multiples of 4 warps (even over the schedulers); ties go to the shape that wastes
    // the least of the persistent grid's last round.
    double best = -1;
    for (int n_cta = 1; n_cta <= 4; n_cta++) {
      const int cap = std::min(fit_threads(n_cta), 2048 / n_cta);
      for (int t = 128; t <= cap; t += 128) {
        const int warps = n_cta * (t / 32);
        const double occ = warps <= 48 ? warps : 48 - 0.25 * (warps - 48);
        const long long tiles = (L.rows + t - 1) / t, slots = (long long)num_sms * n_cta;
        const double tail = (double)tiles / (double)(((tiles + slots - 1) / slots) * slots);  // 1 = no idle round
        const double score = occ * (0.9 + 0.1 * tail);
        if (score > best) { best = score; threads = t; want_per_sm = n_cta; }
      }
    }
    if (threads <= 0) { threads = std::max(32, fit_threads(1)); want_per_sm = 1; }
    while (threads > 32 && (L.rows + threads - 1) / threads < num_sms) threads = ((threads / 2) + 31) & ~31;
  } else if (threads <= 0) {
    int best_warps = 0;
    threads = 0;
    for (int n_cta = 2; n_cta <= 4; n_cta++) {
      const int t = std::min(fit_threads(n_cta), 2048 / n_cta);
      if (t < 64) continue;
      const int warps = std::min(64, n_cta * (t / 32));
      if (warps > best_warps) { best_warps = warps; threads = t & ~31; }
    }
    if (threads == 0) threads = fit_threads(1);
    while (threads > 32 && (L.rows + threads - 1) / threads < num_sms) threads = ((threads / 2) + 31) & ~31;
  }
  threads = std::max(32, std::min(1024, (threads / 32) * 32));
  while (fixed + (size_t)threads * per_item > kMaxSmem && threads > 32) threads = ((threads / 2) + 31) & ~31;
  if (fixed + (size_t)threads * per_item > kMaxSmem) fail(MR_ERR_UNSUPPORTED, "binned tile does not fit in shared memory");
  const size_t smem = fixed + (size_t)threads * per_item;
  if (L.compact) {
    auto go = [&](auto kern) {
      MR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      MR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
      if (per_sm < 1) fail(MR_ERR_CUDA, "compact gbdt kernel does not fit on an SM");
      const int n_tiles = (p.rows + threads - 1) / threads;
      if (want_per_sm > 0) per_sm = std::min(per_sm, want_per_sm);
      static const bool debug = getenv("MR_DEBUG_LAUNCH") != nullptr;
      if (debug)
        fprintf(stderr, "[mr] compact scorer: rows %d tile_cols %d chunks %d x %u B, %d threads x %d CTAs (%d/SM), %zu B smem\n",
                p.rows, F, p.n_chunks, p.chunk_stride, threads, std::max(1, std::min(n_tiles, num_sms * per_sm)), per_sm, smem);
      { ProfScope _ps("gbdt_score_compact_kernel", stream); kern<<<std::max(1, std::min(n_tiles, num_sms * per_sm)), threads, smem, stream>>>(p); }
      MR_CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    };
    const bool aligned = aligned_tile;
    if (L.kind == MR_BOOSTER_XGBOOST) { if (aligned) go(gbdt_score_compact_kernel<float, false, true>); else go(gbdt_score_compact_kernel<float, false, false>); }
    else if (L.has_cat) go(gbdt_score_compact_kernel<double, true, false>);
    else if (aligned) go(gbdt_score_compact_kernel<double, false, true>);
    else go(gbdt_score_compact_kernel<double, false, false>);
    return;
  }
  // generic binned kernel: the fallback for models the compact layout cannot hold (> 1023 columns, oversized
  // categorical tables); two trees in flight per thread
  if (L.kind == MR_BOOSTER_XGBOOST) launch_b<float, false, 2>(p, threads, smem, num_sms, stream);
  else if (L.has_cat) launch_b<double, true, 2>(p, threads, smem, num_sms, stream);
  else launch_b<double, false, 2>(p, threads, smem, num_sms, stream);
}

}  // namespace mr
