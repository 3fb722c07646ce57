// GBDT ensemble scoring on sm_100a — the device side of Booster.predictMat
// (reference call site S/ml/rank/LambdaMARTRanker.scala:348; LightGBM / XGBoost
// prediction semantics restated in oracle/gbdt_oracle.c).
//
// Mapping.  One THREAD owns one item and walks every tree in tree order, so the score
// is accumulated in exactly the order the CPU libraries use (f64 `+=` per tree for
// LightGBM, f32 for XGBoost): scores are bit-identical to the sequential reference,
// which is what makes the final ordering bit-exact.  A CTA owns a tile of W items:
//   * the tile's features live in shared memory TRANSPOSED, xs[f][item]; thread `tid`
//     reads xs[f*W + tid], so any per-lane feature index is bank-conflict free;
//   * the ensemble is streamed through shared memory in CHUNKS by the TMA engine
//     (cp.async.bulk global->shared, mbarrier complete_tx), double-buffered so the
//     copy of chunk c+1 overlaps the traversal of chunk c; a model that fits in one
//     chunk is staged once per CTA and stays resident across item tiles;
//   * CTAs are persistent: grid = min(#tiles, SMs x resident CTAs), tiles strided.
// Node visit = one 16-byte LDS (node) + one LDS (feature) + compare/select.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "gbdt_kernels.cuh"
#include "tma.cuh"

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

namespace mr {

long long g_kernel_launches = 0;

// ------------------------------------------------------------------ per-kernel profile (bench.py)
namespace {
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
struct ProfRec { const char *name; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof;
}  // namespace

bool profile_active() { return g_prof_on.load(std::memory_order_relaxed); }

ProfScope::ProfScope(const char *n, cudaStream_t s) : name(n), stream(s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&e0) != cudaSuccess) { e0 = nullptr; return; }
  cudaEventRecord(e0, stream);
}
ProfScope::~ProfScope() {
  if (!e0) return;
  cudaEvent_t e1 = nullptr;
  if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return; }
  cudaEventRecord(e1, stream);
  std::lock_guard<std::mutex> g(g_prof_mu);
  g_prof.push_back(ProfRec{name, e0, e1});
}
void profile_begin() {
  std::lock_guard<std::mutex> g(g_prof_mu);
  for (auto &r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_on.store(true);
}
void profile_end(std::string &out) {
  g_prof_on.store(false);
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> g(g_prof_mu);
  std::vector<std::pair<std::string, std::pair<long long, double>>> agg;
  for (auto &r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    bool found = false;
    for (auto &a : agg)
      if (a.first == r.name) { a.second.first++; a.second.second += ms; found = true; break; }
    if (!found) agg.push_back({r.name, {1, (double)ms}});
  }
  g_prof.clear();
  out = "[";
  for (size_t i = 0; i < agg.size(); i++) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s{\"kernel\": \"%s\", \"launches\": %lld, \"ms\": %.6f}", i ? ", " : "", agg[i].first.c_str(),
             agg[i].second.first, agg[i].second.second);
    out += buf;
  }
  out += "]";
}

namespace {


struct KParams {
  const uint8_t *model;
  const ChunkDesc *chunks;
  const double *values;
  double *out;
  unsigned long long *visited;
  int n_chunks;
  uint32_t chunk_stride;  // bytes reserved per shared-memory chunk buffer
  int rows, cols, n_features;
  float base_score;
  int model_in_global;  // 1: a chunk exceeds shared memory -> nodes are read from HBM/L2 directly
};

template <typename Real> struct Acc;
template <> struct Acc<double> { using type = double; };
template <> struct Acc<float> { using type = float; };

template <typename Real> __device__ __forceinline__ Real cvt_feature(double v);
template <> __device__ __forceinline__ double cvt_feature<double>(double v) { return v; }
// XGBoost reads the matrix as binary32 (DMatrix of floats): round-to-nearest-even once.
template <> __device__ __forceinline__ float cvt_feature<float>(double v) { return __double2float_rn(v); }

// One decision.  nd = {thr.lo, thr.hi, feature|flags<<24, left|right<<16}.
template <typename Real, bool HAS_CAT, bool HAS_ZERO>
__device__ __forceinline__ int step(const uint4 nd, const Real x, const uint8_t *chunk) {
  const uint32_t flags = nd.z >> 24;
  bool left;
  if constexpr (sizeof(Real) == 8) {
    // LightGBM Tree::NumericalDecision / CategoricalDecision
    const double xd = (double)x;
    if (HAS_CAT && (flags & NF_CATEGORICAL)) {
      left = false;
      if (xd == xd) {
        // static_cast<int>(fval); values outside int range behave as negative (x86 cvttsd2si)
        const bool in_range = (xd < 2147483648.0) && (xd > -2147483649.0);
        const int iv = in_range ? __double2int_rz(xd) : -1;
        if (iv >= 0) {
          const uint32_t w = (uint32_t)iv >> 5;
          if (w < nd.y) left = (reinterpret_cast<const uint32_t *>(chunk)[nd.x + w] >> (iv & 31)) & 1u;
        }
      }
    } else {
      const double thr = __hiloint2double((int)nd.y, (int)nd.x);
      left = xd <= thr;  // false when xd is NaN
      if (HAS_ZERO) {
        // missing_type == Zero: |x| <= kZeroThreshold (= (double)1e-35f) takes the default side
        const double kZero = (double)1e-35f;
        if (((flags >> NF_MISSING_SHIFT) & 3u) == 1u && xd >= -kZero && xd <= kZero)
          left = (flags & NF_DEFAULT_LEFT) != 0;
      }
      if (xd != xd) left = (flags & NF_NAN_LEFT) != 0;
    }
  } else {
    // XGBoost: fvalue < split_cond goes left; missing (NaN) takes default_left
    const float thr = __uint_as_float(nd.x);
    left = x < thr;
    if (x != x) left = (flags & NF_NAN_LEFT) != 0;
  }
  const int l = (int)(short)(nd.w & 0xFFFFu), r = (int)(short)(nd.w >> 16);
  return left ? l : r;
}

// Lock-step: every lane walks tree t at the same time (ILP trees in flight); node loads of a warp stay inside
// one tree (mostly broadcast / conflict free).  A free-running variant (a lane moves on to its next tree as soon as
// it hits a leaf) lost to shared-memory bank conflicts and was removed (profiles/sweep_r1.md).
template <typename Real, bool HAS_CAT, bool HAS_ZERO, int ILP, bool STAGE, bool COUNT>
__global__ void __launch_bounds__(1024) gbdt_score_kernel(const KParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  using AccT = typename Acc<Real>::type;
  const int W = blockDim.x;
  const int tid = threadIdx.x;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  const bool resident = p.n_chunks == 1;
  uint8_t *cbuf0 = smem + 128;
  uint8_t *cbuf1 = cbuf0 + (resident ? 0u : p.chunk_stride);
  Real *xs = reinterpret_cast<Real *>(cbuf1 + p.chunk_stride);

  const int n_tiles = (p.rows + W - 1) / W;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < n_tiles && !p.model_in_global) {
    const ChunkDesc cd = p.chunks[0];
    mbar_arrive_expect_tx(&bars[0], cd.bytes);
    tma_bulk_g2s(cbuf0, p.model + cd.byte_off, cd.bytes, &bars[0]);
  }

  uint32_t it = 0;
  unsigned long long visited = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int item = tile * W + tid;
    const double *row = p.values + (size_t)item * p.cols;
    if (STAGE) {
      // transposed feature tile: xs[f*W + tid]
      if (item < p.rows) {
        if ((p.cols & 1) == 0) {
          const double2 *row2 = reinterpret_cast<const double2 *>(row);
          int f = 0;
          for (; f + 1 < p.n_features; f += 2) {
            const double2 v = __ldg(row2 + (f >> 1));
            xs[f * W + tid] = cvt_feature<Real>(v.x);
            xs[(f + 1) * W + tid] = cvt_feature<Real>(v.y);
          }
          if (f < p.n_features) xs[f * W + tid] = cvt_feature<Real>(__ldg(row + f));
        } else {
          for (int f = 0; f < p.n_features; f++) xs[f * W + tid] = cvt_feature<Real>(__ldg(row + f));
        }
      } else {
        for (int f = 0; f < p.n_features; f++) xs[f * W + tid] = (Real)0;
      }
    }
    __syncthreads();

    AccT acc = (sizeof(Real) == 4) ? (AccT)p.base_score : (AccT)0;
    for (int c = 0; c < p.n_chunks; ++c, ++it) {
      if (!resident && tid == 0 && !p.model_in_global) {
        const bool more = (c + 1 < p.n_chunks) || (tile + (int)gridDim.x < n_tiles);
        if (more) {
          const int nc = (c + 1 < p.n_chunks) ? c + 1 : 0;
          const ChunkDesc cd = p.chunks[nc];
          uint64_t *bar = &bars[(it + 1) & 1];
          fence_proxy_async();  // generic-proxy reads of that buffer (iteration it-1) precede the async write
          mbar_arrive_expect_tx(bar, cd.bytes);
          tma_bulk_g2s(((it + 1) & 1) ? cbuf1 : cbuf0, p.model + cd.byte_off, cd.bytes, bar);
        }
      }
      if (!p.model_in_global && (!resident || it == 0)) mbar_wait(&bars[it & 1], (it >> 1) & 1);
      const uint8_t *cb = p.model_in_global ? p.model + p.chunks[c].byte_off
                                            : ((!resident && (it & 1)) ? cbuf1 : cbuf0);
      const int ntree = (int)*reinterpret_cast<const uint32_t *>(cb);
      const uint2 *tab = reinterpret_cast<const uint2 *>(cb + 16);

      auto feat = [&](uint32_t f) -> Real {
        if (STAGE) return xs[f * W + tid];
        return (item < p.rows) ? cvt_feature<Real>(__ldg(row + f)) : (Real)0;
      };

      {
        int t = 0;
        for (; t + ILP <= ntree; t += ILP) {
          const uint4 *nodes[ILP];
          const Real *leaves[ILP];
          int n[ILP];
#pragma unroll
          for (int k = 0; k < ILP; k++) {
            const uint2 to = tab[t + k];
            nodes[k] = reinterpret_cast<const uint4 *>(cb + to.x);
            leaves[k] = reinterpret_cast<const Real *>(cb + to.y);
            n[k] = 0;
          }
          bool any = true;
          while (any) {
            // all ILP node loads, then all feature loads, are issued before any is consumed;
            // a tree that already reached its leaf re-reads its root (harmless) and keeps n
            uint4 nd[ILP];
            Real x[ILP];
#pragma unroll
            for (int k = 0; k < ILP; k++) nd[k] = nodes[k][n[k] < 0 ? 0 : n[k]];
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = feat(nd[k].z & 0xFFFFFFu);
            any = false;
#pragma unroll
            for (int k = 0; k < ILP; k++) {
              const int nx = step<Real, HAS_CAT, HAS_ZERO>(nd[k], x[k], cb);
              if (COUNT) visited += n[k] >= 0;
              n[k] = n[k] >= 0 ? nx : n[k];
              any |= n[k] >= 0;
            }
          }
#pragma unroll
          for (int k = 0; k < ILP; k++) acc += (AccT)leaves[k][~n[k]];
        }
        for (; t < ntree; t++) {
          const uint2 to = tab[t];
          const uint4 *nodes = reinterpret_cast<const uint4 *>(cb + to.x);
          const Real *leaves = reinterpret_cast<const Real *>(cb + to.y);
          int n = 0;
          do {
            const uint4 nd = nodes[n];
            n = step<Real, HAS_CAT, HAS_ZERO>(nd, feat(nd.z & 0xFFFFFFu), cb);
            if (COUNT) visited++;
          } while (n >= 0);
          acc += (AccT)leaves[~n];
        }
      }
      __syncthreads();  // the buffer may be overwritten by the next prefetch; xs by the next tile
    }
    if (item < p.rows) {
      p.out[item] = (double)acc;
      if (COUNT) atomicAdd(p.visited, visited);
    }
    visited = 0;
  }
}

template <typename Real, bool HAS_CAT, bool HAS_ZERO, int ILP, bool STAGE, bool COUNT>
void launch_inst(const KParams &p, int threads, size_t smem, int num_sms, cudaStream_t stream) {
  auto kern = gbdt_score_kernel<Real, HAS_CAT, HAS_ZERO, ILP, STAGE, COUNT>;
  MR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) fail(MR_ERR_CUDA, "gbdt_score kernel does not fit on an SM (smem %zu, threads %d)", smem, threads);
  const int n_tiles = (p.rows + threads - 1) / threads;
  const int grid = std::max(1, std::min(n_tiles, num_sms * per_sm));
  { ProfScope _ps("gbdt_score_kernel", stream); kern<<<grid, threads, smem, stream>>>(p); }
  MR_CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

// Instantiations kept: trees in flight per thread 2 (narrow rows) or 4 (wide rows: fewer resident warps, more ILP);
// the path-counting build exists once per precision (it only serves mr_model_count_path).
template <typename Real, bool HAS_CAT, bool HAS_ZERO, bool STAGE>
void launch_ilp(const KParams &p, int ilp, bool count, int threads, size_t smem, int num_sms, cudaStream_t s) {
  if (count) return launch_inst<Real, HAS_CAT, HAS_ZERO, 2, STAGE, true>(p, threads, smem, num_sms, s);
  if (ilp >= 4) return launch_inst<Real, HAS_CAT, HAS_ZERO, 4, STAGE, false>(p, threads, smem, num_sms, s);
  return launch_inst<Real, HAS_CAT, HAS_ZERO, 2, STAGE, false>(p, threads, smem, num_sms, s);
}

template <typename Real, bool STAGE>
void launch_flags(const KParams &p, bool cat, bool zero, int ilp, bool count, int threads, size_t smem, int num_sms,
                  cudaStream_t s) {
  if constexpr (sizeof(Real) == 4) {
    return launch_ilp<Real, false, false, STAGE>(p, ilp, count, threads, smem, num_sms, s);
  } else {
    if (cat || zero) return launch_ilp<Real, true, true, STAGE>(p, ilp, count, threads, smem, num_sms, s);
    return launch_ilp<Real, false, false, STAGE>(p, ilp, count, threads, smem, num_sms, s);
  }
}

}  // namespace

void launch_gbdt_score(const ScoreLaunch &L, int num_sms, cudaStream_t stream) {
  if (L.rows <= 0) return;
  KParams p;
  p.model = L.d_model;
  p.chunks = L.d_chunks;
  p.values = L.d_values;
  p.out = L.d_out;
  p.visited = L.d_visited;
  p.n_chunks = L.n_chunks;
  p.chunk_stride = (L.max_chunk_bytes + 127u) & ~127u;
  p.rows = L.rows;
  p.cols = L.cols;
  p.n_features = L.n_features;
  p.base_score = L.base_score;
  p.model_in_global = 0;

  const bool f32 = L.kind == MR_BOOSTER_XGBOOST;
  const size_t real_sz = f32 ? 4 : 8;
  const size_t kMaxSmem = 227 * 1024;
  size_t fixed = 128 + (size_t)p.chunk_stride * (L.n_chunks == 1 ? 1 : 2);
  if (fixed + 32 * (size_t)L.n_features * real_sz > kMaxSmem) {
    // a chunk (= one enormous tree) does not leave room for even a 32-item tile: leave the ensemble
    // in HBM/L2 and walk it with ordinary loads (slow path; keeps every legal model scorable)
    p.model_in_global = 1;
    p.chunk_stride = 0;
    fixed = 128;
  }

  // Occupancy is what this kernel lives on (every node visit is a dependent
  // LDS -> LDS -> compare chain), and shared memory is what limits it: each item in
  // flight pins n_features * sizeof(Real) bytes of tile.  Pick the CTA shape that
  // keeps the most warps resident per SM.
  const size_t per_item = (size_t)L.n_features * real_sz;
  auto fit_threads = [&](int n_cta) -> int {
    const size_t per_cta = kMaxSmem / (size_t)n_cta;
    const size_t reserve = 1024;  // driver-reserved shared memory per CTA
    if (per_cta < fixed + reserve + 32 * per_item) return 0;
    size_t t = (per_cta - fixed - reserve) / std::max<size_t>(per_item, 1);
    t = std::min<size_t>(t, 1024) & ~size_t(31);
    return (int)t;
  };
  int threads = L.threads;
  if (threads <= 0) {
    // Measured on B200 (profiles/sweep_r1.md): two or three mid-sized CTAs per SM beat one huge
    // CTA with more warps, because every chunk ends in a CTA-wide barrier and a second CTA
    // fills the bubble.  So: the most warps among >= 2 CTAs/SM, one CTA/SM only if nothing else fits.
    int best_warps = 0;
    threads = 0;
    for (int n_cta = 2; n_cta <= 4; n_cta++) {
      const int t = fit_threads(n_cta);
      if (t < 64) continue;
      const int warps = std::min(64, n_cta * (t / 32));
      if (warps > best_warps) {
        best_warps = warps;
        threads = t;
      }
    }
    if (threads == 0) threads = fit_threads(1);
    if (threads == 0) threads = 32;
    // small batches: shrink the tile until the persistent grid covers the chip
    while (threads > 32 && (L.rows + threads - 1) / threads < num_sms) threads = ((threads / 2) + 31) & ~31;
  }
  threads = std::max(32, std::min(1024, (threads / 32) * 32));
  bool stage = true;
  while (fixed + (size_t)threads * per_item > kMaxSmem && threads > 32) threads = ((threads / 2) + 31) & ~31;
  if (fixed + (size_t)threads * per_item > kMaxSmem) stage = false;  // very wide rows: read HBM/L1 directly
  const size_t smem = fixed + (stage ? (size_t)threads * per_item : 0);

  const int ilp = L.ilp <= 0 ? (per_item > 384 ? 4 : 2) : L.ilp;
  const bool count = L.d_visited != nullptr;
  if (f32) {
    if (stage) launch_flags<float, true>(p, L.has_cat, L.has_zero, ilp, count, threads, smem, num_sms, stream);
    else launch_flags<float, false>(p, L.has_cat, L.has_zero, ilp, count, threads, smem, num_sms, stream);
  } else {
    if (stage) launch_flags<double, true>(p, L.has_cat, L.has_zero, ilp, count, threads, smem, num_sms, stream);
    else launch_flags<double, false>(p, L.has_cat, L.has_zero, ilp, count, threads, smem, num_sms, stream);
  }
}

}  // namespace mr
