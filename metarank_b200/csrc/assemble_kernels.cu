// Feature-vector assembly on sm_100a: the device side of
//   FeatureValueLoader.fromStateBackend  (S/fstore/FeatureValueLoader.scala:11-25)
//   ItemValue.fromState                  (S/model/ItemValue.scala:25-72)
//   ClickthroughQuery.collectFeatureValues (S/flow/ClickthroughQuery.scala:50-74)
// and of each hot extractor's value()/values().
//
// Kernels (all HBM gather/scan work; no tensor cores anywhere on this path):
//   lookup_kernel   thread per item: request index, hash probe item-id -> item row;
//                   thread per request: user/session rows
//   cosine_kernel   thread per item: bi-encoder cosine in the reference's exact operation order
//   prepass_kernel  CTA per request: sorted tag multisets (interacted_with, diversity strings),
//                   diversity median, cosine min-max / position normalisation
//   assemble_kernel thread per item, uniform loop over the extractor plan -> dense f64 row
//   order_kernel    CTA per request: stable descending rank of the scores
// Java arithmetic is strict IEEE (no fused multiply-add), so every a*b+c below is spelled
// with __dmul_rn/__dadd_rn to keep nvcc from contracting it.
#include "assemble_kernels.cuh"

#include <algorithm>

#include "gbdt_kernels.cuh"  // g_kernel_launches

namespace mr {
namespace {

constexpr uint32_t kNoRow = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t probe(const DTable &t, uint64_t key) {
  if (t.mask == 0 || t.keys == nullptr) return kNoRow;
  if (key == 0) key = 1;
  uint32_t i = (uint32_t)mix64(key) & t.mask;
  for (;;) {
    const uint64_t k = __ldg(t.keys + i);
    if (k == key) return __ldg(t.vals + i);
    if (k == 0) return kNoRow;
    i = (i + 1) & t.mask;
  }
}

__device__ __forceinline__ uint64_t d_hash_combine(uint64_t a, uint64_t b) {
  uint64_t x = a ^ (b + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2));
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
  return x ? x : 1;
}

__device__ __forceinline__ const uint64_t *row_ptr(const DTable &t, uint32_t row) {
  return t.rows + (size_t)row * t.row_words;
}
__device__ __forceinline__ bool present(const uint64_t *row, int bit) {
  return (__ldg(row + (bit >> 6)) >> (bit & 63)) & 1ull;
}
__device__ __forceinline__ double nan_d() { return __longlong_as_double(0x7FF8000000000000ll); }

// java.lang.Double.compare as a sortable key (NaN largest, -0.0 < 0.0)
__device__ __forceinline__ long long total_order_key(double x) {
  if (x != x) return 0x7FFFFFFFFFFFFFFFll;
  long long b = __double_as_longlong(x);
  return b < 0 ? (b ^ 0x7FFFFFFFFFFFFFFFll) : b;
}

// ------------------------------------------------------------------ lookup
__global__ void lookup_kernel(RankArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.total_items) {
    // owning request: last r with off[r] <= i
    int lo = 0, hi = a.n_requests;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(a.item_offsets + mid) <= i) lo = mid; else hi = mid;
    }
    a.item_req[i] = lo;
    a.item_row[i] = probe(a.st.t[SC_ITEM], __ldg(a.item_ids + i));
  }
  if (i < a.n_requests) {
    const uint64_t u = a.user_ids ? __ldg(a.user_ids + i) : 0, s = a.session_ids ? __ldg(a.session_ids + i) : 0;
    a.visitor_row[2 * i] = u ? probe(a.st.t[SC_USER], u) : kNoRow;
    a.visitor_row[2 * i + 1] = s ? probe(a.st.t[SC_SESSION], s) : kNoRow;
  }
  if (i == 0) {  // first kernel of every rank call: reset the per-call device state
    *a.hist_cursor = 0;
    *a.error_flag = 0;
  }
}

// ------------------------------------------------------------------ cosine
// CosineDistance.dist (S/ml/onnx/distance/DistanceFunction.scala:13-27): sequential sums,
// query(i)*query(i) is a FLOAT product, the other two are double products.
__global__ void __launch_bounds__(128) cosine_kernel(RankArgs a) {
  // A warp owns 32 items.  Embedding rows are streamed 32 dims at a time with coalesced 256-byte row
  // segments into a padded shared-memory tile, then every lane walks ITS item's 32 values in index
  // order, so the three sums are accumulated in exactly the reference's sequence.
  __shared__ double s_tile[4][32][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < a.total_items;
  const int r = live ? a.item_req[i] : 0;
  const uint32_t ir = live ? a.item_row[i] : kNoRow;
  int voff = 0;
  for (int f = 0; f < a.n_plan; f++) {
    const DFeature d = a.plan[f];
    if (d.kind != FK_COSINE) continue;
    const int dim = d.aux0;
    const bool q_ok = live && a.req_vec_present && a.req_vec_present[(size_t)r * a.n_req_vec + d.in0];
    const bool ok = q_ok && ir != kNoRow && present(row_ptr(a.st.t[SC_ITEM], ir), d.b[0]);
    const float *q = a.req_vec + (size_t)r * a.vec_stride + voff;
    const double *side = a.st.side[(int)d.uparam];
    if (!side) { voff += dim; continue; }  // stored as binary32: cosine_f32_kernel's
    double top = 0.0, as = 0.0, bs = 0.0;
    for (int d0 = 0; d0 < dim; d0 += 32) {
      const int nd = min(32, dim - d0);
      // cooperative, coalesced: row segment of item j of this warp -> s_tile[warp][j][*]
      for (int j = 0; j < 32; j++) {
        const uint32_t rj = __shfl_sync(0xFFFFFFFFu, ir, j);
        const int okj = __shfl_sync(0xFFFFFFFFu, (int)ok, j);
        if (okj && lane < nd) s_tile[warp][j][lane] = __ldg(side + (size_t)rj * dim + d0 + lane);
      }
      __syncwarp();
      if (ok) {
        for (int k = 0; k < nd; k++) {
          const float qk = __ldg(q + d0 + k);
          const double ek = s_tile[warp][lane][k];
          top = __dadd_rn(top, __dmul_rn((double)qk, ek));
          as = __dadd_rn(as, (double)__fmul_rn(qk, qk));
          bs = __dadd_rn(bs, __dmul_rn(ek, ek));
        }
      }
      __syncwarp();
    }
    if (live) {
      const double out = ok ? __ddiv_rn(top, __dmul_rn(__dsqrt_rn(as), __dsqrt_rn(bs))) : nan_d();
      a.cos[(size_t)d.aux2 * a.total_items + i] = out;
      a.cos[(size_t)(a.n_cos + d.aux2) * a.total_items + i] = out;  // normalised copy (noop default)
    }
    voff += dim;
  }
}

// CosineDistance's aSum (S/ml/onnx/distance/DistanceFunction.scala:21) is a function of the request alone: once per
// (request, feature) instead of once per item.  query(i) * query(i) is a FLOAT product, widened and added in index order.
// A warp per request: the query is staged coalesced, lane 0 walks it.
__global__ void __launch_bounds__(128) cosine_qnorm_kernel(RankArgs a) {
  extern __shared__ float s_qn[];  // [4 warps][vec_stride]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + warp;
  if (r >= a.n_requests) return;
  float *q = s_qn + (size_t)warp * a.vec_stride;
  const float *src = a.req_vec + (size_t)r * a.vec_stride;
  for (int k = lane; k < a.vec_stride; k += 32) q[k] = __ldg(src + k);
  __syncwarp();
  if (lane != 0) return;
  int voff = 0;
  for (int f = 0; f < a.n_plan; f++) {
    const DFeature d = a.plan[f];
    if (d.kind != FK_COSINE) continue;
    double as = 0.0;
    for (int k = 0; k < d.aux0; k++) as = __dadd_rn(as, (double)__fmul_rn(q[voff + k], q[voff + k]));
    a.qnorm[(size_t)r * a.n_cos + d.aux2] = as;
    voff += d.aux0;
  }
}

// Embeddings stored as binary32 (DState::side_f32): the kernel the bi-encoder configs run.  HBM-bound — an item costs
// its 4 * dim bytes and nothing else: aSum comes from lookup_kernel (per request), bSum from the state (per item, computed
// at upsert time), so only topSum = sum of (double)q[k] * (double)e[k] in index order is left, one DMUL + one DADD per element
// (products rounded before the add, like the JVM).  A warp owns 32 items: their rows stream chunk by chunk through cp.async
// (16 bytes per lane, a row segment per instruction, 32 rows in flight) into a padded tile, then lane j walks ITS item's
// values in order.  The query is widened once per CTA into shared memory.
constexpr int kCosChunk = 32;             // floats per row per stage (small stages: 5 CTAs = 20 warps per SM hide the L2 latency)
constexpr int kCosPitch = kCosChunk + 4;  // 144-byte rows: 8 consecutive lanes read 16 bytes from 8 different bank groups
constexpr int kCosStages = 2;

__global__ void __launch_bounds__(128) cosine_f32_kernel(RankArgs a, int dim_max) {
  extern __shared__ __align__(16) uint8_t s_cos_raw[];
  double *s_q = reinterpret_cast<double *>(s_cos_raw);  // [dim_max] the CTA's first request's query, widened
  float *s_tiles = reinterpret_cast<float *>(s_cos_raw + (((size_t)dim_max * 8 + 15) & ~size_t(15)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *tile = s_tiles + (size_t)warp * kCosStages * 32 * kCosPitch;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < a.total_items;
  const int r = live ? a.item_req[i] : 0;
  const int r0 = a.item_req[min(blockIdx.x * blockDim.x, (unsigned)a.total_items - 1u)];  // CTA-uniform
  const uint32_t ir = live ? a.item_row[i] : kNoRow;
  int voff = 0;
  for (int f = 0; f < a.n_plan; f++) {
    const DFeature d = a.plan[f];
    if (d.kind != FK_COSINE) continue;
    const int dim = d.aux0;
    const float *side = a.st.side_f32[(int)d.uparam];
    if (!side) { voff += dim; continue; }  // stored as f64: cosine_kernel's
    // the rows do not wait for anything but the item's row index: start the first chunk before the query is staged and
    // before the presence bits are read (a row without an embedding holds zeros; whether it counts is decided below)
    const int n_chunks = (dim + kCosChunk - 1) / kCosChunk;
    auto issue = [&](int c) {
      if (c < n_chunks) {
        const int d0 = c * kCosChunk, nd = min(kCosChunk, dim - d0);  // a multiple of 4 (f32 mode requires dim % 4 == 0)
        float *st = tile + (size_t)(c % kCosStages) * 32 * kCosPitch;
        // 8 lanes cover a row segment of 32 floats: four rows per instruction
        const int quad = lane >> 3, l8 = lane & 7;
        for (int j = 0; j < 32; j += 4) {
          const uint32_t rj = __shfl_sync(0xFFFFFFFFu, ir, j + quad);
          if (rj != kNoRow && l8 * 4 < nd)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(st + (j + quad) * kCosPitch + l8 * 4)),
                         "l"(side + (size_t)rj * dim + d0 + l8 * 4)
                         : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    __syncthreads();  // s_q of the previous feature is no longer read
    {
      const float *q0 = a.req_vec + (size_t)r0 * a.vec_stride + voff;
      for (int k = threadIdx.x; k < dim; k += blockDim.x) s_q[k] = (double)__ldg(q0 + k);
    }
    __syncthreads();
    const bool q_ok = live && a.req_vec_present && a.req_vec_present[(size_t)r * a.n_req_vec + d.in0];
    const bool ok = q_ok && ir != kNoRow && present(row_ptr(a.st.t[SC_ITEM], ir), d.b[0]);
    const float *q = a.req_vec + (size_t)r * a.vec_stride + voff;
    const bool same_q = r == r0;
    // request boundaries inside a warp are rare (a CTA is 128 consecutive items): the whole warp takes the loop that
    // reads the widened query from shared memory unless one of its lanes belongs to another request
    const bool warp_same = __all_sync(0xFFFFFFFFu, same_q || !ok);
    double top = 0.0;
    for (int c = 0; c < n_chunks; c++) {
      issue(c + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");  // chunk c has landed (chunk c + 1 may still be in flight)
      __syncwarp();
      if (ok) {
        const int d0 = c * kCosChunk, nd = min(kCosChunk, dim - d0);
        const float4 *row = reinterpret_cast<const float4 *>(tile + (size_t)(c % kCosStages) * 32 * kCosPitch + lane * kCosPitch);
        if (warp_same) {
#pragma unroll 2
          for (int k4 = 0; k4 < nd / 4; k4++) {
            const float4 e = row[k4];
            const double *qd = s_q + d0 + k4 * 4;
            const double p0 = __dmul_rn(qd[0], (double)e.x), p1 = __dmul_rn(qd[1], (double)e.y);
            const double p2 = __dmul_rn(qd[2], (double)e.z), p3 = __dmul_rn(qd[3], (double)e.w);
            top = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(top, p0), p1), p2), p3);
          }
        } else {
          for (int k4 = 0; k4 < nd / 4; k4++) {
            const float4 e = row[k4];
            const int k = d0 + k4 * 4;
            const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
              const double qd = same_q ? s_q[k + c4] : (double)__ldg(q + k + c4);
              top = __dadd_rn(top, __dmul_rn(qd, (double)ev[c4]));
            }
          }
        }
      }
      __syncwarp();  // before chunk c + 2 overwrites this stage
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live) {
      double out = nan_d();
      if (ok) {
        const double as = a.qnorm[(size_t)r * a.n_cos + d.aux2], bs = __ldg(a.st.side_bs[(int)d.uparam] + ir);
        out = __ddiv_rn(top, __dmul_rn(__dsqrt_rn(as), __dsqrt_rn(bs)));
      }
      a.cos[(size_t)d.aux2 * a.total_items + i] = out;
      a.cos[(size_t)(a.n_cos + d.aux2) * a.total_items + i] = out;  // normalised copy (noop default)
    }
    voff += dim;
  }
}

// ------------------------------------------------------------------ per-request prepass
// Reserves `n` pool entries for histogram (r,h); returns the offset or 0xFFFFFFFF on overflow.
__device__ uint32_t hist_alloc(const RankArgs &a, uint32_t n) {
  __shared__ uint32_t s_off;
  if (threadIdx.x == 0) {
    uint32_t off = atomicAdd(a.hist_cursor, n);
    if (off + n > a.hist_pool_cap || off + n < off) {
      atomicExch(a.error_flag, -1);
      off = 0xFFFFFFFFu;
    }
    s_off = off;
  }
  __syncthreads();
  const uint32_t o = s_off;
  __syncthreads();
  return o;
}

// A request's tag multiset is READ once per (item, tag) — 100-1000 items x a handful of tags — so it is stored as an
// open-addressing table {tag, count}: a lookup is one or two 16-byte loads instead of the 2 log2(n) dependent 8-byte
// loads of two binary searches over the sorted multiset (ncu_r2_asm: those searches were 35 % of assemble_kernel's stall
// samples, and the kernel's time moves with the NUMBER of divergent loads, not with occupancy).  Layout in the pool, from
// a 16-byte aligned start: `slots` entries (slots = a power of two >= 2 n, so it is never more than half full) and one
// word counting tags equal to the empty marker itself.  Tags are inserted straight from the item rows (global atomics:
// a few hundred per request), the multiset itself is never materialised.
constexpr uint64_t kHistEmpty = ~0ull;
__host__ __device__ inline uint32_t hist_slots(uint32_t n) {
  uint32_t s = 2;
  while (s < 2 * n) s <<= 1;
  return s;
}
__host__ __device__ inline uint32_t hist_words(uint32_t n) { return 1 + 2 * hist_slots(n) + 2; }

// Reserves and clears the table for a multiset of n tags (CTA-collective); returns its first word or 0xFFFFFFFF.
__device__ uint32_t hist_begin(const RankArgs &a, uint32_t n) {
  const uint32_t off = hist_alloc(a, hist_words(n));
  if (off == 0xFFFFFFFFu) return off;
  const uint32_t slots = hist_slots(n), toff = (off + 1) & ~1u;
  unsigned long long *tab = reinterpret_cast<unsigned long long *>(a.hist_pool + toff);
  for (uint32_t i = threadIdx.x; i < slots; i += blockDim.x) { tab[2 * i] = kHistEmpty; tab[2 * i + 1] = 0; }
  if (threadIdx.x == 0) tab[2 * slots] = 0;
  __syncthreads();
  return toff;
}
__device__ __forceinline__ void hist_insert(const RankArgs &a, uint32_t toff, uint32_t slots, unsigned long long tag) {
  unsigned long long *tab = reinterpret_cast<unsigned long long *>(a.hist_pool + toff);
  if (tag == kHistEmpty) { atomicAdd(&tab[2 * slots], 1ull); return; }
  const uint32_t mask = slots - 1;
  for (uint32_t sl = (uint32_t)mix64(tag) & mask;; sl = (sl + 1) & mask) {
    const unsigned long long prev = atomicCAS(&tab[2 * sl], kHistEmpty, tag);
    if (prev == kHistEmpty || prev == tag) { atomicAdd(&tab[2 * sl + 1], 1ull); return; }
  }
}
__device__ void hist_end(const RankArgs &a, int r, int h, uint32_t toff, uint32_t slots) {
  __syncthreads();
  if (threadIdx.x == 0) a.hist_desc[(size_t)r * a.n_hist + h] = make_uint2(toff, slots);
}

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_scan, uint32_t *total) {
  // exclusive scan over blockDim.x (<= 1024) values
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_scan[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < (int)((blockDim.x + 31) >> 5) ? s_scan[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
      if (lane >= o) w += y;
    }
    s_scan[lane] = w;
  }
  __syncthreads();
  const uint32_t base = wid ? s_scan[wid - 1] : 0;
  *total = s_scan[((blockDim.x + 31) >> 5) - 1];
  __syncthreads();
  return base + x - v;
}

__global__ void __launch_bounds__(256) prepass_kernel(RankArgs a) {
  __shared__ uint32_t s_scan[32];
  __shared__ int s_i[4];
  __shared__ double s_d[4];
  const int r = blockIdx.x;
  const int i0 = a.item_offsets[r], n_items = a.item_offsets[r + 1] - i0;
  const DTable &IT = a.st.t[SC_ITEM];

  // grid.y enumerates the plan entries that need a per-request aggregate: one CTA per (request, entry)
  int f_sel = -1;
  {
    int seen = 0;
    for (int f = 0; f < a.n_plan; f++) {
      const int k = a.plan[f].kind;
      const bool agg = k == FK_INTERACTED || k == FK_DIVERSITY || (k == FK_COSINE && a.plan[f].aux1 != 0);
      if (agg && seen++ == (int)blockIdx.y) { f_sel = f; break; }
    }
  }
  if (f_sel < 0) return;
  for (int f = f_sel; f <= f_sel; f++) {
    const DFeature d = a.plan[f];
    if (d.kind == FK_INTERACTED) {
      // visitor's interacted items -> multiset of their values of this field
      // (InteractedWithFeature.values :134-147)
      const int tb = d.scope;  // SC_USER | SC_SESSION
      const uint32_t vr = a.visitor_row[2 * r + (tb == SC_SESSION ? 1 : 0)];
      uint32_t hoff = 0, hlen = 0;
      const uint64_t *hist_items = nullptr;
      if (vr != kNoRow) {
        const uint64_t *vrow = row_ptr(a.st.t[tb], vr);
        if (present(vrow, d.b[1])) {
          const uint64_t desc = vrow[d.w[1]];
          hoff = (uint32_t)desc;
          hlen = (uint32_t)(desc >> 32);
          hist_items = a.st.t[tb].pool + hoff;
        }
      }
      // pass 1: count, pass 2: fill (chunks of blockDim history items)
      uint32_t total = 0;
      for (uint32_t base = 0; base < hlen; base += blockDim.x) {
        const uint32_t j = base + threadIdx.x;
        uint32_t cnt = 0;
        if (j < hlen) {
          const uint32_t row = probe(IT, hist_items[j]);
          if (row != kNoRow) {
            const uint64_t *rp = row_ptr(IT, row);
            if (present(rp, d.b[0])) cnt = (uint32_t)(rp[d.w[0]] >> 32);
          }
        }
        uint32_t chunk_total;
        block_excl_scan(cnt, s_scan, &chunk_total);
        total += chunk_total;
      }
      const uint32_t off = total ? hist_begin(a, total) : 0, slots = hist_slots(total);
      if (off == 0xFFFFFFFFu || total == 0) { if (threadIdx.x == 0) a.hist_desc[(size_t)r * a.n_hist + d.aux0] = make_uint2(0, 0); continue; }
      for (uint32_t base = 0; base < hlen; base += blockDim.x) {
        const uint32_t j = base + threadIdx.x;
        uint32_t cnt = 0, src = 0;
        if (j < hlen) {
          const uint32_t row = probe(IT, hist_items[j]);
          if (row != kNoRow) {
            const uint64_t *rp = row_ptr(IT, row);
            if (present(rp, d.b[0])) { const uint64_t desc = rp[d.w[0]]; cnt = (uint32_t)(desc >> 32); src = (uint32_t)desc; }
          }
        }
        for (uint32_t k = 0; k < cnt; k++) hist_insert(a, off, slots, IT.pool[src + k]);
      }
      hist_end(a, r, d.aux0, off, slots);
    } else if (d.kind == FK_DIVERSITY) {
      // DiversityFeature.values (:67-130): items WITH state, in request order; the first one's
      // scalar type selects the mode; aggregates over the first `top` items of that type.
      double *agg = a.reqagg + ((size_t)r * a.n_reqagg + d.aux2) * 4;
      // head = first item with state
      if (threadIdx.x == 0) { s_i[0] = 0x7FFFFFFF; }
      __syncthreads();
      for (int base = 0; base < n_items; base += blockDim.x) {
        const int j = base + threadIdx.x;
        if (j < n_items) {
          const uint32_t row = a.item_row[i0 + j];
          if (row != kNoRow && present(row_ptr(IT, row), d.b[0])) atomicMin(&s_i[0], j);
        }
        __syncthreads();
        const int seen = s_i[0];
        __syncthreads();  // everyone has read s_i[0] before the next chunk's atomicMin
        if (seen != 0x7FFFFFFF) break;
      }
      const int head = s_i[0];
      __syncthreads();
      int mode = 0;
      if (head != 0x7FFFFFFF) {
        const uint64_t kd = row_ptr(IT, a.item_row[i0 + head])[d.w[0]];
        mode = kd == 1 ? 1 : kd == 2 ? 2 : 0;
      }
      const int top = d.aux0;
      if (mode == 0) {
        if (threadIdx.x == 0) { agg[0] = 0.0; agg[1] = 0.0; agg[2] = 0.0; a.hist_desc[(size_t)r * a.n_hist + d.aux1] = make_uint2(0, 0); }
        continue;
      }
      // ordinal of every matching item (prefix count in request order); selected iff ordinal < top
      if (mode == 1) {
        // collect the selected doubles into the pool (as bits), then LEGACY percentile(50)
        uint32_t n_sel = 0;
        {
          uint32_t run = 0;
          // count first
          for (int base = 0; base < n_items && run < (uint32_t)top; base += blockDim.x) {
            const int j = base + threadIdx.x;
            uint32_t m = 0;
            if (j < n_items) {
              const uint32_t row = a.item_row[i0 + j];
              if (row != kNoRow) { const uint64_t *rp = row_ptr(IT, row); m = present(rp, d.b[0]) && rp[d.w[0]] == 1; }
            }
            uint32_t ct;
            block_excl_scan(m, s_scan, &ct);
            run += ct;
          }
          n_sel = run < (uint32_t)top ? run : (uint32_t)top;
        }
        const uint32_t off = hist_alloc(a, n_sel);
        if (off == 0xFFFFFFFFu) { if (threadIdx.x == 0) { agg[0] = 0; agg[1] = 0; agg[2] = 0; } continue; }
        uint32_t run = 0;
        for (int base = 0; base < n_items && run < (uint32_t)top; base += blockDim.x) {
          const int j = base + threadIdx.x;
          uint32_t m = 0;
          uint64_t bits = 0;
          if (j < n_items) {
            const uint32_t row = a.item_row[i0 + j];
            if (row != kNoRow) {
              const uint64_t *rp = row_ptr(IT, row);
              if (present(rp, d.b[0]) && rp[d.w[0]] == 1) { m = 1; bits = rp[d.w[0] + 1]; }
            }
          }
          uint32_t ct;
          const uint32_t my = block_excl_scan(m, s_scan, &ct);
          if (m && run + my < (uint32_t)top) a.hist_pool[off + run + my] = bits;
          run += ct;
        }
        __syncthreads();
        // commons-math3 Percentile (LEGACY estimation, NaNs removed), p = 50:
        //   n == 1 -> the value;  work = non-NaN values;  pos = 0.5 * (n' + 1)
        //   pos < 1 -> min; pos >= n' -> max; else lower + (pos - floor(pos)) * (upper - lower)
        const uint64_t *vals = a.hist_pool + off;
        if (threadIdx.x == 0) { s_i[1] = 0; s_d[0] = nan_d(); s_d[1] = nan_d(); }
        __syncthreads();
        uint32_t nn = 0;
        for (uint32_t k = threadIdx.x; k < n_sel; k += blockDim.x) { const double v = __longlong_as_double((long long)vals[k]); if (v == v) nn++; }
        atomicAdd(&s_i[1], (int)nn);
        __syncthreads();
        const int np = s_i[1];
        double median = nan_d();
        if (n_sel == 1) {
          median = __longlong_as_double((long long)vals[0]);
        } else if (np > 0) {
          const double pos = 0.5 * (double)(np + 1);
          const double fpos = floor(pos);
          const int ip = (int)fpos;
          int k_lo, k_hi;
          if (pos < 1.0) k_lo = k_hi = 0;
          else if (pos >= (double)np) k_lo = k_hi = np - 1;
          else { k_lo = ip - 1; k_hi = ip; }
          // k-th smallest by rank counting (ties broken by position) over the non-NaN values
          for (uint32_t k = threadIdx.x; k < n_sel; k += blockDim.x) {
            const double v = __longlong_as_double((long long)vals[k]);
            if (v != v) continue;
            int rank = 0;
            for (uint32_t q = 0; q < n_sel; q++) {
              const double u = __longlong_as_double((long long)vals[q]);
              if (u != u) continue;
              rank += (u < v) || (u == v && q < k);
            }
            if (rank == k_lo) s_d[0] = v;
            if (rank == k_hi) s_d[1] = v;
          }
          __syncthreads();
          const double lower = s_d[0], upper = s_d[1];
          median = (k_lo == k_hi) ? lower : __dadd_rn(lower, __dmul_rn(pos - fpos, __dadd_rn(upper, -lower)));
        }
        __syncthreads();
        if (threadIdx.x == 0) { agg[0] = 1.0; agg[1] = median; agg[2] = 0.0; a.hist_desc[(size_t)r * a.n_hist + d.aux1] = make_uint2(0, 0); }
      } else {
        // strings: multiset of the tags of the first `top` string-typed items; sum = its size
        uint32_t total = 0, run = 0;
        for (int base = 0; base < n_items && run < (uint32_t)top; base += blockDim.x) {
          const int j = base + threadIdx.x;
          uint32_t m = 0, cnt = 0;
          if (j < n_items) {
            const uint32_t row = a.item_row[i0 + j];
            if (row != kNoRow) {
              const uint64_t *rp = row_ptr(IT, row);
              if (present(rp, d.b[0]) && rp[d.w[0]] == 2) { m = 1; cnt = (uint32_t)(rp[d.w[0] + 1] >> 32); }
            }
          }
          uint32_t ct, tt;
          const uint32_t ord = block_excl_scan(m, s_scan, &ct);
          if (!(m && run + ord < (uint32_t)top)) cnt = 0;
          block_excl_scan(cnt, s_scan, &tt);
          total += tt;
          run += ct;
        }
        const uint32_t off = total ? hist_begin(a, total) : 0, slots = hist_slots(total);
        if (off == 0xFFFFFFFFu) { if (threadIdx.x == 0) { agg[0] = 0; agg[1] = 0; agg[2] = 0; a.hist_desc[(size_t)r * a.n_hist + d.aux1] = make_uint2(0, 0); } continue; }
        run = 0;
        for (int base = 0; base < n_items && run < (uint32_t)top && total; base += blockDim.x) {
          const int j = base + threadIdx.x;
          uint32_t m = 0, cnt = 0, src = 0;
          if (j < n_items) {
            const uint32_t row = a.item_row[i0 + j];
            if (row != kNoRow) {
              const uint64_t *rp = row_ptr(IT, row);
              if (present(rp, d.b[0]) && rp[d.w[0]] == 2) { m = 1; const uint64_t ds = rp[d.w[0] + 1]; cnt = (uint32_t)(ds >> 32); src = (uint32_t)ds; }
            }
          }
          uint32_t ct;
          const uint32_t ord = block_excl_scan(m, s_scan, &ct);
          if (!(m && run + ord < (uint32_t)top)) cnt = 0;
          for (uint32_t k = 0; k < cnt; k++) hist_insert(a, off, slots, IT.pool[src + k]);
          run += ct;
        }
        if (total) hist_end(a, r, d.aux1, off, slots);
        else if (threadIdx.x == 0) a.hist_desc[(size_t)r * a.n_hist + d.aux1] = make_uint2(0, 0);
        if (threadIdx.x == 0) { agg[0] = 2.0; agg[1] = 0.0; agg[2] = (double)total; }
      }
    } else if (d.kind == FK_COSINE && d.aux1 != 0) {
      // Normalize.scale over the request's N raw values (S/ml/onnx/Normalize.scala:13-46)
      const double *raw = a.cos + (size_t)d.aux2 * a.total_items + i0;
      double *nrm = a.cos + (size_t)(a.n_cos + d.aux2) * a.total_items + i0;
      if (d.aux1 == 1) {
        // MinMax over non-NaN values; none -> unchanged
        if (threadIdx.x == 0) { s_i[0] = 0; }
        __syncthreads();
        double mn = 0, mx = 0;
        bool any = false;
        for (int j = threadIdx.x; j < n_items; j += blockDim.x) {
          const double v = raw[j];
          if (v == v) { if (!any) { mn = mx = v; any = true; } else { mn = fmin(mn, v); mx = fmax(mx, v); } }
        }
        // block reduce through shared memory (tiny): serialise by warp leaders
        for (int o = 16; o > 0; o >>= 1) {
          const double omn = __shfl_down_sync(0xFFFFFFFFu, mn, o), omx = __shfl_down_sync(0xFFFFFFFFu, mx, o);
          const int oany = __shfl_down_sync(0xFFFFFFFFu, (int)any, o);
          if (oany) { if (!any) { mn = omn; mx = omx; any = true; } else { mn = fmin(mn, omn); mx = fmax(mx, omx); } }
        }
        __shared__ double w_mn[8], w_mx[8];
        __shared__ int w_any[8];
        if ((threadIdx.x & 31) == 0) { w_mn[threadIdx.x >> 5] = mn; w_mx[threadIdx.x >> 5] = mx; w_any[threadIdx.x >> 5] = any; }
        __syncthreads();
        if (threadIdx.x == 0) {
          bool A = false; double MN = 0, MX = 0;
          for (int w = 0; w < (int)(blockDim.x >> 5); w++) if (w_any[w]) { if (!A) { MN = w_mn[w]; MX = w_mx[w]; A = true; } else { MN = fmin(MN, w_mn[w]); MX = fmax(MX, w_mx[w]); } }
          s_i[0] = A; s_d[0] = MN; s_d[1] = MX;
        }
        __syncthreads();
        if (s_i[0]) {
          const double MN = s_d[0], MX = s_d[1];
          for (int j = threadIdx.x; j < n_items; j += blockDim.x) nrm[j] = __ddiv_rn(__dadd_rn(raw[j], -MN), __dadd_rn(MX, -MN));
        }
        __syncthreads();
      } else {
        // Position: stable sort by value (NaN last), value -> sortedIndex / size, NaN stays NaN
        const double size = (double)n_items;
        for (int j = threadIdx.x; j < n_items; j += blockDim.x) {
          const double v = raw[j];
          if (v != v) { nrm[j] = v; continue; }
          const long long kj = total_order_key(v);
          int rank = 0;
          for (int q = 0; q < n_items; q++) {
            const long long kq = total_order_key(raw[q]);
            rank += (kq < kj) || (kq == kj && q < j);
          }
          nrm[j] = __ddiv_rn((double)rank, size);
        }
        __syncthreads();
      }
    }
  }
}

// Sum over `n` tags of their counts in histogram (r, h), in tag order (the reference folds left; the terms are small
// integers, but the order is kept anyway).  Four lookups in flight: the four tag loads are independent, then the four
// first probes are — the dependent chain per item is 2 ceil(n / 4) loads instead of 2 n (ncu_r2_final_assemble: the tag
// load and the probe behind it were 45 % of the kernel's samples).
__device__ __forceinline__ double hist_sum(const RankArgs &a, int r, int h, const uint64_t *tags, uint32_t n) {
  const uint2 hd = a.hist_desc[(size_t)r * a.n_hist + h];
  double s = 0.0;
  if (hd.y == 0) return s;  // n additions of +0.0
  const uint32_t mask = hd.y - 1;
  const ulonglong2 *tab = reinterpret_cast<const ulonglong2 *>(a.hist_pool + hd.x);
  for (uint32_t k = 0; k < n; k += 4) {
    uint64_t t[4];
    uint32_t sl[4];
    ulonglong2 e[4];
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = k + i < n ? tags[k + i] : 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { sl[i] = (uint32_t)mix64(t[i]) & mask; e[i] = tab[sl[i]]; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (k + i >= n) break;
      uint32_t c;
      if (t[i] == kHistEmpty) c = (uint32_t)a.hist_pool[hd.x + 2 * (size_t)hd.y];
      else {
        ulonglong2 x = e[i];
        uint32_t q = sl[i];
        while (x.x != t[i] && x.x != kHistEmpty) { q = (q + 1) & mask; x = tab[q]; }
        c = x.x == t[i] ? (uint32_t)x.y : 0u;
      }
      s = __dadd_rn(s, (double)c);
    }
  }
  return s;
}

// ------------------------------------------------------------------ coalesced row gather (fast columns)
// Most columns of a typical config are a plain function of one word of the item's row (numbers, word
// counts, booleans, category indices, counters, windows, stored vectors).  For those a WARP owns a
// group of 32 items and walks them one by one: the whole row (<= 128 words) arrives with up to four
// coalesced loads, lane c picks the word of column c out of the warp with a shuffle, converts it and
// turns it into the scorer's rank code; codes are staged in shared memory and written one 64-byte line
// per column, the f64 row (explain) as one contiguous run per item.
constexpr int kGatherWarps = 4;

// An item row of up to 128 words held across a warp: word w lives in lane (w & 31), register (w >> 5).
struct RowRegs { uint64_t w[4]; };
__device__ __forceinline__ void row_load(RowRegs &r, const uint64_t *rp, int rw, int lane) {
#pragma unroll
  for (int k = 0; k < 4; k++) r.w[k] = (rp && lane + 32 * k < rw) ? __ldg(rp + lane + 32 * k) : 0ull;
}
__device__ __forceinline__ uint64_t row_pick(const RowRegs &r, int word, int rw) {  // rw is warp-uniform
  uint64_t v = __shfl_sync(0xFFFFFFFFu, r.w[0], word & 31);
  if (rw > 32) { const uint64_t h = __shfl_sync(0xFFFFFFFFu, r.w[1], word & 31); if ((word >> 5) == 1) v = h; }
  if (rw > 64) { const uint64_t h = __shfl_sync(0xFFFFFFFFu, r.w[2], word & 31); if ((word >> 5) == 2) v = h; }
  if (rw > 96) { const uint64_t h = __shfl_sync(0xFFFFFFFFu, r.w[3], word & 31); if ((word >> 5) == 3) v = h; }
  return v;
}

// CODES: emit the scorer's u16 codes; OUT: emit the dense f64 row; XGB: XGBoost code semantics.
template <bool CODES, bool OUT, bool XGB>
__global__ void __launch_bounds__(kGatherWarps * 32) row_gather_kernel(RankArgs a) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  // [FastCol table][bucket headers (optional)][per warp: codes tile n_fast x 32 u16]
  FastCol *s_cols = reinterpret_cast<FastCol *>(s_raw);
  const size_t cols_bytes = ((size_t)a.n_fast * sizeof(FastCol) + 15) & ~size_t(15);
  BinParams bin = a.bin;
  size_t off = cols_bytes;
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a.fast_cols);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_raw);
    for (int k = threadIdx.x; k < a.n_fast * (int)(sizeof(FastCol) / 4); k += blockDim.x) dst[k] = __ldg(src + k);
  }
  if (CODES && a.stage_meta) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a.bin.meta);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_raw + off);
    for (int k = threadIdx.x; k < a.dim * (int)(sizeof(BinMeta) / 4); k += blockDim.x) dst[k] = __ldg(src + k);
    bin.meta = reinterpret_cast<const BinMeta *>(s_raw + off);
    off += (size_t)a.dim * sizeof(BinMeta);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // codes tile, item-major with an odd word stride: conflict-free both for the per-item stores
  // (lane = column) and for the per-column read-back (lane = item)
  const int tstride = (a.n_fast | 1) * 2;  // u16 elements per item row: (base, duplicate) per column, odd word stride
  uint16_t *tile = reinterpret_cast<uint16_t *>(s_raw + off) + (size_t)warp * 32 * tstride;
  const int g = blockIdx.x * kGatherWarps + warp;  // group of 32 items
  const int i0 = g * 32;
  if (i0 >= a.total_items) return;
  const DTable &IT = a.st.t[SC_ITEM];
  const int rw = IT.row_words;
  const uint32_t my_row = (i0 + lane < a.total_items) ? a.item_row[i0 + lane] : kNoRow;
  const int n_here = min(32, a.total_items - i0);
  const double kNaN = nan_d();
  for (int c0 = 0; c0 < a.n_fast; c0 += 32) {
    // this lane's column: descriptor and bucket-index header stay in registers for all 32 items
    const int c = c0 + lane;
    const bool act = c < a.n_fast;
    const FastCol fc = s_cols[act ? c : 0];
    BinMeta M{};
    bool cat = false;
    if (CODES) { M = bin.meta[fc.col]; cat = (M.flags & kMetaCat) != 0; }
    const int wsrc = fc.word, psrc = fc.bit >> 6;
    const double vmiss = fc.missing ? 0.0 : kNaN;
    // software pipeline: row j + 1 is in flight while row j is converted
    RowRegs nx;
    uint32_t ir_next = __shfl_sync(0xFFFFFFFFu, my_row, 0);
    row_load(nx, ir_next != kNoRow ? IT.rows + (size_t)ir_next * rw : nullptr, rw, lane);
    for (int j = 0; j < n_here; j++) {
      const uint32_t ir = ir_next;
      const RowRegs cur = nx;
      if (j + 1 < n_here) {
        ir_next = __shfl_sync(0xFFFFFFFFu, my_row, j + 1);
        row_load(nx, ir_next != kNoRow ? IT.rows + (size_t)ir_next * rw : nullptr, rw, lane);
      }
      const uint64_t vw = row_pick(cur, wsrc, rw), pw = row_pick(cur, psrc, rw);
      if (!act) continue;
      const int item = i0 + j;
      double v = vmiss;
      if (ir != kNoRow && ((pw >> (fc.bit & 63)) & 1ull)) {
        v = fc.conv == 0 ? __longlong_as_double((long long)vw)
            : fc.conv == 1 ? (double)(long long)vw : (double)(int)vw;
      }
      if (fc.override_slot >= 0 && a.item_f64) {
        const double o = __ldg(a.item_f64 + (size_t)item * a.n_item_f64 + fc.override_slot);
        if (o == o) v = o;
      }
      if (OUT) a.out_features[(size_t)item * a.dim + fc.col] = v;
      if (CODES) {
        const uint16_t cd = code_of_col_t<XGB>(bin, M, cat, v);
        *reinterpret_cast<uint32_t *>(tile + j * tstride + 2 * c) = (uint32_t)base_code(M, cd) | ((uint32_t)dup_code(cd) << 16);
      }
    }
  }
  if (CODES) {
    __syncwarp();
    for (int c = 0; c < a.n_fast; c++) {
      const int col = s_cols[c].col;
      const uint32_t dup = dup_col(bin.meta[col]);  // warp-uniform
      const uint32_t both = *reinterpret_cast<const uint32_t *>(tile + lane * tstride + 2 * c);
      if (lane < n_here) {
        a.codes[code_index(bin.tile_T, bin.tile_cols, i0 + lane, col)] = (uint16_t)both;
        if (dup != kMetaNoDup) a.codes[code_index(bin.tile_T, bin.tile_cols, i0 + lane, (int)dup)] = (uint16_t)(both >> 16);
      }
    }
  }
}

// ------------------------------------------------------------------ per-model code rows
// The fast columns of an item are a function of its row and of the model's thresholds only, so their
// codes are computed once per (model, row version) into `code_rows` and ranking just copies them:
// 2 bytes per tile column instead of the 8-byte row words + a bucket/threshold lookup per column.
template <bool XGB>
__global__ void __launch_bounds__(kGatherWarps * 32) code_rows_kernel(RankArgs a, uint32_t *code_rows, int crw,
                                                                      uint32_t n_rows, const uint32_t *idx, uint32_t n_idx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_work = idx ? n_idx : n_rows + 1;  // + the unknown-item row
  const uint32_t w0i = (blockIdx.x * kGatherWarps + warp) * 32u;
  if (w0i >= n_work) return;
  const DTable &IT = a.st.t[SC_ITEM];
  const int rw = IT.row_words;
  const BinParams &bin = a.bin;
  uint16_t *out16 = reinterpret_cast<uint16_t *>(code_rows);
  const double kNaN = nan_d();
  const uint32_t n_here = min(32u, n_work - w0i);
  for (int c0 = 0; c0 < a.n_fast; c0 += 32) {
    const int c = c0 + lane;
    const bool act = c < a.n_fast;
    const FastCol fc = a.fast_cols[act ? c : 0];
    const BinMeta M = bin.meta[fc.col];
    const bool cat = (M.flags & kMetaCat) != 0;
    const int wsrc = fc.word, psrc = fc.bit >> 6;
    for (uint32_t j = 0; j < n_here; j++) {
      // work item -> table row (kNoRow for the unknown-item row, which is the last one of a full build)
      uint32_t r = idx ? __ldg(idx + w0i + j) : w0i + j;
      if (!idx && r == n_rows) r = kNoRow;
      RowRegs cur;
      row_load(cur, r != kNoRow ? IT.rows + (size_t)r * rw : nullptr, rw, lane);
      const uint64_t vw = row_pick(cur, wsrc, rw), pw = row_pick(cur, psrc, rw);
      if (!act) continue;
      double v = fc.missing ? 0.0 : kNaN;
      if (r != kNoRow && ((pw >> (fc.bit & 63)) & 1ull)) {
        v = fc.conv == 0 ? __longlong_as_double((long long)vw)
            : fc.conv == 1 ? (double)(long long)vw : (double)(int)vw;
      }
      const uint16_t cd = code_of_col_t<XGB>(bin, M, cat, v);
      uint16_t *dst = out16 + (size_t)(r + 1u) * crw * 2;  // kNoRow + 1 == 0
      dst[fc.col] = base_code(M, cd);
      if (dup_col(M) != kMetaNoDup) dst[dup_col(M)] = dup_code(cd);
    }
  }
}

// Ranking with code rows: a warp copies the code rows of its 32 items (coalesced, independent loads) into
// a shared-memory tile and writes them out column by column in the scorer's [group][column][lane] layout.
__global__ void __launch_bounds__(kGatherWarps * 32) code_gather_kernel(RankArgs a) {
  extern __shared__ __align__(16) uint32_t s_tile[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * kGatherWarps + warp, i0 = g * 32;
  if (i0 >= a.total_items) return;
  const int crw = a.code_row_words, tw = crw | 1;  // odd word stride: conflict-free column reads
  uint32_t *tile = s_tile + (size_t)warp * 32 * tw;
  const int n_here = min(32, a.total_items - i0);
  const uint32_t my = (lane < n_here) ? a.item_row[i0 + lane] + 1u : 0u;  // unknown item (kNoRow) -> row 0
  if (a.total_items <= 4096) {
    // small batch (a single /rank request): latency matters, bandwidth does not — every lane streams its own
    // item's row, so all of a warp's 32 rows are in flight at once instead of one after another
    const uint32_t *src = a.code_rows + (size_t)my * crw;
#pragma unroll 8
    for (int l = 0; l < crw; l++) tile[lane * tw + l] = __ldg(src + l);
  } else {
    // the warp's 32 rows x crw words as ONE flattened range over the lanes: every load instruction is fully populated
    // (a row of 21 words would leave a third of the lanes idle) and there are crw of them instead of 32
    const int total = n_here * crw;
    int j = lane / crw, l = lane - j * crw;   // element e = base + lane lives in row j, word l
#pragma unroll 8
    for (int base = 0; base < total; base += 32) {
      const bool ok = base + lane < total;
      const uint32_t r = __shfl_sync(0xFFFFFFFFu, my, ok ? j : 0);
      if (ok) tile[j * tw + l] = __ldg(a.code_rows + (size_t)r * crw + l);
      l += 32;
      if (crw >= 16) { while (l >= crw) { l -= crw; j++; } }
      else { const int q = l / crw; j += q; l -= q * crw; }
    }
  }
  __syncwarp();
  const uint16_t *t16 = reinterpret_cast<const uint16_t *>(tile);
  if (a.bin.tile_T == 0) {
    uint16_t *out = a.codes + (size_t)g * a.bin.tile_cols * 32;
    if (lane < n_here)
      for (int c = 0; c < a.bin.tile_cols; c++) out[c * 32 + lane] = t16[lane * tw * 2 + c];
  } else if (lane < n_here) {
    // slim layout: a code row IS the sequence of its column pairs — one u32 per pair and lane, 128 bytes per warp store
    const int T = a.bin.tile_T, n_pairs = (a.bin.tile_cols + 1) >> 1, item = i0 + lane;
    uint32_t *out = reinterpret_cast<uint32_t *>(a.codes) + ((size_t)(item / T) * n_pairs) * T + (item % T);
    for (int pr = 0; pr < n_pairs; pr++) out[(size_t)pr * T] = tile[lane * tw + pr];
  }
}

// ------------------------------------------------------------------ assemble
__global__ void __launch_bounds__(128) assemble_kernel(RankArgs a) {
  // the extractor plan is read by every thread for every feature: stage it in shared memory
  extern __shared__ __align__(16) uint8_t s_plan_raw[];
  DFeature *s_plan = reinterpret_cast<DFeature *>(s_plan_raw);
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a.plan);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_plan_raw);
    const int n_words = a.n_plan * (int)(sizeof(DFeature) / 4);
    for (int k = threadIdx.x; k < n_words; k += blockDim.x) dst[k] = __ldg(src + k);
  }
  // ... and so are the per-column bucket-index headers of the binned scorer (32 B each)
  BinParams bin = a.bin;
  if (a.codes && a.stage_meta) {
    uint8_t *s_meta_raw = s_plan_raw + (((size_t)a.n_plan * sizeof(DFeature) + 15) & ~size_t(15));
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a.bin.meta);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_meta_raw);
    const int n_words = a.dim * (int)(sizeof(BinMeta) / 4);
    for (int k = threadIdx.x; k < n_words; k += blockDim.x) dst[k] = __ldg(src + k);
    bin.meta = reinterpret_cast<const BinMeta *>(s_meta_raw);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total_items) return;
  const int r = a.item_req[i];
  const uint32_t ir = a.item_row[i];
  const DTable &IT = a.st.t[SC_ITEM];
  const uint64_t *irow = ir != kNoRow ? row_ptr(IT, ir) : nullptr;
  if (irow) {
    // the extractors below touch the row word by word; without this each of its 32-byte sectors
    // is a separate, serialised HBM miss.  Pull the whole row towards L2/L1 at once.
    for (int w = 0; w < IT.row_words; w += 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(irow + w));
  }
  const double *ov = a.item_f64 ? a.item_f64 + (size_t)i * a.n_item_f64 : nullptr;
  const double kNaN = nan_d();
  // One assembled value: the dense f64 row (explain / f64 scorer) and/or its exact u16 rank code for
  // the binned scorer, laid out [group of 32 items][column][lane] so a warp's store is one 64-byte line.
  struct Emit {
    double *row;
    uint16_t *codes;
    const BinParams *bp;
    int item;
    __device__ __forceinline__ void operator()(int col, double v) const {
      if (row) row[col] = v;
      if (codes) {
        const BinMeta M = bp->meta[col];
        const uint16_t c = code_of_col(*bp, M, (M.flags & kMetaCat) != 0, v);
        codes[code_index(bp->tile_T, bp->tile_cols, item, col)] = base_code(M, c);
        if (dup_col(M) != kMetaNoDup) codes[code_index(bp->tile_T, bp->tile_cols, item, (int)dup_col(M))] = dup_code(c);
      }
    }
  };
  struct OutProxy {  // lets the extractor code below keep writing out[col] = v
    const Emit *e; int col;
    __device__ __forceinline__ void operator=(double v) const { (*e)(col, v); }
  };
  struct OutArr {
    const Emit *e;
    __device__ __forceinline__ OutProxy operator[](int col) const { return OutProxy{e, col}; }
  };
  const Emit emit{a.out_features ? a.out_features + (size_t)i * a.dim : nullptr, a.codes, &bin, i};
  const OutArr out{&emit};

  auto scoped_row = [&](int scope) -> const uint64_t * {
    switch (scope) {
      case SC_ITEM: return irow;
      case SC_GLOBAL: return a.st.t[SC_GLOBAL].n_rows ? a.st.t[SC_GLOBAL].rows : nullptr;
      case SC_USER: { const uint32_t v = a.visitor_row[2 * r]; return v != kNoRow ? row_ptr(a.st.t[SC_USER], v) : nullptr; }
      case SC_SESSION: { const uint32_t v = a.visitor_row[2 * r + 1]; return v != kNoRow ? row_ptr(a.st.t[SC_SESSION], v) : nullptr; }
    }
    return nullptr;
  };

  for (int f = 0; f < a.n_plan; f++) {
    const DFeature &d = s_plan[f];
    if (d.fast && a.n_fast > 0) continue;  // produced by row_gather_kernel
    switch (d.kind) {
      case FK_NUMBER: {
        // NumberFeature.values :84-93 (request-item field override) then value :58-69; WordCountFeature.value :63-68
        double v = kNaN;
        const double o = (d.in0 >= 0 && ov) ? ov[d.in0] : kNaN;
        if (o == o) v = o;
        else {
          const uint64_t *rp = scoped_row(d.scope);
          if (rp) {  // presence word and value word are fetched together (same or adjacent sector), then selected
            const uint64_t pw = __ldg(rp + (d.b[0] >> 6)), vw = __ldg(rp + d.w[0]);
            if ((pw >> (d.b[0] & 63)) & 1ull) v = __longlong_as_double((long long)vw);
          }
        }
        out[d.col] = v;
        break;
      }
      case FK_CATEGORY: {
        // IndexCategoricalEncoder: index + 1, 0 = nil (StringFeature.scala:124-137); CategoryValue -> index.toDouble
        double v = 0.0;
        const double o = (d.in0 >= 0 && ov) ? ov[d.in0] : kNaN;
        if (o == o) v = o;
        else { const uint64_t *rp = scoped_row(d.scope); if (rp && present(rp, d.b[0])) v = (double)(int)rp[d.w[0]]; }
        out[d.col] = v;
        break;
      }
      case FK_ONEHOT: {
        const double o = (d.in0 >= 0 && ov) ? ov[d.in0] : kNaN;
        if (o == o) { for (int k = 0; k < d.dim; k++) out[d.col + k] = ov[d.in0 + k]; break; }
        uint64_t mask = 0;
        const uint64_t *rp = scoped_row(d.scope);
        if (rp && present(rp, d.b[0])) mask = rp[d.w[0]];
        for (int k = 0; k < d.dim; k++) out[d.col + k] = (mask >> k) & 1ull ? 1.0 : 0.0;
        break;
      }
      case FK_COUNT: {
        // InteractionCountFeature.value :44-59 — missing is 0.0, not NaN
        const uint64_t *rp = scoped_row(d.scope);
        out[d.col] = (rp && present(rp, d.b[0])) ? (double)(long long)rp[d.w[0]] : 0.0;
        break;
      }
      case FK_WINDOW: {
        // WindowInteractionCountFeature.value :50-63
        const uint64_t *rp = scoped_row(d.scope);
        const bool ok = rp && present(rp, d.b[0]);
        for (int k = 0; k < d.dim; k++) out[d.col + k] = ok ? (double)(long long)rp[d.w[0] + k] : kNaN;
        break;
      }
      case FK_RATE: {
        // RateFeature.value :290-356
        const uint64_t *tr = nullptr;
        if (d.scope == SC_ITEM) tr = irow;
        else if (d.scope == SC_FIELD) {
          if (irow && present(irow, d.aux2)) {
            const uint32_t row = probe(a.st.t[SC_FIELD], d_hash_combine(d.uparam, irow[d.aux1]));
            if (row != kNoRow) tr = row_ptr(a.st.t[SC_FIELD], row);
          }
        } else {
          const uint64_t fv = a.req_u64 ? a.req_u64[(size_t)r * a.n_req_u64 + d.in1] : 0;
          if (fv) {
            const uint32_t row = probe(a.st.t[SC_IRF], d_hash_combine(d_hash_combine(d.uparam, fv), a.item_ids[i]));
            if (row != kNoRow) tr = row_ptr(a.st.t[SC_IRF], row);
          }
        }
        bool ok = tr && present(tr, d.b[0]) && present(tr, d.b[1]);
        const uint64_t *gr = nullptr;
        if (d.aux0) {
          gr = scoped_row(SC_GLOBAL);
          ok = ok && gr && present(gr, d.b[2]) && present(gr, d.b[3]);
        }
        for (int k = 0; k < d.dim; k++) {
          double v = kNaN;
          if (ok) {
            const long long top = (long long)tr[d.w[0] + k], bot = (long long)tr[d.w[1] + k];
            if (!d.aux0) {
              v = __ddiv_rn((double)top, (double)bot);  // Long / Double: 0/0 -> NaN, x/0 -> +-Inf
            } else {
              const long long tg = (long long)gr[d.w[2] + k], bg = (long long)gr[d.w[3] + k];
              if (tg == 0) {
                atomicExch(a.error_flag, MR_ERR_ARITHMETIC);  // java.lang.ArithmeticException: / by zero
              } else {
                const long long q = (tg == -1) ? -bg : bg / tg;  // Long division truncates toward zero
                v = __ddiv_rn(__dadd_rn(d.dparam, (double)top),
                              __dadd_rn(__dmul_rn(d.dparam, (double)q), (double)bot));
              }
            }
          }
          out[d.col + k] = v;
        }
        break;
      }
      case FK_INTERACTED: {
        // sum over the item's values of this field of the visitor's histogram count (:152-161)
        double cnt = 0.0;
        if (irow && present(irow, d.b[0])) {
          const uint64_t desc = irow[d.w[0]];
          const uint32_t off = (uint32_t)desc, n = (uint32_t)(desc >> 32);
          cnt = hist_sum(a, r, d.aux0, IT.pool + off, n);
        }
        out[d.col] = cnt;
        break;
      }
      case FK_TOKEN_MATCH: {
        // FieldMatchFeature.values (S/feature/FieldMatchFeature.scala:60-93): no query field / no item state -> 0.
        // Jaccard of the two token sets (FieldMatcher.score, S/feature/matcher/FieldMatcher.scala:15-49) or BM25
        // summed over the query's tokens in their order (BM25Matcher.score, matcher/BM25Matcher.scala:19-33;
        // the per-token IDF arrives as the token's weight).
        double v = 0.0;
        int q0 = 0, q1 = 0;
        if (a.req_tok_off) {
          const int32_t *qo = a.req_tok_off + (size_t)r * a.n_req_tok + d.in0;
          q0 = __ldg(qo) - a.req_tok_base;
          q1 = __ldg(qo + 1) - a.req_tok_base;
        }
        if (q1 > q0 && irow && present(irow, d.b[0])) {
          const uint64_t desc = irow[d.w[0]];
          const uint32_t off = (uint32_t)desc, n = (uint32_t)(desc >> 32);
          const uint64_t *doc = IT.pool + off;
          auto contains = [&](uint64_t h) -> bool {  // the stored set is sorted by hash
            uint32_t lo = 0, hi = n;
            while (lo < hi) { const uint32_t m = (lo + hi) >> 1; const uint64_t x = __ldg(doc + m); if (x < h) lo = m + 1; else hi = m; }
            return lo < n && __ldg(doc + lo) == h;
          };
          if (d.aux0 == 0) {
            if (n > 0) {
              int inter = 0;
              for (int q = q0; q < q1; q++) inter += contains(__ldg(a.req_tok_hash + q)) ? 1 : 0;
              v = __ddiv_rn((double)inter, (double)((q1 - q0) + (int)n - inter));
            }
          } else {
            const double K1 = 1.2, B = 0.75;
            // K1 * (1.0 - B + B * (doc.length / avgdl)) is the same for every term of this item
            const double norm = __dmul_rn(K1, __dadd_rn(1.0 - B, __dmul_rn(B, __ddiv_rn((double)n, d.dparam))));
            double sum = 0.0;
            for (int q = q0; q < q1; q++) {
              const double tf = contains(__ldg(a.req_tok_hash + q)) ? 1.0 : 0.0;
              const double idf = __ldg(a.req_tok_w + q);
              sum = __dadd_rn(sum, __ddiv_rn(__dmul_rn(idf, __dmul_rn(tf, K1 + 1.0)), __dadd_rn(tf, norm)));
            }
            v = sum;
          }
        }
        out[d.col] = v;
        break;
      }
      case FK_VECTOR: {
        // NumVectorFeature.value :55-70 — the stored (already reduced) list, or NaN x dim
        const uint64_t *rp = scoped_row(d.scope);
        const bool ok = rp && present(rp, d.b[0]);
        for (int k = 0; k < d.dim; k++) out[d.col + k] = ok ? __longlong_as_double((long long)rp[d.w[0] + k]) : kNaN;
        break;
      }
      case FK_ITEM_AGE: {
        // ItemAgeFeature.value :74-86: updatedAt = Timestamp(math.round(value * 1000));
        // updatedAt.diff(request.timestamp).toSeconds.toDouble  (|delta| millis, truncated to seconds)
        double v = kNaN;
        if (irow && present(irow, d.b[0])) {
          const double ms = __dmul_rn(__longlong_as_double((long long)irow[d.w[0]]), 1000.0);
          long long upd;  // java.lang.Math.round: half up, NaN -> 0, saturating
          if (ms != ms) upd = 0;
          else if (ms >= 9.2233720368547758e18) upd = 0x7FFFFFFFFFFFFFFFll;
          else if (ms <= -9.2233720368547758e18) upd = (long long)0x8000000000000000ull;
          else { const double fl = floor(ms); upd = (long long)fl + ((ms - fl) >= 0.5 ? 1 : 0); }
          const long long req_ts = a.req_u64 ? (long long)a.req_u64[(size_t)r * a.n_req_u64 + d.in1] : 0;
          const long long delta = req_ts - upd;
          const long long ad = delta < 0 ? -delta : delta;  // math.abs
          v = (double)(ad / 1000);
        }
        out[d.col] = v;
        break;
      }
      case FK_RELEVANCY:
        out[d.col] = ov ? ov[d.in0] : kNaN;
        break;
      case FK_POSITION:
        out[d.col] = d.dparam;  // ValueMode.OnlineInference (PositionFeature.scala:32)
        break;
      case FK_DIVERSITY: {
        const double *agg = a.reqagg + ((size_t)r * a.n_reqagg + d.aux2) * 4;
        const int mode = (int)agg[0];
        double v = 0.0;  // emptyResponse
        if (mode == 1) {
          v = (irow && present(irow, d.b[0]) && irow[d.w[0]] == 1)
                  ? __dadd_rn(__longlong_as_double((long long)irow[d.w[0] + 1]), -agg[1]) : kNaN;
        } else if (mode == 2) {
          if (irow && present(irow, d.b[0]) && irow[d.w[0]] == 2) {
            const uint64_t desc = irow[d.w[0] + 1];
            const uint32_t off = (uint32_t)desc, n = (uint32_t)(desc >> 32);
            const double s = hist_sum(a, r, d.aux1, IT.pool + off, n);
            v = __ddiv_rn(s, agg[2]);
          } else v = kNaN;
        }
        out[d.col] = v;
        break;
      }
      case FK_COSINE:
        out[d.col] = a.cos[(size_t)(a.n_cos + d.aux2) * a.total_items + i];
        break;
      case FK_CONST_REQ:
        for (int k = 0; k < d.dim; k++) out[d.col + k] = a.req_f64 ? a.req_f64[(size_t)r * a.n_req_f64 + d.in0 + k] : kNaN;
        break;
    }
  }
}

// ------------------------------------------------------------------ ordering
// Requests of up to kSmallOrder items (the usual /rank page): a WARP sorts one request with a bitonic
// network held entirely in registers — 4 (key, index) pairs per lane, strides below 4 are exchanges inside
// the lane, larger ones `shfl.xor` with the partner lane.  Keys are (total-order key of -score, index): every
// pair is distinct, which is the stability of the reference's sortBy.  No shared memory, no barrier, eight
// requests per CTA.
constexpr int kSmallOrder = 128;
constexpr int kOrderWarps = 8;

__device__ __forceinline__ bool pair_gt(long long ka, int ia, long long kb, int ib) {
  return ka > kb || (ka == kb && ia > ib);
}

__global__ void __launch_bounds__(kOrderWarps * 32) order_small_kernel(const double *scores, const int32_t *offsets,
                                                                       int n_requests, int32_t *order) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kOrderWarps + warp;
  if (r >= n_requests) return;
  const int b = __ldg(offsets + r), n = __ldg(offsets + r + 1) - b;
  if (n <= 0 || n > kSmallOrder) return;
  long long k[4];
  int id[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int e = lane * 4 + m;  // blocked layout: element e lives in lane e / 4
    k[m] = e < n ? total_order_key(-__ldg(scores + b + e)) : 0x7FFFFFFFFFFFFFFFll;
    id[m] = e < n ? e : 0x7FFFFFFF;
  }
#pragma unroll
  for (int size = 2; size <= kSmallOrder; size <<= 1) {
#pragma unroll
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j >= 4) {
        const int pl = j >> 2;  // partner lane distance
        const bool lower = (lane & pl) == 0;
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const long long ok = __shfl_xor_sync(0xFFFFFFFFu, k[m], pl);
          const int oi = __shfl_xor_sync(0xFFFFFFFFu, id[m], pl);
          const bool up = ((lane * 4 + m) & size) == 0;
          const bool mine_gt = pair_gt(k[m], id[m], ok, oi);
          // the lower element of an ascending pair keeps the minimum, the upper one the maximum (and vice versa)
          const bool take_other = (lower == up) ? mine_gt : !mine_gt;
          if (take_other) { k[m] = ok; id[m] = oi; }
        }
      } else {
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const int l = m ^ j;
          if (l > m) {
            const bool up = ((lane * 4 + m) & size) == 0;
            if (pair_gt(k[m], id[m], k[l], id[l]) == up) {
              const long long tk = k[m]; k[m] = k[l]; k[l] = tk;
              const int ti = id[m]; id[m] = id[l]; id[l] = ti;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int e = lane * 4 + m;
    if (e < n) order[b + e] = id[m];
  }
}

__global__ void __launch_bounds__(256) order_kernel(const double *scores, const int32_t *offsets, int n_requests,
                                                    int32_t *order) {
  extern __shared__ long long s_keys[];  // [4096] keys, then [4096] int indices
  const int cap = 4096;
  int *s_idx = reinterpret_cast<int *>(s_keys + cap);
  // a resident grid walks the requests; the ones order_small_kernel ranked are skipped
  for (int r = blockIdx.x; r < n_requests; r += gridDim.x) {
    const int b = __ldg(offsets + r), n = __ldg(offsets + r + 1) - b;
    if (n <= kSmallOrder || n > cap) continue;  // order_small_kernel's / order_big_*'s
    {
      // bitonic sort of (total-order key of -score, request index): the index makes every pair distinct,
      // which is exactly the stability of the reference's sortBy
      int p2 = 1;
      while (p2 < n) p2 <<= 1;
      __syncthreads();  // the previous request's read-out is complete
      for (int j = threadIdx.x; j < p2; j += blockDim.x) {
        s_keys[j] = j < n ? total_order_key(-scores[b + j]) : 0x7FFFFFFFFFFFFFFFll;
        s_idx[j] = j < n ? j : 0x7FFFFFFF;
      }
      __syncthreads();
      for (int k = 2; k <= p2; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
          for (int i = threadIdx.x; i < p2; i += blockDim.x) {
            const int l = i ^ jj;
            if (l > i) {
              const long long ki = s_keys[i], kl = s_keys[l];
              const int ii = s_idx[i], il = s_idx[l];
              const bool gt = ki > kl || (ki == kl && ii > il);
              const bool up = (i & k) == 0;
              if (gt == up) { s_keys[i] = kl; s_keys[l] = ki; s_idx[i] = il; s_idx[l] = ii; }
            }
          }
          __syncthreads();
        }
      }
      for (int j = threadIdx.x; j < n; j += blockDim.x) order[b + j] = s_idx[j];
      continue;
    }
    // larger requests: order_big_* (below) — a whole grid per request instead of one CTA
  }
}

// Mega-requests (more items than the CTA sort holds; BASELINE config #5 is one 10 000-item request), O(n log n) over
// the whole chip instead of one CTA:
//   1. order_big_sort_kernel   every 1024-item chunk of the request is sorted by its own CTA (bitonic network in shared
//                              memory over (total-order key of -score, index in the request) pairs: all distinct, which is
//                              the stability of the reference's sortBy) and written to scratch;
//   2. order_big_merge_kernel  one thread per element: its final rank = its position in its own sorted chunk + for every
//                              other chunk the number of elements that sort before it, found by binary search — keys <= its
//                              key in chunks of LOWER item indices (ties lose to the earlier index), keys < its key in later ones.
constexpr int kBigOrderMin = 4096;    // requests above this size take these paths:
constexpr int kCountOrderMax = 20000; //   (kBigOrderMin, kCountOrderMax] items rank by counting (order_count_kernel), larger ones sort + merge
constexpr int kBigOrderChunk = 1024;  // items per sorted chunk (chunks are aligned to the batch's item index space)

__device__ __forceinline__ int owning_request(const int32_t *offsets, int n_requests, int i) {
  int lo = 0, hi = n_requests;  // last r with off[r] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(offsets + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) order_big_sort_kernel(const double *scores, const int32_t *offsets, int n_requests,
                                                             int total_items, long long *skeys, int32_t *sidx) {
  __shared__ long long s_k[kBigOrderChunk];
  __shared__ int s_i[kBigOrderChunk];
  const int c0 = blockIdx.x * kBigOrderChunk, c1 = min(total_items, c0 + kBigOrderChunk);
  for (int r = owning_request(offsets, n_requests, c0); r < n_requests; r++) {
    const int b = __ldg(offsets + r), n = __ldg(offsets + r + 1) - b;
    if (b >= c1) break;
    if (n <= kCountOrderMax) continue;  // order_small_kernel's / order_kernel's / order_count_kernel's
    const int lo = max(b, c0), hi = min(b + n, c1), m = hi - lo;  // this chunk's part of request r
    if (m <= 0) continue;
    int p2 = 1;
    while (p2 < m) p2 <<= 1;
    __syncthreads();  // the previous segment's read-out is complete
    for (int j = threadIdx.x; j < p2; j += blockDim.x) {
      s_k[j] = j < m ? total_order_key(-scores[lo + j]) : 0x7FFFFFFFFFFFFFFFll;
      s_i[j] = j < m ? lo + j - b : 0x7FFFFFFF;
    }
    __syncthreads();
    for (int k = 2; k <= p2; k <<= 1) {
      for (int jj = k >> 1; jj > 0; jj >>= 1) {
        for (int i = threadIdx.x; i < p2; i += blockDim.x) {
          const int l = i ^ jj;
          if (l > i) {
            const long long ki = s_k[i], kl = s_k[l];
            const int ii = s_i[i], il = s_i[l];
            const bool gt = ki > kl || (ki == kl && ii > il);
            if (gt == ((i & k) == 0)) { s_k[i] = kl; s_k[l] = ki; s_i[i] = il; s_i[l] = ii; }
          }
        }
        __syncthreads();
      }
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) { skeys[lo + j] = s_k[j]; sidx[lo + j] = s_i[j]; }
  }
}

__global__ void __launch_bounds__(256) order_big_merge_kernel(const int32_t *offsets, int n_requests, int total_items,
                                                              const long long *skeys, const int32_t *sidx, int32_t *order) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_items) return;
  const int r = owning_request(offsets, n_requests, g);
  const int b = __ldg(offsets + r), n = __ldg(offsets + r + 1) - b;
  if (n <= kCountOrderMax) return;
  const int my_lo = max(b, (g / kBigOrderChunk) * kBigOrderChunk);
  const long long kj = skeys[g];
  int rank = g - my_lo;  // position inside its own sorted chunk
  for (int c0 = (b / kBigOrderChunk) * kBigOrderChunk; c0 < b + n; c0 += kBigOrderChunk) {
    const int lo = max(b, c0), hi = min(b + n, c0 + kBigOrderChunk);
    if (lo == my_lo) continue;
    const bool ties_before = lo < my_lo;  // the other chunk holds earlier item indices: equal keys sort before this element
    int x = lo, y = hi;                   // first position whose key is > kj (ties_before) or >= kj
    while (x < y) {
      const int mid = (x + y) >> 1;
      const long long km = __ldg(skeys + mid);
      if (ties_before ? km <= kj : km < kj) x = mid + 1; else y = mid;
    }
    rank += x - lo;
  }
  order[b + rank] = sidx[g];
}

// Requests of a few thousand to ~20 000 items (BASELINE config #5: one request of 10 000): the rank of an element IS the
// number of elements that sort before it, and at this size counting them outright is cheaper than sorting — n^2 compares
// spread over the whole chip (four elements per warp, keys staged through shared memory in tiles every warp of the CTA
// reads) against a chunk sort plus ~n log^2 n dependent binary-search loads.  Same order as every other path: by the
// total-order key of -score, ties by the earlier index (the stability of the reference's sortBy).
constexpr int kCountPerWarp = 4;       // elements a warp ranks at once: one shared-memory read of a key serves four compares
constexpr int kCountThreads = 128;      // 4 warps = 16 elements per CTA: 625 CTAs for 10 000 items spread evenly enough over 148 SMs
constexpr int kCountTile = 1024;       // keys per shared-memory tile
constexpr int kCountPerCta = (kCountThreads / 32) * kCountPerWarp;

// n^2 x 8 bytes through the shared-memory pipe is what a warp per element costs (10 000 items: 800 MB = 21 us at
// 128 B/clk/SM; measured 45 us, the same as 8 lanes per element, which was latency-bound at 17 warps per SM) — so a warp
// keeps FOUR consecutive elements in registers and every key it reads is compared against all four: 33-35 us on the
// 10 000-item request, against 52 us for the chunk sort + merge it replaces there (16 half-rate ISETPs per two keys are
// what is left: unrolled staging, predicated adds and smaller CTAs moved it by 2 us).  The tie rule depends on the side of the element a key comes from; the four elements are neighbours,
// so all keys before the first of them take `<=`, all keys after the last take `<`, and only the 4 x 4 block in between
// is compared with the index.
__global__ void __launch_bounds__(kCountThreads) order_count_kernel(const double *scores, const int32_t *offsets, int n_requests,
                                                          int total_items, int32_t *order) {
  __shared__ long long s_k[kCountTile];
  const int g0 = blockIdx.x * kCountPerCta, g1 = min(total_items, g0 + kCountPerCta);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w0 = g0 + warp * kCountPerWarp;  // this warp's elements: w0 .. w0 + 3 (batch item indices)
  for (int r = owning_request(offsets, n_requests, g0); r < n_requests; r++) {
    const int b = __ldg(offsets + r), n = __ldg(offsets + r + 1) - b;
    if (b >= g1) break;
    if (n <= kBigOrderMin || n > kCountOrderMax) continue;
    // the warp's elements that belong to this request: [e_lo, e_hi) of its four
    const int e_lo = max(0, b - w0), e_hi = min(kCountPerWarp, min(g1, b + n) - w0);
    const bool any = e_lo < e_hi;
    long long key[kCountPerWarp];
    int cnt[kCountPerWarp];
#pragma unroll
    for (int e = 0; e < kCountPerWarp; e++) {
      cnt[e] = 0;
      key[e] = (e >= e_lo && e < e_hi) ? total_order_key(-scores[w0 + e]) : 0;
    }
    const int first = w0 + e_lo - b, last = w0 + e_hi - 1 - b;  // request-relative indices of the first / last active element
    for (int t0 = 0; t0 < n; t0 += kCountTile) {
      const int m = min(kCountTile, n - t0);
      __syncthreads();  // the previous tile has been read
      {
        // all of a thread's loads first, then the keys: one memory latency per tile instead of one per element (as a
        // plain loop the staging alone cost ~14 us of the kernel: 40 dependent L2 round trips per thread)
        double v[kCountTile / kCountThreads];
#pragma unroll
        for (int q = 0; q < kCountTile / kCountThreads; q++) {
          const int j = threadIdx.x + q * kCountThreads;
          v[q] = j < m ? scores[b + t0 + j] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < kCountTile / kCountThreads; q++) {
          const int j = threadIdx.x + q * kCountThreads;
          if (j < m) s_k[j] = total_order_key(-v[q]);
        }
      }
      __syncthreads();
      if (!any) continue;
      const int end_a = min(max(first - t0, 0), m);       // tile positions [0, end_a): items before all four -> ties count
      const int beg_b = min(max(last + 1 - t0, 0), m);    // [beg_b, m): items after all four -> ties do not
      // (a 64-bit compare is two ISETPs; the predicated add keeps the count at one instruction — `cnt += k <= key` compiles
      // to an add plus a select)
#pragma unroll 2
      for (int j = lane; j < end_a; j += 32) {
        const long long k = s_k[j];
#pragma unroll
        for (int e = 0; e < kCountPerWarp; e++)
          asm("{\n.reg .pred p;\nsetp.le.s64 p, %1, %2;\n@p add.s32 %0, %0, 1;\n}" : "+r"(cnt[e]) : "l"(k), "l"(key[e]));
      }
#pragma unroll 2
      for (int j = beg_b + lane; j < m; j += 32) {
        const long long k = s_k[j];
#pragma unroll
        for (int e = 0; e < kCountPerWarp; e++)
          asm("{\n.reg .pred p;\nsetp.lt.s64 p, %1, %2;\n@p add.s32 %0, %0, 1;\n}" : "+r"(cnt[e]) : "l"(k), "l"(key[e]));
      }
      const int j = end_a + lane;  // the block between them: at most four positions, compared with their index
      if (j < beg_b) {
        const long long k = s_k[j];
#pragma unroll
        for (int e = 0; e < kCountPerWarp; e++) cnt[e] += k < key[e] || (k == key[e] && t0 + j < w0 + e - b);
      }
    }
#pragma unroll
    for (int e = 0; e < kCountPerWarp; e++) {
      const int c = __reduce_add_sync(0xFFFFFFFFu, cnt[e]);
      if (lane == 0 && e >= e_lo && e < e_hi) order[b + c] = w0 + e - b;
    }
  }
}

}  // namespace

void launch_assemble(const RankArgs &a, const Schema &schema, cudaStream_t stream) {
  if (a.total_items <= 0 && a.n_requests <= 0) return;
  const int n = std::max(a.total_items, a.n_requests);
  { ProfScope _ps("lookup_kernel", stream); lookup_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a); }
  MR_CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
  if (a.total_items <= 0) return;
  if (schema.needs_cosine) {
    bool any_f32 = false, any_f64 = false;
    int dim_max = 0;
    for (auto &d : schema.plan)
      if (d.kind == FK_COSINE) {
        if (a.st.side_f32[(int)d.uparam]) { any_f32 = true; dim_max = std::max(dim_max, (int)d.aux0); }
        else any_f64 = true;
      }
    if (any_f32) {
      { ProfScope _ps("cosine_qnorm_kernel", stream); cosine_qnorm_kernel<<<(a.n_requests + 3) / 4, 128, (size_t)4 * std::max(a.vec_stride, 1) * sizeof(float), stream>>>(a); }
      MR_CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
      const size_t smem = (((size_t)dim_max * 8 + 15) & ~size_t(15)) + (size_t)4 * kCosStages * 32 * kCosPitch * sizeof(float);
      MR_CUDA_CHECK(cudaFuncSetAttribute(cosine_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      { ProfScope _ps("cosine_f32_kernel", stream); cosine_f32_kernel<<<(a.total_items + 127) / 128, 128, smem, stream>>>(a, dim_max); }
      MR_CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    }
    if (any_f64) {
      { ProfScope _ps("cosine_kernel", stream); cosine_kernel<<<(a.total_items + 127) / 128, 128, 0, stream>>>(a); }
      MR_CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    }
  }
  if (schema.needs_prepass && a.n_requests > 0) {
    int n_agg = 0;
    for (auto &d : schema.plan) n_agg += d.kind == FK_INTERACTED || d.kind == FK_DIVERSITY || (d.kind == FK_COSINE && d.aux1 != 0);
    { ProfScope _ps("prepass_kernel", stream); prepass_kernel<<<dim3((unsigned)a.n_requests, (unsigned)std::max(n_agg, 1)), 256, 0, stream>>>(a); }
    MR_CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
  {
    RankArgs b = a;
    const size_t plan_bytes = (((size_t)std::max(a.n_plan, 1) * sizeof(DFeature)) + 15) & ~size_t(15);
    const size_t meta_bytes = a.codes ? (size_t)a.dim * sizeof(BinMeta) : 0;
    b.stage_meta = meta_bytes > 0 && plan_bytes + meta_bytes <= 40 * 1024;
    // row-local columns: coalesced gather kernel; everything else: the generic per-item kernel
    const size_t cols_bytes = ((size_t)a.n_fast * sizeof(FastCol) + 15) & ~size_t(15);
    const size_t gather_smem = cols_bytes + (b.stage_meta ? meta_bytes : 0) +
                               (a.codes ? (size_t)kGatherWarps * 32 * ((a.n_fast | 1) * 2) * sizeof(uint16_t) : 0);
    if (a.n_fast > 0 && gather_smem > 96 * 1024) b.n_fast = 0;  // too many columns for the tile: generic path
    bool any_generic = false;
    for (auto &d : schema.plan) any_generic |= !(d.fast && b.n_fast > 0);
    if (b.n_fast > 0 && a.code_rows && a.codes && !a.out_features && !(a.item_f64 && schema.fast_has_override)) {
      // the fast columns' codes are already materialised per item row: copy them
      const int n_groups = (a.total_items + 31) / 32;
      const size_t smem = (size_t)kGatherWarps * 32 * (a.code_row_words | 1) * sizeof(uint32_t);
      if (smem > 48 * 1024)
        MR_CUDA_CHECK(cudaFuncSetAttribute(code_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      { ProfScope _ps("code_gather_kernel", stream); code_gather_kernel<<<(n_groups + kGatherWarps - 1) / kGatherWarps, kGatherWarps * 32, smem, stream>>>(b); }
      MR_CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    } else if (b.n_fast > 0) {
      const int n_groups = (a.total_items + 31) / 32;
      auto go = [&](auto kern) {
        if (gather_smem > 48 * 1024)
          MR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gather_smem));
        { ProfScope _ps("row_gather_kernel", stream); kern<<<(n_groups + kGatherWarps - 1) / kGatherWarps, kGatherWarps * 32, gather_smem, stream>>>(b); }
        MR_CUDA_CHECK(cudaGetLastError());
        g_kernel_launches++;
      };
      const bool xgb = a.codes && a.bin.xgb;
      if (a.codes && a.out_features) { if (xgb) go(row_gather_kernel<true, true, true>); else go(row_gather_kernel<true, true, false>); }
      else if (a.codes) { if (xgb) go(row_gather_kernel<true, false, true>); else go(row_gather_kernel<true, false, false>); }
      else go(row_gather_kernel<false, true, false>);
    }
    if (any_generic) {
      { ProfScope _ps("assemble_kernel", stream); assemble_kernel<<<(a.total_items + 127) / 128, 128, plan_bytes + (b.stage_meta ? meta_bytes : 0), stream>>>(b); }
      MR_CUDA_CHECK(cudaGetLastError());
      g_kernel_launches++;
    }
  }
}

void launch_code_rows(const RankArgs &a, uint32_t *code_rows, int code_row_words, uint32_t n_rows, const uint32_t *d_idx,
                      uint32_t n_idx, cudaStream_t stream) {
  const uint32_t n_work = d_idx ? n_idx : n_rows + 1;
  if (n_work == 0 || a.n_fast <= 0) return;
  const unsigned grid = (unsigned)((n_work + kGatherWarps * 32 - 1) / (kGatherWarps * 32));
  if (a.bin.xgb) { ProfScope _ps("code_rows_kernel", stream); code_rows_kernel<true><<<grid, kGatherWarps * 32, 0, stream>>>(a, code_rows, code_row_words, n_rows, d_idx, n_idx); }
  else { ProfScope _ps("code_rows_kernel", stream); code_rows_kernel<false><<<grid, kGatherWarps * 32, 0, stream>>>(a, code_rows, code_row_words, n_rows, d_idx, n_idx); }
  MR_CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}

void launch_rank_order(const double *d_scores, const int32_t *d_item_offsets, int n_requests, int total_items,
                       int32_t *d_order, cudaStream_t stream, int max_items_hint, int32_t *d_rank_tmp) {
  if (n_requests <= 0 || total_items <= 0) return;
  const bool unknown = max_items_hint <= 0;
  if (unknown || max_items_hint <= kSmallOrder || n_requests > 1) {
    { ProfScope _ps("order_small_kernel", stream); order_small_kernel<<<(n_requests + kOrderWarps - 1) / kOrderWarps, kOrderWarps * 32, 0, stream>>>(d_scores, d_item_offsets, n_requests, d_order); }
    MR_CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
  const bool only_mega = !unknown && n_requests == 1 && max_items_hint > kBigOrderMin;  // one request, and it is a large one
  if ((unknown || max_items_hint > kSmallOrder) && !only_mega) {  // larger requests: CTA-wide sort (small ones return at once)
    { ProfScope _ps("order_kernel", stream); order_kernel<<<std::min(n_requests, 148 * 4), 256, 4096 * (sizeof(long long) + sizeof(int)), stream>>>(d_scores, d_item_offsets, n_requests, d_order); }
    MR_CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
  if ((unknown && total_items > kBigOrderMin) || max_items_hint > kBigOrderMin) {
    // (kBigOrderMin, kCountOrderMax] items: rank by counting; beyond: chunk sort + merge (each kernel skips the other's requests)
    { ProfScope _ps("order_count_kernel", stream); order_count_kernel<<<(unsigned)((total_items + kCountPerCta - 1) / kCountPerCta), kCountThreads, 0, stream>>>(d_scores, d_item_offsets, n_requests, total_items, d_order); }
    MR_CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
  if ((unknown && total_items > kCountOrderMax) || max_items_hint > kCountOrderMax) {
    int32_t *tmp = d_rank_tmp;  // 3 ints per item: sorted keys (8 B) + their item indices (4 B)
    if (!tmp) MR_CUDA_CHECK(cudaMallocAsync((void **)&tmp, (size_t)total_items * 12 + 16, stream));
    long long *skeys = reinterpret_cast<long long *>(tmp);
    int32_t *sidx = tmp + 2 * (size_t)total_items;
    { ProfScope _ps("order_big_sort_kernel", stream); order_big_sort_kernel<<<(unsigned)((total_items + kBigOrderChunk - 1) / kBigOrderChunk), 256, 0, stream>>>(d_scores, d_item_offsets, n_requests, total_items, skeys, sidx); }
    MR_CUDA_CHECK(cudaGetLastError());
    { ProfScope _ps("order_big_merge_kernel", stream); order_big_merge_kernel<<<(unsigned)((total_items + 255) / 256), 256, 0, stream>>>(d_item_offsets, n_requests, total_items, skeys, sidx, d_order); }
    MR_CUDA_CHECK(cudaGetLastError());
    g_kernel_launches += 2;
    if (!d_rank_tmp) MR_CUDA_CHECK(cudaFreeAsync(tmp, stream));
  }
}

}  // namespace mr
