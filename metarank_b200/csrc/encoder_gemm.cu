// Dense GEMM of the bi-encoder's query forward on the 5th-generation tensor cores — the one place on
// the /rank path where a dense contraction, not shared-memory bandwidth, is the roofline.
//
//   C[M x N] = act( A[M x K] * W[N x K]^T + bias[N] ) (+ residual[M x N])      A, W: binary16, K-major; accumulate f32
//
// A persistent CTA per SM walks 128 x BLOCK_N output tiles; ten warps in three roles (the canonical sm_100 shape):
//   warp 0, one lane   TMA producer: per 64-element K block one `cp.async.bulk.tensor.2d` for the A tile and one for the
//                      W tile, 128-byte swizzled, into a kStages-deep ring; completion on the stage's `full` mbarrier
//   warp 1, one lane   MMA issuer: `tcgen05.mma.cta_group::1.kind::f16` (UMMA 128 x BLOCK_N x 16, operands straight from
//                      shared memory through 64-bit matrix descriptors, accumulator in TENSOR MEMORY); `tcgen05.commit`
//                      frees the stage for the producer and, after the last K block, hands the accumulator — one of
//                      two TMEM buffers — to the epilogue, then starts the next tile in the other
//   warps 2..9         epilogue: `tcgen05.ld` 32 lanes x 32 columns at a time -> bias, exact-erf GELU, residual ->
//                      f32 and/or binary16 rows to global memory (two warps per TMEM lane quarter, half the columns each)
// warp 1 also allocates / frees the TMEM columns.  SASS: UTMALDG (TMA), UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "common.h"
#include "encoder.h"
#include "gbdt_kernels.cuh"  // ProfScope, g_kernel_launches
#include "tma.cuh"

namespace mr {
namespace {

constexpr int kBlockM = 128, kBlockK = 64, kUmmaK = 16;
// ring depth: 192 KB of operand stages either way (6 x 32 KB, 8 x 24 KB) — a whole K = 384 tile ahead of the MMA warp
template <int BLOCK_N> struct StagesFor { static constexpr int value = BLOCK_N == 192 ? 5 : BLOCK_N == 128 ? 6 : 8; };  // 200 / 192 / 192 KB
// TMEM columns are allocated in powers of two: the two accumulators of a 192-column tile sit 256 columns apart
template <int BLOCK_N> struct AccStride { static constexpr uint32_t value = BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256; };

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_out, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 64-bit shared-memory matrix descriptor of a K-major tile whose rows are 128-byte swizzled 64-element (128 B) segments,
// exactly what the TMA box {64, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes: 8-row groups are 1024 B apart (SBO), the
// leading-dimension offset is unused for this layout, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, 16-byte units           bits [0, 14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (ignored here)     bits [16, 30)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: 8 rows x 128 B     bits [32, 46)
  d |= (uint64_t)1 << 46;                           // descriptor version                     bits [46, 48)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B                           bits [61, 64)
  return d;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

struct GemmParams {
  const float *bias;      // [N] or null
  const float *residual;  // [M x N] f32 or null: added after the activation
  float *out_f32;         // [M x N] or null
  __half *out_f16;        // [M x N] or null
  int M, N, K;
  int gelu;
};

// One lane of a converged warp, chosen by the hardware: unlike `lane == 0` the compiler knows the branch holds a single
// thread and emits the tcgen05 / TMA instructions bare instead of wrapping each in an ELECT / BRA.U.ANY loop over the active
// lanes (ncu_r2_gemm3: that wrapping was most of what the MMA-issuing thread executed).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent: one CTA per SM walks output tiles t = blockIdx.x, + gridDim.x, ... (n fastest, so the CTAs running together
// share an A row block in L2).  The accumulator is double-buffered in tensor memory (2 x BLOCK_N columns): the epilogue
// warps drain tile i while the MMA warp already accumulates tile i + 1, and the smem ring never drains between tiles.
// EPI: what the epilogue adds after the bias — compile-time, so that the plain layers carry neither 64 inlined erf
// expansions to branch around (instruction-fetch stalls: ncu_r2_gemm2) nor predicated-off residual loads.
enum { EPI_BIAS = 0, EPI_GELU = 1, EPI_RESIDUAL = 2 };

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(320, 1) encoder_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_w, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // the 128-byte swizzle atoms (8 rows x 128 B) must start on 1024-byte boundaries of the shared-memory address space
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // [A stages: kStages x 128 x 128 B][W stages: kStages x BLOCK_N x 128 B][barriers][tmem address]
  constexpr uint32_t kStages = StagesFor<BLOCK_N>::value;
  constexpr uint32_t kABytes = kBlockM * kBlockK * 2, kWBytes = BLOCK_N * kBlockK * 2;
  constexpr uint32_t kAccStride = AccStride<BLOCK_N>::value, kTmemCols = 2 * kAccStride;  // a power of two >= 32
  uint8_t *sa = smem, *sw = smem + kStages * kABytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(sw + kStages * kWBytes), *empty = full + kStages;
  uint64_t *acc_full = empty + kStages, *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kblocks = p.K / kBlockK;
  const int tiles_n = p.N / BLOCK_N, n_tiles = ((p.M + kBlockM - 1) / kBlockM) * tiles_n;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (uint32_t s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }  // 8 epilogue warps release a buffer
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // Waits name the phase parity they need; a fresh barrier reads as "phase 1 complete", so the first pass over `empty` /
  // `acc_empty` falls through.  Polling waits (mbar_wait_spin): the waiter is normally AHEAD of the data here, where
  // try_wait's suspension costs more than the spin.
  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * kBlockM, n0 = (t % tiles_n) * BLOCK_N;
        for (int kb = 0; kb < n_kblocks; kb++, it++) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait_spin(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], kABytes + kWBytes);
          tma_load_2d(sa + (size_t)s * kABytes, &map_a, kb * kBlockK, m0, &full[s]);
          tma_load_2d(sw + (size_t)s * kWBytes, &map_w, kb * kBlockK, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer.  Instruction descriptor: D = f32, A = B = f16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, lt++) {
        const uint32_t acc = lt & 1, acc_ph = (lt >> 1) & 1;
        mbar_wait_spin(&acc_empty[acc], acc_ph ^ 1);  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < n_kblocks; kb++, it++) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait_spin(&full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_kmajor_sw128_desc(smem_u32(sa + (size_t)s * kABytes));
          const uint64_t dw = make_kmajor_sw128_desc(smem_u32(sw + (size_t)s * kWBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; k++)  // advancing 16 elements = 32 bytes inside the swizzle atom: +2 in 16-byte units
            umma_f16(tmem_d, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          umma_commit(&empty[s]);                                // the stage is free once these MMAs have read it
          if (kb == n_kblocks - 1) umma_commit(&acc_full[acc]);  // ... and the accumulator is complete
        }
      }
    }
  } else {
    // ===== epilogue: 8 warps.  A warp may only read the TMEM lanes of its quarter (warp % 4): rows [32 (w % 4), +32) of the
    // tile; the two warps of a quarter split the columns.  Everything that does not depend on the accumulator is fetched
    // BEFORE the wait — the tile's bias slice (lane l holds column 32 c + l of each chunk, broadcast by shuffle later) and
    // the thread's residual row — so the drain itself is TMEM load -> FADDs -> stores (ncu_r2_gemm: with the loads inside,
    // their latency, 16 dependent round trips per tile, was the whole kernel).
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    constexpr int kChunks = BLOCK_N / 64;  // 32-column chunks per warp
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, lt++) {
      const int m0 = (t / tiles_n) * kBlockM, n0 = (t % tiles_n) * BLOCK_N + half * (BLOCK_N / 2);
      const uint32_t acc = lt & 1, acc_ph = (lt >> 1) & 1;
      const int row = m0 + quarter * 32 + lane;
      float bias_lane[kChunks];
#pragma unroll
      for (int c = 0; c < kChunks; c++) bias_lane[c] = p.bias ? __ldg(p.bias + n0 + 32 * c + lane) : 0.f;
      if (lane == 0) mbar_wait_spin(&acc_full[acc], acc_ph);  // one polling lane per warp (try_wait's suspension oversleeps by microseconds)
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c = 0; c < kChunks; c++) {
        const int n = n0 + 32 * c;
        float4 res[8];
        if (EPI == EPI_RESIDUAL && row < p.M) {
          const float4 *rs = reinterpret_cast<const float4 *>(p.residual + (size_t)row * p.N + n);
#pragma unroll
          for (int j = 0; j < 8; j++) res[j] = __ldg(rs + j);
        }
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kAccStride + (uint32_t)(half * (BLOCK_N / 2) + 32 * c), v);
        if (c == kChunks - 1) {
          // every column this warp owns is in registers: hand the accumulator back before the stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float x = __uint_as_float(v[j]) + __shfl_sync(0xFFFFFFFFu, bias_lane[c], j);
          if (EPI == EPI_GELU) x = gelu_erf(x);
          r[j] = x;
        }
        if (row < p.M) {
          if (EPI == EPI_RESIDUAL) {
#pragma unroll
            for (int j = 0; j < 8; j++) { r[4 * j] += res[j].x; r[4 * j + 1] += res[j].y; r[4 * j + 2] += res[j].z; r[4 * j + 3] += res[j].w; }
          }
          if (p.out_f32) {
            float4 *o = reinterpret_cast<float4 *>(p.out_f32 + (size_t)row * p.N + n);
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
          if (p.out_f16) {
            uint4 *o = reinterpret_cast<uint4 *>(p.out_f16 + (size_t)row * p.N + n);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              __half2 h0 = __floats2half2_rn(r[8 * j], r[8 * j + 1]), h1 = __floats2half2_rn(r[8 * j + 2], r[8 * j + 3]);
              __half2 h2 = __floats2half2_rn(r[8 * j + 4], r[8 * j + 5]), h3 = __floats2half2_rn(r[8 * j + 6], r[8 * j + 7]);
              o[j] = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1), *reinterpret_cast<uint32_t *>(&h2),
                                *reinterpret_cast<uint32_t *>(&h3));
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_free(tmem_base, kTmemCols);
  }
}

// ---- tensor maps (driver entry point fetched through the runtime: libcuda is not linked) -------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  if (!fn) fail(MR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  return fn;
}

// rows x K binary16 matrix, row-major (K contiguous); box = {64 elements of K (128 B), box_rows}, 128-byte swizzle,
// out-of-bounds rows read as zero
CUtensorMap make_map(const __half *base, int rows, int K, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  const cuuint32_t elem[2] = {1, 1};
  const CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, dims, strides, box, elem,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(MR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d matrix", (int)r, rows, K);
  return m;
}

int num_sms() {
  static int n = [] { int dev = 0, v = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); return v > 0 ? v : 148; }();
  return n;
}

template <int BLOCK_N> constexpr size_t smem_bytes() {
  return (size_t)StagesFor<BLOCK_N>::value * (kBlockM + BLOCK_N) * kBlockK * 2 + (2 * StagesFor<BLOCK_N>::value + 4) * 8 + 16 + 1024;  // + alignment slack
}

template <int BLOCK_N, int EPI>
void launch(const __half *A, const __half *W, const GemmParams &p, cudaStream_t stream) {
  const CUtensorMap ma = make_map(A, p.M, p.K, kBlockM), mw = make_map(W, p.N, p.K, BLOCK_N);
  constexpr size_t smem = smem_bytes<BLOCK_N>();
  auto kern = encoder_gemm_kernel<BLOCK_N, EPI>;
  const int n_tiles = ((p.M + kBlockM - 1) / kBlockM) * (p.N / BLOCK_N);
  const unsigned grid = (unsigned)std::min(n_tiles, num_sms());
  { ProfScope _ps("encoder_gemm_kernel", stream); kern<<<grid, 320, smem, stream>>>(ma, mw, p); }
  MR_CUDA_CHECK(cudaGetLastError());
  g_kernel_launches++;
}


template <int BLOCK_N, int EPI> void set_smem() {
  MR_CUDA_CHECK(cudaFuncSetAttribute(encoder_gemm_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<BLOCK_N>()));
}
}  // namespace

void encoder_gemm_init() {
  static std::once_flag once;
  std::call_once(once, [] {
    set_smem<128, EPI_BIAS>(); set_smem<128, EPI_GELU>(); set_smem<128, EPI_RESIDUAL>();
    set_smem<64, EPI_BIAS>(); set_smem<64, EPI_GELU>(); set_smem<64, EPI_RESIDUAL>();
    set_smem<192, EPI_BIAS>(); set_smem<192, EPI_GELU>(); set_smem<192, EPI_RESIDUAL>();
  });
}

void encoder_gemm(const __half *A, const __half *W, const float *bias, const float *residual, float *out_f32, __half *out_f16,
                  int M, int N, int K, bool gelu, cudaStream_t stream) {
  if (M <= 0) return;
  if (K % kBlockK != 0 || N % 64 != 0) fail(MR_ERR_INVALID_ARG, "encoder GEMM needs K %% 64 == 0 and N %% 64 == 0 (got N %d, K %d)", N, K);
  encoder_gemm_init();
  GemmParams p{bias, residual, out_f32, out_f16, M, N, K, gelu ? 1 : 0};
  if (gelu && residual) fail(MR_ERR_INVALID_ARG, "encoder GEMM: GELU and a residual in one epilogue are not built");
  const int epi = gelu ? EPI_GELU : residual ? EPI_RESIDUAL : EPI_BIAS;
  // Tile width: 192 columns where N allows it (all BERT-small shapes: 384, 1152, 1536) — a third fewer re-reads of the A
  // row block than 128 (the K = 384 layers are bound by L2 -> SM operand traffic, ncu_r2_summary.md); with only a few row
  // blocks (a handful of queries) 64-column tiles put more SMs on the weight stream instead.  MR_GEMM_BN=64|128|192 forces one.
  const long long row_blocks = (M + kBlockM - 1) / kBlockM;
  static const int forced = [] { const char *e = getenv("MR_GEMM_BN"); return e ? atoi(e) : 0; }();
  int bn = row_blocks * (N / 128 > 0 ? N / 128 : 1) < num_sms() / 2 ? 64 : N % 192 == 0 ? 192 : N % 128 == 0 ? 128 : 64;
  if (forced && N % forced == 0 && (forced == 64 || forced == 128 || forced == 192)) bn = forced;
  if (bn == 192) {
    if (epi == EPI_GELU) launch<192, EPI_GELU>(A, W, p, stream);
    else if (epi == EPI_RESIDUAL) launch<192, EPI_RESIDUAL>(A, W, p, stream);
    else launch<192, EPI_BIAS>(A, W, p, stream);
  } else if (bn == 128) {
    if (epi == EPI_GELU) launch<128, EPI_GELU>(A, W, p, stream);
    else if (epi == EPI_RESIDUAL) launch<128, EPI_RESIDUAL>(A, W, p, stream);
    else launch<128, EPI_BIAS>(A, W, p, stream);
  } else {
    if (epi == EPI_GELU) launch<64, EPI_GELU>(A, W, p, stream);
    else if (epi == EPI_RESIDUAL) launch<64, EPI_RESIDUAL>(A, W, p, stream);
    else launch<64, EPI_BIAS>(A, W, p, stream);
  }
}

}  // namespace mr
