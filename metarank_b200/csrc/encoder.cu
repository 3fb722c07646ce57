// Bi-encoder query forward: OnnxBiEncoder.embed + avgpool (reference S/ml/onnx/sbert/OnnxBiEncoder.scala:13-60) for a
// BERT-shaped sentence encoder.  The reference hands three int64 tensors to an ONNX Runtime session and mean-pools the
// returned last_hidden_state; here the graph is spelled out (HuggingFace BertModel: embeddings + LayerNorm, per layer
// self-attention / output projection / LayerNorm / GELU feed-forward / LayerNorm), dense layers on the tensor cores
// (encoder_gemm.cu), everything between them in the f32 kernels below.
//
// Device layout (M = batch * seq token rows, H hidden, I intermediate):
//   x_f32 [M x H]   residual stream        x_f16 [M x H]  the same rows rounded once, the next GEMM's A operand
//   qkv   [M x 3H]  binary16, bias added    ctx  [M x H]  binary16 attention output
//   tmp   [M x H]   f32: dense + bias + residual, LayerNorm'ed back into x
//   mid   [M x I]   binary16 GELU(dense)
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mr_b200.h"
#include "common.h"
#include "encoder.h"
#include "gbdt_kernels.cuh"
#include "internal.h"
#include "json.h"

namespace mr {
namespace {

constexpr int kMaxPerLane = 32;  // hidden <= 1024

// LayerNorm over one row held by a warp: v[k] = element lane + 32 k.  Biased variance, eps inside the root (torch / ONNX).
__device__ __forceinline__ void warp_layernorm(float (&v)[kMaxPerLane], int per_lane, int H, float eps, const float *gamma,
                                               const float *beta, int lane, float *out_f32, __half *out_f16) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxPerLane; k++) if (k < per_lane) s += v[k];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  const float mean = s / (float)H;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxPerLane; k++) if (k < per_lane) { const float d = v[k] - mean; q += d * d; }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
  const float rstd = 1.0f / sqrtf(q / (float)H + eps);
#pragma unroll
  for (int k = 0; k < kMaxPerLane; k++)
    if (k < per_lane) {
      const int c = lane + 32 * k;
      const float y = (v[k] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      out_f32[c] = y;
      out_f16[c] = __float2half_rn(y);
    }
}

// BertEmbeddings: LayerNorm(word[id] + position[j] + token_type[t]); one warp per token row
__global__ void __launch_bounds__(128) embed_ln_kernel(const int64_t *ids, const int64_t *types, int M, int S, int H, int vocab,
                                                       int n_types, const float *word, const float *pos, const float *type,
                                                       const float *gamma, const float *beta, float eps, float *x_f32,
                                                       __half *x_f16, int *error_flag) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  int64_t id = ids[row], t = types ? types[row] : 0;
  if (id < 0 || id >= vocab || t < 0 || t >= n_types) {  // ONNX Runtime's Gather fails the run on an index out of range
    if (lane == 0) atomicExch(error_flag, 1);
    id = 0; t = 0;
  }
  const int j = row % S, per_lane = H / 32;
  float v[kMaxPerLane];
#pragma unroll
  for (int k = 0; k < kMaxPerLane; k++)
    if (k < per_lane) {
      const int c = lane + 32 * k;
      v[k] = __ldg(word + (size_t)id * H + c) + __ldg(pos + (size_t)j * H + c) + __ldg(type + (size_t)t * H + c);
    }
  warp_layernorm(v, per_lane, H, eps, gamma, beta, lane, x_f32 + (size_t)row * H, x_f16 + (size_t)row * H);
}

// x = LayerNorm(tmp) (tmp already holds dense + bias + residual from the GEMM epilogue)
__global__ void __launch_bounds__(128) ln_kernel(const float *tmp, int M, int H, const float *gamma, const float *beta, float eps,
                                                 float *x_f32, __half *x_f16) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const int per_lane = H / 32;
  float v[kMaxPerLane];
#pragma unroll
  for (int k = 0; k < kMaxPerLane; k++) if (k < per_lane) v[k] = tmp[(size_t)row * H + lane + 32 * k];
  warp_layernorm(v, per_lane, H, eps, gamma, beta, lane, x_f32 + (size_t)row * H, x_f16 + (size_t)row * H);
}

// BertSelfAttention for one (sequence, head): softmax(Q K^T / sqrt(d) + mask) V in f32.  K and V of the head sit in shared
// memory as f32 rows of pitch D + 1 (conflict-free both when lanes walk keys and when lanes walk dimensions); a warp owns a
// query row at a time: lanes stride the keys for the scores, reduce max / sum with shuffles, then lane = output dimension.
// Keys whose attention_mask is 0 get weight exactly 0 — what the exported graph's additive -FLT_MAX mask yields in f32.
// A row whose every key is masked averages all keys (softmax of a constant row), like the graph.
template <int D>
__global__ void __launch_bounds__(128) attention_kernel(const __half *qkv, const int64_t *mask, int S, int H, float scale,
                                                        __half *ctx) {
  extern __shared__ float sm[];
  const int b = blockIdx.y, head = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int P = D + 1;
  float *sk = sm, *sv = sk + (size_t)S * P, *sp = sv + (size_t)S * P;  // sp: 4 warps x S scores
  unsigned char *smask = reinterpret_cast<unsigned char *>(sp + 4 * (size_t)S);
  const __half *base = qkv + (size_t)b * S * 3 * H + head * D;
  for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
    const int j = i / D, c = i % D;
    sk[j * P + c] = __half2float(base[(size_t)j * 3 * H + H + c]);
    sv[j * P + c] = __half2float(base[(size_t)j * 3 * H + 2 * H + c]);
  }
  int any = 0;
  for (int j = threadIdx.x; j < S; j += blockDim.x) { const unsigned char m = mask[(size_t)b * S + j] != 0; smask[j] = m; any |= m; }
  any = __syncthreads_or(any);
  float *my = sp + (size_t)warp * S;
  for (int q = warp; q < S; q += 4) {
    float qr[D];
#pragma unroll
    for (int c = 0; c < D; c++) qr[c] = __half2float(base[(size_t)q * 3 * H + c]);  // same address across the warp: one broadcast load
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < D; c++) s = fmaf(qr[c], sk[j * P + c], s);
      s *= scale;
      if (!any) s = 0.f;  // every key masked: the additive -FLT_MAX absorbs the scores, softmax is uniform
      else if (!smask[j]) s = -INFINITY;
      my[j] = s;
      mx = fmaxf(mx, s);
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = __expf(my[j] - mx); my[j] = e; sum += e; }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    __syncwarp();
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      float acc = 0.f;
      for (int j = 0; j < S; j++) acc = fmaf(my[j], sv[j * P + c0 + lane], acc);
      ctx[((size_t)b * S + q) * H + head * D + c0 + lane] = __float2half_rn(acc * inv);
    }
    __syncwarp();
  }
}

// Queries are short (a search phrase is 5-20 tokens): for seq <= 32 and head dimension 32 a WARP owns one (sequence, head) and
// nothing touches shared memory.  Lane j holds key row j, lane d holds column d of V (all keys), the query row arrives as a
// broadcast load; scores, the two softmax reductions and the weights stay in registers (weights broadcast by shuffle).
// Same operation order as attention_kernel — dot products by ascending dimension, reductions by the same xor tree, the
// weighted sum by ascending key — so both kernels return the same bits.
__global__ void __launch_bounds__(128, 5) attention_short_kernel(const __half *qkv, const int64_t *mask, int n_pairs, int heads, int S,
                                                              int H, float scale, __half *ctx) {
  constexpr int D = 32;
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= n_pairs) return;
  const int b = w / heads, head = w % heads;
  const __half *base = qkv + (size_t)b * S * 3 * H + head * D;
  const bool has = lane < S;
  float k[D], vt[32];
  {
    const uint4 *kp = reinterpret_cast<const uint4 *>(base + (size_t)(has ? lane : 0) * 3 * H + H);
#pragma unroll
    for (int i = 0; i < D / 8; i++) {
      const uint4 u = __ldg(kp + i);
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 f = __half22float2(h[e]); k[8 * i + 2 * e] = f.x; k[8 * i + 2 * e + 1] = f.y; }
    }
  }
#pragma unroll
  for (int j = 0; j < 32; j++) vt[j] = j < S ? __half2float(base[(size_t)j * 3 * H + 2 * H + lane]) : 0.f;
  const bool m = has && mask[(size_t)b * S + (has ? lane : 0)] != 0;
  const bool any = __ballot_sync(0xFFFFFFFFu, m) != 0;
  uint4 qn[D / 8];  // the next query row, in flight while the current one is worked on
#pragma unroll
  for (int i = 0; i < D / 8; i++) qn[i] = __ldg(reinterpret_cast<const uint4 *>(base) + i);
  for (int q = 0; q < S; q++) {
    uint4 qc[D / 8];
#pragma unroll
    for (int i = 0; i < D / 8; i++) qc[i] = qn[i];
    if (q + 1 < S) {
      const uint4 *qp = reinterpret_cast<const uint4 *>(base + (size_t)(q + 1) * 3 * H);
#pragma unroll
      for (int i = 0; i < D / 8; i++) qn[i] = __ldg(qp + i);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < D / 8; i++) {
      const uint4 u = qc[i];
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 f = __half22float2(h[e]); s = fmaf(f.x, k[8 * i + 2 * e], s); s = fmaf(f.y, k[8 * i + 2 * e + 1], s); }
    }
    s *= scale;
    if (!has) s = -INFINITY;
    else if (!any) s = 0.f;
    else if (!m) s = -INFINITY;
    float mx = s;
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    const float e = has ? __expf(s - mx) : 0.f;
    float sum = e;
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    const float inv = 1.0f / sum;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const float pj = __shfl_sync(0xFFFFFFFFu, e, j);
      if (j < S) acc = fmaf(pj, vt[j], acc);
    }
    ctx[((size_t)b * S + q) * H + head * D + lane] = __float2half_rn(acc * inv);
  }
}

// seq <= 16 (the usual search phrase) and head dimension 32: the whole head is two tensor-core tiles.  A warp owns one
// (sequence, head): S = Q K^T as 2 x 2 `mma.sync.m16n8k16` (Q rows = A fragments, K rows = B fragments, both straight from
// global memory: a fragment register is two adjacent binary16 values of one row), softmax on the accumulator fragments
// (a row lives in the four lanes of a quad: two xor-shuffles per reduction), P re-used in place as the A fragment of
// O = P V (the accumulator layout of two 8-column tiles IS the A layout of a 16-deep step), V gathered as B fragments.
// ~120 instructions per head instead of ~3000 scalar ones.  P is rounded to binary16 for the second product — the same
// rounding the output gets anyway; tolerance in tests/test_encoder_gpu.py.  (tcgen05 has no shape this small: its minimum
// tile is 64 x 8 x 16 per CTA and one accumulator round trip through TMEM costs more than this whole head.)
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&h);
}

__global__ void __launch_bounds__(128) attention_mma16_kernel(const __half *qkv, const int64_t *mask, int n_pairs, int heads, int S,
                                                              int H, float scale, __half *ctx) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= n_pairs) return;
  const int b = w / heads, head = w % heads, g = lane >> 2, t = lane & 3;
  const __half *base = qkv + (size_t)b * S * 3 * H + head * 32;
  const size_t pitch = (size_t)3 * H;
  const int r0 = min(g, S - 1), r1 = min(g + 8, S - 1);  // rows / keys beyond the sequence read a valid row; their results are dropped
  auto ld32 = [](const __half *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); };
  // scores
  float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
  for (int ks = 0; ks < 2; ks++) {
    const int d0 = 16 * ks + 2 * t;
    const uint32_t a[4] = {ld32(base + r0 * pitch + d0), ld32(base + r1 * pitch + d0), ld32(base + r0 * pitch + d0 + 8),
                           ld32(base + r1 * pitch + d0 + 8)};
    mma_16816(sc[0], a, ld32(base + r0 * pitch + H + d0), ld32(base + r0 * pitch + H + d0 + 8));  // keys 0..7: key g
    mma_16816(sc[1], a, ld32(base + r1 * pitch + H + d0), ld32(base + r1 * pitch + H + d0 + 8));  // keys 8..15: key g + 8
  }
  // key validity of this lane's four columns: keys 2t, 2t+1, 8+2t, 9+2t
  bool any_l = false, ok[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int j = (i >> 1) * 8 + 2 * t + (i & 1);
    ok[i] = j < S && mask[(size_t)b * S + min(j, S - 1)] != 0;
    any_l |= ok[i];
  }
  const bool any = __ballot_sync(0xFFFFFFFFu, any_l) != 0;
  float p[2][4];  // [key tile][c0 c1 = row g, c2 c3 = row g + 8]
#pragma unroll
  for (int nt = 0; nt < 2; nt++)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int col = nt * 2 + (i & 1), j = nt * 8 + 2 * t + (i & 1);
      float x = sc[nt][i] * scale;
      if (j >= S) x = -INFINITY;
      else if (!any) x = 0.f;
      else if (!ok[col]) x = -INFINITY;
      p[nt][i] = x;
    }
  float inv[2];
#pragma unroll
  for (int hrow = 0; hrow < 2; hrow++) {  // row g (values 0, 1 of both tiles), row g + 8 (values 2, 3)
    float mx = fmaxf(fmaxf(p[0][2 * hrow], p[0][2 * hrow + 1]), fmaxf(p[1][2 * hrow], p[1][2 * hrow + 1]));
    mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int i = 0; i < 2; i++) { const float e = __expf(p[nt][2 * hrow + i] - mx); p[nt][2 * hrow + i] = e; sum += e; }
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, 1);
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, 2);
    inv[hrow] = 1.0f / sum;
  }
  const uint32_t pa[4] = {pack_half2(p[0][0], p[0][1]), pack_half2(p[0][2], p[0][3]), pack_half2(p[1][0], p[1][1]),
                          pack_half2(p[1][2], p[1][3])};
  // O = P V: B fragment of dimension tile dt = {V[2t][d], V[2t+1][d]}, {V[2t+8][d], V[2t+9][d]} with d = 8 dt + g
  const unsigned short *vb = reinterpret_cast<const unsigned short *>(base + 2 * H);
  const size_t k0 = (size_t)min(2 * t, S - 1) * pitch, k1 = (size_t)min(2 * t + 1, S - 1) * pitch,
               k2 = (size_t)min(2 * t + 8, S - 1) * pitch, k3 = (size_t)min(2 * t + 9, S - 1) * pitch;
#pragma unroll
  for (int dt = 0; dt < 4; dt++) {
    const int d = 8 * dt + g;
    const uint32_t b0 = (uint32_t)__ldg(vb + k0 + d) | (uint32_t)__ldg(vb + k1 + d) << 16;
    const uint32_t b1 = (uint32_t)__ldg(vb + k2 + d) | (uint32_t)__ldg(vb + k3 + d) << 16;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    mma_16816(o, pa, b0, b1);
    const int col = head * 32 + 8 * dt + 2 * t;
    if (g < S) *reinterpret_cast<uint32_t *>(ctx + ((size_t)b * S + g) * H + col) = pack_half2(o[0] * inv[0], o[1] * inv[0]);
    if (g + 8 < S) *reinterpret_cast<uint32_t *>(ctx + ((size_t)b * S + g + 8) * H + col) = pack_half2(o[2] * inv[1], o[3] * inv[1]);
  }
}

// OnnxBiEncoder.avgpool (:38-60): per dimension, a double sum over the first tokenLengths = sum(attention_mask) tokens
// divided by their count, narrowed to float; 0 tokens -> 0.0 / 0 = NaN, as in the reference.
__global__ void __launch_bounds__(128) meanpool_kernel(const float *x, const int64_t *mask, int S, int H, float *out, double *out_f64) {
  __shared__ int s_len;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_len = 0;
  __syncthreads();
  int part = 0;
  for (int j = threadIdx.x; j < S; j += blockDim.x) part += (int)mask[(size_t)b * S + j];
  atomicAdd(&s_len, part);
  __syncthreads();
  int len = s_len;
  if (len > S) len = S;
  if (len < 0) len = 0;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    double sum = 0.0;
    for (int j = 0; j < len; j++) sum += (double)x[((size_t)b * S + j) * H + c];
    const float e = (float)(sum / (double)len);
    out[(size_t)b * H + c] = e;
    if (out_f64) out_f64[(size_t)b * H + c] = (double)e;
  }
}

__global__ void f32_to_f16_kernel(const float *in, __half *out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __float2half_rn(in[i]);
}

// ---- safetensors: u64 little-endian header length, JSON header {name: {dtype, shape, data_offsets: [lo, hi]}}, raw bytes
struct TensorView {
  std::string dtype;
  std::vector<int64_t> shape;
  const uint8_t *data = nullptr;
  size_t bytes = 0;
  size_t count() const { size_t n = 1; for (auto d : shape) n *= (size_t)d; return n; }
};

float half_bits_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 31, man = h & 1023;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else { int e = -1; uint32_t m = man; do { e++; m <<= 1; } while (!(m & 1024)); bits = sign | (uint32_t)(127 - 15 - e) << 23 | (m & 1023) << 13; }
  } else if (exp == 31) bits = sign | 0x7F800000u | man << 13;
  else bits = sign | (exp + 112) << 23 | man << 13;
  float f; memcpy(&f, &bits, 4); return f;
}

std::map<std::string, TensorView> parse_safetensors(const uint8_t *p, size_t len) {
  if (!p || len < 8) fail(MR_ERR_PARSE, "safetensors: shorter than its 8-byte header length");
  uint64_t hl = 0;
  for (int i = 7; i >= 0; i--) hl = hl << 8 | p[i];
  if (hl > len - 8 || hl > (1ull << 27)) fail(MR_ERR_PARSE, "safetensors: header length %llu exceeds the file", (unsigned long long)hl);
  JsonParser jp(p + 8, (size_t)hl);
  const JValue root = jp.parse();
  if (root.kind != JValue::Obj) fail(MR_ERR_PARSE, "safetensors: header is not a JSON object");
  const uint8_t *data = p + 8 + hl;
  const size_t data_len = len - 8 - (size_t)hl;
  std::map<std::string, TensorView> out;
  for (auto &kv : root.obj) {
    if (kv.first == "__metadata__") continue;
    TensorView t;
    t.dtype = kv.second.at("dtype").str;
    for (auto &d : kv.second.at("shape").arr) { const int64_t v = d.as_int(); if (v < 0) fail(MR_ERR_PARSE, "safetensors: negative dimension in '%s'", kv.first.c_str()); t.shape.push_back(v); }
    const JValue &off = kv.second.at("data_offsets");
    if (off.kind != JValue::Arr || off.arr.size() != 2) fail(MR_ERR_PARSE, "safetensors: bad data_offsets of '%s'", kv.first.c_str());
    const int64_t lo = off.arr[0].as_int(), hi = off.arr[1].as_int();
    if (lo < 0 || hi < lo || (uint64_t)hi > data_len) fail(MR_ERR_PARSE, "safetensors: tensor '%s' lies outside the file", kv.first.c_str());
    t.data = data + lo;
    t.bytes = (size_t)(hi - lo);
    const size_t esz = t.dtype == "F32" ? 4 : t.dtype == "F16" ? 2 : t.dtype == "I64" ? 8 : 0;
    if (esz && t.count() * esz != t.bytes) fail(MR_ERR_PARSE, "safetensors: tensor '%s' has %zu bytes for its shape", kv.first.c_str(), t.bytes);
    std::string name = kv.first;
    if (name.rfind("bert.", 0) == 0) name = name.substr(5);
    out[name] = std::move(t);
  }
  return out;
}

std::vector<float> as_f32(const std::map<std::string, TensorView> &T, const std::string &name, const char *legacy,
                          std::initializer_list<int64_t> shape) {
  auto it = T.find(name);
  if (it == T.end() && legacy) it = T.find(legacy);
  if (it == T.end()) fail(MR_ERR_PARSE, "encoder weights: tensor '%s' is missing", name.c_str());
  const TensorView &t = it->second;
  if (t.shape != std::vector<int64_t>(shape)) fail(MR_ERR_PARSE, "encoder weights: tensor '%s' has an unexpected shape", name.c_str());
  std::vector<float> v(t.count());
  if (t.dtype == "F32") memcpy(v.data(), t.data, t.bytes);
  else if (t.dtype == "F16") { for (size_t i = 0; i < v.size(); i++) { uint16_t h; memcpy(&h, t.data + 2 * i, 2); v[i] = half_bits_to_float(h); } }
  else fail(MR_ERR_UNSUPPORTED, "encoder weights: tensor '%s' is %s; F32 and F16 are read", name.c_str(), t.dtype.c_str());
  return v;
}

}  // namespace
}  // namespace mr

using namespace mr;

struct mr_encoder {
  mr_ctx *ctx = nullptr;
  int H = 0, L = 0, I = 0, heads = 0, D = 0, vocab = 0, max_pos = 0, n_types = 0;
  float eps = 1e-12f;
  // f32 tables and vectors in one allocation, binary16 matrices in another
  float *d_f32 = nullptr;
  __half *d_f16 = nullptr;
  size_t f32_count = 0, f16_count = 0;
  const float *word, *pos, *type, *emb_g, *emb_b;
  struct Layer {
    const __half *w_qkv, *w_o, *w_1, *w_2;
    const float *b_qkv, *b_o, *b_1, *b_2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  };
  std::vector<Layer> layers;
  // workspace, grown on demand; one forward at a time
  std::mutex mu;
  size_t ws_rows = 0;
  float *x_f32 = nullptr, *tmp = nullptr;
  __half *x_f16 = nullptr, *qkv = nullptr, *ctxb = nullptr, *mid = nullptr;
  int *d_error = nullptr;
  int *h_error = nullptr;  // pinned
  // host-entry staging
  int64_t *d_in = nullptr; size_t d_in_cap = 0;
  float *d_out = nullptr; size_t d_out_cap = 0;
  cudaStream_t stream = nullptr;
  // A forward is 7 L + 2 launches of a few microseconds each: for one query the launches ARE the latency.  Each
  // (buffers, batch, seq) shape is captured once into a CUDA graph and replayed (tensor maps and pointers are baked
  // into the nodes, so the key holds every pointer; growing the workspace drops the cache).
  struct GraphKey {
    const void *ids, *types, *mask, *out, *out64;
    int B, S;
    bool operator==(const GraphKey &o) const { return ids == o.ids && types == o.types && mask == o.mask && out == o.out && out64 == o.out64 && B == o.B && S == o.S; }
  };
  struct GraphEntry { GraphKey key; cudaGraphExec_t exec; uint64_t used; };
  std::vector<GraphEntry> graphs;
  uint64_t tick = 0;
  static constexpr size_t kMaxGraphs = 16;

  void drop_graphs() {
    for (auto &g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }

  ~mr_encoder() {
    drop_graphs();
    cudaFree(d_f32); cudaFree(d_f16); cudaFree(x_f32); cudaFree(tmp); cudaFree(x_f16); cudaFree(qkv); cudaFree(ctxb); cudaFree(mid);
    cudaFree(d_error); cudaFree(d_in); cudaFree(d_out);
    if (h_error) cudaFreeHost(h_error);
    if (stream) cudaStreamDestroy(stream);
  }

  void reserve(size_t rows) {
    if (rows <= ws_rows) return;
    MR_CUDA_CHECK(cudaDeviceSynchronize());
    drop_graphs();
    cudaFree(x_f32); cudaFree(tmp); cudaFree(x_f16); cudaFree(qkv); cudaFree(ctxb); cudaFree(mid);
    x_f32 = tmp = nullptr; x_f16 = qkv = ctxb = mid = nullptr; ws_rows = 0;
    const size_t r = (rows + 127) & ~(size_t)127;
    MR_CUDA_CHECK(cudaMalloc(&x_f32, r * H * 4));
    MR_CUDA_CHECK(cudaMalloc(&tmp, r * H * 4));
    MR_CUDA_CHECK(cudaMalloc(&x_f16, r * H * 2));
    MR_CUDA_CHECK(cudaMalloc(&qkv, r * 3 * H * 2));
    MR_CUDA_CHECK(cudaMalloc(&ctxb, r * H * 2));
    MR_CUDA_CHECK(cudaMalloc(&mid, r * I * 2));
    ws_rows = r;
  }

  void run(const int64_t *ids, const int64_t *types, const int64_t *mask, int B, int S, float *out, double *out_f64, cudaStream_t st) {
    reserve((size_t)B * S);
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (st == nullptr || profile_active() || cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      forward(ids, types, mask, B, S, out, out_f64, st);  // legacy stream / open profile / caller is capturing: plain launches
      return;
    }
    const GraphKey key{ids, types, mask, out, out_f64, B, S};
    for (auto &g : graphs)
      if (g.key == key) {
        g.used = ++tick;
        MR_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
        g_kernel_launches += 7 * L + 2;
        return;
      }
    const long long before = g_kernel_launches;
    MR_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    cudaGraph_t graph = nullptr;
    try {
      forward(ids, types, mask, B, S, out, out_f64, st);
    } catch (...) {
      cudaStreamEndCapture(st, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    g_kernel_launches = before;
    MR_CUDA_CHECK(cudaStreamEndCapture(st, &graph));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t err = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    MR_CUDA_CHECK(err);
    if (graphs.size() >= kMaxGraphs) {
      size_t lru = 0;
      for (size_t i = 1; i < graphs.size(); i++) if (graphs[i].used < graphs[lru].used) lru = i;
      cudaGraphExecDestroy(graphs[lru].exec);
      graphs.erase(graphs.begin() + lru);
    }
    graphs.push_back({key, exec, ++tick});
    MR_CUDA_CHECK(cudaGraphLaunch(exec, st));
    g_kernel_launches += 7 * L + 2;
  }

  void forward(const int64_t *ids, const int64_t *types, const int64_t *mask, int B, int S, float *out, double *out_f64,
               cudaStream_t st) {
    const int M = B * S;
    { ProfScope _ps("embed_ln_kernel", st);
      embed_ln_kernel<<<(M + 3) / 4, 128, 0, st>>>(ids, types, M, S, H, vocab, n_types, word, pos, type, emb_g, emb_b, eps, x_f32, x_f16, d_error); }
    g_kernel_launches++;
    const size_t att_smem = ((size_t)2 * S * (D + 1) + 4 * (size_t)S) * 4 + (size_t)S;
    const float scale = 1.0f / sqrtf((float)D);
    for (const Layer &ly : layers) {
      encoder_gemm(x_f16, ly.w_qkv, ly.b_qkv, nullptr, nullptr, qkv, M, 3 * H, H, false, st);
      { ProfScope _ps("attention_kernel", st);
        // a warp per head pays once there are enough heads to fill the chip
        if (D == 32 && S <= 16 && B * heads >= 1024) attention_mma16_kernel<<<(B * heads + 3) / 4, 128, 0, st>>>(qkv, mask, B * heads, heads, S, H, scale, ctxb);
        else if (D == 32 && S <= 32 && B * heads >= 4096) attention_short_kernel<<<(B * heads + 3) / 4, 128, 0, st>>>(qkv, mask, B * heads, heads, S, H, scale, ctxb);
        else if (D == 32) attention_kernel<32><<<dim3(heads, B), 128, att_smem, st>>>(qkv, mask, S, H, scale, ctxb);
        else attention_kernel<64><<<dim3(heads, B), 128, att_smem, st>>>(qkv, mask, S, H, scale, ctxb); }
      g_kernel_launches++;
      encoder_gemm(ctxb, ly.w_o, ly.b_o, x_f32, tmp, nullptr, M, H, H, false, st);
      { ProfScope _ps("ln_kernel", st); ln_kernel<<<(M + 3) / 4, 128, 0, st>>>(tmp, M, H, ly.ln1_g, ly.ln1_b, eps, x_f32, x_f16); }
      encoder_gemm(x_f16, ly.w_1, ly.b_1, nullptr, nullptr, mid, M, I, H, true, st);
      encoder_gemm(mid, ly.w_2, ly.b_2, x_f32, tmp, nullptr, M, H, I, false, st);
      { ProfScope _ps("ln_kernel", st); ln_kernel<<<(M + 3) / 4, 128, 0, st>>>(tmp, M, H, ly.ln2_g, ly.ln2_b, eps, x_f32, x_f16); }
      g_kernel_launches += 2;
    }
    { ProfScope _ps("meanpool_kernel", st); meanpool_kernel<<<B, 128, 0, st>>>(x_f32, mask, S, H, out, out_f64); }
    g_kernel_launches++;
    MR_CUDA_CHECK(cudaGetLastError());
  }
};

namespace {

void check_shape(const mr_encoder *e, int32_t batch, int32_t seq) {
  if (!e) fail(MR_ERR_INVALID_ARG, "encoder handle is null");
  if (batch <= 0 || seq <= 0) fail(MR_ERR_INVALID_ARG, "embed needs batch >= 1 and seq >= 1 (got %d x %d)", batch, seq);
  if (seq > e->max_pos) fail(MR_ERR_INVALID_ARG, "sequence of %d tokens exceeds the model's %d positions", seq, e->max_pos);
  if ((int64_t)batch * seq > (1 << 24)) fail(MR_ERR_INVALID_ARG, "embed of %lld token rows exceeds 16 M", (long long)batch * seq);
}

}  // namespace

extern "C" {

mr_status mr_encoder_load(mr_ctx *ctx, const uint8_t *blob, size_t len, int32_t n_heads, double layer_norm_eps, mr_encoder **out) {
  return guard([&] {
    if (!ctx || !out) fail(MR_ERR_INVALID_ARG, "ctx / out is null");
    *out = nullptr;
    const auto T = parse_safetensors(blob, len);
    auto it = T.find("embeddings.word_embeddings.weight");
    if (it == T.end() || it->second.shape.size() != 2) fail(MR_ERR_PARSE, "encoder weights: 'embeddings.word_embeddings.weight' [vocab x hidden] is missing");
    auto e = std::make_unique<mr_encoder>();
    e->ctx = ctx;
    e->vocab = (int)it->second.shape[0];
    e->H = (int)it->second.shape[1];
    auto pit = T.find("embeddings.position_embeddings.weight");
    auto tit = T.find("embeddings.token_type_embeddings.weight");
    if (pit == T.end() || tit == T.end() || pit->second.shape.size() != 2 || tit->second.shape.size() != 2)
      fail(MR_ERR_PARSE, "encoder weights: position / token_type embeddings are missing");
    e->max_pos = (int)pit->second.shape[0];
    e->n_types = (int)tit->second.shape[0];
    while (T.count("encoder.layer." + std::to_string(e->L) + ".attention.self.query.weight")) e->L++;
    if (e->L == 0) fail(MR_ERR_PARSE, "encoder weights: no 'encoder.layer.0.attention.self.query.weight'");
    auto iit = T.find("encoder.layer.0.intermediate.dense.weight");
    if (iit == T.end() || iit->second.shape.size() != 2) fail(MR_ERR_PARSE, "encoder weights: 'encoder.layer.0.intermediate.dense.weight' is missing");
    e->I = (int)iit->second.shape[0];
    e->heads = n_heads;
    if (n_heads <= 0 || e->H % n_heads) fail(MR_ERR_INVALID_ARG, "hidden size %d is not divisible by %d heads", e->H, n_heads);
    e->D = e->H / n_heads;
    if (e->D != 32 && e->D != 64) fail(MR_ERR_UNSUPPORTED, "head dimension %d: 32 and 64 are built", e->D);
    if (e->H % 64 || e->I % 64 || e->H > 32 * kMaxPerLane) fail(MR_ERR_UNSUPPORTED, "hidden %d / intermediate %d: multiples of 64, hidden <= 1024", e->H, e->I);
    if (!(layer_norm_eps > 0) || layer_norm_eps > 1e-3) fail(MR_ERR_INVALID_ARG, "layer_norm_eps %g out of range", layer_norm_eps);
    e->eps = (float)layer_norm_eps;
    const int64_t H = e->H, I = e->I;

    // gather: f32 pack (tables, biases, LayerNorm), matrices as f32 staged for the on-device binary16 rounding
    std::vector<float> f32;
    std::vector<float> mats;
    auto put = [&](std::vector<float> &dst, const std::vector<float> &v) { const size_t o = dst.size(); dst.insert(dst.end(), v.begin(), v.end()); while (dst.size() % 64) dst.push_back(0.f); return o; };
    const size_t o_word = put(f32, as_f32(T, "embeddings.word_embeddings.weight", nullptr, {e->vocab, H}));
    const size_t o_pos = put(f32, as_f32(T, "embeddings.position_embeddings.weight", nullptr, {e->max_pos, H}));
    const size_t o_type = put(f32, as_f32(T, "embeddings.token_type_embeddings.weight", nullptr, {e->n_types, H}));
    const size_t o_eg = put(f32, as_f32(T, "embeddings.LayerNorm.weight", "embeddings.LayerNorm.gamma", {H}));
    const size_t o_eb = put(f32, as_f32(T, "embeddings.LayerNorm.bias", "embeddings.LayerNorm.beta", {H}));
    struct Off { size_t w_qkv, w_o, w_1, w_2, b_qkv, b_o, b_1, b_2, g1, be1, g2, be2; };
    std::vector<Off> offs;
    for (int l = 0; l < e->L; l++) {
      const std::string p = "encoder.layer." + std::to_string(l) + ".";
      Off o{};
      std::vector<float> wqkv, bqkv;
      for (const char *n : {"query", "key", "value"}) {
        auto w = as_f32(T, p + "attention.self." + n + ".weight", nullptr, {H, H});
        auto b = as_f32(T, p + "attention.self." + n + ".bias", nullptr, {H});
        wqkv.insert(wqkv.end(), w.begin(), w.end());
        bqkv.insert(bqkv.end(), b.begin(), b.end());
      }
      o.w_qkv = put(mats, wqkv);
      o.b_qkv = put(f32, bqkv);
      o.w_o = put(mats, as_f32(T, p + "attention.output.dense.weight", nullptr, {H, H}));
      o.b_o = put(f32, as_f32(T, p + "attention.output.dense.bias", nullptr, {H}));
      o.g1 = put(f32, as_f32(T, p + "attention.output.LayerNorm.weight", (p + "attention.output.LayerNorm.gamma").c_str(), {H}));
      o.be1 = put(f32, as_f32(T, p + "attention.output.LayerNorm.bias", (p + "attention.output.LayerNorm.beta").c_str(), {H}));
      o.w_1 = put(mats, as_f32(T, p + "intermediate.dense.weight", nullptr, {I, H}));
      o.b_1 = put(f32, as_f32(T, p + "intermediate.dense.bias", nullptr, {I}));
      o.w_2 = put(mats, as_f32(T, p + "output.dense.weight", nullptr, {H, I}));
      o.b_2 = put(f32, as_f32(T, p + "output.dense.bias", nullptr, {H}));
      o.g2 = put(f32, as_f32(T, p + "output.LayerNorm.weight", (p + "output.LayerNorm.gamma").c_str(), {H}));
      o.be2 = put(f32, as_f32(T, p + "output.LayerNorm.bias", (p + "output.LayerNorm.beta").c_str(), {H}));
      offs.push_back(o);
    }
    MR_CUDA_CHECK(cudaSetDevice(ctx->device));
    encoder_gemm_init();
    MR_CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->f32_count = f32.size();
    e->f16_count = mats.size();
    MR_CUDA_CHECK(cudaMalloc(&e->d_f32, f32.size() * 4));
    MR_CUDA_CHECK(cudaMalloc(&e->d_f16, mats.size() * 2));
    MR_CUDA_CHECK(cudaMemcpy(e->d_f32, f32.data(), f32.size() * 4, cudaMemcpyHostToDevice));
    {
      float *stage = nullptr;
      MR_CUDA_CHECK(cudaMalloc(&stage, mats.size() * 4));
      cudaError_t err = cudaMemcpy(stage, mats.data(), mats.size() * 4, cudaMemcpyHostToDevice);
      if (err == cudaSuccess) {
        f32_to_f16_kernel<<<ctx->num_sms * 4, 256, 0, e->stream>>>(stage, e->d_f16, mats.size());
        err = cudaStreamSynchronize(e->stream);
      }
      cudaFree(stage);
      MR_CUDA_CHECK(err);
    }
    MR_CUDA_CHECK(cudaMalloc(&e->d_error, 4));
    MR_CUDA_CHECK(cudaMemset(e->d_error, 0, 4));
    MR_CUDA_CHECK(cudaMallocHost(&e->h_error, 4));
    e->word = e->d_f32 + o_word; e->pos = e->d_f32 + o_pos; e->type = e->d_f32 + o_type;
    e->emb_g = e->d_f32 + o_eg; e->emb_b = e->d_f32 + o_eb;
    for (auto &o : offs)
      e->layers.push_back({e->d_f16 + o.w_qkv, e->d_f16 + o.w_o, e->d_f16 + o.w_1, e->d_f16 + o.w_2, e->d_f32 + o.b_qkv, e->d_f32 + o.b_o,
                           e->d_f32 + o.b_1, e->d_f32 + o.b_2, e->d_f32 + o.g1, e->d_f32 + o.be1, e->d_f32 + o.g2, e->d_f32 + o.be2});
    // attention keeps a head's K and V in shared memory: 2 * max_pos * (D + 1) f32 + score rows
    const size_t att_max = ((size_t)2 * e->max_pos * (e->D + 1) + 4 * (size_t)e->max_pos) * 4 + (size_t)e->max_pos;
    if (att_max > 227 * 1024) fail(MR_ERR_UNSUPPORTED, "%d positions x head dimension %d does not fit the attention kernel's shared memory", e->max_pos, e->D);
    if (e->D == 32) MR_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_max));
    else MR_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_max));
    *out = e.release();
  });
}

mr_status mr_encoder_info(const mr_encoder *e, int32_t *dim, int32_t *layers, int32_t *max_tokens, int32_t *vocab) {
  return guard([&] {
    if (!e) fail(MR_ERR_INVALID_ARG, "encoder handle is null");
    if (dim) *dim = e->H;
    if (layers) *layers = e->L;
    if (max_tokens) *max_tokens = e->max_pos;
    if (vocab) *vocab = e->vocab;
  });
}

mr_status mr_encoder_embed_device(mr_encoder *e, const int64_t *d_ids, const int64_t *d_types, const int64_t *d_mask, int32_t batch,
                                  int32_t seq, float *d_out, double *d_out_f64, void *cuda_stream) {
  return guard([&] {
    check_shape(e, batch, seq);
    if (!d_ids || !d_mask || !d_out) fail(MR_ERR_INVALID_ARG, "input_ids / attention_mask / out is null");
    std::lock_guard<std::mutex> g(e->mu);
    MR_CUDA_CHECK(cudaSetDevice(e->ctx->device));
    e->run(d_ids, d_types, d_mask, batch, seq, d_out, d_out_f64, (cudaStream_t)cuda_stream);
  });
}

mr_status mr_encoder_embed(mr_encoder *e, const int64_t *ids, const int64_t *types, const int64_t *mask, int32_t batch, int32_t seq,
                           float *out) {
  return guard([&] {
    check_shape(e, batch, seq);
    if (!ids || !mask || !out) fail(MR_ERR_INVALID_ARG, "input_ids / attention_mask / out is null");
    std::lock_guard<std::mutex> g(e->mu);
    MR_CUDA_CHECK(cudaSetDevice(e->ctx->device));
    const size_t n = (size_t)batch * seq;
    if (3 * n > e->d_in_cap) { MR_CUDA_CHECK(cudaStreamSynchronize(e->stream)); cudaFree(e->d_in); e->d_in = nullptr; e->d_in_cap = 0; MR_CUDA_CHECK(cudaMalloc(&e->d_in, 3 * n * 8)); e->d_in_cap = 3 * n; }
    const size_t no = (size_t)batch * e->H;
    if (no > e->d_out_cap) { MR_CUDA_CHECK(cudaStreamSynchronize(e->stream)); cudaFree(e->d_out); e->d_out = nullptr; e->d_out_cap = 0; MR_CUDA_CHECK(cudaMalloc(&e->d_out, no * 4)); e->d_out_cap = no; }
    cudaStream_t st = e->stream;
    MR_CUDA_CHECK(cudaMemcpyAsync(e->d_in, ids, n * 8, cudaMemcpyHostToDevice, st));
    if (types) MR_CUDA_CHECK(cudaMemcpyAsync(e->d_in + n, types, n * 8, cudaMemcpyHostToDevice, st));
    MR_CUDA_CHECK(cudaMemcpyAsync(e->d_in + 2 * n, mask, n * 8, cudaMemcpyHostToDevice, st));
    e->run(e->d_in, types ? e->d_in + n : nullptr, e->d_in + 2 * n, batch, seq, e->d_out, nullptr, st);
    MR_CUDA_CHECK(cudaMemcpyAsync(out, e->d_out, no * 4, cudaMemcpyDeviceToHost, st));
    MR_CUDA_CHECK(cudaMemcpyAsync(e->h_error, e->d_error, 4, cudaMemcpyDeviceToHost, st));
    MR_CUDA_CHECK(cudaStreamSynchronize(st));
    if (*e->h_error) {
      MR_CUDA_CHECK(cudaMemsetAsync(e->d_error, 0, 4, st));
      MR_CUDA_CHECK(cudaStreamSynchronize(st));
      fail(MR_ERR_INVALID_ARG, "input_ids / token_type_ids hold an index outside the model's %d-token vocabulary / %d types", e->vocab, e->n_types);
    }
  });
}

mr_status mr_encoder_free(mr_encoder *e) {
  return guard([&] {
    if (!e) return;
    cudaSetDevice(e->ctx->device);
    cudaDeviceSynchronize();
    delete e;
  });
}

mr_status mr_encoder_gemm_f16(mr_ctx *ctx, const void *d_a, const void *d_w, const float *d_bias, const float *d_residual,
                              float *d_out_f32, void *d_out_f16, int32_t m, int32_t n, int32_t k, int32_t gelu, void *cuda_stream) {
  return guard([&] {
    if (!ctx || !d_a || !d_w) fail(MR_ERR_INVALID_ARG, "ctx / a / w is null");
    if (m < 0 || n <= 0 || k <= 0) fail(MR_ERR_INVALID_ARG, "bad GEMM shape %d x %d x %d", m, n, k);
    MR_CUDA_CHECK(cudaSetDevice(ctx->device));
    encoder_gemm((const __half *)d_a, (const __half *)d_w, d_bias, d_residual, d_out_f32, (__half *)d_out_f16, m, n, k, gelu != 0,
                 (cudaStream_t)cuda_stream);
  });
}

}  // extern "C"
