// Bi-encoder query forward: the transformer behind OnnxBiEncoder.embed
// (reference S/ml/onnx/sbert/OnnxBiEncoder.scala:13-36) as sm_100a kernels — dense layers on tcgen05 tensor cores
// (encoder_gemm.cu), embedding / LayerNorm / attention / mean-pool as fused f32 kernels (encoder.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace mr {

// C[M x N] = act(A[M x K] W[N x K]^T + bias) + residual; A, W binary16 row-major (K contiguous, 16-byte aligned rows),
// f32 accumulation in tensor memory.  K % 64 == 0, N % 64 == 0; any M >= 1.  out_f32 / out_f16 / bias / residual may be null.
void encoder_gemm(const __half *A, const __half *W, const float *bias, const float *residual, float *out_f32, __half *out_f16,
                  int M, int N, int K, bool gelu, cudaStream_t stream);

// Sets the kernels' shared-memory attributes (not a stream operation: done once, outside any graph capture).
void encoder_gemm_init();

}  // namespace mr
