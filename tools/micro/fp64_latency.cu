// Dependent-issue latency of the instructions the in-order sums are made of (B200): DADD, DMUL, F2F.F64.F32, and the
// LDS.64 -> DADD pair of gbdt_sum_kernel's consumer.  nvcc -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, const double *in, int n) {
  __shared__ double s[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = in[i];
  __syncthreads();
  double a = in[0];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = __dadd_rn(a, 1.0);
  long long t1 = clock64();
  double b = in[1];
  for (int i = 0; i < n; i++) b = __dmul_rn(b, 1.0000001);
  long long t2 = clock64();
  double c = in[2];
  for (int i = 0; i < n; i++) c = __dadd_rn(c, s[(i * 32 + threadIdx.x) & 2047]);
  long long t3 = clock64();
  float f = (float)in[3];
  double d = 0;
  for (int i = 0; i < n; i++) { d = __dadd_rn(d, (double)f); f = (float)d; }
  long long t4 = clock64();
  out[threadIdx.x] = a + b + c + d;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main() {
  double *out, *in; long long *cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&in, 8 * 2048); cudaMalloc(&cyc, 64);
  cudaMemset(in, 0, 8 * 2048);
  const int n = 4096;
  for (int warps : {1, 4, 8, 16}) {
    k<<<1, 32 * warps>>>(out, cyc, in, n);
    long long h[4];
    cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d: DADD chain %.1f cyc/op, DMUL chain %.1f, LDS.64+DADD chain %.1f, DADD+F2F(f64->f32->f64) loop %.1f\n", warps,
           (double)h[0] / n, (double)h[1] / n, (double)h[2] / n, (double)h[3] / n);
  }
  return 0;
}
