// Microbenchmark: issue/throughput of packed fp32x2 arithmetic (FMUL2/FFMA2, sm_100) against scalar FMUL/FFMA.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void scalar_mul(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f + 1.0f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float b = 0.9999f;
  for (int i = 0; i < iters; ++i) {
    a0 = __fmul_rn(a0, b); a1 = __fmul_rn(a1, b); a2 = __fmul_rn(a2, b); a3 = __fmul_rn(a3, b);
    a4 = __fmul_rn(a4, b); a5 = __fmul_rn(a5, b); a6 = __fmul_rn(a6, b); a7 = __fmul_rn(a7, b);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 123.456f) out[0] = a0;
}
__global__ void packed_mul(float* out, int iters) {
  float2 a0 = make_float2(threadIdx.x * 1e-3f + 1.0f, 2.f), a1 = a0, a2 = a0, a3 = a0;
  a1.x += 1; a2.x += 2; a3.x += 3;
  const float2 b = make_float2(0.9999f, 0.9998f);
  for (int i = 0; i < iters; ++i) {
    a0 = __fmul2_rn(a0, b); a1 = __fmul2_rn(a1, b); a2 = __fmul2_rn(a2, b); a3 = __fmul2_rn(a3, b);
  }
  if (a0.x + a1.x + a2.x + a3.x + a0.y + a1.y + a2.y + a3.y == 123.456f) out[0] = a0.x;
}
__global__ void packed_mul8(float* out, int iters) {  // same number of INSTRUCTIONS as scalar_mul, twice the flops
  float2 a[8];
  for (int k = 0; k < 8; ++k) a[k] = make_float2(threadIdx.x * 1e-3f + 1.0f + k, 2.f + k);
  const float2 b = make_float2(0.9999f, 0.9998f);
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __fmul2_rn(a[k], b);
  float s = 0;
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
  if (s == 123.456f) out[0] = s;
}
__global__ void packed_fma8(float* out, int iters) {
  float2 a[8];
  for (int k = 0; k < 8; ++k) a[k] = make_float2(threadIdx.x * 1e-3f + 1.0f + k, 2.f + k);
  const float2 b = make_float2(0.9999f, 0.9998f), c = make_float2(0.1f, 0.2f);
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __ffma2_rn(a[k], b, c);
  float s = 0;
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
  if (s == 123.456f) out[0] = s;
}

template <typename K>
void run(const char* name, K kern, float* out, int iters, double flops_per_iter) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    kern<<<148 * 4, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("%-12s %.3f ms  %.2f Tflop/s(mul or fma counted once per element)\n", name, ms,
                    flops_per_iter * iters * 148 * 4 * 256 / (ms * 1e-3) / 1e12);
  }
}
int main() {
  float* out; cudaMalloc(&out, 4);
  const int iters = 100000;
  run("scalar_mul", scalar_mul, out, iters, 8);
  run("packed_mul", packed_mul, out, iters, 8);
  run("packed_mul8", packed_mul8, out, iters, 16);
  run("packed_fma8", packed_fma8, out, iters, 16);
  return 0;
}
