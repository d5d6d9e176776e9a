// Microbenchmark: does a CTA with 5 warps load the 4 SM sub-partitions unevenly (warp id % 4 -> scheduler)?
// Same total work and 20 resident warps per SM either as 5 CTAs x 4 warps or as 4 CTAs x 5 warps.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void fma_loop(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = 1.0001f, c = 0.5f, d = 0.25f, e = 0.125f;
  for (int i = 0; i < iters; ++i) {
    a = fmaf(a, b, c); c = fmaf(c, b, d); d = fmaf(d, b, e); e = fmaf(e, b, a);
  }
  if (a + c + d + e == 123.456f) out[0] = a;
}

int main() {
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 200000;
  struct { int threads, ctas; } cfg[] = {{128, 5}, {160, 4}, {96, 6}, {192, 3}, {256, 2}, {64, 10}, {32, 20}};
  for (auto c : cfg) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      fma_loop<<<148 * c.ctas, c.threads>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("threads %3d x %2d CTAs/SM = %2d warps/SM : %.3f ms  (%.3f ms per warp-per-SM)\n", c.threads, c.ctas,
                      c.threads / 32 * c.ctas, ms, ms / (c.threads / 32 * c.ctas));
    }
  }
  return 0;
}
