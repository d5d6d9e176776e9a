// Shared-memory bandwidth microbenchmark (the roofline denominator for the shared-memory-resident BP kernel; SURVEY.md
// 8(d) asks for a measured figure because MEASURED_PEAKS.json has none).  Conflict-free 128-bit accesses, all SMs.
//   read : LDS.128 only             copy : LDS.128 + STS.128 (the BP kernel's 1:1 mix)
// Prints JSON: GB/s for each and the SM clock implied by clock64().
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 512, kFloat4 = 8192;  // 128 KB buffer per CTA, 1 CTA per SM... two CTAs of 96 KB do not fit: use 1

__global__ void __launch_bounds__(kThreads, 1) smem_read(float* out, int iters, long long* cycles) {
  extern __shared__ float4 buf[];
  for (int i = threadIdx.x; i < kFloat4; i += kThreads) buf[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kFloat4 / kThreads; ++k) {
      float4 v;  // volatile PTX load: the compiler may neither hoist nor drop it
      const unsigned addr = (unsigned)__cvta_generic_to_shared(&buf[k * kThreads + threadIdx.x]);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

__global__ void __launch_bounds__(kThreads, 1) smem_copy(float* out, int iters, long long* cycles) {
  extern __shared__ float4 buf[];
  for (int i = threadIdx.x; i < kFloat4; i += kThreads) buf[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kFloat4 / kThreads / 2; ++k) {  // read the first half-slot, write the mirrored slot
      float4 v;
      const unsigned src = (unsigned)__cvta_generic_to_shared(&buf[(2 * k) * kThreads + threadIdx.x]);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&buf[(2 * k + 1) * kThreads + threadIdx.x]);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  __syncthreads();
  if (buf[threadIdx.x].x == 123.456f) out[0] = buf[threadIdx.x].y;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = kFloat4 * sizeof(float4);
  cudaFuncSetAttribute(smem_read, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(smem_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  double gbs[2], mhz[2];
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) smem_read<<<sms, kThreads, smem>>>(out, iters, cyc);
      else smem_copy<<<sms, kThreads, smem>>>(out, iters, cyc);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)sms * iters * (which == 0 ? kFloat4 : kFloat4) * 16.0;  // copy: half read + half written
      gbs[which] = bytes / (ms * 1e-3) / 1e9;
      mhz[which] = c / (ms * 1e-3) / 1e6;
    }
  printf("{\"sms\": %d, \"smem_read_gbs\": %.1f, \"smem_copy_gbs\": %.1f, \"sm_mhz_read\": %.0f, \"sm_mhz_copy\": %.0f, "
         "\"bytes_per_clk_per_sm_read\": %.1f, \"bytes_per_clk_per_sm_copy\": %.1f}\n",
         sms, gbs[0], gbs[1], mhz[0], mhz[1], gbs[0] * 1e9 / (sms * mhz[0] * 1e6), gbs[1] * 1e9 / (sms * mhz[1] * 1e6));
  return 0;
}
