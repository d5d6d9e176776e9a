// Does a packed fp32x2 instruction (FMUL2 / FFMA2, two FMA-pipe cycles per warp) leave the issue port free in its
// second cycle?  Streams of 8 independent FMUL2 per loop trip, alone and interleaved 1:1 with instructions of another
// pipe (integer ALU, shared-memory loads, MUFU), at the BP kernel's occupancy.  If mixed time ~ max(parts) the gaps are
// usable; if ~ sum(parts) the packed instruction holds the dispatch port for both cycles.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/issue_mix tools/micro/issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

// MODE bits: 64 = single IADD x8, 128 = single IADD x4, 1 = FMUL2 x8, 2 = ALU (LOP3/IADD) x8, 4 = LDS.64 x8, 8 = MUFU x4, 16 = scalar FMUL x16 (instead of FMUL2),
// 32 = FFMA2 with three varying pairs x8 (instead of FMUL2)
template <int MODE>
__global__ void k_mix(float* out, int iters, float bx, unsigned ix) {
  __shared__ float2 sm[1024];
  float2 a[8];
  unsigned u[8];
  float2 l[8];
  float m0 = 1.5f + threadIdx.x, m1 = 2.5f, m2 = 3.5f, m3 = 4.5f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = make_float2(1.0f + 1e-3f * threadIdx.x + k, 2.0f + k);
    u[k] = threadIdx.x * 7u + k;
    l[k] = make_float2(0.f, 0.f);
  }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float2(i, -i);
  __syncthreads();
  const float2 b = make_float2(bx, bx * 0.999f);
  const float2* p = sm + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE & 1) a[k] = __fmul2_rn(a[k], b);
      if (MODE & 16) { a[k].x = __fmul_rn(a[k].x, b.x); a[k].y = __fmul_rn(a[k].y, b.y); }
      if (MODE & 32) a[k] = __ffma2_rn(a[k], a[(k + 1) & 7], a[(k + 2) & 7]);
      if (MODE & 2) u[k] = (u[k] ^ ix) + (u[k] >> 3);   // LOP3 / SHF / IADD on the integer pipe
      if (MODE & 64) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(ix));
      if ((MODE & 128) && (k & 1)) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(ix));
      if (MODE & 4) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p + ((k * 128 + i) & 895)));
        l[k].x += 0.0f * v.x;  // keeps the load alive; folded by no one (volatile asm)
      }
      if ((MODE & 8) && (k & 1)) {
        float& m = k == 1 ? m0 : k == 3 ? m1 : k == 5 ? m2 : m3;
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(m));
      }
    }
  }
  float s = m0 + m1 + m2 + m3;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y + (float)u[k] + l[k].x;
  if (s == 123.456f) out[0] = s;
}

template <typename F>
double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  const int sms = prop.multiProcessorCount;
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 20000;
  printf("{\"device\": \"%s\", \"clock_ghz\": %.3f, \"unit\": \"cycles per loop trip per SM sub-partition warp (8 packed + 8 other instructions)\", \"rows\": [\n", prop.name, ghz);
  for (int ctas : {7, 4, 1}) {
    struct R { const char* name; double ms; };
#define RUN(M) time_ms([&] { k_mix<M><<<sms * ctas, 128>>>(out, iters, 0.9999f, 0x5bd1e995u); })
    R r[] = {{"FMUL2 x8", RUN(1)},
             {"ALU x8 (3 int ops each)", RUN(2)},
             {"FMUL2 x8 + ALU x8", RUN(1 | 2)},
             {"LDS.64 x8", RUN(4)},
             {"FMUL2 x8 + LDS.64 x8", RUN(1 | 4)},
             {"MUFU x4", RUN(8)},
             {"FMUL2 x8 + MUFU x4", RUN(1 | 8)},
             {"FMUL scalar x16", RUN(16)},
             {"FMUL scalar x16 + ALU x8", RUN(16 | 2)},
             {"FFMA2 3-pair x8", RUN(32)},
             {"FFMA2 3-pair x8 + ALU x8", RUN(32 | 2)},
             {"FMUL2 x8 + ALU x8 + LDS.64 x8 + MUFU x4", RUN(1 | 2 | 4 | 8)},
             {"IADD x8", RUN(64)},
             {"FMUL2 x8 + IADD x8", RUN(1 | 64)},
             {"IADD x4", RUN(128)},
             {"FMUL2 x8 + IADD x4", RUN(1 | 128)},
             {"FMUL scalar x16 + IADD x8", RUN(16 | 64)},
             {"FFMA2 3-pair x8 + IADD x8", RUN(32 | 64)},
             {"FMUL2 x8 + IADD x4 + MUFU x4(k odd)", RUN(1 | 128 | 8)}};
#undef RUN
    const int nr = sizeof(r) / sizeof(r[0]);
    for (int i = 0; i < nr; ++i)
      printf("  {\"warps_per_smsp\": %d, \"mix\": \"%s\", \"ms\": %.3f, \"cycles_per_trip\": %.2f}%s\n", ctas, r[i].name, r[i].ms,
             r[i].ms * 1e-3 * ghz * 1e9 / ((double)iters * ctas), (ctas == 1 && i == nr - 1) ? "" : ",");
  }
  printf("]}\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
