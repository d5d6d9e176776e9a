// Arithmetic-only ceiling of the BP tile kernel: the exact FMUL2 / FFMA2 / FADD / MUFU stream of the check-node and
// variable-node updates (bp_kernel.cuh, reference order of operations) on register-resident values, no shared or
// global memory in the loop.  Answers "what would the kernel reach if memory, addressing, barriers and bookkeeping
// were free" and measures the issue cost of the individual packed instructions.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/arith_ceiling tools/micro/arith_ceiling.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../qec_ldpc_b200/csrc/bp_kernel.cuh"

using namespace qldpc;
typedef Pack<2> P2;

template <int DC>
__device__ __forceinline__ void check_math(P2 (&x)[DC], P2 cf) {
  P2 t[DC];
#pragma unroll
  for (int i = 0; i < DC; ++i) t[i] = pfma(P2::splat(-2.0f), x[i], P2::splat(1.0f));
  const P2 half = P2::splat(0.5f);
  P2 pre = t[0];
  {
    P2 p = t[1];
#pragma unroll
    for (int k = 2; k < DC; ++k) p = pmul(p, t[k]);
    x[0] = pfma(cf, p, half);
  }
#pragma unroll
  for (int i = 1; i < DC; ++i) {
    P2 p = pre;
#pragma unroll
    for (int k = i + 1; k < DC; ++k) p = pmul(p, t[k]);
    x[i] = pfma(cf, p, half);
    if (i < DC - 1) pre = pmul(pre, t[i]);
  }
}

// DEN: 0 = scalar adds on negated operands (shipping kernel), 1 = one FFMA2 (num * -1 + (-Q)) with the Q chain
// carried negated
template <int DV, int DEN>
__device__ __forceinline__ void var_math(P2 (&b)[DV], float prior, float omp) {
  P2 pk[DV], om[DV], num[DV], den[DV];
#pragma unroll
  for (int k = 0; k < DV; ++k) {
    pk[k] = b[k];
    om[k] = pfma(pk[k], P2::splat(-1.0f), P2::splat(1.0f));
  }
  P2 preP = P2::splat(prior), preQ = P2::splat(DEN ? -omp : omp);
#pragma unroll
  for (int j = 0; j < DV; ++j) {
    P2 p = preP, q = preQ;
#pragma unroll
    for (int k = j + 1; k < DV; ++k) {
      q = pmul(q, om[k]);
      p = pmul(p, pk[k]);
    }
    num[j] = p;
    den[j] = q;
    if (j < DV - 1) {
      preQ = pmul(preQ, om[j]);
      preP = pmul(preP, pk[j]);
    }
  }
  bool unsafe = false;
#pragma unroll
  for (int j = 0; j < DV; ++j) {
    if (DEN == 0) {
#pragma unroll
      for (int w = 0; w < 2; ++w) den[j].set(w, __fadd_rn(-den[j].get(w), -num[j].get(w)));
    } else {
      den[j] = pfma(num[j], P2::splat(-1.0f), den[j]);
    }
    b[j] = div_fast_pack<0, 2>(num[j], den[j], unsafe);
  }
}

template <int DC>
__global__ void __maxnreg__(72) k_check(float* out, int iters) {
  P2 x[DC];
#pragma unroll
  for (int i = 0; i < DC; ++i) x[i] = P2{make_float2(0.03f + 1e-4f * threadIdx.x + 0.01f * i, 0.04f + 0.02f * i)};
  const P2 cf = P2{make_float2(-0.5f, threadIdx.x & 1 ? 0.5f : -0.5f)};
  for (int it = 0; it < iters; ++it) check_math<DC>(x, cf);
  float s = 0;
#pragma unroll
  for (int i = 0; i < DC; ++i) s += x[i].a.x + x[i].a.y;
  if (s == 123.456f) out[0] = s;
}

template <int DV, int DEN>
__global__ void __maxnreg__(72) k_var(float* out, int iters, float prior) {
  P2 b[DV];
#pragma unroll
  for (int i = 0; i < DV; ++i) b[i] = P2{make_float2(0.3f + 1e-4f * threadIdx.x + 0.01f * i, 0.4f + 0.02f * i)};
  const float omp = 1.0f - prior;
  for (int it = 0; it < iters; ++it) var_math<DV, DEN>(b, prior, omp);
  float s = 0;
#pragma unroll
  for (int i = 0; i < DV; ++i) s += b[i].a.x + b[i].a.y;
  if (s == 123.456f) out[0] = s;
}

// ---- single-instruction issue probes: 8 independent accumulators per thread ----
// MODE 0 FMUL2 a=a*b | 1 FFMA2 a=a*b+c (b,c loop-invariant pairs) | 2 FFMA2 a=b*c+a | 3 FFMA2 a_k = a_k * a_{k+1} + a_{k+2}
// (three varying pairs) | 4 scalar FFMA a=a*b+c | 5 scalar FMUL | 6 MUFU.RCP | 7 FADD2
template <int MODE>
__global__ void k_probe(float* out, int iters, float bx, float cx) {
  float2 a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = make_float2(1.0f + 1e-3f * threadIdx.x + k, 2.0f + k);
  const float2 b = make_float2(bx, bx * 0.999f), c = make_float2(cx, cx * 1.001f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) a[k] = __fmul2_rn(a[k], b);
      if (MODE == 1) a[k] = __ffma2_rn(a[k], b, c);
      if (MODE == 2) a[k] = __ffma2_rn(b, c, a[k]);
      if (MODE == 3) a[k] = __ffma2_rn(a[k], a[(k + 1) & 7], a[(k + 2) & 7]);
      if (MODE == 4) { a[k].x = __fmaf_rn(a[k].x, b.x, c.x); a[k].y = __fmaf_rn(a[k].y, b.y, c.y); }
      if (MODE == 5) { a[k].x = __fmul_rn(a[k].x, b.x); a[k].y = __fmul_rn(a[k].y, b.y); }
      if (MODE == 6) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[k].x)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[k].y)); }
      if (MODE == 7) a[k] = __fadd2_rn(a[k], b);
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
  if (s == 123.456f) out[0] = s;
}

static int g_sms = 148;
static double g_clk_ghz = 1.965;

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
  g_sms = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  g_clk_ghz = khz * 1e-6;
  float* out;
  cudaMalloc(&out, 4);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f,\n", prop.name, g_sms, g_clk_ghz);
  const int iters = 20000;
  // issue probes at 28 warps per SM (7 CTAs x 128 threads, the shipping occupancy) and at 16 warps per SM
  const char* names[8] = {"FMUL2 a*b", "FFMA2 a*b+c (b,c invariant)", "FFMA2 b*c+a", "FFMA2 three varying pairs",
                          "FFMA scalar x2", "FMUL scalar x2", "MUFU.RCP x2", "FADD2 a+b"};
  printf(" \"probes\": [\n");
  for (int ctas : {7, 4}) {
    double ms[8];
    ms[0] = time_ms([&] { k_probe<0><<<g_sms * ctas, 128>>>(out, iters, 0.9999f, 0.1f); });
    ms[1] = time_ms([&] { k_probe<1><<<g_sms * ctas, 128>>>(out, iters, 0.9999f, 0.1f); });
    ms[2] = time_ms([&] { k_probe<2><<<g_sms * ctas, 128>>>(out, iters, 0.9999f, 0.1f); });
    ms[3] = time_ms([&] { k_probe<3><<<g_sms * ctas, 128>>>(out, iters, 0.9999f, 0.1f); });
    ms[4] = time_ms([&] { k_probe<4><<<g_sms * ctas, 128>>>(out, iters, 0.9999f, 0.1f); });
    ms[5] = time_ms([&] { k_probe<5><<<g_sms * ctas, 128>>>(out, iters, 0.9999f, 0.1f); });
    ms[6] = time_ms([&] { k_probe<6><<<g_sms * ctas, 128>>>(out, iters, 0.9999f, 0.1f); });
    ms[7] = time_ms([&] { k_probe<7><<<g_sms * ctas, 128>>>(out, iters, 0.9999f, 0.1f); });
    for (int m = 0; m < 8; ++m) {
      // warp-instructions per SM sub-partition: ctas * 4 warps / 4 sub-partitions = ctas warps, each iters * 8 (x2 for
      // the scalar forms) instructions
      const double inst = (double)ctas * iters * 8 * ((m >= 4 && m <= 6) ? 2 : 1);
      const double cyc = ms[m] * 1e-3 * g_clk_ghz * 1e9;
      printf("  {\"warps_per_sm\": %d, \"op\": \"%s\", \"ms\": %.3f, \"cycles_per_warp_instruction_per_smsp\": %.3f}%s\n",
             ctas * 4, names[m], ms[m], cyc / inst, (ctas == 4 && m == 7) ? "" : ",");
    }
  }
  printf(" ],\n \"streams\": [\n");
  const float prior = 2.0f / 3.0f * 0.05f;
  struct R { const char* name; double ms; double edges; };
  for (int ctas : {7, 5, 4}) {
    const double thr = (double)g_sms * ctas * 128;
    R r[6];
    r[0] = {"check dc=10", time_ms([&] { k_check<10><<<g_sms * ctas, 128>>>(out, iters); }), thr * iters * 10 * 2};
    r[1] = {"var dv=4 (scalar den adds)", time_ms([&] { k_var<4, 0><<<g_sms * ctas, 128>>>(out, iters, prior); }), thr * iters * 4 * 2};
    r[2] = {"var dv=5 (scalar den adds)", time_ms([&] { k_var<5, 0><<<g_sms * ctas, 128>>>(out, iters, prior); }), thr * iters * 5 * 2};
    r[3] = {"var dv=4 (FFMA2 den)", time_ms([&] { k_var<4, 1><<<g_sms * ctas, 128>>>(out, iters, prior); }), thr * iters * 4 * 2};
    r[4] = {"var dv=5 (FFMA2 den)", time_ms([&] { k_var<5, 1><<<g_sms * ctas, 128>>>(out, iters, prior); }), thr * iters * 5 * 2};
    r[5] = {"check dc=8", time_ms([&] { k_check<8><<<g_sms * ctas, 128>>>(out, iters); }), thr * iters * 8 * 2};
    for (int i = 0; i < 6; ++i)
      printf("  {\"warps_per_sm\": %d, \"stream\": \"%s\", \"ms\": %.3f, \"edge_halves_per_s\": %.4e},\n", ctas * 4, r[i].name,
             r[i].ms, r[i].edges / (r[i].ms * 1e-3));
    // an edge-update = one check-side half + one variable-side half
    const double tc = r[0].ms * 1e-3 / r[0].edges;
    for (int v = 1; v <= 4; ++v) {
      const double tv = r[v].ms * 1e-3 / r[v].edges;
      printf("  {\"warps_per_sm\": %d, \"ceiling\": \"check dc=10 + %s\", \"edge_updates_per_s\": %.4e}%s\n", ctas * 4,
             r[v].name, 1.0 / (tc + tv), (ctas == 4 && v == 4) ? "" : ",");
    }
  }
  printf(" ]}\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
