// Compares the packed (fp32x2) forms used by the BP kernel with their scalar forms on BP-like operands.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../qec_ldpc_b200/csrc/bp_kernel.cuh"
using namespace qldpc;

__device__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ float prob(uint32_t& s) {  // a message-like value in [0,1]
  uint32_t k = rng(s) % 8;
  if (k == 0) return 0.0f;
  if (k == 1) return 1.0f;
  if (k == 2) return __uint_as_float(((90u + rng(s) % 37u) << 23) | (rng(s) & 0x7FFFFFu));  // tiny
  if (k == 3) return 1.0f - __uint_as_float(((100u + rng(s) % 27u) << 23) | (rng(s) & 0x7FFFFFu));
  return (rng(s) >> 8) * (1.0f / 16777216.0f);
}
__global__ void check(unsigned long long* out) {
  uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long bad_div = 0, bad_om = 0, bad_chain = 0;
  for (int i = 0; i < 20000; ++i) {
    float x0 = prob(s) * prob(s), x1 = prob(s) * prob(s), y0 = x0 + prob(s), y1 = x1 + prob(s);
    bool u = false;
    float q0 = div_fast<0>(x0, y0, u), q1 = div_fast<0>(x1, y1, u);
    Pack<2> q = div_fast_pack<0, 2>(Pack<2>{make_float2(x0, x1)}, Pack<2>{make_float2(-y0, -y1)}, u);
    auto same = [](float a, float b) { return (a != a && b != b) || __float_as_uint(a) == __float_as_uint(b); };
    bad_div += !same(q.get(0), q0) + !same(q.get(1), q1);
    Pack<2> om = pfma(Pack<2>{make_float2(x0, x1)}, Pack<2>::splat(-1.0f), Pack<2>::splat(1.0f));
    bad_om += !same(om.get(0), __fsub_rn(1.0f, x0)) + !same(om.get(1), __fsub_rn(1.0f, x1));
    float a0 = prob(s), a1 = prob(s), b0 = prob(s), b1 = prob(s);
    Pack<2> t = pfma(Pack<2>::splat(-2.0f), Pack<2>{make_float2(a0, a1)}, Pack<2>::splat(1.0f));
    Pack<2> t2 = pfma(Pack<2>::splat(-2.0f), Pack<2>{make_float2(b0, b1)}, Pack<2>::splat(1.0f));
    Pack<2> p = pmul(t, t2);
    Pack<2> r = pfma(Pack<2>{make_float2(-0.5f, 0.5f)}, p, Pack<2>::splat(0.5f));
    float s0 = __fmul_rn(__fmaf_rn(-2.0f, a0, 1.0f), __fmaf_rn(-2.0f, b0, 1.0f));
    float s1 = __fmul_rn(__fmaf_rn(-2.0f, a1, 1.0f), __fmaf_rn(-2.0f, b1, 1.0f));
    bad_chain += !same(r.get(0), __fmaf_rn(-0.5f, s0, 0.5f)) + !same(r.get(1), __fmaf_rn(0.5f, s1, 0.5f));
  }
  atomicAdd(&out[0], bad_div); atomicAdd(&out[1], bad_om); atomicAdd(&out[2], bad_chain);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
  check<<<148, 256>>>(d);
  unsigned long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("mismatches: div %llu  one_minus %llu  check_chain %llu\n", h[0], h[1], h[2]);
  return 0;
}
