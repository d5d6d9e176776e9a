// Are the halves of FMUL2 / FFMA2 / FADD2 bit-identical to scalar round-to-nearest ops, including denormals, NaN, 0?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ float pick(uint32_t& s) {
  uint32_t r = rng(s), k = rng(s) % 10;
  if (k == 0) return 0.0f;
  if (k == 1) return __uint_as_float(r & 0x007FFFFFu);               // denormal
  if (k == 2) return __uint_as_float((r & 0x007FFFFFu) | 0x00800000u);  // tiny normal
  if (k == 3) return __uint_as_float(0x7FC00000u);                   // NaN
  if (k == 4) return 1.0f;
  return __uint_as_float(((60u + r % 70u) << 23) | (rng(s) & 0x007FFFFFu));  // 2^-67 .. 2^2
}
__global__ void check(unsigned long long* out) {
  uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long bad_mul = 0, bad_fma = 0, bad_add = 0, den = 0;
  for (int i = 0; i < 20000; ++i) {
    float a0 = pick(s), a1 = pick(s), b0 = pick(s), b1 = pick(s), c0 = pick(s), c1 = pick(s);
    float2 m = __fmul2_rn(make_float2(a0, a1), make_float2(b0, b1));
    float2 f = __ffma2_rn(make_float2(a0, a1), make_float2(b0, b1), make_float2(c0, c1));
    float2 d = __fadd2_rn(make_float2(a0, a1), make_float2(c0, c1));
    float m0 = __fmul_rn(a0, b0), m1 = __fmul_rn(a1, b1), f0 = __fmaf_rn(a0, b0, c0), f1 = __fmaf_rn(a1, b1, c1);
    float d0 = __fadd_rn(a0, c0), d1 = __fadd_rn(a1, c1);
    auto same = [](float x, float y) { return (x != x && y != y) || __float_as_uint(x) == __float_as_uint(y); };
    bad_mul += !same(m.x, m0) + !same(m.y, m1);
    bad_fma += !same(f.x, f0) + !same(f.y, f1);
    bad_add += !same(d.x, d0) + !same(d.y, d1);
    uint32_t e = __float_as_uint(m0) & 0x7F800000u;
    den += (e == 0 && m0 != 0.0f);
  }
  atomicAdd(&out[0], bad_mul); atomicAdd(&out[1], bad_fma); atomicAdd(&out[2], bad_add); atomicAdd(&out[3], den);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
  check<<<148, 256>>>(d);
  unsigned long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("mismatches: mul2 %llu  fma2 %llu  add2 %llu   (denormal scalar products seen: %llu)\n", h[0], h[1], h[2], h[3]);
  return 0;
}
