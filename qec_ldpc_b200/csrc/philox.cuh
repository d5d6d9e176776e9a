// Counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)
// and the integer-threshold depolarizing sampler built on it.  Host and device share this header so the
// generator is defined once; oracle/oracle.c restates the same definition independently for the parity tests.
//
//   key     = (seed lo, seed hi)
//   counter = (frame lo, frame hi, qubit >> 2, QLDPC_PHILOX_TAG)     one call serves four consecutive qubits
//   r       = out[qubit & 3];  T = floor(p * 2^32), t1 = T/3, t2 = 2T/3
//   r < t1 -> X,  t1 <= r < t2 -> Y (X and Z),  t2 <= r < T -> Z     (type->bit mapping: DecoderCPU.h:456-457)
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define QLDPC_HD __host__ __device__ __forceinline__
#else
#define QLDPC_HD inline
#endif

namespace qldpc {

constexpr uint32_t kPhiloxTag = 0x51454331u;  // "QEC1"

struct Thresholds {
  uint32_t t1, t2, T;
};

inline Thresholds depolarizing_thresholds(float p) {
  const double s = (double)p * 4294967296.0;
  const uint64_t T = s <= 0 ? 0ull : s >= 4294967295.0 ? 4294967295ull : (uint64_t)s;
  Thresholds t;
  t.t1 = (uint32_t)(T / 3);
  t.t2 = (uint32_t)(2 * T / 3);
  t.T = (uint32_t)T;
  return t;
}

QLDPC_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                            uint32_t (&out)[4]) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
#if defined(__CUDA_ARCH__)
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// Four qubits of one frame: bit w of the returned nibbles = x / z error on qubit 4*block + w.
QLDPC_HD void depolarizing_block(uint64_t seed, uint64_t frame, uint32_t block, const Thresholds& t, uint32_t& xn,
                                 uint32_t& zn) {
  uint32_t r[4];
  philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), block, kPhiloxTag, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  xn = zn = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    xn |= (uint32_t)(r[w] < t.t2) << w;                // X or Y
    zn |= (uint32_t)(r[w] >= t.t1 && r[w] < t.T) << w;  // Y or Z
  }
}

}  // namespace qldpc
