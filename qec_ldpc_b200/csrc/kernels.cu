// Device code: BP tile-kernel instantiations and dispatch, Philox depolarizing-error generator, sparse
// syndrome kernel, bit pack/unpack, and the on-device statistics reduction (CodeStatistics counters).
#include "kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../../include/qldpc_b200.h"

namespace qldpc {

// =====================================================================================================
// BP dispatch
// =====================================================================================================

typedef void (*BpKernel)(const BpArgs);
constexpr int kMaxT = 512;

// Multiprocessor count of the current device (grids of the grid-stride kernels are sized in multiples of it).
static int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    int v = 0;
    cached = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0 ? v : 148;
    cached_dev = dev;
  }
  return cached;
}
// one translation unit per shape (bp_shape_<dc>_<dv>.cu)
#define QLDPC_SHAPES(X) X(6, 3) X(10, 4) X(10, 5) X(8, 4) X(8, 3) X(12, 6) X(10, 3) X(12, 3) X(12, 4) X(12, 5) X(4, 2) X(6, 2) X(8, 2) X(10, 2) X(12, 2)
#define QLDPC_DECL(DC, DV) BpKernel bp_shape_##DC##_##DV(int vec, int guard, int m, int threads);
QLDPC_SHAPES(QLDPC_DECL)
#undef QLDPC_DECL

// QLDPC_NO_SPECIALIZE=1 in the environment forces the generic kernels (check count as a run-time value) even where an
// instantiation with a compile-time check count exists; the parity tests run both.
static bool specialization_enabled() {
  const char* e = getenv("QLDPC_NO_SPECIALIZE");
  return !(e && e[0] && e[0] != '0');
}

static BpKernel lookup_kernel(int dc, int dv, int vec, int guard, int m, int threads = 0) {
  if (!specialization_enabled()) m = 0;
#define QLDPC_CASE(DC, DV) \
  if (dc == DC && dv == DV) return bp_shape_##DC##_##DV(vec, guard, m, threads);
  QLDPC_SHAPES(QLDPC_CASE)
#undef QLDPC_CASE
  return nullptr;
}

bool bp_configure(int dc, int dv, int m_code, int n, int qc_P, int num_sms, BpLaunch& cfg, const char** why) {
  static const char* kNoShape = "no compiled BP kernel for this (check degree, variable degree)";
  static const char* kNoFit = "one frame of messages does not fit in shared memory (such codes use the HBM-resident path)";
  static const char* kBadCfg = "invalid launch configuration";
  // the specialised instantiations lay their variable phase out by circulant column blocks (bp_kernel.cuh:var_phase)
  const int spec_m = qc_P > 0 && qc_P * dv == m_code && n % qc_P == 0 ? m_code : 0;
  cfg.spec_m = spec_m;
  if (!lookup_kernel(dc, dv, 1, 0, spec_m)) { *why = kNoShape; return false; }
  const int m = m_code, E = m * dc, mw = (m + 31) / 32, nw = (n + 31) / 32;
  if (E >= 65536) { *why = kNoFit; return false; }
  int dev = 0, smem_optin = 0, smem_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  // Candidate (tile width, warps per CTA) pairs are scored by the warps they keep resident per SM (shared memory and
  // the register file both limit the CTA count; beyond ~16 warps the kernel gains nothing) and by
  // the warp-rounds wasted when the check phase (ceil(m/32) warp-tasks) and the variable phase (ceil(n/32)) do not
  // divide evenly among the CTA's warps.  Wider tiles win ties (fewer shared-memory instructions per edge-update).
  auto regs_of = [&](int v) {
    int r = 0;
    for (int guard : {0, 1, 3, 4, 5, 7})
      for (int thr : {0, 128}) {  // generic instantiation and, where there is one, the one specialised for this code
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, lookup_kernel(dc, dv, v, guard, spec_m, thr)) == cudaSuccess) r = std::max(r, fa.numRegs);
      }
    cudaGetLastError();
    return std::max(r, 32);
  };
  int vec = cfg.vec, threads = cfg.threads;
  if (vec != 0 && !(vec == 1 || vec == 2 || vec == 4)) { *why = kBadCfg; return false; }
  if (vec == 0 || threads == 0) {
    const int wc = (m + 31) / 32, wv = (n + 31) / 32;
    double best = -1.0;
    int best_v = 0, best_w = 0;
    for (int v : {4, 2, 1}) {
      if (vec != 0 && v != vec) continue;
      const int sm_bytes = (int)bp_smem_bytes(v, E, m, n, mw, nw);
      if (sm_bytes > smem_optin) continue;
      const int smem_ctas = std::max(1, smem_sm / (sm_bytes + 1024));
      const int regs_v = (regs_of(v) + 7) / 8 * 8;
      for (int w = 1; w <= kMaxT / 32; ++w) {
        if (threads != 0 && 32 * w != threads) continue;
        const int reg_ctas = 65536 / (regs_v * 32 * w);
        if (reg_ctas < 1) continue;
        const int resident = std::min(std::min(smem_ctas, reg_ctas), 32) * w;
        const double waste = (double)((wc + w - 1) / w * w - wc) * dc + (double)((wv + w - 1) / w * w - wv) * dv * 2;
        const double work = (double)wc * dc + (double)wv * dv * 2;
        // A CTA whose warp count is not a multiple of 4 loads one SM sub-partition with two of its warps; since the
        // packed-FP32 kernel keeps the FMA pipe ~2/3 busy, that sub-partition then paces every phase of the CTA
        // (measured on the Z side of J4K5L10P61: 4 warps 9.85e11 vs 5 warps 9.33e11 edge-updates/s).
        // Single-warp CTAs (tiny codes: one warp covers a whole phase) are spread over the sub-partitions by the CTA
        // scheduler instead (measured on J3K3L6P7: 1 warp 6.9e11 vs 2 warps 4.6e11 edge-updates/s).
        const double balance = w % 4 == 0 || w == 1 ? 1.0 : 0.92;
        // beyond 16 warps extra residency still buys a few per cent (measured: 28 warps of the 2-slot tile beat 20 warps
        // of the 4-slot tile by 2-4%), which outweighs the wider tile's saving in shared-memory instructions
        // (tiles of 2 or 4 slots use the packed fp32x2 arithmetic, the 1-slot tile cannot)
        const double score = std::min(resident, 16) / 16.0 * (work / (work + waste)) * balance + 0.001 * v +
                             (v >= 2 ? 0.02 : 0.0) + 0.001 * std::min(resident, 32);
        if (score > best) { best = score; best_v = v; best_w = w; }
      }
    }
    if (best_v == 0) { *why = threads != 0 || vec != 0 ? kBadCfg : kNoFit; return false; }
    vec = best_v;
    threads = 32 * best_w;
  }
  const int smem = (int)bp_smem_bytes(vec, E, m, n, mw, nw);
  if (smem > smem_optin) { *why = kBadCfg; return false; }
  // The attribute is a per-kernel ceiling shared by every decoder of the process (several codes can use the same
  // instantiation with different tile sizes), so it is raised to the device limit rather than to this tile's size.
  for (int guard : {0, 1, 3, 4, 5, 7})
    for (int thr : {0, 128}) {
      if (cudaFuncSetAttribute(lookup_kernel(dc, dv, vec, guard, spec_m, thr), cudaFuncAttributeMaxDynamicSharedMemorySize,
                               smem_optin) != cudaSuccess) {
        cudaGetLastError();
        *why = kNoFit;
        return false;
      }
    }
  const int regs = regs_of(vec);
  BpKernel k = lookup_kernel(dc, dv, vec, 3, spec_m, threads);
  if (threads % 32 || threads < 32 || threads > kMaxT) { *why = kBadCfg; return false; }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    *why = kNoFit;
    return false;
  }
  if (cfg.ctas_per_sm > 0) occ = std::min(occ, cfg.ctas_per_sm);
  cfg.vec = vec;
  cfg.threads = threads;
  cfg.ctas_per_sm = occ;
  cfg.grid = occ * num_sms;
  cfg.smem = smem;
  cfg.regs = regs;
  cfg.max_threads = kMaxT;
  return true;
}

cudaError_t bp_launch(int dc, int dv, const BpLaunch& cfg, const BpArgs& args, int nframes, int guard, cudaStream_t st) {
  if (args.trace_q || args.trace_r) guard += 4;  // the same instantiation with the message taps compiled in
  BpKernel k = lookup_kernel(dc, dv, cfg.vec, guard, cfg.spec_m, cfg.threads);
  if (!k) return cudaErrorInvalidDeviceFunction;
  const int tiles = (nframes + cfg.vec - 1) / cfg.vec;
  const int grid = std::max(1, std::min(cfg.grid, tiles));
  k<<<grid, cfg.threads, cfg.smem, st>>>(args);
  return cudaGetLastError();
}

// =====================================================================================================
// Error generation + syndrome in one kernel: device-side counter-based Philox depolarizing noise (replaces the
// mt19937 weight-W generator of DecoderCPU.h:394-396,446-459 / RandomErrorGenerator.h:31-44 for the Monte-Carlo path)
// and s = H e (mod 2) of both sides (Quantum_LDPC_Code::GetSyndromeX/Z, Quantum_LDPC_Code.h:94-124) -- the reference's
// intended "generate syndrome" step of its statistics kernel (kernels.cu:252-268).  One warp per frame; lane b handles
// Philox block b (qubits 4b..4b+3): the erroneous qubits flip their dv checks in the warp's shared-memory syndrome
// words straight from the registers that hold the fresh error bits, so the errors are written to HBM once (the
// statistics kernel needs them) and never read back; eight lanes assemble one 32-bit error word.
// =====================================================================================================
// Syndrome of one side from the frame's error words in shared memory: the positions of the set bits are compacted into
// a list (warp prefix sum of the word popcounts), then the (error, k-th check of its variable) pairs are dealt out
// evenly over the 32 lanes, so the scatter runs without lane divergence however the errors cluster.
__device__ __forceinline__ void scatter_side(const uint32_t* errw, int nw, uint16_t* list, const uint16_t* __restrict__ vchk,
                                             int n, int dv, uint32_t* syn, int lane) {
  int count = 0;
  for (int w0 = 0; w0 < nw; w0 += 32) {
    const int w = w0 + lane;
    uint32_t word = w < nw ? errw[w] : 0u;
    const int c = __popc(word);
    int incl = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += t;
    }
    int pos = count + incl - c;
    while (word) {
      list[pos++] = (uint16_t)(w * 32 + __ffs((int)word) - 1);
      word &= word - 1u;
    }
    count += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  const int items = count * dv;
  for (int i = lane; i < items; i += 32) {
    const int j = i / dv, k = i - j * dv;
    const int e = vchk[k * n + list[j]];
    atomicXor(&syn[e >> 5], 1u << (e & 31));
  }
}

__global__ void __launch_bounds__(256) generate_syndrome_kernel(uint64_t seed, uint64_t first_frame, int nframes, int n,
                                                                int nw, Thresholds thr, uint32_t* __restrict__ errX,
                                                                uint32_t* __restrict__ errZ,
                                                                const uint16_t* __restrict__ vchkX, int dvX, int mwX,
                                                                uint32_t* __restrict__ synX,
                                                                const uint16_t* __restrict__ vchkZ, int dvZ, int mwZ,
                                                                uint32_t* __restrict__ synZ) {
  extern __shared__ uint32_t sh[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nblocks = (n + 3) >> 2;
  // per warp: error words x, z [nw each], syndrome words [mwX + mwZ], position list [n] (16-bit)
  const int per_warp = 2 * nw + mwX + mwZ + (n + 1) / 2;
  uint32_t* ex = sh + (size_t)wib * per_warp;
  uint32_t* ez = ex + nw;
  uint32_t* smX = ez + nw;
  uint32_t* smZ = smX + mwX;
  uint16_t* list = reinterpret_cast<uint16_t*>(smZ + mwZ);
  for (int f = warp; f < nframes; f += nwarps) {
    const uint64_t frame = first_frame + (uint64_t)f;
    for (int w = lane; w < mwX + mwZ; w += 32) smX[w] = 0u;
    for (int b0 = 0; b0 < nblocks; b0 += 32) {
      const int b = b0 + lane;
      uint32_t xn = 0, zn = 0;
      if (b < nblocks) {
        depolarizing_block(seed, frame, (uint32_t)b, thr, xn, zn);
        const int valid = n - 4 * b;  // qubits beyond n do not exist
        if (valid < 4) {
          const uint32_t mask = (1u << valid) - 1u;
          xn &= mask;
          zn &= mask;
        }
      }
      uint32_t xv = xn << ((lane & 7) * 4), zv = zn << ((lane & 7) * 4);
#pragma unroll
      for (int s = 1; s < 8; s <<= 1) {
        xv |= __shfl_xor_sync(0xffffffffu, xv, s);
        zv |= __shfl_xor_sync(0xffffffffu, zv, s);
      }
      const int w = b >> 3;
      if ((lane & 7) == 0 && w < nw) {
        errX[(size_t)f * nw + w] = xv;
        errZ[(size_t)f * nw + w] = zv;
        ex[w] = xv;
        ez[w] = zv;
      }
    }
    __syncwarp();
    if (synX) {
      scatter_side(ex, nw, list, vchkX, n, dvX, smX, lane);
      __syncwarp();
      scatter_side(ez, nw, list, vchkZ, n, dvZ, smZ, lane);
      __syncwarp();
      for (int w = lane; w < mwX; w += 32) synX[(size_t)f * mwX + w] = smX[w];
      for (int w = lane; w < mwZ; w += 32) synZ[(size_t)f * mwZ + w] = smZ[w];
    }
    __syncwarp();
  }
}

cudaError_t launch_generate_syndrome(uint64_t seed, uint64_t first_frame, int nframes, int n, int nw, Thresholds thr,
                                     uint32_t* errX, uint32_t* errZ, const uint16_t* vchkX, int dvX, int mwX,
                                     uint32_t* synX, const uint16_t* vchkZ, int dvZ, int mwZ, uint32_t* synZ,
                                     cudaStream_t st) {
  if (nframes <= 0) return cudaSuccess;
  const size_t per_warp = (size_t)(2 * nw + mwX + mwZ + (n + 1) / 2) * sizeof(uint32_t);
  if (per_warp > 48 * 1024) return cudaErrorInvalidValue;
  const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (48 * 1024) / per_warp));
  const int blocks = std::min((nframes + warps - 1) / warps, sm_count() * 16);
  generate_syndrome_kernel<<<blocks, 32 * warps, warps * per_warp, st>>>(seed, first_frame, nframes, n, nw, thr, errX, errZ,
                                                                       vchkX, dvX, mwX, synX, vchkZ, dvZ, mwZ, synZ);
  return cudaGetLastError();
}

// =====================================================================================================
// Syndrome s = H e mod 2 (Quantum_LDPC_Code::GetSyndromeX/Z, Quantum_LDPC_Code.h:94-124).  The reference scans the
// dense row of every check; here the work is proportional to the error weight: one warp per frame walks the set
// bits of the bit-packed error and flips the dv checks of each erroneous qubit in shared memory (vchk = CSC table).
// =====================================================================================================
__device__ __forceinline__ void syndrome_side(const uint32_t* __restrict__ err, int nw, const uint16_t* __restrict__ vchk,
                                              int n, int dv, int mw, uint32_t* sm, uint32_t* __restrict__ out, int lane) {
  for (int w = lane; w < mw; w += 32) sm[w] = 0u;
  __syncwarp();
  for (int w = lane; w < nw; w += 32) {
    uint32_t word = err[w];
    while (word) {
      const int v = w * 32 + __ffs((int)word) - 1;
      word &= word - 1u;
      for (int k = 0; k < dv; ++k) {
        const int e = vchk[k * n + v];
        atomicXor(&sm[e >> 5], 1u << (e & 31));
      }
    }
  }
  __syncwarp();
  for (int w = lane; w < mw; w += 32) out[w] = sm[w];
  __syncwarp();
}

__global__ void __launch_bounds__(256) syndrome_kernel(const uint32_t* __restrict__ errX, const uint32_t* __restrict__ errZ,
                                                       int nframes, int n, int nw, const uint16_t* __restrict__ vchkX,
                                                       int dvX, int mwX, uint32_t* __restrict__ synX,
                                                       const uint16_t* __restrict__ vchkZ, int dvZ, int mwZ,
                                                       uint32_t* __restrict__ synZ) {
  extern __shared__ uint32_t sh[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  uint32_t* sm = sh + (size_t)wib * max(mwX, mwZ);
  for (int f = warp; f < nframes; f += nwarps) {
    syndrome_side(errX + (size_t)f * nw, nw, vchkX, n, dvX, mwX, sm, synX + (size_t)f * mwX, lane);
    syndrome_side(errZ + (size_t)f * nw, nw, vchkZ, n, dvZ, mwZ, sm, synZ + (size_t)f * mwZ, lane);
  }
}

cudaError_t launch_syndrome(const uint32_t* errX, const uint32_t* errZ, int nframes, int n, int nw,
                            const uint16_t* vchkX, int dvX, int mwX, uint32_t* synX, const uint16_t* vchkZ, int dvZ,
                            int mwZ, uint32_t* synZ, cudaStream_t st) {
  if (nframes <= 0) return cudaSuccess;
  // one warp per frame with max(mwX, mwZ) words of shared memory each; fewer warps per block for very wide codes
  const size_t per_warp = (size_t)std::max(mwX, mwZ) * sizeof(uint32_t);
  const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (48 * 1024) / per_warp));
  const int blocks = std::min((nframes + warps - 1) / warps, sm_count() * 16);
  syndrome_kernel<<<blocks, 32 * warps, warps * per_warp, st>>>(errX, errZ, nframes, n, nw, vchkX, dvX, mwX, synX, vchkZ,
                                                               dvZ, mwZ, synZ);
  return cudaGetLastError();
}

// =====================================================================================================
// Bit pack / unpack between one-element-per-bit rows (the reference's int / byte vectors) and packed words
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) pack_kernel(const T* __restrict__ src, int64_t rows, int cols, int words,
                                                   uint32_t* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t total = rows * words;
  for (int64_t t = warp; t < total; t += nwarps) {
    const int64_t r = t / words;
    const int w = (int)(t - r * words);
    const int c = w * 32 + lane;
    const unsigned bit = c < cols ? (unsigned)(src[r * cols + c] != 0) : 0u;
    const unsigned word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) dst[t] = word;
  }
}

__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t* __restrict__ src, int64_t rows, int cols, int words,
                                                     uint8_t* __restrict__ dst) {
  const int64_t total = rows * cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / cols;
    const int c = (int)(t - r * cols);
    dst[t] = (uint8_t)((src[r * words + (c >> 5)] >> (c & 31)) & 1u);
  }
}

cudaError_t launch_pack(const void* src, int elem_size, int64_t rows, int cols, int words, uint32_t* dst, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int blocks = (int)std::min<int64_t>((rows * words + 7) / 8, sm_count() * 32);
  if (elem_size == 1) pack_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)src, rows, cols, words, dst);
  else pack_kernel<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)src, rows, cols, words, dst);
  return cudaGetLastError();
}

cudaError_t launch_unpack(const uint32_t* src, int64_t rows, int cols, int words, uint8_t* dst, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int blocks = (int)std::min<int64_t>((rows * cols + 255) / 256, sm_count() * 32);
  unpack_kernel<<<blocks, 256, 0, st>>>(src, rows, cols, words, dst);
  return cudaGetLastError();
}

// Small-batch helpers of the low-latency Decode path: both syndrome rows packed by one launch, and everything a
// frame returns (corrections of both sides as bytes, ErrorCode bits, iteration counts) produced by one launch.
__global__ void __launch_bounds__(256) pack2_kernel(const uint8_t* __restrict__ srcX, int colsX, int wordsX,
                                                    uint32_t* __restrict__ dstX, const uint8_t* __restrict__ srcZ,
                                                    int colsZ, int wordsZ, uint32_t* __restrict__ dstZ, int rows) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int per = wordsX + wordsZ, total = rows * per;
  for (int t = warp; t < total; t += nwarps) {
    const int r = t / per, w = t - r * per;
    const bool z = w >= wordsX;
    const int ww = z ? w - wordsX : w, cols = z ? colsZ : colsX;
    const uint8_t* src = z ? srcZ : srcX;
    const int c = ww * 32 + lane;
    const unsigned bit = c < cols ? (unsigned)(src[(size_t)r * cols + c] != 0) : 0u;
    const unsigned word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) (z ? dstZ : dstX)[(size_t)r * (z ? wordsZ : wordsX) + ww] = word;
  }
}

cudaError_t launch_pack2(const uint8_t* srcX, int colsX, int wordsX, uint32_t* dstX, const uint8_t* srcZ, int colsZ,
                         int wordsZ, uint32_t* dstZ, int rows, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int blocks = std::min((rows * (wordsX + wordsZ) + 7) / 8, sm_count() * 8);
  pack2_kernel<<<blocks, 256, 0, st>>>(srcX, colsX, wordsX, dstX, srcZ, colsZ, wordsZ, dstZ, rows);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) finish_small_kernel(const uint32_t* __restrict__ decX,
                                                           const uint32_t* __restrict__ decZ, int rows, int cols, int words,
                                                           uint8_t* __restrict__ outX, uint8_t* __restrict__ outZ,
                                                           const uint8_t* __restrict__ sfX, const uint8_t* __restrict__ sfZ,
                                                           uint8_t* __restrict__ flags, const uint32_t* __restrict__ itX,
                                                           const uint32_t* __restrict__ itZ, uint32_t* __restrict__ iters) {
  const int total = rows * cols;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int r = t / cols, c = t - r * cols;
    outX[t] = (uint8_t)((decX[(size_t)r * words + (c >> 5)] >> (c & 31)) & 1u);
    outZ[t] = (uint8_t)((decZ[(size_t)r * words + (c >> 5)] >> (c & 31)) & 1u);
    if (c == 0) {
      const unsigned sx = sfX[r], sz = sfZ[r];
      flags[r] = (uint8_t)((sx & 1u) | ((sz & 1u) << 1) | (((sx >> 1) & 1u) << 2) | (((sz >> 1) & 1u) << 3));
      iters[2 * r] = itX[r];
      iters[2 * r + 1] = itZ[r];
    }
  }
}

cudaError_t launch_finish_small(const uint32_t* decX, const uint32_t* decZ, int rows, int cols, int words, uint8_t* outX,
                                uint8_t* outZ, const uint8_t* sfX, const uint8_t* sfZ, uint8_t* flags, const uint32_t* itX,
                                const uint32_t* itZ, uint32_t* iters, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int blocks = std::min((rows * cols + 255) / 256, sm_count() * 8);
  finish_small_kernel<<<blocks, 256, 0, st>>>(decX, decZ, rows, cols, words, outX, outZ, sfX, sfZ, flags, itX, itZ, iters);
  return cudaGetLastError();
}

// =====================================================================================================
// Statistics: the per-frame bookkeeping of GetStatistics (DecoderCPU.h:461-521) and the CodeStatistics counters
// (CodeStatistics.h:5-20) as a warp-aggregated on-device reduction.  One warp per frame.
// Logical check = CheckLogicalError (Quantum_LDPC_Code.h:126-142) on the residual [x^xhat | z^zhat] with
// bit-packed, transposed rows (lane = row, coalesced), skipped when the residual is zero.
// =====================================================================================================
__device__ __forceinline__ bool any_odd_row(const uint32_t* __restrict__ LT, int rows, const uint32_t* r, int words,
                                            int lane) {
  const int rows_pad = (rows + 31) & ~31;
  for (int g = 0; g < rows; g += 32) {
    uint32_t acc = 0;
    for (int w = 0; w < words; ++w) acc ^= LT[(size_t)w * rows_pad + g + lane] & r[w];
    if (__any_sync(0xffffffffu, __popc(acc) & 1)) return true;
  }
  return false;
}

__global__ void __launch_bounds__(256) stats_kernel(const StatsArgs a) {
  extern __shared__ uint32_t sh[];
  __shared__ unsigned long long blk[QLDPC_NUM_COUNTERS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nw = a.nw;
  uint32_t* res = sh + (size_t)wib * 2 * nw;  // residual words: x part, z part
  if (threadIdx.x < QLDPC_NUM_COUNTERS) blk[threadIdx.x] = 0ull;
  __syncthreads();
  unsigned long long k[QLDPC_NUM_COUNTERS];
#pragma unroll
  for (int i = 0; i < QLDPC_NUM_COUNTERS; ++i) k[i] = 0ull;
  for (int f = warp; f < a.nframes; f += nwarps) {
    unsigned ex = 0, ez = 0, rx = 0, rz = 0;
    for (int w = lane; w < nw; w += 32) {
      const uint32_t x = a.errX[(size_t)f * nw + w], z = a.errZ[(size_t)f * nw + w];
      const uint32_t dx = x ^ a.decX[(size_t)f * nw + w], dz = z ^ a.decZ[(size_t)f * nw + w];
      res[w] = dx;
      res[nw + w] = dz;
      ex |= x; ez |= z; rx |= dx; rz |= dz;
    }
    __syncwarp();
    const bool anyx = __any_sync(0xffffffffu, ex != 0), anyz = __any_sync(0xffffffffu, ez != 0);
    const bool resx = __any_sync(0xffffffffu, rx != 0), resz = __any_sync(0xffffffffu, rz != 0);
    const unsigned sx = a.sfX[f], sz = a.sfZ[f];
    const bool synx = sx & 1u, synz = sz & 1u;
    unsigned fl = (sx & 1u) | ((sz & 1u) << 1) | (((sx >> 1) & 1u) << 2) | (((sz >> 1) & 1u) << 3);
    const bool nan = ((sx | sz) >> 2) & 1u;
    bool logical = false, corrected = false;
    if (!(synx || synz)) {  // DecoderCPU.h:492
      if (resx && a.lx_rows) logical = any_odd_row(a.lx, a.lx_rows, res, nw, lane);
      if (!logical && resz && a.lz_rows) logical = any_odd_row(a.lz, a.lz_rows, res + nw, nw, lane);
      if (!logical && (resx || resz) && a.lm_rows) logical = any_odd_row(a.lm, a.lm_rows, res, 2 * nw, lane);
      corrected = !logical;
    }
    fl |= (logical ? QLDPC_FRAME_LOGICAL : 0) | (corrected ? QLDPC_FRAME_CORRECTED : 0) | (nan ? QLDPC_FRAME_NAN : 0);
    if (lane == 0) {
      k[QLDPC_C_FRAMES] += 1;
      k[QLDPC_C_XTESTED] += anyx;  // DecoderCPU.h:464-473
      k[QLDPC_C_ZTESTED] += anyz;
      k[QLDPC_C_CORRECTED] += corrected;
      k[QLDPC_C_SYNX] += synx;
      k[QLDPC_C_SYNZ] += synz;
      k[QLDPC_C_LOGICAL] += logical;
      k[QLDPC_C_CVX] += (sx >> 1) & 1u;  // counted independently of the outcome, DecoderCPU.h:514-521
      k[QLDPC_C_CVZ] += (sz >> 1) & 1u;
      k[QLDPC_C_ITERSX] += a.itX[f];
      k[QLDPC_C_ITERSZ] += a.itZ[f];
      k[QLDPC_C_NANFRAMES] += nan;
      if (a.fflags) a.fflags[f] = (uint8_t)fl;
    }
    __syncwarp();
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < QLDPC_NUM_COUNTERS; ++i)
      if (k[i]) atomicAdd(&blk[i], k[i]);
  }
  __syncthreads();
  if (threadIdx.x < QLDPC_NUM_COUNTERS && blk[threadIdx.x]) atomicAdd(&a.counters[threadIdx.x], blk[threadIdx.x]);
}

cudaError_t launch_stats(const StatsArgs& a, cudaStream_t st) {
  if (a.nframes <= 0) return cudaSuccess;
  // one warp per frame with 2 * nw words of shared memory each (the residual); for very long codes fewer warps per
  // block, and beyond 48 KB per block the opt-in shared-memory ceiling
  const size_t per_warp = (size_t)2 * a.nw * sizeof(uint32_t);
  size_t cap = 48 * 1024;
  if (per_warp * 4 > cap) {
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cap = (size_t)std::max(optin - 1024, 48 * 1024);
    if (per_warp > cap) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
    if (e != cudaSuccess) return e;
  }
  const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, cap / per_warp));
  const int blocks = std::min((a.nframes + warps - 1) / warps, sm_count() * 8);
  stats_kernel<<<blocks, 32 * warps, warps * per_warp, st>>>(a);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) merge_flags_kernel(const uint8_t* __restrict__ sfX, const uint8_t* __restrict__ sfZ,
                                                          int nframes, uint8_t* __restrict__ out) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nframes; f += gridDim.x * blockDim.x) {
    const unsigned sx = sfX[f], sz = sfZ[f];
    out[f] = (uint8_t)((sx & 1u) | ((sz & 1u) << 1) | (((sx >> 1) & 1u) << 2) | (((sz >> 1) & 1u) << 3));
  }
}

cudaError_t launch_merge_flags(const uint8_t* sfX, const uint8_t* sfZ, int nframes, uint8_t* out, cudaStream_t st) {
  if (nframes <= 0) return cudaSuccess;
  merge_flags_kernel<<<std::min((nframes + 255) / 256, sm_count() * 8), 256, 0, st>>>(sfX, sfZ, nframes, out);
  return cudaGetLastError();
}


// =====================================================================================================
// Parity tap for div_fast (bp_kernel.cuh): random operand pairs 0 <= x <= y from Philox, wide exponent spread,
// compared bit for bit with __fdiv_rn wherever div_fast does not flag the pair as unsafe.
// =====================================================================================================
__global__ void __launch_bounds__(256) division_check_kernel(uint64_t seed, long long npairs, unsigned long long* out) {
  unsigned long long mism = 0, unsafe_n = 0, zero_n = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += stride) {
    uint32_t r[4];
    philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0x44495631u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    // y: random mantissa, exponent in [2^-104, 2^1]; x = y scaled down by a random factor in (0, 1], or 0, or y
    const uint32_t ey = 23u + r[2] % 106u;                       // biased exponent 23..128
    const float y = __uint_as_float((ey << 23) | (r[0] & 0x7FFFFFu));
    float x;
    const uint32_t sel = r[3] & 15u;
    if (sel == 0) x = 0.0f;
    else if (sel == 1) x = y;
    else {
      const uint32_t ex = 1u + (r[3] >> 4) % ey;                 // biased exponent 1..ey
      x = __uint_as_float((ex << 23) | (r[1] & 0x7FFFFFu));
      if (x > y) x = y;
    }
    bool unsafe = false;
    const float q = div_fast<3>(x, y, unsafe);
    const float want = __fdiv_rn(x, y);
    if (unsafe) ++unsafe_n;
    else if (__float_as_uint(q) != __float_as_uint(want)) ++mism;
    if (x == 0.0f) ++zero_n;
  }
  atomicAdd(&out[0], mism);
  atomicAdd(&out[1], unsafe_n);
  atomicAdd(&out[2], zero_n);
}

cudaError_t launch_division_check(uint64_t seed, long long npairs, unsigned long long* out, cudaStream_t st) {
  division_check_kernel<<<sm_count() * 8, 256, 0, st>>>(seed, npairs, out);
  return cudaGetLastError();
}

}  // namespace qldpc
