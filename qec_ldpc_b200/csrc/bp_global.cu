// Global-memory ("HBM-resident") belief-propagation path: the fallback for codes the shared-memory tile kernel
// (bp_kernel.cuh) does not cover -- a (check degree, variable degree) pair without a compiled instantiation, or a
// frame whose messages do not fit in shared memory.  Same arithmetic, operation for operation, as the tile kernel
// and the reference (EqNodeUpdate / VarNodeUpdate / CheckConvergence / Decode tail, DecoderCPU.h:150-390); only
// the residence of the message state differs.
//
// Layout: msg[row][F] with row = i*m + e (i = position of the edge in its check) and the FRAME index innermost, so a
// warp that processes one node for 32 consecutive frames reads and writes fully coalesced 128-byte lines.  A batch
// of F frames advances in lock step (all frames share the iteration number, so the n % 10 checkpoints coincide);
// frames that have stopped are frozen by a per-frame flag.  One check kernel + one variable kernel per iteration,
// a control kernel at checkpoints, then decision and syndrome-check kernels.  Bound by HBM: 16 bytes per
// edge-update; degrees are runtime values up to MAXD (8, 16 or 32).
#include <algorithm>
#include <cstdint>

#include "kernels.cuh"

namespace qldpc {

namespace {

__device__ __forceinline__ bool unconverged_g(float x) {
  constexpr uint32_t lo = 0x3C23D70Au, hi = 0x3F7D70A4u;  // 0.01f, 0.99f (DecoderCPU.h:260-261)
  return (__float_as_uint(x) - (lo + 1u)) < (hi - lo - 1u);
}

// syndrome bits -> bytes [e][F]; messages <- prior; per-frame state reset
__global__ void __launch_bounds__(256) g_init(const uint32_t* __restrict__ syn, int mw, int m, int E, int F, int nf,
                                              float prior, float* __restrict__ msg, uint8_t* __restrict__ synb,
                                              uint8_t* __restrict__ active, uint8_t* __restrict__ bad) {
  const long long total = (long long)E * F;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    msg[t] = prior;  // InitVarNodes, DecoderCPU.h:135-148
    if (t < (long long)m * F) {
      const int e = (int)(t / F), f = (int)(t % F);
      synb[t] = f < nf ? (uint8_t)((syn[(size_t)f * mw + (e >> 5)] >> (e & 31)) & 1u) : 0;
    }
    if (t < F) {
      active[t] = t < nf;
      bad[t] = 0;
    }
  }
}

template <int MAXD>
__global__ void __launch_bounds__(128) g_check(float* __restrict__ msg, const uint8_t* __restrict__ synb,
                                               const uint8_t* __restrict__ active, int m, int dc, int F) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (f >= F || !active[f]) return;
  float t[MAXD];
#pragma unroll
  for (int i = 0; i < MAXD; ++i)
    if (i < dc) t[i] = __fmaf_rn(-2.0f, msg[((size_t)i * m + e) * F + f], 1.0f);  // 1 - 2q (DecoderCPU.h:175)
  const float cf = synb[(size_t)e * F + f] ? 0.5f : -0.5f;  // DecoderCPU.h:178-183, see bp_kernel.cuh
  float pre = 1.0f;
#pragma unroll
  for (int i = 0; i < MAXD; ++i) {
    if (i < dc) {
      float p = pre;  // reference order: 1.0f * t0 * ... skipping i, left to right (DecoderCPU.h:168-176)
#pragma unroll
      for (int k = i + 1; k < MAXD; ++k)
        if (k < dc) p = __fmul_rn(p, t[k]);
      msg[((size_t)i * m + e) * F + f] = __fmaf_rn(cf, p, 0.5f);
      pre = __fmul_rn(pre, t[i]);
    }
  }
}

template <int MAXD>
__global__ void __launch_bounds__(128) g_var(float* __restrict__ msg, const uint32_t* __restrict__ vrow,
                                             const uint8_t* __restrict__ active, uint8_t* __restrict__ bad, int n, int dv,
                                             int F, float prior, int last, int ck) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (f >= F || !active[f]) return;
  float pk[MAXD], om[MAXD];
  uint32_t row[MAXD];
#pragma unroll
  for (int k = 0; k < MAXD; ++k)
    if (k < dv) {
      row[k] = vrow[(size_t)k * n + v];
      pk[k] = msg[(size_t)row[k] * F + f];
      om[k] = __fsub_rn(1.0f, pk[k]);
    }
  float preP = prior, preQ = __fsub_rn(1.0f, prior);  // DecoderCPU.h:209-210
  float fullP = preP, fullQ = preQ;
  if (last) {
#pragma unroll
    for (int k = 0; k < MAXD; ++k)
      if (k < dv) {
        fullQ = __fmul_rn(fullQ, om[k]);
        fullP = __fmul_rn(fullP, pk[k]);
      }
  }
  bool anybad = false;
#pragma unroll
  for (int j = 0; j < MAXD; ++j) {
    if (j < dv) {
      float P = preP, Q = preQ;
#pragma unroll
      for (int k = j + 1; k < MAXD; ++k)
        if (k < dv) {
          Q = __fmul_rn(Q, om[k]);
          P = __fmul_rn(P, pk[k]);
        }
      if (last) { P = fullP; Q = fullQ; }
      const float q = __fdiv_rn(P, __fadd_rn(Q, P));  // DecoderCPU.h:223
      msg[(size_t)row[j] * F + f] = q;
      anybad |= unconverged_g(q);
      preQ = __fmul_rn(preQ, om[j]);
      preP = __fmul_rn(preP, pk[j]);
    }
  }
  if (ck && anybad) bad[f] = 1;
}

// BeliefPropogation loop control (DecoderCPU.h:280-291) for the whole batch after iteration n.
__global__ void __launch_bounds__(256) g_control(uint8_t* __restrict__ active, uint8_t* __restrict__ bad,
                                                 uint8_t* __restrict__ convfail, uint32_t* __restrict__ iters, int F, int n,
                                                 int last, unsigned int* __restrict__ remaining) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  if (active[f]) {
    if (last || !bad[f]) {
      active[f] = 0;
      convfail[f] = bad[f];  // CONVERGENCE_FAIL = !CheckConvergence(final messages), DecoderCPU.h:375-378
      iters[f] = (uint32_t)(n + 1);
    } else {
      atomicAdd(remaining, 1u);
    }
  }
  bad[f] = 0;
}

// hard decision (any edge message >= 0.5f, DecoderCPU.h:354-373) + NaN flag
__global__ void __launch_bounds__(128) g_decide(const float* __restrict__ msg, const uint32_t* __restrict__ vrow, int n,
                                                int dv, int F, int nf, int nw, uint32_t* __restrict__ dec,
                                                uint8_t* __restrict__ nanflag) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (f >= nf) return;
  bool bit = false, nan = false;
  for (int k = 0; k < dv; ++k) {
    const float x = msg[(size_t)vrow[(size_t)k * n + v] * F + f];
    bit |= x >= 0.5f;
    nan |= x != x;
  }
  if (bit) atomicOr(&dec[(size_t)f * nw + (v >> 5)], 1u << (v & 31));
  if (nan) nanflag[f] = 1;
}

// syndrome of the decision against the input syndrome (DecoderCPU.h:380-384) and the per-frame outputs
__global__ void __launch_bounds__(128) g_verify(const uint32_t* __restrict__ dec, const uint32_t* __restrict__ cvar,
                                                const uint8_t* __restrict__ synb, int m, int dc, int F, int nf, int nw,
                                                uint8_t* __restrict__ mismatch) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (f >= nf) return;
  unsigned par = synb[(size_t)e * F + f];
  for (int i = 0; i < dc; ++i) {
    const uint32_t v = cvar[(size_t)i * m + e];
    par ^= (dec[(size_t)f * nw + (v >> 5)] >> (v & 31)) & 1u;
  }
  if (par) mismatch[f] = 1;
}

__global__ void __launch_bounds__(256) g_flags(const uint8_t* __restrict__ mismatch, const uint8_t* __restrict__ convfail,
                                               const uint8_t* __restrict__ nanflag, const uint32_t* __restrict__ it, int nf,
                                               uint8_t* __restrict__ flags, uint32_t* __restrict__ iters) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  flags[f] = (uint8_t)((mismatch[f] & 1u) | ((convfail[f] & 1u) << 1) | ((nanflag[f] & 1u) << 2));
  iters[f] = it[f];
}

template <int MAXD>
cudaError_t run_batch(const GlobalBpArgs& a, const uint32_t* syn, uint32_t* dec, uint8_t* flags, uint32_t* iters, int nf,
                      cudaStream_t st) {
  const int F = a.batch, m = a.m, n = a.n, E = a.m * a.dc;
  float* msg = a.msg;
  uint8_t* synb = a.bytes;            // [m][F]
  uint8_t* active = synb + (size_t)m * F;
  uint8_t* bad = active + F;
  uint8_t* convfail = bad + F;
  uint8_t* nanflag = convfail + F;
  uint8_t* mismatch = nanflag + F;
  uint32_t* it = a.words;             // [F]
  unsigned int* remaining = a.words + F;
  cudaMemsetAsync(convfail, 0, (size_t)3 * F, st);
  cudaMemsetAsync(dec, 0, (size_t)nf * a.nw * sizeof(uint32_t), st);
  g_init<<<std::min<long long>(((long long)E * F + 255) / 256, 148 * 32), 256, 0, st>>>(syn, a.mw, m, E, F, nf, a.prior, msg,
                                                                                     synb, active, bad);
  const dim3 gc((F + 127) / 128, m), gv((F + 127) / 128, n);
  for (int it_n = 0; it_n < a.maxit; ++it_n) {
    const int last = it_n == a.maxit - 1, ck = last || it_n % 10 == 0;
    g_check<MAXD><<<gc, 128, 0, st>>>(msg, synb, active, m, a.dc, F);
    g_var<MAXD><<<gv, 128, 0, st>>>(msg, a.vrow, active, bad, n, a.dv, F, a.prior, last, ck);
    if (ck) {
      cudaMemsetAsync(remaining, 0, sizeof(unsigned int), st);
      g_control<<<(F + 255) / 256, 256, 0, st>>>(active, bad, convfail, it, F, it_n, last, remaining);
      unsigned int left = 0;
      cudaError_t e = cudaMemcpyAsync(&left, remaining, sizeof left, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return e;
      if (left == 0) break;  // every frame of the batch has stopped
    }
  }
  const dim3 gd((nf + 127) / 128, n), ge((nf + 127) / 128, m);
  g_decide<<<gd, 128, 0, st>>>(msg, a.vrow, n, a.dv, F, nf, a.nw, dec, nanflag);
  g_verify<<<ge, 128, 0, st>>>(dec, a.cvar, synb, m, a.dc, F, nf, a.nw, mismatch);
  g_flags<<<(nf + 255) / 256, 256, 0, st>>>(mismatch, convfail, nanflag, it, nf, flags, iters);
  return cudaGetLastError();
}

}  // namespace

size_t global_bp_bytes(int m, int dc, int batch, size_t* msg_bytes, size_t* byte_bytes, size_t* word_bytes) {
  *msg_bytes = (size_t)m * dc * batch * sizeof(float);
  *byte_bytes = (size_t)m * batch + (size_t)5 * batch;
  *word_bytes = ((size_t)batch + 4) * sizeof(uint32_t);
  return *msg_bytes + *byte_bytes + *word_bytes;
}

cudaError_t global_bp_run(const GlobalBpArgs& a, const uint32_t* syn, uint32_t* dec, uint8_t* flags, uint32_t* iters,
                          int nframes, int* launches, cudaStream_t st) {
  const int maxd = std::max(a.dc, a.dv);
  if (maxd > 32) return cudaErrorInvalidValue;
  for (int off = 0; off < nframes; off += a.batch) {
    const int nf = std::min(a.batch, nframes - off);
    cudaError_t e;
    if (maxd <= 8) e = run_batch<8>(a, syn + (size_t)off * a.mw, dec + (size_t)off * a.nw, flags + off, iters + off, nf, st);
    else if (maxd <= 16) e = run_batch<16>(a, syn + (size_t)off * a.mw, dec + (size_t)off * a.nw, flags + off, iters + off, nf, st);
    else e = run_batch<32>(a, syn + (size_t)off * a.mw, dec + (size_t)off * a.nw, flags + off, iters + off, nf, st);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 1;
  }
  return cudaSuccess;
}

}  // namespace qldpc
