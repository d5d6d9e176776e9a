// Global-memory ("HBM-resident") belief-propagation path: the fallback for codes the shared-memory tile kernel
// (bp_kernel.cuh) does not cover -- a (check degree, variable degree) pair without a compiled instantiation, or a
// frame whose messages do not fit in shared memory.  Same arithmetic, operation for operation, as the tile kernel
// and the reference (EqNodeUpdate / VarNodeUpdate / CheckConvergence / Decode tail, DecoderCPU.h:150-390); only
// the residence of the message state differs.
//
// Layout: msg[row][S] with row = i*m + e (i = position of the edge in its check) and the SLOT index innermost; a
// thread handles one node for W (4, 2 or 1) consecutive slots with one 4*W-byte access per message row, so a warp
// moves up to 512 contiguous bytes per row and the path is bound by HBM at 16 bytes per edge-update.
// S frame slots are decoded concurrently.  Like the tile kernel, every slot has its own iteration counter, n % 10
// convergence cadence and `last` iteration; a slot that stops is syndrome-checked, written out and handed the next
// frame of a queue in the same pass, so slots never idle while frames remain (no lock-step batches).
//
// One pass = g_check, g_var (+ its last-iteration variant), g_control, then the three kernels that serve the slots
// which just stopped: g_verify_pack (syndrome of the decision; decision words of the frame out), g_handover (flags and
// iteration count out, next frame in) and g_fill (its syndrome bits).  Everything a stopped slot needs is produced by the passes themselves in BIT-PACKED,
// slot-innermost arrays -- one 32-bit word holds the bit of 32 consecutive slots: synw[check][S/32] (input syndrome),
// decw[variable][S/32] (hard decision, written by the variable kernel for every slot at a convergence checkpoint) --
// so serving the scattered stopped slots costs (m + E + n) * S / 8 bytes per pass, about 1% of the message traffic.
// A fresh slot is not initialised at all: its first check phase substitutes the prior for the messages.
// The host never synchronises inside a run: the completion count is mirrored into mapped pinned memory by g_control
// and read there; passes are enqueued at most 8 ahead of the device.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"

namespace qldpc {

namespace {

// kFresh = running its iteration 0: the messages are implicitly the prior (InitVarNodes, DecoderCPU.h:135-148)
enum : uint8_t { kIdle = 0, kRun = 1, kDone = 2, kFresh = 3 };
__device__ __forceinline__ bool running(uint8_t st) { return st == kRun || st == kFresh; }

constexpr int kGroup = 256;  // slots per control block

struct Slots {
  float* msg;          // [E][S]
  uint32_t* synw;      // [m][S/32] input syndrome bits of the frames in the slots (bit = slot % 32)
  uint32_t* decw;      // [n][S/32] hard decision after the last checkpoint variable phase
  uint32_t* mismatchw; // [S/32] decision syndrome != input syndrome, for the slots that just stopped
  uint32_t* donew;     // [S/32] slots in state kDone
  uint8_t* state;      // [S]
  uint8_t* bad;        // [S] an unconverged message was seen in the last variable phase
  uint8_t* nanflag;    // [S] a NaN message was seen in the last checkpoint variable phase
  int32_t* frame;      // [S] frame id in the slot
  int32_t* iter;       // [S] iteration index n of the slot
  unsigned int* ctr;   // [0] next frame to hand out, [1] frames completed, [2], [3] lengths of the two `lastq` lists
  uint32_t* lastq;     // [2][S] thread indices (slot / W) of the variable kernel with a slot entering its last iteration
  unsigned int* host_done;  // mapped pinned memory: mirror of ctr[1] for the host
};

// Every slot starts "stopped" with no frame to write out: g_handover gives it its first frame.
__global__ void __launch_bounds__(kGroup) g_start(Slots s, int S) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < S) {
    s.state[t] = kDone;
    s.frame[t] = -1;
    s.iter[t] = 0;
    s.bad[t] = s.nanflag[t] = 0;
  }
  if (t < S / 32) {
    s.donew[t] = 0xFFFFFFFFu;
    s.mismatchw[t] = 0u;
  }
  if (t < 4) s.ctr[t] = 0;
  if (t == 0) *s.host_done = 0u;
}

// W consecutive slots per thread, moved with one 4*W-byte access per message row: a pass is bound by HBM, and wide
// accesses are what keeps enough bytes in flight per SM.
template <int W> __device__ __forceinline__ void ld_f(float (&d)[W], const float* p) {
  if constexpr (W == 4) { const float4 v = *reinterpret_cast<const float4*>(p); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
  else if constexpr (W == 2) { const float2 v = *reinterpret_cast<const float2*>(p); d[0] = v.x; d[1] = v.y; }
  else d[0] = p[0];
}
template <int W> __device__ __forceinline__ void st_f(float* p, const float (&d)[W]) {
  if constexpr (W == 4) *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
  else if constexpr (W == 2) *reinterpret_cast<float2*>(p) = make_float2(d[0], d[1]);
  else p[0] = d[0];
}
template <int W> __device__ __forceinline__ void ld_b(uint8_t (&d)[W], const uint8_t* p) {
  if constexpr (W == 4) { const uchar4 v = *reinterpret_cast<const uchar4*>(p); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
  else if constexpr (W == 2) { const uchar2 v = *reinterpret_cast<const uchar2*>(p); d[0] = v.x; d[1] = v.y; }
  else d[0] = p[0];
}
template <int W> __device__ __forceinline__ void ld_i(int (&d)[W], const int32_t* p) {
  if constexpr (W == 4) { const int4 v = *reinterpret_cast<const int4*>(p); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
  else if constexpr (W == 2) { const int2 v = *reinterpret_cast<const int2*>(p); d[0] = v.x; d[1] = v.y; }
  else d[0] = p[0];
}

template <int W> __device__ __forceinline__ void st_b(uint8_t* p, const uint8_t (&d)[W]) {
  if constexpr (W == 4) *reinterpret_cast<uchar4*>(p) = make_uchar4(d[0], d[1], d[2], d[3]);
  else if constexpr (W == 2) *reinterpret_cast<uchar2*>(p) = make_uchar2(d[0], d[1]);
  else p[0] = d[0];
}

// A thread whose W slots are not all running still computes and stores all W lanes: the other lanes are idle slots
// (stopped slots only exist between g_control and g_handover), whose messages nobody reads.
// (7 resident blocks for the common degrees: 72 registers, no spills; +0.3% over 6 blocks, 8 blocks spill and lose 1%)
template <int MAXC, int W, bool EXACT>
__global__ void __launch_bounds__(128, (EXACT && W == 4 && MAXC <= 12) ? 7 : 1) g_check(Slots s, int m, int dc_rt, int S, float prior) {
  const int dc = EXACT ? MAXC : dc_rt;
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  const int e = blockIdx.y;
  if (f >= S) return;
  uint8_t st[W];
  ld_b<W>(st, s.state + f);
  bool any = false, anyfresh = false;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    any |= running(st[w]);
    anyfresh |= st[w] == kFresh;
  }
  if (!any) return;
  float t[MAXC][W];
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (i < dc) {
      ld_f<W>(t[i], s.msg + ((size_t)i * m + e) * S + f);
#pragma unroll
      for (int w = 0; w < W; ++w) t[i][w] = __fmaf_rn(-2.0f, t[i][w], 1.0f);  // 1 - 2q (DecoderCPU.h:175)
    }
  if (anyfresh) {
    const float tp = __fmaf_rn(-2.0f, prior, 1.0f);
#pragma unroll
    for (int i = 0; i < MAXC; ++i)
      if (i < dc) {
#pragma unroll
        for (int w = 0; w < W; ++w)
          if (st[w] == kFresh) t[i][w] = tp;
      }
  }
  const uint32_t sw = s.synw[(size_t)e * (S >> 5) + (f >> 5)] >> (f & 31);  // W divides 32: one word holds the W bits
  float cf[W], pre[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    cf[w] = (sw >> w) & 1u ? 0.5f : -0.5f;  // DecoderCPU.h:178-183, see bp_kernel.cuh
    pre[w] = 1.0f;
  }
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    if (i < dc) {
      float p[W];  // reference order: 1.0f * t0 * ... skipping i, left to right (DecoderCPU.h:168-176)
#pragma unroll
      for (int w = 0; w < W; ++w) p[w] = pre[w];
#pragma unroll
      for (int k = i + 1; k < MAXC; ++k)
        if (k < dc) {
#pragma unroll
          for (int w = 0; w < W; ++w) p[w] = __fmul_rn(p[w], t[k][w]);
        }
#pragma unroll
      for (int w = 0; w < W; ++w) {
        p[w] = __fmaf_rn(cf[w], p[w], 0.5f);
        pre[w] = __fmul_rn(pre[w], t[i][w]);
      }
      st_f<W>(s.msg + ((size_t)i * m + e) * S + f, p);
    }
  }
}

// Two launches per pass: LAST = false serves the threads none of whose slots is in its final iteration (no
// full-product registers: higher occupancy for a bandwidth-bound kernel), LAST = true the others (DecoderCPU.h:284).
// The latter works from the list g_control compiled in the previous pass (`lastq`, usually empty or short), or
// over all threads if `lastq` is null (one-iteration runs, where every slot is in its last iteration).
// GUARD: the division range tests of the hot variant, as in the tile kernel (decoder.cu:division_guard): 0 = the host
// proved that neither can fire, 1 = only the numerator test could, and the product chains run scaled by 2^64 instead
// (exact, bp_kernel.cuh:var_phase), 3 = both tests with the __fdiv_rn fallback.  The tests and the fallback's control flow
// are a fifth of this kernel's instructions, and it runs at 60% issue utilisation next to its memory traffic.
template <int MAXV, int W, bool EXACT, bool LAST, int GUARD = 3>
__device__ __forceinline__ void g_var_body(const Slots& s, const uint32_t* __restrict__ vrow, int n, int dv_rt, int S,
                                           float prior, int last_it, int t, int v) {
  const int dv = EXACT ? MAXV : dv_rt;
  const int f = t * W;
  const int SW = S >> 5;
  uint8_t st[W];
  int it[W];
  bool any = false, anylast = false, anyck = false, last[W], ck[W];
#pragma unroll
  for (int w = 0; w < W; ++w) { st[w] = kIdle; it[w] = 0; last[w] = ck[w] = false; }
  if (f < S) {
    ld_b<W>(st, s.state + f);
    ld_i<W>(it, s.iter + f);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const bool run = running(st[w]);
      last[w] = run && it[w] == last_it;              // DecoderCPU.h:284
      ck[w] = run && (last[w] || it[w] % 10 == 0);    // DecoderCPU.h:287
      any |= run;
      anylast |= last[w];
      anyck |= ck[w];
    }
  }
  const bool active = any && anylast == LAST;
  // The LAST = false launch keeps whole warps alive: the decision bits of 32 consecutive slots share a word, which
  // the lanes assemble together below.  The list-driven LAST = true launch has no such alignment and uses atomics.
  if (!active) {
    if (LAST) return;
    anyck = false;
#pragma unroll
    for (int w = 0; w < W; ++w) ck[w] = false;
  }
  uint8_t bit[W];
#pragma unroll
  for (int w = 0; w < W; ++w) bit[w] = 0;
  if (active) {
  // Only the messages are kept in registers; the complements 1 - p are recomputed where they are used (one FADD,
  // same rounding), which is cheaper than the occupancy their registers would cost a bandwidth-bound kernel.
  float pk[MAXV][W];
  uint32_t row[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
    if (k < dv) {
      row[k] = vrow[(size_t)k * n + v];
      ld_f<W>(pk[k], s.msg + (size_t)row[k] * S + f);
    }
  const float prior1 = __fsub_rn(1.0f, prior);  // DecoderCPU.h:209-210
  float preP[W], preQ[W], fullP[LAST ? W : 1], fullQ[LAST ? W : 1];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    preP[w] = prior;
    preQ[w] = prior1;
    if (LAST) { fullP[w] = prior; fullQ[w] = prior1; }
  }
  if (LAST) {
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
      if (k < dv) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
          fullQ[w] = __fmul_rn(fullQ[w], __fsub_rn(1.0f, pk[k][w]));
          fullP[w] = __fmul_rn(fullP[w], pk[k][w]);
        }
      }
  }
  bool anybad[W], anynan[W];
#pragma unroll
  for (int w = 0; w < W; ++w) anybad[w] = anynan[w] = false;
  if constexpr (!LAST && W % 2 == 0) {
    // Hot variant: two slots per instruction with the packed fp32x2 arithmetic of sm_100 (bp_kernel.cuh: Pack<2>), each
    // half an independent IEEE operation.  The variable kernel issues ~60% of its slots with scalar arithmetic, which
    // keeps it below the HBM roofline; packing halves the multiplies and the division's FMAs.
    constexpr int H = W / 2;
    typedef Pack<2> P2;
    P2 pk2[MAXV][H], om2[MAXV][H], pP[H], pQ[H];
    uint32_t lowest[W];
#pragma unroll
    for (int w = 0; w < W; ++w) lowest[w] = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
      if (k < dv) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          pk2[k][h] = P2::load(&pk[k][2 * h]);
          om2[k][h] = pfma(pk2[k][h], P2::splat(-1.0f), P2::splat(1.0f));  // 1 - p, one rounding (DecoderCPU.h:220)
        }
      }
    constexpr int kTests = GUARD == 1 ? 0 : GUARD;
    const float scale = GUARD == 1 ? 18446744073709551616.0f : 1.0f;  // 2^64
#pragma unroll
    for (int h = 0; h < H; ++h) { pP[h] = P2::splat(__fmul_rn(prior, scale)); pQ[h] = P2::splat(__fmul_rn(prior1, scale)); }
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      if (j < dv) {
        float q[W];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          P2 P = pP[h], Q = pQ[h];
#pragma unroll
          for (int k = j + 1; k < MAXV; ++k)
            if (k < dv) {
              Q = pmul(Q, om2[k][h]);
              P = pmul(P, pk2[k][h]);
            }
          P2 nden;  // -(Q + P) by scalar adds: ptxas would contract a packed mul + add into one FFMA2 (bp_kernel.cuh)
          nden.set(0, __fadd_rn(-Q.get(0), -P.get(0)));
          nden.set(1, __fadd_rn(-Q.get(1), -P.get(1)));
          bool unsafe = false;
          P2 out = div_fast_pack<kTests, 2>(P, nden, unsafe);  // == P / (Q + P), DecoderCPU.h:223
          if (kTests != 0 && unsafe) {
            out.set(0, __fdiv_rn(P.get(0), -nden.get(0)));
            out.set(1, __fdiv_rn(P.get(1), -nden.get(1)));
          }
          q[2 * h] = out.get(0);
          q[2 * h + 1] = out.get(1);
          if (j < dv - 1) {
            pQ[h] = pmul(pQ[h], om2[j][h]);
            pP[h] = pmul(pP[h], pk2[j][h]);
          }
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
          // saturation test as a running minimum in a volatile asm statement (bp_kernel.cuh:sat_update): without the
          // range tests' control flow around it, the plain comparison came out of CUDA 12.9 reporting converged slots
          // as unconverged here as well (10 extra iterations; the parity tests of this path catch it)
          sat_update(lowest[w], q[w]);
          anynan[w] |= q[w] != q[w];
          bit[w] |= q[w] >= 0.5f;  // hard decision: any edge message >= 0.5f (DecoderCPU.h:354-373)
        }
        st_f<W>(s.msg + (size_t)row[j] * S + f, q);
      }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) anybad[w] = lowest[w] < (0x3F7D70A4u - 0x3C23D70Au - 1u);  // == OR of unconverged(q)
  } else {
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    if (j < dv) {
      float P[W], Q[W];
#pragma unroll
      for (int w = 0; w < W; ++w) { P[w] = preP[w]; Q[w] = preQ[w]; }
#pragma unroll
      for (int k = j + 1; k < MAXV; ++k)
        if (k < dv) {
#pragma unroll
          for (int w = 0; w < W; ++w) {
            Q[w] = __fmul_rn(Q[w], __fsub_rn(1.0f, pk[k][w]));
            P[w] = __fmul_rn(P[w], pk[k][w]);
          }
        }
      float q[W];
#pragma unroll
      for (int w = 0; w < W; ++w) {
        if (LAST && last[w]) { P[w] = fullP[w]; Q[w] = fullQ[w]; }
        const float den = __fadd_rn(Q[w], P[w]);
        bool unsafe = false;
        q[w] = div_fast<3>(P[w], den, unsafe);  // == P / (Q + P), DecoderCPU.h:223
        if (unsafe) q[w] = __fdiv_rn(P[w], den);
        anybad[w] |= unconverged(q[w]);
        anynan[w] |= q[w] != q[w];
        bit[w] |= q[w] >= 0.5f;  // hard decision: any edge message >= 0.5f (DecoderCPU.h:354-373)
        preQ[w] = __fmul_rn(preQ[w], __fsub_rn(1.0f, pk[j][w]));
        preP[w] = __fmul_rn(preP[w], pk[j][w]);
      }
      st_f<W>(s.msg + (size_t)row[j] * S + f, q);
    }
  }
  }
#pragma unroll
  for (int w = 0; w < W; ++w)
    if (ck[w]) {
      if (anybad[w]) s.bad[f + w] = 1;
      if (anynan[w]) s.nanflag[f + w] = 1;
    }
  }  // active
  // A slot can only stop at a checkpoint, so this is where its decision is recorded: bit (slot % 32) of
  // decw[v][slot / 32], only for the slots that are at a checkpoint in this pass.
  uint32_t val = 0, mask = 0;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    val |= (uint32_t)(bit[w] & 1u) << w;
    mask |= (uint32_t)ck[w] << w;
  }
  if (LAST) {
    if (mask) {
      uint32_t* word = s.decw + (size_t)v * SW + (f >> 5);
      const int sh = f & 31;
      atomicAnd(word, ~((mask & ~val) << sh));
      atomicOr(word, (mask & val) << sh);
    }
    return;
  }
  if (!__any_sync(0xffffffffu, anyck)) return;
  constexpr int G = 32 / W;  // lanes per word
  const int lane = threadIdx.x & 31, sh = W * (lane % G);
  val <<= sh;
  mask <<= sh;
#pragma unroll
  for (int d = 1; d < G; d <<= 1) {
    val |= __shfl_xor_sync(0xffffffffu, val, d);
    mask |= __shfl_xor_sync(0xffffffffu, mask, d);
  }
  if (lane % G == 0 && mask && f < S) {
    uint32_t* word = s.decw + (size_t)v * SW + (f >> 5);
    *word = (*word & ~mask) | (val & mask);
  }
}

// Resident blocks per SM asked of the compiler for the hot variants: the kernel is bound by the bytes it keeps in flight
// (a thread moves only dv * 16 B, and spends most of its lifetime waiting or computing with no load outstanding), so
// occupancy is worth more than instructions or a few spilled registers here.  Measured on J4K5L10P61 (both sides
// interleaved): 6 / 7 / 8 / 9 / 10 blocks of 128 threads -> 0.74 / 0.786 / 0.80 / 0.78 / 0.77 of the HBM roofline
// (8 blocks = 64 registers, 8-32 B of spills).
constexpr int var_min_blocks(int maxv, int w, bool last) { return !last && w == 4 && maxv <= 5 ? 8 : 1; }
template <int MAXV, int W, bool EXACT, bool LAST, int GUARD = 3>
__global__ void __launch_bounds__(128, var_min_blocks(MAXV, W, LAST)) g_var(Slots s, const uint32_t* __restrict__ vrow, int n, int dv_rt, int S,
                                             float prior, int last_it, const uint32_t* __restrict__ lastq,
                                             const unsigned int* __restrict__ lastq_len) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (LAST && lastq) {
    // The list-driven launch uses a small grid and strides over the list: it is short or empty in almost every pass,
    // and a full-size grid of threads that only find that out costs as much as a tenth of a pass.
    const int len = (int)*lastq_len;
    for (int li = t; li < len; li += gridDim.x * blockDim.x)
      g_var_body<MAXV, W, EXACT, LAST, GUARD>(s, vrow, n, dv_rt, S, prior, last_it, (int)lastq[li], blockIdx.y);
  } else {
    g_var_body<MAXV, W, EXACT, LAST, GUARD>(s, vrow, n, dv_rt, S, prior, last_it, t, blockIdx.y);
  }
}

// BeliefPropogation loop control (DecoderCPU.h:280-291), per slot, after the variable phase.  Also publishes which
// slots stopped (one bit per slot, for g_verify / g_finish) and mirrors the completion count for the host.
__global__ void __launch_bounds__(kGroup) g_control(Slots s, int S, int last_it, int wv, int parity) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f == 0) {
    s.ctr[2 + (parity ^ 1)] = 0;  // the list the next pass's g_control appends to; its reader has run
    *s.host_done = s.ctr[1];      // every frame counted here has all its outputs written (earlier kernels)
  }
  bool next_is_last = false, done = false;
  if (f < S && running(s.state[f])) {
    const int it = s.iter[f];
    const bool last = it == last_it, ck = last || it % 10 == 0;
    if (last || (ck && !s.bad[f])) {
      s.state[f] = kDone;  // bad / nanflag are kept: CONVERGENCE_FAIL and the NaN bit describe the final state
      done = true;
    } else {
      s.state[f] = kRun;
      s.iter[f] = it + 1;
      s.bad[f] = s.nanflag[f] = 0;
      next_is_last = it + 1 == last_it;
    }
  }
  const int lane = threadIdx.x & 31;
  const unsigned donemask = __ballot_sync(0xffffffffu, done);
  if (lane == 0 && f < S) s.donew[f >> 5] = donemask;
  // one list entry per variable-kernel thread (wv consecutive slots): the lowest flagged slot of the group appends
  const unsigned flagged = __ballot_sync(0xffffffffu, next_is_last);
  const int first = lane & ~(wv - 1);
  const unsigned group = (flagged >> first) & ((1u << wv) - 1u);
  if (next_is_last && (group & ((1u << (lane - first)) - 1u)) == 0)
    s.lastq[(size_t)parity * S + atomicAdd(&s.ctr[2 + parity], 1u)] = (uint32_t)(f / wv);
}

// Serves the slots that just stopped, 32 slots (one bit-word) per thread, blockIdx.y striding over work units:
//  * units [0, m): syndrome of the decision against the input syndrome (DecoderCPU.h:380-384) -- the thread XORs the
//    decision words of the dc variables of check `unit` onto its syndrome word; any bit left set is a mismatch
//    (partial results meet in mismatchw by atomicOr);
//  * units [m, m + nw): word `unit - m` of the packed decision rows of the frames leaving -- a 32 x 32 bit transpose
//    of decw[32 variables][this word], one output word per stopped slot.
__global__ void __launch_bounds__(128) g_verify_pack(Slots s, const uint32_t* __restrict__ cvar, int m, int dc, int n,
                                                     int nw, int S, uint32_t* __restrict__ dec) {
  const int SW = S >> 5, wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= SW) return;
  const uint32_t done = s.donew[wi];
  if (!done) return;
  uint32_t any = 0;
  for (int u = blockIdx.y; u < m + nw; u += gridDim.y) {
    if (u < m) {
      uint32_t par = s.synw[(size_t)u * SW + wi];
      for (int i = 0; i < dc; ++i) par ^= s.decw[(size_t)cvar[(size_t)i * m + u] * SW + wi];
      any |= par;
    } else {
      const int word = u - m, v0 = word * 32, cnt = min(32, n - v0);
      uint32_t col[32];
#pragma unroll
      for (int b = 0; b < 32; ++b) col[b] = b < cnt ? s.decw[(size_t)(v0 + b) * SW + wi] : 0u;
      uint32_t left = done;
      while (left) {  // one output word per stopped slot of this thread's 32
        const int sl = __ffs((int)left) - 1;
        left &= left - 1u;
        const int fr = s.frame[wi * 32 + sl];
        if (fr < 0) continue;
        uint32_t out = 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) out |= ((col[b] >> sl) & 1u) << b;
        dec[(size_t)fr * nw + word] = out;
      }
    }
  }
  any &= done;
  if (any) atomicOr(&s.mismatchw[wi], any);
}

// Per-frame outputs of the stopped slots, then the hand-over to the next frame of the queue; one warp per 32 slots
// (lane = slot), one queue atomic per warp.  Leaves the mask of the slots that received a frame in donew (g_fill's input).
__global__ void __launch_bounds__(kGroup) g_handover(Slots s, int S, int nframes, uint8_t* __restrict__ flags,
                                                     uint32_t* __restrict__ iters) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  if (f - lane >= S) return;  // S is a multiple of 32: whole warps
  const int wi = f >> 5;
  const uint32_t donemask = s.donew[wi];
  if (!donemask) return;
  const bool done = (donemask >> lane) & 1u;
  const int fr = done ? s.frame[f] : -1;
  if (fr >= 0) {
    // CONVERGENCE_FAIL = !CheckConvergence(final messages), DecoderCPU.h:375-378
    const uint32_t mis = (s.mismatchw[wi] >> lane) & 1u;
    flags[fr] = (uint8_t)(mis | ((s.bad[f] & 1u) << 1) | ((s.nanflag[f] & 1u) << 2));
    iters[fr] = (uint32_t)(s.iter[f] + 1);
  }
  const unsigned leaving = __ballot_sync(0xffffffffu, fr >= 0);
  unsigned base = 0;
  if (lane == 0) {
    if (leaving) atomicAdd(&s.ctr[1], (unsigned)__popc(leaving));
    base = atomicAdd(&s.ctr[0], (unsigned)__popc(donemask));
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  int next = -1;
  if (done) {
    const unsigned ticket = base + (unsigned)__popc(donemask & ((1u << lane) - 1u));
    next = ticket < (unsigned)nframes ? (int)ticket : -1;
    s.bad[f] = s.nanflag[f] = 0;
    s.iter[f] = 0;
    s.frame[f] = next;
    s.state[f] = next >= 0 ? kFresh : kIdle;
  }
  const unsigned fresh = __ballot_sync(0xffffffffu, next >= 0);
  if (lane == 0) {
    s.mismatchw[wi] = 0u;
    s.donew[wi] = fresh;  // consumed (and cleared for the next pass by g_control's rewrite) by g_fill
  }
}

// Syndrome bits of the frames just handed out: bit `slot % 32` of synw[e][slot / 32]; the other slots of the word keep
// theirs.  One thread per word, blockIdx.y striding over the checks.
__global__ void __launch_bounds__(128) g_fill(Slots s, const uint32_t* __restrict__ syn, int mw, int m, int S) {
  const int SW = S >> 5, wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= SW) return;
  const uint32_t fresh = s.donew[wi];
  if (!fresh) return;
  for (int ew = blockIdx.y; ew < mw; ew += gridDim.y) {  // 32 checks per step: one syndrome word of every fresh frame
    uint32_t col[32];
    uint32_t left = fresh;
#pragma unroll
    for (int b = 0; b < 32; ++b) col[b] = 0u;
    while (left) {
      const int sl = __ffs((int)left) - 1;
      left &= left - 1u;
      const uint32_t w = syn[(size_t)s.frame[wi * 32 + sl] * mw + ew];
#pragma unroll
      for (int b = 0; b < 32; ++b) col[b] |= ((w >> b) & 1u) << sl;
    }
    const int cnt = min(32, m - ew * 32);
#pragma unroll
    for (int b = 0; b < 32; ++b)
      if (b < cnt) {
        uint32_t* word = s.synw + (size_t)(ew * 32 + b) * SW + wi;
        *word = (*word & ~fresh) | col[b];
      }
  }
}

using CheckFn = void (*)(Slots, int, int, int, float);
using VarFn = void (*)(Slots, const uint32_t*, int, int, int, float, int, const uint32_t*, const unsigned int*);
constexpr int check_width(int maxc) { return maxc <= 16 ? 4 : 2; }
constexpr int var_width(int maxv) { return maxv <= 8 ? 4 : maxv <= 16 ? 2 : 1; }

// Common degrees get an instantiation with the degree as a compile-time constant (fully unrolled, messages in
// registers); larger ones a bounded instantiation with the degree tested at run time.
void pick_check(int dc, CheckFn& fn, int& w) {
#define QLDPC_E(N) if (dc == N) { fn = g_check<N, check_width(N), true>; w = check_width(N); return; }
#define QLDPC_B(N) if (dc <= N) { fn = g_check<N, check_width(N), false>; w = check_width(N); return; }
  QLDPC_E(2) QLDPC_E(3) QLDPC_E(4) QLDPC_E(5) QLDPC_E(6) QLDPC_E(7) QLDPC_E(8) QLDPC_E(9) QLDPC_E(10) QLDPC_E(11)
  QLDPC_E(12) QLDPC_E(13) QLDPC_E(14) QLDPC_E(15) QLDPC_E(16) QLDPC_B(24) QLDPC_B(32)
#undef QLDPC_E
#undef QLDPC_B
}
void pick_var(int dv, int guard, VarFn& fn, VarFn& fn_last, int& w) {
  // the common small degrees also come without the division range tests / with scaled chains (g_var_body: GUARD)
#define QLDPC_G(N)                                                                                        \
  if (dv == N && guard != 3) {                                                                            \
    fn = guard == 0 ? g_var<N, var_width(N), true, false, 0> : g_var<N, var_width(N), true, false, 1>;    \
    fn_last = g_var<N, var_width(N), true, true>;                                                         \
    w = var_width(N);                                                                                     \
    return;                                                                                               \
  }
  QLDPC_G(2) QLDPC_G(3) QLDPC_G(4) QLDPC_G(5) QLDPC_G(6)
#undef QLDPC_G
#define QLDPC_E(N) \
  if (dv == N) { fn = g_var<N, var_width(N), true, false>; fn_last = g_var<N, var_width(N), true, true>; w = var_width(N); return; }
#define QLDPC_B(N) \
  if (dv <= N) { fn = g_var<N, var_width(N), false, false>; fn_last = g_var<N, var_width(N), false, true>; w = var_width(N); return; }
  QLDPC_B(1) QLDPC_E(2) QLDPC_E(3) QLDPC_E(4) QLDPC_E(5) QLDPC_E(6) QLDPC_E(7) QLDPC_E(8) QLDPC_E(9) QLDPC_E(10)
  QLDPC_E(11) QLDPC_E(12) QLDPC_B(16) QLDPC_B(32)
#undef QLDPC_E
#undef QLDPC_B
}

// One run of one side, driven pass by pass so that the two sides of a decoder can be interleaved on two streams.
struct GlobalRun {
  static constexpr int kAhead = 8;  // passes enqueued beyond the one the host has seen finish
  GlobalBpArgs a;
  const uint32_t* syn = nullptr;
  uint32_t* dec = nullptr;
  uint8_t* flags = nullptr;
  uint32_t* iters = nullptr;
  int nframes = 0, S = 0, sb = 0, last_it = 0, wv = 1;
  cudaStream_t st = nullptr;
  Slots s;
  CheckFn check = nullptr;
  VarFn var = nullptr, var_last = nullptr;
  dim3 gc, gv, ge, gf;
  cudaEvent_t ev[kAhead] = {};
  bool finished = false, begun = false, use_graphs = false;
  cudaGraphExec_t gexec[4] = {};

  cudaError_t begin(const GlobalBpArgs& a_, const uint32_t* syn_, uint32_t* dec_, uint8_t* flags_, uint32_t* iters_,
                    int nframes_, cudaStream_t st_) {
    a = a_; syn = syn_; dec = dec_; flags = flags_; iters = iters_; nframes = nframes_; st = st_;
    begun = true;
    use_graphs = nframes >= 4096 && getenv("QLDPC_GLOBAL_NO_GRAPH") == nullptr;  // short runs: not worth four captures
    const int m = a.m, n = a.n, E = a.m * a.dc;
    // Slots in flight: enough that one pass moves ~1 GB (launch overhead out of sight), about a quarter of the frames
    // so that every slot is refilled a few times and the straggler tail stays short, at most what was allocated.
    const long long want = std::max<long long>(((long long)nframes + 3) / 4, (long long)(1.0e9 / (16.0 * E)));
    // (heuristic: at most 48k slots -- measured optimum of the n=610 code at 1M frames; beyond it the tail grows faster
    // than the per-pass overheads shrink)
    const long long asked = a.slots > 0 ? ((long long)a.slots + 31) / 32 * 32 : (std::min<long long>(want, 49152) + 127) / 128 * 128;
    S = (int)std::max<long long>(32, std::min<long long>(a.batch, asked));
    const int SW = S / 32;
    s.msg = a.msg;
    s.state = a.bytes;
    s.bad = s.state + S;
    s.nanflag = s.bad + S;
    s.synw = a.words;
    s.decw = s.synw + (size_t)m * SW;
    s.mismatchw = s.decw + (size_t)n * SW;
    s.donew = s.mismatchw + SW;
    s.frame = (int32_t*)(a.words + ((size_t)(m + n + 2) * SW + 3) / 4 * 4);  // 16-byte aligned for the 4-slot loads
    s.iter = s.frame + S;
    s.ctr = (unsigned int*)(s.iter + S);
    s.lastq = s.ctr + 4;
    s.host_done = a.host_done;
    int wc = 1;
    pick_check(a.dc, check, wc);
    pick_var(a.dv, a.guard, var, var_last, wv);
    sb = (S + kGroup - 1) / kGroup;
    gc = dim3((S + 128 * wc - 1) / (128 * wc), m);
    gv = dim3((S + 128 * wv - 1) / (128 * wv), n);
    // the bit-word kernels: SW / 128 blocks of words, times enough y-strides over their units to fill the machine
    const int vx = (SW + 127) / 128;
    ge = dim3(vx, std::max(1, std::min(m + a.nw, 4096 / std::max(vx, 1))));
    gf = dim3(vx, std::max(1, std::min(a.mw, 4096 / std::max(vx, 1))));
    last_it = a.maxit - 1;
    for (int i = 0; i < kAhead; ++i) {
      cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    g_start<<<sb, kGroup, 0, st>>>(s, S);
    return cudaGetLastError();
  }

  // A pass first serves the slots that stopped in the previous pass (before the first pass: all of them, with no
  // frame to write out), then runs one BP iteration on every running slot.  Completion is read from mapped pinned
  // memory (g_control mirrors the count there); the host enqueues at most kAhead passes beyond the one it has seen
  // finish, so the device never waits for the host and the host never synchronises the stream inside a run.
  // Passes enqueued after the last frame has left find every slot idle and return at once.
  // Returns with `finished` set once the host has seen every frame leave.
  cudaError_t step(long long pass) {
    if (finished) return cudaSuccess;
    if (pass >= kAhead) {
      cudaError_t err = cudaEventSynchronize(ev[pass % kAhead]);  // pass - kAhead has finished
      if (err != cudaSuccess) return err;
      if (*(volatile unsigned int*)a.host_done >= (unsigned)nframes) {
        finished = true;
        return cudaSuccess;
      }
    }
    // From the second pass on a pass is one of four fixed launch sequences (with / without the list-driven `last`
    // launch, list parity 0 / 1): each is captured once per run and replayed as a CUDA graph, which takes the launch
    // gaps between its 6-7 kernels from ~5 us to ~1 us.
    const int gi = (pass >= last_it ? 2 : 0) + (int)(pass & 1);
    if (use_graphs && pass > 0 && last_it > 0) {
      cudaError_t err = cudaSuccess;
      if (!gexec[gi]) {
        cudaGraph_t g = nullptr;
        err = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (err == cudaSuccess) {
          launch_pass(pass);
          err = cudaStreamEndCapture(st, &g);
        }
        if (err == cudaSuccess) err = cudaGraphInstantiate(&gexec[gi], g, 0);
        if (g) cudaGraphDestroy(g);
        if (err != cudaSuccess) return err;
      }
      err = cudaGraphLaunch(gexec[gi], st);
      if (err != cudaSuccess) return err;
    } else {
      launch_pass(pass);
    }
    cudaError_t err = cudaEventRecord(ev[pass % kAhead], st);
    if (err != cudaSuccess) return err;
    if (pass == 0) return cudaGetLastError();  // a bad launch configuration shows here
    return cudaSuccess;
  }

  void launch_pass(long long pass) {
    const int n = a.n, m = a.m;
    if (pass > 0) g_verify_pack<<<ge, 128, 0, st>>>(s, a.cvar, m, a.dc, n, a.nw, S, dec);
    g_handover<<<sb, kGroup, 0, st>>>(s, S, nframes, flags, iters);
    g_fill<<<gf, 128, 0, st>>>(s, syn, a.mw, m, S);
    check<<<gc, 128, 0, st>>>(s, m, a.dc, S, a.prior);
    const int parity = (int)(pass & 1);
    var<<<gv, 128, 0, st>>>(s, a.vrow, n, a.dv, S, a.prior, last_it, nullptr, nullptr);
    if (last_it == 0)
      var_last<<<gv, 128, 0, st>>>(s, a.vrow, n, a.dv, S, a.prior, last_it, nullptr, nullptr);
    else if (pass >= last_it)  // no slot is that old before; the list was written by the previous pass's g_control
      var_last<<<dim3(std::min<unsigned>(gv.x, 8u), n), 128, 0, st>>>(s, a.vrow, n, a.dv, S, a.prior, last_it,
                                                                       s.lastq + (size_t)(parity ^ 1) * S,
                                                                       s.ctr + 2 + (parity ^ 1));
    g_control<<<sb, kGroup, 0, st>>>(s, S, last_it, wv, parity);
  }

  cudaError_t end(cudaError_t err) {
    if (err == cudaSuccess && begun) err = cudaStreamSynchronize(st);  // drains the (empty) passes enqueued ahead
    for (int i = 0; i < kAhead; ++i)
      if (ev[i]) cudaEventDestroy(ev[i]);
    for (int i = 0; i < 4; ++i)
      if (gexec[i]) cudaGraphExecDestroy(gexec[i]);
    if (err != cudaSuccess) return err;
    return cudaGetLastError();
  }
};

}  // namespace

size_t global_bp_bytes(int m, int n, int dc, int batch, size_t* msg_bytes, size_t* byte_bytes, size_t* word_bytes) {
  *msg_bytes = (size_t)m * dc * batch * sizeof(float);
  *byte_bytes = (size_t)3 * batch;  // state, bad, nanflag
  // synw, decw [m + n][batch / 32], mismatchw, donew [batch / 32], frame, iter [batch], 4 counters, lastq [2][batch]
  *word_bytes = ((size_t)(m + n + 2) * (batch / 32) + 4 + (size_t)4 * batch + 4) * sizeof(uint32_t);
  return *msg_bytes + *byte_bytes + *word_bytes;
}

cudaError_t global_bp_run(const GlobalBpArgs& a, const uint32_t* syn, uint32_t* dec, uint8_t* flags, uint32_t* iters,
                          int nframes, int* launches, cudaStream_t st) {
  const int maxd = std::max(a.dc, a.dv);
  if (maxd > 32) return cudaErrorInvalidValue;
  if (nframes <= 0) return cudaSuccess;
  if (launches) *launches += 1;
  GlobalRun r;
  cudaError_t err = r.begin(a, syn, dec, flags, iters, nframes, st);
  for (long long pass = 0; err == cudaSuccess && !r.finished; ++pass) err = r.step(pass);
  return r.end(err);
}

// Both sides of a decoder at once, each on its own stream, their passes enqueued alternately by the one host thread:
// the service kernels, launch gaps and straggler tail of one side are covered by the other side's message traffic.
cudaError_t global_bp_run_pair(const GlobalBpArgs& ax, const uint32_t* synX, uint32_t* decX, uint8_t* flagsX, uint32_t* itersX,
                               cudaStream_t stX, const GlobalBpArgs& az, const uint32_t* synZ, uint32_t* decZ,
                               uint8_t* flagsZ, uint32_t* itersZ, cudaStream_t stZ, int nframes) {
  if (std::max(std::max(ax.dc, ax.dv), std::max(az.dc, az.dv)) > 32) return cudaErrorInvalidValue;
  if (nframes <= 0) return cudaSuccess;
  GlobalRun r[2];
  cudaError_t err = r[0].begin(ax, synX, decX, flagsX, itersX, nframes, stX);
  if (err == cudaSuccess) err = r[1].begin(az, synZ, decZ, flagsZ, itersZ, nframes, stZ);
  for (long long pass = 0; err == cudaSuccess && !(r[0].finished && r[1].finished); ++pass) {
    err = r[0].step(pass);
    if (err == cudaSuccess) err = r[1].step(pass);
  }
  const cudaError_t e0 = r[0].end(err), e1 = r[1].end(err);
  return e0 != cudaSuccess ? e0 : e1;
}

}  // namespace qldpc
