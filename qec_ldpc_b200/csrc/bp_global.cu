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
// One pass = g_check, g_var (+ its last-iteration variant), g_control, then the kernels that serve the slots which
// just stopped: g_verify, g_pack, g_handover, g_fill.  Everything a stopped slot needs is produced in slot-innermost
// arrays by the passes themselves (the variable kernel writes the hard decision of every slot that is at a
// convergence checkpoint, one byte per variable), so these kernels stay coalesced even though the stopped slots are
// scattered; a fresh slot is not initialised at all -- its first check phase substitutes the prior for the messages.
#include <algorithm>
#include <cstdint>

#include "kernels.cuh"

namespace qldpc {

namespace {

// kFresh = running its iteration 0: the messages are implicitly the prior (InitVarNodes, DecoderCPU.h:135-148)
enum : uint8_t { kIdle = 0, kRun = 1, kDone = 2, kFresh = 3 };
__device__ __forceinline__ bool running(uint8_t st) { return st == kRun || st == kFresh; }

constexpr int kGroup = 256;  // slots per control block

struct Slots {
  float* msg;          // [E][S]
  uint8_t* synb;       // [m][S] input syndrome bit of the frame in the slot
  uint8_t* state;      // [S]
  uint8_t* bad;        // [S] an unconverged message was seen in the last variable phase
  uint8_t* nanflag;    // [S] a NaN message was seen in the last checkpoint variable phase
  uint8_t* mismatch;   // [S]
  uint8_t* decb;       // [n][S] hard decision after the last checkpoint variable phase, one byte per bit
  int32_t* frame;      // [S] frame id in the slot
  int32_t* iter;       // [S] iteration index n of the slot
  unsigned int* ctr;   // [0] next frame to hand out, [1] frames completed, [2], [3] lengths of the two `lastq` lists
  uint32_t* lastq;     // [2][S] thread indices (slot / W) of the variable kernel with a slot entering its last iteration
};

// Every slot starts "stopped" with no frame to write out: g_handover gives it its first frame.
__global__ void __launch_bounds__(kGroup) g_start(Slots s, int S) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < S) {
    s.state[t] = kDone;
    s.frame[t] = -1;
    s.iter[t] = 0;
    s.bad[t] = s.nanflag[t] = s.mismatch[t] = 0;
  }
  if (t < 4) s.ctr[t] = 0;
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
template <int MAXC, int W, bool EXACT>
__global__ void __launch_bounds__(128) g_check(Slots s, int m, int dc_rt, int S, float prior) {
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
  uint8_t sb[W];
  ld_b<W>(sb, s.synb + (size_t)e * S + f);
  float cf[W], pre[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    cf[w] = sb[w] ? 0.5f : -0.5f;  // DecoderCPU.h:178-183, see bp_kernel.cuh
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
template <int MAXV, int W, bool EXACT, bool LAST>
__global__ void __launch_bounds__(128) g_var(Slots s, const uint32_t* __restrict__ vrow, int n, int dv_rt, int S,
                                             float prior, int last_it, const uint32_t* __restrict__ lastq,
                                             const unsigned int* __restrict__ lastq_len) {
  const int dv = EXACT ? MAXV : dv_rt;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (LAST && lastq) {
    if (t >= (int)*lastq_len) return;
    t = (int)lastq[t];
  }
  const int f = t * W;
  const int v = blockIdx.y;
  if (f >= S) return;
  uint8_t st[W];
  ld_b<W>(st, s.state + f);
  int it[W];
  ld_i<W>(it, s.iter + f);
  bool any = false, anylast = false, anyck = false, last[W], ck[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const bool run = running(st[w]);
    last[w] = run && it[w] == last_it;              // DecoderCPU.h:284
    ck[w] = run && (last[w] || it[w] % 10 == 0);    // DecoderCPU.h:287
    any |= run;
    anylast |= last[w];
    anyck |= ck[w];
  }
  if (!any || anylast != LAST) return;
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
  uint8_t bit[W];
#pragma unroll
  for (int w = 0; w < W; ++w) { anybad[w] = anynan[w] = false; bit[w] = 0; }
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
  if (!anyck) return;
  // A slot can only stop at a checkpoint, so this is where its decision is recorded (lanes that are not at one get
  // a value nobody reads).
  st_b<W>(s.decb + (size_t)v * S + f, bit);
#pragma unroll
  for (int w = 0; w < W; ++w)
    if (ck[w]) {
      if (anybad[w]) s.bad[f + w] = 1;
      if (anynan[w]) s.nanflag[f + w] = 1;
    }
}

// BeliefPropogation loop control (DecoderCPU.h:280-291), per slot, after the variable phase.
__global__ void __launch_bounds__(kGroup) g_control(Slots s, int S, int last_it, int wv, int parity) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f == 0) s.ctr[2 + (parity ^ 1)] = 0;  // the list the next pass's g_control appends to; its reader has run
  bool next_is_last = false;
  if (f < S && running(s.state[f])) {
    const int it = s.iter[f];
    const bool last = it == last_it, ck = last || it % 10 == 0;
    if (last || (ck && !s.bad[f])) {
      s.state[f] = kDone;  // bad / nanflag are kept: CONVERGENCE_FAIL and the NaN bit describe the final state
    } else {
      s.state[f] = kRun;
      s.iter[f] = it + 1;
      s.bad[f] = s.nanflag[f] = 0;
      next_is_last = it + 1 == last_it;
    }
  }
  // one list entry per variable-kernel thread (wv consecutive slots): the lowest flagged slot of the group appends
  const unsigned flagged = __ballot_sync(0xffffffffu, next_is_last);
  const int lane = threadIdx.x & 31, first = lane & ~(wv - 1);
  const unsigned group = (flagged >> first) & ((1u << wv) - 1u);
  if (next_is_last && (group & ((1u << (lane - first)) - 1u)) == 0)
    s.lastq[(size_t)parity * S + atomicAdd(&s.ctr[2 + parity], 1u)] = (uint32_t)(f / wv);
}

// syndrome of the decision against the input syndrome (DecoderCPU.h:380-384), for the slots that just stopped
__global__ void __launch_bounds__(128) g_verify(Slots s, const uint32_t* __restrict__ cvar, int m, int dc, int S) {
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (f >= S) return;
  uint8_t st[4];
  ld_b<4>(st, s.state + f);
  if (st[0] != kDone && st[1] != kDone && st[2] != kDone && st[3] != kDone) return;
  unsigned any = 0;
  for (int e = blockIdx.y; e < m; e += gridDim.y) {
    unsigned par = *reinterpret_cast<const uint32_t*>(s.synb + (size_t)e * S + f);  // 4 slots, one byte each
    for (int i = 0; i < dc; ++i)
      par ^= *reinterpret_cast<const uint32_t*>(s.decb + (size_t)cvar[(size_t)i * m + e] * S + f);
    any |= par;
  }
#pragma unroll
  for (int w = 0; w < 4; ++w)
    if (st[w] == kDone && ((any >> (8 * w)) & 1u)) s.mismatch[f + w] = 1;
}

// decision bytes -> bit-packed words of the frame, for the slots that just stopped
__global__ void __launch_bounds__(128) g_pack(Slots s, int n, int S, int nw, uint32_t* __restrict__ dec) {
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (f >= S) return;
  uint8_t st[4];
  ld_b<4>(st, s.state + f);
  int fr[4];
  ld_i<4>(fr, s.frame + f);
  bool any = false;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    if (st[w] != kDone) fr[w] = -1;
    any |= fr[w] >= 0;
  }
  if (!any) return;
  for (int word = blockIdx.y; word < nw; word += gridDim.y) {
    uint32_t out[4] = {0, 0, 0, 0};
    for (int b = 0; b < 32; ++b) {
      const int v = word * 32 + b;
      if (v >= n) break;
      const uint32_t d = *reinterpret_cast<const uint32_t*>(s.decb + (size_t)v * S + f);
#pragma unroll
      for (int w = 0; w < 4; ++w) out[w] |= ((d >> (8 * w)) & 1u) << b;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w)
      if (fr[w] >= 0) dec[(size_t)fr[w] * nw + word] = out[w];
  }
}

// per-frame outputs of the stopped slots, then the hand-over to the next frame of the queue
__global__ void __launch_bounds__(kGroup) g_handover(Slots s, int S, int nframes, uint8_t* __restrict__ flags,
                                                     uint32_t* __restrict__ iters) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= S || s.state[f] != kDone) return;
  const int fr = s.frame[f];
  if (fr >= 0) {
    // CONVERGENCE_FAIL = !CheckConvergence(final messages), DecoderCPU.h:375-378
    flags[fr] = (uint8_t)((s.mismatch[f] & 1u) | ((s.bad[f] & 1u) << 1) | ((s.nanflag[f] & 1u) << 2));
    iters[fr] = (uint32_t)(s.iter[f] + 1);
    atomicAdd(&s.ctr[1], 1u);
  }
  const unsigned next = atomicAdd(&s.ctr[0], 1u);
  s.bad[f] = s.nanflag[f] = s.mismatch[f] = 0;
  s.iter[f] = 0;
  if (next < (unsigned)nframes) {
    s.frame[f] = (int)next;
    s.state[f] = kFresh;
  } else {
    s.frame[f] = -1;
    s.state[f] = kIdle;
  }
}

// syndrome bits of the frames just handed out -> one byte per (check, slot).  Runs right after g_handover, when
// kFresh marks exactly the new arrivals (g_control turns kFresh into kRun after their first iteration).
__global__ void __launch_bounds__(128) g_fill(Slots s, const uint32_t* __restrict__ syn, int mw, int m, int S) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= S || s.state[f] != kFresh) return;
  const uint32_t* row = syn + (size_t)s.frame[f] * mw;
  for (int e = blockIdx.y; e < m; e += gridDim.y) s.synb[(size_t)e * S + f] = (uint8_t)((row[e >> 5] >> (e & 31)) & 1u);
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
void pick_var(int dv, VarFn& fn, VarFn& fn_last, int& w) {
#define QLDPC_E(N) \
  if (dv == N) { fn = g_var<N, var_width(N), true, false>; fn_last = g_var<N, var_width(N), true, true>; w = var_width(N); return; }
#define QLDPC_B(N) \
  if (dv <= N) { fn = g_var<N, var_width(N), false, false>; fn_last = g_var<N, var_width(N), false, true>; w = var_width(N); return; }
  QLDPC_B(1) QLDPC_E(2) QLDPC_E(3) QLDPC_E(4) QLDPC_E(5) QLDPC_E(6) QLDPC_E(7) QLDPC_E(8) QLDPC_E(9) QLDPC_E(10)
  QLDPC_E(11) QLDPC_E(12) QLDPC_B(16) QLDPC_B(32)
#undef QLDPC_E
#undef QLDPC_B
}

cudaError_t run(const GlobalBpArgs& a, const uint32_t* syn, uint32_t* dec, uint8_t* flags, uint32_t* iters, int nframes,
                cudaStream_t st) {
  const int m = a.m, n = a.n, E = a.m * a.dc;
  // Slots in flight: enough that one pass moves ~1 GB (launch overhead out of sight), about a quarter of the frames
  // so that every slot is refilled a few times and the straggler tail stays short, at most what was allocated.
  const long long want = std::max<long long>(((long long)nframes + 3) / 4, (long long)(1.0e9 / (16.0 * E)));
  const long long asked = a.slots > 0 ? ((long long)a.slots + 31) / 32 * 32 : (want + 127) / 128 * 128;
  const int S = (int)std::max<long long>(32, std::min<long long>(a.batch, asked));
  Slots s;
  s.msg = a.msg;
  s.synb = a.bytes;
  s.state = s.synb + (size_t)m * S;
  s.bad = s.state + S;
  s.nanflag = s.bad + S;
  s.mismatch = s.nanflag + S;
  s.decb = s.mismatch + S;
  s.frame = (int32_t*)a.words;
  s.iter = s.frame + S;
  s.ctr = (unsigned int*)(s.iter + S);
  s.lastq = s.ctr + 4;
  CheckFn check = nullptr;
  VarFn var = nullptr, var_last = nullptr;
  int wc = 1, wv = 1;
  pick_check(a.dc, check, wc);
  pick_var(a.dv, var, var_last, wv);
  const int sb = (S + kGroup - 1) / kGroup, s1 = (S + 127) / 128, s4 = (S + 511) / 512;
  const dim3 gc((S + 128 * wc - 1) / (128 * wc), m), gv((S + 128 * wv - 1) / (128 * wv), n);
  const dim3 ge(s4, std::min(m, 32)), gp(s4, std::min(a.nw, 32)), gf(s1, std::min(m, 16));
  const int last_it = a.maxit - 1;
  g_start<<<sb, kGroup, 0, st>>>(s, S);
  // A pass first serves the slots that stopped in the previous pass (before the first pass: all of them, with no
  // frame to write out), then runs one BP iteration on every running slot.
  for (long long pass = 0;; ++pass) {
    if (pass > 0) {
      g_verify<<<ge, 128, 0, st>>>(s, a.cvar, m, a.dc, S);
      g_pack<<<gp, 128, 0, st>>>(s, n, S, a.nw, dec);
    }
    g_handover<<<sb, kGroup, 0, st>>>(s, S, nframes, flags, iters);
    g_fill<<<gf, 128, 0, st>>>(s, syn, a.mw, m, S);
    if (pass % 8 == 0) {  // completion is polled every few passes (a device-to-host copy and a stream sync)
      unsigned int completed = 0;
      cudaError_t e = cudaMemcpyAsync(&completed, s.ctr + 1, sizeof completed, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return e;
      if (completed >= (unsigned)nframes) break;
    }
    check<<<gc, 128, 0, st>>>(s, m, a.dc, S, a.prior);
    const int parity = (int)(pass & 1);
    var<<<gv, 128, 0, st>>>(s, a.vrow, n, a.dv, S, a.prior, last_it, nullptr, nullptr);
    if (last_it == 0)
      var_last<<<gv, 128, 0, st>>>(s, a.vrow, n, a.dv, S, a.prior, last_it, nullptr, nullptr);
    else if (pass >= last_it)  // no slot is that old before; the list was written by the previous pass's g_control
      var_last<<<gv, 128, 0, st>>>(s, a.vrow, n, a.dv, S, a.prior, last_it, s.lastq + (size_t)(parity ^ 1) * S,
                                   s.ctr + 2 + (parity ^ 1));
    g_control<<<sb, kGroup, 0, st>>>(s, S, last_it, wv, parity);
  }
  return cudaGetLastError();
}

}  // namespace

size_t global_bp_bytes(int m, int n, int dc, int batch, size_t* msg_bytes, size_t* byte_bytes, size_t* word_bytes) {
  *msg_bytes = (size_t)m * dc * batch * sizeof(float);
  *byte_bytes = ((size_t)m + 4 + n) * batch;
  *word_bytes = ((size_t)4 * batch + 4) * sizeof(uint32_t);
  return *msg_bytes + *byte_bytes + *word_bytes;
}

cudaError_t global_bp_run(const GlobalBpArgs& a, const uint32_t* syn, uint32_t* dec, uint8_t* flags, uint32_t* iters,
                          int nframes, int* launches, cudaStream_t st) {
  const int maxd = std::max(a.dc, a.dv);
  if (maxd > 32) return cudaErrorInvalidValue;
  if (nframes <= 0) return cudaSuccess;
  if (launches) *launches += 1;
  return run(a, syn, dec, flags, iters, nframes, st);
}

}  // namespace qldpc
