// Global-memory ("HBM-resident") belief-propagation path: the fallback for codes the shared-memory tile kernel
// (bp_kernel.cuh) does not cover -- a (check degree, variable degree) pair without a compiled instantiation, or a
// frame whose messages do not fit in shared memory.  Same arithmetic, operation for operation, as the tile kernel
// and the reference (EqNodeUpdate / VarNodeUpdate / CheckConvergence / Decode tail, DecoderCPU.h:150-390); only
// the residence of the message state differs.
//
// Layout: msg[row][S] with row = i*m + e (i = position of the edge in its check) and the SLOT index innermost, so a
// warp that processes one node for 32 consecutive slots reads and writes fully coalesced 128-byte lines: the path is
// bound by HBM at 16 bytes per edge-update.  S frame slots are decoded concurrently.  Like the tile
// kernel, every slot has its own iteration counter, n % 10 convergence cadence and `last` iteration; a slot that stops
// is hard-decided, syndrome-checked, written out and refilled from a frame queue in the same pass, so slots never idle
// while frames remain (no lock-step batches).  One pass = check kernel, variable kernel, and four small kernels that
// only touch the slots which just stopped.  Degrees are runtime values up to 32; the kernels are
// instantiated for a ladder of degree bounds and W = 4, 2 or 1 slots per thread.
#include <algorithm>
#include <cstdint>

#include "kernels.cuh"

namespace qldpc {

namespace {

enum : uint8_t { kIdle = 0, kRun = 1, kDone = 2, kFresh = 3 };

struct Slots {
  float* msg;          // [E][S]
  uint8_t* synb;       // [m][S] input syndrome bit of the frame in the slot
  uint8_t* state;      // [S] kIdle / kRun / kDone / kFresh
  uint8_t* bad;        // [S] an unconverged message was seen in the last variable phase
  uint8_t* nanflag;    // [S]
  uint8_t* mismatch;   // [S]
  int32_t* frame;      // [S] frame id in the slot
  int32_t* iter;       // [S] iteration index n of the slot
  unsigned int* ctr;   // [0] next frame to hand out, [1] frames completed
};

// Slots start as kDone-less kFresh candidates: every slot asks the queue for its first frame.
__global__ void __launch_bounds__(256) g_start(Slots s, int S) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < S) {
    s.state[t] = kDone;   // "stopped" with no frame to write out: g_handover gives it its first frame
    s.frame[t] = -1;
    s.bad[t] = s.nanflag[t] = s.mismatch[t] = 0;
  }
  if (t < 2) s.ctr[t] = 0;
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

// A thread whose W slots are not all running still computes and stores all W lanes: the other lanes are idle slots
// (stopped and fresh slots only exist between g_control and g_activate), whose messages nobody reads.
template <int MAXC, int W, bool EXACT>
__global__ void __launch_bounds__(128) g_check(Slots s, int m, int dc_rt, int S) {
  const int dc = EXACT ? MAXC : dc_rt;
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  const int e = blockIdx.y;
  if (f >= S) return;
  uint8_t st[W];
  ld_b<W>(st, s.state + f);
  bool any = false;
#pragma unroll
  for (int w = 0; w < W; ++w) any |= st[w] == kRun;
  if (!any) return;
  float t[MAXC][W];
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (i < dc) {
      ld_f<W>(t[i], s.msg + ((size_t)i * m + e) * S + f);
#pragma unroll
      for (int w = 0; w < W; ++w) t[i][w] = __fmaf_rn(-2.0f, t[i][w], 1.0f);  // 1 - 2q (DecoderCPU.h:175)
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

template <int MAXV, int W, bool EXACT>
__global__ void __launch_bounds__(128) g_var(Slots s, const uint32_t* __restrict__ vrow, int n, int dv_rt, int S,
                                             float prior, int last_it) {
  const int dv = EXACT ? MAXV : dv_rt;
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  const int v = blockIdx.y;
  if (f >= S) return;
  uint8_t st[W];
  ld_b<W>(st, s.state + f);
  bool any = false;
#pragma unroll
  for (int w = 0; w < W; ++w) any |= st[w] == kRun;
  if (!any) return;
  int it[W];
  ld_i<W>(it, s.iter + f);
  bool last[W], anylast = false;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    last[w] = it[w] == last_it;  // DecoderCPU.h:284,287
    anylast |= last[w];
  }
  float pk[MAXV][W], om[MAXV][W];
  uint32_t row[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
    if (k < dv) {
      row[k] = vrow[(size_t)k * n + v];
      ld_f<W>(pk[k], s.msg + (size_t)row[k] * S + f);
#pragma unroll
      for (int w = 0; w < W; ++w) om[k][w] = __fsub_rn(1.0f, pk[k][w]);
    }
  const float prior1 = __fsub_rn(1.0f, prior);  // DecoderCPU.h:209-210
  float preP[W], preQ[W], fullP[W], fullQ[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    preP[w] = fullP[w] = prior;
    preQ[w] = fullQ[w] = prior1;
  }
  if (anylast) {
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
      if (k < dv) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
          fullQ[w] = __fmul_rn(fullQ[w], om[k][w]);
          fullP[w] = __fmul_rn(fullP[w], pk[k][w]);
        }
      }
  }
  bool anybad[W];
#pragma unroll
  for (int w = 0; w < W; ++w) anybad[w] = false;
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
            Q[w] = __fmul_rn(Q[w], om[k][w]);
            P[w] = __fmul_rn(P[w], pk[k][w]);
          }
        }
      float q[W];
#pragma unroll
      for (int w = 0; w < W; ++w) {
        if (last[w]) { P[w] = fullP[w]; Q[w] = fullQ[w]; }
        const float den = __fadd_rn(Q[w], P[w]);
        bool unsafe = false;
        q[w] = div_fast<3>(P[w], den, unsafe);  // == P / (Q + P), DecoderCPU.h:223
        if (unsafe) q[w] = __fdiv_rn(P[w], den);
        anybad[w] |= unconverged(q[w]);
        preQ[w] = __fmul_rn(preQ[w], om[j][w]);
        preP[w] = __fmul_rn(preP[w], pk[j][w]);
      }
      st_f<W>(s.msg + (size_t)row[j] * S + f, q);
    }
  }
#pragma unroll
  for (int w = 0; w < W; ++w)
    if (st[w] == kRun && (last[w] || it[w] % 10 == 0) && anybad[w]) s.bad[f + w] = 1;
}

// BeliefPropogation loop control (DecoderCPU.h:280-291), per slot, after the variable phase.
__global__ void __launch_bounds__(256) g_control(Slots s, int S, int last_it) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= S || s.state[f] != kRun) return;
  const int it = s.iter[f];
  const bool last = it == last_it, ck = last || it % 10 == 0;
  if (last || (ck && !s.bad[f])) s.state[f] = kDone;  // bad[f] is kept: CONVERGENCE_FAIL of the final state
  else {
    s.iter[f] = it + 1;
    s.bad[f] = 0;
  }
}

// The three kernels below only work for slots that stopped in this pass; blockIdx.y strides over the nodes so that a
// warp without such a slot retires after one byte load.

// hard decision (any edge message >= 0.5f, DecoderCPU.h:354-373) + NaN flag
__global__ void __launch_bounds__(128) g_decide(Slots s, const uint32_t* __restrict__ vrow, int n, int dv, int S, int nw,
                                                uint32_t* __restrict__ dec) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= S || s.state[f] != kDone || s.frame[f] < 0) return;
  uint32_t* d = dec + (size_t)s.frame[f] * nw;
  bool nan = false;
  for (int v = blockIdx.y; v < n; v += gridDim.y) {
    bool bit = false;
    for (int k = 0; k < dv; ++k) {
      const float x = s.msg[(size_t)vrow[(size_t)k * n + v] * S + f];
      bit |= x >= 0.5f;
      nan |= x != x;
    }
    if (bit) atomicOr(&d[v >> 5], 1u << (v & 31));
  }
  if (nan) s.nanflag[f] = 1;
}

// syndrome of the decision against the input syndrome (DecoderCPU.h:380-384)
__global__ void __launch_bounds__(128) g_verify(Slots s, const uint32_t* __restrict__ cvar, int m, int dc, int S, int nw,
                                                const uint32_t* __restrict__ dec) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= S || s.state[f] != kDone || s.frame[f] < 0) return;
  const uint32_t* d = dec + (size_t)s.frame[f] * nw;
  unsigned any = 0;
  for (int e = blockIdx.y; e < m; e += gridDim.y) {
    unsigned par = s.synb[(size_t)e * S + f];
    for (int i = 0; i < dc; ++i) {
      const uint32_t v = cvar[(size_t)i * m + e];
      par ^= (d[v >> 5] >> (v & 31)) & 1u;
    }
    any |= par;
  }
  if (any) s.mismatch[f] = 1;
}

// per-frame outputs of the stopped slots, then the hand-over to the next frame of the queue
__global__ void __launch_bounds__(256) g_handover(Slots s, int S, int nframes, uint8_t* __restrict__ flags,
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

// fresh slots: messages <- prior (InitVarNodes, DecoderCPU.h:135-148), syndrome bits -> bytes
__global__ void __launch_bounds__(128) g_fill(Slots s, const uint32_t* __restrict__ syn, int mw, int m, int E, int S,
                                              float prior) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= S || s.state[f] != kFresh) return;
  const int fr = s.frame[f];
  for (int r = blockIdx.y; r < E; r += gridDim.y) {
    s.msg[(size_t)r * S + f] = prior;
    if (r < m) s.synb[(size_t)r * S + f] = (uint8_t)((syn[(size_t)fr * mw + (r >> 5)] >> (r & 31)) & 1u);
  }
}

__global__ void __launch_bounds__(256) g_activate(Slots s, int S) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < S && s.state[f] == kFresh) s.state[f] = kRun;
}

using CheckFn = void (*)(Slots, int, int, int);
using VarFn = void (*)(Slots, const uint32_t*, int, int, int, float, int);
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
void pick_var(int dv, VarFn& fn, int& w) {
#define QLDPC_E(N) if (dv == N) { fn = g_var<N, var_width(N), true>; w = var_width(N); return; }
#define QLDPC_B(N) if (dv <= N) { fn = g_var<N, var_width(N), false>; w = var_width(N); return; }
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
  s.frame = (int32_t*)a.words;
  s.iter = s.frame + S;
  s.ctr = (unsigned int*)(s.iter + S);
  cudaMemsetAsync(dec, 0, (size_t)nframes * a.nw * sizeof(uint32_t), st);
  const int sb = (S + 255) / 256, sx = (S + 127) / 128;
  g_start<<<sb, 256, 0, st>>>(s, S);
  CheckFn check = nullptr;
  VarFn var = nullptr;
  int wc = 1, wv = 1;
  pick_check(a.dc, check, wc);
  pick_var(a.dv, var, wv);
  const dim3 gc((S + 128 * wc - 1) / (128 * wc), m), gv((S + 128 * wv - 1) / (128 * wv), n), gd(sx, std::min(n, 32)), ge(sx, std::min(m, 32)), gf(sx, std::min(E, 64));
  const int last_it = a.maxit - 1;
  // A pass first hands frames to the slots that stopped in the previous pass (all of them before the first), then
  // runs one BP iteration on every running slot.
  for (long long pass = 0;; ++pass) {
    g_decide<<<gd, 128, 0, st>>>(s, a.vrow, n, a.dv, S, a.nw, dec);
    g_verify<<<ge, 128, 0, st>>>(s, a.cvar, m, a.dc, S, a.nw, dec);
    g_handover<<<sb, 256, 0, st>>>(s, S, nframes, flags, iters);
    g_fill<<<gf, 128, 0, st>>>(s, syn, a.mw, m, E, S, a.prior);
    g_activate<<<sb, 256, 0, st>>>(s, S);
    if (pass % 8 == 0) {  // completion is polled every few passes (a device-to-host copy and a stream sync)
      unsigned int completed = 0;
      cudaError_t e = cudaMemcpyAsync(&completed, s.ctr + 1, sizeof completed, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return e;
      if (completed >= (unsigned)nframes) break;
    }
    check<<<gc, 128, 0, st>>>(s, m, a.dc, S);
    var<<<gv, 128, 0, st>>>(s, a.vrow, n, a.dv, S, a.prior, last_it);
    g_control<<<sb, 256, 0, st>>>(s, S, last_it);
  }
  return cudaGetLastError();
}

}  // namespace

size_t global_bp_bytes(int m, int dc, int batch, size_t* msg_bytes, size_t* byte_bytes, size_t* word_bytes) {
  *msg_bytes = (size_t)m * dc * batch * sizeof(float);
  *byte_bytes = (size_t)m * batch + (size_t)4 * batch;
  *word_bytes = ((size_t)2 * batch + 4) * sizeof(uint32_t);
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
