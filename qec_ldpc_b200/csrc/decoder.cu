// Host side of the decoder and the C ABI (include/qldpc_b200.h).  Everything that decodes runs on the GPU;
// without a usable device the calls fail with QLDPC_ERR_NO_DEVICE -- there is no CPU path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/qldpc_b200.h"
#include "code.h"
#include <memory>

#include "host_pack.h"
#include "kernels.cuh"

using namespace qldpc;

struct qldpc_code {
  Code* c;
};

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(expr)                                                                                  \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      cudaGetLastError();                                                                             \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? QLDPC_ERR_NO_DEVICE  \
                                                                               : QLDPC_ERR_CUDA,      \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));                                \
    }                                                                                                 \
  } while (0)

struct DevSide {
  int m = 0, dc = 0, dv = 0, E = 0, mw = 0;
  uint16_t* vrow = nullptr;  // [dv][n]
  uint16_t* vchk = nullptr;  // [dv][n] check index of the k-th edge of variable v (CSC)
  // [dv][L] circulant exponents of a quasi-cyclic side: the BP kernel generates its row table from them on the device
  // (the syndrome kernels keep the CSC table: two instructions per index against eight for the closed form)
  int32_t* hexp = nullptr;
  int P = 0, L = 0;
  BpLaunch cfg, user;        // resolved configuration / user overrides
  bool cfg_ok = false;
  std::string cfg_err;
  // global-memory fallback path (bp_global.cu): shapes without a tile-kernel instantiation, frames too large for
  // shared memory, or forced with frames_per_tile = -1
  bool use_global = false, force_global = false;
  uint32_t* gvrow = nullptr;  // [dv][n] row i*m+e
  uint32_t* gcvar = nullptr;  // [dc][m] variable of the i-th edge of check e
  float* gmsg = nullptr;
  uint8_t* gbytes = nullptr;
  uint32_t* gwords = nullptr;
  unsigned int* ghost_done = nullptr;  // mapped pinned completion counter of the HBM-resident path
  int gbatch = 0;
};

template <typename T>
cudaError_t dev_alloc(T*& p, size_t count) {
  return cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
}

}  // namespace

struct qldpc_decoder {
  Code code;
  int device = 0, num_sms = 0, chunk = 0;
  int n = 0, nw = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  DevSide s[2];
  uint32_t *errX = nullptr, *errZ = nullptr, *synX = nullptr, *synZ = nullptr, *decX = nullptr, *decZ = nullptr;
  uint8_t *sfX = nullptr, *sfZ = nullptr, *fflags = nullptr;
  uint32_t *itX = nullptr, *itZ = nullptr;
  unsigned long long* counters = nullptr;
  unsigned int* queues = nullptr;  // [2]
  uint32_t *lx = nullptr, *lz = nullptr, *lm = nullptr;
  int lx_rows = 0, lz_rows = 0, lm_rows = 0;
  void* stage = nullptr;  // device staging for one-element-per-bit I/O
  size_t stage_bytes = 0;
  // host-buffer entry points run as a pipeline: H2D of the next slice and D2H of the previous one overlap the decode
  cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
  // the Z-side BP launch of a small batch or pipeline slice runs beside the X side on its own stream, so that the
  // straggler tail of one kernel overlaps the start of the other
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap_sides = false;  // set by the sliced / small-batch entry points for the duration of the call
  uint8_t* small_pin = nullptr;  // pinned staging of the low-latency Decode path
  size_t small_pin_bytes = 0;
  // the low-latency path replays a captured CUDA graph when the call repeats the previous one's shape (a per-frame
  // Decode loop): one graph launch instead of nine stream operations
  struct SmallGraph {
    cudaGraphExec_t exec = nullptr;
    int nf = 0, maxit = 0;
    uint32_t p_bits = 0;
    const void *stage = nullptr, *pin = nullptr;
    int cfg_epoch = -1;
  } small_graph;
  int cfg_epoch = 0;  // bumped by every qldpc_decoder_configure
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_out_free[2] = {nullptr, nullptr};
  uint32_t* pin = nullptr;  // pinned host staging: weight-W generator, host-packed rows
  size_t pin_words = 0;
  // host-buffer entry points pack one-element-per-bit rows on the host before they cross the link (host_pack.h);
  // host_threads: -1 = default_host_threads(), 0 = off (raw rows are copied and packed on the device)
  int host_threads = -1;
  std::unique_ptr<qldpc::HostPacker> packer;
  // measurement: launches per kernel class, and (when enabled) CUDA-event pairs on the launching stream
  bool timing = false;
  uint64_t launches[QLDPC_NUM_TIMERS] = {0, 0, 0, 0, 0, 0};
  double ms[QLDPC_NUM_TIMERS] = {0, 0, 0, 0, 0, 0};
  struct Pending { int cls; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;

  ~qldpc_decoder() {
    cudaSetDevice(device);
    for (int i = 0; i < 2; ++i) {
      cudaFree(s[i].vrow); cudaFree(s[i].vchk); cudaFree(s[i].hexp); cudaFree(s[i].gvrow); cudaFree(s[i].gcvar);
      cudaFree(s[i].gmsg); cudaFree(s[i].gbytes); cudaFree(s[i].gwords);
      if (s[i].ghost_done) cudaFreeHost(s[i].ghost_done);
    }
    cudaFree(errX); cudaFree(errZ); cudaFree(synX); cudaFree(synZ); cudaFree(decX); cudaFree(decZ);
    cudaFree(sfX); cudaFree(sfZ); cudaFree(fflags); cudaFree(itX); cudaFree(itZ);
    cudaFree(counters); cudaFree(queues); cudaFree(lx); cudaFree(lz); cudaFree(lm); cudaFree(stage);
    if (pin) cudaFreeHost(pin);
    if (small_graph.exec) cudaGraphExecDestroy(small_graph.exec);
    if (small_pin) cudaFreeHost(small_pin);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side_stream) cudaStreamDestroy(side_stream);
    for (int i = 0; i < 2; ++i) {
      if (ev_ready[i]) cudaEventDestroy(ev_ready[i]);
      if (ev_free[i]) cudaEventDestroy(ev_free[i]);
      if (ev_done[i]) cudaEventDestroy(ev_done[i]);
      if (ev_out_free[i]) cudaEventDestroy(ev_out_free[i]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (d2h_stream) cudaStreamDestroy(d2h_stream);
    for (auto& p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : pool) cudaEventDestroy(e);
    if (own_stream) cudaStreamDestroy(own_stream);
  }
};

namespace {

// Counts a launch of kernel class `cls`; with timing enabled brackets it with events on the launch stream.
struct Timed {
  qldpc_decoder* d;
  int cls;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  Timed(qldpc_decoder* d_, int cls_, int n = 1, cudaStream_t st_ = nullptr) : d(d_), cls(cls_), st(st_ ? st_ : d_->stream) {
    d->launches[cls] += (uint64_t)n;
    if (!d->timing) return;
    auto get = [&]() {
      cudaEvent_t e = nullptr;
      if (!d->pool.empty()) { e = d->pool.back(); d->pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, st);
  }
  ~Timed() {
    if (!a) return;
    cudaEventRecord(b, st);
    d->pending.push_back({cls, a, b});
  }
};

void drain_timing(qldpc_decoder* d) {
  for (auto& p : d->pending) {
    float t = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) d->ms[p.cls] += t;
    d->pool.push_back(p.a);
    d->pool.push_back(p.b);
  }
  d->pending.clear();
  cudaGetLastError();
}

// Largest slice of the host-buffer pipelines (frames).  The first slice cannot overlap anything (the host packs it
// while the device idles), so the slices start small and grow by a quarter per slice up to this size (slice_frames):
// packing slice i+1 then takes no longer than decoding slice i (host packing runs at 0.7-0.9 of the decode rate), so
// the device idles only for the first, small slice.
const int kPipeFrames = 1 << 17;
const int kFirstSlice = 1 << 13;
// batches up to this many frames take the low-latency Decode path (one stream, no host threads, no pipeline)
const int kSmallBatch = 2048;

int64_t slice_frames(int index, int64_t max_slice) {
  int64_t s = kFirstSlice;
  for (int i = 0; i < index && s < max_slice; ++i) s = (s + s / 4 + 1023) / 1024 * 1024;
  return std::min<int64_t>(s, max_slice);
}

struct OverlapScope {  // X and Z side by side for the calls made while one of these is alive
  qldpc_decoder* d;
  explicit OverlapScope(qldpc_decoder* d_) : d(d_) { d->overlap_sides = true; }
  ~OverlapScope() { d->overlap_sides = false; }
};

// Waits for everything a pipelined call may still have in flight (error exits: the caller's buffers and the pinned
// staging must not be touched by asynchronous copies after the call has returned).
void quiesce(qldpc_decoder* d) {
  if (d->stream) cudaStreamSynchronize(d->stream);
  if (d->copy_stream) cudaStreamSynchronize(d->copy_stream);
  if (d->d2h_stream) cudaStreamSynchronize(d->d2h_stream);
  if (d->side_stream) cudaStreamSynchronize(d->side_stream);
  cudaGetLastError();
}

int ensure_side_stream(qldpc_decoder* d) {
  if (d->side_stream) return QLDPC_OK;
  CU_TRY(cudaStreamCreateWithFlags(&d->side_stream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
  return QLDPC_OK;
}

int ensure_pipeline(qldpc_decoder* d) {
  if (d->copy_stream) return QLDPC_OK;
  CU_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&d->d2h_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CU_TRY(cudaEventCreateWithFlags(&d->ev_ready[i], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&d->ev_free[i], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&d->ev_done[i], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&d->ev_out_free[i], cudaEventDisableTiming));
  }
  return QLDPC_OK;
}

int ensure_pin(qldpc_decoder* d, size_t words) {
  if (words <= d->pin_words) return QLDPC_OK;
  if (d->pin) cudaFreeHost(d->pin);
  d->pin = nullptr;
  d->pin_words = 0;
  CU_TRY(cudaMallocHost((void**)&d->pin, words * sizeof(uint32_t)));
  d->pin_words = words;
  return QLDPC_OK;
}

// Host threads the host-buffer entry points use for rows of `elem`-byte elements (0 = raw rows over the link, packed on
// the device).  An explicit qldpc_decoder_set_host_threads setting is taken as given.  By default the rows are packed
// on the host when this process has enough cores to itself to outrun the raw copy:
//  * int32 rows (4 bytes per bit): the raw copy is bound by the GPU's own host link (about 10.5 M frames/s per GPU for
//    n = 610) and scales with the number of GPUs until host DRAM saturates, whereas packing is bound by what the
//    host's cores can stream in total (140-190 GB/s on the 32-core B200 hosts, i.e. 29-39 M frames/s per BOX).  With
//    one or two ranks per box (>= 10 threads each) packing wins; from four ranks on the raw copy does.
//  * byte rows: the link carries them at the decode rate; packing only helps with >= 8 threads.
// Measured: profiles/r2/host_pack_e2e.jsonl, host_pack_ranks.jsonl, bench_C2_{1,2,4,8}gpu.json (both series in every line).
int host_threads_for(const qldpc_decoder* d, int elem) {
  if (d->host_threads >= 0) return d->host_threads;
  const int want = default_host_threads();
  return want >= (elem == 4 ? 10 : 8) ? want : 0;
}

// The host packer, or null when host-side packing is switched off.
HostPacker* host_packer(qldpc_decoder* d, int elem) {
  const int want = host_threads_for(d, elem);
  if (want <= 0) return nullptr;
  if (!d->packer || d->packer->threads() != want) d->packer.reset(new HostPacker(want));
  return d->packer.get();
}

int ensure_stage(qldpc_decoder* d, size_t bytes) {
  if (bytes <= d->stage_bytes) return QLDPC_OK;
  cudaFree(d->stage);
  d->stage = nullptr;
  d->stage_bytes = 0;
  CU_TRY(cudaMalloc(&d->stage, bytes));
  d->stage_bytes = bytes;
  return QLDPC_OK;
}

// logical rows [r0, r0+rows): x-part words then z-part words (or the z-part alone), transposed and row-padded LT[w][rows_pad]
int upload_logical(const BitMatrix& L, int r0, int rows, int words, int n, bool z_part, uint32_t*& dst) {
  dst = nullptr;
  if (rows == 0) return QLDPC_OK;
  const int rows_pad = (rows + 31) & ~31;
  std::vector<uint32_t> t((size_t)words * rows_pad, 0u);
  const int nw = (n + 31) / 32;
  for (int r = 0; r < rows; ++r)
    for (int col = 0; col < 2 * n; ++col)
      if (L.get(r0 + r, col)) {
        int w, b;
        if (col < n) {
          if (z_part) continue;
          w = col >> 5; b = col & 31;
        } else {
          const int zc = col - n;
          w = (z_part ? 0 : nw) + (zc >> 5); b = zc & 31;
        }
        if (w < words) t[(size_t)w * rows_pad + r] |= 1u << b;
      }
  CU_TRY(dev_alloc(dst, t.size()));
  CU_TRY(cudaMemcpy(dst, t.data(), t.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return QLDPC_OK;
}

// Device tables and work buffers of the global-memory path, built on first need.
int ensure_global(qldpc_decoder* d, int side) {
  DevSide& s = d->s[side];
  if (s.gmsg) return QLDPC_OK;
  const SideTables& t = d->code.side[side];
  const int n = d->n;
  if (std::max(s.dc, s.dv) > 32) return fail(QLDPC_ERR_UNSUPPORTED, "node degree above 32");
  // the HBM-resident kernels put the node index on gridDim.y (at most 65535)
  if (n > 65535 || s.m > 65535) return fail(QLDPC_ERR_UNSUPPORTED, "more than 65535 variables or checks per side");
  std::vector<uint32_t> vrow((size_t)t.E), cvar((size_t)t.E);
  for (int v = 0; v < n; ++v)
    for (int k = 0; k < t.dv; ++k) {
      const int edge = t.var_edge[(size_t)v * t.dv + k];
      vrow[(size_t)k * n + v] = (uint32_t)((edge % t.dc) * t.m + edge / t.dc);
    }
  for (int e = 0; e < t.m; ++e)
    for (int i = 0; i < t.dc; ++i) cvar[(size_t)i * t.m + e] = (uint32_t)t.chk_var[(size_t)e * t.dc + i];
  // batch: as many frames as fit in ~2 GB of messages, a multiple of 32, at most the decoder's chunk
  long long batch = (long long)(2.0e9 / ((double)t.E * 4.0)) / 32 * 32;
  batch = std::max<long long>(32, std::min<long long>(batch, ((long long)d->chunk + 31) / 32 * 32));
  s.gbatch = (int)std::min<long long>(batch, 1 << 16);
  size_t mb, bb, wb;
  global_bp_bytes(s.m, n, s.dc, s.gbatch, &mb, &bb, &wb);
  CU_TRY(dev_alloc(s.gvrow, vrow.size()));
  CU_TRY(dev_alloc(s.gcvar, cvar.size()));
  CU_TRY(cudaMemcpy(s.gvrow, vrow.data(), vrow.size() * 4, cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(s.gcvar, cvar.data(), cvar.size() * 4, cudaMemcpyHostToDevice));
  CU_TRY(cudaMalloc((void**)&s.gbytes, bb));
  CU_TRY(cudaMalloc((void**)&s.gwords, wb));
  CU_TRY(cudaHostAlloc((void**)&s.ghost_done, sizeof(unsigned int), cudaHostAllocMapped));
  *s.ghost_done = 0u;
  CU_TRY(cudaMalloc((void**)&s.gmsg, mb));
  return QLDPC_OK;
}

int resolve_config(qldpc_decoder* d, int side) {
  DevSide& s = d->s[side];
  BpLaunch cfg = s.user;
  const char* why = "";
  s.use_global = false;
  s.cfg_ok = !s.force_global && s.E < 65536 && bp_configure(s.dc, s.dv, s.m, d->n, s.hexp ? s.P : 0, d->num_sms, cfg, &why);
  if (!s.cfg_ok) {
    // no tile kernel for this side: the HBM-resident path takes over (still on the GPU; slower, see bp_global.cu)
    if (s.m > 65535) {
      s.cfg_err = "more than 65535 checks per side";
      return fail(QLDPC_ERR_UNSUPPORTED, s.cfg_err);
    }
    int rc = ensure_global(d, side);
    if (rc) {
      s.cfg_err = qldpc_last_error();
      return rc;
    }
    s.use_global = true;
    s.cfg_ok = true;
    s.cfg = BpLaunch();
    s.cfg.vec = -1;
    s.cfg.threads = 128;
    s.cfg.grid = s.gbatch;
    return QLDPC_OK;
  }
  s.cfg = cfg;
  return QLDPC_OK;
}

// Which range tests the kernel's branch-free division needs outside the `last` iteration (bp_kernel.cuh:div_fast).
// A numerator is prior * (dv-1 check-to-variable messages), each 0 or >= 2^-25; a denominator adds
// (1-prior) * (dv-1 complements), each 0 or >= 2^-24.  If the smallest non-zero value already clears the
// threshold the test can never fire and is compiled out (0).  If a numerator could fall below 2^-100 while neither
// chain can leave the normal range (>= 2^-126), the kernel runs both chains scaled by 2^64 instead of testing (1: exact,
// see var_phase).  Otherwise both tests stay in (3).
int division_guard(float prior, int dv) {
  if (!(prior > 0.0f && prior < 1.0f)) return 3;
  const int nf = dv - 1;
  const double xmin = std::ldexp((double)prior, -25 * nf);
  const double ymin = std::min(xmin, std::ldexp(1.0 - (double)prior, -24 * nf));
  int g = 0;
  if (xmin < std::ldexp(1.0, -100)) g |= 1;
  if (ymin < std::ldexp(1.0, -126)) g = 3;
  return g;
}

// BP on both sides over nf frames whose bit-packed syndromes are resident on the device.
int run_bp(qldpc_decoder* d, const uint32_t* synX, const uint32_t* synZ, int nf, float errorProbability, int maxIterations,
           uint32_t* decX, uint32_t* decZ, uint8_t* sfX, uint8_t* sfZ, uint32_t* itX, uint32_t* itZ, int only_side = -1,
           float* trace_q = nullptr, float* trace_r = nullptr, int trace_cap = 0) {
  if (nf <= 0) return QLDPC_OK;
  const float prior = 2.0f / 3.0f * errorProbability;  // DecoderCPU.h:259, same float expression
  CU_TRY(cudaMemsetAsync(d->queues, 0, 2 * sizeof(unsigned int), d->stream));
  // Slices and small batches (the host-buffer entry points): the Z side runs beside the X side on a second stream (its
  // CTAs move in as the X side's drain); the device-resident statistics calls launch back to back -- their launches
  // are large, the tail is negligible there, and the per-kernel timing stays clean.
  const bool overlap = d->overlap_sides && only_side < 0 && !trace_q && !trace_r && !d->s[0].use_global &&
                       !d->s[1].use_global && ensure_side_stream(d) == QLDPC_OK;
  if (overlap) {
    CU_TRY(cudaEventRecord(d->ev_fork, d->stream));
    CU_TRY(cudaStreamWaitEvent(d->side_stream, d->ev_fork, 0));
  }
  auto global_args = [&](const DevSide& s) {
    GlobalBpArgs g;
    g.m = s.m; g.n = d->n; g.dc = s.dc; g.dv = s.dv; g.mw = s.mw; g.nw = d->nw;
    g.maxit = maxIterations; g.batch = s.gbatch; g.prior = prior;
    g.slots = s.force_global ? s.user.threads : 0;
    g.vrow = s.gvrow; g.cvar = s.gcvar; g.msg = s.gmsg; g.bytes = s.gbytes; g.words = s.gwords;
    g.host_done = s.ghost_done;
    g.guard = division_guard(prior, s.dv);
    return g;
  };
  // HBM-resident path on both sides: the two runs are interleaved on two streams (bp_global.cu:global_bp_run_pair); one
  // timing bracket covers both, booked on the X side (the Z side counts its launch and no time of its own).
  static const bool pair_runs = getenv("QLDPC_GLOBAL_SERIAL") == nullptr;
  if (pair_runs && only_side < 0 && !trace_q && !trace_r && d->s[0].use_global && d->s[1].use_global && d->s[0].cfg_ok &&
      d->s[1].cfg_ok && ensure_side_stream(d) == QLDPC_OK) {
    d->launches[QLDPC_T_BP_Z] += 1;
    Timed t(d, QLDPC_T_BP_X, 1, d->stream);
    CU_TRY(cudaEventRecord(d->ev_fork, d->stream));
    CU_TRY(cudaStreamWaitEvent(d->side_stream, d->ev_fork, 0));
    CU_TRY(global_bp_run_pair(global_args(d->s[0]), synX, decX, sfX, itX, d->stream, global_args(d->s[1]), synZ, decZ, sfZ,
                              itZ, d->side_stream, nf));
    CU_TRY(cudaEventRecord(d->ev_join, d->side_stream));
    CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_join, 0));
    return QLDPC_OK;
  }
  for (int side = 0; side < 2; ++side) {
    if (only_side >= 0 && side != only_side) continue;
    DevSide& s = d->s[side];
    if (!s.cfg_ok) return fail(QLDPC_ERR_UNSUPPORTED, s.cfg_err);
    cudaStream_t st = overlap && side == 1 ? d->side_stream : d->stream;
    BpArgs a;
    a.syn = side ? synZ : synX;
    a.dec = side ? decZ : decX;
    a.flags = side ? sfZ : sfX;
    a.iters = side ? itZ : itX;
    a.vrow = s.vrow;
    a.hexp = s.hexp; a.P = s.P; a.L = s.L;
    a.queue = d->queues + side;
    a.m = s.m; a.n = d->n; a.mw = s.mw; a.nw = d->nw;
    a.nframes = nf;
    a.maxit = maxIterations;
    a.prior = prior;
    a.trace_q = trace_q; a.trace_r = trace_r; a.trace_cap = trace_cap;
    Timed t(d, side ? QLDPC_T_BP_Z : QLDPC_T_BP_X, 1, st);
    if (s.use_global) {
      if (trace_q || trace_r) return fail(QLDPC_ERR_UNSUPPORTED, "message taps are not available on the global-memory path");
      CU_TRY(global_bp_run(global_args(s), a.syn, a.dec, a.flags, a.iters, nf, nullptr, d->stream));
      continue;
    }
    CU_TRY(bp_launch(s.dc, s.dv, s.cfg, a, nf, division_guard(prior, s.dv), st));
  }
  if (overlap) {
    CU_TRY(cudaEventRecord(d->ev_join, d->side_stream));
    CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_join, 0));
  }
  return QLDPC_OK;
}

int run_stats(qldpc_decoder* d, int nf, uint8_t* fflags) {
  StatsArgs a;
  a.errX = d->errX; a.errZ = d->errZ; a.decX = d->decX; a.decZ = d->decZ;
  a.sfX = d->sfX; a.sfZ = d->sfZ; a.itX = d->itX; a.itZ = d->itZ;
  a.lx = d->lx; a.lz = d->lz; a.lm = d->lm;
  a.lx_rows = d->lx_rows; a.lz_rows = d->lz_rows; a.lm_rows = d->lm_rows;
  a.nframes = nf; a.nw = d->nw;
  a.counters = d->counters;
  a.fflags = fflags;
  Timed t(d, QLDPC_T_STATS);
  CU_TRY(launch_stats(a, d->stream));
  return QLDPC_OK;
}

int run_syndrome(qldpc_decoder* d, int nf) {
  Timed t(d, QLDPC_T_SYNDROME);
  CU_TRY(launch_syndrome(d->errX, d->errZ, nf, d->n, d->nw, d->s[0].vchk, d->s[0].dv, d->s[0].mw, d->synX, d->s[1].vchk,
                         d->s[1].dv, d->s[1].mw, d->synZ, d->stream));
  return QLDPC_OK;
}

// Philox depolarizing errors of nf frames and their syndromes, one kernel (errors only when with_syndrome is false).
int run_syndrome(qldpc_decoder* d, int nf);

int run_generate(qldpc_decoder* d, uint64_t seed, uint64_t first_frame, int nf, const Thresholds& thr) {
  // the fused kernel lists error positions in 16 bits; longer codes generate and form syndromes in two launches
  const bool fused = d->n <= 65535;
  {
    Timed t(d, QLDPC_T_GENERATE);
    CU_TRY(launch_generate_syndrome(seed, first_frame, nf, d->n, d->nw, thr, d->errX, d->errZ, d->s[0].vchk, d->s[0].dv,
                                    d->s[0].mw, fused ? d->synX : nullptr, d->s[1].vchk, d->s[1].dv, d->s[1].mw, d->synZ,
                                    d->stream));
  }
  return fused ? QLDPC_OK : run_syndrome(d, nf);
}

int check_common(qldpc_decoder* d, int64_t nframes, int maxIterations) {
  if (!d) return fail(QLDPC_ERR_ARG, "null decoder");
  if (nframes < 0) return fail(QLDPC_ERR_ARG, "negative frame count");
  if (maxIterations < 1) return fail(QLDPC_ERR_ARG, "maxIterations must be >= 1");
  CU_TRY(cudaSetDevice(d->device));
  return QLDPC_OK;
}

// After errors + syndromes of one chunk are resident: decode, reduce, copy the optional per-frame outputs.
int finish_chunk(qldpc_decoder* d, int nf, int64_t off, float ep, int maxit, uint8_t* perFrameFlags, uint32_t* perFrameIters) {
  int rc = run_bp(d, d->synX, d->synZ, nf, ep, maxit, d->decX, d->decZ, d->sfX, d->sfZ, d->itX, d->itZ);
  if (rc) return rc;
  rc = run_stats(d, nf, d->fflags);
  if (rc) return rc;
  if (perFrameFlags) CU_TRY(cudaMemcpyAsync(perFrameFlags + off, d->fflags, (size_t)nf, cudaMemcpyDeviceToHost, d->stream));
  if (perFrameIters) {
    std::vector<uint32_t> hx(nf), hz(nf);
    CU_TRY(cudaMemcpyAsync(hx.data(), d->itX, (size_t)nf * 4, cudaMemcpyDeviceToHost, d->stream));
    CU_TRY(cudaMemcpyAsync(hz.data(), d->itZ, (size_t)nf * 4, cudaMemcpyDeviceToHost, d->stream));
    CU_TRY(cudaStreamSynchronize(d->stream));
    for (int f = 0; f < nf; ++f) {
      perFrameIters[2 * (off + f)] = hx[f];
      perFrameIters[2 * (off + f) + 1] = hz[f];
    }
  }
  return QLDPC_OK;
}

int read_counters(qldpc_decoder* d, uint64_t* counters) {
  unsigned long long h[QLDPC_NUM_COUNTERS];
  CU_TRY(cudaMemcpyAsync(h, d->counters, sizeof h, cudaMemcpyDeviceToHost, d->stream));
  CU_TRY(cudaStreamSynchronize(d->stream));
  if (counters)
    for (int i = 0; i < QLDPC_NUM_COUNTERS; ++i) counters[i] = (uint64_t)h[i];
  return QLDPC_OK;
}

template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::string& s) {
    return fail(QLDPC_ERR_ARG, s);
  } catch (const std::bad_alloc&) {
    return fail(QLDPC_ERR_ARG, "out of host memory");
  } catch (const std::exception& e) {
    return fail(QLDPC_ERR_ARG, e.what());
  }
}

}  // namespace

extern "C" {

const char* qldpc_version(void) { return "qldpc_b200 0.1 (sm_100a)"; }
const char* qldpc_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------------------------------------------- code

int qldpc_code_create_qc(int J, int K, int L, int P, int sigma, int tau, qldpc_code** out) {
  if (!out) return fail(QLDPC_ERR_ARG, "null out pointer");
  return guarded([&] {
    *out = new qldpc_code{code_from_qc(J, K, L, P, sigma, tau)};
    return QLDPC_OK;
  });
}

int qldpc_code_create_dense(int J, int K, int L, int P, int sigma, int tau, const int32_t* pcmX, const int32_t* pcmZ,
                            const int32_t* iMinusP, qldpc_code** out) {
  if (!out) return fail(QLDPC_ERR_ARG, "null out pointer");
  return guarded([&] {
    *out = new qldpc_code{code_from_dense(J, K, L, P, sigma, tau, pcmX, pcmZ, iMinusP)};
    return QLDPC_OK;
  });
}

int qldpc_code_create_from_file(const char* path, qldpc_code** out) {
  if (!out || !path) return fail(QLDPC_ERR_ARG, "null argument");
  try {
    *out = new qldpc_code{code_from_file(path)};
    return QLDPC_OK;
  } catch (const std::string& s) {
    return fail(s.rfind("Unable to", 0) == 0 ? QLDPC_ERR_IO : QLDPC_ERR_ARG, s);
  } catch (const std::exception& e) {
    return fail(QLDPC_ERR_ARG, e.what());
  }
}

int qldpc_code_write_file(const qldpc_code* code, const char* path) {
  if (!code || !path) return fail(QLDPC_ERR_ARG, "null argument");
  try {
    code_write_file(*code->c, path);
    return QLDPC_OK;
  } catch (const std::string& s) {
    return fail(QLDPC_ERR_IO, s);
  }
}

void qldpc_code_destroy(qldpc_code* code) {
  if (code) {
    delete code->c;
    delete code;
  }
}

int qldpc_code_get_info(const qldpc_code* code, qldpc_code_info* o) {
  if (!code || !o) return fail(QLDPC_ERR_ARG, "null argument");
  const Code& c = *code->c;
  o->J = c.J; o->K = c.K; o->L = c.L; o->P = c.P; o->sigma = c.sigma; o->tau = c.tau;
  o->n = c.n; o->mX = c.side[0].m; o->mZ = c.side[1].m;
  o->dcX = c.side[0].dc; o->dcZ = c.side[1].dc; o->dvX = c.side[0].dv; o->dvZ = c.side[1].dv;
  o->EX = c.side[0].E; o->EZ = c.side[1].E;
  o->logical_rows = c.logical.rows;
  o->is_qc = c.is_qc;
  o->logical_from_file = c.logical_from_file;
  return QLDPC_OK;
}

int qldpc_code_name(const qldpc_code* code, char* out, int cap) {
  if (!code || !out || cap < 1) return fail(QLDPC_ERR_ARG, "null argument");
  const std::string s = code->c->name();
  snprintf(out, (size_t)cap, "%s", s.c_str());
  return QLDPC_OK;
}

int qldpc_code_exponents(const qldpc_code* code, int side, int32_t* out) {
  if (!code || !out || side < 0 || side > 1) return fail(QLDPC_ERR_ARG, "bad argument");
  const auto& h = code->c->side[side].hexp;
  if (h.empty()) return fail(QLDPC_ERR_ARG, "code is not quasi-cyclic in its (J,K,L,P,sigma,tau)");
  std::copy(h.begin(), h.end(), out);
  return QLDPC_OK;
}

int qldpc_code_csr(const qldpc_code* code, int side, int32_t* chk_var) {
  if (!code || !chk_var || side < 0 || side > 1) return fail(QLDPC_ERR_ARG, "bad argument");
  const auto& t = code->c->side[side];
  std::copy(t.chk_var.begin(), t.chk_var.end(), chk_var);
  return QLDPC_OK;
}

int qldpc_code_csc(const qldpc_code* code, int side, int32_t* var_chk, int32_t* var_edge) {
  if (!code || !var_chk || side < 0 || side > 1) return fail(QLDPC_ERR_ARG, "bad argument");
  const auto& t = code->c->side[side];
  std::copy(t.var_chk.begin(), t.var_chk.end(), var_chk);
  if (var_edge) std::copy(t.var_edge.begin(), t.var_edge.end(), var_edge);
  return QLDPC_OK;
}

int qldpc_code_dense(const qldpc_code* code, int which, int32_t* out) {
  if (!code || !out || which < 0 || which > 3) return fail(QLDPC_ERR_ARG, "bad argument");
  const Code& c = *code->c;
  if (which < 2) {
    c.dense_pcm(which, out);
  } else if (which == 3) {
    // the reference's shape contract: iMinusP is 2n x 2n (Quantum_LDPC_Code.h:16,60-74).  The matrix as it was
    // supplied when there is one, else the generated logical-check rows followed by zero rows (same kernel).
    const BitMatrix& src = c.iminusp_raw.rows ? c.iminusp_raw : c.logical;
    const int w = 2 * c.n;
    std::fill(out, out + (size_t)w * w, 0);
    for (int r = 0; r < src.rows && r < w; ++r)
      for (int col = 0; col < w; ++col) out[(size_t)r * w + col] = src.get(r, col);
  } else {
    for (int r = 0; r < c.logical.rows; ++r)
      for (int col = 0; col < 2 * c.n; ++col) out[(size_t)r * 2 * c.n + col] = c.logical.get(r, col);
  }
  return QLDPC_OK;
}

int qldpc_code_is_css(const qldpc_code* code) {
  if (!code) return fail(QLDPC_ERR_ARG, "null argument");
  return code->c->is_css() ? 1 : 0;
}

int qldpc_code_syndrome(const qldpc_code* code, int side, const int32_t* errors, int32_t* syndrome) {
  if (!code || !errors || !syndrome || side < 0 || side > 1) return fail(QLDPC_ERR_ARG, "bad argument");
  code->c->syndrome(side, errors, syndrome);
  return QLDPC_OK;
}

int qldpc_code_check_logical(const qldpc_code* code, const int32_t* errors2n) {
  if (!code || !errors2n) return fail(QLDPC_ERR_ARG, "null argument");
  return code->c->check_logical(errors2n) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------- decoder

int qldpc_decoder_create(const qldpc_code* code, int device_ordinal, int max_frames, qldpc_decoder** out) {
  if (!code || !out) return fail(QLDPC_ERR_ARG, "null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(QLDPC_ERR_NO_DEVICE,
                std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU decode path");
  }
  if (device_ordinal < 0) CU_TRY(cudaGetDevice(&device_ordinal));
  if (device_ordinal >= ndev) return fail(QLDPC_ERR_ARG, "device ordinal out of range");
  CU_TRY(cudaSetDevice(device_ordinal));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device_ordinal));
  if (prop.major < 10)
    return fail(QLDPC_ERR_NO_DEVICE, std::string("device ") + prop.name + " is not sm_100 class; kernels are built for sm_100a only");

  qldpc_decoder* d = new qldpc_decoder;
  d->code = *code->c;
  d->device = device_ordinal;
  d->num_sms = prop.multiProcessorCount;
  d->chunk = max_frames > 0 ? max_frames : (1 << 18);
  d->n = d->code.n;
  d->nw = (d->n + 31) / 32;
  auto bail = [&](int rc) {
    delete d;
    return rc;
  };
#define D_TRY(expr)                          \
  do {                                       \
    int rc_ = [&]() -> int {                 \
      CU_TRY(expr);                          \
      return QLDPC_OK;                       \
    }();                                     \
    if (rc_) return bail(rc_);               \
  } while (0)
  D_TRY(cudaStreamCreateWithFlags(&d->own_stream, cudaStreamNonBlocking));
  d->stream = d->own_stream;
  const int n = d->n;
  for (int side = 0; side < 2; ++side) {
    const SideTables& t = d->code.side[side];
    DevSide& s = d->s[side];
    s.m = t.m; s.dc = t.dc; s.dv = t.dv; s.E = t.E; s.mw = (t.m + 31) / 32;
    // the device tables hold check and message-row indices in 16 bits (syndrome_kernel, generate_syndrome_kernel and
    // the BP kernels all read them): a side beyond that is rejected here, before anything can be launched on it
    if (t.m > 65535) return bail(fail(QLDPC_ERR_UNSUPPORTED, "more than 65535 checks per side"));
    std::vector<uint16_t> vrow((size_t)t.E), vchk((size_t)t.E);
    for (int v = 0; v < n; ++v)
      for (int k = 0; k < t.dv; ++k) {
        const int edge = t.var_edge[(size_t)v * t.dv + k];
        const int e = edge / t.dc, i = edge % t.dc;
        vrow[(size_t)k * n + v] = (uint16_t)(i * t.m + e);
        vchk[(size_t)k * n + v] = (uint16_t)e;
      }
    D_TRY(dev_alloc(s.vrow, vrow.size()));
    D_TRY(dev_alloc(s.vchk, vchk.size()));
    D_TRY(cudaMemcpy(s.vrow, vrow.data(), vrow.size() * 2, cudaMemcpyHostToDevice));
    D_TRY(cudaMemcpy(s.vchk, vchk.data(), vchk.size() * 2, cudaMemcpyHostToDevice));
    if (!t.hexp.empty() && (int)t.hexp.size() == t.dv * d->code.L && t.m == t.dv * d->code.P) {
      // quasi-cyclic side: the kernels generate their index tables from the exponents instead of reading them
      D_TRY(dev_alloc(s.hexp, t.hexp.size()));
      D_TRY(cudaMemcpy(s.hexp, t.hexp.data(), t.hexp.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      s.P = d->code.P;
      s.L = d->code.L;
    }
    resolve_config(d, side);  // failure is reported when the side is first used
  }
  const size_t F = (size_t)d->chunk;
  D_TRY(dev_alloc(d->errX, F * d->nw));
  D_TRY(dev_alloc(d->errZ, F * d->nw));
  D_TRY(dev_alloc(d->decX, F * d->nw));
  D_TRY(dev_alloc(d->decZ, F * d->nw));
  D_TRY(dev_alloc(d->synX, F * d->s[0].mw));
  D_TRY(dev_alloc(d->synZ, F * d->s[1].mw));
  D_TRY(dev_alloc(d->sfX, F));
  D_TRY(dev_alloc(d->sfZ, F));
  D_TRY(dev_alloc(d->fflags, F));
  D_TRY(dev_alloc(d->itX, F));
  D_TRY(dev_alloc(d->itZ, F));
  D_TRY(dev_alloc(d->counters, (size_t)QLDPC_NUM_COUNTERS));
  D_TRY(dev_alloc(d->queues, (size_t)2));
  const Code& c = d->code;
  d->lx_rows = c.lx; d->lz_rows = c.lz; d->lm_rows = c.lm;
  int rc = upload_logical(c.logical, 0, c.lx, d->nw, n, false, d->lx);
  if (!rc) rc = upload_logical(c.logical, c.lx, c.lz, d->nw, n, true, d->lz);
  if (!rc) rc = upload_logical(c.logical, c.lx + c.lz, c.lm, 2 * d->nw, n, false, d->lm);
  if (rc) return bail(rc);
#undef D_TRY
  *out = d;
  return QLDPC_OK;
}

void qldpc_decoder_destroy(qldpc_decoder* dec) { delete dec; }

int qldpc_decoder_set_stream(qldpc_decoder* dec, void* cuda_stream) {
  if (!dec) return fail(QLDPC_ERR_ARG, "null decoder");
  dec->stream = cuda_stream ? (cudaStream_t)cuda_stream : dec->own_stream;
  return QLDPC_OK;
}

int qldpc_decoder_configure(qldpc_decoder* dec, int side, int frames_per_tile, int threads_per_cta, int ctas_per_sm) {
  if (!dec || side < 0 || side > 1) return fail(QLDPC_ERR_ARG, "bad argument");
  CU_TRY(cudaSetDevice(dec->device));
  BpLaunch keep = dec->s[side].user;
  const bool keep_force = dec->s[side].force_global;
  ++dec->cfg_epoch;
  dec->s[side].force_global = frames_per_tile < 0;
  if (frames_per_tile < 0) frames_per_tile = 0;
  dec->s[side].user = BpLaunch();
  dec->s[side].user.vec = frames_per_tile;
  dec->s[side].user.threads = threads_per_cta;
  dec->s[side].user.ctas_per_sm = ctas_per_sm;
  int rc = resolve_config(dec, side);
  if (!rc && !dec->s[side].force_global && dec->s[side].use_global && (frames_per_tile > 0 || threads_per_cta > 0))
    rc = QLDPC_ERR_ARG;  // an explicit tile shape that does not fit must not silently become the fallback path
  if (rc) {
    dec->s[side].user = keep;
    dec->s[side].force_global = keep_force;
    resolve_config(dec, side);
    return fail(rc, "configuration rejected");
  }
  return QLDPC_OK;
}

int qldpc_decoder_set_host_threads(qldpc_decoder* dec, int threads) {
  if (!dec) return fail(QLDPC_ERR_ARG, "null decoder");
  if (threads > 64) return fail(QLDPC_ERR_ARG, "at most 64 host threads");
  dec->host_threads = threads < 0 ? -1 : threads;
  return QLDPC_OK;
}

int qldpc_default_host_threads(void) { return default_host_threads(); }

int qldpc_decoder_host_threads_in_use(qldpc_decoder* dec, int elem_size) {
  if (!dec || (elem_size != 1 && elem_size != 4)) return fail(QLDPC_ERR_ARG, "bad argument");
  return host_threads_for(dec, elem_size);
}

int qldpc_decoder_launch_info(qldpc_decoder* dec, int side, int32_t out[8]) {
  if (!dec || !out || side < 0 || side > 1) return fail(QLDPC_ERR_ARG, "bad argument");
  const DevSide& s = dec->s[side];
  if (!s.cfg_ok) return fail(QLDPC_ERR_UNSUPPORTED, s.cfg_err);
  int v[8] = {s.cfg.vec, s.cfg.threads, s.cfg.ctas_per_sm, s.cfg.grid, s.cfg.smem, s.cfg.regs, dec->num_sms, dec->chunk};
  std::copy(v, v + 8, out);
  return QLDPC_OK;
}

int qldpc_decoder_enable_timing(qldpc_decoder* dec, int on) {
  if (!dec) return fail(QLDPC_ERR_ARG, "null decoder");
  dec->timing = on != 0;
  return QLDPC_OK;
}

int qldpc_decoder_get_timing(qldpc_decoder* dec, double* ms, uint64_t* launches, int reset) {
  if (!dec) return fail(QLDPC_ERR_ARG, "null decoder");
  CU_TRY(cudaSetDevice(dec->device));
  drain_timing(dec);
  for (int i = 0; i < QLDPC_NUM_TIMERS; ++i) {
    if (ms) ms[i] = dec->ms[i];
    if (launches) launches[i] = dec->launches[i];
    if (reset) { dec->ms[i] = 0; dec->launches[i] = 0; }
  }
  return QLDPC_OK;
}

int qldpc_decode_batch_device(qldpc_decoder* dec, const uint32_t* d_synX, const uint32_t* d_synZ, int64_t nframes,
                              float errorProbability, int maxIterations, uint32_t* d_outX, uint32_t* d_outZ,
                              uint8_t* d_outFlags, uint32_t* d_outIters) {
  int rc = check_common(dec, nframes, maxIterations);
  if (rc) return rc;
  if (!d_synX || !d_synZ || !d_outX || !d_outZ || !d_outFlags) return fail(QLDPC_ERR_ARG, "null buffer");
  qldpc_decoder* d = dec;
  for (int64_t off = 0; off < nframes; off += d->chunk) {
    const int nf = (int)std::min<int64_t>(d->chunk, nframes - off);
    rc = run_bp(d, d_synX + off * d->s[0].mw, d_synZ + off * d->s[1].mw, nf, errorProbability, maxIterations,
                d_outX + off * d->nw, d_outZ + off * d->nw, d->sfX, d->sfZ, d->itX, d->itZ);
    if (rc) return rc;
    {
      Timed t(d, QLDPC_T_PACK);
      CU_TRY(launch_merge_flags(d->sfX, d->sfZ, nf, d_outFlags + off, d->stream));
    }
    if (d_outIters) {
      CU_TRY(cudaMemcpy2DAsync(d_outIters + 2 * off, 8, d->itX, 4, 4, (size_t)nf, cudaMemcpyDeviceToDevice, d->stream));
      CU_TRY(cudaMemcpy2DAsync(d_outIters + 2 * off + 1, 8, d->itZ, 4, 4, (size_t)nf, cudaMemcpyDeviceToDevice, d->stream));
    }
  }
  CU_TRY(cudaStreamSynchronize(d->stream));
  return QLDPC_OK;
}

// Low-latency Decode for small batches (the call a reference-style per-frame loop makes, DecoderCPU.h:477): no host
// threads, no pipeline -- the syndromes go through one pinned staging buffer and one H2D copy, one launch packs both
// sides, the X and Z BP kernels run side by side on two streams, one launch produces everything the frames return,
// one D2H copy brings it back.
static int decode_small(qldpc_decoder* d, const uint8_t* synX, const uint8_t* synZ, int nf, float errorProbability,
                        int maxIterations, uint8_t* outX, uint8_t* outZ, uint8_t* outFlags, uint32_t* outIters) {
  const int n = d->n, mX = d->s[0].m, mZ = d->s[1].m;
  const size_t in_bytes = ((size_t)(mX + mZ) * nf + 15) / 16 * 16;
  const size_t it_off = in_bytes, x_off = it_off + (size_t)8 * nf, z_off = x_off + (size_t)n * nf,
               f_off = z_off + (size_t)n * nf, total = (f_off + nf + 15) / 16 * 16;
  if (total > d->small_pin_bytes) {
    if (d->small_pin) cudaFreeHost(d->small_pin);
    d->small_pin = nullptr;
    d->small_pin_bytes = 0;
    const size_t want = std::max<size_t>(total, (size_t)64 << 10);
    CU_TRY(cudaMallocHost((void**)&d->small_pin, want));
    d->small_pin_bytes = want;
  }
  int rc = ensure_stage(d, total);
  if (rc) return rc;
  uint8_t* h = d->small_pin;
  uint8_t* g = (uint8_t*)d->stage;
  std::memcpy(h, synX, (size_t)mX * nf);
  std::memcpy(h + (size_t)mX * nf, synZ, (size_t)mZ * nf);
  auto enqueue = [&]() -> int {
    CU_TRY(cudaMemcpyAsync(g, h, (size_t)(mX + mZ) * nf, cudaMemcpyHostToDevice, d->stream));
    {
      Timed t(d, QLDPC_T_PACK);
      CU_TRY(launch_pack2(g, mX, d->s[0].mw, d->synX, g + (size_t)mX * nf, mZ, d->s[1].mw, d->synZ, nf, d->stream));
    }
    int r = run_bp(d, d->synX, d->synZ, nf, errorProbability, maxIterations, d->decX, d->decZ, d->sfX, d->sfZ, d->itX, d->itZ);
    if (r) return r;
    {
      Timed t(d, QLDPC_T_PACK);
      CU_TRY(launch_finish_small(d->decX, d->decZ, nf, n, d->nw, g + x_off, g + z_off, d->sfX, d->sfZ, g + f_off, d->itX,
                                 d->itZ, (uint32_t*)(g + it_off), d->stream));
    }
    CU_TRY(cudaMemcpyAsync(h + it_off, g + it_off, f_off + nf - it_off, cudaMemcpyDeviceToHost, d->stream));
    return QLDPC_OK;
  };
  // Graph replay: valid while the shape (frames, prior, iteration limit), the staging buffers and the launch
  // configuration are those of the captured call.  Not used with per-kernel timing (its events are host-recorded) or
  // on the HBM-resident path (host-polled passes).
  uint32_t p_bits;
  std::memcpy(&p_bits, &errorProbability, 4);
  const bool graphable = !d->timing && !d->s[0].use_global && !d->s[1].use_global && ensure_side_stream(d) == QLDPC_OK;
  qldpc_decoder::SmallGraph& sg = d->small_graph;
  if (graphable && sg.exec && sg.nf == nf && sg.maxit == maxIterations && sg.p_bits == p_bits && sg.stage == g &&
      sg.pin == h && sg.cfg_epoch == d->cfg_epoch) {
    CU_TRY(cudaGraphLaunch(sg.exec, d->stream));
    d->launches[QLDPC_T_PACK] += 2;  // the replayed graph holds the same four kernels the capture counted
    d->launches[QLDPC_T_BP_X] += 1;
    d->launches[QLDPC_T_BP_Z] += 1;
  } else if (graphable) {
    if (sg.exec) cudaGraphExecDestroy(sg.exec);
    sg.exec = nullptr;
    cudaGraph_t graph = nullptr;
    CU_TRY(cudaStreamBeginCapture(d->stream, cudaStreamCaptureModeThreadLocal));
    rc = enqueue();
    const cudaError_t ce = cudaStreamEndCapture(d->stream, &graph);
    if (rc || ce != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc) return rc;
      rc = enqueue();  // capture refused (e.g. a legacy-stream interaction): plain stream operations
      if (rc) return rc;
    } else {
      const cudaError_t ie = cudaGraphInstantiate(&sg.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) {
        cudaGetLastError();
        sg.exec = nullptr;
        rc = enqueue();
        if (rc) return rc;
      } else {
        sg.nf = nf; sg.maxit = maxIterations; sg.p_bits = p_bits; sg.stage = g; sg.pin = h; sg.cfg_epoch = d->cfg_epoch;
        CU_TRY(cudaGraphLaunch(sg.exec, d->stream));
      }
    }
  } else {
    rc = enqueue();
    if (rc) return rc;
  }
  CU_TRY(cudaStreamSynchronize(d->stream));
  std::memcpy(outX, h + x_off, (size_t)n * nf);
  std::memcpy(outZ, h + z_off, (size_t)n * nf);
  std::memcpy(outFlags, h + f_off, (size_t)nf);
  if (outIters) std::memcpy(outIters, h + it_off, (size_t)8 * nf);
  return QLDPC_OK;
}

static int decode_batch_impl(qldpc_decoder* dec, const uint8_t* synX, const uint8_t* synZ, int64_t nframes,
                             float errorProbability, int maxIterations, uint8_t* outX, uint8_t* outZ, uint8_t* outFlags,
                             uint32_t* outIters);

int qldpc_decode_batch(qldpc_decoder* dec, const uint8_t* synX, const uint8_t* synZ, int64_t nframes,
                       float errorProbability, int maxIterations, uint8_t* outX, uint8_t* outZ, uint8_t* outFlags,
                       uint32_t* outIters) {
  int rc = check_common(dec, nframes, maxIterations);
  if (rc) return rc;
  if (!synX || !synZ || !outX || !outZ || !outFlags) return fail(QLDPC_ERR_ARG, "null buffer");
  if (nframes == 0) return QLDPC_OK;
  OverlapScope overlap(dec);
  if (nframes <= std::min<int64_t>(kSmallBatch, dec->chunk))
    rc = decode_small(dec, synX, synZ, (int)nframes, errorProbability, maxIterations, outX, outZ, outFlags, outIters);
  else
    rc = decode_batch_impl(dec, synX, synZ, nframes, errorProbability, maxIterations, outX, outZ, outFlags, outIters);
  if (rc) {
    const std::string keep = g_err;
    quiesce(dec);
    g_err = keep;
  }
  return rc;
}

static int decode_batch_impl(qldpc_decoder* dec, const uint8_t* synX, const uint8_t* synZ, int64_t nframes,
                             float errorProbability, int maxIterations, uint8_t* outX, uint8_t* outZ, uint8_t* outFlags,
                             uint32_t* outIters) {
  int rc = QLDPC_OK;
  // Three-stage pipeline over slices of at most kPipeFrames frames, two staging buffers per direction:
  //   copy stream : H2D of slice i+1          (waits until slice i-1 has been packed out of that buffer)
  //   main stream : pack, BP X/Z, unpack, flags of slice i   (waits until slice i-2 has left the output buffer)
  //   d2h stream  : D2H of slice i-1
  qldpc_decoder* d = dec;
  const int n = d->n, mX = d->s[0].m, mZ = d->s[1].m;
  const int64_t slice = std::min<int64_t>(std::min<int64_t>(d->chunk, kPipeFrames), std::max<int64_t>(nframes, 1));
  if (HostPacker* hp = host_packer(d, 1)) {
    // Host-packed variant of the same pipeline: syndromes are packed on the host before the H2D copy and the
    // corrections come back as packed words that the host threads expand to one byte per bit, so the link carries
    // 1/8 of the bytes in both directions and pageable user buffers cost nothing extra.  While the device works on
    // slice i the host packs slice i+1 and unpacks slice i-1.
    const int nw = d->nw, mwX = d->s[0].mw, mwZ = d->s[1].mw;
    const size_t in_w = (size_t)slice * (mwX + mwZ), out_w = (size_t)slice * 2 * nw;
    const size_t dev_one = (in_w + out_w) * sizeof(uint32_t) + (size_t)9 * slice + 64;
    rc = ensure_pin(d, 2 * (in_w + out_w));
    if (rc) return rc;
    rc = ensure_stage(d, 2 * ((dev_one + 15) / 16 * 16));
    if (rc) return rc;
    rc = ensure_pipeline(d);
    if (rc) return rc;
    const size_t dev_stride = (dev_one + 15) / 16 * 16;
    std::vector<int64_t> starts;  // first frame of every slice issued so far
    auto unpack_slice = [&](int j) {  // corrections of slice j: pinned words -> the caller's byte rows
      const int64_t o = starts[(size_t)j];
      const int cnt = (int)std::min<int64_t>(slice_frames(j, slice), nframes - o);
      const uint32_t* hout = d->pin + 2 * in_w + (size_t)(j & 1) * out_w;
      hp->unpack(hout, cnt, n, nw, outX + o * n);
      hp->unpack(hout + (size_t)cnt * nw, cnt, n, nw, outZ + o * n);
    };
    int i = 0;
    for (int64_t off = 0; off < nframes; off += slice_frames(i, slice), ++i) {
      const int nf = (int)std::min<int64_t>(slice_frames(i, slice), nframes - off);
      starts.push_back(off);
      const int b = i & 1;
      uint32_t* hin = d->pin + (size_t)b * in_w;
      uint32_t* hout = d->pin + 2 * in_w + (size_t)b * out_w;
      uint8_t* base = (uint8_t*)d->stage + (size_t)b * dev_stride;
      uint32_t* din = (uint32_t*)base;
      uint32_t* dout = din + in_w;
      uint32_t* its = dout + out_w;               // [nf][2]
      uint8_t* oF = (uint8_t*)(its + 2 * slice);  // [nf]
      const size_t xw = (size_t)nf * mwX, zw = (size_t)nf * mwZ, ow = (size_t)nf * nw;
      if (i >= 2) CU_TRY(cudaEventSynchronize(d->ev_ready[b]));  // H2D out of this pinned buffer has finished
      hp->pack(synX + off * mX, 1, nf, mX, mwX, hin);
      hp->pack(synZ + off * mZ, 1, nf, mZ, mwZ, hin + xw);
      if (i >= 2) CU_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_free[b], 0));
      CU_TRY(cudaMemcpyAsync(din, hin, (xw + zw) * sizeof(uint32_t), cudaMemcpyHostToDevice, d->copy_stream));
      CU_TRY(cudaEventRecord(d->ev_ready[b], d->copy_stream));
      CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_ready[b], 0));
      CU_TRY(cudaMemcpyAsync(d->synX, din, xw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
      CU_TRY(cudaMemcpyAsync(d->synZ, din + xw, zw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
      CU_TRY(cudaEventRecord(d->ev_free[b], d->stream));
      rc = run_bp(d, d->synX, d->synZ, nf, errorProbability, maxIterations, d->decX, d->decZ, d->sfX, d->sfZ, d->itX, d->itZ);
      if (rc) return rc;
      if (i >= 2) CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_out_free[b], 0));
      CU_TRY(cudaMemcpyAsync(dout, d->decX, ow * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
      CU_TRY(cudaMemcpyAsync(dout + ow, d->decZ, ow * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
      CU_TRY(launch_merge_flags(d->sfX, d->sfZ, nf, oF, d->stream));
      if (outIters) {
        CU_TRY(cudaMemcpy2DAsync(its, 8, d->itX, 4, 4, (size_t)nf, cudaMemcpyDeviceToDevice, d->stream));
        CU_TRY(cudaMemcpy2DAsync(its + 1, 8, d->itZ, 4, 4, (size_t)nf, cudaMemcpyDeviceToDevice, d->stream));
      }
      CU_TRY(cudaEventRecord(d->ev_done[b], d->stream));
      CU_TRY(cudaStreamWaitEvent(d->d2h_stream, d->ev_done[b], 0));
      // pinned out-buffer b was last read by unpack_slice(i - 2), which ran (on this thread) during iteration i - 1
      CU_TRY(cudaMemcpyAsync(hout, dout, 2 * ow * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->d2h_stream));
      CU_TRY(cudaMemcpyAsync(outFlags + off, oF, (size_t)nf, cudaMemcpyDeviceToHost, d->d2h_stream));
      if (outIters)
        CU_TRY(cudaMemcpyAsync(outIters + 2 * off, its, (size_t)nf * 8, cudaMemcpyDeviceToHost, d->d2h_stream));
      CU_TRY(cudaEventRecord(d->ev_out_free[b], d->d2h_stream));
      if (i >= 1) {
        CU_TRY(cudaEventSynchronize(d->ev_out_free[b ^ 1]));
        unpack_slice(i - 1);
      }
    }
    CU_TRY(cudaStreamSynchronize(d->stream));
    CU_TRY(cudaStreamSynchronize(d->d2h_stream));
    if (i >= 1) unpack_slice(i - 1);
    return QLDPC_OK;
  }
  const size_t in_bytes = (size_t)(mX + mZ) * slice;
  const size_t out_bytes = ((size_t)2 * n + 1) * slice + 15;
  const size_t it_bytes = (size_t)8 * slice;
  const size_t one = (in_bytes + out_bytes + it_bytes + 63) / 16 * 16;
  rc = ensure_stage(d, 2 * one);
  if (rc) return rc;
  rc = ensure_pipeline(d);
  if (rc) return rc;
  int i = 0;
  for (int64_t off = 0; off < nframes; off += slice_frames(i, slice), ++i) {
    const int nf = (int)std::min<int64_t>(slice_frames(i, slice), nframes - off);
    const int b = i & 1;
    uint8_t* base = (uint8_t*)d->stage + (size_t)b * one;
    uint32_t* its = (uint32_t*)base;          // [nf][2], kept first for alignment
    uint8_t* inX = base + it_bytes;
    uint8_t* inZ = inX + (size_t)mX * nf;
    uint8_t* oX = base + it_bytes + in_bytes;
    uint8_t* oZ = oX + (size_t)n * nf;
    uint8_t* oF = oZ + (size_t)n * nf;
    if (i >= 2) CU_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_free[b], 0));
    CU_TRY(cudaMemcpyAsync(inX, synX + off * mX, (size_t)nf * mX, cudaMemcpyHostToDevice, d->copy_stream));
    CU_TRY(cudaMemcpyAsync(inZ, synZ + off * mZ, (size_t)nf * mZ, cudaMemcpyHostToDevice, d->copy_stream));
    CU_TRY(cudaEventRecord(d->ev_ready[b], d->copy_stream));
    CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_ready[b], 0));
    {
      Timed t(d, QLDPC_T_PACK, 2);
      CU_TRY(launch_pack(inX, 1, nf, mX, d->s[0].mw, d->synX, d->stream));
      CU_TRY(launch_pack(inZ, 1, nf, mZ, d->s[1].mw, d->synZ, d->stream));
    }
    CU_TRY(cudaEventRecord(d->ev_free[b], d->stream));
    rc = run_bp(d, d->synX, d->synZ, nf, errorProbability, maxIterations, d->decX, d->decZ, d->sfX, d->sfZ, d->itX, d->itZ);
    if (rc) return rc;
    if (i >= 2) CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_out_free[b], 0));
    {
      Timed t(d, QLDPC_T_PACK, 3);
      CU_TRY(launch_unpack(d->decX, nf, n, d->nw, oX, d->stream));
      CU_TRY(launch_unpack(d->decZ, nf, n, d->nw, oZ, d->stream));
      CU_TRY(launch_merge_flags(d->sfX, d->sfZ, nf, oF, d->stream));
    }
    if (outIters) {
      CU_TRY(cudaMemcpy2DAsync(its, 8, d->itX, 4, 4, (size_t)nf, cudaMemcpyDeviceToDevice, d->stream));
      CU_TRY(cudaMemcpy2DAsync(its + 1, 8, d->itZ, 4, 4, (size_t)nf, cudaMemcpyDeviceToDevice, d->stream));
    }
    CU_TRY(cudaEventRecord(d->ev_done[b], d->stream));
    CU_TRY(cudaStreamWaitEvent(d->d2h_stream, d->ev_done[b], 0));
    CU_TRY(cudaMemcpyAsync(outX + off * n, oX, (size_t)nf * n, cudaMemcpyDeviceToHost, d->d2h_stream));
    CU_TRY(cudaMemcpyAsync(outZ + off * n, oZ, (size_t)nf * n, cudaMemcpyDeviceToHost, d->d2h_stream));
    CU_TRY(cudaMemcpyAsync(outFlags + off, oF, (size_t)nf, cudaMemcpyDeviceToHost, d->d2h_stream));
    if (outIters)
      CU_TRY(cudaMemcpyAsync(outIters + 2 * off, its, (size_t)nf * 8, cudaMemcpyDeviceToHost, d->d2h_stream));
    CU_TRY(cudaEventRecord(d->ev_out_free[b], d->d2h_stream));
  }
  CU_TRY(cudaStreamSynchronize(d->stream));
  CU_TRY(cudaStreamSynchronize(d->d2h_stream));
  return QLDPC_OK;
}

int qldpc_get_statistics_depolarizing(qldpc_decoder* dec, uint64_t seed, uint64_t first_frame, int64_t nframes,
                                      float p, int maxIterations, uint64_t* counters, uint8_t* perFrameFlags,
                                      uint32_t* perFrameIters) {
  int rc = check_common(dec, nframes, maxIterations);
  if (rc) return rc;
  qldpc_decoder* d = dec;
  if (p != p) return fail(QLDPC_ERR_ARG, "error probability is NaN");
  const Thresholds thr = depolarizing_thresholds(p);
  CU_TRY(cudaMemsetAsync(d->counters, 0, QLDPC_NUM_COUNTERS * sizeof(unsigned long long), d->stream));
  for (int64_t off = 0; off < nframes; off += d->chunk) {
    const int nf = (int)std::min<int64_t>(d->chunk, nframes - off);
    rc = run_generate(d, seed, first_frame + (uint64_t)off, nf, thr);
    if (rc) return rc;
    rc = finish_chunk(d, nf, off, p, maxIterations, perFrameFlags, perFrameIters);
    if (rc) return rc;
  }
  return read_counters(d, counters);
}

// GetStatistics(errorWeight, numErrors, ...) (DecoderCPU.h:392-530): the reference's fixed-weight generator
// (WeightWGenerator, host_pack.h: one serial mt19937 stream, as published) feeding the device decoder.  Slices are
// generated into two pinned buffers while the device decodes the previous slice.
static int weightw_impl(qldpc_decoder* dec, int errorWeight, int64_t numErrors, float errorProbability, int maxIterations,
                        uint32_t seed, uint64_t* counters, uint8_t* perFrameFlags, uint32_t* perFrameIters);

int qldpc_get_statistics_weightw(qldpc_decoder* dec, int errorWeight, int64_t numErrors, float errorProbability,
                                 int maxIterations, uint32_t seed, uint64_t* counters, uint8_t* perFrameFlags,
                                 uint32_t* perFrameIters) {
  if (!dec) return fail(QLDPC_ERR_ARG, "null decoder");
  OverlapScope overlap(dec);
  const int rc = weightw_impl(dec, errorWeight, numErrors, errorProbability, maxIterations, seed, counters, perFrameFlags,
                              perFrameIters);
  if (rc && dec) {
    const std::string keep = g_err;
    quiesce(dec);
    g_err = keep;
  }
  return rc;
}

static int weightw_impl(qldpc_decoder* dec, int errorWeight, int64_t numErrors, float errorProbability, int maxIterations,
                        uint32_t seed, uint64_t* counters, uint8_t* perFrameFlags, uint32_t* perFrameIters) {
  int rc = check_common(dec, numErrors, maxIterations);
  if (rc) return rc;
  if (errorWeight < 0) return fail(QLDPC_ERR_ARG, "negative error weight");
  qldpc_decoder* d = dec;
  const int n = d->n, nw = d->nw;
  const int64_t slice = std::min<int64_t>(std::min<int64_t>(d->chunk, 1 << 16), std::max<int64_t>(numErrors, 1));
  const size_t hwords = 2 * (size_t)slice * nw;  // x rows then z rows of one slice
  rc = ensure_pin(d, 2 * hwords);
  if (rc) return rc;
  rc = ensure_stage(d, 2 * hwords * sizeof(uint32_t));
  if (rc) return rc;
  rc = ensure_pipeline(d);
  if (rc) return rc;
  WeightWGenerator gen(seed, n, errorWeight);  // DecoderCPU.h:394
  HostPacker* pool = host_packer(d, 4);
  std::unique_ptr<HostPacker> own_pool;  // the weight-W generator always needs workers for its mapping stage
  if (!pool) {
    own_pool.reset(new HostPacker(std::max(1, default_host_threads())));
    pool = own_pool.get();
  }
  CU_TRY(cudaMemsetAsync(d->counters, 0, QLDPC_NUM_COUNTERS * sizeof(unsigned long long), d->stream));
  int i = 0;
  for (int64_t off = 0; off < numErrors; off += slice, ++i) {
    const int nf = (int)std::min<int64_t>(slice, numErrors - off);
    const int b = i & 1;
    uint32_t* hbuf = d->pin + (size_t)b * hwords;
    uint32_t* dbuf = (uint32_t*)d->stage + (size_t)b * hwords;
    const size_t xw = (size_t)nf * nw;
    if (i >= 2) CU_TRY(cudaEventSynchronize(d->ev_ready[b]));  // the copy out of this pinned buffer has finished
    gen.next(nf, nw, hbuf, hbuf + xw, pool);
    if (i >= 2) CU_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_free[b], 0));
    CU_TRY(cudaMemcpyAsync(dbuf, hbuf, 2 * xw * sizeof(uint32_t), cudaMemcpyHostToDevice, d->copy_stream));
    CU_TRY(cudaEventRecord(d->ev_ready[b], d->copy_stream));
    CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_ready[b], 0));
    CU_TRY(cudaMemcpyAsync(d->errX, dbuf, xw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
    CU_TRY(cudaMemcpyAsync(d->errZ, dbuf + xw, xw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
    CU_TRY(cudaEventRecord(d->ev_free[b], d->stream));
    rc = run_syndrome(d, nf);
    if (rc) return rc;
    rc = finish_chunk(d, nf, off, errorProbability, maxIterations, perFrameFlags, perFrameIters);
    if (rc) return rc;
  }
  return read_counters(d, counters);
}

// Host-supplied error patterns are processed in slices of at most kPipeFrames frames: slice i+1 is copied to the device
// on a second stream while slice i is packed, decoded and reduced (two staging buffers, events for hand-over), so
// with pinned host memory the PCIe transfer hides behind the decode (or vice versa).
static int stats_from_errors_impl(qldpc_decoder* d, const void* xErrors, const void* zErrors, int elem, int64_t numErrors,
                                  float errorProbability, int maxIterations, uint64_t* counters, uint8_t* perFrameFlags,
                                  uint32_t* perFrameIters);

static int stats_from_errors(qldpc_decoder* d, const void* xErrors, const void* zErrors, int elem, int64_t numErrors,
                             float errorProbability, int maxIterations, uint64_t* counters, uint8_t* perFrameFlags,
                             uint32_t* perFrameIters) {
  if (!d) return fail(QLDPC_ERR_ARG, "null decoder");
  OverlapScope overlap(d);
  const int rc = stats_from_errors_impl(d, xErrors, zErrors, elem, numErrors, errorProbability, maxIterations, counters,
                                        perFrameFlags, perFrameIters);
  if (rc && d) {
    const std::string keep = g_err;
    quiesce(d);
    g_err = keep;
  }
  return rc;
}

static int stats_from_errors_impl(qldpc_decoder* d, const void* xErrors, const void* zErrors, int elem, int64_t numErrors,
                                  float errorProbability, int maxIterations, uint64_t* counters, uint8_t* perFrameFlags,
                                  uint32_t* perFrameIters) {
  int rc = check_common(d, numErrors, maxIterations);
  if (rc) return rc;
  if (!xErrors || !zErrors) return fail(QLDPC_ERR_ARG, "null buffer");
  const int n = d->n;
  const size_t row = (size_t)n * elem;
  const int64_t slice = std::min<int64_t>(std::min<int64_t>(d->chunk, kPipeFrames), std::max<int64_t>(numErrors, 1));
  if (HostPacker* hp = host_packer(d, elem)) {
    // Host-packed pipeline: while the device decodes slice i the host threads pack slice i+1 into pinned memory;
    // only nw words per row cross the link (1/32 of the int layout), through a double-buffered device stage.
    const int nw = d->nw;
    const size_t hwords = 2 * (size_t)slice * nw;  // x rows then z rows of one slice
    rc = ensure_pin(d, 2 * hwords);
    if (rc) return rc;
    rc = ensure_stage(d, 2 * hwords * sizeof(uint32_t));
    if (rc) return rc;
    rc = ensure_pipeline(d);
    if (rc) return rc;
    CU_TRY(cudaMemsetAsync(d->counters, 0, QLDPC_NUM_COUNTERS * sizeof(unsigned long long), d->stream));
    int i = 0;
    for (int64_t off = 0; off < numErrors; off += slice_frames(i, slice), ++i) {
      const int nf = (int)std::min<int64_t>(slice_frames(i, slice), numErrors - off);
      const int b = i & 1;
      uint32_t* hbuf = d->pin + (size_t)b * hwords;
      uint32_t* dbuf = (uint32_t*)d->stage + (size_t)b * hwords;
      const size_t xw = (size_t)nf * nw;
      if (i >= 2) CU_TRY(cudaEventSynchronize(d->ev_ready[b]));  // the copy out of this pinned buffer has finished
      hp->pack((const uint8_t*)xErrors + off * row, elem, nf, n, nw, hbuf);
      hp->pack((const uint8_t*)zErrors + off * row, elem, nf, n, nw, hbuf + xw);
      if (i >= 2) CU_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_free[b], 0));
      CU_TRY(cudaMemcpyAsync(dbuf, hbuf, 2 * xw * sizeof(uint32_t), cudaMemcpyHostToDevice, d->copy_stream));
      CU_TRY(cudaEventRecord(d->ev_ready[b], d->copy_stream));
      CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_ready[b], 0));
      CU_TRY(cudaMemcpyAsync(d->errX, dbuf, xw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
      CU_TRY(cudaMemcpyAsync(d->errZ, dbuf + xw, xw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d->stream));
      CU_TRY(cudaEventRecord(d->ev_free[b], d->stream));
      rc = run_syndrome(d, nf);
      if (rc) return rc;
      rc = finish_chunk(d, nf, off, errorProbability, maxIterations, perFrameFlags, perFrameIters);
      if (rc) return rc;
    }
    return read_counters(d, counters);
  }
  const size_t half = 2 * row * (size_t)slice;  // one staging buffer: x rows then z rows
  rc = ensure_stage(d, 2 * half);
  if (rc) return rc;
  rc = ensure_pipeline(d);
  if (rc) return rc;
  CU_TRY(cudaMemsetAsync(d->counters, 0, QLDPC_NUM_COUNTERS * sizeof(unsigned long long), d->stream));
  int i = 0;
  for (int64_t off = 0; off < numErrors; off += slice_frames(i, slice), ++i) {
    const int nf = (int)std::min<int64_t>(slice_frames(i, slice), numErrors - off);
    const int b = i & 1;
    uint8_t* st0 = (uint8_t*)d->stage + (size_t)b * half;
    uint8_t* st1 = st0 + row * nf;
    if (i >= 2) CU_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_free[b], 0));
    CU_TRY(cudaMemcpyAsync(st0, (const uint8_t*)xErrors + off * row, row * nf, cudaMemcpyHostToDevice, d->copy_stream));
    CU_TRY(cudaMemcpyAsync(st1, (const uint8_t*)zErrors + off * row, row * nf, cudaMemcpyHostToDevice, d->copy_stream));
    CU_TRY(cudaEventRecord(d->ev_ready[b], d->copy_stream));
    CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_ready[b], 0));
    {
      Timed t(d, QLDPC_T_PACK, 2);
      CU_TRY(launch_pack(st0, elem, nf, n, d->nw, d->errX, d->stream));
      CU_TRY(launch_pack(st1, elem, nf, n, d->nw, d->errZ, d->stream));
    }
    CU_TRY(cudaEventRecord(d->ev_free[b], d->stream));
    rc = run_syndrome(d, nf);
    if (rc) return rc;
    rc = finish_chunk(d, nf, off, errorProbability, maxIterations, perFrameFlags, perFrameIters);
    if (rc) return rc;
  }
  return read_counters(d, counters);
}

int qldpc_get_stats_from_errors_i32(qldpc_decoder* dec, const int32_t* xErrors, const int32_t* zErrors,
                                    int64_t numErrors, float errorProbability, int maxIterations, uint64_t* counters,
                                    uint8_t* perFrameFlags, uint32_t* perFrameIters) {
  return stats_from_errors(dec, xErrors, zErrors, 4, numErrors, errorProbability, maxIterations, counters, perFrameFlags,
                           perFrameIters);
}

int qldpc_get_stats_from_errors_u8(qldpc_decoder* dec, const uint8_t* xErrors, const uint8_t* zErrors,
                                   int64_t numErrors, float errorProbability, int maxIterations, uint64_t* counters,
                                   uint8_t* perFrameFlags, uint32_t* perFrameIters) {
  return stats_from_errors(dec, xErrors, zErrors, 1, numErrors, errorProbability, maxIterations, counters, perFrameFlags,
                           perFrameIters);
}

// -------------------------------------------------------------------------------------------------- taps

int qldpc_debug_weightw_patterns(uint32_t seed, int errorWeight, int n, int64_t nframes, int threads, uint32_t* xWords,
                                 uint32_t* zWords) {
  if (errorWeight < 0 || n < 1 || nframes < 0 || threads < 1 || threads > 64 || !xWords || !zWords)
    return fail(QLDPC_ERR_ARG, "bad argument");
  HostPacker pool(threads);
  WeightWGenerator gen(seed, n, errorWeight);
  const int nw = (n + 31) / 32;
  // two calls, so that the hand-over of the stream position between slices is exercised as well
  const int64_t first = nframes / 3;
  gen.next(first, nw, xWords, zWords, &pool);
  gen.next(nframes - first, nw, xWords + (size_t)first * nw, zWords + (size_t)first * nw, &pool);
  return QLDPC_OK;
}

int qldpc_debug_host_pack(const void* src, int elem_size, int64_t rows, int cols, uint32_t* dst, int threads) {
  if (!src || !dst || rows < 0 || cols < 1 || (elem_size != 1 && elem_size != 4) || threads < 1 || threads > 64)
    return fail(QLDPC_ERR_ARG, "bad argument");
  HostPacker hp(threads);
  hp.pack(src, elem_size, rows, cols, (cols + 31) / 32, dst);
  return QLDPC_OK;
}

int qldpc_debug_host_read_gbs(const void* src, int64_t bytes, int threads, int repeats, double* gbs) {
  if (!src || bytes < 1 || threads < 1 || threads > 64 || repeats < 1 || !gbs) return fail(QLDPC_ERR_ARG, "bad argument");
  HostPacker hp(threads);
  volatile uint64_t sink = hp.read_all(src, (size_t)bytes);  // first pass: page in, wake the workers
  double best = 0.0;
  for (int r = 0; r < repeats; ++r) {
    const auto t0 = std::chrono::steady_clock::now();
    sink = sink | hp.read_all(src, (size_t)bytes);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    best = std::max(best, (double)bytes / sec / 1e9);
  }
  (void)sink;
  *gbs = best;
  return QLDPC_OK;
}

int qldpc_debug_host_unpack(const uint32_t* src, int64_t rows, int cols, uint8_t* dst, int threads) {
  if (!src || !dst || rows < 0 || cols < 1 || threads < 1 || threads > 64) return fail(QLDPC_ERR_ARG, "bad argument");
  HostPacker hp(threads);
  hp.unpack(src, rows, cols, (cols + 31) / 32, dst);
  return QLDPC_OK;
}

int qldpc_debug_generate(qldpc_decoder* dec, uint64_t seed, uint64_t first_frame, int64_t nframes, float p,
                         uint8_t* xerr, uint8_t* zerr, uint8_t* synX, uint8_t* synZ) {
  int rc = check_common(dec, nframes, 1);
  if (rc) return rc;
  qldpc_decoder* d = dec;
  const int n = d->n, mX = d->s[0].m, mZ = d->s[1].m;
  const size_t per = (size_t)std::max(n, std::max(mX, mZ));
  rc = ensure_stage(d, per * std::min<int64_t>(d->chunk, std::max<int64_t>(nframes, 1)));
  if (rc) return rc;
  if (p != p) return fail(QLDPC_ERR_ARG, "error probability is NaN");
  const Thresholds thr = depolarizing_thresholds(p);
  for (int64_t off = 0; off < nframes; off += d->chunk) {
    const int nf = (int)std::min<int64_t>(d->chunk, nframes - off);
    rc = run_generate(d, seed, first_frame + (uint64_t)off, nf, thr);
    if (rc) return rc;
    struct { const uint32_t* src; int cols, words; uint8_t* dst; } jobs[4] = {
        {d->errX, n, d->nw, xerr}, {d->errZ, n, d->nw, zerr}, {d->synX, mX, d->s[0].mw, synX}, {d->synZ, mZ, d->s[1].mw, synZ}};
    for (auto& j : jobs) {
      if (!j.dst) continue;
      CU_TRY(launch_unpack(j.src, nf, j.cols, j.words, (uint8_t*)d->stage, d->stream));
      CU_TRY(cudaMemcpyAsync(j.dst + off * j.cols, d->stage, (size_t)nf * j.cols, cudaMemcpyDeviceToHost, d->stream));
      CU_TRY(cudaStreamSynchronize(d->stream));
    }
  }
  return QLDPC_OK;
}

int qldpc_debug_bp_trace(qldpc_decoder* dec, int side, const uint8_t* syn, int nframes, float errorProbability,
                         int maxIterations, int cap_iters, float* q_trace, float* r_trace, uint32_t* iters) {
  int rc = check_common(dec, nframes, maxIterations);
  if (rc) return rc;
  if (side < 0 || side > 1 || !syn || cap_iters < 1) return fail(QLDPC_ERR_ARG, "bad argument");
  qldpc_decoder* d = dec;
  if (nframes > d->chunk) return fail(QLDPC_ERR_ARG, "trace batch larger than max_frames");
  const DevSide& s = d->s[side];
  const size_t tsz = (size_t)nframes * cap_iters * s.E;
  rc = ensure_stage(d, (size_t)nframes * s.m);
  if (rc) return rc;
  float *dq = nullptr, *dr = nullptr;
  CU_TRY(dev_alloc(dq, tsz));
  CU_TRY(dev_alloc(dr, tsz));
  auto cleanup = [&](int r) {
    cudaFree(dq);
    cudaFree(dr);
    return r;
  };
  uint32_t* dsyn = side ? d->synZ : d->synX;
  rc = [&]() -> int {
    CU_TRY(cudaMemsetAsync(dq, 0, tsz * 4, d->stream));
    CU_TRY(cudaMemsetAsync(dr, 0, tsz * 4, d->stream));
    CU_TRY(cudaMemcpyAsync(d->stage, syn, (size_t)nframes * s.m, cudaMemcpyHostToDevice, d->stream));
    CU_TRY(launch_pack(d->stage, 1, nframes, s.m, s.mw, dsyn, d->stream));
    int r = run_bp(d, d->synX, d->synZ, nframes, errorProbability, maxIterations, d->decX, d->decZ, d->sfX, d->sfZ, d->itX,
                   d->itZ, side, dq, dr, cap_iters);
    if (r) return r;
    if (q_trace) CU_TRY(cudaMemcpyAsync(q_trace, dq, tsz * 4, cudaMemcpyDeviceToHost, d->stream));
    if (r_trace) CU_TRY(cudaMemcpyAsync(r_trace, dr, tsz * 4, cudaMemcpyDeviceToHost, d->stream));
    if (iters)
      CU_TRY(cudaMemcpyAsync(iters, side ? d->itZ : d->itX, (size_t)nframes * 4, cudaMemcpyDeviceToHost, d->stream));
    CU_TRY(cudaStreamSynchronize(d->stream));
    return QLDPC_OK;
  }();
  return cleanup(rc);
}

int qldpc_debug_division_check(qldpc_decoder* dec, uint64_t seed, int64_t npairs, uint64_t out[3]) {
  if (!dec || !out || npairs < 0) return fail(QLDPC_ERR_ARG, "bad argument");
  CU_TRY(cudaSetDevice(dec->device));
  unsigned long long* d = nullptr;
  CU_TRY(dev_alloc(d, (size_t)3));
  int rc = [&]() -> int {
    CU_TRY(cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), dec->stream));
    CU_TRY(launch_division_check(seed, (long long)npairs, d, dec->stream));
    unsigned long long h[3];
    CU_TRY(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, dec->stream));
    CU_TRY(cudaStreamSynchronize(dec->stream));
    for (int i = 0; i < 3; ++i) out[i] = (uint64_t)h[i];
    return QLDPC_OK;
  }();
  cudaFree(d);
  return rc;
}

}  // extern "C"
