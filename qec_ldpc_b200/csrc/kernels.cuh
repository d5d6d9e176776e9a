// Launch interface of the device code (kernels.cu) used by the host decoder (decoder.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "bp_kernel.cuh"
#include "philox.cuh"

namespace qldpc {

struct BpLaunch {
  int vec = 0;       // frame slots per CTA tile
  int threads = 0;   // threads per CTA
  int ctas_per_sm = 0;
  int grid = 0;
  int smem = 0;      // dynamic shared memory per CTA
  int regs = 0;
  int max_threads = 0;
  int spec_m = 0;    // check count the kernel is specialised for (0: generic instantiation), set by bp_configure
};

// Picks a kernel instantiation for (dc, dv), sizes the tile and fills `cfg` (zero fields = heuristic).
// Returns false if no compiled instantiation covers the shape or the tile does not fit in shared memory.
// qc_P: circulant size of a quasi-cyclic side (dv block rows, n / qc_P column blocks), 0 for any other code; the
// instantiations specialised for a check count assume that structure.
bool bp_configure(int dc, int dv, int m, int n, int qc_P, int num_sms, BpLaunch& cfg, const char** why);
// guard: division range tests compiled in (0 none, 1 numerator, 3 both), see bp_kernel.cuh:div_fast
cudaError_t bp_launch(int dc, int dv, const BpLaunch& cfg, const BpArgs& args, int nframes, int guard, cudaStream_t st);

// Global-memory fallback path (bp_global.cu).
struct GlobalBpArgs {
  int m, n, dc, dv, mw, nw, maxit, batch;
  int slots = 0;         // frame slots in flight (0 = heuristic), at most `batch`
  int guard = 3;         // division range tests outside the last iteration: 0 / 1 / 3 (decoder.cu:division_guard)
  float prior;
  const uint32_t* vrow;  // [dv][n]
  const uint32_t* cvar;  // [dc][m]
  float* msg;            // [E][batch]
  uint8_t* bytes;        // 3 per-slot byte arrays
  uint32_t* words;       // bit-packed syndrome / decision words, frame ids, iteration indices, counters, lists
  unsigned int* host_done = nullptr;  // mapped pinned word: completion count mirrored for the host (no stream syncs)
};
size_t global_bp_bytes(int m, int n, int dc, int batch, size_t* msg_bytes, size_t* byte_bytes, size_t* word_bytes);
cudaError_t global_bp_run(const GlobalBpArgs& a, const uint32_t* syn, uint32_t* dec, uint8_t* flags, uint32_t* iters,
                          int nframes, int* launches, cudaStream_t st);
// X and Z side of one decoder concurrently on two streams (same results as two global_bp_run calls).
cudaError_t global_bp_run_pair(const GlobalBpArgs& ax, const uint32_t* synX, uint32_t* decX, uint8_t* flagsX, uint32_t* itersX,
                               cudaStream_t stX, const GlobalBpArgs& az, const uint32_t* synZ, uint32_t* decZ,
                               uint8_t* flagsZ, uint32_t* itersZ, cudaStream_t stZ, int nframes);

// Philox depolarizing errors, bit-packed: errX, errZ [nframes][nw], and their syndromes synX [nframes][mwX],
// synZ [nframes][mwZ] in the same kernel (synX == nullptr: errors only).
cudaError_t launch_generate_syndrome(uint64_t seed, uint64_t first_frame, int nframes, int n, int nw, Thresholds thr,
                                     uint32_t* errX, uint32_t* errZ, const uint16_t* vchkX, int dvX, int mwX,
                                     uint32_t* synX, const uint16_t* vchkZ, int dvZ, int mwZ, uint32_t* synZ,
                                     cudaStream_t st);
// s = H e (mod 2) for both sides from bit-packed errors; vchk = CSC tables [dv][n] (check index of the k-th edge).
cudaError_t launch_syndrome(const uint32_t* errX, const uint32_t* errZ, int nframes, int n, int nw,
                            const uint16_t* vchkX, int dvX, int mwX, uint32_t* synX, const uint16_t* vchkZ, int dvZ,
                            int mwZ, uint32_t* synZ, cudaStream_t st);
// rows of `bits` one-per-element (elem_size 1 or 4 bytes) <-> bit-packed words
cudaError_t launch_pack(const void* src, int elem_size, int64_t rows, int cols, int words, uint32_t* dst, cudaStream_t st);
cudaError_t launch_unpack(const uint32_t* src, int64_t rows, int cols, int words, uint8_t* dst, cudaStream_t st);

// low-latency Decode path: both syndrome rows (one byte per bit) packed by one launch; corrections as bytes, ErrorCode
// bits and iteration counts [rows][2] produced by one launch
cudaError_t launch_pack2(const uint8_t* srcX, int colsX, int wordsX, uint32_t* dstX, const uint8_t* srcZ, int colsZ,
                         int wordsZ, uint32_t* dstZ, int rows, cudaStream_t st);
cudaError_t launch_finish_small(const uint32_t* decX, const uint32_t* decZ, int rows, int cols, int words, uint8_t* outX,
                                uint8_t* outZ, const uint8_t* sfX, const uint8_t* sfZ, uint8_t* flags, const uint32_t* itX,
                                const uint32_t* itZ, uint32_t* iters, cudaStream_t st);

struct StatsArgs {
  const uint32_t *errX, *errZ, *decX, *decZ;  // [nframes][nw]
  const uint8_t *sfX, *sfZ;                   // per-side flags from the BP kernel
  const uint32_t *itX, *itZ;
  // logical check rows, transposed and padded to a multiple of 32 rows: LT[w][rows_pad]
  const uint32_t *lx, *lz, *lm;
  int lx_rows, lz_rows, lm_rows;
  int nframes, nw;
  unsigned long long* counters;  // [QLDPC_NUM_COUNTERS]
  uint8_t* fflags;               // [nframes] final per-frame flags (may be null)
};
cudaError_t launch_stats(const StatsArgs& a, cudaStream_t st);
// ErrorCode bits only (decode_batch): flags[f] = synX | synZ<<1 | cvX<<2 | cvZ<<3
cudaError_t launch_merge_flags(const uint8_t* sfX, const uint8_t* sfZ, int nframes, uint8_t* out, cudaStream_t st);

// out[3] = {mismatches vs __fdiv_rn among pairs div_fast accepts, pairs flagged unsafe, pairs with x == 0}
cudaError_t launch_division_check(uint64_t seed, long long npairs, unsigned long long* out, cudaStream_t st);

}  // namespace qldpc
