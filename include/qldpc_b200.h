/* qldpc_b200 -- C ABI of the B200-native Monte-Carlo BP decoder for quasi-cyclic quantum CSS LDPC codes.
 *
 * This is the drop-in boundary for the decode path of cantwellc/QEC_LDPC.  The reference has no FFI: its
 * boundary is the C++ virtual class `Decoder` (QEC_LDPC/Decoder.h:7-48) with `DecoderCPU` / `DecoderGPU`
 * behind it.  Every entry point below names the reference interface it replaces (file:line relative to
 * /root/reference/QEC_LDPC/); the C++ classes of the same names in qec_ldpc_b200/cpp/ are thin header-only
 * wrappers over this ABI (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, opaque handles, `int` status returns
 * (0 = ok, negative = error; text via qldpc_last_error()), no exceptions cross the boundary, no torch types.
 * There is NO CPU decode path: every decode / statistics call runs hand-written sm_100a CUDA kernels and
 * returns QLDPC_ERR_NO_DEVICE (never a silent fallback) when no CUDA device is usable.
 * One decoder handle per host thread / GPU; calls on one handle are serialised (stream-ordered).
 */
#ifndef QLDPC_B200_H
#define QLDPC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define QLDPC_OK 0
#define QLDPC_ERR_ARG (-1)         /* bad argument / inconsistent sizes            */
#define QLDPC_ERR_CUDA (-2)        /* a CUDA runtime call or kernel failed          */
#define QLDPC_ERR_NO_DEVICE (-3)   /* no usable CUDA device: there is no CPU path   */
#define QLDPC_ERR_UNSUPPORTED (-4) /* code shape not covered by the compiled kernels */
#define QLDPC_ERR_IO (-5)          /* file could not be read / written              */

typedef struct qldpc_code qldpc_code;       /* Quantum_LDPC_Code (Quantum_LDPC_Code.h:7-150) + edge tables */
typedef struct qldpc_decoder qldpc_decoder; /* DecoderGPU (DecoderGPU.h:11-281) state on one device          */

/* Per-frame result bits.  Bits 0..3 are Decoder::ErrorCode (Decoder.h:14-23); bits 4..6 are the per-frame
 * outcome GetStatistics derives (DecoderCPU.h:492-509) plus a NaN diagnostic. */
#define QLDPC_SYNDROME_FAIL_X 1
#define QLDPC_SYNDROME_FAIL_Z 2
#define QLDPC_CONVERGENCE_FAIL_X 4
#define QLDPC_CONVERGENCE_FAIL_Z 8
#define QLDPC_FRAME_LOGICAL 16
#define QLDPC_FRAME_CORRECTED 32
#define QLDPC_FRAME_NAN 64

/* Counter vector filled by the statistics calls (CodeStatistics, CodeStatistics.h:5-20, plus iteration sums). */
enum {
  QLDPC_C_FRAMES = 0,    /* numErrorsTested      */
  QLDPC_C_XTESTED = 1,   /* numXErrorsTested     */
  QLDPC_C_ZTESTED = 2,   /* numZErrorsTested     */
  QLDPC_C_CORRECTED = 3, /* corrected            */
  QLDPC_C_SYNX = 4,      /* syndromeErrorsX      */
  QLDPC_C_SYNZ = 5,      /* syndromeErrorsZ      */
  QLDPC_C_LOGICAL = 6,   /* logicalErrors        */
  QLDPC_C_CVX = 7,       /* convergenceFailX     */
  QLDPC_C_CVZ = 8,       /* convergenceFailZ     */
  QLDPC_C_ITERSX = 9,    /* sum of executed BP iterations, X side */
  QLDPC_C_ITERSZ = 10,   /* sum of executed BP iterations, Z side */
  QLDPC_C_NANFRAMES = 11,/* frames whose final messages hold a NaN */
  QLDPC_NUM_COUNTERS = 12
};

typedef struct qldpc_code_info {
  int32_t J, K, L, P, sigma, tau;
  int32_t n, mX, mZ;     /* n = L*P, numEqsX = J*P, numEqsZ = K*P (Quantum_LDPC_Code.h:84) */
  int32_t dcX, dcZ;      /* check degrees    (numVarsPerEq, DecoderCPU.h:299)                */
  int32_t dvX, dvZ;      /* variable degrees (numEqsPerVar, DecoderCPU.h:300)                */
  int32_t EX, EZ;        /* Tanner-graph edges per side                                      */
  int32_t logical_rows;  /* rows of the (row-reduced) logical-check matrix in use            */
  int32_t is_qc;         /* 1 if pcmX/pcmZ equal the Hagiwara-Imai expansion of (J..tau)     */
  int32_t logical_from_file; /* 1 if iMinusP came from a code file / caller, 0 if generated */
} qldpc_code_info;

const char* qldpc_version(void);
const char* qldpc_last_error(void); /* thread-local text of the last error on this thread */

/* ---- code construction -------------------------------------------------------------------------------- */

/* QC_LDPC_CSS::QC_LDPC_CSS(J,K,L,P,sigma,tau), commented-out in the reference: QC_LDPC_CSS.h:146-156,
 * formulas QEC_LDPC_CSS.cu:37-131.  Emits packed CSR/CSC edge tables by closed-form circulant indexing.
 * The logical check (the reference reads iMinusP from file only) is generated: a basis of ker(pcmX) and of
 * ker(pcmZ), which has the same kernel as the file's iMinusP and so gives identical decisions. */
int qldpc_code_create_qc(int J, int K, int L, int P, int sigma, int tau, qldpc_code** out);

/* Quantum_LDPC_Code(J,K,L,P,sigma,tau,pcmX,pcmZ,imp), Quantum_LDPC_Code.h:82-88.  Dense row-major 0/1 ints:
 * pcmX [J*P x L*P], pcmZ [K*P x L*P], iMinusP [2LP x 2LP] or NULL (then the logical check is generated as above).
 * Tables are built as DecoderCPU::InitIndexArrays does (DecoderCPU.h:41-84): ascending scan, regular degrees. */
int qldpc_code_create_dense(int J, int K, int L, int P, int sigma, int tau, const int32_t* pcmX, const int32_t* pcmZ,
                            const int32_t* iMinusP, qldpc_code** out);

/* Quantum_LDPC_Code::createFromFile, Quantum_LDPC_Code.h:26-80: the 4-line text format
 * (`J K L P sigma tau` / pcmX / pcmZ / iMinusP, whitespace separated). */
int qldpc_code_create_from_file(const char* path, qldpc_code** out);
int qldpc_code_write_file(const qldpc_code* code, const char* path);
void qldpc_code_destroy(qldpc_code* code);

int qldpc_code_get_info(const qldpc_code* code, qldpc_code_info* out);
/* operator<<(ostream, Quantum_LDPC_Code), Quantum_LDPC_Code.h:145-150: "[J=..,K=..,...][[n=..,k=..]]". */
int qldpc_code_name(const qldpc_code* code, char* out, int cap);
/* side: 0 = X (pcmX, hHC), 1 = Z (pcmZ, hHD). */
int qldpc_code_exponents(const qldpc_code* code, int side, int32_t* out /* [J*L] or [K*L]; needs is_qc */);
int qldpc_code_csr(const qldpc_code* code, int side, int32_t* chk_var /* [m*dc], ascending variable */);
int qldpc_code_csc(const qldpc_code* code, int side, int32_t* var_chk /* [n*dv], ascending check */,
                   int32_t* var_edge /* [n*dv] check-major edge id e*dc+i, may be NULL */);
/* which: 0 pcmX, 1 pcmZ, 2 logical-check matrix in use [logical_rows x 2n] (row-reduced), 3 iMinusP in the reference's
 * shape [2n x 2n]: the matrix as supplied (file / caller) when there is one, else the generated rows padded with zero rows. */
int qldpc_code_dense(const qldpc_code* code, int which, int32_t* out);
/* 1 if pcmX * pcmZ^T == 0 (mod 2), 0 if not (the reference never checks). */
int qldpc_code_is_css(const qldpc_code* code);
/* Host-side single-vector helpers of the code object (not the decode path):
 * GetSyndromeX/Z (Quantum_LDPC_Code.h:94-124) and CheckLogicalError (Quantum_LDPC_Code.h:126-142). */
int qldpc_code_syndrome(const qldpc_code* code, int side, const int32_t* errors /* [n] */, int32_t* syndrome /* [m] */);
int qldpc_code_check_logical(const qldpc_code* code, const int32_t* errors2n /* [2n] */);

/* ---- decoder ------------------------------------------------------------------------------------------ */

/* DecoderGPU::DecoderGPU(code), DecoderGPU.h:117-130: uploads the edge tables, allocates frame buffers for up to
 * max_frames frames per launch (larger requests are processed in chunks).  device_ordinal < 0 = current device.
 * A decoder owns its device buffers, streams and host threads: use one decoder per host thread (several decoders,
 * also on the same device, may run concurrently; a code object may be shared by any number of decoders). */
int qldpc_decoder_create(const qldpc_code* code, int device_ordinal, int max_frames, qldpc_decoder** out);
void qldpc_decoder_destroy(qldpc_decoder* dec);
/* Run all work of this handle on the given cudaStream_t (NULL = the handle's own stream). */
int qldpc_decoder_set_stream(qldpc_decoder* dec, void* cuda_stream);
/* Tuning knobs (0 = heuristic): frames per CTA tile (1, 2 or 4) and threads per CTA for side 0/1.
 * frames_per_tile = -1 selects the global-memory (HBM-resident) BP path, which is otherwise used only for shapes the
 * shared-memory tile kernel does not cover (no instantiation for the degrees, or a frame larger than shared memory);
 * with -1, threads_per_cta is the number of frame slots kept in flight (0 = heuristic) and ctas_per_sm is ignored. */
int qldpc_decoder_configure(qldpc_decoder* dec, int side, int frames_per_tile, int threads_per_cta, int ctas_per_sm);
/* Host-buffer entry points (qldpc_get_stats_from_errors_*, qldpc_decode_batch) convert the reference's
 * one-element-per-bit rows to packed words on the host with `threads` worker threads, so that 1/32 (int) or 1/8 (byte)
 * of the bytes cross the host-device link.  threads < 0: default (environment QLDPC_HOST_THREADS, else
 * min(16, hardware threads / processes on this host as announced by the launcher); with fewer than 10 (int32 rows) or
 * 8 (byte rows) the raw rows are copied instead, see qldpc_decoder_host_threads_in_use); 0: off -- raw rows
 * are copied and packed on the device.  Batches of at most 2048 frames of qldpc_decode_batch take a low-latency path
 * without host threads.  Marshalling only: the decode itself never runs on the host. */
int qldpc_decoder_set_host_threads(qldpc_decoder* dec, int threads);
int qldpc_default_host_threads(void); /* min(16, cores / ranks) for this process */
/* Threads the host-buffer entry points will use for rows of elem_size-byte elements (1 or 4) with the current setting;
 * 0 = the raw rows are copied and packed on the device.  With the default setting the library packs on the host only
 * when that beats the raw copy: at least 10 threads for int32 rows, at least 8 for byte rows. */
int qldpc_decoder_host_threads_in_use(qldpc_decoder* dec, int elem_size);
/* Launch geometry in use: out[0..7] = vec (-1: global-memory path), threads, ctas_per_sm, grid, dyn_smem_bytes, regs,
 * SM count, frames per launch. */
int qldpc_decoder_launch_info(qldpc_decoder* dec, int side, int32_t out[8]);

/* Decoder::Decode / DecoderCPU::Decode on a batch of frames (Decoder.h:40-43, DecoderCPU.h:317-390;
 * the stubbed DecoderGPU::Decode is DecoderGPU.h:136-191).  Host buffers, one byte per bit:
 *   synX [nframes x mX], synZ [nframes x mZ]  ->  outX, outZ [nframes x n], outFlags [nframes] (ErrorCode bits 0..3),
 *   outIters [nframes x 2] executed iterations X,Z (may be NULL).  H2D / D2H copies are part of the call. */
int qldpc_decode_batch(qldpc_decoder* dec, const uint8_t* synX, const uint8_t* synZ, int64_t nframes,
                       float errorProbability, int maxIterations, uint8_t* outX, uint8_t* outZ, uint8_t* outFlags,
                       uint32_t* outIters);
/* Same with DEVICE pointers and bit-packed rows (LSB-first 32-bit words): synX [nframes x ceil(mX/32)], ...,
 * outX/outZ [nframes x ceil(n/32)]; outFlags, outIters device pointers (outIters may be NULL). */
int qldpc_decode_batch_device(qldpc_decoder* dec, const uint32_t* d_synX, const uint32_t* d_synZ, int64_t nframes,
                              float errorProbability, int maxIterations, uint32_t* d_outX, uint32_t* d_outZ,
                              uint8_t* d_outFlags, uint32_t* d_outIters);

/* DecoderCPU::GetStatistics(errorWeight, numErrors, errorProbability, maxIterations, seed), DecoderCPU.h:392-530
 * (pure virtual Decoder.h:44-47; stub DecoderGPU.h:230-273): the reference's fixed-weight-W error model, drawn
 * from one std::mt19937 stream with the distribution mapping the published results were made with (MSVC).
 * Exactly numErrors frames are tested (the reference tests (numErrors / nThreads) * nThreads).
 * counters[QLDPC_NUM_COUNTERS]; perFrameFlags [numErrors] and perFrameIters [numErrors x 2] may be NULL. */
int qldpc_get_statistics_weightw(qldpc_decoder* dec, int errorWeight, int64_t numErrors, float errorProbability,
                                 int maxIterations, uint32_t seed, uint64_t* counters, uint8_t* perFrameFlags,
                                 uint32_t* perFrameIters);

/* North-star replacement of the error model: depolarizing(p) noise generated ON DEVICE by counter-based
 * Philox4x32-10 keyed by (seed, global frame id, qubit); frames [first_frame, first_frame + nframes).
 * Results are independent of batch shape and GPU count.  BP prior stays 2/3*p (DecoderCPU.h:259). */
int qldpc_get_statistics_depolarizing(qldpc_decoder* dec, uint64_t seed, uint64_t first_frame, int64_t nframes,
                                      float p, int maxIterations, uint64_t* counters, uint8_t* perFrameFlags,
                                      uint32_t* perFrameIters);

/* DecoderGPU::GetStats(errorWeight, numErrors, errorProbability, maxIterations, seed, xErrors, zErrors),
 * DecoderGPU.h:193-228: pre-generated error patterns, frame-major [numErrors x n], from HOST memory.
 * `_i32` takes the reference's layout (one int per bit); `_u8` one byte per bit. */
int qldpc_get_stats_from_errors_i32(qldpc_decoder* dec, const int32_t* xErrors, const int32_t* zErrors,
                                    int64_t numErrors, float errorProbability, int maxIterations, uint64_t* counters,
                                    uint8_t* perFrameFlags, uint32_t* perFrameIters);
int qldpc_get_stats_from_errors_u8(qldpc_decoder* dec, const uint8_t* xErrors, const uint8_t* zErrors,
                                   int64_t numErrors, float errorProbability, int maxIterations, uint64_t* counters,
                                   uint8_t* perFrameFlags, uint32_t* perFrameIters);

/* ---- measurement -------------------------------------------------------------------------------------- */

/* Per-kernel device timing with CUDA events on the launching stream (off by default: enabling it adds two
 * event records per launch).  Kernel classes: */
enum {
  QLDPC_T_GENERATE = 0, /* Philox error generator                 */
  QLDPC_T_SYNDROME = 1, /* sparse syndrome                        */
  QLDPC_T_BP_X = 2,     /* BP tile kernel, X side                 */
  QLDPC_T_BP_Z = 3,     /* BP tile kernel, Z side                 */
  QLDPC_T_STATS = 4,    /* residual / logical check / counters    */
  QLDPC_T_PACK = 5,     /* pack / unpack / flag merge             */
  QLDPC_NUM_TIMERS = 6
};
int qldpc_decoder_enable_timing(qldpc_decoder* dec, int on);
/* Accumulated since the last reset: ms[QLDPC_NUM_TIMERS] device milliseconds and launches[QLDPC_NUM_TIMERS]
 * kernel launches per class (launches are counted whether or not timing is enabled).  reset != 0 clears them. */
int qldpc_decoder_get_timing(qldpc_decoder* dec, double* ms, uint64_t* launches, int reset);

/* ---- parity taps (tests) ------------------------------------------------------------------------------- */

/* Test taps for the host-side marshalling (host_pack.h); they run without a GPU.  pack: dst[r][w] bit b =
 * (src[r][32 w + b] != 0), rows of ceil(cols/32) words; unpack: the inverse, one byte per bit. */
int qldpc_debug_weightw_patterns(uint32_t seed, int errorWeight, int n, int64_t nframes, int threads, uint32_t* xWords,
                                 uint32_t* zWords); /* the weight-W error stream of qldpc_get_statistics_weightw, packed rows */
int qldpc_debug_host_pack(const void* src, int elem_size, int64_t rows, int cols, uint32_t* dst, int threads);
int qldpc_debug_host_unpack(const uint32_t* src, int64_t rows, int cols, uint8_t* dst, int threads);
/* Streaming read of a host buffer with `threads` of the packer's worker threads (its access pattern without the
 * arithmetic), best of `repeats`: the measured host-memory bandwidth behind bench.py's e2e.host_mem_roofline. */
int qldpc_debug_host_read_gbs(const void* src, int64_t bytes, int threads, int repeats, double* gbs);
/* Device Philox generator + syndrome kernel, unpacked to host bytes: xerr, zerr [nframes x n],
 * synX [nframes x mX], synZ [nframes x mZ] (any may be NULL). */
int qldpc_debug_generate(qldpc_decoder* dec, uint64_t seed, uint64_t first_frame, int64_t nframes, float p,
                         uint8_t* xerr, uint8_t* zerr, uint8_t* synX, uint8_t* synZ);
/* Runs the production BP kernel on `nframes` syndromes of one side ([nframes x m] bytes) and copies out the
 * messages after every iteration, check-major (edge = e*dc + i) as the oracle does:
 * q_trace, r_trace [nframes x cap_iters x E] (rows beyond the executed iterations are left untouched),
 * iters [nframes]. */
int qldpc_debug_bp_trace(qldpc_decoder* dec, int side, const uint8_t* syn, int nframes, float errorProbability,
                         int maxIterations, int cap_iters, float* q_trace, float* r_trace, uint32_t* iters);

/* Checks the branch-free division used by the BP kernel against IEEE division (__fdiv_rn) on `npairs` random
 * operand pairs 0 <= x <= y: out[0] = mismatching pairs (must be 0), out[1] = pairs the fast path defers to
 * __fdiv_rn (tiny numerators / denormal denominators), out[2] = pairs with x == 0. */
int qldpc_debug_division_check(qldpc_decoder* dec, uint64_t seed, int64_t npairs, uint64_t out[3]);

#ifdef __cplusplus
}
#endif
#endif /* QLDPC_B200_H */
