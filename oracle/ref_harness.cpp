// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// Compiles the UNMODIFIED reference CPU decoder (DecoderCPU.h, Quantum_LDPC_Code.h, Decoder.h,
// CodeStatistics.h) from where it lies under /root/reference, by include path, into
// oracle/_ref/libqldpc_ref.so with a small C ABI so that tests and bench.py's reference arm can drive it
// through ctypes.  No reference source is copied into this repository.  Recipe: oracle/Makefile.
//
// Glue needed:
//  * oracle/shim/cusp, oracle/shim/thrust : container stand-ins for the un-vendored cusp 0.6.
//  * `static enum ErrorCode` (Decoder.h:14) is MSVC-only -> `#define static` around that include.
//  * The results files (K1..K5) were produced with MSVC's uniform_int_distribution; its mapping is
//    restated below (rejection + modulo) and substituted by macro so the reference's own
//    GetStatistics (DecoderCPU.h:392-537) reproduces the published counters bit for bit.
//  * `#define private public` exposes EqNodeUpdate / VarNodeUpdate / CheckConvergence / node arrays
//    for the per-iteration message taps (qref_bp_trace).
#include <random>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <chrono>

namespace std {
// MSVC-style uniform_int mapping for a 32-bit engine: accept u iff it falls below the largest
// multiple of R, return u % R.  Needed only to reproduce the reference's published results files.
template <class T>
struct msvc_uniform_int {
  T lo, hi;
  msvc_uniform_int(T a = 0, T b = 9) : lo(a), hi(b) {}
  template <class Engine>
  T operator()(Engine& eng) {
    const uint32_t R = (uint32_t)(hi - lo) + 1u;
    for (;;) {
      const uint32_t u = (uint32_t)eng();
      if (u / R < 0xFFFFFFFFu / R || 0xFFFFFFFFu % R == R - 1) return (T)(u % R) + lo;
    }
  }
};
}  // namespace std

#include "QEC_LDPC/Quantum_LDPC_Code.h"
#define static
#include "QEC_LDPC/Decoder.h"
#undef static
#define uniform_int_distribution msvc_uniform_int
#define private public
#include "QEC_LDPC/DecoderCPU.h"
#undef private
#undef uniform_int_distribution

namespace {
struct SilenceCout {
  std::streambuf* old;
  std::ostringstream sink;
  SilenceCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~SilenceCout() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

void* qref_code_from_file(const char* path) {
  SilenceCout q;
  try {
    return new Quantum_LDPC_Code(Quantum_LDPC_Code::createFromFile(path));
  } catch (std::string&) {
    return nullptr;
  }
}
void qref_code_free(void* c) { delete (Quantum_LDPC_Code*)c; }

void qref_code_dims(void* c, int out[9]) {
  auto* code = (Quantum_LDPC_Code*)c;
  int v[9] = {code->J, code->K, code->L, code->P, code->sigma, code->tau, code->n, code->numEqsX, code->numEqsZ};
  memcpy(out, v, sizeof v);
}

// which: 0 pcmX (mX x n), 1 pcmZ (mZ x n), 2 iMinusP (2n x 2n); row-major ints
void qref_code_dense(void* c, int which, int* out) {
  auto* code = (Quantum_LDPC_Code*)c;
  const IntArray2d_h& m = which == 0 ? code->pcmX : which == 1 ? code->pcmZ : code->iMinusP;
  std::copy(m.values.begin(), m.values.end(), out);
}

void qref_code_name(void* c, char* out, int cap) {
  std::ostringstream s;
  s << *(Quantum_LDPC_Code*)c;
  strncpy(out, s.str().c_str(), cap - 1);
  out[cap - 1] = 0;
}

void* qref_decoder_create(void* c) { return new DecoderCPU(*(Quantum_LDPC_Code*)c); }
void qref_decoder_free(void* d) { delete (DecoderCPU*)d; }

// The reference's own Monte-Carlo loop (weight-W errors from mt19937), DecoderCPU.h:392-530.
// out[0..9] = numErrorsTested, numXErrorsTested, numZErrorsTested, errorWeight, corrected,
//             syndromeErrorsX, syndromeErrorsZ, logicalErrors, convergenceFailX, convergenceFailZ
long long qref_get_statistics(void* d, int W, int COUNT, float p, int maxit, unsigned seed, int nthreads,
                              unsigned out[10]) {
  SilenceCout q;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  CodeStatistics s = ((DecoderCPU*)d)->GetStatistics(W, COUNT, p, maxit, seed);
  unsigned v[10] = {s.numErrorsTested, s.numXErrorsTested, s.numZErrorsTested, s.errorWeight, s.corrected,
                    s.syndromeErrorsX, s.syndromeErrorsZ, s.logicalErrors, s.convergenceFailX, s.convergenceFailZ};
  memcpy(out, v, sizeof v);
  return s.durationMicroSeconds;
}

// The results-file text of a CodeStatistics record (CodeStatistics.h:22-37), for the writer KAT.
int qref_format_statistics(void* c, const unsigned v[10], unsigned seed, long long dur, char* out, int cap) {
  CodeStatistics s = {*(Quantum_LDPC_Code*)c, seed, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], dur};
  std::ostringstream os;
  os << s;
  strncpy(out, os.str().c_str(), cap - 1);
  out[cap - 1] = 0;
  return (int)os.str().size();
}

int qref_decode(void* d, const int* synX, const int* synZ, float p, int maxit, int* outX, int* outZ) {
  auto* dec = (DecoderCPU*)d;
  const int n = dec->_numVars;
  std::vector<int> sx(synX, synX + dec->_numEqsX), sz(synZ, synZ + dec->_numEqsZ), ox(n, 0), oz(n, 0);
  int code = (int)dec->Decode(sx, sz, p, maxit, ox, oz);
  std::copy(ox.begin(), ox.end(), outX);
  std::copy(oz.begin(), oz.end(), outZ);
  return code;
}

void qref_syndrome(void* c, int side, const int* err, int* out) {
  auto* code = (Quantum_LDPC_Code*)c;
  std::vector<int> e(err, err + code->n);
  IntArray1d_h s = side == 0 ? code->GetSyndromeX(e) : code->GetSyndromeZ(e);
  std::copy(s.begin(), s.end(), out);
}

int qref_check_logical(void* c, const int* err2n) {
  auto* code = (Quantum_LDPC_Code*)c;
  std::vector<int> e(err2n, err2n + 2 * code->n);
  return code->CheckLogicalError(e) ? 1 : 0;
}

// Per-iteration taps: drives the reference's own EqNodeUpdate / VarNodeUpdate / CheckConvergence in the
// order BeliefPropogation does (DecoderCPU.h:259-291) and copies out the edge messages after every
// iteration, check-major (edge = eq*dc + i, i-th neighbour in ascending variable order):
//   r_out[it][edge] = eqNodes[eq*numVars+var]   (check -> variable, after EqNodeUpdate of iteration it)
//   q_out[it][edge] = varNodes[var*numEqs+eq]   (variable -> check, after VarNodeUpdate of iteration it)
// Returns the number of iterations executed; conv_out[it] = result of the n%10==0 convergence test (-1 if not run).
int qref_bp_trace(void* d, int side, const int* syn, float errorProbability, int maxit, int cap_iters, float* q_out,
                  float* r_out, int* conv_out) {
  auto* dec = (DecoderCPU*)d;
  const int numVars = dec->_numVars;
  const int numEqs = side == 0 ? dec->_numEqsX : dec->_numEqsZ;
  const int dc = side == 0 ? dec->_numVarsPerEqX : dec->_numVarsPerEqZ;
  const int dv = side == 0 ? dec->_numEqsPerVarX : dec->_numEqsPerVarZ;
  auto& varNodes = side == 0 ? dec->_varNodesX : dec->_varNodesZ;
  auto& eqNodes = side == 0 ? dec->_eqNodesX : dec->_eqNodesZ;
  auto& eqIdx = side == 0 ? dec->_eqNodeVarIndicesX : dec->_eqNodeVarIndicesZ;
  auto& varIdx = side == 0 ? dec->_varNodeEqIndicesX : dec->_varNodeEqIndicesZ;
  auto& eqPtrs = side == 0 ? dec->_eqNodeVarPtrsX : dec->_eqNodeVarPtrsZ;
  auto& varPtrs = side == 0 ? dec->_varNodeEqPtrsX : dec->_varNodeEqPtrsZ;
  std::vector<int> s(syn, syn + numEqs);
  float p = 2.0f / 3.0f * errorProbability;
  std::fill(varNodes.begin(), varNodes.end(), 0.0f);
  DecoderCPU::InitVarNodes(varNodes, eqIdx, p, dc, numEqs);
  const int E = numEqs * dc;
  bool converge = false;
  int it = 0;
  for (int n = 0; n < maxit; ++n) {
    if (converge) break;
    DecoderCPU::EqNodeUpdate(&eqNodes[0], eqPtrs, &eqIdx[0], &s[0], numEqs, numVars, dc);
    DecoderCPU::VarNodeUpdate(&varNodes[0], varPtrs, &varIdx[0], p, n == maxit - 1, numEqs, numVars, dv);
    int cv = -1;
    if (n % 10 == 0) {
      converge = DecoderCPU::CheckConvergence(&varNodes[0], 0.99f, 0.01f, numVars, numEqs);
      cv = converge ? 1 : 0;
    }
    if (it < cap_iters) {
      for (int e = 0; e < numEqs; ++e)
        for (int i = 0; i < dc; ++i) {
          int v = eqIdx[e * dc + i];
          if (r_out) r_out[(size_t)it * E + e * dc + i] = eqNodes[(size_t)e * numVars + v];
          if (q_out) q_out[(size_t)it * E + e * dc + i] = varNodes[(size_t)v * numEqs + e];
        }
      if (conv_out) conv_out[it] = cv;
    }
    ++it;
  }
  return it;
}

// Frame runner for supplied error patterns (depolarizing patterns generated elsewhere): the per-frame
// bookkeeping of GetStatistics (DecoderCPU.h:461-521) around the reference's public GetSyndromeX/Z, Decode and
// CheckLogicalError, one DecoderCPU per OpenMP thread as DecoderCPU.h:419-431 does.  xerr/zerr are
// [nframes x n] bytes (0/1).  counters[8] = xTested, zTested, corrected, synX, synZ, logical, cvX, cvZ.
// flags[f] = ErrorCode bits (Decoder.h:14-23) | 16 if logical error | 32 if corrected.
// outX/outZ (optional) = decoded patterns [nframes x n] bytes.  Returns wall-clock seconds of the frame loop.
double qref_run_frames(void* c, const unsigned char* xerr, const unsigned char* zerr, int nframes, float p, int maxit,
                       int nthreads, unsigned long long counters[8], unsigned char* flags, unsigned char* outX,
                       unsigned char* outZ) {
  auto* code = (Quantum_LDPC_Code*)c;
  const int n = code->n;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  unsigned long long xT = 0, zT = 0, cor = 0, sX = 0, sZ = 0, lg = 0, cX = 0, cZ = 0;
  double seconds = 0;
#pragma omp parallel reduction(+ : xT, zT, cor, sX, sZ, lg, cX, cZ)
  {
    DecoderCPU decoder(*code);
    std::vector<int> xe(n), ze(n), xd(n), zd(n);
#pragma omp barrier
    auto t0 = std::chrono::high_resolution_clock::now();
#pragma omp for schedule(dynamic, 16)
    for (int f = 0; f < nframes; ++f) {
      bool anyx = false, anyz = false;
      for (int i = 0; i < n; ++i) {
        xe[i] = xerr[(size_t)f * n + i];
        ze[i] = zerr[(size_t)f * n + i];
        anyx |= xe[i] != 0;
        anyz |= ze[i] != 0;
      }
      std::fill(xd.begin(), xd.end(), 0);
      std::fill(zd.begin(), zd.end(), 0);
      auto sx = code->GetSyndromeX(xe);
      auto sz = code->GetSyndromeZ(ze);
      std::vector<int> sx1(sx.begin(), sx.end()), sz1(sz.begin(), sz.end());
      xT += anyx;
      zT += anyz;
      int ec = (int)decoder.Decode(sx1, sz1, p, maxit, xd, zd);
      unsigned char fl = (unsigned char)ec;
      bool dEX = ec & Decoder::SYNDROME_FAIL_X, dEZ = ec & Decoder::SYNDROME_FAIL_Z;
      sX += dEX;
      sZ += dEZ;
      if (!(dEX || dEZ)) {
        std::vector<int> errors(2 * n);
        for (int i = 0; i < n; ++i) {
          errors[i] = (xe[i] + xd[i]) % 2;
          errors[n + i] = (ze[i] + zd[i]) % 2;
        }
        if (code->CheckLogicalError(errors)) { lg++; fl |= 16; } else { cor++; fl |= 32; }
      }
      if (ec & Decoder::CONVERGENCE_FAIL_X) cX++;
      if (ec & Decoder::CONVERGENCE_FAIL_Z) cZ++;
      if (flags) flags[f] = fl;
      if (outX) for (int i = 0; i < n; ++i) outX[(size_t)f * n + i] = (unsigned char)xd[i];
      if (outZ) for (int i = 0; i < n; ++i) outZ[(size_t)f * n + i] = (unsigned char)zd[i];
    }
    auto t1 = std::chrono::high_resolution_clock::now();
#pragma omp master
    seconds = std::chrono::duration<double>(t1 - t0).count();
  }
  unsigned long long v[8] = {xT, zT, cor, sX, sZ, lg, cX, cZ};
  memcpy(counters, v, sizeof v);
  return seconds;
}

int qref_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
