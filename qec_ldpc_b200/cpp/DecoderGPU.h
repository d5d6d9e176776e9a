// DecoderGPU: the reference's GPU decoder class (QEC_LDPC/DecoderGPU.h:11-281, a stub there: Decode has its BP removed
// :155-179 and GetStats launches nothing :220) implemented on the C ABI of include/qldpc_b200.h, i.e. on the
// hand-written sm_100a kernels.  Same method names, argument meaning and return types; errors of the library surface
// as std::string exceptions, the way the reference reports failures (Quantum_LDPC_Code.h:78, main.cu:106).
// There is no CPU decode path behind this class.
#pragma once
#include <chrono>
#include <random>
#include <string>
#include <vector>

#include "Decoder.h"

class DecoderGPU : public Decoder {
  struct Owner {
    qldpc_decoder* h;
    explicit Owner(qldpc_decoder* p) : h(p) {}
    ~Owner() { qldpc_decoder_destroy(h); }
  };
  std::shared_ptr<Owner> _dec;

  static void check(int rc) {
    if (rc != QLDPC_OK) throw std::string(qldpc_last_error());
  }
  static long long microsSince(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - t0).count();
  }
  CodeStatistics pack(const uint64_t* k, unsigned seed, int W, long long micros) const {
    CodeStatistics s = {_code, seed, (unsigned)k[QLDPC_C_FRAMES], (unsigned)k[QLDPC_C_XTESTED], (unsigned)k[QLDPC_C_ZTESTED],
                        (unsigned)W, (unsigned)k[QLDPC_C_CORRECTED], (unsigned)k[QLDPC_C_SYNX], (unsigned)k[QLDPC_C_SYNZ],
                        (unsigned)k[QLDPC_C_LOGICAL], (unsigned)k[QLDPC_C_CVX], (unsigned)k[QLDPC_C_CVZ], micros};
    s.iterationsX = k[QLDPC_C_ITERSX];
    s.iterationsZ = k[QLDPC_C_ITERSZ];
    return s;
  }

 public:
  // DecoderGPU.h:117-130.  device < 0: current CUDA device; maxFrames: frames per launch (0 = library default).
  explicit DecoderGPU(Quantum_LDPC_Code code, int device = -1, int maxFrames = 0) : Decoder(code) {
    qldpc_decoder* h = nullptr;
    check(qldpc_decoder_create(_code.handle(), device, maxFrames, &h));
    _dec = std::make_shared<Owner>(h);
  }
  qldpc_decoder* handle() const { return _dec->h; }

  // Decoder.h:40-43 / DecoderGPU.h:136-191 (the reference's derived classes take std::vector<int>, which is what
  // IntArray1d_h is here, so this both overrides the base and matches the derived signature).
  ErrorCode Decode(const IntArray1d_h& syndromeX, const IntArray1d_h& syndromeZ, float errorProbability, int maxIterations,
                   IntArray1d_h& outErrorsX, IntArray1d_h& outErrorsZ) override {
    std::vector<uint8_t> sx(syndromeX.begin(), syndromeX.end()), sz(syndromeZ.begin(), syndromeZ.end());
    if ((int)sx.size() != _code.numEqsX || (int)sz.size() != _code.numEqsZ) throw std::string("Decode: syndrome size mismatch");
    std::vector<uint8_t> ox((size_t)_code.n), oz((size_t)_code.n);
    uint8_t flags = 0;
    check(qldpc_decode_batch(_dec->h, sx.data(), sz.data(), 1, errorProbability, maxIterations, ox.data(), oz.data(), &flags,
                             nullptr));
    outErrorsX.assign(ox.begin(), ox.end());
    outErrorsZ.assign(oz.begin(), oz.end());
    return static_cast<ErrorCode>(flags & 15);
  }

  // Many frames at once: syndromes [nframes x numEqs] bytes -> decisions [nframes x n] bytes, ErrorCode per frame.
  void DecodeBatch(const std::vector<uint8_t>& syndromesX, const std::vector<uint8_t>& syndromesZ, long long nframes,
                   float errorProbability, int maxIterations, std::vector<uint8_t>& outErrorsX,
                   std::vector<uint8_t>& outErrorsZ, std::vector<uint8_t>& outCodes) {
    outErrorsX.resize((size_t)nframes * _code.n);
    outErrorsZ.resize((size_t)nframes * _code.n);
    outCodes.resize((size_t)nframes);
    check(qldpc_decode_batch(_dec->h, syndromesX.data(), syndromesZ.data(), nframes, errorProbability, maxIterations,
                             outErrorsX.data(), outErrorsZ.data(), outCodes.data(), nullptr));
  }

  // DecoderGPU.h:193-228: pre-generated error patterns [numErrors x n], one int per qubit.
  CodeStatistics GetStats(int errorWeight, int numErrors, float errorProbability, int maxIterations, int seed,
                          std::vector<int>& xErrors, std::vector<int>& zErrors) {
    auto t0 = std::chrono::high_resolution_clock::now();
    uint64_t k[QLDPC_NUM_COUNTERS];
    if (xErrors.size() < (size_t)numErrors * _code.n || zErrors.size() < (size_t)numErrors * _code.n)
      throw std::string("GetStats: error arrays smaller than numErrors x n");
    check(qldpc_get_stats_from_errors_i32(_dec->h, xErrors.data(), zErrors.data(), numErrors, errorProbability, maxIterations,
                                          k, nullptr, nullptr));
    return pack(k, (unsigned)seed, errorWeight, microsSince(t0));
  }

  // Decoder.h:44-47 / DecoderCPU.h:392-530 / DecoderGPU.h:230-273: fixed-weight-W errors from std::mt19937(seed).
  CodeStatistics GetStatistics(int errorWeight, int numErrors, float errorProbability, int maxIterations,
                               unsigned int seed) override {
    auto t0 = std::chrono::high_resolution_clock::now();
    uint64_t k[QLDPC_NUM_COUNTERS];
    check(qldpc_get_statistics_weightw(_dec->h, errorWeight, numErrors, errorProbability, maxIterations, seed, k, nullptr,
                                       nullptr));
    return pack(k, seed, errorWeight, microsSince(t0));
  }
  CodeStatistics GetStatistics(int errorWeight, int numErrors, float errorProbability, int maxIterations) override {
    std::random_device rd;  // DecoderCPU.h:532-537
    return GetStatistics(errorWeight, numErrors, errorProbability, maxIterations, rd());
  }

  // Depolarizing(p) noise generated on the device (counter-based Philox, frames firstFrame .. firstFrame+numFrames-1).
  // errorWeight is reported as 0.
  CodeStatistics GetStatisticsDepolarizing(long long numFrames, float p, int maxIterations, unsigned long long seed,
                                           unsigned long long firstFrame = 0, uint64_t* countersOut = nullptr) {
    auto t0 = std::chrono::high_resolution_clock::now();
    uint64_t k[QLDPC_NUM_COUNTERS];
    check(qldpc_get_statistics_depolarizing(_dec->h, seed, firstFrame, numFrames, p, maxIterations, k, nullptr, nullptr));
    if (countersOut)
      for (int i = 0; i < QLDPC_NUM_COUNTERS; ++i) countersOut[i] = k[i];
    return pack(k, (unsigned)seed, 0, microsSince(t0));
  }
};
