// The decoder interface a driver programs against: same type, member and enumerator names as the reference's
// abstract class (QEC_LDPC/Decoder.h:7-48), so code written for it compiles unchanged; the implementation behind it in
// this framework is DecoderGPU (DecoderGPU.h), which runs on the sm_100a kernels through include/qldpc_b200.h.
//
//   Decode(...)          one frame: X- and Z-syndromes in, X- and Z-corrections out, ErrorCode bit mask returned
//   GetStatistics(...)   Monte-Carlo run over `numErrors` random error patterns of weight `errorWeight`
//
// ErrorCode is a bit mask.  A frame can fail the syndrome check (the correction does not reproduce the measured
// syndrome) and/or the convergence check (some message is still inside (0.01, 0.99)) independently on either side;
// the two are reported separately because a non-converged frame may still decode correctly (DecoderCPU.h:511-521).
#pragma once
#include "CodeStatistics.h"
#include "HostDeviceArray.h"
#include "Quantum_LDPC_Code.h"

class Decoder {
 public:
  // Values as in Decoder.h:14-23 (declared there as `static enum`, an MSVC extension).
  enum ErrorCode {
    SUCCESS = 0,
    SYNDROME_FAIL_X = 1,      // bit 0
    SYNDROME_FAIL_Z = 2,      // bit 1
    SYNDROME_FAIL_XZ = 3,
    CONVERGENCE_FAIL_X = 4,   // bit 2
    CONVERGENCE_FAIL_Z = 8,   // bit 3
    CONVERGENCE_FAIL_XZ = 12
  };

  explicit Decoder(Quantum_LDPC_Code code) : _code(code) {}
  virtual ~Decoder() {}

  // Not pure, like the reference's (Decoder.h:40-43): the base does nothing and reports SUCCESS.
  virtual ErrorCode Decode(const IntArray1d_h& syndromeX, const IntArray1d_h& syndromeZ, float errorProbability,
                           int maxIterations, IntArray1d_h& outErrorsX, IntArray1d_h& outErrorsZ) {
    (void)syndromeX; (void)syndromeZ; (void)errorProbability; (void)maxIterations; (void)outErrorsX; (void)outErrorsZ;
    return SUCCESS;
  }

  // Seeded and unseeded forms (the unseeded one draws its seed from std::random_device, DecoderCPU.h:532-537).
  virtual CodeStatistics GetStatistics(int errorWeight, int numErrors, float errorProbability, int maxIterations,
                                       unsigned int seed) = 0;
  virtual CodeStatistics GetStatistics(int errorWeight, int numErrors, float errorProbability, int maxIterations) = 0;

  const Quantum_LDPC_Code& code() const { return _code; }

 protected:
  Quantum_LDPC_Code _code;  // held by value like the reference (Decoder.h:11); copies share the native handle
};

// Flag algebra on ErrorCode (Decoder.h:25-30).
inline Decoder::ErrorCode operator|(Decoder::ErrorCode lhs, Decoder::ErrorCode rhs) {
  return Decoder::ErrorCode(static_cast<int>(lhs) | static_cast<int>(rhs));
}
inline Decoder::ErrorCode operator&(Decoder::ErrorCode lhs, Decoder::ErrorCode rhs) {
  return Decoder::ErrorCode(static_cast<int>(lhs) & static_cast<int>(rhs));
}
