// Abstract decoder interface of the reference (QEC_LDPC/Decoder.h:7-48): the boundary the decode path sits behind.
#pragma once
#include "CodeStatistics.h"
#include "HostDeviceArray.h"
#include "Quantum_LDPC_Code.h"

class Decoder {
 protected:
  Quantum_LDPC_Code _code;

 public:
  // Decoder.h:14-23 (the reference spells it `static enum`, an MSVC extension)
  enum ErrorCode {
    SUCCESS = 0,
    SYNDROME_FAIL_X = 1 << 0,
    SYNDROME_FAIL_Z = 1 << 1,
    SYNDROME_FAIL_XZ = SYNDROME_FAIL_X | SYNDROME_FAIL_Z,
    CONVERGENCE_FAIL_X = 1 << 2,
    CONVERGENCE_FAIL_Z = 1 << 3,
    CONVERGENCE_FAIL_XZ = CONVERGENCE_FAIL_X | CONVERGENCE_FAIL_Z
  };
  friend inline ErrorCode operator|(const ErrorCode& a, const ErrorCode& b) { return static_cast<ErrorCode>(int(a) | int(b)); }
  friend inline ErrorCode operator&(const ErrorCode& a, const ErrorCode& b) { return static_cast<ErrorCode>(int(a) & int(b)); }

  explicit Decoder(Quantum_LDPC_Code code) : _code(code) {}
  virtual ~Decoder() {}

  // Decoder.h:40-47
  virtual ErrorCode Decode(const IntArray1d_h& syndromeX, const IntArray1d_h& syndromeZ, float errorProbability,
                           int maxIterations, IntArray1d_h& outErrorsX, IntArray1d_h& outErrorsZ) = 0;
  virtual CodeStatistics GetStatistics(int errorWeight, int numErrors, float errorProbability, int maxIterations) = 0;
  virtual CodeStatistics GetStatistics(int errorWeight, int numErrors, float errorProbability, int maxIterations,
                                       unsigned int seed) = 0;
  const Quantum_LDPC_Code& code() const { return _code; }
};
