// CodeStatistics: result record of a Monte-Carlo run and its text form, which is the reference's results-file format
// (QEC_LDPC/CodeStatistics.h:5-37).  Two extra fields carry the iteration sums the device reduction also returns;
// they are not printed, so records stay byte-compatible.
#pragma once
#include <ostream>

#include "Quantum_LDPC_Code.h"

struct CodeStatistics {
  Quantum_LDPC_Code code;
  unsigned int randSeed;
  unsigned int numErrorsTested;
  unsigned int numXErrorsTested;
  unsigned int numZErrorsTested;
  unsigned int errorWeight;
  unsigned int corrected;
  unsigned int syndromeErrorsX;
  unsigned int syndromeErrorsZ;
  unsigned int logicalErrors;
  unsigned int convergenceFailX;
  unsigned int convergenceFailZ;
  long long durationMicroSeconds;
  unsigned long long iterationsX = 0, iterationsZ = 0;  // sums of executed BP iterations (not printed)
};

inline std::ostream& operator<<(std::ostream& os, CodeStatistics const& s) {
  const struct { const char* label; long long value; } rows[] = {
      {"Rand Seed", (long long)s.randSeed},           {"Duration(micro-s)", s.durationMicroSeconds},
      {"Errors Tested", (long long)s.numErrorsTested}, {"Errors With X", (long long)s.numXErrorsTested},
      {"Errors With Z", (long long)s.numZErrorsTested}, {"Error Weight", (long long)s.errorWeight},
      {"Corrected", (long long)s.corrected},           {"Syndrome Errors X", (long long)s.syndromeErrorsX},
      {"Syndrome Errors Z", (long long)s.syndromeErrorsZ}, {"Logical Errors", (long long)s.logicalErrors},
      {"Convergence Fail X", (long long)s.convergenceFailX}, {"Convergence Fail Z", (long long)s.convergenceFailZ}};
  os << "Code: " << s.code << std::endl;
  for (const auto& r : rows) os << r.label << ": " << r.value << std::endl;
  return os;
}
