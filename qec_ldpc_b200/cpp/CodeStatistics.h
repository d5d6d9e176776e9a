// Result record of one Monte-Carlo run.  Field names, order and the text form follow the reference
// (QEC_LDPC/CodeStatistics.h:5-37) because (a) drivers brace-initialise the struct positionally
// (DecoderCPU.h:527-529) and (b) the text form IS the results-file format that the reference's checked-in results use
// (tests/test_cpp.py compares byte for byte).  The on-device reduction also returns the executed-iteration sums; they
// ride along in two trailing members that are not printed, so records stay format-compatible.
#pragma once
#include <ostream>

#include "Quantum_LDPC_Code.h"

struct CodeStatistics {
  Quantum_LDPC_Code code;             // the code the run was made on (printed through its operator<<)
  unsigned int randSeed;              // seed of the error stream
  unsigned int numErrorsTested;       // frames decoded
  unsigned int numXErrorsTested;      // frames whose pattern had at least one X component
  unsigned int numZErrorsTested;      // ... at least one Z component
  unsigned int errorWeight;           // W of the fixed-weight generator (0 for depolarizing runs)
  unsigned int corrected;             // no syndrome failure and residual is not a logical error
  unsigned int syndromeErrorsX;       // decision does not reproduce the X syndrome
  unsigned int syndromeErrorsZ;
  unsigned int logicalErrors;         // no syndrome failure, residual fails the logical check
  unsigned int convergenceFailX;      // counted independently of the outcome
  unsigned int convergenceFailZ;
  long long durationMicroSeconds;     // wall clock of the whole run
  unsigned long long iterationsX = 0, iterationsZ = 0;  // sums of executed BP iterations (extension, not printed)
};

inline std::ostream& operator<<(std::ostream& os, CodeStatistics const& s) {
  struct Row { const char* label; long long value; };
  const Row rows[] = {{"Rand Seed", s.randSeed},
                      {"Duration(micro-s)", s.durationMicroSeconds},
                      {"Errors Tested", s.numErrorsTested},
                      {"Errors With X", s.numXErrorsTested},
                      {"Errors With Z", s.numZErrorsTested},
                      {"Error Weight", s.errorWeight},
                      {"Corrected", s.corrected},
                      {"Syndrome Errors X", s.syndromeErrorsX},
                      {"Syndrome Errors Z", s.syndromeErrorsZ},
                      {"Logical Errors", s.logicalErrors},
                      {"Convergence Fail X", s.convergenceFailX},
                      {"Convergence Fail Z", s.convergenceFailZ}};
  os << "Code: " << s.code << '\n';
  for (const Row& r : rows) os << r.label << ": " << r.value << '\n';
  return os.flush();
}
