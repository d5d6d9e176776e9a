// Host-side code object behind `qldpc_code` (include/qldpc_b200.h).
// Reference semantics: Quantum_LDPC_Code (QEC_LDPC/Quantum_LDPC_Code.h:7-150) and the commented-out
// QC_LDPC_CSS constructor (QEC_LDPC/QEC_LDPC_CSS.cu:37-131).  Holds packed edge tables instead of the
// reference's dense int matrices; dense views are materialised on request only.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace qldpc {

struct SideTables {
  int m = 0, dc = 0, dv = 0, E = 0;
  std::vector<int32_t> chk_var;   // CSR  [m*dc]  neighbours of check e, ascending variable index
  std::vector<int32_t> var_chk;   // CSC  [n*dv]  neighbours of variable v, ascending check index
  std::vector<int32_t> var_edge;  //      [n*dv]  check-major edge id (e*dc + i) of the k-th edge of v
  std::vector<int32_t> hexp;      // circulant exponents [rows/P x L] when the side is quasi-cyclic
};

// A GF(2) matrix with bit-packed rows (LSB-first 32-bit words).
struct BitMatrix {
  int rows = 0, cols = 0, words = 0;
  std::vector<uint32_t> w;  // [rows x words]
  BitMatrix() {}
  BitMatrix(int r, int c) : rows(r), cols(c), words((c + 31) / 32), w((size_t)r * ((c + 31) / 32), 0u) {}
  bool get(int r, int c) const { return (w[(size_t)r * words + (c >> 5)] >> (c & 31)) & 1u; }
  void set(int r, int c) { w[(size_t)r * words + (c >> 5)] |= 1u << (c & 31); }
  uint32_t* row(int r) { return &w[(size_t)r * words]; }
  const uint32_t* row(int r) const { return &w[(size_t)r * words]; }
};

// Row-reduce in place to a basis of the row space; returns the rank (rows is shrunk to it).
int row_reduce(BitMatrix& a);
// Basis of the right null space {u : a u = 0}.
BitMatrix null_space(const BitMatrix& a);

struct Code {
  int J = 0, K = 0, L = 0, P = 0, sigma = 0, tau = 0, n = 0;
  SideTables side[2];
  bool is_qc = false;
  bool logical_from_file = false;
  // Logical check in use: e = [x-part | z-part] is a logical error iff some row has odd overlap with it
  // (Quantum_LDPC_Code::CheckLogicalError, Quantum_LDPC_Code.h:126-142).  Row-reduced; rows supported on the
  // x-part only come first (lx of them), then z-part-only rows (lz), then mixed rows (lm).
  BitMatrix logical;  // [lx+lz+lm x 2n], column c < n = x bit c, column n + c = z bit c
  int lx = 0, lz = 0, lm = 0;
  // iMinusP exactly as supplied (file line 4 / caller), kept for write_file round trips; empty if generated.
  BitMatrix iminusp_raw;

  std::string name() const;  // Quantum_LDPC_Code.h:145-150
  void dense_pcm(int s, int32_t* out) const;
  bool is_css() const;
  void syndrome(int s, const int32_t* err, int32_t* syn) const;
  bool check_logical(const int32_t* err2n) const;
};

// Throws std::string on invalid input (caught at the C ABI).
Code* code_from_qc(int J, int K, int L, int P, int sigma, int tau);
Code* code_from_dense(int J, int K, int L, int P, int sigma, int tau, const int32_t* pcmX, const int32_t* pcmZ,
                      const int32_t* iMinusP);
Code* code_from_file(const std::string& path);
void code_write_file(const Code& c, const std::string& path);
void qc_exponents(int J, int K, int L, int P, int sigma, int tau, std::vector<int32_t>& hHC, std::vector<int32_t>& hHD);

}  // namespace qldpc
