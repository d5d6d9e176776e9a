// Quantum_LDPC_Code with the reference's public surface (QEC_LDPC/Quantum_LDPC_Code.h:7-150), backed by the
// qldpc_code handle of the C ABI (include/qldpc_b200.h), which keeps packed CSR/CSC edge tables next to the dense
// matrices the reference exposes.  Copies share the handle (the reference copies the dense matrices by value into
// every decoder and every call, Decoder.h:32 / DecoderCPU.h:431).
#pragma once
#include <memory>
#include <ostream>
#include <string>

#include "HostDeviceArray.h"
#include "qldpc_b200.h"

class Quantum_LDPC_Code {
 protected:
  struct Owner {
    qldpc_code* h;
    explicit Owner(qldpc_code* p) : h(p) {}
    ~Owner() { qldpc_code_destroy(h); }
  };
  std::shared_ptr<Owner> _handle;

  static qldpc_code_info infoOf(qldpc_code* h) {
    qldpc_code_info i;
    qldpc_code_get_info(h, &i);
    return i;
  }
  static IntArray2d_h denseOf(qldpc_code* h, int which, int rows, int cols) {
    IntArray2d_h m((size_t)rows, (size_t)cols);
    if (rows > 0) qldpc_code_dense(h, which, m.values.data());
    return m;
  }
  Quantum_LDPC_Code(qldpc_code* h, const qldpc_code_info& i)
      : _handle(std::make_shared<Owner>(h)), J(i.J), K(i.K), L(i.L), P(i.P), sigma(i.sigma), tau(i.tau), n(i.n),
        numEqsX(i.mX), numEqsZ(i.mZ), pcmX(denseOf(h, 0, i.mX, i.n)), pcmZ(denseOf(h, 1, i.mZ, i.n)),
        iMinusP(denseOf(h, 3, 2 * i.n, 2 * i.n)) {}
  explicit Quantum_LDPC_Code(qldpc_code* h) : Quantum_LDPC_Code(h, infoOf(h)) {}

 public:
  const int J, K, L, P, sigma, tau;
  const int n;  // number of physical qubits = L*P
  const int numEqsX, numEqsZ;
  IntArray2d_h pcmX, pcmZ;
  // 2n x 2n like the reference's (Quantum_LDPC_Code.h:16): the iMinusP of the code file / constructor argument as it was
  // given, or -- for codes built from (J,K,L,P,sigma,tau) alone -- the generated logical-check rows (a basis of
  // ker(pcmX) and of ker(pcmZ): same kernel, hence the same decisions) followed by zero rows.
  IntArray2d_h iMinusP;

  // Quantum_LDPC_Code.h:26-80; throws std::string like the reference (:78).
  static Quantum_LDPC_Code createFromFile(std::string file) {
    qldpc_code* h = nullptr;
    if (qldpc_code_create_from_file(file.c_str(), &h) != QLDPC_OK) throw std::string(qldpc_last_error());
    return Quantum_LDPC_Code(h);
  }

  // Quantum_LDPC_Code.h:82-88 (dense 0/1 matrices; imp may be empty: the logical check is then generated).
  Quantum_LDPC_Code(int J, int K, int L, int P, int sigma, int tau, IntArray2d_h pcmX, IntArray2d_h pcmZ, IntArray2d_h imp)
      : Quantum_LDPC_Code(fromDense(J, K, L, P, sigma, tau, pcmX, pcmZ, imp)) {}

  void writeFile(const std::string& file) const {
    if (qldpc_code_write_file(_handle->h, file.c_str()) != QLDPC_OK) throw std::string(qldpc_last_error());
  }

  // Quantum_LDPC_Code.h:94-124
  IntArray1d_h GetSyndromeX(IntArray1d_h errors) const { return syndrome(0, errors, numEqsX); }
  IntArray1d_h GetSyndromeZ(IntArray1d_h errors) const { return syndrome(1, errors, numEqsZ); }
  // Quantum_LDPC_Code.h:126-142; errors = {x1..xn, z1..zn}
  bool CheckLogicalError(IntArray1d_h errors) const {
    if ((int)errors.size() != 2 * n) throw std::string("CheckLogicalError: expected 2n entries");
    return qldpc_code_check_logical(_handle->h, errors.data()) == 1;
  }

  bool isCSS() const { return qldpc_code_is_css(_handle->h) == 1; }
  qldpc_code* handle() const { return _handle->h; }
  std::string name() const {
    char buf[192];
    qldpc_code_name(_handle->h, buf, (int)sizeof buf);
    return buf;
  }

 private:
  static qldpc_code* fromDense(int J, int K, int L, int P, int sigma, int tau, const IntArray2d_h& x, const IntArray2d_h& z,
                               const IntArray2d_h& imp) {
    const size_t nn = (size_t)L * (size_t)P;
    auto shaped = [](const IntArray2d_h& m, size_t rows, size_t cols) {
      return m.num_rows == rows && m.num_cols == cols && m.values.size() == rows * cols;
    };
    if (J < 1 || K < 1 || L < 1 || P < 1) throw std::string("Quantum_LDPC_Code: J, K, L, P must be positive");
    if (!shaped(x, (size_t)J * P, nn)) throw std::string("Quantum_LDPC_Code: pcmX must be (J*P) x (L*P)");
    if (!shaped(z, (size_t)K * P, nn)) throw std::string("Quantum_LDPC_Code: pcmZ must be (K*P) x (L*P)");
    if (!imp.values.empty() && !shaped(imp, 2 * nn, 2 * nn)) throw std::string("Quantum_LDPC_Code: iMinusP must be (2*L*P) x (2*L*P)");
    qldpc_code* h = nullptr;
    const int32_t* ip = imp.values.empty() ? nullptr : imp.values.data();
    if (qldpc_code_create_dense(J, K, L, P, sigma, tau, x.values.data(), z.values.data(), ip, &h) != QLDPC_OK)
      throw std::string(qldpc_last_error());
    return h;
  }
  IntArray1d_h syndrome(int side, const IntArray1d_h& errors, int m) const {
    if ((int)errors.size() != n) throw std::string("GetSyndrome: expected n entries");
    IntArray1d_h s((size_t)m);
    qldpc_code_syndrome(_handle->h, side, errors.data(), s.data());
    return s;
  }
  friend class QC_LDPC_CSS;
};

// Quantum_LDPC_Code.h:145-150 (k is printed as numEqsZ - numEqsX there; results file names depend on it, main.cu:94)
inline std::ostream& operator<<(std::ostream& stream, Quantum_LDPC_Code const& code) { return stream << code.name(); }
