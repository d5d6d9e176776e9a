// QC_LDPC_CSS(J, K, L, P, sigma, tau): the Hagiwara-Imai quasi-cyclic CSS construction that is commented out in the
// reference (QEC_LDPC/QC_LDPC_CSS.h:146-156, formulas QEC_LDPC/QEC_LDPC_CSS.cu:37-131), restored as a code
// constructor.  Edge tables come from the closed-form circulant index functions; the logical check is generated.
#pragma once
#include "Quantum_LDPC_Code.h"

class QC_LDPC_CSS : public Quantum_LDPC_Code {
  static qldpc_code* build(int J, int K, int L, int P, int sigma, int tau) {
    qldpc_code* h = nullptr;
    if (qldpc_code_create_qc(J, K, L, P, sigma, tau, &h) != QLDPC_OK) throw std::string(qldpc_last_error());
    return h;
  }

 public:
  QC_LDPC_CSS(int J, int K, int L, int P, int sigma, int tau) : Quantum_LDPC_Code(build(J, K, L, P, sigma, tau)) {}

  // Circulant exponent matrices hHC (J x L) and hHD (K x L), QEC_LDPC_CSS.cu:43-90.
  IntArray2d_h exponentsX() const { return exps(0, J); }
  IntArray2d_h exponentsZ() const { return exps(1, K); }

 private:
  IntArray2d_h exps(int side, int rows) const {
    IntArray2d_h m((size_t)rows, (size_t)L);
    qldpc_code_exponents(handle(), side, m.values.data());
    return m;
  }
};
