// Host-side checks of the C++ class surface (no GPU needed): code construction, file round trip, helper methods and
// the results-file text format.  Prints lines the Python test compares with golden data.
#include <iostream>
#include <sstream>

#include "CodeStatistics.h"
#include "QC_LDPC_CSS.h"
#include "RandomErrorGenerator.h"

int main(int argc, char** argv) {
  try {
    QC_LDPC_CSS code(4, 5, 10, 61, 9, 49);
    std::cout << "name " << code << std::endl;
    std::cout << "dims " << code.n << " " << code.numEqsX << " " << code.numEqsZ << " " << code.pcmX.num_rows << "x"
              << code.pcmX.num_cols << " " << code.iMinusP.num_rows << "x" << code.iMinusP.num_cols << std::endl;
    IntArray2d_h hc = code.exponentsX();
    std::cout << "hHC0";
    for (int l = 0; l < code.L; ++l) std::cout << " " << hc(0, l);
    std::cout << std::endl;
    // syndrome of a single X error on qubit 0 has weight J; a row of pcmX is a stabilizer (never a logical error)
    IntArray1d_h e(code.n, 0);
    e[0] = 1;
    int wt = 0;
    for (int s : code.GetSyndromeX(e)) wt += s;
    std::cout << "syndrome_weight " << wt << std::endl;
    IntArray1d_h stab(2 * code.n, 0);
    for (int v = 0; v < code.n; ++v) stab[v] = code.pcmX(3, v);
    std::cout << "row_is_logical " << code.CheckLogicalError(stab) << " single_is_logical ";
    IntArray1d_h single(2 * code.n, 0);
    single[5] = 1;
    std::cout << code.CheckLogicalError(single) << std::endl;
    // results-file text
    CodeStatistics s = {code, 655687811u, 1000u, 1000u, 1000u, 30u, 949u, 17u, 33u, 1u, 0u, 0u, 1258131ll};
    std::cout << "BEGIN_STATS" << std::endl << s << "END_STATS" << std::endl;
    // weight-W generator replays the reference's stream
    RandomErrorGenerator gen(code.n, 655687811u);
    std::vector<int> x(code.n, 0), z(code.n, 0);
    gen.GenerateError(x, z, 30);
    std::cout << "first_error";
    for (int v = 0; v < code.n; ++v)
      if (x[v] || z[v]) std::cout << " " << v << (x[v] && z[v] ? "Y" : x[v] ? "X" : "Z");
    std::cout << std::endl;
    if (argc > 1) {  // file round trip
      code.writeFile(argv[1]);
      Quantum_LDPC_Code back = Quantum_LDPC_Code::createFromFile(argv[1]);
      std::cout << "roundtrip " << (back.pcmX.values == code.pcmX.values && back.pcmZ.values == code.pcmZ.values) << " "
                << back << std::endl;
      try {
        Quantum_LDPC_Code::createFromFile(std::string(argv[1]) + ".missing");
      } catch (std::string& msg) {
        std::cout << "missing: " << msg << std::endl;
      }
    }
  } catch (std::string& s) {
    std::cout << "EXCEPTION " << s << std::endl;
    return 1;
  }
  return 0;
}
