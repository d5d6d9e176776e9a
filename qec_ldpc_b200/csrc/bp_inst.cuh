// Explicit instantiation of the BP tile kernel for one (check degree, variable degree) shape: tile widths 1/2/4, the
// three division-guard variants, each also with the per-iteration message taps of the parity tests (guard + 4), so that
// the taps run exactly the arithmetic of the production instantiation.  One translation unit per shape so that nvcc builds them in parallel.
// A shape may add instantiations with the number of checks as a compile-time constant (QLDPC_DEFINE_SHAPE_M): the
// check-phase addresses of such a kernel are immediates and its phase loops are unrolled for 128 threads per CTA; it is
// picked when the code's check count and the launch shape match.
#pragma once
#include "bp_kernel.cuh"

namespace qldpc {
typedef void (*BpKernel)(const BpArgs);

// guard: 0 / 1 / 3 as division_guard (decoder.cu) returns them; + 4 selects the same kernel with the message taps
template <int DC, int DV, int M>
BpKernel bp_kernel_for(int vec, int guard) {
#define QLDPC_V(V)                                                                  \
  if (vec == V) {                                                                   \
    if (guard == 0) return bp_tile_kernel<DC, DV, V, 0, M>;           \
    if (guard == 1) return bp_tile_kernel<DC, DV, V, 1, M>;           \
    if (guard == 3) return bp_tile_kernel<DC, DV, V, 3, M>;           \
    if (guard == 4) return bp_tile_kernel<DC, DV, V, 0, M, true>;     \
    if (guard == 5) return bp_tile_kernel<DC, DV, V, 1, M, true>;     \
    return bp_tile_kernel<DC, DV, V, 3, M, true>;                     \
  }
  QLDPC_V(4) QLDPC_V(2) QLDPC_V(1)
#undef QLDPC_V
  return nullptr;
}
}  // namespace qldpc

#define QLDPC_DEFINE_SHAPE(DC, DV) \
  namespace qldpc { BpKernel bp_shape_##DC##_##DV(int vec, int guard, int, int) { return bp_kernel_for<DC, DV, 0>(vec, guard); } }
#define QLDPC_DEFINE_SHAPE_M(DC, DV, M)                                       \
  namespace qldpc {                                                           \
  BpKernel bp_shape_##DC##_##DV(int vec, int guard, int m, int threads) {     \
    if (m == M && threads == 128) return bp_kernel_for<DC, DV, M>(vec, guard); \
    return bp_kernel_for<DC, DV, 0>(vec, guard);                              \
  }                                                                           \
  }
