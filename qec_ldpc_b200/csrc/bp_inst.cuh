// Explicit instantiation of the BP tile kernel for one (check degree, variable degree) shape: tile widths 1/2/4 and
// the three division-guard variants.  One translation unit per shape so that nvcc builds them in parallel.
#pragma once
#include "bp_kernel.cuh"

namespace qldpc {
typedef void (*BpKernel)(const BpArgs);

template <int DC, int DV>
BpKernel bp_kernel_for(int vec, int guard) {
#define QLDPC_V(V)                                                                  \
  if (vec == V) {                                                                   \
    if (guard == 0) return bp_tile_kernel<DC, DV, V, 0>;              \
    if (guard == 1) return bp_tile_kernel<DC, DV, V, 1>;              \
    return bp_tile_kernel<DC, DV, V, 3>;                              \
  }
  QLDPC_V(4) QLDPC_V(2) QLDPC_V(1)
#undef QLDPC_V
  return nullptr;
}
}  // namespace qldpc

#define QLDPC_DEFINE_SHAPE(DC, DV) \
  namespace qldpc { BpKernel bp_shape_##DC##_##DV(int vec, int guard) { return bp_kernel_for<DC, DV>(vec, guard); } }
