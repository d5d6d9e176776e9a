// TEST INFRASTRUCTURE ONLY (oracle/): the reference calls thrust::fill on std::vector
// (DecoderCPU.h:441-444); this stand-in keeps the CPU oracle free of the CUDA toolkit.
#pragma once
#include <algorithm>
namespace thrust {
template <class It, class T>
void fill(It first, It last, const T& v) { std::fill(first, last, v); }
}  // namespace thrust
