// Container aliases of the reference's class surface (QEC_LDPC/HostDeviceArray.h:6-13), without the cusp dependency:
// the reference uses cusp::array1d / array2d purely as containers, so plain std::vector-backed types stand in.
#pragma once
#include <cstddef>
#include <vector>

typedef std::vector<int> IntArray1d_h;

template <typename T>
struct Array2d_h {
  size_t num_rows = 0, num_cols = 0, num_entries = 0;
  std::vector<T> values;  // row-major
  Array2d_h() {}
  Array2d_h(size_t rows, size_t cols, T fill = T()) : num_rows(rows), num_cols(cols), num_entries(rows * cols), values(rows * cols, fill) {}
  T& operator()(size_t r, size_t c) { return values[r * num_cols + c]; }
  const T& operator()(size_t r, size_t c) const { return values[r * num_cols + c]; }
};
typedef Array2d_h<int> IntArray2d_h;
typedef Array2d_h<float> FloatArray2d_h;
