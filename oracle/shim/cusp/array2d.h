// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the un-vendored cusp 0.6
// container library so that the reference headers under /root/reference compile
// unmodified with g++.  cusp is used by the reference purely as a container
// (HostDeviceArray.h:6-13); no arithmetic lives here.
#pragma once
#include <vector>
#include <iostream>
#include <sstream>
#include <algorithm>
#include <cstddef>
namespace cusp {
struct host_memory {};
struct device_memory {};
struct row_major {};
template <class T, class MemorySpace>
struct array1d : public std::vector<T> {
  using std::vector<T>::vector;
  array1d() {}
  array1d(const std::vector<T>& v) : std::vector<T>(v) {}
};
template <class T, class MemorySpace, class Orientation>
struct array2d {
  size_t num_rows = 0, num_cols = 0, num_entries = 0;
  array1d<T, MemorySpace> values;
  array2d() {}
  array2d(size_t r, size_t c) : num_rows(r), num_cols(c), num_entries(r * c), values(r * c) {}
  array2d(size_t r, size_t c, T v) : num_rows(r), num_cols(c), num_entries(r * c), values(r * c, v) {}
};
}  // namespace cusp
