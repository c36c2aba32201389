/**
 * @file equal.hxx
 * @brief `util::equal(d_ptr, h_ptr, n, error_op, verbose)`: count the positions where a device
 * array differs from a host array under `error_op` (reference include/loops/util/equal.hxx:44-68).
 */
#pragma once

#include <cstddef>
#include <iomanip>
#include <iostream>
#include <limits>

#include <cuda_runtime.h>
#include <thrust/host_vector.h>

namespace loops {
namespace util {

namespace detail {
struct not_equal_t {
  template <typename a_t, typename b_t>
  bool operator()(const a_t& a, const b_t& b) const { return a != b; }
};
}  // namespace detail

template <typename type_t, typename comp_t = detail::not_equal_t>
std::size_t equal(const type_t* d_ptr, const type_t* h_ptr, const std::size_t n, comp_t error_op = comp_t(),
                  const bool verbose = false) {
  thrust::host_vector<type_t> got(n);
  if (n) cudaMemcpy(got.data(), d_ptr, n * sizeof(type_t), cudaMemcpyDeviceToHost);
  std::size_t bad = 0;
  for (std::size_t i = 0; i < n; ++i) {
    if (!error_op(got[i], h_ptr[i])) continue;
    if (verbose)
      std::cout << "Error[" << i << "]: " << std::setw(10) << std::fixed
                << std::setprecision(std::numeric_limits<type_t>::digits10) << got[i] << " != " << std::setw(10)
                << h_ptr[i] << std::endl;
    ++bad;
  }
  return bad;
}

}  // namespace util
}  // namespace loops
