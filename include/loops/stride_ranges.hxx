/**
 * @file stride_ranges.hxx
 * @brief Grid-/block-/custom-stride ranges (reference
 * include/loops/stride_ranges.hxx:16-62).
 */
#pragma once
#include <loops/range.hxx>

namespace loops {

/// Thread g = blockIdx.x*blockDim.x+threadIdx.x visits begin+g, then strides
/// by the whole grid.
template <typename T>
LOOPS_D step_range<T> grid_stride_range(T begin, T end) {
  const T lane = static_cast<T>(blockDim.x * blockIdx.x + threadIdx.x);
  return step_range<T>(begin + lane, end,
                       static_cast<T>(gridDim.x * blockDim.x));
}

/// Same start for every thread, stride blockDim.x.
template <typename T>
LOOPS_D step_range<T> block_stride_range(T begin, T end) {
  return step_range<T>(begin, end, static_cast<T>(blockDim.x));
}

template <typename T>
LOOPS_D step_range<T> custom_stride_range(T begin, T end, T stride) {
  return step_range<T>(begin, end, stride);
}

}  // namespace loops
