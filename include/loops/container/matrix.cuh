/**
 * @file matrix.cuh
 * @brief Dense row-major matrix (reference include/loops/container/matrix.cuh):
 * owning storage plus a raw pointer that survives by-value copies into kernels;
 * `operator()(r, c)` and `operator[](i)` address `cols * r + c`.
 */
#pragma once

#include <cstddef>

#include <loops/container/vector.hxx>
#include <loops/memory.hxx>

namespace loops {
using namespace memory;

template <typename value_t, memory_space_t space = memory_space_t::device>
struct matrix_t {
  std::size_t rows = 0, cols = 0;
  vector_t<value_t, space> m_data;
  value_t* m_data_ptr = nullptr;

  matrix_t() = default;
  matrix_t(std::size_t r, std::size_t c)
      : rows(r), cols(c), m_data(r * c), m_data_ptr(thrust::raw_pointer_cast(m_data.data())) {}

  /// Shallow copy (what a kernel receives): same storage, no ownership.
  __host__ __device__ matrix_t(const matrix_t<value_t, space>& other)
      : rows(other.rows), cols(other.cols), m_data_ptr(other.m_data_ptr) {}

  __host__ __device__ __forceinline__ value_t operator()(int r, int c) const { return m_data_ptr[cols * r + c]; }
  __host__ __device__ __forceinline__ value_t& operator()(int r, int c) { return m_data_ptr[cols * r + c]; }
  __host__ __device__ __forceinline__ value_t operator[](std::size_t i) const { return m_data_ptr[i]; }
  __host__ __device__ __forceinline__ value_t& operator[](std::size_t i) { return m_data_ptr[i]; }
};

}  // namespace loops
