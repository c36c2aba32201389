/**
 * @file thread_mapped.cuh
 * @brief `loops::algorithms::spmm::thread_mapped(csr, B, C, stream)` -- same name
 * and arguments as the reference (include/loops/algorithms/spmm/thread_mapped.cuh:55-80),
 * a thin call into `loopsb_spmm_csr_f32` (warp-per-row kernel with coalesced reads of B).
 * Synchronous like the reference wrapper; C is fully overwritten.
 */
#pragma once

#include <cuda_runtime.h>

#include <loops/container/formats.hxx>
#include <loops/container/matrix.cuh>
#include <loops/error.hxx>
#include <loopsb.h>

namespace loops {
namespace algorithms {
namespace spmm {

inline void thread_mapped(csr_t<int, int, float>& csr, matrix_t<float>& B, matrix_t<float>& C,
                          cudaStream_t stream = 0) {
  error::throw_if_exception(B.rows != csr.cols || C.rows != csr.rows || B.cols != C.cols,
                            "spmm::thread_mapped: shapes must be A[m x k] B[k x n] C[m x n]");
  const loopsb_layout_t lay = csr.layout().descriptor();
  error::throw_if_status(
      loopsb_spmm_csr_f32(&lay, thrust::raw_pointer_cast(csr.values.data()),
                          thrust::raw_pointer_cast(csr.indices.data()), B.m_data_ptr, C.m_data_ptr,
                          static_cast<int32_t>(csr.rows), static_cast<int32_t>(csr.cols),
                          static_cast<int32_t>(B.cols), stream),
      "loopsb_spmm_csr_f32");
  cudaStreamSynchronize(stream);
}

}  // namespace spmm
}  // namespace algorithms
}  // namespace loops
