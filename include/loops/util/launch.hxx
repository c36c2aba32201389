/**
 * @file launch.hxx
 * @brief Kernel launch helpers under the reference's names (reference
 * include/loops/util/launch.hxx:32-75): `launch::non_cooperative(stream, kernel, grid,
 * block, args...)` and `launch::cooperative(stream, kernel, blocks, threads, args...)`.
 * User kernels written against `schedule::setup<>` (e.g. the reference's
 * examples/spmv/custom_layout.cu:230) launch through these; the library's own SpMV
 * entry points go through the C ABI instead.
 */
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <utility>

#include <loops/backend/xpu.hxx>

namespace loops {
namespace launch {

/// `kernel<<<grid, block, 0, stream>>>(args...)`.
template <typename func_t, typename... args_t>
void non_cooperative(cudaStream_t stream, const func_t& kernel, dim3 number_of_blocks, dim3 threads_per_block,
                     args_t&&... args) {
  kernel<<<number_of_blocks, threads_per_block, 0, stream>>>(std::forward<args_t>(args)...);
}

/// Cooperative launch (grid-wide sync available to the kernel): the argument pack is
/// handed to cudaLaunchCooperativeKernel as an array of addresses.
template <typename func_t, typename... args_t>
void cooperative(cudaStream_t stream, const func_t& kernel, std::size_t number_of_blocks,
                 std::size_t threads_per_block, args_t&&... args) {
  void* slots[sizeof...(args_t) == 0 ? 1 : sizeof...(args_t)] = {
      const_cast<void*>(static_cast<const void*>(&args))...};
  cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&kernel), dim3(static_cast<unsigned>(number_of_blocks)),
                              dim3(static_cast<unsigned>(threads_per_block)), slots, 0, stream);
}

}  // namespace launch
}  // namespace loops
