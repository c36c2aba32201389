/**
 * @file launch_box.hxx
 * @brief Reference-visible schedule geometry for SpMV (reference
 * include/loops/algorithms/spmv/launch_box.hxx:63-90). loops-b200 targets
 * sm_100a only, so there is one entry: 128 threads x 8 items for 4-byte values
 * (x 4 for 8-byte ones) -- the reference's sm_90|sm_100 row. These numbers fix
 * the merge-path coordinates; the sm_100a kernels behind the C ABI pick their
 * own CTA geometry on top (see loops_b200/csrc/spmv_merge.cuh).
 */
#pragma once
#include <cstddef>

namespace loops {
namespace algorithms {
namespace spmv {

template <typename type_t>
struct launch_t {
  static constexpr std::size_t block_size = 128;
  static constexpr std::size_t items_per_thread = sizeof(type_t) > 4 ? 4 : 8;
  static constexpr std::size_t shared_memory_bytes = 0;
};

}  // namespace spmv
}  // namespace algorithms
}  // namespace loops
