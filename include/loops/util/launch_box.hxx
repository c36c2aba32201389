/**
 * @file launch_box.hxx
 * @brief Compile-time launch-parameter selection under the reference's names (reference
 * include/loops/util/launch_box.hxx:48-239): `sm_flag_t`, `launch_params_t<flags, block,
 * items, smem>`, `launch_box_t<params...>` (first entry whose flags match the target
 * architecture, `fallback` matches anything) and `occupancy_grid`. loops-b200 is built
 * for sm_100a only, so the target flag is `sm_100` unless LOOPS_TARGET_ARCH says otherwise;
 * boxes written for the reference (sm_80 | sm_86, ..., fallback) select the same entry
 * they would select in a reference build pinned to that architecture.
 */
#pragma once

#include <cstddef>
#include <type_traits>

#include <cuda_runtime.h>

#include <loops/util/device.hxx>

namespace loops {
namespace launch_box {

enum sm_flag_t : unsigned int {
  fallback = 1u << 0,
  sm_70 = 1u << 1,
  sm_72 = 1u << 2,
  sm_75 = 1u << 3,
  sm_80 = 1u << 4,
  sm_86 = 1u << 5,
  sm_89 = 1u << 6,
  sm_90 = 1u << 7,
  sm_100 = 1u << 8,
  // AMD flags keep their reference bit positions so boxes naming them still compile;
  // they can never match here (no HIP backend).
  gfx906 = 1u << 16, gfx908 = 1u << 17, gfx90a = 1u << 18, gfx942 = 1u << 19, gfx950 = 1u << 20,
  gfx1030 = 1u << 21, gfx1100 = 1u << 22, gfx1200 = 1u << 23, gfx1201 = 1u << 24,
};

constexpr sm_flag_t operator|(sm_flag_t a, sm_flag_t b) {
  return static_cast<sm_flag_t>(static_cast<unsigned int>(a) | static_cast<unsigned int>(b));
}
constexpr sm_flag_t operator&(sm_flag_t a, sm_flag_t b) {
  return static_cast<sm_flag_t>(static_cast<unsigned int>(a) & static_cast<unsigned int>(b));
}

/// Flag of a compute capability given as major*10+minor; unknown values carry no bit.
constexpr sm_flag_t flag_of(int cc) {
  return cc == 70 ? sm_70 : cc == 72 ? sm_72 : cc == 75 ? sm_75 : cc == 80 ? sm_80 : cc == 86 ? sm_86
       : cc == 89 ? sm_89 : cc == 90 ? sm_90 : cc == 100 ? sm_100 : static_cast<sm_flag_t>(0u);
}

#ifndef LOOPS_TARGET_ARCH
#define LOOPS_TARGET_ARCH 100   // the only architecture this library is compiled for
#endif
constexpr sm_flag_t target_flag = flag_of(LOOPS_TARGET_ARCH);

template <sm_flag_t sm_flags_, std::size_t block_size_, std::size_t items_per_thread_ = 1,
          std::size_t shared_memory_bytes_ = 0>
struct launch_params_t {
  static constexpr sm_flag_t sm_flags = sm_flags_;
  static constexpr std::size_t block_size = block_size_;
  static constexpr std::size_t items_per_thread = items_per_thread_;
  static constexpr std::size_t shared_memory_bytes = shared_memory_bytes_;
};

namespace detail {
template <sm_flag_t target>
struct no_match_t {
  static_assert(target != target, "launch_box_t: no entry matches the target architecture (add a fallback entry)");
};

template <sm_flag_t target, typename... entries_t>
struct first_match { using type = no_match_t<target>; };

template <sm_flag_t target, typename head_t, typename... tail_t>
struct first_match<target, head_t, tail_t...> {
  static constexpr bool hit = (head_t::sm_flags & target) != 0u || (head_t::sm_flags & fallback) != 0u;
  using type = std::conditional_t<hit, head_t, typename first_match<target, tail_t...>::type>;
};
}  // namespace detail

/// `using box_t = launch_box_t<launch_params_t<sm_90 | sm_100, 128, 8>, launch_params_t<fallback, 128, 7>>;`
template <typename... entries_t>
struct launch_box_t : detail::first_match<target_flag, entries_t...>::type {};

/// Resident blocks per SM (occupancy API) x SM count: the grid a persistent / grid-stride kernel covers the
/// device with in one wave.
template <typename kernel_t>
inline std::size_t occupancy_grid(const kernel_t& kernel, int block_size, std::size_t dynamic_shared_memory_bytes = 0) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block_size, dynamic_shared_memory_bytes);
  if (per_sm < 1) per_sm = 1;
  return static_cast<std::size_t>(per_sm) * static_cast<std::size_t>(device::multi_processor_count());
}

}  // namespace launch_box
}  // namespace loops
