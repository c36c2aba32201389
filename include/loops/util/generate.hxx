/**
 * @file generate.hxx
 * @brief Input generators under the reference's names (reference include/loops/util/generate.hxx:
 * `generate::random::hash` :33-41, `uniform_distribution` :54-79, `random::csr` :94-113).
 * The x recipe of every example -- `uniform_distribution(x.begin(), x.end(), 1, 10, 42u)` -- must
 * reproduce the reference's vector bit for bit (it is what --validate compares against), so the
 * per-index engine and distributions are thrust's own: element i is the first draw of a
 * `thrust::default_random_engine` seeded with `hash(i) * seed`, through
 * `uniform_int_distribution` for integral bounds (the examples pass ints: x holds 1..10) and
 * `uniform_real_distribution` for floating-point ones.
 */
#pragma once

#include <chrono>
#include <cstddef>
#include <type_traits>

#include <thrust/distance.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/random.h>
#include <thrust/transform.h>

#include <loops/container/formats.hxx>
#include <loops/memory.hxx>

namespace loops {
namespace generate {
namespace random {

/// Integer mixing hash (Bob Jenkins' 6-shift variant), applied to the element index.
__forceinline__ __host__ __device__ unsigned int hash(unsigned int a) {
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

namespace detail {
template <typename type_t>
struct draw_t {
  type_t lo, hi;
  unsigned int useed;
  __host__ __device__ type_t operator()(std::size_t i) const {
    thrust::default_random_engine rng(hash(static_cast<unsigned int>(i)) * useed);
    if constexpr (std::is_floating_point_v<type_t>) {
      thrust::uniform_real_distribution<type_t> dist(lo, hi);
      return dist(rng);
    } else {
      thrust::uniform_int_distribution<type_t> dist(lo, hi);
      return dist(rng);
    }
  }
};
}  // namespace detail

/// Fill [begin_it, end_it) with independent draws from [begin, end]; the element type of the
/// DRAW is the type of the bounds (ints -> integers stored into whatever the range holds).
template <typename iterator_t, typename type_t>
void uniform_distribution(iterator_t begin_it, iterator_t end_it, type_t begin, type_t end,
                          unsigned int useed = static_cast<unsigned int>(
                              std::chrono::system_clock::now().time_since_epoch().count())) {
  const std::size_t n = static_cast<std::size_t>(thrust::distance(begin_it, end_it));
  thrust::transform(thrust::make_counting_iterator<std::size_t>(0), thrust::make_counting_iterator<std::size_t>(n),
                    begin_it, detail::draw_t<type_t>{begin, end, useed});
}

/// Uniformly random CSR with about sparsity * rows * cols entries (duplicates dropped).
template <typename index_t, typename offset_t, typename value_t>
void csr(std::size_t rows, std::size_t cols, float sparsity, csr_t<index_t, offset_t, value_t>& matrix) {
  const std::size_t nnzs = static_cast<std::size_t>(sparsity * float(rows * cols));
  coo_t<index_t, value_t, memory::memory_space_t::host> coo(rows, cols, nnzs);
  uniform_distribution(coo.row_indices.begin(), coo.row_indices.end(), index_t(0), index_t(rows ? rows - 1 : 0));
  uniform_distribution(coo.col_indices.begin(), coo.col_indices.end(), index_t(0), index_t(cols ? cols - 1 : 0));
  uniform_distribution(coo.values.begin(), coo.values.end(), value_t(0.0), value_t(1.0));
  coo.remove_duplicates();
  matrix = csr_t<index_t, offset_t, value_t>(coo);
}

}  // namespace random
}  // namespace generate
}  // namespace loops
