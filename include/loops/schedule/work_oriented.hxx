/**
 * @file work_oriented.hxx
 * @brief Even share of (tiles + atoms) per THREAD (reference
 * include/loops/schedule/work_oriented.hxx:45-190).
 *
 * With W = tiles + atoms and N = gridDim.x * threads_per_block threads, thread
 * g owns the merge-path items [min(w g, W), min(w g + w, W)), w = ceil(W / N).
 * `init()` turns both ends into (tile, atom) coordinates by diagonal search;
 * the body then walks whole tiles (`tiles(m)` / `atoms(t, m)`) and finally the
 * partial tile the share ends in (`remainder_tiles(m)` / `remainder_atoms(m)`).
 */
#pragma once

#include <loops/stride_ranges.hxx>
#include <loops/util/math.hxx>
#include <loops/util/search.hxx>
#include <loops/container/layout.hxx>

namespace loops {
namespace schedule {

/// {first, second} aggregate with the member names the reference's
/// thrust::pair-based map exposes (m.first.first, m.second.second, ...).
template <typename a_t, typename b_t>
struct pair_t {
  a_t first;
  b_t second;
};

template <std::size_t THREADS_PER_BLOCK,
          std::size_t ITEMS_PER_THREAD,
          typename tiles_type,
          typename atoms_type,
          typename tile_size_type,
          typename atom_size_type,
          typename layout_type>
class setup<algorithms_t::work_oriented,
            THREADS_PER_BLOCK,
            ITEMS_PER_THREAD,
            tiles_type,
            atoms_type,
            tile_size_type,
            atom_size_type,
            layout_type> {
 public:
  using tiles_t = tiles_type;
  using atoms_t = atoms_type;
  using tiles_iterator_t = tiles_t*;
  using atoms_iterator_t = atoms_t*;
  using tile_size_t = tile_size_type;
  using atom_size_t = atom_size_type;
  using layout_t = layout_type;
  using point_t = pair_t<atom_size_t, atom_size_t>;  ///< (tile, atom)
  using map_t = pair_t<point_t, point_t>;            ///< (start, end)

  enum : unsigned int {
    threads_per_block = THREADS_PER_BLOCK,
    items_per_thread = ITEMS_PER_THREAD,
    items_per_tile = threads_per_block * items_per_thread,
  };

  __device__ __forceinline__ setup(tiles_iterator_t _tiles,
                                   tile_size_t _num_tiles,
                                   atom_size_t _num_atoms)
      : view_(_tiles,
              static_cast<typename layout_t::tile_id_t>(_num_tiles),
              static_cast<typename layout_t::atom_id_t>(_num_atoms)) {
    derive();
  }

  __device__ __forceinline__ explicit setup(layout_t _layout) : view_(_layout) {
    derive();
  }

  /// Both ends of the calling thread's share as merge-grid coordinates.
  __device__ __forceinline__ map_t init() {
    const std::size_t g = threadIdx.x + std::size_t(blockIdx.x) * blockDim.x;
    std::size_t from = share_ * g;
    if (from > work_)
      from = work_;
    std::size_t to = from + share_;
    if (to > work_)
      to = work_;
    map_t m;
    m.first = locate(from);
    m.second = locate(to);
    return m;
  }

  /// Tiles that END inside the share.
  template <typename map_type>
  __device__ __forceinline__ step_range_t<tiles_t> tiles(map_type& m) const {
    return custom_stride_range(tiles_t(m.first.first), tiles_t(m.second.first),
                               tiles_t(1));
  }

  /// Atoms of `t` from the share's cursor to the tile's end; moves the cursor.
  template <typename map_type>
  __device__ __forceinline__ step_range_t<atoms_t> atoms(tiles_t t,
                                                         map_type& m) {
    const atoms_t stop = view_.tile_end(t);
    const atoms_t start = static_cast<atoms_t>(m.first.second);
    m.first.second += (stop - start);
    return custom_stride_range(start, stop, atoms_t(1));
  }

  /// The one tile the share ends in the middle of.
  template <typename map_type>
  __device__ __forceinline__ step_range_t<tiles_t> remainder_tiles(
      map_type& m) const {
    return custom_stride_range(tiles_t(m.second.first),
                               tiles_t(m.second.first + 1), tiles_t(1));
  }

  /// Its leading atoms that still belong to this share.
  template <typename map_type>
  __device__ __forceinline__ step_range_t<atoms_t> remainder_atoms(
      map_type& m) const {
    return custom_stride_range(atoms_t(m.first.second),
                               atoms_t(m.second.second), atoms_t(1));
  }

  __host__ __device__ const layout_t& layout() const { return view_; }

 private:
  __device__ __forceinline__ void derive() {
    work_ = std::size_t(view_.num_tiles()) + std::size_t(view_.num_atoms());
    threads_ = std::size_t(gridDim.x) * threads_per_block;
    share_ = math::ceil_div(work_, threads_);
  }

  __device__ __forceinline__ point_t locate(std::size_t diagonal) const {
    const auto c = search::diagonal_split(
        static_cast<long long>(diagonal), view_.tile_end_iter(),
        counting_iterator<long long>(0),
        static_cast<long long>(view_.num_tiles()),
        static_cast<long long>(view_.num_atoms()));
    point_t p;
    p.first = c.x;
    p.second = c.y;
    return p;
  }

  layout_t view_;
  std::size_t work_;
  std::size_t threads_;
  std::size_t share_;
};

}  // namespace schedule
}  // namespace loops
