/**
 * @file thread_mapped.hxx
 * @brief One thread per tile (reference
 * include/loops/schedule/thread_mapped.hxx:40-129).
 *
 * Thread g = blockIdx.x*blockDim.x + threadIdx.x is handed tiles g, g + G,
 * g + 2G, ... (G = threads in the grid) and, per tile, the atoms
 * tile_begin(t) .. tile_end(t)-1 in ascending order. Host- and
 * device-constructible so the host wrapper can build it once and pass it by
 * value.
 */
#pragma once

#include <loops/stride_ranges.hxx>
#include <loops/container/layout.hxx>

namespace loops {
namespace schedule {

template <typename tiles_type,
          typename atoms_type,
          typename tile_size_type,
          typename atom_size_type,
          typename layout_type>
class setup<algorithms_t::thread_mapped,
            1,
            1,
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

  __host__ __device__ setup() : view_() {}

  /// CSR shortcut: offsets pointer + counts (valid when layout_t is
  /// offsets-shaped).
  __host__ __device__ setup(tiles_t* offsets,
                            tile_size_t num_tiles,
                            atom_size_t num_atoms)
      : view_(offsets,
              static_cast<typename layout_t::tile_id_t>(num_tiles),
              static_cast<typename layout_t::atom_id_t>(num_atoms)) {}

  __host__ __device__ explicit setup(layout_t view) : view_(view) {}

  /// Tiles of the calling thread (grid stride).
  __device__ step_range_t<tile_size_t> tiles() const {
    return grid_stride_range(tile_size_t(0),
                             static_cast<tile_size_t>(view_.num_tiles()));
  }

  /// All atoms of `tile`, ascending.
  __device__ auto atoms(const tile_size_t& tile) {
    return loops::range(view_.tile_begin(tile), view_.tile_end(tile));
  }

  /// Atoms of `tile` starting from a caller-computed position (resume).
  template <typename first_atom_fn_t>
  __device__ auto atoms(const tile_size_t& tile, first_atom_fn_t first_atom) {
    return loops::range(first_atom(tile), view_.tile_end(tile));
  }

  __host__ __device__ const layout_t& layout() const { return view_; }

 private:
  layout_t view_;
};

}  // namespace schedule
}  // namespace loops
