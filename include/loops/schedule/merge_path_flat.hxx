/**
 * @file merge_path_flat.hxx
 * @brief Merge-path (flat) schedule: the merge of "tile ends" with "atom ids"
 * is cut into equal pieces, first per block, then per thread (reference
 * include/loops/schedule/merge_path_flat.hxx:33-393).
 *
 * Arithmetic kept identical to the reference so the (tile, atom) stream every
 * (block, thread, item) sees is bit-exact:
 *   W = tiles + atoms, I = TPB * IPT, M = ceil(W / I)
 *   block b = blockIdx.x * gridDim.y + blockIdx.y owns diagonals [b I, b I + I)
 *   s = S(b I), e = S(b I + I)        (search::diagonal_split on the layout)
 *   E[j] = tile_end[min(s.x + j, T - 1)],  j < (e.x - s.x) + IPT   (shared)
 *   thread k starts at S_local(k IPT) over E vs counting(s.y)
 *
 * What differs is how the data gets there: when the layout's tile_end_iter()
 * is a real array (csr / csc / bcsr), the E window is pulled into shared memory
 * with one 1-D bulk async copy (TMA, `cp.async.bulk` + mbarrier) issued by a
 * single thread instead of a strided loop of per-thread loads; arithmetic
 * layouts (coo / ell / dia / partitioner) compute E in registers. The per-block
 * coordinates come from a `preprocess_t` that OWNS its device buffer through a
 * shared handle, so by-value copies handed to kernels keep a valid pointer.
 */
#pragma once

#include <cstdint>
#include <memory>
#include <type_traits>

#include <cuda_runtime.h>

#include <loops/stride_ranges.hxx>
#include <loops/util/math.hxx>
#include <loops/util/search.hxx>
#include <loops/util/tma.hxx>
#include <loops/container/coordinate.hxx>
#include <loops/container/layout.hxx>

namespace loops {
namespace schedule {

using coord_t = coordinate_t<unsigned int>;

/// Coordinate value init() hands to blocks that have no merge tile.
static constexpr unsigned int invalid_coordinate = 0xffffffffu;

namespace merge_path {

/// One thread per merge tile boundary: coords[b] = S(b * TPB * IPT), b = 0..M.
template <std::size_t THREADS_PER_BLOCK,
          std::size_t ITEMS_PER_THREAD,
          typename layout_t,
          typename tile_size_t,
          typename atom_size_t>
__global__ void generate_search_coordinates(layout_t layout,
                                            tile_size_t num_tiles,
                                            atom_size_t num_atoms,
                                            std::size_t num_merge_tiles,
                                            coord_t* d_tile_coordinates) {
  const std::size_t b = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b > num_merge_tiles)
    return;
  constexpr long long items_per_tile =
      (long long)(THREADS_PER_BLOCK * ITEMS_PER_THREAD);
  d_tile_coordinates[b] = search::diagonal_split(
      (long long)b * items_per_tile, layout.tile_end_iter(),
      counting_iterator<long long>(0), (long long)num_tiles,
      (long long)num_atoms);
}

/// Host-side helper that precomputes the per-block coordinates once.
template <std::size_t THREADS_PER_BLOCK,
          std::size_t ITEMS_PER_THREAD,
          typename tiles_type,
          typename atoms_type,
          typename tile_size_type,
          typename atom_size_type,
          typename layout_type = layout::csr<tiles_type, atoms_type>>
class preprocess_t {
 public:
  using tiles_t = tiles_type;
  using atoms_t = atoms_type;
  using tiles_iterator_t = tiles_t*;
  using atoms_iterator_t = atoms_t*;
  using tile_size_t = tile_size_type;
  using atom_size_t = atom_size_type;
  using layout_t = layout_type;

  preprocess_t(tiles_iterator_t _tiles,
               tile_size_t _num_tiles,
               atom_size_t _num_atoms,
               cudaStream_t stream = 0)
      : preprocess_t(
            layout_t(_tiles,
                     static_cast<typename layout_t::tile_id_t>(_num_tiles),
                     static_cast<typename layout_t::atom_id_t>(_num_atoms)),
            stream) {}

  explicit preprocess_t(layout_t _layout, cudaStream_t stream = 0)
      : total_work(std::size_t(_layout.num_tiles()) +
                   std::size_t(_layout.num_atoms())),
        num_merge_tiles(math::ceil_div(
            total_work, std::size_t(THREADS_PER_BLOCK * ITEMS_PER_THREAD))),
        d_tile_coordinates(nullptr) {
    if (num_merge_tiles == 0)
      return;
    coord_t* raw = nullptr;
    if (cudaMalloc(&raw, (num_merge_tiles + 1) * sizeof(coord_t)) !=
        cudaSuccess) {
      (void)cudaGetLastError();
      return;  // kernels fall back to the in-block search
    }
    owner_ = std::shared_ptr<coord_t>(raw, [](coord_t* p) { cudaFree(p); });
    d_tile_coordinates = raw;
    constexpr unsigned block = 128;
    const unsigned grid =
        static_cast<unsigned>((num_merge_tiles + 1 + block - 1) / block);
    generate_search_coordinates<THREADS_PER_BLOCK, ITEMS_PER_THREAD, layout_t,
                                tile_size_t, atom_size_t>
        <<<grid, block, 0, stream>>>(
            _layout, static_cast<tile_size_t>(_layout.num_tiles()),
            static_cast<atom_size_t>(_layout.num_atoms()), num_merge_tiles,
            d_tile_coordinates);
  }

  __host__ __device__ coord_t* data() const { return d_tile_coordinates; }
  __host__ __device__ std::size_t merge_tiles() const { return num_merge_tiles; }

 private:
  std::size_t total_work;
  std::size_t num_merge_tiles;
  coord_t* d_tile_coordinates;
  std::shared_ptr<coord_t> owner_;  ///< host-side lifetime only
};

}  // namespace merge_path

template <std::size_t THREADS_PER_BLOCK,
          std::size_t ITEMS_PER_THREAD,
          typename tiles_type,
          typename atoms_type,
          typename tile_size_type,
          typename atom_size_type,
          typename layout_type>
class setup<algorithms_t::merge_path_flat,
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
  using meta_t = merge_path::preprocess_t<THREADS_PER_BLOCK,
                                          ITEMS_PER_THREAD,
                                          tiles_type,
                                          atoms_type,
                                          tile_size_type,
                                          atom_size_type,
                                          layout_type>;

  enum : unsigned int {
    threads_per_block = THREADS_PER_BLOCK,
    items_per_thread = ITEMS_PER_THREAD,
    items_per_tile = threads_per_block * items_per_thread,
    window_capacity = ((items_per_thread + items_per_tile + 1 + 3 + 3) / 4) * 4,
  };

  /// Window of tile ends staged in shared memory. Indexing starts at the
  /// block's first tile; `skew` absorbs the 16-byte alignment the bulk copy
  /// needs on the global side.
  struct __align__(16) window_t {
    tiles_t raw[window_capacity];
    unsigned int skew;
    __device__ __forceinline__ tiles_t& operator[](unsigned int i) {
      return raw[skew + i];
    }
    __device__ __forceinline__ const tiles_t& operator[](unsigned int i) const {
      return raw[skew + i];
    }
    __device__ __forceinline__ const tiles_t* data() const { return raw + skew; }
  };

  /// Shared memory one block needs.
  struct __align__(16) storage_t {
    window_t tile_end_offset;
    coord_t tile_coords[2];
    unsigned long long copy_done;  ///< mbarrier for the bulk copy
  };

  storage_t& buffer;
  meta_t& meta;

  counting_iterator<atoms_t> atoms_counting_it;
  counting_iterator<tiles_t> tiles_counting_it;

  __device__ __forceinline__ setup(meta_t& _meta,
                                   storage_t& _buffer,
                                   tiles_iterator_t _tiles,
                                   tile_size_t _num_tiles,
                                   atom_size_t _num_atoms)
      : buffer(_buffer),
        meta(_meta),
        view_(_tiles,
              static_cast<typename layout_t::tile_id_t>(_num_tiles),
              static_cast<typename layout_t::atom_id_t>(_num_atoms)) {
    derive();
  }

  __device__ __forceinline__ setup(meta_t& _meta,
                                   storage_t& _buffer,
                                   layout_t _layout)
      : buffer(_buffer), meta(_meta), view_(_layout) {
    derive();
  }

  /// Block-level cut, window staging, thread-level cut. Returns the calling
  /// thread's starting coordinate RELATIVE to the block's (or the invalid
  /// sentinel for blocks past the last merge tile).
  __device__ __forceinline__ coord_t init() {
    const std::size_t b = std::size_t(blockIdx.x) * gridDim.y + blockIdx.y;
    if (b >= merge_tiles_)
      return coord_t{invalid_coordinate, invalid_coordinate};

    const long long T = static_cast<long long>(view_.num_tiles());
    const long long A = static_cast<long long>(view_.num_atoms());

    if (threadIdx.x < 2) {
      const coord_t* pre = meta.data();
      buffer.tile_coords[threadIdx.x] =
          pre ? pre[b + threadIdx.x]
              : search::diagonal_split(
                    static_cast<long long>(b + threadIdx.x) * items_per_tile,
                    view_.tile_end_iter(), counting_iterator<long long>(0), T,
                    A);
      if (threadIdx.x == 0 && use_bulk_copy())
        tma::barrier_init(
            reinterpret_cast<uint64_t*>(&buffer.copy_done), 1);
    }
    __syncthreads();

    const coord_t s = buffer.tile_coords[0];
    const coord_t e = buffer.tile_coords[1];
    tile_num_tiles = e.x - s.x;
    tile_num_atoms = e.y - s.y;

    stage_window(s, T);

    tiles_counting_it = counting_iterator<tiles_t>(static_cast<tiles_t>(s.x));
    atoms_counting_it = counting_iterator<atoms_t>(static_cast<atoms_t>(s.y));

    return search::diagonal_split(
        static_cast<long long>(threadIdx.x) * items_per_thread,
        buffer.tile_end_offset.data(),
        counting_iterator<long long>(static_cast<long long>(s.y)),
        static_cast<long long>(tile_num_tiles),
        static_cast<long long>(tile_num_atoms));
  }

  __device__ __forceinline__ bool is_valid_accessor(coord_t& coord) const {
    return coord.x != invalid_coordinate && coord.y != invalid_coordinate;
  }

  /// 0 .. items_per_thread.
  __device__ __forceinline__ step_range_t<int> virtual_idx() const {
    return custom_stride_range(int(0), int(items_per_thread), int(1));
  }

  /// Atom under the cursor (clamped so the speculative load on a row-advance
  /// step stays in bounds).
  __device__ __forceinline__ atoms_t atom_idx(int, coord_t& coord) {
    const atoms_t a = atoms_counting_it[coord.y];
    const atoms_t last = static_cast<atoms_t>(view_.num_atoms()) - 1;
    return a < last ? a : last;
  }

  /// Tile under the cursor.
  __device__ __forceinline__ tiles_t tile_idx(coord_t& coord) const {
    return tiles_counting_it[coord.x];
  }

  __device__ __forceinline__ tile_size_t num_tiles() const {
    return tile_num_tiles;
  }
  __device__ __forceinline__ atom_size_t num_atoms() const {
    return tile_num_atoms;
  }

  __host__ __device__ const layout_t& layout() const { return view_; }

 private:
  using end_iter_t = typename layout_t::tile_end_iterator_t;
  static constexpr bool ends_are_array =
      std::is_pointer<end_iter_t>::value &&
      sizeof(typename std::remove_pointer<end_iter_t>::type) == 4 &&
      sizeof(tiles_t) == 4;

  __device__ __forceinline__ static constexpr bool use_bulk_copy() {
#if LOOPS_HAS_BULK_COPY
    return ends_are_array;
#else
    return false;
#endif
  }

  __device__ __forceinline__ void derive() {
    total_work_ = std::size_t(view_.num_tiles()) + std::size_t(view_.num_atoms());
    merge_tiles_ = math::ceil_div(total_work_, std::size_t(items_per_tile));
  }

  /// Fill buffer.tile_end_offset[j] = tile_end[min(s.x + j, T - 1)] for
  /// j < tile_num_tiles + items_per_thread, then barrier.
  __device__ __forceinline__ void stage_window(const coord_t& s, long long T) {
    const int want = static_cast<int>(tile_num_tiles) + int(items_per_thread);
    const auto ends = view_.tile_end_iter();
    if constexpr (ends_are_array) {
#if LOOPS_HAS_BULK_COPY
      // Entries that exist in the array; the rest replicate the last end.
      long long have = T - static_cast<long long>(s.x);
      if (have > want) have = want;
      if (have < 0) have = 0;
      const tiles_t* first = reinterpret_cast<const tiles_t*>(ends) + s.x;
      const unsigned skew =
          static_cast<unsigned>((reinterpret_cast<uintptr_t>(first) & 15u) >> 2);
      const int covered = static_cast<int>(skew + have);  // raw[] entries
      const int bulk = covered & ~3;                      // whole 16-B chunks
      uint64_t* bar = reinterpret_cast<uint64_t*>(&buffer.copy_done);
      if (threadIdx.x == 0) {
        buffer.tile_end_offset.skew = skew;
        if (bulk > 0) {
          tma::barrier_arrive_expect_tx(bar, uint32_t(bulk) * 4u);
          tma::bulk_g2s(buffer.tile_end_offset.raw, first - skew,
                        uint32_t(bulk) * 4u, bar);
        }
      }
      // Ragged tail of the array part, and the replicated clamp entries.
      const tiles_t last_end = (T > 0) ? ends[T - 1] : tiles_t(0);
      for (int r = bulk + int(threadIdx.x); r < int(skew) + want;
           r += int(threads_per_block)) {
        buffer.tile_end_offset.raw[r] =
            (r < covered) ? (first - skew)[r] : last_end;
      }
      if (bulk > 0)
        tma::barrier_wait(bar, 0);
      __syncthreads();
      return;
#endif
    }
    if (threadIdx.x == 0)
      buffer.tile_end_offset.skew = 0;
    for (int j = int(threadIdx.x); j < want; j += int(threads_per_block)) {
      long long at = static_cast<long long>(s.x) + j;
      if (at > T - 1) at = T - 1;
      buffer.tile_end_offset.raw[j] = (at >= 0) ? ends[at] : tiles_t(0);
    }
    __syncthreads();
  }

  layout_t view_;
  std::size_t total_work_;
  std::size_t merge_tiles_;
  tile_size_t tile_num_tiles;
  atom_size_t tile_num_atoms;
};

}  // namespace schedule
}  // namespace loops
