/**
 * @file group_mapped.hxx
 * @brief A group of threads (warp, block, or any power-of-two slice of the
 * block) shares the atoms of the tiles its members own (reference
 * include/loops/schedule/group_mapped.hxx:39-210).
 *
 * Every thread of the grid owns tile `global thread rank` (or none past the
 * end). Within a group the owned tiles' sizes are prefix-summed; the resulting
 * flat "virtual atom" range [0, aggregate) is dealt round-robin to the group's
 * threads (rank r takes r, r + size, ...). A virtual atom is mapped back to
 * its tile with an upper-bound search over the prefix sums and to its real
 * atom id by offsetting into that tile.
 *
 * Same public surface as the reference (storage_t, partition(),
 * atom_accessor(), tile_accessor(), is_valid_accessor(), tile_id(), atom_id(),
 * get_length(), warp_mapped / block_mapped aliases). The group object returned
 * by partition() is a plain value type and the prefix sum is a warp-shuffle
 * scan with one shared-memory hop across warps -- no cooperative-groups
 * dependency.
 */
#pragma once

#include <loops/stride_ranges.hxx>
#include <loops/container/layout.hxx>

namespace loops {
namespace schedule {

/// Handle for "my slice of the block": `group_threads` consecutive threads.
template <unsigned group_threads, unsigned block_threads>
struct thread_group_t {
  static_assert(group_threads > 0 && (group_threads & (group_threads - 1)) == 0,
                "group size must be a power of two");
  static_assert(block_threads % group_threads == 0,
                "groups must tile the block");

  __device__ __forceinline__ unsigned thread_rank() const {
    return threadIdx.x % group_threads;
  }
  __device__ __forceinline__ unsigned size() const { return group_threads; }
  __device__ __forceinline__ unsigned meta_group_rank() const {
    return threadIdx.x / group_threads;
  }
  __device__ __forceinline__ unsigned meta_group_size() const {
    return block_threads / group_threads;
  }
  /// Barrier + memory ordering among the group's threads.
  __device__ __forceinline__ void sync() const {
    if (group_threads == block_threads) {
      __syncthreads();
    } else if (group_threads <= 32) {
      const unsigned lane = threadIdx.x & 31u;
      const unsigned base = lane & ~(group_threads - 1u);
      const unsigned mask =
          (0xffffffffu >> (32u - (group_threads < 32 ? group_threads : 32u)))
          << base;
      __syncwarp(mask);
    } else {
      // Multi-warp slice of the block: one named barrier per group
      // (ids 1..15; id 0 is __syncthreads()).
      asm volatile("bar.sync %0, %1;" ::"r"(1u + meta_group_rank()),
                   "r"(group_threads)
                   : "memory");
    }
  }
};

template <std::size_t THREADS_PER_BLOCK,
          std::size_t THREADS_PER_TILE,
          typename tiles_type,
          typename atoms_type,
          typename tile_size_type,
          typename atom_size_type,
          typename layout_type>
class setup<algorithms_t::group_mapped,
            THREADS_PER_BLOCK,
            THREADS_PER_TILE,
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
  using group_t = thread_group_t<unsigned(THREADS_PER_TILE),
                                 unsigned(THREADS_PER_BLOCK)>;

  enum : unsigned int {
    threads_per_block = THREADS_PER_BLOCK,
    threads_per_tile = THREADS_PER_TILE,
    tiles_per_block = THREADS_PER_BLOCK / THREADS_PER_TILE,
    warps_per_block = (THREADS_PER_BLOCK + 31) / 32,
  };

  /// Shared-memory scratch of one block.
  struct __align__(16) storage_t {
    atoms_t tile_aggregates[tiles_per_block];   ///< atoms owned per group
    atoms_t atoms_offsets[threads_per_block];   ///< exclusive prefix per group
    tiles_t tiles_indices[threads_per_block];   ///< owned tile id, or -1
    atoms_t warp_totals[warps_per_block];       ///< cross-warp scan hop
  };

  storage_t& buffer;

  __device__ __forceinline__ setup(storage_t& _buffer,
                                   tiles_iterator_t _tiles,
                                   tile_size_t _num_tiles,
                                   atom_size_t _num_atoms)
      : buffer(_buffer),
        view_(_tiles,
              static_cast<typename layout_t::tile_id_t>(_num_tiles),
              static_cast<typename layout_t::atom_id_t>(_num_atoms)) {}

  __device__ __forceinline__ setup(storage_t& _buffer, layout_t _layout)
      : buffer(_buffer), view_(_layout) {}

  /// Record which tile each thread owns; hand back the group handle.
  __device__ __forceinline__ group_t partition() {
    const long long me = global_rank();
    buffer.tiles_indices[threadIdx.x] =
        (me < static_cast<long long>(view_.num_tiles()))
            ? static_cast<tiles_t>(me)
            : static_cast<tiles_t>(-1);
    return group_t();
  }

  /// Prefix-sum the group's tile sizes; return this thread's share of the
  /// flattened atom range.
  template <typename partition_t>
  __device__ step_range_t<atoms_t> atom_accessor(partition_t& p) {
    const long long me = global_rank();
    atoms_t mine = 0;
    if (me < static_cast<long long>(view_.num_tiles()))
      mine = view_.tile_size(static_cast<tiles_t>(me));

    const atoms_t before = group_exclusive_sum(mine, p);
    buffer.atoms_offsets[threadIdx.x] = before;
    if (p.thread_rank() == p.size() - 1)
      buffer.tile_aggregates[p.meta_group_rank()] = before + mine;
    p.sync();

    return custom_stride_range(atoms_t(p.thread_rank()),
                               buffer.tile_aggregates[p.meta_group_rank()],
                               atoms_t(p.size()));
  }

  /// How many of the group's ranks own a real tile.
  template <typename partition_t>
  __device__ __forceinline__ int get_length(partition_t& p) {
    const long long group_first = global_rank() - p.thread_rank();
    long long past = group_first + p.size();
    const long long tiles = static_cast<long long>(view_.num_tiles());
    if (tiles < past)
      past = tiles;
    return static_cast<int>(past - group_first);
  }

  /// Group-local rank of the tile that owns `virtual_atom`:
  /// (first prefix entry greater than it) - 1.
  template <typename partition_t>
  __device__ __forceinline__ tiles_t tile_accessor(atoms_t& virtual_atom,
                                                   partition_t& p) {
    const atoms_t* prefix = group_prefix(p);
    int first = 0;
    int count = get_length(p);
    while (count > 0) {
      const int half = count >> 1;
      if (!(virtual_atom < prefix[first + half])) {
        first += half + 1;
        count -= half + 1;
      } else {
        count = half;
      }
    }
    return static_cast<tiles_t>(first - 1);
  }

  template <typename partition_t>
  __device__ __forceinline__ bool is_valid_accessor(tiles_t& tile_id,
                                                    partition_t& p) {
    return tile_id < get_length(p);
  }

  /// Group-local tile rank -> real tile id.
  template <typename partition_t>
  __device__ __forceinline__ tiles_t tile_id(tiles_t& v_tile_id,
                                             partition_t& p) {
    return buffer.tiles_indices[v_tile_id + p.meta_group_rank() * p.size()];
  }

  /// Virtual atom -> real atom id inside `tile_id`.
  template <typename partition_t>
  __device__ __forceinline__ atoms_t atom_id(atoms_t& v_atom,
                                             tiles_t& tile_id,
                                             tiles_t& v_tile_id,
                                             partition_t& p) {
    return view_.tile_begin(tile_id) + v_atom - group_prefix(p)[v_tile_id];
  }

  __host__ __device__ const layout_t& layout() const { return view_; }

 private:
  __device__ __forceinline__ long long global_rank() const {
    return static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  }

  template <typename partition_t>
  __device__ __forceinline__ atoms_t* group_prefix(partition_t& p) {
    return buffer.atoms_offsets + p.meta_group_rank() * threads_per_tile;
  }

  /// Exclusive sum over the group. Sub-warp and warp groups: shuffle scan.
  /// Wider groups: per-warp shuffle scan, warp totals through shared memory.
  template <typename partition_t>
  __device__ __forceinline__ atoms_t group_exclusive_sum(atoms_t v,
                                                         partition_t& p) {
    const unsigned lane = threadIdx.x & 31u;
    constexpr unsigned seg = threads_per_tile < 32 ? threads_per_tile : 32;
    const unsigned seg_lane = lane & (seg - 1u);
    atoms_t inc = v;
#pragma unroll
    for (unsigned d = 1; d < seg; d <<= 1) {
      const atoms_t up = __shfl_up_sync(0xffffffffu, inc, d, seg);
      if (seg_lane >= d)
        inc += up;
    }
    atoms_t before = inc - v;
    if (threads_per_tile > 32) {
      const unsigned warp = threadIdx.x >> 5;
      if (lane == 31)
        buffer.warp_totals[warp] = inc;
      p.sync();
      constexpr unsigned warps_per_group = threads_per_tile / 32;
      const unsigned first_warp = p.meta_group_rank() * warps_per_group;
      atoms_t carry = 0;
      for (unsigned w = first_warp; w < warp; ++w)
        carry += buffer.warp_totals[w];
      before += carry;
    }
    return before;
  }

  layout_t view_;
};

template <std::size_t threads_per_block, typename tiles_t, typename atoms_t>
using warp_mapped =
    setup<algorithms_t::group_mapped, threads_per_block, 32, tiles_t, atoms_t>;

template <std::size_t threads_per_block, typename tiles_t, typename atoms_t>
using block_mapped = setup<algorithms_t::group_mapped,
                           threads_per_block,
                           threads_per_block,
                           tiles_t,
                           atoms_t>;

}  // namespace schedule
}  // namespace loops
