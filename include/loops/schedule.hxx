/**
 * @file schedule.hxx
 * @brief Load-balancing schedules: the `schedule::setup<...>` entry point.
 *
 * Same enum order and template parameter list as the reference (reference
 * include/loops/schedule.hxx:26-32,55-63) so user kernels written against
 * `schedule::setup<scheme, TPB, TPT|IPT, tiles_t, atoms_t, tile_size_t,
 * atom_size_t, layout_t>` compile unchanged. The four specialisations live in
 * loops/schedule/. They are the generic, layout-agnostic device path (and the
 * instrument `loopsb_emit_schedule` records); the tuned sm_100a SpMV kernels
 * behind include/loopsb.h use the same partition arithmetic.
 */
#pragma once

#include <cstddef>

#include <loops/container/layout.hxx>

namespace loops {
namespace schedule {

enum algorithms_t {
  merge_path_flat,  ///< even split of (tiles + atoms) over blocks, then threads
  work_oriented,    ///< even split of (tiles + atoms) over all threads
  thread_mapped,    ///< one thread per tile
  group_mapped,     ///< a group of threads shares the atoms of its tiles
  bucketing,        ///< reserved (unimplemented upstream as well)
};

template <algorithms_t scheme,
          std::size_t threads_per_block,
          std::size_t threads_per_tile,
          typename tiles_t,
          typename atoms_t,
          typename tile_size_t = std::size_t,
          typename atom_size_t = std::size_t,
          typename layout_type = layout::csr<tiles_t, atoms_t>>
class setup;

}  // namespace schedule
}  // namespace loops

#include <loops/schedule/thread_mapped.hxx>
#include <loops/schedule/group_mapped.hxx>
#include <loops/schedule/work_oriented.hxx>
#include <loops/schedule/merge_path_flat.hxx>
