/**
 * @file coordinate.hxx
 * @brief (x, y) pair on the merge grid: x counts tiles, y counts atoms
 * (reference include/loops/container/coordinate.hxx:14-18).
 */
#pragma once
namespace loops {
template <typename index_t>
struct coordinate_t {
  index_t x;
  index_t y;
};
}  // namespace loops
