/**
 * @file search.hxx
 * @brief Merge-path diagonal search.
 *
 * Splits the merge of two sorted lists -- A = tile end offsets (length a_len)
 * and B = a counting sequence of atom ids (length b_len) -- at diagonal d:
 * returns (x, y), x + y = d, such that the first x tile-ends and first y atoms
 * precede the split. Same contract and corner cases as the reference's
 * search::_binary_search (reference include/loops/util/search.hxx:34-60):
 *   x in [max(d - b_len, 0), min(d, a_len)], first x with a[x] > b[d - x - 1];
 *   an inverted interval (d > a_len + b_len) yields (a_len, d - x_min);
 *   the result is a coordinate_t<unsigned int>.
 * Hand-rolled (no thrust), 64-bit interval arithmetic so tiles + atoms may
 * exceed 2^31 on this path.
 */
#pragma once

#include <loops/range.hxx>
#include <loops/container/coordinate.hxx>

namespace loops {
namespace search {

/// Core routine: `b` only needs operator[].
template <typename a_iterator_t, typename b_iterator_t>
LOOPS_HD coordinate_t<unsigned int> diagonal_split(long long diagonal,
                                                   a_iterator_t a,
                                                   b_iterator_t b,
                                                   long long a_len,
                                                   long long b_len) {
  long long lo = diagonal - b_len;
  if (lo < 0)
    lo = 0;
  long long hi = diagonal < a_len ? diagonal : a_len;
  const long long x_min = lo;
  while (lo < hi) {
    const long long mid = lo + ((hi - lo) >> 1);
    // Tile `mid` ends at or before the atom that sits opposite on the
    // diagonal -> the split lies further down the tile list.
    if (static_cast<long long>(a[mid]) <=
        static_cast<long long>(b[diagonal - mid - 1]))
      lo = mid + 1;
    else
      hi = mid;
  }
  if (hi < x_min)  // inverted interval: leave the cursor at its start
    lo = x_min;
  coordinate_t<unsigned int> c;
  c.x = static_cast<unsigned int>(lo < a_len ? lo : a_len);
  c.y = static_cast<unsigned int>(diagonal - lo);
  return c;
}

/// Reference spelling / argument order.
template <typename offset_t, typename xit_t, typename yit_t>
LOOPS_HD coordinate_t<unsigned int> _binary_search(const offset_t& diagonal,
                                                   const xit_t a,
                                                   const yit_t b,
                                                   const offset_t& a_len,
                                                   const offset_t& b_len) {
  return diagonal_split(static_cast<long long>(diagonal), a, b,
                        static_cast<long long>(a_len),
                        static_cast<long long>(b_len));
}

}  // namespace search
}  // namespace loops
