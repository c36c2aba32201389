/**
 * @file math.hxx
 * @brief Integer helpers (reference include/loops/util/math.hxx:19-29).
 */
#pragma once
#include <loops/range.hxx>
namespace loops {
namespace math {
/// Smallest q with q * d >= n, in the numerator's type.
template <class numerator_t, class denominator_t>
LOOPS_HD constexpr numerator_t ceil_div(numerator_t const& n,
                                        denominator_t const& d) {
  return static_cast<numerator_t>((n / d) + ((n % d) ? 1 : 0));
}
}  // namespace math
}  // namespace loops
