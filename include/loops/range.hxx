/**
 * @file range.hxx
 * @brief Range-for helpers for loops-b200 kernels.
 *
 * API parity with the reference's `loops::range(begin, end)[.step(s)]`
 * (reference include/loops/range.hxx:52-116,181-189): a half-open integer
 * interval usable in `for (auto i : ...)`, optionally strided. The strided
 * form terminates as soon as the cursor reaches OR PASSES the end, so ends
 * that are not a multiple of the stride are safe (reference :78-80).
 *
 * One class serves both forms here (stride 1 is just a stride), the cursor is
 * a plain value type and nothing depends on <iterator>.
 */
#pragma once
#include <initializer_list>
#include <type_traits>
#include <utility>

#include <cstddef>

#ifndef LOOPS_HD
#if defined(__CUDACC__)
#define LOOPS_HD __host__ __device__ __forceinline__
#define LOOPS_D __device__ __forceinline__
#else
#define LOOPS_HD inline
#define LOOPS_D inline
#endif
#endif

namespace loops {

/// Strided half-open interval [first, last). `operator!=` means "not done".
template <typename T>
class step_range {
 public:
  class cursor {
   public:
    LOOPS_HD cursor(T at, T by) : at_(at), by_(by) {}
    LOOPS_HD T operator*() const { return at_; }
    LOOPS_HD cursor& operator++() {
      at_ += by_;
      return *this;
    }
    /// "Reached the sentinel": past-or-equal for forward strides,
    /// strictly-below for backward ones.
    LOOPS_HD bool operator!=(cursor const& sentinel) const {
      return by_ > T(0) ? at_ < sentinel.at_ : !(at_ < sentinel.at_);
    }
    LOOPS_HD bool operator==(cursor const& sentinel) const {
      return !(*this != sentinel);
    }

   private:
    T at_;
    T by_;
  };

  LOOPS_HD step_range(T first, T last, T by = T(1))
      : first_(first), last_(last), by_(by) {}

  LOOPS_HD cursor begin() const { return cursor(first_, by_); }
  LOOPS_HD cursor end() const { return cursor(last_, by_); }

  /// Re-stride: `range(a, b).step(s)`.
  LOOPS_HD step_range step(T by) const { return step_range(first_, last_, by); }

 private:
  T first_, last_, by_;
};

/// Unbounded counterpart: `range(begin)` / `range(begin).step(s)`.
template <typename T>
class open_range {
 public:
  class cursor {
   public:
    LOOPS_HD cursor(T at, T by) : at_(at), by_(by) {}
    LOOPS_HD T operator*() const { return at_; }
    LOOPS_HD cursor& operator++() {
      at_ += by_;
      return *this;
    }
    LOOPS_HD bool operator!=(cursor const&) const { return true; }

   private:
    T at_, by_;
  };
  LOOPS_HD explicit open_range(T first, T by = T(1)) : first_(first), by_(by) {}
  LOOPS_HD cursor begin() const { return cursor(first_, by_); }
  LOOPS_HD cursor end() const { return cursor(first_, by_); }
  LOOPS_HD open_range step(T by) const { return open_range(first_, by); }

 private:
  T first_, by_;
};

template <typename T>
LOOPS_HD step_range<T> range(T begin, T end) {
  return step_range<T>(begin, end, T(1));
}

template <typename T>
LOOPS_HD open_range<T> range(T begin) {
  return open_range<T>(begin);
}

/// Spelling used by the reference for the strided proxy type.
template <typename T>
using step_range_t = step_range<T>;

/// `for (auto i : indices(arr))` over a C array.
template <typename T, std::size_t N>
LOOPS_HD step_range<std::size_t> indices(T (&)[N]) {
  return step_range<std::size_t>(0, N, 1);
}

namespace traits {
/// Does `C` have a `size()` returning an integer?
template <typename C, typename = void>
struct has_size : std::false_type {};
template <typename C>
struct has_size<C, std::enable_if_t<std::is_integral<decltype(std::declval<const C&>().size())>::value>>
    : std::true_type {};
}  // namespace traits

/// `for (auto i : indices(container))` -- host only (a container's size() is a host function).
template <typename C, typename = std::enable_if_t<traits::has_size<C>::value>>
__host__ auto indices(const C& cont) -> step_range<decltype(cont.size())> {
  using size_type = decltype(cont.size());
  return step_range<size_type>(size_type(0), cont.size(), size_type(1));
}

/// `for (auto i : indices({a, b, c}))`.
template <typename T>
LOOPS_HD step_range<typename std::initializer_list<T>::size_type> indices(std::initializer_list<T>&& cont) {
  using size_type = typename std::initializer_list<T>::size_type;
  return step_range<size_type>(size_type(0), cont.size(), size_type(1));
}

}  // namespace loops
