// tests/cpp/catch2_mini/catch2/catch_test_macros.hpp -- TEST INFRASTRUCTURE.
// A small stand-in for Catch2 v3's <catch2/catch_test_macros.hpp> (Catch2 is not installed and
// cannot be fetched here), enough to compile the REFERENCE's own unit tests
// (/root/reference/unittests/test_*.cu) UNCHANGED against this repo's include/ tree:
// TEST_CASE, SECTION (each leaf section gets its own pass over the test case, like Catch2),
// CHECK / REQUIRE / CHECK_FALSE / REQUIRE_FALSE / CHECK_THROWS / REQUIRE_THROWS / INFO / CAPTURE,
// and a main() that runs every registered case and returns the number of failed ones.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>
#include <vector>

namespace catch_mini {
struct test_case { const char* name; void (*fn)(); };
inline std::vector<test_case>& registry() { static std::vector<test_case> r; return r; }
struct registrar { registrar(const char* n, void (*f)()) { registry().push_back({n, f}); } };
struct abort_case : std::exception {};
struct state_t {
  int failures = 0, assertions = 0;
  int target = 0;      // index of the top-level section this pass executes
  int seen = 0;        // top-level sections met so far in this pass
  int depth = 0;
};
inline state_t& state() { static state_t s; return s; }
inline void report(bool ok, const char* kind, const char* expr, const char* file, int line) {
  state().assertions++;
  if (!ok) { state().failures++; std::fprintf(stderr, "%s:%d: %s( %s ) FAILED\n", file, line, kind, expr); }
}
struct section_guard {
  bool run;
  explicit section_guard(const char*) {
    state_t& s = state();
    if (s.depth == 0) { run = (s.seen == s.target); s.seen++; } else run = true;   // nested sections all run
    if (run) s.depth++;
  }
  ~section_guard() { if (run) state().depth--; }
  explicit operator bool() const { return run; }
};
inline int run_all() {
  int failed_cases = 0;
  for (const test_case& t : registry()) {
    state_t& s = state();
    const int before = s.failures;
    s.target = 0;
    for (;;) {
      s.seen = 0; s.depth = 0;
      try { t.fn(); }
      catch (const abort_case&) {}
      catch (const std::exception& e) { s.failures++; std::fprintf(stderr, "%s: unexpected exception: %s\n", t.name, e.what()); }
      catch (...) { s.failures++; std::fprintf(stderr, "%s: unexpected exception\n", t.name); }
      if (s.target + 1 >= s.seen) break;    // no further top-level section to visit
      s.target++;
    }
    const bool ok = s.failures == before;
    std::printf("[%s] %s\n", ok ? "  OK  " : "FAILED", t.name);
    if (!ok) failed_cases++;
  }
  std::printf("test cases: %zu | failed: %d | assertions: %d\n", registry().size(), failed_cases, state().assertions);
  return failed_cases;
}
}  // namespace catch_mini

#define CATCH_MINI_CAT2(a, b) a##b
#define CATCH_MINI_CAT(a, b) CATCH_MINI_CAT2(a, b)
#define CATCH_MINI_TEST(fn, ...)                                                      \
  static void fn();                                                                   \
  static ::catch_mini::registrar CATCH_MINI_CAT(fn, _reg)(                            \
      [] { static const char* n[] = {__VA_ARGS__}; return n[0]; }(), &fn);            \
  static void fn()
#define TEST_CASE(...) CATCH_MINI_TEST(CATCH_MINI_CAT(catch_mini_case_, __COUNTER__), __VA_ARGS__)
#define SECTION(...) if (::catch_mini::section_guard CATCH_MINI_CAT(catch_mini_sec_, __COUNTER__){"" __VA_ARGS__})
#define CHECK(...) ::catch_mini::report(static_cast<bool>(__VA_ARGS__), "CHECK", #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK_FALSE(...) ::catch_mini::report(!static_cast<bool>(__VA_ARGS__), "CHECK_FALSE", #__VA_ARGS__, __FILE__, __LINE__)
#define REQUIRE(...)                                                                  \
  do {                                                                                \
    const bool catch_mini_ok = static_cast<bool>(__VA_ARGS__);                        \
    ::catch_mini::report(catch_mini_ok, "REQUIRE", #__VA_ARGS__, __FILE__, __LINE__); \
    if (!catch_mini_ok) throw ::catch_mini::abort_case();                             \
  } while (0)
#define REQUIRE_FALSE(...) REQUIRE(!(__VA_ARGS__))
#define CHECK_THROWS(...)                                                             \
  do {                                                                                \
    bool catch_mini_threw = false;                                                    \
    try { static_cast<void>(__VA_ARGS__); } catch (...) { catch_mini_threw = true; }  \
    ::catch_mini::report(catch_mini_threw, "CHECK_THROWS", #__VA_ARGS__, __FILE__, __LINE__); \
  } while (0)
#define REQUIRE_THROWS(...) CHECK_THROWS(__VA_ARGS__)
#define CHECK_NOTHROW(...)                                                            \
  do {                                                                                \
    bool catch_mini_threw = false;                                                    \
    try { static_cast<void>(__VA_ARGS__); } catch (...) { catch_mini_threw = true; }  \
    ::catch_mini::report(!catch_mini_threw, "CHECK_NOTHROW", #__VA_ARGS__, __FILE__, __LINE__); \
  } while (0)
#define INFO(...) do { } while (0)
#define CAPTURE(...) do { } while (0)
#define SUCCEED(...) ::catch_mini::report(true, "SUCCEED", "", __FILE__, __LINE__)
#define FAIL(...) do { ::catch_mini::report(false, "FAIL", "" #__VA_ARGS__, __FILE__, __LINE__); throw ::catch_mini::abort_case(); } while (0)

#ifndef CATCH_MINI_NO_MAIN
int main() { return ::catch_mini::run_all(); }
#endif
