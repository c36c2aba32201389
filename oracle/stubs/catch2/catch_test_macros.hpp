// oracle/stubs/catch2/catch_test_macros.hpp -- TEST INFRASTRUCTURE.
// Catch2 is not installed (and cannot be fetched); the reference's fixture
// header unittests/test_helpers.hxx only needs REQUIRE from it. This stub lets
// oracle/ref_battery.cu include that header UNMODIFIED to regenerate the
// reference's own SpMV test battery.
#pragma once
#include <cstdio>
#include <cstdlib>
#define REQUIRE(cond)                                                    \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::fprintf(stderr, "REQUIRE failed: %s\n", #cond);               \
      std::abort();                                                      \
    }                                                                    \
  } while (0)
#define CHECK(cond) REQUIRE(cond)
#define INFO(msg) do { } while (0)
