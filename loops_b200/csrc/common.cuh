// loops_b200/csrc/common.cuh -- shared helpers for the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <loopsb.h>

namespace loopsb {

// Per-thread last error text (returned by loopsb_last_error()).
char* last_error_buffer();
void set_error(const char* fmt, ...);

#define LOOPSB_CUDA_TRY(expr)                                              \
  do {                                                                     \
    cudaError_t e__ = (expr);                                              \
    if (e__ != cudaSuccess) {                                              \
      ::loopsb::set_error("%s failed: %s (%s:%d)", #expr,                  \
                          cudaGetErrorString(e__), __FILE__, __LINE__);    \
      return LOOPSB_ERR_CUDA;                                              \
    }                                                                      \
  } while (0)

#define LOOPSB_REQUIRE(cond, msg)                                          \
  do {                                                                     \
    if (!(cond)) {                                                         \
      ::loopsb::set_error("invalid argument: %s (%s)", msg, #cond);        \
      return LOOPSB_ERR_INVALID;                                           \
    }                                                                      \
  } while (0)

struct device_props {
  int sm_count = 0;
  int cc_major = 0;
  int cc_minor = 0;
  int max_smem_optin = 0;
  bool valid = false;
};
// Cached per device ordinal; returns nullptr (and sets the error) if no device.
const device_props* current_device();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool is_offsets_kind(int kind) {
  return kind == LOOPSB_LAYOUT_CSR || kind == LOOPSB_LAYOUT_CSC ||
         kind == LOOPSB_LAYOUT_BCSR;
}
inline bool is_pitch_kind(int kind) {
  return kind == LOOPSB_LAYOUT_ELL || kind == LOOPSB_LAYOUT_DIA;
}

}  // namespace loopsb
