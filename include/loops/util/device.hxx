/**
 * @file device.hxx
 * @brief Device queries under the reference's names (reference include/loops/util/device.hxx):
 * `device::set / get`, `properties_t`, `multi_processor_count`, `compute_capability`.
 * loops-b200 runs on sm_100a only and calls the CUDA runtime directly (no xpu shim); the
 * per-ordinal attribute cache is guarded so that several host threads (one per GPU) may
 * make their first query at the same time.
 */
#pragma once

#include <cuda_runtime.h>

#include <mutex>

namespace loops {
namespace device {

using device_id_t = int;

inline void set(device_id_t ordinal) { cudaSetDevice(ordinal); }
inline device_id_t get() {
  device_id_t d = 0;
  cudaGetDevice(&d);
  return d;
}

namespace detail {
struct attrs_t {
  int sm_count = 0, cc_major = 0, cc_minor = 0, max_smem_optin = 0;
  bool known = false;
};
inline const attrs_t& attributes(device_id_t ordinal) {
  static attrs_t table[64];
  static std::mutex mu;
  static const attrs_t none{};
  if (ordinal < 0 || ordinal >= 64) return none;
  std::lock_guard<std::mutex> lock(mu);
  attrs_t& a = table[ordinal];
  if (!a.known) {
    cudaDeviceGetAttribute(&a.sm_count, cudaDevAttrMultiProcessorCount, ordinal);
    cudaDeviceGetAttribute(&a.cc_major, cudaDevAttrComputeCapabilityMajor, ordinal);
    cudaDeviceGetAttribute(&a.cc_minor, cudaDevAttrComputeCapabilityMinor, ordinal);
    cudaDeviceGetAttribute(&a.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ordinal);
    a.known = true;
  }
  return a;
}
}  // namespace detail

/// Cheap value copy of the cached attributes of one device.
struct properties_t {
  int multi_processor_count = 0;
  int major = 0, minor = 0;
  int shared_memory_per_block_optin = 0;
  explicit properties_t(device_id_t ordinal = device::get()) {
    const detail::attrs_t& a = detail::attributes(ordinal);
    multi_processor_count = a.sm_count;
    major = a.cc_major;
    minor = a.cc_minor;
    shared_memory_per_block_optin = a.max_smem_optin;
  }
};

inline int multi_processor_count(device_id_t ordinal = device::get()) { return detail::attributes(ordinal).sm_count; }

/// major * 10 + minor (100 on B200).
inline int compute_capability(device_id_t ordinal = device::get()) {
  const detail::attrs_t& a = detail::attributes(ordinal);
  return a.cc_major * 10 + a.cc_minor;
}

}  // namespace device
}  // namespace loops
