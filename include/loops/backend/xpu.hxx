/**
 * @file xpu.hxx
 * @brief The `loops::xpu::*` runtime names user code written for the reference touches
 * (reference include/loops/backend/{xpu,cuda}.hxx: `xpu::stream_t`, `xpu::stream_synchronize`,
 * `xpu::memcpy`, events ...), bound DIRECTLY to the CUDA runtime. loops-b200 has one backend
 * (sm_100a); this is a set of aliases so that e.g. examples/spmv/custom_layout.cu:232 compiles,
 * not a vendor dispatch layer.
 */
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

namespace loops {
namespace xpu {

using error_t = cudaError_t;
using stream_t = cudaStream_t;
using event_t = cudaEvent_t;
using device_properties_t = cudaDeviceProp;
using memcpy_kind_t = cudaMemcpyKind;
using device_attribute_t = cudaDeviceAttr;

inline constexpr error_t success = cudaSuccess;
inline constexpr memcpy_kind_t memcpy_host_to_device = cudaMemcpyHostToDevice;
inline constexpr memcpy_kind_t memcpy_device_to_host = cudaMemcpyDeviceToHost;
inline constexpr memcpy_kind_t memcpy_device_to_device = cudaMemcpyDeviceToDevice;
inline constexpr device_attribute_t attr_multiprocessor_count = cudaDevAttrMultiProcessorCount;
inline constexpr device_attribute_t attr_compute_capability_major = cudaDevAttrComputeCapabilityMajor;
inline constexpr device_attribute_t attr_compute_capability_minor = cudaDevAttrComputeCapabilityMinor;
inline constexpr device_attribute_t attr_max_grid_dim_x = cudaDevAttrMaxGridDimX;

inline error_t set_device(int ordinal) { return cudaSetDevice(ordinal); }
inline error_t get_device(int* ordinal) { return cudaGetDevice(ordinal); }
inline error_t get_device_properties(device_properties_t* p, int ordinal) { return cudaGetDeviceProperties(p, ordinal); }
inline error_t device_get_attribute(int* value, device_attribute_t attr, int ordinal) {
  return cudaDeviceGetAttribute(value, attr, ordinal);
}
inline error_t device_synchronize() { return cudaDeviceSynchronize(); }
inline error_t malloc(void** ptr, std::size_t bytes) { return cudaMalloc(ptr, bytes); }
inline error_t free(void* ptr) { return cudaFree(ptr); }
inline error_t memcpy(void* dst, const void* src, std::size_t bytes, memcpy_kind_t kind) {
  return cudaMemcpy(dst, src, bytes, kind);
}
inline error_t stream_synchronize(stream_t stream = 0) { return cudaStreamSynchronize(stream); }
inline error_t event_create(event_t* e) { return cudaEventCreate(e); }
inline error_t event_destroy(event_t e) { return cudaEventDestroy(e); }
inline error_t event_record(event_t e, stream_t stream = 0) { return cudaEventRecord(e, stream); }
inline error_t event_synchronize(event_t e) { return cudaEventSynchronize(e); }
inline error_t event_elapsed_time(float* ms, event_t start, event_t stop) { return cudaEventElapsedTime(ms, start, stop); }
inline const char* get_error_string(error_t status) { return cudaGetErrorString(status); }

template <typename func_t>
inline error_t occupancy_max_active_blocks_per_multiprocessor(int* blocks, func_t kernel, int block_size,
                                                              std::size_t dynamic_smem_bytes) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kernel, block_size, dynamic_smem_bytes);
}

template <typename func_t>
inline error_t launch_cooperative_kernel(const func_t* kernel, std::size_t blocks, std::size_t threads, void** args,
                                         std::size_t shared_bytes, stream_t stream) {
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(static_cast<unsigned>(blocks)),
                                     dim3(static_cast<unsigned>(threads)), args, shared_bytes, stream);
}

}  // namespace xpu
}  // namespace loops
