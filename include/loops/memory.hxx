/**
 * @file memory.hxx
 * @brief Memory spaces (reference include/loops/memory.hxx:22-38).
 */
#pragma once
#include <thrust/device_ptr.h>

namespace loops {
namespace memory {

enum memory_space_t { device, host, managed };

template <typename type_t>
inline type_t* raw_pointer_cast(thrust::device_ptr<type_t> p) {
  return thrust::raw_pointer_cast(p);
}
template <typename type_t>
__host__ __device__ inline type_t* raw_pointer_cast(type_t* p) {
  return p;
}

}  // namespace memory
}  // namespace loops
