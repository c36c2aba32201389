/**
 * @file timer.hxx
 * @brief CUDA-event stopwatch (reference include/loops/util/timer.hxx:19-52),
 * recording on a caller-chosen stream instead of always stream 0.
 */
#pragma once
#include <cuda_runtime.h>

namespace loops {
namespace util {

struct timer_t {
  float time = 0.0f;
  cudaStream_t stream = 0;
  cudaEvent_t start_ = nullptr, stop_ = nullptr;

  explicit timer_t(cudaStream_t s = 0) : stream(s) {
    cudaEventCreate(&start_);
    cudaEventCreate(&stop_);
  }
  timer_t(const timer_t&) = delete;
  timer_t& operator=(const timer_t&) = delete;
  timer_t(timer_t&& o) noexcept : time(o.time), stream(o.stream), start_(o.start_), stop_(o.stop_) {
    o.start_ = o.stop_ = nullptr;
  }
  ~timer_t() {
    if (start_) cudaEventDestroy(start_);
    if (stop_) cudaEventDestroy(stop_);
  }

  void start() { cudaEventRecord(start_, stream); }
  float stop() {
    cudaEventRecord(stop_, stream);
    cudaEventSynchronize(stop_);
    cudaEventElapsedTime(&time, start_, stop_);
    return milliseconds();
  }
  float seconds() const { return time * 1e-3f; }
  float milliseconds() const { return time; }
};

}  // namespace util
}  // namespace loops
