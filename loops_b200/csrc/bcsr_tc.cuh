// loops_b200/csrc/bcsr_tc.cuh -- BCSR 4x4 bf16 SpMV on the tcgen05 tensor
// cores (placeholder until the tensor-core kernel lands).
#pragma once
#include "common.cuh"

namespace loopsb {
namespace bcsr_tc {

struct plan_data {
  long long bytes = 0;
};

inline int create(plan_data** out, const loopsb_layout_t*, int, cudaStream_t) {
  *out = new plan_data();
  return LOOPSB_OK;
}
inline void destroy(plan_data* p) { delete p; }
inline long long workspace_bytes(const plan_data* p) { return p ? p->bytes : 0; }
inline int run(plan_data*, const loopsb_layout_t*, const uint16_t*, const int32_t*,
               const uint16_t*, float*, int32_t, cudaStream_t) {
  set_error("BCSR 4x4 bf16 tcgen05 kernel not built yet");
  return LOOPSB_ERR_UNSUPPORTED;
}

}  // namespace bcsr_tc
}  // namespace loopsb
