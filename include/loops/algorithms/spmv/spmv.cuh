/**
 * @file spmv.cuh
 * @brief `loops::algorithms::spmv::*` host entry points -- same names,
 * argument lists and return types as the reference
 * (include/loops/algorithms/spmv/{merge_path_flat,work_oriented,thread_mapped,
 * group_mapped,coo_thread_mapped,ell_thread_mapped,ell_merge_path,
 * bcsr_thread_mapped}.cuh), each a thin call into the C ABI of
 * include/loopsb.h (libloopsb200.so) where the sm_100a kernels live.
 *
 * Like the reference wrappers they are synchronous (the stream is synchronised
 * before returning); the ones that return `util::timer_t` time the SpMV
 * launches only, excluding the per-matrix preprocess (reference
 * merge_path_flat.cuh:111 precedes :121-122). Differences, both relaxations:
 * `y` does not have to be zeroed by the caller, and failures surface as
 * `error::exception_t` instead of being dropped.
 */
#pragma once

#include <cstdlib>

#include <cuda_runtime.h>

#include <loops/error.hxx>
#include <loops/schedule.hxx>
#include <loops/container/formats.hxx>
#include <loops/container/vector.hxx>
#include <loops/util/timer.hxx>
#include <loops/algorithms/spmv/launch_box.hxx>
#include <loopsb.h>

namespace loops {
namespace algorithms {
namespace spmv {
namespace detail {

/// RAII handle of a loopsb plan (the reference's preprocess_t).
struct plan_guard {
  loopsb_plan_t* p = nullptr;
  plan_guard(const loopsb_layout_t& lay, int schedule, cudaStream_t stream) {
    error::throw_if_status(loopsb_plan_create(&p, &lay, schedule, stream), "loopsb_plan_create");
  }
  ~plan_guard() { loopsb_plan_destroy(p); }
  plan_guard(const plan_guard&) = delete;
  plan_guard& operator=(const plan_guard&) = delete;
};

template <typename vec_t>
auto* raw(vec_t& v) { return thrust::raw_pointer_cast(v.data()); }

/// One SpMV through the plan cached on the container (created on the first call). `k_rows` is
/// the third array the plan's validity hangs on (CSR offsets / COO row ids / null for ELL).
inline util::timer_t run(::loops::detail::plan_cache_t& cache, const loopsb_layout_t& lay, int schedule,
                         const float* values, const int* cols, const int* rows_idx, const void* k_rows,
                         const float* x, float* y, std::size_t nrows, std::size_t ncols, std::size_t nnzs,
                         cudaStream_t stream) {
  auto& e = cache.get(schedule, lay, k_rows, cols, values, nrows, ncols, nnzs, std::size_t(lay.pitch), stream);
  util::timer_t timer(stream);
  timer.start();
  error::throw_if_status(loopsb_spmv_f32(e.plan, values, cols, rows_idx, x, y, static_cast<int32_t>(nrows),
                                         static_cast<int32_t>(ncols), stream),
                         "loopsb_spmv_f32");
  cudaStreamSynchronize(stream);
  timer.stop();
  ++e.calls;
  return timer;
}

/// LOOPSB_TILED: "0" = the cached merge_path_flat plan never takes a band-tiled copy of the matrix,
/// "1" = on the first call, anything else / unset = once the calls made on the plan reach the
/// break-even count of loopsb_plan_tile_breakeven (build time / per-call saving on B200).
/// One SpMV on a plan made for this call only (layout views that are not cached on a container).
inline util::timer_t run_once(const loopsb_layout_t& lay, int schedule, const float* values, const int* cols,
                              const int* rows_idx, const float* x, float* y, std::size_t nrows, std::size_t ncols,
                              cudaStream_t stream) {
  plan_guard plan(lay, schedule, stream);
  util::timer_t timer(stream);
  timer.start();
  error::throw_if_status(loopsb_spmv_f32(plan.p, values, cols, rows_idx, x, y, static_cast<int32_t>(nrows),
                                         static_cast<int32_t>(ncols), stream),
                         "loopsb_spmv_f32");
  cudaStreamSynchronize(stream);
  timer.stop();
  return timer;
}

inline util::timer_t run_coo(coo_t<int, float>& coo, int schedule, vector_t<float>& x, vector_t<float>& y,
                             cudaStream_t stream) {
  layout::coo<int, int> lay(static_cast<int>(coo.nnzs));
  return run(coo.plans(), lay.descriptor(), schedule, raw(coo.values), raw(coo.col_indices), raw(coo.row_indices),
             raw(coo.row_indices), raw(x), raw(y), coo.rows, coo.cols, coo.nnzs, stream);
}

inline util::timer_t run_ell(ell_t<int, float>& ell, int schedule, vector_t<float>& x, vector_t<float>& y,
                             cudaStream_t stream) {
  return run(ell.plans(), ell.layout().descriptor(), schedule, raw(ell.values), raw(ell.indices), nullptr, nullptr,
             raw(x), raw(y), ell.rows, ell.cols, ell.nnzs, stream);
}

inline int tiling_mode() {
  const char* e = std::getenv("LOOPSB_TILED");
  if (e && e[0] == '0' && e[1] == 0) return 0;
  if (e && e[0] == '1' && e[1] == 0) return 1;
  return 2;
}

}  // namespace detail

/**
 * reference algorithms/spmv/merge_path_flat.cuh:96-139. The plan (merge-path coordinates) is
 * cached on `csr` (container/formats.hxx: detail::plan_cache_t), so calls after the first
 * allocate nothing; once the calls made on it reach the break-even count of the band-tiled
 * copy (loopsb_plan_tile_breakeven; LOOPSB_TILED=1 or csr.plans().tiling = 1: at once, 0: never) the plan takes that copy
 * and later calls run the band-tiled kernel. The copy holds VALUES: after writing csr.values /
 * csr.indices in place call csr.values_changed(), after writing csr.offsets in place
 * csr.structure_changed(); re-assigned arrays are noticed by themselves.
 */
inline util::timer_t merge_path_flat(csr_t<int, int, float>& csr, vector_t<float>& x,
                                     vector_t<float>& y, cudaStream_t stream = 0) {
  const loopsb_layout_t lay = csr.layout().descriptor();
  const float* values = detail::raw(csr.values);
  const int* indices = detail::raw(csr.indices);
  auto& e = csr.plans().get(LOOPSB_SCHED_MERGE_PATH_FLAT, lay, detail::raw(csr.offsets), indices, values, csr.rows,
                            csr.cols, csr.nnzs, std::size_t(lay.pitch), stream);
  const int mode = csr.plans().tiling >= 0 ? csr.plans().tiling : detail::tiling_mode();
  const bool forced = mode == 1;
  if (!e.tiled && mode != 0 && csr.nnzs > 0 && !(forced ? e.declined_forced : e.declined)) {
    long long after = 0;
    if (!forced) {
      if (e.tile_after == -2) {
        int64_t n = -1;
        error::throw_if_status(loopsb_plan_tile_breakeven(e.plan, static_cast<int32_t>(csr.cols), &n),
                               "loopsb_plan_tile_breakeven");
        e.tile_after = static_cast<long long>(n);
      }
      after = e.tile_after;
    }
    if (after >= 0 && e.calls >= after) {
      const int rc = loopsb_plan_tile_csr(e.plan, indices, values, static_cast<int32_t>(csr.cols),
                                          forced ? LOOPSB_TILE_FORCE : 0, stream);
      if (rc != LOOPSB_OK && rc != LOOPSB_ERR_UNSUPPORTED) error::throw_if_status(rc, "loopsb_plan_tile_csr");
      e.tiled = rc == LOOPSB_OK;
      if (!e.tiled) (forced ? e.declined_forced : e.declined) = true;   // do not ask again for these arrays
    }
  }
  return detail::run(csr.plans(), lay, LOOPSB_SCHED_MERGE_PATH_FLAT, values, indices, nullptr,
                     detail::raw(csr.offsets), detail::raw(x), detail::raw(y), csr.rows, csr.cols, csr.nnzs, stream);
}

/**
 * @brief A merge_path_flat plan kept across calls -- what the reference's
 * `schedule::merge_path::preprocess_t` is to one launch
 * (schedule/merge_path_flat.hxx:92-172), made re-usable for iterative solvers.
 * With `band_tiled` (default) the plan also asks the library for a band-tiled
 * copy of the matrix (loopsb_plan_tile_csr); the library declines when its cost
 * model prefers the plain CSR kernel, `force` overrides that. The csr_t must
 * outlive the plan and keep its arrays (the copy is keyed by their addresses).
 *
 *     algorithms::spmv::merge_path_plan_t plan(csr);      // once per matrix
 *     for (...) plan(x, y);                                // every iteration
 */
class merge_path_plan_t {
 public:
  explicit merge_path_plan_t(csr_t<int, int, float>& csr, cudaStream_t stream = 0, bool band_tiled = true,
                             bool force = false)
      : csr_(csr), plan_(csr.layout().descriptor(), LOOPSB_SCHED_MERGE_PATH_FLAT, stream) {
    if (band_tiled) {
      const int rc = loopsb_plan_tile_csr(plan_.p, detail::raw(csr.indices), detail::raw(csr.values),
                                          static_cast<int32_t>(csr.cols), force ? LOOPSB_TILE_FORCE : 0, stream);
      if (rc != LOOPSB_OK && rc != LOOPSB_ERR_UNSUPPORTED) error::throw_if_status(rc, "loopsb_plan_tile_csr");
      tiled_ = rc == LOOPSB_OK;
    }
  }
  /// True when SpMV calls on this plan run the band-tiled kernel.
  bool band_tiled() const { return tiled_; }
  /// y = A x; synchronous like the reference wrappers, times the SpMV launch only.
  util::timer_t operator()(vector_t<float>& x, vector_t<float>& y, cudaStream_t stream = 0) {
    util::timer_t timer(stream);
    timer.start();
    error::throw_if_status(
        loopsb_spmv_f32(plan_.p, detail::raw(csr_.values), detail::raw(csr_.indices), nullptr, detail::raw(x),
                        detail::raw(y), static_cast<int32_t>(csr_.rows), static_cast<int32_t>(csr_.cols), stream),
        "loopsb_spmv_f32");
    cudaStreamSynchronize(stream);
    timer.stop();
    return timer;
  }

 private:
  csr_t<int, int, float>& csr_;
  detail::plan_guard plan_;
  bool tiled_ = false;
};

inline void work_oriented(csr_t<int, int, float>& csr, vector_t<float>& x, vector_t<float>& y,
                          cudaStream_t stream = 0) {
  detail::run(csr.plans(), csr.layout().descriptor(), LOOPSB_SCHED_WORK_ORIENTED, detail::raw(csr.values),
              detail::raw(csr.indices), nullptr, detail::raw(csr.offsets), detail::raw(x), detail::raw(y), csr.rows,
              csr.cols, csr.nnzs, stream);
}

inline void thread_mapped(csr_t<int, int, float>& csr, vector_t<float>& x, vector_t<float>& y,
                          cudaStream_t stream = 0) {
  detail::run(csr.plans(), csr.layout().descriptor(), LOOPSB_SCHED_THREAD_MAPPED, detail::raw(csr.values),
              detail::raw(csr.indices), nullptr, detail::raw(csr.offsets), detail::raw(x), detail::raw(y), csr.rows,
              csr.cols, csr.nnzs, stream);
}

inline void group_mapped(csr_t<int, int, float>& csr, vector_t<float>& x, vector_t<float>& y,
                         cudaStream_t stream = 0) {
  detail::run(csr.plans(), csr.layout().descriptor(), LOOPSB_SCHED_GROUP_MAPPED, detail::raw(csr.values),
              detail::raw(csr.indices), nullptr, detail::raw(csr.offsets), detail::raw(x), detail::raw(y), csr.rows,
              csr.cols, csr.nnzs, stream);
}

inline util::timer_t coo_thread_mapped(coo_t<int, float>& coo, vector_t<float>& x, vector_t<float>& y,
                                       cudaStream_t stream = 0) {
  return detail::run_coo(coo, LOOPSB_SCHED_THREAD_MAPPED, x, y, stream);
}

inline void ell_thread_mapped(ell_t<int, float>& ell, vector_t<float>& x, vector_t<float>& y,
                              cudaStream_t stream = 0) {
  detail::run_ell(ell, LOOPSB_SCHED_THREAD_MAPPED, x, y, stream);
}

inline util::timer_t ell_merge_path(ell_t<int, float>& ell, vector_t<float>& x, vector_t<float>& y,
                                    cudaStream_t stream = 0) {
  return detail::run_ell(ell, LOOPSB_SCHED_MERGE_PATH_FLAT, x, y, stream);
}

// ---- the five (schedule x layout) cells of BASELINE configs[2] the reference has no kernel for:
// schedule::setup<scheme, ..., layout> over the COO / ELL view with the format's per-atom body
// (loops_b200/csrc/spmv_generic.cu). New entry points, same argument meaning. ----
inline util::timer_t coo_group_mapped(coo_t<int, float>& coo, vector_t<float>& x, vector_t<float>& y,
                                      cudaStream_t stream = 0) {
  return detail::run_coo(coo, LOOPSB_SCHED_GROUP_MAPPED, x, y, stream);
}
inline util::timer_t coo_work_oriented(coo_t<int, float>& coo, vector_t<float>& x, vector_t<float>& y,
                                       cudaStream_t stream = 0) {
  return detail::run_coo(coo, LOOPSB_SCHED_WORK_ORIENTED, x, y, stream);
}
inline util::timer_t coo_merge_path(coo_t<int, float>& coo, vector_t<float>& x, vector_t<float>& y,
                                    cudaStream_t stream = 0) {
  return detail::run_coo(coo, LOOPSB_SCHED_MERGE_PATH_FLAT, x, y, stream);
}
inline util::timer_t ell_group_mapped(ell_t<int, float>& ell, vector_t<float>& x, vector_t<float>& y,
                                      cudaStream_t stream = 0) {
  return detail::run_ell(ell, LOOPSB_SCHED_GROUP_MAPPED, x, y, stream);
}
inline util::timer_t ell_work_oriented(ell_t<int, float>& ell, vector_t<float>& x, vector_t<float>& y,
                                       cudaStream_t stream = 0) {
  return detail::run_ell(ell, LOOPSB_SCHED_WORK_ORIENTED, x, y, stream);
}

/// reference algorithms/spmv/original.cuh:55-72 -- the plain one-thread-per-row kernel; same
/// arithmetic as thread_mapped (sequential per row), so it is the same sm_100a kernel.
inline void original(csr_t<int, int, float>& csr, vector_t<float>& x, vector_t<float>& y,
                     cudaStream_t stream = 0) {
  thread_mapped(csr, x, y, stream);
}

/// reference algorithms/spmv/csc_thread_mapped.cuh:54-84 (one thread per column, atomics into y).
inline util::timer_t csc_thread_mapped(csc_t<int, int, float>& csc, vector_t<float>& x, vector_t<float>& y,
                                       cudaStream_t stream = 0) {
  return detail::run_once(csc.layout().descriptor(), LOOPSB_SCHED_THREAD_MAPPED, detail::raw(csc.values),
                     detail::raw(csc.indices), nullptr, detail::raw(x), detail::raw(y), csc.rows, csc.cols, stream);
}

/// reference algorithms/spmv/dia_thread_mapped.cuh:65-98 (one thread per row over the stored diagonals).
inline util::timer_t dia_thread_mapped(dia_t<int, int, float>& dia, vector_t<float>& x, vector_t<float>& y,
                                       cudaStream_t stream = 0) {
  util::timer_t timer(stream);
  timer.start();
  error::throw_if_status(
      loopsb_spmv_dia_f32(static_cast<int32_t>(dia.rows), static_cast<int32_t>(dia.cols),
                          static_cast<int64_t>(dia.stride), static_cast<int32_t>(dia.num_diagonals),
                          detail::raw(dia.diag_offsets), detail::raw(dia.values), detail::raw(x), detail::raw(y),
                          stream),
      "loopsb_spmv_dia_f32");
  cudaStreamSynchronize(stream);
  timer.stop();
  return timer;
}

/// reference algorithms/spmv/flat_partitioned.cuh:73-107: thread_mapped over
/// layout::flat_uniform_occupancy<K, layout::csr> (windows of K nonzeros).
template <std::size_t K = 8>
util::timer_t flat_partitioned(csr_t<int, int, float>& csr, vector_t<float>& x, vector_t<float>& y,
                               cudaStream_t stream = 0) {
  layout::flat_uniform_occupancy<K, layout::csr<int, int>> lay(csr.layout());
  return detail::run_once(lay.descriptor(), LOOPSB_SCHED_THREAD_MAPPED, detail::raw(csr.values), detail::raw(csr.indices),
                          nullptr, detail::raw(x), detail::raw(y), csr.rows, csr.cols, stream);
}

template <std::size_t R, std::size_t C>
util::timer_t bcsr_thread_mapped(bcsr_t<R, C, int, int, float>& bcsr, vector_t<float>& x,
                                 vector_t<float>& y, cudaStream_t stream = 0) {
  const loopsb_layout_t lay = bcsr.layout().descriptor();
  util::timer_t timer(stream);
  timer.start();
  error::throw_if_status(
      loopsb_spmv_bcsr_f32(int32_t(R), int32_t(C), &lay, detail::raw(bcsr.values),
                           detail::raw(bcsr.block_col_indices), detail::raw(x), detail::raw(y),
                           static_cast<int32_t>(bcsr.rows), stream),
      "loopsb_spmv_bcsr_f32");
  cudaStreamSynchronize(stream);
  timer.stop();
  return timer;
}

// ---- fp64: the reference's entry points are templates on type_t and
// its examples are also built for double (examples/spmv/CMakeLists.txt:29) ----
namespace detail {
inline util::timer_t run_f64(csr_t<int, int, double>& csr, int schedule, vector_t<double>& x, vector_t<double>& y,
                             cudaStream_t stream) {
  const loopsb_layout_t lay = csr.layout().descriptor();
  util::timer_t timer(stream);
  timer.start();
  error::throw_if_status(loopsb_spmv_f64(&lay, schedule, raw(csr.values), raw(csr.indices), raw(x), raw(y),
                                         static_cast<int32_t>(csr.rows), static_cast<int32_t>(csr.cols), stream),
                         "loopsb_spmv_f64");
  cudaStreamSynchronize(stream);
  timer.stop();
  return timer;
}
}  // namespace detail

inline util::timer_t merge_path_flat(csr_t<int, int, double>& csr, vector_t<double>& x, vector_t<double>& y,
                                     cudaStream_t stream = 0) {
  return detail::run_f64(csr, LOOPSB_SCHED_MERGE_PATH_FLAT, x, y, stream);
}
inline void work_oriented(csr_t<int, int, double>& csr, vector_t<double>& x, vector_t<double>& y,
                          cudaStream_t stream = 0) {
  detail::run_f64(csr, LOOPSB_SCHED_WORK_ORIENTED, x, y, stream);
}
inline void thread_mapped(csr_t<int, int, double>& csr, vector_t<double>& x, vector_t<double>& y,
                          cudaStream_t stream = 0) {
  detail::run_f64(csr, LOOPSB_SCHED_THREAD_MAPPED, x, y, stream);
}
inline void group_mapped(csr_t<int, int, double>& csr, vector_t<double>& x, vector_t<double>& y,
                         cudaStream_t stream = 0) {
  detail::run_f64(csr, LOOPSB_SCHED_GROUP_MAPPED, x, y, stream);
}

// ---- fp64 for the other in-tree formats (every reference example is also built as .f64) ----
namespace detail {
inline util::timer_t run_layout_f64(const loopsb_layout_t& lay, int schedule, const double* values, const int* cols,
                                    const int* rows_idx, const double* x, double* y, std::size_t nrows,
                                    std::size_t ncols, cudaStream_t stream) {
  util::timer_t timer(stream);
  timer.start();
  error::throw_if_status(loopsb_spmv_layout_f64(&lay, schedule, values, cols, rows_idx, x, y, static_cast<int32_t>(nrows),
                                                static_cast<int32_t>(ncols), stream),
                         "loopsb_spmv_layout_f64");
  cudaStreamSynchronize(stream);
  timer.stop();
  return timer;
}
}  // namespace detail

inline void original(csr_t<int, int, double>& csr, vector_t<double>& x, vector_t<double>& y, cudaStream_t stream = 0) {
  thread_mapped(csr, x, y, stream);
}
inline util::timer_t coo_thread_mapped(coo_t<int, double>& coo, vector_t<double>& x, vector_t<double>& y,
                                       cudaStream_t stream = 0) {
  layout::coo<int, int> lay(static_cast<int>(coo.nnzs));
  return detail::run_layout_f64(lay.descriptor(), LOOPSB_SCHED_THREAD_MAPPED, detail::raw(coo.values),
                                detail::raw(coo.col_indices), detail::raw(coo.row_indices), detail::raw(x),
                                detail::raw(y), coo.rows, coo.cols, stream);
}
inline void ell_thread_mapped(ell_t<int, double>& ell, vector_t<double>& x, vector_t<double>& y,
                              cudaStream_t stream = 0) {
  detail::run_layout_f64(ell.layout().descriptor(), LOOPSB_SCHED_THREAD_MAPPED, detail::raw(ell.values),
                         detail::raw(ell.indices), nullptr, detail::raw(x), detail::raw(y), ell.rows, ell.cols, stream);
}
inline util::timer_t ell_merge_path(ell_t<int, double>& ell, vector_t<double>& x, vector_t<double>& y,
                                    cudaStream_t stream = 0) {
  return detail::run_layout_f64(ell.layout().descriptor(), LOOPSB_SCHED_MERGE_PATH_FLAT, detail::raw(ell.values),
                                detail::raw(ell.indices), nullptr, detail::raw(x), detail::raw(y), ell.rows, ell.cols,
                                stream);
}
inline util::timer_t csc_thread_mapped(csc_t<int, int, double>& csc, vector_t<double>& x, vector_t<double>& y,
                                       cudaStream_t stream = 0) {
  return detail::run_layout_f64(csc.layout().descriptor(), LOOPSB_SCHED_THREAD_MAPPED, detail::raw(csc.values),
                                detail::raw(csc.indices), nullptr, detail::raw(x), detail::raw(y), csc.rows, csc.cols,
                                stream);
}
template <std::size_t K = 8>
util::timer_t flat_partitioned(csr_t<int, int, double>& csr, vector_t<double>& x, vector_t<double>& y,
                               cudaStream_t stream = 0) {
  layout::flat_uniform_occupancy<K, layout::csr<int, int>> lay(csr.layout());
  return detail::run_layout_f64(lay.descriptor(), LOOPSB_SCHED_THREAD_MAPPED, detail::raw(csr.values),
                                detail::raw(csr.indices), nullptr, detail::raw(x), detail::raw(y), csr.rows, csr.cols,
                                stream);
}
inline util::timer_t dia_thread_mapped(dia_t<int, int, double>& dia, vector_t<double>& x, vector_t<double>& y,
                                       cudaStream_t stream = 0) {
  util::timer_t timer(stream);
  timer.start();
  error::throw_if_status(
      loopsb_spmv_dia_f64(static_cast<int32_t>(dia.rows), static_cast<int32_t>(dia.cols),
                          static_cast<int64_t>(dia.stride), static_cast<int32_t>(dia.num_diagonals),
                          detail::raw(dia.diag_offsets), detail::raw(dia.values), detail::raw(x), detail::raw(y), stream),
      "loopsb_spmv_dia_f64");
  cudaStreamSynchronize(stream);
  timer.stop();
  return timer;
}
template <std::size_t R, std::size_t C>
util::timer_t bcsr_thread_mapped(bcsr_t<R, C, int, int, double>& bcsr, vector_t<double>& x, vector_t<double>& y,
                                 cudaStream_t stream = 0) {
  const loopsb_layout_t lay = bcsr.layout().descriptor();
  util::timer_t timer(stream);
  timer.start();
  error::throw_if_status(loopsb_spmv_bcsr_f64(int32_t(R), int32_t(C), &lay, detail::raw(bcsr.values),
                                              detail::raw(bcsr.block_col_indices), detail::raw(x), detail::raw(y),
                                              static_cast<int32_t>(bcsr.rows), stream),
                         "loopsb_spmv_bcsr_f64");
  cudaStreamSynchronize(stream);
  timer.stop();
  return timer;
}

/// The schedule the library would pick for this matrix (the
/// reference publishes its heuristic's outcomes in plots/data/heuristics.csv).
/// The widest row is computed on the device.
inline schedule::algorithms_t select_schedule(csr_t<int, int, float>& csr) {
  int widest = -1, pick = LOOPSB_SCHED_MERGE_PATH_FLAT;
  if (csr.rows > 0)
    error::throw_if_status(loopsb_csr_max_degree(int32_t(csr.rows), detail::raw(csr.offsets), &widest, nullptr),
                           "loopsb_csr_max_degree");
  error::throw_if_status(loopsb_select_schedule(int32_t(csr.rows), int32_t(csr.cols), int64_t(csr.nnzs), widest, &pick),
                         "loopsb_select_schedule");
  return pick == LOOPSB_SCHED_THREAD_MAPPED ? schedule::algorithms_t::thread_mapped
                                            : schedule::algorithms_t::merge_path_flat;
}

/// SpMV through the schedule select_schedule() picks.
inline void automatic(csr_t<int, int, float>& csr, vector_t<float>& x, vector_t<float>& y, cudaStream_t stream = 0) {
  if (select_schedule(csr) == schedule::algorithms_t::thread_mapped) thread_mapped(csr, x, y, stream);
  else merge_path_flat(csr, x, y, stream);
}

}  // namespace spmv
}  // namespace algorithms
}  // namespace loops
