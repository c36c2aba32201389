/**
 * @file reference.hxx
 * @brief Host-side validators for SpMV results, with the reference's public names
 * (reference include/loops/util/reference.hxx:61-76,116-131,150-198,203-243,278-337,
 * 358-388): `reference::spmv` (row sums in value_t, the order a sequential CPU loop
 * adds them), `spmv_f64` (double accumulation), `row_l1_products`, `default_tolerance`,
 * `unit_roundoff`, `rigorous_report` / `rigorously_validate_spmv` (per-row Wilkinson
 * bound max(atol_floor, K * nnz_r * u * L1_r) against the double-precision sums) and
 * `count_errors`. They accept containers in either memory space; everything is pulled
 * to the host first. Not part of the GPU path -- these are what `--validate` prints.
 */
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>
#include <thrust/host_vector.h>

#include <loops/container/formats.hxx>
#include <loops/container/vector.hxx>
#include <loops/memory.hxx>

namespace loops {
namespace reference {

namespace detail {
/// One pass over the rows of a host CSR: `visit(row, begin, end)`.
template <typename csr_host_t, typename fn_t>
void for_each_row(const csr_host_t& a, fn_t&& visit) {
  for (std::size_t r = 0; r < a.rows; ++r)
    visit(r, static_cast<std::size_t>(a.offsets[r]), static_cast<std::size_t>(a.offsets[r + 1]));
}

/// Row-wise reduction of the products values[k] * x[indices[k]] in accumulator type acc_t,
/// each product passed through `shape` first (identity, or |.| for the L1 mass).
template <typename acc_t, typename index_t, typename offset_t, typename value_t, memory_space_t space,
          typename shape_t>
vector_t<value_t, memory_space_t::host> reduce_rows(const csr_t<index_t, offset_t, value_t, space>& csr,
                                                    const vector_t<value_t, space>& x, shape_t shape) {
  csr_t<index_t, offset_t, value_t, memory_space_t::host> a(csr);
  vector_t<value_t, memory_space_t::host> xh(x);
  vector_t<value_t, memory_space_t::host> out(a.rows, value_t{0});
  for_each_row(a, [&](std::size_t r, std::size_t b, std::size_t e) {
    acc_t acc = acc_t{0};
    for (std::size_t k = b; k < e; ++k)
      acc += shape(static_cast<acc_t>(a.values[k]) * static_cast<acc_t>(xh[a.indices[k]]));
    out[r] = static_cast<value_t>(acc);
  });
  return out;
}
}  // namespace detail

/// y = A x with value_t accumulation, rows summed left to right.
template <typename index_t, typename offset_t, typename value_t, memory_space_t space>
vector_t<value_t, memory_space_t::host> spmv(const csr_t<index_t, offset_t, value_t, space>& csr,
                                             const vector_t<value_t, space>& x) {
  return detail::reduce_rows<value_t>(csr, x, [](value_t p) { return p; });
}

/// y = A x accumulated in double, rounded once to value_t.
template <typename index_t, typename offset_t, typename value_t, memory_space_t space>
vector_t<value_t, memory_space_t::host> spmv_f64(const csr_t<index_t, offset_t, value_t, space>& csr,
                                                 const vector_t<value_t, space>& x) {
  return detail::reduce_rows<double>(csr, x, [](double p) { return p; });
}

/// sum_k |values[k] * x[indices[k]]| per row (what fp32 summation error scales with).
template <typename index_t, typename offset_t, typename value_t, memory_space_t space>
vector_t<value_t, memory_space_t::host> row_l1_products(const csr_t<index_t, offset_t, value_t, space>& csr,
                                                        const vector_t<value_t, space>& x) {
  return detail::reduce_rows<double>(csr, x, [](double p) { return std::abs(p); });
}

/// The coarse mismatch predicate of the examples' --validate: |a - b| > 1e-2 + 1e-3 |b|.
template <typename value_t>
struct default_tolerance {
  static constexpr value_t atol() { return value_t(1e-2); }
  static constexpr value_t rtol() { return value_t(1e-3); }
  static __host__ __device__ bool ne(value_t a, value_t b) {
    const value_t d = a > b ? a - b : b - a;
    const value_t m = b < value_t(0) ? -b : b;
    return d > atol() + rtol() * m;
  }
};

/// eps / 2 for round-to-nearest.
template <typename value_t>
constexpr value_t unit_roundoff();
template <>
constexpr float unit_roundoff<float>() { return 5.9604644775390625e-08f; }            // 2^-24
template <>
constexpr double unit_roundoff<double>() { return 1.1102230246251565404e-16; }        // 2^-53

struct rigorous_report {
  std::size_t total_rows = 0;
  std::size_t naive_mismatches = 0;        ///< rows default_tolerance flags against the value_t sums
  std::size_t f32_baseline_overruns = 0;   ///< rows where the value_t CPU sum itself leaves the bound
  std::size_t gpu_overruns = 0;            ///< rows where the candidate leaves the bound
  double max_gpu_abs_error = 0.0;          ///< max |y - y_f64|
  double max_gpu_rel_error = 0.0;          ///< max |y - y_f64| / max(|y_f64|, 1)
  double wilkinson_k = 0.0;
};

namespace detail {
template <typename value_t>
thrust::host_vector<value_t> fetch(const value_t* p, std::size_t n) {
  thrust::host_vector<value_t> h(n);
  if (n) {
    cudaPointerAttributes at{};
    const bool on_device = cudaPointerGetAttributes(&at, p) == cudaSuccess &&
                           (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
    (void)cudaGetLastError();
    if (on_device) cudaMemcpy(h.data(), p, n * sizeof(value_t), cudaMemcpyDeviceToHost);
    else std::copy(p, p + n, h.begin());
  }
  return h;
}
}  // namespace detail

/**
 * @brief Classify every row of a candidate y as round-off or bug: compare with the
 * double-precision sums under the row's Wilkinson bound, relaxed by `wilkinson_k`.
 * `d_y_gpu` may be a device or a host pointer.
 */
template <typename index_t, typename offset_t, typename value_t, memory_space_t space>
rigorous_report rigorously_validate_spmv(const csr_t<index_t, offset_t, value_t, space>& csr,
                                         const vector_t<value_t, space>& x, const value_t* d_y_gpu,
                                         double wilkinson_k = 8.0, double atol_floor = 1e-3,
                                         bool verbose = false) {
  csr_t<index_t, offset_t, value_t, memory_space_t::host> a(csr);
  vector_t<value_t, memory_space_t::host> xh(x);
  const auto y_native = spmv(a, xh);
  const auto y_wide = spmv_f64(a, xh);
  const auto mass = row_l1_products(a, xh);
  const auto y = detail::fetch(d_y_gpu, a.rows);
  const double u = static_cast<double>(unit_roundoff<value_t>());

  rigorous_report rep;
  rep.total_rows = a.rows;
  rep.wilkinson_k = wilkinson_k;
  detail::for_each_row(a, [&](std::size_t r, std::size_t b, std::size_t e) {
    const double bound = std::max(atol_floor, wilkinson_k * double(e - b) * u * double(mass[r]));
    const double wide = double(y_wide[r]);
    const double err_native = std::abs(double(y_native[r]) - wide);
    const double err = std::abs(double(y[r]) - wide);
    if (default_tolerance<value_t>::ne(y[r], y_native[r])) ++rep.naive_mismatches;
    if (err_native > bound) ++rep.f32_baseline_overruns;
    if (err > bound) {
      ++rep.gpu_overruns;
      if (verbose)
        std::printf("GPU_OVERRUN row=%zu nnz=%zu L1=%.6g y_gpu=%.8g y_f64=%.8g abs_err=%.6g bound=%.6g\n", r, e - b,
                    double(mass[r]), double(y[r]), wide, err, bound);
    }
    rep.max_gpu_abs_error = std::max(rep.max_gpu_abs_error, err);
    rep.max_gpu_rel_error = std::max(rep.max_gpu_rel_error, err / std::max(std::abs(wide), 1.0));
  });
  return rep;
}

/// Number of positions where `ne(d_y[i], h_ref[i])`; `d_y` may live on the device.
template <typename value_t, typename ne_t>
std::size_t count_errors(const value_t* d_y, const value_t* h_ref, std::size_t n, ne_t ne, bool verbose = false) {
  const auto y = detail::fetch(d_y, n);
  std::size_t bad = 0;
  for (std::size_t i = 0; i < n; ++i) {
    if (!ne(y[i], h_ref[i])) continue;
    if (verbose) std::printf("Error[%zu]: %.8g != %.8g\n", i, double(y[i]), double(h_ref[i]));
    ++bad;
  }
  return bad;
}

template <typename value_t>
std::size_t count_errors(const value_t* d_y, const value_t* h_ref, std::size_t n, bool verbose = false) {
  return count_errors(d_y, h_ref, n, [](value_t a, value_t b) { return default_tolerance<value_t>::ne(a, b); },
                      verbose);
}

}  // namespace reference
}  // namespace loops
