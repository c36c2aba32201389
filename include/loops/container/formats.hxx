/**
 * @file formats.hxx
 * @brief Sparse containers. Same struct names, members, dtypes and padding
 * rules as the reference (include/loops/container/{csr,coo,ell,bcsr}.hxx);
 * they are the INPUT MEMORY FORMATS of the SpMV path. Conversions between two
 * DEVICE containers of (int, int, float) run on the device through the C ABI
 * (loopsb_csr_to_* / loopsb_coo_to_csr, csrc/convert.cu -- the reference does
 * these with host loops); every other combination runs on the
 * host with std:: algorithms, with the same results.
 */
#pragma once

#include <algorithm>
#include <cstddef>
#include <numeric>
#include <vector>

#include <type_traits>

#include <cuda_runtime.h>

#include <loops/container/vector.hxx>
#include <loops/container/layout.hxx>
#include <loops/error.hxx>
#include <loops/memory.hxx>
#include <loopsb.h>

namespace loops {
using namespace memory;

namespace detail {
/// Both sides on the device with the C ABI's types: convert on the device.
template <typename index_t, typename offset_t, typename value_t, memory_space_t a, memory_space_t b>
inline constexpr bool device_conversion_v =
    a == memory_space_t::device && b == memory_space_t::device && std::is_same_v<index_t, int> &&
    std::is_same_v<offset_t, int> && std::is_same_v<value_t, float>;
template <typename vec_t>
auto* raw(vec_t& v) { return thrust::raw_pointer_cast(v.data()); }
template <typename vec_t>
const auto* raw(const vec_t& v) { return thrust::raw_pointer_cast(v.data()); }
}  // namespace detail

namespace detail {
/**
 * Plans kept on a container between SpMV calls. The reference runs its per-matrix
 * preprocess inside every algorithms::spmv::* call (merge_path_flat.cuh:111,
 * schedule/merge_path_flat.hxx:92-172: allocate + search + free, each call); here the first
 * call on a container creates the plan and later calls re-use it, so the steady state has
 * no allocation and no search. Entries are keyed by the ADDRESSES and sizes of the
 * container's arrays: re-assigning an array (csr.values = other) is seen and rebuilds the
 * plan; a change of the arrays' CONTENTS in place cannot be seen from the host -- call
 * values_changed() (values / column ids) or structure_changed() (offsets, row ids) on the
 * container after such a write. A plan never survives a copy of its container.
 * Not thread-safe per container (like the thrust vectors beside it).
 */
class plan_cache_t {
 public:
  struct entry {
    loopsb_plan_t* plan = nullptr;
    int schedule = -1;
    const void* key[3] = {nullptr, nullptr, nullptr};
    std::size_t dims[4] = {0, 0, 0, 0};
    long long calls = 0;         ///< SpMV calls since the plan was made / last invalidated
    long long tile_after = -2;   ///< -2 not asked yet, -1 never, else calls after which a band-tiled copy pays
    bool tiled = false;          ///< the plan holds a band-tiled copy of the matrix (loopsb_plan_tile_csr)
    bool declined = false;       ///< the library's cost model declined the copy for these arrays
    bool declined_forced = false;  ///< ... even when forced (the format cannot hold this matrix)
  };

  /// merge_path_flat on CSR: -1 = follow LOOPSB_TILED (default: after the break-even number of
  /// calls), 0 = never take a band-tiled copy, 1 = on the next call, 2 = after break-even.
  int tiling = -1;

  plan_cache_t() = default;
  plan_cache_t(const plan_cache_t& rhs) : tiling(rhs.tiling) {}
  plan_cache_t& operator=(const plan_cache_t& rhs) {
    if (this != &rhs) {
      clear();
      tiling = rhs.tiling;
    }
    return *this;
  }
  ~plan_cache_t() { clear(); }

  /// The plan of `schedule` over these arrays, created on the first call (and again after the arrays moved).
  entry& get(int schedule, const loopsb_layout_t& lay, const void* k0, const void* k1, const void* k2,
             std::size_t d0, std::size_t d1, std::size_t d2, std::size_t d3, cudaStream_t stream) {
    entry* slot = nullptr;
    for (auto& e : entries_)
      if (e.schedule == schedule) slot = &e;
    if (slot && slot->plan && slot->key[0] == k0 && slot->key[1] == k1 && slot->key[2] == k2 &&
        slot->dims[0] == d0 && slot->dims[1] == d1 && slot->dims[2] == d2 && slot->dims[3] == d3)
      return *slot;
    if (!slot) {
      entries_.emplace_back();
      slot = &entries_.back();
    }
    if (slot->plan) loopsb_plan_destroy(slot->plan);
    *slot = entry();
    slot->schedule = schedule;
    error::throw_if_status(loopsb_plan_create(&slot->plan, &lay, schedule, stream), "loopsb_plan_create");
    slot->key[0] = k0; slot->key[1] = k1; slot->key[2] = k2;
    slot->dims[0] = d0; slot->dims[1] = d1; slot->dims[2] = d2; slot->dims[3] = d3;
    return *slot;
  }

  /// Values / column ids were written in place: drop every plan-owned copy of them
  /// (loopsb_plan_invalidate); the plans themselves (built from the offsets) stay.
  void values_changed() {
    for (auto& e : entries_) {
      if (e.plan) loopsb_plan_invalidate(e.plan);
      e.tiled = false;
      e.calls = 0;
    }
  }
  /// Offsets / row ids were written in place: every plan is void.
  void clear() {
    for (auto& e : entries_)
      if (e.plan) loopsb_plan_destroy(e.plan);
    entries_.clear();
  }
  std::size_t size() const { return entries_.size(); }
  const entry* find(int schedule) const {
    for (auto& e : entries_)
      if (e.schedule == schedule) return &e;
    return nullptr;
  }

 private:
  std::vector<entry> entries_;
};
}  // namespace detail

template <typename index_t, typename value_t, memory_space_t space = memory_space_t::device>
struct coo_t;
template <typename index_t, typename offset_t, typename value_t, memory_space_t space = memory_space_t::device>
struct csr_t;

/// COO: row_indices / col_indices / values, each nnzs long.
template <typename index_t, typename value_t, memory_space_t space>
struct coo_t {
  std::size_t rows = 0, cols = 0, nnzs = 0;
  vector_t<index_t, space> row_indices;
  vector_t<index_t, space> col_indices;
  vector_t<value_t, space> values;

  coo_t() = default;
  coo_t(std::size_t r, std::size_t c, std::size_t nnz)
      : rows(r), cols(c), nnzs(nnz), row_indices(nnz), col_indices(nnz), values(nnz) {}

  template <auto rhs_space>
  coo_t(const coo_t<index_t, value_t, rhs_space>& rhs)
      : rows(rhs.rows), cols(rhs.cols), nnzs(rhs.nnzs),
        row_indices(rhs.row_indices), col_indices(rhs.col_indices), values(rhs.values) {}

  /// Expand CSR row offsets into one row id per nonzero.
  template <auto rhs_space, typename offset_t>
  coo_t(const csr_t<index_t, offset_t, value_t, rhs_space>& csr);

  /// Row-major (row, col) ordering; ties keep their input order.
  void sort_by_row() { reorder(true); }

  /// Column-major (col, row) ordering (reference container/coo.hxx:116-122); the values
  /// travel with their coordinates.
  void sort_by_column() { reorder(false); }

  /// Sort by (row, col) and keep the first entry of every coordinate pair
  /// (reference container/coo.hxx:128-145).
  void remove_duplicates() {
    sort_by_row();
    thrust::host_vector<index_t> r(row_indices), c(col_indices);
    thrust::host_vector<value_t> v(values);
    std::size_t kept = 0;
    for (std::size_t i = 0; i < nnzs; ++i) {
      if (kept > 0 && r[i] == r[kept - 1] && c[i] == c[kept - 1]) continue;
      r[kept] = r[i]; c[kept] = c[i]; v[kept] = v[i];
      ++kept;
    }
    r.resize(kept); c.resize(kept); v.resize(kept);
    nnzs = kept;
    row_indices = r; col_indices = c; values = v;
    plans_.clear();
  }

  /// In-place writes to the arrays (see detail::plan_cache_t): the cached SpMV plans are rebuilt.
  void values_changed() { plans_.values_changed(); }
  void structure_changed() { plans_.clear(); }
  detail::plan_cache_t& plans() const { return plans_; }

 private:
  mutable detail::plan_cache_t plans_;
  void reorder(bool by_row) {
    thrust::host_vector<index_t> r(row_indices), c(col_indices);
    thrust::host_vector<value_t> v(values);
    std::vector<std::size_t> perm(nnzs);
    std::iota(perm.begin(), perm.end(), std::size_t(0));
    std::stable_sort(perm.begin(), perm.end(), [&](std::size_t a, std::size_t b) {
      const index_t ka = by_row ? r[a] : c[a], kb = by_row ? r[b] : c[b];
      const index_t ta = by_row ? c[a] : r[a], tb = by_row ? c[b] : r[b];
      return ka != kb ? ka < kb : ta < tb;
    });
    thrust::host_vector<index_t> r2(nnzs), c2(nnzs);
    thrust::host_vector<value_t> v2(nnzs);
    for (std::size_t i = 0; i < nnzs; ++i) {
      r2[i] = r[perm[i]]; c2[i] = c[perm[i]]; v2[i] = v[perm[i]];
    }
    row_indices = r2; col_indices = c2; values = v2;
    plans_.clear();
  }
};

/// CSR: offsets[rows+1], indices[nnzs], values[nnzs].
template <typename index_t, typename offset_t, typename value_t, memory_space_t space>
struct csr_t {
  std::size_t rows = 0, cols = 0, nnzs = 0;
  vector_t<offset_t, space> offsets;
  vector_t<index_t, space> indices;
  vector_t<value_t, space> values;

  csr_t() = default;
  csr_t(std::size_t r, std::size_t c, std::size_t nnz)
      : rows(r), cols(c), nnzs(nnz), offsets(r + 1), indices(nnz), values(nnz) {}

  template <auto rhs_space>
  csr_t(const csr_t<index_t, offset_t, value_t, rhs_space>& rhs)
      : rows(rhs.rows), cols(rhs.cols), nnzs(rhs.nnzs),
        offsets(rhs.offsets), indices(rhs.indices), values(rhs.values) {}

  /// From COO: sort by (row, col), then count rows into offsets.
  template <auto rhs_space>
  csr_t(const coo_t<index_t, value_t, rhs_space>& coo)
      : rows(coo.rows), cols(coo.cols), nnzs(coo.nnzs) {
    if constexpr (detail::device_conversion_v<index_t, offset_t, value_t, space, rhs_space>) {
      offsets.resize(rows + 1); indices.resize(nnzs); values.resize(nnzs);
      error::throw_if_status(
          loopsb_coo_to_csr(int(rows), int64_t(nnzs), detail::raw(coo.row_indices), detail::raw(coo.col_indices),
                            detail::raw(coo.values), detail::raw(offsets), detail::raw(indices), detail::raw(values),
                            nullptr),
          "loopsb_coo_to_csr");
      return;
    }
    coo_t<index_t, value_t, memory_space_t::host> sorted(coo);
    sorted.sort_by_row();
    thrust::host_vector<offset_t> off(rows + 1, offset_t(0));
    for (std::size_t i = 0; i < nnzs; ++i)
      off[sorted.row_indices[i] + 1] += 1;
    for (std::size_t r = 0; r < rows; ++r)
      off[r + 1] += off[r];
    offsets = off;
    indices = sorted.col_indices;
    values = sorted.values;
  }

  layout::csr<index_t, offset_t> layout() const {
    return layout::csr<index_t, offset_t>(
        thrust::raw_pointer_cast(offsets.data()), static_cast<index_t>(rows),
        static_cast<offset_t>(nnzs));
  }

  /// Call after writing `values` / `indices` IN PLACE (a kernel, thrust::transform, values[i] = ...):
  /// the SpMV plans cached on this container drop their re-ordered copies of the matrix
  /// (detail::plan_cache_t above). Re-assigning a whole array needs no call.
  void values_changed() { plans_.values_changed(); }
  /// Call after writing `offsets` in place: every cached plan is rebuilt on its next use.
  void structure_changed() { plans_.clear(); }
  detail::plan_cache_t& plans() const { return plans_; }

 private:
  mutable detail::plan_cache_t plans_;
};

template <typename index_t, typename value_t, memory_space_t space>
template <auto rhs_space, typename offset_t>
coo_t<index_t, value_t, space>::coo_t(const csr_t<index_t, offset_t, value_t, rhs_space>& csr)
    : rows(csr.rows), cols(csr.cols), nnzs(csr.nnzs), col_indices(csr.indices), values(csr.values) {
  if constexpr (detail::device_conversion_v<index_t, offset_t, value_t, space, rhs_space>) {
    row_indices.resize(nnzs);
    error::throw_if_status(
        loopsb_csr_to_coo(int(rows), int64_t(nnzs), detail::raw(csr.offsets), detail::raw(row_indices), nullptr),
        "loopsb_csr_to_coo");
    return;
  }
  thrust::host_vector<offset_t> off(csr.offsets);
  thrust::host_vector<index_t> r(nnzs);
  for (std::size_t row = 0; row < rows; ++row)
    for (offset_t k = off[row]; k < off[row + 1]; ++k)
      r[k] = static_cast<index_t>(row);
  row_indices = r;
}

/// ELL: row-major rows*pitch slabs; padding = column sentinel() (-1), value 0.
template <typename index_t, typename value_t, memory_space_t space = memory_space_t::device>
struct ell_t {
  std::size_t rows = 0, cols = 0, nnzs = 0, pitch = 0;
  vector_t<index_t, space> indices;
  vector_t<value_t, space> values;

  static __host__ __device__ index_t sentinel() { return static_cast<index_t>(-1); }

  ell_t() = default;
  ell_t(std::size_t r, std::size_t c, std::size_t nnz, std::size_t p)
      : rows(r), cols(c), nnzs(nnz), pitch(p), indices(r * p, sentinel()), values(r * p, value_t(0)) {}

  template <auto rhs_space>
  ell_t(const ell_t<index_t, value_t, rhs_space>& rhs)
      : rows(rhs.rows), cols(rhs.cols), nnzs(rhs.nnzs), pitch(rhs.pitch),
        indices(rhs.indices), values(rhs.values) {}

  template <typename offset_t, auto rhs_space>
  static std::size_t max_nnz_per_row(const csr_t<index_t, offset_t, value_t, rhs_space>& csr) {
    if constexpr (detail::device_conversion_v<index_t, offset_t, value_t, rhs_space, rhs_space>) {
      int widest = 0;
      error::throw_if_status(loopsb_csr_max_degree(int(csr.rows), detail::raw(csr.offsets), &widest, nullptr),
                             "loopsb_csr_max_degree");
      return std::size_t(widest);
    }
    thrust::host_vector<offset_t> off(csr.offsets);
    std::size_t widest = 0;
    for (std::size_t r = 0; r < csr.rows; ++r)
      widest = std::max(widest, static_cast<std::size_t>(off[r + 1] - off[r]));
    return widest;
  }

  /// pitch = widest row; row r's k-th stored entry lands in slot r*pitch + k.
  template <typename offset_t, auto rhs_space>
  ell_t(const csr_t<index_t, offset_t, value_t, rhs_space>& csr)
      : rows(csr.rows), cols(csr.cols), nnzs(csr.nnzs), pitch(max_nnz_per_row(csr)) {
    if constexpr (detail::device_conversion_v<index_t, offset_t, value_t, space, rhs_space>) {
      indices.resize(rows * pitch); values.resize(rows * pitch);
      error::throw_if_status(
          loopsb_csr_to_ell(int(rows), int(pitch), detail::raw(csr.offsets), detail::raw(csr.indices),
                            detail::raw(csr.values), detail::raw(indices), detail::raw(values), nullptr),
          "loopsb_csr_to_ell");
      error::throw_if_exception(cudaStreamSynchronize(nullptr) != cudaSuccess, "csr -> ell failed on the device");
      return;
    }
    thrust::host_vector<offset_t> off(csr.offsets);
    thrust::host_vector<index_t> idx(csr.indices);
    thrust::host_vector<value_t> val(csr.values);
    thrust::host_vector<index_t> e_idx(rows * pitch, sentinel());
    thrust::host_vector<value_t> e_val(rows * pitch, value_t(0));
    for (std::size_t r = 0; r < rows; ++r) {
      std::size_t slot = r * pitch;
      for (offset_t k = off[r]; k < off[r + 1]; ++k, ++slot) {
        e_idx[slot] = idx[k];
        e_val[slot] = val[k];
      }
    }
    indices = e_idx;
    values = e_val;
  }

  layout::ell<index_t, index_t> layout() const {
    return layout::ell<index_t, index_t>(static_cast<index_t>(rows), static_cast<index_t>(pitch));
  }

  /// In-place writes to the arrays (see detail::plan_cache_t). ELL plans depend on rows and pitch only.
  void values_changed() { plans_.values_changed(); }
  void structure_changed() { plans_.clear(); }
  detail::plan_cache_t& plans() const { return plans_; }

 private:
  mutable detail::plan_cache_t plans_;
};

/// BCSR: R x C dense blocks, values[b*R*C + i*C + j]; block columns ascending
/// inside a block-row; CSR entries are assigned into their block, padding 0.
template <std::size_t R, std::size_t C, typename index_t, typename offset_t, typename value_t,
          memory_space_t space = memory_space_t::device>
struct bcsr_t {
  static_assert(R > 0 && C > 0, "BCSR block dims must be positive.");
  static constexpr std::size_t kBlockRows = R, kBlockCols = C, kBlockSize = R * C;

  std::size_t rows = 0, cols = 0, nnzs = 0;
  std::size_t num_block_rows = 0, num_block_cols = 0, num_blocks = 0;
  vector_t<offset_t, space> block_offsets;
  vector_t<index_t, space> block_col_indices;
  vector_t<value_t, space> values;

  bcsr_t() = default;

  template <auto rhs_space>
  bcsr_t(const bcsr_t<R, C, index_t, offset_t, value_t, rhs_space>& rhs)
      : rows(rhs.rows), cols(rhs.cols), nnzs(rhs.nnzs), num_block_rows(rhs.num_block_rows),
        num_block_cols(rhs.num_block_cols), num_blocks(rhs.num_blocks),
        block_offsets(rhs.block_offsets), block_col_indices(rhs.block_col_indices), values(rhs.values) {}

  template <auto rhs_space, typename csr_offset_t>
  bcsr_t(const csr_t<index_t, csr_offset_t, value_t, rhs_space>& csr)
      : rows(csr.rows), cols(csr.cols), nnzs(csr.nnzs),
        num_block_rows((csr.rows + R - 1) / R), num_block_cols((csr.cols + C - 1) / C) {
    if constexpr (detail::device_conversion_v<index_t, csr_offset_t, value_t, space, rhs_space> &&
                  std::is_same_v<offset_t, int>) {
      block_offsets.resize(num_block_rows + 1);
      vector_t<int, memory_space_t::device> atom_block(nnzs);
      int64_t nb = 0;
      error::throw_if_status(
          loopsb_csr_to_bcsr_count(int(R), int(C), int(rows), int(cols), int64_t(nnzs), detail::raw(csr.offsets),
                                   detail::raw(csr.indices), detail::raw(block_offsets), detail::raw(atom_block), &nb,
                                   nullptr),
          "loopsb_csr_to_bcsr_count");
      num_blocks = std::size_t(nb);
      block_col_indices.resize(num_blocks); values.resize(num_blocks * kBlockSize);
      error::throw_if_status(
          loopsb_csr_to_bcsr_fill(int(R), int(C), int(rows), int(cols), int64_t(nnzs), detail::raw(csr.offsets),
                                  detail::raw(csr.indices), detail::raw(csr.values), detail::raw(atom_block), nb,
                                  detail::raw(block_col_indices), detail::raw(values), 0, nullptr),
          "loopsb_csr_to_bcsr_fill");
      return;
    }
    thrust::host_vector<csr_offset_t> off(csr.offsets);
    thrust::host_vector<index_t> idx(csr.indices);
    thrust::host_vector<value_t> val(csr.values);
    std::vector<offset_t> b_off(num_block_rows + 1, offset_t(0));
    std::vector<index_t> b_col;
    std::vector<value_t> b_val;
    std::vector<index_t> seen;  // block columns touched by the current block-row
    for (std::size_t br = 0; br < num_block_rows; ++br) {
      const std::size_t lo = br * R, hi = std::min(lo + R, rows);
      seen.clear();
      for (csr_offset_t a = off[lo]; a < off[hi]; ++a)
        seen.push_back(static_cast<index_t>(idx[a] / C));
      std::sort(seen.begin(), seen.end());
      seen.erase(std::unique(seen.begin(), seen.end()), seen.end());
      const std::size_t first_block = b_col.size();
      b_col.insert(b_col.end(), seen.begin(), seen.end());
      b_val.resize(b_val.size() + seen.size() * kBlockSize, value_t(0));
      for (std::size_t r = lo; r < hi; ++r)
        for (csr_offset_t a = off[r]; a < off[r + 1]; ++a) {
          const index_t bc = static_cast<index_t>(idx[a] / C);
          const std::size_t local = std::lower_bound(seen.begin(), seen.end(), bc) - seen.begin();
          b_val[(first_block + local) * kBlockSize + (r - lo) * C + (idx[a] % C)] = val[a];
        }
      b_off[br + 1] = static_cast<offset_t>(b_col.size());
    }
    num_blocks = b_col.size();
    block_offsets = thrust::host_vector<offset_t>(b_off.begin(), b_off.end());
    block_col_indices = thrust::host_vector<index_t>(b_col.begin(), b_col.end());
    values = thrust::host_vector<value_t>(b_val.begin(), b_val.end());
  }

  layout::bcsr<index_t, offset_t> layout() const {
    return layout::bcsr<index_t, offset_t>(
        thrust::raw_pointer_cast(block_offsets.data()), static_cast<index_t>(num_block_rows),
        static_cast<offset_t>(num_blocks));
  }
};

/// CSC: offsets[cols+1], indices[nnzs] = ROW ids, values[nnzs]; entries ordered
/// by (column, row) (reference container/csc.hxx:88-102).
template <typename index_t, typename offset_t, typename value_t, memory_space_t space = memory_space_t::device>
struct csc_t {
  std::size_t rows = 0, cols = 0, nnzs = 0;
  vector_t<offset_t, space> offsets;
  vector_t<index_t, space> indices;
  vector_t<value_t, space> values;

  csc_t() = default;
  csc_t(std::size_t r, std::size_t c, std::size_t nnz)
      : rows(r), cols(c), nnzs(nnz), offsets(c + 1), indices(nnz), values(nnz) {}

  template <auto rhs_space>
  csc_t(const csc_t<index_t, offset_t, value_t, rhs_space>& rhs)
      : rows(rhs.rows), cols(rhs.cols), nnzs(rhs.nnzs),
        offsets(rhs.offsets), indices(rhs.indices), values(rhs.values) {}

  /// From COO (reference container/csc.hxx:85-93): entries ordered by (column, row),
  /// column offsets by counting.
  template <auto rhs_space>
  csc_t(const coo_t<index_t, value_t, rhs_space>& coo) : rows(coo.rows), cols(coo.cols), nnzs(coo.nnzs) {
    coo_t<index_t, value_t, memory_space_t::host> sorted(coo);
    sorted.sort_by_column();
    thrust::host_vector<offset_t> c_off(cols + 1, offset_t(0));
    for (std::size_t a = 0; a < nnzs; ++a) c_off[sorted.col_indices[a] + 1] += 1;
    for (std::size_t c = 0; c < cols; ++c) c_off[c + 1] += c_off[c];
    offsets = c_off;
    indices = sorted.row_indices;
    values = sorted.values;
  }

  /// From CSR: stable counting sort by column over CSR order.
  template <auto rhs_space>
  csc_t(const csr_t<index_t, offset_t, value_t, rhs_space>& csr)
      : rows(csr.rows), cols(csr.cols), nnzs(csr.nnzs) {
    if constexpr (detail::device_conversion_v<index_t, offset_t, value_t, space, rhs_space>) {
      offsets.resize(cols + 1); indices.resize(nnzs); values.resize(nnzs);
      error::throw_if_status(
          loopsb_csr_to_csc(int(rows), int(cols), int64_t(nnzs), detail::raw(csr.offsets), detail::raw(csr.indices),
                            detail::raw(csr.values), detail::raw(offsets), detail::raw(indices), detail::raw(values),
                            nullptr),
          "loopsb_csr_to_csc");
      return;
    }
    thrust::host_vector<offset_t> off(csr.offsets);
    thrust::host_vector<index_t> idx(csr.indices);
    thrust::host_vector<value_t> val(csr.values);
    thrust::host_vector<offset_t> c_off(cols + 1, offset_t(0));
    for (std::size_t a = 0; a < nnzs; ++a) c_off[idx[a] + 1] += 1;
    for (std::size_t c = 0; c < cols; ++c) c_off[c + 1] += c_off[c];
    std::vector<offset_t> cur(c_off.begin(), c_off.end() - 1);
    thrust::host_vector<index_t> c_row(nnzs);
    thrust::host_vector<value_t> c_val(nnzs);
    for (std::size_t r = 0; r < rows; ++r)
      for (offset_t a = off[r]; a < off[r + 1]; ++a) {
        const offset_t p = cur[idx[a]]++;
        c_row[p] = static_cast<index_t>(r);
        c_val[p] = val[a];
      }
    offsets = c_off;
    indices = c_row;
    values = c_val;
  }

  layout::csc<index_t, offset_t> layout() const {
    return layout::csc<index_t, offset_t>(thrust::raw_pointer_cast(offsets.data()),
                                          static_cast<index_t>(cols), static_cast<offset_t>(nnzs));
  }
};

/// DIA: distinct (col - row) offsets ascending, values column-major
/// values[d * stride + r], stride = rows (reference container/dia.hxx:56-62,135-188).
template <typename index_t, typename offset_t, typename value_t, memory_space_t space = memory_space_t::device>
struct dia_t {
  std::size_t rows = 0, cols = 0, nnzs = 0, stride = 0, num_diagonals = 0;
  vector_t<index_t, space> diag_offsets;
  vector_t<value_t, space> values;

  dia_t() = default;

  template <auto rhs_space>
  dia_t(const dia_t<index_t, offset_t, value_t, rhs_space>& rhs)
      : rows(rhs.rows), cols(rhs.cols), nnzs(rhs.nnzs), stride(rhs.stride), num_diagonals(rhs.num_diagonals),
        diag_offsets(rhs.diag_offsets), values(rhs.values) {}

  template <auto rhs_space, typename csr_offset_t>
  static std::size_t count_diagonals(const csr_t<index_t, csr_offset_t, value_t, rhs_space>& csr) {
    return sorted_offsets(csr).size();
  }

  template <auto rhs_space, typename csr_offset_t>
  dia_t(const csr_t<index_t, csr_offset_t, value_t, rhs_space>& csr)
      : rows(csr.rows), cols(csr.cols), nnzs(csr.nnzs), stride(csr.rows) {
    if constexpr (detail::device_conversion_v<index_t, csr_offset_t, value_t, space, rhs_space>) {
      int nd = 0;
      error::throw_if_status(
          loopsb_csr_to_dia_count(int(rows), int(cols), int64_t(nnzs), detail::raw(csr.offsets),
                                  detail::raw(csr.indices), &nd, nullptr),
          "loopsb_csr_to_dia_count");
      num_diagonals = std::size_t(nd);
      diag_offsets.resize(num_diagonals); values.resize(num_diagonals * stride);
      error::throw_if_status(
          loopsb_csr_to_dia_fill(int(rows), int(cols), int64_t(nnzs), detail::raw(csr.offsets),
                                 detail::raw(csr.indices), detail::raw(csr.values), nd, detail::raw(diag_offsets),
                                 detail::raw(values), nullptr),
          "loopsb_csr_to_dia_fill");
      return;
    }
    const std::vector<index_t> offs = sorted_offsets(csr);
    num_diagonals = offs.size();
    thrust::host_vector<csr_offset_t> off(csr.offsets);
    thrust::host_vector<index_t> idx(csr.indices);
    thrust::host_vector<value_t> val(csr.values);
    thrust::host_vector<value_t> d_val(num_diagonals * stride, value_t(0));
    for (std::size_t r = 0; r < rows; ++r)
      for (csr_offset_t a = off[r]; a < off[r + 1]; ++a) {
        const index_t o = idx[a] - static_cast<index_t>(r);
        const std::size_t d = std::lower_bound(offs.begin(), offs.end(), o) - offs.begin();
        d_val[d * stride + r] = val[a];   // assigned, not added (as the reference)
      }
    diag_offsets = thrust::host_vector<index_t>(offs.begin(), offs.end());
    values = d_val;
  }

  layout::dia<std::size_t, std::size_t> layout() const {
    return layout::dia<std::size_t, std::size_t>(rows, num_diagonals);
  }

 private:
  template <auto rhs_space, typename csr_offset_t>
  static std::vector<index_t> sorted_offsets(const csr_t<index_t, csr_offset_t, value_t, rhs_space>& csr) {
    thrust::host_vector<csr_offset_t> off(csr.offsets);
    thrust::host_vector<index_t> idx(csr.indices);
    std::vector<index_t> all;
    all.reserve(csr.nnzs);
    for (std::size_t r = 0; r < csr.rows; ++r)
      for (csr_offset_t a = off[r]; a < off[r + 1]; ++a) all.push_back(idx[a] - static_cast<index_t>(r));
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    return all;
  }
};

}  // namespace loops
