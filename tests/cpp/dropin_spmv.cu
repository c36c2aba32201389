// tests/cpp/dropin_spmv.cu -- source-level drop-in check of include/loops/.
//
// Written the way user code is written against the reference: owning
// containers, algorithms::spmv::<name>(container, x, y), and hand-written
// kernels over schedule::setup<...> -- including one with a USER-DEFINED layout
// (the pattern of reference examples/spmv/custom_layout.cu:64-224) and one that
// is the reference's own merge-path kernel body verbatim in form
// (algorithms/spmv/merge_path_flat.cuh:63-82: init / is_valid_accessor /
// virtual_idx / atom_idx / tile_idx / atoms_counting_it / tile_end_offset).
// Prints one "OK <case>" or "FAIL <case>" line per check; exit code = failures.
#include <loops/schedule.hxx>
#include <loops/container/formats.hxx>
#include <loops/algorithms/spmv/merge_path_flat.cuh>
#include <loops/algorithms/spmv/work_oriented.cuh>
#include <loops/algorithms/spmv/thread_mapped.cuh>
#include <loops/algorithms/spmv/group_mapped.cuh>
#include <loops/algorithms/spmv/coo_thread_mapped.cuh>
#include <loops/algorithms/spmv/ell_thread_mapped.cuh>
#include <loops/algorithms/spmv/ell_merge_path.cuh>
#include <loops/algorithms/spmv/bcsr_thread_mapped.cuh>
#include <loops/algorithms/spmv/csc_thread_mapped.cuh>
#include <loops/algorithms/spmv/dia_thread_mapped.cuh>
#include <loops/algorithms/spmv/flat_partitioned.cuh>
#include <loops/algorithms/spmv/original.cuh>
#include <loops/algorithms/spmm/thread_mapped.cuh>

#include <thrust/fill.h>
#include <thrust/transform.h>

#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

using namespace loops;
using csr_host_t = csr_t<int, int, float, memory_space_t::host>;

// ---- a user layout: every tile owns `width` atoms except that the tiles are
// described by an explicit array the user keeps (here: the CSR offsets, shifted
// view) -- anything with the six contract methods works.
struct banded_layout {
  using tile_id_t = int;
  using atom_id_t = int;
  using tile_end_iterator_t = const int*;
  const int* ends;  // ends[t] = end of tile t; begin of tile 0 is 0
  int n_tiles, n_atoms;
  __host__ __device__ int num_tiles() const { return n_tiles; }
  __host__ __device__ int num_atoms() const { return n_atoms; }
  __host__ __device__ int tile_begin(int t) const { return t == 0 ? 0 : ends[t - 1]; }
  __host__ __device__ int tile_end(int t) const { return ends[t]; }
  __host__ __device__ int tile_size(int t) const { return tile_end(t) - tile_begin(t); }
  __host__ __device__ tile_end_iterator_t tile_end_iter() const { return ends; }
};

template <typename setup_t>
__global__ void user_thread_mapped(setup_t config, const int* indices, const float* values, const float* x,
                                   float* y) {
  for (auto row : config.tiles()) {
    float sum = 0;
    for (auto nz : config.atoms(row))
      sum += values[nz] * x[indices[nz]];
    y[row] = sum;
  }
}

template <std::size_t TPB, std::size_t IPT, typename meta_t>
__global__ void __launch_bounds__(int(TPB))
    user_merge_path(meta_t meta, std::size_t rows, std::size_t nnz, int* offsets, int* indices,
                    const float* values, const float* x, float* y) {
  using setup_t = schedule::setup<schedule::algorithms_t::merge_path_flat, TPB, IPT, int, int, std::size_t,
                                  std::size_t>;
  using storage_t = typename setup_t::storage_t;
  __shared__ storage_t temporary_storage;
  setup_t config(meta, temporary_storage, offsets, rows, nnz);
  auto map = config.init();
  if (!config.is_valid_accessor(map))
    return;
  for (auto item : config.virtual_idx()) {
    auto nz = config.atom_idx(item, map);
    auto row = config.tile_idx(map);
    float nonzero = values[nz] * x[indices[nz]];
    if (config.atoms_counting_it[map.y] < temporary_storage.tile_end_offset[map.x]) {
      atomicAdd(&(y[row]), nonzero);
      map.y++;
    } else {
      map.x++;
    }
  }
}

struct twice_fn { __host__ __device__ float operator()(float v) const { return 2.0f * v; } };
struct halve_fn { __host__ __device__ float operator()(float v) const { return 0.5f * v; } };

static int failures = 0;
static void check(const char* name, const thrust::host_vector<float>& y, const std::vector<float>& ref) {
  double worst = 0;
  for (std::size_t i = 0; i < ref.size(); ++i)
    worst = std::max(worst, std::abs(double(y[i]) - double(ref[i])) / std::max(1.0, std::abs(double(ref[i]))));
  const bool ok = worst <= 1e-5;
  std::printf("%s %s (max rel err %.3g)\n", ok ? "OK" : "FAIL", name, worst);
  failures += ok ? 0 : 1;
}

int main() {
  // a skewed matrix: one heavy row, empty rows, random light rows
  const int rows = 5000, cols = 4096;
  std::mt19937 rng(3);
  std::uniform_real_distribution<float> val(0.5f, 1.5f);
  std::vector<int> off(rows + 1, 0), idx;
  std::vector<float> vals;
  for (int r = 0; r < rows; ++r) {
    int deg = (r % 7 == 0) ? 0 : (r == 11 ? 3000 : int(rng() % 24));
    int step = std::max(1, cols / std::max(deg, 1));
    for (int k = 0; k < deg && k * step < cols; ++k) {
      idx.push_back(k * step + int(rng() % step));
      vals.push_back(val(rng));
    }
    off[r + 1] = int(idx.size());
  }
  const int nnz = int(idx.size());
  csr_host_t h(rows, cols, nnz);
  std::copy(off.begin(), off.end(), h.offsets.begin());
  std::copy(idx.begin(), idx.end(), h.indices.begin());
  std::copy(vals.begin(), vals.end(), h.values.begin());
  std::vector<float> xs(cols);
  for (auto& v : xs) v = val(rng);
  std::vector<float> ref(rows, 0.0f);
  for (int r = 0; r < rows; ++r) {
    float s = 0;
    for (int k = off[r]; k < off[r + 1]; ++k) s += vals[k] * xs[idx[k]];
    ref[r] = s;
  }

  csr_t<int, int, float> csr(h);                 // host -> device, as in the examples
  vector_t<float> x(xs.begin(), xs.end());
  vector_t<float> y(rows);

  try {
    auto t = algorithms::spmv::merge_path_flat(csr, x, y);
    check("algorithms::spmv::merge_path_flat", y, ref);
    {
      // the reference signature keeps its plan on the container: the second call re-uses it
      // (no allocation), a forced tiling switches later calls to the band-tiled kernel, and
      // values_changed() after an in-place write keeps the result in step with the live arrays
      const auto* e1 = csr.plans().find(LOOPSB_SCHED_MERGE_PATH_FLAT);
      const loopsb_plan_t* p1 = e1 ? e1->plan : nullptr;
      thrust::fill(y.begin(), y.end(), -1.0f);
      algorithms::spmv::merge_path_flat(csr, x, y);
      const auto* e2 = csr.plans().find(LOOPSB_SCHED_MERGE_PATH_FLAT);
      bool ok = p1 && e2 && e2->plan == p1 && e2->calls == 2 && !e2->tiled && csr.plans().size() == 1;
      csr.plans().tiling = 1;
      thrust::fill(y.begin(), y.end(), -1.0f);
      algorithms::spmv::merge_path_flat(csr, x, y);
      loopsb_tiled_info_t ti;
      ok = ok && e2->plan == p1 && e2->tiled && loopsb_plan_tiled_info(e2->plan, &ti) == LOOPSB_OK && ti.grid_blocks > 0;
      if (!ok) { std::printf("FAIL cached plan state (reuse / forced tiling)\n"); ++failures; }
      check("algorithms::spmv::merge_path_flat (cached plan, band-tiled kernel)", y, ref);
      // in-place update of the values: 2*A
      thrust::transform(csr.values.begin(), csr.values.end(), csr.values.begin(), twice_fn());
      csr.values_changed();
      std::vector<float> ref2(ref);
      for (auto& v : ref2) v *= 2.0f;
      thrust::fill(y.begin(), y.end(), -1.0f);
      algorithms::spmv::merge_path_flat(csr, x, y);      // re-tiles (tiling = 1) from the live values
      check("algorithms::spmv::merge_path_flat after values_changed()", y, ref2);
      thrust::transform(csr.values.begin(), csr.values.end(), csr.values.begin(), halve_fn());
      csr.values_changed();
      csr.plans().tiling = 0;
      thrust::fill(y.begin(), y.end(), -1.0f);
      algorithms::spmv::merge_path_flat(csr, x, y);
      check("algorithms::spmv::merge_path_flat (tiling off again)", y, ref);
      csr.plans().tiling = -1;
    }
    // the re-usable plan with the band-tiled copy (forced: this matrix is below the cost model's size)
    algorithms::spmv::merge_path_plan_t plan(csr, 0, true, true);
    for (int rep = 0; rep < 3; ++rep) {
      thrust::fill(y.begin(), y.end(), -1.0f);
      plan(x, y);
    }
    if (!plan.band_tiled()) { std::printf("FAIL merge_path_plan_t did not tile\n"); ++failures; }
    check("algorithms::spmv::merge_path_plan_t (band-tiled)", y, ref);
    // the remaining in-tree entry points: original, csc, dia (on a banded matrix), flat_partitioned
    thrust::fill(y.begin(), y.end(), -1.0f);
    algorithms::spmv::original(csr, x, y);
    check("algorithms::spmv::original", y, ref);
    csc_t<int, int, float> csc(csr);
    thrust::fill(y.begin(), y.end(), -1.0f);
    algorithms::spmv::csc_thread_mapped(csc, x, y);
    check("algorithms::spmv::csc_thread_mapped", y, ref);
    thrust::fill(y.begin(), y.end(), -1.0f);
    algorithms::spmv::flat_partitioned<8>(csr, x, y);
    check("algorithms::spmv::flat_partitioned<8>", y, ref);
    {
      const int n = 600;
      csr_host_t hb(n, n, 0);
      std::vector<int> boff(n + 1, 0), bidx;
      std::vector<float> bval;
      for (int r = 0; r < n; ++r) {
        for (int d : {-7, -1, 0, 2, 31})
          if (r + d >= 0 && r + d < n && (r + d) % 5 != 0) { bidx.push_back(r + d); bval.push_back(val(rng)); }
        boff[r + 1] = int(bidx.size());
      }
      csr_host_t hb2(n, n, bidx.size());
      std::copy(boff.begin(), boff.end(), hb2.offsets.begin());
      std::copy(bidx.begin(), bidx.end(), hb2.indices.begin());
      std::copy(bval.begin(), bval.end(), hb2.values.begin());
      std::vector<float> bref(n, 0.0f);
      for (int r = 0; r < n; ++r) {
        float sum = 0;
        for (int k = boff[r]; k < boff[r + 1]; ++k) sum += bval[k] * xs[bidx[k]];
        bref[r] = sum;
      }
      csr_t<int, int, float> band(hb2);
      dia_t<int, int, float> dia(band);
      vector_t<float> yb(n);
      thrust::fill(yb.begin(), yb.end(), -1.0f);
      algorithms::spmv::dia_thread_mapped(dia, x, yb);
      check("algorithms::spmv::dia_thread_mapped", yb, bref);
    }
    {
      // SpMM: C = A B with 5 dense columns; column j of B is x scaled by (j + 1)
      const int n = 5;
      matrix_t<float> B(cols, n), Cm(rows, n);
      thrust::host_vector<float> hB(std::size_t(cols) * n);
      for (int r = 0; r < cols; ++r)
        for (int j = 0; j < n; ++j) hB[std::size_t(r) * n + j] = xs[r] * float(j + 1);
      B.m_data = hB;
      B.m_data_ptr = thrust::raw_pointer_cast(B.m_data.data());
      algorithms::spmm::thread_mapped(csr, B, Cm);
      thrust::host_vector<float> hC(Cm.m_data);
      for (int j = 0; j < n; ++j) {
        thrust::host_vector<float> colj(rows);
        std::vector<float> refj(rows);
        for (int r = 0; r < rows; ++r) { colj[r] = hC[std::size_t(r) * n + j]; refj[r] = ref[r] * float(j + 1); }
        if (j == 0 || j == n - 1) check(j == 0 ? "algorithms::spmm::thread_mapped (col 0)" : "algorithms::spmm::thread_mapped (col 4)", colj, refj);
      }
    }
    std::printf("   timer_t: %.3f ms\n", t.milliseconds());
    algorithms::spmv::work_oriented(csr, x, y);   check("algorithms::spmv::work_oriented", y, ref);
    algorithms::spmv::thread_mapped(csr, x, y);   check("algorithms::spmv::thread_mapped", y, ref);
    algorithms::spmv::group_mapped(csr, x, y);    check("algorithms::spmv::group_mapped", y, ref);
    algorithms::spmv::automatic(csr, x, y);       check("algorithms::spmv::automatic", y, ref);
    {  // the same entry points in double (the reference's .f64 example builds)
      csr_t<int, int, double, memory_space_t::host> hd(rows, cols, nnz);
      std::copy(off.begin(), off.end(), hd.offsets.begin());
      std::copy(idx.begin(), idx.end(), hd.indices.begin());
      std::copy(vals.begin(), vals.end(), hd.values.begin());
      csr_t<int, int, double> cd(hd);
      std::vector<double> xd_h(xs.begin(), xs.end());
      vector_t<double> xd(xd_h.begin(), xd_h.end()), yd(rows);
      std::vector<double> refd(rows, 0.0);
      for (int r = 0; r < rows; ++r) {
        double s = 0;
        for (int k = off[r]; k < off[r + 1]; ++k) s += double(vals[k]) * double(xs[idx[k]]);
        refd[r] = s;
      }
      auto check_d = [&](const char* name) {
        thrust::host_vector<double> got(yd);
        double worst = 0;
        for (int r = 0; r < rows; ++r) worst = std::max(worst, std::fabs(got[r] - refd[r]) / std::max(1.0, std::fabs(refd[r])));
        const bool ok = worst <= 1e-13;
        failures += ok ? 0 : 1;
        std::printf("%s %s (max rel err %.3g)\n", ok ? "OK" : "FAIL", name, worst);
      };
      algorithms::spmv::merge_path_flat(cd, xd, yd); check_d("algorithms::spmv::merge_path_flat<double>");
      algorithms::spmv::thread_mapped(cd, xd, yd);   check_d("algorithms::spmv::thread_mapped<double>");
    }
    coo_t<int, float> coo(csr);
    algorithms::spmv::coo_thread_mapped(coo, x, y); check("algorithms::spmv::coo_thread_mapped", y, ref);
    ell_t<int, float> ell(csr);
    algorithms::spmv::ell_thread_mapped(ell, x, y); check("algorithms::spmv::ell_thread_mapped", y, ref);
    algorithms::spmv::ell_merge_path(ell, x, y);    check("algorithms::spmv::ell_merge_path", y, ref);
    bcsr_t<4, 4, int, int, float> bcsr(csr);
    vector_t<float> xp(bcsr.num_block_cols * 4, 0.0f);
    thrust::copy(x.begin(), x.end(), xp.begin());
    algorithms::spmv::bcsr_thread_mapped(bcsr, xp, y); check("algorithms::spmv::bcsr_thread_mapped<4,4>", y, ref);
    // device-space constructors convert on the device (csrc/convert.cu); host-space
    // ones run the host loops: the arrays must be identical
    {
      auto same = [](const auto& dv, const auto& hv) {
        thrust::host_vector<typename std::decay_t<decltype(hv)>::value_type> back(dv);
        return back.size() == hv.size() && std::equal(back.begin(), back.end(), hv.begin());
      };
      coo_t<int, float, memory_space_t::host> coo_h(h);
      ell_t<int, float, memory_space_t::host> ell_h(h);
      bcsr_t<4, 4, int, int, float, memory_space_t::host> bcsr_h(h);
      csc_t<int, int, float, memory_space_t::host> csc_h(h);
      csc_t<int, int, float> csc_d(csr);
      csr_t<int, int, float> back_d(coo);            // coo -> csr on the device
      bool ok = same(coo.row_indices, coo_h.row_indices) && ell.pitch == ell_h.pitch &&
                same(ell.indices, ell_h.indices) && same(ell.values, ell_h.values) &&
                bcsr.num_blocks == bcsr_h.num_blocks && same(bcsr.block_offsets, bcsr_h.block_offsets) &&
                same(bcsr.block_col_indices, bcsr_h.block_col_indices) && same(bcsr.values, bcsr_h.values) &&
                same(csc_d.offsets, csc_h.offsets) && same(csc_d.indices, csc_h.indices) &&
                same(csc_d.values, csc_h.values) && same(back_d.offsets, h.offsets) &&
                same(back_d.indices, h.indices) && same(back_d.values, h.values);
      if (!ok) ++failures;
      std::printf("%s device conversions == host conversions (coo, ell, bcsr<4,4>, csc, coo->csr)\n", ok ? "OK" : "FAIL");
    }
  } catch (const error::exception_t& e) {
    std::printf("FAIL exception: %s\n", e.what());
    return 100;
  }

  // user kernel on schedule::setup<thread_mapped> with the default CSR layout
  {
    using setup_t = schedule::setup<schedule::algorithms_t::thread_mapped, 1, 1, int, int>;
    setup_t config(thrust::raw_pointer_cast(csr.offsets.data()), csr.rows, csr.nnzs);
    thrust::fill(y.begin(), y.end(), -1.0f);
    user_thread_mapped<<<(rows + 127) / 128, 128>>>(config, thrust::raw_pointer_cast(csr.indices.data()),
                                                    thrust::raw_pointer_cast(csr.values.data()),
                                                    thrust::raw_pointer_cast(x.data()),
                                                    thrust::raw_pointer_cast(y.data()));
    cudaDeviceSynchronize();
    check("user kernel: setup<thread_mapped> (csr)", y, ref);
  }
  // ... and with a user-defined layout passed as the last template argument
  {
    using setup_t = schedule::setup<schedule::algorithms_t::thread_mapped, 1, 1, int, int, std::size_t,
                                    std::size_t, banded_layout>;
    banded_layout lay{thrust::raw_pointer_cast(csr.offsets.data()) + 1, rows, nnz};
    setup_t config(lay);
    thrust::fill(y.begin(), y.end(), -1.0f);
    user_thread_mapped<<<(rows + 127) / 128, 128>>>(config, thrust::raw_pointer_cast(csr.indices.data()),
                                                    thrust::raw_pointer_cast(csr.values.data()),
                                                    thrust::raw_pointer_cast(x.data()),
                                                    thrust::raw_pointer_cast(y.data()));
    cudaDeviceSynchronize();
    check("user kernel: setup<thread_mapped, ..., custom layout>", y, ref);
  }
  // the reference's merge-path kernel body against our setup<merge_path_flat> + preprocess_t
  {
    constexpr std::size_t TPB = 128, IPT = 8;
    using meta_t = schedule::merge_path::preprocess_t<TPB, IPT, int, int, std::size_t, std::size_t>;
    meta_t meta(thrust::raw_pointer_cast(csr.offsets.data()), csr.rows, csr.nnzs, 0);
    const int M = int((rows + nnz + TPB * IPT - 1) / (TPB * IPT));
    thrust::fill(y.begin(), y.end(), 0.0f);
    user_merge_path<TPB, IPT, meta_t><<<M, TPB>>>(meta, csr.rows, csr.nnzs,
                                                 thrust::raw_pointer_cast(csr.offsets.data()),
                                                 thrust::raw_pointer_cast(csr.indices.data()),
                                                 thrust::raw_pointer_cast(csr.values.data()),
                                                 thrust::raw_pointer_cast(x.data()),
                                                 thrust::raw_pointer_cast(y.data()));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) std::printf("FAIL user merge kernel: %s\n", cudaGetErrorString(e));
    check("user kernel: reference merge-path body on setup<merge_path_flat>", y, ref);
  }
  std::printf("failures: %d\n", failures);
  return failures;
}
