// oracle/ref_gpu.cu -- TEST INFRASTRUCTURE, not product code.
//
// Runs the UNMODIFIED gunrock/loops schedule templates and SpMV kernels
// (-I/root/reference/include -DLOOPS_TARGET_ARCH=100) on the GPU box so that
//   * the schedule index streams of oracle/loops_oracle.c and of the product
//     (loopsb_emit_schedule) can be pinned against the reference's own
//     schedule::setup<> classes, and
//   * the reference's kernels can be timed beside the product ("kernel to
//     beat", SURVEY.md section 6 / BASELINE.md section 3).
// The recorder kernels below call the reference's setup API in exactly the
// order its SpMV kernels do and store (thread, step, tile) per atom instead of
// doing the multiply:
//   thread_mapped   algorithms/spmv/thread_mapped.cuh:27-56
//   group_mapped    algorithms/spmv/group_mapped.cuh:27-61
//   work_oriented   algorithms/spmv/work_oriented.cuh:35-89
//   merge_path_flat algorithms/spmv/merge_path_flat.cuh:38-83
// Built by oracle/Makefile into oracle/_ref/libloopsref_gpu.so.

#include <loops/schedule.hxx>
#include <loops/container/formats.hxx>
#include <loops/container/coo.hxx>
#include <loops/container/ell.hxx>
#include <loops/container/bcsr.hxx>
#include <loops/container/vector.hxx>
#include <loops/algorithms/spmv/thread_mapped.cuh>
#include <loops/algorithms/spmv/group_mapped.cuh>
#include <loops/algorithms/spmv/work_oriented.cuh>
#include <loops/algorithms/spmv/merge_path_flat.cuh>
#include <loops/algorithms/spmv/coo_thread_mapped.cuh>
#include <loops/algorithms/spmv/ell_thread_mapped.cuh>
#include <loops/algorithms/spmv/ell_merge_path.cuh>
#include <loops/algorithms/spmv/bcsr_thread_mapped.cuh>
#include <loops/util/launch_box.hxx>

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

using namespace loops;

namespace {

struct rec_t {
  int* visitor;
  int* step;
  int* tile;
  int* visits;
  __device__ void operator()(int a, int g, int s, int t) const {
    visitor[a] = g;
    step[a] = s;
    tile[a] = t;
    atomicAdd(&visits[a], 1);
  }
};

// ---- thread_mapped --------------------------------------------------------
template <typename setup_t>
__global__ void k_thread_mapped(setup_t config, rec_t rec) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0;
  for (auto t : config.tiles())
    for (auto a : config.atoms(t))
      rec((int)a, g, s++, (int)t);
}

// ---- group_mapped (block_mapped<128>) ---------------------------------------
template <typename setup_t, typename layout_t>
__global__ void __launch_bounds__(128) k_group_mapped(layout_t lay, rec_t rec) {
  using storage_t = typename setup_t::storage_t;
  __shared__ storage_t temporary_storage;
  setup_t config(temporary_storage, lay);
  auto p = config.partition();
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0;
  for (auto virtual_atom : config.atom_accessor(p)) {
    auto virtual_tile = config.tile_accessor(virtual_atom, p);
    int my_step = s++;
    if (!(config.is_valid_accessor(virtual_tile, p)))
      continue;
    auto row = config.tile_id(virtual_tile, p);
    auto nz_idx = config.atom_id(virtual_atom, row, virtual_tile, p);
    rec((int)nz_idx, g, my_step, (int)row);
  }
}

// ---- work_oriented --------------------------------------------------------
template <typename setup_t, typename layout_t>
__global__ void __launch_bounds__(128)
    k_work_oriented(layout_t lay, rec_t rec, int* commit, int* map_out) {
  setup_t config(lay);
  auto map = config.init();
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  map_out[4 * g + 0] = (int)map.first.first;
  map_out[4 * g + 1] = (int)map.first.second;
  map_out[4 * g + 2] = (int)map.second.first;
  map_out[4 * g + 3] = (int)map.second.second;
  int s = 0;
  bool first_tile = true;
  for (auto row : config.tiles(map)) {
    for (auto nz : config.atoms(row, map)) {
      rec((int)nz, g, s++, (int)row);
      commit[nz] = first_tile ? 0 : 1;
    }
    first_tile = false;
  }
  __syncthreads();
  for (auto row : config.remainder_tiles(map)) {
    for (auto nz : config.remainder_atoms(map)) {
      rec((int)nz, g, s++, (int)row);
      commit[nz] = 2;
    }
  }
}

// ---- merge_path_flat --------------------------------------------------------
template <typename setup_t, typename meta_t, typename layout_t>
__global__ void __launch_bounds__(int(setup_t::threads_per_block))
    k_merge_path(meta_t meta, layout_t lay, rec_t rec, int* d_tile, int* d_atom,
                 int* d_emit, int* thread_start) {
  using storage_t = typename setup_t::storage_t;
  __shared__ storage_t temporary_storage;
  setup_t config(meta, temporary_storage, lay);
  auto map = config.init();
  if (!config.is_valid_accessor(map))
    return;
  int b = blockIdx.x * gridDim.y + blockIdx.y;
  int g = b * blockDim.x + threadIdx.x;
  thread_start[2 * g] = (int)map.x;
  thread_start[2 * g + 1] = (int)map.y;
  int s = 0;
  for (auto item : config.virtual_idx()) {
    auto nz = config.atom_idx(item, map);
    auto row = config.tile_idx(map);
    int slot = g * int(setup_t::items_per_thread) + s;
    d_tile[slot] = (int)row;
    d_atom[slot] = (int)nz;
    if (config.atoms_counting_it[map.y] <
        temporary_storage.tile_end_offset[map.x]) {
      d_emit[slot] = 1;
      rec((int)nz, g, s, (int)row);
      map.y++;
    } else {
      d_emit[slot] = 0;
      map.x++;
    }
    s++;
  }
}

template <typename T>
struct dbuf {
  T* p = nullptr;
  size_t n = 0;
  explicit dbuf(size_t n_, int fill = -1) : n(n_) {
    cudaMalloc(&p, (n ? n : 1) * sizeof(T));
    cudaMemset(p, fill, (n ? n : 1) * sizeof(T));
  }
  ~dbuf() { cudaFree(p); }
  void to_host(T* h) const {
    if (h && n)
      cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost);
  }
};

enum { KIND_CSR = 0, KIND_COO = 1, KIND_ELL = 2 };

template <typename layout_t>
int emit_with_layout(layout_t lay, int schedule, int grid_blocks, int tpb,
                     int ipt, long long A, int* visitor, int* step, int* tile,
                     int* visits, int* extra_a, int* extra_b, int* d_tile,
                     int* d_atom, int* d_emit, long long dense_len) {
  dbuf<int> v(A), s(A), t(A), c(A, 0);
  rec_t rec{v.p, s.p, t.p, c.p};
  using tm_t = schedule::setup<schedule::algorithms_t::thread_mapped, 1, 1, int,
                               int, std::size_t, std::size_t, layout_t>;
  using gm_t = schedule::setup<schedule::algorithms_t::group_mapped, 128, 128,
                               int, int, std::size_t, std::size_t, layout_t>;
  using wo_t = schedule::setup<schedule::algorithms_t::work_oriented, 128, 1,
                               int, int, std::size_t, std::size_t, layout_t>;
  if (schedule == schedule::algorithms_t::thread_mapped) {
    tm_t config(lay);
    k_thread_mapped<tm_t><<<grid_blocks, tpb>>>(config, rec);
  } else if (schedule == schedule::algorithms_t::group_mapped) {
    k_group_mapped<gm_t, layout_t><<<grid_blocks, 128>>>(lay, rec);
  } else if (schedule == schedule::algorithms_t::work_oriented) {
    dbuf<int> commit(A), map((size_t)grid_blocks * 128 * 4);
    k_work_oriented<wo_t, layout_t>
        <<<grid_blocks, 128>>>(lay, rec, commit.p, map.p);
    cudaDeviceSynchronize();
    commit.to_host(extra_a);
    map.to_host(extra_b);
  } else if (schedule == schedule::algorithms_t::merge_path_flat) {
    dbuf<int> dt(dense_len), da(dense_len), de(dense_len),
        ts((size_t)(dense_len / ipt) * 2);
    long long W = (long long)lay.num_tiles() + (long long)lay.num_atoms();
    auto run = [&](auto tpb_c, auto ipt_c) {
      constexpr std::size_t TPB = decltype(tpb_c)::value;
      constexpr std::size_t IPT = decltype(ipt_c)::value;
      using meta_t =
          schedule::merge_path::preprocess_t<TPB, IPT, int, int, std::size_t,
                                             std::size_t, layout_t>;
      using mp_t = schedule::setup<schedule::algorithms_t::merge_path_flat, TPB,
                                   IPT, int, int, std::size_t, std::size_t,
                                   layout_t>;
      meta_t meta(lay, 0);
      int M = (int)((W + TPB * IPT - 1) / (TPB * IPT));
      if (M > 0)
        k_merge_path<mp_t, meta_t, layout_t>
            <<<dim3(M, 1, 1), TPB>>>(meta, lay, rec, dt.p, da.p, de.p, ts.p);
      cudaDeviceSynchronize();
    };
    if (tpb == 128 && ipt == 8)
      run(std::integral_constant<std::size_t, 128>{},
          std::integral_constant<std::size_t, 8>{});
    else if (tpb == 128 && ipt == 7)
      run(std::integral_constant<std::size_t, 128>{},
          std::integral_constant<std::size_t, 7>{});
    else if (tpb == 128 && ipt == 5)
      run(std::integral_constant<std::size_t, 128>{},
          std::integral_constant<std::size_t, 5>{});
    else
      return 3;
    dt.to_host(d_tile);
    da.to_host(d_atom);
    de.to_host(d_emit);
    ts.to_host(extra_b);
  } else {
    return 2;
  }
  cudaError_t e = cudaDeviceSynchronize();
  v.to_host(visitor);
  s.to_host(step);
  t.to_host(tile);
  c.to_host(visits);
  return e == cudaSuccess ? 0 : 100 + (int)e;
}

}  // namespace

extern "C" {

// kind: 0 csr (h_offsets, T, A), 1 coo (A), 2 ell (T, pitch).
// schedule: reference enum order (schedule.hxx:26-32).
// Outputs are host arrays: visitor/step/tile/visits [A]; for work_oriented
// extra_a = commit[A], extra_b = map[grid*128*4]; for merge_path_flat
// d_tile/d_atom/d_emit [M*tpb*ipt], extra_b = thread_start[M*tpb*2].
int ref_gpu_emit(int kind, const int* h_offsets, int T, int A_in, int pitch,
                 int schedule, int grid_blocks, int tpb, int ipt, int* visitor,
                 int* step, int* tile, int* visits, int* extra_a, int* extra_b,
                 int* d_tile, int* d_atom, int* d_emit, long long dense_len) {
  if (kind == KIND_CSR) {
    dbuf<int> off((size_t)T + 1);
    cudaMemcpy(off.p, h_offsets, ((size_t)T + 1) * sizeof(int),
               cudaMemcpyHostToDevice);
    layout::csr<int, int> lay(off.p, T, A_in);
    return emit_with_layout(lay, schedule, grid_blocks, tpb, ipt, A_in, visitor,
                            step, tile, visits, extra_a, extra_b, d_tile,
                            d_atom, d_emit, dense_len);
  } else if (kind == KIND_COO) {
    layout::coo<int, int> lay(A_in);
    return emit_with_layout(lay, schedule, grid_blocks, tpb, ipt, A_in, visitor,
                            step, tile, visits, extra_a, extra_b, d_tile,
                            d_atom, d_emit, dense_len);
  } else if (kind == KIND_ELL) {
    layout::ell<int, int> lay(T, pitch);
    return emit_with_layout(lay, schedule, grid_blocks, tpb, ipt,
                            (long long)T * pitch, visitor, step, tile, visits,
                            extra_a, extra_b, d_tile, d_atom, d_emit,
                            dense_len);
  }
  return 1;
}

// The grid the reference's own work_oriented wrapper would launch
// (algorithms/spmv/work_oriented.cuh:112-113, util/launch_box.hxx:229-239).
int ref_gpu_work_oriented_grid() {
  auto kernel = algorithms::spmv::__work_oriented<128, int, int, float>;
  return (int)launch_box::occupancy_grid(kernel, 128);
}

// Run the reference's own SpMV wrappers on device containers. which:
// 0 merge_path_flat, 1 work_oriented, 2 thread_mapped, 3 group_mapped (csr);
// 4 coo_thread_mapped, 5 ell_thread_mapped, 6 ell_merge_path.
// y is zero-filled before every run (the atomic kernels require it). Returns
// the best CUDA-event time over `reps` in *ms_best (whole wrapper incl. the
// reference's own preprocess for merge-path; *ms_inner = the timer_t the
// reference itself reports where it returns one, else -1).
int ref_gpu_spmv(int which, int rows, int cols, int nnz, const int* h_off,
                 const int* h_idx, const float* h_val, const float* h_x,
                 float* h_y, int reps, float* ms_best, float* ms_inner) {
  using csr_h = csr_t<int, int, float, memory_space_t::host>;
  csr_h hc(rows, cols, nnz);
  std::copy(h_off, h_off + rows + 1, hc.offsets.begin());
  std::copy(h_idx, h_idx + nnz, hc.indices.begin());
  std::copy(h_val, h_val + nnz, hc.values.begin());
  csr_t<int, int, float> csr(hc);
  vector_t<float> x(h_x, h_x + cols);
  vector_t<float> y(rows, 0.0f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f, inner = -1.0f;
  coo_t<int, float> coo;
  ell_t<int, float> ell;
  if (which == 4)
    coo = coo_t<int, float>(csr);
  if (which == 5 || which == 6)
    ell = ell_t<int, float>(csr);
  for (int r = 0; r < reps; ++r) {
    thrust::fill(y.begin(), y.end(), 0.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    switch (which) {
      case 0: {
        auto t = algorithms::spmv::merge_path_flat(csr, x, y);
        float m = t.milliseconds();
        inner = (inner < 0 || m < inner) ? m : inner;
      } break;
      case 1: algorithms::spmv::work_oriented(csr, x, y); break;
      case 2: algorithms::spmv::thread_mapped(csr, x, y); break;
      case 3: algorithms::spmv::group_mapped(csr, x, y); break;
      case 4: {
        auto t = algorithms::spmv::coo_thread_mapped(coo, x, y);
        float m = t.milliseconds();
        inner = (inner < 0 || m < inner) ? m : inner;
      } break;
      case 5: algorithms::spmv::ell_thread_mapped(ell, x, y); break;
      case 6: {
        auto t = algorithms::spmv::ell_merge_path(ell, x, y);
        float m = t.milliseconds();
        inner = (inner < 0 || m < inner) ? m : inner;
      } break;
      default: return 1;
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best)
      best = ms;
  }
  cudaError_t e = cudaDeviceSynchronize();
  thrust::copy(y.begin(), y.end(), h_y);
  *ms_best = best;
  *ms_inner = inner;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return e == cudaSuccess ? 0 : 100 + (int)e;
}

// BCSR R=C in {2,3,4}, fp32 (the reference kernel has no bf16 path).
int ref_gpu_spmv_bcsr(int R, int rows, int cols, int nnz, const int* h_off,
                      const int* h_idx, const float* h_val,
                      const float* h_x_padded, int x_len, float* h_y) {
  using csr_h = csr_t<int, int, float, memory_space_t::host>;
  csr_h hc(rows, cols, nnz);
  std::copy(h_off, h_off + rows + 1, hc.offsets.begin());
  std::copy(h_idx, h_idx + nnz, hc.indices.begin());
  std::copy(h_val, h_val + nnz, hc.values.begin());
  vector_t<float> x(h_x_padded, h_x_padded + x_len);
  vector_t<float> y(rows, 0.0f);
  auto go = [&](auto rc) {
    constexpr std::size_t RC = decltype(rc)::value;
    bcsr_t<RC, RC, int, int, float, memory_space_t::host> hb(hc);
    bcsr_t<RC, RC, int, int, float> b(hb);
    algorithms::spmv::bcsr_thread_mapped(b, x, y);
  };
  if (R == 2) go(std::integral_constant<std::size_t, 2>{});
  else if (R == 3) go(std::integral_constant<std::size_t, 3>{});
  else if (R == 4) go(std::integral_constant<std::size_t, 4>{});
  else return 1;
  cudaError_t e = cudaDeviceSynchronize();
  thrust::copy(y.begin(), y.end(), h_y);
  return e == cudaSuccess ? 0 : 100 + (int)e;
}

}  // extern "C"
