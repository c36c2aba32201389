// loops_b200/csrc/spmv_generic.cu -- SpMV for the (schedule x layout) cells of BASELINE
// configs[2] that have NO kernel in the reference tree: coo x {group_mapped, work_oriented,
// merge_path_flat} and ell x {group_mapped, work_oriented}. The reference's schedules take any
// layout_type (schedule/group_mapped.hxx:39-199, work_oriented.hxx:45-190,
// merge_path_flat.hxx:186-391), so the definition of these cells is:
// schedule::setup<scheme, ..., layout> handing out (tile, atom) pairs, with the per-atom body of
// the same format's thread_mapped kernel (coo_thread_mapped.cuh:37-51: atomicAdd(y[row[a]],
// v[a] * x[col[a]]); ell_thread_mapped.cuh:28-43: skip col < 0). That is exactly what runs here:
// the loops-b200 setup classes of include/loops/schedule/*.hxx -- the ones whose emitted index
// streams are pinned bit for bit to the reference templates for these layouts
// (tests/test_gpu_streams.py) -- drive the body. y is zeroed by the library first.
#include "common.cuh"

#include <loops/schedule.hxx>

#include <memory>
#include <new>

using namespace loops;

namespace loopsb {
namespace generic {

struct coo_body {
  const int* row;
  const int* col;
  const float* val;
  const float* x;
  float* y;
  __device__ __forceinline__ void operator()(long long /*tile*/, long long a) const {
    atomicAdd(y + row[a], __fmul_rn(val[a], __ldg(x + col[a])));
  }
};
struct ell_body {
  const int* col;
  const float* val;
  const float* x;
  float* y;
  __device__ __forceinline__ void operator()(long long tile, long long a) const {
    const int c = col[a];
    if (c >= 0) atomicAdd(y + tile, __fmul_rn(val[a], __ldg(x + c)));
  }
};

template <typename setup_t, typename layout_t, typename body_t>
__global__ void __launch_bounds__(128) group_kernel(layout_t lay, body_t body) {
  __shared__ typename setup_t::storage_t scratch;
  setup_t config(scratch, lay);
  auto p = config.partition();
  for (auto virtual_atom : config.atom_accessor(p)) {
    auto virtual_tile = config.tile_accessor(virtual_atom, p);
    if (!config.is_valid_accessor(virtual_tile, p)) continue;
    auto row = config.tile_id(virtual_tile, p);
    auto nz = config.atom_id(virtual_atom, row, virtual_tile, p);
    body((long long)row, (long long)nz);
  }
}

template <typename setup_t, typename layout_t, typename body_t>
__global__ void __launch_bounds__(128) work_kernel(layout_t lay, body_t body) {
  setup_t config(lay);
  auto map = config.init();
  for (auto row : config.tiles(map))
    for (auto nz : config.atoms(row, map)) body((long long)row, (long long)nz);
  for (auto row : config.remainder_tiles(map))
    for (auto nz : config.remainder_atoms(map)) body((long long)row, (long long)nz);
}

template <typename setup_t, typename meta_t, typename layout_t, typename body_t>
__global__ void __launch_bounds__(int(setup_t::threads_per_block)) merge_kernel(meta_t meta, layout_t lay, body_t body) {
  __shared__ typename setup_t::storage_t scratch;
  setup_t config(meta, scratch, lay);
  auto map = config.init();
  if (!config.is_valid_accessor(map)) return;
  for (auto item : config.virtual_idx()) {
    auto nz = config.atom_idx(item, map);
    auto row = config.tile_idx(map);
    if (config.atoms_counting_it[map.y] < scratch.tile_end_offset[map.x]) {
      body((long long)row, (long long)nz);
      map.y++;
    } else {
      map.x++;
    }
  }
}

constexpr std::size_t kTPB = 128, kIPT = 8;   // reference launch_t<float> on sm_100 (launch_box.hxx:66-68)
using coo_layout_t = layout::coo<int, int>;
using ell_layout_t = layout::ell<int, int>;
using coo_meta_t = schedule::merge_path::preprocess_t<kTPB, kIPT, int, int, std::size_t, std::size_t, coo_layout_t>;

struct state {
  std::unique_ptr<coo_meta_t> coo_meta;   // merge_path_flat over COO: tile coordinates, once per plan
  int wo_grid = 0;
};

int create(state** out, const loopsb_layout_t* lay, int schedule, int wo_grid, cudaStream_t s) {
  state* st = new (std::nothrow) state();
  if (!st) { set_error("host allocation failed"); return LOOPSB_ERR_ALLOC; }
  st->wo_grid = wo_grid;
  if (schedule == LOOPSB_SCHED_MERGE_PATH_FLAT && lay->kind == LOOPSB_LAYOUT_COO && lay->num_atoms > 0) {
    st->coo_meta.reset(new (std::nothrow) coo_meta_t(coo_layout_t(lay->num_atoms), s));
    if (!st->coo_meta) { delete st; set_error("host allocation failed"); return LOOPSB_ERR_ALLOC; }
    if (cudaStreamSynchronize(s) != cudaSuccess) { delete st; set_error("coordinate kernel failed"); return LOOPSB_ERR_CUDA; }
  }
  *out = st;
  return LOOPSB_OK;
}

void destroy(state* st) { delete st; }

bool supports(int kind, int schedule) {
  if (kind == LOOPSB_LAYOUT_COO)
    return schedule == LOOPSB_SCHED_GROUP_MAPPED || schedule == LOOPSB_SCHED_WORK_ORIENTED ||
           schedule == LOOPSB_SCHED_MERGE_PATH_FLAT;
  if (kind == LOOPSB_LAYOUT_ELL) return schedule == LOOPSB_SCHED_GROUP_MAPPED || schedule == LOOPSB_SCHED_WORK_ORIENTED;
  return false;
}

template <typename layout_t, typename body_t>
int run_for(state* st, layout_t lay, int schedule, body_t body, cudaStream_t s) {
  using gm_t = schedule::setup<schedule::algorithms_t::group_mapped, 128, 128, int, int, std::size_t, std::size_t, layout_t>;
  using wo_t = schedule::setup<schedule::algorithms_t::work_oriented, 128, 1, int, int, std::size_t, std::size_t, layout_t>;
  const long long T = lay.num_tiles();
  if (schedule == LOOPSB_SCHED_GROUP_MAPPED) {
    const unsigned blocks = unsigned((T + 127) / 128);
    if (blocks) group_kernel<gm_t, layout_t, body_t><<<blocks, 128, 0, s>>>(lay, body);
  } else if (schedule == LOOPSB_SCHED_WORK_ORIENTED) {
    work_kernel<wo_t, layout_t, body_t><<<st->wo_grid > 0 ? st->wo_grid : 1, 128, 0, s>>>(lay, body);
  } else {
    set_error("schedule not available for this layout");
    return LOOPSB_ERR_UNSUPPORTED;
  }
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}

int run(state* st, const loopsb_layout_t* lay, int schedule, const float* values, const int* cols, const int* rows,
        const float* x, float* y, int num_rows, cudaStream_t s) {
  LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(float), s));
  if (lay->num_atoms == 0) return LOOPSB_OK;
  if (lay->kind == LOOPSB_LAYOUT_COO) {
    LOOPSB_REQUIRE(rows != nullptr, "COO needs row_indices");
    coo_layout_t v(lay->num_atoms);
    coo_body body{rows, cols, values, x, y};
    if (schedule == LOOPSB_SCHED_MERGE_PATH_FLAT) {
      using setup_t = schedule::setup<schedule::algorithms_t::merge_path_flat, kTPB, kIPT, int, int, std::size_t,
                                      std::size_t, coo_layout_t>;
      LOOPSB_REQUIRE(st->coo_meta != nullptr, "plan holds no merge-path coordinates");
      const long long M = (long long)st->coo_meta->merge_tiles();
      int max_x = 0, dev = 0;
      LOOPSB_CUDA_TRY(cudaGetDevice(&dev));
      LOOPSB_CUDA_TRY(cudaDeviceGetAttribute(&max_x, cudaDevAttrMaxGridDimX, dev));
      const unsigned gx = unsigned(M < max_x ? M : max_x), gy = unsigned((M + max_x - 1) / max_x);
      merge_kernel<setup_t, coo_meta_t, coo_layout_t, coo_body><<<dim3(gx, gy, 1), kTPB, 0, s>>>(*st->coo_meta, v, body);
      LOOPSB_CUDA_TRY(cudaGetLastError());
      return LOOPSB_OK;
    }
    return run_for(st, v, schedule, body, s);
  }
  if (lay->kind == LOOPSB_LAYOUT_ELL) {
    ell_layout_t v(lay->num_tiles, lay->pitch);
    return run_for(st, v, schedule, ell_body{cols, values, x, y}, s);
  }
  set_error("generic SpMV cells exist for COO and ELL");
  return LOOPSB_ERR_UNSUPPORTED;
}

}  // namespace generic
}  // namespace loopsb
