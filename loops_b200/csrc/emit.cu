// loops_b200/csrc/emit.cu -- loopsb_emit_schedule: run the loops-b200
// schedule::setup<> iterators (include/loops/schedule/*.hxx) with a recording
// body. Parity instrument only; see include/loopsb.h for the record format.
//
// Each recorder calls the setup API in the order the reference's SpMV kernels
// do (reference algorithms/spmv/{thread_mapped.cuh:27-56, group_mapped.cuh:
// 27-61, work_oriented.cuh:35-89, merge_path_flat.cuh:38-83}) so that what is
// compared is exactly what a user kernel would be handed.

#include "common.cuh"

#include <loops/schedule.hxx>

using namespace loops;

namespace {

struct recorder {
  int32_t* visitor;
  int32_t* step;
  int32_t* tile;
  int32_t* visits;
  __device__ __forceinline__ void operator()(long long a, int g, int s,
                                             long long t) const {
    visitor[a] = g;
    step[a] = s;
    tile[a] = static_cast<int32_t>(t);
    atomicAdd(&visits[a], 1);
  }
};

template <typename setup_t>
__global__ void emit_thread_mapped(setup_t config, recorder rec) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0;
  for (auto t : config.tiles())
    for (auto a : config.atoms(t))
      rec(a, g, s++, (long long)t);
}

template <typename setup_t, typename layout_t>
__global__ void __launch_bounds__(128)
    emit_group_mapped(layout_t lay, recorder rec) {
  __shared__ typename setup_t::storage_t scratch;
  setup_t config(scratch, lay);
  auto p = config.partition();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0;
  for (auto virtual_atom : config.atom_accessor(p)) {
    auto virtual_tile = config.tile_accessor(virtual_atom, p);
    const int mine = s++;
    if (!config.is_valid_accessor(virtual_tile, p))
      continue;
    auto row = config.tile_id(virtual_tile, p);
    auto nz = config.atom_id(virtual_atom, row, virtual_tile, p);
    rec(nz, g, mine, row);
  }
}

template <typename setup_t, typename layout_t>
__global__ void __launch_bounds__(128)
    emit_work_oriented(layout_t lay, recorder rec, int32_t* commit,
                       int32_t* map_out) {
  setup_t config(lay);
  auto map = config.init();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  map_out[4 * g + 0] = (int32_t)map.first.first;
  map_out[4 * g + 1] = (int32_t)map.first.second;
  map_out[4 * g + 2] = (int32_t)map.second.first;
  map_out[4 * g + 3] = (int32_t)map.second.second;
  int s = 0;
  bool first_tile = true;
  for (auto row : config.tiles(map)) {
    for (auto nz : config.atoms(row, map)) {
      rec(nz, g, s++, row);
      commit[nz] = first_tile ? 0 : 1;
    }
    first_tile = false;
  }
  __syncthreads();
  for (auto row : config.remainder_tiles(map)) {
    for (auto nz : config.remainder_atoms(map)) {
      rec(nz, g, s++, row);
      commit[nz] = 2;
    }
  }
}

template <typename setup_t, typename meta_t, typename layout_t>
__global__ void __launch_bounds__(int(setup_t::threads_per_block))
    emit_merge_path(meta_t meta, layout_t lay, recorder rec, int32_t* d_tile,
                    int32_t* d_atom, int32_t* d_emit, int32_t* thread_start) {
  __shared__ typename setup_t::storage_t scratch;
  setup_t config(meta, scratch, lay);
  auto map = config.init();
  if (!config.is_valid_accessor(map))
    return;
  const long long b = (long long)blockIdx.x * gridDim.y + blockIdx.y;
  const long long g = b * blockDim.x + threadIdx.x;
  thread_start[2 * g] = (int32_t)map.x;
  thread_start[2 * g + 1] = (int32_t)map.y;
  int s = 0;
  for (auto item : config.virtual_idx()) {
    auto nz = config.atom_idx(item, map);
    auto row = config.tile_idx(map);
    const long long slot = g * (long long)setup_t::items_per_thread + s;
    d_tile[slot] = (int32_t)row;
    d_atom[slot] = (int32_t)nz;
    if (config.atoms_counting_it[map.y] < scratch.tile_end_offset[map.x]) {
      d_emit[slot] = 1;
      rec(nz, (int)g, s, row);
      map.y++;
    } else {
      d_emit[slot] = 0;
      map.x++;
    }
    s++;
  }
}

template <std::size_t TPB, std::size_t IPT, typename layout_t>
int launch_merge(layout_t lay, recorder rec, int32_t* d_tile, int32_t* d_atom,
                 int32_t* d_emit, int32_t* thread_start, int64_t dense_len,
                 cudaStream_t stream) {
  using meta_t = schedule::merge_path::preprocess_t<TPB, IPT, int, int,
                                                    std::size_t, std::size_t,
                                                    layout_t>;
  using setup_t =
      schedule::setup<schedule::algorithms_t::merge_path_flat, TPB, IPT, int,
                      int, std::size_t, std::size_t, layout_t>;
  const long long W = (long long)lay.num_tiles() + (long long)lay.num_atoms();
  const long long M = (W + (long long)(TPB * IPT) - 1) / (long long)(TPB * IPT);
  LOOPSB_REQUIRE(dense_len >= M * (long long)(TPB * IPT),
                 "dense_len too small for M*TPB*IPT records");
  if (M == 0)
    return LOOPSB_OK;
  meta_t meta(lay, stream);
  // Same 2-D grid folding as the reference wrapper
  // (algorithms/spmv/merge_path_flat.cuh:125-127).
  int max_x = 0, dev = 0;
  LOOPSB_CUDA_TRY(cudaGetDevice(&dev));
  LOOPSB_CUDA_TRY(cudaDeviceGetAttribute(&max_x, cudaDevAttrMaxGridDimX, dev));
  const unsigned gx = (unsigned)(M < max_x ? M : max_x);
  const unsigned gy = (unsigned)((M + max_x - 1) / max_x);
  emit_merge_path<setup_t, meta_t, layout_t><<<dim3(gx, gy, 1), TPB, 0, stream>>>(
      meta, lay, rec, d_tile, d_atom, d_emit, thread_start);
  LOOPSB_CUDA_TRY(cudaGetLastError());
  LOOPSB_CUDA_TRY(cudaStreamSynchronize(stream));
  return LOOPSB_OK;
}

template <typename layout_t>
int emit_for_layout(layout_t lay, int schedule, int grid_blocks, int tpb,
                    int ipt, recorder rec, int32_t* extra_a, int32_t* extra_b,
                    int32_t* d_tile, int32_t* d_atom, int32_t* d_emit,
                    int64_t dense_len, cudaStream_t stream) {
  using tm_t = schedule::setup<schedule::algorithms_t::thread_mapped, 1, 1, int,
                               int, std::size_t, std::size_t, layout_t>;
  using gm_t = schedule::setup<schedule::algorithms_t::group_mapped, 128, 128,
                               int, int, std::size_t, std::size_t, layout_t>;
  using wo_t = schedule::setup<schedule::algorithms_t::work_oriented, 128, 1,
                               int, int, std::size_t, std::size_t, layout_t>;
  switch (schedule) {
    case LOOPSB_SCHED_THREAD_MAPPED: {
      LOOPSB_REQUIRE(grid_blocks > 0 && tpb > 0 && tpb <= 1024,
                     "thread_mapped needs grid_blocks and threads_per_block");
      tm_t config(lay);
      emit_thread_mapped<tm_t><<<grid_blocks, tpb, 0, stream>>>(config, rec);
    } break;
    case LOOPSB_SCHED_GROUP_MAPPED: {
      const long long T = lay.num_tiles();
      const unsigned blocks = (unsigned)((T + 127) / 128);
      if (blocks > 0)
        emit_group_mapped<gm_t, layout_t><<<blocks, 128, 0, stream>>>(lay, rec);
    } break;
    case LOOPSB_SCHED_WORK_ORIENTED: {
      LOOPSB_REQUIRE(grid_blocks > 0, "work_oriented needs grid_blocks");
      LOOPSB_REQUIRE(extra_a && extra_b, "work_oriented needs commit and map");
      emit_work_oriented<wo_t, layout_t>
          <<<grid_blocks, 128, 0, stream>>>(lay, rec, extra_a, extra_b);
    } break;
    case LOOPSB_SCHED_MERGE_PATH_FLAT: {
      LOOPSB_REQUIRE(d_tile && d_atom && d_emit && extra_b,
                     "merge_path_flat needs dense_* and thread_start");
      if (tpb == 128 && ipt == 8)
        return launch_merge<128, 8>(lay, rec, d_tile, d_atom, d_emit, extra_b,
                                    dense_len, stream);
      if (tpb == 128 && ipt == 7)
        return launch_merge<128, 7>(lay, rec, d_tile, d_atom, d_emit, extra_b,
                                    dense_len, stream);
      if (tpb == 128 && ipt == 5)
        return launch_merge<128, 5>(lay, rec, d_tile, d_atom, d_emit, extra_b,
                                    dense_len, stream);
      loopsb::set_error("merge_path_flat emit supports (128,8) (128,7) (128,5)");
      return LOOPSB_ERR_UNSUPPORTED;
    }
    default:
      loopsb::set_error("unknown schedule %d", schedule);
      return LOOPSB_ERR_INVALID;
  }
  LOOPSB_CUDA_TRY(cudaGetLastError());
  LOOPSB_CUDA_TRY(cudaStreamSynchronize(stream));
  return LOOPSB_OK;
}

}  // namespace

extern "C" int loopsb_emit_schedule(const loopsb_layout_t* lay, int schedule,
                                    int32_t grid_blocks,
                                    int32_t threads_per_block,
                                    int32_t items_per_thread, int32_t* visitor,
                                    int32_t* step, int32_t* tile,
                                    int32_t* visits, int32_t* extra_a,
                                    int32_t* extra_b, int32_t* dense_tile,
                                    int32_t* dense_atom, int32_t* dense_emit,
                                    int64_t dense_len, void* stream) {
  LOOPSB_REQUIRE(lay != nullptr, "layout descriptor is null");
  LOOPSB_REQUIRE(lay->num_tiles >= 0 && lay->num_atoms >= 0, "negative sizes");
  if (loopsb::current_device() == nullptr)
    return LOOPSB_ERR_CUDA;
  const bool has_atoms = lay->num_atoms > 0;
  LOOPSB_REQUIRE(!has_atoms || (visitor && step && tile && visits),
                 "record arrays are null");
  recorder rec{visitor, step, tile, visits};
  cudaStream_t s = loopsb::as_stream(stream);
  if (loopsb::is_offsets_kind(lay->kind)) {
    LOOPSB_REQUIRE(lay->offsets != nullptr, "offsets-kind layout without offsets");
    layout::csr<int, int> v(lay->offsets, lay->num_tiles, lay->num_atoms);
    return emit_for_layout(v, schedule, grid_blocks, threads_per_block,
                           items_per_thread, rec, extra_a, extra_b, dense_tile,
                           dense_atom, dense_emit, dense_len, s);
  }
  if (lay->kind == LOOPSB_LAYOUT_COO) {
    layout::coo<int, int> v(lay->num_atoms);
    return emit_for_layout(v, schedule, grid_blocks, threads_per_block,
                           items_per_thread, rec, extra_a, extra_b, dense_tile,
                           dense_atom, dense_emit, dense_len, s);
  }
  if (loopsb::is_pitch_kind(lay->kind)) {
    layout::ell<int, int> v(lay->num_tiles, lay->pitch);
    return emit_for_layout(v, schedule, grid_blocks, threads_per_block,
                           items_per_thread, rec, extra_a, extra_b, dense_tile,
                           dense_atom, dense_emit, dense_len, s);
  }
  loopsb::set_error("emit: unsupported layout kind %d", lay->kind);
  return LOOPSB_ERR_UNSUPPORTED;
}
