// loops_b200/csrc/spmv_f64.cu -- fp64 CSR SpMV (the reference's
// examples are also built for double, examples/spmv/CMakeLists.txt:29, with the
// merge-path tuning 128 x 4 = 512 items per tile, algorithms/spmv/launch_box.hxx:68).
//
// Two kernels behind loopsb_spmv_f64:
//  * thread_mapped: one thread per row, un-fused multiply and add in atom order --
//    the arithmetic of reference::spmv<double> (util/reference.hxx:61-76), so y is
//    bit-equal to the CPU validator (algorithms/spmv/thread_mapped.cuh:27-56);
//  * merge_path_flat (also serves work_oriented and group_mapped, whose fp64
//    results differ from it only in summation order): CTA tiles of 512 merge items
//    cut by two diagonal searches over row_offsets (schedule/merge_path_flat.hxx:
//    267-335); the tile's products are formed with coalesced loads into shared
//    memory, one thread then sums each row of the tile in atom order. Rows that lie
//    inside one tile are stored; rows cut by a tile boundary are accumulated with
//    fp64 atomics into a y zeroed by the library (the reference accumulates EVERY
//    nonzero that way, merge_path_flat.cuh:71-82).
// fp64 is not a BASELINE config: these kernels are correct and coalesced, not tuned.
#include "common.cuh"

using namespace loopsb;

namespace {

constexpr int kThreads = 128;
constexpr int kItems = 512;   // merge items (row ends + atoms) per CTA tile

// first i in [max(d - A, 0), min(d, T)) with row_end(i) > d - i - 1  (util/search.hxx:34-60)
__device__ __forceinline__ void diagonal_search(long long d, const int* __restrict__ off, long long T, long long A,
                                                long long& ox, long long& oy) {
  long long lo = d - A > 0 ? d - A : 0;
  long long hi = d < T ? d : T;
  while (lo < hi) {
    const long long mid = lo + ((hi - lo) >> 1);
    if ((long long)off[mid + 1] <= d - mid - 1) lo = mid + 1; else hi = mid;
  }
  ox = lo < T ? lo : T;
  oy = d - lo;
}

__global__ void __launch_bounds__(kThreads)
    spmv_f64_thread_mapped_kernel(const int* __restrict__ off, const int* __restrict__ idx,
                                  const double* __restrict__ val, const double* __restrict__ x,
                                  double* __restrict__ y, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double sum = 0.0;
  for (int a = off[r]; a < off[r + 1]; ++a) sum = __dadd_rn(sum, __dmul_rn(val[a], x[idx[a]]));
  y[r] = sum;
}

__global__ void __launch_bounds__(kThreads)
    spmv_f64_merge_kernel(const int* __restrict__ off, const int* __restrict__ idx, const double* __restrict__ val,
                          const double* __restrict__ x, double* __restrict__ y, int rows, long long nnz) {
  __shared__ double prod[kItems];
  __shared__ long long coord[4];
  const long long T = rows, A = nnz, W = T + A;
  if (threadIdx.x < 2) {
    long long d = (long long)(blockIdx.x + threadIdx.x) * kItems;
    if (d > W) d = W;
    diagonal_search(d, off, T, A, coord[2 * threadIdx.x], coord[2 * threadIdx.x + 1]);
  }
  __syncthreads();
  const long long sx = coord[0], sy = coord[1], ex = coord[2], ey = coord[3];
  const int na = int(ey - sy);
  for (int i = threadIdx.x; i < na; i += kThreads) {
    const long long a = sy + i;
    prod[i] = __dmul_rn(val[a], x[idx[a]]);
  }
  __syncthreads();
  const long long last = ex < T - 1 ? ex : T - 1;
  for (long long r = sx + threadIdx.x; r <= last; r += kThreads) {
    const long long rb = off[r], re = off[r + 1];
    const long long a0 = rb > sy ? rb : sy, a1 = re < ey ? re : ey;
    double sum = 0.0;
    for (long long a = a0; a < a1; ++a) sum = __dadd_rn(sum, prod[a - sy]);
    if (rb >= sy && re <= ey) y[r] = sum;          // the whole row lies in this tile (0 for an empty row)
    else if (a1 > a0) atomicAdd(&y[r], sum);       // cut by a tile boundary
  }
}

}  // namespace

extern "C" int loopsb_spmv_f64(const loopsb_layout_t* lay, int schedule, const double* values,
                               const int32_t* col_indices, const double* x, double* y, int32_t num_rows,
                               int32_t num_cols, void* stream) {
  LOOPSB_REQUIRE(lay != nullptr && lay->kind == LOOPSB_LAYOUT_CSR, "CSR layout required");
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0 && lay->num_tiles == num_rows && lay->num_atoms >= 0, "bad dimensions");
  LOOPSB_REQUIRE(schedule == LOOPSB_SCHED_MERGE_PATH_FLAT || schedule == LOOPSB_SCHED_WORK_ORIENTED ||
                     schedule == LOOPSB_SCHED_THREAD_MAPPED || schedule == LOOPSB_SCHED_GROUP_MAPPED,
                 "unknown schedule");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  if (num_rows == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(y != nullptr && lay->offsets != nullptr, "null y / offsets");
  LOOPSB_REQUIRE(lay->num_atoms == 0 || (values && col_indices && x), "null matrix / x pointer");
  cudaStream_t s = as_stream(stream);
  if (schedule == LOOPSB_SCHED_THREAD_MAPPED) {
    spmv_f64_thread_mapped_kernel<<<(num_rows + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        lay->offsets, col_indices, values, x, y, num_rows);
  } else {
    const long long W = (long long)num_rows + lay->num_atoms;
    const long long tiles = (W + kItems - 1) / kItems;
    LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(double), s));
    spmv_f64_merge_kernel<<<unsigned(tiles), kThreads, 0, s>>>(lay->offsets, col_indices, values, x, y, num_rows,
                                                              (long long)lay->num_atoms);
  }
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}
