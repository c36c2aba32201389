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

// ---- the other in-tree layouts in double (the reference builds every example as .f64 too,
// examples/spmv/CMakeLists.txt:29). Same thread->work maps as the fp32 kernels of
// spmv_schedules.cuh, un-fused multiply / add; the scatter formats accumulate with fp64 atomics
// into a y the library zeroes, exactly where the reference's kernels use atomics.
__global__ void __launch_bounds__(kThreads)
    coo_f64_kernel(const int* __restrict__ row, const int* __restrict__ col, const double* __restrict__ val,
                   const double* __restrict__ x, double* __restrict__ y, long long nnz) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a < nnz) atomicAdd(&y[row[a]], __dmul_rn(val[a], x[col[a]]));
}

__global__ void __launch_bounds__(kThreads)
    ell_f64_thread_kernel(const int* __restrict__ idx, const double* __restrict__ val, const double* __restrict__ x,
                          double* __restrict__ y, int rows, int pitch) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const long long first = (long long)r * pitch;
  double sum = 0.0;
  for (int s = 0; s < pitch; ++s) {
    const int c = idx[first + s];
    if (c >= 0) sum = __dadd_rn(sum, __dmul_rn(val[first + s], x[c]));
  }
  y[r] = sum;
}

// ell_merge_path in double: a warp per row, lanes stride over the row's slots (coalesced), shuffle tree.
__global__ void __launch_bounds__(kThreads)
    ell_f64_warp_kernel(const int* __restrict__ idx, const double* __restrict__ val, const double* __restrict__ x,
                        double* __restrict__ y, int rows, int pitch) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const long long first = (long long)r * pitch;
  double sum = 0.0;
  for (int s = lane; s < pitch; s += 32) {
    const int c = idx[first + s];
    if (c >= 0) sum = __dadd_rn(sum, __dmul_rn(val[first + s], x[c]));
  }
  for (int d = 16; d > 0; d >>= 1) sum = __dadd_rn(sum, __shfl_down_sync(0xffffffffu, sum, d));
  if (lane == 0) y[r] = sum;
}

__global__ void __launch_bounds__(kThreads)
    csc_f64_kernel(const int* __restrict__ off, const int* __restrict__ row, const double* __restrict__ val,
                   const double* __restrict__ x, double* __restrict__ y, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const double xc = x[c];
  for (int a = off[c]; a < off[c + 1]; ++a) atomicAdd(&y[row[a]], __dmul_rn(val[a], xc));
}

__global__ void __launch_bounds__(kThreads)
    flat_f64_kernel(const int* __restrict__ off, const int* __restrict__ idx, const double* __restrict__ val,
                    const double* __restrict__ x, double* __restrict__ y, int rows, int nnz, int K) {
  const long long windows = ((long long)nnz + K - 1) / K;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= windows) return;
  const int a0 = int(t * K), a1 = min(nnz, a0 + K);
  int lo = 0, hi = rows;                       // offsets[lo] <= a0 < offsets[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= a0) lo = mid; else hi = mid;
  }
  int r = lo, r_end = off[r + 1];
  double acc = 0.0;
  for (int a = a0; a < a1; ++a) {
    while (a >= r_end) {
      if (acc != 0.0) atomicAdd(&y[r], acc);
      acc = 0.0;
      ++r;
      r_end = off[r + 1];
    }
    acc = __dadd_rn(acc, __dmul_rn(val[a], x[idx[a]]));
  }
  if (acc != 0.0) atomicAdd(&y[r], acc);
}

__global__ void __launch_bounds__(kThreads)
    dia_f64_kernel(const int* __restrict__ diag, const double* __restrict__ val, const double* __restrict__ x,
                   double* __restrict__ y, int rows, int cols, long long stride, int nd) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double acc = 0.0;
  for (int d = 0; d < nd; ++d) {
    const long long c = (long long)r + diag[d];
    if (c >= 0 && c < cols) acc = __dadd_rn(acc, __dmul_rn(val[(long long)d * stride + r], x[c]));
  }
  y[r] = acc;
}

template <int R, int C>
__global__ void __launch_bounds__(kThreads)
    bcsr_f64_kernel(const int* __restrict__ boff, const int* __restrict__ bcol, const double* __restrict__ val,
                    const double* __restrict__ x, double* __restrict__ y, int block_rows, int rows) {
  const int br = blockIdx.x * blockDim.x + threadIdx.x;
  if (br >= block_rows) return;
  double acc[R];
#pragma unroll
  for (int i = 0; i < R; ++i) acc[i] = 0.0;
  for (int b = boff[br]; b < boff[br + 1]; ++b) {
    const long long bc = bcol[b];
    const double* blk = val + (long long)b * R * C;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) acc[i] = __dadd_rn(acc[i], __dmul_rn(blk[i * C + j], x[bc * C + j]));
  }
#pragma unroll
  for (int i = 0; i < R; ++i)
    if ((long long)br * R + i < rows) y[(long long)br * R + i] = acc[i];
}

}  // namespace

extern "C" int loopsb_spmv_layout_f64(const loopsb_layout_t* lay, int schedule, const double* values,
                                      const int32_t* col_indices, const int32_t* row_indices, const double* x,
                                      double* y, int32_t num_rows, int32_t num_cols, void* stream) {
  LOOPSB_REQUIRE(lay != nullptr, "layout is null");
  if (lay->kind == LOOPSB_LAYOUT_CSR)
    return loopsb_spmv_f64(lay, schedule, values, col_indices, x, y, num_rows, num_cols, stream);
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0 && lay->num_atoms >= 0 && lay->num_tiles >= 0, "bad dimensions");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  if (num_rows == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(y != nullptr, "y is null");
  cudaStream_t s = as_stream(stream);
  const long long A = lay->num_atoms;
  if (A == 0) {
    LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(double), s));
    return LOOPSB_OK;
  }
  LOOPSB_REQUIRE(values && col_indices && x, "null matrix / x pointer");
  auto blocks = [](long long n) { return unsigned((n + kThreads - 1) / kThreads); };
  switch (lay->kind) {
    case LOOPSB_LAYOUT_COO:
      LOOPSB_REQUIRE(schedule == LOOPSB_SCHED_THREAD_MAPPED && row_indices != nullptr, "COO: thread_mapped with row ids");
      LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(double), s));
      coo_f64_kernel<<<blocks(A), kThreads, 0, s>>>(row_indices, col_indices, values, x, y, A);
      break;
    case LOOPSB_LAYOUT_ELL:
      LOOPSB_REQUIRE(lay->num_tiles == num_rows && lay->pitch >= 0, "ELL: tiles must equal rows");
      if (schedule == LOOPSB_SCHED_THREAD_MAPPED)
        ell_f64_thread_kernel<<<blocks(num_rows), kThreads, 0, s>>>(col_indices, values, x, y, num_rows, lay->pitch);
      else if (schedule == LOOPSB_SCHED_MERGE_PATH_FLAT)
        ell_f64_warp_kernel<<<blocks((long long)num_rows * 32), kThreads, 0, s>>>(col_indices, values, x, y, num_rows,
                                                                                   lay->pitch);
      else { set_error("ELL in double: thread_mapped or merge_path_flat"); return LOOPSB_ERR_UNSUPPORTED; }
      break;
    case LOOPSB_LAYOUT_CSC:
      LOOPSB_REQUIRE(schedule == LOOPSB_SCHED_THREAD_MAPPED && lay->offsets && lay->num_tiles == num_cols, "CSC: thread_mapped over columns");
      LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(double), s));
      csc_f64_kernel<<<blocks(num_cols), kThreads, 0, s>>>(lay->offsets, col_indices, values, x, y, num_cols);
      break;
    case LOOPSB_LAYOUT_FLAT:
      LOOPSB_REQUIRE(schedule == LOOPSB_SCHED_THREAD_MAPPED && lay->offsets && lay->pitch > 0, "flat_uniform_occupancy: thread_mapped, K > 0");
      LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(double), s));
      flat_f64_kernel<<<blocks((A + lay->pitch - 1) / lay->pitch), kThreads, 0, s>>>(lay->offsets, col_indices, values, x, y,
                                                                                     num_rows, int(A), lay->pitch);
      break;
    default:
      set_error("no fp64 SpMV for this layout kind here (DIA / BCSR have their own entry points)");
      return LOOPSB_ERR_UNSUPPORTED;
  }
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}

extern "C" int loopsb_spmv_dia_f64(int32_t num_rows, int32_t num_cols, int64_t stride, int32_t num_diagonals,
                                   const int32_t* diag_offsets, const double* values, const double* x, double* y,
                                   void* stream) {
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0 && num_diagonals >= 0 && stride >= num_rows, "bad dimensions");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  if (num_rows == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(y != nullptr && (num_diagonals == 0 || (diag_offsets && values && x)), "null pointer");
  dia_f64_kernel<<<(num_rows + kThreads - 1) / kThreads, kThreads, 0, as_stream(stream)>>>(
      diag_offsets, values, x, y, num_rows, num_cols, stride, num_diagonals);
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}

extern "C" int loopsb_spmv_bcsr_f64(int32_t R, int32_t C, const loopsb_layout_t* lay, const double* values,
                                    const int32_t* block_col_indices, const double* x_padded, double* y,
                                    int32_t num_rows, void* stream) {
  LOOPSB_REQUIRE(lay != nullptr && lay->kind == LOOPSB_LAYOUT_BCSR && lay->offsets != nullptr, "BCSR layout required");
  LOOPSB_REQUIRE(num_rows >= 0, "negative rows");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  if (num_rows == 0 || lay->num_tiles == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(y != nullptr && (lay->num_atoms == 0 || (values && block_col_indices && x_padded)), "null pointer");
  const int br = lay->num_tiles;
  const unsigned grid = unsigned((br + kThreads - 1) / kThreads);
  cudaStream_t s = as_stream(stream);
  if (R == 2 && C == 2) bcsr_f64_kernel<2, 2><<<grid, kThreads, 0, s>>>(lay->offsets, block_col_indices, values, x_padded, y, br, num_rows);
  else if (R == 3 && C == 3) bcsr_f64_kernel<3, 3><<<grid, kThreads, 0, s>>>(lay->offsets, block_col_indices, values, x_padded, y, br, num_rows);
  else if (R == 4 && C == 4) bcsr_f64_kernel<4, 4><<<grid, kThreads, 0, s>>>(lay->offsets, block_col_indices, values, x_padded, y, br, num_rows);
  else { set_error("BCSR block shape %dx%d not instantiated (2x2, 3x3, 4x4)", R, C); return LOOPSB_ERR_UNSUPPORTED; }
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}

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
