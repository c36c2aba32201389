// loops_b200/csrc/spmv_schedules.cuh -- SpMV kernels for the thread_mapped,
// group_mapped and work_oriented schedules on CSR, plus the COO / ELL / BCSR
// thread_mapped kernels. Each keeps the reference's thread->work map (so the
// index streams of include/loops/schedule/*.hxx describe them) and changes how
// the data moves:
//
//   thread_mapped (ref algorithms/spmv/thread_mapped.cuh:27-56)
//       sequential left-to-right dot per row with un-fused mul/add: y is
//       bit-identical to the reference CPU validator (util/reference.hxx:61-76).
//   group_mapped  (ref group_mapped.cuh:27-61, block_mapped<128>)
//       the block's 129 row offsets arrive by one bulk async copy; the prefix
//       sums the reference builds with a CG scan ARE those offsets minus the
//       first; the block's atoms are one contiguous, coalesced range; per-atom
//       atomics to y are replaced by a warp-shuffle segmented reduce into 128
//       shared accumulators and one coalesced store.
//   work_oriented (ref work_oriented.cuh:35-89)
//       the two diagonal searches of every thread run on a row-offset window
//       staged in shared memory by a bulk async copy (block-level coordinates
//       bound the window); commit rules (first complete row / later rows /
//       remainder) follow the reference, y is zeroed by the library first.
//   coo_thread_mapped (ref coo_thread_mapped.cuh:37-51)
//       one thread per nonzero; equal-row runs inside a warp are reduced with
//       shuffles before one atomic per run.
//   ell_thread_mapped (ref ell_thread_mapped.cuh:28-43), bcsr_thread_mapped
//       (ref bcsr_thread_mapped.cuh:36-74): same maps, read-only-path loads.
#pragma once

#include "common.cuh"

#include <loops/util/tma.hxx>

namespace loopsb {
namespace sk {

// ------------------------------------------------------------------ thread --
// thread_mapped keeps the reference's map (thread g owns rows g, g+G, ...; each
// row summed left to right by its thread -- ref thread_mapped.cuh:27-56), but the
// 32 rows of a warp are one CONTIGUOUS range of CSR atoms, so the warp reads that
// range coalesced (indices, values, then the x gathers, all 32 lanes wide), parks
// the un-fused products in shared memory, and every lane then adds ITS row's
// products in order. Same adds in the same order as one-thread-per-row (bit-exact
// vs the CPU validator), without the per-lane strided streams.
constexpr int kThreadChunk = 256;   // atoms staged per warp and round

__global__ void __launch_bounds__(128)
    spmv_thread_mapped_csr(const int* __restrict__ offsets,
                           const int* __restrict__ indices,
                           const float* __restrict__ values,
                           const float* __restrict__ x, float* __restrict__ y,
                           int rows) {
  __shared__ float prod[4][kThreadChunk];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stride = gridDim.x * blockDim.x;
  // whole warps iterate together (rows beyond the end behave like empty rows)
  for (int row0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; row0 < rows; row0 += stride) {
    const int row = row0 + lane;
    const int last = min(row0 + 32, rows);            // one past the warp's last row
    const int a_end = __ldg(offsets + last);
    const int beg = row < rows ? __ldg(offsets + row) : a_end;
    const int end = row < rows ? __ldg(offsets + row + 1) : a_end;
    const int a0 = __shfl_sync(0xffffffffu, beg, 0);
    float sum = 0.0f;
    int cur = beg;
    for (int base = a0; base < a_end; base += kThreadChunk) {
#pragma unroll
      for (int j = 0; j < kThreadChunk / 32; ++j) {
        const int a = base + j * 32 + lane;
        if (a < a_end) prod[warp][j * 32 + lane] = __fmul_rn(__ldg(values + a), __ldg(x + __ldg(indices + a)));
      }
      __syncwarp();
      const int stop = min(end, base + kThreadChunk);
      for (; cur < stop; ++cur) sum = __fadd_rn(sum, prod[warp][cur - base]);
      __syncwarp();
    }
    if (row < rows) y[row] = sum;
  }
}

// ------------------------------------------------------------------- group --
// Sum `v` over runs of equal `key` among CONSECUTIVE lanes; the first lane of
// each run ends up with the run total. A run ends where the key changes: keys
// need not be sorted (COO triples arrive in file order), so an equal key further
// up the warp with a different one in between starts a new run -- the sum is a
// true segmented one, bounded by the next run head (ballot), not by key equality.
__device__ __forceinline__ float warp_run_sum(float v, int key, unsigned mask) {
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(mask, key, 1);
  const unsigned heads = __ballot_sync(mask, lane == 0 || prev != key);
  const unsigned above = lane == 31 ? 0u : heads & (0xfffffffeu << lane);
  const int stop = above ? __ffs(above) - 1 : 32;     // first lane of the next run
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float ov = __shfl_down_sync(mask, v, d);
    if (lane + d < stop) v += ov;
  }
  return v;
}

constexpr int kGroupThreads = 128;

__global__ void __launch_bounds__(kGroupThreads)
    spmv_group_mapped_csr(const int* __restrict__ offsets,
                          const int* __restrict__ indices,
                          const float* __restrict__ values,
                          const float* __restrict__ x, float* __restrict__ y,
                          int rows) {
  __shared__ __align__(16) int offs[kGroupThreads + 4];  // offsets[base..base+len]
  __shared__ float acc[kGroupThreads];
  __shared__ unsigned long long bar;
  const int r = threadIdx.x;
  const int base = blockIdx.x * kGroupThreads;
  int len = rows - base;
  if (len > kGroupThreads) len = kGroupThreads;
  const int need = len + 1;
  const int* src = offsets + base;
  const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
  const int bulk = aligned ? (need & ~3) : 0;
  uint64_t* mb = reinterpret_cast<uint64_t*>(&bar);
  if (r == 0) {
    loops::tma::barrier_init(mb, 1);
    if (bulk > 0) {
      loops::tma::barrier_arrive_expect_tx(mb, 4u * bulk);
      loops::tma::bulk_g2s(offs, src, 4u * bulk, mb);
    }
  }
  acc[r] = 0.0f;
  for (int i = bulk + r; i < need; i += kGroupThreads) offs[i] = __ldg(src + i);
  __syncthreads();
  if (bulk > 0) loops::tma::barrier_wait(mb, 0);
  __syncthreads();

  const int first_atom = offs[0];
  const int total = offs[len] - first_atom;  // == tile_aggregates of the block
  int vt = 0;                                // tiles are visited in order
  for (int vbase = 0; vbase < total; vbase += kGroupThreads) {
    const int v = vbase + r;
    const bool live = v < total;
    float p = 0.0f;
    int key = -1;
    if (live) {
      // upper_bound over the prefix (offs[i] - first_atom), resumed at vt
      int lo = vt, cnt = len - vt;
      while (cnt > 0) {
        const int half = cnt >> 1;
        if (!(v < offs[lo + half] - first_atom)) { lo += half + 1; cnt -= half + 1; }
        else cnt = half;
      }
      vt = lo - 1;
      key = vt;
      const int nz = first_atom + v;
      p = __fmul_rn(__ldg(values + nz), __ldg(x + __ldg(indices + nz)));
    }
    const float run = warp_run_sum(p, key, 0xffffffffu);
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    if (live && ((r & 31) == 0 || prev != key)) atomicAdd(&acc[key], run);
  }
  __syncthreads();
  if (r < len) y[base + r] = acc[r];
}

// ----------------------------------------------------------- work oriented --
constexpr int kWorkThreads = 128;
constexpr int kWorkWindow = 1016;  // staged row ends per block (4 KB: keeps 16 blocks/SM inside the 100 KB carve-out)

__device__ __forceinline__ void work_search(long long d, const int* ends,
                                            long long ends_first, long long lo_x,
                                            long long hi_x, long long T,
                                            long long A, long long& ox,
                                            long long& oy) {
  // ends[i - ends_first] == tile_end(i); search restricted to [lo_x, hi_x]
  long long lo = d - A; if (lo < 0) lo = 0;
  long long hi = d < T ? d : T;
  if (lo < lo_x) lo = lo_x;
  if (hi > hi_x) hi = hi_x;
  const long long x_min = lo;
  while (lo < hi) {
    const long long mid = lo + ((hi - lo) >> 1);
    if ((long long)ends[mid - ends_first] <= d - mid - 1) lo = mid + 1; else hi = mid;
  }
  if (hi < x_min) lo = x_min;
  ox = lo < T ? lo : T;
  oy = d - lo;
}

__global__ void __launch_bounds__(kWorkThreads)
    spmv_work_oriented_csr(const int* __restrict__ offsets,
                           const int* __restrict__ indices,
                           const float* __restrict__ values,
                           const float* __restrict__ x, float* __restrict__ y,
                           int rows, int nnz) {
  __shared__ __align__(16) int window[kWorkWindow + 8];
  __shared__ long long bounds[4];
  __shared__ unsigned long long bar;
  const long long T = rows, A = nnz, W = T + A;
  const long long N = (long long)gridDim.x * kWorkThreads;
  const long long w = (W + N - 1) / N;
  const int* row_end = offsets + 1;
  const long long g = (long long)blockIdx.x * kWorkThreads + threadIdx.x;

  // block-level coordinates: two threads, global search
  if (threadIdx.x < 2) {
    long long d = w * (long long)(blockIdx.x + threadIdx.x) * kWorkThreads;
    if (d > W) d = W;
    long long bx, by;
    work_search(d, row_end, 0, 0, T, T, A, bx, by);
    bounds[2 * threadIdx.x] = bx;
    bounds[2 * threadIdx.x + 1] = by;
  }
  uint64_t* mb = reinterpret_cast<uint64_t*>(&bar);
  if (threadIdx.x == 0) loops::tma::barrier_init(mb, 1);
  __syncthreads();
  const long long bx0 = bounds[0], bx1 = bounds[2];
  // row ends needed by this block: indices bx0 .. min(bx1, T-1)
  long long last = bx1 < T - 1 ? bx1 : T - 1;
  const long long need = last - bx0 + 1;
  const bool staged = need > 0 && need <= kWorkWindow;
  const int* ends = row_end;
  long long ends_first = 0;
  if (staged) {
    const int* src = row_end + bx0;
    const int skew = int((reinterpret_cast<uintptr_t>(src) & 15u) >> 2);
    const int covered = skew + int(need);
    // stay inside the offsets array: it has T + 1 entries, row_end = offsets+1
    const long long room = (long long)skew + (T - bx0);
    int bulk = covered & ~3;
    if (bulk > (room & ~3LL)) bulk = int(room & ~3LL);
    if (threadIdx.x == 0 && bulk > 0) {
      loops::tma::barrier_arrive_expect_tx(mb, 4u * bulk);
      loops::tma::bulk_g2s(window, src - skew, 4u * bulk, mb);
    }
    for (int i = bulk + threadIdx.x; i < covered; i += kWorkThreads)
      window[i] = __ldg(src - skew + i);
    if (bulk > 0) loops::tma::barrier_wait(mb, 0);
    __syncthreads();
    ends = window + skew;
    ends_first = bx0;
  }

  long long d0 = w * g; if (d0 > W) d0 = W;
  long long d1 = d0 + w; if (d1 > W) d1 = W;
  long long sx, sy, ex, ey;
  work_search(d0, ends, ends_first, bx0, bx1, T, A, sx, sy);
  work_search(d1, ends, ends_first, bx0, bx1, T, A, ex, ey);

  float sum = 0.0f;
  bool first_tile = true;
  long long cur = sy;
  for (long long row = sx; row < ex; ++row) {
    const long long stop = ends[row - ends_first];
    long long nz = cur;
    for (; nz + 4 <= stop; nz += 4) {   // four independent index->x chains
      const int c0 = __ldg(indices + nz), c1 = __ldg(indices + nz + 1);
      const int c2 = __ldg(indices + nz + 2), c3 = __ldg(indices + nz + 3);
      const float x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
      sum += __ldg(values + nz) * x0;
      sum += __ldg(values + nz + 1) * x1;
      sum += __ldg(values + nz + 2) * x2;
      sum += __ldg(values + nz + 3) * x3;
    }
    for (; nz < stop; ++nz)
      sum += __ldg(values + nz) * __ldg(x + __ldg(indices + nz));
    cur = stop;
    if (first_tile) {
      if (sum != 0.0f) atomicAdd(y + row, sum);
      first_tile = false;
    } else {
      y[row] = sum;
    }
    sum = 0.0f;
  }
  __syncthreads();
  for (long long nz = cur; nz < ey; ++nz)
    sum += __ldg(values + nz) * __ldg(x + __ldg(indices + nz));
  if (sum != 0.0f) atomicAdd(y + ex, sum);
}

// --------------------------------------------------------------------- coo --
__global__ void __launch_bounds__(128)
    spmv_coo_thread_mapped(const int* __restrict__ row_indices,
                           const int* __restrict__ col_indices,
                           const float* __restrict__ values,
                           const float* __restrict__ x, float* __restrict__ y,
                           int nnz) {
  const int stride = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // whole warps iterate together so the shuffles stay convergent
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < nnz;
       base += stride) {
    const int a = base + lane;
    const bool live = a < nnz;
    int row = -1;
    float p = 0.0f;
    if (live) {
      row = __ldg(row_indices + a);
      p = __fmul_rn(__ldg(values + a), __ldg(x + __ldg(col_indices + a)));
    }
    const float run = warp_run_sum(p, row, 0xffffffffu);
    const int prev = __shfl_up_sync(0xffffffffu, row, 1);
    if (live && (lane == 0 || prev != row)) atomicAdd(y + row, run);
  }
}

// --------------------------------------------------------------------- ell --
__global__ void __launch_bounds__(128)
    spmv_ell_thread_mapped(const int* __restrict__ indices,
                           const float* __restrict__ values,
                           const float* __restrict__ x, float* __restrict__ y,
                           int rows, int pitch) {
  const int stride = gridDim.x * blockDim.x;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < rows;
       row += stride) {
    const long long first = (long long)row * pitch;
    float sum = 0.0f;
    for (int s = 0; s < pitch; ++s) {
      const int col = __ldg(indices + first + s);
      if (col >= 0)
        sum = __fadd_rn(sum, __fmul_rn(__ldg(values + first + s), __ldg(x + col)));
    }
    y[row] = sum;
  }
}

// --------------------------------------------------------------------- csc --
// Replaces reference algorithms/spmv/csc_thread_mapped.cuh:29-41: one thread
// per column (tile), x[col] read once, one atomic per stored entry (the format
// scatters rows, there is nothing to reduce locally). y is zeroed by the caller
// side of the C ABI.
__global__ void __launch_bounds__(128)
    spmv_csc_thread_mapped(const int* __restrict__ offsets,
                           const int* __restrict__ row_indices,
                           const float* __restrict__ values,
                           const float* __restrict__ x, float* __restrict__ y,
                           int cols) {
  const int stride = gridDim.x * blockDim.x;
  for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < cols; col += stride) {
    const float xc = __ldg(x + col);
    const int end = __ldg(offsets + col + 1);
    for (int a = __ldg(offsets + col); a < end; ++a)
      atomicAdd(y + __ldg(row_indices + a), __fmul_rn(__ldg(values + a), xc));
  }
}

// --------------------------------------------------------------------- dia --
// Replaces reference algorithms/spmv/dia_thread_mapped.cuh:33-54: one thread
// per row, diagonals in ascending offset order (= ascending column), values
// column-major values[d * stride + r] so a warp reads 128 contiguous bytes per
// diagonal. Un-fused multiply/add: bit-identical to the CPU validator.
__global__ void __launch_bounds__(128)
    spmv_dia_thread_mapped(const int* __restrict__ diag_offsets,
                           const float* __restrict__ values,
                           const float* __restrict__ x, float* __restrict__ y,
                           int rows, int cols, long long stride_elems, int num_diagonals) {
  const int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    float acc = 0.0f;
    for (int d = 0; d < num_diagonals; ++d) {
      const long long c = (long long)r + __ldg(diag_offsets + d);
      if (c >= 0 && c < cols)
        acc = __fadd_rn(acc, __fmul_rn(__ldg(values + (long long)d * stride_elems + r), __ldg(x + c)));
    }
    y[r] = acc;
  }
}

// ----------------------------------------------------- flat_uniform_occupancy --
// Replaces reference algorithms/spmv/flat_partitioned.cuh:46-60: one thread per
// window of K consecutive atoms (layout::flat_uniform_occupancy<K, csr>). The
// reference finds the row of EVERY atom with an upper_bound search and issues
// one atomic per atom; here the search runs once per window, the rows are then
// walked, and one atomic goes out per (window, row) run.
__global__ void __launch_bounds__(128)
    spmv_flat_partitioned(const int* __restrict__ offsets,
                          const int* __restrict__ indices,
                          const float* __restrict__ values,
                          const float* __restrict__ x, float* __restrict__ y,
                          int rows, int nnz, int K) {
  const long long windows = ((long long)nnz + K - 1) / K;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < windows; t += stride) {
    const int a0 = int(t * K);
    const int a1 = min(nnz, a0 + K);
    // row of atom a0: last row whose begin offset is <= a0 (base.tile_of)
    int lo = 0, hi = rows;   // invariant: offsets[lo] <= a0 < offsets[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(offsets + mid) <= a0) lo = mid; else hi = mid;
    }
    int row = lo;
    int row_end = __ldg(offsets + row + 1);
    float acc = 0.0f;
    for (int a = a0; a < a1; ++a) {
      while (a >= row_end) {            // close the run (empty rows are skipped)
        if (acc != 0.0f) atomicAdd(y + row, acc);
        acc = 0.0f;
        ++row;
        row_end = __ldg(offsets + row + 1);
      }
      acc = __fadd_rn(acc, __fmul_rn(__ldg(values + a), __ldg(x + __ldg(indices + a))));
    }
    if (acc != 0.0f) atomicAdd(y + row, acc);
  }
}

// -------------------------------------------------------------------- spmm --
// C[row, :] = sum_nz values[nz] * B[indices[nz], :]   (B, C row-major, n columns)
// Replaces reference algorithms/spmm/thread_mapped.cuh:28-53, whose one thread per
// row walks B column by column (stride-n reads). Here a warp owns one row and a
// slab of 32 columns: every stored entry costs one broadcast (value, index) load
// and ONE coalesced 128-byte read of B. Per (row, column) the adds run in
// ascending nz order with un-fused arithmetic, like the reference's inner loop.
__global__ void __launch_bounds__(128)
    spmm_csr_row_warp(const int* __restrict__ offsets, const int* __restrict__ indices,
                      const float* __restrict__ values, const float* __restrict__ B,
                      float* __restrict__ C, int rows, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  const int col = blockIdx.y * 32 + lane;
  if (row >= rows) return;
  const int begin = __ldg(offsets + row), end = __ldg(offsets + row + 1);
  float acc = 0.0f;
  for (int base = begin; base < end; base += 32) {
    // one coalesced read of up to 32 (index, value) pairs, then broadcast by shuffle
    const int mine = base + lane;
    int idx = 0;
    float val = 0.0f;
    if (mine < end) { idx = __ldg(indices + mine); val = __ldg(values + mine); }
    const int cnt = min(32, end - base);
    for (int k = 0; k < cnt; ++k) {
      const int i = __shfl_sync(0xffffffffu, idx, k);
      const float v = __shfl_sync(0xffffffffu, val, k);
      if (col < n) acc = __fadd_rn(acc, __fmul_rn(v, __ldg(B + (long long)i * n + col)));
    }
  }
  if (col < n) C[(long long)row * n + col] = acc;
}

// -------------------------------------------------------------------- bcsr --
template <int R, int C>
__global__ void __launch_bounds__(128)
    spmv_bcsr_thread_mapped(const int* __restrict__ block_offsets,
                            const int* __restrict__ block_cols,
                            const float* __restrict__ values,
                            const float* __restrict__ x, float* __restrict__ y,
                            int block_rows, int rows) {
  const int stride = gridDim.x * blockDim.x;
  for (int br = blockIdx.x * blockDim.x + threadIdx.x; br < block_rows;
       br += stride) {
    float acc[R];
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = 0.0f;
    const int stop = __ldg(block_offsets + br + 1);
    for (int b = __ldg(block_offsets + br); b < stop; ++b) {
      const long long bc = __ldg(block_cols + b);
      const float* blk = values + (long long)b * R * C;
      float xs[C];
#pragma unroll
      for (int j = 0; j < C; ++j) xs[j] = __ldg(x + bc * C + j);
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j)
          acc[i] = __fadd_rn(acc[i], __fmul_rn(__ldg(blk + i * C + j), xs[j]));
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const long long row = (long long)br * R + i;
      if (row < rows) y[row] = acc[i];
    }
  }
}

}  // namespace sk
}  // namespace loopsb
