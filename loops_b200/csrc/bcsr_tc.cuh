// loops_b200/csrc/bcsr_tc.cuh -- BCSR 4x4, bf16 values and x, fp32 accumulate:
// SpMV on the tcgen05 tensor cores (BASELINE.json config 4). Replaces, for this
// dtype, reference algorithms/spmv/bcsr_thread_mapped.cuh:36-74 (one thread per
// block-row doing 16 scalar FMAs per block).
//
// Mapping (a block really is a dense 4x4 contraction, a block-ROW is a
// contraction over all of its blocks):
//   * a GROUP is 32 block-rows; its 128 scalar rows are the M dimension of one
//     UMMA tile: m = 4*b + i  (b = block-row slot 0..31, i = row inside block);
//   * one K-step covers 4 consecutive blocks of every block-row of the group:
//     k = 4*slot + j (slot 0..3, j = column inside block), K = 16 (bf16 UMMA K);
//         A[m][k] = block(row_b, 4s + slot)[i][j]           (0 past the row end)
//         B[n][k] = x[4*bcol(row_n, 4s + slot) + j]          (n = block-row slot)
//     D[m][n] += sum_k A[m][k] B[n][k]; the wanted y values are the block
//     diagonal n == m/4, everything else is discarded (N = 32 is the price of
//     giving every block-row its own gathered x; tensor time stays far below
//     HBM time: 16 cycles per 128 blocks);
//   * the accumulator D (128 lanes x 32 columns fp32) lives in TMEM for the
//     whole K loop of the group and is read once (tcgen05.ld) at the end;
//   * block-rows are ordered by length (plan, data-independent of values) so
//     the 32 rows of a group need nearly the same number of K-steps; groups are
//     dealt longest-first to persistent CTAs.
// A and B tiles are staged in shared memory in the canonical K-major
// no-swizzle UMMA layout (8-row x 16-byte core matrices: LBO = 128 B between
// the two K halves, SBO = 256 B between 8-row groups), double-buffered; reuse
// is gated by tcgen05.commit -> mbarrier.
#pragma once

#include "common.cuh"

#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <loops/util/tma.hxx>

namespace loopsb {
namespace bcsr_tc {

constexpr int kThreads = 128;       // 32 block-row slots x 4 block slots
constexpr int kGroupRows = 32;
constexpr int kTmemCols = 32;       // N = 32 fp32 accumulator columns
constexpr int kATileBytes = 128 * 16 * 2;  // 4 KB
constexpr int kBTileBytes = 32 * 16 * 2;   // 1 KB

constexpr int kChunkSteps = 16;     // K-steps (of 4 blocks per row) per work item

struct plan_data {
  int num_block_rows = 0;
  int num_groups = 0;
  int* order = nullptr;       // block-row ids, longest first
  int* lengths = nullptr;     // sorted lengths (device)
  // Work items: a group's K loop cut into chunks of <= kChunkSteps steps so
  // that a 1024-block row does not serialise inside one CTA.
  int num_items = 0;
  int4* items = nullptr;      // {group, first step, past-last step, partial slot or -1}
  int4* item_rows = nullptr;  // [item][32 block-row slots] {first block, blocks, block-row id or -1, 0}
  int num_split = 0;
  int4* split = nullptr;      // {group, first partial slot, number of chunks, 0}
  float* partial = nullptr;   // [slots][128] fp32
  int sm_count = 0;
  int ctas_per_sm = 0;        // persistent grid = resident CTAs per SM x SMs
  long long bytes = 0;
  // Packed copy of the block values (pack()): one 4 KB UMMA A-tile image per K-step of
  // every work item, in the order the kernel consumes them, so a step's values arrive
  // as ONE bulk copy instead of 128 LSU sector requests.
  uint16_t* packed = nullptr;       // [total_tiles]{A-tile image 4 KB, 128 block columns}
  long long* item_tile = nullptr;   // [num_items] first tile of the item
  long long total_tiles = 0;
  const void* packed_key = nullptr; // the values pointer the copy was made from
  int* work_counter = nullptr;      // two counters handing out work items (packed kernel: dynamic, longest first); launch n
                                    // uses counter n & 1 and re-arms the other one for launch n + 1 (no memset per launch)
  int launch_parity = 0;
  bool packed_launched = false;     // a first launch on the current packed copy has been enqueued (later ones may use PDL)
};

__global__ void row_lengths_kernel(const int* __restrict__ off, int n,
                                   int* __restrict__ len, int* __restrict__ id) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) { len[r] = off[r + 1] - off[r]; id[r] = r; }
}

// One int4 per (work item, block-row slot): what a thread needs to start on an
// item, in ONE load whose address depends only on the item number -- so it can be
// requested an item ahead instead of walking items -> order -> offsets.
__global__ void item_rows_kernel(const int4* __restrict__ items, int num_items, const int* __restrict__ order,
                                 const int* __restrict__ off, int num_block_rows, int4* __restrict__ out) {
  const int it = blockIdx.x, b = threadIdx.x;
  if (it >= num_items) return;
  const int slot_row = items[it].x * 32 + b;
  int4 v = make_int4(0, 0, -1, 0);
  if (slot_row < num_block_rows) {
    const int r = order[slot_row];
    v.x = off[r];
    v.y = off[r + 1] - v.x;
    v.z = r;
  }
  out[(long long)it * 32 + b] = v;
}

inline void destroy(plan_data* p) {
  if (!p) return;
  if (p->order) cudaFree(p->order);
  if (p->lengths) cudaFree(p->lengths);
  if (p->items) cudaFree(p->items);
  if (p->item_rows) cudaFree(p->item_rows);
  if (p->split) cudaFree(p->split);
  if (p->partial) cudaFree(p->partial);
  if (p->packed) cudaFree(p->packed);
  if (p->item_tile) cudaFree(p->item_tile);
  if (p->work_counter) cudaFree(p->work_counter);
  delete p;
}

inline long long workspace_bytes(const plan_data* p) { return p ? p->bytes : 0; }

inline int create(plan_data** out, const loopsb_layout_t* lay, int sm_count,
                  cudaStream_t stream) {
  *out = nullptr;
  plan_data* p = new plan_data();
  p->num_block_rows = lay->num_tiles;
  p->num_groups = (lay->num_tiles + kGroupRows - 1) / kGroupRows;
  p->sm_count = sm_count;
  const int n = lay->num_tiles;
  if (n == 0) { *out = p; return LOOPSB_OK; }
  int *len_in = nullptr, *id_in = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  auto fail = [&](const char* what) {
    set_error("bcsr tensor-core plan: %s failed: %s", what, cudaGetErrorString(cudaGetLastError()));
    cudaFree(len_in); cudaFree(id_in); cudaFree(tmp);
    destroy(p);
    return LOOPSB_ERR_CUDA;
  };
  if (cudaMalloc(&len_in, size_t(n) * 4) != cudaSuccess || cudaMalloc(&id_in, size_t(n) * 4) != cudaSuccess ||
      cudaMalloc(&p->order, size_t(n) * 4) != cudaSuccess || cudaMalloc(&p->lengths, size_t(n) * 4) != cudaSuccess)
    return fail("cudaMalloc");
  row_lengths_kernel<<<(n + 255) / 256, 256, 0, stream>>>(lay->offsets, n, len_in, id_in);
  if (cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, len_in, p->lengths, id_in, p->order, n,
                                                0, 32, stream) != cudaSuccess)
    return fail("cub sizing");
  if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16) != cudaSuccess) return fail("cudaMalloc(tmp)");
  if (cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, len_in, p->lengths, id_in, p->order, n, 0, 32,
                                                stream) != cudaSuccess)
    return fail("cub sort");
  if (cudaStreamSynchronize(stream) != cudaSuccess) return fail("sync");
  cudaFree(len_in); cudaFree(id_in); cudaFree(tmp);
  len_in = id_in = nullptr; tmp = nullptr;
  // Work-item tables (host): only the longest row of every group matters.
  std::vector<int> lens(static_cast<size_t>(n), 0);
  if (cudaMemcpy(lens.data(), p->lengths, size_t(n) * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail("cudaMemcpy(lengths)");
  std::vector<int4> items, split;
  int slots = 0;
  for (int g = 0; g < p->num_groups; ++g) {
    const int steps = (lens[size_t(g) * kGroupRows] + 3) >> 2;
    const int nch = steps > kChunkSteps ? (steps + kChunkSteps - 1) / kChunkSteps : 1;
    if (nch == 1) {
      items.push_back(make_int4(g, 0, steps, -1));
    } else {
      split.push_back(make_int4(g, slots, nch, 0));
      for (int c = 0; c < nch; ++c) {
        const int s0 = c * kChunkSteps;
        const int s1 = s0 + kChunkSteps < steps ? s0 + kChunkSteps : steps;
        items.push_back(make_int4(g, s0, s1, slots++));
      }
    }
  }
  p->num_items = int(items.size());
  p->num_split = int(split.size());
  if (cudaMalloc(&p->items, items.size() * sizeof(int4)) != cudaSuccess) return fail("cudaMalloc(items)");
  if (cudaMemcpy(p->items, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail("cudaMemcpy(items)");
  if (cudaMalloc(&p->item_rows, (items.size() + 1) * 32 * sizeof(int4)) != cudaSuccess) return fail("cudaMalloc(item_rows)");
  item_rows_kernel<<<p->num_items, 32, 0, stream>>>(p->items, p->num_items, p->order, lay->offsets, n, p->item_rows);
  if (cudaStreamSynchronize(stream) != cudaSuccess) return fail("item_rows_kernel");
  if (!split.empty()) {
    if (cudaMalloc(&p->split, split.size() * sizeof(int4)) != cudaSuccess ||
        cudaMalloc(&p->partial, size_t(slots) * 128 * sizeof(float)) != cudaSuccess)
      return fail("cudaMalloc(split)");
    if (cudaMemcpy(p->split, split.data(), split.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess)
      return fail("cudaMemcpy(split)");
  }
  p->bytes = (long long)n * 8 + (long long)items.size() * (16 + 512) + (long long)split.size() * 16 + (long long)slots * 512;
  *out = p;
  return LOOPSB_OK;
}

// ---- tcgen05 / TMEM wrappers (inline PTX) ---------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   loops::tma::smem_addr(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   loops::tma::smem_addr(bar)) : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns of this warp's TMEM quadrant.
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, no swizzle: 8x16B core matrices, LBO between K halves, SBO between
// 8-row groups (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t make_smem_desc(const void* tile, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint64_t addr = loops::tma::smem_addr(tile);
  return ((addr >> 4) & 0x3FFFull) | (uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t tile_offset(int row, int slot) {
  // byte offset of the 8-byte piece holding k = 4*slot .. 4*slot+3 of `row`
  return uint32_t((row >> 3) * 256 + (row & 7) * 16 + (slot >> 1) * 128 + (slot & 1) * 8);
}

struct __align__(1024) tc_shared {
  unsigned char a[2][kATileBytes];
  unsigned char b[2][kBTileBytes];
  unsigned long long mma_done[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads)
    spmv_bcsr4x4_bf16_kernel(const int* __restrict__ block_offsets, const int* __restrict__ block_cols,
                             const uint16_t* __restrict__ values, const uint16_t* __restrict__ x,
                             float* __restrict__ y, const int* __restrict__ order, int num_block_rows,
                             const int4* __restrict__ items, int num_items, float* __restrict__ partial,
                             int num_rows, const int4* __restrict__ item_rows) {
  __shared__ tc_shared sm;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int b = tid >> 2;      // block-row slot inside the group
  const int slot = tid & 3;    // block slot inside the K-step

  if (warp == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
  if (tid == 0) {
    loops::tma::barrier_init(reinterpret_cast<uint64_t*>(&sm.mma_done[0]), 1);
    loops::tma::barrier_init(reinterpret_cast<uint64_t*>(&sm.mma_done[1]), 1);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  constexpr uint32_t idesc = make_idesc(128, 32);

  uint32_t t = 0;  // K-steps staged so far by this CTA: step t uses buffer t & 1, whose previous use was step t - 2

  for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
    // Both descriptors of the NEXT item go to L2 now (their addresses depend only on
    // the item number), so the two loads below hit L2 one item later. (Carrying
    // them in registers was tried: the item fields are warp-uniform, ptxas moves
    // them to uniform registers right behind the load and the prefetch turns
    // synchronous.)
    if (it + int(gridDim.x) < num_items && lane == 0) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(items + it + gridDim.x));
      if (warp < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(item_rows + (long long)(it + gridDim.x) * 32 + 8 * warp));
    }
    const int4 item = __ldg(items + it);
    const int4 row = __ldg(item_rows + (long long)it * 32 + b);
    const int start = row.x, len = row.y;
    // rows are sorted by length; the plan cut the longest row's K loop into
    // [item.y, item.z)
    const int steps = item.z - item.y;

    // Two thread->data maps per K-step (tid = 4*b + slot):
    //  * B tile (x slices) and block columns: thread (b, slot) owns block 4s+slot of
    //    block-row slot b -- its 8-byte x slice lands conflict-free in the B tile;
    //  * A tile (values): in round j = 0..3 a warp moves the blocks of TWO block-row
    //    slots (8w + 2j + lane/16), lane -> (slot_a = (lane/4)%4, row i = lane%4):
    //    one coalesced 256-byte load (32 lanes x 8 B, every sector fully used) and one
    //    conflict-free 256-byte shared store into the canonical UMMA layout per
    //    round. (One thread per whole block -- 2 x LDG.128 and 4 strided 8-byte
    //    stores -- cost 8-way bank conflicts and half-used sectors.)
    // Three register stages, two K-steps of look-ahead: while step s is staged and
    // multiplied, the x slice of step s+1 is in flight (its block column arrived one
    // step ago) and the column + values of step s+2 are requested.
    const int slot_a = (lane >> 2) & 3, row_i = lane & 3;
    int a_start[4], a_len[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int src_lane = 4 * (2 * j + (lane >> 4));   // a lane whose b is the block-row slot of round j
      a_start[j] = __shfl_sync(0xffffffffu, start, src_lane);
      a_len[j] = __shfl_sync(0xffffffffu, len, src_lane);
    }
    struct stage_regs { uint2 va[4]; uint2 xs; int bc; bool valid; };
    stage_regs st[3];
    auto load_cv = [&](stage_regs& r, int step) {
      const int k = 4 * step + slot;
      r.valid = k < len;
      r.xs = make_uint2(0, 0); r.bc = 0;
      if (r.valid) r.bc = __ldg(block_cols + (long long)start + k);
      const int ka = 4 * step + slot_a;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r.va[j] = make_uint2(0, 0);
        if (ka < a_len[j])
          r.va[j] = __ldg(reinterpret_cast<const uint2*>(values + ((long long)a_start[j] + ka) * 16 + row_i * 4));
      }
    };
    auto load_x = [&](stage_regs& r) {
      if (r.valid) r.xs = __ldg(reinterpret_cast<const uint2*>(x + (long long)r.bc * 4));
    };
    auto process = [&](const stage_regs& r, int s) {
      const uint32_t buf = t & 1u;
      // the MMA that last read this buffer (step t - 2, its (t/2 - 1)-th use) must have completed
      if (t >= 2u)
        loops::tma::barrier_wait(reinterpret_cast<uint64_t*>(&sm.mma_done[buf]), ((t >> 1) - 1u) & 1u);
      unsigned char* A = sm.a[buf];
      unsigned char* B = sm.b[buf];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ba = 8 * warp + 2 * j + (lane >> 4);       // block-row slot moved by this lane in round j
        *reinterpret_cast<uint2*>(A + tile_offset(4 * ba + row_i, slot_a)) = r.va[j];
      }
      *reinterpret_cast<uint2*>(B + tile_offset(b, slot)) = r.xs;
      loops::tma::fence_proxy_async();   // generic-proxy smem writes -> tensor-core (async) proxy
      tc_fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after_sync();
        umma_bf16(tmem, make_smem_desc(A, 128, 256), make_smem_desc(B, 128, 256), idesc, s > item.y ? 1u : 0u);
        umma_commit(reinterpret_cast<uint64_t*>(&sm.mma_done[buf]));
      }
      ++t;
    };
    if (steps > 0) {
      int s = item.y;
      load_cv(st[0], s);
      if (s + 1 < item.z) load_cv(st[1], s + 1);
      load_x(st[0]);
#define LOOPSB_BCSR_STEP(A_, B_, C_)                       \
      if (s + 1 < item.z) load_x(st[B_]);                  \
      if (s + 2 < item.z) load_cv(st[C_], s + 2);          \
      process(st[A_], s);
      while (true) {
        LOOPSB_BCSR_STEP(0, 1, 2) if (++s >= item.z) break;
        LOOPSB_BCSR_STEP(1, 2, 0) if (++s >= item.z) break;
        LOOPSB_BCSR_STEP(2, 0, 1) if (++s >= item.z) break;
      }
#undef LOOPSB_BCSR_STEP
    }

    if (steps > 0) {
      // accumulator complete when the last commit lands
      const uint32_t last = (t - 1u) & 1u;
      loops::tma::barrier_wait(reinterpret_cast<uint64_t*>(&sm.mma_done[last]), ((t - 1u) >> 1) & 1u);
      // the other buffer's commit (if any) was issued earlier, so it has landed too
      tc_fence_after_sync();
      // lane l of warp w holds scalar row m = 32w + l; its block-row slot is
      // m/4 = 8w + l/4, i.e. column (l/4) of the 8-column window at 8w.
      uint32_t acc[8];
      tmem_ld_32x32b_x8(tmem + (uint32_t(32 * warp) << 16) + uint32_t(8 * warp), acc);
      const int m = 32 * warp + lane;
      float out = 0.0f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if ((lane >> 2) == q) out = __uint_as_float(acc[q]);
      if (item.w >= 0) {
        partial[(long long)item.w * 128 + m] = out;   // this chunk's share, reduced in order later
      } else {
        // scalar row m belongs to block-row slot m/4 == b, the slot whose descriptor this thread holds
        if (row.z >= 0) {
          const long long yrow = (long long)row.z * 4 + (m & 3);
          if (yrow < num_rows) y[yrow] = out;
        }
      }
      tc_fence_before_sync();
    } else {
      // every block-row of the group is empty
      const int m = 32 * warp + lane;
      if (row.z >= 0) {
        const long long yrow = (long long)row.z * 4 + (m & 3);
        if (yrow < num_rows) y[yrow] = 0.0f;
      }
    }
    __syncthreads();   // TMEM and both buffers are free for the next group
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------
// Packed variant: the A tiles come from the plan's packed copy by TMA bulk copy.
// ---------------------------------------------------------------------------
constexpr int kPackStages = 5;   // tiles in flight per CTA (23 KB); x slices are gathered two steps ahead
constexpr int kPackBStages = 4;  // B tiles: steps s, s+1, s+2 being filled / read, s-1 still under its MMA
constexpr int kPackTileBytes = kATileBytes + kThreads * 4;   // A-tile image + the step's 128 block columns (-1 = none)

// grid = work items, 128 threads (tid = 4*b + slot as in the SpMV kernel): writes the
// item's A-tile images, zero where a block-row has fewer blocks than the step needs.
__global__ void __launch_bounds__(kThreads)
    bcsr_pack_kernel(const int4* __restrict__ items, const int4* __restrict__ item_rows,
                     const long long* __restrict__ item_tile, const uint16_t* __restrict__ values,
                     const int* __restrict__ block_cols, uint16_t* __restrict__ packed) {
  const int it = blockIdx.x, tid = threadIdx.x, b = tid >> 2, slot = tid & 3;
  const int4 item = items[it];
  const int4 row = item_rows[(long long)it * 32 + b];
  unsigned char* base = reinterpret_cast<unsigned char*>(packed) + size_t(item_tile[it]) * kPackTileBytes;
  for (int s = item.y; s < item.z; ++s) {
    const int k = 4 * s + slot;
    uint2 v[4] = {make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0)};
    int bc = -1;
    if (k < row.y) {
      bc = __ldg(block_cols + (long long)row.x + k);
      const uint4* src = reinterpret_cast<const uint4*>(values + ((long long)row.x + k) * 16);
      const uint4 lo = __ldg(src), hi = __ldg(src + 1);
      v[0] = make_uint2(lo.x, lo.y); v[1] = make_uint2(lo.z, lo.w);
      v[2] = make_uint2(hi.x, hi.y); v[3] = make_uint2(hi.z, hi.w);
    }
    unsigned char* tile = base + size_t(s - item.y) * kPackTileBytes;
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint2*>(tile + tile_offset(4 * b + i, slot)) = v[i];
    reinterpret_cast<int*>(tile + kATileBytes)[tid] = bc;
  }
}

struct __align__(128) packed_stage {
  unsigned char a[kATileBytes];
  int cols[kThreads];
};
struct __align__(128) tc_shared_packed {
  packed_stage st[kPackStages];               // filled by one bulk copy of kPackTileBytes
  unsigned char b[kPackBStages][kBTileBytes];
  unsigned long long full[kPackStages];       // the stage's A tile has landed (TMA complete_tx)
  unsigned long long mma_done[kPackStages];   // the MMA that read the stage has completed
  uint32_t tmem_base;
  int next_item;                              // handed out by the work counter, one item ahead
};

__global__ void __launch_bounds__(kThreads)
    spmv_bcsr4x4_bf16_packed_kernel(const uint16_t* __restrict__ packed, const long long* __restrict__ item_tile,
                                    const uint16_t* __restrict__ x, float* __restrict__ y,
                                    const int4* __restrict__ items, int num_items, float* __restrict__ partial,
                                    int num_rows, const int4* __restrict__ item_rows, int* __restrict__ counters,
                                    int parity, int pdl) {
  __shared__ tc_shared_packed sm;
  int* const work_counter = counters + parity;
  // pdl: launched with programmatic stream serialization -- TMEM, barriers and the first A tiles of the
  // CTA's first item (plan-owned constants) are set up while the previous kernel of the stream drains;
  // x, y, the partial rows and the work counters are only touched after griddepcontrol.wait.
  if (pdl) asm volatile("griddepcontrol.launch_dependents;");
  bool ordered = false;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int b = tid >> 2;      // block-row slot inside the group
  const int slot = tid & 3;    // block slot inside the K-step
  constexpr uint32_t NS = kPackStages, NB = kPackBStages;

  if (warp == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
  if (tid == 0) {
    for (int k = 0; k < kPackStages; ++k) {
      loops::tma::barrier_init(reinterpret_cast<uint64_t*>(&sm.full[k]), 1);
      loops::tma::barrier_init(reinterpret_cast<uint64_t*>(&sm.mma_done[k]), 1);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  constexpr uint32_t idesc = make_idesc(128, 32);
  const unsigned char* packed_bytes = reinterpret_cast<const unsigned char*>(packed);

  uint32_t t = 0;  // K-steps staged so far by this CTA: step t lives in stage t % NS (its (t / NS)-th use)

  // The x slice of step tt goes from global memory straight into its B tile with an
  // 8-byte cp.async (zero-filled when the slot holds no block). Nothing lands in
  // registers, so the generic->async proxy fence of an earlier step does not have to
  // wait for the gathers of later steps (with register-staged slices it did: ncu showed
  // the fence stalled on the long scoreboard of the look-ahead loads).
  auto issue_x = [&](uint32_t tt) {
    const uint32_t st = tt % NS, bst = tt % NB;
    loops::tma::barrier_wait(reinterpret_cast<uint64_t*>(&sm.full[st]), (tt / NS) & 1u);   // the tile (and its columns) is in
    if (tt >= NB) {   // the MMA that last read this B tile (step tt - NB) has completed
      const uint32_t pt = tt - NB;
      loops::tma::barrier_wait(reinterpret_cast<uint64_t*>(&sm.mma_done[pt % NS]), (pt / NS) & 1u);
    }
    const int bc = sm.st[st].cols[tid];
    const uint32_t dst = loops::tma::smem_addr(sm.b[bst] + tile_offset(b, slot));
    const uint16_t* src = x + (long long)(bc >= 0 ? bc : 0) * 4;
    const uint32_t nbytes = bc >= 0 ? 8u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  };
  auto commit_x = []() { asm volatile("cp.async.commit_group;" ::: "memory"); };

  // Work items are sorted longest first and handed out dynamically (a static deal leaves
  // the busiest CTA with ~20 % more K-steps than the average): the first item is the
  // CTA's own index, every further one comes from the counter, fetched one item ahead.
  for (int it = blockIdx.x; it < num_items;) {
    const int4 item = __ldg(items + it);
    const int4 row = __ldg(item_rows + (long long)it * 32 + b);
    const int steps = item.z - item.y;
    const unsigned char* tiles = packed_bytes + size_t(__ldg(item_tile + it)) * kPackTileBytes;

    // every MMA of the previous item has completed (its epilogue waited for the last
    // commit), so all stages are free: request the first NS tiles of this item
    if (tid == 0) {
      for (int j = 0; j < int(NS) && j < steps; ++j) {
        const uint32_t st = (t + uint32_t(j)) % NS;
        uint64_t* bar = reinterpret_cast<uint64_t*>(&sm.full[st]);
        loops::tma::barrier_arrive_expect_tx(bar, kPackTileBytes);
        loops::tma::bulk_g2s(&sm.st[st], tiles + size_t(j) * kPackTileBytes, kPackTileBytes, bar);
      }
    }

    if (!ordered) {
      if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
      if (blockIdx.x == 0 && tid == 0) counters[parity ^ 1] = 0;   // the previous launch (its last user) is complete
      ordered = true;
    }
    // the first tiles of this CTA's NEXT item go to L2 now, so that its start pays an L2
    // hit instead of an HBM round trip (one bulk prefetch per tile, issued by one thread)
    if (tid == 32) {
      const int nxt = int(gridDim.x) + atomicAdd(work_counter, 1);
      sm.next_item = nxt;   // read by everybody after the item's last barrier
      if (nxt < num_items) {
        const int4 nitem = __ldg(items + nxt);
        const unsigned char* ntiles = packed_bytes + size_t(__ldg(item_tile + nxt)) * kPackTileBytes;
        const int n = min(3, nitem.z - nitem.y);
        for (int j = 0; j < n; ++j)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ntiles + size_t(j) * kPackTileBytes),
                       "r"(uint32_t(kPackTileBytes)) : "memory");
        asm volatile("prefetch.global.L2 [%0];" ::"l"(item_rows + (long long)nxt * 32));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(item_rows + (long long)nxt * 32 + 8));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(item_rows + (long long)nxt * 32 + 16));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(item_rows + (long long)nxt * 32 + 24));
      }
    }

    if (steps > 0) {
      // one cp.async group per K-step, two steps of look-ahead
      issue_x(t); commit_x();
      if (item.y + 1 < item.z) issue_x(t + 1u);
      commit_x();
      for (int s = item.y; s < item.z; ++s) {
        if (s + 2 < item.z) issue_x(t + 2u);
        commit_x();
        asm volatile("cp.async.wait_group 2;" ::: "memory");   // this step's slice has landed
        const uint32_t st = t % NS, bst = t % NB;
        loops::tma::fence_proxy_async();   // generic-proxy smem writes -> tensor-core (async) proxy
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after_sync();
          umma_bf16(tmem, make_smem_desc(sm.st[st].a, 128, 256), make_smem_desc(sm.b[bst], 128, 256), idesc,
                    s > item.y ? 1u : 0u);
          umma_commit(reinterpret_cast<uint64_t*>(&sm.mma_done[st]));
          // refill the stage of the PREVIOUS step (its MMA was committed one iteration ago)
          // with the tile NS - 1 steps ahead of this one
          const int ahead = s - 1 + int(NS);
          if (s > item.y && ahead < item.z) {
            const uint32_t pt = t - 1u, pst = pt % NS, puse = pt / NS;
            loops::tma::barrier_wait(reinterpret_cast<uint64_t*>(&sm.mma_done[pst]), puse & 1u);
            uint64_t* bar = reinterpret_cast<uint64_t*>(&sm.full[pst]);
            loops::tma::barrier_arrive_expect_tx(bar, kPackTileBytes);
            loops::tma::bulk_g2s(&sm.st[pst], tiles + size_t(ahead - item.y) * kPackTileBytes, kPackTileBytes, bar);
          }
        }
        ++t;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }

    if (steps > 0) {
      // accumulator complete when the last commit lands (commits complete in issue order)
      const uint32_t lt = t - 1u;
      loops::tma::barrier_wait(reinterpret_cast<uint64_t*>(&sm.mma_done[lt % NS]), (lt / NS) & 1u);
      tc_fence_after_sync();
      uint32_t acc[8];
      tmem_ld_32x32b_x8(tmem + (uint32_t(32 * warp) << 16) + uint32_t(8 * warp), acc);
      const int m = 32 * warp + lane;
      const int q = lane >> 2;
      uint32_t o = acc[0];
      o = q == 1 ? acc[1] : o; o = q == 2 ? acc[2] : o; o = q == 3 ? acc[3] : o; o = q == 4 ? acc[4] : o;
      o = q == 5 ? acc[5] : o; o = q == 6 ? acc[6] : o; o = q == 7 ? acc[7] : o;
      const float out = __uint_as_float(o);
      if (item.w >= 0) {
        partial[(long long)item.w * 128 + m] = out;
      } else if (row.z >= 0) {
        const long long yrow = (long long)row.z * 4 + (m & 3);
        if (yrow < num_rows) y[yrow] = out;
      }
      tc_fence_before_sync();
    } else {
      const int m = 32 * warp + lane;
      if (row.z >= 0) {
        const long long yrow = (long long)row.z * 4 + (m & 3);
        if (yrow < num_rows) y[yrow] = 0.0f;
      }
    }
    __syncthreads();   // TMEM and all stages are free for the next item
    it = sm.next_item;
    __syncthreads();   // everybody has read it before thread 32 overwrites it
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// Build (or rebuild) the packed copy from `values`. One pass over the blocks.
// Forget the packed copy (the caller changed the block values in place): SpMV calls take the
// direct kernel, which reads the live arrays, until pack() is called again.
inline void unpack(plan_data* p) {
  if (p) p->packed_key = nullptr;
}

inline int pack(plan_data* p, const uint16_t* values, const int* block_cols, cudaStream_t stream) {
  LOOPSB_REQUIRE(p != nullptr, "null plan");
  if (p->num_items == 0) { p->packed_key = values; return LOOPSB_OK; }
  LOOPSB_REQUIRE(values != nullptr && block_cols != nullptr && (reinterpret_cast<uintptr_t>(values) & 15u) == 0,
                 "values must be 16-byte aligned");
  if (!p->item_tile) {
    std::vector<int4> items(size_t(p->num_items));
    LOOPSB_CUDA_TRY(cudaMemcpy(items.data(), p->items, items.size() * sizeof(int4), cudaMemcpyDeviceToHost));
    std::vector<long long> first(items.size());
    long long total = 0;
    for (size_t i = 0; i < items.size(); ++i) { first[i] = total; total += items[i].z - items[i].y; }
    p->total_tiles = total;
    LOOPSB_CUDA_TRY(cudaMalloc(&p->item_tile, first.size() * sizeof(long long)));
    LOOPSB_CUDA_TRY(cudaMemcpy(p->item_tile, first.data(), first.size() * sizeof(long long), cudaMemcpyHostToDevice));
    LOOPSB_CUDA_TRY(cudaMalloc(&p->packed, size_t(total > 0 ? total : 1) * kPackTileBytes));
    p->bytes += (long long)first.size() * 8 + total * kPackTileBytes;
  }
  bcsr_pack_kernel<<<p->num_items, kThreads, 0, stream>>>(p->items, p->item_rows, p->item_tile, values, block_cols,
                                                          p->packed);
  LOOPSB_CUDA_TRY(cudaGetLastError());
  p->packed_key = values;
  p->packed_launched = false;   // the next launch must be fully ordered behind the pack kernel
  return LOOPSB_OK;
}

// y[row] = sum over the chunks of a split group, in chunk order.
__global__ void __launch_bounds__(128)
    bcsr_split_reduce_kernel(const int4* __restrict__ split, const float* __restrict__ partial,
                             const int* __restrict__ order, int num_block_rows, int num_rows,
                             float* __restrict__ y) {
  asm volatile("griddepcontrol.launch_dependents;");   // a PDL-launched SpMV behind this kernel may start its set-up
  const int4 s = __ldg(split + blockIdx.x);
  const int m = threadIdx.x;
  float acc = 0.0f;
  for (int c = 0; c < s.z; ++c) acc = __fadd_rn(acc, partial[(long long)(s.y + c) * 128 + m]);
  const int rr_slot = s.x * kGroupRows + (m >> 2);
  if (rr_slot < num_block_rows) {
    const long long row = (long long)__ldg(order + rr_slot) * 4 + (m & 3);
    if (row < num_rows) y[row] = acc;
  }
}

inline int run(plan_data* p, const loopsb_layout_t* lay, const uint16_t* values, const int32_t* block_cols,
               const uint16_t* x, float* y, int32_t num_rows, cudaStream_t stream) {
  LOOPSB_REQUIRE(p != nullptr && lay != nullptr, "null plan");
  LOOPSB_REQUIRE(num_rows >= 0, "negative rows");
  if (num_rows == 0 || p->num_block_rows == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(y != nullptr, "y is null");
  LOOPSB_REQUIRE(lay->num_atoms == 0 || (values && block_cols && x), "null matrix / x pointer");
  LOOPSB_REQUIRE((reinterpret_cast<uintptr_t>(values) & 15u) == 0 && (reinterpret_cast<uintptr_t>(x) & 7u) == 0,
                 "values must be 16-byte and x 8-byte aligned");
  if (p->ctas_per_sm == 0) {
    // persistent grid: 8 CTAs per SM measured best on B200 for the direct kernel (tools/bcsr_bench.py: 6 -> 137 us,
    // 8 -> 130 us, 10 -> 147 us, 12 -> 139 us)
    // (cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for this kernel on
    // B200 -- it is not used; 56 registers x 128 threads and 11 KB leave room for 9.)
    int per_sm = 8;
    if (const char* e = getenv("LOOPSB_BCSR_CTAS")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
    p->ctas_per_sm = per_sm;
  }
  const bool use_packed = p->packed && p->packed_key == values;
  // the packed kernel holds 28 KB of shared memory per CTA: 7 fit on an SM
  int grid = p->sm_count * (use_packed && !getenv("LOOPSB_BCSR_CTAS") ? 7 : p->ctas_per_sm);
  if (grid > p->num_items) grid = p->num_items;
  if (use_packed) {
    static bool carved = false;   // 22 KB of static shared memory per CTA: ask for the large carve-out once
    if (!carved) {
      cudaFuncSetAttribute(spmv_bcsr4x4_bf16_packed_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
      carved = true;
    }
    if (!p->work_counter) {
      LOOPSB_CUDA_TRY(cudaMalloc(&p->work_counter, 8));
      LOOPSB_CUDA_TRY(cudaMemsetAsync(p->work_counter, 0, 8, stream));   // once; afterwards launch n re-arms n + 1's counter
      p->launch_parity = 0;
    }
    // LOOPSB_BCSR_PDL (default 1): launches after the first on a packed copy are programmatic dependent
    // launches (see the kernel); a kernel of another kind in front of it simply never triggers early.
    static const bool pdl_env = !getenv("LOOPSB_BCSR_PDL") || atoi(getenv("LOOPSB_BCSR_PDL")) != 0;
    const int pdl = pdl_env && p->packed_launched ? 1 : 0;
    const int parity = p->launch_parity;
    p->launch_parity ^= 1;
    p->packed_launched = true;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    LOOPSB_CUDA_TRY(cudaLaunchKernelEx(&cfg, spmv_bcsr4x4_bf16_packed_kernel, (const uint16_t*)p->packed,
                                       (const long long*)p->item_tile, x, y, (const int4*)p->items, p->num_items,
                                       p->partial, num_rows, (const int4*)p->item_rows, p->work_counter, parity, pdl));
  } else
    spmv_bcsr4x4_bf16_kernel<<<grid, kThreads, 0, stream>>>(lay->offsets, block_cols, values, x, y, p->order,
                                                          p->num_block_rows, p->items, p->num_items, p->partial,
                                                          num_rows, p->item_rows);
  LOOPSB_CUDA_TRY(cudaGetLastError());
  if (p->num_split > 0) {
    bcsr_split_reduce_kernel<<<p->num_split, 128, 0, stream>>>(p->split, p->partial, p->order, p->num_block_rows,
                                                             num_rows, y);
    LOOPSB_CUDA_TRY(cudaGetLastError());
  }
  return LOOPSB_OK;
}

}  // namespace bcsr_tc
}  // namespace loopsb
