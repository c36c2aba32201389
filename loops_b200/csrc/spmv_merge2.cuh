// loops_b200/csrc/spmv_merge2.cuh -- merge_path_flat SpMV for sm_100a, second
// generation ("flag kernel"). Replaces reference
// algorithms/spmv/merge_path_flat.cuh:38-83 and ell_merge_path.cuh:32-69 on the
// CSR / ELL arrays themselves (no plan-owned copy of the matrix).
//
// What bounds this kernel (profiles/microbench_r01_partial.txt, ncu_merge_path_r0*.txt):
// every x[col] with a random col is its own 32-byte sector request and an SM's
// L1 issues ~1 such request per clock, so a CTA design is judged by how close it
// keeps that path to 100 % busy. The first-generation kernel (spmv_merge.cuh)
// reached 0.67 of that floor: its gathers were in flight for only part of each
// tile's life (park products -> barrier -> per-thread diagonal search + merge
// walk -> scan -> barrier), and it spent ~80 thread instructions and 0.37
// shared-memory wavefronts per nonzero on the walk.
//
// Same partition as the reference (merge tiles of 128*8 = 1024 items, start
// coordinates S(b*1024) from merge_coords_kernel), different execution:
//   * persistent CTAs, tiles dealt round-robin; one thread pulls a tile's three
//     contiguous streams (row-end window, column ids, values) into a 2-deep
//     shared-memory stage with 1-D TMA bulk copies (cp.async.bulk, mbarrier
//     completion), two tiles ahead of their use;
//   * SOFTWARE PIPELINE: in iteration n a thread first issues the 8 x-gathers of
//     tile n+1 (128-bit shared loads of its column ids / values), THEN reduces
//     tile n, whose gathers were issued one iteration earlier -- every thread
//     always has 8..16 gathers in flight, across the CTA barrier and the
//     reduction (two ping-pong register sets, loop unrolled by two so no
//     register moves wait on the loads);
//   * row structure without searches: the <= 1024 row ends of the window become
//     one 16-bit mark per atom position ("row i of the window ends after this
//     atom"), written by one thread per row end; rows with no atoms in the tile
//     are stored as 0 right there;
//   * reduction in registers: a thread owns 4 consecutive atoms per chunk
//     (exactly the 16 bytes it loaded; one 8-byte shared load brings their marks),
//     sums them left to right between marks, stores rows that lie inside its
//     chunk, and a ballot + 5-step warp-shuffle segmented scan joins rows that
//     span threads; a warp owns 256 consecutive atoms, chains its two chunk
//     slots in registers, and the per-warp aggregates are combined through one
//     16-byte shared word after the single CTA barrier;
//   * both pipes count: an SM's L1 serves ONE wavefront per clock, a random
//     gather or a 128-byte shared-memory access alike (ncu: gather sectors +
//     shared wavefronts = 0.9 of the elapsed cycles in both generations), so the
//     design spends ~0.1 shared wavefronts per nonzero where the first
//     generation spent 0.37;
//   * the partial row a tile ends in goes out as one (row, value) carry, folded
//     in by spmv_merge_fixup_kernel (spmv_merge.cuh) -- no atomics, y fully
//     overwritten, bit-stable run to run.
// Products are rounded to fp32 before the adds (__fmul_rn / __fadd_rn), so every
// term equals the reference's `values[nz] * x[indices[nz]]`.
//
// Requires 16-byte aligned `indices` and `values` base pointers (then the two
// streams of a tile share one skew = s.y & 3 and the staged 16-byte chunks are
// whole); other pointers take the first-generation kernel.
#pragma once

#include "common.cuh"
#include "spmv_merge.cuh"

#include <loops/util/tma.hxx>

namespace loopsb {
namespace mp2 {

constexpr int kTile = mp::kRefItemsPerMergeTile;   // 1024 items per merge tile

// STAGES: depth of the bulk-copy stage. A tile's staged streams are consumed the
// moment they are first looked at (the early part of the pipeline), so ONE buffer
// re-filled right after the CTA barrier has a whole iteration (~4 us with 6 CTAs
// sharing an SM) to land; the 8 KB it saves per CTA go to the L1, whose lines are
// what outstanding gather misses are parked in.
template <int THREADS, int STAGES = 1>
struct shared_t {
  static constexpr int kStageInts = 2 * kTile + 32;
  static constexpr int kWarps = THREADS / 32;
  static constexpr int kChunks = kTile / (4 * THREADS);   // 16-byte chunks per thread
  alignas(16) int stage[STAGES][kStageInts];
  // mark[pos] = 1 + window-relative id of the row whose last atom sits at staged
  // position pos (0 = no row ends here); read 4 at a time, cleared by the reader
  alignas(16) unsigned short mark[2][kTile + 8];
  alignas(16) float spill[2][4];        // products of the 257th chunk (na + skew > 1024)
  alignas(16) float wt_val[2][kWarps];  // per-warp aggregate: sum since the last closed row
  alignas(4) unsigned char wt_any[2][kWarps];   // ... and whether the warp closed a row at all
  unsigned long long full[STAGES];
  alignas(16) int4 coord[STAGES];            // {s.x, s.y, e.x, e.y} of the staged tile
};

// Everything a thread needs to know about a staged tile follows from its two
// merge coordinates (one 16-byte shared load) and the array alignments.
struct tile_geom {
  int sx, sy, nt, na, skew, cap, skew_r, bulk_a, bulk_r;
};
template <bool ARRAY_ENDS>
__device__ __forceinline__ tile_geom make_geom(int4 c, const int* row_end, int T, int A) {
  tile_geom g;
  g.sx = c.x; g.sy = c.y; g.nt = c.z - c.x; g.na = c.w - c.y;
  g.skew = c.y & 3;                                   // base pointers are 16-byte aligned
  g.cap = g.na > 0 ? (g.skew + g.na + 3) & ~3 : 0;    // whole chunks holding the tile's atoms
  const int room_a = (g.skew + (A - c.y)) & ~3;
  g.bulk_a = g.cap < room_a ? g.cap : room_a;
  g.skew_r = 0;
  g.bulk_r = 0;
  if (ARRAY_ENDS) {
    g.skew_r = int((reinterpret_cast<uintptr_t>(row_end + c.x) & 15u) >> 2);
    if (g.nt > 0) {
      const int need = (g.skew_r + g.nt + 3) & ~3;
      const int room = (g.skew_r + (T - c.x)) & ~3;
      g.bulk_r = need < room ? need : room;
    }
  }
  return g;
}

// One thread: launch the bulk copies of the tile between coordinates s and e.
template <int THREADS, int STAGES, bool ARRAY_ENDS>
__device__ __forceinline__ void issue_tile(shared_t<THREADS, STAGES>& sm, int st, int2 s, int2 e,
                                           const int* __restrict__ row_end,
                                           const int* __restrict__ indices,
                                           const float* __restrict__ values, int T, int A,
                                           uint64_t stream_policy) {
  const int4 c = make_int4(s.x, s.y, e.x, e.y);
  const tile_geom g = make_geom<ARRAY_ENDS>(c, row_end, T, A);
  sm.coord[st] = c;
  uint64_t* bar = reinterpret_cast<uint64_t*>(&sm.full[st]);
  const uint32_t bytes = 4u * uint32_t(2 * g.bulk_a + g.bulk_r);
  if (bytes == 0) {
    loops::tma::barrier_arrive(bar);
    return;
  }
  loops::tma::barrier_arrive_expect_tx(bar, bytes);
  if (g.bulk_a > 0) {
    loops::tma::bulk_g2s_hint(&sm.stage[st][0], indices + (s.y - g.skew), 4u * uint32_t(g.bulk_a), bar,
                              stream_policy);
    loops::tma::bulk_g2s_hint(&sm.stage[st][g.cap], values + (s.y - g.skew), 4u * uint32_t(g.bulk_a), bar,
                              stream_policy);
  }
  if (g.bulk_r > 0)
    loops::tma::bulk_g2s(&sm.stage[st][2 * g.cap], row_end + (s.x - g.skew_r), 4u * uint32_t(g.bulk_r), bar);
}

// Registers a thread carries for one tile between "gathers issued" and "reduced".
template <int CH>
struct tile_regs {
  float4 v[CH];      // values of the thread's chunks
  float x[CH][4];    // gathered x
  int sx, nt, na, skew;
};

// ACCUM: y += A x (every row the tile closes is read-modified-written by its one
// owner thread; rows without atoms are left alone) -- used for the second column
// block of a multi-GPU shard.
template <int THREADS, int MINB, int STAGES, bool ARRAY_ENDS, bool ACCUM>
__global__ void __launch_bounds__(THREADS, MINB)
    spmv_merge2_kernel(const int* __restrict__ row_end, int pitch, const int* __restrict__ indices,
                       const float* __restrict__ values, const float* __restrict__ x,
                       float* __restrict__ y, const int2* __restrict__ coords, int M, int T, int A,
                       int num_tiles, int* __restrict__ carry_row, float* __restrict__ carry_val) {
  using smem_t = shared_t<THREADS, STAGES>;
  constexpr int CH = smem_t::kChunks;
  constexpr int WARPS = smem_t::kWarps;
  constexpr int SPILL = 4 * CH * THREADS;    // first staged position of the 257th chunk (= 1024)
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  smem_t& sm = *reinterpret_cast<smem_t*>(smem_raw);

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int warp = t >> 5;
  const int first = blockIdx.x;
  const int stride = gridDim.x;
  if (first >= num_tiles) return;
  const int my_tiles = (num_tiles - first + stride - 1) / stride;
  // a warp owns 32*CH consecutive chunks; chunk slot u of this lane:
  const int chunk0 = warp * (32 * CH) + lane;

  uint64_t policy = 0;
  int issued = 0;   // thread 0: tiles put in flight so far
  auto issue_next = [&](int st) {
    const int j = first + issued * stride;
    const int2 s = __ldg(coords + j);
    const int2 e = __ldg(coords + (j + 1 < M ? j + 1 : M));
    issue_tile<THREADS, STAGES, ARRAY_ENDS>(sm, st, s, e, row_end, indices, values, T, A, policy);
    ++issued;
  };

  if (t == 0) {
    policy = loops::tma::policy_evict_first();
    for (int q = 0; q < STAGES; ++q) loops::tma::barrier_init(reinterpret_cast<uint64_t*>(&sm.full[q]), 1);
    issue_next(0);
    if (STAGES > 1 && my_tiles > 1) issue_next(1);
  }
  for (int w = t; w < 2 * (kTile + 8) / 2; w += THREADS) reinterpret_cast<unsigned*>(&sm.mark[0][0])[w] = 0u;
  __syncthreads();

  // ---- early part of tile n: row-end marks, then issue its gathers -----------
  auto early = [&](tile_regs<CH>& R, int n) {
    const int st = n & 1;                    // parity of the tile: marks, spill, warp aggregates
    const int sg = STAGES > 1 ? st : 0;      // bulk-copy stage it arrives in
    loops::tma::barrier_wait(reinterpret_cast<uint64_t*>(&sm.full[sg]),
                             uint32_t(STAGES > 1 ? n >> 1 : n) & 1u);
    const tile_geom m = make_geom<ARRAY_ENDS>(sm.coord[sg], row_end, T, A);
    R.sx = m.sx; R.nt = m.nt; R.na = m.na; R.skew = m.skew;
    int* sidx = &sm.stage[sg][0];
    float* sval = reinterpret_cast<float*>(&sm.stage[sg][m.cap]);
    int* sre = &sm.stage[sg][2 * m.cap];
    const bool ragged = m.bulk_a < m.cap || (ARRAY_ENDS && m.nt > 0 && m.bulk_r < m.skew_r + m.nt);
    if (ragged) {   // the end of an array (last tile only): finish the stage with ordinary loads
      for (int r = m.bulk_a + t; r < m.cap; r += THREADS) {
        const bool in = r < m.skew + m.na;
        sidx[r] = in ? indices[m.sy + (r - m.skew)] : 0;
        sval[r] = in ? values[m.sy + (r - m.skew)] : 0.0f;
      }
      if (ARRAY_ENDS)
        for (int r = m.bulk_r + t; r < m.skew_r + m.nt; r += THREADS)
          sre[r] = row_end[m.sx + (r - m.skew_r)];
      __syncthreads();
    }
    auto rel_end = [&](int i) -> int {
      if (ARRAY_ENDS) return sre[m.skew_r + i] - m.sy;
      return (m.sx + i + 1) * pitch - m.sy;
    };
    // one thread per row end: mark the row's last atom, or store the empty row
    unsigned short* MK = sm.mark[st];
    for (int i = t; i < m.nt; i += THREADS) {
      const int e = rel_end(i);
      const int b = i > 0 ? rel_end(i - 1) : 0;
      if (e > b) MK[m.skew + e - 1] = static_cast<unsigned short>(i + 1);
      else if (!ACCUM) y[m.sx + i] = 0.0f;
    }
    // gathers
    const int nchunks = m.cap >> 2;
    const int4* sidx4 = reinterpret_cast<const int4*>(sidx);
    const float4* sval4 = reinterpret_cast<const float4*>(sval);
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const int c = chunk0 + 32 * u;
      const int cc = c < nchunks ? c : 0;           // absent chunks re-read chunk 0 (masked later)
      int4 ci = make_int4(0, 0, 0, 0);
      float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nchunks > 0) { ci = sidx4[cc]; cv = sval4[cc]; }
      R.v[u] = cv;
      const int cols[4] = {ci.x, ci.y, ci.z, ci.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ARRAY_ENDS) R.x[u][q] = __ldg(x + cols[q]);
        else R.x[u][q] = cols[q] >= 0 ? __ldg(x + cols[q]) : 0.0f;
      }
    }
    // 257th chunk (only when fewer than `skew` rows end in the tile): done on the spot
    if (4 * nchunks > SPILL && t == THREADS - 1) {
      const int4 ci = sidx4[SPILL / 4];
      const float4 cv = sval4[SPILL / 4];
      const int cols[4] = {ci.x, ci.y, ci.z, ci.w};
      const float vals[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float pr = 0.0f;
        if (SPILL + q < m.skew + m.na) {
          const float xx = (ARRAY_ENDS || cols[q] >= 0) ? __ldg(x + cols[q]) : 0.0f;
          pr = __fmul_rn(vals[q], xx);
        }
        sm.spill[st][q] = pr;
      }
    }
  };

  // ---- reduce tile n (its gathers were issued one iteration ago) -------------
  auto reduce = [&](tile_regs<CH>& R, int n, bool refill) {
    const int st = n & 1;
    unsigned short* MK = sm.mark[st];
    float* yb = y + R.sx;
    const int lo = R.skew, hi = R.skew + R.na;       // valid staged positions [lo, hi)
    float head[CH], excl[CH], acc_val[CH];
    int head_row[CH];
    bool any[CH], open[CH], acc_any[CH];
    // running fold over this warp's earlier chunk slots: (sum since the last closed row, closed any?)
    float wv = 0.0f;
    bool wa = false;
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const int c = chunk0 + 32 * u;
      const int p0 = 4 * c;
      float p[4];
      const float vals[4] = {R.v[u].x, R.v[u].y, R.v[u].z, R.v[u].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) p[q] = __fmul_rn(vals[q], R.x[u][q]);
      if (p0 < lo || p0 + 4 > hi) {                  // boundary chunks: drop what is not the tile's
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (p0 + q < lo || p0 + q >= hi) p[q] = 0.0f;
      }
      uint2 mk = *reinterpret_cast<const uint2*>(MK + p0);
      if (mk.x | mk.y) *reinterpret_cast<uint2*>(MK + p0) = make_uint2(0u, 0u);   // single reader: clear for tile n+2
      const unsigned mks[4] = {mk.x & 0xffffu, mk.x >> 16, mk.y & 0xffffu, mk.y >> 16};
      float r = 0.0f;
      bool got = false;
      float hd = 0.0f;
      int hrow = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        r = __fadd_rn(r, p[q]);
        if (mks[q]) {
          const int row = int(mks[q]) - 1;
          if (!got) { hd = r; hrow = row; got = true; }
          else yb[row] = ACCUM ? __fadd_rn(yb[row], r) : r;     // row lies wholly inside this chunk
          r = 0.0f;
        }
      }
      // the very last chunk is followed by the spill chunk (rare)
      if (u == CH - 1 && t == THREADS - 1 && hi > SPILL) {
        const uint2 sk = *reinterpret_cast<const uint2*>(MK + SPILL);
        if (sk.x | sk.y) *reinterpret_cast<uint2*>(MK + SPILL) = make_uint2(0u, 0u);
        const unsigned sks[4] = {sk.x & 0xffffu, sk.x >> 16, sk.y & 0xffffu, sk.y >> 16};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          r = __fadd_rn(r, sm.spill[st][q]);
          if (sks[q]) {
            const int row = int(sks[q]) - 1;
            if (!got) { hd = r; hrow = row; got = true; }
            else yb[row] = ACCUM ? __fadd_rn(yb[row], r) : r;
            r = 0.0f;
          }
        }
      }
      head[u] = hd; head_row[u] = hrow; any[u] = got;
      // warp-level segmented scan of the tails (segments start at lanes that closed a row)
      const unsigned B = __ballot_sync(kFull, got);
      const unsigned below = B & ((1u << lane) - 1u);
      const int start = got ? lane : (below ? 31 - __clz(below) : 0);
      float v = r;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float tv = __shfl_up_sync(kFull, v, d);
        if (lane - d >= start) v = __fadd_rn(tv, v);
      }
      float ex = __shfl_up_sync(kFull, v, 1);
      if (lane == 0) ex = 0.0f;
      excl[u] = ex;
      open[u] = below == 0u;
      acc_val[u] = wv; acc_any[u] = wa;              // what this warp's earlier slots hand to this one
      const float last = __shfl_sync(kFull, v, 31);
      if (B != 0u) { wv = last; wa = true; }
      else wv = __fadd_rn(wv, last);
    }
    if (lane == 31) {
      sm.wt_val[st][warp] = wv;
      sm.wt_any[st][warp] = wa ? 1 : 0;
    }
    __syncthreads();
    if (t == 0 && refill) {
      // every thread is past its reads of the stage tile n+1 arrived in (its early part)
      loops::tma::fence_proxy_async();
      issue_next(STAGES > 1 ? (n + 1) & 1 : 0);
    }
    // fold the warp aggregates in sequence order
    float wvals[WARPS];
#pragma unroll
    for (int q = 0; q < WARPS; q += 4) {
      const float4 f = *reinterpret_cast<const float4*>(&sm.wt_val[st][q]);
      wvals[q] = f.x; wvals[q + 1] = f.y; wvals[q + 2] = f.z; wvals[q + 3] = f.w;
    }
    unsigned anyw[(WARPS + 3) / 4];
#pragma unroll
    for (int q = 0; q < WARPS; q += 4) anyw[q / 4] = *reinterpret_cast<const unsigned*>(&sm.wt_any[st][q]);
    float run = 0.0f, before = 0.0f;
#pragma unroll
    for (int q = 0; q < WARPS; ++q) {
      if (q == warp) before = run;
      run = ((anyw[q / 4] >> (8 * (q & 3))) & 0xffu) ? wvals[q] : __fadd_rn(run, wvals[q]);
    }
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      if (any[u]) {
        float carry_in = excl[u];
        if (open[u]) {
          const float prior = acc_any[u] ? acc_val[u] : __fadd_rn(before, acc_val[u]);
          carry_in = __fadd_rn(prior, excl[u]);
        }
        const float out = __fadd_rn(carry_in, head[u]);
        yb[head_row[u]] = ACCUM ? __fadd_rn(yb[head_row[u]], out) : out;
      }
    }
    if (t == 0) {
      const int j = first + n * stride;
      carry_row[j] = R.sx + R.nt;
      carry_val[j] = run;
    }
  };

  // tiles in flight ahead of the one being reduced: its successor (being consumed by
  // `early`) plus STAGES more in the bulk-copy stage(s)
  constexpr int AHEAD = STAGES + 1;
  tile_regs<CH> R0, R1;
  early(R0, 0);
  __syncthreads();
  if (t == 0 && my_tiles > STAGES) {
    loops::tma::fence_proxy_async();
    issue_next(0);
  }
  for (int n = 0; n < my_tiles; n += 2) {
    // the stage tile n+1 arrived in is re-filled with tile n+1+STAGES after the barrier in reduce(n)
    if (n + 1 < my_tiles) early(R1, n + 1);
    reduce(R0, n, n + AHEAD < my_tiles);
    if (n + 1 >= my_tiles) break;
    if (n + 2 < my_tiles) early(R0, n + 2);
    reduce(R1, n + 1, n + 1 + AHEAD < my_tiles);
  }
}

}  // namespace mp2
}  // namespace loopsb
