// loops_b200/csrc/spmv_merge.cuh -- merge_path_flat SpMV for sm_100a.
//
// Replaces reference algorithms/spmv/merge_path_flat.cuh:38-83 (kernel) and
// ell_merge_path.cuh:32-69 (ELL variant). Same partition arithmetic
// (W = tiles + atoms cut into merge tiles of TPB*IPT = 1024 items; coordinates
// S(b*1024) are the reference's generate_search_coordinates values), but a
// different execution plan:
//
//   * persistent CTAs (grid = resident CTAs/SM x SM count); CTA tile = G
//     consecutive merge tiles (TILE = G*1024 items), dealt round-robin;
//   * the tile's three streams -- row-end window, column indices, values --
//     are contiguous ranges of the CSR arrays, so ONE thread pulls them into a
//     double-buffered shared-memory stage with 1-D bulk async copies
//     (cp.async.bulk / UBLKCP) completing on an mbarrier; the next tile is in
//     flight while the current one is processed;
//   * gather/multiply phase: 128-bit shared loads of indices/values, x gathered
//     through the read-only path (16 independent loads per thread), products
//     parked in a padded shared array;
//   * reduce phase: each thread walks TILE/THREADS consecutive merge items
//     (per-thread diagonal search runs on the staged row-end window), keeps a
//     head/tail partial, a warp-shuffle segmented scan joins rows that span
//     threads, completed rows are written with coalesced stores;
//   * the partial row a tile ends in is emitted as one (row, value) carry per
//     tile and folded in by a tiny fix-up kernel -- no atomics, y is fully
//     overwritten and the result is deterministic.
//
// Products are rounded to fp32 before the adds (no FMA contraction), so every
// term equals the reference's `values[nz] * x[indices[nz]]`; only the order of
// the adds differs from the CPU validator.
#pragma once

#include "common.cuh"

#include <loops/util/tma.hxx>

namespace loopsb {
namespace mp {

constexpr int kRefItemsPerMergeTile = 128 * 8;  // launch_box.hxx:66-68 (sm_100, f32)

struct tile_meta {
  int sx, sy;          // tile start coordinate (tile id, atom id)
  int nt, na;          // row ends / atoms inside the tile
  int skew_i, skew_v, skew_r;   // leading alignment slack of each staged stream
  int off_v, off_r;    // int offsets of the value / row-end regions in the stage
  int bulk_i, bulk_v, bulk_r;   // elements delivered by the bulk copies
};

template <int THREADS, int TILE, int STAGES>
struct merge_shared {
  static constexpr int kStageInts = 2 * TILE + 32;
  static constexpr int kProdWords = TILE + TILE / 32 + 20;  // + speculative reads past na
  alignas(16) int stage[STAGES][kStageInts];
  alignas(16) float prod[kProdWords];
  unsigned long long full[STAGES];
  tile_meta meta[STAGES];
  float warp_val[THREADS / 32];
  int warp_flag[THREADS / 32];
};

__device__ __forceinline__ int pad_index(int p) { return p + (p >> 5); }

// How the tile-end sequence is obtained.
struct ends_array {          // csr / csc / bcsr: row_end = offsets + 1
  const int* row_end;
};
struct ends_pitch {          // ell / dia: tile_end(k) = (k + 1) * pitch
  int pitch;
};

// ---------------------------------------------------------------------------
// One thread: describe tile `j` and launch its bulk copies into stage `st`.
// ---------------------------------------------------------------------------
template <int THREADS, int TILE, int STAGES, bool ARRAY_ENDS>
__device__ __forceinline__ void issue_tile(
    merge_shared<THREADS, TILE, STAGES>& sm, int st, int2 s, int2 e,
    const int* __restrict__ row_end, const int* __restrict__ indices,
    const float* __restrict__ values, int T, int A, uint64_t stream_policy) {
  tile_meta m;
  m.sx = s.x; m.sy = s.y; m.nt = e.x - s.x; m.na = e.y - s.y;
  const int* gi = indices + s.y;
  const float* gv = values + s.y;
  m.skew_i = int((reinterpret_cast<uintptr_t>(gi) & 15u) >> 2);
  m.skew_v = int((reinterpret_cast<uintptr_t>(gv) & 15u) >> 2);
  auto clamp_bulk = [](int skew, int want, int avail) {
    // whole 16-byte chunks that cover [0, skew + want) without running past
    // the end of the array (avail = elements from the aligned start to its end)
    if (want <= 0) return 0;
    int need = (skew + want + 3) & ~3;
    int room = (skew + avail) & ~3;
    return need < room ? need : room;
  };
  m.bulk_i = clamp_bulk(m.skew_i, m.na, A - s.y);
  m.bulk_v = clamp_bulk(m.skew_v, m.na, A - s.y);
  const int cap_i = (m.skew_i + m.na + 3) & ~3;
  const int cap_v = (m.skew_v + m.na + 3) & ~3;
  m.off_v = cap_i;
  m.off_r = cap_i + cap_v;
  m.skew_r = 0;
  m.bulk_r = 0;
  const int* gr = nullptr;
  if (ARRAY_ENDS) {
    gr = row_end + s.x;
    m.skew_r = int((reinterpret_cast<uintptr_t>(gr) & 15u) >> 2);
    m.bulk_r = clamp_bulk(m.skew_r, m.nt, T - s.x);
  }
  sm.meta[st] = m;
  uint64_t* bar = reinterpret_cast<uint64_t*>(&sm.full[st]);
  const uint32_t bytes = 4u * uint32_t(m.bulk_i + m.bulk_v + m.bulk_r);
  if (bytes == 0) {
    loops::tma::barrier_arrive(bar);
    return;
  }
  loops::tma::barrier_arrive_expect_tx(bar, bytes);
  if (m.bulk_i > 0)
    loops::tma::bulk_g2s_hint(&sm.stage[st][0], gi - m.skew_i,
                              4u * uint32_t(m.bulk_i), bar, stream_policy);
  if (m.bulk_v > 0)
    loops::tma::bulk_g2s_hint(&sm.stage[st][m.off_v], gv - m.skew_v,
                              4u * uint32_t(m.bulk_v), bar, stream_policy);
  if (m.bulk_r > 0)
    loops::tma::bulk_g2s(&sm.stage[st][m.off_r], gr - m.skew_r,
                         4u * uint32_t(m.bulk_r), bar);
}

// ---------------------------------------------------------------------------
// The SpMV kernel.
//   coords      : S(b*1024), b = 0..M (int2 = {tile, atom})
//   G           : merge tiles per CTA tile (TILE == G*1024)
//   carry_row/val[num_cta_tiles] : the partial row each CTA tile ends in
// ELL (ARRAY_ENDS == false): padding slots carry column -1 and contribute 0
// (reference ell_merge_path.cuh:60).
// ---------------------------------------------------------------------------
template <int THREADS, int TILE, int STAGES, int MINB, bool ARRAY_ENDS>
__global__ void __launch_bounds__(THREADS, MINB)
    spmv_merge_kernel(const int* __restrict__ row_end, int pitch,
                      const int* __restrict__ indices,
                      const float* __restrict__ values,
                      const float* __restrict__ x, float* __restrict__ y,
                      const int2* __restrict__ coords, int M, int G, int T,
                      int A, int num_cta_tiles, int* __restrict__ carry_row,
                      float* __restrict__ carry_val,
                      long long* __restrict__ phase_cycles) {
  using shared_t = merge_shared<THREADS, TILE, STAGES>;
  constexpr int IPT = TILE / THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  shared_t& sm = *reinterpret_cast<shared_t*>(smem_raw);

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int warp = t >> 5;
  const int first_tile = blockIdx.x;
  const int stride = gridDim.x;
  if (first_tile >= num_cta_tiles)
    return;

  uint64_t policy = 0;
  int2 nxt_s = make_int2(0, 0), nxt_e = make_int2(0, 0);
  int issue_tile_id = first_tile;  // next tile thread 0 will put in flight

  auto load_coords = [&](int j) {
    const int b0 = j * G;
    int b1 = b0 + G;
    if (b1 > M) b1 = M;
    nxt_s = __ldg(coords + b0);
    nxt_e = __ldg(coords + b1);
  };

  if (t == 0) {
    policy = loops::tma::policy_evict_first();
    for (int s = 0; s < STAGES; ++s)
      loops::tma::barrier_init(reinterpret_cast<uint64_t*>(&sm.full[s]), 1);
    // Prologue: fill every stage that has a tile.
    for (int s = 0; s < STAGES && issue_tile_id < num_cta_tiles; ++s) {
      load_coords(issue_tile_id);
      issue_tile<THREADS, TILE, STAGES, ARRAY_ENDS>(
          sm, s, nxt_s, nxt_e, row_end, indices, values, T, A, policy);
      issue_tile_id += stride;
    }
    if (issue_tile_id < num_cta_tiles)
      load_coords(issue_tile_id);
  }
  __syncthreads();

  // Optional per-phase cycle accounting (thread 0 of every CTA; tuning aid).
  long long ph[6] = {0, 0, 0, 0, 0, 0};
  long long tick = 0;
  const bool timing = (phase_cycles != nullptr) && (t == 0);
  auto lap = [&](int which) {
    if (timing) { const long long now = clock64(); ph[which] += now - tick; tick = now; }
  };
  if (timing) tick = clock64();

  int k = 0;
  for (int j = first_tile; j < num_cta_tiles; j += stride, ++k) {
    const int st = k % STAGES;
    const uint32_t parity = uint32_t(k / STAGES) & 1u;
    loops::tma::barrier_wait(reinterpret_cast<uint64_t*>(&sm.full[st]), parity);
    lap(0);  // waiting for the stage

    const tile_meta m = sm.meta[st];
    int* sidx = &sm.stage[st][0];
    float* sval = reinterpret_cast<float*>(&sm.stage[st][m.off_v]);
    int* sre = &sm.stage[st][m.off_r];

    // Ragged end of an array (only the very last tile): finish the stage with
    // ordinary loads so the hot loops below never branch on it.
    const bool ragged = (m.bulk_i < m.skew_i + m.na) ||
                        (m.bulk_v < m.skew_v + m.na) ||
                        (ARRAY_ENDS && m.bulk_r < m.skew_r + m.nt);
    if (ragged) {
      for (int r = m.bulk_i + t; r < ((m.skew_i + m.na + 3) & ~3); r += THREADS)
        sidx[r] = (r < m.skew_i + m.na) ? indices[m.sy + (r - m.skew_i)] : 0;
      for (int r = m.bulk_v + t; r < m.skew_v + m.na; r += THREADS)
        sval[r] = values[m.sy + (r - m.skew_v)];
      if (ARRAY_ENDS)
        for (int r = m.bulk_r + t; r < m.skew_r + m.nt; r += THREADS)
          sre[r] = row_end[m.sx + (r - m.skew_r)];
      __syncthreads();
    }

    // Row end of tile-local row i, relative to the tile's first atom.
    auto rel_end = [&](int i) -> int {
      if (ARRAY_ENDS) return sre[m.skew_r + i] - m.sy;
      return (m.sx + i + 1) * pitch - m.sy;
    };
    const int items = m.nt + m.na;
    int d = t * IPT;
    if (d > items) d = items;
    int row, atom0;
    // Per-thread diagonal search on the staged row-end window. Independent of
    // the x gathers, so it is placed between their issue and their use and
    // runs while they are in flight.
    auto thread_search = [&]() {
      int lo = d - m.na; if (lo < 0) lo = 0;
      int hi = d < m.nt ? d : m.nt;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rel_end(mid) <= d - mid - 1) lo = mid + 1; else hi = mid;
      }
      row = lo;
      atom0 = d - lo;
    };

    // ---------------- gather / multiply ----------------
    if (m.skew_i == m.skew_v) {
      const int skew = m.skew_i;
      const int nchunks = m.na > 0 ? (skew + m.na + 3) >> 2 : 0;
      const int4* sidx4 = reinterpret_cast<const int4*>(sidx);
      const float4* sval4 = reinterpret_cast<const float4*>(sval);
      constexpr int U = TILE / (4 * THREADS);   // 16-byte chunks per thread
      const unsigned una = unsigned(m.na);
      int4 ci[U];
      float4 cv[U];
      float xv[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = u * THREADS + t;
        // chunks past the end read chunk 0 and are masked at the store; every
        // index inside a present chunk is a real column id (neighbouring
        // atoms) or was sanitised above.
        const int cc = (c < nchunks) ? c : 0;
        ci[u] = sidx4[cc];
        cv[u] = sval4[cc];
      }
      if (nchunks > 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int cols[4] = {ci[u].x, ci[u].y, ci[u].z, ci[u].w};
          const int cq = u * THREADS + t;
          const int pq = (cq < nchunks ? 4 * cq : -8) - skew;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // words outside the tile's atoms (alignment slack in front of the first
            // atom, the tail of the last chunk) may be anything when the caller's
            // arrays are views that do not start on a 16-byte boundary: never
            // gather through them
            const int col = unsigned(pq + q) < una ? cols[q] : 0;
            if (ARRAY_ENDS) xv[u][q] = __ldg(x + col);
            else xv[u][q] = col >= 0 ? __ldg(x + col) : 0.0f;
          }
        }
      }
      thread_search();
      if (nchunks > 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int c = u * THREADS + t;
          const int p0 = (c < nchunks ? 4 * c : -8) - skew;
          const float vals[4] = {cv[u].x, cv[u].y, cv[u].z, cv[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int p = p0 + q;
            if (unsigned(p) < una)
              sm.prod[pad_index(p)] = __fmul_rn(vals[q], xv[u][q]);
          }
        }
        // a non-zero skew spills at most one chunk past U*THREADS
        if (t == 0 && nchunks > U * THREADS) {
          const int c = U * THREADS;
          const int4 i4 = sidx4[c];
          const float4 v4 = sval4[c];
          const int cols[4] = {i4.x, i4.y, i4.z, i4.w};
          const float vals[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int p = 4 * c - skew + q;
            if (unsigned(p) < una) {
              const float xx = (ARRAY_ENDS || cols[q] >= 0) ? __ldg(x + cols[q]) : 0.0f;
              sm.prod[pad_index(p)] = __fmul_rn(vals[q], xx);
            }
          }
        }
      }
    } else {
      // Streams with different 16-byte phases (caller passed oddly aligned
      // arrays): scalar shared loads.
      for (int p = t; p < m.na; p += THREADS) {
        const int col = sidx[m.skew_i + p];
        const float v = sval[m.skew_v + p];
        const float xx = (ARRAY_ENDS || col >= 0) ? __ldg(x + col) : 0.0f;
        sm.prod[pad_index(p)] = __fmul_rn(v, xx);
      }
      thread_search();
    }
    __syncthreads();
    lap(1);  // gather / multiply / search (incl. barrier)

    // ---------------- per-thread merge walk ----------------
    int budget = items - d;
    if (budget > IPT) budget = IPT;
    // All candidate products up front: IPT independent shared loads instead
    // of a load->compare->load chain (reads past the thread's own atoms are
    // harmless and unused).
    float pr[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q) pr[q] = sm.prod[pad_index(atom0 + q)];
    // Two row ends are kept in registers: the one being watched and the one
    // after it. Closing a row shifts them and re-loads the look-ahead with a
    // shared load whose latency is not needed until the NEXT closing, so the
    // common case (at most one row ends at a given atom) is straight-line,
    // predicated code without a dependent load.
    constexpr int kNever = 0x7fffffff;
    int rend = (row < m.nt) ? rel_end(row) : kNever;
    int rend1 = (row + 1 < m.nt) ? rel_end(row + 1) : kNever;
    float sum = 0.0f, head = 0.0f;
    int first_row = -1;
    auto close_row = [&]() {
      if (first_row < 0) { head = sum; first_row = row; }
      else y[m.sx + row] = sum;        // row lies wholly inside this thread
      sum = 0.0f;
      ++row;
      rend = rend1;
      rend1 = (row + 1 < m.nt) ? rel_end(row + 1) : kNever;
      --budget;
    };
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
      // rows that end at or before atom (atom0 + q) are closed first -- the
      // same order the merge path visits them in
      if (budget > 0 && rend <= atom0 + q) close_row();
      while (budget > 0 && rend <= atom0 + q) close_row();   // empty rows
      if (budget > 0) {
        sum = __fadd_rn(sum, pr[q]);
        --budget;
      }
    }
    lap(2);  // walk

    // ---------------- segmented scan of (flag, value) over threads -------
    // combine(a, b) = (a.f | b.f, b.f ? b.v : a.v + b.v)
    int f = (first_row >= 0) ? 1 : 0;
    float v = sum;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const float pv = __shfl_up_sync(0xffffffffu, v, dlt);
      const int pf = __shfl_up_sync(0xffffffffu, f, dlt);
      if (lane >= dlt) {
        if (!f) v = __fadd_rn(pv, v);
        f |= pf;
      }
    }
    // exclusive (what precedes this thread inside the warp)
    float ev = __shfl_up_sync(0xffffffffu, v, 1);
    int ef = __shfl_up_sync(0xffffffffu, f, 1);
    if (lane == 0) { ev = 0.0f; ef = 0; }
    if (lane == 31) { sm.warp_val[warp] = v; sm.warp_flag[warp] = f; }
    __syncthreads();
    // Every thread is done with stage `st` (indices/values before the first
    // barrier, row ends before this one): put the next tile in flight now.
    if (t == 0 && issue_tile_id < num_cta_tiles) {
      loops::tma::fence_proxy_async();
      issue_tile<THREADS, TILE, STAGES, ARRAY_ENDS>(
          sm, st, nxt_s, nxt_e, row_end, indices, values, T, A, policy);
      issue_tile_id += stride;
      if (issue_tile_id < num_cta_tiles)
        load_coords(issue_tile_id);
    }
    lap(3);  // scan barrier + refill issue
    // what precedes this warp inside the CTA
    float wv = 0.0f; int wf = 0;
    for (int w = 0; w < warp; ++w) {
      const float bv = sm.warp_val[w];
      const int bf = sm.warp_flag[w];
      wv = bf ? bv : __fadd_rn(wv, bv);
      wf |= bf;
    }
    const float carry_in = ef ? ev : __fadd_rn(wv, ev);
    if (first_row >= 0)
      y[m.sx + first_row] = __fadd_rn(carry_in, head);
    if (t == THREADS - 1) {
      // CTA-wide inclusive aggregate = the partial of the row the tile ends in.
      const float total = f ? v : __fadd_rn(wv, v);
      carry_row[j] = m.sx + m.nt;
      carry_val[j] = total;
    }
    lap(4);  // cross-warp combine + row stores
    // No barrier here: the next tile's first barrier orders the scratch reuse
    // (warp_val/warp_flag are rewritten only after it, prod only before it by
    // threads that have all passed the barrier above).
  }
  if (timing) {
    for (int q = 0; q < 6; ++q) phase_cycles[blockIdx.x * 8 + q] = ph[q];
    phase_cycles[blockIdx.x * 8 + 6] = k;
  }
}

// Fold the per-tile carries into y: all carries naming the same row are summed
// left to right by the first of them, then added once.
__global__ void spmv_merge_fixup_kernel(const int* __restrict__ carry_row,
                                        const float* __restrict__ carry_val,
                                        int n, int T, float* __restrict__ y) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int row = carry_row[j];
  if (row >= T) return;
  if (j > 0 && carry_row[j - 1] == row) return;
  float acc = carry_val[j];
  for (int q = j + 1; q < n && carry_row[q] == row; ++q)
    acc = __fadd_rn(acc, carry_val[q]);
  y[row] = __fadd_rn(acc, y[row]);
}

// coords[b] = S(b * items_per_merge_tile), b = 0..M, for array / pitch ends.
// With `diagonals` (M+1 ascending values) the cut points are arbitrary -- used
// by work_oriented, whose block boundaries are multiples of its per-thread
// share rather than of the merge-tile size.
template <bool ARRAY_ENDS>
__global__ void merge_coords_kernel(const int* __restrict__ row_end, int pitch,
                                    int T, int A, long long items, int M,
                                    int2* __restrict__ coords,
                                    const long long* __restrict__ diagonals = nullptr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > M) return;
  const long long d = diagonals ? diagonals[b] : (long long)b * items;
  long long lo = d - A; if (lo < 0) lo = 0;
  long long hi = d < T ? d : (long long)T;
  const long long x_min = lo;
  while (lo < hi) {
    const long long mid = lo + ((hi - lo) >> 1);
    const long long end_mid = ARRAY_ENDS ? (long long)__ldg(row_end + mid)
                                         : (mid + 1) * (long long)pitch;
    if (end_mid <= d - mid - 1) lo = mid + 1; else hi = mid;
  }
  if (hi < x_min) lo = x_min;
  coords[b] = make_int2(int(lo < T ? lo : T), int(d - lo));
}

}  // namespace mp
}  // namespace loopsb
