// loops_b200/csrc/spmv_tiled.cuh -- band-tiled CSR SpMV for sm_100a.
//
// Why this exists: with column ids spread over a 4 MB x,
// every x[col] read of the merge-path kernel is its own 32-byte sector request
// and an SM issues about one such request per clock (tools/microbench.cu), so
// the CSR kernels top out near 120 us on BASELINE config 2 however the rest is
// written -- a third of the HBM roofline. Shared memory serves 32 gathers per
// clock. This path therefore keeps BOTH vectors of the product in shared
// memory: the plan (the reference's merge_path::preprocess_t, generalised --
// schedule/merge_path_flat.hxx:92-172) owns a re-ordered copy of the matrix cut
// into (row block x column band) tiles, and the kernel streams that copy once
// (8 bytes per nonzero, the same bytes CSR costs) while
//   * the CTA's y rows live in shared memory for the whole kernel,
//   * the x band of the current tile is a TMA bulk copy into a small ring,
//   * the nonzero stream arrives through per-warp TMA rings (1 KB steps).
// The public contract is unchanged: the caller passes CSR arrays, x and y.
//
// Geometry. rows are cut into `nb` row blocks of `rb` rows, columns into `q`
// parts of `cq` columns; CTA (block, part) owns one row block x one column part
// and `warps` consumer warps split its rows into sub-blocks of `rw` rows, so no
// two warps ever touch the same y row. A column part is walked in bands of
// `cb` columns that cycle through an `xb`-deep x ring. With q > 1 each CTA
// emits a partial y; the last CTA of a row block to finish adds the `q` partials
// in a fixed order (deterministic, no floating-point atomics).
//
// Stream format (one stream per consumer warp, built by the plan). Entries are
// ordered (band, row, column) -- i.e. CSR order inside a band -- and packed in
// STEPS of 128 entries = 1 KB: words [0,128) are packed ids, words [128,256)
// the fp32 values, entry (lane, j) at word lane*4 + j, so a lane's four entries
// are one conflict-free 128-bit shared load and are consecutive in CSR order.
//   id bits  0..15  position of x[col] in the x ring ((band % xb)*cb + col - band start)
//      bits 16..30  row inside the CTA's row block
//      bit  31      set on the entries of a row that the fast path cannot
//                   combine inside the step (see "flags" in build_host); any
//                   such entry sends the step down the general path
// Padding entries have value 0, point at a zero word behind the x ring and at a
// scratch row behind the y rows, so they need no predicate.
// A step only holds bands of one window of `xb` consecutive bands (the plan
// pads to a step boundary otherwise), which makes the x ring deadlock-free.
#pragma once

#include "common.cuh"

#include <loops/util/tma.hxx>

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace loopsb {
namespace bt {

constexpr int kLanes = 32;
constexpr int kPerLane = 4;
constexpr int kStep = kLanes * kPerLane;  // entries per step
constexpr int kStepWords = 2 * kStep;     // 256 words = 1 KB
constexpr uint32_t kFlagBit = 0x80000000u;
constexpr int kMaxRowsPerBlock = 32767;   // 15 row bits, one value kept for the scratch row
constexpr int kMaxRingFloats = 65532;     // 16 column bits, zero word behind the ring

struct geom {
  // chosen
  int nb = 0, q = 0, warps = 0, cb = 0, xb = 0, es = 0;
  // derived
  int rb = 0, rw = 0, cq = 0, nband = 0;
  int rows = 0, cols = 0;
  int nstreams() const { return nb * q * warps; }
  int grid() const { return nb * q; }
  int cta_threads() const { return (warps + 1) * 32; }
  int zero_slot() const { return xb * cb; }
  // shared-memory carve-up (bytes), identical in the kernel
  int bar_count() const { return (2 * xb + warps * es + 1) & ~1; }
  int ys_words() const { return (rb + 1 + 3) & ~3; }
  int xs_words() const { return xb * cb + 4; }
  int table_bytes() const { return ((2 * warps * nband * 2 + 4) + 15) & ~15; }
  int smem_bytes() const {
    return bar_count() * 8 + warps * es * kStepWords * 4 + xs_words() * 4 + ys_words() * 4 + table_bytes();
  }
};

inline int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

// Fill the derived fields; returns false (with a reason) when the format's
// field widths cannot hold the geometry.
inline bool derive(geom& g, int rows, int cols, const char** why) {
  static const char* reasons[] = {"geometry fields must be positive", "cb must be a multiple of 4",
                                  "x ring exceeds 16 bits of position", "row block exceeds 15 bits of row",
                                  "more than 31 consumer warps", "more than 8 column parts"};
  g.rows = rows; g.cols = cols;
  if (g.nb < 1 || g.q < 1 || g.warps < 1 || g.cb < 4 || g.xb < 2 || g.es < 2) { *why = reasons[0]; return false; }
  if (g.cb % 4) { *why = reasons[1]; return false; }
  if (g.xb * g.cb > kMaxRingFloats) { *why = reasons[2]; return false; }
  if (g.warps > 31) { *why = reasons[4]; return false; }
  if (g.q > 8) { *why = reasons[5]; return false; }
  g.rb = std::max(1, ceil_div(rows, g.nb));
  if (g.rb > kMaxRowsPerBlock) { *why = reasons[3]; return false; }
  g.rw = std::max(1, ceil_div(g.rb, g.warps));
  g.cq = std::max(4, (ceil_div(cols, g.q) + 3) & ~3);
  g.nband = std::max(1, ceil_div(g.cq, g.cb));
  return true;
}

// Host image of the tiled copy.
struct host_image {
  geom g;
  std::vector<uint32_t> steps;       // total_steps * 256 words
  std::vector<int32_t> stream_base;  // nstreams + 1, in steps
  std::vector<uint16_t> fs;          // [stream][band] first step holding the band
  std::vector<uint16_t> le;          // [stream][band] one past the last step holding it
  long long total_steps = 0, real_entries = 0, pad_entries = 0, flagged_entries = 0, flagged_steps = 0;
};

// Build the image from HOST CSR arrays. Pure host code (unit-tested on CPU).
// Returns LOOPSB_OK / LOOPSB_ERR_UNSUPPORTED / LOOPSB_ERR_INVALID.
inline int build_host(host_image& im, geom g, int rows, int cols, const int* off, const int* idx,
                      const float* val) {
  const char* why = "";
  if (!derive(g, rows, cols, &why)) { set_error("band-tiled plan: %s", why); return LOOPSB_ERR_UNSUPPORTED; }
  im.g = g;
  const int ns = g.nstreams(), nband = g.nband;
  const long long nnz = rows > 0 ? off[rows] : 0;
  std::vector<int32_t> count(size_t(ns) * nband, 0);
  // ---- pass 1: entries per (stream, band) ----
  for (int r = 0; r < rows; ++r) {
    const int rbi = r / g.rb, lr = r - rbi * g.rb, w = lr / g.rw;
    const size_t sb = size_t(rbi) * g.q;
    for (int a = off[r]; a < off[r + 1]; ++a) {
      const int c = idx[a];
      if (c < 0 || c >= cols) { set_error("band-tiled plan: column id %d out of range at atom %d", c, a); return LOOPSB_ERR_INVALID; }
      const int qi = c / g.cq, cl = c - qi * g.cq, b = cl / g.cb;
      ++count[((sb + qi) * g.warps + w) * nband + b];
    }
  }
  // ---- layout: padded start of every band in every stream ----
  std::vector<int32_t> start(size_t(ns) * nband, 0);
  im.stream_base.assign(size_t(ns) + 1, 0);
  im.fs.assign(size_t(ns) * nband, 0);
  im.le.assign(size_t(ns) * nband, 0);
  long long total = 0;
  for (int s = 0; s < ns; ++s) {
    long long pos = 0;
    int step_min_band = -1;  // smallest band with an entry in the step that holds `pos`
    const int32_t* cnt = &count[size_t(s) * nband];
    int32_t* st = &start[size_t(s) * nband];
    std::vector<long long> end(nband, 0);
    for (int b = 0; b < nband; ++b) {
      if (cnt[b] == 0) { st[b] = int32_t(pos); end[b] = pos; continue; }
      if (pos % kStep != 0 && b > step_min_band + g.xb - 1) pos = (pos + kStep - 1) / kStep * kStep;
      if (pos % kStep == 0) step_min_band = b;
      st[b] = int32_t(pos);
      const long long e = pos + cnt[b];
      if ((e - 1) / kStep > pos / kStep) step_min_band = b;  // the step `e` lands in started inside band b
      pos = e;
      end[b] = e;
    }
    const long long nsteps = (pos + kStep - 1) / kStep;
    if (nsteps > 65535) { set_error("band-tiled plan: a warp stream needs %lld steps (> 65535)", nsteps); return LOOPSB_ERR_UNSUPPORTED; }
    // band tables: empty bands borrow the first step of the next non-empty one
    int next_fs = int(nsteps);
    for (int b = nband - 1; b >= 0; --b) {
      uint16_t& f = im.fs[size_t(s) * nband + b];
      uint16_t& l = im.le[size_t(s) * nband + b];
      if (cnt[b] > 0) {
        next_fs = int(st[b] / kStep);
        f = uint16_t(next_fs);
        l = uint16_t((end[b] - 1) / kStep + 1);
      } else {
        f = uint16_t(next_fs);
        l = uint16_t(next_fs);
      }
    }
    im.stream_base[s] = int32_t(total);
    total += nsteps;
  }
  im.stream_base[ns] = int32_t(total);
  if (total > 0x7fffffffLL) { set_error("band-tiled plan: too many steps"); return LOOPSB_ERR_UNSUPPORTED; }
  im.total_steps = total;
  // ---- pass 2: scatter (padding pre-filled) ----
  const uint32_t pad_id = (uint32_t(g.rb) << 16) | uint32_t(g.zero_slot());
  im.steps.assign(size_t(total) * kStepWords, 0u);
  for (long long s = 0; s < total; ++s) {
    uint32_t* w = &im.steps[size_t(s) * kStepWords];
    for (int i = 0; i < kStep; ++i) w[i] = pad_id;
  }
  std::vector<int32_t>& cursor = start;  // consumed in place
  for (int r = 0; r < rows; ++r) {
    const int rbi = r / g.rb, lr = r - rbi * g.rb, w = lr / g.rw;
    const size_t sb = size_t(rbi) * g.q;
    for (int a = off[r]; a < off[r + 1]; ++a) {
      const int c = idx[a];
      const int qi = c / g.cq, cl = c - qi * g.cq, b = cl / g.cb, lc = cl - b * g.cb;
      const size_t stream = (sb + qi) * g.warps + w;
      const int32_t p = cursor[stream * nband + b]++;
      uint32_t* sw = &im.steps[(size_t(im.stream_base[stream]) + size_t(p / kStep)) * kStepWords];
      const int slot = p % kStep;  // lane*4 + j
      sw[slot] = (uint32_t(lr) << 16) | uint32_t((b % g.xb) * g.cb + lc);
      float v = val[a];
      uint32_t vb; memcpy(&vb, &v, 4);
      sw[kStep + slot] = vb;
    }
  }
  im.real_entries = nnz;
  im.pad_entries = total * kStep - nnz;
  // ---- pass 3: flags. A step is "clean" when every row it touches sits in ONE
  // contiguous range of slots (slot = lane*4 + j, i.e. CSR order) that spans at
  // most two adjacent lanes -- what the kernel's fast path can combine with a
  // per-lane run sum and one neighbour shuffle. Entries of rows that break the
  // rule (a row seen in two bands of the same step, or a run over 3+ lanes) get
  // bit 31 and send the whole step down the general path. ----
  {
    std::vector<int64_t> stamp(size_t(g.rb) + 1, -1);
    std::vector<int32_t> first(size_t(g.rb) + 1, 0), last(size_t(g.rb) + 1, 0), cnt(size_t(g.rb) + 1, 0);
    for (long long s = 0; s < total; ++s) {
      uint32_t* w = &im.steps[size_t(s) * kStepWords];
      for (int p = 0; p < kStep; ++p) {
        const int lr = int((w[p] >> 16) & 0x7fff);
        if (lr == g.rb) continue;  // padding -> scratch row, harmless
        if (stamp[lr] != s) { stamp[lr] = s; first[lr] = last[lr] = p; cnt[lr] = 1; }
        else { last[lr] = p; ++cnt[lr]; }
      }
      bool any = false;
      for (int p = 0; p < kStep; ++p) {
        const int lr = int((w[p] >> 16) & 0x7fff);
        if (lr == g.rb) continue;
        const bool bad = (last[lr] - first[lr] + 1 != cnt[lr]) ||
                         (last[lr] / kPerLane - first[lr] / kPerLane >= 2);
        if (bad) { w[p] |= kFlagBit; ++im.flagged_entries; any = true; }
      }
      if (any) ++im.flagged_steps;
    }
  }
  return LOOPSB_OK;
}

// ---------------------------------------------------------------------------
// Device side
// ---------------------------------------------------------------------------
struct params {
  const uint32_t* steps;
  const int32_t* stream_base;
  const uint16_t* fs;
  const uint16_t* le;
  const float* x;
  float* y;
  float* partial;       // [q][nb*rb] when q > 1
  unsigned* counters;   // [nb] when q > 1
  int rows, cols, rb, cq, cb, xb, es, nband, q, nb;
  long long* prof;      // PROFILE builds: 8 counters per consumer warp
};

// General y update for one slot of a dirty step: any lanes may hold the same
// row. Lanes are grouped by row (match.any); the lowest lane of each group
// updates, the rest retry, so equal rows are applied one after another in
// lane order. Padding (scratch row) is skipped.
__device__ __forceinline__ void rmw_general(float* ys, int r, float p, int scratch) {
  bool todo = r != scratch;
  unsigned pend = __ballot_sync(0xffffffffu, todo);
  while (pend) {
    if (todo) {
      const unsigned grp = __match_any_sync(pend, r);
      if ((__ffs(grp) - 1) == int(threadIdx.x & 31)) { ys[r] += p; todo = false; }
    }
    __syncwarp();
    pend = __ballot_sync(0xffffffffu, todo);
  }
}

template <int WARPS, bool PROFILE = false>
__global__ void __launch_bounds__((WARPS + 1) * 32, 1) spmv_bt_kernel(const params p) {
  extern __shared__ __align__(16) unsigned char bt_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int rbi = cta / p.q, qi = cta - rbi * p.q;
  const int row0 = rbi * p.rb;
  const int rows_here = min(p.rb, p.rows - row0);
  if (rows_here <= 0) return;  // whole row block past the end: nothing to write

  // ---- carve shared memory (geom::smem_bytes order) ----
  uint64_t* xfull = reinterpret_cast<uint64_t*>(bt_smem);
  uint64_t* xempty = xfull + p.xb;
  uint64_t* efull = xempty + p.xb;
  const int nbar = (2 * p.xb + WARPS * p.es + 1) & ~1;
  uint32_t* ering = reinterpret_cast<uint32_t*>(xfull + nbar);
  float* xs = reinterpret_cast<float*>(ering + WARPS * p.es * kStepWords);
  float* ys = xs + (p.xb * p.cb + 4);
  const int ys_words = (p.rb + 1 + 3) & ~3;
  uint16_t* fs_s = reinterpret_cast<uint16_t*>(ys + ys_words);
  uint16_t* le_s = fs_s + WARPS * p.nband;
  int* last_flag = reinterpret_cast<int*>(le_s + WARPS * p.nband);

  if (tid == 0) {
    for (int k = 0; k < p.xb; ++k) {
      loops::tma::barrier_init(&xfull[k], 1);
      loops::tma::barrier_init(&xempty[k], WARPS);
    }
    for (int i = 0; i < WARPS * p.es; ++i) loops::tma::barrier_init(&efull[i], 1);
  }
  for (int i = tid; i < ys_words; i += (WARPS + 1) * 32) ys[i] = 0.f;
  if (tid < 4) xs[p.xb * p.cb + tid] = 0.f;
  {
    const size_t tb = size_t(cta) * WARPS * p.nband;
    for (int i = tid; i < WARPS * p.nband; i += (WARPS + 1) * 32) {
      fs_s[i] = p.fs[tb + i];
      le_s[i] = p.le[tb + i];
    }
  }
  __syncthreads();

  if (warp == WARPS) {
    // ---- x producer: one thread walks the bands of this column part ----
    if (lane == 0) {
      const uint64_t keep = loops::tma::policy_evict_last();
      const int col0 = qi * p.cq;
      const int part_cols = min(p.cq, p.cols - col0);
      for (int b = 0; b < p.nband; ++b) {
        const int k = b % p.xb;
        if (b >= p.xb) loops::tma::barrier_wait(&xempty[k], uint32_t((b / p.xb) - 1) & 1u);
        const int c0 = b * p.cb;
        int n = min(p.cb, part_cols - c0);
        if (n < 0) n = 0;
        const int n16 = n & ~3;
        float* dst = xs + k * p.cb;
        const float* src = p.x + col0 + c0;
        for (int t = n16; t < n; ++t) dst[t] = src[t];  // ragged tail of the last band
        if (n16 > 0) {
          loops::tma::barrier_arrive_expect_tx(&xfull[k], uint32_t(n16) * 4u);
          loops::tma::bulk_g2s_hint(dst, src, uint32_t(n16) * 4u, &xfull[k], keep);
        } else {
          loops::tma::barrier_arrive(&xfull[k]);
        }
      }
    }
  } else {
    // ---- consumer warp: its own stream, its own rows ----
    const int ws = cta * WARPS + warp;
    const int sbase = p.stream_base[ws];
    const int nsteps = p.stream_base[ws + 1] - sbase;
    uint32_t* ring = ering + warp * p.es * kStepWords;
    uint64_t* ef = efull + warp * p.es;
    const uint32_t* src = p.steps + size_t(sbase) * kStepWords;
    const uint64_t stream_policy = loops::tma::policy_evict_first();
    if (lane == 0) {
      const int pre = min(p.es, nsteps);
      for (int s = 0; s < pre; ++s) {
        loops::tma::barrier_arrive_expect_tx(&ef[s], kStepWords * 4u);
        loops::tma::bulk_g2s_hint(ring + s * kStepWords, src + size_t(s) * kStepWords, kStepWords * 4u, &ef[s],
                                  stream_policy);
      }
    }
    const uint32_t* refill = src + size_t(p.es) * kStepWords;  // next step to request
    const uint16_t* fsw = fs_s + warp * p.nband;
    const uint16_t* lew = le_s + warp * p.nband;
    constexpr int kNever = 0x7fffffff;
    // x-ring bookkeeping without divisions: slot and parity of the next band to
    // acquire / release, and the step numbers at which that happens.
    int acq = 0, rel = 0, acq_k = 0, rel_k = 0;
    uint32_t acq_par = 0;
    int nfs = p.nband > 0 ? int(fsw[0]) : kNever;  // first step that needs band `acq`
    int nle = kNever;                              // band `rel` is finished once s + 1 >= nle
    auto release_upto = [&](int limit) {           // release bands whose le <= limit
      while (nle <= limit) {
        if (lane == 0) loops::tma::barrier_arrive(&xempty[rel_k]);
        ++rel;
        if (++rel_k == p.xb) rel_k = 0;
        nle = rel < acq ? int(lew[rel]) : kNever;
      }
    };
    auto acquire_upto = [&](int s) {               // acquire every band first needed at step <= s
      while (nfs <= s) {
        loops::tma::barrier_wait(&xfull[acq_k], acq_par);
        if (rel == acq) nle = int(lew[acq]);
        ++acq;
        if (++acq_k == p.xb) { acq_k = 0; acq_par ^= 1u; }
        nfs = acq < p.nband ? int(fsw[acq]) : kNever;
        release_upto(s);
      }
    };
    int st = 0;
    uint32_t ph = 0;
    long long t_stream = 0, t_band = 0, t_gather = 0, t_fast = 0, t_slow = 0, n_slow = 0, t0 = 0, t1 = 0;
    const long long t_begin = PROFILE ? clock64() : 0;
    const int scratch = p.rb;
    for (int s = 0; s < nsteps; ++s) {
      if (PROFILE) t0 = clock64();
      loops::tma::barrier_wait(&ef[st], ph);
      if (PROFILE) { t1 = clock64(); t_stream += t1 - t0; t0 = t1; }
      uint32_t* stage = ring + st * kStepWords;
      const uint4 I = reinterpret_cast<const uint4*>(stage)[lane];
      const float4 V = reinterpret_cast<const float4*>(stage + kStep)[lane];
      acquire_upto(s);
      if (PROFILE) { t1 = clock64(); t_band += t1 - t0; t0 = t1; }
      const float p0 = __fmul_rn(V.x, xs[I.x & 0xffffu]);
      const float p1 = __fmul_rn(V.y, xs[I.y & 0xffffu]);
      const float p2 = __fmul_rn(V.z, xs[I.z & 0xffffu]);
      const float p3 = __fmul_rn(V.w, xs[I.w & 0xffffu]);
      const int r0 = int((I.x >> 16) & 0x7fffu), r1 = int((I.y >> 16) & 0x7fffu);
      const int r2 = int((I.z >> 16) & 0x7fffu), r3 = int((I.w >> 16) & 0x7fffu);
      const uint32_t fl = (I.x | I.y | I.z | I.w) & kFlagBit;
      const bool dirty = __any_sync(0xffffffffu, fl != 0u);  // also: every lane is done with the stage
      if (lane == 0 && s + p.es < nsteps) {
        loops::tma::barrier_arrive_expect_tx(&ef[st], kStepWords * 4u);
        loops::tma::bulk_g2s_hint(stage, refill, kStepWords * 4u, &ef[st], stream_policy);
      }
      refill += kStepWords;
      if (PROFILE) { t1 = clock64(); t_gather += t1 - t0; t0 = t1; }
      if (!dirty) {
        // Clean step: a row's entries are one contiguous slot range over at most
        // two adjacent lanes. Sum runs inside the lane, hand a run that started in
        // the previous lane to that lane, then update distinct y rows independently.
        const bool c1 = r1 == r0, c2 = r2 == r1, c3 = r3 == r2;
        const float v0 = p0;
        const float v1 = c1 ? v0 + p1 : p1;
        const float v2 = c2 ? v1 + p2 : p2;
        float v3 = c3 ? v2 + p3 : p3;
        bool o0 = !c1, o1 = !c2, o2 = !c3, o3 = true;   // slot closes its run
        const int prev_last = __shfl_up_sync(0xffffffffu, r3, 1);
        const bool hc = lane > 0 && prev_last == r0;     // my first run continues the previous lane's last
        const float hv = o0 ? v0 : (o1 ? v1 : (o2 ? v2 : v3));
        float recv = __shfl_down_sync(0xffffffffu, hc ? hv : 0.f, 1);
        if (lane == 31) recv = 0.f;
        if (hc) {                                        // that run is written by the previous lane
          if (o0) o0 = false; else if (o1) o1 = false; else if (o2) o2 = false; else o3 = false;
        }
        v3 += recv;
        const float y0 = ys[o0 ? r0 : scratch], y1 = ys[o1 ? r1 : scratch];
        const float y2 = ys[o2 ? r2 : scratch], y3 = ys[o3 ? r3 : scratch];
        if (o0) ys[r0] = y0 + v0;
        if (o1) ys[r1] = y1 + v1;
        if (o2) ys[r2] = y2 + v2;
        if (o3) ys[r3] = y3 + v3;
        if (PROFILE) { t1 = clock64(); t_fast += t1 - t0; t0 = t1; }
      } else {
        rmw_general(ys, r0, p0, scratch);
        rmw_general(ys, r1, p1, scratch);
        rmw_general(ys, r2, p2, scratch);
        rmw_general(ys, r3, p3, scratch);
        if (PROFILE) { t1 = clock64(); t_slow += t1 - t0; t0 = t1; ++n_slow; }
      }
      __syncwarp();   // y rows of this step are settled before the next step's loads
      release_upto(s + 1);
      if (++st == p.es) { st = 0; ph ^= 1u; }
    }
    // bands after the warp's last step: keep the ring protocol going
    acquire_upto(kNever - 1);
    release_upto(kNever - 1);
    if (PROFILE && lane == 0 && p.prof) {
      long long* o = p.prof + size_t(ws) * 8;
      o[0] = t_stream; o[1] = t_band; o[2] = t_gather; o[3] = t_fast; o[4] = t_slow; o[5] = n_slow;
      o[6] = clock64() - t_begin; o[7] = nsteps;
    }
  }
  __syncthreads();

  // ---- write-out ----
  constexpr int NT = (WARPS + 1) * 32;
  if (p.q == 1) {
    for (int i = tid; i < rows_here; i += NT) p.y[row0 + i] = ys[i];
    return;
  }
  // q > 1: publish this CTA's partial rows; the last of the q CTAs of the row
  // block to arrive adds the partials in part order and writes y. Partials are
  // padded to 4 rows per block so they move as 128-bit words.
  const int rb4 = (p.rb + 3) & ~3;
  const size_t pitch = size_t(p.nb) * size_t(rb4);
  const int n4 = (rows_here + 3) >> 2;
  float4* mine = reinterpret_cast<float4*>(p.partial + size_t(qi) * pitch + size_t(rbi) * rb4);
  const float4* ys4 = reinterpret_cast<const float4*>(ys);
  for (int i = tid; i < n4; i += NT) __stcg(mine + i, ys4[i]);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(&p.counters[rbi], 1u);
    *last_flag = (t == unsigned(p.q - 1));
  }
  __syncthreads();
  if (*last_flag) {
    __threadfence();
    const float4* part = reinterpret_cast<const float4*>(p.partial + size_t(rbi) * rb4);
    const size_t pitch4 = pitch >> 2;
    for (int i = tid; i < n4; i += NT) {
      float4 v[8];
#pragma unroll
      for (int qq = 0; qq < 8; ++qq)
        if (qq < p.q && qq != qi) v[qq] = __ldcg(part + size_t(qq) * pitch4 + i);
      float4 acc = (qi == 0) ? ys4[i] : v[0];
#pragma unroll
      for (int qq = 1; qq < 8; ++qq)
        if (qq < p.q) {
          const float4 t = (qq == qi) ? ys4[i] : v[qq];
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
      const int r = 4 * i;
      float* out = p.y + row0 + r;
      out[0] = acc.x;
      if (r + 1 < rows_here) out[1] = acc.y;
      if (r + 2 < rows_here) out[2] = acc.z;
      if (r + 3 < rows_here) out[3] = acc.w;
    }
    if (tid == 0) p.counters[rbi] = 0u;  // ready for the next launch
  }
}

// ---------------------------------------------------------------------------
// Device-resident plan data + launch
// ---------------------------------------------------------------------------
struct plan_data {
  geom g;
  uint32_t* steps = nullptr;
  int32_t* stream_base = nullptr;
  uint16_t* fs = nullptr;
  uint16_t* le = nullptr;
  float* partial = nullptr;
  unsigned* counters = nullptr;
  const void* key_indices = nullptr;  // the CSR arrays this copy was made from
  const void* key_values = nullptr;
  long long total_steps = 0, real_entries = 0, pad_entries = 0, flagged_entries = 0, flagged_steps = 0;
  long long bytes = 0;
  int smem = 0;
  long long* prof = nullptr;  // LOOPSB_DEBUG_PHASES: 8 counters per consumer warp
};

inline void destroy(plan_data* d) {
  if (!d) return;
  cudaFree(d->steps); cudaFree(d->stream_base); cudaFree(d->fs); cudaFree(d->le);
  cudaFree(d->partial); cudaFree(d->counters); cudaFree(d->prof);
  delete d;
}

using kernel_fn = void (*)(const params);
inline kernel_fn kernel_for(int warps, bool profile = false) {
  switch (warps) {
    case 4: return profile ? spmv_bt_kernel<4, true> : spmv_bt_kernel<4>;
    case 8: return profile ? spmv_bt_kernel<8, true> : spmv_bt_kernel<8>;
    case 12: return profile ? spmv_bt_kernel<12, true> : spmv_bt_kernel<12>;
    case 16: return profile ? spmv_bt_kernel<16, true> : spmv_bt_kernel<16>;
    case 20: return profile ? spmv_bt_kernel<20, true> : spmv_bt_kernel<20>;
    case 24: return profile ? spmv_bt_kernel<24, true> : spmv_bt_kernel<24>;
    default: return nullptr;
  }
}

// Default geometry for a matrix on a device with `sms` SMs and `max_smem`
// bytes of opt-in shared memory. LOOPSB_TILED_GEOM="nb,q,warps,cb,xb,es"
// overrides (0 keeps the default of that field).
inline geom choose_geom(int rows, int cols, int sms, int max_smem) {
  geom g;
  g.q = (sms % 4 == 0) ? 4 : (sms % 2 == 0 ? 2 : 1);
  g.warps = 16;
  g.xb = 4;
  g.es = 2;
  int over[6] = {0, 0, 0, 0, 0, 0};
  if (const char* e = getenv("LOOPSB_TILED_GEOM"))
    sscanf(e, "%d,%d,%d,%d,%d,%d", &over[0], &over[1], &over[2], &over[3], &over[4], &over[5]);
  if (over[1] > 0) g.q = over[1];
  if (over[2] > 0) g.warps = over[2];
  if (over[4] > 0) g.xb = over[4];
  if (over[5] > 0) g.es = over[5];
  if (over[0] > 0) {
    g.nb = over[0];
  } else {
    // whole waves of CTAs (grid = nb*q = waves * sms); as few row blocks as the
    // y rows in shared memory allow, keeping >= 48 KB for the x ring
    const int per_wave = std::max(1, sms / g.q);
    const int ring = g.warps * g.es * kStepWords * 4;
    int rb_cap = (max_smem - ring - 48 * 1024 - 4096) / 4;
    rb_cap = std::max(256, std::min(rb_cap, kMaxRowsPerBlock));
    int waves = 1;
    while (ceil_div(rows, (long long)per_wave * waves) > rb_cap) ++waves;
    g.nb = per_wave * waves;
  }
  if (over[3] > 0) {
    g.cb = over[3];
  } else {
    // widest band (<= 4096 columns) whose ring still fits beside y and the stream rings
    const int cq = std::max(4, (ceil_div(cols, g.q) + 3) & ~3);
    int cb = std::min(std::min(kMaxRingFloats / g.xb, 4096), (cq + 63) & ~63) & ~3;
    for (; cb > 64; cb -= 64) {
      geom t = g;
      t.cb = cb;
      const char* why;
      if (derive(t, rows, cols, &why) && t.smem_bytes() <= max_smem) break;
    }
    g.cb = std::max(cb, 4);
  }
  return g;
}

}  // namespace bt
}  // namespace loopsb
