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
//      bit  31      slots 1..3: zero. Slot 0: one bit of the step's 32-bit control
//                   word, bit `lane` in lane `lane` (the warp reads it with one
//                   ballot): dirty flag and the x-ring events of the step, see
//                   meta_* below
// Padding entries have value 0, point at a zero word behind the x ring and at a
// scratch row behind the y rows, so they need no predicate.
// A step only holds bands of one window of `xb` consecutive bands (the plan
// pads to a step boundary otherwise), which makes the x ring deadlock-free.
#pragma once

#include "common.cuh"

#include <loops/util/tma.hxx>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>
#include <vector>

namespace loopsb {
namespace bt {

constexpr int kLanes = 32;
constexpr int kPerLane = 4;
constexpr int kStep = kLanes * kPerLane;  // entries per step
constexpr int kStepWords = 2 * kStep;     // 256 words = 1 KB
constexpr uint32_t kFlagBit = 0x80000000u;
constexpr int kL2Ahead = 2;               // steps between the L2 prefetch and the register prefetch (swept 2..12 on B200 with PDL launches: 2 is best warm and cold, profiles/tiled_l2ahead_r02.txt)
// Step control word (ballot of bit 31 of the slot-0 ids):
//   bit 0       dirty: some row of the step is NOT one contiguous range of cells (general path)
//   bit 29      long : every row is contiguous, but some run covers three or more lanes
//               (fast path with a segmented scan over the lanes instead of one hand-off)
//   bits 1..12  lead : empty bands to acquire and release at once before the step
//   bits 13..16 plen : bands that follow, first to last band STARTING in this step
//   bits 17..24 pat  : bit i = band i of those has entries (held), 0 = empty (released at once)
//   bits 25..28 nrel : held bands whose last entry is in this step (released after it, oldest first)
constexpr uint32_t kMetaLongBit = 1u << 29;
constexpr int kMetaLeadShift = 1, kMetaLeadBits = 12, kMetaPlenShift = 13, kMetaPatShift = 17, kMetaRelShift = 25;
constexpr int kMaxRowsPerBlock = 32767;   // 15 row bits, one value kept for the scratch row
constexpr int kMaxRingFloats = 65532;     // 16 column bits, zero word behind the ring

struct geom {
  // chosen
  int nb = 0, q = 0, warps = 0, cb = 0, xb = 0, es = 0;
  int pack = 1;   // bank-aware placement of the entries inside a step (LOOPSB_TILED_PACK=0 turns it off)
  int midpoint = 1; // consumer-warp row ranges are cut at row midpoints (0 = round-1 rule: at row starts; LOOPSB_TILED_SPLIT=0)
  int quantum = 1;  // a stream's step count is rounded up to a multiple of this (1 = none; LOOPSB_TILED_ROUND=1 restores
                    // the round-1 format, whole prefetch groups of `es` steps: +1.2 % traffic, slowest warp of a CTA +1.7 %)
  // derived
  int rb = 0, rw = 0, cq = 0, nband = 0;
  int rows = 0, cols = 0;
  int nstreams() const { return nb * q * warps; }
  int grid() const { return nb * q; }
  int cta_threads() const { return (warps + 1) * 32; }
  int zero_slot() const { return xb * cb; }
  // shared-memory carve-up (bytes), identical in the kernel
  int bar_count() const { return (2 * xb + 1) & ~1; }
  int ys_words() const { return (rb + 1 + 3) & ~3; }
  int xs_words() const { return xb * cb + 4; }
  int table_bytes() const { return 16; }   // the last-arriver flag of the q-way reduction
  int smem_bytes() const {
    return bar_count() * 8 + xs_words() * 4 + ys_words() * 4 + table_bytes();
  }
};

inline int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

// Fill the derived fields; returns false (with a reason) when the format's
// field widths cannot hold the geometry.
inline bool derive(geom& g, int rows, int cols, const char** why) {
  static const char* reasons[] = {"geometry fields must be positive", "cb must be a multiple of 4",
                                  "x ring exceeds 16 bits of position", "row block exceeds 15 bits of row",
                                  "more than 31 consumer warps", "more than 4 column parts"};
  g.rows = rows; g.cols = cols;
  if (g.nb < 1 || g.q < 1 || g.warps < 1 || g.cb < 4 || g.xb < 2 || g.es < 2) { *why = reasons[0]; return false; }
  if (g.cb % 4) { *why = reasons[1]; return false; }
  if (g.xb * g.cb > kMaxRingFloats) { *why = reasons[2]; return false; }
  if (g.warps > 31) { *why = reasons[4]; return false; }
  if (g.q > 4) { *why = reasons[5]; return false; }
  if (g.xb > 8) { *why = "x ring deeper than 8 bands"; return false; }
  // rb is an UPPER BOUND here (row blocks are cut by nonzero count, so a block
  // of light rows may hold up to ~15 % more rows than rows/nb); build_host
  // replaces it by the largest block actually cut.
  const int even = std::max(1, ceil_div(rows, g.nb));
  if (even > kMaxRowsPerBlock) { *why = reasons[3]; return false; }
  g.rb = std::min(kMaxRowsPerBlock, even + even / 7 + 8);
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
  std::vector<int32_t> blk_begin;    // nb + 1: first row of every row block (cut by nonzero count)
  std::vector<int32_t> warp_begin;   // [cta][warps + 1]: block-local first row of every consumer warp
  long long total_steps = 0, real_entries = 0, pad_entries = 0, flagged_entries = 0, flagged_steps = 0;
  long long long_steps = 0;   // steps whose runs need the segmented scan (contiguous, >= 3 lanes)
};

// Static split of [0, n) over the host threads (LOOPSB_HOST_THREADS, default
// min(hardware threads, 16)); fn(begin, end, thread index).
template <typename F>
inline void parallel_for(long long n, F fn) {
  int nt = 0;
  if (const char* e = getenv("LOOPSB_HOST_THREADS")) nt = atoi(e);
  if (nt <= 0) nt = int(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
  nt = int(std::min<long long>(nt, std::max<long long>(1, n)));
  if (nt == 1) { fn(0LL, n, 0); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t)
    pool.emplace_back([=]() { fn(n * t / nt, n * (t + 1) / nt, t); });
  for (auto& th : pool) th.join();
}

// Build the image from HOST CSR arrays. Pure host code (unit-tested on CPU),
// every pass split over the host threads by rows, row blocks, streams or steps.
// Returns LOOPSB_OK / LOOPSB_ERR_UNSUPPORTED / LOOPSB_ERR_INVALID.
inline int build_host(host_image& im, geom g, int rows, int cols, const int* off, const int* idx,
                      const float* val) {
  const char* why = "";
  if (!derive(g, rows, cols, &why)) { set_error("band-tiled plan: %s", why); return LOOPSB_ERR_UNSUPPORTED; }
  const int ns = g.nstreams(), nband = g.nband;
  const long long nnz = rows > 0 ? off[rows] : 0;
  // ---- pass 0: nonzeros per (row, column part); validates the column ids ----
  std::vector<int32_t> rowpart(size_t(rows) * g.q, 0);
  std::atomic<long long> bad_atom(-1);
  parallel_for(rows, [&](long long r0, long long r1, int) {
    for (long long r = r0; r < r1; ++r)
      for (int a = off[r]; a < off[r + 1]; ++a) {
        const int c = idx[a];
        if (c < 0 || c >= cols) { bad_atom.store(a); return; }
        ++rowpart[size_t(r) * g.q + c / g.cq];
      }
  });
  if (bad_atom.load() >= 0) {
    set_error("band-tiled plan: column id %d out of range at atom %lld", idx[bad_atom.load()], bad_atom.load());
    return LOOPSB_ERR_INVALID;
  }
  // ---- row blocks: equal nonzero counts, at most g.rb rows each ----
  const int rb_cap = g.rb;
  im.blk_begin.assign(size_t(g.nb) + 1, 0);
  {
    int prev = 0;
    for (int k = 1; k <= g.nb; ++k) {
      const long long target = nnz * k / g.nb;
      int cut = int(std::lower_bound(off, off + rows + 1, target) - off);   // first row r with off[r] >= target
      const long long lo = std::max<long long>(prev, (long long)rows - (long long)(g.nb - k) * rb_cap);
      const long long hi = std::min<long long>(rows, (long long)prev + rb_cap);
      cut = int(std::min<long long>(std::max<long long>(cut, lo), hi));
      if (k == g.nb) cut = rows;
      im.blk_begin[k] = cut;
      prev = cut;
    }
  }
  int rb_max = 1;
  for (int k = 0; k < g.nb; ++k) rb_max = std::max(rb_max, im.blk_begin[k + 1] - im.blk_begin[k]);
  g.rb = rb_max;
  // ---- consumer warps: contiguous row ranges of the block with equal nonzero
  // counts inside this CTA's column part ----
  im.warp_begin.assign(size_t(g.grid()) * (g.warps + 1), 0);
  std::vector<uint8_t> wmap(size_t(rows) * g.q, 0);   // warp that owns (row, part)
  std::vector<int> rw_of_block(size_t(g.nb), 1);
  parallel_for(g.nb, [&](long long k0, long long k1, int) {
    for (int rbi = int(k0); rbi < int(k1); ++rbi) {
      const int b0 = im.blk_begin[rbi], b1 = im.blk_begin[rbi + 1];
      for (int qi = 0; qi < g.q; ++qi) {
        long long tot = 0;
        for (int r = b0; r < b1; ++r) tot += rowpart[size_t(r) * g.q + qi];
        int32_t* wb = &im.warp_begin[(size_t(rbi) * g.q + qi) * (g.warps + 1)];
        long long acc = 0;
        int w = 0;
        wb[0] = 0;
        for (int r = b0; r < b1; ++r) {
          // row r belongs to the warp whose share of the nonzeros holds the row's MIDPOINT:
          // w(r) = min(warps-1, floor((2 acc + v) warps / (2 tot))). Rows are atomic (no two warps
          // share a y row), so a boundary is off by at most half a row -- with rows of up to 256
          // entries per part that is one step of a ~74-step stream instead of two.
          const long long v = rowpart[size_t(r) * g.q + qi];
          const long long mid = g.midpoint ? v : 0;
          while (w + 1 < g.warps && (2 * acc + mid) * g.warps >= 2 * tot * (w + 1) && tot > 0) wb[++w] = r - b0;
          wmap[size_t(r) * g.q + qi] = uint8_t(w);
          acc += v;
        }
        while (w + 1 <= g.warps) wb[++w] = b1 - b0;
        for (int k = 0; k < g.warps; ++k) rw_of_block[rbi] = std::max(rw_of_block[rbi], wb[k + 1] - wb[k]);
      }
    }
  });
  g.rw = *std::max_element(rw_of_block.begin(), rw_of_block.end());
  im.g = g;
  // ---- pass 1: entries per (stream, band) ----
  std::vector<int32_t> count(size_t(ns) * nband, 0);
  parallel_for(g.nb, [&](long long k0, long long k1, int) {
    for (int rbi = int(k0); rbi < int(k1); ++rbi) {
      const size_t sb = size_t(rbi) * g.q;
      for (int r = im.blk_begin[rbi]; r < im.blk_begin[rbi + 1]; ++r)
        for (int a = off[r]; a < off[r + 1]; ++a) {
          const int c = idx[a];
          const int qi = c / g.cq, cl = c - qi * g.cq, b = cl / g.cb;
          ++count[((sb + qi) * g.warps + wmap[size_t(r) * g.q + qi]) * nband + b];
        }
    }
  });
  // ---- layout: padded start of every band in every stream ----
  std::vector<int32_t> start(size_t(ns) * nband, 0);
  std::vector<int32_t> steps_of(size_t(ns), 0);
  im.stream_base.assign(size_t(ns) + 1, 0);
  im.fs.assign(size_t(ns) * nband, 0);
  im.le.assign(size_t(ns) * nband, 0);
  std::atomic<long long> too_long(0);
  parallel_for(ns, [&](long long s0, long long s1, int) {
    std::vector<long long> end(nband, 0);
    for (long long s = s0; s < s1; ++s) {
      long long pos = 0;
      int step_min_band = -1;  // smallest band with an entry in the step that holds `pos`
      const int32_t* cnt = &count[size_t(s) * nband];
      int32_t* st = &start[size_t(s) * nband];
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
      long long nsteps = (pos + kStep - 1) / kStep;
      nsteps = (nsteps + g.quantum - 1) / g.quantum * g.quantum;   // g.quantum = es: whole prefetch groups (all-padding steps at the end)
      if (nsteps > 65535) { too_long.store(nsteps); nsteps = 0; }
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
      steps_of[size_t(s)] = int32_t(nsteps);
    }
  });
  if (too_long.load() > 0) {
    set_error("band-tiled plan: a warp stream needs %lld steps (> 65535)", too_long.load());
    return LOOPSB_ERR_UNSUPPORTED;
  }
  long long total = 0;
  for (int s = 0; s < ns; ++s) { im.stream_base[s] = int32_t(total); total += steps_of[s]; }
  im.stream_base[ns] = int32_t(total);
  if (total > 0x7fffffffLL / 2) { set_error("band-tiled plan: too many steps"); return LOOPSB_ERR_UNSUPPORTED; }
  im.total_steps = total;
  // ---- pass 2: scatter (padding pre-filled) ----
  const uint32_t pad_id = (uint32_t(g.rb) << 16) | uint32_t(g.zero_slot());
  im.steps.resize(size_t(total + g.es) * kStepWords);   // + es steps the last prefetch may touch
  parallel_for(total + g.es, [&](long long s0, long long s1, int) {
    for (long long s = s0; s < s1; ++s) {
      uint32_t* w = &im.steps[size_t(s) * kStepWords];
      for (int i = 0; i < kStep; ++i) w[i] = pad_id;
      for (int i = kStep; i < kStepWords; ++i) w[i] = 0u;
    }
  });
  std::vector<int32_t>& cursor = start;  // consumed in place
  parallel_for(g.nb, [&](long long k0, long long k1, int) {
    for (int rbi = int(k0); rbi < int(k1); ++rbi) {
      const size_t sb = size_t(rbi) * g.q;
      for (int r = im.blk_begin[rbi]; r < im.blk_begin[rbi + 1]; ++r) {
        const int lr = r - im.blk_begin[rbi];
        for (int a = off[r]; a < off[r + 1]; ++a) {
          const int c = idx[a];
          const int qi = c / g.cq, cl = c - qi * g.cq, b = cl / g.cb, lc = cl - b * g.cb;
          const size_t stream = (sb + qi) * g.warps + wmap[size_t(r) * g.q + qi];
          const int32_t p = cursor[stream * nband + b]++;
          uint32_t* sw = &im.steps[(size_t(im.stream_base[stream]) + size_t(p / kStep)) * kStepWords];
          const int slot = p % kStep;  // lane*4 + j
          sw[slot] = (uint32_t(lr) << 16) | uint32_t((b % g.xb) * g.cb + lc);
          float v = val[a];
          uint32_t vb; memcpy(&vb, &v, 4);
          sw[kStep + slot] = vb;
        }
      }
    }
  });
  im.real_entries = nnz;
  im.pad_entries = total * kStep - nnz;
  // ---- pass 2b: pack every step for the shared-memory banks. Which of the 128
  // cells (lane, slot j) of a step an entry sits in is free as long as a row's
  // entries stay together, so entries are placed to spread the 32 lanes of each
  // slot j over distinct banks of the x ring (bank = ring position mod 32) and of
  // the y rows (bank = row mod 32, counted for the cell that closes the run):
  // an instruction over 32 random banks costs ~3.4 shared-memory wavefronts, a
  // packed one ~2.2. All entries of a row inside the step become ONE run (also
  // when they come from two bands), longest runs are placed first, single
  // entries go to the slot column where their two banks are least used. ----
  if (g.pack) parallel_for(total, [&](long long step0, long long step1, int) {
    struct ent { uint32_t id, val; };
    std::vector<int64_t> stamp(size_t(g.rb) + 1, -1);
    std::vector<int32_t> unit_of(size_t(g.rb) + 1, 0);
    std::vector<ent> in(kStep), out(kStep);
    std::vector<int> unit_start, unit_len, order;
    std::vector<ent> unit_ents;
    for (long long s = step0; s < step1; ++s) {
      uint32_t* w = &im.steps[size_t(s) * kStepWords];
      // gather units (all entries of one row), preserving the stream order inside a unit
      unit_len.clear();
      int nreal = 0;
      for (int p = 0; p < kStep; ++p) {
        in[p] = ent{w[p], w[kStep + p]};
        const int lr = int((w[p] >> 16) & 0x7fff);
        if (lr == g.rb) continue;
        ++nreal;
        if (stamp[lr] != s) { stamp[lr] = s; unit_of[lr] = int(unit_len.size()); unit_len.push_back(1); }
        else ++unit_len[unit_of[lr]];
      }
      if (nreal == 0) continue;
      const int nu = int(unit_len.size());
      unit_start.assign(nu + 1, 0);
      for (int u = 0; u < nu; ++u) unit_start[u + 1] = unit_start[u] + unit_len[u];
      unit_ents.resize(nreal);
      order.assign(nu, 0);   // per-unit fill cursor first
      for (int p = 0; p < kStep; ++p) {
        const int lr = int((in[p].id >> 16) & 0x7fff);
        if (lr == g.rb) continue;
        const int u = unit_of[lr];
        unit_ents[unit_start[u] + order[u]++] = in[p];
      }
      // units by decreasing length (counting sort on min(len, 9))
      order.clear();
      for (int L = 9; L >= 1; --L)
        for (int u = 0; u < nu; ++u)
          if (std::min(unit_len[u], 9) == L) order.push_back(u);
      bool used[kLanes][kPerLane] = {};
      int free_in_col[kPerLane] = {kLanes, kLanes, kLanes, kLanes};
      int cx[kPerLane][32] = {}, cy[kPerLane][32] = {};
      for (int p = 0; p < kStep; ++p) out[p] = ent{pad_id, 0u};
      auto xbank = [](const ent& e) { return int(e.id & 31u); };
      auto ybank = [](const ent& e) { return int((e.id >> 16) & 31u); };
      auto put = [&](int lane, int j, const ent& e, bool closes) {
        used[lane][j] = true;
        --free_in_col[j];
        out[lane * kPerLane + j] = e;
        ++cx[j][xbank(e)];
        if (closes) ++cy[j][ybank(e)];
      };
      bool failed = false;
      for (int u : order) {
        const ent* e = &unit_ents[unit_start[u]];
        const int L = unit_len[u];
        if (L == 1) {
          int best = -1, best_cost = 1 << 30;
          for (int j = 0; j < kPerLane; ++j) {
            if (free_in_col[j] == 0) continue;
            const int cost = (cx[j][xbank(e[0])] + cy[j][ybank(e[0])]) * 64 - free_in_col[j];
            if (cost < best_cost) { best_cost = cost; best = j; }
          }
          if (best < 0) { failed = true; break; }
          int lane = 0;
          while (used[lane][best]) ++lane;
          put(lane, best, e[0], true);
          continue;
        }
        if (L <= kPerLane) {
          int bl = -1, bj = -1, best_cost = 1 << 30;
          for (int lane = 0; lane < kLanes; ++lane)
            for (int j = 0; j + L <= kPerLane; ++j) {
              bool ok = true;
              int cost = 0;
              for (int i = 0; i < L; ++i) { ok = ok && !used[lane][j + i]; cost += cx[j + i][xbank(e[i])]; }
              if (!ok) continue;
              cost += cy[j + L - 1][ybank(e[L - 1])];
              if (cost < best_cost) { best_cost = cost; bl = lane; bj = j; }
            }
          if (bl >= 0) {
            for (int i = 0; i < L; ++i) put(bl, bj + i, e[i], i == L - 1);
            continue;
          }
        }
        // longer runs: lane-aligned first (a run of up to 8 then covers two lanes and the
        // step stays on the one-hand-off path), else the first free contiguous range of
        // cells wherever it starts -- the kernel's segmented scan handles any range
        int c0 = -1;
        for (int p0 = 0; p0 + L <= kStep && c0 < 0; p0 += kPerLane) {
          bool ok = true;
          for (int i = 0; i < L && ok; ++i) ok = !used[(p0 + i) / kPerLane][(p0 + i) % kPerLane];
          if (ok) c0 = p0;
        }
        for (int p0 = 0; p0 + L <= kStep && c0 < 0; ++p0) {
          bool ok = true;
          for (int i = 0; i < L && ok; ++i) ok = !used[(p0 + i) / kPerLane][(p0 + i) % kPerLane];
          if (ok) c0 = p0;
        }
        if (c0 < 0) { failed = true; break; }
        for (int i = 0; i < L; ++i) put((c0 + i) / kPerLane, (c0 + i) % kPerLane, e[i], i == L - 1);
      }
      if (failed) {
        // out of room (lane alignment leaves holes): lay the units back to back in the
        // same order instead -- every row still one contiguous range, no holes, always fits
        for (int p = 0; p < kStep; ++p) out[p] = ent{pad_id, 0u};
        int cell = 0;
        for (int u : order)
          for (int i = 0; i < unit_len[u]; ++i) out[cell++] = unit_ents[unit_start[u] + i];
      }
      for (int p = 0; p < kStep; ++p) { w[p] = out[p].id; w[kStep + p] = out[p].val; }
    }
  });
  // ---- pass 3: control words. The kernel's fast path needs every row of a step
  // in ONE contiguous range of cells (cell = lane*4 + j): per-lane run sums, then
  // either one neighbour hand-off (runs over at most two lanes) or a segmented
  // scan over the lanes (bit 29: some run covers three or more lanes). A row in
  // two separate ranges marks the step dirty (bit 0, general path). The x-ring events of the step (which bands the warp must wait
  // for before it, which it is done with after it) come from the band tables. ----
  std::vector<uint32_t> meta(size_t(total), 0u);
  std::atomic<long long> flagged_entries(0), flagged_steps(0), long_steps(0);
  parallel_for(total, [&](long long step0, long long step1, int) {
    std::vector<int64_t> stamp(size_t(g.rb) + 1, -1);
    std::vector<int32_t> first(size_t(g.rb) + 1, 0), last(size_t(g.rb) + 1, 0), cnt(size_t(g.rb) + 1, 0);
    long long fe = 0, fst = 0, lst = 0;
    for (long long s = step0; s < step1; ++s) {
      const uint32_t* w = &im.steps[size_t(s) * kStepWords];
      for (int p = 0; p < kStep; ++p) {
        const int lr = int((w[p] >> 16) & 0x7fff);
        if (lr == g.rb) continue;  // padding -> scratch row, harmless
        if (stamp[lr] != s) { stamp[lr] = s; first[lr] = last[lr] = p; cnt[lr] = 1; }
        else { last[lr] = p; ++cnt[lr]; }
      }
      bool any = false, any_long = false;
      for (int p = 0; p < kStep; ++p) {
        const int lr = int((w[p] >> 16) & 0x7fff);
        if (lr == g.rb) continue;
        if (last[lr] - first[lr] + 1 != cnt[lr]) { ++fe; any = true; }                    // not contiguous
        else if (last[lr] / kPerLane - first[lr] / kPerLane >= 2) any_long = true;       // three or more lanes
      }
      if (any) { ++fst; meta[size_t(s)] |= 1u; }
      if (any_long) { ++lst; meta[size_t(s)] |= kMetaLongBit; }
    }
    flagged_entries += fe;
    flagged_steps += fst;
    long_steps += lst;
  });
  im.flagged_entries = flagged_entries.load();
  im.flagged_steps = flagged_steps.load();
  im.long_steps = long_steps.load();
  std::atomic<int> meta_overflow(0);
  parallel_for(ns, [&](long long st0, long long st1, int) {
    for (long long st = st0; st < st1; ++st) {
      const long long base = im.stream_base[st];
      const int nsteps = im.stream_base[st + 1] - im.stream_base[st];
      const int32_t* cnt = &count[size_t(st) * nband];
      const uint16_t* fs = &im.fs[size_t(st) * nband];
      const uint16_t* le = &im.le[size_t(st) * nband];
      int b = 0;
      while (b < nband) {
        const int step = fs[b];
        if (step >= nsteps) break;   // trailing empty bands: acquired and released after the last step
        // bands b .. e-1 all start (or, being empty, are borrowed) at `step`
        int e = b;
        while (e < nband && fs[e] == step) ++e;
        int lead = 0;
        while (b + lead < e && cnt[b + lead] == 0) ++lead;
        const int plen = e - (b + lead);
        if (lead >= (1 << kMetaLeadBits) || plen > 8) { meta_overflow.store(1); break; }
        uint32_t pat = 0;
        for (int i = 0; i < plen; ++i)
          if (cnt[b + lead + i] > 0) pat |= 1u << i;
        meta[size_t(base + step)] |= (uint32_t(lead) << kMetaLeadShift) | (uint32_t(plen) << kMetaPlenShift) |
                                     (pat << kMetaPatShift);
        b = e;
      }
      for (int bb = 0; bb < nband; ++bb)
        if (cnt[bb] > 0) meta[size_t(base + le[bb] - 1)] += 1u << kMetaRelShift;   // nrel <= xb <= 8 by the window rule
    }
  });
  if (meta_overflow.load()) {
    set_error("band-tiled plan: too many empty / starting bands in one step for the control word");
    return LOOPSB_ERR_UNSUPPORTED;
  }
  parallel_for(total, [&](long long step0, long long step1, int) {
    for (long long s = step0; s < step1; ++s) {
      uint32_t* w = &im.steps[size_t(s) * kStepWords];
      const uint32_t m = meta[size_t(s)];
      for (int lane = 0; lane < kLanes; ++lane)
        if ((m >> lane) & 1u) w[lane * kPerLane] |= kFlagBit;
    }
  });
  return LOOPSB_OK;
}

// ---------------------------------------------------------------------------
// Device side
// ---------------------------------------------------------------------------
struct params {
  const uint32_t* steps;
  const uint32_t* steps_end;  // one past the last word of the step buffer (L2 prefetch guard)
  const int32_t* stream_base;
  const int32_t* blk_begin;   // nb + 1 row-block boundaries
  const float* x;
  float* y;
  float* partial;       // [q][nb*rb] when q > 1
  unsigned* counters;   // [2 * nb] when q > 1 (arrivals, departures)
  int rows, cols, rb, cq, cb, xb, es, nband, q, nb;
  int l2_ahead;         // steps between the L2 prefetch and the register prefetch
  int l2_guard;         // 1: the L2 prefetch stops at the end of the warp's own stream
  int peers;            // 1: every CTA resident at once, the q CTAs of a row block share the reduction
  int pdl;              // 1: launched with programmatic stream serialization (see launch_tiled): everything that
                        //    does not depend on the previous kernel of the stream -- barrier set-up, the first
                        //    matrix-stream requests, the zeroed y tile -- runs before griddepcontrol.wait
  long long* prof;      // PROFILE builds: 8 counters per consumer warp, then 4 wall-clock stamps per CTA
};

// General y update for one slot of a dirty step: any lanes may hold the same
// row. Lanes are grouped by row (match.any); the lowest lane of each group
// updates, the rest retry, so equal rows are applied one after another in
// lane order. Padding (scratch row) is skipped.
__device__ __noinline__ void rmw_general(float* ys, int r, float p, int scratch) {
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

// One 1 KB step of a warp's stream, held in registers: ids and values of the
// lane's four entries. Loaded straight from HBM with two coalesced 128-bit
// streaming loads per lane, DEPTH steps ahead of their use.
struct step_regs {
  uint4 I;
  float4 V;
};

__device__ __forceinline__ void load_step(step_regs& r, const uint32_t* step, int lane) {
  r.I = __ldcs(reinterpret_cast<const uint4*>(step) + lane);
  r.V = __ldcs(reinterpret_cast<const float4*>(step + kStep) + lane);
}

// Long runs: what lane l receives from the lanes after it. T_l = S_l + C_l * T_{l+1}
// is what lane l hands down -- its own head-run sum S_l plus, when the whole lane is
// one run that came from above (C_l), everything that passes through it; lane l
// receives T_{l+1}. A segmented suffix scan over the 32 lanes (5 shuffle rounds).
__device__ __forceinline__ float long_run_recv(float S, bool C) {
  const int lane = threadIdx.x & 31;
  float T = S;
  bool P = C;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    const float Td = __shfl_down_sync(0xffffffffu, T, dd);
    const bool Pd = __shfl_down_sync(0xffffffffu, int(P), dd) != 0;
    if (lane + dd < 32) { if (P) T += Td; P = P & Pd; }
  }
  return __shfl_down_sync(0xffffffffu, T, 1);
}

// The two rare kinds of step, kept out of line so the hot loop stays small:
//  * dirty (a row in separate cell ranges): slot by slot through rmw_general;
//  * long runs (every row contiguous, some over three or more lanes): the fast
//    path's arithmetic with the segmented scan in place of the single hand-off.
__device__ __noinline__ void cold_update(float* ys, int r0, int r1, int r2, int r3, float p0, float p1, float p2,
                                         float p3, int scratch, bool non_contiguous) {
  if (non_contiguous) {
    rmw_general(ys, r0, p0, scratch);
    rmw_general(ys, r1, p1, scratch);
    rmw_general(ys, r2, p2, scratch);
    rmw_general(ys, r3, p3, scratch);
    return;
  }
  const int lane = threadIdx.x & 31;
  const bool c1 = r1 == r0, c2 = r2 == r1, c3 = r3 == r2;
  const float v0 = p0;
  const float v1 = c1 ? v0 + p1 : p1;
  const float v2 = c2 ? v1 + p2 : p2;
  float v3 = c3 ? v2 + p3 : p3;
  const int prev_last = __shfl_up_sync(0xffffffffu, r3, 1);
  const bool hc = (lane > 0) & (prev_last == r0);
  const bool h0 = !c1, h1 = c1 & !c2, h2 = c1 & c2 & !c3, h3 = c1 & c2 & c3;
  const float hv = h0 ? v0 : (h1 ? v1 : (h2 ? v2 : v3));
  float recv = long_run_recv(hc ? hv : 0.f, hc & h3);
  recv = lane == 31 ? 0.f : recv;
  v3 += recv;
  const int a0 = (!c1 & !(hc & h0)) ? r0 : scratch;
  const int a1 = (!c2 & !(hc & h1)) ? r1 : scratch;
  const int a2 = (!c3 & !(hc & h2)) ? r2 : scratch;
  const int a3 = !(hc & h3) ? r3 : scratch;
  const float y0 = ys[a0], y1 = ys[a1], y2 = ys[a2], y3 = ys[a3];
  ys[a0] = y0 + v0;
  ys[a1] = y1 + v1;
  ys[a2] = y2 + v2;
  ys[a3] = y3 + v3;
}

#ifndef LOOPSB_STAMPS_ONLY
#define LOOPSB_STAMPS_ONLY 0   // 1: PROFILE builds keep only wall-clock stamps (no per-step cycle counters)
#endif
template <int WARPS, int DEPTH, bool PROFILE = false>
__global__ void __launch_bounds__((WARPS + 1) * 32, 1) spmv_bt_kernel(const params p) {
  constexpr bool COUNT = PROFILE && !LOOPSB_STAMPS_ONLY;   // per-step cycle counters (slow the loop ~2x)
  constexpr bool STAMP = PROFILE && LOOPSB_STAMPS_ONLY;    // per-warp wall-clock stamps only
  extern __shared__ __align__(16) unsigned char bt_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int rbi = cta / p.q, qi = cta - rbi * p.q;
  const int row0 = p.blk_begin[rbi];
  const int rows_here = p.blk_begin[rbi + 1] - row0;
  if (rows_here <= 0) return;  // empty row block: nothing to write
  auto wall_ns = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  long long* stamps = (PROFILE && p.prof) ? p.prof + size_t(gridDim.x) * WARPS * 8 + size_t(cta) * 4 : nullptr;
  // One-thread work (barrier set-up, stamps) is given to the producer warp, so no
  // consumer warp ever enters its loop with a lane on a different control path.
  const bool boss = tid == WARPS * 32;
  if (PROFILE && stamps && boss) stamps[0] = wall_ns();

  // ---- carve shared memory (geom::smem_bytes order) ----
  uint64_t* xfull = reinterpret_cast<uint64_t*>(bt_smem);
  uint64_t* xempty = xfull + p.xb;
  const int nbar = (2 * p.xb + 1) & ~1;
  float* xs = reinterpret_cast<float*>(xfull + nbar);
  float* ys = xs + (p.xb * p.cb + 4);
  const int ys_words = (p.rb + 1 + 3) & ~3;
  int* last_flag = reinterpret_cast<int*>(ys + ys_words);

  // The kernel is one wave of CTAs that all start together, so whatever precedes the
  // first stream data is time HBM sits idle. Order of the prologue: the producer
  // thread arms the ring and requests the first xb bands; every consumer warp requests
  // the first DEPTH steps of its stream (registers) and the steps behind them (L2);
  // only then is the y tile zeroed, under those loads.
  const uint64_t keep = loops::tma::policy_evict_last();
  const int col0 = qi * p.cq;
  const int part_cols = min(p.cq, p.cols - col0);
  auto request_band = [&](int b, int k) {   // producer thread only
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
  };
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;");   // the next kernel of the stream may start filling freed SMs
  if (boss) {
    for (int k = 0; k < p.xb; ++k) {
      loops::tma::barrier_init(&xfull[k], 1);
      loops::tma::barrier_init(&xempty[k], WARPS);
    }
    if (!p.pdl) for (int b = 0; b < min(p.xb, p.nband); ++b) request_band(b, b);
  }
  const bool consumer = warp < WARPS;
  const int ws = consumer ? cta * WARPS + warp : 0;
  const int sbase = p.stream_base[ws];
  const int nsteps = consumer ? p.stream_base[ws + 1] - sbase : 0;
  const uint32_t* src = p.steps + size_t(sbase) * kStepWords;
  step_regs buf[DEPTH];
  if (consumer) {
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) load_step(buf[k], src + size_t(k) * kStepWords, lane);
    // steps DEPTH .. DEPTH + l2_ahead - 1: 128-byte lines, four steps per pass of the warp
    for (int t = 0; t < p.l2_ahead; t += 4) {
      const uint32_t* far = src + size_t(DEPTH + t) * kStepWords + lane * 32;
      if (lane < 8 * (p.l2_ahead - t) && far < p.steps_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(far));
    }
  }
  for (int i = tid; i < ys_words; i += (WARPS + 1) * 32) ys[i] = 0.f;
  if (tid < 4) xs[p.xb * p.cb + tid] = 0.f;
  if (p.pdl && boss) {
    // x (and y, the partial rows, the counters) belong to the previous kernel of the stream until it
    // has completed: the producer thread waits for that here, every other thread is ordered behind
    // it by the barrier below and by the x ring (no consumer passes its first band before this).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int b = 0; b < min(p.xb, p.nband); ++b) request_band(b, b);
  }
  __syncthreads();
  if (PROFILE && stamps && boss) stamps[1] = wall_ns();

  if (warp == WARPS) {
    // ---- x producer: one thread walks the remaining bands of this column part ----
    if (lane == 0) {
      int k = 0;
      uint32_t par = 0;  // parity of the PREVIOUS use of slot k
      for (int b = p.xb; b < p.nband; ++b) {
        loops::tma::barrier_wait_suspend(&xempty[k], par);
        request_band(b, k);
        if (++k == p.xb) { k = 0; par ^= 1u; }
      }
    }
  } else {
    // ---- consumer warp: its own stream, its own rows ----
    __syncwarp();
    const uint32_t* refill = src + size_t(DEPTH) * kStepWords;  // next step to request
    const size_t l2_ahead_words = size_t(p.l2_ahead) * kStepWords;
    const uint32_t* l2_end = p.l2_guard ? src + size_t(nsteps) * kStepWords : p.steps_end;   // A/B: stop the L2 prefetch at the stream's end
    // x-ring bookkeeping: ring slot / parity of the next band to acquire, and a
    // FIFO (3 bits per entry) of the slots of the bands this warp still holds.
    // What to do at each step comes from the step's control word.
    int acq = 0, acq_k = 0, nheld = 0;
    uint32_t acq_par = 0, held = 0;
    auto advance = [&]() {
      ++acq;
      if (++acq_k == p.xb) { acq_k = 0; acq_par ^= 1u; }
    };
    auto band_events = [&](uint32_t m) {   // before the step's x gathers
      const int lead = int((m >> kMetaLeadShift) & ((1u << kMetaLeadBits) - 1u));
      const int plen = int((m >> kMetaPlenShift) & 0xfu);
      const uint32_t pat = (m >> kMetaPatShift) & 0xffu;
#pragma unroll 1
      for (int i = 0; i < lead; ++i) {      // bands without entries for this warp
        loops::tma::barrier_wait_suspend(&xfull[acq_k], acq_par);
        if (lane == 0) loops::tma::barrier_arrive(&xempty[acq_k]);
        advance();
      }
#pragma unroll 1
      for (int i = 0; i < plen; ++i) {
        loops::tma::barrier_wait_suspend(&xfull[acq_k], acq_par);
        if ((pat >> i) & 1u) { held |= uint32_t(acq_k) << (3 * nheld); ++nheld; }
        else if (lane == 0) loops::tma::barrier_arrive(&xempty[acq_k]);
        advance();
      }
    };
    auto release = [&](int n) {             // after the step: its x values are in registers
#pragma unroll 1
      for (int i = 0; i < n; ++i) {
        if (lane == 0) loops::tma::barrier_arrive(&xempty[held & 7u]);
        held >>= 3;
        --nheld;
      }
    };
    long long w_first = 0;
    long long t_stream = 0, t_band = 0, t_gather = 0, t_rmw = 0, n_slow = 0, t0 = 0;
    const long long t_begin = COUNT ? clock64() : 0;
    const int scratch = p.rb;
    // (Issuing the x gathers of step s+1 ahead of the y updates of step s was
    // tried and measured slower: a warp that then blocks on the x ring sits on
    // y updates it could have retired. Steps are processed one at a time.)
    // DEPTH spare steps follow the last stream, so prefetches need no guard. A stream's
    // step count need not be a multiple of DEPTH: the last group stops at the stream's
    // end (warp-uniform test); what its unused registers hold is the head of the next stream.
    for (int s0 = 0; s0 < nsteps; s0 += DEPTH) {
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) {
        if (s0 + k >= nsteps) break;
        // (the step is used in place and its registers are re-loaded only when the
        // step is done: copying them out first made ptxas rotate registers with
        // moves that wait on the load just issued -- a synchronous "prefetch")
        const uint4 I = buf[k].I;
        const float4 V = buf[k].V;
        if (COUNT) t0 = clock64();
        const uint32_t m = __ballot_sync(0xffffffffu, (I.x & kFlagBit) != 0u);   // the step's control word
        if (STAMP && s0 == 0 && k == 0) w_first = wall_ns();
        if (COUNT) { const long long t1 = clock64(); t_stream += t1 - t0; t0 = t1; }
        if (m >> kMetaLeadShift) band_events(m);
        if (COUNT) { const long long t1 = clock64(); t_band += t1 - t0; t0 = t1; }
        const float p0 = __fmul_rn(V.x, xs[I.x & 0xffffu]);
        const float p1 = __fmul_rn(V.y, xs[I.y & 0xffffu]);
        const float p2 = __fmul_rn(V.z, xs[I.z & 0xffffu]);
        const float p3 = __fmul_rn(V.w, xs[I.w & 0xffffu]);
        const int r0 = int((I.x >> 16) & 0x7fffu), r1 = int((I.y >> 16) & 0x7fffu);
        const int r2 = int((I.z >> 16) & 0x7fffu), r3 = int((I.w >> 16) & 0x7fffu);
        const bool dirty = (m & (1u | kMetaLongBit)) != 0u;   // anything but the one-hand-off fast path
        if (COUNT) { const long long t1 = clock64(); t_gather += (p0 + p1 + p2 + p3 == 12345.f) ? 1 : t1 - t0; t0 = t1; }
        if (!dirty) {
          // Every row of the step is one contiguous range of cells. Sum runs inside the
          // lane, hand the part of a run that lies in later lanes back to the lane where
          // the run starts, then update distinct y rows independently.
          const bool c1 = r1 == r0, c2 = r2 == r1, c3 = r3 == r2;
          const float v0 = p0;
          const float v1 = c1 ? v0 + p1 : p1;
          const float v2 = c2 ? v1 + p2 : p2;
          float v3 = c3 ? v2 + p3 : p3;
          const int prev_last = __shfl_up_sync(0xffffffffu, r3, 1);
          const bool hc = (lane > 0) & (prev_last == r0);  // my first run continues the previous lane's last
          // slot that closes the lane's FIRST run (exactly one of h0..h3 is set)
          const bool h0 = !c1, h1 = c1 & !c2, h2 = c1 & c2 & !c3, h3 = c1 & c2 & c3;
          const float hv = h0 ? v0 : (h1 ? v1 : (h2 ? v2 : v3));
          float recv = __shfl_down_sync(0xffffffffu, hc ? hv : 0.f, 1);   // runs end in the next lane at the latest
          recv = lane == 31 ? 0.f : recv;
          v3 += recv;
          // a slot writes y when it closes a run that this lane owns; everything
          // else is pointed at the scratch row (branch-free, no predicates)
          const int a0 = (!c1 & !(hc & h0)) ? r0 : scratch;
          const int a1 = (!c2 & !(hc & h1)) ? r1 : scratch;
          const int a2 = (!c3 & !(hc & h2)) ? r2 : scratch;
          const int a3 = !(hc & h3) ? r3 : scratch;
          const float y0 = ys[a0], y1 = ys[a1], y2 = ys[a2], y3 = ys[a3];
          ys[a0] = y0 + v0;
          ys[a1] = y1 + v1;
          ys[a2] = y2 + v2;
          ys[a3] = y3 + v3;
        } else {
          cold_update(ys, r0, r1, r2, r3, p0, p1, p2, p3, scratch, (m & 1u) != 0u);
          if (COUNT) ++n_slow;
        }
        if (COUNT) t_rmw += clock64() - t0;
        __syncwarp();   // y rows of this step are settled before the next step's loads
        release(int(m >> kMetaRelShift) & 0xf);
        // step s + DEPTH into the registers just freed, and step s + DEPTH + kL2Ahead
        // from HBM into L2 (eight 128-byte lines, one per lane 0..7) so that the
        // register prefetch only pays an L2 hit. (Stopping them at the end of the warp's
        // own stream was measured twice -- both prefetches, and the L2 one alone via
        // LOOPSB_TILED_L2GUARD=1: ~3 % slower either way. Running a few steps into the
        // neighbour's stream leaves the head of the streams in L2 for the next launch.)
        load_step(buf[k], refill, lane);
        {
          const uint32_t* far = refill + l2_ahead_words + lane * 32;
          if (lane < 8 && far < l2_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(far));
        }
        refill += kStepWords;
      }
    }
    // bands after the warp's last step (none of them has entries for this warp):
    // keep the ring protocol going
    release(nheld);
    while (acq < p.nband) {
      loops::tma::barrier_wait_suspend(&xfull[acq_k], acq_par);
      if (lane == 0) loops::tma::barrier_arrive(&xempty[acq_k]);
      advance();
    }
    if (STAMP && lane == 0 && p.prof) {
      long long* o = p.prof + size_t(ws) * 8;
      o[0] = w_first; o[1] = wall_ns(); o[7] = nsteps;
    }
    if (COUNT && lane == 0 && p.prof) {
      long long* o = p.prof + size_t(ws) * 8;
      o[0] = t_stream; o[1] = t_band; o[2] = t_gather; o[3] = t_rmw; o[4] = 0; o[5] = n_slow;
      o[6] = clock64() - t_begin; o[7] = nsteps;
    }
  }
  __syncthreads();
  if (PROFILE && stamps && boss) stamps[2] = wall_ns();

  // ---- write-out ----
  constexpr int NT = (WARPS + 1) * 32;
  if (p.q == 1) {
    for (int i = tid; i < rows_here; i += NT) p.y[row0 + i] = ys[i];
    return;
  }
  // q > 1: publish this CTA's partial rows, then add the q partials of the row
  // block in part order. Partials are padded to 4 rows per block so they move as
  // 128-bit words. Two protocols:
  //  * peers (cooperative launch, every CTA resident): each of the q CTAs waits
  //    until all q partials are published and reduces its own 1/q of the rows --
  //    the reduction is latency-bound (L2 round trips), so four CTAs doing one
  //    trip each beat one CTA doing them all;
  //  * last arriver (any grid): the last CTA to publish reduces the whole block.
  const int rb4 = (p.rb + 3) & ~3;
  const size_t pitch = size_t(p.nb) * size_t(rb4);
  const int n4 = (rows_here + 3) >> 2;
  float4* mine = reinterpret_cast<float4*>(p.partial + size_t(qi) * pitch + size_t(rbi) * rb4);
  const float4* ys4 = reinterpret_cast<const float4*>(ys);
  int lo = 0, hi = n4;
  if (p.peers) {   // the rows this CTA reduces itself never leave its shared memory
    const int chunk = (n4 + p.q - 1) / p.q;
    lo = min(n4, qi * chunk);
    hi = min(n4, lo + chunk);
    for (int i = tid; i < n4 - (hi - lo); i += NT) {
      const int j = i < lo ? i : i + (hi - lo);
      __stcg(mine + j, ys4[j]);
    }
  } else {
    for (int i = tid; i < n4; i += NT) __stcg(mine + i, ys4[i]);
  }
  // release: the CTA's stores are ordered before thread 0's fence by the barrier, and
  // the fence (cumulative, device scope) orders them before the counter update
  __syncthreads();
  if (STAMP && tid == 0 && p.prof) p.prof[size_t(cta) * WARPS * 8 + 3] = wall_ns();   // partial rows stored
  if (p.peers) {
    if (tid == 0) {
      __threadfence();
      atomicAdd(&p.counters[rbi], 1u);
      volatile unsigned* c = p.counters + rbi;
      while (*c < unsigned(p.q)) __nanosleep(32);
      __threadfence();   // acquire: the peers' partials are read after the count was seen
      if (STAMP && p.prof) p.prof[size_t(cta) * WARPS * 8 + 2] = wall_ns();             // all q partials visible
    }
    __syncthreads();
  } else {
    if (tid == 0) {
      __threadfence();
      const unsigned t = atomicAdd(&p.counters[rbi], 1u);
      *last_flag = (t == unsigned(p.q - 1));
      __threadfence();
    }
    __syncthreads();
    if (!*last_flag) hi = 0;
  }
  if (hi > lo) {
    const float4* part = reinterpret_cast<const float4*>(p.partial + size_t(rbi) * rb4);
    const size_t pitch4 = pitch >> 2;
    // three row groups per thread and trip, all their partial loads issued
    // before the first add, so the L2 round trips overlap
    constexpr int U = 3;
    for (int i0 = lo + tid; i0 < hi; i0 += NT * U) {
      float4 v[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NT;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          if (qq < p.q && qq != qi && i < hi) v[u][qq] = __ldcg(part + size_t(qq) * pitch4 + i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NT;
        if (i >= hi) break;
        float4 acc = (qi == 0) ? ys4[i] : v[u][0];
#pragma unroll
        for (int qq = 1; qq < 4; ++qq)
          if (qq < p.q) {
            const float4 t = (qq == qi) ? ys4[i] : v[u][qq];
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
        const int r = 4 * i;
        float* out = p.y + row0 + r;
        out[0] = acc.x;
        if (r + 1 < rows_here) out[1] = acc.y;
        if (r + 2 < rows_here) out[2] = acc.z;
        if (r + 3 < rows_here) out[3] = acc.w;
      }
    }
  }
  // re-arm the counters for the next launch
  if (p.peers) {
    __syncthreads();   // every thread of this CTA has read the peers' partials
    if (tid == 0) {
      const unsigned d = atomicAdd(&p.counters[p.nb + rbi], 1u);
      if (d == unsigned(p.q - 1)) { p.counters[rbi] = 0u; p.counters[p.nb + rbi] = 0u; }
    }
  } else if (hi > lo && tid == 0) {
    p.counters[rbi] = 0u;
  }
  if (PROFILE && stamps) { __syncthreads(); if (boss) stamps[3] = wall_ns(); }
}

// ---------------------------------------------------------------------------
// Device-resident plan data + launch
// ---------------------------------------------------------------------------
struct plan_data {
  geom g;
  uint32_t* steps = nullptr;
  int32_t* stream_base = nullptr;
  int32_t* blk_begin = nullptr;
  float* partial = nullptr;
  unsigned* counters = nullptr;
  const void* key_indices = nullptr;  // the CSR arrays this copy was made from
  const void* key_values = nullptr;
  long long total_steps = 0, real_entries = 0, pad_entries = 0, flagged_entries = 0, flagged_steps = 0;
  long long long_steps = 0;
  long long bytes = 0;
  int smem = 0;
  int peers = 0;   // the grid fits the device in one wave: launch cooperatively, share the q-way reduction
  bool launched = false;   // a first launch on this copy has been enqueued (launch_tiled: later ones may use PDL)
  long long* prof = nullptr;  // LOOPSB_DEBUG_PHASES: 8 counters per consumer warp
};

inline void destroy(plan_data* d) {
  if (!d) return;
  cudaFree(d->steps); cudaFree(d->stream_base); cudaFree(d->blk_begin); 
  cudaFree(d->partial); cudaFree(d->counters); cudaFree(d->prof);
  delete d;
}

using kernel_fn = void (*)(const params);
template <int WARPS>
inline kernel_fn kernel_for_depth(int depth, bool profile) {
  switch (depth) {
    case 2: return profile ? spmv_bt_kernel<WARPS, 2, true> : spmv_bt_kernel<WARPS, 2>;
    case 3: return profile ? spmv_bt_kernel<WARPS, 3, true> : spmv_bt_kernel<WARPS, 3>;
    case 4: return profile ? spmv_bt_kernel<WARPS, 4, true> : spmv_bt_kernel<WARPS, 4>;
    default: return nullptr;
  }
}
// warps: consumer warps per CTA; depth (geom::es): steps prefetched in registers.
inline kernel_fn kernel_for(int warps, int depth, bool profile = false) {
  switch (warps) {
    case 4: return kernel_for_depth<4>(depth, profile);
    case 8: return kernel_for_depth<8>(depth, profile);
    case 12: return kernel_for_depth<12>(depth, profile);
    case 16: return kernel_for_depth<16>(depth, profile);
    case 20: return kernel_for_depth<20>(depth, profile);
    case 24: return kernel_for_depth<24>(depth, profile);
    default: return nullptr;
  }
}

// Default geometry for a matrix on a device with `sms` SMs and `max_smem`
// bytes of opt-in shared memory. LOOPSB_TILED_GEOM="nb,q,warps,cb,xb,es"
// overrides (0 keeps the default of that field).
inline geom choose_geom(int rows, int cols, int sms, int max_smem) {
  geom g;
  g.q = (sms % 4 == 0) ? 4 : (sms % 2 == 0 ? 2 : 1);
  // measured best on B200 for BASELINE config 2 (tools/tiled_sweep.py, profiles/):
  // 24 consumer warps, 3 x 7168-column bands in the x ring, 3 steps prefetched
  g.warps = 24;
  g.xb = 3;
  g.es = 3;
  int over[6] = {0, 0, 0, 0, 0, 0};
  if (const char* e = getenv("LOOPSB_TILED_GEOM"))
    sscanf(e, "%d,%d,%d,%d,%d,%d", &over[0], &over[1], &over[2], &over[3], &over[4], &over[5]);
  if (over[1] > 0) g.q = over[1];
  if (over[2] > 0) g.warps = over[2];
  if (over[4] > 0) g.xb = over[4];
  if (over[5] > 0) g.es = over[5];
  if (const char* e = getenv("LOOPSB_TILED_PACK")) g.pack = atoi(e) != 0;
  if (const char* e = getenv("LOOPSB_TILED_ROUND")) g.quantum = atoi(e) != 0 ? g.es : 1;
  if (const char* e = getenv("LOOPSB_TILED_SPLIT")) g.midpoint = atoi(e) != 0;
  if (over[0] > 0) {
    g.nb = over[0];
  } else {
    // whole waves of CTAs (grid = nb*q = waves * sms); as few row blocks as the
    // y rows in shared memory allow, keeping >= 64 KB for the x ring
    const int per_wave = std::max(1, sms / g.q);
    int rb_cap = (max_smem - 64 * 1024 - 8192) / 4;
    rb_cap = std::max(256, std::min(rb_cap, kMaxRowsPerBlock));
    int waves = 1;
    while (ceil_div(rows, (long long)per_wave * waves) > rb_cap) ++waves;
    g.nb = per_wave * waves;
  }
  if (over[3] > 0) {
    g.cb = over[3];
  } else {
    // widest band (<= 7168 columns) whose ring still fits beside the y rows
    const int cq = std::max(4, (ceil_div(cols, g.q) + 3) & ~3);
    int cb = std::min(std::min(kMaxRingFloats / g.xb, 7168), (cq + 63) & ~63) & ~3;
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
