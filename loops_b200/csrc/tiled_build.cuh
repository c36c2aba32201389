// loops_b200/csrc/tiled_build.cuh -- device-side builder of the band-tiled copy
// (spmv_tiled.cuh). Same algorithm, same tie-breaks and therefore the SAME IMAGE,
// byte for byte, as bt::build_host (tests/test_gpu_tiled_build.py compares them):
//
//   rowpart     nonzeros per (row, column part)                 thread per row
//   row blocks  equal-nnz cuts of the row offsets               host, on a 4-byte/row copy of offsets
//   warp map    equal-nnz row ranges of the consumer warps      block per (row block, part): reduce + scan
//   keys        key(atom) = (stream, band), counts per key      thread per row, global atomics
//   sort        stable radix sort of (key, atom)                cub::DeviceRadixSort
//   layout      padded start of every band in every stream,     thread per stream (sequential over bands,
//               band tables, steps per stream                    exactly the host loop)
//   scatter     atom -> (step, cell) in CSR order inside a band thread per sorted atom
//   pack        bank-aware placement inside every step, dirty   one thread per step, arrays interleaved in
//               flag, control word                              shared memory ([i][thread]), host loop verbatim
//
// The matrix never leaves the device; only the row offsets (4 bytes per row) and
// a few counters cross PCIe.
#pragma once

#include "spmv_tiled.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace loopsb {
namespace bt {
namespace dev {

// ---- geometry constants handed to the kernels ----
struct gconst {
  int rows, cols, q, warps, cb, xb, es, cq, nband, nb, rb, ns, quantum, midpoint;
};

__global__ void rowpart_kernel(const int* __restrict__ off, const int* __restrict__ idx, int rows, int cols, int q,
                               int cq, int* __restrict__ rowpart, int* __restrict__ bad_atom) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int cnt[4] = {0, 0, 0, 0};
  for (int a = off[r]; a < off[r + 1]; ++a) {
    const int c = idx[a];
    if (c < 0 || c >= cols) { atomicMin(bad_atom, a); continue; }
    ++cnt[c / cq];
  }
  for (int k = 0; k < q; ++k) rowpart[(long long)r * q + k] = cnt[k];
}

// Block per (row block, part): warp that owns every row of the block inside this
// part -- w(r) = min(warps-1, floor((2 acc_before(r) + v(r)) * warps / (2 tot))), the closed
// form of the host's "the warp whose share holds the row's midpoint" loop.
__global__ void __launch_bounds__(256)
    warp_map_kernel(const int* __restrict__ rowpart, const int* __restrict__ blk_begin, gconst g,
                    unsigned char* __restrict__ wmap) {
  const int rbi = blockIdx.x / g.q, qi = blockIdx.x % g.q;
  const int b0 = blk_begin[rbi], b1 = blk_begin[rbi + 1];
  __shared__ long long s_part[256];
  __shared__ long long s_run;
  // total of the block
  long long mine = 0;
  for (int r = b0 + threadIdx.x; r < b1; r += 256) mine += rowpart[(long long)r * g.q + qi];
  s_part[threadIdx.x] = mine;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) s_part[threadIdx.x] += s_part[threadIdx.x + st];
    __syncthreads();
  }
  const long long tot = s_part[0];
  __syncthreads();
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  // chunked exclusive scan over the rows of the block
  for (int base = b0; base < b1; base += 256) {
    const int r = base + threadIdx.x;
    const long long v = r < b1 ? rowpart[(long long)r * g.q + qi] : 0;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int st = 1; st < 256; st <<= 1) {   // Hillis-Steele inclusive scan
      const long long add = threadIdx.x >= st ? s_part[threadIdx.x - st] : 0;
      __syncthreads();
      s_part[threadIdx.x] += add;
      __syncthreads();
    }
    const long long before = s_run + s_part[threadIdx.x] - v;
    if (r < b1) {
      int w = 0;
      if (tot > 0) {
        const long long k = (2 * before + (g.midpoint ? v : 0)) * g.warps / (2 * tot);   // the row's midpoint (host loop's closed form)
        w = int(k < g.warps - 1 ? k : g.warps - 1);
      }
      wmap[(long long)r * g.q + qi] = (unsigned char)w;
    }
    __syncthreads();
    if (threadIdx.x == 255) s_run += s_part[255];
    __syncthreads();
  }
}

__device__ __forceinline__ int block_of_row(const int* __restrict__ blk_begin, int nb, int r) {
  int lo = 0, hi = nb;   // blk_begin[lo] <= r < blk_begin[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (blk_begin[mid] <= r) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void key_kernel(const int* __restrict__ off, const int* __restrict__ idx, const int* __restrict__ blk_begin,
                           const unsigned char* __restrict__ wmap, gconst g, unsigned* __restrict__ keys,
                           unsigned* __restrict__ atoms, int* __restrict__ count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= g.rows) return;
  const int rbi = block_of_row(blk_begin, g.nb, r);
  for (int a = off[r]; a < off[r + 1]; ++a) {
    const int c = idx[a];
    const int qi = c / g.cq, cl = c - qi * g.cq, b = cl / g.cb;
    const int w = wmap[(long long)r * g.q + qi];
    const unsigned key = unsigned(((rbi * g.q + qi) * g.warps + w) * g.nband + b);
    keys[a] = key;
    atoms[a] = unsigned(a);
    atomicAdd(count + key, 1);
  }
}

// Thread per stream: the host layout loop verbatim.
__global__ void layout_kernel(const int* __restrict__ count, gconst g, int* __restrict__ start,
                              unsigned short* __restrict__ fs, unsigned short* __restrict__ le,
                              int* __restrict__ steps_of, int* __restrict__ too_long) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= g.ns) return;
  const int* cnt = count + (long long)s * g.nband;
  int* st = start + (long long)s * g.nband;
  long long pos = 0;
  int step_min_band = -1;
  for (int b = 0; b < g.nband; ++b) {
    if (cnt[b] == 0) { st[b] = int(pos); continue; }
    if (pos % kStep != 0 && b > step_min_band + g.xb - 1) pos = (pos + kStep - 1) / kStep * kStep;
    if (pos % kStep == 0) step_min_band = b;
    st[b] = int(pos);
    const long long e = pos + cnt[b];
    if ((e - 1) / kStep > pos / kStep) step_min_band = b;
    pos = e;
  }
  long long nsteps = (pos + kStep - 1) / kStep;
  nsteps = (nsteps + g.quantum - 1) / g.quantum * g.quantum;
  if (nsteps > 65535) { atomicMax(too_long, 1); nsteps = 0; }
  int next_fs = int(nsteps);
  for (int b = g.nband - 1; b >= 0; --b) {
    unsigned short f, l;
    if (cnt[b] > 0) {
      next_fs = st[b] / kStep;
      f = (unsigned short)next_fs;
      l = (unsigned short)((st[b] + cnt[b] - 1) / kStep + 1);
    } else {
      f = (unsigned short)next_fs;
      l = (unsigned short)next_fs;
    }
    fs[(long long)s * g.nband + b] = f;
    le[(long long)s * g.nband + b] = l;
  }
  steps_of[s] = int(nsteps);
}

__global__ void fill_pad_kernel(uint32_t* __restrict__ steps, long long total_steps, uint32_t pad_id) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per word
  if (i >= total_steps * kStepWords) return;
  steps[i] = (i % kStepWords) < kStep ? pad_id : 0u;
}

__global__ void scatter_kernel(const unsigned* __restrict__ skeys, const unsigned* __restrict__ satoms, long long nnz,
                               const int* __restrict__ off, const int* __restrict__ idx, const float* __restrict__ val,
                               const int* __restrict__ blk_begin, const int* __restrict__ start,
                               const int* __restrict__ keyfirst, const int* __restrict__ stream_base, gconst g,
                               uint32_t* __restrict__ steps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const unsigned key = skeys[i];
  const int a = int(satoms[i]);
  const int stream = int(key / unsigned(g.nband)), b = int(key % unsigned(g.nband));
  const int cta = stream / g.warps, rbi = cta / g.q, qi = cta % g.q;
  // row of atom a: last row with off[row] <= a
  int lo = blk_begin[rbi], hi = blk_begin[rbi + 1];
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= a) lo = mid; else hi = mid;
  }
  const int lr = lo - blk_begin[rbi];
  const int c = idx[a];
  const int lc = (c - qi * g.cq) - b * g.cb;
  const int p = start[key] + int(i - keyfirst[key]);
  uint32_t* sw = steps + ((long long)stream_base[stream] + p / kStep) * kStepWords;
  const int slot = p % kStep;
  sw[slot] = (uint32_t(lr) << 16) | uint32_t((b % g.xb) * g.cb + lc);
  sw[kStep + slot] = __float_as_uint(val[a]);
}

// Thread per stream: x-ring events of every step into meta[] (host loop verbatim).
__global__ void meta_kernel(const int* __restrict__ count, const unsigned short* __restrict__ fs,
                            const unsigned short* __restrict__ le, const int* __restrict__ stream_base, gconst g,
                            uint32_t* __restrict__ meta, int* __restrict__ overflow) {
  const int st = blockIdx.x * blockDim.x + threadIdx.x;
  if (st >= g.ns) return;
  const long long base = stream_base[st];
  const int nsteps = stream_base[st + 1] - stream_base[st];
  const int* cnt = count + (long long)st * g.nband;
  const unsigned short* f = fs + (long long)st * g.nband;
  const unsigned short* l = le + (long long)st * g.nband;
  int b = 0;
  while (b < g.nband) {
    const int step = f[b];
    if (step >= nsteps) break;
    int e = b;
    while (e < g.nband && f[e] == step) ++e;
    int lead = 0;
    while (b + lead < e && cnt[b + lead] == 0) ++lead;
    const int plen = e - (b + lead);
    if (lead >= (1 << kMetaLeadBits) || plen > 8) { atomicMax(overflow, 1); break; }
    uint32_t pat = 0;
    for (int i = 0; i < plen; ++i)
      if (cnt[b + lead + i] > 0) pat |= 1u << i;
    meta[base + step] |= (uint32_t(lead) << kMetaLeadShift) | (uint32_t(plen) << kMetaPlenShift) | (pat << kMetaPatShift);
    b = e;
  }
  for (int bb = 0; bb < g.nband; ++bb)
    if (cnt[bb] > 0) meta[base + l[bb] - 1] += 1u << kMetaRelShift;
}

// ---- pack + dirty + control-word bits: one thread per step -----------------
// All per-thread arrays live in shared memory as [index][thread] so that threads
// working on the same index hit different banks.
constexpr int kPackThreads = 32;
struct pack_smem {
  uint32_t in_id[kStep][kPackThreads];
  uint32_t in_val[kStep][kPackThreads];
  uint32_t out_id[kStep][kPackThreads];
  uint32_t out_val[kStep][kPackThreads];
  unsigned char unit_of[kStep][kPackThreads];     // unit of input entry p (0xff: padding)
  unsigned char unit_len[kStep][kPackThreads];
  unsigned char unit_start[kStep + 1][kPackThreads];
  unsigned char unit_fill[kStep][kPackThreads];
  unsigned char ent_of[kStep][kPackThreads];      // input positions grouped by unit
  unsigned char order[kStep][kPackThreads];
  unsigned char cx[kPerLane][32][kPackThreads];
  unsigned char cy[kPerLane][32][kPackThreads];
  unsigned char used[kLanes][kPackThreads];       // 4 bits per lane
};

__global__ void __launch_bounds__(kPackThreads)
    pack_kernel(uint32_t* __restrict__ steps, long long total_steps, const uint32_t* __restrict__ meta, int rb,
                uint32_t pad_id, int do_pack, unsigned long long* __restrict__ stats /* dirty entries, dirty steps, long-run steps */) {
  extern __shared__ __align__(16) unsigned char pk_raw[];
  pack_smem& sm = *reinterpret_cast<pack_smem*>(pk_raw);
  const int t = threadIdx.x;
  const long long s = (long long)blockIdx.x * kPackThreads + t;
  if (s >= total_steps) return;
  uint32_t* w = steps + s * kStepWords;
  // ---- load the step, group the entries of a row into units (first-appearance order) ----
  int nreal = 0, nu = 0;
  for (int p = 0; p < kStep; ++p) {
    const uint32_t id = w[p];
    sm.in_id[p][t] = id;
    sm.in_val[p][t] = w[kStep + p];
    const int lr = int((id >> 16) & 0x7fff);
    if (lr == rb) { sm.unit_of[p][t] = 0xff; continue; }
    ++nreal;
    int u = -1;
    for (int e = 0; e < p; ++e)
      if (sm.unit_of[e][t] != 0xff && int((sm.in_id[e][t] >> 16) & 0x7fff) == lr) { u = sm.unit_of[e][t]; break; }
    if (u < 0) { u = nu++; sm.unit_len[u][t] = 0; }
    sm.unit_of[p][t] = (unsigned char)u;
    ++sm.unit_len[u][t];
  }
  bool packed = false;
  if (do_pack && nreal > 0) {
    int acc = 0;
    for (int u = 0; u < nu; ++u) { sm.unit_start[u][t] = (unsigned char)acc; acc += sm.unit_len[u][t]; sm.unit_fill[u][t] = 0; }
    sm.unit_start[nu][t] = (unsigned char)acc;
    for (int p = 0; p < kStep; ++p) {
      const int u = sm.unit_of[p][t];
      if (u == 0xff) continue;
      sm.ent_of[sm.unit_start[u][t] + sm.unit_fill[u][t]++][t] = (unsigned char)p;
    }
    int no = 0;
    for (int L = 9; L >= 1; --L)
      for (int u = 0; u < nu; ++u)
        if (min(int(sm.unit_len[u][t]), 9) == L) sm.order[no++][t] = (unsigned char)u;
    for (int l = 0; l < kLanes; ++l) sm.used[l][t] = 0;
    for (int j = 0; j < kPerLane; ++j)
      for (int k = 0; k < 32; ++k) { sm.cx[j][k][t] = 0; sm.cy[j][k][t] = 0; }
    for (int p = 0; p < kStep; ++p) { sm.out_id[p][t] = pad_id; sm.out_val[p][t] = 0u; }
    int free_in_col[kPerLane] = {kLanes, kLanes, kLanes, kLanes};
    auto ent_id = [&](int u, int i) { return sm.in_id[sm.ent_of[sm.unit_start[u][t] + i][t]][t]; };
    auto ent_val = [&](int u, int i) { return sm.in_val[sm.ent_of[sm.unit_start[u][t] + i][t]][t]; };
    auto put = [&](int lane, int j, uint32_t id, uint32_t v, bool closes) {
      sm.used[lane][t] |= (unsigned char)(1u << j);
      --free_in_col[j];
      sm.out_id[lane * kPerLane + j][t] = id;
      sm.out_val[lane * kPerLane + j][t] = v;
      ++sm.cx[j][id & 31u][t];
      if (closes) ++sm.cy[j][(id >> 16) & 31u][t];
    };
    bool failed = false;
    for (int oi = 0; oi < no && !failed; ++oi) {
      const int u = sm.order[oi][t];
      const int L = sm.unit_len[u][t];
      if (L == 1) {
        const uint32_t id = ent_id(u, 0);
        int best = -1, best_cost = 1 << 30;
        for (int j = 0; j < kPerLane; ++j) {
          if (free_in_col[j] == 0) continue;
          const int cost = (int(sm.cx[j][id & 31u][t]) + int(sm.cy[j][(id >> 16) & 31u][t])) * 64 - free_in_col[j];
          if (cost < best_cost) { best_cost = cost; best = j; }
        }
        if (best < 0) { failed = true; break; }
        int lane = 0;
        while (sm.used[lane][t] & (1u << best)) ++lane;
        put(lane, best, id, ent_val(u, 0), true);
        continue;
      }
      if (L <= kPerLane) {
        int bl = -1, bj = -1, best_cost = 1 << 30;
        for (int lane = 0; lane < kLanes; ++lane)
          for (int j = 0; j + L <= kPerLane; ++j) {
            bool ok = true;
            int cost = 0;
            for (int i = 0; i < L; ++i) {
              ok = ok && !(sm.used[lane][t] & (1u << (j + i)));
              cost += sm.cx[j + i][ent_id(u, i) & 31u][t];
            }
            if (!ok) continue;
            cost += sm.cy[j + L - 1][(ent_id(u, L - 1) >> 16) & 31u][t];
            if (cost < best_cost) { best_cost = cost; bl = lane; bj = j; }
          }
        if (bl >= 0) {
          for (int i = 0; i < L; ++i) put(bl, bj + i, ent_id(u, i), ent_val(u, i), i == L - 1);
          continue;
        }
      }
      // longer runs: lane-aligned first, else the first free contiguous range of cells
      int c0 = -1;
      for (int p0 = 0; p0 + L <= kStep && c0 < 0; p0 += kPerLane) {
        bool ok = true;
        for (int i = 0; i < L && ok; ++i) ok = !(sm.used[(p0 + i) / kPerLane][t] & (1u << ((p0 + i) % kPerLane)));
        if (ok) c0 = p0;
      }
      for (int p0 = 0; p0 + L <= kStep && c0 < 0; ++p0) {
        bool ok = true;
        for (int i = 0; i < L && ok; ++i) ok = !(sm.used[(p0 + i) / kPerLane][t] & (1u << ((p0 + i) % kPerLane)));
        if (ok) c0 = p0;
      }
      if (c0 < 0) { failed = true; break; }
      for (int i = 0; i < L; ++i) put((c0 + i) / kPerLane, (c0 + i) % kPerLane, ent_id(u, i), ent_val(u, i), i == L - 1);
    }
    if (failed) {
      // out of room: lay the units back to back in the same order (always fits)
      for (int p = 0; p < kStep; ++p) { sm.out_id[p][t] = pad_id; sm.out_val[p][t] = 0u; }
      int cell = 0;
      for (int oi = 0; oi < no; ++oi) {
        const int u = sm.order[oi][t];
        for (int i = 0; i < int(sm.unit_len[u][t]); ++i) {
          sm.out_id[cell][t] = ent_id(u, i);
          sm.out_val[cell][t] = ent_val(u, i);
          ++cell;
        }
      }
    }
    packed = true;
  }
  // the final cell contents are in out_* when packed, in in_* otherwise
  // ---- dirty check: every row one contiguous range of cells over at most two lanes ----
  unsigned long long flagged_entries = 0;
  bool any = false, any_long = false;
  for (int p = 0; p < kStep; ++p) {
    const uint32_t id = packed ? sm.out_id[p][t] : sm.in_id[p][t];
    const int lr = int((id >> 16) & 0x7fff);
    if (lr == rb) continue;
    int first = p, last = p, cnt = 0;
    for (int e = 0; e < kStep; ++e) {
      const uint32_t ie = packed ? sm.out_id[e][t] : sm.in_id[e][t];
      if (int((ie >> 16) & 0x7fff) == lr) { if (e < first) first = e; if (e > last) last = e; ++cnt; }
    }
    if (last - first + 1 != cnt) { ++flagged_entries; any = true; }
    else if (last / kPerLane - first / kPerLane >= 2) any_long = true;
  }
  const uint32_t m = meta[s] | (any ? 1u : 0u) | (any_long ? kMetaLongBit : 0u);
  // ---- write back: cells, then bit `lane` of the control word into slot 0 of lane `lane` ----
  for (int p = 0; p < kStep; ++p) {
    uint32_t id = packed ? sm.out_id[p][t] : sm.in_id[p][t];
    if ((p % kPerLane) == 0 && ((m >> (p / kPerLane)) & 1u)) id |= kFlagBit;
    w[p] = id;
    w[kStep + p] = packed ? sm.out_val[p][t] : sm.in_val[p][t];
  }
  if (any) { atomicAdd(stats + 0, flagged_entries); atomicAdd(stats + 1, 1ull); }
  if (any_long) atomicAdd(stats + 2, 1ull);
}

}  // namespace dev

// ---------------------------------------------------------------------------
// Host driver. On success fills d->steps / stream_base / blk_begin (device) and
// the statistics, and returns the final geometry in d->g.
// ---------------------------------------------------------------------------
inline int build_device(plan_data* d, geom g, int rows, int cols, const int* d_off, const int* d_idx,
                        const float* d_val, cudaStream_t s) {
  const char* why = "";
  if (!derive(g, rows, cols, &why)) { set_error("band-tiled plan: %s", why); return LOOPSB_ERR_UNSUPPORTED; }
  const int ns = g.nstreams(), nband = g.nband;
  if ((long long)ns * nband >= (1ll << 31)) { set_error("band-tiled plan: too many (stream, band) keys"); return LOOPSB_ERR_UNSUPPORTED; }
  // row offsets on the host: the row-block cuts are a handful of binary searches
  std::vector<int> off(size_t(rows) + 1);
  LOOPSB_CUDA_TRY(cudaMemcpyAsync(off.data(), d_off, off.size() * 4, cudaMemcpyDeviceToHost, s));
  LOOPSB_CUDA_TRY(cudaStreamSynchronize(s));
  const long long nnz = off[rows];
  std::vector<int> blk_begin(size_t(g.nb) + 1, 0);
  {
    const int rb_cap = g.rb;
    int prev = 0;
    for (int k = 1; k <= g.nb; ++k) {
      const long long target = nnz * k / g.nb;
      int cut = int(std::lower_bound(off.begin(), off.end(), target) - off.begin());
      const long long lo = std::max<long long>(prev, (long long)rows - (long long)(g.nb - k) * rb_cap);
      const long long hi = std::min<long long>(rows, (long long)prev + rb_cap);
      cut = int(std::min<long long>(std::max<long long>(cut, lo), hi));
      if (k == g.nb) cut = rows;
      blk_begin[k] = cut;
      prev = cut;
    }
  }
  int rb_max = 1;
  for (int k = 0; k < g.nb; ++k) rb_max = std::max(rb_max, blk_begin[k + 1] - blk_begin[k]);
  g.rb = rb_max;
  g.rw = 0;   // (rows per warp is only reported by the host builder)
  dev::gconst gc{rows, cols, g.q, g.warps, g.cb, g.xb, g.es, g.cq, g.nband, g.nb, g.rb, ns, g.quantum, g.midpoint};

  // ---- temporaries ----
  int *rowpart = nullptr, *count = nullptr, *start = nullptr, *keyfirst = nullptr, *steps_of = nullptr, *flags = nullptr;
  unsigned char* wmap = nullptr;
  unsigned *keys = nullptr, *atoms = nullptr, *skeys = nullptr, *satoms = nullptr;
  unsigned short *fs = nullptr, *le = nullptr;
  uint32_t* meta = nullptr;
  unsigned long long* stats = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(rowpart); cudaFree(count); cudaFree(start); cudaFree(keyfirst); cudaFree(steps_of); cudaFree(flags);
    cudaFree(wmap); cudaFree(keys); cudaFree(atoms); cudaFree(skeys); cudaFree(satoms); cudaFree(fs); cudaFree(le);
    cudaFree(meta); cudaFree(stats); cudaFree(tmp);
  };
  auto fail = [&](int code, const char* what) {
    if (what) set_error("band-tiled plan (device builder): %s: %s", what, cudaGetErrorString(cudaGetLastError()));
    cleanup();
    return code;
  };
  const size_t nkeys = size_t(ns) * nband;
  const size_t n1 = size_t(nnz > 0 ? nnz : 1);
  if (cudaMalloc(&rowpart, size_t(rows) * g.q * 4) != cudaSuccess || cudaMalloc(&wmap, size_t(rows) * g.q) != cudaSuccess ||
      cudaMalloc(&count, nkeys * 4) != cudaSuccess || cudaMalloc(&start, nkeys * 4) != cudaSuccess ||
      cudaMalloc(&keyfirst, nkeys * 4) != cudaSuccess || cudaMalloc(&steps_of, size_t(ns) * 4) != cudaSuccess ||
      cudaMalloc(&flags, 16) != cudaSuccess || cudaMalloc(&keys, n1 * 4) != cudaSuccess ||
      cudaMalloc(&atoms, n1 * 4) != cudaSuccess || cudaMalloc(&skeys, n1 * 4) != cudaSuccess ||
      cudaMalloc(&satoms, n1 * 4) != cudaSuccess || cudaMalloc(&fs, nkeys * 2) != cudaSuccess ||
      cudaMalloc(&le, nkeys * 2) != cudaSuccess || cudaMalloc(&stats, 32) != cudaSuccess ||
      cudaMalloc(&d->blk_begin, blk_begin.size() * 4) != cudaSuccess ||
      cudaMalloc(&d->stream_base, (size_t(ns) + 1) * 4) != cudaSuccess)
    return fail(LOOPSB_ERR_ALLOC, "cudaMalloc");
  const int h_flags[4] = {0x7fffffff, 0, 0, 0};   // {first bad atom, stream too long, control-word overflow, -}
  cudaMemcpyAsync(flags, h_flags, 16, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d->blk_begin, blk_begin.data(), blk_begin.size() * 4, cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(count, 0, nkeys * 4, s);
  cudaMemsetAsync(stats, 0, 32, s);

  dev::rowpart_kernel<<<(rows + 127) / 128, 128, 0, s>>>(d_off, d_idx, rows, cols, g.q, g.cq, rowpart, flags + 0);
  dev::warp_map_kernel<<<g.grid(), 256, 0, s>>>(rowpart, d->blk_begin, gc, wmap);
  dev::key_kernel<<<(rows + 127) / 128, 128, 0, s>>>(d_off, d_idx, d->blk_begin, wmap, gc, keys, atoms, count);
  int h_bad = 0;
  cudaMemcpyAsync(&h_bad, flags, 4, cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LOOPSB_ERR_CUDA, "count pass");
  if (h_bad != 0x7fffffff) {
    set_error("band-tiled plan: column id out of range at atom %d", h_bad);
    return fail(LOOPSB_ERR_INVALID, nullptr);
  }
  // stable sort of the atoms by (stream, band)
  int key_bits = 1;
  while ((1ull << key_bits) < nkeys) ++key_bits;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, skeys, atoms, satoms, int(nnz), 0, key_bits, s);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, count, keyfirst, int(nkeys), s);
  size_t scan2_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan2_bytes, steps_of, d->stream_base, ns, s);
  tmp_bytes = std::max(tmp_bytes, std::max(scan_bytes, scan2_bytes));
  if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16) != cudaSuccess) return fail(LOOPSB_ERR_ALLOC, "cudaMalloc(cub)");
  if (nnz > 0 && cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, skeys, atoms, satoms, int(nnz), 0, key_bits, s) != cudaSuccess)
    return fail(LOOPSB_ERR_CUDA, "radix sort");
  if (cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, count, keyfirst, int(nkeys), s) != cudaSuccess)
    return fail(LOOPSB_ERR_CUDA, "scan(count)");
  dev::layout_kernel<<<(ns + 127) / 128, 128, 0, s>>>(count, gc, start, fs, le, steps_of, flags + 1);
  if (cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, steps_of, d->stream_base, ns, s) != cudaSuccess)
    return fail(LOOPSB_ERR_CUDA, "scan(steps)");
  int h_last[2] = {0, 0}, h_flags2[4] = {0, 0, 0, 0};
  cudaMemcpyAsync(&h_last[0], d->stream_base + (ns - 1), 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(&h_last[1], steps_of + (ns - 1), 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(h_flags2, flags, 16, cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LOOPSB_ERR_CUDA, "layout pass");
  if (h_flags2[1]) { set_error("band-tiled plan: a warp stream needs more than 65535 steps"); return fail(LOOPSB_ERR_UNSUPPORTED, nullptr); }
  const long long total = (long long)h_last[0] + h_last[1];
  if (total > 0x7fffffffLL / 2) { set_error("band-tiled plan: too many steps"); return fail(LOOPSB_ERR_UNSUPPORTED, nullptr); }
  const int h_total = int(total);
  cudaMemcpyAsync(d->stream_base + ns, &h_total, 4, cudaMemcpyHostToDevice, s);
  const size_t steps_words = size_t(total + g.es) * kStepWords;
  if (cudaMalloc(&d->steps, steps_words * 4) != cudaSuccess || cudaMalloc(&meta, size_t(total > 0 ? total : 1) * 4) != cudaSuccess)
    return fail(LOOPSB_ERR_ALLOC, "cudaMalloc(steps)");
  const uint32_t pad_id = (uint32_t(g.rb) << 16) | uint32_t(g.zero_slot());
  dev::fill_pad_kernel<<<unsigned((steps_words + 255) / 256), 256, 0, s>>>(d->steps, total + g.es, pad_id);
  cudaMemsetAsync(meta, 0, size_t(total > 0 ? total : 1) * 4, s);
  if (nnz > 0)
    dev::scatter_kernel<<<unsigned((nnz + 255) / 256), 256, 0, s>>>(skeys, satoms, nnz, d_off, d_idx, d_val, d->blk_begin,
                                                                   start, keyfirst, d->stream_base, gc, d->steps);
  dev::meta_kernel<<<(ns + 127) / 128, 128, 0, s>>>(count, fs, le, d->stream_base, gc, meta, flags + 2);
  if (total > 0) {
    if (cudaFuncSetAttribute(dev::pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(dev::pack_smem))) != cudaSuccess)
      return fail(LOOPSB_ERR_CUDA, "pack kernel shared memory");
    dev::pack_kernel<<<unsigned((total + dev::kPackThreads - 1) / dev::kPackThreads), dev::kPackThreads,
                       sizeof(dev::pack_smem), s>>>(d->steps, total, meta, g.rb, pad_id, g.pack, stats);
  }
  unsigned long long h_stats[4] = {0, 0, 0, 0};
  cudaMemcpyAsync(h_stats, stats, 32, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(h_flags2, flags, 16, cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail(LOOPSB_ERR_CUDA, "scatter / pack pass");
  if (h_flags2[2]) {
    set_error("band-tiled plan: too many empty / starting bands in one step for the control word");
    return fail(LOOPSB_ERR_UNSUPPORTED, nullptr);
  }
  cleanup();
  d->g = g;
  d->total_steps = total;
  d->real_entries = nnz;
  d->pad_entries = total * kStep - nnz;
  d->flagged_entries = (long long)h_stats[0];
  d->flagged_steps = (long long)h_stats[1];
  d->long_steps = (long long)h_stats[2];
  return LOOPSB_OK;
}

}  // namespace bt
}  // namespace loopsb
