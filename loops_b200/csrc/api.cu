// loops_b200/csrc/api.cu -- the C ABI declared in include/loopsb.h:
// plan management (the reference's merge_path::preprocess_t generalised) and
// SpMV dispatch by (schedule, layout kind).

#include "common.cuh"
#include "spmv_merge.cuh"
#include "spmv_merge2.cuh"
#include "spmv_schedules.cuh"
#include "bcsr_tc.cuh"
#include "spmv_tiled.cuh"
#include "tiled_build.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <vector>
#include <mutex>
#include <new>

namespace loopsb {

char* last_error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
}

const device_props* current_device() {
  static device_props cache[64];
  static std::mutex mu;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    (void)cudaGetLastError();
    set_error("no usable CUDA device: %s (loops-b200 has no CPU fallback)",
              cudaGetErrorString(e));
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(mu);
  device_props& p = cache[dev];
  if (!p.valid) {
    if (cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&p.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("cudaDeviceGetAttribute failed on device %d", dev);
      return nullptr;
    }
    p.valid = true;
  }
  return &p;
}

}  // namespace loopsb

using namespace loopsb;

// spmv_generic.cu: the (schedule x layout) cells without a reference kernel
namespace loopsb { namespace generic {
struct state;
int create(state** out, const loopsb_layout_t* lay, int schedule, int wo_grid, cudaStream_t s);
void destroy(state* st);
bool supports(int kind, int schedule);
int run(state* st, const loopsb_layout_t* lay, int schedule, const float* values, const int* cols, const int* rows,
        const float* x, float* y, int num_rows, cudaStream_t s);
}}  // namespace loopsb::generic

// Merge-path SpMV geometry variants: {CTA threads, merge tiles per CTA tile,
// stages, target CTAs/SM}. Variant 0 is the default; LOOPSB_MERGE_VARIANT
// selects another one at plan creation (tuning aid).
namespace {
struct merge_variant {
  int threads, g, stages, ctas_per_sm, smem;
  int carveout_kb;  // preferred shared-memory carve-out per SM (keeps the L1 that tracks gather misses)
  void (*launch)(bool array_ends, int grid, int smem, cudaStream_t s, const int* row_end, int pitch,
                 const int* indices, const float* values, const float* x, float* y, const int2* coords,
                 int M, int T, int A, int nct, int* carry_row, float* carry_val, long long* phases);
  cudaError_t (*prepare)(int smem);
};

template <int THREADS, int G, int STAGES, int MINB>
struct merge_inst {
  static constexpr int TILE = G * mp::kRefItemsPerMergeTile;
  using shared_t = mp::merge_shared<THREADS, TILE, STAGES>;
  static void launch(bool array_ends, int grid, int smem, cudaStream_t s, const int* row_end, int pitch,
                     const int* indices, const float* values, const float* x, float* y, const int2* coords,
                     int M, int T, int A, int nct, int* carry_row, float* carry_val, long long* phases) {
    if (array_ends)
      mp::spmv_merge_kernel<THREADS, TILE, STAGES, MINB, true><<<grid, THREADS, smem, s>>>(
          row_end, pitch, indices, values, x, y, coords, M, G, T, A, nct, carry_row, carry_val, phases);
    else
      mp::spmv_merge_kernel<THREADS, TILE, STAGES, MINB, false><<<grid, THREADS, smem, s>>>(
          row_end, pitch, indices, values, x, y, coords, M, G, T, A, nct, carry_row, carry_val, phases);
  }
  // Smallest carve-out bucket (KB) that holds MINB CTAs (+1 KB driver reserve each).
  static constexpr int carveout_kb() {
    const int need = (MINB * (int(sizeof(shared_t)) + 1024) + 1023) / 1024;
    const int buckets[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
    for (int b : buckets) if (need <= b) return b;
    return 228;
  }
  static cudaError_t prepare(int smem) {
    auto ka = mp::spmv_merge_kernel<THREADS, TILE, STAGES, MINB, true>;
    auto kp = mp::spmv_merge_kernel<THREADS, TILE, STAGES, MINB, false>;
    const int pct = (carveout_kb() * 100 + 227) / 228;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ka, cudaFuncAttributePreferredSharedMemoryCarveout, pct)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(kp, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  }
  static constexpr merge_variant variant() {
    return merge_variant{THREADS, G, STAGES, MINB, int(sizeof(shared_t)), carveout_kb(), &launch, &prepare};
  }
};

const merge_variant kMergeVariants[] = {
    merge_inst<256, 4, 2, 2>::variant(),  // 0: 256 thr, 4096 items, 2 stages, 2 CTAs/SM
    merge_inst<128, 2, 2, 5>::variant(),  // 1: 128 thr, 2048 items, 2 stages, 5 CTAs/SM
    merge_inst<128, 2, 3, 4>::variant(),  // 2: 128 thr, 2048 items, 3 stages, 4 CTAs/SM
    merge_inst<256, 2, 2, 5>::variant(),  // 3: 256 thr, 2048 items (8/thread), 5 CTAs/SM
    merge_inst<512, 4, 2, 2>::variant(),  // 4: 512 thr, 4096 items (8/thread), 2 CTAs/SM
    merge_inst<128, 1, 3, 8>::variant(),  // 5: 128 thr, 1024 items (8/thread), 8 CTAs/SM
    merge_inst<256, 4, 3, 1>::variant(),  // 6: 256 thr, 4096 items, 3 stages, 1 CTA/SM
    merge_inst<256, 2, 2, 3>::variant(),  // 7: 256 thr, 2048 items, 2 stages, 3 CTAs/SM (132 KB)
    merge_inst<256, 2, 2, 4>::variant(),  // 8: 256 thr, 2048 items, 2 stages, 4 CTAs/SM (196 KB)
    merge_inst<128, 1, 2, 6>::variant(),  // 9: 128 thr, 1024 items, 2 stages, 6 CTAs/SM (132 KB)
    merge_inst<128, 2, 2, 3>::variant(),  // 10: 128 thr, 2048 items, 2 stages, 3 CTAs/SM (132 KB)
    merge_inst<128, 1, 2, 7>::variant(),  // 11: 128 thr, 1024 items, 2 stages, 7 CTAs/SM (164 KB)
    merge_inst<256, 1, 2, 6>::variant(),  // 12: 256 thr, 1024 items (4/thread), 6 CTAs/SM
};
constexpr int kNumMergeVariants = int(sizeof(kMergeVariants) / sizeof(kMergeVariants[0]));

int merge_variant_from_env() {
  // Default: 128-thread CTAs over one 1024-item merge tile each, 6 CTAs/SM
  // inside the 132 KB carve-out (measured best on B200, profiles/).
  constexpr int kDefault = 9;
  const char* e = getenv("LOOPSB_MERGE_VARIANT");
  if (!e) return kDefault;
  int v = atoi(e);
  return (v >= 0 && v < kNumMergeVariants) ? v : kDefault;
}

// Second-generation merge-path kernel (spmv_merge2.cuh): {CTA threads, target CTAs/SM}.
struct merge2_variant {
  int threads, ctas_per_sm, smem, carveout_kb;
  void (*launch)(bool array_ends, bool accumulate, int grid, int smem, cudaStream_t s, const int* row_end, int pitch,
                 const int* indices, const float* values, const float* x, float* y, const int2* coords,
                 int M, int T, int A, int nct, int* carry_row, float* carry_val);
  cudaError_t (*prepare)(int smem);
};

template <int THREADS, int MINB, int STAGES>
struct merge2_inst {
  using shared_t = mp2::shared_t<THREADS, STAGES>;
  static void launch(bool array_ends, bool accumulate, int grid, int smem, cudaStream_t s, const int* row_end,
                     int pitch, const int* indices, const float* values, const float* x, float* y,
                     const int2* coords, int M, int T, int A, int nct, int* carry_row, float* carry_val) {
    if (array_ends && accumulate)
      mp2::spmv_merge2_kernel<THREADS, MINB, STAGES, true, true><<<grid, THREADS, smem, s>>>(
          row_end, pitch, indices, values, x, y, coords, M, T, A, nct, carry_row, carry_val);
    else if (array_ends)
      mp2::spmv_merge2_kernel<THREADS, MINB, STAGES, true, false><<<grid, THREADS, smem, s>>>(
          row_end, pitch, indices, values, x, y, coords, M, T, A, nct, carry_row, carry_val);
    else
      mp2::spmv_merge2_kernel<THREADS, MINB, STAGES, false, false><<<grid, THREADS, smem, s>>>(
          row_end, pitch, indices, values, x, y, coords, M, T, A, nct, carry_row, carry_val);
  }
  static constexpr int carveout_kb() {
    const int need = (MINB * (int(sizeof(shared_t)) + 1024) + 1023) / 1024;
    const int buckets[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
    for (int b : buckets) if (need <= b) return b;
    return 228;
  }
  static cudaError_t prepare(int smem) {
    const void* ks[] = {reinterpret_cast<const void*>(mp2::spmv_merge2_kernel<THREADS, MINB, STAGES, true, false>),
                        reinterpret_cast<const void*>(mp2::spmv_merge2_kernel<THREADS, MINB, STAGES, true, true>),
                        reinterpret_cast<const void*>(mp2::spmv_merge2_kernel<THREADS, MINB, STAGES, false, false>)};
    const int pct = (carveout_kb() * 100 + 227) / 228;
    for (const void* k : ks) {
      cudaError_t e;
      if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct)) != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  static constexpr merge2_variant variant() {
    return merge2_variant{THREADS, MINB, int(sizeof(shared_t)), carveout_kb(), &launch, &prepare};
  }
};

const merge2_variant kMerge2Variants[] = {
    merge2_inst<128, 6, 2>::variant(),  // 0: 128 thr, 2 chunks/thread, 6 CTAs/SM, two bulk-copy stages (default)
    merge2_inst<128, 5, 2>::variant(),  // 1
    merge2_inst<128, 4, 2>::variant(),  // 2: fewer CTAs, more L1 -- ahead when x is far larger than an L2 partition
    merge2_inst<128, 6, 1>::variant(),  // 3: one stage (measured: the re-fill does not land in time, 183 vs 161 us)
    merge2_inst<256, 3, 2>::variant(),  // 4: 256 thr, 1 chunk/thread
    merge2_inst<128, 7, 2>::variant(),  // 5
};
constexpr int kNumMerge2Variants = int(sizeof(kMerge2Variants) / sizeof(kMerge2Variants[0]));

// LOOPSB_MERGE_KERNEL=1 keeps the first-generation kernel (tuning / A-B aid).
int merge_generation_from_env() {
  const char* e = getenv("LOOPSB_MERGE_KERNEL");
  return (e && atoi(e) == 1) ? 1 : 2;
}
int merge2_variant_from_env() {
  // -1 = decide per call from the size of x (see merge2_pick)
  const char* e = getenv("LOOPSB_MERGE2_VARIANT");
  if (!e) return -1;
  int v = atoi(e);
  return (v >= 0 && v < kNumMerge2Variants) ? v : -1;
}
// Measured on B200 (profiles/merge_probe_r02.txt, block_probe_r02.txt): 6 CTAs/SM for config 2
// (x = 4 MB, 161 vs 170 us) and for the column blocks of a multi-GPU shard (x slices of
// 8..56 MB: 361 vs 380 us per 1/8 shard), 4 CTAs/SM when the gathers range over all of a
// 64 MB x at once (345 vs 354 us).
inline int merge2_pick(int forced, size_t x_bytes) {
  if (forced >= 0) return forced;
  return x_bytes >= (size_t(60) << 20) ? 2 : 0;
}
}  // namespace

struct loopsb_plan {
  loopsb_layout_t lay;
  int schedule;
  int device;
  int sm_count;
  // merge_path_flat
  int2* coords = nullptr;     // S(b*1024), b = 0..M
  long long M = 0;
  int num_cta_tiles = 0;
  int variant = 0;
  int gen = 1;        // merge kernel generation (2 = spmv_merge2.cuh, needs 16-byte aligned arrays)
  int variant2 = -1;  // index into kMerge2Variants forced by LOOPSB_MERGE2_VARIANT, -1 = pick per call
  long long x_span_bytes = -1;   // loopsb_plan_hint_x_bytes: how much of x the matrix's columns touch
  generic::state* cell = nullptr; // schedule::setup<>-driven cells without a reference kernel (spmv_generic.cu)
  int wo_grid = 0;   // work_oriented: reference-style grid (blocks of 128 threads)
  long long* phases = nullptr;  // LOOPSB_DEBUG_PHASES=1: per-CTA phase cycle counters
  int* carry_row = nullptr;
  float* carry_val = nullptr;
  // launch geometry of the SpMV kernel
  int grid = 0;
  int cta_threads = 0;
  int smem_bytes = 0;
  int launches = 1;
  long long workspace_bytes = 0;
  // bcsr tensor-core path
  bcsr_tc::plan_data* tc = nullptr;
  // band-tiled copy of a CSR matrix (loopsb_plan_tile_csr)
  bt::plan_data* tiled = nullptr;
  // kernel-time probes (loopsb_plan_probe_*)
  cudaEvent_t* probe_ev = nullptr;  // 2 * probe_cap events
  int probe_cap = 0;
  int probe_n = 0;
  bool probing = false;
};

namespace {
// Brackets the dominant kernel of one SpMV call with an event pair.
struct probe_scope {
  loopsb_plan* p;
  cudaStream_t s;
  bool on;
  probe_scope(loopsb_plan* plan, cudaStream_t stream)
      : p(plan), s(stream), on(plan->probing && plan->probe_n < plan->probe_cap) {
    if (on) cudaEventRecord(p->probe_ev[2 * p->probe_n], s);
  }
  void close() {
    if (on) {
      cudaEventRecord(p->probe_ev[2 * p->probe_n + 1], s);
      p->probe_n++;
      on = false;
    }
  }
};
void free_probes(loopsb_plan* p) {
  if (p->probe_ev) {
    for (int i = 0; i < 2 * p->probe_cap; ++i)
      if (p->probe_ev[i]) cudaEventDestroy(p->probe_ev[i]);
    delete[] p->probe_ev;
  }
  p->probe_ev = nullptr;
  p->probe_cap = p->probe_n = 0;
  p->probing = false;
}
}  // namespace

namespace {
cudaError_t prepare_merge2(int forced) {
  for (int v = 0; v < kNumMerge2Variants; ++v) {
    if (forced >= 0 ? v != forced : (v != merge2_pick(-1, 0) && v != merge2_pick(-1, size_t(1) << 40))) continue;
    cudaError_t e = kMerge2Variants[v].prepare(kMerge2Variants[v].smem);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Second-generation merge-path launch (merge_path_flat / ell_merge_path / work_oriented tiles).
void launch_merge2(const loopsb_plan* plan, bool array_ends, bool accumulate, cudaStream_t s,
                   const int32_t* col_indices, const float* values, const float* x, float* y, int32_t num_cols) {
  const size_t x_bytes = plan->x_span_bytes >= 0 ? size_t(plan->x_span_bytes) : size_t(num_cols) * sizeof(float);
  const merge2_variant& m2 = kMerge2Variants[merge2_pick(plan->variant2, x_bytes)];
  const int grid = std::min(m2.ctas_per_sm * plan->sm_count, plan->num_cta_tiles);
  m2.launch(array_ends, accumulate, grid, m2.smem, s, array_ends ? plan->lay.offsets + 1 : nullptr, plan->lay.pitch,
            col_indices, values, x, y, plan->coords, int(plan->M), plan->lay.num_tiles, plan->lay.num_atoms,
            plan->num_cta_tiles, plan->carry_row, plan->carry_val);
}

inline bool aligned16(const void* a, const void* b) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
}

int launch_tiled(bt::plan_data* d, const float* x, float* y, cudaStream_t s) {
  bt::params p;
  p.steps = d->steps; p.steps_end = d->steps + size_t(d->total_steps + d->g.es) * bt::kStepWords; p.stream_base = d->stream_base; p.blk_begin = d->blk_begin; 
  p.x = x; p.y = y; p.partial = d->partial; p.counters = d->counters;
  p.rows = d->g.rows; p.cols = d->g.cols; p.rb = d->g.rb; p.cq = d->g.cq; p.cb = d->g.cb;
  p.xb = d->g.xb; p.es = d->g.es; p.nband = d->g.nband; p.q = d->g.q; p.nb = d->g.nb;
  p.prof = d->prof;
  p.peers = d->peers;
  {
    static const int ahead = getenv("LOOPSB_TILED_L2AHEAD") ? atoi(getenv("LOOPSB_TILED_L2AHEAD")) : bt::kL2Ahead;
    p.l2_ahead = ahead;
    static const int guard = getenv("LOOPSB_TILED_L2GUARD") ? atoi(getenv("LOOPSB_TILED_L2GUARD")) : 0;
    p.l2_guard = guard;
  }
  bt::kernel_fn k = bt::kernel_for(d->g.warps, d->g.es, d->prof != nullptr);
  // Launches after the first on a copy use PROGRAMMATIC DEPENDENT LAUNCH instead of a cooperative launch:
  // the CTAs of launch n+1 take the SMs the CTAs of launch n leave and run their prologue (barrier set-up,
  // first matrix-stream requests, y tile) while the rest of launch n is still reducing; nothing launch n
  // reads or writes is touched before griddepcontrol.wait (measured on config 2: 61.1 -> 57.9 us per SpMV
  // back to back). The first launch after a (re)build stays fully ordered behind the builder kernels.
  // Without the cooperative launch's co-residency guarantee the q CTAs of a row block still wait for each
  // other, which is safe as long as no OTHER chain of these grids can be partially resident at the same time
  // (two such grids could hold each other's SMs). On one stream launch n+1 only starts once every CTA of
  // launch n has started, so a single stream is always safe; the rule for several streams: an event is
  // recorded after every band-tiled launch, and a launch is PDL only when every earlier launch made on a
  // DIFFERENT stream of this device is known to have completed; otherwise it is cooperative, which cannot
  // start until all of its CTAs fit. LOOPSB_TILED_PDL=0 turns PDL off (e.g. processes sharing the GPU via MPS).
  static const int pdl_env = getenv("LOOPSB_TILED_PDL") ? atoi(getenv("LOOPSB_TILED_PDL")) : 1;
  struct launch_tracker {
    cudaStream_t cur = nullptr;         // stream of the latest band-tiled launch
    cudaEvent_t cur_ev = nullptr;       // recorded after it
    std::vector<cudaEvent_t> foreign;   // last events of the streams used before `cur`, while still pending
    std::vector<cudaEvent_t> spare;
    bool broken = false;                // an event could not be created or recorded: stay cooperative
  };
  static std::mutex pdl_mu;
  static launch_tracker trackers[64];
  int dev = 0;
  LOOPSB_CUDA_TRY(cudaGetDevice(&dev));
  const bool track = pdl_env == 1 && dev >= 0 && dev < 64;   // (2 = PDL without the tracking events: measurement only)
  p.pdl = 0;
  bool coop_ok = true;
  // decision, launch and event record are one critical section: two host threads launching on two
  // streams of the same device must not both see "nothing foreign in flight"
  std::unique_lock<std::mutex> pdl_lock(pdl_mu, std::defer_lock);
  if (track) {
    pdl_lock.lock();
    launch_tracker& T = trackers[dev];
    if (T.cur_ev && T.cur != s) {          // stream switch: the old stream's tail becomes a foreign event
      T.foreign.push_back(T.cur_ev);
      T.cur_ev = nullptr;
    }
    for (size_t f = 0; f < T.foreign.size();) {
      const cudaError_t qe = cudaEventQuery(T.foreign[f]);
      if (qe == cudaSuccess) {
        T.spare.push_back(T.foreign[f]);
        T.foreign[f] = T.foreign.back();
        T.foreign.pop_back();
      } else {
        (void)cudaGetLastError();
        ++f;
      }
    }
    coop_ok = T.foreign.empty() && !T.broken;
  }
  if (pdl_env != 0 && d->peers && d->launched && !d->prof && (pdl_env == 2 || (track && coop_ok))) p.pdl = 1;
  auto note_launch = [&]() {   // (pdl_lock is held)
    if (!track) return;
    launch_tracker& T = trackers[dev];
    if (!T.cur_ev) {
      if (!T.spare.empty()) { T.cur_ev = T.spare.back(); T.spare.pop_back(); }
      else if (cudaEventCreateWithFlags(&T.cur_ev, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); T.cur_ev = nullptr; }
    }
    T.cur = s;
    if (!T.cur_ev || cudaEventRecord(T.cur_ev, s) != cudaSuccess) { (void)cudaGetLastError(); T.broken = true; }
  };
  if (p.pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(d->g.grid());
    cfg.blockDim = dim3(d->g.cta_threads());
    cfg.dynamicSmemBytes = size_t(d->smem);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    LOOPSB_CUDA_TRY(cudaLaunchKernelEx(&cfg, k, p));
    note_launch();
    return LOOPSB_OK;
  }
  d->launched = true;
  if (d->peers) {
    // every CTA resident at once (checked at plan time): the CTAs of a row block may wait for each other
    void* args[] = {&p};
    LOOPSB_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(k), dim3(d->g.grid()),
                                                dim3(d->g.cta_threads()), args, size_t(d->smem), s));
    note_launch();
  } else {
    k<<<d->g.grid(), d->g.cta_threads(), d->smem, s>>>(p);
    LOOPSB_CUDA_TRY(cudaGetLastError());
  }
  return LOOPSB_OK;
}

void fill_tiled_info(loopsb_tiled_info_t* o, const bt::geom& g, long long total_steps, long long real_entries,
                     long long pad_entries, long long flagged_entries, long long flagged_steps, long long bytes,
                     long long long_steps) {
  memset(o, 0, sizeof(*o));
  o->nb = g.nb; o->q = g.q; o->warps = g.warps; o->cb = g.cb; o->xb = g.xb; o->es = g.es;
  o->rb = g.rb; o->rw = g.rw; o->cq = g.cq; o->nband = g.nband;
  o->grid_blocks = g.grid(); o->cta_threads = g.cta_threads(); o->smem_bytes = g.smem_bytes();
  o->total_steps = total_steps; o->real_entries = real_entries; o->pad_entries = pad_entries;
  o->flagged_entries = flagged_entries; o->flagged_steps = flagged_steps; o->bytes = bytes;
  o->long_steps = int32_t(long_steps);
}
}  // namespace

struct loopsb_tiled_image {
  bt::host_image im;
};

namespace {
// A dense x that does not fit shared memory is gathered from L2; when it is large
// (the 64 MB x of the multi-GPU shards) the matrix stream pushes it out of the
// 126 MB L2 between uses. Pin it for the duration of the launch: persisting-L2
// access-policy window over x on the launching stream (misses stream through),
// replaced by the caller's previous window right after the launch, so the stream is left as it was.
// Measured on the 1/8 shard of BASELINE configs[4]: 430 -> 379 us.
struct l2_pin_scope {
  cudaStream_t s;
  bool on = false;
  cudaStreamAttrValue prev{};     // the caller's own window on this stream, restored afterwards
  l2_pin_scope(const loopsb_plan* p, const float* x, size_t bytes, cudaStream_t stream) : s(stream) {
    static const bool off = getenv("LOOPSB_NO_L2_PIN") != nullptr;
    if (off || bytes < (size_t(16) << 20)) return;
    // limits are per device; the persisting carve-out is raised once per device, under a lock
    // (several ranks / host threads may make their first call at the same time)
    static std::mutex mu;
    static int max_persist[64], max_window[64];
    static bool known[64];
    const int dev = p->device;
    if (dev < 0 || dev >= 64) return;
    int persist, window;
    {
      std::lock_guard<std::mutex> lock(mu);
      if (!known[dev]) {
        max_persist[dev] = max_window[dev] = 0;
        cudaDeviceGetAttribute(&max_persist[dev], cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window[dev], cudaDevAttrMaxAccessPolicyWindowSize, dev);
        if (max_persist[dev] > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, size_t(max_persist[dev]));
        (void)cudaGetLastError();
        known[dev] = true;
      }
      persist = max_persist[dev];
      window = max_window[dev];
    }
    if (persist <= 0 || window <= 0) return;
    if (cudaStreamGetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &prev) != cudaSuccess) {
      (void)cudaGetLastError();
      return;
    }
    cudaStreamAttrValue v{};
    v.accessPolicyWindow.base_ptr = const_cast<float*>(x);
    v.accessPolicyWindow.num_bytes = std::min(size_t(window), bytes);
    const double ratio = double(std::min(size_t(persist), bytes)) / double(v.accessPolicyWindow.num_bytes);
    v.accessPolicyWindow.hitRatio = ratio > 1.0 ? 1.0f : float(ratio);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    on = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v) == cudaSuccess;
    if (!on) (void)cudaGetLastError();
  }
  ~l2_pin_scope() {
    if (!on) return;
    cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &prev);   // what the caller had (usually none)
    (void)cudaGetLastError();
  }
};
}  // namespace

extern "C" {

int loopsb_tiled_image_build_host(int32_t num_rows, int32_t num_cols, const int32_t* host_offsets,
                                  const int32_t* host_indices, const float* host_values,
                                  const int32_t geometry[6], loopsb_tiled_image_t** out) {
  LOOPSB_REQUIRE(out != nullptr && geometry != nullptr && host_offsets != nullptr, "null argument");
  *out = nullptr;
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0, "negative dimensions");
  LOOPSB_REQUIRE(num_rows == 0 || host_offsets[num_rows] == 0 || (host_indices && host_values), "null matrix arrays");
  loopsb_tiled_image* img = new (std::nothrow) loopsb_tiled_image();
  if (!img) { set_error("host allocation failed"); return LOOPSB_ERR_ALLOC; }
  bt::geom g;
  g.nb = geometry[0]; g.q = geometry[1]; g.warps = geometry[2];
  g.cb = geometry[3]; g.xb = geometry[4]; g.es = geometry[5];
  if (const char* e = getenv("LOOPSB_TILED_PACK")) g.pack = atoi(e) != 0;
  if (const char* e = getenv("LOOPSB_TILED_ROUND")) g.quantum = atoi(e) != 0 ? g.es : 1;
  if (const char* e = getenv("LOOPSB_TILED_SPLIT")) g.midpoint = atoi(e) != 0;
  int rc = LOOPSB_OK;
  try {
    rc = bt::build_host(img->im, g, num_rows, num_cols, host_offsets, host_indices, host_values);
  } catch (const std::bad_alloc&) {
    set_error("host allocation failed while tiling");
    rc = LOOPSB_ERR_ALLOC;
  }
  if (rc != LOOPSB_OK) { delete img; return rc; }
  *out = img;
  return LOOPSB_OK;
}

int loopsb_tiled_image_info(const loopsb_tiled_image_t* img, loopsb_tiled_info_t* info) {
  LOOPSB_REQUIRE(img != nullptr && info != nullptr, "null argument");
  const bt::host_image& im = img->im;
  fill_tiled_info(info, im.g, im.total_steps, im.real_entries, im.pad_entries, im.flagged_entries,
                  im.flagged_steps, (long long)im.steps.size() * 4, im.long_steps);
  return LOOPSB_OK;
}

int loopsb_tiled_image_arrays(const loopsb_tiled_image_t* img, const uint32_t** steps,
                              const int32_t** stream_base, const uint16_t** first_step,
                              const uint16_t** last_step_end, const int32_t** block_begin,
                              const int32_t** warp_begin) {
  LOOPSB_REQUIRE(img != nullptr, "null argument");
  if (block_begin) *block_begin = img->im.blk_begin.data();
  if (warp_begin) *warp_begin = img->im.warp_begin.data();
  if (steps) *steps = img->im.steps.data();
  if (stream_base) *stream_base = img->im.stream_base.data();
  if (first_step) *first_step = img->im.fs.data();
  if (last_step_end) *last_step_end = img->im.le.data();
  return LOOPSB_OK;
}

int loopsb_tiled_image_free(loopsb_tiled_image_t* img) {
  delete img;
  return LOOPSB_OK;
}

int loopsb_plan_untile(loopsb_plan_t* plan) {
  LOOPSB_REQUIRE(plan != nullptr, "plan is null");
  if (plan->tiled) { bt::destroy(plan->tiled); plan->tiled = nullptr; }
  return LOOPSB_OK;
}

int loopsb_plan_tiled_info(const loopsb_plan_t* plan, loopsb_tiled_info_t* info) {
  LOOPSB_REQUIRE(plan != nullptr && info != nullptr, "null argument");
  if (!plan->tiled) { set_error("the plan holds no band-tiled copy"); return LOOPSB_ERR_UNSUPPORTED; }
  const bt::plan_data* d = plan->tiled;
  fill_tiled_info(info, d->g, d->total_steps, d->real_entries, d->pad_entries, d->flagged_entries,
                  d->flagged_steps, d->bytes, d->long_steps);
  return LOOPSB_OK;
}

int loopsb_plan_tiled_download(const loopsb_plan_t* plan, uint32_t* host_steps, int64_t capacity_words,
                               int32_t* host_stream_base, int32_t* host_block_begin) {
  LOOPSB_REQUIRE(plan != nullptr, "plan is null");
  if (!plan->tiled) { set_error("the plan holds no band-tiled copy"); return LOOPSB_ERR_UNSUPPORTED; }
  const bt::plan_data* d = plan->tiled;
  const long long words = (d->total_steps + d->g.es) * bt::kStepWords;
  LOOPSB_CUDA_TRY(cudaDeviceSynchronize());
  if (host_steps) {
    LOOPSB_REQUIRE(capacity_words >= words, "host buffer too small");
    LOOPSB_CUDA_TRY(cudaMemcpy(host_steps, d->steps, size_t(words) * 4, cudaMemcpyDeviceToHost));
  }
  if (host_stream_base)
    LOOPSB_CUDA_TRY(cudaMemcpy(host_stream_base, d->stream_base, (size_t(d->g.nstreams()) + 1) * 4, cudaMemcpyDeviceToHost));
  if (host_block_begin)
    LOOPSB_CUDA_TRY(cudaMemcpy(host_block_begin, d->blk_begin, (size_t(d->g.nb) + 1) * 4, cudaMemcpyDeviceToHost));
  return LOOPSB_OK;
}

int loopsb_plan_tile_csr(loopsb_plan_t* plan, const int32_t* col_indices, const float* values,
                         int32_t num_cols, int32_t flags, void* stream) {
  LOOPSB_REQUIRE(plan != nullptr, "plan is null");
  LOOPSB_REQUIRE(plan->schedule == LOOPSB_SCHED_MERGE_PATH_FLAT && plan->lay.kind == LOOPSB_LAYOUT_CSR,
                 "band tiling applies to merge_path_flat plans over CSR");
  LOOPSB_REQUIRE(num_cols >= 0, "negative columns");
  const device_props* dp = current_device();
  if (!dp) return LOOPSB_ERR_CUDA;
  const int rows = plan->lay.num_tiles, nnz = plan->lay.num_atoms;
  if (rows == 0 || nnz == 0 || num_cols == 0) { set_error("nothing to tile"); return LOOPSB_ERR_UNSUPPORTED; }
  LOOPSB_REQUIRE(col_indices != nullptr && values != nullptr, "null matrix arrays");
  cudaStream_t s = as_stream(stream);
  const bool force = (flags & LOOPSB_TILE_FORCE) != 0;

  bt::geom g = bt::choose_geom(rows, num_cols, dp->sm_count, dp->max_smem_optin);
  const char* why = "";
  if (!bt::derive(g, rows, num_cols, &why)) { set_error("band-tiled plan: %s", why); return LOOPSB_ERR_UNSUPPORTED; }
  if (g.smem_bytes() > dp->max_smem_optin) {
    set_error("band-tiled plan needs %d bytes of shared memory (device allows %d)", g.smem_bytes(), dp->max_smem_optin);
    return LOOPSB_ERR_UNSUPPORTED;
  }
  if (!bt::kernel_for(g.warps, g.es)) {
    set_error("band-tiled plan: no kernel for %d consumer warps / prefetch depth %d", g.warps, g.es);
    return LOOPSB_ERR_UNSUPPORTED;
  }
  if (!force) {
    // Cost model: every row block re-reads its column part of x from L2
    // (nb * cols * 4 bytes in total) -- worth it only while that stays below
    // the matrix stream itself, and only for matrices big enough to fill the chip.
    const double x_traffic = double(g.nb) * double(num_cols) * 4.0;
    const double stream_bytes = double(nnz) * 8.0;
    if (nnz < (1 << 22) || x_traffic > stream_bytes) {
      set_error("band-tiled plan not profitable (nnz %d, x re-read %.0f MB vs stream %.0f MB)", nnz,
                x_traffic / 1e6, stream_bytes / 1e6);
      return LOOPSB_ERR_UNSUPPORTED;
    }
  }

  bt::plan_data* d = new (std::nothrow) bt::plan_data();
  if (!d) { set_error("host allocation failed"); return LOOPSB_ERR_ALLOC; }
  auto fail = [&](int code) { bt::destroy(d); return code; };
  int rc = LOOPSB_OK;
  if (!getenv("LOOPSB_TILED_HOST_BUILD")) {
    // Device builder (tiled_build.cuh): the matrix stays in HBM, only the row
    // offsets and a few counters cross PCIe. Same image as the host builder.
    rc = bt::build_device(d, g, rows, num_cols, plan->lay.offsets, col_indices, values, s);
    if (rc != LOOPSB_OK) return fail(rc);
  } else {
    // Host builder (the reference implementation of the format; also what the
    // CPU format tests run): download CSR, tile, upload.
    std::vector<int32_t> h_off, h_idx;
    std::vector<float> h_val;
    bt::host_image im;
    try {
      h_off.resize(size_t(rows) + 1); h_idx.resize(size_t(nnz)); h_val.resize(size_t(nnz));
      if (cudaStreamSynchronize(s) != cudaSuccess ||
          cudaMemcpy(h_off.data(), plan->lay.offsets, h_off.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(h_idx.data(), col_indices, h_idx.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(h_val.data(), values, h_val.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("download of the CSR arrays failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(LOOPSB_ERR_CUDA);
      }
      if (h_off[rows] != nnz) { set_error("invalid argument: offsets[rows] must equal num_atoms"); return fail(LOOPSB_ERR_INVALID); }
      rc = bt::build_host(im, g, rows, num_cols, h_off.data(), h_idx.data(), h_val.data());
    } catch (const std::bad_alloc&) {
      set_error("host allocation failed while tiling");
      rc = LOOPSB_ERR_ALLOC;
    }
    if (rc != LOOPSB_OK) return fail(rc);
    const size_t steps_b = im.steps.size() * 4, base_b = im.stream_base.size() * 4;
    if (cudaMalloc(&d->steps, steps_b ? steps_b : 16) != cudaSuccess ||
        cudaMalloc(&d->stream_base, base_b) != cudaSuccess ||
        cudaMalloc(&d->blk_begin, im.blk_begin.size() * 4) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("device allocation of the band-tiled copy failed (%zu bytes)", steps_b);
      return fail(LOOPSB_ERR_ALLOC);
    }
    if (cudaMemcpy(d->steps, im.steps.data(), steps_b, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d->stream_base, im.stream_base.data(), base_b, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d->blk_begin, im.blk_begin.data(), im.blk_begin.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("upload of the band-tiled copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      return fail(LOOPSB_ERR_CUDA);
    }
    d->g = im.g;
    d->total_steps = im.total_steps; d->real_entries = im.real_entries; d->pad_entries = im.pad_entries;
    d->flagged_entries = im.flagged_entries; d->flagged_steps = im.flagged_steps; d->long_steps = im.long_steps;
  }
  if (!force && d->flagged_steps * 20 > d->total_steps) {
    // rows with long runs inside a band (or too many bands per step) push steps
    // onto the general y-update path; above 5 % of the steps the plain kernel wins
    set_error("band-tiled plan not profitable (%lld of %lld steps need the general y-update path)",
              d->flagged_steps, d->total_steps);
    return fail(LOOPSB_ERR_UNSUPPORTED);
  }
  const bt::geom& fg = d->g;   // the final geometry (row blocks cut, rb known)
  d->smem = fg.smem_bytes();
  const size_t steps_b = size_t(d->total_steps + fg.es) * bt::kStepWords * 4, base_b = (size_t(fg.nstreams()) + 1) * 4;
  const size_t part_b = fg.q > 1 ? size_t(fg.q) * fg.nb * ((fg.rb + 3) & ~3) * 4 : 0;
  if ((part_b && cudaMalloc(&d->partial, part_b) != cudaSuccess) ||
      (part_b && cudaMalloc(&d->counters, size_t(fg.nb) * 8) != cudaSuccess) ||
      (part_b && cudaMemset(d->counters, 0, size_t(fg.nb) * 8) != cudaSuccess)) {
    (void)cudaGetLastError();
    set_error("device allocation of the partial-row workspace failed (%zu bytes)", part_b);
    return fail(LOOPSB_ERR_ALLOC);
  }
  if (getenv("LOOPSB_DEBUG_PHASES")) {
    if (cudaMalloc(&d->prof, (size_t(fg.nstreams()) * 8 + size_t(fg.grid()) * 4) * sizeof(long long)) != cudaSuccess) { (void)cudaGetLastError(); d->prof = nullptr; }
    else cudaMemset(d->prof, 0, (size_t(fg.nstreams()) * 8 + size_t(fg.grid()) * 4) * sizeof(long long));
  }
  bt::kernel_fn k = bt::kernel_for(fg.warps, fg.es, d->prof != nullptr);
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, d->smem) != cudaSuccess) {
    set_error("cannot opt in to %d bytes of dynamic shared memory", d->smem);
    (void)cudaGetLastError();
    return fail(LOOPSB_ERR_CUDA);
  }
  {
    // one wave? then the launch can be cooperative and the q CTAs of a row block share the reduction
    int per_sm = 0, coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, plan->device);
    if (coop && fg.q > 1 && !getenv("LOOPSB_TILED_NO_PEERS") &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, fg.cta_threads(), size_t(d->smem)) == cudaSuccess &&
        per_sm * dp->sm_count >= fg.grid())
      d->peers = 1;
    (void)cudaGetLastError();
  }
  d->key_indices = col_indices;
  d->key_values = values;
  d->bytes = (long long)(steps_b + base_b + part_b + (part_b ? size_t(fg.nb) * 8 : 0));
  if (plan->tiled) bt::destroy(plan->tiled);
  plan->tiled = d;
  return LOOPSB_OK;
}

int loopsb_version(void) { return LOOPSB_VERSION; }

const char* loopsb_status_string(int status) {
  switch (status) {
    case LOOPSB_OK: return "ok";
    case LOOPSB_ERR_INVALID: return "invalid argument";
    case LOOPSB_ERR_CUDA: return "CUDA error";
    case LOOPSB_ERR_UNSUPPORTED: return "unsupported (schedule, layout, dtype)";
    case LOOPSB_ERR_ALLOC: return "allocation failed";
    default: return "unknown status";
  }
}

const char* loopsb_last_error(void) { return last_error_buffer(); }

int loopsb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  const device_props* p = current_device();
  if (!p) return LOOPSB_ERR_CUDA;
  if (sm_count) *sm_count = p->sm_count;
  if (cc_major) *cc_major = p->cc_major;
  if (cc_minor) *cc_minor = p->cc_minor;
  return LOOPSB_OK;
}

int loopsb_work_oriented_grid(int32_t* grid_blocks) {
  LOOPSB_REQUIRE(grid_blocks != nullptr, "grid_blocks is null");
  const device_props* p = current_device();
  if (!p) return LOOPSB_ERR_CUDA;
  int per_sm = 0;
  LOOPSB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &per_sm, sk::spmv_work_oriented_csr, sk::kWorkThreads, 0));
  if (per_sm < 1) per_sm = 1;
  *grid_blocks = per_sm * p->sm_count;
  return LOOPSB_OK;
}

// Schedule selection. The reference's paper picks thread-mapped /
// group-mapped / merge-path from the matrix sizes (plots/data/heuristics.csv: its
// `kernel` column is merge-path exactly when nnz >= 10,000 on all but 4 of 4831
// matrices). Thresholds here are from tools/heuristic_sweep.py on B200
// (profiles/heuristic_sweep_r01.log): below ~10^4 nonzeros every kernel is one
// launch latency and thread_mapped is the cheapest launch; very light rows
// (average degree <= 4) also favour thread_mapped, but only when no row is heavy
// (one lane walks a row); everything else goes to merge_path_flat, whose plan
// additionally takes the band-tiled path when its cost model accepts the matrix.
int loopsb_select_schedule(int32_t num_rows, int32_t num_cols, int64_t nnz, int32_t max_degree,
                           int32_t* schedule) {
  LOOPSB_REQUIRE(schedule != nullptr, "schedule is null");
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0 && nnz >= 0, "negative dimensions");
  int pick = LOOPSB_SCHED_MERGE_PATH_FLAT;
  if (nnz < 10000) {
    pick = LOOPSB_SCHED_THREAD_MAPPED;
  } else if (max_degree >= 0 && max_degree <= 32 && nnz <= 4 * int64_t(num_rows)) {
    pick = LOOPSB_SCHED_THREAD_MAPPED;
  }
  *schedule = pick;
  return LOOPSB_OK;
}

int loopsb_plan_destroy(loopsb_plan_t* plan) {
  if (!plan) return LOOPSB_OK;
  if (plan->coords) cudaFree(plan->coords);
  if (plan->carry_row) cudaFree(plan->carry_row);
  if (plan->carry_val) cudaFree(plan->carry_val);
  if (plan->phases) cudaFree(plan->phases);
  if (plan->tc) bcsr_tc::destroy(plan->tc);
  if (plan->tiled) bt::destroy(plan->tiled);
  if (plan->cell) generic::destroy(plan->cell);
  free_probes(plan);
  delete plan;
  return LOOPSB_OK;
}

int loopsb_plan_create(loopsb_plan_t** out, const loopsb_layout_t* lay,
                       int schedule, void* stream) {
  LOOPSB_REQUIRE(out != nullptr && lay != nullptr, "null argument");
  *out = nullptr;
  LOOPSB_REQUIRE(lay->num_tiles >= 0 && lay->num_atoms >= 0, "negative sizes");
  LOOPSB_REQUIRE(schedule >= LOOPSB_SCHED_MERGE_PATH_FLAT &&
                     schedule <= LOOPSB_SCHED_GROUP_MAPPED,
                 "unknown schedule");
  if (is_offsets_kind(lay->kind))
    LOOPSB_REQUIRE(lay->offsets != nullptr, "offsets-kind layout without offsets");
  if (is_pitch_kind(lay->kind))
    LOOPSB_REQUIRE(lay->pitch >= 0, "negative pitch");
  LOOPSB_REQUIRE((long long)lay->num_tiles + (long long)lay->num_atoms < 0x7fffffffLL,
                 "tiles + atoms must stay below 2^31");
  const device_props* dp = current_device();
  if (!dp) return LOOPSB_ERR_CUDA;
  cudaStream_t s = as_stream(stream);

  loopsb_plan* p = new (std::nothrow) loopsb_plan();
  if (!p) { set_error("host allocation failed"); return LOOPSB_ERR_ALLOC; }
  p->lay = *lay;
  p->schedule = schedule;
  p->sm_count = dp->sm_count;
  cudaGetDevice(&p->device);
  const int T = lay->num_tiles, A = lay->num_atoms;

  auto fail = [&](int code) { loopsb_plan_destroy(p); return code; };

  if (generic::supports(lay->kind, schedule)) {
    // coo x {group, work, merge}, ell x {group, work}: no kernel in the reference tree; the cell is
    // the schedule's setup<> over that layout with the format's per-atom body
    int g = 0;
    if (schedule == LOOPSB_SCHED_WORK_ORIENTED) {
      int rc = loopsb_work_oriented_grid(&g);
      if (rc != LOOPSB_OK) return fail(rc);
      p->wo_grid = g;
    }
    int rc = generic::create(&p->cell, lay, schedule, g, s);
    if (rc != LOOPSB_OK) return fail(rc);
    p->cta_threads = 128;
    p->launches = 2;
    p->grid = schedule == LOOPSB_SCHED_WORK_ORIENTED ? g : (T + 127) / 128;
    *out = p;
    return LOOPSB_OK;
  }

  if (schedule == LOOPSB_SCHED_MERGE_PATH_FLAT) {
    const bool array_ends = lay->kind == LOOPSB_LAYOUT_CSR;
    if (!array_ends && lay->kind != LOOPSB_LAYOUT_ELL) {
      set_error("merge_path_flat SpMV exists for CSR and ELL (reference has no other)");
      return fail(LOOPSB_ERR_UNSUPPORTED);
    }
    const long long W = (long long)T + A;
    p->M = (W + mp::kRefItemsPerMergeTile - 1) / mp::kRefItemsPerMergeTile;
    p->gen = merge_generation_from_env();
    p->variant = merge_variant_from_env();
    if (p->gen == 2 && kMergeVariants[p->variant].g != 1) p->variant = 9;   // fallback kernel shares the 1-tile carries
    const merge_variant& mv = kMergeVariants[p->variant];
    p->num_cta_tiles = int((p->M + mv.g - 1) / mv.g);
    p->cta_threads = mv.threads;
    p->smem_bytes = mv.smem;
    p->launches = 2;
    int grid = mv.ctas_per_sm * dp->sm_count;
    if (grid > p->num_cta_tiles) grid = p->num_cta_tiles;
    p->grid = grid;
    if (p->gen == 2) {
      p->variant2 = merge2_variant_from_env();
    }
    if (p->M > 0) {
      const size_t cbytes = size_t(p->M + 1) * sizeof(int2);
      const size_t nct = size_t(p->num_cta_tiles);
      if (cudaMalloc(&p->coords, cbytes) != cudaSuccess ||
          cudaMalloc(&p->carry_row, nct * sizeof(int)) != cudaSuccess ||
          cudaMalloc(&p->carry_val, nct * sizeof(float)) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("device allocation of the merge-path workspace failed");
        return fail(LOOPSB_ERR_ALLOC);
      }
      p->workspace_bytes = (long long)(cbytes + nct * 8);
      const int threads = 128;
      const int blocks = int((p->M + 1 + threads - 1) / threads);
      if (array_ends)
        mp::merge_coords_kernel<true><<<blocks, threads, 0, s>>>(
            lay->offsets + 1, 0, T, A, mp::kRefItemsPerMergeTile, int(p->M), p->coords);
      else
        mp::merge_coords_kernel<false><<<blocks, threads, 0, s>>>(
            nullptr, lay->pitch, T, A, mp::kRefItemsPerMergeTile, int(p->M), p->coords);
      if (cudaGetLastError() != cudaSuccess) {
        set_error("merge_coords_kernel launch failed");
        return fail(LOOPSB_ERR_CUDA);
      }
      if (getenv("LOOPSB_DEBUG_PHASES")) {
        if (cudaMalloc(&p->phases, size_t(p->grid) * 8 * sizeof(long long)) != cudaSuccess) {
          (void)cudaGetLastError();
          p->phases = nullptr;
        } else {
          cudaMemsetAsync(p->phases, 0, size_t(p->grid) * 8 * sizeof(long long), s);
        }
      }
      if (mv.prepare(p->smem_bytes) != cudaSuccess ||
          (p->gen == 2 && prepare_merge2(p->variant2) != cudaSuccess)) {
        set_error("cannot opt in to %d bytes of dynamic shared memory", p->smem_bytes);
        (void)cudaGetLastError();
        return fail(LOOPSB_ERR_CUDA);
      }
      // The coordinates are written on `s`; SpMV calls may come on any stream
      // (the Python mirror caches plans per container): finish them here.
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        set_error("merge-path plan set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(LOOPSB_ERR_CUDA);
      }
    }
  } else if (schedule == LOOPSB_SCHED_THREAD_MAPPED) {
    if (!(lay->kind == LOOPSB_LAYOUT_CSR || lay->kind == LOOPSB_LAYOUT_COO ||
          lay->kind == LOOPSB_LAYOUT_ELL || lay->kind == LOOPSB_LAYOUT_BCSR ||
          lay->kind == LOOPSB_LAYOUT_CSC || lay->kind == LOOPSB_LAYOUT_FLAT)) {
      set_error("thread_mapped SpMV exists for CSR, CSC, COO, ELL, BCSR and flat_uniform_occupancy here "
                "(DIA: loopsb_spmv_dia_f32)");
      return fail(LOOPSB_ERR_UNSUPPORTED);
    }
    if (lay->kind == LOOPSB_LAYOUT_FLAT && !(lay->offsets != nullptr && lay->pitch > 0)) {
      set_error("flat_uniform_occupancy needs the base CSR offsets and K > 0");
      return fail(LOOPSB_ERR_INVALID);
    }
    p->cta_threads = 128;
    p->grid = (T + 127) / 128;
    p->launches = (lay->kind == LOOPSB_LAYOUT_COO || lay->kind == LOOPSB_LAYOUT_CSC ||
                   lay->kind == LOOPSB_LAYOUT_FLAT) ? 2 : 1;
    if (lay->kind == LOOPSB_LAYOUT_BCSR) {
      int rc = bcsr_tc::create(&p->tc, lay, dp->sm_count, s);
      if (rc != LOOPSB_OK) return fail(rc);
      p->workspace_bytes = bcsr_tc::workspace_bytes(p->tc);
    }
  } else if (schedule == LOOPSB_SCHED_GROUP_MAPPED) {
    if (lay->kind != LOOPSB_LAYOUT_CSR) {
      set_error("group_mapped SpMV exists for CSR (reference has no other)");
      return fail(LOOPSB_ERR_UNSUPPORTED);
    }
    p->cta_threads = sk::kGroupThreads;
    p->grid = (T + sk::kGroupThreads - 1) / sk::kGroupThreads;
  } else {  // work_oriented
    if (lay->kind != LOOPSB_LAYOUT_CSR) {
      set_error("work_oriented SpMV exists for CSR (reference has no other)");
      return fail(LOOPSB_ERR_UNSUPPORTED);
    }
    // Schedule geometry as the reference: N = occupancy grid x 128 threads,
    // every thread owns w = ceil(W / N) consecutive merge items, so block B
    // owns items [128 w B, 128 w (B+1)). The SpMV runs each block's range
    // through the cooperative merge-tile kernel in pieces of <= 1024 items:
    // same partition at block granularity, deterministic, y overwritten.
    int g = 0;
    int rc = loopsb_work_oriented_grid(&g);
    if (rc != LOOPSB_OK) return fail(rc);
    p->wo_grid = g;
    const long long W = (long long)T + A;
    const long long N = (long long)g * sk::kWorkThreads;
    const long long w = (W + N - 1) / N;
    const long long per_block = w * sk::kWorkThreads;
    std::vector<long long> diags;
    if (W > 0) {
      for (long long d0 = 0; d0 < W; d0 += per_block) {
        const long long d1 = d0 + per_block < W ? d0 + per_block : W;
        for (long long d = d0; d < d1; d += mp::kRefItemsPerMergeTile) diags.push_back(d);
      }
      diags.push_back(W);
    }
    p->gen = merge_generation_from_env();
    p->variant = merge_variant_from_env();
    const merge_variant& mv = kMergeVariants[p->variant];
    if (mv.g != 1) { p->variant = 9; }
    const merge_variant& mv1 = kMergeVariants[p->variant];
    p->M = diags.empty() ? 0 : (long long)diags.size() - 1;
    p->num_cta_tiles = int(p->M);
    p->cta_threads = mv1.threads;
    p->smem_bytes = mv1.smem;
    p->launches = 2;
    int grid = mv1.ctas_per_sm * dp->sm_count;
    if (grid > p->num_cta_tiles) grid = p->num_cta_tiles;
    p->grid = grid;
    if (p->gen == 2) {
      p->variant2 = merge2_variant_from_env();
    }
    if (p->M > 0) {
      long long* d_diags = nullptr;
      const size_t nct = size_t(p->num_cta_tiles);
      if (cudaMalloc(&d_diags, diags.size() * sizeof(long long)) != cudaSuccess ||
          cudaMalloc(&p->coords, size_t(p->M + 1) * sizeof(int2)) != cudaSuccess ||
          cudaMalloc(&p->carry_row, nct * sizeof(int)) != cudaSuccess ||
          cudaMalloc(&p->carry_val, nct * sizeof(float)) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(d_diags);
        set_error("device allocation of the work_oriented workspace failed");
        return fail(LOOPSB_ERR_ALLOC);
      }
      cudaMemcpyAsync(d_diags, diags.data(), diags.size() * sizeof(long long), cudaMemcpyHostToDevice, s);
      const int threads = 128;
      mp::merge_coords_kernel<true><<<int((p->M + 1 + threads - 1) / threads), threads, 0, s>>>(
          lay->offsets + 1, 0, T, A, 0, int(p->M), p->coords, d_diags);
      cudaError_t e = cudaStreamSynchronize(s);   // diags (host vector) and d_diags die here
      cudaFree(d_diags);
      if (e != cudaSuccess || mv1.prepare(p->smem_bytes) != cudaSuccess ||
          (p->gen == 2 && prepare_merge2(p->variant2) != cudaSuccess)) {
        set_error("work_oriented plan set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(LOOPSB_ERR_CUDA);
      }
      p->workspace_bytes = (long long)(size_t(p->M + 1) * sizeof(int2) + nct * 8);
    }
  }
  *out = p;
  return LOOPSB_OK;
}

int loopsb_plan_invalidate(loopsb_plan_t* plan) {
  LOOPSB_REQUIRE(plan != nullptr, "plan is null");
  // the caller changed values / column ids IN PLACE: every plan-owned copy of them is stale.
  // Dropping them is always safe -- SpMV calls fall back to the kernels that read the live arrays.
  if (plan->tiled) { bt::destroy(plan->tiled); plan->tiled = nullptr; }
  if (plan->tc) bcsr_tc::unpack(plan->tc);
  return LOOPSB_OK;
}

int loopsb_plan_tile_breakeven(const loopsb_plan_t* plan, int32_t num_cols, int64_t* calls) {
  LOOPSB_REQUIRE(plan != nullptr && calls != nullptr, "null argument");
  *calls = -1;
  if (plan->schedule != LOOPSB_SCHED_MERGE_PATH_FLAT || plan->lay.kind != LOOPSB_LAYOUT_CSR) return LOOPSB_OK;
  const device_props* dp = current_device();
  if (!dp) return LOOPSB_ERR_CUDA;
  const double nnz = double(plan->lay.num_atoms), rows = double(plan->lay.num_tiles);
  if (nnz < double(1 << 22) || rows == 0 || num_cols <= 0) return LOOPSB_OK;
  bt::geom g = bt::choose_geom(plan->lay.num_tiles, num_cols, dp->sm_count, dp->max_smem_optin);
  const char* why = "";
  if (!bt::derive(g, plan->lay.num_tiles, num_cols, &why)) return LOOPSB_OK;
  if (double(g.nb) * double(num_cols) * 4.0 > nnz * 8.0) return LOOPSB_OK;      // the cost model would decline
  // Measured on B200 (profiles/): device build ~2.2 ns per nonzero + ~20 ms fixed; plain CSR kernel
  // ~205 Gnnz/s; band-tiled kernel ~4.6 TB/s of its 8.4 B per nonzero.
  const double build_s = 0.020 + 2.2e-9 * nnz;
  const double plain_s = nnz / 205e9, tiled_s = nnz * 8.4 / 4.6e12;
  if (plain_s <= tiled_s) return LOOPSB_OK;
  *calls = int64_t(build_s / (plain_s - tiled_s)) + 1;
  return LOOPSB_OK;
}

int loopsb_plan_hint_x_bytes(loopsb_plan_t* plan, int64_t bytes) {
  LOOPSB_REQUIRE(plan != nullptr, "plan is null");
  plan->x_span_bytes = bytes;
  return LOOPSB_OK;
}

int loopsb_plan_info(const loopsb_plan_t* plan, loopsb_plan_info_t* info) {
  LOOPSB_REQUIRE(plan != nullptr && info != nullptr, "null argument");
  memset(info, 0, sizeof(*info));
  info->schedule = plan->schedule;
  info->layout_kind = plan->lay.kind;
  info->threads_per_block = 128;
  info->items_per_thread = plan->schedule == LOOPSB_SCHED_MERGE_PATH_FLAT ? 8 : 1;
  info->num_merge_tiles = plan->M;
  const bool gen2 = plan->gen == 2 && (plan->schedule == LOOPSB_SCHED_MERGE_PATH_FLAT ||
                                       plan->schedule == LOOPSB_SCHED_WORK_ORIENTED);
  const merge2_variant& m2 = kMerge2Variants[merge2_pick(plan->variant2, 0)];   // geometry for an L2-sized x
  info->grid_blocks = gen2 ? std::min(m2.ctas_per_sm * plan->sm_count, plan->num_cta_tiles) : plan->grid;
  info->cta_threads = gen2 ? m2.threads : plan->cta_threads;
  info->launches_per_spmv = plan->launches;
  info->smem_bytes = gen2 ? m2.smem : plan->smem_bytes;
  info->workspace_bytes = plan->workspace_bytes;
  return LOOPSB_OK;
}

int loopsb_plan_merge_coords_host(const loopsb_plan_t* plan, int32_t* host_xy,
                                  int64_t capacity_pairs) {
  LOOPSB_REQUIRE(plan != nullptr && host_xy != nullptr, "null argument");
  LOOPSB_REQUIRE(plan->schedule == LOOPSB_SCHED_MERGE_PATH_FLAT, "not a merge-path plan");
  LOOPSB_REQUIRE(capacity_pairs >= plan->M + 1, "host buffer too small");
  if (plan->M == 0) { host_xy[0] = 0; host_xy[1] = 0; return LOOPSB_OK; }
  LOOPSB_CUDA_TRY(cudaDeviceSynchronize());
  LOOPSB_CUDA_TRY(cudaMemcpy(host_xy, plan->coords, size_t(plan->M + 1) * sizeof(int2),
                             cudaMemcpyDeviceToHost));
  return LOOPSB_OK;
}

int loopsb_plan_debug_phases_host(const loopsb_plan_t* plan, int64_t* host_out,
                                  int64_t capacity_ctas) {
  LOOPSB_REQUIRE(plan != nullptr && host_out != nullptr, "null argument");
  if (plan->tiled && plan->tiled->prof) {   // band-tiled kernel: 8 counters per consumer warp
    // 8 counters per consumer warp, then (if the buffer has room) 4 wall-clock stamps per CTA
    const long long n = plan->tiled->g.nstreams();
    const long long extra = (plan->tiled->g.grid() * 4 + 7) / 8;
    LOOPSB_REQUIRE(capacity_ctas >= n, "host buffer too small");
    const long long take = capacity_ctas >= n + extra ? n * 8 + plan->tiled->g.grid() * 4 : n * 8;
    LOOPSB_CUDA_TRY(cudaDeviceSynchronize());
    LOOPSB_CUDA_TRY(cudaMemcpy(host_out, plan->tiled->prof, size_t(take) * sizeof(long long), cudaMemcpyDeviceToHost));
    return LOOPSB_OK;
  }
  if (!plan->phases) { set_error("phase counters are off (set LOOPSB_DEBUG_PHASES=1 before creating the plan)"); return LOOPSB_ERR_UNSUPPORTED; }
  LOOPSB_REQUIRE(capacity_ctas >= plan->grid, "host buffer too small");
  LOOPSB_CUDA_TRY(cudaDeviceSynchronize());
  LOOPSB_CUDA_TRY(cudaMemcpy(host_out, plan->phases, size_t(plan->grid) * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  return LOOPSB_OK;
}

int loopsb_plan_probe_begin(loopsb_plan_t* plan, int32_t capacity) {
  LOOPSB_REQUIRE(plan != nullptr && capacity > 0 && capacity <= (1 << 20), "bad probe request");
  free_probes(plan);
  plan->probe_ev = new (std::nothrow) cudaEvent_t[2 * size_t(capacity)]();
  if (!plan->probe_ev) { set_error("host allocation failed"); return LOOPSB_ERR_ALLOC; }
  plan->probe_cap = capacity;
  for (int i = 0; i < 2 * capacity; ++i)
    LOOPSB_CUDA_TRY(cudaEventCreate(&plan->probe_ev[i]));
  plan->probing = true;
  return LOOPSB_OK;
}

int loopsb_plan_probe_collect(loopsb_plan_t* plan, float* host_ms,
                              int32_t capacity, int32_t* n) {
  LOOPSB_REQUIRE(plan != nullptr && host_ms != nullptr && n != nullptr, "null argument");
  *n = 0;
  plan->probing = false;
  int count = plan->probe_n < capacity ? plan->probe_n : capacity;
  for (int i = 0; i < count; ++i) {
    LOOPSB_CUDA_TRY(cudaEventSynchronize(plan->probe_ev[2 * i + 1]));
    LOOPSB_CUDA_TRY(cudaEventElapsedTime(&host_ms[i], plan->probe_ev[2 * i], plan->probe_ev[2 * i + 1]));
  }
  *n = count;
  free_probes(plan);
  return LOOPSB_OK;
}

int loopsb_spmv_f32(loopsb_plan_t* plan, const float* values,
                    const int32_t* col_indices, const int32_t* row_indices,
                    const float* x, float* y, int32_t num_rows,
                    int32_t num_cols, void* stream) {
  LOOPSB_REQUIRE(plan != nullptr, "plan is null");
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0, "negative dimensions");
  const loopsb_layout_t& lay = plan->lay;
  const int T = lay.num_tiles, A = lay.num_atoms;
  LOOPSB_REQUIRE(num_rows == 0 || y != nullptr, "y is null");
  LOOPSB_REQUIRE(A == 0 || (values && col_indices && x), "null matrix / x pointer");
  cudaStream_t s = as_stream(stream);
  if (num_rows == 0) return LOOPSB_OK;
  if (A == 0) {  // nothing stored: y = 0
    LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(float), s));
    return LOOPSB_OK;
  }

  if (plan->cell) {
    probe_scope probe(plan, s);
    const int rc = generic::run(plan->cell, &lay, plan->schedule, values, col_indices, row_indices, x, y, num_rows, s);
    probe.close();
    return rc;
  }

  switch (plan->schedule) {
    case LOOPSB_SCHED_MERGE_PATH_FLAT: {
      LOOPSB_REQUIRE(T == num_rows, "layout tiles must equal num_rows");
      if (plan->tiled && plan->tiled->key_indices == col_indices && plan->tiled->key_values == values &&
          plan->tiled->g.cols == num_cols && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
        probe_scope probe(plan, s);
        int rc = launch_tiled(plan->tiled, x, y, s);
        probe.close();
        return rc;
      }
      const int nct = plan->num_cta_tiles;
      l2_pin_scope pin(plan, x, size_t(num_cols) * sizeof(float), s);
      probe_scope probe(plan, s);
      const merge_variant& mv = kMergeVariants[plan->variant];
      const bool array_ends = lay.kind == LOOPSB_LAYOUT_CSR;
      if (plan->gen == 2 && !plan->phases && aligned16(col_indices, values))
        launch_merge2(plan, array_ends, false, s, col_indices, values, x, y, num_cols);
      else
        mv.launch(array_ends, plan->grid, plan->smem_bytes, s, array_ends ? lay.offsets + 1 : nullptr,
                  lay.pitch, col_indices, values, x, y, plan->coords, int(plan->M), T, A, nct,
                  plan->carry_row, plan->carry_val, plan->phases);
      probe.close();
      LOOPSB_CUDA_TRY(cudaGetLastError());
      mp::spmv_merge_fixup_kernel<<<(nct + 255) / 256, 256, 0, s>>>(
          plan->carry_row, plan->carry_val, nct, T, y);
      LOOPSB_CUDA_TRY(cudaGetLastError());
      return LOOPSB_OK;
    }
    case LOOPSB_SCHED_THREAD_MAPPED: {
      if (lay.kind == LOOPSB_LAYOUT_COO) {
        LOOPSB_REQUIRE(row_indices != nullptr, "COO needs row_indices");
        LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(float), s));
      }
      if (lay.kind == LOOPSB_LAYOUT_CSC || lay.kind == LOOPSB_LAYOUT_FLAT)   // scatter kernels accumulate into y
        LOOPSB_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(num_rows) * sizeof(float), s));
      probe_scope probe(plan, s);
      if (lay.kind == LOOPSB_LAYOUT_CSR) {
        LOOPSB_REQUIRE(T == num_rows, "layout tiles must equal num_rows");
        sk::spmv_thread_mapped_csr<<<plan->grid, 128, 0, s>>>(
            lay.offsets, col_indices, values, x, y, num_rows);
      } else if (lay.kind == LOOPSB_LAYOUT_COO) {
        sk::spmv_coo_thread_mapped<<<(A + 127) / 128, 128, 0, s>>>(
            row_indices, col_indices, values, x, y, A);
      } else if (lay.kind == LOOPSB_LAYOUT_ELL) {
        LOOPSB_REQUIRE(T == num_rows, "layout tiles must equal num_rows");
        sk::spmv_ell_thread_mapped<<<plan->grid, 128, 0, s>>>(
            col_indices, values, x, y, num_rows, lay.pitch);
      } else if (lay.kind == LOOPSB_LAYOUT_CSC) {
        // tiles are columns; `col_indices` carries the ROW index of every stored entry
        LOOPSB_REQUIRE(T == num_cols, "CSC layout tiles must equal num_cols");
        sk::spmv_csc_thread_mapped<<<plan->grid, 128, 0, s>>>(lay.offsets, col_indices, values, x, y, num_cols);
      } else if (lay.kind == LOOPSB_LAYOUT_FLAT) {
        // tiles are windows of K = pitch atoms over the base CSR whose offsets the descriptor carries
        sk::spmv_flat_partitioned<<<plan->grid, 128, 0, s>>>(lay.offsets, col_indices, values, x, y, num_rows, A,
                                                             lay.pitch);
      } else {
        set_error("use loopsb_spmv_bcsr_f32 / loopsb_spmv_bcsr4x4_bf16 for BCSR");
        return LOOPSB_ERR_UNSUPPORTED;
      }
      probe.close();
      LOOPSB_CUDA_TRY(cudaGetLastError());
      return LOOPSB_OK;
    }
    case LOOPSB_SCHED_GROUP_MAPPED: {
      LOOPSB_REQUIRE(T == num_rows, "layout tiles must equal num_rows");
      probe_scope probe(plan, s);
      sk::spmv_group_mapped_csr<<<plan->grid, sk::kGroupThreads, 0, s>>>(
          lay.offsets, col_indices, values, x, y, num_rows);
      probe.close();
      LOOPSB_CUDA_TRY(cudaGetLastError());
      return LOOPSB_OK;
    }
    case LOOPSB_SCHED_WORK_ORIENTED: {
      LOOPSB_REQUIRE(T == num_rows, "layout tiles must equal num_rows");
      const int nct = plan->num_cta_tiles;
      const merge_variant& mv = kMergeVariants[plan->variant];
      l2_pin_scope pin(plan, x, size_t(num_cols) * sizeof(float), s);
      probe_scope probe(plan, s);
      if (plan->gen == 2 && aligned16(col_indices, values))
        launch_merge2(plan, true, false, s, col_indices, values, x, y, num_cols);
      else
        mv.launch(true, plan->grid, plan->smem_bytes, s, lay.offsets + 1, 0, col_indices, values, x, y,
                  plan->coords, int(plan->M), T, A, nct, plan->carry_row, plan->carry_val, nullptr);
      probe.close();
      LOOPSB_CUDA_TRY(cudaGetLastError());
      mp::spmv_merge_fixup_kernel<<<(nct + 255) / 256, 256, 0, s>>>(plan->carry_row, plan->carry_val, nct, T, y);
      LOOPSB_CUDA_TRY(cudaGetLastError());
      return LOOPSB_OK;
    }
    default:
      set_error("unknown schedule in plan");
      return LOOPSB_ERR_INVALID;
  }
}

int loopsb_spmv_acc_f32(loopsb_plan_t* plan, const float* values, const int32_t* col_indices, const float* x,
                        float* y, int32_t num_rows, int32_t num_cols, void* stream) {
  LOOPSB_REQUIRE(plan != nullptr, "plan is null");
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0, "negative dimensions");
  LOOPSB_REQUIRE(plan->schedule == LOOPSB_SCHED_MERGE_PATH_FLAT && plan->lay.kind == LOOPSB_LAYOUT_CSR,
                 "y += A x exists for merge_path_flat plans over CSR");
  const int T = plan->lay.num_tiles, A = plan->lay.num_atoms;
  LOOPSB_REQUIRE(T == num_rows, "layout tiles must equal num_rows");
  if (num_rows == 0 || A == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(values && col_indices && x && y, "null matrix / x / y pointer");
  if (plan->gen != 2 || !aligned16(col_indices, values)) {
    set_error("y += A x needs the second-generation merge kernel and 16-byte aligned indices / values");
    return LOOPSB_ERR_UNSUPPORTED;
  }
  cudaStream_t s = as_stream(stream);
  const int nct = plan->num_cta_tiles;
  l2_pin_scope pin(plan, x, size_t(num_cols) * sizeof(float), s);
  probe_scope probe(plan, s);
  launch_merge2(plan, true, true, s, col_indices, values, x, y, num_cols);
  probe.close();
  LOOPSB_CUDA_TRY(cudaGetLastError());
  mp::spmv_merge_fixup_kernel<<<(nct + 255) / 256, 256, 0, s>>>(plan->carry_row, plan->carry_val, nct, T, y);
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}

int loopsb_spmv_bcsr_f32(int32_t R, int32_t C, const loopsb_layout_t* lay,
                         const float* values, const int32_t* block_col_indices,
                         const float* x_padded, float* y, int32_t num_rows,
                         void* stream) {
  LOOPSB_REQUIRE(lay != nullptr && lay->kind == LOOPSB_LAYOUT_BCSR, "BCSR layout required");
  LOOPSB_REQUIRE(lay->offsets != nullptr, "block offsets are null");
  LOOPSB_REQUIRE(num_rows >= 0, "negative rows");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  if (num_rows == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(y != nullptr, "y is null");
  const int br = lay->num_tiles;
  LOOPSB_REQUIRE(lay->num_atoms == 0 || (values && block_col_indices && x_padded),
                 "null matrix / x pointer");
  cudaStream_t s = as_stream(stream);
  const int grid = (br + 127) / 128;
  if (grid == 0) return LOOPSB_OK;
  if (R == 2 && C == 2)
    sk::spmv_bcsr_thread_mapped<2, 2><<<grid, 128, 0, s>>>(lay->offsets, block_col_indices, values, x_padded, y, br, num_rows);
  else if (R == 3 && C == 3)
    sk::spmv_bcsr_thread_mapped<3, 3><<<grid, 128, 0, s>>>(lay->offsets, block_col_indices, values, x_padded, y, br, num_rows);
  else if (R == 4 && C == 4)
    sk::spmv_bcsr_thread_mapped<4, 4><<<grid, 128, 0, s>>>(lay->offsets, block_col_indices, values, x_padded, y, br, num_rows);
  else {
    set_error("BCSR block shape %dx%d not instantiated (2x2, 3x3, 4x4)", R, C);
    return LOOPSB_ERR_UNSUPPORTED;
  }
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}

int loopsb_spmv_dia_f32(int32_t num_rows, int32_t num_cols, int64_t stride, int32_t num_diagonals,
                        const int32_t* diag_offsets, const float* values, const float* x, float* y,
                        void* stream) {
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0 && num_diagonals >= 0 && stride >= num_rows,
                 "negative size or stride shorter than the rows");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  if (num_rows == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(y != nullptr, "y is null");
  LOOPSB_REQUIRE(num_diagonals == 0 || (diag_offsets && values && x), "null matrix / x pointer");
  cudaStream_t s = as_stream(stream);
  sk::spmv_dia_thread_mapped<<<(num_rows + 127) / 128, 128, 0, s>>>(diag_offsets, values, x, y, num_rows, num_cols,
                                                                   stride, num_diagonals);
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}

int loopsb_spmm_csr_f32(const loopsb_layout_t* lay, const float* values, const int32_t* col_indices,
                        const float* B, float* C, int32_t num_rows, int32_t num_cols, int32_t n,
                        void* stream) {
  LOOPSB_REQUIRE(lay != nullptr && lay->kind == LOOPSB_LAYOUT_CSR && lay->offsets != nullptr, "CSR layout required");
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0 && n >= 0 && lay->num_tiles == num_rows, "bad dimensions");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  if (num_rows == 0 || n == 0) return LOOPSB_OK;
  LOOPSB_REQUIRE(C != nullptr, "C is null");
  LOOPSB_REQUIRE(lay->num_atoms == 0 || (values && col_indices && B), "null matrix / B pointer");
  const dim3 grid((num_rows + 3) / 4, (n + 31) / 32);
  LOOPSB_REQUIRE(grid.y <= 65535u, "more than 2M dense columns");
  sk::spmm_csr_row_warp<<<grid, 128, 0, as_stream(stream)>>>(lay->offsets, col_indices, values, B, C, num_rows, n);
  LOOPSB_CUDA_TRY(cudaGetLastError());
  return LOOPSB_OK;
}

int loopsb_spmv_bcsr4x4_bf16(loopsb_plan_t* plan, const uint16_t* values_bf16,
                             const int32_t* block_col_indices,
                             const uint16_t* x_bf16_padded, float* y,
                             int32_t num_rows, void* stream) {
  LOOPSB_REQUIRE(plan != nullptr && plan->tc != nullptr,
                 "plan was not created from a BCSR layout with thread_mapped");
  return bcsr_tc::run(plan->tc, &plan->lay, values_bf16, block_col_indices,
                      x_bf16_padded, y, num_rows, as_stream(stream));
}

int loopsb_plan_pack_bcsr4x4(loopsb_plan_t* plan, const uint16_t* values_bf16, const int32_t* block_col_indices,
                             void* stream) {
  LOOPSB_REQUIRE(plan != nullptr && plan->tc != nullptr,
                 "plan was not created from a BCSR layout with thread_mapped");
  return bcsr_tc::pack(plan->tc, values_bf16, block_col_indices, as_stream(stream));
}

int loopsb_spmv_csr_host_f32(int schedule, int32_t num_rows, int32_t num_cols,
                             int32_t nnz, const int32_t* host_offsets,
                             const int32_t* host_indices,
                             const float* host_values, const float* host_x,
                             float* host_y, float* kernel_ms) {
  LOOPSB_REQUIRE(num_rows >= 0 && num_cols >= 0 && nnz >= 0, "negative sizes");
  LOOPSB_REQUIRE(host_offsets && (nnz == 0 || (host_indices && host_values)) &&
                     (num_cols == 0 || host_x) && (num_rows == 0 || host_y),
                 "null host pointer");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  int *d_off = nullptr, *d_idx = nullptr;
  float *d_val = nullptr, *d_x = nullptr, *d_y = nullptr;
  loopsb_plan_t* plan = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = LOOPSB_OK;
  auto cleanup = [&]() {
    if (plan) loopsb_plan_destroy(plan);
    cudaFree(d_off); cudaFree(d_idx); cudaFree(d_val); cudaFree(d_x); cudaFree(d_y);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define HOST_TRY(expr)                                                        \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      set_error("%s failed: %s", #expr, cudaGetErrorString(e__));             \
      cleanup();                                                              \
      return LOOPSB_ERR_CUDA;                                                 \
    }                                                                         \
  } while (0)
  HOST_TRY(cudaMalloc(&d_off, size_t(num_rows + 1) * 4));
  HOST_TRY(cudaMalloc(&d_idx, size_t(nnz ? nnz : 1) * 4));
  HOST_TRY(cudaMalloc(&d_val, size_t(nnz ? nnz : 1) * 4));
  HOST_TRY(cudaMalloc(&d_x, size_t(num_cols ? num_cols : 1) * 4));
  HOST_TRY(cudaMalloc(&d_y, size_t(num_rows ? num_rows : 1) * 4));
  HOST_TRY(cudaMemcpy(d_off, host_offsets, size_t(num_rows + 1) * 4, cudaMemcpyHostToDevice));
  if (nnz) {
    HOST_TRY(cudaMemcpy(d_idx, host_indices, size_t(nnz) * 4, cudaMemcpyHostToDevice));
    HOST_TRY(cudaMemcpy(d_val, host_values, size_t(nnz) * 4, cudaMemcpyHostToDevice));
  }
  if (num_cols)
    HOST_TRY(cudaMemcpy(d_x, host_x, size_t(num_cols) * 4, cudaMemcpyHostToDevice));
  loopsb_layout_t lay{};
  lay.kind = LOOPSB_LAYOUT_CSR;
  lay.num_tiles = num_rows;
  lay.num_atoms = nnz;
  lay.offsets = d_off;
  rc = loopsb_plan_create(&plan, &lay, schedule, nullptr);
  if (rc != LOOPSB_OK) { cleanup(); return rc; }
  HOST_TRY(cudaEventCreate(&e0));
  HOST_TRY(cudaEventCreate(&e1));
  HOST_TRY(cudaEventRecord(e0, 0));
  rc = loopsb_spmv_f32(plan, d_val, d_idx, nullptr, d_x, d_y, num_rows, num_cols, nullptr);
  if (rc != LOOPSB_OK) { cleanup(); return rc; }
  HOST_TRY(cudaEventRecord(e1, 0));
  HOST_TRY(cudaEventSynchronize(e1));
  if (kernel_ms) HOST_TRY(cudaEventElapsedTime(kernel_ms, e0, e1));
  if (num_rows)
    HOST_TRY(cudaMemcpy(host_y, d_y, size_t(num_rows) * 4, cudaMemcpyDeviceToHost));
#undef HOST_TRY
  cleanup();
  return LOOPSB_OK;
}

}  // extern "C"
