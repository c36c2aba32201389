/* include/loopsb.h -- the C ABI of loops-b200 (libloopsb200.so).
 *
 * gunrock/loops is a header-only C++/CUDA template library: it has no FFI of
 * its own. The "drop-in boundary" is therefore the set of host entry points in
 * reference include/loops/algorithms/spmv/ (each takes an owning container, x, y
 * and a stream) plus the schedule/layout template API those kernels are written
 * against. This header is what those host entry points bind to in loops-b200:
 * plain pointers and sizes, no C++ or torch types, one shared object of
 * hand-written sm_100a kernels behind it. The C++ mirror of the reference API
 * that calls these functions lives in include/loops/ (same namespaces and
 * names); the Python/ctypes binding used by tests/ and bench.py lives in
 * loops_b200/. INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name says `host`.
 *  - ids are int32 (the reference instantiates index_t = offset_t = int
 *    everywhere: examples/spmv/merge_path.cu:18-20); tiles+atoms < 2^31
 *    (reference util/search.hxx:46-47 has the same cap).
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *    Calls are asynchronous on that stream unless stated otherwise.
 *  - Every function returns a loopsb_status_t; nothing throws across the
 *    boundary. loopsb_last_error() gives the CUDA / argument detail.
 *  - y is always fully OVERWRITTEN (rows with no atoms get 0); the reference's
 *    "y must be pre-zeroed" precondition for its atomic kernels
 *    (algorithms/spmv/coo_thread_mapped.cuh:14-15, ell_merge_path.cuh:113-114)
 *    is tolerated but not required.
 *  - There is no CPU fallback: without a CUDA device every compute entry point
 *    returns LOOPSB_ERR_CUDA.
 */
#ifndef LOOPSB_H_
#define LOOPSB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOOPSB_VERSION 100 /* 0.1.0 */

typedef enum {
  LOOPSB_OK = 0,
  LOOPSB_ERR_INVALID = 1,     /* bad argument (null pointer, negative size ...) */
  LOOPSB_ERR_CUDA = 2,        /* CUDA runtime error, see loopsb_last_error()   */
  LOOPSB_ERR_UNSUPPORTED = 3, /* (schedule, layout, dtype) has no kernel       */
  LOOPSB_ERR_ALLOC = 4
} loopsb_status_t;

/* Same order as reference schedule::algorithms_t (schedule.hxx:26-32). */
typedef enum {
  LOOPSB_SCHED_MERGE_PATH_FLAT = 0,
  LOOPSB_SCHED_WORK_ORIENTED = 1,
  LOOPSB_SCHED_THREAD_MAPPED = 2,
  LOOPSB_SCHED_GROUP_MAPPED = 3
} loopsb_schedule_t;

/* Which in-tree layout view a descriptor stands for (container/layout.hxx). */
typedef enum {
  LOOPSB_LAYOUT_CSR = 0,  /* offsets[T+1]                     layout.hxx:87-149  */
  LOOPSB_LAYOUT_COO = 1,  /* tiles == atoms == nnz            layout.hxx:385-421 */
  LOOPSB_LAYOUT_ELL = 2,  /* pitch atoms per tile             layout.hxx:443-496 */
  LOOPSB_LAYOUT_BCSR = 3, /* offsets over block-rows          layout.hxx:239-285 */
  LOOPSB_LAYOUT_CSC = 4,  /* offsets over columns             layout.hxx:312-359 */
  LOOPSB_LAYOUT_DIA = 5,  /* pitch = number of diagonals      layout.hxx:166-217 */
  LOOPSB_LAYOUT_FLAT = 6  /* windows of `pitch` (=K) atoms    partitioning.hxx:71-141 */
} loopsb_layout_kind_t;

/* POD image of a layout view: enough to answer the six contract questions
 * (num_tiles, num_atoms, tile_begin, tile_end, tile_size, tile_end_iter). */
typedef struct loopsb_layout {
  int32_t kind;           /* loopsb_layout_kind_t                              */
  int32_t num_tiles;
  int32_t num_atoms;
  int32_t pitch;          /* ELL/DIA: atoms per tile; FLAT: K; else unused     */
  const int32_t* offsets; /* CSR/CSC/BCSR: device array [num_tiles+1]; else 0  */
} loopsb_layout_t;

/* ---------------------------------------------------------------------------
 * Library / device
 * ------------------------------------------------------------------------- */
int loopsb_version(void);
const char* loopsb_status_string(int status);
const char* loopsb_last_error(void);
/* SM count and max resident threads of the current device. */
int loopsb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---------------------------------------------------------------------------
 * Plan: what the reference calls schedule::merge_path::preprocess_t
 * (schedule/merge_path_flat.hxx:92-172) generalised to every schedule -- the
 * per-matrix, data-independent-of-values set-up: merge-path tile coordinates,
 * work-oriented per-block coordinates, carry-out workspace, launch geometry.
 * Borrowing rule: `lay->offsets` must stay valid while the plan is used.
 * ------------------------------------------------------------------------- */
typedef struct loopsb_plan loopsb_plan_t;

typedef struct loopsb_plan_info {
  int32_t schedule;
  int32_t layout_kind;
  int32_t threads_per_block;   /* reference-visible schedule geometry          */
  int32_t items_per_thread;
  int64_t num_merge_tiles;     /* merge_path_flat: ceil((T+A)/(TPB*IPT))       */
  int32_t grid_blocks;         /* CTAs the SpMV kernel launches                */
  int32_t cta_threads;         /* threads per CTA of the SpMV kernel           */
  int32_t launches_per_spmv;   /* kernels enqueued by one loopsb_spmv_f32 call */
  int32_t smem_bytes;          /* dynamic shared memory per CTA                */
  int64_t workspace_bytes;     /* device memory owned by the plan              */
} loopsb_plan_info_t;

int loopsb_plan_create(loopsb_plan_t** out, const loopsb_layout_t* lay,
                       int schedule, void* stream);
int loopsb_plan_destroy(loopsb_plan_t* plan);
int loopsb_plan_info(const loopsb_plan_t* plan, loopsb_plan_info_t* info);
/* STALENESS CONTRACT of plan-owned copies. Two optional accelerators keep a re-ordered COPY
 * of the matrix inside the plan (loopsb_plan_tile_csr: indices + values; loopsb_plan_pack_bcsr4x4:
 * block values + block columns). They are keyed by the caller's array ADDRESSES, so a change of
 * the arrays' CONTENTS in place cannot be seen by the library. After such a change call
 * loopsb_plan_invalidate: it drops every copy, SpMV calls on the plan then run the kernels
 * that read the live arrays (always correct), and the copy can be rebuilt with another
 * loopsb_plan_tile_csr / loopsb_plan_pack_bcsr4x4 when the values have settled. Plans WITHOUT
 * such a copy never hold matrix values. (The Python mirror tracks torch's in-place version
 * counters and calls this by itself; the C++ mirror exposes csr.values_changed().) */
int loopsb_plan_invalidate(loopsb_plan_t* plan);
/* How many SpMV calls it takes for loopsb_plan_tile_csr to pay for itself on this matrix
 * (build time / per-call saving, from rates measured on B200); -1 = never (the cost model
 * declines the matrix). The C++ mirror tiles a cached plan once that many calls were made. */
int loopsb_plan_tile_breakeven(const loopsb_plan_t* plan, int32_t num_cols, int64_t* calls);
/* Tuning hint: the number of bytes of x the matrix's columns range over when that is
 * less than num_cols * sizeof(value) (a column block of a larger matrix). Only the launch
 * geometry depends on it (resident CTAs per SM vs L1 capacity), never the result. */
int loopsb_plan_hint_x_bytes(loopsb_plan_t* plan, int64_t bytes);
/* Copy the merge-path tile start coordinates S(b * TPB*IPT), b = 0..M, to a
 * HOST array of 2*(M+1) int32 (x0,y0,x1,y1,...). Synchronises. This is what
 * the reference's generate_search_coordinates (merge_path_flat.hxx:45-76)
 * materialises; exposed for parity tests. */
int loopsb_plan_merge_coords_host(const loopsb_plan_t* plan, int32_t* host_xy,
                                  int64_t capacity_pairs);

/* Kernel-time probes (measurement aid, off by default). After
 * loopsb_plan_probe_begin(plan, capacity) every loopsb_spmv_f32 call on this
 * plan brackets its DOMINANT kernel (the merge-path kernel, or the single
 * kernel of the other schedules) with a cudaEvent pair recorded on the call's
 * stream. loopsb_plan_probe_collect synchronises, writes up to `capacity`
 * per-launch durations in milliseconds to a HOST array, returns the count in
 * *n and switches the probes off again. */
int loopsb_plan_probe_begin(loopsb_plan_t* plan, int32_t capacity);
/* Tuning aid: when the environment variable LOOPSB_DEBUG_PHASES is set at plan
 * creation, the merge-path kernel accumulates, per CTA (thread 0), the SM
 * cycles spent in {stage wait, gather, walk, scan, store, refill} and the
 * number of tiles; this copies the last launch's 8 x int64 per CTA to the host
 * (grid_blocks CTAs). LOOPSB_ERR_UNSUPPORTED when the counters are off. */
int loopsb_plan_debug_phases_host(const loopsb_plan_t* plan, int64_t* host_out,
                                  int64_t capacity_ctas);
int loopsb_plan_probe_collect(loopsb_plan_t* plan, float* host_ms,
                              int32_t capacity, int32_t* n);

/* ---------------------------------------------------------------------------
 * SpMV  y = A * x, fp32 values, int32 ids. Replaces the host entry points
 *   algorithms::spmv::merge_path_flat   merge_path_flat.cuh:96-139
 *   algorithms::spmv::work_oriented     work_oriented.cuh:102-120
 *   algorithms::spmv::thread_mapped     thread_mapped.cuh:69-91
 *   algorithms::spmv::group_mapped      group_mapped.cuh:72-104
 *   algorithms::spmv::coo_thread_mapped coo_thread_mapped.cuh:61-89
 *   algorithms::spmv::ell_thread_mapped ell_thread_mapped.cuh:52-76
 *   algorithms::spmv::ell_merge_path    ell_merge_path.cuh:76-126
 *   algorithms::spmv::csc_thread_mapped csc_thread_mapped.cuh:54-84   (CSC + thread_mapped:
 *       tiles are columns, `col_indices` carries the ROW index of every entry)
 *   algorithms::spmv::flat_partitioned  flat_partitioned.cuh:73-107   (FLAT + thread_mapped:
 *       descriptor = {kind FLAT, num_tiles = ceil(nnz/K), num_atoms = nnz,
 *       pitch = K, offsets = the base CSR row offsets})
 *   algorithms::spmv::original          original.cuh:55-72            (= CSR + thread_mapped)
 * selected by (plan schedule, plan layout kind). `row_indices` is read only
 * for COO. For ELL, `col_indices`/`values` are the row-major rows*pitch slabs
 * with column -1 in padding slots (container/ell.hxx:31-36).
 * ------------------------------------------------------------------------- */
int loopsb_spmv_f32(loopsb_plan_t* plan, const float* values,
                    const int32_t* col_indices, const int32_t* row_indices,
                    const float* x, float* y, int32_t num_rows,
                    int32_t num_cols, void* stream);

/* y += A x for a merge_path_flat plan over CSR (same kernel, read-modify-write of the
 * rows it closes; deterministic). Used for the second and later column blocks of a
 * multi-GPU shard (loopsb_dist_*). LOOPSB_ERR_UNSUPPORTED for other plans or for
 * indices / values that are not 16-byte aligned. */
int loopsb_spmv_acc_f32(loopsb_plan_t* plan, const float* values, const int32_t* col_indices,
                        const float* x, float* y, int32_t num_rows, int32_t num_cols, void* stream);

/* BCSR R x C dense blocks (values[b*R*C + i*C + j], container/bcsr.hxx:14-22),
 * one thread per block-row, fp32 FMA. Replaces
 * algorithms::spmv::bcsr_thread_mapped (bcsr_thread_mapped.cuh:88-123).
 * R == C in {2,3,4}. x must be padded to num_block_cols*C. */
int loopsb_spmv_bcsr_f32(int32_t R, int32_t C, const loopsb_layout_t* lay,
                         const float* values, const int32_t* block_col_indices,
                         const float* x_padded, float* y, int32_t num_rows,
                         void* stream);

/* DIA: diagonals in ascending offset order, values column-major
 * values[d * stride + r] (container/dia.hxx:56-62), one thread per row.
 * Replaces algorithms::spmv::dia_thread_mapped (dia_thread_mapped.cuh:65-98). */
int loopsb_spmv_dia_f32(int32_t num_rows, int32_t num_cols, int64_t stride,
                        int32_t num_diagonals, const int32_t* diag_offsets,
                        const float* values, const float* x, float* y,
                        void* stream);

/* SpMM  C = A * B, A in CSR (fp32), B [num_cols x n] and C [num_rows x n] dense
 * row-major. Replaces algorithms::spmm::thread_mapped (spmm/thread_mapped.cuh:55-80)
 * over container/matrix.cuh. C is fully overwritten. */
int loopsb_spmm_csr_f32(const loopsb_layout_t* lay, const float* values,
                        const int32_t* col_indices, const float* B, float* C,
                        int32_t num_rows, int32_t num_cols, int32_t n, void* stream);

/* BCSR 4x4, bf16 values and x, fp32 accumulate and y, on the tcgen05 tensor
 * cores (BASELINE.json config 4). The plan must have been created from a
 * LOOPSB_LAYOUT_BCSR descriptor with LOOPSB_SCHED_THREAD_MAPPED. */
int loopsb_spmv_bcsr4x4_bf16(loopsb_plan_t* plan, const uint16_t* values_bf16,
                             const int32_t* block_col_indices,
                             const uint16_t* x_bf16_padded, float* y,
                             int32_t num_rows, void* stream);
/* Optional, once per matrix (like loopsb_plan_tile_csr): the plan keeps a copy
 * of the block values (and block columns) re-ordered into the tensor-core
 * operand tiles the kernel consumes, so that a K-step's operands arrive as one
 * TMA bulk copy. Costs 36 bytes per (padded) block of device memory and one
 * pass over the matrix;
 * loopsb_spmv_bcsr4x4_bf16 uses the copy whenever it is called with the same
 * `values` pointer, and the plain path otherwise. Call again after changing
 * the values or columns in place. Results are bit-identical either way. */
int loopsb_plan_pack_bcsr4x4(loopsb_plan_t* plan, const uint16_t* values_bf16,
                             const int32_t* block_col_indices, void* stream);

/* ---------------------------------------------------------------------------
 * Band-tiled plan (optional accelerator for merge_path_flat / CSR).
 * Still the reference's merge_path::preprocess_t idea -- per-matrix set-up that
 * the timed SpMV re-uses (schedule/merge_path_flat.hxx:92-172; the reference's
 * timer excludes it, algorithms/spmv/merge_path_flat.cuh:111,121-122) -- taken
 * one step further: the plan keeps a re-ordered COPY of the matrix cut into
 * (row block x column band) tiles (8 bytes per nonzero, as CSR) so that the
 * SpMV kernel can hold its y rows and the current x band in shared memory
 * (loops_b200/csrc/spmv_tiled.cuh). The call signature of the SpMV does not
 * change: loopsb_spmv_f32 on this plan takes the tiled kernel whenever it is
 * handed the SAME col_indices / values pointers the copy was made from (and a
 * 16-byte aligned x); any other pointers run the plain CSR kernel. If the
 * caller changes values in place it must call loopsb_plan_tile_csr again.
 * A plan serves ONE SpMV at a time: the tiled kernel keeps per-plan partial-row
 * and arrival-counter workspaces, so calls on the same plan must be ordered on
 * one stream (concurrent streams need one plan each).
 *
 * flags bit 0 (LOOPSB_TILE_FORCE): build even when the cost model says the
 * plain kernel is the better choice (x too large for the band walk to pay).
 * Returns LOOPSB_ERR_UNSUPPORTED (plan unchanged, still usable) when the
 * matrix does not fit the format or is not worth tiling. Synchronises.
 * The copy is built ON THE DEVICE (loops_b200/csrc/tiled_build.cuh; only the
 * row offsets cross PCIe); LOOPSB_TILED_HOST_BUILD=1 selects the host builder,
 * which produces the same image.
 * ------------------------------------------------------------------------- */
#define LOOPSB_TILE_FORCE 1

typedef struct loopsb_tiled_info {
  int32_t nb, q, warps, cb, xb, es;   /* row blocks, column parts, consumer warps,
                                         band width, x-ring depth, stream-ring depth */
  int32_t rb, rw, cq, nband;          /* max rows/block, max rows/warp, columns/part, bands/part */
  int32_t grid_blocks, cta_threads, smem_bytes;
  int32_t long_steps;                 /* steps with a run over >= 3 lanes (segmented-scan fast path) */
  int64_t total_steps;                /* 1 KB steps (128 entries) in the copy   */
  int64_t real_entries, pad_entries;  /* nnz and padding entries                */
  int64_t flagged_entries, flagged_steps; /* rows split into separate cell ranges (general path) */
  int64_t bytes;                      /* device memory held by the copy         */
} loopsb_tiled_info_t;

int loopsb_plan_tile_csr(loopsb_plan_t* plan, const int32_t* col_indices,
                         const float* values, int32_t num_cols, int32_t flags,
                         void* stream);
int loopsb_plan_untile(loopsb_plan_t* plan);
/* LOOPSB_ERR_UNSUPPORTED when the plan holds no tiled copy. */
int loopsb_plan_tiled_info(const loopsb_plan_t* plan, loopsb_tiled_info_t* info);

/* Copy the plan's tiled image back to HOST arrays (format tests: the device
 * builder must produce the host builder's image byte for byte). steps:
 * (total_steps + es) * 256 words; stream_base: nb*q*warps + 1; block_begin: nb + 1.
 * Any pointer may be NULL. Synchronises. */
int loopsb_plan_tiled_download(const loopsb_plan_t* plan, uint32_t* host_steps,
                               int64_t capacity_words, int32_t* host_stream_base,
                               int32_t* host_block_begin);

/* The same builder on HOST arrays, no device involved: returns the image the
 * plan would upload, for format tests. geometry = {nb,q,warps,cb,xb,es}. */
typedef struct loopsb_tiled_image loopsb_tiled_image_t;
int loopsb_tiled_image_build_host(int32_t num_rows, int32_t num_cols,
                                  const int32_t* host_offsets,
                                  const int32_t* host_indices,
                                  const float* host_values,
                                  const int32_t geometry[6],
                                  loopsb_tiled_image_t** out);
int loopsb_tiled_image_info(const loopsb_tiled_image_t* img, loopsb_tiled_info_t* info);
/* steps[total_steps*256], stream_base[nb*q*warps+1], first_step / last_step_end
 * [nb*q*warps][nband], block_begin[nb+1] (row-block cuts, by nonzero count),
 * warp_begin[nb*q][warps+1] (block-local row cuts between consumer warps). */
int loopsb_tiled_image_arrays(const loopsb_tiled_image_t* img, const uint32_t** steps,
                              const int32_t** stream_base, const uint16_t** first_step,
                              const uint16_t** last_step_end, const int32_t** block_begin,
                              const int32_t** warp_begin);
int loopsb_tiled_image_free(loopsb_tiled_image_t* img);

/* Host-buffer convenience with the flow of the reference's example mains
 * (examples/spmv/merge_path.cu:17-50): upload CSR + x, run, download y,
 * synchronise. *kernel_ms (optional) receives the CUDA-event time of the SpMV
 * launches only (what the reference's util::timer_t reports). */
int loopsb_spmv_csr_host_f32(int schedule, int32_t num_rows, int32_t num_cols,
                             int32_t nnz, const int32_t* host_offsets,
                             const int32_t* host_indices,
                             const float* host_values, const float* host_x,
                             float* host_y, float* kernel_ms);

/* ---------------------------------------------------------------------------
 * Schedule index-stream emission (parity instrument; not on the SpMV path).
 * Runs the loops-b200 schedule::setup<> iterators (include/loops/schedule/)
 * with a recording body and writes, for every atom a:
 *   visitor[a] = global thread id that was handed the atom
 *   step[a]    = ordinal of that hand-out within the thread's own sequence
 *   tile[a]    = tile id handed out with it
 *   visits[a] += 1
 * (caller pre-fills visitor/step/tile with -1 and visits with 0).
 *   work_oriented : extra_a = commit[A] (0 first complete tile -> atomic-if-
 *                   nonzero, 1 later complete tile -> store, 2 remainder),
 *                   extra_b = map[grid*TPB*4] = {st.x, st.y, en.x, en.y}
 *   merge_path_flat: dense_{tile,atom,emit}[M*TPB*IPT] per (block,thread,item)
 *                   exactly as the reference kernel sees them
 *                   (algorithms/spmv/merge_path_flat.cuh:71-82),
 *                   extra_b = thread_start[M*TPB*2]
 * grid_blocks is only read for thread_mapped and work_oriented (the other two
 * derive their grid from the layout like the reference wrappers do).
 * (TPB, IPT) for merge_path_flat: (128,8) = reference sm_100 f32 tuning
 * (algorithms/spmv/launch_box.hxx:66-68), (128,7) fallback, (128,5) ELL.
 * Synchronises before returning.
 * ------------------------------------------------------------------------- */
int loopsb_emit_schedule(const loopsb_layout_t* lay, int schedule,
                         int32_t grid_blocks, int32_t threads_per_block,
                         int32_t items_per_thread, int32_t* visitor,
                         int32_t* step, int32_t* tile, int32_t* visits,
                         int32_t* extra_a, int32_t* extra_b,
                         int32_t* dense_tile, int32_t* dense_atom,
                         int32_t* dense_emit, int64_t dense_len, void* stream);

/* ---------------------------------------------------------------------------
 * Format conversions on the device. The reference
 * converts on the host and copies the result up (container/ell.hxx:113-145,
 * bcsr.hxx:111-194, dia.hxx:135-188) or sorts zipped COO triples with thrust
 * (coo.hxx:87-98, csr.hxx:86-94, csc.hxx:86-108, detail/convert.hxx:36-78).
 * All pointers are DEVICE pointers; outputs are caller-allocated and bit-equal
 * to the reference converters' for CSR input with unique columns per row. The
 * calls use temporary device memory and synchronise the stream before
 * returning (except loopsb_csr_to_ell, which is a single asynchronous launch).
 * ------------------------------------------------------------------------- */
/* row_indices[nnz]: the row of every atom (detail/convert.hxx:36-60). */
int loopsb_csr_to_coo(int32_t num_rows, int64_t nnz, const int32_t* offsets,
                      int32_t* row_indices, void* stream);
/* Sort by (row, col) and compress the rows (csr.hxx:86-94); input order free. */
int loopsb_coo_to_csr(int32_t num_rows, int64_t nnz, const int32_t* row_indices,
                      const int32_t* col_indices, const float* values,
                      int32_t* offsets /*[num_rows+1]*/, int32_t* indices /*[nnz]*/,
                      float* out_values /*[nnz]*/, void* stream);
/* Structural transpose, entries ordered by (column, row) (csc.hxx:86-108). */
int loopsb_csr_to_csc(int32_t num_rows, int32_t num_cols, int64_t nnz,
                      const int32_t* offsets, const int32_t* indices, const float* values,
                      int32_t* csc_offsets /*[num_cols+1]*/, int32_t* csc_row_indices /*[nnz]*/,
                      float* csc_values /*[nnz]*/, void* stream);
/* ELL pitch = widest row (ell.hxx:121-126); *max_degree is a HOST int. */
int loopsb_csr_max_degree(int32_t num_rows, const int32_t* offsets, int32_t* max_degree,
                          void* stream);
/* Row-major rows*pitch slabs, padding column -1 / value 0 (ell.hxx:128-140). */
int loopsb_csr_to_ell(int32_t num_rows, int32_t pitch, const int32_t* offsets,
                      const int32_t* indices, const float* values,
                      int32_t* ell_indices /*[rows*pitch]*/, float* ell_values /*[rows*pitch]*/,
                      void* stream);
/* BCSR R x C (bcsr.hxx:111-194) in two calls. count: block_offsets[ceil(rows/R)+1],
 * atom_block[nnz] = the block every atom lands in, *num_blocks (HOST). fill:
 * block_col_indices[num_blocks] ascending per block-row, block_values
 * [num_blocks*R*C] (fp32, or bf16 when bf16_values != 0), zero padding. */
int loopsb_csr_to_bcsr_count(int32_t R, int32_t C, int32_t num_rows, int32_t num_cols, int64_t nnz,
                             const int32_t* offsets, const int32_t* indices,
                             int32_t* block_offsets, int32_t* atom_block, int64_t* num_blocks,
                             void* stream);
int loopsb_csr_to_bcsr_fill(int32_t R, int32_t C, int32_t num_rows, int32_t num_cols, int64_t nnz,
                            const int32_t* offsets, const int32_t* indices, const float* values,
                            const int32_t* atom_block, int64_t num_blocks,
                            int32_t* block_col_indices, void* block_values, int32_t bf16_values,
                            void* stream);
/* DIA (dia.hxx:135-188) in two calls. count: *num_diagonals (HOST) = distinct
 * (col - row). fill: diag_offsets[num_diagonals] ascending, dia_values
 * [num_diagonals*rows] with values[d*rows + r], zero padding. */
int loopsb_csr_to_dia_count(int32_t num_rows, int32_t num_cols, int64_t nnz,
                            const int32_t* offsets, const int32_t* indices,
                            int32_t* num_diagonals, void* stream);
int loopsb_csr_to_dia_fill(int32_t num_rows, int32_t num_cols, int64_t nnz,
                           const int32_t* offsets, const int32_t* indices, const float* values,
                           int32_t num_diagonals, int32_t* diag_offsets, float* dia_values,
                           void* stream);

/* Column-block split of a CSR matrix (the multi-GPU shards of SURVEY.md section 8e): the
 * columns are cut into `num_chunks` equal ranges of `chunk_cols` (one per source rank of
 * the x all-gather) and host_block_of_chunk[c] in [0, num_blocks) names the output block
 * the entries of chunk c go to. count: block_offsets[num_blocks * (num_rows + 1)] = every
 * block's own CSR offsets (each starting at 0), host_block_nnz[num_blocks] (HOST). fill:
 * out_indices / out_values [nnz + 4 * num_blocks] = the blocks one after the other, each
 * starting on a 16-byte boundary (start_0 = 0, start_b = round_up(start_{b-1} + nnz_{b-1}, 4)),
 * CSR order kept inside every block, column ids stay GLOBAL. */
int loopsb_csr_split_columns_count(int32_t num_rows, int64_t nnz, const int32_t* offsets,
                                   const int32_t* indices, int32_t chunk_cols, int32_t num_chunks,
                                   const int32_t* host_block_of_chunk, int32_t num_blocks,
                                   int32_t* block_offsets, int64_t* host_block_nnz, void* stream);
int loopsb_csr_split_columns_fill(int32_t num_rows, int64_t nnz, const int32_t* offsets,
                                  const int32_t* indices, const float* values, int32_t chunk_cols,
                                  int32_t num_chunks, const int32_t* host_block_of_chunk,
                                  int32_t num_blocks, const int32_t* block_offsets,
                                  const int64_t* host_block_nnz, int32_t* out_indices,
                                  float* out_values, void* stream);

/* ---------------------------------------------------------------------------
 * Row-partitioned multi-GPU SpMV (BASELINE configs[4]).
 * The reference is single-GPU (memory.hxx:24 mentions multi-GPU only as a TODO); this
 * is the north star's design: rank r owns a contiguous row range with GLOBAL column
 * ids and an equal shard of x (num_cols % world == 0); the only thing that crosses
 * GPUs is the dense x, all-gathered with NCCL over NVLink. One process (or host
 * thread) per GPU, the current CUDA device is the rank's GPU. NCCL is loaded with
 * dlopen("libnccl.so.2") (LOOPSB_NCCL_LIB overrides), so libloopsb200.so does not link it.
 *
 *   loopsb_dist_unique_id : rank 0 makes the 128-byte NCCL id; the host launcher hands
 *                           it to every rank (MPI / torch.distributed / a file).
 *   loopsb_dist_create    : collective. `groups` = NULL/0: every step is ONE
 *                           ncclAllGather on the caller's stream followed by one
 *                           merge-path SpMV over the caller's arrays (borrowed).
 *                           groups = {g1, g2, ...} (sum = world - 1): the all-gather is
 *                           issued as ring-shifted NCCL send/recv phases on a side
 *                           stream -- phase i brings the next g_i chunks in ring order
 *                           r+1, r+2, ... -- and the plan keeps the shard as column
 *                           blocks (own columns, then one block per phase; 8 bytes per
 *                           nonzero of device memory), so block i's SpMV (y += A_i x)
 *                           runs while phase i+1 is in flight and gathers from a slice
 *                           of x that stays L2-resident. Same y on exact inputs. The
 *                           phases are copy-engine pulls from the peers' CUDA-IPC
 *                           mapped staging buffers, ordered by stream memory operations
 *                           (no SM, no host barrier); LOOPSB_DIST_TRANSPORT=nccl keeps
 *                           them as NCCL send/recv groups (also the automatic fallback).
 *   loopsb_dist_spmv      : y_shard = A_shard * allgather(x_shard); asynchronous on
 *                           `stream`; x_shard must stay untouched until the call's work
 *                           on `stream` has completed. Collective: every rank calls it.
 * ------------------------------------------------------------------------- */
typedef struct loopsb_dist loopsb_dist_t;
#define LOOPSB_DIST_ID_BYTES 128
typedef struct loopsb_dist_info {
  int32_t world, rank, local_rows, num_cols, num_blocks, nccl_version;
  int32_t transport;          /* 0 = one ncclAllGather, 1 = NCCL send/recv phases, 2 = copy-engine pulls over CUDA IPC */
  int32_t graphs_cached;      /* CUDA graphs of whole steps held for replay (0 = steps are enqueued call by call) */
  int64_t local_nnz;
  int64_t block_nnz[8];
  int64_t bytes;              /* device memory held by the object (gathered x, column blocks) */
} loopsb_dist_info_t;
int loopsb_dist_unique_id(void* id128);
int loopsb_dist_create(loopsb_dist_t** out, const void* id128, int32_t world, int32_t rank,
                       int32_t local_rows, int32_t num_cols, int64_t local_nnz,
                       const int32_t* offsets, const int32_t* col_indices, const float* values,
                       const int32_t* groups, int32_t num_groups, void* stream);
int loopsb_dist_spmv(loopsb_dist_t* dist, const float* x_shard, float* y_shard, void* stream);
int loopsb_dist_info(const loopsb_dist_t* dist, loopsb_dist_info_t* info);
/* The gathered x of the last step (device pointer owned by the object). */
int loopsb_dist_x_full(const loopsb_dist_t* dist, const float** x_full);
/* Breakdown probes: with probing on, every step records CUDA events around the
 * all-gather (first send/recv issued -> last chunk landed) and around every block's
 * SpMV; loopsb_dist_probe_read synchronises the device and returns the LAST step's
 * comm_ms, the sum of its block kernels and (optionally) each block's ms. */
int loopsb_dist_probe(loopsb_dist_t* dist, int32_t enable);
int loopsb_dist_probe_read(loopsb_dist_t* dist, float* comm_ms, float* kernel_ms, float* block_ms,
                           int32_t capacity);
int loopsb_dist_destroy(loopsb_dist_t* dist);

/* fp64 CSR SpMV (the reference builds its examples
 * for double too, examples/spmv/CMakeLists.txt:29). thread_mapped is bit-equal to
 * reference::spmv<double>; the other three schedules share one merge-path kernel
 * (512 items per tile, algorithms/spmv/launch_box.hxx:68) whose rows cut by a tile
 * boundary are accumulated with fp64 atomics. y is fully overwritten. No plan. */
int loopsb_spmv_f64(const loopsb_layout_t* lay, int schedule, const double* values,
                    const int32_t* col_indices, const double* x, double* y,
                    int32_t num_rows, int32_t num_cols, void* stream);

/* The other in-tree layouts in double -- COO / CSC / flat_uniform_occupancy (thread_mapped),
 * ELL (thread_mapped, merge_path_flat), CSR (forwards to loopsb_spmv_f64); same argument meaning
 * as loopsb_spmv_f32 -- and DIA / BCSR like their fp32 entry points. The reference's entry points
 * are templates on type_t and every example is also built as .f64
 * (examples/spmv/CMakeLists.txt:29). Same thread->work maps as the fp32 kernels, un-fused
 * arithmetic, y fully overwritten; not tuned (fp64 is not a BASELINE config). */
int loopsb_spmv_layout_f64(const loopsb_layout_t* lay, int schedule, const double* values,
                           const int32_t* col_indices, const int32_t* row_indices, const double* x,
                           double* y, int32_t num_rows, int32_t num_cols, void* stream);
int loopsb_spmv_dia_f64(int32_t num_rows, int32_t num_cols, int64_t stride, int32_t num_diagonals,
                        const int32_t* diag_offsets, const double* values, const double* x, double* y,
                        void* stream);
int loopsb_spmv_bcsr_f64(int32_t R, int32_t C, const loopsb_layout_t* lay, const double* values,
                         const int32_t* block_col_indices, const double* x_padded, double* y,
                         int32_t num_rows, void* stream);

/* Which schedule to run for a CSR matrix (the
 * reference publishes the outcome of its heuristic per matrix in
 * plots/data/heuristics.csv). max_degree < 0 = unknown (then only nnz decides,
 * like the reference's rule); loopsb_csr_max_degree computes it on the device.
 * Pure host function: no device is touched. */
int loopsb_select_schedule(int32_t num_rows, int32_t num_cols, int64_t nnz,
                           int32_t max_degree, int32_t* schedule);

/* Grid the reference-compatible work_oriented launch uses on this device:
 * resident blocks per SM (occupancy API) x SM count, 128 threads per block
 * (algorithms/spmv/work_oriented.cuh:112-113). */
int loopsb_work_oriented_grid(int32_t* grid_blocks);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* LOOPSB_H_ */
