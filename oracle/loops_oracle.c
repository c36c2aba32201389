/* oracle/loops_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked, imported or
 * executed by the product path; see DESIGN.md "Oracle").
 *
 * Plain-C CPU restatement of the gunrock/loops SpMV hot path: the layout
 * contract, the four schedules' iterators (as ordered index streams), the CPU
 * validator, the x recipe and the host-side format conversions. Every function
 * cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Parity status: PINNED.
 *   - x recipe, reference::spmv / spmv_f64 / row_l1_products, CSR->ELL/COO/BCSR
 *     are checked against the unmodified reference headers compiled into
 *     oracle/_ref/libloopsref_host.so (tests/test_oracle_vs_ref.py, runs in the
 *     build container) and against the chesapeake known answers.
 *   - schedule index streams are checked against the reference's own
 *     schedule::setup<> templates run on the B200 (oracle/ref_gpu.cu ->
 *     oracle/_ref/libloopsref_gpu.so; tests/test_gpu_ref_streams.py) and
 *     against the golden streams captured from that run (tests/golden/).
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC -> oracle/libloops_oracle.so)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t i32;
typedef int64_t i64;

/* ------------------------------------------------------------------------ */
/* Layout contract: include/loops/container/layout.hxx:16-55                 */
/* kinds: 0 = offsets-array (csr :87-149, bcsr :239-285, csc :312-359)       */
/*        1 = coo (:385-421)   2 = uniform pitch (ell :443-496, dia :166-217) */
/*        3 = flat_uniform_occupancy<K> (container/partitioning.hxx:71-141)  */
/* ------------------------------------------------------------------------ */
typedef struct {
  i32 kind;
  const i32* offsets; /* kind 0: length num_tiles+1 */
  i32 num_tiles;
  i32 num_atoms;
  i32 pitch; /* kind 2: atoms per tile; kind 3: K */
} orc_layout;

enum { ORC_OFFSETS = 0, ORC_COO = 1, ORC_PITCH = 2, ORC_FLAT = 3 };

i32 orc_num_tiles(const orc_layout* l) {
  switch (l->kind) {
    case ORC_COO: return l->num_atoms;                      /* layout.hxx:396 */
    case ORC_FLAT: return (l->num_atoms + l->pitch - 1) / l->pitch; /* partitioning.hxx:102-105 */
    default: return l->num_tiles;
  }
}
i32 orc_num_atoms(const orc_layout* l) {
  if (l->kind == ORC_PITCH) return l->num_tiles * l->pitch; /* layout.hxx:468-470 */
  return l->num_atoms;
}
i32 orc_tile_begin(const orc_layout* l, i32 t) {
  switch (l->kind) {
    case ORC_OFFSETS: return l->offsets[t];                 /* layout.hxx:115 */
    case ORC_COO: return t;                                 /* :401 */
    case ORC_PITCH: return t * l->pitch;                    /* :472 */
    default: return t * l->pitch;                           /* partitioning.hxx:109 */
  }
}
i32 orc_tile_end(const orc_layout* l, i32 t) {
  switch (l->kind) {
    case ORC_OFFSETS: return l->offsets[t + 1];             /* layout.hxx:118 */
    case ORC_COO: return t + 1;                             /* :404 */
    case ORC_PITCH: return (t + 1) * l->pitch;              /* :475 */
    default: {                                              /* partitioning.hxx:113-117 */
      i32 e = (t + 1) * l->pitch;
      return e < l->num_atoms ? e : l->num_atoms;
    }
  }
}
i32 orc_tile_size(const orc_layout* l, i32 t) {
  return orc_tile_end(l, t) - orc_tile_begin(l, t);
}
/* tile_of: layout.hxx:137-148 (upper-bound flavour), :418, :493 */
i32 orc_tile_of(const orc_layout* l, i32 a) {
  if (l->kind == ORC_COO) return a;
  if (l->kind == ORC_PITCH || l->kind == ORC_FLAT) return a / l->pitch;
  i32 lo = 0, hi = l->num_tiles;
  while (lo < hi) {
    i32 mid = lo + ((hi - lo) >> 1);
    if (l->offsets[mid + 1] <= a) lo = mid + 1; else hi = mid;
  }
  return lo;
}
/* tile_end_iter()[k] == tile_end(k): layout.hxx:126-128, :411-413, :483-486 */
static inline i64 tile_end_at(const orc_layout* l, i64 k) {
  return (i64)orc_tile_end(l, (i32)k);
}

/* ------------------------------------------------------------------------ */
/* Diagonal search: util/search.hxx:34-60 (and the private copy in           */
/* schedule/work_oriented.hxx:156-179). thrust::lower_bound(seq) over the    */
/* counting range [x_min, x_max) with pred  a[i] <= b0 + d - i - 1 ; an       */
/* inverted range leaves the iterator at x_min. Returns (min(i,a_len), d-i). */
/* `a` is either the layout's tile_end_iter (loc == NULL) or a staged array. */
/* ------------------------------------------------------------------------ */
static void diag_search(i64 d, const orc_layout* lay, const i32* loc, i64 b0,
                        i64 a_len, i64 b_len, i64* ox, i64* oy) {
  i64 x_min = d - b_len; if (x_min < 0) x_min = 0;
  i64 x_max = d < a_len ? d : a_len;
  i64 first = x_min, len = x_max - x_min;
  while (len > 0) {
    i64 half = len >> 1, mid = first + half;
    i64 av = loc ? (i64)loc[mid] : tile_end_at(lay, mid);
    if (av <= b0 + d - mid - 1) { first = mid + 1; len = len - half - 1; }
    else len = half;
  }
  *ox = first < a_len ? first : a_len;
  *oy = d - first;
}

void orc_diag_search(const orc_layout* lay, i64 d, i64* ox, i64* oy) {
  diag_search(d, lay, NULL, 0, orc_num_tiles(lay), orc_num_atoms(lay), ox, oy);
}

/* ------------------------------------------------------------------------ */
/* Per-atom stream record shared by all four schedules:                      */
/*   visitor[a] = global thread id that touches atom a                        */
/*   step[a]    = ordinal of that touch in the thread's own sequence          */
/*   tile[a]    = tile id the schedule hands to the body for that atom        */
/* Arrays are pre-filled with -1 by the caller; `visits[a]` counts touches    */
/* (the "every atom exactly once" property of                                */
/* unittests/test_schedule_coverage.cu:58-111).                               */
/* ------------------------------------------------------------------------ */
static inline void rec(i32* visitor, i32* step, i32* tile, i32* visits, i64 a,
                       i64 g, i64 s, i64 t) {
  visitor[a] = (i32)g; step[a] = (i32)s; tile[a] = (i32)t; visits[a] += 1;
}

/* thread_mapped: stride_ranges.hxx:28-32 + schedule/thread_mapped.hxx:96-109.
 * Thread g walks tiles g, g+G, ... < T (G = gridDim.x*blockDim.x); per tile,
 * atoms tile_begin..tile_end-1 ascending. */
void orc_emit_thread_mapped(const orc_layout* lay, i64 G, i32* visitor,
                            i32* step, i32* tile, i32* visits) {
  i64 T = orc_num_tiles(lay);
  for (i64 g = 0; g < G; ++g) {
    i64 s = 0;
    for (i64 t = g; t < T; t += G)
      for (i64 a = orc_tile_begin(lay, (i32)t); a < orc_tile_end(lay, (i32)t); ++a)
        rec(visitor, step, tile, visits, a, g, s++, t);
  }
}

/* group_mapped as SpMV uses it (block_mapped<TPB>, group == whole block):
 * schedule/group_mapped.hxx:104-192, algorithms/spmv/group_mapped.cuh:40-60.
 * off[] = exclusive scan of tile_size over the block's ranks (0 for ranks past
 * T), agg = off[TPB-1]+size[TPB-1]; rank r walks v = r, r+TPB, ... < agg;
 * vt = upper_bound(off[0..len), v) - 1; skipped when vt >= len. */
void orc_emit_group_mapped(const orc_layout* lay, i32 TPB, i32* visitor,
                           i32* step, i32* tile, i32* visits) {
  i64 T = orc_num_tiles(lay);
  i64 blocks = (T + TPB - 1) / TPB;
  i32* off = (i32*)malloc(sizeof(i32) * (size_t)TPB);
  for (i64 B = 0; B < blocks; ++B) {
    i64 base = B * TPB;
    i64 len = (T < base + TPB ? T : base + TPB) - base; /* get_length :145-158 */
    i32 run = 0, last_size = 0;
    for (i32 r = 0; r < TPB; ++r) {
      i32 sz = (base + r < T) ? orc_tile_size(lay, (i32)(base + r)) : 0;
      off[r] = run; run += sz; last_size = sz;
    }
    i32 agg = off[TPB - 1] + last_size;
    for (i32 r = 0; r < TPB; ++r) {
      for (i32 v = r; v < agg; v += TPB) {
        /* thrust::upper_bound(seq, off, off+len, v) */
        i64 lo = 0, n = len;
        while (n > 0) {
          i64 half = n >> 1;
          if (!(v < off[lo + half])) { lo += half + 1; n -= half + 1; } else n = half;
        }
        i64 vt = lo - 1;
        if (!(vt < len)) continue;
        i64 t = base + vt;                                   /* tile_id :172-176 */
        i64 a = (i64)orc_tile_begin(lay, (i32)t) + v - off[vt]; /* atom_id :184-192 */
        rec(visitor, step, tile, visits, a, base + r, (v - r) / TPB, t);
      }
    }
  }
  free(off);
}

/* work_oriented: schedule/work_oriented.hxx:78-142 with the SpMV body of
 * algorithms/spmv/work_oriented.cuh:52-88. N = gridDim.x*TPB threads,
 * w = ceil((T+A)/N). map[g] = {st.x, st.y, en.x, en.y}. commit[a]: 0 = atom
 * belongs to the thread's first complete tile (atomicAdd-if-nonzero), 1 = a
 * later complete tile (plain store), 2 = remainder tile (atomicAdd-if-nonzero). */
void orc_emit_work_oriented(const orc_layout* lay, i64 N, i32* visitor,
                            i32* step, i32* tile, i32* visits, i32* commit,
                            i32* map /* [N*4] or NULL */) {
  i64 T = orc_num_tiles(lay), A = orc_num_atoms(lay);
  i64 W = T + A;
  i64 w = N > 0 ? (W / N + (W % N != 0 ? 1 : 0)) : 0;      /* math.hxx ceil_div */
  for (i64 g = 0; g < N; ++g) {
    i64 d0 = w * g < W ? w * g : W;                          /* :97-100 */
    i64 d1 = d0 + w < W ? d0 + w : W;
    i64 sx, sy, ex, ey;
    diag_search(d0, lay, NULL, 0, T, A, &sx, &sy);
    diag_search(d1, lay, NULL, 0, T, A, &ex, &ey);
    if (map) { map[4*g] = (i32)sx; map[4*g+1] = (i32)sy; map[4*g+2] = (i32)ex; map[4*g+3] = (i32)ey; }
    i64 cur = sy, s = 0; int first = 1;
    for (i64 t = sx; t < ex; ++t) {                          /* tiles(m) :117-120 */
      i64 e = orc_tile_end(lay, (i32)t);                     /* atoms(t,m) :122-128 */
      for (i64 a = cur; a < e; ++a) {
        rec(visitor, step, tile, visits, a, g, s++, t);
        if (commit) commit[a] = first ? 0 : 1;
      }
      cur += (e - cur);
      first = 0;
    }
    for (i64 a = cur; a < ey; ++a) {                         /* remainder :131-142 */
      rec(visitor, step, tile, visits, a, g, s++, ex);
      if (commit) commit[a] = 2;
    }
  }
}

/* merge_path_flat: schedule/merge_path_flat.hxx:267-371 with the SpMV body of
 * algorithms/spmv/merge_path_flat.cuh:63-82 and grid shape :114-127.
 * I = TPB*IPT, M = ceil((T+A)/I). For block b: s = S(b*I), e = S((b+1)*I);
 * E[j] = end[min(s.x+j, T-1)], j < nt+IPT; thread k: c = S(k*IPT over E vs
 * counting(s.y)); then IPT steps.
 * Outputs (dense, index (b*TPB + k)*IPT + item):
 *   d_tile = tile_idx(map), d_atom = atom_idx(map) (clamped to A-1),
 *   d_emit = 1 when the step commits the atom, 0 when it advances the row.
 * coords[2*b], coords[2*b+1] = S(b*I) for b = 0..M (what
 * generate_search_coordinates :45-76 materialises). thread_start[(b*TPB+k)*2]
 * = the per-thread coordinate returned by init().
 * Returns M. */
i64 orc_emit_merge_path(const orc_layout* lay, i32 TPB, i32 IPT, i32* visitor,
                        i32* step, i32* tile, i32* visits, i32* d_tile,
                        i32* d_atom, i32* d_emit, i32* coords,
                        i32* thread_start) {
  i64 T = orc_num_tiles(lay), A = orc_num_atoms(lay);
  i64 I = (i64)TPB * IPT, W = T + A;
  i64 M = W / I + (W % I != 0 ? 1 : 0);
  i32* E = (i32*)malloc(sizeof(i32) * (size_t)(I + IPT + 1));
  if (coords)
    for (i64 b = 0; b <= M; ++b) {
      i64 x, y; diag_search(b * I, lay, NULL, 0, T, A, &x, &y);
      coords[2*b] = (i32)x; coords[2*b+1] = (i32)y;
    }
  for (i64 b = 0; b < M; ++b) {
    i64 sx, sy, ex, ey;
    diag_search(b * I, lay, NULL, 0, T, A, &sx, &sy);
    diag_search((b + 1) * I, lay, NULL, 0, T, A, &ex, &ey);
    i64 nt = ex - sx, na = ey - sy;
    for (i64 j = 0; j < nt + IPT; ++j) {                     /* :309-316 */
      i64 o = sx + j < T - 1 ? sx + j : T - 1;
      E[j] = o >= 0 ? (i32)tile_end_at(lay, o) : 0;
    }
    for (i32 k = 0; k < TPB; ++k) {
      i64 cx, cy;
      diag_search((i64)k * IPT, lay, E, sy, nt, na, &cx, &cy); /* :328-330 */
      i64 base = (b * TPB + k) * IPT;
      if (thread_start) { thread_start[(b*TPB+k)*2] = (i32)cx; thread_start[(b*TPB+k)*2+1] = (i32)cy; }
      i64 s = 0;
      for (i32 item = 0; item < IPT; ++item) {
        i64 nz = sy + cy < A - 1 ? sy + cy : A - 1;          /* atom_idx :358-361 */
        i64 row = sx + cx;                                   /* tile_idx :369-371 */
        int commit = (sy + cy) < (i64)E[cx];                 /* merge_path_flat.cuh:75-76 */
        if (d_tile) { d_tile[base+item] = (i32)row; d_atom[base+item] = (i32)nz; d_emit[base+item] = commit; }
        if (commit) {
          if (visitor) rec(visitor, step, tile, visits, nz, b * TPB + k, s, row);
          cy++;
        } else cx++;
        s++;
      }
    }
  }
  free(E);
  return M;
}

/* ------------------------------------------------------------------------ */
/* x recipe: util/generate.hxx:33-41 (hash), :54-79 (uniform_distribution),  */
/* with thrust::default_random_engine == minstd_rand (a=48271, m=2^31-1),    */
/* thrust uniform_real_distribution<double> and the uniform_int -> real+1    */
/* cast of thrust/random/detail/uniform_int_distribution.inl (CUDA 12.9).    */
/* ------------------------------------------------------------------------ */
uint32_t orc_hash(uint32_t a) {
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}
static uint32_t minstd_first(uint32_t seed) {
  const uint64_t m = 2147483647ull;
  uint64_t x = seed % m;
  if (x == 0) x = 1;                    /* linear_congruential_engine::seed */
  return (uint32_t)((48271ull * x) % m);
}
static double unit_real(uint32_t draw) { /* (urng()-min)/(1+(max-min)) */
  return (double)(draw - 1u) / (1.0 + (double)(2147483646u - 1u));
}
void orc_x_recipe_int(i32 n, i32 lo, i32 hi, uint32_t useed, float* out) {
  for (i32 i = 0; i < n; ++i) {
    uint32_t seed = orc_hash((uint32_t)i) * useed;
    double u = unit_real(minstd_first(seed));
    double r = u * (((double)hi + 1.0) - (double)lo) + (double)lo;
    out[i] = (float)(i32)r;
  }
}
void orc_x_recipe_float(i32 n, float lo, float hi, uint32_t useed, float* out) {
  for (i32 i = 0; i < n; ++i) {
    uint32_t seed = orc_hash((uint32_t)i) * useed;
    /* uniform_real_distribution<float>: arithmetic in float */
    float result = (float)(minstd_first(seed) - 1u);
    result /= (1.0f + (float)(2147483646u - 1u));
    out[i] = (result * (hi - lo)) + lo;
  }
}

/* ------------------------------------------------------------------------ */
/* CPU validator: util/reference.hxx:61-76 (spmv), :150-166 (spmv_f64),      */
/* :182-198 (row_l1_products), :116-131 (default_tolerance),                 */
/* :278-337 (rigorously_validate_spmv).                                      */
/* ------------------------------------------------------------------------ */
void orc_spmv_f32(i32 rows, const i32* off, const i32* idx, const float* val,
                  const float* x, float* y) {
  for (i32 r = 0; r < rows; ++r) {
    volatile float sum = 0.0f; /* volatile: forbid FMA contraction / reassoc */
    for (i32 k = off[r]; k < off[r + 1]; ++k) {
      volatile float p = val[k] * x[idx[k]];
      sum = sum + p;
    }
    y[r] = sum;
  }
}
void orc_spmv_f64(i32 rows, const i32* off, const i32* idx, const float* val,
                  const float* x, float* y) {
  for (i32 r = 0; r < rows; ++r) {
    double sum = 0.0;
    for (i32 k = off[r]; k < off[r + 1]; ++k)
      sum += (double)val[k] * (double)x[idx[k]];
    y[r] = (float)sum;
  }
}
/* reference::spmv instantiated for double (util/reference.hxx:61-76): the same
 * loop, every operand and the accumulator in fp64, no contraction. */
void orc_spmv_d(i32 rows, const i32* off, const i32* idx, const double* val,
                const double* x, double* y) {
  for (i32 r = 0; r < rows; ++r) {
    volatile double sum = 0.0;
    for (i32 k = off[r]; k < off[r + 1]; ++k) {
      volatile double p = val[k] * x[idx[k]];
      sum = sum + p;
    }
    y[r] = sum;
  }
}
void orc_row_l1(i32 rows, const i32* off, const i32* idx, const float* val,
                const float* x, float* l1) {
  for (i32 r = 0; r < rows; ++r) {
    double sum = 0.0;
    for (i32 k = off[r]; k < off[r + 1]; ++k)
      sum += fabs((double)val[k] * (double)x[idx[k]]);
    l1[r] = (float)sum;
  }
}
int orc_tolerance_ne(float a, float b) {
  return fabsf(a - b) > 1e-2f + 1e-3f * fabsf(b);
}
i64 orc_count_errors(const float* y, const float* ref, i64 n) {
  i64 e = 0;
  for (i64 i = 0; i < n; ++i) e += orc_tolerance_ne(y[i], ref[i]);
  return e;
}
typedef struct {
  i64 total_rows, naive_mismatches, f32_baseline_overruns, gpu_overruns;
  double max_gpu_abs_error, max_gpu_rel_error, wilkinson_k;
} orc_rigorous_report;

void orc_rigorous_validate(i32 rows, const i32* off, const i32* idx,
                           const float* val, const float* x,
                           const float* y_gpu, double wilkinson_k,
                           double atol_floor, orc_rigorous_report* rep) {
  float* y32 = (float*)malloc(sizeof(float) * (size_t)(rows > 0 ? rows : 1));
  float* y64 = (float*)malloc(sizeof(float) * (size_t)(rows > 0 ? rows : 1));
  float* l1 = (float*)malloc(sizeof(float) * (size_t)(rows > 0 ? rows : 1));
  orc_spmv_f32(rows, off, idx, val, x, y32);
  orc_spmv_f64(rows, off, idx, val, x, y64);
  orc_row_l1(rows, off, idx, val, x, l1);
  memset(rep, 0, sizeof(*rep));
  rep->total_rows = rows; rep->wilkinson_k = wilkinson_k;
  const double eps = 5.96046447753906e-08; /* 2^-24, reference.hxx:204-207 */
  for (i32 r = 0; r < rows; ++r) {
    double nnz_r = (double)(off[r + 1] - off[r]);
    double bound = wilkinson_k * nnz_r * eps * (double)l1[r];
    if (bound < atol_floor) bound = atol_floor;
    double ref = (double)y64[r];
    double f32_err = fabs((double)y32[r] - ref);
    double gpu_err = fabs((double)y_gpu[r] - ref);
    double scale = fabs(ref) > 1.0 ? fabs(ref) : 1.0;
    if (orc_tolerance_ne(y_gpu[r], y32[r])) rep->naive_mismatches++;
    if (f32_err > bound) rep->f32_baseline_overruns++;
    if (gpu_err > bound) rep->gpu_overruns++;
    if (gpu_err > rep->max_gpu_abs_error) rep->max_gpu_abs_error = gpu_err;
    if (gpu_err / scale > rep->max_gpu_rel_error) rep->max_gpu_rel_error = gpu_err / scale;
  }
  free(y32); free(y64); free(l1);
}

/* ------------------------------------------------------------------------ */
/* Format conversions (host): container/coo.hxx:87-98 + detail/convert.hxx:   */
/* 36-66 (offsets->row ids), container/ell.hxx:113-145, bcsr.hxx:111-194.    */
/* ------------------------------------------------------------------------ */
void orc_csr_to_coo_rows(i32 rows, const i32* off, i32* row_idx) {
  for (i32 r = 0; r < rows; ++r)
    for (i32 k = off[r]; k < off[r + 1]; ++k) row_idx[k] = r;
}
i32 orc_ell_pitch(i32 rows, const i32* off) {
  i32 p = 0;
  for (i32 r = 0; r < rows; ++r) if (off[r + 1] - off[r] > p) p = off[r + 1] - off[r];
  return p;
}
/* padding: column sentinel -1, value 0 (ell.hxx:31-36,57-59) */
void orc_csr_to_ell(i32 rows, const i32* off, const i32* idx, const float* val,
                    i32 pitch, i32* e_idx, float* e_val) {
  for (i64 i = 0; i < (i64)rows * pitch; ++i) { e_idx[i] = -1; e_val[i] = 0.0f; }
  for (i32 r = 0; r < rows; ++r)
    for (i32 k = off[r]; k < off[r + 1]; ++k) {
      i64 slot = (i64)r * pitch + (k - off[r]);
      e_idx[slot] = idx[k]; e_val[slot] = val[k];
    }
}
static int cmp_i32(const void* a, const void* b) {
  i32 x = *(const i32*)a, y = *(const i32*)b; return (x > y) - (x < y);
}
/* Two-call protocol: b_off==NULL sizes (*num_blocks), second call fills.
 * Block columns sorted ascending per block-row (bcsr.hxx:148-153); entries
 * are ASSIGNED, not accumulated (:175); padding stays 0. */
int orc_csr_to_bcsr(i32 R, i32 C, i32 rows, i32 cols, const i32* off,
                    const i32* idx, const float* val, i32* num_blocks,
                    i32* b_off, i32* b_col, float* b_val) {
  i32 nbr = (rows + R - 1) / R, nbc = (cols + C - 1) / C;
  i32* local = (i32*)malloc(sizeof(i32) * (size_t)(nbc > 0 ? nbc : 1));
  i32* opened = (i32*)malloc(sizeof(i32) * (size_t)(nbc > 0 ? nbc : 1));
  for (i32 i = 0; i < nbc; ++i) local[i] = -1;
  i32 total = 0;
  if (b_off) b_off[0] = 0;
  for (i32 br = 0; br < nbr; ++br) {
    i32 lo = br * R, hi = lo + R < rows ? lo + R : rows, n = 0;
    for (i32 r = lo; r < hi; ++r)
      for (i32 a = off[r]; a < off[r + 1]; ++a) {
        i32 bc = idx[a] / C;
        if (local[bc] == -1) { local[bc] = n; opened[n++] = bc; }
      }
    qsort(opened, (size_t)n, sizeof(i32), cmp_i32);
    for (i32 k = 0; k < n; ++k) local[opened[k]] = k;
    if (b_off) {
      for (i32 k = 0; k < n; ++k) {
        b_col[total + k] = opened[k];
        memset(b_val + (i64)(total + k) * R * C, 0, sizeof(float) * (size_t)(R * C));
      }
      for (i32 r = lo; r < hi; ++r)
        for (i32 a = off[r]; a < off[r + 1]; ++a) {
          i32 bc = idx[a] / C, j = idx[a] % C;
          b_val[(i64)(total + local[bc]) * R * C + (r - lo) * C + j] = val[a];
        }
      b_off[br + 1] = total + n;
    }
    for (i32 k = 0; k < n; ++k) local[opened[k]] = -1;
    total += n;
  }
  *num_blocks = total;
  free(local); free(opened);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Per-format SpMV bodies (sequential restatements of the kernels' maths):   */
/* coo_thread_mapped.cuh:37-51, ell_thread_mapped.cuh:28-43,                 */
/* bcsr_thread_mapped.cuh:36-74.                                             */
/* ------------------------------------------------------------------------ */
void orc_spmv_coo(i64 nnz, const i32* row, const i32* col, const float* val,
                  const float* x, float* y /* pre-zeroed, rows */) {
  for (i64 a = 0; a < nnz; ++a) {
    volatile float p = val[a] * x[col[a]];
    y[row[a]] = y[row[a]] + p;
  }
}
void orc_spmv_ell(i32 rows, i32 pitch, const i32* e_idx, const float* e_val,
                  const float* x, float* y) {
  for (i32 r = 0; r < rows; ++r) {
    volatile float sum = 0.0f;
    for (i32 s = 0; s < pitch; ++s) {
      i32 c = e_idx[(i64)r * pitch + s];
      if (c >= 0) { volatile float p = e_val[(i64)r * pitch + s] * x[c]; sum = sum + p; }
    }
    y[r] = sum;
  }
}
/* fp32 accumulate in ascending block id, i then j inner order. With bf16
 * storage (config 4) the caller passes values/x already rounded to bf16 and
 * widened back to float: products of two bf16 are exact in fp32. */
void orc_spmv_bcsr(i32 R, i32 C, i32 rows, i32 nbr, const i32* b_off,
                   const i32* b_col, const float* b_val,
                   const float* x /* padded to nbc*C */, float* y) {
  for (i32 br = 0; br < nbr; ++br) {
    float acc[16];
    for (i32 i = 0; i < R; ++i) acc[i] = 0.0f;
    for (i32 b = b_off[br]; b < b_off[br + 1]; ++b) {
      const float* blk = b_val + (i64)b * R * C;
      i64 bc = b_col[b];
      for (i32 i = 0; i < R; ++i)
        for (i32 j = 0; j < C; ++j) {
          volatile float p = blk[i * C + j] * x[bc * C + j];
          acc[i] = acc[i] + p;
        }
    }
    for (i32 i = 0; i < R; ++i)
      if ((i64)br * R + i < rows) y[(i64)br * R + i] = acc[i];
  }
}
/* round-to-nearest-even float -> bf16 -> float */
float orc_bf16_round(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return f;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&f, &u, 4);
  return f;
}

/* ---- CSC / DIA / flat_uniform_occupancy ----------------------
 * csc_t(csr_t): container/csc.hxx:88-102 -- the CSR entries re-sorted by
 * (column, row) (coo.hxx:116-122 sorts the zipped (col, row) keys) and the
 * column ids compressed into offsets (detail/convert.hxx:70-78). A stable
 * counting sort over CSR order gives exactly that order when (row, col) pairs
 * are unique. */
void orc_csr_to_csc(i32 rows, i32 cols, const i32* off, const i32* idx, const float* val,
                    i32* c_off /* cols+1 */, i32* c_row /* nnz */, float* c_val /* nnz */) {
  for (i32 c = 0; c <= cols; ++c) c_off[c] = 0;
  for (i64 a = 0; a < off[rows]; ++a) c_off[idx[a] + 1] += 1;
  for (i32 c = 0; c < cols; ++c) c_off[c + 1] += c_off[c];
  i32* cur = (i32*)malloc(sizeof(i32) * (size_t)(cols > 0 ? cols : 1));
  for (i32 c = 0; c < cols; ++c) cur[c] = c_off[c];
  for (i32 r = 0; r < rows; ++r)
    for (i32 a = off[r]; a < off[r + 1]; ++a) {
      i32 p = cur[idx[a]]++;
      c_row[p] = r;
      c_val[p] = val[a];
    }
  free(cur);
}
/* algorithms/spmv/csc_thread_mapped.cuh:29-41 (sequential restatement: columns
 * ascending, entries of a column in stored order; y pre-zeroed). */
void orc_spmv_csc(i32 cols, const i32* c_off, const i32* c_row, const float* c_val,
                  const float* x, float* y /* pre-zeroed, rows */) {
  for (i32 c = 0; c < cols; ++c)
    for (i32 a = c_off[c]; a < c_off[c + 1]; ++a) {
      volatile float p = c_val[a] * x[c];
      y[c_row[a]] = y[c_row[a]] + p;
    }
}
/* dia_t(csr_t): container/dia.hxx:135-188 -- distinct (col - row) offsets in
 * ascending order, stride = rows, values[d*stride + r] assigned (not added). */
i32 orc_dia_offsets(i32 rows, const i32* off, const i32* idx, i32* out /* cap >= nnz, or NULL to count */,
                    i32 cap) {
  i64 nnz = off[rows];
  i32* all = (i32*)malloc(sizeof(i32) * (size_t)(nnz > 0 ? nnz : 1));
  for (i32 r = 0; r < rows; ++r)
    for (i32 a = off[r]; a < off[r + 1]; ++a) all[a] = idx[a] - r;
  qsort(all, (size_t)nnz, sizeof(i32), cmp_i32);
  i32 n = 0;
  for (i64 a = 0; a < nnz; ++a)
    if (a == 0 || all[a] != all[a - 1]) { if (out && n < cap) out[n] = all[a]; ++n; }
  free(all);
  return n;
}
void orc_csr_to_dia(i32 rows, const i32* off, const i32* idx, const float* val, i32 num_diagonals,
                    const i32* diag_offsets, float* d_val /* num_diagonals * rows, pre-zeroed */) {
  for (i32 r = 0; r < rows; ++r)
    for (i32 a = off[r]; a < off[r + 1]; ++a) {
      i32 o = idx[a] - r, lo = 0, hi = num_diagonals;   /* lower_bound over the sorted offsets */
      while (lo < hi) { i32 mid = (lo + hi) / 2; if (diag_offsets[mid] < o) lo = mid + 1; else hi = mid; }
      d_val[(i64)lo * rows + r] = val[a];
    }
}
/* algorithms/spmv/dia_thread_mapped.cuh:33-54: per row, diagonals ascending. */
void orc_spmv_dia(i32 rows, i32 cols, i64 stride, i32 num_diagonals, const i32* diag_offsets,
                  const float* d_val, const float* x, float* y) {
  for (i32 r = 0; r < rows; ++r) {
    volatile float acc = 0.0f;
    for (i32 d = 0; d < num_diagonals; ++d) {
      i64 c = (i64)r + diag_offsets[d];
      if (c >= 0 && c < cols) { volatile float p = d_val[(i64)d * stride + r] * x[c]; acc = acc + p; }
    }
    y[r] = acc;
  }
}
/* algorithms/spmv/flat_partitioned.cuh:46-60 over layout::flat_uniform_occupancy<K, csr>
 * (container/partitioning.hxx:71-141): windows of K atoms, row = base.tile_of(atom)
 * (layout.hxx:137-148), one add per atom in window order; y pre-zeroed. */
void orc_spmv_flat_partitioned(i32 rows, i32 K, const i32* off, const i32* idx, const float* val,
                               const float* x, float* y /* pre-zeroed */) {
  i64 nnz = off[rows];
  for (i64 t = 0; t * K < nnz; ++t)
    for (i64 a = t * K; a < (t + 1) * K && a < nnz; ++a) {
      i32 lo = 0, hi = rows;                       /* upper_bound(offsets, a) - 1 */
      while (hi - lo > 1) { i32 mid = (lo + hi) / 2; if (off[mid] <= a) lo = mid; else hi = mid; }
      volatile float p = val[a] * x[idx[a]];
      y[lo] = y[lo] + p;
    }
}
/* algorithms/spmm/thread_mapped.cuh:28-53: per (row, col) a sequential sum over the
 * row's atoms in ascending order; B [cols x n], C [rows x n] row-major. */
void orc_spmm(i32 rows, i32 n, const i32* off, const i32* idx, const float* val, const float* B, float* C) {
  for (i32 r = 0; r < rows; ++r)
    for (i32 c = 0; c < n; ++c) {
      volatile float sum = 0.0f;
      for (i32 a = off[r]; a < off[r + 1]; ++a) { volatile float p = val[a] * B[(i64)idx[a] * n + c]; sum = sum + p; }
      C[(i64)r * n + c] = sum;
    }
}
