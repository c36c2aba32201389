// oracle/ref_host.cu -- TEST INFRASTRUCTURE, not product code.
//
// Thin extern "C" driver around the UNMODIFIED gunrock/loops reference headers
// (compiled from where they lie: -I/root/reference/include; outputs only into
// oracle/_ref/). Everything here runs on the HOST (memory_space_t::host), so
// it works in the GPU-less build container and on the GPU box's CPU cores.
//
// Used for two things only:
//   * pinning the C restatement in oracle/loops_oracle.c (tests/),
//   * the CPU baseline leg of bench.py ("kind": "reference").
//
// Reference entry points exercised (file:line relative to /root/reference):
//   include/loops/util/reference.hxx:61-76     reference::spmv
//   include/loops/util/reference.hxx:150-166   reference::spmv_f64
//   include/loops/util/reference.hxx:182-198   reference::row_l1_products
//   include/loops/util/generate.hxx:54-79      generate::random::uniform_distribution
//   include/loops/container/market.hxx:100-177 matrix_market_t::load
//   include/loops/container/csr.hxx:86-94      csr_t(coo_t)
//   include/loops/container/ell.hxx:113-145    ell_t(csr_t)
//   include/loops/container/bcsr.hxx:111-194   bcsr_t(csr_t)
//   include/loops/container/coo.hxx:87-98      coo_t(csr_t)
//   include/loops/container/csc.hxx:88-102     csc_t(csr_t)
//   include/loops/container/dia.hxx:135-188    dia_t(csr_t)

#include <loops/container/formats.hxx>
#include <loops/container/market.hxx>
#include <loops/container/vector.hxx>
#include <loops/memory.hxx>
#include <loops/util/generate.hxx>
#include <loops/util/reference.hxx>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

using namespace loops;
using namespace loops::memory;

using csr_h = csr_t<int, int, float, memory_space_t::host>;
using coo_h = coo_t<int, float, memory_space_t::host>;
using ell_h = ell_t<int, float, memory_space_t::host>;
using vec_h = vector_t<float, memory_space_t::host>;

static csr_h make_csr(int rows, int cols, int nnz, const int* off,
                      const int* idx, const float* val) {
  csr_h c(rows, cols, nnz);
  std::copy(off, off + rows + 1, c.offsets.begin());
  std::copy(idx, idx + nnz, c.indices.begin());
  std::copy(val, val + nnz, c.values.begin());
  return c;
}

static csr_h g_loaded;  // last matrix loaded through ref_load_mtx

template <std::size_t R, std::size_t C>
static int bcsr_convert(const csr_h& c, int* num_blocks, int* b_off,
                        int* b_col, float* b_val, int cap_blocks) {
  bcsr_t<R, C, int, int, float, memory_space_t::host> b(c);
  *num_blocks = (int)b.num_blocks;
  if (b_off == nullptr)
    return 0;
  if ((int)b.num_blocks > cap_blocks)
    return 2;
  std::copy(b.block_offsets.begin(), b.block_offsets.end(), b_off);
  std::copy(b.block_col_indices.begin(), b.block_col_indices.end(), b_col);
  std::copy(b.values.begin(), b.values.end(), b_val);
  return 0;
}

extern "C" {

// ---- Matrix Market (config 1) -------------------------------------------
int ref_load_mtx(const char* path, int* rows, int* cols, int* nnz) {
  try {
    matrix_market_t<int, int, float> mtx;
    coo_h coo = mtx.load(path);
    g_loaded = csr_h(coo);
    *rows = (int)g_loaded.rows;
    *cols = (int)g_loaded.cols;
    *nnz = (int)g_loaded.nnzs;
    return 0;
  } catch (...) {
    return 1;
  }
}

void ref_loaded_csr(int* off, int* idx, float* val) {
  std::copy(g_loaded.offsets.begin(), g_loaded.offsets.end(), off);
  std::copy(g_loaded.indices.begin(), g_loaded.indices.end(), idx);
  std::copy(g_loaded.values.begin(), g_loaded.values.end(), val);
}

// ---- x recipe: examples/spmv/merge_path.cu:33-34 --------------------------
// Note the reference passes int bounds (1, 10), so type_t deduces to int and
// the integer branch of the generator runs; results are stored into floats.
void ref_x_recipe_int(int n, int lo, int hi, unsigned seed, float* out) {
  vec_h x(n);
  generate::random::uniform_distribution(x.begin(), x.end(), lo, hi, seed);
  std::copy(x.begin(), x.end(), out);
}

void ref_x_recipe_float(int n, float lo, float hi, unsigned seed, float* out) {
  vec_h x(n);
  generate::random::uniform_distribution(x.begin(), x.end(), lo, hi, seed);
  std::copy(x.begin(), x.end(), out);
}

unsigned ref_hash(unsigned a) { return generate::random::hash(a); }

// ---- CPU validator ---------------------------------------------------------
void ref_spmv_f32(int rows, int cols, int nnz, const int* off, const int* idx,
                  const float* val, const float* x, float* y) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  vec_h xv(x, x + cols);
  auto r = reference::spmv(c, xv);
  std::copy(r.begin(), r.end(), y);
}

void ref_spmv_f64(int rows, int cols, int nnz, const int* off, const int* idx,
                  const float* val, const float* x, float* y) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  vec_h xv(x, x + cols);
  auto r = reference::spmv_f64(c, xv);
  std::copy(r.begin(), r.end(), y);
}

// reference::spmv with value_t = double (what the .f64 example builds validate against)
void ref_spmv_d(int rows, int cols, int nnz, const int* off, const int* idx,
                const double* val, const double* x, double* y) {
  csr_t<int, int, double, memory_space_t::host> c(rows, cols, nnz);
  std::copy(off, off + rows + 1, c.offsets.begin());
  std::copy(idx, idx + nnz, c.indices.begin());
  std::copy(val, val + nnz, c.values.begin());
  vector_t<double, memory_space_t::host> xv(x, x + cols);
  auto r = reference::spmv(c, xv);
  std::copy(r.begin(), r.end(), y);
}

void ref_row_l1(int rows, int cols, int nnz, const int* off, const int* idx,
                const float* val, const float* x, float* l1) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  vec_h xv(x, x + cols);
  auto r = reference::row_l1_products(c, xv);
  std::copy(r.begin(), r.end(), l1);
}

int ref_default_tolerance_ne(float a, float b) {
  return reference::default_tolerance<float>::ne(a, b) ? 1 : 0;
}

// ---- format conversions ----------------------------------------------------
int ref_ell_pitch(int rows, int cols, int nnz, const int* off, const int* idx,
                  const float* val) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  return (int)ell_h::max_nnz_per_row(c);
}

void ref_csr_to_ell(int rows, int cols, int nnz, const int* off,
                    const int* idx, const float* val, int* e_idx,
                    float* e_val) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  ell_h e(c);
  std::copy(e.indices.begin(), e.indices.end(), e_idx);
  std::copy(e.values.begin(), e.values.end(), e_val);
}

void ref_csr_to_coo_rows(int rows, int cols, int nnz, const int* off,
                         const int* idx, const float* val, int* row_idx) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  coo_h o(c);
  std::copy(o.row_indices.begin(), o.row_indices.end(), row_idx);
}

// R==C in {2,3,4}. Call once with b_off==NULL to size, then again to fill.
int ref_csr_to_bcsr(int R, int rows, int cols, int nnz, const int* off,
                    const int* idx, const float* val, int* num_blocks,
                    int* b_off, int* b_col, float* b_val, int cap_blocks) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  switch (R) {
    case 2:
      return bcsr_convert<2, 2>(c, num_blocks, b_off, b_col, b_val, cap_blocks);
    case 3:
      return bcsr_convert<3, 3>(c, num_blocks, b_off, b_col, b_val, cap_blocks);
    case 4:
      return bcsr_convert<4, 4>(c, num_blocks, b_off, b_col, b_val, cap_blocks);
    default:
      return 1;
  }
}

// csc_t(csr_t): column offsets, row indices and values in the container's order.
void ref_csr_to_csc(int rows, int cols, int nnz, const int* off, const int* idx,
                    const float* val, int* c_off, int* c_row, float* c_val) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  csc_t<int, int, float, memory_space_t::host> s(c);
  std::copy(s.offsets.begin(), s.offsets.end(), c_off);
  std::copy(s.indices.begin(), s.indices.end(), c_row);
  std::copy(s.values.begin(), s.values.end(), c_val);
}

// dia_t(csr_t): returns num_diagonals; fills the arrays when they are non-null
// and large enough (diag_offsets[num_diagonals], values[num_diagonals * rows]).
int ref_csr_to_dia(int rows, int cols, int nnz, const int* off, const int* idx,
                   const float* val, int* diag_offsets, float* d_val, int cap_diagonals) {
  csr_h c = make_csr(rows, cols, nnz, off, idx, val);
  dia_t<int, int, float, memory_space_t::host> d(c);
  const int n = (int)d.num_diagonals;
  if (diag_offsets && d_val && n <= cap_diagonals) {
    std::copy(d.diag_offsets.begin(), d.diag_offsets.end(), diag_offsets);
    std::copy(d.values.begin(), d.values.end(), d_val);
  }
  return n;
}

// ---- CPU baseline timing ---------------------------------------------------
// Times reference::spmv (as shipped: includes its internal host copies,
// reference.hxx:64-66) on `threads` host threads, each thread calling the
// unmodified function on its own contiguous row slice. threads==1 is the
// reference exactly as shipped. Returns seconds per SpMV (best of `reps`).
double ref_time_spmv(int rows, int cols, int nnz, const int* off,
                     const int* idx, const float* val, const float* x,
                     int reps, int threads, float* y_out) {
  if (threads < 1)
    threads = 1;
  std::vector<csr_h> parts;
  std::vector<int> row_lo;
  for (int t = 0; t < threads; ++t) {
    int lo = (int)((long long)rows * t / threads);
    int hi = (int)((long long)rows * (t + 1) / threads);
    int a = off[lo], b = off[hi];
    csr_h p(hi - lo, cols, b - a);
    for (int r = lo; r <= hi; ++r)
      p.offsets[r - lo] = off[r] - a;
    std::copy(idx + a, idx + b, p.indices.begin());
    std::copy(val + a, val + b, p.values.begin());
    parts.push_back(std::move(p));
    row_lo.push_back(lo);
  }
  vec_h xv(x, x + cols);
  double best = 1e30;
  for (int rep = 0; rep < reps; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) {
      th.emplace_back([&, t]() {
        auto r = reference::spmv(parts[t], xv);
        if (y_out)
          std::copy(r.begin(), r.end(), y_out + row_lo[t]);
      });
    }
    for (auto& q : th)
      q.join();
    auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  return best;
}

}  // extern "C"
