// oracle/ref_battery.cu -- TEST INFRASTRUCTURE, not product code.
//
// Regenerates the reference's own SpMV test battery by including its fixture
// headers unmodified (reference unittests/test_helpers.hxx:55-278,
// unittests/test_spmv_battery.hxx:52-65): the 9 matrices of standard_battery(),
// make_input_vector (mt19937 seed 23) and reference_spmv. Host-only.
// Built into oracle/_ref/libloopsref_battery.so; tests/golden/make_golden.py
// dumps it into tests/golden/battery.npz.
#include "test_spmv_battery.hxx"

#include <cstring>
#include <string>
#include <vector>

using namespace loops::testing;

static std::vector<battery_case>& cases() {
  static std::vector<battery_case> b = standard_battery();
  return b;
}

extern "C" {

int ref_battery_count() { return (int)cases().size(); }

const char* ref_battery_name(int i) { return cases()[i].name.c_str(); }

void ref_battery_dims(int i, int* rows, int* cols, int* nnz) {
  auto& c = cases()[i].csr;
  *rows = (int)c.rows; *cols = (int)c.cols; *nnz = (int)c.nnzs;
}

void ref_battery_get(int i, int* off, int* idx, float* val, float* x, float* y) {
  auto& c = cases()[i].csr;
  std::copy(c.offsets.begin(), c.offsets.end(), off);
  std::copy(c.indices.begin(), c.indices.end(), idx);
  std::copy(c.values.begin(), c.values.end(), val);
  auto xv = make_input_vector(c);
  std::copy(xv.begin(), xv.end(), x);
  auto yv = reference_spmv(c, xv);
  std::copy(yv.begin(), yv.end(), y);
}

int ref_nearly_equal(float a, float b) { return nearly_equal(a, b) ? 1 : 0; }

}  // extern "C"
