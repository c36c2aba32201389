// tests/cpp/market_load.cu -- host-only check of include/loops/container/market.hxx:
//   market_load FILE  -> prints "rows cols nnz" then one "row col value" line per COO
//   entry in the loader's order, then the CSR arrays of csr_t(coo). Exit 2 with the
//   exception text on stderr when the loader rejects the file. No device is touched.
#include <cstdio>
#include <loops/container/market.hxx>

int main(int argc, char** argv) {
  using namespace loops;
  if (argc < 2) return 64;
  try {
    matrix_market_t<int, int, float> mtx;
    coo_t<int, float, memory_space_t::host> coo = mtx.load(argv[1]);
    std::printf("%zu %zu %zu\n", coo.rows, coo.cols, coo.nnzs);
    for (std::size_t i = 0; i < coo.nnzs; ++i)
      std::printf("%d %d %.9g\n", int(coo.row_indices[i]), int(coo.col_indices[i]), double(coo.values[i]));
    csr_t<int, int, float, memory_space_t::host> csr(coo);
    for (std::size_t i = 0; i <= csr.rows; ++i) std::printf("%d ", int(csr.offsets[i]));
    std::printf("\n");
    for (std::size_t i = 0; i < csr.nnzs; ++i) std::printf("%d ", int(csr.indices[i]));
    std::printf("\n");
    for (std::size_t i = 0; i < csr.nnzs; ++i) std::printf("%.9g ", double(csr.values[i]));
    std::printf("\n");
  } catch (const error::exception_t& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  return 0;
}
