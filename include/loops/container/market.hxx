/**
 * @file market.hxx
 * @brief Matrix Market (coordinate) loader -> host COO, with the semantics of the
 * reference loader (include/loops/container/market.hxx:100-289 and
 * detail/mtx_parser.hxx:150-211): object `matrix`, format `coordinate`, field
 * real / integer / pattern (pattern -> value 1), symmetry general / symmetric
 * (off-diagonal entries mirrored right after their source entry); complex,
 * hermitian, skew-symmetric and dense `array` files are rejected; entries keep
 * file order (csr_t(coo) then sorts by (row, col), duplicates preserved).
 * Own implementation: one buffered read and a hand-rolled tokenizer; no mmap,
 * no thrust algorithms. Host-only set-up code, not on the SpMV hot path.
 */
#pragma once

#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include <loops/container/formats.hxx>
#include <loops/error.hxx>
#include <loops/memory.hxx>

namespace loops {

using namespace memory;

namespace detail {
struct mm_typecode_t {
  bool is_matrix = false, is_coordinate = false;
  bool is_real = false, is_integer = false, is_pattern = false, is_complex = false;
  bool is_general = false, is_symmetric = false, is_skew = false, is_hermitian = false;
};
inline std::string lower(std::string s) {
  for (auto& ch : s) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
  return s;
}
}  // namespace detail

template <typename index_t, typename offset_t, typename type_t>
struct matrix_market_t {
  std::string filename;
  std::string dataset;
  detail::mm_typecode_t code;

  coo_t<index_t, type_t, memory_space_t::host> load(std::string _filename) {
    filename = std::move(_filename);
    {
      const std::size_t slash = filename.find_last_of("/\\");
      const std::string base = slash == std::string::npos ? filename : filename.substr(slash + 1);
      const std::size_t dot = base.find_last_of('.');
      dataset = dot == std::string::npos ? base : base.substr(0, dot);
    }
    std::ifstream in(filename, std::ios::binary);
    error::throw_if_exception(!in.good(), "matrix-market: cannot open " + filename);
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    error::throw_if_exception(text.empty(), "matrix-market: empty file " + filename);
    const char* p = text.data();
    const char* end = p + text.size();

    // ---- banner ----
    {
      const char* eol = p;
      while (eol < end && *eol != '\n') ++eol;
      std::istringstream banner(std::string(p, eol));
      std::string tag, object, format, field, symmetry;
      banner >> tag >> object >> format >> field >> symmetry;
      error::throw_if_exception(detail::lower(tag) != "%%matrixmarket",
                                "matrix-market: missing %%MatrixMarket banner in " + filename);
      object = detail::lower(object); format = detail::lower(format);
      field = detail::lower(field); symmetry = detail::lower(symmetry);
      code = detail::mm_typecode_t{};
      code.is_matrix = object == "matrix";
      code.is_coordinate = format == "coordinate";
      code.is_real = field == "real"; code.is_integer = field == "integer";
      code.is_pattern = field == "pattern"; code.is_complex = field == "complex";
      code.is_general = symmetry == "general"; code.is_symmetric = symmetry == "symmetric";
      code.is_skew = symmetry == "skew-symmetric"; code.is_hermitian = symmetry == "hermitian";
      p = eol < end ? eol + 1 : end;
    }
    error::throw_if_exception(!code.is_matrix, "matrix-market: object must be 'matrix' in " + filename);
    error::throw_if_exception(!code.is_coordinate,
                              "matrix-market: only the coordinate (sparse) format is supported in " + filename);
    error::throw_if_exception(code.is_complex, "matrix-market: complex values not supported in " + filename);
    error::throw_if_exception(code.is_hermitian || code.is_skew,
                              "matrix-market: hermitian / skew-symmetric not supported in " + filename);
    error::throw_if_exception(!(code.is_general || code.is_symmetric),
                              "matrix-market: missing or unrecognized symmetry tag in " + filename);
    error::throw_if_exception(!(code.is_real || code.is_integer || code.is_pattern),
                              "matrix-market: missing or unrecognized field tag in " + filename);

    auto skip_ws = [&](const char* q) { while (q < end && std::isspace(static_cast<unsigned char>(*q))) ++q; return q; };
    auto skip_blank = [&](const char* q) { while (q < end && (*q == ' ' || *q == '\t' || *q == '\r')) ++q; return q; };
    auto skip_to_eol = [&](const char* q) { while (q < end && *q != '\n') ++q; return q < end ? q + 1 : end; };
    auto parse_size = [&](const char*& q, std::size_t& out) {
      const char* s = q;
      std::size_t v = 0;
      while (q < end && *q >= '0' && *q <= '9') { v = v * 10 + std::size_t(*q - '0'); ++q; }
      out = v;
      return q != s;
    };
    // ---- comments, dimension line ----
    for (;;) {
      p = skip_ws(p);
      if (p < end && *p == '%') p = skip_to_eol(p); else break;
    }
    std::size_t num_rows = 0, num_cols = 0, header_nnz = 0;
    error::throw_if_exception(!parse_size(p, num_rows), "matrix-market: expected dimension line (first integer)");
    p = skip_blank(p);
    error::throw_if_exception(!parse_size(p, num_cols), "matrix-market: expected dimension line (second integer)");
    p = skip_blank(p);
    error::throw_if_exception(!parse_size(p, header_nnz), "matrix-market: expected dimension line (third integer)");
    p = skip_to_eol(p);
    error::throw_if_exception(num_rows >= std::size_t(std::numeric_limits<index_t>::max()) ||
                                  num_cols >= std::size_t(std::numeric_limits<index_t>::max()),
                              "matrix-market: index_t overflow (rows or cols >= INT_MAX) in " + filename);

    // ---- body: one pass into growing vectors (mirrors appended in place) ----
    std::vector<index_t> r_idx, c_idx;
    std::vector<type_t> vals;
    r_idx.reserve(header_nnz * (code.is_symmetric ? 2 : 1));
    c_idx.reserve(r_idx.capacity());
    vals.reserve(r_idx.capacity());
    for (std::size_t i = 0; i < header_nnz; ++i) {
      p = skip_ws(p);
      std::size_t row1 = 0, col1 = 0;
      error::throw_if_exception(!parse_size(p, row1), "matrix-market: expected row index in body");
      p = skip_blank(p);
      error::throw_if_exception(!parse_size(p, col1), "matrix-market: expected column index in body");
      double weight = 1.0;
      if (!code.is_pattern) {
        p = skip_blank(p);
        char* after = nullptr;
        weight = std::strtod(p, &after);
        error::throw_if_exception(after == p, "matrix-market: expected value in body");
        p = after;
      }
      p = skip_to_eol(p);
      error::throw_if_exception(row1 == 0 || col1 == 0,
                                "matrix-market: zero-indexed entry (Matrix Market is 1-indexed)");
      error::throw_if_exception(row1 > num_rows || col1 > num_cols,
                                "matrix-market: entry outside the declared dimensions in " + filename);
      const index_t r = static_cast<index_t>(row1 - 1), c = static_cast<index_t>(col1 - 1);
      const type_t v = static_cast<type_t>(weight);
      r_idx.push_back(r); c_idx.push_back(c); vals.push_back(v);
      if (code.is_symmetric && r != c) { r_idx.push_back(c); c_idx.push_back(r); vals.push_back(v); }
    }
    error::throw_if_exception(r_idx.size() >= std::size_t(std::numeric_limits<offset_t>::max()),
                              "matrix-market: offset_t overflow (final nnz exceeds offset_t max) in " + filename);
    coo_t<index_t, type_t, memory_space_t::host> coo(num_rows, num_cols, r_idx.size());
    std::copy(r_idx.begin(), r_idx.end(), coo.row_indices.begin());
    std::copy(c_idx.begin(), c_idx.end(), coo.col_indices.begin());
    std::copy(vals.begin(), vals.end(), coo.values.begin());
    return coo;
  }
};

}  // namespace loops
