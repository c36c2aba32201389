/**
 * @file filepath.hxx
 * @brief Path helpers of the example mains (reference include/loops/util/filepath.hxx:18-35).
 */
#pragma once

#include <string>

namespace loops {

/// "a/b/c.mtx" -> "c.mtx".
inline std::string extract_filename(std::string path, std::string delim = "/") {
  const std::size_t cut = path.find_last_of(delim);
  return cut == std::string::npos ? path : path.substr(cut + 1);
}

/// "c.mtx" -> "c".
inline std::string extract_dataset(std::string filename) {
  const std::size_t dot = filename.find_last_of('.');
  return dot == std::string::npos ? filename : filename.substr(0, dot);
}

namespace detail {
inline bool ends_with(const std::string& s, const char* tail) {
  const std::string t(tail);
  return s.size() >= t.size() && s.compare(s.size() - t.size(), t.size(), t) == 0;
}
}  // namespace detail

inline bool is_market(std::string filename) {
  return detail::ends_with(filename, ".mtx") || detail::ends_with(filename, ".mmio");
}

inline bool is_binary_csr(std::string filename) { return detail::ends_with(filename, ".csr"); }

}  // namespace loops
