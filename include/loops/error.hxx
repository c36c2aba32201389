/**
 * @file error.hxx
 * @brief Exceptions (reference include/loops/error.hxx:22-47) plus the
 * translation of C-ABI status codes into them.
 */
#pragma once
#include <exception>
#include <string>
#include <loopsb.h>

namespace loops {
namespace error {

struct exception_t : std::exception {
  std::string report;
  explicit exception_t(std::string message) : report(std::move(message)) {}
  const char* what() const noexcept override { return report.c_str(); }
};

inline void throw_if_exception(bool is_exception, std::string message = "") {
  if (is_exception)
    throw exception_t(message);
}

/// loopsb_status_t -> exception_t (the C ABI never throws by itself).
inline void throw_if_status(int status, const char* where) {
  if (status != LOOPSB_OK)
    throw exception_t(std::string(where) + ": " + loopsb_status_string(status) +
                      " (" + loopsb_last_error() + ")");
}

}  // namespace error
}  // namespace loops
