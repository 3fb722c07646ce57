// Shared host-side helpers for libmrgpu (status codes, error plumbing).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/mr_b200.h"

namespace mr {

// Internal exception: never crosses the C ABI (api.cu catches and converts to a status).
struct Error : std::runtime_error {
  mr_status code;
  Error(mr_status c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

[[noreturn]] inline void fail(mr_status code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

#define MR_CUDA_CHECK(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      ::mr::fail(MR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                 __LINE__);                                                                  \
  } while (0)

}  // namespace mr
