// dgs_internal.h -- shared host-side helpers of libdgs_b200.so (not part of the public C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "dgs_b200.h"

namespace dgs {

void set_error(const char* fmt, ...);

#define DGS_CUDA_OK(expr)                                                                 \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::dgs::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return DGS_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// after a kernel launch: always check the launch itself; in debug mode also synchronise
#define DGS_LAUNCH_OK(stream, debug)                                   \
  do {                                                                 \
    DGS_CUDA_OK(cudaGetLastError());                                   \
    if (debug) DGS_CUDA_OK(cudaStreamSynchronize(stream));             \
  } while (0)

#define DGS_REQUIRE(cond, ...)             \
  do {                                     \
    if (!(cond)) {                         \
      ::dgs::set_error(__VA_ARGS__);       \
      return DGS_ERR_INVALID_ARGUMENT;     \
    }                                      \
  } while (0)

// Carves typed, 256-byte aligned sub-buffers out of one arena (or just measures, base == nullptr).
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  size_t bytes() const { return (off + 255) & ~size_t(255); }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace dgs
