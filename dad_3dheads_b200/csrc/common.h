// Shared host-side helpers for libdad3d.so (error string, launch counter, CUDA error checks).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <string>

namespace dad3d {

void set_error(const std::string& msg);
extern std::atomic<unsigned long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(static_cast<unsigned long long>(n), std::memory_order_relaxed); }

#define DAD3D_CUDA_OK(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::dad3d::set_error(std::string(#expr) + " -> " + cudaGetErrorString(_e));                 \
      return DAD3D_ERR_CUDA;                                                                    \
    }                                                                                           \
  } while (0)

#define DAD3D_REQUIRE(cond, msg)                      \
  do {                                                \
    if (!(cond)) {                                    \
      ::dad3d::set_error(std::string("invalid argument: ") + (msg)); \
      return DAD3D_ERR_INVALID;                       \
    }                                                 \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace dad3d
