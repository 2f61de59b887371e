// Shared helpers for the pydmrg_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <atomic>

#include "../../include/pydmrg_b200.h"

namespace pdmrg {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define PDMRG_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      pdmrg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

#define PDMRG_CHECK_LAUNCH()                                                                \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      pdmrg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

#define PDMRG_REQUIRE(cond, msg)                                                            \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      pdmrg::set_error("%s (%s:%d)", msg, __FILE__, __LINE__);                              \
      return 2;                                                                             \
    }                                                                                       \
  } while (0)

inline size_t elem_bytes(int dtype) { return dtype == PDMRG_C128 ? 16 : 8; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int num_sms();

// ---- complex helpers on double2 --------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ void cfma(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfmac(double2& acc, double2 a, double2 b) {  // acc += conj(a) * b
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

}  // namespace pdmrg
