// common.cuh -- shared host/device helpers for libgulon_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gulon_b200.h"

namespace gulon {

typedef unsigned long long u64;
typedef long long i64;

// ---- errors -----------------------------------------------------------------------------
inline std::string &err_slot() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  err_slot() = buf;
  return code;
}
#define GCU(x)                                                                                   \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess)                                                                       \
      return ::gulon::fail(e_ == cudaErrorMemoryAllocation ? GULON_ENOMEM : GULON_ECUDA,         \
                           "%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_));     \
  } while (0)
#define GCHECK(x)                                                                                \
  do {                                                                                           \
    int s_ = (x);                                                                                \
    if (s_ != GULON_OK) return s_;                                                               \
  } while (0)
#define GREQUIRE(cond, ...)                                                                      \
  do {                                                                                           \
    if (!(cond)) return ::gulon::fail(GULON_EINVAL, __VA_ARGS__);                                \
  } while (0)

// counts kernels launched by this library (bench.py reports it as gpu_launches)
inline std::atomic<long long> &launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}
#define GLAUNCH(kernel, grid, block, smem, stream, ...)                                          \
  do {                                                                                           \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                  \
    ::gulon::launch_counter().fetch_add(1, std::memory_order_relaxed);                           \
    GCU(cudaGetLastError());                                                                     \
  } while (0)

// ---- growable device buffer ---------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return GULON_OK;
    if (p) {
      GCU(cudaDeviceSynchronize());
      GCU(cudaFree(p));
      p = nullptr;
      cap = 0;
    }
    size_t want = bytes + (bytes >> 3) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(GULON_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return GULON_OK;
  }
  template <typename T>
  T *as() const { return reinterpret_cast<T *>(p); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

__host__ __device__ inline i64 round_up(i64 x, i64 m) { return (x + m - 1) / m * m; }
__host__ __device__ inline i64 ceil_div(i64 x, i64 m) { return (x + m - 1) / m; }

// ---- (distance, id) keys --------------------------------------------------------------------
// A candidate is one u64: order-preserving image of the fp32 distance in the high word, row id in
// the low word, so that u64 '<' is exactly (distance asc, id asc).  SENT marks an empty slot.
constexpr u64 KEY_SENT = ~0ULL;

__host__ __device__ inline uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  // every NaN orders as the largest value (after +inf), whatever its sign or payload: the keys then
  // form a total order.  (The reference heap's treatment of NaN is order-dependent, G/TopKHeap.scala:69-79.)
  if ((u & 0x7fffffffu) > 0x7f800000u) u = 0x7fffffffu;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
__host__ __device__ inline u64 make_key(float d, uint32_t id) {
  return ((u64)f2ord(d) << 32) | (u64)id;
}

}  // namespace gulon
