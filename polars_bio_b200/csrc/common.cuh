// common.cuh -- error plumbing, stream-ordered allocation, launch accounting, search helpers.
// Part of libpbgpu.so (sm_100a only).  No torch, no Thrust/CUB: every kernel here is ours.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/pbgpu.h"

namespace pbgpu {

// ---- thread-local error text ------------------------------------------------------------
extern thread_local char g_err[512];
int set_error(int code, const char *fmt, ...);

#define PB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return pbgpu::set_error(PBGPU_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                            \
  } while (0)

#define PB_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != PBGPU_OK) return _rc; \
  } while (0)

// ---- launch accounting (bench.py reports gpu_launches from this) --------------------------
extern std::atomic<uint64_t> g_launches;
#define PB_LAUNCH(kernel, grid, block, smem, stream, ...)                     \
  do {                                                                        \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);               \
    pbgpu::g_launches.fetch_add(1, std::memory_order_relaxed);                \
  } while (0)

#define PB_CHECK_LAUNCH() PB_CUDA(cudaGetLastError())

// ---- stream-ordered scratch -------------------------------------------------------------
int dev_alloc(void **p, size_t bytes, cudaStream_t s);
void dev_free(void *p, cudaStream_t s);

template <typename T>
inline int dev_alloc_t(T **p, size_t count, cudaStream_t s) {
  return dev_alloc(reinterpret_cast<void **>(p), sizeof(T) * (count ? count : 1), s);
}

// RAII holder for scratch freed on the same stream
struct Scratch {
  cudaStream_t s;
  void *ptrs[32];
  int n = 0;
  explicit Scratch(cudaStream_t st) : s(st) {}
  ~Scratch() {
    for (int i = 0; i < n; ++i) dev_free(ptrs[i], s);
  }
  template <typename T>
  int get(T **p, size_t count) {
    int rc = dev_alloc_t(p, count, s);
    if (rc == PBGPU_OK && n < 32) ptrs[n++] = *p;
    return rc;
  }
};

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------------
constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// first index in [lo,hi) with a[idx] >= x
__device__ __forceinline__ int32_t lower_bound_i32(const int32_t *__restrict__ a, int32_t lo, int32_t hi, int32_t x) {
  while (lo < hi) {
    int32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// first index in [lo,hi) with a[idx] > x
__device__ __forceinline__ int32_t upper_bound_i32(const int32_t *__restrict__ a, int32_t lo, int32_t hi, int32_t x) {
  while (lo < hi) {
    int32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(a + mid) <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Lanes of the warp holding the same 9-bit value (bit 8 = "no element"): the ballot form of __match_any_sync.  The MATCH
// instruction walks the distinct values of the warp one by one -- with random 8-bit digits that is ~28 rounds, and ncu
// (r2h) showed the radix passes stalled on its result (short scoreboard 32 %, DRAM 27 % busy).  Nine VOTE + LOP3 pairs
// cost the same whatever the values are.
__device__ __forceinline__ unsigned match9(unsigned d) {
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 9; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned v = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? v : ~v;
  }
  return peers;
}

// the same over 8-bit values (every lane holds an element)
__device__ __forceinline__ unsigned match8(unsigned d) {
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned v = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? v : ~v;
  }
  return peers;
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace pbgpu
