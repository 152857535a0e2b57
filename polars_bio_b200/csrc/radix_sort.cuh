// radix_sort.cuh -- hand-written stable LSD radix sort of (uint64 key, uint64 value) pairs.
//
// One pass = 8 key bits: per-tile digit histogram -> device-wide exclusive scan of the
// [digit][tile] counts -> stable scatter.  The scatter ranks keys inside a tile with
// warp-level __match_any_sync multisplit (one shared-memory counter row per warp), so equal
// digits keep their input order; that stability is what makes the composite
// (contig | start) sort double as "radix partition by contig + segmented sort by start"
// and keeps (start,row) order among equal ends in the second (end) sort.
//
// HBM traffic per pass: 8 B/key histogram read + 16 B read + 16 B write per element.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace pbgpu {

constexpr int kRsThreads = 512;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 8;  // per thread
constexpr int kRsTile = kRsThreads * kRsItems;
constexpr int kRsRadix = 256;

__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                             uint32_t *__restrict__ hist /*[256][nblk]*/, int64_t nblk) {
  __shared__ uint32_t h[kRsRadix];
  for (int i = threadIdx.x; i < kRsRadix; i += kRsThreads) h[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll
  for (int j = 0; j < kRsItems; ++j) {
    int64_t i = base + (int64_t)j * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRsRadix; i += kRsThreads) hist[(int64_t)i * nblk + blockIdx.x] = h[i];
}

__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const uint64_t *__restrict__ keys_in,
                                                                const uint64_t *__restrict__ vals_in,
                                                                uint64_t *__restrict__ keys_out,
                                                                uint64_t *__restrict__ vals_out, int64_t n, int shift,
                                                                const uint32_t *__restrict__ hist_scanned, int64_t nblk) {
  __shared__ uint32_t wcnt[kRsWarps][kRsRadix];  // per-warp digit counters -> exclusive warp offsets
  __shared__ uint32_t dbase[kRsRadix];           // global offset of (digit, this tile)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kRsWarps * kRsRadix; i += kRsThreads) (&wcnt[0][0])[i] = 0;
  for (int i = threadIdx.x; i < kRsRadix; i += kRsThreads) dbase[i] = hist_scanned[(int64_t)i * nblk + blockIdx.x];
  __syncthreads();

  // warp w owns the contiguous slice [base + w*256, +256): round r covers 32 consecutive keys
  const int64_t wbase = (int64_t)blockIdx.x * kRsTile + (int64_t)warp * (32 * kRsItems);
  uint64_t k[kRsItems], v[kRsItems];
  uint32_t rank[kRsItems];
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool ok = i < n;
    k[r] = ok ? keys_in[i] : 0;
    v[r] = ok ? vals_in[i] : 0;
    const unsigned d = ok ? (unsigned)((k[r] >> shift) & 0xff) : 0x100u;  // 0x100: out-of-range lanes group together
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (ok && lane == leader) {
      old = wcnt[warp][d];
      wcnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over warps, per digit
  for (int d = threadIdx.x; d < kRsRadix; d += kRsThreads) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      uint32_t t = wcnt[w][d];
      wcnt[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const unsigned d = (unsigned)((k[r] >> shift) & 0xff);
      const uint32_t dst = dbase[d] + wcnt[warp][d] + rank[r];
      keys_out[dst] = k[r];
      vals_out[dst] = v[r];
    }
  }
}

// Sorts n pairs by the low `bits` bits of the key.  keys/vals are overwritten with the sorted
// result (internally ping-pongs with scratch).  n < 2^32.
inline int radix_sort_pairs(uint64_t *keys, uint64_t *vals, int64_t n, int bits, cudaStream_t s) {
  if (n <= 1 || bits <= 0) return PBGPU_OK;
  const int64_t nblk = cdiv(n, kRsTile);
  Scratch sc(s);
  uint64_t *k2 = nullptr, *v2 = nullptr;
  uint32_t *hist = nullptr;
  PB_TRY(sc.get(&k2, (size_t)n));
  PB_TRY(sc.get(&v2, (size_t)n));
  PB_TRY(sc.get(&hist, (size_t)(nblk * kRsRadix)));
  uint64_t *ki = keys, *vi = vals, *ko = k2, *vo = v2;
  for (int shift = 0; shift < bits; shift += 8) {
    PB_LAUNCH(rs_hist_kernel, (unsigned)nblk, kRsThreads, 0, s, ki, n, shift, hist, nblk);
    PB_TRY((device_scan<SumU32, false>(hist, hist, nblk * kRsRadix, nullptr, s)));
    PB_LAUNCH(rs_scatter_kernel, (unsigned)nblk, kRsThreads, 0, s, ki, vi, ko, vo, n, shift, hist, nblk);
    PB_CHECK_LAUNCH();
    uint64_t *t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  if (ki != keys) {
    PB_CUDA(cudaMemcpyAsync(keys, ki, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    PB_CUDA(cudaMemcpyAsync(vals, vi, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  }
  return PBGPU_OK;
}

}  // namespace pbgpu
