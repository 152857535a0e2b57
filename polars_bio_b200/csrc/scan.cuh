// scan.cuh -- hand-written device-wide prefix scans (reduce -> scan partials -> rescan).
// Used for radix-sort digit offsets (uint32 sum), the running max of `end` per contig
// (uint64 max over contig-tagged keys) and the pair-offset block bases (uint64 sum).
#pragma once
#include "common.cuh"

namespace pbgpu {

struct SumU32 { using T = uint32_t; __device__ static T id() { return 0u; } __device__ static T op(T a, T b) { return a + b; } };
struct SumU64 { using T = unsigned long long; __device__ static T id() { return 0ull; } __device__ static T op(T a, T b) { return a + b; } };
struct MaxU64 { using T = unsigned long long; __device__ static T id() { return 0ull; } __device__ static T op(T a, T b) { return a > b ? a : b; } };

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Exclusive scan of one value per thread across the block; returns the block total via smem.
template <typename Op, int THREADS>
__device__ __forceinline__ typename Op::T block_exclusive(typename Op::T v, typename Op::T *warp_tot /*[THREADS/32+1]*/) {
  using T = typename Op::T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl = Op::op(o, incl);
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    constexpr int NW = THREADS / 32;
    T w = lane < NW ? warp_tot[lane] : Op::id();
    T wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi = Op::op(o, wi);
    }
    T we = __shfl_up_sync(0xffffffffu, wi, 1);
    if (lane == 0) we = Op::id();
    if (lane < NW) warp_tot[lane] = we;
    if (lane == NW - 1) warp_tot[NW] = wi;  // block total
  }
  __syncthreads();
  T excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = Op::id();
  return Op::op(warp_tot[warp], excl);
}

template <typename Op>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const typename Op::T *__restrict__ in, int64_t n,
                                                                   typename Op::T *__restrict__ partial) {
  using T = typename Op::T;
  __shared__ T wt[kScanThreads / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  T acc = Op::id();
  // striped (coalesced) loads: order does not matter for a commutative reduction
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + (int64_t)j * kScanThreads + threadIdx.x;
    if (i < n) acc = Op::op(acc, in[i]);
  }
  (void)block_exclusive<Op, kScanThreads>(acc, wt);
  if (threadIdx.x == 0) partial[blockIdx.x] = wt[kScanThreads / 32];
}

// single block: exclusive scan of the partials in place (looping over chunks)
template <typename Op>
__global__ void __launch_bounds__(1024) scan_partials_kernel(typename Op::T *__restrict__ partial, int64_t nparts,
                                                             typename Op::T *__restrict__ total_out,
                                                             volatile unsigned long long *mailbox = nullptr,
                                                             unsigned long long mailbox_seq = 0) {
  using T = typename Op::T;
  __shared__ T wt[1024 / 32 + 1];
  __shared__ T carry_s;
  if (threadIdx.x == 0) carry_s = Op::id();
  __syncthreads();
  for (int64_t base = 0; base < nparts; base += 1024) {
    int64_t i = base + threadIdx.x;
    T v = i < nparts ? partial[i] : Op::id();
    T ex = block_exclusive<Op, 1024>(v, wt);
    T carry = carry_s;
    if (i < nparts) partial[i] = Op::op(carry, ex);
    __syncthreads();
    if (threadIdx.x == 0) carry_s = Op::op(carry, wt[1024 / 32]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (total_out) *total_out = carry_s;
    if (mailbox) {  // host-mapped words (pbgpu.cu): total, then the sequence number
      mailbox[0] = (unsigned long long)carry_s;
      __threadfence_system();
      mailbox[15] = mailbox_seq;
    }
  }
}

template <typename Op, bool INCLUSIVE>
__global__ void __launch_bounds__(kScanThreads) scan_final_kernel(const typename Op::T *__restrict__ in,
                                                                  typename Op::T *__restrict__ out, int64_t n,
                                                                  const typename Op::T *__restrict__ partial) {
  using T = typename Op::T;
  __shared__ T wt[kScanThreads / 32 + 1];
  __shared__ T stage[kScanTile];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  // coalesced load into smem, then blocked per-thread runs
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + (int64_t)j * kScanThreads + threadIdx.x;
    stage[j * kScanThreads + threadIdx.x] = i < n ? in[i] : Op::id();
  }
  __syncthreads();
  T v[kScanItems];
  T acc = Op::id();
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = stage[threadIdx.x * kScanItems + j];
    acc = Op::op(acc, v[j]);
  }
  T ex = block_exclusive<Op, kScanThreads>(acc, wt);
  T run = Op::op(partial[blockIdx.x], ex);
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    T nxt = Op::op(run, v[j]);
    stage[threadIdx.x * kScanItems + j] = INCLUSIVE ? nxt : run;
    run = nxt;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + (int64_t)j * kScanThreads + threadIdx.x;
    if (i < n) out[i] = stage[j * kScanThreads + threadIdx.x];
  }
}

// whole scan in one block (one launch instead of three) for small arrays (contig tables, a few thousand
// block totals); its serial loop costs ~2 us per 4096 entries, so larger inputs take the three-kernel path
// mailbox (optional; uint64 totals only): host-mapped words, see pbgpu.cu -- the total goes to word 0, then the
// sequence number to word 15, so the host has it without a further launch or copy.
template <typename Op, bool INCLUSIVE, int ITEMS = 4>
__global__ void __launch_bounds__(1024) scan_single_block_kernel(const typename Op::T *__restrict__ in, typename Op::T *__restrict__ out,
                                                                 int64_t n, typename Op::T *__restrict__ total_out,
                                                                 volatile unsigned long long *mailbox = nullptr,
                                                                 unsigned long long mailbox_seq = 0) {
  using T = typename Op::T;
  __shared__ T wt[1024 / 32 + 1];
  __shared__ T carry_s;
  if (threadIdx.x == 0) carry_s = Op::id();
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024 * ITEMS) {
    T v[ITEMS];
    T acc = Op::id();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int64_t i = base + (int64_t)threadIdx.x * ITEMS + j;
      v[j] = i < n ? in[i] : Op::id();
      acc = Op::op(acc, v[j]);
    }
    T run = Op::op(carry_s, block_exclusive<Op, 1024>(acc, wt));
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int64_t i = base + (int64_t)threadIdx.x * ITEMS + j;
      const T nxt = Op::op(run, v[j]);
      if (i < n) out[i] = INCLUSIVE ? nxt : run;
      run = nxt;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s = Op::op(carry_s, wt[1024 / 32]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (total_out) *total_out = carry_s;
    if (mailbox) {
      mailbox[0] = (unsigned long long)carry_s;
      __threadfence_system();
      mailbox[15] = mailbox_seq;
    }
  }
}

// out may alias in.  d_total (optional, device) receives the grand total.
// mailbox/mailbox_seq: when given, the kernel that computes the total also posts it to the host mailbox and
// *posted is set; otherwise the caller fetches d_total itself.
template <typename Op, bool INCLUSIVE>
int device_scan(const typename Op::T *in, typename Op::T *out, int64_t n, typename Op::T *d_total, cudaStream_t s,
                unsigned long long *mailbox = nullptr, unsigned long long mailbox_seq = 0, bool *posted = nullptr) {
  using T = typename Op::T;
  if (posted) *posted = false;
  if (n <= 0) {
    if (d_total) PB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(T), s));
    return PBGPU_OK;
  }
  if (n <= 8192) {
    PB_LAUNCH((scan_single_block_kernel<Op, INCLUSIVE, 4>), 1, 1024, 0, s, in, out, n, d_total, mailbox, mailbox_seq);
    PB_CHECK_LAUNCH();
    if (posted) *posted = mailbox != nullptr;
    return PBGPU_OK;
  }

  const int64_t nblk = cdiv(n, kScanTile);
  Scratch sc(s);
  T *partial = nullptr;
  PB_TRY(sc.get(&partial, (size_t)nblk));
  PB_LAUNCH(scan_reduce_kernel<Op>, (unsigned)nblk, kScanThreads, 0, s, in, n, partial);
  // the total is known after the second launch: posted from there, the host has it while the third still runs
  // (a single block over all n entries was tried for n <= 64K: 48 us at 39K entries against 13 us for these three)
  PB_LAUNCH(scan_partials_kernel<Op>, 1, 1024, 0, s, partial, nblk, d_total, mailbox, mailbox_seq);
  if (posted) *posted = mailbox != nullptr;
  PB_LAUNCH((scan_final_kernel<Op, INCLUSIVE>), (unsigned)nblk, kScanThreads, 0, s, in, out, n, partial);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

}  // namespace pbgpu
