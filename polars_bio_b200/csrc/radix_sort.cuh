// radix_sort.cuh -- hand-written stable LSD radix sort of (uint64 key, uint64 value) pairs.
//
// One pass = 8 key bits: per-tile digit histogram (+ per-digit totals by atomics) -> offsets kernel (one
// warp per digit: base = sum of the smaller digits' totals, then a running scan along its [digit][tile] row;
// replaces a device-wide scan that was 35 us per pass at 1M rows) -> stable scatter.  The scatter ranks keys inside a tile with
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

#ifndef PBGPU_RS_THREADS
#define PBGPU_RS_THREADS 512
#endif
constexpr int kRsThreads = PBGPU_RS_THREADS;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 8;  // per thread
constexpr int kRsTile = kRsThreads * kRsItems;
constexpr int kRsRadix = 256;

__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                             uint32_t *__restrict__ hist /*[256][nblk]*/, int64_t nblk,
                                                             uint32_t *__restrict__ digit_totals /*[256], zeroed*/) {
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
  for (int i = threadIdx.x; i < kRsRadix; i += kRsThreads) {
    hist[(int64_t)i * nblk + blockIdx.x] = h[i];
    if (h[i]) atomicAdd(digit_totals + i, h[i]);
  }
}

// one warp per digit: turn its row of per-tile counts into global scatter offsets (in place)
__global__ void __launch_bounds__(256) rs_offsets_kernel(uint32_t *__restrict__ hist, int64_t nblk,
                                                         const uint32_t *__restrict__ digit_totals) {
  const int lane = threadIdx.x & 31;
  const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (d >= kRsRadix) return;
  uint32_t base = 0;
  for (int k = lane; k < d; k += 32) base += digit_totals[k];
#pragma unroll
  for (int o = 16; o; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
  uint32_t *row = hist + (int64_t)d * nblk;
  uint32_t run = base;
  for (int64_t b0 = 0; b0 < nblk; b0 += 32 * 4) {
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int64_t b = b0 + j * 32 + lane; v[j] = b < nblk ? row[b] : 0u; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t incl = v[j];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      const int64_t b = b0 + j * 32 + lane;
      if (b < nblk) row[b] = run + incl - v[j];
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
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

// ---- single-kernel pass: decoupled look-back ("onesweep") with shared-memory staged stores -----------------------
// The digit totals of EVERY pass are permutation-invariant, so one kernel histograms them all up front
// (rs_hist_all_kernel).  A pass is then ONE kernel: a tile takes a ticket (tiles start in ticket order, so a tile only
// ever waits on tiles that are already running), ranks its keys exactly like rs_scatter_kernel, publishes its per-digit
// counts as  count | AGGREGATE  in status[tile][digit], walks back over its predecessors -- a window of kLbWindow
// status words in flight at a time -- adding their words until it meets an  inclusive-prefix | PREFIX  word, and
// publishes its own inclusive prefix.  Flag and count share one 32-bit word (2 + 30 bits, hence n < 2^30 on this
// path), so no fence is needed between them.
// Stores: a thread-per-key scatter writes 8 bytes to 32 different lines per warp instruction, and the L1TEX t-stage
// replays a divergent store once per line (~2 cycles each): 2 x n such stores per pass were the whole cost of a pass
// (r01n: 30 us at 1M rows).  Here the tile is first sorted by digit INSIDE shared memory (key at its local rank), then
// written out position by position: consecutive positions of one digit go to consecutive global addresses, so a warp
// store covers a few runs instead of 32 lines.  Values follow through the same staging buffer.
// HBM traffic per pass: 16 B read + 16 B write per element (+ 8 B once for the histogram).
constexpr uint32_t kLbAgg = 1u << 30, kLbPre = 2u << 30, kLbMask = (1u << 30) - 1u;
constexpr int kRsMaxPasses = 8;
#ifndef PBGPU_LB_WINDOW
#define PBGPU_LB_WINDOW 8
#endif
constexpr int kLbWindow = PBGPU_LB_WINDOW;  // predecessor status words a look-back step keeps in flight per digit (r2r A/B on config 3: 8 -> build 7.07 ms, 16 -> 7.29, 32 -> 8.03: the walk is not what bounds a pass)
// resident blocks per SM the single-kernel passes are compiled for (register cap = 65536 / (512 * OCC)): both variants
// are built, PBGPU_RS_OCC=2|3 picks at run time (A/B; default below)
// PBGPU_MATCH=hw at compile time (-DPBGPU_MATCH_HW) keeps the MATCH instruction (A/B builds)
__device__ __forceinline__ unsigned rs_match(unsigned d) {
#ifdef PBGPU_MATCH_HW
  return __match_any_sync(0xffffffffu, d);
#else
  return match9(d);
#endif
}
__device__ __forceinline__ unsigned rs_match8(unsigned d) {
#ifdef PBGPU_MATCH_HW
  return __match_any_sync(0xffffffffu, d);
#else
  return match8(d);
#endif
}
static inline int rs_occ() {
  static int v = [] { const char *e = getenv("PBGPU_RS_OCC"); return (e && e[0] == '2') ? 2 : 3; }();
  return v;
}

__global__ void __launch_bounds__(kRsThreads) rs_hist_all_kernel(const uint64_t *__restrict__ keys, int64_t n, int passes,
                                                                 uint32_t *__restrict__ digit_totals /*[passes][256], zeroed*/) {
  __shared__ uint32_t h[kRsMaxPasses][kRsRadix];
  for (int i = threadIdx.x; i < passes * kRsRadix; i += kRsThreads) (&h[0][0])[i] = 0;
  __syncthreads();
  constexpr int U = 4;  // loads in flight per thread (a one-load loop was latency-bound: 12 us at 1M keys)
  const int64_t stride = (int64_t)gridDim.x * kRsThreads;
  for (int64_t i0 = (int64_t)blockIdx.x * kRsThreads + threadIdx.x; i0 < n; i0 += stride * U) {
    uint64_t k[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t i = i0 + u * stride; k[u] = i < n ? keys[i] : 0; }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < n)
        for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(k[u] >> (8 * p)) & 0xff], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kRsRadix; i += kRsThreads) {
    const uint32_t v = (&h[0][0])[i];
    if (v) atomicAdd(digit_totals + i, v);
  }
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// V: value type moved along with the 64-bit key (uint64_t: end | row of the generic build; uint32_t: the end column of the
// global-key build, where the row id rides in the low half of the key).
// Register diet (round 2; ncu r2e: 64 registers -> 2 blocks of 512 per SM, warps 49 % active, short-scoreboard + barrier
// stalls, DRAM 35 % busy): the eight MATCH instructions of a thread are issued back to back before the serial counter
// chain that consumes them, and the values are loaded only after the keys have left for global memory, so keys and
// values are never live together.
// Shared memory of one pass (declared once in the kernel; the tile body below is instantiated twice)
struct RsShared {
  uint64_t stage[kRsTile];            // the tile in digit order: keys, then values
  uint16_t wcnt[kRsWarps][kRsRadix];  // per-warp digit counters (<= 256 keys per warp) -> exclusive warp offsets
  uint32_t dbase[kRsRadix];           // global position of local position 0 of digit d: dst = dbase[d] + local
  uint32_t toff[kRsRadix];            // first local position of digit d
  uint32_t wt[kRsThreads / 32 + 1];
  uint32_t tile_s;
};
// FULL: the tile holds kRsTile keys -- every bounds test and the "no element" bit of the digit match fold away (all tiles
// but the last one; the kernel branches once per block; r2x A/B: build 6.78 -> 6.74 ms on config 3).  W: predecessor status
// words a look-back step keeps in flight per digit (r2r A/B on config 3: 8 -> build 7.07 ms, 16 -> 7.29, 32 -> 8.03; r2x:
// W = 32 for sorts whose tiles are all resident at once does not help either -- 1 M rows: 0.180 vs 0.186 ms per build --
// a pass there is bound by the per-block instruction chain, not by the walk).
template <typename V, bool FULL, int W>
__device__ __forceinline__ void rs_onesweep_tile(RsShared &sm, const uint64_t *__restrict__ keys_in, const V *__restrict__ vals_in,
                                                 uint64_t *__restrict__ keys_out, V *__restrict__ vals_out, int shift,
                                                 const uint32_t *__restrict__ digit_totals, uint32_t *status, uint32_t tile,
                                                 int64_t tbase, int tile_n_rt) {
  const int tile_n = FULL ? kRsTile : tile_n_rt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp w owns the contiguous slice [tbase + w*256, +256): round r covers 32 consecutive keys
  const int wofs = warp * (32 * kRsItems);
  uint64_t k[kRsItems];
  uint32_t q[kRsItems];  // rank inside (warp, digit), later the local position in the tile
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int li = wofs + r * 32 + lane;
    k[r] = (FULL || li < tile_n) ? keys_in[tbase + li] : 0;
  }
  {
    unsigned peers[kRsItems];
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
      if (FULL) peers[r] = rs_match8((unsigned)((k[r] >> shift) & 0xff));
      else {
        const bool ok = wofs + r * 32 + lane < tile_n;
        const unsigned d = ok ? (unsigned)((k[r] >> shift) & 0xff) : 0x100u;  // 0x100: out-of-range lanes group together
        peers[r] = rs_match(d);
      }
    }
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
      const bool ok = FULL || wofs + r * 32 + lane < tile_n;
      const unsigned d = (unsigned)((k[r] >> shift) & 0xff);
      const int leader = __ffs(peers[r]) - 1;
      uint32_t old = 0;
      if (ok && lane == leader) {
        old = sm.wcnt[warp][d];
        sm.wcnt[warp][d] = (uint16_t)(old + __popc(peers[r]));
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      q[r] = old + __popc(peers[r] & lt);
      __syncwarp();
    }
  }
  __syncthreads();
  // per digit: exclusive prefix over warps; run = the tile's count of digit d
  uint32_t run = 0;
  if (threadIdx.x < kRsRadix) {
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      const uint32_t t = sm.wcnt[w][threadIdx.x];
      sm.wcnt[w][threadIdx.x] = (uint16_t)run;
      run += t;
    }
  }
  // first local position of every digit (exclusive scan of the tile's counts): everything the staging needs.  The global
  // bases (which wait for the predecessors) are only needed by the copy-out, so the keys go to shared memory first.
  const uint32_t lbase = block_exclusive<SumU32, kRsThreads>(run, sm.wt);
  if (threadIdx.x < kRsRadix) sm.toff[threadIdx.x] = lbase;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    if (FULL || wofs + r * 32 + lane < tile_n) {
      const unsigned d = (unsigned)((k[r] >> shift) & 0xff);
      q[r] = sm.toff[d] + sm.wcnt[warp][d] + q[r];
      sm.stage[q[r]] = k[r];
    }
  }
  const uint32_t gbase = block_exclusive<SumU32, kRsThreads>(threadIdx.x < kRsRadix ? digit_totals[threadIdx.x] : 0u, sm.wt);  // (barriers inside)
  if (threadIdx.x < kRsRadix) {
    const int d = threadIdx.x;
    uint32_t *mine = status + (size_t)tile * kRsRadix + d;
    uint32_t excl = 0;
    if (tile == 0) st_volatile_u32(mine, run | kLbPre);
    else {
      st_volatile_u32(mine, run | kLbAgg);
      int64_t t = (int64_t)tile - 1;
      bool done = false;
      while (!done) {  // tile 0 always ends the walk with a PREFIX word; positions below it read as an empty PREFIX
        uint32_t w[W];
#pragma unroll
        for (int j = 0; j < W; ++j) w[j] = (t - j >= 0) ? ld_volatile_u32(status + (size_t)(t - j) * kRsRadix + d) : kLbPre;
#pragma unroll
        for (int j = 0; j < W; ++j) {
          if (!done) {
            while ((w[j] >> 30) == 0u) w[j] = ld_volatile_u32(status + (size_t)(t - j) * kRsRadix + d);  // not published yet
            excl += w[j] & kLbMask;
            done = (w[j] & kLbPre) != 0u;
          }
        }
        t -= W;
      }
      st_volatile_u32(mine, (excl + run) | kLbPre);
    }
    sm.dbase[d] = gbase + excl - lbase;  // may wrap below zero: dbase[d] + local position is exact mod 2^32
  }
  // values: loaded while the look-back is under way (the warps without a digit have nothing else to do)
  V v[kRsItems];
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int li = wofs + r * 32 + lane;
    v[r] = (FULL || li < tile_n) ? vals_in[tbase + li] : V(0);
  }
  __syncthreads();
  uint32_t dst[kRsItems];
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int p = r * kRsThreads + threadIdx.x;
    if (FULL || p < tile_n) {
      const uint64_t kk = sm.stage[p];
      dst[r] = sm.dbase[(unsigned)((kk >> shift) & 0xff)] + (uint32_t)p;
      keys_out[dst[r]] = kk;
    }
  }
  __syncthreads();
  V *vstage = reinterpret_cast<V *>(sm.stage);
#pragma unroll
  for (int r = 0; r < kRsItems; ++r)
    if (FULL || wofs + r * 32 + lane < tile_n) vstage[q[r]] = v[r];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int p = r * kRsThreads + threadIdx.x;
    if (FULL || p < tile_n) vals_out[dst[r]] = vstage[p];
  }
}

template <typename V, int OCC, int W>
__global__ void __launch_bounds__(kRsThreads, OCC * (512 / kRsThreads)) rs_onesweep_kernel(const uint64_t *__restrict__ keys_in,
                                                                    const V *__restrict__ vals_in,
                                                                    uint64_t *__restrict__ keys_out,
                                                                    V *__restrict__ vals_out, int64_t n, int shift,
                                                                    const uint32_t *__restrict__ digit_totals /*[256] of this pass*/,
                                                                    uint32_t *status /*[nblk][256], zeroed*/,
                                                                    uint32_t *ticket /*zeroed*/) {
  __shared__ RsShared sm;
  if (threadIdx.x == 0) sm.tile_s = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < kRsWarps * kRsRadix / 2; i += kRsThreads) ((uint32_t *)&sm.wcnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = sm.tile_s;
  const int64_t tbase = (int64_t)tile * kRsTile;
  const int tile_n = (int)((n - tbase) < (int64_t)kRsTile ? (n - tbase) : (int64_t)kRsTile);
#ifdef PBGPU_RS_NOFULL
  if (false) {}
#else
  if (tile_n == kRsTile) rs_onesweep_tile<V, true, W>(sm, keys_in, vals_in, keys_out, vals_out, shift, digit_totals, status, tile, tbase, tile_n);
#endif
  else rs_onesweep_tile<V, false, W>(sm, keys_in, vals_in, keys_out, vals_out, shift, digit_totals, status, tile, tbase, tile_n);
}

// PBGPU_SORT=3k: the three-kernels-per-pass sort (the first implementation; kept for A/B runs and for n >= 2^30)
static inline bool rs_three_kernel() {
  static bool v = [] { const char *e = getenv("PBGPU_SORT"); return e && !strcmp(e, "3k"); }();
  return v;
}

template <typename V>
struct SortedKV { uint64_t *keys; V *vals; };
using SortedPairs = SortedKV<uint64_t>;

// Stable LSD sort of n (key, value) pairs over the listed 8-bit digit positions (digit p = key bits [8p, 8p+8)), lowest
// first.  Ping-pongs between (keys, vals) and (k2, v2); *out says where the sorted pairs ended up (no copy back).
// totals_by_pos: digit totals [kRsMaxPasses][256] indexed by digit position, when the caller already has them
// (the build's prep kernel); NULL = histogram here.
template <typename V>
inline int radix_sort_digits(uint64_t *keys, V *vals, uint64_t *k2, V *v2, int64_t n, const int *digit_pos, int npass,
                             const uint32_t *totals_by_pos, cudaStream_t s, SortedKV<V> *out) {
  out->keys = keys;
  out->vals = vals;
  if (n <= 1 || npass <= 0) return PBGPU_OK;
  const int64_t nblk = cdiv(n, kRsTile);
  int max_pos = 0;
  for (int p = 0; p < npass; ++p) {
    if (digit_pos[p] < 0 || digit_pos[p] >= kRsMaxPasses) return set_error(PBGPU_EINVAL, "bad digit position %d", digit_pos[p]);
    if (digit_pos[p] > max_pos) max_pos = digit_pos[p];
  }
  Scratch sc(s);
  uint64_t *ki = keys, *ko = k2;
  V *vi = vals, *vo = v2;
  if (n < (int64_t)kLbMask && (!rs_three_kernel() || sizeof(V) != 8)) {
    uint32_t *work = nullptr;  // [kRsMaxPasses][256] digit totals (when not supplied) | tickets (256) | [npass][nblk][256] status
    const size_t tot_w = totals_by_pos ? 0 : (size_t)kRsMaxPasses * kRsRadix, tick_w = kRsRadix, stat_w = (size_t)npass * (size_t)nblk * kRsRadix;
    PB_TRY(sc.get(&work, tot_w + tick_w + stat_w));
    PB_CUDA(cudaMemsetAsync(work, 0, sizeof(uint32_t) * (tot_w + tick_w + stat_w), s));
    uint32_t *tickets = work + tot_w, *status = work + tot_w + tick_w;
    if (!totals_by_pos) {
      int64_t hgrid = cdiv(n, (int64_t)kRsThreads * 4);
      if (hgrid > kSMs * 4) hgrid = kSMs * 4;
      PB_LAUNCH(rs_hist_all_kernel, (unsigned)hgrid, kRsThreads, 0, s, keys, n, max_pos + 1, work);
      totals_by_pos = work;
    }
    for (int p = 0; p < npass; ++p) {
      const uint32_t *tot = totals_by_pos + (size_t)digit_pos[p] * kRsRadix;
      uint32_t *stat = status + (size_t)p * (size_t)nblk * kRsRadix;
      if (rs_occ() == 3)
        PB_LAUNCH((rs_onesweep_kernel<V, 3, kLbWindow>), (unsigned)nblk, kRsThreads, 0, s, ki, vi, ko, vo, n, 8 * digit_pos[p], tot, stat, tickets + p);
      else
        PB_LAUNCH((rs_onesweep_kernel<V, 2, kLbWindow>), (unsigned)nblk, kRsThreads, 0, s, ki, vi, ko, vo, n, 8 * digit_pos[p], tot, stat, tickets + p);
      uint64_t *t = ki; ki = ko; ko = t;
      V *u = vi; vi = vo; vo = u;
    }
    PB_CHECK_LAUNCH();
  } else {
    if (sizeof(V) != 8) return set_error(PBGPU_ERANGE, "table too large for the single-kernel radix pass");
    uint32_t *hist = nullptr, *totals = nullptr;
    PB_TRY(sc.get(&hist, (size_t)(nblk * kRsRadix)));
    PB_TRY(sc.get(&totals, (size_t)npass * kRsRadix));
    PB_CUDA(cudaMemsetAsync(totals, 0, sizeof(uint32_t) * (size_t)npass * kRsRadix, s));
    for (int p = 0; p < npass; ++p) {
      uint32_t *tot = totals + p * kRsRadix;
      const int shift = 8 * digit_pos[p];
      PB_LAUNCH(rs_hist_kernel, (unsigned)nblk, kRsThreads, 0, s, ki, n, shift, hist, nblk, tot);
      PB_LAUNCH(rs_offsets_kernel, kRsRadix / 8, 256, 0, s, hist, nblk, tot);
      PB_LAUNCH(rs_scatter_kernel, (unsigned)nblk, kRsThreads, 0, s, ki, (const uint64_t *)vi, ko, (uint64_t *)vo, n, shift, hist, nblk);
      PB_CHECK_LAUNCH();
      uint64_t *t = ki; ki = ko; ko = t;
      V *u = vi; vi = vo; vo = u;
    }
  }
  out->keys = ki;
  out->vals = vi;
  return PBGPU_OK;
}

// Sorts n pairs by the low `bits` bits of the key.  keys/vals are overwritten with the sorted result.  n < 2^32.
inline int radix_sort_pairs(uint64_t *keys, uint64_t *vals, int64_t n, int bits, cudaStream_t s) {
  if (n <= 1 || bits <= 0) return PBGPU_OK;
  if (bits > 8 * kRsMaxPasses) bits = 8 * kRsMaxPasses;
  Scratch sc(s);
  uint64_t *k2 = nullptr, *v2 = nullptr;
  PB_TRY(sc.get(&k2, (size_t)n));
  PB_TRY(sc.get(&v2, (size_t)n));
  int pos[kRsMaxPasses];
  const int npass = (bits + 7) / 8;
  for (int p = 0; p < npass; ++p) pos[p] = p;
  SortedPairs r;
  PB_TRY(radix_sort_digits(keys, vals, k2, v2, n, pos, npass, nullptr, s, &r));
  if (r.keys != keys) {
    PB_CUDA(cudaMemcpyAsync(keys, r.keys, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    PB_CUDA(cudaMemcpyAsync(vals, r.vals, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  }
  return PBGPU_OK;
}

}  // namespace pbgpu
