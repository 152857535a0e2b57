// bins.cuh -- probe-side coordinate partition for indexes that do not fit the L2 (BASELINE config 3: 90 M indexed rows,
// 3 GB rank directory).
//
// With probes in arbitrary order every directory lookup is a random 32-byte DRAM sector: measured 47.7 G sectors/s on
// B200 (scripts/randsector_probe.cu: 2.1 ms per 100 M lookups, whatever the number of loads in flight) -- the ceiling
// of round 1's count kernel (4.3 ms per 100 M probes, two records per 150-bp read).  The same lookups grouped into
// coarse coordinate bins (random INSIDE a bin of a few MB of directory) run at 152 G/s: the DRAM pages of a bin stay
// open and its sectors stay in the L2 while the grid sweeps over it.  So the probes are partitioned ONCE by the top
// bits of their global-axis start coordinate (stable: one radix pass in the style of radix_sort.cuh's onesweep kernel,
// ticket order + decoupled look-back, tile re-ordered in shared memory so every (tile, bin) run leaves as consecutive
// 16-byte records), and count / pass 1 / pass 2 run over the partitioned records.  Pairs are emitted in bin order
// (their order is free: the reference's own tests sort before comparing); counts go back to row order through the
// saved destination of every row (a gather whose sources are the same runs, so it reads whole sectors).
//
// Record (16 bytes, one LDG.128 per probe in every later kernel):
//   proper probe      : x = global-axis start (clamped into the contig's slice), y = global-axis end, w = contig code
//   w = -1            : no indexed rows on its contig / null key (count 0, nothing to look up); bin 0
//   w = -2 - contig   : empty or inverted probe interval: x = RAW start, y = RAW end (bare predicate over the candidate
//                       window of the generic kernels); bin 0 (rare; their lookups are not worth a bin)
//   z = the id reported for the probe: its row, or the caller's id column (multi-GPU: global row ids, so no translation
//       pass over the pair buffer is needed)
#pragma once
#include "common.cuh"
#include "index.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "sweep.cuh"

namespace pbgpu {

#ifndef PBGPU_BIN_THREADS
#define PBGPU_BIN_THREADS 512
#endif
constexpr int kBinThreads = PBGPU_BIN_THREADS, kBinItems = 8, kBinTile = kBinThreads * kBinItems, kBinRadix = 256, kBinWarps = kBinThreads / 32;

__device__ __forceinline__ int4 make_probe_rec(const IndexView &ix, int32_t c, int32_t s, int32_t e, uint32_t row, bool strict) {
  if (c < 0 || c >= ix.n_contigs) return make_int4(0, 0, (int)row, -1);
  const ContigMap cm = ix.cmap[c];
  if (!cm.has) return make_int4(0, 0, (int)row, -1);
  long long ls = s, le = e;  // clamp into the contig's slice: order against every indexed coordinate is preserved
  ls = ls < cm.lo_m1 ? cm.lo_m1 : (ls > cm.hi_p1 ? cm.hi_p1 : ls);
  le = le < cm.lo_m1 ? cm.lo_m1 : (le > cm.hi_p1 ? cm.hi_p1 : le);
  const uint32_t g_s = cm.off + (uint32_t)(ls - cm.lo_m1), g_e = cm.off + (uint32_t)(le - cm.lo_m1);
  const bool proper = strict ? (s < e) : (s <= e);
  if (!proper) return make_int4(s, e, (int)row, -2 - c);
  return make_int4((int)g_s, (int)g_e, (int)row, c);
}
__device__ __forceinline__ unsigned rec_bin(const int4 r, int bin_shift) { return r.w >= 0 ? ((uint32_t)r.x >> bin_shift) : 0u; }
// bin of a probe: top bits of its global start; probes that look nothing up go to bin 0
__device__ __forceinline__ uint32_t probe_bin(const IndexView &ix, int32_t c, int32_t s, int bin_shift) {
  if (c < 0 || c >= ix.n_contigs) return 0u;
  const ContigMap cm = ix.cmap[c];
  if (!cm.has) return 0u;
  long long ls = s;
  ls = ls < cm.lo_m1 ? cm.lo_m1 : (ls > cm.hi_p1 ? cm.hi_p1 : ls);
  return (cm.off + (uint32_t)(ls - cm.lo_m1)) >> bin_shift;
}

__global__ void __launch_bounds__(512) bin_hist_kernel(IndexView ix, const int32_t *__restrict__ pc, const int32_t *__restrict__ ps,
                                                       const int32_t *__restrict__ pe, int64_t n, int bin_shift, int strict,
                                                       uint32_t *__restrict__ totals /*[256], zeroed*/) {
  __shared__ uint32_t h[kBinRadix];
  for (int i = threadIdx.x; i < kBinRadix; i += 512) h[i] = 0;
  __syncthreads();
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * 512;
  for (int64_t i0 = (int64_t)blockIdx.x * 512 + threadIdx.x; i0 < n; i0 += stride * U) {
    int32_t c[U], s[U], e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int64_t i = i0 + u * stride; const bool ok = i < n; c[u] = ok ? pc[i] : -1; s[u] = ok ? ps[i] : 0; e[u] = ok ? pe[i] : 0; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= n) continue;
      const bool proper = strict ? (s[u] < e[u]) : (s[u] <= e[u]);
      const uint32_t d = proper ? probe_bin(ix, c[u], s[u], bin_shift) : 0u;
      // neighbouring lanes often share a bin only by chance (random probes): plain shared-memory atomics
      atomicAdd(&h[d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBinRadix; i += 512) if (h[i]) atomicAdd(totals + i, h[i]);
}

// One stable radix pass over the probes (see radix_sort.cuh: rs_onesweep_kernel for the scheme).  Dynamic shared memory:
// the tile's records in bin order (64 KB).  WRITE_POS: also pos[row] = destination of the row (count_overlaps un-binning).
template <bool WRITE_POS, int OCC>
__global__ void __launch_bounds__(kBinThreads, OCC * (512 / kBinThreads)) bin_partition_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                       const int32_t *__restrict__ ps, const int32_t *__restrict__ pe,
                                                                       const uint32_t *__restrict__ ids /*NULL: the row*/,
                                                                       int64_t n, int bin_shift, int strict,
                                                                       const uint32_t *__restrict__ totals /*[256]*/,
                                                                       uint32_t *status /*[tiles][256], zeroed*/, uint32_t *ticket /*zeroed*/,
                                                                       int4 *__restrict__ recs, uint32_t *__restrict__ pos, int bulk_store) {
  extern __shared__ __align__(16) unsigned char bin_smem[];
  int4 *stage = reinterpret_cast<int4 *>(bin_smem);
  __shared__ uint16_t wcnt[kBinWarps][kBinRadix];
  __shared__ uint32_t dbase[kBinRadix];
  __shared__ uint32_t toff[kBinRadix];
  __shared__ uint32_t wt[kBinThreads / 32 + 1];
  __shared__ uint32_t tile_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < kBinWarps * kBinRadix / 2; i += kBinThreads) ((uint32_t *)&wcnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = tile_s;
  const int64_t tbase = (int64_t)tile * kBinTile;
  const int tile_n = (int)((n - tbase) < (int64_t)kBinTile ? (n - tbase) : (int64_t)kBinTile);
  const int wofs = warp * (32 * kBinItems);
  int4 rec[kBinItems];
  uint32_t q[kBinItems];
  const unsigned lt = lanemask_lt();
  {
    int32_t c[kBinItems], s[kBinItems], e[kBinItems];
#pragma unroll
    for (int r = 0; r < kBinItems; ++r) {
      const int li = wofs + r * 32 + lane;
      const bool ok = li < tile_n;
      c[r] = ok ? pc[tbase + li] : -1; s[r] = ok ? ps[tbase + li] : 0; e[r] = ok ? pe[tbase + li] : 0;
    }
#pragma unroll
    for (int r = 0; r < kBinItems; ++r) {
      const int li = wofs + r * 32 + lane;
      const uint32_t id = (ids && li < tile_n) ? ids[tbase + li] : (uint32_t)(tbase + li);
      rec[r] = make_probe_rec(ix, c[r], s[r], e[r], id, strict != 0);
    }
  }
  // bins + the eight MATCH instructions first, then the serial counter chain that consumes them (radix_sort.cuh)
  unsigned dg[kBinItems];
  {
    unsigned peers[kBinItems];
#pragma unroll
    for (int r = 0; r < kBinItems; ++r) {
      const bool ok = wofs + r * 32 + lane < tile_n;
      dg[r] = ok ? rec_bin(rec[r], bin_shift) : 0x100u;
      peers[r] = rs_match(dg[r]);
    }
#pragma unroll
    for (int r = 0; r < kBinItems; ++r) {
      const bool ok = dg[r] != 0x100u;
      const int leader = __ffs(peers[r]) - 1;
      uint32_t old = 0;
      if (ok && lane == leader) {
        old = wcnt[warp][dg[r]];
        wcnt[warp][dg[r]] = (uint16_t)(old + __popc(peers[r]));
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      q[r] = old + __popc(peers[r] & lt);
      __syncwarp();
    }
  }
  __syncthreads();
  uint32_t run = 0;
  if (threadIdx.x < kBinRadix) {
#pragma unroll
    for (int w = 0; w < kBinWarps; ++w) {
      const uint32_t t = wcnt[w][threadIdx.x];
      wcnt[w][threadIdx.x] = (uint16_t)run;
      run += t;
    }
  }
  // local bin offsets are all the staging needs: the records leave the registers before the look-back starts
  const uint32_t lbase = block_exclusive<SumU32, kBinThreads>(run, wt);
  if (threadIdx.x < kBinRadix) toff[threadIdx.x] = lbase;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kBinItems; ++r) {
    if (dg[r] != 0x100u) {
      q[r] = toff[dg[r]] + wcnt[warp][dg[r]] + q[r];
      stage[q[r]] = rec[r];
    }
  }
  // the staged records are read by the bulk-copy (async proxy) engine below: order this thread's shared-memory writes
  // before it; the barriers inside the next block scan make them visible to the issuing threads
  if (bulk_store) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t gbase = block_exclusive<SumU32, kBinThreads>(threadIdx.x < kBinRadix ? totals[threadIdx.x] : 0u, wt);
  if (threadIdx.x < kBinRadix) {
    const int d = threadIdx.x;
    uint32_t *mine = status + (size_t)tile * kBinRadix + d;
    uint32_t excl = 0;
    if (tile == 0) st_volatile_u32(mine, run | kLbPre);
    else {
      st_volatile_u32(mine, run | kLbAgg);
      int64_t t = (int64_t)tile - 1;
      bool done = false;
      while (!done) {
        uint32_t w[kLbWindow];
#pragma unroll
        for (int j = 0; j < kLbWindow; ++j) w[j] = (t - j >= 0) ? ld_volatile_u32(status + (size_t)(t - j) * kBinRadix + d) : kLbPre;
#pragma unroll
        for (int j = 0; j < kLbWindow; ++j) {
          if (!done) {
            while ((w[j] >> 30) == 0u) w[j] = ld_volatile_u32(status + (size_t)(t - j) * kBinRadix + d);
            excl += w[j] & kLbMask;
            done = (w[j] & kLbPre) != 0u;
          }
        }
        t -= kLbWindow;
      }
      st_volatile_u32(mine, (excl + run) | kLbPre);
    }
    dbase[d] = gbase + excl - lbase;  // may wrap below zero: dbase[d] + local position is exact mod 2^32
  }
  __syncthreads();
  if (WRITE_POS) {
#pragma unroll
    for (int r = 0; r < kBinItems; ++r)
      if (dg[r] != 0x100u) pos[tbase + wofs + r * 32 + lane] = dbase[dg[r]] + q[r];
  }
  if (bulk_store) {
    // TMA bulk stores (cp.async.bulk shared -> global; SASS: UBLKCP): the records of (tile, bin d) are ONE run in shared
    // memory and ONE run in global memory, both 16-byte aligned (16-byte records): thread d hands its bin's run to the
    // copy engine as a whole instead of the block copying record by record through registers.
    if (threadIdx.x < kBinRadix && run) {
      const uint32_t first = toff[threadIdx.x];
      const uint64_t gdst = (uint64_t)(uintptr_t)(recs + (dbase[threadIdx.x] + first));
      const uint32_t ssrc = (uint32_t)__cvta_generic_to_shared(stage + first);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(run * 16u) : "memory");
    }
    if (threadIdx.x < kBinRadix) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the block's shared memory goes away when it exits
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < kBinItems; ++r) {
    const int p = r * kBinThreads + threadIdx.x;
    if (p < tile_n) {
      const int4 v = stage[p];
      recs[dbase[rec_bin(v, bin_shift)] + (uint32_t)p] = v;
    }
  }
}

// four consecutive (end, row) entries, 32-byte aligned, in one request
__device__ __forceinline__ void ld_er4(const uint2 *__restrict__ p, uint32_t (&w)[8]) {
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
      : "l"(p));
}

// count of one partitioned probe; hi_out as fast_count()
template <bool STRICT>
__device__ __forceinline__ uint32_t binned_count(const IndexView &ix, const int4 r, uint32_t &hi_out) {
  hi_out = 0;
  if (r.w >= 0) {
    uint32_t hi, re;
    jdir_ranks<STRICT>(ix, (uint32_t)r.x, (uint32_t)r.y, hi, re);
    hi_out = hi;
    return hi - re;
  }
  if (r.w == -1) return 0;
  hi_out = kGenericProbe;  // empty / inverted probe: bare predicate over its raw coordinates
  return probe_count<STRICT>(ix, -2 - r.w, r.x, r.y);
}

template <bool STRICT, int ITEMS>
__global__ void __launch_bounds__(kSweepThreads) binned_count_kernel(IndexView ix, const int4 *__restrict__ recs, int64_t n, uint32_t *__restrict__ cnt_b) {
  const int64_t base = (int64_t)blockIdx.x * (kSweepThreads * ITEMS) + threadIdx.x;
  int4 r[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * kSweepThreads;
    r[j] = i < n ? recs[i] : make_int4(0, 0, 0, -1);
  }
  uint32_t cnt[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) { uint32_t hi; cnt[j] = binned_count<STRICT>(ix, r[j], hi); }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * kSweepThreads;
    if (i < n) cnt_b[i] = cnt[j];
  }
}

// counts back to row order: out[row] = cnt_b[pos[row]].  The sources of a tile of rows are its 256 runs in cnt_b.
template <typename OutT>
__global__ void __launch_bounds__(256) unbin_counts_kernel(const uint32_t *__restrict__ cnt_b, const uint32_t *__restrict__ pos, int64_t n,
                                                           OutT *__restrict__ out) {
  constexpr int U = 4;
  const int64_t base = (int64_t)blockIdx.x * (256 * U) + threadIdx.x;
  uint32_t p[U], v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { const int64_t i = base + u * 256; p[u] = i < n ? pos[i] : 0u; }
#pragma unroll
  for (int u = 0; u < U; ++u) { const int64_t i = base + u * 256; v[u] = i < n ? __ldg(cnt_b + p[u]) : 0u; }
#pragma unroll
  for (int u = 0; u < U; ++u) { const int64_t i = base + u * 256; if (i < n) out[i] = (OutT)v[u]; }
}

// pass 1 over partitioned probes: (count, start rank) per position + the raw total of every 256 positions (scanned by the
// caller) + the offset of every 32-position group inside its block (flat pass 2)
template <bool STRICT, int ITEMS>
__global__ void __launch_bounds__(kSweepThreads) binned_p1_kernel(IndexView ix, const int4 *__restrict__ recs, int64_t n, uint32_t *__restrict__ counts,
                                                                  uint32_t *__restrict__ his, unsigned long long *__restrict__ block_base,
                                                                  unsigned long long *__restrict__ warp_off) {
  __shared__ unsigned long long wt[ITEMS][kSweepThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * (kSweepThreads * ITEMS) + threadIdx.x;
  int4 r[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * kSweepThreads;
    r[j] = i < n ? recs[i] : make_int4(0, 0, 0, -1);
  }
  uint32_t cnt[ITEMS], hi[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) cnt[j] = binned_count<STRICT>(ix, r[j], hi[j]);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * kSweepThreads;
    if (i < n) { counts[i] = cnt[j]; his[i] = hi[j]; }
    const unsigned long long v = warp_sum_u32(i < n ? cnt[j] : 0u);
    if ((threadIdx.x & 31) == 0) wt[j][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (warp_off && threadIdx.x >= 32 && threadIdx.x < 32 + ITEMS * (kSweepThreads / 32)) {
    const int q = threadIdx.x - 32;
    const int j = q / (kSweepThreads / 32), w = q % (kSweepThreads / 32);
    unsigned long long t = 0;
    for (int k = 0; k < w; ++k) t += wt[j][k];
    const int64_t g = ((int64_t)blockIdx.x * ITEMS + j) * (kSweepThreads / 32) + w;
    if (g * 32 < n) warp_off[g] = t;
  }
  if (threadIdx.x < ITEMS) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < kSweepThreads / 32; ++w) t += wt[threadIdx.x][w];
    const int64_t blk = (int64_t)blockIdx.x * ITEMS + threadIdx.x;
    if (blk * kSweepThreads < n) block_base[blk] = t;
  }
}

// ---- pass 2 with warp-staged, coalesced pair stores (indexes with nested intervals) ---------------------------------
// The hits of a probe are the `cnt` entries below its start rank whose end reaches past the probe start: every lane walks
// down its own window, but instead of storing each pair where it belongs (32 lanes -> 32 scattered 4-byte stores per
// column: partial sectors, one L1TEX replay per line) it drops the pair into the warp's shared-memory window at its
// slot, and the warp then copies the window out with consecutive lanes on consecutive pairs.  A window longer than the
// staging buffer is produced in rounds, highest slots first: a lane finds its hits from the top of its window down, so
// its walk simply pauses between rounds.  Probes with very many hits are walked by the whole warp straight to global
// memory (ballot compaction: those stores are contiguous already).
// BINNED: probes come from partitioned records (probe id = rec.z, start = un-shifted global start); else from the columns.
constexpr int kEmitStagePairs = 384;   // per warp: 3 KB; 8 warps -> 24 KB per block
constexpr uint32_t kEmitHeavy = 192;   // hits of ONE probe from which the whole warp walks its window
template <bool STRICT, bool BINNED>
__global__ void __launch_bounds__(kSweepThreads) overlap_emit_staged_kernel(IndexView ix, const int4 *__restrict__ recs,
                                                                            const int32_t *__restrict__ pc, const int32_t *__restrict__ ps,
                                                                            const int32_t *__restrict__ pe, const uint32_t *__restrict__ ids, int64_t n,
                                                                            const uint32_t *__restrict__ counts, const uint32_t *__restrict__ his,
                                                                            const unsigned long long *__restrict__ block_base, int64_t blk0,
                                                                            uint32_t *__restrict__ out_probe, uint32_t *__restrict__ out_build) {
  __shared__ unsigned long long wt[kSweepThreads / 32 + 1];
  __shared__ uint2 stage[kSweepThreads / 32][kEmitStagePairs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t blk = blk0 + blockIdx.x;
  const int64_t i = blk * kSweepThreads + threadIdx.x;
  uint32_t cnt = 0, hi = 0, pid = (uint32_t)i;
  long long s = 0;  // probe start in the index's own coordinates (raw, or a stand-in with the same order against every indexed end)
  int32_t c = -1, e_raw = 0;
  if (i < n) {
    cnt = counts[i]; hi = his[i];
    if (BINNED) {
      const int4 r = recs[i];
      pid = (uint32_t)r.z;  // already the caller's id when an id column was given
      if (r.w >= 0) { c = r.w; const ContigMap cm = ix.cmap[c]; s = (long long)(uint32_t)r.x - (long long)cm.off + cm.lo_m1; }
      else if (r.w < -1) { c = -2 - r.w; s = r.x; e_raw = r.y; }
    } else {
      s = ps[i];
      if (ids) pid = ids[i];
    }
  }
  const unsigned long long pos = block_base[blk] - block_base[blk0] + block_exclusive<SumU64, kSweepThreads>((unsigned long long)cnt, wt);
  const bool generic = cnt && hi == kGenericProbe;
  const bool heavy = cnt >= kEmitHeavy && !generic;
  const unsigned lt = lanemask_lt();
  const long long thr = s - (STRICT ? 0 : 1);  // hit <=> indexed end > thr   (Strict: end > start; Weak: end >= start)
  // Flat fast path.  cnt = (starts before the probe end) - (ends at or before the probe start), so the window [hi-cnt, hi)
  // holds exactly cnt entries; when nothing BELOW it reaches past the probe start -- one load of the running maximum at
  // hi-cnt-1 decides -- every hit lies inside it, hence every entry of it is a hit: the probe's pairs are a plain expansion
  // of (hi, cnt), no candidate is looked at.  With shallow nesting (config 3: 10 % short indels) that is ~99.9 % of the
  // probes; a warp whose 32 probes are all of that kind writes its pairs 32 consecutive slots at a time straight from the
  // row column (the scheme of overlap_emit_flat_kernel) and skips the walk and the staging window.
#ifndef PBGPU_EMIT_NOFLAT
  {
    bool clean = false;
    const uint32_t first = hi - cnt;
    if (cnt && !generic) {
      const int32_t cc = BINNED ? c : pc[i];
      clean = first <= (uint32_t)__ldg(ix.seg + cc) || (long long)__ldg(ix.pmax + first - 1) <= thr;
    }
    if (__all_sync(0xffffffffu, cnt == 0 || clean)) {
      uint32_t incl_f = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl_f, d); if (lane >= d) incl_f += o; }
      const uint32_t total = __shfl_sync(0xffffffffu, incl_f, 31);
      const uint32_t excl_f = incl_f - cnt;
      for (uint32_t j0 = 0; j0 < total; j0 += 32) {
        const uint32_t j = j0 + lane;
        int p = 0;
#pragma unroll
        for (int step = 16; step; step >>= 1) {
          const int cand = p + step;
          const uint32_t e = __shfl_sync(0xffffffffu, excl_f, cand & 31);
          if (e <= j) p = cand;  // the last lane whose exclusive offset is <= j owns slot j (cand <= 31 always)
        }
        const uint32_t e_p = __shfl_sync(0xffffffffu, excl_f, p);
        const uint32_t f_p = __shfl_sync(0xffffffffu, first, p);
        const uint32_t id_p = __shfl_sync(0xffffffffu, pid, p);
        const unsigned long long pos_p = __shfl_sync(0xffffffffu, pos, p);
        if (j < total) {
          const unsigned long long g = pos_p + (j - e_p);
          out_probe[g] = id_p;
          out_build[g] = __ldg(ix.row + (f_p + (j - e_p)));
        }
      }
      return;
    }
  }
#endif
  // the warp's window holds the pairs of the lanes that go through the staging buffer
  const uint32_t mine = (cnt && !heavy) ? cnt : 0u;
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
  const uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t excl = incl - mine;
  // walk state: candidates j (descending) down to jlo; generic probes (empty / inverted interval) walk their bare window
  int64_t j = (int64_t)hi - 1, jlo = 0;
  if (generic) {
    const int32_t cc = BINNED ? c : pc[i];
    const int32_t ee = BINNED ? e_raw : pe[i];
    int32_t glo, ghi;
    probe_window<STRICT>(ix, ix.seg[cc], ix.seg[cc + 1], (int32_t)s, ee, glo, ghi);
    j = (int64_t)ghi - 1; jlo = glo;
  }
  uint32_t k = 0;  // hits found so far
  if (wtotal) {
    for (int64_t c0 = (int64_t)((wtotal - 1) / kEmitStagePairs) * kEmitStagePairs; c0 >= 0; c0 -= kEmitStagePairs) {
      const uint32_t lo_slot = (uint32_t)c0;
      const uint32_t hi_slot = lo_slot + kEmitStagePairs < wtotal ? lo_slot + kEmitStagePairs : wtotal;
      if (mine && excl < hi_slot && excl + mine > lo_slot) {
        while (k < cnt && j >= jlo) {
          uint32_t slot = excl + (cnt - 1 - k);  // slot of the next hit to be found
          if (slot < lo_slot) break;            // belongs to a later round
          if (generic) {
            const int32_t ev = __ldg(ix.en + j);
            const uint32_t rv = __ldg(ix.row + j);
            --j;
            if ((long long)ev > thr) { stage[warp][slot - lo_slot] = make_uint2(pid, rv); ++k; }
            continue;
          }
          // four (end, row) candidates per request: the L1TEX serves one line per lane and request whatever its width,
          // and that request rate -- not bytes -- bounds this kernel (438 M single-entry loads = 3.1 ms on config 3)
          const int64_t g0 = j & ~(int64_t)3;
          uint32_t w[8];
          ld_er4(ix.er + g0, w);
          bool paused = false;
#pragma unroll
          for (int t = 3; t >= 0; --t) {
            const int64_t jj = g0 + t;
            if (jj <= j && jj >= jlo && k < cnt && !paused) {
              slot = excl + (cnt - 1 - k);
              if (slot < lo_slot) { paused = true; j = jj; }
              else if ((long long)(int32_t)w[2 * t] > thr) { stage[warp][slot - lo_slot] = make_uint2(pid, w[2 * t + 1]); ++k; }
            }
          }
          if (paused) break;
          j = g0 - 1;
        }
      }
      __syncwarp();
      // copy-out: consecutive lanes on consecutive window slots; the owner of a slot is the first lane whose inclusive
      // offset exceeds it (lanes without staged pairs repeat their predecessor's offset and are never chosen)
      for (uint32_t t0 = lo_slot; t0 < hi_slot; t0 += 32) {
        const uint32_t slot = t0 + lane;
        int p = 0;
#pragma unroll
        for (int step = 16; step; step >>= 1) {
          const int cand = p + step;
          const uint32_t v = __shfl_sync(0xffffffffu, incl, cand - 1);
          if (v <= slot) p = cand;
        }
        const uint32_t e_o = __shfl_sync(0xffffffffu, excl, p);
        const unsigned long long pos_o = __shfl_sync(0xffffffffu, pos, p);
        if (slot < hi_slot) {
          const uint2 v = stage[warp][slot - lo_slot];
          const unsigned long long g = pos_o + (slot - e_o);
          out_probe[g] = v.x;
          out_build[g] = v.y;
        }
      }
      __syncwarp();
    }
  }
  // probes with very many hits: the whole warp walks the window, 32 candidates per step, ballot-compacted stores
  unsigned hm = __ballot_sync(0xffffffffu, heavy);
  while (hm) {
    const int src = __ffs(hm) - 1;
    hm &= hm - 1;
    const uint32_t h = __shfl_sync(0xffffffffu, hi, src), c_all = __shfl_sync(0xffffffffu, cnt, src);
    const long long tt = __shfl_sync(0xffffffffu, thr, src);
    const unsigned long long p0 = __shfl_sync(0xffffffffu, pos, src);
    const uint32_t pi = __shfl_sync(0xffffffffu, pid, src);
    uint32_t found = 0;
    for (int64_t top = (int64_t)h - 1; found < c_all && top >= 0; top -= 32) {
      const int64_t jj = top - lane;  // lane 0 = highest position
      const bool ok = jj >= 0 && (long long)__ldg(ix.en + jj) > tt;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint32_t kk = found + __popc(m & lt);
        if (kk < c_all) {
          const unsigned long long pp = p0 + (c_all - 1 - kk);
          out_probe[pp] = pi;
          out_build[pp] = __ldg(ix.row + jj);
        }
      }
      found += __popc(m);
    }
  }
}

}  // namespace pbgpu
