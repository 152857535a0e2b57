// sweep.cuh -- probe-side kernels: count (pass 1), emit (pass 2), count_overlaps, coverage,
// nearest.  One thread per probe row, coalesced int32 loads of (contig,start,end); every
// search is bounded to the probe's contig segment of the index.
//
// Predicate (docs/developers.md:549-552), never rewritten as end+1 (int32 domain):
//   Strict  hit <=> b.start <  a.end && b.end >  a.start
//   Weak    hit <=> b.start <= a.end && b.end >= a.start
#pragma once
#include "common.cuh"
#include "index.cuh"
#include "scan.cuh"

namespace pbgpu {

constexpr int kSweepThreads = 256;
constexpr int kHeavyWindow = 64;  // windows at least this long are walked by the whole warp

template <bool STRICT>
__device__ __forceinline__ bool end_hits(int32_t b_end, int32_t a_start) {
  return STRICT ? (b_end > a_start) : (b_end >= a_start);
}
template <bool STRICT>
__device__ __forceinline__ bool is_hit(int32_t as, int32_t ae, int32_t bs, int32_t be) {
  return STRICT ? (as < be && ae > bs) : (as <= be && ae >= bs);
}

// Candidate window [lo,hi) of a probe in (st,en,row) order: hi = first start past the probe
// end, lo = first position whose running max end reaches past the probe start.
template <bool STRICT>
__device__ __forceinline__ void probe_window(const IndexView &ix, int32_t seg_lo, int32_t seg_hi, int32_t s, int32_t e,
                                             int32_t &lo, int32_t &hi) {
  hi = STRICT ? lower_bound_i32(ix.st, seg_lo, seg_hi, e) : upper_bound_i32(ix.st, seg_lo, seg_hi, e);
  lo = STRICT ? upper_bound_i32(ix.pmax, seg_lo, hi, s) : lower_bound_i32(ix.pmax, seg_lo, hi, s);
}

template <bool STRICT>
__device__ __forceinline__ uint32_t probe_count(const IndexView &ix, int32_t c, int32_t s, int32_t e) {
  if (c < 0 || c >= ix.n_contigs) return 0;
  const int32_t seg_lo = ix.seg[c], seg_hi = ix.seg[c + 1];
  if (seg_lo >= seg_hi) return 0;
  const bool proper = STRICT ? (s < e) : (s <= e);
  if (proper && !ix.has_inverted && ix.en_sorted) {
    // rank identity: every indexed row ending before the probe starts also starts before it ends
    const int32_t hi = STRICT ? lower_bound_i32(ix.st, seg_lo, seg_hi, e) : upper_bound_i32(ix.st, seg_lo, seg_hi, e);
    const int32_t re = STRICT ? upper_bound_i32(ix.en_sorted, seg_lo, seg_hi, s) : lower_bound_i32(ix.en_sorted, seg_lo, seg_hi, s);
    return (uint32_t)(hi - re);
  }
  int32_t lo, hi;
  probe_window<STRICT>(ix, seg_lo, seg_hi, s, e, lo, hi);
  uint32_t n = 0;
  for (int32_t j = lo; j < hi; ++j) n += end_hits<STRICT>(__ldg(ix.en + j), s);
  return n;
}

// ---- count_overlaps: int64 count per iterated row -------------------------------------------
template <bool STRICT, typename OutT>
__global__ void __launch_bounds__(kSweepThreads) count_overlaps_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                       const int32_t *__restrict__ ps,
                                                                       const int32_t *__restrict__ pe, int64_t n,
                                                                       OutT *__restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * kSweepThreads + threadIdx.x;
  if (i >= n) return;
  counts[i] = (OutT)probe_count<STRICT>(ix, pc[i], ps[i], pe[i]);
}

// ---- overlap pass 1: uint32 count per probe + uint64 total per block -------------------------
template <bool STRICT>
__global__ void __launch_bounds__(kSweepThreads) overlap_count_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                      const int32_t *__restrict__ ps,
                                                                      const int32_t *__restrict__ pe, int64_t n,
                                                                      uint32_t *__restrict__ counts,
                                                                      unsigned long long *__restrict__ block_totals) {
  __shared__ unsigned long long wt[kSweepThreads / 32];
  int64_t i = (int64_t)blockIdx.x * kSweepThreads + threadIdx.x;
  uint32_t cnt = 0;
  if (i < n) {
    cnt = probe_count<STRICT>(ix, pc[i], ps[i], pe[i]);
    counts[i] = cnt;
  }
  unsigned long long v = cnt;
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) wt[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < kSweepThreads / 32; ++w) t += wt[w];
    block_totals[blockIdx.x] = t;
  }
}

// ---- overlap pass 2: write (probe_row, build_row) pairs at exact offsets ---------------------
// Offsets = scanned block base + in-block exclusive scan of the pass-1 counts.  Short windows
// are walked by their own thread; long ones by the whole warp with ballot compaction so the
// stores of one probe are contiguous.
template <bool STRICT>
__global__ void __launch_bounds__(kSweepThreads) overlap_emit_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                     const int32_t *__restrict__ ps,
                                                                     const int32_t *__restrict__ pe, int64_t n,
                                                                     const uint32_t *__restrict__ counts,
                                                                     const unsigned long long *__restrict__ block_base,
                                                                     int64_t blk0,
                                                                     uint32_t *__restrict__ out_probe,
                                                                     uint32_t *__restrict__ out_build) {
  __shared__ unsigned long long wt[kSweepThreads / 32 + 1];
  const int64_t blk = blk0 + blockIdx.x;  // block range [blk0, blk0+grid): the output buffer starts at the first pair of block blk0
  const int64_t i = blk * kSweepThreads + threadIdx.x;
  const bool in_range = i < n;
  const uint32_t cnt = in_range ? counts[i] : 0u;
  unsigned long long pos = block_base[blk] - block_base[blk0] + block_exclusive<SumU64, kSweepThreads>((unsigned long long)cnt, wt);

  int32_t lo = 0, hi = 0, s = 0;
  if (cnt) {
    const int32_t c = pc[i];
    s = ps[i];
    probe_window<STRICT>(ix, ix.seg[c], ix.seg[c + 1], s, pe[i], lo, hi);
  }
  const bool heavy = cnt && (hi - lo) >= kHeavyWindow;
  if (cnt && !heavy) {
    for (int32_t j = lo; j < hi; ++j) {
      if (end_hits<STRICT>(__ldg(ix.en + j), s)) {
        out_probe[pos] = (uint32_t)i;
        out_build[pos] = __ldg(ix.row + j);
        ++pos;
      }
    }
  }
  unsigned hm = __ballot_sync(0xffffffffu, heavy);
  const int lane = threadIdx.x & 31;
  const unsigned lt = lanemask_lt();
  while (hm) {
    const int src = __ffs(hm) - 1;
    hm &= hm - 1;
    const int32_t l = __shfl_sync(0xffffffffu, lo, src), h = __shfl_sync(0xffffffffu, hi, src);
    const int32_t ss = __shfl_sync(0xffffffffu, s, src);
    unsigned long long p = __shfl_sync(0xffffffffu, pos, src);
    const uint32_t pi = (uint32_t)__shfl_sync(0xffffffffu, (unsigned long long)i, src);
    for (int32_t j0 = l; j0 < h; j0 += 32) {
      const int32_t j = j0 + lane;
      const bool ok = j < h && end_hits<STRICT>(__ldg(ix.en + j), ss);
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const unsigned long long k = p + __popc(m & lt);
        out_probe[k] = pi;
        out_build[k] = __ldg(ix.row + j);
      }
      p += __popc(m);
    }
  }
}


// =============================================================================================
// Fast path: joint rank directory over the global coordinate axis (index.cuh).  One 32-byte sector per
// probe (two for probes longer than a bucket); the count is the sweep-line identity  |{st < a.end}| - |{en <= a.start}|  (Strict)
// /  |{st <= a.end}| - |{en < a.start}|  (Weak)  (polars_bio/range_op.py:548-594).
// =============================================================================================
// one 256-bit request per record (LDG.E.256): two 128-bit loads cost two L2 sector requests each time the first is
// still in flight; no L1 allocation -- the directory is touched at random, nothing is reused
__device__ __forceinline__ void ld_jrec(const JRec *__restrict__ p, uint32_t (&w)[8]) {
  asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
      : "l"(p));
}
// number of the record's 12 fields below t (t <= 0x7FFF).  Packed compare: every field has bit 15 clear, so
// field + (0x8000 - t) carries into bit 15 exactly when field >= t and never into the neighbouring field; the six
// words' guard bits are moved to distinct positions and counted with one POPC.
__device__ __forceinline__ uint32_t jrec_nlt(const uint32_t (&w)[8], uint32_t t) {
  const uint32_t c = (0x8000u - t) * 0x00010001u;
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) acc |= ((w[2 + i] + c) & 0x80008000u) >> i;
  return (uint32_t)kJKeys - (uint32_t)__popc(acc);
}
__device__ __forceinline__ uint32_t search_g(const uint32_t *__restrict__ g, uint32_t a, uint32_t e, uint32_t x) {
  while (a < e) { const uint32_t mid = a + ((e - a) >> 1); if (__ldg(g + mid) < x) a = mid + 1; else e = mid; }
  return a;
}
// Both ranks of a proper probe with global-axis coordinates g_s <= g_e:
//   hi = number of indexed starts before the probe end   (Strict: gs <  g_e ; Weak: gs <= g_e)
//   re = number of indexed ends   before the probe start (Strict: ge <= g_s ; Weak: ge <  g_s)
// written as "keys below xS / xE" so one packed compare serves both predicates.
template <bool STRICT>
__device__ __forceinline__ void jdir_ranks(const IndexView &ix, uint32_t g_s, uint32_t g_e, uint32_t &hi, uint32_t &re) {
  const uint32_t xS = g_e + (STRICT ? 0u : 1u), xE = g_s + (STRICT ? 1u : 0u);  // the axis ends below 2^32-16: no wrap
  const int sh = ix.shift;
  const uint32_t b = g_s >> sh, lo = b << sh;
  uint32_t w[8];
  ld_jrec(ix.jdir + b, w);
  const uint32_t dS = xS - lo;
  const bool same = dS <= (2u << sh);  // the record covers the starts of [lo, lo+2W)
  if (!(w[0] & 0x80000000u)) {
    re = w[1] + jrec_nlt(w, 0x4000u + (xE - lo));
    if (same) { hi = w[0] + jrec_nlt(w, dS); return; }
  } else {  // crowded bucket: search inside its rank range (or count, when the bucket's ends are not sorted)
    if (ix.ge_sorted) re = search_g(ix.ge, w[1], w[3], xE);
    else { re = w[1]; for (uint32_t j = w[1]; j < w[3]; ++j) re += __ldg(ix.ge + j) < xE; }
    if (same) { hi = search_g(ix.gs, w[0] & 0x7fffffffu, w[2], xS); return; }
  }
  const uint32_t b2 = g_e >> sh, lo2 = b2 << sh;  // probe longer than the overlap: the record of its end bucket
  ld_jrec(ix.jdir + b2, w);
  if (!(w[0] & 0x80000000u)) hi = w[0] + jrec_nlt(w, xS - lo2);
  else hi = search_g(ix.gs, w[0] & 0x7fffffffu, w[2], xS);
}

constexpr uint32_t kGenericProbe = 0xFFFFFFFFu;  // pass-1 marker: this probe must take the generic window path

// count of one probe on the fast path; hi_out = its rank among the starts (end of its candidate window)
template <bool STRICT>
__device__ __forceinline__ uint32_t fast_count(const IndexView &ix, int32_t c, int32_t s, int32_t e, uint32_t &hi_out) {
  hi_out = 0;
  if (c < 0 || c >= ix.n_contigs) return 0;
  const ContigMap32 cm = ld_cmap32(ix.cmap32 + c);
  if (!cm.has) return 0;
  if (!(STRICT ? (s < e) : (s <= e))) {  // empty / inverted probe: identity not valid, bare predicate instead
    hi_out = kGenericProbe;
    return probe_count<STRICT>(ix, c, s, e);
  }
  // clamped into the contig's slice: order against every indexed coordinate is preserved
  const uint32_t g_s = global_of(cm, s), g_e = global_of(cm, e);
  uint32_t hi, re;
  jdir_ranks<STRICT>(ix, g_s, g_e, hi, re);
  hi_out = hi;
  return hi - re;
}

// Candidate window [lo,hi) and end-rank `lend` of one probe via the rank directories (nearest / coverage).
// hi and lend are global positions == positions in the start-ordered / end-ordered arrays.  The window start is
// hi-cnt when nothing below it reaches past the probe start (no nested intervals there: one load), else a search
// over the running max inside the contig segment.  Returns false when the probe must take the generic path.
template <bool STRICT>
__device__ __forceinline__ bool fast_window(const IndexView &ix, int32_t c, int32_t s, int32_t e, int32_t seg_lo,
                                            int32_t &lo, int32_t &hi, int32_t &lend) {
  if (!ix.jdir || !(STRICT ? (s < e) : (s <= e))) return false;
  const ContigMap32 cm = ld_cmap32(ix.cmap32 + c);
  const uint32_t g_s = global_of(cm, s), g_e = global_of(cm, e);
  uint32_t uh, ur;
  jdir_ranks<STRICT>(ix, g_s, g_e, uh, ur);
  hi = (int32_t)uh;
  lend = (int32_t)ur;
  const int32_t p = lend;  // = hi - cnt
  if (p >= hi) { lo = hi; return true; }
  if (p <= seg_lo || !end_hits<STRICT>(__ldg(ix.pmax + p - 1), s)) lo = p;
  else lo = STRICT ? upper_bound_i32(ix.pmax, seg_lo, p, s) : lower_bound_i32(ix.pmax, seg_lo, p, s);
  return true;
}

// ITEMS probes per thread (strided by the block size, so loads stay coalesced): their directory lookups are
// independent, which keeps more L2 requests in flight per warp.
template <bool STRICT, typename OutT, int ITEMS>  // OutT: int64_t (the ABI's count column) or uint32_t (half the D2H for the Arrow bridge)
__global__ void __launch_bounds__(kSweepThreads) count_overlaps_fast_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                            const int32_t *__restrict__ ps,
                                                                            const int32_t *__restrict__ pe, int64_t n,
                                                                            OutT *__restrict__ counts) {
  const int64_t base = (int64_t)blockIdx.x * (kSweepThreads * ITEMS) + threadIdx.x;
  int32_t c[ITEMS], s[ITEMS], e[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * kSweepThreads;
    const bool ok = i < n;
    c[j] = ok ? pc[i] : -1; s[j] = ok ? ps[i] : 0; e[j] = ok ? pe[i] : 0;
  }
  uint32_t cnt[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) { uint32_t hi; cnt[j] = fast_count<STRICT>(ix, c[j], s[j], e[j], hi); }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * kSweepThreads;
    if (i < n) counts[i] = (OutT)cnt[j];
  }
}

// Exact 64-bit sum of one uint32 per lane with two REDUX instructions (16-bit halves cannot overflow in a warp):
// ten SHFLs for a 64-bit butterfly would share the MIO/LSU path these L1TEX-bound kernels are limited by.
__device__ __forceinline__ unsigned long long warp_sum_u32(uint32_t v) {
  const uint32_t lo = __reduce_add_sync(0xffffffffu, v & 0xffffu), hi = __reduce_add_sync(0xffffffffu, v >> 16);
  return (unsigned long long)lo + ((unsigned long long)hi << 16);
}
// Pass 1 on the fast path.  Besides the per-probe (count, start rank) it produces the pair offsets pass 2 needs, in
// the same launch: the offset of every 32-probe group inside its 256-probe block (warp_off), and -- LOOKBACK -- the
// exclusive pair offset of every 256-probe block plus the grand total by a decoupled look-back over the launch's
// blocks (ticket order; status word = 2 flag bits | 62-bit pair count; warp 0 inspects 32 predecessors per step).
// That replaces the three-kernel device scan + a host fetch: the block holding the last ticket posts the total
// straight into the caller's host mailbox (pbgpu.cu) when one is given.  Without LOOKBACK the raw block totals are
// left in block_base for a separate scan (A/B: PBGPU_P1SCAN=kernels).
constexpr unsigned long long kP1Agg = 1ull << 62, kP1Pre = 2ull << 62, kP1Mask = (1ull << 62) - 1ull;
template <bool STRICT, int ITEMS, bool LOOKBACK>
__global__ void __launch_bounds__(kSweepThreads) overlap_count_fast_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                           const int32_t *__restrict__ ps,
                                                                           const int32_t *__restrict__ pe, int64_t n,
                                                                           uint32_t *__restrict__ counts, uint32_t *__restrict__ his,
                                                                           unsigned long long *__restrict__ block_base /*[nblk+1]*/,
                                                                           unsigned long long *__restrict__ warp_off /*[n/32] or NULL*/,
                                                                           unsigned long long *status /*[grid], zeroed*/,
                                                                           unsigned int *ticket /*zeroed*/,
                                                                           volatile unsigned long long *mailbox, unsigned long long mailbox_seq) {
  __shared__ unsigned long long wt[ITEMS][kSweepThreads / 32];
  __shared__ unsigned int tile_s;
  unsigned int tile = blockIdx.x;
  if (LOOKBACK) {
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    __syncthreads();
    tile = tile_s;
  }
  const int64_t base = (int64_t)tile * (kSweepThreads * ITEMS) + threadIdx.x;
  int32_t c[ITEMS], s[ITEMS], e[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * kSweepThreads;
    const bool ok = i < n;
    c[j] = ok ? pc[i] : -1; s[j] = ok ? ps[i] : 0; e[j] = ok ? pe[i] : 0;
  }
  uint32_t cnt[ITEMS], hi[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) cnt[j] = fast_count<STRICT>(ix, c[j], s[j], e[j], hi[j]);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * kSweepThreads;
    if (i < n) { counts[i] = cnt[j]; his[i] = hi[j]; }
    const unsigned long long v = warp_sum_u32(i < n ? cnt[j] : 0u);
    if ((threadIdx.x & 31) == 0) wt[j][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  // offset of every 32-probe group inside its 256-probe block: lets pass 2 run warp by warp with no block-wide barrier
  if (warp_off && threadIdx.x >= 32 && threadIdx.x < 32 + ITEMS * (kSweepThreads / 32)) {
    const int q = threadIdx.x - 32;
    const int j = q / (kSweepThreads / 32), w = q % (kSweepThreads / 32);
    unsigned long long t = 0;
    for (int k = 0; k < w; ++k) t += wt[j][k];
    const int64_t g = ((int64_t)tile * ITEMS + j) * (kSweepThreads / 32) + w;
    if (g * 32 < n) warp_off[g] = t;
  }
  if (threadIdx.x >= 32) return;
  // warp 0: one total per 256 consecutive probes (lane j < ITEMS holds the total of sub-block j)
  const int lane = threadIdx.x;
  unsigned long long t = 0;
  if (lane < ITEMS) {
#pragma unroll
    for (int w = 0; w < kSweepThreads / 32; ++w) t += wt[lane][w];
  }
  const int64_t blk = (int64_t)tile * ITEMS + lane;
  if (!LOOKBACK) {
    if (lane < ITEMS && blk * kSweepThreads < n) block_base[blk] = t;
    return;
  }
  unsigned long long total = 0, before = 0;  // pairs of the whole tile; pairs of the sub-blocks ahead of this lane's
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const unsigned long long tj = __shfl_sync(0xffffffffu, t, j);
    total += tj;
    if (j < lane) before += tj;
  }
  unsigned long long excl = 0;
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(status, total | kP1Pre);
  } else {
    if (lane == 0) st_volatile_u64(status + tile, total | kP1Agg);
    for (long long top = (long long)tile - 1;; top -= 32) {
      const long long idx = top - lane;
      unsigned long long w = kP1Pre;  // below tile 0: an empty prefix ends the walk
      if (idx >= 0) { do { w = ld_volatile_u64(status + idx); } while ((w >> 62) == 0ull); }
      const unsigned pre = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
      const int first = pre ? __ffs(pre) - 1 : 31;
      unsigned long long v = lane <= first ? (w & kP1Mask) : 0ull;
#pragma unroll
      for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      excl += v;
      if (pre) break;
    }
    if (lane == 0) st_volatile_u64(status + tile, (excl + total) | kP1Pre);
  }
  if (lane < ITEMS && blk * kSweepThreads < n) block_base[blk] = excl + before;
  if (tile == gridDim.x - 1 && lane == 0) {
    const int64_t nblk = (n + kSweepThreads - 1) / kSweepThreads;
    block_base[nblk] = excl + total;
    if (mailbox) {
      mailbox[0] = excl + total;
      __threadfence_system();
      mailbox[15] = mailbox_seq;
    }
  }
}
// pass 2 on the fast path: the hits of a probe are the `cnt` entries below its start-rank `hi` whose end
// reaches past the probe start, so walk down from hi-1 until cnt of them are found (exactly cnt steps when
// the indexed intervals do not nest) and write them back to front: output stays ordered by (start,row).
constexpr uint32_t kHeavyCount = 32;
template <bool STRICT>
__global__ void __launch_bounds__(kSweepThreads) overlap_emit_fast_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                          const int32_t *__restrict__ ps,
                                                                          const int32_t *__restrict__ pe, int64_t n,
                                                                          const uint32_t *__restrict__ counts,
                                                                          const uint32_t *__restrict__ his,
                                                                          const unsigned long long *__restrict__ block_base,
                                                                          int64_t blk0,
                                                                          uint32_t *__restrict__ out_probe,
                                                                          uint32_t *__restrict__ out_build) {
  __shared__ unsigned long long wt[kSweepThreads / 32 + 1];
  const int64_t blk = blk0 + blockIdx.x;  // block range [blk0, blk0+grid): the output buffer starts at the first pair of block blk0
  const int64_t i = blk * kSweepThreads + threadIdx.x;
  // all three coalesced loads are issued before the block scan so their latency hides behind it
  uint32_t cnt = 0, hi = 0;
  int32_t s = 0;
  if (i < n) { cnt = counts[i]; hi = his[i]; s = ps[i]; }
  const unsigned long long pos = block_base[blk] - block_base[blk0] + block_exclusive<SumU64, kSweepThreads>((unsigned long long)cnt, wt);
  const bool generic = cnt && hi == kGenericProbe;
  const bool heavy = cnt >= kHeavyCount && !generic;
  if (generic) {  // rare: empty / inverted probe interval
    int32_t lo, h2;
    const int32_t c = pc[i];
    probe_window<STRICT>(ix, ix.seg[c], ix.seg[c + 1], s, pe[i], lo, h2);
    unsigned long long p = pos;
    for (int32_t j = lo; j < h2; ++j)
      if (end_hits<STRICT>(__ldg(ix.en + j), s)) { out_probe[p] = (uint32_t)i; out_build[p] = __ldg(ix.row + j); ++p; }
  } else if (cnt && !heavy) {
    uint32_t k = 0;
    for (int64_t j = (int64_t)hi - 1; k < cnt && j >= 0; --j) {
      const uint2 v = __ldg(ix.er + j);  // (end, row) interleaved: one 8-byte random load per candidate
      const int32_t ev = (int32_t)v.x;
      const uint32_t rv = v.y;
      if (end_hits<STRICT>(ev, s)) {
        const unsigned long long p = pos + (cnt - 1 - k);
        out_probe[p] = (uint32_t)i;
        out_build[p] = rv;
        ++k;
      }
    }
  }
  unsigned hm = __ballot_sync(0xffffffffu, heavy);
  const int lane = threadIdx.x & 31;
  const unsigned lt = lanemask_lt();
  while (hm) {  // many hits: the whole warp walks the window, 32 candidates per step, ballot-compacted stores
    const int src = __ffs(hm) - 1;
    hm &= hm - 1;
    const uint32_t h = __shfl_sync(0xffffffffu, hi, src), c_all = __shfl_sync(0xffffffffu, cnt, src);
    const int32_t ss = __shfl_sync(0xffffffffu, s, src);
    const unsigned long long p0 = __shfl_sync(0xffffffffu, pos, src);
    const uint32_t pi = (uint32_t)__shfl_sync(0xffffffffu, (unsigned long long)i, src);
    uint32_t found = 0;
    for (int64_t top = (int64_t)h - 1; found < c_all && top >= 0; top -= 32) {
      const int64_t j = top - lane;  // lane 0 = highest position
      const bool ok = j >= 0 && end_hits<STRICT>(__ldg(ix.en + j), ss);
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint32_t k = found + __popc(m & lt);
        if (k < c_all) {
          const unsigned long long p = p0 + (c_all - 1 - k);
          out_probe[p] = pi;
          out_build[p] = __ldg(ix.row + j);
        }
      }
      found += __popc(m);
    }
  }
}

// pass 2 when the indexed intervals do not nest (ends ascend with starts inside every contig -- SNVs, reads, exons of
// one transcript): the hits of a proper probe are exactly the positions [hi-cnt, hi) of the start order, so pass 2 is
// a pure expansion of (cnt, hi) into pairs and never looks at a coordinate.  Warp-granular: a warp owns 32
// consecutive probes, output slot j of the warp belongs to the last probe whose exclusive offset is <= j (five-step
// shuffle search), so the 32 lanes write 32 CONSECUTIVE pairs per step whatever the individual counts are -- stores
// coalesce, the work is balanced across lanes and long hit lists need no separate path.  No block-wide barrier:
// block base (scanned block totals) + the group's offset inside the block (pass 1) + the warp scan (32-bit: the
// caller takes this kernel only when 32 x indexed rows < 2^32).
template <bool STRICT, int ITEMS>
__global__ void __launch_bounds__(kSweepThreads) overlap_emit_flat_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                          const int32_t *__restrict__ ps,
                                                                          const int32_t *__restrict__ pe, int64_t n,
                                                                          const uint32_t *__restrict__ counts,
                                                                          const uint32_t *__restrict__ his,
                                                                          const unsigned long long *__restrict__ block_base,
                                                                          const unsigned long long *__restrict__ warp_off,
                                                                          int64_t blk0, int64_t blk_hi,
                                                                          uint32_t *__restrict__ out_probe,
                                                                          uint32_t *__restrict__ out_build,
                                                                          const uint32_t *__restrict__ ids /*NULL: probe row*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long base0 = block_base[blk0];
  uint32_t cnt[ITEMS], hi[ITEMS];
  unsigned long long bbase[ITEMS], woff[ITEMS];
  int64_t idx[ITEMS];
  uint32_t pid[ITEMS];
#pragma unroll
  for (int t = 0; t < ITEMS; ++t) {  // every load of the warp's ITEMS groups is in flight before the first use
    const int64_t blk = blk0 + (int64_t)blockIdx.x * ITEMS + t;
    const int64_t i = blk * kSweepThreads + threadIdx.x;
    idx[t] = i;
    const bool ok = blk < blk_hi && i < n;
    cnt[t] = ok ? counts[i] : 0u;
    hi[t] = ok ? his[i] : 0u;
    pid[t] = (ok && ids) ? ids[i] : (uint32_t)i;
    const bool gok = blk < blk_hi && (i - lane) < n;
    woff[t] = gok ? warp_off[blk * (kSweepThreads / 32) + warp] : 0ull;
    bbase[t] = gok ? block_base[blk] : base0;
  }
#pragma unroll
  for (int t = 0; t < ITEMS; ++t) {
    uint32_t incl = cnt[t];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;  // warp-uniform
    const uint32_t excl = incl - cnt[t];
    const unsigned long long wpos = bbase[t] - base0 + woff[t];
    const uint32_t first = hi[t] - cnt[t];  // first hit position (garbage for generic probes: never used)
    const bool generic = cnt[t] && hi[t] == kGenericProbe;
    if (generic) {  // rare: empty / inverted probe interval, bare predicate over its window
      const int64_t i = idx[t];
      int32_t lo, h2;
      const int32_t c = pc[i], s = ps[i];
      probe_window<STRICT>(ix, ix.seg[c], ix.seg[c + 1], s, pe[i], lo, h2);
      unsigned long long p = wpos + excl;
      for (int32_t j = lo; j < h2; ++j)
        if (end_hits<STRICT>(__ldg(ix.en + j), s)) { out_probe[p] = pid[t]; out_build[p] = __ldg(ix.row + j); ++p; }
    }
    for (uint32_t j0 = 0; j0 < total; j0 += 32) {
      const uint32_t j = j0 + lane;
      int p = 0;
#pragma unroll
      for (int step = 16; step; step >>= 1) {
        const int cand = p + step;
        const uint32_t e = __shfl_sync(0xffffffffu, excl, cand & 31);
        if (e <= j) p = cand;  // cand <= 31 always: p < 32 - step before the step
      }
      const uint32_t e_p = __shfl_sync(0xffffffffu, excl, p);
      const uint32_t f_p = __shfl_sync(0xffffffffu, first, p);
      const bool g_p = __shfl_sync(0xffffffffu, (int)generic, p) != 0;
      const uint32_t id_p = __shfl_sync(0xffffffffu, pid[t], p);
      if (j < total && !g_p) {
        out_probe[wpos + j] = id_p;
        out_build[wpos + j] = __ldg(ix.row + (f_p + (j - e_p)));
      }
    }
  }
}

// ---- coverage: positions of the probe covered by the union of indexed rows -------------------
template <bool STRICT>
__global__ void __launch_bounds__(kSweepThreads) coverage_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                 const int32_t *__restrict__ ps,
                                                                 const int32_t *__restrict__ pe, int64_t n,
                                                                 int64_t *__restrict__ cov) {
  int64_t i = (int64_t)blockIdx.x * kSweepThreads + threadIdx.x;
  if (i >= n) return;
  const int32_t c = pc[i];
  long long total = 0;
  if (c >= 0 && c < ix.n_contigs) {
    const int32_t s = ps[i], e = pe[i];
    int32_t lo, hi, lend;
    const int32_t seg_lo = ix.seg[c], seg_hi = ix.seg[c + 1];
    if (seg_lo >= seg_hi) { cov[i] = 0; return; }
    if (!fast_window<STRICT>(ix, c, s, e, seg_lo, lo, hi, lend)) probe_window<STRICT>(ix, seg_lo, seg_hi, s, e, lo, hi);
    long long cs = 0, ce = 0;
    bool open = false;
    for (int32_t j = lo; j < hi; ++j) {  // hits arrive start-sorted: merge clipped pieces on the fly
      const int32_t be_raw = __ldg(ix.en + j);
      if (!end_hits<STRICT>(be_raw, s)) continue;
      long long bs = __ldg(ix.st + j), be = be_raw;
      if (bs < s) bs = s;
      if (be > e) be = e;
      if (!STRICT) be += 1;  // closed -> half-open in 64 bit
      if (be <= bs) continue;
      if (!open) { cs = bs; ce = be; open = true; }
      else if (bs <= ce) { if (be > ce) ce = be; }
      else { total += ce - cs; cs = bs; ce = be; }
    }
    if (open) total += ce - cs;
  }
  cov[i] = total;
}

// ---- nearest ------------------------------------------------------------------------------
// Ordering: overlapping partners first (when included) in (start,row) order, then the merge of
//   upstream   rows ending before the probe starts, by descending end; equal ends in (start,row) order
//   downstream rows starting after the probe ends, in (start,row) order
// by (distance, start, row).  Mirrors oracle/interval_oracle.c:pbo_nearest decision for decision.
__device__ __forceinline__ long long gap_of(int32_t as, int32_t ae, int32_t bs, int32_t be) {
  long long d1 = (long long)bs - (long long)ae, d2 = (long long)as - (long long)be;
  long long d = d1 > d2 ? d1 : d2;
  return d > 0 ? d : 0;
}

template <bool STRICT>
__global__ void __launch_bounds__(kSweepThreads) nearest_kernel(IndexView ix, const int32_t *__restrict__ pc,
                                                                const int32_t *__restrict__ ps,
                                                                const int32_t *__restrict__ pe, int64_t n, int64_t k,
                                                                int include_overlaps, uint32_t *__restrict__ partner,
                                                                int64_t *__restrict__ dist) {
  const int64_t i = (int64_t)blockIdx.x * kSweepThreads + threadIdx.x;
  if (i >= n) return;
  uint32_t *ob = partner + i * k;
  int64_t *od = dist ? dist + i * k : nullptr;
  for (int64_t j = 0; j < k; ++j) { ob[j] = PBGPU_NO_PARTNER; if (od) od[j] = -1; }
  const int32_t c = pc[i];
  if (c < 0 || c >= ix.n_contigs) return;
  const int32_t seg_lo = ix.seg[c], seg_hi = ix.seg[c + 1];
  if (seg_lo >= seg_hi) return;
  const int32_t qs = ps[i], qe = pe[i];
  int32_t lo, hi, lend;
  if (!fast_window<STRICT>(ix, c, qs, qe, seg_lo, lo, hi, lend)) {
    probe_window<STRICT>(ix, seg_lo, seg_hi, qs, qe, lo, hi);
    lend = STRICT ? upper_bound_i32(ix.en_sorted, seg_lo, seg_hi, qs) : lower_bound_i32(ix.en_sorted, seg_lo, seg_hi, qs);
  }
  int64_t got = 0;
  if (include_overlaps) {
    for (int32_t j = lo; j < hi && got < k; ++j) {
      if (end_hits<STRICT>(__ldg(ix.en + j), qs)) { ob[got] = __ldg(ix.row + j); if (od) od[got] = 0; ++got; }
    }
  }
  if (got >= k) return;
  int32_t rp = hi;  // downstream cursor: first start past the probe end
  int32_t lg_hi = lend, lg_lo = lend, lcur = lend;  // current equal-end group [lg_lo,lg_hi), cursor lcur
  for (;;) {
    bool have_l = false, have_r = false;
    int32_t lpos = 0;
    long long ld = 0, rd = 0;
    for (;;) {
      if (lcur >= lg_hi) {  // open the next (smaller end) group
        if (lg_lo <= seg_lo) break;
        lg_hi = lg_lo;
        const int32_t v = __ldg(ix.en_sorted + lg_hi - 1);
        int32_t g = lg_hi - 1;  // start of the equal-end group: groups are tiny, walk back before searching
        for (int st = 0; g > seg_lo && __ldg(ix.en_sorted + g - 1) == v; ) {
          --g;
          if (++st == 8) { g = lower_bound_i32(ix.en_sorted, seg_lo, g, v); break; }
        }
        lg_lo = g;
        lcur = lg_lo;
      }
      lpos = ix.en_pos ? (int32_t)__ldg(ix.en_pos + lcur) : lcur;  // NULL: end order == start order
      const int32_t bs = __ldg(ix.st + lpos), be = __ldg(ix.en + lpos);
      if (is_hit<STRICT>(qs, qe, bs, be)) { ++lcur; continue; }
      have_l = true;
      ld = gap_of(qs, qe, bs, be);
      break;
    }
    while (rp < seg_hi) {
      const int32_t bs = __ldg(ix.st + rp), be = __ldg(ix.en + rp);
      if (is_hit<STRICT>(qs, qe, bs, be) || (STRICT ? (be <= qs) : (be < qs))) { ++rp; continue; }
      have_r = true;
      rd = gap_of(qs, qe, bs, be);
      break;
    }
    if (!have_l && !have_r) break;
    bool take_left;
    if (have_l && have_r) {
      const int32_t lst = __ldg(ix.st + lpos), rst = __ldg(ix.st + rp);
      if (ld != rd) take_left = ld < rd;
      else if (lst != rst) take_left = lst < rst;
      else take_left = __ldg(ix.row + lpos) < __ldg(ix.row + rp);
    } else take_left = have_l;
    if (take_left) { ob[got] = __ldg(ix.row + lpos); if (od) od[got] = ld; ++lcur; }
    else { ob[got] = __ldg(ix.row + rp); if (od) od[got] = rd; ++rp; }
    if (++got >= k) break;
  }
}

}  // namespace pbgpu
