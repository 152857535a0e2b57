// index.cuh -- the search structure over the indexed ("build") table.
//
// Replaces the per-contig COITrees the reference's providers build from their indexed side
// (docs/developers.md:629-649; constructor call sites src/operation.rs:146-158,253-263,331-340).
//
// HBM layout (m = rows with a valid contig code; all arrays int32/uint32, m entries):
//   seg[n_contigs+1]  first position of every contig in the sorted order (radix partition)
//   st, en, row       rows sorted by (contig, start, row): start, end, original row id
//   pmax              running maximum of `en` inside the contig  (monotone -> searchable:
//                     every hit of a probe lies in [first pmax > a.start, first st >= a.end))
//   en_sorted         ends sorted by (contig, end, start, row)  (rank identity:
//                     count(a) = |{st < a.end}| - |{en <= a.start}|, range_op.py:548-594)
//   en_pos            position in (st,en,row) order of every en_sorted entry (nearest upstream)
// Fast path (no inverted rows, total coordinate span < 2^32): every contig is shifted to its own slice of
// one global uint32 axis (G(c,p) = off[c] + clamp(p, lo_c-1, hi_c+1) - (lo_c-1)), which makes the two sorted
// arrays globally monotone, and both share ONE bucketed *joint rank directory*:
//   gs, ge            global-axis starts (start order) / ends (end order), uint32[m]
//   jdir              one 32-byte record per W = 2^shift wide bucket b of the axis (JRec below): the rank of the
//                     bucket's first start and first end, then up to 12 in-bucket keys as 15-bit offsets: the starts
//                     of [bW, bW+2W) -- the record OVERLAPS the next bucket -- and the ends of [bW, bW+W).
//                     A probe needs rank_S(a.end) and rank_E(a.start); a.start picks the bucket, and whenever
//                     a.end < bW+2W (always when the probe is no longer than W) the same record answers both:
//                     ONE random 32-byte sector per probe (the L1TEX t-stage serves one divergent request per
//                     clock per SM, so requests per probe is what bounds the count kernels), no search.  Longer
//                     probes read the record of a.end's bucket too; crowded buckets (more than 12 keys) keep the
//                     rank range instead of keys and are searched.
#pragma once
#include <mutex>

#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace pbgpu {
struct ContigMap {   // per contig: position of its slice on the global axis
  long long lo_m1;   // (smallest indexed coordinate) - 1
  long long hi_p1;   // (largest indexed coordinate) + 1
  uint32_t off;      // global coordinate of lo_m1
  int32_t has;       // contig has indexed rows
};
// The same slice in 32-bit form, one 16-byte load per probe (round 2; the 64-bit clamp against lo_m1 / hi_p1 and the three
// 8-byte loads of a ContigMap were a tenth of the instructions of a count-kernel probe):
//   G(c, p) = clamp(p, lo, hi) + delta  (mod 2^32),  lo = max(lo_m1, INT32_MIN), hi = min(hi_p1, INT32_MAX) -- an int32
//   coordinate cannot lie beyond those anyway -- and delta = off - lo_m1 (mod 2^32).
struct alignas(16) ContigMap32 {
  int32_t lo, hi;
  uint32_t delta;
  int32_t has;
};
// Joint rank record of bucket b (W = 2^shift, lo = b*W):
//   w[0]  rank of the first start >= lo (#{gs < lo}); bit 31 = crowded
//   normal : w[1] = #{ge < lo} - nS   (nS = number of start keys stored, so that w[1] + #{fields < 0x4000+t} is
//                                      the end rank: every start field is < 0x4000)
//            w[2..7] = 12 x 16-bit fields, ascending: starts of [lo, lo+2W) as (g - lo) < 0x4000, then ends of
//                      [lo, lo+W) as 0x4000 | (g - lo), padded with 0x7FFF.  Bit 15 of every field is clear: that
//                      is the guard bit of the packed compare in jrec_nlt().
//   crowded: w[1] = #{ge < lo}, w[2] = #{gs < lo+2W}, w[3] = #{ge < lo+W}: rank ranges to search inside
struct alignas(32) JRec {
  uint32_t w[8];
};
constexpr int kJKeys = 12;
constexpr int kJMaxShift = 13;  // 2W <= 0x4000
}  // namespace pbgpu

struct pbgpu_index {
  int64_t m = 0;  // valid rows
  int64_t m_in = 0;
  int32_t n_contigs = 0;
  int device = 0;
  int has_inverted = 0;  // some indexed row has start > end: rank identity disabled
  int nested = 0;        // some indexed interval contains a later-starting one (ends not ascending in start order)
  int32_t *seg = nullptr;
  int32_t *st = nullptr, *en = nullptr, *pmax = nullptr, *en_sorted = nullptr;
  uint32_t *row = nullptr, *en_pos = nullptr;
  // fast path
  int fast = 0;
  int shift = 0;
  uint32_t n_buckets = 0;
  uint32_t axis_span = 0;  // length of the global axis (fast path)
  pbgpu::ContigMap *cmap = nullptr;
  pbgpu::ContigMap32 *cmap32 = nullptr;
  uint32_t *gs = nullptr, *ge = nullptr;
  pbgpu::JRec *jdir = nullptr;
  uint2 *er = nullptr;  // (end, row) interleaved in start order: one 8-byte load per emitted pair in pass 2
  void *slab = nullptr, *slab2 = nullptr, *slab_n = nullptr;  // stream-ordered allocations backing every array above
  // Round 2: with nested intervals the ends are no longer sorted at build time.  `ge` then holds the global-axis ends
  // grouped by directory bucket (ge_sorted = 0: unordered inside a bucket) and en_sorted / en_pos stay NULL until a
  // caller needs the end order (nearest): built once, on that call's stream, into slab_e (ensure_end_order, pbgpu.cu).
  int ge_sorted = 1;
  int32_t min_end = 0, max_end = 0;
  void *slab_e = nullptr;
  cudaEvent_t end_ready = nullptr;
  cudaStream_t end_stream = nullptr;
  std::mutex mu;
  // without nested intervals pmax and en_sorted alias `en` and en_pos is NULL (= identity)
  size_t bytes = 0;
};

namespace pbgpu {

struct IndexView {
  const ContigMap *__restrict__ cmap;
  const ContigMap32 *__restrict__ cmap32;
  const uint32_t *__restrict__ gs;
  const uint32_t *__restrict__ ge;
  const JRec *__restrict__ jdir;
  const uint2 *__restrict__ er;
  int shift;
  const int32_t *__restrict__ seg;
  const int32_t *__restrict__ st;
  const int32_t *__restrict__ en;
  const int32_t *__restrict__ pmax;
  const int32_t *__restrict__ en_sorted;
  const uint32_t *__restrict__ row;
  const uint32_t *__restrict__ en_pos;
  int32_t n_contigs;
  int has_inverted;
  int ge_sorted;  // 0: the ends of a crowded directory record are counted one by one
};

inline IndexView view_of(const pbgpu_index *ix) {
  return IndexView{ix->cmap, ix->cmap32, ix->gs, ix->ge, ix->jdir, ix->er, ix->shift, ix->seg, ix->st, ix->en, ix->pmax, ix->en_sorted, ix->row, ix->en_pos, ix->n_contigs, ix->has_inverted, ix->ge_sorted};
}

struct BuildStats {  // device-side reduction target
  int min_start, max_start, min_end, max_end;
  unsigned long long inverted, valid;
  unsigned long long max_len;  // longest (end - start) over the non-inverted rows
  unsigned long long blocks_done;  // prep_kernel: blocks that have added their share (the last one posts to the host)
};

__device__ __forceinline__ ContigMap32 ld_cmap32(const ContigMap32 *__restrict__ p) {
  const int4 v = __ldg(reinterpret_cast<const int4 *>(p));
  return ContigMap32{v.x, v.y, (uint32_t)v.z, v.w};
}
// global-axis coordinate of an int32 position on a contig that has indexed rows
__device__ __forceinline__ uint32_t global_of(const ContigMap32 &cm, int32_t p) {
  return (uint32_t)min(max(p, cm.lo), cm.hi) + cm.delta;
}
__global__ void __launch_bounds__(256) cmap32_kernel(const ContigMap *__restrict__ cmap, int32_t n_contigs, ContigMap32 *__restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n_contigs) return;
  const ContigMap m = cmap[c];
  ContigMap32 r{0, 0, 0u, 0};
  if (m.has) {
    r.lo = (int32_t)(m.lo_m1 < (long long)INT32_MIN ? (long long)INT32_MIN : m.lo_m1);
    r.hi = (int32_t)(m.hi_p1 > (long long)INT32_MAX ? (long long)INT32_MAX : m.hi_p1);
    r.delta = m.off - (uint32_t)(unsigned long long)m.lo_m1;
    r.has = 1;
  }
  out[c] = r;
}

// ---- prep: the one pass over the raw input ----------------------------------------------------------------------
// Per row: sort key  contig << 32 | (start ^ 0x80000000)  (null-keyed rows: contig = n_contigs, so the partition drops
// them behind the last real contig) and value  end << 32 | row;  per block: coordinate statistics and the histogram of
// every 8-bit key digit (the radix passes' totals), flushed with one atomic per non-empty bin.  Two rows per thread
// and iteration so six independent loads are in flight.
constexpr int kPrepThreads = 512;
__global__ void __launch_bounds__(kPrepThreads) prep_kernel(const int32_t *__restrict__ c, const int32_t *__restrict__ s,
                                                            const int32_t *__restrict__ e, int64_t n, int32_t n_contigs, int n_digits,
                                                            BuildStats *st, uint64_t *__restrict__ keys, uint64_t *__restrict__ vals,
                                                            uint32_t *__restrict__ digit_totals /*[kRsMaxPasses][256], zeroed*/,
                                                            volatile unsigned long long *mailbox, unsigned long long mailbox_seq,
                                                            const uint32_t *__restrict__ row_ids /*NULL: the row's position*/) {
  __shared__ uint32_t h[kRsMaxPasses][kRsRadix];
  __shared__ int sm[4][kPrepThreads / 32];
  __shared__ unsigned su[3][kPrepThreads / 32];
  for (int i = threadIdx.x; i < n_digits * kRsRadix; i += kPrepThreads) (&h[0][0])[i] = 0;
  __syncthreads();
  int mn_s = INT32_MAX, mx_s = INT32_MIN, mn_e = INT32_MAX, mx_e = INT32_MIN;
  unsigned inv = 0, val = 0, mlen = 0;
  const int64_t stride = (int64_t)gridDim.x * kPrepThreads;
  for (int64_t i0 = (int64_t)blockIdx.x * kPrepThreads + threadIdx.x; i0 < n; i0 += 2 * stride) {
    int32_t cc[2], ss[2], ee[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      const bool in = i < n;
      cc[u] = in ? c[i] : -1; ss[u] = in ? s[i] : 0; ee[u] = in ? e[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n) continue;
      const bool ok = cc[u] >= 0 && cc[u] < n_contigs;
      const uint64_t key = ((ok ? (uint64_t)cc[u] : (uint64_t)n_contigs) << 32) | (ok ? ((uint32_t)ss[u] ^ 0x80000000u) : 0u);
      keys[i] = key;
      vals[i] = ((uint64_t)(uint32_t)ee[u] << 32) | (row_ids ? row_ids[i] : (uint32_t)i);
      for (int p = 0; p < n_digits; ++p) atomicAdd(&h[p][(key >> (8 * p)) & 0xff], 1u);
      if (ok) {
        mn_s = min(mn_s, ss[u]); mx_s = max(mx_s, ss[u]);
        mn_e = min(mn_e, ee[u]); mx_e = max(mx_e, ee[u]);
        inv += ss[u] > ee[u];
        if (ee[u] >= ss[u]) mlen = max(mlen, (unsigned)((long long)ee[u] - (long long)ss[u]));
        ++val;
      }
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    mlen = max(mlen, __shfl_xor_sync(0xffffffffu, mlen, d));
    mn_s = min(mn_s, __shfl_xor_sync(0xffffffffu, mn_s, d));
    mx_s = max(mx_s, __shfl_xor_sync(0xffffffffu, mx_s, d));
    mn_e = min(mn_e, __shfl_xor_sync(0xffffffffu, mn_e, d));
    mx_e = max(mx_e, __shfl_xor_sync(0xffffffffu, mx_e, d));
    inv += __shfl_xor_sync(0xffffffffu, inv, d);
    val += __shfl_xor_sync(0xffffffffu, val, d);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sm[0][w] = mn_s; sm[1][w] = mx_s; sm[2][w] = mn_e; sm[3][w] = mx_e; su[0][w] = inv; su[1][w] = val; su[2][w] = mlen; }
  __syncthreads();
  if (threadIdx.x == 0) {  // one set of atomics per block
    for (int k = 1; k < kPrepThreads / 32; ++k) {
      sm[0][0] = min(sm[0][0], sm[0][k]); sm[1][0] = max(sm[1][0], sm[1][k]);
      sm[2][0] = min(sm[2][0], sm[2][k]); sm[3][0] = max(sm[3][0], sm[3][k]);
      su[0][0] += su[0][k]; su[1][0] += su[1][k]; su[2][0] = max(su[2][0], su[2][k]);
    }
    if (su[1][0]) {
      atomicMin(&st->min_start, sm[0][0]); atomicMax(&st->max_start, sm[1][0]);
      atomicMin(&st->min_end, sm[2][0]); atomicMax(&st->max_end, sm[3][0]);
      if (su[0][0]) atomicAdd(&st->inverted, (unsigned long long)su[0][0]);
      atomicAdd(&st->valid, (unsigned long long)su[1][0]);
      atomicMax(&st->max_len, (unsigned long long)su[2][0]);
    }
    if (mailbox) {  // the block that finishes last posts the five statistics words to the host mailbox (pbgpu.cu)
      __threadfence();
      if (atomicAdd(&st->blocks_done, 1ull) == (unsigned long long)gridDim.x - 1ull) {
        __threadfence();
        const volatile unsigned long long *w = reinterpret_cast<const volatile unsigned long long *>(st);
        for (int k = 0; k < 5; ++k) mailbox[k] = w[k];
        __threadfence_system();
        mailbox[15] = mailbox_seq;
      }
    }
  }
  for (int i = threadIdx.x; i < n_digits * kRsRadix; i += kPrepThreads) {
    const uint32_t v = (&h[0][0])[i];
    if (v) atomicAdd(digit_totals + i, v);
  }
}

// unpack the start-sorted pairs into SoA, find the contig segments (boundary detection: the thread that sees a
// contig change writes seg[] for every contig in between, empty ones included) and flag end inversions
// (en[i] < en[i-1] inside a contig; none <=> no nested intervals <=> ends are already sorted in start order).
__global__ void __launch_bounds__(256) unpack_sorted_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals,
                                                            int64_t m, int pos_bits, uint32_t bias_s, int32_t n_contigs,
                                                            int32_t *__restrict__ st, int32_t *__restrict__ en,
                                                            uint32_t *__restrict__ row, uint2 *__restrict__ er,
                                                            int32_t *__restrict__ seg,
                                                            unsigned long long *__restrict__ inversions) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned inv = 0;
  if (i < m) {
    const uint64_t k = keys[i], v = vals[i];
    const long long contig = (long long)(k >> pos_bits);
    const uint32_t sp = (uint32_t)(k & ((1ull << pos_bits) - 1ull));
    const int32_t e = (int32_t)(uint32_t)(v >> 32);
    st[i] = (int32_t)(sp ^ bias_s);
    en[i] = e;
    row[i] = (uint32_t)v;
    er[i] = make_uint2((uint32_t)e, (uint32_t)v);
    long long prev = -1;
    if (i > 0) {
      prev = (long long)(keys[i - 1] >> pos_bits);
      if (prev == contig) inv = e < (int32_t)(uint32_t)(vals[i - 1] >> 32);
    }
    for (long long c = prev + 1; c <= contig; ++c) seg[c] = (int32_t)i;
    if (i == m - 1) for (long long c = contig + 1; c <= n_contigs; ++c) seg[c] = (int32_t)m;
  }
  // only "any inversion at all?" matters.  One update per BLOCK that saw one, and only while the flag still reads 0:
  // an atomic per warp on this single address serialised 2.8M updates (1.2 ms at 90M nested rows)
  const int any = __syncthreads_or((int)inv);
  if (any && threadIdx.x == 0 && *(volatile unsigned long long *)inversions == 0ull) atomicMax(inversions, 1ull);
}

// nested case only: contig-tagged end keys for the running max and the (contig | end) keys of the second sort
__global__ void __launch_bounds__(256) make_end_keys_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ en, int64_t m,
                                                            int pos_bits, uint32_t bias_e, uint64_t *__restrict__ pm_keys,
                                                            uint64_t *__restrict__ ekeys, uint64_t *__restrict__ evals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t contig = keys[i] >> pos_bits;
  const int32_t e = en[i];
  pm_keys[i] = (contig << 32) | ((uint32_t)e ^ 0x80000000u);
  ekeys[i] = (contig << pos_bits) | ((uint32_t)e ^ bias_e);
  evals[i] = (uint64_t)i;
}

__global__ void __launch_bounds__(256) unpack_pmax_kernel(const uint64_t *__restrict__ pm_keys, int64_t m,
                                                          int32_t *__restrict__ pmax) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) pmax[i] = (int32_t)((uint32_t)pm_keys[i] ^ 0x80000000u);
}

__global__ void __launch_bounds__(256) unpack_ends_kernel(const uint64_t *__restrict__ ekeys, const uint64_t *__restrict__ evals,
                                                          int64_t m, int pos_bits, uint32_t bias_e,
                                                          int32_t *__restrict__ en_sorted, uint32_t *__restrict__ en_pos) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  en_sorted[i] = (int32_t)((uint32_t)(ekeys[i] & ((1ull << pos_bits) - 1ull)) ^ bias_e);
  en_pos[i] = (uint32_t)evals[i];
}

// ---- fast-path construction ------------------------------------------------------------------
// per contig: coordinate range of its indexed rows -> width of its slice on the global axis
__global__ void __launch_bounds__(128) contig_span_kernel(const int32_t *__restrict__ seg, const int32_t *__restrict__ st,
                                                          long long max_len, int32_t n_contigs,
                                                          ContigMap *__restrict__ cmap, unsigned long long *__restrict__ span) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_contigs) return;
  const int32_t lo = seg[c], hi = seg[c + 1];
  ContigMap m;
  m.off = 0;
  if (lo >= hi) { m.lo_m1 = 0; m.hi_p1 = 0; m.has = 0; span[c] = 0; }
  else {
    m.lo_m1 = (long long)st[lo] - 1;                 // rows are not inverted on this path: min coordinate = min start
    m.hi_p1 = (long long)st[hi - 1] + max_len + 1;   // upper bound of every end: largest start + longest interval
    m.has = 1;
    span[c] = (unsigned long long)(m.hi_p1 - m.lo_m1 + 1);
  }
  cmap[c] = m;
}
// small contig tables (<= 1024): spans, their exclusive scan (= slice offsets), the total span and the finished
// ContigMap in ONE single-block launch (instead of span kernel + scan + copy + offset kernel)
__global__ void __launch_bounds__(1024) contig_layout_kernel(const int32_t *__restrict__ seg, const int32_t *__restrict__ st,
                                                             long long max_len, int32_t n_contigs,
                                                             ContigMap *__restrict__ cmap, unsigned long long *meta /*[0] nested flag (in), [1] total span (out)*/,
                                                             volatile unsigned long long *mailbox, unsigned long long mailbox_seq) {
  __shared__ unsigned long long wt[1024 / 32 + 1];
  const int c = threadIdx.x;
  ContigMap m;
  m.off = 0; m.lo_m1 = 0; m.hi_p1 = 0; m.has = 0;
  unsigned long long span = 0;
  if (c < n_contigs) {
    const int32_t lo = seg[c], hi = seg[c + 1];
    if (lo < hi) {
      m.lo_m1 = (long long)st[lo] - 1;
      m.hi_p1 = (long long)st[hi - 1] + max_len + 1;
      m.has = 1;
      span = (unsigned long long)(m.hi_p1 - m.lo_m1 + 1);
    }
  }
  const unsigned long long off = block_exclusive<SumU64, 1024>(span, wt);
  if (c < n_contigs) { m.off = (uint32_t)off; cmap[c] = m; }  // off is only used when the total fits 32 bits
  if (threadIdx.x == 0) {
    const unsigned long long total = wt[1024 / 32];
    meta[1] = total;
    if (mailbox) {  // both meta words straight to the host mailbox (pbgpu.cu)
      mailbox[0] = meta[0];
      mailbox[1] = total;
      __threadfence_system();
      mailbox[15] = mailbox_seq;
    }
  }
}
__global__ void __launch_bounds__(128) contig_off_kernel(const unsigned long long *__restrict__ off, int32_t n_contigs,
                                                         ContigMap *__restrict__ cmap) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_contigs) cmap[c].off = (uint32_t)off[c];
}
// global-axis coordinate of every sorted key (contig taken from the packed sort key)
__global__ void __launch_bounds__(256) global_coord_kernel(const uint64_t *__restrict__ keys, int pos_bits, const int32_t *__restrict__ pos,
                                                           int64_t m, const ContigMap *__restrict__ cmap, uint32_t *__restrict__ g) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const ContigMap cm = cmap[keys[i] >> pos_bits];
  g[i] = cm.off + (uint32_t)((long long)pos[i] - cm.lo_m1);
}
// first index in [lo,hi) with g[idx] >= x (x may be 2^32: 64-bit compare)
__device__ __forceinline__ uint32_t lower_bound_g(const uint32_t *__restrict__ g, uint32_t lo, uint32_t hi, uint64_t x) {
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if ((uint64_t)__ldg(g + mid) < x) lo = mid + 1; else hi = mid; }
  return lo;
}
// one thread per bucket: ranks of its first start / first end by binary search (neighbouring buckets share their
// search paths, so the loads coalesce), then the keys of its window packed into the record (JRec above)
__global__ void __launch_bounds__(256) build_jdir_kernel(const uint32_t *__restrict__ gs, const uint32_t *__restrict__ ge, int64_t m,
                                                         int shift, uint32_t n_buckets, JRec *__restrict__ dir) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_buckets) return;  // record n_buckets: every key is below it (only reached by thresholds == its lo)
  const uint64_t lo = (uint64_t)b << shift, W = 1ull << shift;  // the axis is shorter than 2^32 on this path
  const uint32_t mm = (uint32_t)m;
  uint32_t ls = 0, hs = mm, le = 0, he = mm;  // two interleaved searches: their loads overlap
  while (ls < hs || le < he) {
    if (ls < hs) { const uint32_t mid = ls + ((hs - ls) >> 1); if ((uint64_t)__ldg(gs + mid) < lo) ls = mid + 1; else hs = mid; }
    if (le < he) { const uint32_t mid = le + ((he - le) >> 1); if ((uint64_t)__ldg(ge + mid) < lo) le = mid + 1; else he = mid; }
  }
  const uint32_t base_s = ls, base_e = le;
  uint32_t ns = 0, ne = 0;  // keys of the window, counted up to one past the capacity
  while (ns <= (uint32_t)kJKeys && base_s + ns < mm && (uint64_t)__ldg(gs + base_s + ns) < lo + 2 * W) ++ns;
  while (ns + ne <= (uint32_t)kJKeys && base_e + ne < mm && (uint64_t)__ldg(ge + base_e + ne) < lo + W) ++ne;
  JRec r;
  if (ns + ne > (uint32_t)kJKeys) {  // crowded: keep the rank ranges
    r.w[0] = base_s | 0x80000000u;
    r.w[1] = base_e;
    r.w[2] = lower_bound_g(gs, base_s + ns, mm, lo + 2 * W);
    r.w[3] = lower_bound_g(ge, base_e + ne, mm, lo + W);
    r.w[4] = r.w[5] = r.w[6] = r.w[7] = 0x7FFF7FFFu;
  } else {
    r.w[0] = base_s;
    r.w[1] = base_e - ns;
    const uint32_t lo32 = (uint32_t)lo;
#pragma unroll
    for (int k = 0; k < kJKeys / 2; ++k) {
      uint32_t f[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t slot = 2 * k + h;
        uint32_t v = 0x7FFFu;
        if (slot < ns) v = __ldg(gs + base_s + slot) - lo32;
        else if (slot < ns + ne) v = 0x4000u | (__ldg(ge + base_e + (slot - ns)) - lo32);
        f[h] = v;
      }
      r.w[2 + k] = f[0] | (f[1] << 16);
    }
  }
  dir[b] = r;
}

// ---- streaming construction of the joint directory (default; PBGPU_JDIR=search keeps the kernel above) ---------
// The sorted global-axis arrays already hold every rank: #{g < b*W} is the position of the first entry whose bucket is
// >= b.  jdir_mark_kernel: one thread per sorted position i computes gs[i], ge[i] (what global_coord_kernel did) and,
// where the bucket number steps up between i-1 and i, writes i into rank_s[b] / rank_e[b] (compact scratch arrays:
// neighbouring positions write neighbouring buckets) for every b in (bucket(i-1), bucket(i)]; position m-1 also closes
// the tail (bucket(m-1), n_buckets] with m.  Total writes = number of records, O(m) loads -- no binary search
// (build_jdir_kernel: 2 x log2(m) dependent loads per bucket).  Runs longer than 8 buckets are filled by the whole
// warp so one large gap cannot serialise a thread.
__device__ __forceinline__ void jdir_fill_run(uint32_t *__restrict__ rank, long long first, long long last_incl,
                                              uint32_t val, bool active) {
  const int lane = threadIdx.x & 31;
  const long long len = active ? last_incl - first + 1 : 0;
  const bool big = len > 8;
  if (len > 0 && !big)
    for (long long b = first; b <= last_incl; ++b) rank[b] = val;
  unsigned mask = __ballot_sync(0xffffffffu, big);
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    const long long f = __shfl_sync(0xffffffffu, first, src), l = __shfl_sync(0xffffffffu, last_incl, src);
    const uint32_t v = __shfl_sync(0xffffffffu, val, src);
    for (long long b = f + lane; b <= l; b += 32) rank[b] = v;
  }
}
__device__ __forceinline__ uint32_t global_coord_of(const uint64_t *__restrict__ keys, int pos_bits, const int32_t *__restrict__ pos,
                                                    int64_t i, const ContigMap *__restrict__ cmap) {
  const ContigMap cm = cmap[keys[i] >> pos_bits];
  return cm.off + (uint32_t)((long long)__ldg(pos + i) - cm.lo_m1);
}
__global__ void __launch_bounds__(256) jdir_mark_kernel(const uint64_t *__restrict__ skeys, const uint64_t *__restrict__ ekeys, int pos_bits,
                                                        const int32_t *__restrict__ st, const int32_t *__restrict__ en_sorted, int64_t m,
                                                        const ContigMap *__restrict__ cmap, int shift, uint32_t n_buckets,
                                                        uint32_t *__restrict__ gs, uint32_t *__restrict__ ge,
                                                        uint32_t *__restrict__ rank_s, uint32_t *__restrict__ rank_e) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // no early return: the warp fills runs together
  const bool ok = i < m;
  long long s_prev = -1, s_cur = -1, e_prev = -1, e_cur = -1;
  if (ok) {
    const uint32_t g_s = global_coord_of(skeys, pos_bits, st, i, cmap), g_e = global_coord_of(ekeys, pos_bits, en_sorted, i, cmap);
    gs[i] = g_s;
    ge[i] = g_e;
    s_cur = (long long)(g_s >> shift);
    e_cur = (long long)(g_e >> shift);
    if (i > 0) {
      s_prev = (long long)(global_coord_of(skeys, pos_bits, st, i - 1, cmap) >> shift);
      e_prev = (long long)(global_coord_of(ekeys, pos_bits, en_sorted, i - 1, cmap) >> shift);
    }
  }
  jdir_fill_run(rank_s, s_prev + 1, s_cur, (uint32_t)i, ok);
  jdir_fill_run(rank_e, e_prev + 1, e_cur, (uint32_t)i, ok);
  const bool last = i == m - 1;
  jdir_fill_run(rank_s, s_cur + 1, (long long)n_buckets, (uint32_t)m, last);
  jdir_fill_run(rank_e, e_cur + 1, (long long)n_buckets, (uint32_t)m, last);
}
// one thread per record: ranks from jdir_mark_kernel; pack the keys of its window (same record layout and the same
// crowded rule as build_jdir_kernel).
__global__ void __launch_bounds__(256) jdir_pack_kernel(const uint32_t *__restrict__ gs, const uint32_t *__restrict__ ge, int64_t m,
                                                        int shift, uint32_t n_buckets, const uint32_t *__restrict__ rank_s,
                                                        const uint32_t *__restrict__ rank_e, JRec *__restrict__ dir) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_buckets) return;
  const uint64_t lo = (uint64_t)b << shift, W = 1ull << shift;
  const uint32_t mm = (uint32_t)m;
  const uint32_t base_s = rank_s[b], base_e = rank_e[b];
  // the first keys of both windows are fetched at once (a record holds ~3 on average): one round trip instead of a
  // dependent load per key; only fuller records go on one key at a time
  constexpr int kAhead = 4;
  uint32_t fs[kAhead], fe[kAhead];
#pragma unroll
  for (int k = 0; k < kAhead; ++k) {
    fs[k] = base_s + k < mm ? __ldg(gs + base_s + k) : 0xFFFFFFFFu;
    fe[k] = base_e + k < mm ? __ldg(ge + base_e + k) : 0xFFFFFFFFu;
  }
  const uint32_t lo32 = (uint32_t)lo;
  uint32_t ns = 0, ne = 0;
  bool more = true;
#pragma unroll
  for (int k = 0; k < kAhead; ++k) {
    if (more && base_s + k < mm && (uint64_t)fs[k] < lo + 2 * W) ++ns; else more = false;
  }
  if (more)  // all prefetched starts are inside: continue one by one (counted up to one past the capacity)
    while (ns <= (uint32_t)kJKeys && base_s + ns < mm && (uint64_t)__ldg(gs + base_s + ns) < lo + 2 * W) ++ns;
  more = true;
#pragma unroll
  for (int k = 0; k < kAhead; ++k) {
    if (more && base_e + k < mm && (uint64_t)fe[k] < lo + W) ++ne; else more = false;
  }
  if (more)
    while (ns + ne <= (uint32_t)kJKeys && base_e + ne < mm && (uint64_t)__ldg(ge + base_e + ne) < lo + W) ++ne;
  JRec r;
  if (ns + ne > (uint32_t)kJKeys) {  // crowded: keep the rank ranges
    r.w[0] = base_s | 0x80000000u;
    r.w[1] = base_e;
    r.w[2] = lower_bound_g(gs, base_s + ns, mm, lo + 2 * W);
    r.w[3] = lower_bound_g(ge, base_e + ne, mm, lo + W);
    r.w[4] = r.w[5] = r.w[6] = r.w[7] = 0x7FFF7FFFu;
  } else {
    r.w[0] = base_s;
    r.w[1] = base_e - ns;
#pragma unroll
    for (int k = 0; k < kJKeys / 2; ++k) {
      uint32_t f[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t slot = 2 * k + h;
        uint32_t v = 0x7FFFu;
        if (slot < ns) v = (slot < (uint32_t)kAhead ? fs[slot % kAhead] : __ldg(gs + base_s + slot)) - lo32;  // slot is a constant here
        else if (slot < ns + ne) {
          const uint32_t j = slot - ns;
          uint32_t e = 0;
          bool got = false;
#pragma unroll
          for (int a = 0; a < kAhead; ++a) if (j == (uint32_t)a) { e = fe[a]; got = true; }
          if (!got) e = __ldg(ge + base_e + j);
          v = 0x4000u | (e - lo32);
        }
        f[h] = v;
      }
      r.w[2 + k] = f[0] | (f[1] << 16);
    }
  }
  dir[b] = r;
}

// ---- round 2: no end sort on the fast path -----------------------------------------------------------------------
// contig of sorted position i: the largest c with seg[c] <= i < seg[c+1] (empty contigs have seg[c] == seg[c+1])
__device__ __forceinline__ int32_t contig_of_pos(const int32_t *__restrict__ seg, int32_t n_contigs, int64_t i) {
  int32_t lo = 0, hi = n_contigs;  // first c with seg[c+1] > i
  while (lo < hi) { const int32_t mid = lo + ((hi - lo) >> 1); if ((int64_t)__ldg(seg + mid + 1) <= i) lo = mid + 1; else hi = mid; }
  return lo;
}

// Running maximum of `en` inside every contig in ONE pass (decoupled look-back over 2048-row tiles; replaces key
// construction + three-kernel scan + unpack: 56 -> 8 bytes per row).  The scanned value is contig << 32 | biased end:
// the contig never decreases along the sorted rows, so the running maximum of that word restarts by itself at every
// contig boundary.  Status word: 2 flag bits | 62-bit value (contig codes are below 2^30).
constexpr int kPmThreads = 256, kPmItems = 8, kPmTile = kPmThreads * kPmItems;
__global__ void __launch_bounds__(kPmThreads) pmax_lookback_kernel(const int32_t *__restrict__ seg, int32_t n_contigs,
                                                                   const int32_t *__restrict__ en, int64_t m, int32_t *__restrict__ pmax,
                                                                   unsigned long long *status /*[tiles], zeroed*/, unsigned int *ticket /*zeroed*/) {
  constexpr unsigned long long kAgg = 1ull << 62, kPre = 2ull << 62, kMask = (1ull << 62) - 1ull;
  __shared__ unsigned long long wt[kPmThreads / 32 + 1];
  __shared__ unsigned long long prefix_s;
  __shared__ unsigned int tile_s;
  if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned int tile = tile_s;
  const int64_t i0 = (int64_t)tile * kPmTile + (int64_t)threadIdx.x * kPmItems;
  unsigned long long v[kPmItems];
  unsigned long long acc = 0ull;
  if (i0 < m) {
    int32_t c = contig_of_pos(seg, n_contigs, i0);
    int64_t next = __ldg(seg + c + 1);
    int32_t e[kPmItems];
    if (i0 + kPmItems <= m) {
      const int4 a = *reinterpret_cast<const int4 *>(en + i0), b = *reinterpret_cast<const int4 *>(en + i0 + 4);
      e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < kPmItems; ++j) e[j] = i0 + j < m ? en[i0 + j] : INT32_MIN;
    }
#pragma unroll
    for (int j = 0; j < kPmItems; ++j) {
      const int64_t i = i0 + j;
      while (i >= next && c + 1 < n_contigs) { ++c; next = __ldg(seg + c + 1); }
      const unsigned long long w = i < m ? (((unsigned long long)(uint32_t)c << 32) | ((uint32_t)e[j] ^ 0x80000000u)) : 0ull;
      acc = acc > w ? acc : w;
      v[j] = acc;  // inclusive running max inside the thread
    }
  } else {
#pragma unroll
    for (int j = 0; j < kPmItems; ++j) v[j] = 0ull;
  }
  const unsigned long long excl_thread = block_exclusive<MaxU64, kPmThreads>(acc, wt);
  const unsigned long long total = wt[kPmThreads / 32];
  if (threadIdx.x < 32) {  // warp 0: publish, look back
    const int lane = threadIdx.x;
    unsigned long long excl = 0ull;
    if (tile == 0) {
      if (lane == 0) st_volatile_u64(status, total | kPre);
    } else {
      if (lane == 0) st_volatile_u64(status + tile, total | kAgg);
      for (long long top = (long long)tile - 1;; top -= 32) {
        const long long idx = top - lane;
        unsigned long long w = kPre;  // below tile 0: an empty prefix ends the walk
        if (idx >= 0) { do { w = ld_volatile_u64(status + idx); } while ((w >> 62) == 0ull); }
        const unsigned pre = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
        const int first = pre ? __ffs(pre) - 1 : 31;
        unsigned long long x = lane <= first ? (w & kMask) : 0ull;
#pragma unroll
        for (int d = 16; d; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, x, d); x = x > o ? x : o; }
        excl = excl > x ? excl : x;
        if (pre) break;
      }
      if (lane == 0) st_volatile_u64(status + tile, (excl > total ? excl : total) | kPre);
    }
    if (lane == 0) prefix_s = excl;
  }
  __syncthreads();
  if (i0 >= m) return;
  unsigned long long before = prefix_s;
  before = before > excl_thread ? before : excl_thread;
  int32_t o[kPmItems];
#pragma unroll
  for (int j = 0; j < kPmItems; ++j) {
    const unsigned long long w = v[j] > before ? v[j] : before;
    o[j] = (int32_t)((uint32_t)w ^ 0x80000000u);
  }
  if (i0 + kPmItems <= m) {
    *reinterpret_cast<int4 *>(pmax + i0) = make_int4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<int4 *>(pmax + i0 + 4) = make_int4(o[4], o[5], o[6], o[7]);
  } else {
#pragma unroll
    for (int j = 0; j < kPmItems; ++j) if (i0 + j < m) pmax[i0 + j] = o[j];
  }
}

// Directory over an index with nested intervals WITHOUT sorting the ends: the record of bucket b needs (1) the number
// of ends below the bucket (rank), (2) the ends inside it -- as a SET: the packed compare counts fields below a
// threshold whatever their order.  So the ends only have to be grouped by bucket: counting sort with the buckets as
// bins.  jdir_mark_nested_kernel: global coordinates of both ends of every row (gs in start order, the end's coordinate
// into a scratch column), start ranks of the buckets by run filling (as jdir_mark_kernel), and one atomic per row on
// the bucket counter of its end (rows arrive in start order, so the counters touched by a warp are neighbours);
// an exclusive scan turns the counters into cursors; jdir_place_ends_kernel drops every end at its bucket's cursor.
// Afterwards cursor[b] = number of ends below bucket b+1.
__global__ void __launch_bounds__(256) jdir_mark_nested_kernel(const uint64_t *__restrict__ skeys, int pos_bits, const int32_t *__restrict__ st,
                                                               const int32_t *__restrict__ en, int64_t m, const ContigMap *__restrict__ cmap,
                                                               int shift, uint32_t n_buckets, uint32_t *__restrict__ gs,
                                                               uint32_t *__restrict__ ge_tmp, uint32_t *__restrict__ rank_s,
                                                               uint32_t *__restrict__ ecnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // no early return: the warp fills runs together
  const bool ok = i < m;
  long long s_prev = -1, s_cur = -1;
  if (ok) {
    const ContigMap cm = cmap[skeys[i] >> pos_bits];
    const uint32_t g_s = cm.off + (uint32_t)((long long)__ldg(st + i) - cm.lo_m1);
    const uint32_t g_e = cm.off + (uint32_t)((long long)__ldg(en + i) - cm.lo_m1);
    gs[i] = g_s;
    ge_tmp[i] = g_e;
    atomicAdd(ecnt + (g_e >> shift), 1u);
    s_cur = (long long)(g_s >> shift);
    if (i > 0) s_prev = (long long)(global_coord_of(skeys, pos_bits, st, i - 1, cmap) >> shift);
  }
  jdir_fill_run(rank_s, s_prev + 1, s_cur, (uint32_t)i, ok);
  jdir_fill_run(rank_s, s_cur + 1, (long long)n_buckets, (uint32_t)m, i == m - 1);
}
__global__ void __launch_bounds__(256) jdir_place_ends_kernel(const uint32_t *__restrict__ ge_tmp, int64_t m, int shift,
                                                              uint32_t *__restrict__ cursor, uint32_t *__restrict__ ge) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t g = ge_tmp[i];
  ge[atomicAdd(cursor + (g >> shift), 1u)] = g;
}
// One thread per record, both kinds of index.  rank_s[b] = starts below bucket b; the ends of bucket b are
// ge[e_lo, e_hi) with  e_lo = ENDS_BY_CURSOR ? cursor[b-1] : rank_e[b],  e_hi = ENDS_BY_CURSOR ? cursor[b] : rank_e[b+1]
// (rank arrays have n_buckets + 1 entries; the window of the last records is clamped).  No counting loops: the window
// sizes are rank differences.  max_crowded_ends: largest number of ends in a crowded record (the nested build sorts
// nothing inside a bucket; probes count linearly there, so the host wants to know how bad it can get).
template <bool ENDS_BY_CURSOR>
__global__ void __launch_bounds__(256) jdir_pack2_kernel(const uint32_t *__restrict__ gs, const uint32_t *__restrict__ ge, int64_t m,
                                                         int shift, uint32_t n_buckets, const uint32_t *__restrict__ rank_s,
                                                         const uint32_t *__restrict__ rank_e, JRec *__restrict__ dir,
                                                         unsigned int *__restrict__ max_crowded_ends) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_buckets) return;
  const uint32_t mm = (uint32_t)m;
  const uint32_t base_s = rank_s[b];
  const uint32_t top_s = b + 2 <= n_buckets ? rank_s[b + 2] : mm;
  uint32_t base_e, top_e;
  if (ENDS_BY_CURSOR) { base_e = b ? rank_e[b - 1] : 0u; top_e = rank_e[b]; }
  else { base_e = rank_e[b]; top_e = b + 1 <= n_buckets ? rank_e[b + 1] : mm; }
  const uint32_t ns = top_s - base_s, ne = top_e - base_e;
  const uint32_t lo32 = b << shift;
  JRec r;
  if (ns + ne > (uint32_t)kJKeys) {  // crowded: keep the rank ranges
    r.w[0] = base_s | 0x80000000u;
    r.w[1] = base_e;
    r.w[2] = top_s;
    r.w[3] = top_e;
    r.w[4] = r.w[5] = r.w[6] = r.w[7] = 0x7FFF7FFFu;
    if (max_crowded_ends && ne > 1) atomicMax(max_crowded_ends, ne);
  } else {
    r.w[0] = base_s;
    r.w[1] = base_e - ns;
    uint32_t f[kJKeys];
#pragma unroll
    for (int k = 0; k < kJKeys; ++k) {
      uint32_t v = 0x7FFFu;
      if ((uint32_t)k < ns) v = __ldg(gs + base_s + k) - lo32;
      else if ((uint32_t)k < ns + ne) v = 0x4000u | (__ldg(ge + base_e + ((uint32_t)k - ns)) - lo32);
      f[k] = v;
    }
#pragma unroll
    for (int k = 0; k < kJKeys / 2; ++k) r.w[2 + k] = f[2 * k] | (f[2 * k + 1] << 16);
  }
  dir[b] = r;
}

// (contig | end) sort keys + start-order positions for the LAZY end order (nearest, generic kernels); the contig comes
// from the segment table
__global__ void __launch_bounds__(256) make_end_keys_seg_kernel(const int32_t *__restrict__ seg, int32_t n_contigs, const int32_t *__restrict__ en,
                                                                int64_t m, int pos_bits, uint32_t bias_e, uint64_t *__restrict__ ekeys,
                                                                uint64_t *__restrict__ evals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t contig = (uint64_t)contig_of_pos(seg, n_contigs, i);
  ekeys[i] = (contig << pos_bits) | ((uint32_t)en[i] ^ bias_e);
  evals[i] = (uint64_t)i;
}

// ---- round 2: global-key build (fast path; <= 1024 contigs, no inverted rows, axis < 2^32) ----------------------------
// The generic build sorts 16-byte (contig | start, end | row) pairs over every varying digit of the 37..64-bit key:
// five passes for the human genome (four start bytes + the contig byte).  When the contig slices of the global axis are
// known BEFORE the sort, the key is the 32-bit global start: four passes whatever the number of contigs, the row id
// rides in the low half of the 64-bit key and the end travels as a 4-byte value -- 12 bytes per row and pass instead
// of 16, and the sorted keys ARE the gs column.  Price: one extra streaming pass over the input for the per-contig
// coordinate range (the slices), before the keys can be formed.
struct GStats {
  unsigned long long valid, inverted, max_len;
  long long min_end, max_end;   // as 64-bit so that atomics on them are plain 64-bit min / max
  unsigned long long blocks_done;
};
__global__ void __launch_bounds__(256) gstats_init_kernel(int32_t *__restrict__ cmin, int32_t *__restrict__ cmax, int32_t n_contigs, GStats *st) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < n_contigs) { cmin[c] = INT32_MAX; cmax[c] = INT32_MIN; }
  if (c == 0) { st->valid = 0; st->inverted = 0; st->max_len = 0; st->min_end = INT32_MAX; st->max_end = INT32_MIN; st->blocks_done = 0; }
}
// per contig: smallest and largest start; overall: valid / inverted rows, longest interval, range of the ends.
// A row touches its contig's shared-memory slots with an atomic only when it would improve them (a plain load decides):
// after the first few hundred rows of a block almost none does.  (First version: MATCH + REDUX per row, 1.3 ms per
// 90 M rows; the loads are conflict-free broadcasts.)
__global__ void __launch_bounds__(512) contig_stats_kernel(const int32_t *__restrict__ c, const int32_t *__restrict__ s,
                                                           const int32_t *__restrict__ e, int64_t n, int32_t n_contigs,
                                                           int32_t *__restrict__ cmin, int32_t *__restrict__ cmax, GStats *st) {
  __shared__ int32_t smin[1024], smax[1024];
  __shared__ unsigned long long sred[3][512 / 32];
  __shared__ int sred_e[2][512 / 32];
  for (int i = threadIdx.x; i < n_contigs; i += 512) { smin[i] = INT32_MAX; smax[i] = INT32_MIN; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned inv = 0, val = 0, mlen = 0;
  int mn_e = INT32_MAX, mx_e = INT32_MIN;
  const int64_t stride = (int64_t)gridDim.x * 512;
  constexpr int U = 4;
  for (int64_t i0 = (int64_t)blockIdx.x * 512 + threadIdx.x; i0 < n; i0 += U * stride) {
    int32_t cc[U], ss[U], ee[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      const bool in = i < n;
      cc[u] = in ? c[i] : -1; ss[u] = in ? s[i] : 0; ee[u] = in ? e[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (cc[u] >= 0 && cc[u] < n_contigs) {
        if (ss[u] < *(volatile int32_t *)&smin[cc[u]]) atomicMin(&smin[cc[u]], ss[u]);
        if (ss[u] > *(volatile int32_t *)&smax[cc[u]]) atomicMax(&smax[cc[u]], ss[u]);
        mn_e = min(mn_e, ee[u]); mx_e = max(mx_e, ee[u]);
        inv += ss[u] > ee[u];
        if (ee[u] >= ss[u]) mlen = max(mlen, (unsigned)((long long)ee[u] - (long long)ss[u]));
        ++val;
      }
    }
  }
  inv = __reduce_add_sync(0xffffffffu, inv); val = __reduce_add_sync(0xffffffffu, val); mlen = __reduce_max_sync(0xffffffffu, mlen);
  mn_e = __reduce_min_sync(0xffffffffu, mn_e); mx_e = __reduce_max_sync(0xffffffffu, mx_e);
  const int w = threadIdx.x >> 5;
  if (lane == 0) { sred[0][w] = inv; sred[1][w] = val; sred[2][w] = mlen; sred_e[0][w] = mn_e; sred_e[1][w] = mx_e; }
  __syncthreads();
  for (int i = threadIdx.x; i < n_contigs; i += 512) {
    if (smin[i] != INT32_MAX) atomicMin(cmin + i, smin[i]);
    if (smax[i] != INT32_MIN) atomicMax(cmax + i, smax[i]);
  }
  if (threadIdx.x == 0) {
    unsigned long long a = 0, b = 0, m2 = 0;
    int e0 = INT32_MAX, e1 = INT32_MIN;
    for (int k = 0; k < 512 / 32; ++k) { a += sred[0][k]; b += sred[1][k]; m2 = max(m2, sred[2][k]); e0 = min(e0, sred_e[0][k]); e1 = max(e1, sred_e[1][k]); }
    if (b) {
      if (a) atomicAdd(&st->inverted, a);
      atomicAdd(&st->valid, b);
      atomicMax(&st->max_len, m2);
      atomicMin(&st->min_end, (long long)e0);
      atomicMax(&st->max_end, (long long)e1);
    }
  }
}
// slices of the global axis from the per-contig start ranges (one block, <= 1024 contigs): ContigMap, total span, and
// the six statistics words straight to the host mailbox (valid, inverted, max_len, min_end, max_end, total span)
__global__ void __launch_bounds__(1024) contig_layout_mm_kernel(const int32_t *__restrict__ cmin, const int32_t *__restrict__ cmax,
                                                                const GStats *__restrict__ st, int32_t n_contigs,
                                                                ContigMap *__restrict__ cmap, unsigned long long *__restrict__ d_out /*[6]*/,
                                                                volatile unsigned long long *mailbox, unsigned long long mailbox_seq) {
  __shared__ unsigned long long wt[1024 / 32 + 1];
  const int c = threadIdx.x;
  const long long max_len = (long long)st->max_len;
  ContigMap m;
  m.off = 0; m.lo_m1 = 0; m.hi_p1 = 0; m.has = 0;
  unsigned long long span = 0;
  if (c < n_contigs && cmin[c] != INT32_MAX) {
    m.lo_m1 = (long long)cmin[c] - 1;
    m.hi_p1 = (long long)cmax[c] + max_len + 1;
    m.has = 1;
    span = (unsigned long long)(m.hi_p1 - m.lo_m1 + 1);
  }
  const unsigned long long off = block_exclusive<SumU64, 1024>(span, wt);
  if (c < n_contigs) { m.off = (uint32_t)off; cmap[c] = m; }  // off is only used when the total fits 32 bits
  if (threadIdx.x == 0) {
    const unsigned long long w[6] = {st->valid, st->inverted, st->max_len, (unsigned long long)st->min_end, (unsigned long long)st->max_end,
                                     wt[1024 / 32]};
    for (int k = 0; k < 6; ++k) d_out[k] = w[k];
    if (mailbox) {
      for (int k = 0; k < 6; ++k) mailbox[k] = w[k];
      __threadfence_system();
      mailbox[15] = mailbox_seq;
    }
  }
}
// key = global start << 32 | row, value = end; digit totals of key bytes 4..7.  Null-keyed rows (and rows of contigs
// beyond the table) get global start 0xFFFFFFFF: they sort behind every valid row and are cut off.
__global__ void __launch_bounds__(kPrepThreads) gkeys_kernel(const int32_t *__restrict__ c, const int32_t *__restrict__ s,
                                                             const int32_t *__restrict__ e, int64_t n, int32_t n_contigs,
                                                             const ContigMap *__restrict__ cmap, uint64_t *__restrict__ keys,
                                                             uint32_t *__restrict__ vals, uint32_t *__restrict__ digit_totals /*[kRsMaxPasses][256], zeroed*/,
                                                             const uint32_t *__restrict__ row_ids /*NULL: the row's position*/) {
  __shared__ uint32_t h[4][kRsRadix];
  for (int i = threadIdx.x; i < 4 * kRsRadix; i += kPrepThreads) (&h[0][0])[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * kPrepThreads;
  for (int64_t i0 = (int64_t)blockIdx.x * kPrepThreads + threadIdx.x; i0 < n; i0 += 2 * stride) {
    int32_t cc[2], ss[2], ee[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      const bool in = i < n;
      cc[u] = in ? c[i] : -1; ss[u] = in ? s[i] : 0; ee[u] = in ? e[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n) continue;
      uint32_t g = 0xFFFFFFFFu;
      if (cc[u] >= 0 && cc[u] < n_contigs) {
        const ContigMap cm = cmap[cc[u]];
        g = cm.off + (uint32_t)((long long)ss[u] - cm.lo_m1);
      }
      keys[i] = ((uint64_t)g << 32) | (row_ids ? row_ids[i] : (uint32_t)i);
      vals[i] = (uint32_t)ee[u];
#pragma unroll
      for (int p = 0; p < 4; ++p) atomicAdd(&h[p][(g >> (8 * p)) & 0xff], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * kRsRadix; i += kPrepThreads) {
    const uint32_t v = (&h[0][0])[i];
    if (v) atomicAdd(digit_totals + 4 * kRsRadix + i, v);  // digit positions 4..7 of the 64-bit key
  }
}
// contig of a global coordinate: the last contig whose slice starts at or below it (empty contigs have empty slices)
__device__ __forceinline__ int32_t contig_of_g(const ContigMap *__restrict__ cmap, int32_t n_contigs, uint32_t g) {
  int32_t lo = 0, hi = n_contigs;  // first c with off[c] > g
  while (lo < hi) { const int32_t mid = lo + ((hi - lo) >> 1); if (cmap[mid].off <= g) lo = mid + 1; else hi = mid; }
  return lo - 1;
}
// unpack of the global-key sort: SoA columns, gs, contig segments by boundary detection, "any end inversion?".
// The contig of a row: one search per block (its first row), then a walk forward -- rows are sorted, a block of 256
// almost never spans more than two contigs; the predecessor's contig comes from the neighbouring lane.
__global__ void __launch_bounds__(256) unpack_gsorted_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t m,
                                                             const ContigMap *__restrict__ cmap, int32_t n_contigs,
                                                             int32_t *__restrict__ st, int32_t *__restrict__ en, uint32_t *__restrict__ row,
                                                             uint2 *__restrict__ er, uint32_t *__restrict__ gs, int32_t *__restrict__ seg,
                                                             unsigned long long *__restrict__ inversions) {
  __shared__ int32_t c_first;
  const int64_t b0 = (int64_t)blockIdx.x * blockDim.x;
  const int64_t i = b0 + threadIdx.x;
  if (threadIdx.x == 0) c_first = contig_of_g(cmap, n_contigs, (uint32_t)(keys[b0 > 0 ? b0 - 1 : 0] >> 32));  // contig of the row before the block
  __syncthreads();
  unsigned inv = 0;
  const bool ok = i < m;
  uint64_t k = 0;
  uint32_t e = 0;
  int32_t contig = c_first;
  if (ok) {
    k = keys[i];
    e = vals[i];
    const uint32_t g = (uint32_t)(k >> 32);
    while (contig + 1 < n_contigs && cmap[contig + 1].off <= g) ++contig;
  }
  // predecessor: the lane below; lane 0 reads it (the row before the block has contig c_first by construction)
  int32_t prev = __shfl_up_sync(0xffffffffu, contig, 1);
  uint32_t prev_e = __shfl_up_sync(0xffffffffu, e, 1);
  if ((threadIdx.x & 31) == 0 && ok) {
    if (i == 0) prev = -1;
    else {
      prev_e = vals[i - 1];
      prev = c_first;
      const uint32_t pg = (uint32_t)(keys[i - 1] >> 32);
      while (prev + 1 < n_contigs && cmap[prev + 1].off <= pg) ++prev;
    }
  }
  if (ok) {
    const uint32_t g = (uint32_t)(k >> 32);
    const ContigMap cm = cmap[contig];
    st[i] = (int32_t)((long long)(g - cm.off) + cm.lo_m1);
    en[i] = (int32_t)e;
    row[i] = (uint32_t)k;
    er[i] = make_uint2(e, (uint32_t)k);
    gs[i] = g;
    if (prev == contig) inv = (int32_t)e < (int32_t)prev_e;
    for (int32_t c = prev + 1; c <= contig; ++c) seg[c] = (int32_t)i;
    if (i == m - 1) for (int32_t c = contig + 1; c <= n_contigs; ++c) seg[c] = (int32_t)m;
  }
  const int any = __syncthreads_or((int)inv);
  if (any && threadIdx.x == 0 && *(volatile unsigned long long *)inversions == 0ull) atomicMax(inversions, 1ull);
}
// gs of the generic build's sorted rows (contig from the packed sort key)
__global__ void __launch_bounds__(256) gs_from_keys_kernel(const uint64_t *__restrict__ keys, int pos_bits, const int32_t *__restrict__ st, int64_t m,
                                                           const ContigMap *__restrict__ cmap, uint32_t *__restrict__ gs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) gs[i] = global_coord_of(keys, pos_bits, st, i, cmap);
}
// Directory marks from the gs column: no contig lookups -- a row's end sits (en - st) past its start on the global axis.
// NESTED = false: ends ascend in start order; ge and both rank arrays by run filling.
// NESTED = true : ge_tmp + one atomic per row on the bucket counter of its end (see jdir_mark_nested_kernel).
template <bool NESTED>
__global__ void __launch_bounds__(256) jdir_mark_g_kernel(const uint32_t *__restrict__ gs, const int32_t *__restrict__ st,
                                                          const int32_t *__restrict__ en, int64_t m, int shift, uint32_t n_buckets,
                                                          uint32_t *__restrict__ ge_out, uint32_t *__restrict__ rank_s,
                                                          uint32_t *__restrict__ rank_e_or_cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // no early return: the warp fills runs together
  const bool ok = i < m;
  long long s_prev = -1, s_cur = -1, e_prev = -1, e_cur = -1;
  if (ok) {
    const uint32_t g_s = gs[i];
    const uint32_t g_e = g_s + (uint32_t)(en[i] - st[i]);
    ge_out[i] = g_e;
    s_cur = (long long)(g_s >> shift);
    e_cur = (long long)(g_e >> shift);
    if (NESTED) atomicAdd(rank_e_or_cnt + (g_e >> shift), 1u);
    if (i > 0) {
      const uint32_t p_s = gs[i - 1];
      s_prev = (long long)(p_s >> shift);
      if (!NESTED) e_prev = (long long)((p_s + (uint32_t)(en[i - 1] - st[i - 1])) >> shift);
    }
  }
  const bool last = i == m - 1;
  jdir_fill_run(rank_s, s_prev + 1, s_cur, (uint32_t)i, ok);
  jdir_fill_run(rank_s, s_cur + 1, (long long)n_buckets, (uint32_t)m, last);
  if (!NESTED) {
    jdir_fill_run(rank_e_or_cnt, e_prev + 1, e_cur, (uint32_t)i, ok);
    jdir_fill_run(rank_e_or_cnt, e_cur + 1, (long long)n_buckets, (uint32_t)m, last);
  }
}

// Running maximum of `en` inside every contig, two levels (the single look-back pass of the first round-2 version spent
// 68 % of its time in barriers: with 2048-row tiles a tile's walk over its ~1000 resident predecessors costs ten times
// its own work).  Level 1: maximum of every tile; a scan of the 44 K tile maxima; level 2: the tile's own running maximum
// on top of its prefix.  The scanned word is contig << 32 | biased end: it restarts by itself at contig boundaries.
__device__ __forceinline__ void pmax_tile_words(const int32_t *__restrict__ seg, int32_t n_contigs, const int32_t *__restrict__ en,
                                                int64_t m, int64_t i0, unsigned long long (&v)[kPmItems], unsigned long long &acc) {
  acc = 0ull;
  if (i0 < m) {
    int32_t c = contig_of_pos(seg, n_contigs, i0);
    int64_t next = __ldg(seg + c + 1);
    int32_t e[kPmItems];
    if (i0 + kPmItems <= m) {
      const int4 a = *reinterpret_cast<const int4 *>(en + i0), b = *reinterpret_cast<const int4 *>(en + i0 + 4);
      e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < kPmItems; ++j) e[j] = i0 + j < m ? en[i0 + j] : INT32_MIN;
    }
#pragma unroll
    for (int j = 0; j < kPmItems; ++j) {
      const int64_t i = i0 + j;
      while (i >= next && c + 1 < n_contigs) { ++c; next = __ldg(seg + c + 1); }
      const unsigned long long w = i < m ? (((unsigned long long)(uint32_t)c << 32) | ((uint32_t)e[j] ^ 0x80000000u)) : 0ull;
      acc = acc > w ? acc : w;
      v[j] = acc;  // inclusive running max inside the thread
    }
  } else {
#pragma unroll
    for (int j = 0; j < kPmItems; ++j) v[j] = 0ull;
  }
}
__global__ void __launch_bounds__(kPmThreads) pmax_tile_max_kernel(const int32_t *__restrict__ seg, int32_t n_contigs, const int32_t *__restrict__ en,
                                                                   int64_t m, unsigned long long *__restrict__ tile_max) {
  __shared__ unsigned long long wt[kPmThreads / 32 + 1];
  unsigned long long v[kPmItems], acc;
  pmax_tile_words(seg, n_contigs, en, m, (int64_t)blockIdx.x * kPmTile + (int64_t)threadIdx.x * kPmItems, v, acc);
  (void)block_exclusive<MaxU64, kPmThreads>(acc, wt);
  if (threadIdx.x == 0) tile_max[blockIdx.x] = wt[kPmThreads / 32];
}
__global__ void __launch_bounds__(kPmThreads) pmax_tile_final_kernel(const int32_t *__restrict__ seg, int32_t n_contigs, const int32_t *__restrict__ en,
                                                                     int64_t m, const unsigned long long *__restrict__ tile_prefix /*exclusive*/,
                                                                     int32_t *__restrict__ pmax) {
  __shared__ unsigned long long wt[kPmThreads / 32 + 1];
  unsigned long long v[kPmItems], acc;
  const int64_t i0 = (int64_t)blockIdx.x * kPmTile + (int64_t)threadIdx.x * kPmItems;
  pmax_tile_words(seg, n_contigs, en, m, i0, v, acc);
  unsigned long long before = block_exclusive<MaxU64, kPmThreads>(acc, wt);
  const unsigned long long pre = tile_prefix[blockIdx.x];
  before = before > pre ? before : pre;
  if (i0 >= m) return;
  int32_t o[kPmItems];
#pragma unroll
  for (int j = 0; j < kPmItems; ++j) {
    const unsigned long long w = v[j] > before ? v[j] : before;
    o[j] = (int32_t)((uint32_t)w ^ 0x80000000u);
  }
  if (i0 + kPmItems <= m) {
    *reinterpret_cast<int4 *>(pmax + i0) = make_int4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<int4 *>(pmax + i0 + 4) = make_int4(o[4], o[5], o[6], o[7]);
  } else {
#pragma unroll
    for (int j = 0; j < kPmItems; ++j) if (i0 + j < m) pmax[i0 + j] = o[j];
  }
}

static inline int bit_length_u32(uint32_t x) {
  int b = 0;
  while (x) { ++b; x >>= 1; }
  return b;
}

}  // namespace pbgpu
