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
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

struct pbgpu_index {
  int64_t m = 0;  // valid rows
  int64_t m_in = 0;
  int32_t n_contigs = 0;
  int device = 0;
  int has_inverted = 0;  // some indexed row has start > end: rank identity disabled
  int32_t *seg = nullptr;
  int32_t *st = nullptr, *en = nullptr, *pmax = nullptr, *en_sorted = nullptr;
  uint32_t *row = nullptr, *en_pos = nullptr;
  size_t bytes = 0;
};

namespace pbgpu {

struct IndexView {
  const int32_t *__restrict__ seg;
  const int32_t *__restrict__ st;
  const int32_t *__restrict__ en;
  const int32_t *__restrict__ pmax;
  const int32_t *__restrict__ en_sorted;
  const uint32_t *__restrict__ row;
  const uint32_t *__restrict__ en_pos;
  int32_t n_contigs;
  int has_inverted;
};

inline IndexView view_of(const pbgpu_index *ix) {
  return IndexView{ix->seg, ix->st, ix->en, ix->pmax, ix->en_sorted, ix->row, ix->en_pos, ix->n_contigs, ix->has_inverted};
}

struct BuildStats {  // device-side reduction target
  int min_start, max_start, min_end, max_end;
  unsigned long long inverted, valid;
};

__global__ void __launch_bounds__(256) build_stats_kernel(const int32_t *__restrict__ c, const int32_t *__restrict__ s,
                                                          const int32_t *__restrict__ e, int64_t n, int32_t n_contigs,
                                                          BuildStats *st) {
  int mn_s = INT32_MAX, mx_s = INT32_MIN, mn_e = INT32_MAX, mx_e = INT32_MIN;
  unsigned inv = 0, val = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t cc = c[i];
    if (cc < 0 || cc >= n_contigs) continue;
    int32_t ss = s[i], ee = e[i];
    mn_s = min(mn_s, ss); mx_s = max(mx_s, ss);
    mn_e = min(mn_e, ee); mx_e = max(mx_e, ee);
    inv += ss > ee;
    ++val;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    mn_s = min(mn_s, __shfl_xor_sync(0xffffffffu, mn_s, d));
    mx_s = max(mx_s, __shfl_xor_sync(0xffffffffu, mx_s, d));
    mn_e = min(mn_e, __shfl_xor_sync(0xffffffffu, mn_e, d));
    mx_e = max(mx_e, __shfl_xor_sync(0xffffffffu, mx_e, d));
    inv += __shfl_xor_sync(0xffffffffu, inv, d);
    val += __shfl_xor_sync(0xffffffffu, val, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&st->min_start, mn_s); atomicMax(&st->max_start, mx_s);
    atomicMin(&st->min_end, mn_e); atomicMax(&st->max_end, mx_e);
    atomicAdd(&st->inverted, (unsigned long long)inv);
    atomicAdd(&st->valid, (unsigned long long)val);
  }
}

// key = contig << pos_bits | biased start ; value = end << 32 | row.  Null-keyed rows get
// contig = n_contigs so the partition drops them behind the last real contig.
__global__ void __launch_bounds__(256) make_start_keys_kernel(const int32_t *__restrict__ c, const int32_t *__restrict__ s,
                                                              const int32_t *__restrict__ e, int64_t n, int32_t n_contigs,
                                                              int pos_bits, uint32_t bias, uint64_t *__restrict__ keys,
                                                              uint64_t *__restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t cc = c[i];
  bool ok = cc >= 0 && cc < n_contigs;
  uint64_t cpart = ok ? (uint64_t)cc : (uint64_t)n_contigs;
  uint32_t sp = ok ? ((uint32_t)s[i] ^ bias) : 0u;
  keys[i] = (cpart << pos_bits) | sp;
  vals[i] = ((uint64_t)(uint32_t)e[i] << 32) | (uint32_t)i;
}

__global__ void __launch_bounds__(64) find_segments_kernel(const uint64_t *__restrict__ keys, int64_t n, int pos_bits,
                                                           int32_t n_contigs, int32_t *__restrict__ seg) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_contigs) return;
  const uint64_t target = (uint64_t)c << pos_bits;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = lo + ((hi - lo) >> 1);
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  seg[c] = (int32_t)lo;
}

// unpack the start-sorted pairs; also emit the contig-tagged end keys for the running max
// and the (contig | end) keys of the second sort.
__global__ void __launch_bounds__(256) unpack_sorted_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals,
                                                            int64_t m, int pos_bits, uint32_t bias_s, uint32_t bias_e,
                                                            int32_t *__restrict__ st, int32_t *__restrict__ en,
                                                            uint32_t *__restrict__ row, uint64_t *__restrict__ pm_keys,
                                                            uint64_t *__restrict__ ekeys, uint64_t *__restrict__ evals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t k = keys[i], v = vals[i];
  const uint64_t contig = k >> pos_bits;
  const uint32_t sp = (uint32_t)(k & ((1ull << pos_bits) - 1ull));
  const int32_t e = (int32_t)(uint32_t)(v >> 32);
  st[i] = (int32_t)(sp ^ bias_s);
  en[i] = e;
  row[i] = (uint32_t)v;
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

static inline int bit_length_u32(uint32_t x) {
  int b = 0;
  while (x) { ++b; x >>= 1; }
  return b;
}

}  // namespace pbgpu
