// unary.cuh -- the sweeps that reuse the partition + sort of the index build: merge, cluster, subtract (complement is
// subtract with the view table on the left).  Included at the end of pbgpu.cu (it uses index_build_impl, fetch_words
// and the block cache of that file).
//
// Reference: MergeProvider / ClusterProvider / ComplementProvider / SubtractProvider of the un-vendored crate
// datafusion-bio-function-ranges v0.11.0, constructed at /root/reference/src/operation.rs:352-510; behaviour pinned by
// tests/_expected.py:174-181, tests/test_coordinate_system_metadata.py:1032-1054 and
// tests/test_partitioned_range_operation_regressions.py:24-59 (restated in oracle/unary_np.py).
//
//   rows sorted by (contig, start, row) with the running max of the ends  (the index build, sweep_only)
//   fresh[i]  = first row of its contig, or start[i] beyond the reach of everything before it:
//               Strict (0-based half-open): start >= pmax[i-1] + min_dist;  Weak (1-based closed): start > pmax[i-1] + min_dist
//   run id    = inclusive scan of fresh - 1;  a run is a merged interval / a cluster
// The running max of the whole contig can stand in for the running max of the run: a fresh row starts beyond
// everything before it, so whatever reaches a later row of the run belongs to the run itself.
#pragma once
#include "common.cuh"
#include "index.cuh"
#include "scan.cuh"
#include "sweep.cuh"

struct pbgpu_intervals {
  int64_t rows = 0;
  int device = 0;
  void *slab = nullptr;
  int32_t *contig = nullptr;  // merge
  uint32_t *row = nullptr;    // subtract: row of the left table the piece comes from
  int32_t *start = nullptr, *end = nullptr;
  int64_t *count = nullptr;   // merge: n_intervals
};

namespace pbgpu {

// contig of sorted position i: the largest c with seg[c] <= i (seg ascends, seg[n_contigs] = m > i; empty contigs
// repeat the value of the next non-empty one, which is the larger index)
__device__ __forceinline__ int32_t contig_of_pos(const int32_t *__restrict__ seg, int32_t n_contigs, int32_t i) {
  int32_t lo = 0, hi = n_contigs;
  while (hi - lo > 1) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(seg + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

template <bool STRICT>
__global__ void __launch_bounds__(256) sweep_fresh_kernel(const int32_t *__restrict__ seg, const int32_t *__restrict__ st,
                                                          const int32_t *__restrict__ pmax, int64_t m, int32_t n_contigs,
                                                          long long min_dist, unsigned long long *__restrict__ fresh) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int32_t c = contig_of_pos(seg, n_contigs, (int32_t)i);
  bool f = (int32_t)i == __ldg(seg + c);
  if (!f) {
    const long long reach = (long long)pmax[i - 1] + min_dist, s = st[i];
    f = STRICT ? (s >= reach) : (s > reach);
  }
  fresh[i] = f ? 1ull : 0ull;
}

// first position in [lo, m) whose inclusive run count reaches v (incl is non-decreasing)
__device__ __forceinline__ int64_t first_reaching(const unsigned long long *__restrict__ incl, int64_t lo, int64_t m, unsigned long long v) {
  int64_t hi = m;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(incl + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// one thread per sorted row; the first row of every run writes the run (it finds the run's last row by a search
// over the scanned flags).  own_end: the runs' ends are finished by run_end_max_kernel (inverted rows present).
__global__ void __launch_bounds__(256) runs_emit_kernel(const int32_t *__restrict__ seg, const int32_t *__restrict__ st,
                                                        const int32_t *__restrict__ en, const int32_t *__restrict__ pmax,
                                                        const unsigned long long *__restrict__ incl, int64_t m, int32_t n_contigs,
                                                        int own_end, int32_t *__restrict__ out_contig, int32_t *__restrict__ out_start,
                                                        int32_t *__restrict__ out_end, int64_t *__restrict__ out_count) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const unsigned long long r = incl[i];
  if (i > 0 && incl[i - 1] == r) return;  // not the first row of its run
  const int64_t last = first_reaching(incl, i, m, r + 1ull) - 1;
  const int64_t k = (int64_t)r - 1;
  if (out_contig) out_contig[k] = contig_of_pos(seg, n_contigs, (int32_t)i);
  out_start[k] = st[i];
  out_end[k] = own_end ? en[i] : pmax[last];
  if (out_count) out_count[k] = last - i + 1;
}
// with inverted rows (start > end) the running max of the contig may exceed the max of the run: take the max of the
// run's own ends
__global__ void __launch_bounds__(256) run_end_max_kernel(const int32_t *__restrict__ en, const unsigned long long *__restrict__ incl,
                                                          int64_t m, int32_t *__restrict__ out_end) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < m) atomicMax(out_end + ((int64_t)incl[i] - 1), en[i]);
}

// cluster: per input row the id / start / end of its run; null-keyed rows keep the defaults written by cluster_init_kernel
__global__ void __launch_bounds__(256) cluster_init_kernel(int64_t n, int64_t *__restrict__ cid, int32_t *__restrict__ cs, int32_t *__restrict__ ce) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) { cid[i] = -1; cs[i] = 0; ce[i] = 0; }
}
__global__ void __launch_bounds__(256) cluster_scatter_kernel(const uint32_t *__restrict__ row, const unsigned long long *__restrict__ incl,
                                                              int64_t m, const int32_t *__restrict__ run_start, const int32_t *__restrict__ run_end,
                                                              int64_t *__restrict__ cid, int32_t *__restrict__ cs, int32_t *__restrict__ ce) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int64_t k = (int64_t)incl[i] - 1;
  const uint32_t r = row[i];
  cid[r] = k;
  cs[r] = __ldg(run_start + k);
  ce[r] = __ldg(run_end + k);
}

// subtract: right rows with start > end cover nothing -> null key
__global__ void __launch_bounds__(256) drop_inverted_kernel(const int32_t *__restrict__ c, const int32_t *__restrict__ s,
                                                            const int32_t *__restrict__ e, int64_t m, int32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < m) out[i] = s[i] <= e[i] ? c[i] : -1;
}

// Positions p, q of a left row among the merged runs of its contig (an index without nested or inverted rows, so
// the ends ascend with the starts): q = runs starting before the row ends, p = runs ending before the row starts;
// the runs p .. q-1 overlap the row.  Rank directory when the index has one, two bounded searches otherwise.
template <bool STRICT>
__device__ __forceinline__ void run_ranks(const IndexView &ix, int32_t c, int32_t s, int32_t e, int32_t &p, int32_t &q) {
  const int32_t seg_lo = ix.seg[c], seg_hi = ix.seg[c + 1];
  if (seg_lo >= seg_hi) { p = q = seg_lo; return; }
  if (ix.jdir && (STRICT ? (s < e) : (s <= e))) {
    const ContigMap32 cm = ld_cmap32(ix.cmap32 + c);
    uint32_t hi, re;
    jdir_ranks<STRICT>(ix, global_of(cm, s), global_of(cm, e), hi, re);
    q = (int32_t)hi;
    p = (int32_t)re;
    return;
  }
  q = STRICT ? lower_bound_i32(ix.st, seg_lo, seg_hi, e) : upper_bound_i32(ix.st, seg_lo, seg_hi, e);
  p = STRICT ? upper_bound_i32(ix.en, seg_lo, seg_hi, s) : lower_bound_i32(ix.en, seg_lo, seg_hi, s);
}

// pass 1: pieces left of every left row.  No overlapping run (or an inverted row): the row passes through as one
// piece; else (k-1) gaps between the k runs + the piece in front of the first run + the piece behind the last one.
template <bool STRICT>
__global__ void __launch_bounds__(256) subtract_count_kernel(IndexView ix, const int32_t *__restrict__ lc, const int32_t *__restrict__ ls,
                                                             const int32_t *__restrict__ le, int64_t n,
                                                             unsigned long long *__restrict__ pieces, int2 *__restrict__ pk) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int32_t c = lc[i];
  unsigned long long f = 0;
  int32_t p = 0, k = 0;
  if (c >= 0 && c < ix.n_contigs) {
    const int32_t s = ls[i], e = le[i];
    if (s <= e) {
      int32_t q;
      run_ranks<STRICT>(ix, c, s, e, p, q);
      k = q > p ? q - p : 0;
    }
    if (k == 0) f = 1;
    else f = (unsigned long long)(k - 1) + (s < __ldg(ix.st + p) ? 1u : 0u) + (__ldg(ix.en + p + k - 1) < e ? 1u : 0u);
  }
  pieces[i] = f;
  pk[i] = make_int2(p, k);
}
// pass 2: write the pieces of every left row at its exclusive offset.  ADJ = 0 (half-open) / 1 (closed coordinates:
// the piece in front of a run ends at start-1, the piece behind it begins at end+1).
template <bool STRICT>
__global__ void __launch_bounds__(256) subtract_emit_kernel(IndexView ix, const int32_t *__restrict__ lc, const int32_t *__restrict__ ls,
                                                            const int32_t *__restrict__ le, int64_t n,
                                                            const unsigned long long *__restrict__ offs, const int2 *__restrict__ pk,
                                                            uint32_t *__restrict__ out_row, int32_t *__restrict__ out_start,
                                                            int32_t *__restrict__ out_end) {
  constexpr int32_t ADJ = STRICT ? 0 : 1;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int32_t c = lc[i];
  if (c < 0 || c >= ix.n_contigs) return;
  unsigned long long o = offs[i];
  const int2 v = pk[i];
  const int32_t s = ls[i], e = le[i], p = v.x, k = v.y;
  if (k == 0) { out_row[o] = (uint32_t)i; out_start[o] = s; out_end[o] = e; return; }
  int32_t ms = __ldg(ix.st + p);
  if (s < ms) { out_row[o] = (uint32_t)i; out_start[o] = s; out_end[o] = ms - ADJ; ++o; }
  int32_t me = __ldg(ix.en + p);
  for (int32_t j = 1; j < k; ++j) {  // the gap between run p+j-1 and run p+j (never empty: merged runs do not touch)
    ms = __ldg(ix.st + p + j);
    out_row[o] = (uint32_t)i; out_start[o] = me + ADJ; out_end[o] = ms - ADJ; ++o;
    me = __ldg(ix.en + p + j);
  }
  if (me < e) { out_row[o] = (uint32_t)i; out_start[o] = me + ADJ; out_end[o] = e; }
}

// Sorted rows + run ids of one table.  On success *ix_out owns the sorted rows (release with index_free_on) and
// *incl_out (dev_free) holds the inclusive run count of every sorted position; *n_runs is on the host.
static int sweep_runs(const int32_t *d_c, const int32_t *d_s, const int32_t *d_e, int64_t m, int32_t n_contigs, bool strict_pred,
                      long long min_dist, cudaStream_t s, pbgpu_index **ix_out, unsigned long long **incl_out, int64_t *n_runs) {
  *ix_out = nullptr;
  *incl_out = nullptr;
  *n_runs = 0;
  pbgpu_index *ix = new (std::nothrow) pbgpu_index();
  if (!ix) return set_error(PBGPU_ENOMEM, "host allocation failed");
  int rc = index_build_impl(ix, d_c, d_s, d_e, m, n_contigs, s, /*sweep_only=*/true);
  if (rc != PBGPU_OK) { index_free_on(ix, s); return rc; }
  *ix_out = ix;
  const int64_t mv = ix->m;
  if (mv == 0) return PBGPU_OK;
  unsigned long long *incl = nullptr;
  rc = dev_alloc_t(&incl, (size_t)mv + 1, s);  // [mv] = the grand total
  if (rc != PBGPU_OK) return rc;
  *incl_out = incl;
  const unsigned grid = (unsigned)cdiv(mv, 256);
  if (strict_pred) PB_LAUNCH(sweep_fresh_kernel<true>, grid, 256, 0, s, ix->seg, ix->st, ix->pmax, mv, n_contigs, min_dist, incl);
  else PB_LAUNCH(sweep_fresh_kernel<false>, grid, 256, 0, s, ix->seg, ix->st, ix->pmax, mv, n_contigs, min_dist, incl);
  PB_CHECK_LAUNCH();
  PB_TRY((device_scan<SumU64, true>(incl, incl, mv, incl + mv, s)));
  unsigned long long total = 0;
  PB_TRY(fetch_words(incl + mv, 1, &total, s));
  *n_runs = (int64_t)total;
  return PBGPU_OK;
}

// the runs of `ix` as (contig?, start, end, count?) arrays of n_runs entries
static int emit_runs(const pbgpu_index *ix, const unsigned long long *incl, int64_t n_runs, int32_t *out_contig, int32_t *out_start,
                     int32_t *out_end, int64_t *out_count, cudaStream_t s) {
  if (n_runs == 0) return PBGPU_OK;
  const unsigned grid = (unsigned)cdiv(ix->m, 256);
  PB_LAUNCH(runs_emit_kernel, grid, 256, 0, s, ix->seg, ix->st, ix->en, ix->pmax, incl, ix->m, ix->n_contigs, ix->has_inverted, out_contig,
            out_start, out_end, out_count);
  if (ix->has_inverted) PB_LAUNCH(run_end_max_kernel, grid, 256, 0, s, ix->en, incl, ix->m, out_end);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

static int check_unary_args(const int32_t *c, const int32_t *s, const int32_t *e, int64_t m, int32_t n_contigs, int filter_op) {
  if (m < 0 || n_contigs < 0) return set_error(PBGPU_EINVAL, "negative size");
  if (m > 0 && (!c || !s || !e)) return set_error(PBGPU_EINVAL, "NULL column");
  if (m >= (int64_t)INT32_MAX) return set_error(PBGPU_ERANGE, "table has %lld rows; limit is 2^31-2", (long long)m);
  if (filter_op != PBGPU_FILTER_WEAK && filter_op != PBGPU_FILTER_STRICT) return set_error(PBGPU_EINVAL, "bad filter_op %d", filter_op);
  return PBGPU_OK;
}

}  // namespace pbgpu

extern "C" {

int64_t pbgpu_intervals_rows(const pbgpu_intervals *t) { return t ? t->rows : 0; }

int pbgpu_intervals_columns(const pbgpu_intervals *t, const int32_t **d_contig, const uint32_t **d_row, const int32_t **d_start,
                            const int32_t **d_end, const int64_t **d_count) {
  if (!t) return set_error(PBGPU_EINVAL, "table is NULL");
  if (d_contig) *d_contig = t->contig;
  if (d_row) *d_row = t->row;
  if (d_start) *d_start = t->start;
  if (d_end) *d_end = t->end;
  if (d_count) *d_count = t->count;
  return PBGPU_OK;
}

// copies the result's columns into caller-owned device buffers of pbgpu_intervals_rows() entries (NULL = skip)
int pbgpu_intervals_copy(const pbgpu_intervals *t, int32_t *d_contig, uint32_t *d_row, int32_t *d_start, int32_t *d_end,
                         int64_t *d_count, void *stream) {
  if (!t) return set_error(PBGPU_EINVAL, "table is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)t->rows;
  if (n == 0) return PBGPU_OK;
  if ((d_contig && !t->contig) || (d_row && !t->row) || (d_count && !t->count)) return set_error(PBGPU_EINVAL, "the result has no such column");
  if (d_contig) PB_CUDA(cudaMemcpyAsync(d_contig, t->contig, 4 * n, cudaMemcpyDeviceToDevice, s));
  if (d_row) PB_CUDA(cudaMemcpyAsync(d_row, t->row, 4 * n, cudaMemcpyDeviceToDevice, s));
  if (d_start) PB_CUDA(cudaMemcpyAsync(d_start, t->start, 4 * n, cudaMemcpyDeviceToDevice, s));
  if (d_end) PB_CUDA(cudaMemcpyAsync(d_end, t->end, 4 * n, cudaMemcpyDeviceToDevice, s));
  if (d_count) PB_CUDA(cudaMemcpyAsync(d_count, t->count, 8 * n, cudaMemcpyDeviceToDevice, s));
  return PBGPU_OK;
}

void pbgpu_intervals_free(pbgpu_intervals *t, void *stream) {
  if (!t) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != t->device) cudaSetDevice(t->device);
  dev_free(t->slab, (cudaStream_t)stream);
  if (cur != t->device) cudaSetDevice(cur);
  delete t;
}

// MergeProvider (operation.rs:352-380)
int pbgpu_merge(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t m, int32_t n_contigs, int filter_op,
                int64_t min_dist, void *stream, pbgpu_intervals **out) {
  if (!out) return set_error(PBGPU_EINVAL, "out is NULL");
  *out = nullptr;
  PB_TRY(check_unary_args(d_contig, d_start, d_end, m, n_contigs, filter_op));
  if (min_dist < 0) return set_error(PBGPU_EINVAL, "min_dist must be >= 0");
  cudaStream_t s = (cudaStream_t)stream;
  pbgpu_intervals *t = new (std::nothrow) pbgpu_intervals();
  if (!t) return set_error(PBGPU_ENOMEM, "host allocation failed");
  cudaGetDevice(&t->device);
  pbgpu_index *ix = nullptr;
  unsigned long long *incl = nullptr;
  int64_t k = 0;
  int rc = sweep_runs(d_contig, d_start, d_end, m, n_contigs, filter_op == PBGPU_FILTER_STRICT, (long long)min_dist, s, &ix, &incl, &k);
  if (rc == PBGPU_OK && k > 0) {
    const size_t a4 = (4 * (size_t)k + 255) & ~(size_t)255, a8 = (8 * (size_t)k + 255) & ~(size_t)255;
    rc = dev_alloc(&t->slab, 3 * a4 + a8, s);
    if (rc == PBGPU_OK) {
      char *b = (char *)t->slab;
      t->count = (int64_t *)b;
      t->contig = (int32_t *)(b + a8);
      t->start = (int32_t *)(b + a8 + a4);
      t->end = (int32_t *)(b + a8 + 2 * a4);
      t->rows = k;
      rc = emit_runs(ix, incl, k, t->contig, t->start, t->end, t->count, s);
    }
  }
  dev_free(incl, s);
  index_free_on(ix, s);
  if (rc != PBGPU_OK) { pbgpu_intervals_free(t, stream); return rc; }
  *out = t;
  return PBGPU_OK;
}

// ClusterProvider (operation.rs:382-430): d_cluster int64[m] (-1 for null-keyed rows), d_cluster_start / _end int32[m]
int pbgpu_cluster(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t m, int32_t n_contigs, int filter_op,
                  int64_t min_dist, int64_t *d_cluster, int32_t *d_cluster_start, int32_t *d_cluster_end, int64_t *n_clusters,
                  void *stream) {
  PB_TRY(check_unary_args(d_contig, d_start, d_end, m, n_contigs, filter_op));
  if (min_dist < 0) return set_error(PBGPU_EINVAL, "min_dist must be >= 0");
  if (m > 0 && (!d_cluster || !d_cluster_start || !d_cluster_end)) return set_error(PBGPU_EINVAL, "NULL output column");
  cudaStream_t s = (cudaStream_t)stream;
  if (n_clusters) *n_clusters = 0;
  if (m == 0) return PBGPU_OK;
  pbgpu_index *ix = nullptr;
  unsigned long long *incl = nullptr;
  int32_t *rs = nullptr;  // run start | run end
  int64_t k = 0;
  int rc = sweep_runs(d_contig, d_start, d_end, m, n_contigs, filter_op == PBGPU_FILTER_STRICT, (long long)min_dist, s, &ix, &incl, &k);
  if (rc == PBGPU_OK) {
    PB_LAUNCH(cluster_init_kernel, (unsigned)cdiv(m, 256), 256, 0, s, m, d_cluster, d_cluster_start, d_cluster_end);
    if (k > 0) {
      rc = dev_alloc_t(&rs, 2 * (size_t)k, s);
      if (rc == PBGPU_OK) rc = emit_runs(ix, incl, k, nullptr, rs, rs + k, nullptr, s);
      if (rc == PBGPU_OK) {
        PB_LAUNCH(cluster_scatter_kernel, (unsigned)cdiv(ix->m, 256), 256, 0, s, ix->row, incl, ix->m, rs, rs + k, d_cluster, d_cluster_start,
                  d_cluster_end);
        if (cudaGetLastError() != cudaSuccess) rc = set_error(PBGPU_ECUDA, "cluster kernels failed to launch");
      }
    }
  }
  dev_free(rs, s);
  dev_free(incl, s);
  index_free_on(ix, s);
  if (rc == PBGPU_OK && n_clusters) *n_clusters = k;
  return rc;
}

// SubtractProvider (operation.rs:463-510); ComplementProvider (operation.rs:432-461) = the view table on the left
int pbgpu_subtract(const int32_t *l_contig, const int32_t *l_start, const int32_t *l_end, int64_t n, const int32_t *r_contig,
                   const int32_t *r_start, const int32_t *r_end, int64_t m, int32_t n_contigs, int filter_op, void *stream,
                   pbgpu_intervals **out) {
  if (!out) return set_error(PBGPU_EINVAL, "out is NULL");
  *out = nullptr;
  PB_TRY(check_unary_args(l_contig, l_start, l_end, n, n_contigs, filter_op));
  PB_TRY(check_unary_args(r_contig, r_start, r_end, m, n_contigs, filter_op));
  cudaStream_t s = (cudaStream_t)stream;
  const bool strict = filter_op == PBGPU_FILTER_STRICT;
  pbgpu_intervals *t = new (std::nothrow) pbgpu_intervals();
  if (!t) return set_error(PBGPU_ENOMEM, "host allocation failed");
  cudaGetDevice(&t->device);
  pbgpu_index *rx = nullptr, *mx = nullptr;  // sorted right rows; index over their merged runs
  unsigned long long *incl = nullptr, *pieces = nullptr;
  int32_t *rc2 = nullptr, *runs = nullptr;  // right contigs with inverted rows dropped; run contig | start | end
  int2 *pk = nullptr;
  int64_t k = 0;
  int rc = PBGPU_OK;
  do {
    // 1. right side -> disjoint runs that do not even touch (Strict: start <= reach; Weak: start <= reach + 1)
    if (m > 0) {
      if ((rc = dev_alloc_t(&rc2, (size_t)m, s)) != PBGPU_OK) break;
      PB_LAUNCH(drop_inverted_kernel, (unsigned)cdiv(m, 256), 256, 0, s, r_contig, r_start, r_end, m, rc2);
    }
    if ((rc = sweep_runs(rc2, r_start, r_end, m, n_contigs, /*strict_pred=*/false, strict ? 0 : 1, s, &rx, &incl, &k)) != PBGPU_OK) break;
    if ((rc = dev_alloc_t(&runs, 3 * (size_t)(k ? k : 1), s)) != PBGPU_OK) break;
    if ((rc = emit_runs(rx, incl, k, runs, runs + k, runs + 2 * k, nullptr, s)) != PBGPU_OK) break;
    // 2. index over the runs (no nesting, no inverted rows: ends ascend with starts; rank directory when the span allows)
    mx = new (std::nothrow) pbgpu_index();
    if (!mx) { rc = set_error(PBGPU_ENOMEM, "host allocation failed"); break; }
    if ((rc = index_build_impl(mx, runs, runs + k, runs + 2 * k, k, n_contigs, s)) != PBGPU_OK) break;
    if (n == 0) break;
    // 3. pieces per left row, offsets, total
    if ((rc = dev_alloc_t(&pieces, (size_t)n + 1, s)) != PBGPU_OK) break;
    if ((rc = dev_alloc_t(&pk, (size_t)n, s)) != PBGPU_OK) break;
    const unsigned grid = (unsigned)cdiv(n, 256);
    if (strict) PB_LAUNCH(subtract_count_kernel<true>, grid, 256, 0, s, view_of(mx), l_contig, l_start, l_end, n, pieces, pk);
    else PB_LAUNCH(subtract_count_kernel<false>, grid, 256, 0, s, view_of(mx), l_contig, l_start, l_end, n, pieces, pk);
    if ((rc = device_scan<SumU64, false>(pieces, pieces, n, pieces + n, s)) != PBGPU_OK) break;
    unsigned long long total = 0;
    if ((rc = fetch_words(pieces + n, 1, &total, s)) != PBGPU_OK) break;
    if (total >= 0xFFFFFFFFull) { rc = set_error(PBGPU_ERANGE, "subtract leaves %llu pieces; limit is 2^32-2", total); break; }
    if (total == 0) break;
    // 4. exact-sized result
    const size_t a4 = (4 * (size_t)total + 255) & ~(size_t)255;
    if ((rc = dev_alloc(&t->slab, 3 * a4, s)) != PBGPU_OK) break;
    t->row = (uint32_t *)t->slab;
    t->start = (int32_t *)((char *)t->slab + a4);
    t->end = (int32_t *)((char *)t->slab + 2 * a4);
    t->rows = (int64_t)total;
    if (strict) PB_LAUNCH(subtract_emit_kernel<true>, grid, 256, 0, s, view_of(mx), l_contig, l_start, l_end, n, pieces, pk, t->row, t->start, t->end);
    else PB_LAUNCH(subtract_emit_kernel<false>, grid, 256, 0, s, view_of(mx), l_contig, l_start, l_end, n, pieces, pk, t->row, t->start, t->end);
    if (cudaGetLastError() != cudaSuccess) rc = set_error(PBGPU_ECUDA, "subtract kernels failed to launch");
  } while (0);
  dev_free(pk, s);
  dev_free(pieces, s);
  index_free_on(mx, s);
  dev_free(runs, s);
  dev_free(incl, s);
  index_free_on(rx, s);
  dev_free(rc2, s);
  if (rc != PBGPU_OK) { pbgpu_intervals_free(t, stream); return rc; }
  *out = t;
  return PBGPU_OK;
}

}  // extern "C"
