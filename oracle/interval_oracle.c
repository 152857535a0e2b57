/*
 * oracle/interval_oracle.c -- CPU ORACLE for the interval-join hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (polars_bio_b200 + libpbgpu.so) never links or calls anything here.
 *
 * What it restates.  The reference (biodatageeks/polars-bio @ f32af94) delegates the
 * arithmetic of pb.overlap / pb.count_overlaps / pb.nearest / pb.coverage to the
 * un-vendored crate datafusion-bio-function-ranges v0.11.0 (Cargo.toml:65, Cargo.lock:1830-1842)
 * which indexes the build side in coitrees 0.4.0 (Cargo.lock:1091-1094): one augmented
 * interval tree per contig (nodes ordered by start, every node carrying the maximum end of
 * its subtree), queried once per probe row.  That source is not under /root/reference, so
 * the published algorithm is restated here from its call sites and in-repo specification:
 *   - predicate               docs/developers.md:549-552 (Strict: a.start <  b.end && a.end >  b.start,
 *                                                        Weak:   a.start <= b.end && a.end >= b.start)
 *   - FilterOp numbering      src/option.rs:96-99  (Weak = 0, Strict = 1)
 *   - build / probe roles     docs/developers.md:629-649, src/operation.rs:143-158, 253-263, 316-340
 *   - count identity          polars_bio/range_op.py:548-594 (starts-rank minus ends-rank)
 *   - nearest distance        tests/_expected.py:162 ([100,200] vs [234,300] -> 34), bioframe parity
 *                             tests/test_bioframe.py:171-186 (distance only, partner ties arbitrary)
 *   - coverage adjacency      tests/test_coordinate_system_metadata.py:1577-1623
 * Pinned by tests/test_oracle_golden.py against every fixture listed in SURVEY.md 8(c).
 *
 * Parity unpinned (no reference test constrains it; choices documented in DESIGN.md):
 *   nearest tie-break (here: overlapping partners first, then (distance, start, row));
 *   rows with a negative contig code (null key) never match;  start > end rows are
 *   evaluated by the bare predicate.
 *
 * Layout: every table is three int32 columns (contig code, start, end); row ids are
 * positions in those columns.  Contig codes come from a dictionary shared by both sides.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PBO_API __attribute__((visibility("default")))

struct pbo_node { int32_t st, en, left_max, right_max; };  /* max end over the left / right subtree of this slot */
typedef struct {
  int64_t m;          /* indexed (build) rows kept: contig code >= 0                          */
  int32_t n_contigs;  /* contig codes are 0 .. n_contigs-1                                   */
  int64_t *seg;       /* [n_contigs+1] segment offsets into the sorted arrays                 */
  int32_t *st;        /* starts, sorted by (contig, start, row)                               */
  int32_t *en;        /* ends in the same order                                               */
  uint32_t *row;      /* original row ids in the same order                                   */
  int32_t *sub_max;   /* augmented key: max end over the implicit subtree rooted at this slot */
  struct pbo_node *node; /* (start, end, max end of the left subtree, of the right subtree) of every slot side by side:
                          * one cache line per visited node, and a child that cannot hit is never loaded */
  int32_t *en_sorted; /* ends sorted by (contig, end, start, row)                             */
  uint32_t *en_pos;   /* position in st/en/row order of each en_sorted entry                  */
} pbo_index;

/* Stable LSD radix sort of records by a 64-bit key (8 passes of 8 bits, skipping constant bytes).  The
 * reference sorts its nodes once per contig when it builds a COITree; a comparison qsort here would make
 * the CPU baseline needlessly slow. */
typedef struct { uint64_t key; uint32_t a, b, c; } pbo_srec; /* a,b,c: payload */
static void radix_sort_srec(pbo_srec *v, int64_t n) {
  if (n < 2) return;
  pbo_srec *tmp = (pbo_srec *)malloc(sizeof(pbo_srec) * (size_t)n), *src = v, *dst = tmp;
  for (int pass = 0; pass < 8; ++pass) {
    int64_t cnt[256] = {0};
    const int sh = pass * 8;
    for (int64_t i = 0; i < n; ++i) cnt[(src[i].key >> sh) & 0xff]++;
    int skip = 0;
    for (int d = 0; d < 256; ++d) if (cnt[d] == n) { skip = 1; break; }
    if (skip) continue;
    int64_t run = 0;
    for (int d = 0; d < 256; ++d) { int64_t t = cnt[d]; cnt[d] = run; run += t; }
    for (int64_t i = 0; i < n; ++i) dst[cnt[(src[i].key >> sh) & 0xff]++] = src[i];
    pbo_srec *t = src; src = dst; dst = t;
  }
  if (src != v) memcpy(v, src, sizeof(pbo_srec) * (size_t)n);
  free(tmp);
}

/* Implicit balanced tree over a sorted slice [lo,hi): root = midpoint, children = the two
 * halves.  sub_max[mid] = max end over [lo,hi)  (the COITrees/IITree augmentation).        */
static int32_t build_aug(const int32_t *en, int32_t *sub_max, int64_t lo, int64_t hi) {
  if (lo >= hi) return INT32_MIN;
  int64_t mid = lo + ((hi - lo) >> 1);
  int32_t m = en[mid];
  int32_t l = build_aug(en, sub_max, lo, mid);
  int32_t r = build_aug(en, sub_max, mid + 1, hi);
  if (l > m) m = l;
  if (r > m) m = r;
  sub_max[mid] = m;
  return m;
}

PBO_API void pbo_index_free(pbo_index *ix) {
  if (!ix) return;
  free(ix->seg); free(ix->st); free(ix->en); free(ix->row);
  free(ix->sub_max); free(ix->node); free(ix->en_sorted); free(ix->en_pos); free(ix);
}

/* node[mid] of the implicit tree over [lo,hi): the children are the midpoints of [lo,mid) and [mid+1,hi) */
static void fill_nodes(pbo_index *ix, int64_t lo, int64_t hi) {
  if (lo >= hi) return;
  const int64_t mid = lo + ((hi - lo) >> 1);
  struct pbo_node *nd = &ix->node[mid];
  nd->st = ix->st[mid]; nd->en = ix->en[mid];
  nd->left_max = lo < mid ? ix->sub_max[lo + ((mid - lo) >> 1)] : INT32_MIN;
  nd->right_max = mid + 1 < hi ? ix->sub_max[mid + 1 + ((hi - mid - 1) >> 1)] : INT32_MIN;
  fill_nodes(ix, lo, mid);
  fill_nodes(ix, mid + 1, hi);
}

PBO_API pbo_index *pbo_index_build(const int32_t *c, const int32_t *s, const int32_t *e,
                                   int64_t m_in, int32_t n_contigs) {
  pbo_index *ix = (pbo_index *)calloc(1, sizeof(pbo_index));
  pbo_srec *rec = (pbo_srec *)malloc(sizeof(pbo_srec) * (size_t)(m_in > 0 ? m_in : 1));
  int64_t m = 0;
  for (int64_t i = 0; i < m_in; ++i)
    if (c[i] >= 0 && c[i] < n_contigs) {   /* key = (contig, start biased to unsigned); stable => ties keep row order */
      rec[m].key = ((uint64_t)(uint32_t)c[i] << 32) | ((uint32_t)s[i] ^ 0x80000000u);
      rec[m].a = (uint32_t)s[i]; rec[m].b = (uint32_t)e[i]; rec[m].c = (uint32_t)i; ++m;
    }
  radix_sort_srec(rec, m);
  ix->m = m; ix->n_contigs = n_contigs;
  size_t mm = (size_t)(m > 0 ? m : 1);
  ix->seg = (int64_t *)calloc((size_t)n_contigs + 1, sizeof(int64_t));
  ix->st = (int32_t *)malloc(4 * mm); ix->en = (int32_t *)malloc(4 * mm);
  ix->row = (uint32_t *)malloc(4 * mm); ix->sub_max = (int32_t *)malloc(4 * mm);
  ix->en_sorted = (int32_t *)malloc(4 * mm); ix->en_pos = (uint32_t *)malloc(4 * mm);
  for (int64_t i = 0; i < m; ++i) {
    ix->st[i] = (int32_t)rec[i].a; ix->en[i] = (int32_t)rec[i].b; ix->row[i] = rec[i].c;
    ix->seg[(rec[i].key >> 32) + 1]++;
  }
  for (int32_t k = 0; k < n_contigs; ++k) ix->seg[k + 1] += ix->seg[k];
  for (int32_t k = 0; k < n_contigs; ++k) build_aug(ix->en, ix->sub_max, ix->seg[k], ix->seg[k + 1]);
  ix->node = (struct pbo_node *)malloc(sizeof(struct pbo_node) * mm);
  for (int32_t k = 0; k < n_contigs; ++k) fill_nodes(ix, ix->seg[k], ix->seg[k + 1]);
  /* end order: stable sort of the start-ordered records by (contig, end) => ties keep (start,row) order */
  for (int64_t i = 0; i < m; ++i) {
    rec[i].key = (rec[i].key & 0xffffffff00000000ull) | (rec[i].b ^ 0x80000000u);
    rec[i].c = (uint32_t)i; /* position in start order */
  }
  radix_sort_srec(rec, m);
  for (int64_t i = 0; i < m; ++i) { ix->en_sorted[i] = (int32_t)rec[i].b; ix->en_pos[i] = rec[i].c; }
  free(rec);
  return ix;
}

static inline int hit(int strict, int32_t as, int32_t ae, int32_t bs, int32_t be) {
  return strict ? (as < be && ae > bs) : (as <= be && ae >= bs);
}

/* Tree query: visit every indexed interval of the contig slice that satisfies the predicate,
 * in (start,row) order.  cb_pos receives positions in sorted order; returns count.          */
#define PBO_LEAF 16
typedef struct { int64_t lo, hi; } pbo_span;
static int64_t tree_query(const pbo_index *ix, int64_t lo0, int64_t hi0, int strict,
                          int32_t qs, int32_t qe, uint32_t *out_pos, int64_t cap) {
  /* explicit stack; depth <= 64.  In-order traversal so results come out start-sorted.  A node carries the max end of
   * each child's subtree, so a child that cannot hit is pruned without being loaded; and, like COITrees, the bottom of
   * the tree is not descended node by node: a subtree of at most PBO_LEAF slots is a contiguous, start-sorted run of
   * the node array and is scanned linearly. */
  const struct pbo_node *nd = ix->node;
  struct { int64_t lo, hi; int stage; } stk[70];
  int sp = 0; int64_t n = 0;
  if (lo0 >= hi0) return 0;
  {
    const int32_t mx = ix->sub_max[lo0 + ((hi0 - lo0) >> 1)];
    if (strict ? (mx <= qs) : (mx < qs)) return 0;
  }
  stk[sp].lo = lo0; stk[sp].hi = hi0; stk[sp].stage = 0; ++sp;
  while (sp > 0) {
    const int64_t lo = stk[sp - 1].lo, hi = stk[sp - 1].hi; const int stage = stk[sp - 1].stage;
    if (hi - lo <= PBO_LEAF) {   /* whole subtree (already known to reach past the query start): linear scan */
      --sp;
      for (int64_t j = lo; j < hi; ++j) {
        const int32_t bs = nd[j].st;
        if (strict ? (bs >= qe) : (bs > qe)) break;   /* sorted by start: nothing further can hit */
        if (hit(strict, qs, qe, bs, nd[j].en)) { if (out_pos && n < cap) out_pos[n] = (uint32_t)j; ++n; }
      }
      continue;
    }
    const int64_t mid = lo + ((hi - lo) >> 1);
    const struct pbo_node x = nd[mid];
    if (stage == 0) {
      stk[sp - 1].stage = 1;
      /* left subtree first, unless nothing in it ends after (at) the query start */
      if (strict ? (x.left_max > qs) : (x.left_max >= qs)) { stk[sp].lo = lo; stk[sp].hi = mid; stk[sp].stage = 0; ++sp; }
      continue;
    }
    --sp;
    /* node itself, then right subtree, only if node start is still before the query end */
    if (strict ? (x.st < qe) : (x.st <= qe)) {
      if (hit(strict, qs, qe, x.st, x.en)) { if (out_pos && n < cap) out_pos[n] = (uint32_t)mid; ++n; }
      if (strict ? (x.right_max > qs) : (x.right_max >= qs)) { stk[sp].lo = mid + 1; stk[sp].hi = hi; stk[sp].stage = 0; ++sp; }
    }
  }
  return n;
}

/* ---- count_overlaps: per iterated row, number of indexed rows that overlap it ------------ */
PBO_API void pbo_count_overlaps(const pbo_index *ix, const int32_t *c, const int32_t *s, const int32_t *e,
                                int64_t n, int strict, int64_t *counts, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 4096)
#endif
  for (int64_t i = 0; i < n; ++i) {
    int32_t cc = c[i];
    if (cc < 0 || cc >= ix->n_contigs) { counts[i] = 0; continue; }
    counts[i] = tree_query(ix, ix->seg[cc], ix->seg[cc + 1], strict, s[i], e[i], NULL, 0);
  }
}

/* ---- overlap: all (probe_row, build_row) pairs -------------------------------------------
 * One tree query per probe row, the way the reference's probe loop works (docs/developers.md:629-649):
 * every thread walks a contiguous slice of the probe rows and appends its hits to a growable buffer;
 * slices are concatenated in order, so pairs come out ordered by probe row, then by (start,row) of
 * the indexed partner.  Returns the total pair count; writes min(total, cap) pairs.  With NULL output
 * buffers it only counts.                                                                     */
typedef struct { uint32_t *p, *b; int64_t n, cap; } pbo_buf;
static void buf_reserve(pbo_buf *o, int64_t extra) {
  if (o->n + extra <= o->cap) return;
  int64_t nc = o->cap ? o->cap * 2 : 4096;
  while (nc < o->n + extra) nc *= 2;
  o->p = (uint32_t *)realloc(o->p, 4 * (size_t)nc);
  o->b = (uint32_t *)realloc(o->b, 4 * (size_t)nc);
  o->cap = nc;
}
PBO_API int64_t pbo_overlap_pairs(const pbo_index *ix, const int32_t *c, const int32_t *s, const int32_t *e,
                                  int64_t n, int strict, uint32_t *out_probe, uint32_t *out_build,
                                  int64_t cap, int threads) {
  int nt = 1;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
  nt = omp_get_max_threads();
#endif
  const int want = out_probe && out_build;
  pbo_buf *bufs = (pbo_buf *)calloc((size_t)nt, sizeof(pbo_buf));
  int64_t *tot = (int64_t *)calloc((size_t)nt + 1, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    pbo_buf *o = &bufs[t];
    int64_t count = 0;
    for (int64_t i = lo; i < hi; ++i) {
      const int32_t cc = c[i];
      if (cc < 0 || cc >= ix->n_contigs) continue;
      if (!want) { count += tree_query(ix, ix->seg[cc], ix->seg[cc + 1], strict, s[i], e[i], NULL, 0); continue; }
      buf_reserve(o, 64);
      int64_t k = tree_query(ix, ix->seg[cc], ix->seg[cc + 1], strict, s[i], e[i], o->b + o->n, o->cap - o->n);
      if (k > o->cap - o->n) {  /* more hits than room: grow and repeat this probe */
        buf_reserve(o, k);
        k = tree_query(ix, ix->seg[cc], ix->seg[cc + 1], strict, s[i], e[i], o->b + o->n, o->cap - o->n);
      }
      for (int64_t j = 0; j < k; ++j) { o->b[o->n + j] = ix->row[o->b[o->n + j]]; o->p[o->n + j] = (uint32_t)i; }
      o->n += k;
      count += k;
    }
    tot[t + 1] = count;
  }
  for (int t = 0; t < nt; ++t) tot[t + 1] += tot[t];
  const int64_t total = tot[nt];
  if (want) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static, 1)
#endif
    for (int t = 0; t < nt; ++t) {
      int64_t off = tot[t], k = bufs[t].n;
      if (off < cap) {
        if (off + k > cap) k = cap - off;
        memcpy(out_probe + off, bufs[t].p, 4 * (size_t)k);
        memcpy(out_build + off, bufs[t].b, 4 * (size_t)k);
      }
    }
  }
  for (int t = 0; t < nt; ++t) { free(bufs[t].p); free(bufs[t].b); }
  free(bufs); free(tot);
  return total;
}

/* ---- coverage: per iterated row, positions covered by the union of indexed rows ----------
 * Strict (half-open): sum of max(0, min(ae,be') - max(as,bs')) over merged runs.
 * Weak (closed):      sum of max(0, min(ae,be') - max(as,bs') + 1).
 * Pinned by tests/test_coordinate_system_metadata.py:1577-1623 (adjacency: 0 vs 1).         */
PBO_API void pbo_coverage(const pbo_index *ix, const int32_t *c, const int32_t *s, const int32_t *e,
                          int64_t n, int strict, int64_t *cov, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
  {
    int64_t cap = 1024; uint32_t *buf = (uint32_t *)malloc(4 * (size_t)cap);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1024)
#endif
    for (int64_t i = 0; i < n; ++i) {
      int32_t cc = c[i]; cov[i] = 0;
      if (cc < 0 || cc >= ix->n_contigs) continue;
      int64_t k = tree_query(ix, ix->seg[cc], ix->seg[cc + 1], strict, s[i], e[i], buf, cap);
      if (k > cap) { cap = k; buf = (uint32_t *)realloc(buf, 4 * (size_t)cap);
                     tree_query(ix, ix->seg[cc], ix->seg[cc + 1], strict, s[i], e[i], buf, cap); }
      /* hits arrive start-sorted: sweep and merge clipped pieces */
      int64_t total = 0, cur_s = 0, cur_e = 0; int open = 0;
      for (int64_t j = 0; j < k; ++j) {
        int64_t bs = ix->st[buf[j]], be = ix->en[buf[j]];
        if (bs < s[i]) bs = s[i];
        if (be > e[i]) be = e[i];
        if (!strict) be += 1;            /* closed -> half-open in 64-bit, no overflow */
        if (be <= bs) continue;
        if (!open) { cur_s = bs; cur_e = be; open = 1; }
        else if (bs <= cur_e) { if (be > cur_e) cur_e = be; }
        else { total += cur_e - cur_s; cur_s = bs; cur_e = be; }
      }
      if (open) total += cur_e - cur_s;
      cov[i] = total;
    }
    free(buf);
  }
}

/* ---- nearest ------------------------------------------------------------------------------
 * For every iterated row: up to k indexed rows of the same contig ordered by
 *   (overlapping first [if include_overlaps], then distance, then start, then row)
 * distance = 0 for overlapping partners, else max(b.start - a.end, a.start - b.end) (>= 0).
 * out_build[i*k + j] = indexed row id or UINT32_MAX (no partner); out_dist likewise / -1.    */
static inline int64_t gap(int32_t as, int32_t ae, int32_t bs, int32_t be) {
  int64_t d1 = (int64_t)bs - (int64_t)ae, d2 = (int64_t)as - (int64_t)be;
  int64_t d = d1 > d2 ? d1 : d2;
  return d > 0 ? d : 0;
}

PBO_API void pbo_nearest(const pbo_index *ix, const int32_t *c, const int32_t *s, const int32_t *e,
                         int64_t n, int strict, int64_t k, int include_overlaps,
                         uint32_t *out_build, int64_t *out_dist, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
  {
    uint32_t *buf = (uint32_t *)malloc(4 * (size_t)(k > 0 ? k : 1));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4096)
#endif
    for (int64_t i = 0; i < n; ++i) {
      uint32_t *ob = out_build + i * k; int64_t *od = out_dist + i * k;
      for (int64_t j = 0; j < k; ++j) { ob[j] = UINT32_MAX; od[j] = -1; }
      int32_t cc = c[i];
      if (cc < 0 || cc >= ix->n_contigs) continue;
      int64_t lo = ix->seg[cc], hi = ix->seg[cc + 1];
      if (lo >= hi) continue;
      int32_t qs = s[i], qe = e[i];
      int64_t got = 0;
      if (include_overlaps) {
        int64_t h = tree_query(ix, lo, hi, strict, qs, qe, buf, k);
        if (h > k) h = k;
        for (int64_t j = 0; j < h; ++j) { ob[j] = ix->row[buf[j]]; od[j] = 0; }
        got = h;
      }
      if (got >= k) continue;
      /* right stream: first position (start,row order) whose start is past the query end.
       * Entries past the end can still overlap only if inverted; skip overlapping ones.       */
      int64_t a = lo, b = hi;
      while (a < b) { int64_t mid = a + ((b - a) >> 1);
        if (strict ? (ix->st[mid] < qe) : (ix->st[mid] <= qe)) a = mid + 1; else b = mid; }
      int64_t rp = a;
      /* left stream: en_sorted positions with end before (at) the query start, walked by
       * descending end; inside one end-value group ascending (start,row).                    */
      a = lo; b = hi;
      while (a < b) { int64_t mid = a + ((b - a) >> 1);
        if (strict ? (ix->en_sorted[mid] <= qs) : (ix->en_sorted[mid] < qs)) a = mid + 1; else b = mid; }
      int64_t lend = a;           /* en_sorted[lo..lend) are left candidates */
      int64_t lg_hi = lend, lg_lo = lend, lcur = lend; /* current group [lg_lo,lg_hi), cursor lcur */
      /* advance helpers inline */
      for (;;) {
        /* settle the left cursor on a non-overlapping entry */
        int have_l = 0, have_r = 0; int64_t lpos = 0, ld = 0, rd = 0;
        for (;;) {
          if (lcur >= lg_hi) {              /* open next (smaller end) group */
            if (lg_lo <= lo) break;
            lg_hi = lg_lo; int32_t v = ix->en_sorted[lg_hi - 1];
            int64_t g = lg_hi - 1; while (g > lo && ix->en_sorted[g - 1] == v) --g;
            lg_lo = g; lcur = g;
          }
          lpos = ix->en_pos[lcur];
          if (hit(strict, qs, qe, ix->st[lpos], ix->en[lpos])) { ++lcur; continue; }
          have_l = 1; ld = gap(qs, qe, ix->st[lpos], ix->en[lpos]); break;
        }
        /* skip overlapping entries and degenerate ones already owned by the left stream */
        while (rp < hi && (hit(strict, qs, qe, ix->st[rp], ix->en[rp]) ||
                           (strict ? (ix->en[rp] <= qs) : (ix->en[rp] < qs)))) ++rp;
        if (rp < hi) { have_r = 1; rd = gap(qs, qe, ix->st[rp], ix->en[rp]); }
        if (!have_l && !have_r) break;
        int take_left;
        if (have_l && have_r) {
          if (ld != rd) take_left = ld < rd;
          else if (ix->st[lpos] != ix->st[rp]) take_left = ix->st[lpos] < ix->st[rp];
          else take_left = ix->row[lpos] < ix->row[rp];
        } else take_left = have_l;
        if (take_left) { ob[got] = ix->row[lpos]; od[got] = ld; ++lcur; }
        else { ob[got] = ix->row[rp]; od[got] = rd; ++rp; }
        if (++got >= k) break;
      }
    }
    free(buf);
  }
}

/* ---- brute force (ground truth for the ground truth; O(N*M); tiny inputs only) ---------- */
PBO_API int64_t pbo_brute_pairs(const int32_t *lc, const int32_t *ls, const int32_t *le, int64_t n,
                                const int32_t *rc, const int32_t *rs, const int32_t *re, int64_t m,
                                int strict, uint32_t *out_probe, uint32_t *out_build, int64_t cap) {
  int64_t p = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (lc[i] < 0) continue;
    for (int64_t j = 0; j < m; ++j) {
      if (rc[j] != lc[i]) continue;
      if (hit(strict, ls[i], le[i], rs[j], re[j])) {
        if (p < cap && out_probe) { out_probe[p] = (uint32_t)i; out_build[p] = (uint32_t)j; }
        ++p;
      }
    }
  }
  return p;
}

PBO_API int pbo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
