"""CPU oracle for the unary sweeps merge / cluster / complement / subtract -- TEST INFRASTRUCTURE ONLY.

The arithmetic lives in the reference's un-vendored dependency (datafusion-bio-function-ranges v0.11.0:
MergeProvider / ClusterProvider / ComplementProvider / SubtractProvider, constructed at
/root/reference/src/operation.rs:352-510); what is restated here is the behaviour the reference's own tests pin:

* merge: rows sorted by (contig, start); a row joins the running cluster iff  start <  cluster_end + min_dist  (Strict,
  0-based half-open) or  start <= cluster_end + min_dist  (Weak, 1-based closed) -- adjacent intervals stay apart
  0-based and merge 1-based (tests/test_coordinate_system_metadata.py:1032-1054; the sweep-line formulation with
  end + min_dist events, polars_bio/range_op.py:600-655 docs); output (contig, start, end, n_intervals), Int64
  (tests/_expected.py:174-181, tests/test_bioframe.py:122-126 = bioframe.merge(min_dist=None)).
* cluster: the same clusters, numbered from 0 over contigs in lexicographic name order then by start
  (bioframe.cluster numbering, compared column for column at tests/test_bioframe.py:392-411), reported per input row
  with the cluster's start / end (tests/test_partitioned_range_operation_regressions.py:49-59).
* subtract: for every left row the parts not covered by any right row of the same contig, left to right; a fully
  covered row yields nothing (…regressions.py:41-47 == bioframe.subtract, tests/test_bioframe.py:517-529).
* complement: view regions minus the intervals (…regressions.py:33-39 == bioframe.complement with a view,
  tests/test_bioframe.py:455-480); without a view every contig present spans [0, INT64_MAX)
  (polars_bio/range_op.py:726-729).

Parity unpinned (no reference test constrains it; deterministic choices shared with the CUDA path): min_dist > 0;
closed-coordinate (Weak) fragment arithmetic of subtract / complement ([s, e] minus [s2, e2] leaves [s, s2-1] and
[e2+1, e]); rows with a null key are dropped; in subtract / complement a right row with start > end covers nothing
and a left row with start > end passes through unchanged; merge / cluster sweep such rows by the bare comparisons.

Plain loops on purpose: this is the checker, written to be read, and it shares no code with the CUDA path.
"""
from __future__ import annotations

import numpy as np

I64_MAX = np.iinfo(np.int64).max


def _valid_sorted(c, s, e, n_contigs):
    c = np.asarray(c, np.int64); s = np.asarray(s, np.int64); e = np.asarray(e, np.int64)
    rows = np.nonzero((c >= 0) & (c < n_contigs))[0]
    order = rows[np.lexsort((rows, s[rows], c[rows]))]  # (contig, start, row)
    return c, s, e, order


def _clusters(c, s, e, order, strict: bool, min_dist: int):
    """Yield (contig, start, end, [rows]) of every cluster in (contig, start) order."""
    cur = None
    for r in order:
        cc, ss, ee = int(c[r]), int(s[r]), int(e[r])
        if cur is not None and cur[0] == cc and ((ss < cur[2] + min_dist) if strict else (ss <= cur[2] + min_dist)):
            cur[2] = max(cur[2], ee)
            cur[3].append(int(r))
        else:
            if cur is not None:
                yield tuple(cur)
            cur = [cc, ss, ee, [int(r)]]
    if cur is not None:
        yield tuple(cur)


def merge(c, s, e, n_contigs: int, strict: bool, min_dist: int = 0):
    """-> (contig, start, end, n_intervals) int64 arrays ordered by (contig code, start)."""
    c, s, e, order = _valid_sorted(c, s, e, n_contigs)
    out = [(cc, ss, ee, len(rows)) for cc, ss, ee, rows in _clusters(c, s, e, order, strict, min_dist)]
    a = np.array(out, dtype=np.int64).reshape(-1, 4)
    return a[:, 0], a[:, 1], a[:, 2], a[:, 3]


def cluster(c, s, e, n_contigs: int, strict: bool, min_dist: int = 0, contig_rank=None):
    """-> per input row (cluster id, cluster_start, cluster_end), int64; -1 / 0 / 0 for null-key rows.
    Ids count clusters over contigs in `contig_rank` order (rank of each code; default: code order), then start."""
    c, s, e, order = _valid_sorted(c, s, e, n_contigs)
    cl = list(_clusters(c, s, e, order, strict, min_dist))
    rank = np.arange(n_contigs) if contig_rank is None else np.asarray(contig_rank)
    cl.sort(key=lambda t: (int(rank[t[0]]), t[1]))  # stable: equal (contig, start) cannot happen for two clusters
    cid = np.full(len(c), -1, np.int64); cs = np.zeros(len(c), np.int64); ce = np.zeros(len(c), np.int64)
    for k, (_, ss, ee, rows) in enumerate(cl):
        cid[rows] = k; cs[rows] = ss; ce[rows] = ee
    return cid, cs, ce


def subtract(lc, ls, le, rc, rs, re, n_contigs: int, strict: bool):
    """-> (left_row, start, end) int64 arrays ordered by (left row, start)."""
    lc = np.asarray(lc, np.int64); ls = np.asarray(ls, np.int64); le = np.asarray(le, np.int64)
    rc_, rs_, re_, order = _valid_sorted(rc, rs, re, n_contigs)
    by_contig = {}
    for r in order:
        if rs_[r] <= re_[r]:  # an inverted right row covers nothing
            by_contig.setdefault(int(rc_[r]), []).append((int(rs_[r]), int(re_[r])))
    out = []
    for i in range(len(lc)):
        cc = int(lc[i])
        if cc < 0 or cc >= n_contigs:
            continue
        a_s, a_e = int(ls[i]), int(le[i])
        hits = [(b_s, b_e) for b_s, b_e in by_contig.get(cc, ())  # start order
                if a_s <= a_e and ((a_s < b_e and a_e > b_s) if strict else (a_s <= b_e and a_e >= b_s))]
        if not hits:  # untouched rows pass through as they are (zero-length / inverted ones included)
            out.append((i, a_s, a_e))
            continue
        cur = a_s  # first position of the row not yet accounted for
        for b_s, b_e in hits:
            if strict:
                if cur < b_s:
                    out.append((i, cur, b_s))
                cur = max(cur, b_e)
            else:
                if cur <= b_s - 1:
                    out.append((i, cur, b_s - 1))
                cur = max(cur, b_e + 1)
        if (cur < a_e) if strict else (cur <= a_e):
            out.append((i, cur, a_e))
    a = np.array(out, dtype=np.int64).reshape(-1, 3)
    return a[:, 0], a[:, 1], a[:, 2]


def subtract_ranks(lc, ls, le, rc, rs, re, n_contigs: int, strict: bool):
    """Second restatement, vectorised (the shape of the device algorithm, in numpy): merge the right rows that touch
    (Strict: start <= running end; Weak: start <= running end + 1) into disjoint, non-adjacent runs M; a left row
    overlapping the runs p..q-1 (two rank searches) leaves  (q-p-1) + [a.start before M_p] + [M_{q-1} ends before
    a.end]  pieces, each computable on its own from the row and the two runs around it."""
    lc = np.asarray(lc, np.int64); ls = np.asarray(ls, np.int64); le = np.asarray(le, np.int64)
    adj = 0 if strict else 1
    rc = np.where(np.asarray(rs, np.int64) <= np.asarray(re, np.int64), np.asarray(rc, np.int64), -1)  # inverted right rows: dropped
    mc, ms, me, _ = merge(rc, rs, re, n_contigs, strict=False, min_dist=adj)  # start <= end + adj
    BIG = np.int64(1) << 40  # contig-major composite keys: both arrays ascend (runs are disjoint inside a contig)
    off = np.int64(1) << 33  # shifts int32 coordinates to non-negative
    ks, ke = mc * BIG + (ms + off), mc * BIG + (me + off)
    ok = (lc >= 0) & (lc < n_contigs)
    qa_s, qa_e = lc * BIG + (ls + off), lc * BIG + (le + off)
    if strict:
        q = np.searchsorted(ks, qa_e, side="left")    # runs starting before the row ends
        p = np.searchsorted(ke, qa_s, side="right")   # runs ending at or before the row starts
    else:
        q = np.searchsorted(ks, qa_e, side="right")
        p = np.searchsorted(ke, qa_s, side="left")
    k = np.maximum(q - p, 0)
    k[~ok | (ls > le)] = 0  # inverted left rows pass through untouched
    out = []
    for i in np.nonzero(ok)[0]:
        a_s, a_e = int(ls[i]), int(le[i])
        if k[i] == 0:
            out.append((i, a_s, a_e))
            continue
        first, last = int(p[i]), int(q[i]) - 1
        if a_s < ms[first]:
            out.append((i, a_s, int(ms[first]) - adj))
        for j in range(first, last):
            out.append((i, int(me[j]) + adj, int(ms[j + 1]) - adj))
        if me[last] < a_e:
            out.append((i, int(me[last]) + adj, a_e))
    a = np.array(out, dtype=np.int64).reshape(-1, 3)
    return a[:, 0], a[:, 1], a[:, 2]


def complement(c, s, e, n_contigs: int, strict: bool, view=None):
    """-> (contig, start, end) int64 ordered by (view row, start).  `view` = (contig, start, end) arrays of the view
    regions; None = one region [0, INT64_MAX) per contig present among the valid rows (reported with end INT64_MAX)."""
    cc = np.asarray(c, np.int64)
    if view is None:
        present = np.unique(cc[(cc >= 0) & (cc < n_contigs)])
        vc, vs, ve = present, np.zeros(len(present), np.int64), np.full(len(present), I64_MAX, np.int64)
    else:
        vc, vs, ve = (np.asarray(x, np.int64) for x in view)
    row, fs, fe = subtract(vc, vs, ve, c, s, e, n_contigs, strict)
    return vc[row], fs, fe


# ---- brute-force twins for tiny inputs (independent of the sweeps above) ---------------------------------------
def merge_bruteforce(c, s, e, n_contigs: int, strict: bool, min_dist: int = 0):
    """Connected components of the 'touch' graph by union-find over all pairs, O(n^2)."""
    c = np.asarray(c, np.int64); s = np.asarray(s, np.int64); e = np.asarray(e, np.int64)
    rows = [i for i in range(len(c)) if 0 <= c[i] < n_contigs]
    parent = {i: i for i in rows}

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for i in rows:
        for j in rows:
            if i < j and c[i] == c[j]:
                lo, hi = (i, j) if (s[i], i) <= (s[j], j) else (j, i)  # lo starts first
                touch = (s[hi] < e[lo] + min_dist) if strict else (s[hi] <= e[lo] + min_dist)
                if touch:
                    parent[find(i)] = find(j)
    comp = {}
    for i in rows:
        comp.setdefault(find(i), []).append(i)
    out = sorted((int(c[m[0]]), int(min(s[m])), int(max(e[m])), len(m)) for m in comp.values())
    a = np.array(out, dtype=np.int64).reshape(-1, 4)
    return a[:, 0], a[:, 1], a[:, 2], a[:, 3]


def subtract_bruteforce(lc, ls, le, rc, rs, re, n_contigs: int, strict: bool):
    """Position bitmaps (small non-negative coordinates only): the runs of a left row's positions no right row covers."""
    out = []
    for i in range(len(lc)):
        if not (0 <= lc[i] < n_contigs):
            continue
        a_s, a_e = int(ls[i]), int(le[i]) if strict else int(le[i]) + 1  # half-open positions
        if a_e <= a_s:
            continue
        free = np.ones(a_e - a_s, bool)
        for j in range(len(rc)):
            if rc[j] != lc[i]:
                continue
            b_s, b_e = int(rs[j]), int(re[j]) if strict else int(re[j]) + 1
            lo, hi = max(b_s, a_s), min(b_e, a_e)
            if hi > lo:
                free[lo - a_s:hi - a_s] = False
        d = np.diff(np.concatenate(([0], free.astype(np.int8), [0])))
        for st, en in zip(np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]):
            out.append((i, a_s + int(st), a_s + int(en) if strict else a_s + int(en) - 1))
    a = np.array(out, dtype=np.int64).reshape(-1, 3)
    return a[:, 0], a[:, 1], a[:, 2]
