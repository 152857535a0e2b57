"""numpy twin of the CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

An independent restatement (shares no code with interval_oracle.c) built on the identity the
reference itself uses for its second count_overlaps algorithm
(polars_bio/range_op.py:548-594: ``count = starts_rank - ends_rank`` over the two sorted event
lists, tie order flipping with the coordinate system) and on the bare predicate of
docs/developers.md:549-552.  Pure-Python loops appear only in the brute-force nearest /
coverage helpers, which are meant for inputs of a few hundred rows.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def _hit(strict: bool, a_s, a_e, b_s, b_e):
    if strict:
        return (a_s < b_e) & (a_e > b_s)
    return (a_s <= b_e) & (a_e >= b_s)


def count_overlaps(lc, ls, le, rc, rs, re, strict: bool) -> np.ndarray:
    """Per (lc,ls,le) row: number of (rc,rs,re) rows on the same contig that overlap it.

    Rank-difference identity; rows where it is not valid (empty/inverted intervals on either
    side) are recounted with the bare predicate.
    """
    lc, ls, le, rc, rs, re = (np.asarray(x, dtype=np.int64) for x in (lc, ls, le, rc, rs, re))
    out = np.zeros(len(lc), dtype=np.int64)
    for c in np.unique(lc[lc >= 0]):
        qi = np.nonzero(lc == c)[0]
        bi = np.nonzero(rc == c)[0]
        if len(bi) == 0:
            continue
        S = np.sort(rs[bi])
        E = np.sort(re[bi])
        if strict:
            cnt = np.searchsorted(S, le[qi], side="left") - np.searchsorted(E, ls[qi], side="right")
            bad_q = ~(ls[qi] < le[qi])
        else:
            cnt = np.searchsorted(S, le[qi], side="right") - np.searchsorted(E, ls[qi], side="left")
            bad_q = ~(ls[qi] <= le[qi])
        if np.any(rs[bi] > re[bi]):
            bad_q[:] = True
        for j in np.nonzero(bad_q)[0]:
            cnt[j] = int(np.count_nonzero(_hit(strict, ls[qi[j]], le[qi[j]], rs[bi], re[bi])))
        out[qi] = cnt
    return out


def overlap_pairs(lc, ls, le, rc, rs, re, strict: bool) -> Tuple[np.ndarray, np.ndarray]:
    """All (probe_row, build_row) pairs, sorted by (probe_row, build_row)."""
    lc, ls, le, rc, rs, re = (np.asarray(x, dtype=np.int64) for x in (lc, ls, le, rc, rs, re))
    pa, pb = [], []
    for c in np.unique(lc[lc >= 0]):
        qi = np.nonzero(lc == c)[0]
        bi = np.nonzero(rc == c)[0]
        if len(bi) == 0:
            continue
        order = np.argsort(rs[bi], kind="stable")
        bi = bi[order]
        S, E = rs[bi], re[bi]
        pm = np.maximum.accumulate(E)
        if strict:
            hi = np.searchsorted(S, le[qi], side="left")
            lo = np.searchsorted(pm, ls[qi], side="right")
        else:
            hi = np.searchsorted(S, le[qi], side="right")
            lo = np.searchsorted(pm, ls[qi], side="left")
        lo = np.minimum(lo, hi)
        w = hi - lo
        tot = int(w.sum())
        if tot == 0:
            continue
        qrep = np.repeat(np.arange(len(qi)), w)
        base = np.repeat(lo - np.concatenate(([0], np.cumsum(w)[:-1])), w)
        cand = np.arange(tot) + base
        keep = _hit(strict, ls[qi][qrep], le[qi][qrep], S[cand], E[cand])
        pa.append(qi[qrep[keep]])
        pb.append(bi[cand[keep]])
    if not pa:
        return np.empty(0, np.uint32), np.empty(0, np.uint32)
    a = np.concatenate(pa)
    b = np.concatenate(pb)
    o = np.lexsort((b, a))
    return a[o].astype(np.uint32), b[o].astype(np.uint32)


def brute_nearest(lc, ls, le, rc, rs, re, strict: bool, k: int = 1, include_overlaps: bool = True):
    """Pure-Python k-nearest with the oracle's documented ordering:
    overlapping partners first (if included), then (distance, start, row)."""
    n = len(lc)
    ob = np.full((n, k), 0xFFFFFFFF, dtype=np.uint32)
    od = np.full((n, k), -1, dtype=np.int64)
    for i in range(n):
        if lc[i] < 0:
            continue
        cand = []
        for j in range(len(rc)):
            if rc[j] != lc[i]:
                continue
            h = bool(_hit(strict, int(ls[i]), int(le[i]), int(rs[j]), int(re[j])))
            if h and not include_overlaps:
                continue
            d = 0 if h else max(0, int(rs[j]) - int(le[i]), int(ls[i]) - int(re[j]))
            cand.append((0 if h else 1, d, int(rs[j]), j))
        cand.sort()
        for t, (_, d, _, j) in enumerate(cand[:k]):
            ob[i, t] = j
            od[i, t] = d
    return ob, od


def brute_coverage(lc, ls, le, rc, rs, re, strict: bool) -> np.ndarray:
    """Positions of each (lc,ls,le) row covered by the union of same-contig (rc,rs,re) rows."""
    out = np.zeros(len(lc), dtype=np.int64)
    for i in range(len(lc)):
        if lc[i] < 0:
            continue
        pieces = []
        for j in range(len(rc)):
            if rc[j] != lc[i] or not bool(_hit(strict, int(ls[i]), int(le[i]), int(rs[j]), int(re[j]))):
                continue
            s = max(int(ls[i]), int(rs[j]))
            e = min(int(le[i]), int(re[j])) + (0 if strict else 1)
            if e > s:
                pieces.append((s, e))
        pieces.sort()
        tot, cs, ce = 0, None, None
        for s, e in pieces:
            if cs is None:
                cs, ce = s, e
            elif s <= ce:
                ce = max(ce, e)
            else:
                tot += ce - cs
                cs, ce = s, e
        if cs is not None:
            tot += ce - cs
        out[i] = tot
    return out
