"""CPU oracle for the interval-join hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``polars_bio_b200``) never does; it fails loudly when its CUDA library is missing.

Two independent restatements live here:

* :mod:`oracle.interval_oracle` (C, ``liboracle.so``) -- augmented-interval-tree queries,
  the published algorithm of the reference's third-party dependency (coitrees 0.4.0 behind
  datafusion-bio-function-ranges v0.11.0; Cargo.lock:1091-1094,1830-1842).
* :mod:`oracle.oracle_np` (numpy) -- sort + rank-difference / window restatement that shares
  no code with the C one; the two are cross-checked in ``tests/test_oracle_golden.py`` and
  both are pinned by the reference's fixtures.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile ``liboracle.so`` with the committed Makefile (gcc only, no reference sources)."""
    src = os.path.join(_HERE, "interval_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    lib = ctypes.CDLL(_LIB_PATH)
    lib.pbo_index_build.restype = ctypes.c_void_p
    lib.pbo_index_build.argtypes = [_i32p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int32]
    lib.pbo_index_free.argtypes = [ctypes.c_void_p]
    lib.pbo_count_overlaps.argtypes = [ctypes.c_void_p, _i32p, _i32p, _i32p, ctypes.c_int64,
                                       ctypes.c_int, _i64p, ctypes.c_int]
    lib.pbo_overlap_pairs.restype = ctypes.c_int64
    lib.pbo_overlap_pairs.argtypes = [ctypes.c_void_p, _i32p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int,
                                      _u32p, _u32p, ctypes.c_int64, ctypes.c_int]
    lib.pbo_coverage.argtypes = [ctypes.c_void_p, _i32p, _i32p, _i32p, ctypes.c_int64,
                                 ctypes.c_int, _i64p, ctypes.c_int]
    lib.pbo_nearest.argtypes = [ctypes.c_void_p, _i32p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int,
                                ctypes.c_int64, ctypes.c_int, _u32p, _i64p, ctypes.c_int]
    lib.pbo_brute_pairs.restype = ctypes.c_int64
    lib.pbo_brute_pairs.argtypes = [_i32p, _i32p, _i32p, ctypes.c_int64, _i32p, _i32p, _i32p, ctypes.c_int64,
                                    ctypes.c_int, _u32p, _u32p, ctypes.c_int64]
    lib.pbo_max_threads.restype = ctypes.c_int
    _lib = lib
    return lib


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def max_threads() -> int:
    return int(_load().pbo_max_threads())


class Index:
    """Per-contig augmented interval trees over the indexed (build) table."""

    def __init__(self, chrom, start, end, n_contigs: int):
        lib = _load()
        self._c, self._s, self._e = _c(chrom), _c(start), _c(end)
        self.n_contigs = int(n_contigs)
        self._h = lib.pbo_index_build(_p(self._c, _i32p), _p(self._s, _i32p), _p(self._e, _i32p),
                                      len(self._c), self.n_contigs)

    def __del__(self):
        if getattr(self, "_h", None):
            _load().pbo_index_free(self._h)
            self._h = None

    def count_overlaps(self, chrom, start, end, strict: bool, threads: int = 1) -> np.ndarray:
        c, s, e = _c(chrom), _c(start), _c(end)
        out = np.zeros(len(c), dtype=np.int64)
        _load().pbo_count_overlaps(self._h, _p(c, _i32p), _p(s, _i32p), _p(e, _i32p), len(c),
                                   int(strict), _p(out, _i64p), threads)
        return out

    def coverage(self, chrom, start, end, strict: bool, threads: int = 1) -> np.ndarray:
        c, s, e = _c(chrom), _c(start), _c(end)
        out = np.zeros(len(c), dtype=np.int64)
        _load().pbo_coverage(self._h, _p(c, _i32p), _p(s, _i32p), _p(e, _i32p), len(c),
                             int(strict), _p(out, _i64p), threads)
        return out

    def overlap_total(self, chrom, start, end, strict: bool, threads: int = 1) -> int:
        c, s, e = _c(chrom), _c(start), _c(end)
        return int(_load().pbo_overlap_pairs(self._h, _p(c, _i32p), _p(s, _i32p), _p(e, _i32p), len(c),
                                             int(strict), None, None, 0, threads))

    def overlap_pairs(self, chrom, start, end, strict: bool, threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """(probe_row, build_row) uint32 arrays, ordered by probe row then (start,row) of the partner."""
        c, s, e = _c(chrom), _c(start), _c(end)
        lib = _load()
        args = (self._h, _p(c, _i32p), _p(s, _i32p), _p(e, _i32p), len(c), int(strict))
        cap = max(1024, 2 * len(c))  # one pass when the guess holds; the C side reports the true total
        while True:
            a = np.empty(cap, dtype=np.uint32)
            b = np.empty(cap, dtype=np.uint32)
            total = int(lib.pbo_overlap_pairs(*args, _p(a, _u32p), _p(b, _u32p), cap, threads))
            if total <= cap:
                return a[:total], b[:total]
            cap = total

    def nearest(self, chrom, start, end, strict: bool, k: int = 1, include_overlaps: bool = True,
                threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """(partner[n,k] uint32 with 0xFFFFFFFF = none, distance[n,k] int64 with -1 = none)."""
        c, s, e = _c(chrom), _c(start), _c(end)
        n = len(c)
        ob = np.empty((n, k), dtype=np.uint32)
        od = np.empty((n, k), dtype=np.int64)
        _load().pbo_nearest(self._h, _p(c, _i32p), _p(s, _i32p), _p(e, _i32p), n, int(strict), k,
                            int(include_overlaps), _p(ob, _u32p), _p(od, _i64p), threads)
        return ob, od


def brute_pairs(lc, ls, le, rc, rs, re, strict: bool) -> Tuple[np.ndarray, np.ndarray]:
    """O(N*M) nested loop over the bare predicate; tiny inputs only."""
    lc, ls, le, rc, rs, re = map(_c, (lc, ls, le, rc, rs, re))
    lib = _load()
    args = (_p(lc, _i32p), _p(ls, _i32p), _p(le, _i32p), len(lc), _p(rc, _i32p), _p(rs, _i32p), _p(re, _i32p),
            len(rc), int(strict))
    total = int(lib.pbo_brute_pairs(*args, None, None, 0))
    a = np.empty(total, dtype=np.uint32)
    b = np.empty(total, dtype=np.uint32)
    if total:
        lib.pbo_brute_pairs(*args, _p(a, _u32p), _p(b, _u32p), total)
    return a, b


def encode_contigs(*cols):
    """Shared dictionary over several contig columns -> (int32 code arrays..., names).

    ``None`` / NaN entries get code -1 (null key: never matches).
    """
    import pandas as pd

    cat = pd.unique(pd.concat([pd.Series(c, dtype="object") for c in cols], ignore_index=True).dropna())
    lut = {v: i for i, v in enumerate(cat)}
    outs = [np.fromiter((lut.get(v, -1) for v in c), dtype=np.int32, count=len(c)) for c in cols]
    return (*outs, list(cat))
