"""Device-level driver: HBM-resident int32 columns in, torch tensors out.

torch is plumbing here (device memory, the current CUDA stream, ``torch.distributed`` for the
multi-GPU exchange); every computation is a kernel of libpbgpu.so reached through the C ABI of
``include/pbgpu.h``.  The classes mirror the three providers polars-bio's plan builder
constructs (/root/reference/src/operation.rs:146-158, 253-263, 331-340): an index over the
"indexed" table plus one method per provider.
"""
from __future__ import annotations

import ctypes
from typing import Iterator, Optional, Tuple

import torch

from . import _native
from ._native import check

FILTER_WEAK = 0    # FilterOp::Weak   (src/option.rs:97) 1-based closed intervals
FILTER_STRICT = 1  # FilterOp::Strict (src/option.rs:98) 0-based half-open intervals
NO_PARTNER = 0xFFFFFFFF


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _col(t: torch.Tensor, device) -> torch.Tensor:
    if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous():
        raise TypeError("device-level API takes contiguous int32 CUDA tensors")
    if t.device != device:
        raise ValueError(f"column on {t.device}, index on {device}")
    return t


class DeviceIndex:
    """Search structure over the indexed (build) table -- the COITrees replacement.

    ``contig`` holds dictionary codes shared with the probe side (0..n_contigs-1; anything else
    is a null key that never matches); ``start``/``end`` are int32 coordinates.
    """

    def __init__(self, contig: torch.Tensor, start: torch.Tensor, end: torch.Tensor, n_contigs: int,
                 row_ids: Optional[torch.Tensor] = None):
        """``row_ids`` (int32 storage of uint32 ids, one per row): what indexed row i is called in pair buffers and
        nearest partners instead of i -- the global row ids of a sharded table, so results need no translation."""
        if not torch.cuda.is_available():
            raise RuntimeError("polars_bio_b200 needs a CUDA device (no CPU fallback)")
        self.device = contig.device
        self._L = _native.lib()
        self._h = ctypes.c_void_p()
        self.n_contigs = int(n_contigs)
        c, s, e = (_col(x, self.device) for x in (contig, start, end))
        ids = None if row_ids is None else _col(row_ids, self.device)
        if ids is not None and ids.numel() != c.numel():
            raise ValueError("row_ids must have one entry per row")
        with torch.cuda.device(self.device):
            check(self._L.pbgpu_index_build_ids(c.data_ptr(), s.data_ptr(), e.data_ptr(), ids.data_ptr() if ids is not None else None,
                                                c.numel(), self.n_contigs, _stream_ptr(self.device), ctypes.byref(self._h)))

    def close(self):
        """Stream-ordered release on the CURRENT stream of the index's device: the memory is reused only after the
        kernels enqueued there so far (the ones reading this index) have run -- also on non-blocking side streams."""
        if getattr(self, "_h", None) is not None and self._h.value:
            try:
                sp = _stream_ptr(self.device)
            except Exception:  # interpreter shutdown: torch may be gone
                sp = None
            if sp is None:
                self._L.pbgpu_index_free(self._h)
            else:
                self._L.pbgpu_index_free_async(self._h, sp)
            self._h = ctypes.c_void_p()

    __del__ = close

    @property
    def rows(self) -> int:
        return int(self._L.pbgpu_index_rows(self._h))

    @property
    def nbytes(self) -> int:
        return int(self._L.pbgpu_index_bytes(self._h))

    # -- CountOverlapsProvider ---------------------------------------------------------------
    def count_overlaps(self, contig, start, end, filter_op: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        c, s, e = (_col(x, self.device) for x in (contig, start, end))
        n = c.numel()
        if out is None:
            out = torch.empty(n, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._L.pbgpu_count_overlaps(self._h, c.data_ptr(), s.data_ptr(), e.data_ptr(), n, filter_op,
                                               out.data_ptr(), _stream_ptr(self.device)))
        return out

    def coverage(self, contig, start, end, filter_op: int) -> torch.Tensor:
        c, s, e = (_col(x, self.device) for x in (contig, start, end))
        n = c.numel()
        out = torch.empty(n, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._L.pbgpu_coverage(self._h, c.data_ptr(), s.data_ptr(), e.data_ptr(), n, filter_op,
                                         out.data_ptr(), _stream_ptr(self.device)))
        return out

    # -- OverlapProvider ---------------------------------------------------------------------
    def overlap_pairs(self, contig, start, end, filter_op: int, probe_ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Two-pass count-then-emit.  Returns (probe_rows, build_rows): int32 tensors holding uint32 row ids (or the
        ids given by ``probe_ids`` / the index's ``row_ids``).  The pairs of one probe are contiguous and ordered by
        (start, row) of the indexed partner; probes come in row order, or grouped by coordinate bin when the index is far
        beyond the L2 (csrc/bins.cuh) -- pair order is not part of the contract (the reference's tests sort)."""
        c, s, e = (_col(x, self.device) for x in (contig, start, end))
        n = c.numel()
        ids = None if probe_ids is None else _col(probe_ids, self.device)
        if ids is not None and ids.numel() != n:
            raise ValueError("probe_ids must have one entry per probe row")
        plan = ctypes.c_void_p()
        total = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            sp = _stream_ptr(self.device)
            check(self._L.pbgpu_overlap_count_ids(self._h, c.data_ptr(), s.data_ptr(), e.data_ptr(),
                                                  ids.data_ptr() if ids is not None else None, n, filter_op, sp,
                                                  ctypes.byref(plan), ctypes.byref(total)))
            try:
                p = torch.empty(total.value, dtype=torch.int32, device=self.device)
                b = torch.empty(total.value, dtype=torch.int32, device=self.device)
                check(self._L.pbgpu_overlap_emit(plan, p.data_ptr(), b.data_ptr(), sp))
            finally:
                # the plan's scratch is read by the emit kernel: released in stream order behind it, no host sync
                self._L.pbgpu_overlap_plan_free_async(plan, sp)
        return p, b

    def overlap_pairs_stream(self, contig, start, end, filter_op: int, max_pairs: int = 1 << 24) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Streaming sink (SURVEY.md 7 step 6 / BASELINE config 5): the same pairs as :meth:`overlap_pairs`, in the
        same order, yielded chunk by chunk from TWO device buffers of at most ``max_pairs`` pairs (a chunk is a run of
        whole 256-probe blocks; one oversized block still gets its own chunk).  Pass 1 runs once over all probes;
        pass 2 runs per chunk.  The yielded tensors are views of the ring slot: consume (or copy) them before
        advancing the generator twice."""
        import numpy as np

        c, s, e = (_col(x, self.device) for x in (contig, start, end))
        n = c.numel()
        plan = ctypes.c_void_p()
        total = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            sp = _stream_ptr(self.device)
            check(self._L.pbgpu_overlap_count(self._h, c.data_ptr(), s.data_ptr(), e.data_ptr(), n, filter_op, sp,
                                              ctypes.byref(plan), ctypes.byref(total)))
            try:
                nblk = int(self._L.pbgpu_overlap_plan_blocks(plan))
                offs = np.zeros(nblk + 1, dtype=np.uint64)
                check(self._L.pbgpu_overlap_plan_block_offsets(plan, offs.ctypes.data, sp))
                offs = offs.astype(np.int64)
                cap = int(max(max_pairs, int(np.diff(offs).max()) if nblk else 0, 1))
                ring = [(torch.empty(cap, dtype=torch.int32, device=self.device),
                         torch.empty(cap, dtype=torch.int32, device=self.device)) for _ in range(2)]
                lo, k = 0, 0
                while lo < nblk:
                    # largest hi with offs[hi] - offs[lo] <= cap (at least one block)
                    hi = int(np.searchsorted(offs, offs[lo] + cap, side="right")) - 1
                    hi = max(hi, lo + 1)
                    cnt = int(offs[hi] - offs[lo])
                    if cnt:
                        p, b = ring[k & 1]
                        check(self._L.pbgpu_overlap_emit_blocks(plan, lo, hi, p.data_ptr(), b.data_ptr(), sp))
                        k += 1
                        yield p[:cnt], b[:cnt]
                    lo = hi
                torch.cuda.current_stream(self.device).synchronize()
            finally:
                torch.cuda.current_stream(self.device).synchronize()
                self._L.pbgpu_overlap_plan_free(plan)

    # -- NearestProvider ---------------------------------------------------------------------
    def nearest(self, contig, start, end, filter_op: int, k: int = 1, include_overlaps: bool = True,
                compute_distance: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        c, s, e = (_col(x, self.device) for x in (contig, start, end))
        n = c.numel()
        partner = torch.empty((n, k), dtype=torch.int32, device=self.device)
        dist = torch.empty((n, k), dtype=torch.int64, device=self.device) if compute_distance else None
        with torch.cuda.device(self.device):
            check(self._L.pbgpu_nearest(self._h, c.data_ptr(), s.data_ptr(), e.data_ptr(), n, filter_op, k,
                                        int(include_overlaps), partner.data_ptr(),
                                        dist.data_ptr() if dist is not None else None, _stream_ptr(self.device)))
        return partner, dist


# ---- unary sweeps (MergeProvider / ClusterProvider / SubtractProvider, operation.rs:352-510) ------------------
def _intervals_out(L, handle, device, want):
    """Copies the requested columns of a pbgpu_intervals result into torch tensors and releases it in stream order."""
    sp = _stream_ptr(device)
    try:
        n = int(L.pbgpu_intervals_rows(handle))
        out = {w: torch.empty(n, dtype=torch.int64 if w == "count" else torch.int32, device=device) for w in want}
        args = [out[w].data_ptr() if w in out and n else None for w in ("contig", "row", "start", "end", "count")]
        check(L.pbgpu_intervals_copy(handle, *args, sp))
        return tuple(out[w] for w in want)
    finally:
        L.pbgpu_intervals_free(handle, sp)


def merge_intervals(contig, start, end, n_contigs: int, filter_op: int, min_dist: int = 0):
    """Merged intervals of one table, ordered by (contig code, start): (contig, start, end) int32, n_intervals int64."""
    device = contig.device
    c, s, e = (_col(x, device) for x in (contig, start, end))
    L = _native.lib()
    h = ctypes.c_void_p()
    with torch.cuda.device(device):
        check(L.pbgpu_merge(c.data_ptr(), s.data_ptr(), e.data_ptr(), c.numel(), int(n_contigs), filter_op, int(min_dist),
                            _stream_ptr(device), ctypes.byref(h)))
        return _intervals_out(L, h, device, ("contig", "start", "end", "count"))


def cluster_intervals(contig, start, end, n_contigs: int, filter_op: int, min_dist: int = 0):
    """Per input row: cluster id (int64; numbered in (contig code, start) order, -1 for null-keyed rows) and the
    cluster's start / end (int32).  Returns (cluster, cluster_start, cluster_end, n_clusters)."""
    device = contig.device
    c, s, e = (_col(x, device) for x in (contig, start, end))
    m = c.numel()
    cid = torch.empty(m, dtype=torch.int64, device=device)
    cs = torch.empty(m, dtype=torch.int32, device=device)
    ce = torch.empty(m, dtype=torch.int32, device=device)
    k = ctypes.c_int64(0)
    with torch.cuda.device(device):
        check(_native.lib().pbgpu_cluster(c.data_ptr(), s.data_ptr(), e.data_ptr(), m, int(n_contigs), filter_op, int(min_dist),
                                          cid.data_ptr(), cs.data_ptr(), ce.data_ptr(), ctypes.byref(k), _stream_ptr(device)))
    return cid, cs, ce, int(k.value)


def subtract_intervals(l_contig, l_start, l_end, r_contig, r_start, r_end, n_contigs: int, filter_op: int):
    """Pieces of every left row that no right row covers, ordered by (left row, start): (left_row, start, end) int32
    tensors (left_row holds uint32 row ids).  complement = the view table on the left."""
    device = l_contig.device
    lc, ls, le = (_col(x, device) for x in (l_contig, l_start, l_end))
    rc, rs, re = (_col(x, device) for x in (r_contig, r_start, r_end))
    L = _native.lib()
    h = ctypes.c_void_p()
    with torch.cuda.device(device):
        check(L.pbgpu_subtract(lc.data_ptr(), ls.data_ptr(), le.data_ptr(), lc.numel(), rc.data_ptr(), rs.data_ptr(), re.data_ptr(),
                               rc.numel(), int(n_contigs), filter_op, _stream_ptr(device), ctypes.byref(h)))
        return _intervals_out(L, h, device, ("row", "start", "end"))
