"""ctypes binding of libpbgpu.so -- the only way this package computes anything.

There is deliberately no CPU fallback: if the CUDA library is missing or fails to load, import
of the product path raises.  (The CPU oracle under ``oracle/`` is test infrastructure and is
never imported from here.)

The prototypes below are exactly ``include/pbgpu.h``; ``tests/test_abi_symbols.py`` checks that
every symbol declared there is exported by the built library.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PBGPU_LIB") or os.path.join(_HERE, "libpbgpu.so")  # PBGPU_LIB: an experimental build (A/B of compile-time parameters)

_i32p = ctypes.c_void_p  # device pointers travel as integers
_lib = None


class PbgpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"pbgpu error {code}: {message}")
        self.code = code


class PbRangeOptions(ctypes.Structure):
    """Mirror of ``PbRangeOptions`` (include/pbgpu.h), itself a mirror of RangeOptions
    (/root/reference/src/option.rs:8-41)."""

    _fields_ = [
        ("range_op", ctypes.c_int32),
        ("filter_op", ctypes.c_int32),
        ("output_mode", ctypes.c_int32),
        ("emit", ctypes.c_int32),
        ("cols1", ctypes.c_char_p * 3),
        ("cols2", ctypes.c_char_p * 3),
        ("suffixes", ctypes.c_char_p * 2),
        ("nearest_k", ctypes.c_uint64),
        ("include_overlaps", ctypes.c_int32),
        ("compute_distance", ctypes.c_int32),
        ("limit", ctypes.c_uint64),
        ("max_batch_rows", ctypes.c_uint32),
        ("device", ctypes.c_int32),
        ("sink_pairs", ctypes.c_uint64),
        ("min_dist", ctypes.c_int64),
    ]


class PbPeerStep(ctypes.Structure):
    """Mirror of ``pbgpu_peer_step`` (include/pbgpu.h)."""

    _fields_ = [
        ("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("n_tables", ctypes.c_int32), ("n_contigs", ctypes.c_int32),
        ("phases", ctypes.c_int32), ("reserved", ctypes.c_int32), ("step", ctypes.c_uint64),
        ("contig", ctypes.c_void_p * 4), ("start", ctypes.c_void_p * 4), ("end", ctypes.c_void_p * 4),
        ("rows", ctypes.c_int64 * 4),
        ("arena_base", ctypes.c_void_p), ("ctl_base", ctypes.c_void_p), ("cap_rows", ctypes.c_void_p),
        ("d_hist", ctypes.c_void_p), ("d_owner", ctypes.c_void_p), ("d_dst", ctypes.c_void_p),
        ("d_result", ctypes.c_void_p), ("h_result", ctypes.c_void_p), ("d_status", ctypes.c_void_p),
        ("d_scratch", ctypes.c_void_p), ("scratch_bytes", ctypes.c_uint64),
    ]


class StageTimes(ctypes.Structure):
    _fields_ = [("partition_sort_ns", ctypes.c_uint64), ("count_ns", ctypes.c_uint64),
                ("scan_ns", ctypes.c_uint64), ("emit_ns", ctypes.c_uint64), ("count_overlaps_ns", ctypes.c_uint64),
                ("bin_ns", ctypes.c_uint64), ("unbin_ns", ctypes.c_uint64)]


def lib() -> ctypes.CDLL:
    """Load libpbgpu.so (built in-tree by ``__graft_entry__.build()``); fail loudly otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "polars_bio_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
    L.pbgpu_last_error.restype = ctypes.c_char_p
    L.pbgpu_version.restype = ctypes.c_char_p
    L.pbgpu_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.pbgpu_launch_count.restype = u64
    L.pbgpu_index_build.argtypes = [vp, vp, vp, i64, i32, vp, ctypes.POINTER(vp)]
    L.pbgpu_index_build_ids.argtypes = [vp, vp, vp, vp, i64, i32, vp, ctypes.POINTER(vp)]
    L.pbgpu_index_free.argtypes = [vp]
    L.pbgpu_index_free.restype = None
    L.pbgpu_index_free_async.argtypes = [vp, vp]
    L.pbgpu_index_free_async.restype = None
    L.pbgpu_index_rows.argtypes = [vp]
    L.pbgpu_index_rows.restype = i64
    L.pbgpu_index_bytes.argtypes = [vp]
    L.pbgpu_index_bytes.restype = ctypes.c_size_t
    L.pbgpu_count_overlaps.argtypes = [vp, vp, vp, vp, i64, ctypes.c_int, vp, vp]
    L.pbgpu_coverage.argtypes = [vp, vp, vp, vp, i64, ctypes.c_int, vp, vp]
    L.pbgpu_overlap_count.argtypes = [vp, vp, vp, vp, i64, ctypes.c_int, vp, ctypes.POINTER(vp), ctypes.POINTER(i64)]
    L.pbgpu_overlap_count_ids.argtypes = [vp, vp, vp, vp, vp, i64, ctypes.c_int, vp, ctypes.POINTER(vp), ctypes.POINTER(i64)]
    L.pbgpu_overlap_emit.argtypes = [vp, vp, vp, vp]
    L.pbgpu_overlap_plan_blocks.argtypes = [vp]
    L.pbgpu_overlap_plan_blocks.restype = i64
    L.pbgpu_overlap_plan_block_offsets.argtypes = [vp, vp, vp]
    L.pbgpu_overlap_emit_blocks.argtypes = [vp, i64, i64, vp, vp, vp]
    L.pbgpu_overlap_plan_counts.argtypes = [vp]
    L.pbgpu_overlap_plan_counts.restype = vp
    L.pbgpu_overlap_plan_free.argtypes = [vp]
    L.pbgpu_overlap_plan_free.restype = None
    L.pbgpu_overlap_plan_free_async.argtypes = [vp, vp]
    L.pbgpu_overlap_plan_free_async.restype = None
    L.pbgpu_nearest.argtypes = [vp, vp, vp, vp, i64, ctypes.c_int, i64, ctypes.c_int, vp, vp, vp]
    L.pbgpu_pack_by_owner.argtypes = [vp, vp, vp, i64, vp, i32, i32, u32, vp, vp, vp]
    L.pbgpu_gather_i32.argtypes = [vp, vp, i64, vp, vp]
    L.pbgpu_contig_histogram.argtypes = [vp, i64, i32, vp, vp]
    L.pbgpu_unpack_records.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.pbgpu_translate_rows.argtypes = [vp, i64, vp, vp, vp]
    L.pbgpu_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp), vp]
    L.pbgpu_peer_free.argtypes = [vp]
    L.pbgpu_peer_open.argtypes = [vp, ctypes.POINTER(vp)]
    L.pbgpu_peer_close.argtypes = [vp]
    L.pbgpu_peer_plan.argtypes = [vp, vp, u64, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.pbgpu_peer_begin.argtypes = [ctypes.POINTER(PbPeerStep), vp]
    L.pbgpu_peer_scratch_bytes.argtypes = [ctypes.POINTER(PbPeerStep)]
    L.pbgpu_peer_scratch_bytes.restype = ctypes.c_size_t
    L.pbgpu_peer_table.argtypes = [ctypes.POINTER(PbPeerStep), i32, vp]
    L.pbgpu_peer_ctl_bytes.argtypes = [i32, i32, i32]
    L.pbgpu_peer_ctl_bytes.restype = ctypes.c_size_t
    L.pbgpu_peer_scatter.argtypes = [vp, vp, vp, i64, vp, i32, i32, vp, vp, vp, vp]
    L.pbgpu_intervals_rows.argtypes = [vp]
    L.pbgpu_intervals_rows.restype = i64
    L.pbgpu_intervals_columns.argtypes = [vp] + [ctypes.POINTER(vp)] * 5
    L.pbgpu_intervals_copy.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.pbgpu_intervals_free.argtypes = [vp, vp]
    L.pbgpu_intervals_free.restype = None
    L.pbgpu_merge.argtypes = [vp, vp, vp, i64, i32, ctypes.c_int, i64, vp, ctypes.POINTER(vp)]
    L.pbgpu_cluster.argtypes = [vp, vp, vp, i64, i32, ctypes.c_int, i64, vp, vp, vp, ctypes.POINTER(i64), vp]
    L.pbgpu_subtract.argtypes = [vp, vp, vp, i64, vp, vp, vp, i64, i32, ctypes.c_int, vp, ctypes.POINTER(vp)]
    L.pbgpu_last_stage_times.argtypes = [ctypes.POINTER(StageTimes)]
    L.pbgpu_range_op.argtypes = [vp, vp, ctypes.POINTER(PbRangeOptions), vp]
    L.pbgpu_range_open.argtypes = [vp, ctypes.POINTER(PbRangeOptions), ctypes.POINTER(vp)]
    L.pbgpu_range_probe.argtypes = [vp, vp, vp]
    L.pbgpu_range_close.argtypes = [vp]
    L.pbgpu_range_close.restype = None
    L.pbgpu_pinned_stats.argtypes = [ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.c_int]
    L.pbgpu_pinned_stats.restype = None
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise PbgpuError(rc, lib().pbgpu_last_error().decode("utf-8", "replace"))


def pinned_stats(reset_peak: bool = False):
    """(bytes of page-locked staging in use, high-water mark since the last reset) of the Arrow level."""
    busy, peak = ctypes.c_uint64(0), ctypes.c_uint64(0)
    lib().pbgpu_pinned_stats(ctypes.byref(busy), ctypes.byref(peak), 1 if reset_peak else 0)
    return int(busy.value), int(peak.value)


def launch_count() -> int:
    return int(lib().pbgpu_launch_count())


def stage_times() -> dict:
    t = StageTimes()
    check(lib().pbgpu_last_stage_times(ctypes.byref(t)))
    return {k: int(getattr(t, k)) for k, _ in StageTimes._fields_}
