"""Host side of the unary sweeps: merge / cluster / complement / subtract.

Mirrors the Rust glue of /root/reference/src/operation.rs:352-510 (do_merge / do_cluster / do_complement /
do_subtract) over the device-level C ABI (``pbgpu_merge`` / ``pbgpu_cluster`` / ``pbgpu_subtract`` in
include/pbgpu.h): Arrow tables in, contig strings dictionary-encoded (codes in lexicographic name order, so that
results come out ordered by contig name and cluster ids count the way bioframe's do,
tests/test_bioframe.py:392-411), int32 columns to HBM through pinned memory, kernels, result columns back, the
reference's output schemas assembled with pyarrow:

  merge       (contig, start, end) named like the input's interval columns, Int64 positions, + ``n_intervals`` Int64
              (tests/_expected.py:174-181)
  cluster     every input column + ``cluster``, ``cluster_start``, ``cluster_end`` Int64
              (tests/test_partitioned_range_operation_regressions.py:49-59)
  complement  (contig, start, end) named like the input's interval columns, Int64 (…regressions.py:33-39)
  subtract    every df1 column, the interval columns holding the remaining pieces as Int64 (…regressions.py:41-47)

Rows with a null contig / start / end take no part (dropped from cluster's output as well): parity unpinned.
There is no CPU fallback: without libpbgpu.so and a CUDA device these calls raise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from . import _native
from .options import FilterOp

I64_MAX = np.iinfo(np.int64).max
I32_MAX = np.iinfo(np.int32).max


def _contig_strings(col: pa.ChunkedArray) -> Tuple[np.ndarray, List[str]]:
    """-> (int64 local codes with -1 for nulls, local names)."""
    arr = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    if pa.types.is_dictionary(arr.type):
        names = arr.dictionary.cast(pa.large_string()).to_pylist()
        codes = arr.indices.cast(pa.int64())
    else:
        if not (pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type) or pa.types.is_string_view(arr.type)):
            raise _native.PbgpuError(5, f"contig column has type {arr.type}; a string type is required")
        enc = pc.dictionary_encode(arr.cast(pa.large_string()))
        names = enc.dictionary.to_pylist()
        codes = enc.indices.cast(pa.int64())
    np_codes = codes.fill_null(-1).to_numpy(zero_copy_only=False).astype(np.int64, copy=False)
    return np_codes, names


def _positions(col: pa.ChunkedArray, what: str) -> Tuple[np.ndarray, np.ndarray]:
    """-> (int32 values, validity) with the int32-domain check of the Arrow bridge (PBGPU_ERANGE)."""
    arr = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    if not pa.types.is_integer(arr.type):
        raise _native.PbgpuError(5, f"{what} column has type {arr.type}; an integer type is required")
    valid = np.ones(len(arr), bool) if arr.null_count == 0 else arr.is_valid().to_numpy(zero_copy_only=False)
    v = arr.fill_null(0).to_numpy(zero_copy_only=False)
    if v.dtype != np.int32:
        if len(v) and (int(v.max()) > I32_MAX or (v.dtype.kind == "i" and int(v.min()) < -I32_MAX - 1)):
            raise _native.PbgpuError(4, f"{what} holds a coordinate outside the int32 domain")
        v = v.astype(np.int32)
    return v, valid


class _Keys:
    """(contig code, start, end) of one table as int32 numpy columns; null-keyed rows carry code -1."""

    def __init__(self, table: pa.Table, cols: Sequence[str], side: str):
        for c in cols:
            if c not in table.column_names:
                raise _native.PbgpuError(5, f"{side} table has no column '{c}'")
        self.local, self.names = _contig_strings(table.column(cols[0]))
        self.start, v1 = _positions(table.column(cols[1]), f"{side}.{cols[1]}")
        self.end, v2 = _positions(table.column(cols[2]), f"{side}.{cols[2]}")
        self.local = np.where(v1 & v2, self.local, -1)
        self.code: Optional[np.ndarray] = None

    def bind(self, names_sorted: List[str]):
        rank = {n: i for i, n in enumerate(names_sorted)}
        lut = np.array([rank[n] for n in self.names] + [-1], dtype=np.int32)  # [-1] -> -1
        self.code = lut[self.local]


def _shared_names(*keys: _Keys) -> List[str]:
    names = sorted(set().union(*[set(k.names) for k in keys]))
    for k in keys:
        k.bind(names)
    return names


def _to_device(*cols: np.ndarray):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("polars_bio_b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    out = []
    for c in cols:
        a = np.ascontiguousarray(c, dtype=np.int32)
        if not a.flags.writeable:  # zero-copy views of Arrow buffers are read-only; torch wants to own a writable array
            a = a.copy()
        out.append(torch.from_numpy(a).pin_memory().to(dev, non_blocking=True))
    return out


def _contig_column(names: List[str], codes: np.ndarray, like: pa.DataType) -> pa.Array:
    out = pa.array(names, type=pa.large_string()).take(pa.array(codes.astype(np.int64)))
    return out.cast(like) if (pa.types.is_string(like) or pa.types.is_large_string(like)) else out


def _fo(filter_op: FilterOp) -> int:
    return int(filter_op)


def merge_table(table: pa.Table, cols: Sequence[str], filter_op: FilterOp, min_dist: int = 0) -> pa.Table:
    from . import engine

    k = _Keys(table, cols, "input")
    names = _shared_names(k)
    c, s, e = _to_device(k.code, k.start, k.end)
    mc, ms, me, mn = engine.merge_intervals(c, s, e, len(names), _fo(filter_op), min_dist)
    mc, ms, me, mn = (x.cpu().numpy() for x in (mc, ms, me, mn))
    return pa.table({cols[0]: _contig_column(names, mc, table.schema.field(cols[0]).type),
                     cols[1]: pa.array(ms.astype(np.int64)), cols[2]: pa.array(me.astype(np.int64)),
                     "n_intervals": pa.array(mn.astype(np.int64))})


def cluster_table(table: pa.Table, cols: Sequence[str], filter_op: FilterOp, min_dist: int = 0) -> pa.Table:
    from . import engine

    k = _Keys(table, cols, "input")
    names = _shared_names(k)
    c, s, e = _to_device(k.code, k.start, k.end)
    cid, cs, ce, _ = engine.cluster_intervals(c, s, e, len(names), _fo(filter_op), min_dist)
    cid, cs, ce = (x.cpu().numpy() for x in (cid, cs, ce))
    out = table.append_column("cluster", pa.array(cid)).append_column("cluster_start", pa.array(cs.astype(np.int64))) \
               .append_column("cluster_end", pa.array(ce.astype(np.int64)))
    if (cid < 0).any():  # null-keyed rows take no part
        out = out.filter(pa.array(cid >= 0))
    return out


def _subtract(lk: _Keys, rk: _Keys, n_contigs: int, filter_op: FilterOp):
    from . import engine

    lc, ls, le = _to_device(lk.code, lk.start, lk.end)
    rc, rs, re = _to_device(rk.code, rk.start, rk.end)
    row, fs, fe = engine.subtract_intervals(lc, ls, le, rc, rs, re, n_contigs, _fo(filter_op))
    return row.cpu().numpy().view(np.uint32), fs.cpu().numpy().astype(np.int64), fe.cpu().numpy().astype(np.int64)


def subtract_table(left: pa.Table, right: pa.Table, cols1: Sequence[str], cols2: Sequence[str], filter_op: FilterOp) -> pa.Table:
    lk, rk = _Keys(left, cols1, "left"), _Keys(right, cols2, "right")
    names = _shared_names(lk, rk)
    row, fs, fe = _subtract(lk, rk, len(names), filter_op)
    out = left.take(pa.array(row.astype(np.int64)))
    i1, i2 = out.column_names.index(cols1[1]), out.column_names.index(cols1[2])
    out = out.set_column(i1, cols1[1], pa.array(fs)).set_column(i2, cols1[2], pa.array(fe))
    return out


def complement_table(table: pa.Table, cols: Sequence[str], filter_op: FilterOp, view: Optional[pa.Table] = None,
                     view_cols: Optional[Sequence[str]] = None) -> pa.Table:
    k = _Keys(table, cols, "input")
    if view is None:
        # every contig present spans [0, INT64_MAX) (polars_bio/range_op.py:726-729): swept as [0, INT32_MAX], the
        # open end put back afterwards (no int32 coordinate reaches it)
        present = sorted({k.names[i] for i in np.unique(k.local[k.local >= 0])})
        view = pa.table({"c": pa.array(present, type=pa.large_string()), "s": pa.array(np.zeros(len(present), np.int64)),
                         "e": pa.array(np.full(len(present), I32_MAX, np.int64))})
        view_cols, open_end = ("c", "s", "e"), True
    else:
        view_cols, open_end = tuple(view_cols or cols), False
    vk = _Keys(view, view_cols, "view")
    names = _shared_names(vk, k)
    row, fs, fe = _subtract(vk, k, len(names), filter_op)
    if open_end:
        fe = np.where(fe == I32_MAX, I64_MAX, fe)
    contig = _contig_column(names, vk.code[row.astype(np.int64)], table.schema.field(cols[0]).type)
    return pa.table({cols[0]: contig, cols[1]: pa.array(fs), cols[2]: pa.array(fe)})
