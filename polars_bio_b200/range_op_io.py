"""Input/outputs of the range-operation boundary.

Mirrors /root/reference/polars_bio/range_op_io.py (``_df_to_reader`` :398-418, the Arrow C stream
export, the result adapters) and the PyO3 entry point ``range_operation_frame``
(/root/reference/src/lib.rs:79-88), which here is a ctypes call into ``pbgpu_range_op``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterator, List, Optional

import pyarrow as pa

from . import _native
from .constants import BATCH_SIZE, INTERVAL_JOIN_LOW_MEMORY
from .options import FilterOp, OverlapOutputMode, RangeOp, RangeOptions

try:  # optional front ends
    import pandas as pd
except ImportError:  # pragma: no cover
    pd = None
try:
    import polars as pl
except ImportError:
    pl = None


class _CStream(ctypes.Structure):  # struct ArrowArrayStream: five pointers
    _fields_ = [("get_schema", ctypes.c_void_p), ("get_next", ctypes.c_void_p), ("get_last_error", ctypes.c_void_p),
                ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p)]


def _read_path(path: str) -> pa.Table:
    """Path inputs (lib.rs:216-270 -> scan.rs:615-627): Parquet (file or directory) and CSV with header."""
    low = path.lower()
    if low.endswith(".csv"):
        import pyarrow.csv as pcsv

        return pcsv.read_csv(path)
    if low.endswith(".bed"):
        import pyarrow.csv as pcsv

        return pcsv.read_csv(path, read_options=pcsv.ReadOptions(column_names=["chrom", "start", "end"]),
                             parse_options=pcsv.ParseOptions(delimiter="\t"))
    if low.endswith(".parquet") or os.path.isdir(path) or "*" in path:
        import pyarrow.parquet as pq

        if "*" in path:
            import glob

            return pa.concat_tables([pq.read_table(p) for p in sorted(glob.glob(path))])
        return pq.read_table(path)
    raise ValueError(f"unsupported input path '{path}' (Parquet, CSV and BED are supported)")


def _df_to_reader(df, contig_col: Optional[str] = None) -> pa.RecordBatchReader:
    """Any supported frame -> Arrow RecordBatchReader (zero-copy where the source allows)."""
    if isinstance(df, pa.RecordBatchReader):
        return df
    if isinstance(df, pa.Table):
        return df.to_reader()
    if isinstance(df, pa.RecordBatch):
        return pa.Table.from_batches([df]).to_reader()
    if isinstance(df, str):
        return _read_path(df).to_reader()
    if pl is not None and isinstance(df, pl.LazyFrame):
        df = df.collect()
    if pl is not None and isinstance(df, pl.DataFrame):
        return df.to_arrow().to_reader()
    if pd is not None and isinstance(df, pd.DataFrame):
        return pa.Table.from_pandas(df, preserve_index=False).to_reader()
    if hasattr(df, "__arrow_c_stream__"):
        return pa.RecordBatchReader.from_stream(df)
    raise TypeError(f"unsupported input type {type(df).__name__}")


def _input_schema(df) -> pa.Schema:
    if isinstance(df, (pa.Table, pa.RecordBatch, pa.RecordBatchReader)):
        return df.schema
    return _df_to_reader(df).schema


class RangeResult:
    """What ``range_operation_frame`` returns.  Exposes the five methods the reference's Python layer
    calls on its datafusion.DataFrame result (seam B1): ``to_polars`` / ``to_pandas`` /
    ``schema`` / ``select`` / ``execute_stream`` (+ ``to_arrow``).  Backed by the engine's output
    ArrowArrayStream, consumed lazily batch by batch."""

    def __init__(self, reader: pa.RecordBatchReader, columns: Optional[List[str]] = None):
        self._reader = reader
        self._columns = columns

    def schema(self) -> pa.Schema:
        s = self._reader.schema
        return s if self._columns is None else pa.schema([s.field(c) for c in self._columns])

    def select(self, *columns) -> "RangeResult":
        cols = list(columns[0]) if len(columns) == 1 and isinstance(columns[0], (list, tuple)) else list(columns)
        return RangeResult(self._reader, cols)

    def execute_stream(self) -> Iterator[pa.RecordBatch]:
        for b in self._reader:
            yield b if self._columns is None else b.select(self._columns)

    def to_arrow(self) -> pa.Table:
        t = self._reader.read_all()
        return t if self._columns is None else t.select(self._columns)

    def to_pandas(self):
        return self.to_arrow().to_pandas()

    def to_polars(self):
        if pl is None:
            raise ImportError("polars is not installed")
        return pl.from_arrow(self.to_arrow())

    def count(self) -> int:
        return sum(b.num_rows for b in self._reader)


def _c_opts(ro: RangeOptions, emit: int, limit: Optional[int], ctx) -> _native.PbRangeOptions:
    o = _native.PbRangeOptions()
    o.range_op = int(ro.range_op)
    o.filter_op = int(ro.filter_op if ro.filter_op is not None else FilterOp.Weak)
    mode = ro.overlap_output if ro.overlap_output is not None else OverlapOutputMode.Join
    # (Join, _) -> Join; (Left, False) -> Left; (Left, True) -> LeftDistinct   (operation.rs:229-233)
    o.output_mode = 0 if mode == OverlapOutputMode.Join else (2 if ro.distinct_output else 1)
    o.emit = emit
    c1 = ro.columns_1 or ["chrom", "start", "end"]
    c2 = ro.columns_2 or ["chrom", "start", "end"]
    for i in range(3):
        o.cols1[i] = str(c1[i]).encode()
        o.cols2[i] = str(c2[i]).encode()
    sfx = ro.suffixes or ("_1", "_2")
    o.suffixes[0] = sfx[0].encode()
    o.suffixes[1] = sfx[1].encode()
    o.nearest_k = int(ro.nearest_k or 1)
    o.include_overlaps = 1 if (ro.include_overlaps is None or ro.include_overlaps) else 0
    o.compute_distance = 1 if (ro.compute_distance is None or ro.compute_distance) else 0
    o.limit = int(limit or 0)
    batch = 1 << 20
    low_mem = ro.overlap_low_memory
    if low_mem is None and ctx is not None:
        low_mem = (ctx.get_option(INTERVAL_JOIN_LOW_MEMORY) or "").lower() == "true"
    if low_mem:  # low_memory caps the output batch size (range_op.py:168)
        batch = int((ctx.get_option(BATCH_SIZE) if ctx is not None else None) or 8192)
    o.max_batch_rows = batch
    o.device = -1
    return o


def range_operation_frame(py_ctx, df1, df2, range_options: RangeOptions, limit: Optional[int] = None,
                          emit: int = 0) -> RangeResult:
    """``polars_bio.polars_bio.range_operation_frame`` (src/lib.rs:79-88): two Arrow stream exporters in,
    a lazily consumed result out.  The engine moves both input streams (released before this returns)."""
    if range_options.range_op not in (RangeOp.Overlap, RangeOp.Nearest, RangeOp.Coverage, RangeOp.CountOverlapsNaive):
        raise ValueError(f"{range_options.range_op!r} is not on the GPU hot path")
    r1, r2 = _df_to_reader(df1), _df_to_reader(df2)
    s1, s2, so = _CStream(), _CStream(), _CStream()
    r1._export_to_c(ctypes.addressof(s1))
    r2._export_to_c(ctypes.addressof(s2))
    opts = _c_opts(range_options, emit, limit, py_ctx)
    rc = _native.lib().pbgpu_range_op(ctypes.addressof(s1), ctypes.addressof(s2), ctypes.byref(opts), ctypes.addressof(so))
    _native.check(rc)
    return RangeResult(pa.RecordBatchReader._import_from_c(ctypes.addressof(so)))
