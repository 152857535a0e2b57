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


def _bed_layout(path: str):
    """(lines to skip, column names) of a BED file: `track` / `browser` / `#` lines are skipped, the first data line
    decides the column count (BED3 .. BED12; the first three are chrom / start / end)."""
    names = ["chrom", "start", "end", "name", "score", "strand", "thickStart", "thickEnd", "itemRgb", "blockCount",
             "blockSizes", "blockStarts"]
    skip, ncol = 0, 3
    with open(path, "rt") as f:
        for line in f:
            if line.startswith(("track", "browser", "#")) or not line.strip():
                skip += 1
                continue
            ncol = len(line.rstrip("\r\n").split("\t"))
            break
    if ncol < 3:
        raise ValueError(f"'{path}': a BED line needs at least 3 tab-separated columns")
    cols = names[:ncol] if ncol <= len(names) else names + [f"column_{i + 1}" for i in range(len(names), ncol)]
    return skip, cols


def _path_reader(path: str, batch_rows: Optional[int] = None, read_options=None) -> pa.RecordBatchReader:
    """Path inputs (lib.rs:216-270 -> scan.rs:586-627) as a STREAMING reader: Parquet (file, directory or glob), CSV with
    header, BED.  Batches are produced on demand, so a probe-side file is never materialised as a whole."""
    low = path.lower()
    if low.endswith((".csv", ".csv.gz")):
        import pyarrow.csv as pcsv

        return pcsv.open_csv(path)
    if low.endswith((".bed", ".bed.gz")):
        import pyarrow.csv as pcsv

        skip, cols = _bed_layout(path) if not low.endswith(".gz") else (0, ["chrom", "start", "end"])
        return pcsv.open_csv(path, read_options=pcsv.ReadOptions(column_names=cols, skip_rows=skip),
                             parse_options=pcsv.ParseOptions(delimiter="\t"))
    if low.endswith(".parquet") or os.path.isdir(path) or "*" in path:
        import pyarrow.dataset as ds

        if "*" in path:
            import glob

            files = sorted(glob.glob(path))
            if not files:
                raise FileNotFoundError(path)
            dset = ds.dataset(files, format="parquet")
        else:
            dset = ds.dataset(path, format="parquet")
        kw = {"batch_size": int(batch_rows)} if batch_rows else {}
        return dset.scanner(**kw).to_reader()
    raise ValueError(f"unsupported input path '{path}' (Parquet, CSV and BED are supported)")


def _read_path(path: str) -> pa.Table:
    return _path_reader(path).read_all()


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
    o.min_dist = int(ro.min_dist or 0)
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


def range_operation_unary(py_ctx, df1, df2, range_options: RangeOptions) -> RangeResult:
    """merge / cluster / complement / subtract through the same C entry as the binary operations (``pbgpu_range_op``;
    reference: do_merge / do_cluster / do_complement / do_subtract, src/operation.rs:352-510).  ``df2``: the second
    table of subtract, the view table of complement (its interval columns are ``columns_2``), else None."""
    if range_options.range_op not in (RangeOp.Merge, RangeOp.Cluster, RangeOp.Complement, RangeOp.Subtract):
        raise ValueError(f"{range_options.range_op!r} is not a unary sweep")
    s1, s2, so = _CStream(), _CStream(), _CStream()
    _df_to_reader(df1)._export_to_c(ctypes.addressof(s1))
    if df2 is not None:
        _df_to_reader(df2)._export_to_c(ctypes.addressof(s2))
    opts = _c_opts(range_options, 0, None, py_ctx)
    rc = _native.lib().pbgpu_range_op(ctypes.addressof(s1), ctypes.addressof(s2) if df2 is not None else None, ctypes.byref(opts),
                                      ctypes.addressof(so))
    _native.check(rc)
    return RangeResult(pa.RecordBatchReader._import_from_c(ctypes.addressof(so)))


# ------------------------------------------------------------------------------------------------
# Streamed entry points: range_operation_lazy / range_operation_scan (src/lib.rs:154-166, 216-228)
# ------------------------------------------------------------------------------------------------
PROBE_CHUNK_ROWS = "bio.gpu_probe_chunk_rows"  # session option: iterated-side rows per pbgpu_range_probe call
DEFAULT_PROBE_CHUNK_ROWS = 1 << 22


def _as_reader(stream, schema: Optional[pa.Schema] = None) -> pa.RecordBatchReader:
    """Anything that exports the Arrow C stream protocol (or already is a reader) -> RecordBatchReader."""
    if isinstance(stream, pa.RecordBatchReader):
        return stream
    if isinstance(stream, pa.Table):
        return stream.to_reader()
    if hasattr(stream, "__arrow_c_stream__"):
        return pa.RecordBatchReader.from_stream(stream, schema=schema) if schema is not None else pa.RecordBatchReader.from_stream(stream)
    return _df_to_reader(stream)


def _chunks(reader: pa.RecordBatchReader, chunk_rows: int):
    """Group the reader's batches into chunks of about ``chunk_rows`` rows (a larger batch is sliced): the unit of one
    probe call.  Yields lists of record batches; holds at most one chunk."""
    buf, rows = [], 0
    for b in reader:
        off = 0
        while off < b.num_rows:
            take = min(b.num_rows - off, chunk_rows - rows)
            buf.append(b.slice(off, take))
            rows += take
            off += take
            if rows >= chunk_rows:
                yield buf
                buf, rows = [], 0
    if buf:
        yield buf


class RangeSession:
    """``pbgpu_range_open`` / ``_probe`` / ``_close``: the indexed side resident on the device, probed chunk by chunk."""

    def __init__(self, indexed: pa.RecordBatchReader, opts: "_native.PbRangeOptions"):
        self._L = _native.lib()
        self._h = ctypes.c_void_p()
        self._opts = opts  # keeps the byte strings alive
        s = _CStream()
        indexed._export_to_c(ctypes.addressof(s))
        _native.check(self._L.pbgpu_range_open(ctypes.addressof(s), ctypes.byref(opts), ctypes.byref(self._h)))

    def probe(self, chunk: pa.RecordBatchReader) -> pa.RecordBatchReader:
        si, so = _CStream(), _CStream()
        chunk._export_to_c(ctypes.addressof(si))
        _native.check(self._L.pbgpu_range_probe(self._h, ctypes.addressof(si), ctypes.addressof(so)))
        return pa.RecordBatchReader._import_from_c(ctypes.addressof(so))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.pbgpu_range_close(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close


def _chunk_rows_option(py_ctx) -> int:
    v = py_ctx.get_option(PROBE_CHUNK_ROWS) if py_ctx is not None else None
    try:
        return max(1, int(v)) if v else DEFAULT_PROBE_CHUNK_ROWS
    except ValueError:
        return DEFAULT_PROBE_CHUNK_ROWS


def range_operation_lazy(py_ctx, stream1, stream2, schema1, schema2, range_options: RangeOptions, limit: Optional[int] = None,
                         emit: int = 0) -> RangeResult:
    """``polars_bio.polars_bio.range_operation_lazy`` (src/lib.rs:154-166): Arrow C streams in, a lazily produced result
    out.  The indexed side (df2 for overlap / nearest, df1 for count_overlaps / coverage -- the same roles as the eager
    call) is collected and indexed once; the iterated side is pulled batch by batch WHILE the result is consumed: every
    ``bio.gpu_probe_chunk_rows`` rows (default 4 Mi) become one probe call against the resident index, so host staging and
    pinned memory are bounded by the chunk, not by the table (scan.rs:103-139, 320-357)."""
    if range_options.range_op not in (RangeOp.Overlap, RangeOp.Nearest, RangeOp.Coverage, RangeOp.CountOverlapsNaive):
        raise ValueError(f"{range_options.range_op!r} is not on the GPU hot path")
    r1, r2 = _as_reader(stream1, schema1), _as_reader(stream2, schema2)
    iter_is_left = range_options.range_op in (RangeOp.Overlap, RangeOp.Nearest)
    indexed, iterated = (r2, r1) if iter_is_left else (r1, r2)
    opts = _c_opts(range_options, emit, None, py_ctx)
    session = RangeSession(indexed, opts)
    chunk_rows = _chunk_rows_option(py_ctx)
    try:
        out_schema = session.probe(pa.RecordBatchReader.from_batches(iterated.schema, [])).schema
    except Exception:
        session.close()
        raise

    def produce():
        left = None if limit is None else int(limit)
        base = 0  # emit=1: row ids of the iterated side are relative to the chunk
        try:
            for chunk in _chunks(iterated, chunk_rows):
                n_chunk = sum(b.num_rows for b in chunk)
                res = session.probe(pa.RecordBatchReader.from_batches(iterated.schema, chunk))
                for b in res:
                    if emit == 1 and base:
                        b = _shift_rows(b, base, iter_is_left)
                    if left is not None:
                        if left <= 0:
                            return
                        if b.num_rows > left:
                            b = b.slice(0, left)
                        left -= b.num_rows
                    yield b
                base += n_chunk
                if left is not None and left <= 0:
                    return
        finally:
            session.close()

    return RangeResult(pa.RecordBatchReader.from_batches(out_schema, produce()))


def _shift_rows(batch: pa.RecordBatch, base: int, iter_is_left: bool) -> pa.RecordBatch:
    """emit=1 (index pairs): chunk-relative row ids of the iterated side -> table row ids."""
    import pyarrow.compute as pc

    col = 0 if iter_is_left else 1
    arrays = list(batch.columns)
    arrays[col] = pc.add(arrays[col].cast(pa.uint64()), pa.scalar(base, pa.uint64())).cast(arrays[col].type)
    return pa.RecordBatch.from_arrays(arrays, schema=batch.schema)


def range_operation_scan(py_ctx, df_path_or_table1: str, df_path_or_table2: str, range_options: RangeOptions,
                         read_options1=None, read_options2=None, limit: Optional[int] = None, emit: int = 0) -> RangeResult:
    """``polars_bio.polars_bio.range_operation_scan`` (src/lib.rs:216-228 -> scan.rs:586-627): both inputs are paths
    (Parquet file / directory / glob, CSV, BED).  Each file is opened as a streaming reader; the iterated one is never
    materialised (it goes through :func:`range_operation_lazy`)."""
    if not isinstance(df_path_or_table1, str) or not isinstance(df_path_or_table2, str):
        raise TypeError("range_operation_scan takes two paths")
    batch = int((py_ctx.get_option(BATCH_SIZE) if py_ctx is not None else None) or 8192)
    r1 = _path_reader(df_path_or_table1, max(batch, 1 << 16), read_options1)
    r2 = _path_reader(df_path_or_table2, max(batch, 1 << 16), read_options2)
    return range_operation_lazy(py_ctx, r1, r2, r1.schema, r2.schema, range_options, limit, emit)


# ------------------------------------------------------------------------------------------------
# Lazy sources: range_lazy_scan / _range_source / _prepare_lazy_stream_input (range_op_io.py:31-283)
# ------------------------------------------------------------------------------------------------
def _is_lazyframe_like(df) -> bool:
    """Polars LazyFrames or wrappers exposing ``collect_batches`` (range_op_io.py:177-182)."""
    if pl is not None and isinstance(df, pl.LazyFrame):
        return True
    return hasattr(df, "collect_batches") and hasattr(df, "collect_schema")


def _prepare_lazy_stream_input(df, contig_col: str, batch_size: Optional[int] = None):
    """(arrow schema, stream factory) of one input (range_op_io.py:185-283).  A factory, because an Arrow C stream can be
    consumed once and the lazy result may be collected several times."""
    if isinstance(df, str):
        rd = _path_reader(df, batch_size)
        schema = rd.schema
        del rd
        return schema, (lambda: _path_reader(df, batch_size))
    if _is_lazyframe_like(df):
        schema = df.collect_schema()
        arrow_schema = schema.to_arrow() if hasattr(schema, "to_arrow") else schema

        def stream_factory():
            batches = df.collect_batches(lazy=True, engine="streaming", chunk_size=batch_size)
            return getattr(batches, "_inner", batches)

        return arrow_schema, stream_factory
    if isinstance(df, pa.RecordBatchReader):
        raise ValueError("a RecordBatchReader can be consumed once: pass a Table, a frame or a path for lazy execution")
    table = _df_to_reader(df, contig_col).read_all()
    return table.schema, (lambda: table.to_reader())


class LazyRangeSource:
    """What :func:`range_lazy_scan` returns when polars is not importable: the same deferred execution (nothing runs until
    it is consumed; every consumption re-executes from fresh streams) behind a minimal frame-like surface."""

    def __init__(self, source, schema: pa.Schema):
        self._source, self._schema = source, schema
        self._columns, self._n_rows = None, None

    def collect_schema(self) -> pa.Schema:
        s = self._schema
        return s if self._columns is None else pa.schema([s.field(c) for c in self._columns])

    schema = property(collect_schema)

    def select(self, *columns) -> "LazyRangeSource":
        cols = list(columns[0]) if len(columns) == 1 and isinstance(columns[0], (list, tuple)) else list(columns)
        out = LazyRangeSource(self._source, self._schema)
        out._columns, out._n_rows = cols, self._n_rows
        return out

    def head(self, n: int) -> "LazyRangeSource":
        out = LazyRangeSource(self._source, self._schema)
        out._columns, out._n_rows = self._columns, int(n)
        return out

    limit = head

    def collect_batches(self, **_):
        return self._source(self._columns, None, self._n_rows, None)

    def collect(self) -> pa.Table:
        return pa.Table.from_batches(list(self.collect_batches()), schema=self.collect_schema())

    def lazy(self) -> "LazyRangeSource":
        return self


def range_lazy_scan(df_1, df_2, schema, range_options: RangeOptions, ctx, read_options1=None, read_options2=None,
                    projection_pushdown: bool = True):
    """Deferred range operation (range_op_io.py:31-174): returns a ``polars.LazyFrame`` backed by an IO-plugin source
    when polars is importable, else a :class:`LazyRangeSource`.  Nothing is read or computed until the result is
    consumed; paths use :func:`range_operation_scan`, everything else fresh Arrow streams through
    :func:`range_operation_lazy`; ``with_columns`` (projection pushdown) and ``n_rows`` reach the engine, ``predicate``
    is applied per batch."""
    use_file_paths = isinstance(df_1, str) and isinstance(df_2, str)
    batch_size = int(ctx.get_option(BATCH_SIZE) or 8192)
    if use_file_paths:
        lazy_sources = None
    else:
        col1, col2 = range_options.columns_1[0], range_options.columns_2[0]
        lazy_sources = (_prepare_lazy_stream_input(df_1, col1, batch_size), _prepare_lazy_stream_input(df_2, col2, batch_size))

    def _range_source(with_columns, predicate, _n_rows, _batch_size):
        projected = None
        if projection_pushdown and with_columns is not None:
            projected = _column_names(with_columns)
        alg = getattr(range_options, "overlap_alg", None)
        if alg is not None:
            from .logging import logger

            logger.info("Optimizing into IntervalJoinExec using %s algorithm", alg)
        if use_file_paths:
            res = range_operation_scan(ctx, df_1, df_2, range_options, read_options1, read_options2, _n_rows)
        else:
            (schema1, factory1), (schema2, factory2) = lazy_sources
            res = range_operation_lazy(ctx, factory1(), factory2(), schema1, schema2, range_options, _n_rows)
        pushed = False
        if projected:
            try:
                res = res.select(projected)
                pushed = True
            except Exception:
                pushed = False
        for b in res.execute_stream():
            if pl is not None:
                df = pl.DataFrame(b)
                if predicate is not None:
                    df = df.filter(predicate)
                if with_columns is not None and not pushed:
                    df = df.select(with_columns)
                yield df
            else:
                if predicate is not None:
                    b = b.filter(predicate)  # a pyarrow.compute Expression
                if with_columns is not None and not pushed:
                    b = b.select(_column_names(with_columns))
                yield b

    if pl is not None:
        from polars.io.plugins import register_io_source

        return register_io_source(_range_source, schema=schema)
    return LazyRangeSource(_range_source, schema)


def _column_names(with_columns) -> List[str]:
    """Column names of a projection: a list of names, or polars expressions naming columns."""
    out = []
    for c in (with_columns if isinstance(with_columns, (list, tuple)) else [with_columns]):
        if isinstance(c, str):
            out.append(c)
        elif hasattr(c, "meta"):
            out.extend(c.meta.root_names())
        else:
            out.append(str(c))
    return out
