"""Dispatcher between the façade and the native boundary -- mirror of
/root/reference/polars_bio/range_op_helpers.py:171-399 (``range_operation`` + ``_validate_overlap_input``).

Differences forced by the environment, not by design: the result object is this package's
``RangeResult`` instead of a datafusion.DataFrame, ``"pyarrow.Table"`` is an additional output type,
and ``"polars.*"`` outputs need polars to be importable (it is optional here).
"""
from __future__ import annotations

from typing import Optional

import pyarrow as pa

from ._metadata import set_coordinate_system
from .options import FilterOp, OverlapOutputMode, RangeOp, RangeOptions
from .range_op_io import (RangeResult, _input_schema, pd, pl, range_lazy_scan, range_operation_frame, range_operation_lazy,
                          range_operation_scan)

OUTPUT_TYPES = ["polars.LazyFrame", "polars.DataFrame", "pandas.DataFrame", "datafusion.DataFrame", "pyarrow.Table",
                "pyarrow.RecordBatchReader"]


def _validate_overlap_input(col1, col2, on_cols, suffixes, output_type):
    # range_op_helpers.py:379-399
    assert on_cols is None, "on_cols is not supported yet"
    assert output_type in OUTPUT_TYPES, "Only polars.LazyFrame, polars.DataFrame and pandas DataFrame are supported"


def _tag(result, zero_based: bool):
    try:
        return set_coordinate_system(result, zero_based)
    except TypeError:
        return result


def _get_zero_based_from_filter_op(filter_op) -> bool:
    """FilterOp.Strict = 0-based half-open, FilterOp.Weak = 1-based closed (range_op_helpers.py:27-33)."""
    return filter_op == FilterOp.Strict


def _arrow_schema(df, read_options=None) -> pa.Schema:
    """Arrow schema of any supported input without materialising it (range_op_io.py:311-377 ``_get_schema``): paths are
    opened, not read (the reference reads the whole Parquet file just for its schema)."""
    from .range_op_io import _is_lazyframe_like, _path_reader

    if isinstance(df, str):
        return _path_reader(df, None, read_options).schema
    if isinstance(df, (pa.Table, pa.RecordBatch, pa.RecordBatchReader)):
        return df.schema
    if _is_lazyframe_like(df):
        sch = df.collect_schema()
        return sch.to_arrow() if hasattr(sch, "to_arrow") else sch
    if pl is not None and isinstance(df, pl.DataFrame):
        return df.head(0).to_arrow().schema
    if pd is not None and isinstance(df, pd.DataFrame):
        return pa.Schema.from_pandas(df, preserve_index=False)
    return _input_schema(df)


def _generate_overlap_schema(df1_schema: pa.Schema, df2_schema: pa.Schema, range_options: RangeOptions) -> pa.Schema:
    """Output schema of overlap / nearest with the reference's suffix rule (range_op_helpers.py:56-77): Left mode keeps
    df1's fields; Join mode is every df1 field + suffix 1, then every df2 field + suffix 2."""
    if range_options.overlap_output == OverlapOutputMode.Left:
        return df1_schema
    sfx = range_options.suffixes or ("_1", "_2")
    fields = [pa.field(f"{f.name}{sfx[0]}", f.type) for f in df1_schema] + [pa.field(f"{f.name}{sfx[1]}", f.type) for f in df2_schema]
    return pa.schema(fields)


def _result_schema(df1, df2, range_options: RangeOptions, read_options1=None, read_options2=None) -> pa.Schema:
    """Schema synthesis for deferred results (range_op_helpers.py:200-262, 312-346), as Arrow types.  The engine's own
    output stream is the authority on dtypes (string columns come back as large_utf8); this is what a LazyFrame shows
    before anything ran."""
    op = range_options.range_op
    if op in (RangeOp.CountOverlapsNaive, RangeOp.Coverage):
        # df2's schema: range_op.py swaps df1/df2 for these operations and the engine returns the rows of its `right`
        base = _arrow_schema(df2, read_options2)
        return pa.schema(list(base) + [pa.field("count" if op == RangeOp.CountOverlapsNaive else "coverage", pa.int64())])
    merged = _generate_overlap_schema(_arrow_schema(df1, read_options1), _arrow_schema(df2, read_options2), range_options)
    if op == RangeOp.Nearest and (range_options.compute_distance is None or range_options.compute_distance):
        merged = pa.schema(list(merged) + [pa.field("distance", pa.int64())])
    return merged


def _to_polars_schema(schema: pa.Schema):
    return pl.from_arrow(schema.empty_table()).schema


def _is_lazy_input(df) -> bool:
    from .range_op_io import _is_lazyframe_like

    return _is_lazyframe_like(df) or hasattr(df, "_base_lf")


def range_operation(df1, df2, range_options: RangeOptions, output_type: str, ctx, read_options1=None,
                    read_options2=None, projection_pushdown: bool = True, limit: Optional[int] = None):
    """Runs one binary range operation and converts the result -- the input-kind dispatch of the reference
    (range_op_helpers.py:171-376):

    * two paths            -> ``range_operation_scan`` (streamed files);  LazyFrame output -> ``range_lazy_scan``
    * LazyFrame output     -> ``range_lazy_scan`` (deferred; fresh Arrow streams per execution, ``range_operation_lazy``)
    * anything else, eager -> ``range_operation_frame`` (LazyFrame inputs are collected first, :181-189)

    ``"pyarrow.RecordBatchReader"`` is an extra output type: the streamed result of the lazy path without polars."""
    ctx.sync_options()
    zero_based = _get_zero_based_from_filter_op(range_options.filter_op)
    lazy_out = output_type == "polars.LazyFrame"
    if not lazy_out and output_type != "pyarrow.RecordBatchReader":
        if _is_lazy_input(df1):
            df1 = df1.collect()
        if _is_lazy_input(df2):
            df2 = df2.collect()
    if isinstance(df1, str) and isinstance(df2, str):
        if lazy_out:
            if pl is None:
                raise ImportError("polars is not installed in this environment; use output_type='pyarrow.RecordBatchReader' for a deferred result")
            schema = _to_polars_schema(_result_schema(df1, df2, range_options, read_options1, read_options2))
            return _tag(range_lazy_scan(df1, df2, schema, range_options, ctx, read_options1, read_options2, projection_pushdown), zero_based)
        result = range_operation_scan(ctx, df1, df2, range_options, read_options1, read_options2, limit)
        return convert_result(result, output_type, zero_based)
    if lazy_out:
        if pl is None:
            raise ImportError("polars is not installed in this environment; use output_type='pyarrow.RecordBatchReader' for a deferred result")
        schema = _to_polars_schema(_result_schema(df1, df2, range_options, read_options1, read_options2))
        return _tag(range_lazy_scan(df1, df2, schema, range_options, ctx, projection_pushdown=projection_pushdown), zero_based)
    if output_type == "pyarrow.RecordBatchReader":
        from .range_op_io import _prepare_lazy_stream_input

        (s1, f1), (s2, f2) = (_prepare_lazy_stream_input(df1, range_options.columns_1[0]),
                              _prepare_lazy_stream_input(df2, range_options.columns_2[0]))
        return range_operation_lazy(ctx, f1(), f2(), s1, s2, range_options, limit)._reader
    result: RangeResult = range_operation_frame(ctx, df1, df2, range_options, limit)
    return convert_result(result, output_type, zero_based)


def convert_result(result: RangeResult, output_type: str, zero_based: bool):
    """``convert_result`` of the reference (range_op_helpers.py:359-376): result object -> the requested frame kind."""
    if output_type == "datafusion.DataFrame":
        return result
    if output_type == "pyarrow.Table":
        return _tag(result.to_arrow(), zero_based)
    if output_type == "pandas.DataFrame":
        if pd is None:
            raise ImportError("pandas is not installed. Install pandas or use `polars-bio[pandas]`.")
        return _tag(result.to_pandas(), zero_based)
    if output_type in ("polars.DataFrame", "polars.LazyFrame"):
        if pl is None:
            raise ImportError("polars is not installed in this environment; use output_type='pandas.DataFrame' or 'pyarrow.Table'")
        out = result.to_polars()
        return _tag(out.lazy() if output_type == "polars.LazyFrame" else out, zero_based)
    raise ValueError("Only polars.LazyFrame, polars.DataFrame and pandas.DataFrame are supported")


def unary_operation(df1, df2, range_options: RangeOptions, output_type: str, ctx, view_df=None):
    """merge / cluster / complement / subtract (range_op.py:599-868 -> operation.rs:352-510): the tables go through
    ``pbgpu_range_op`` like the binary operations (Arrow streams in, Arrow stream out; csrc/arrow_bridge.cpp
    ``run_unary``); ``unary_op.py`` holds the same host logic over the device-level calls (tests compare the two)."""
    import dataclasses

    from .range_op_io import range_operation_unary

    ctx.sync_options()
    zero_based = range_options.filter_op == FilterOp.Strict
    op = range_options.range_op
    cols1 = list(range_options.columns_1 or ["chrom", "start", "end"])
    if op in (RangeOp.Merge, RangeOp.Cluster):
        second = None
        opts = dataclasses.replace(range_options, columns_1=cols1, columns_2=cols1)
    elif op == RangeOp.Complement:
        second = view_df
        opts = dataclasses.replace(range_options, columns_1=cols1, columns_2=list(range_options.view_columns or cols1))
    elif op == RangeOp.Subtract:
        second = df2
        opts = dataclasses.replace(range_options, columns_1=cols1, columns_2=list(range_options.columns_2 or cols1))
    else:
        raise ValueError(f"{op!r} is not a unary sweep")
    return convert_result(range_operation_unary(ctx, df1, second, opts), output_type, zero_based)
