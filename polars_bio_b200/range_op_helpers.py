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
from .options import FilterOp, RangeOp, RangeOptions
from .range_op_io import RangeResult, pd, pl, range_operation_frame

OUTPUT_TYPES = ["polars.LazyFrame", "polars.DataFrame", "pandas.DataFrame", "datafusion.DataFrame", "pyarrow.Table"]


def _validate_overlap_input(col1, col2, on_cols, suffixes, output_type):
    # range_op_helpers.py:379-399
    assert on_cols is None, "on_cols is not supported yet"
    assert output_type in OUTPUT_TYPES, "Only polars.LazyFrame, polars.DataFrame and pandas DataFrame are supported"


def _tag(result, zero_based: bool):
    try:
        return set_coordinate_system(result, zero_based)
    except TypeError:
        return result


def range_operation(df1, df2, range_options: RangeOptions, output_type: str, ctx, read_options1=None,
                    read_options2=None, projection_pushdown: bool = True, limit: Optional[int] = None):
    """Runs one binary range operation and converts the result (range_op_helpers.py:171-376).

    Every input kind (path / pandas / polars eager or lazy / pyarrow) funnels into
    ``range_operation_frame``; LazyFrame output wraps the eager result lazily (streaming the engine's
    output batches through a Polars IO plugin needs polars >= 1.0, used when available)."""
    ctx.sync_options()
    zero_based = range_options.filter_op == FilterOp.Strict
    if output_type == "datafusion.DataFrame":
        return range_operation_frame(ctx, df1, df2, range_options, limit)
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
    """merge / cluster / complement / subtract (range_op.py:599-868 -> operation.rs:352-510): every input kind is
    collected into an Arrow table, the sweep runs on the device (unary_op.py), the result table is converted."""
    from . import unary_op
    from .range_op_io import _df_to_reader

    ctx.sync_options()
    zero_based = range_options.filter_op == FilterOp.Strict
    t1 = _df_to_reader(df1).read_all()
    cols1 = list(range_options.columns_1 or ["chrom", "start", "end"])
    op = range_options.range_op
    if op == RangeOp.Merge:
        out = unary_op.merge_table(t1, cols1, range_options.filter_op, int(range_options.min_dist or 0))
    elif op == RangeOp.Cluster:
        out = unary_op.cluster_table(t1, cols1, range_options.filter_op, int(range_options.min_dist or 0))
    elif op == RangeOp.Complement:
        view = None if view_df is None else _df_to_reader(view_df).read_all()
        out = unary_op.complement_table(t1, cols1, range_options.filter_op, view, range_options.view_columns)
    elif op == RangeOp.Subtract:
        t2 = _df_to_reader(df2).read_all()
        out = unary_op.subtract_table(t1, t2, cols1, list(range_options.columns_2 or cols1), range_options.filter_op)
    else:
        raise ValueError(f"{op!r} is not a unary sweep")
    return convert_result(RangeResult(out.to_reader()), output_type, zero_based)
