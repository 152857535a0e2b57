"""User-facing interval operations on the GPU hot path -- the API of
/root/reference/polars_bio/range_op.py:114-597 (``IntervalOperations.overlap / nearest / coverage /
count_overlaps``), same argument names, defaults, validation and input swapping.

The coordinate system (0-based -> FilterOp.Strict, 1-based -> FilterOp.Weak; range_op.py:83-84) is read
from frame metadata exactly like the reference.  ``algorithm`` is accepted for source compatibility;
whatever its value the join runs on the B200 engine (the reference's choices -- Coitrees, IntervalTree,
ArrayIntervalTree, Lapper, SuperIntervals -- all produce the same rows, tests/test_overlap_algorithms.py).
"""
from __future__ import annotations

from typing import Literal, Union

from ._metadata import validate_coordinate_system_single, validate_coordinate_systems
from .constants import DEFAULT_INTERVAL_COLUMNS
from .context import ctx
from .logging import logger
from .options import FilterOp, OverlapOutputMode, RangeOp, RangeOptions
from .range_op_helpers import _validate_overlap_input, range_operation, unary_operation

__all__ = ["overlap", "nearest", "count_overlaps", "coverage", "merge", "cluster", "complement", "subtract"]

DEFAULT_OUTPUT_TYPE = "polars.LazyFrame"


def _parse_overlap_output_mode(overlap_output: str) -> OverlapOutputMode:
    normalized = overlap_output.lower()
    if normalized == "join":
        return OverlapOutputMode.Join
    if normalized == "left":
        return OverlapOutputMode.Left
    raise ValueError("overlap_output must be either 'join' or 'left'")  # tests/test_overlap_output_mode.py:190-197


# `algorithm=` of pb.overlap (range_op.py:124,167; src/operation.rs:39-50 stores it in bio.interval_join_algorithm): the
# reference's CPU search structures plus "gpu".  Every accepted name runs on the B200 engine -- they all produce the same
# rows (tests/test_overlap_algorithms.py) -- and the name is recorded in the session like the reference does.
KNOWN_ALGORITHMS = ("gpu", "coitrees", "intervaltree", "arrayintervaltree", "lapper", "superintervals", "coitreesnearest")


def _select_algorithm(algorithm) -> str:
    name = "gpu" if algorithm is None else str(algorithm)
    if name.lower() not in KNOWN_ALGORITHMS:
        raise ValueError(f"unknown interval join algorithm {algorithm!r}; available: gpu, Coitrees, IntervalTree, "
                         "ArrayIntervalTree, Lapper, SuperIntervals")
    from .constants import INTERVAL_JOIN_ALGORITHM

    ctx.set_option(INTERVAL_JOIN_ALGORITHM, name)
    return name


def _filter_op(df1, df2) -> FilterOp:
    return FilterOp.Strict if validate_coordinate_systems(df1, df2, ctx) else FilterOp.Weak


def _filter_op_single(df) -> FilterOp:  # range_op.py:87-111
    return FilterOp.Strict if validate_coordinate_system_single(df, ctx) else FilterOp.Weak


class IntervalOperations:

    @staticmethod
    def overlap(df1, df2, suffixes: tuple = ("_1", "_2"), on_cols=None, cols1=["chrom", "start", "end"],
                cols2=["chrom", "start", "end"], algorithm: str = "Coitrees", low_memory: bool = False,
                overlap_output: Literal["join", "left"] = "join", distinct_output: bool = False,
                output_type: str = DEFAULT_OUTPUT_TYPE, read_options1=None, read_options2=None,
                projection_pushdown: bool = True):
        """Find pairs of overlapping genomic intervals (range_op.py:117-256).

        ``overlap_output="join"``: every df1 column suffixed ``suffixes[0]`` then every df2 column suffixed
        ``suffixes[1]`` (operation.rs:272-292).  ``"left"``: df1 rows that overlap, original names, one row per
        matching df2 row; with ``distinct_output=True`` each overlapping df1 row once, by row identity."""
        _validate_overlap_input(cols1, cols2, on_cols, suffixes, output_type)
        mode = _parse_overlap_output_mode(overlap_output)
        filter_op = _filter_op(df1, df2)
        cols1 = DEFAULT_INTERVAL_COLUMNS if cols1 is None else list(cols1)
        cols2 = DEFAULT_INTERVAL_COLUMNS if cols2 is None else list(cols2)
        algorithm = _select_algorithm(algorithm)
        logger.info("Optimizing into IntervalJoinExec using %s algorithm (B200 engine)", algorithm)
        opts = RangeOptions(range_op=RangeOp.Overlap, filter_op=filter_op, suffixes=tuple(suffixes), columns_1=cols1,
                            columns_2=cols2, overlap_alg=algorithm, overlap_low_memory=low_memory,
                            overlap_output=mode, distinct_output=distinct_output)
        return range_operation(df1, df2, opts, output_type, ctx, read_options1, read_options2, projection_pushdown)

    @staticmethod
    def nearest(df1, df2, suffixes: tuple = ("_1", "_2"), on_cols=None, cols1=["chrom", "start", "end"],
                cols2=["chrom", "start", "end"], k: int = 1, overlap: bool = True, distance: bool = True,
                output_type: str = DEFAULT_OUTPUT_TYPE, read_options=None, projection_pushdown: bool = True):
        """For every df1 row the k closest df2 rows on the same contig (range_op.py:259-340): df1 columns
        (suffix 1), df2 columns (suffix 2), then ``distance`` (operation.rs:170-195)."""
        _validate_overlap_input(cols1, cols2, on_cols, suffixes, output_type)
        filter_op = _filter_op(df1, df2)
        cols1 = DEFAULT_INTERVAL_COLUMNS if cols1 is None else list(cols1)
        cols2 = DEFAULT_INTERVAL_COLUMNS if cols2 is None else list(cols2)
        opts = RangeOptions(range_op=RangeOp.Nearest, filter_op=filter_op, suffixes=tuple(suffixes), columns_1=cols1,
                            columns_2=cols2, nearest_k=k, include_overlaps=overlap, compute_distance=distance)
        return range_operation(df1, df2, opts, output_type, ctx, read_options, projection_pushdown=projection_pushdown)

    @staticmethod
    def coverage(df1, df2, suffixes: tuple = ("_1", "_2"), on_cols=None, cols1=["chrom", "start", "end"],
                 cols2=["chrom", "start", "end"], output_type: str = DEFAULT_OUTPUT_TYPE, read_options=None,
                 projection_pushdown: bool = True):
        """df1 rows + ``coverage``: positions of each df1 interval covered by df2 (range_op.py:343-415).
        Inputs are swapped on the way down, as in the reference (:407-409)."""
        _validate_overlap_input(cols1, cols2, on_cols, suffixes, output_type)
        filter_op = _filter_op(df1, df2)
        cols1 = DEFAULT_INTERVAL_COLUMNS if cols1 is None else list(cols1)
        cols2 = DEFAULT_INTERVAL_COLUMNS if cols2 is None else list(cols2)
        # after the swap columns_1 must describe the engine's `left` (= df2) and columns_2 its `right` (= df1)
        opts = RangeOptions(range_op=RangeOp.Coverage, filter_op=filter_op, suffixes=tuple(suffixes), columns_1=cols2,
                            columns_2=cols1)
        return range_operation(df2, df1, opts, output_type, ctx, read_options, projection_pushdown=projection_pushdown)

    @staticmethod
    def count_overlaps(df1, df2, suffixes: tuple = ("", "_"), cols1=["chrom", "start", "end"],
                       cols2=["chrom", "start", "end"], on_cols=None, output_type: str = DEFAULT_OUTPUT_TYPE,
                       naive_query: bool = True, projection_pushdown: bool = True):
        """df1 rows + ``count``: number of df2 intervals overlapping each df1 interval (range_op.py:418-597).

        Both of the reference's algorithms (``naive_query=True``: CountOverlapsProvider; ``False``: the
        sweep-line window query, :512-597) produce the same frame (tests/test_pandas.py:73-105); here both
        map onto the same kernel, which evaluates the sweep-line identity ``starts_rank - ends_rank``."""
        _validate_overlap_input(cols1, cols2, on_cols, suffixes, output_type)
        filter_op = _filter_op(df1, df2)
        cols1 = DEFAULT_INTERVAL_COLUMNS if cols1 is None else list(cols1)
        cols2 = DEFAULT_INTERVAL_COLUMNS if cols2 is None else list(cols2)
        opts = RangeOptions(range_op=RangeOp.CountOverlapsNaive, filter_op=filter_op, suffixes=tuple(suffixes),
                            columns_1=cols2, columns_2=cols1)
        return range_operation(df2, df1, opts, output_type, ctx)

    # ---- unary sweeps (range_op.py:599-868) -------------------------------------------------
    @staticmethod
    def merge(df, min_dist: int = 0, cols=["chrom", "start", "end"], on_cols=None, output_type: str = DEFAULT_OUTPUT_TYPE,
              projection_pushdown: bool = True):
        """Merge overlapping intervals (range_op.py:600-655): (contig, start, end) as Int64 + ``n_intervals``.
        0-based: adjacent intervals stay apart; 1-based: intervals sharing an end point merge
        (tests/test_coordinate_system_metadata.py:1032-1054)."""
        _validate_overlap_input(cols, cols, on_cols, ("_1", "_2"), output_type)
        filter_op = _filter_op_single(df)
        cols = DEFAULT_INTERVAL_COLUMNS if cols is None else list(cols)
        opts = RangeOptions(range_op=RangeOp.Merge, filter_op=filter_op, columns_1=cols, columns_2=cols, min_dist=min_dist)
        return unary_operation(df, df, opts, output_type, ctx)

    @staticmethod
    def cluster(df, min_dist: int = 0, cols=["chrom", "start", "end"], output_type: str = DEFAULT_OUTPUT_TYPE,
                projection_pushdown: bool = True):
        """Every input row + ``cluster`` (id), ``cluster_start``, ``cluster_end`` of the merged interval it belongs to
        (range_op.py:657-712)."""
        _validate_overlap_input(cols, cols, None, ("_1", "_2"), output_type)
        filter_op = _filter_op_single(df)
        cols = DEFAULT_INTERVAL_COLUMNS if cols is None else list(cols)
        opts = RangeOptions(range_op=RangeOp.Cluster, filter_op=filter_op, columns_1=cols, columns_2=cols, min_dist=min_dist)
        return unary_operation(df, df, opts, output_type, ctx)

    @staticmethod
    def complement(df, view_df=None, cols=["chrom", "start", "end"], view_cols=None, output_type: str = DEFAULT_OUTPUT_TYPE,
                   projection_pushdown: bool = True):
        """The gaps between the intervals, inside ``view_df``'s regions when given, else inside [0, i64::MAX) of every
        contig present (range_op.py:714-789)."""
        _validate_overlap_input(cols, cols, None, ("_1", "_2"), output_type)
        filter_op = _filter_op_single(df)
        cols = DEFAULT_INTERVAL_COLUMNS if cols is None else list(cols)
        view_cols = cols if view_cols is None else list(view_cols)
        if view_df is None:
            logger.warning("No view_df provided — complement will span [0, i64::MAX) per contig. "
                           "Pass a view_df with contig boundaries (e.g., chromosome sizes) for meaningful results.")
        opts = RangeOptions(range_op=RangeOp.Complement, filter_op=filter_op, columns_1=cols, columns_2=cols,
                            view_table=None if view_df is None else "_view", view_columns=view_cols)
        return unary_operation(df, df, opts, output_type, ctx, view_df=view_df)

    @staticmethod
    def subtract(df1, df2, cols1=["chrom", "start", "end"], cols2=["chrom", "start", "end"],
                 output_type: str = DEFAULT_OUTPUT_TYPE, projection_pushdown: bool = True):
        """df1's intervals with every part covered by df2 removed; df1's other columns are kept (range_op.py:791-868)."""
        _validate_overlap_input(cols1, cols2, None, ("_1", "_2"), output_type)
        filter_op = _filter_op(df1, df2)
        cols1 = DEFAULT_INTERVAL_COLUMNS if cols1 is None else list(cols1)
        cols2 = DEFAULT_INTERVAL_COLUMNS if cols2 is None else list(cols2)
        opts = RangeOptions(range_op=RangeOp.Subtract, filter_op=filter_op, columns_1=cols1, columns_2=cols2)
        return unary_operation(df1, df2, opts, output_type, ctx)


overlap = IntervalOperations.overlap
nearest = IntervalOperations.nearest
coverage = IntervalOperations.coverage
count_overlaps = IntervalOperations.count_overlaps
merge = IntervalOperations.merge
cluster = IntervalOperations.cluster
complement = IntervalOperations.complement
subtract = IntervalOperations.subtract
