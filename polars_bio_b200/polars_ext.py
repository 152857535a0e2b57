"""``LazyFrame.pb`` namespace -- thin aliases onto the façade, as in
/root/reference/polars_bio/polars_ext.py:9-97.  Registered only when polars is importable."""
from __future__ import annotations

from .range_op import IntervalOperations

try:
    import polars as pl
except ImportError:  # polars is optional in this environment
    pl = None


class PolarsRangesOperations:
    def __init__(self, ldf):
        self._ldf = ldf

    def overlap(self, other_df, suffixes=("_1", "_2"), cols1=["chrom", "start", "end"], cols2=["chrom", "start", "end"],
                algorithm="Coitrees", low_memory=False, overlap_output="join", distinct_output=False,
                projection_pushdown=True):
        return IntervalOperations.overlap(self._ldf, other_df, suffixes=suffixes, cols1=cols1, cols2=cols2,
                                          algorithm=algorithm, low_memory=low_memory, overlap_output=overlap_output,
                                          distinct_output=distinct_output, projection_pushdown=projection_pushdown)

    def nearest(self, other_df, suffixes=("_1", "_2"), cols1=["chrom", "start", "end"], cols2=["chrom", "start", "end"],
                k=1, overlap=True, distance=True, projection_pushdown=True):
        return IntervalOperations.nearest(self._ldf, other_df, suffixes=suffixes, cols1=cols1, cols2=cols2, k=k,
                                          overlap=overlap, distance=distance, projection_pushdown=projection_pushdown)

    def count_overlaps(self, other_df, suffixes=("", "_"), cols1=["chrom", "start", "end"],
                       cols2=["chrom", "start", "end"], on_cols=None, naive_query=True, projection_pushdown=True):
        return IntervalOperations.count_overlaps(self._ldf, other_df, suffixes=suffixes, cols1=cols1, cols2=cols2,
                                                 on_cols=on_cols, naive_query=naive_query,
                                                 projection_pushdown=projection_pushdown)

    def coverage(self, other_df, suffixes=("_1", "_2"), cols1=["chrom", "start", "end"],
                 cols2=["chrom", "start", "end"], projection_pushdown=True):
        return IntervalOperations.coverage(self._ldf, other_df, suffixes=suffixes, cols1=cols1, cols2=cols2,
                                           projection_pushdown=projection_pushdown)

    def merge(self, min_dist=0, cols=["chrom", "start", "end"], projection_pushdown=True):
        return IntervalOperations.merge(self._ldf, min_dist=min_dist, cols=cols, projection_pushdown=projection_pushdown)

    def cluster(self, min_dist=0, cols=["chrom", "start", "end"], projection_pushdown=True):
        return IntervalOperations.cluster(self._ldf, min_dist=min_dist, cols=cols, projection_pushdown=projection_pushdown)

    def complement(self, view_df=None, cols=["chrom", "start", "end"], view_cols=None, projection_pushdown=True):
        return IntervalOperations.complement(self._ldf, view_df=view_df, cols=cols, view_cols=view_cols,
                                             projection_pushdown=projection_pushdown)

    def subtract(self, other_df, cols1=["chrom", "start", "end"], cols2=["chrom", "start", "end"], projection_pushdown=True):
        return IntervalOperations.subtract(self._ldf, other_df, cols1=cols1, cols2=cols2,
                                           projection_pushdown=projection_pushdown)


if pl is not None:  # pragma: no cover - exercised only where polars exists
    pl.api.register_lazyframe_namespace("pb")(PolarsRangesOperations)
    pl.api.register_dataframe_namespace("pb")(PolarsRangesOperations)
