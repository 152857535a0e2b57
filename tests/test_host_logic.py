"""Host-side logic that needs no GPU: option mirrors, coordinate-system metadata rules, argument validation,
input adapters -- the behaviours pinned by the reference's tests/test_coordinate_system_metadata.py,
test_warnings.py, test_context_options.py and range_op_helpers.py:379-399."""
import warnings

import pandas as pd
import pyarrow as pa
import pytest

import polars_bio_b200 as pb
from polars_bio_b200 import _metadata, range_op_io
from polars_bio_b200.constants import POLARS_BIO_COORDINATE_SYSTEM_CHECK, POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED


def _df(zero_based=None):
    df = pd.DataFrame({"chrom": ["chr1"], "start": [100], "end": [200]})
    if zero_based is not None:
        df.attrs["coordinate_system_zero_based"] = zero_based
    return df


def test_enum_values_match_option_rs():
    assert int(pb.FilterOp.Weak) == 0 and int(pb.FilterOp.Strict) == 1          # option.rs:96-99
    assert int(pb.RangeOp.Overlap) == 0 and int(pb.RangeOp.Nearest) == 3 and int(pb.RangeOp.Coverage) == 4
    assert int(pb.RangeOp.CountOverlapsNaive) == 6 and int(pb.RangeOp.Merge) == 7  # option.rs:103-112
    assert int(pb.OverlapOutputMode.Join) == 0 and int(pb.OverlapOutputMode.Left) == 1
    ro = pb.RangeOptions(range_op=pb.RangeOp.Overlap)
    assert ro.filter_op is None and ro.nearest_k is None and ro.distinct_output is None


def test_defaults_match_reference_context():
    assert pb.get_option(POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED) == "false"   # context.py:45
    assert pb.get_option(POLARS_BIO_COORDINATE_SYSTEM_CHECK) == "false"        # context.py:48
    assert pb.get_option("datafusion.execution.target_partitions") == "1"      # context.py:36
    pb.set_option("datafusion.execution.target_partitions", 4)                  # numeric coercion (test_context_options.py)
    assert pb.get_option("datafusion.execution.target_partitions") == "4"
    pb.set_option("datafusion.execution.target_partitions", 1)
    pb.set_option("some.unknown.key", True)                                     # unknown keys are swallowed (context.rs:94-97)
    assert pb.get_option("some.unknown.key") == "true"


def test_metadata_mismatch_and_missing():
    with pytest.raises(pb.CoordinateSystemMismatchError):
        _metadata.validate_coordinate_systems(_df(True), _df(False))
    assert _metadata.validate_coordinate_systems(_df(True), _df(True)) is True
    assert _metadata.validate_coordinate_systems(_df(False), _df(False)) is False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert _metadata.validate_coordinate_systems(_df(), _df(False)) is False  # lenient: global default (1-based)
        assert any("metadata is missing" in str(x.message) for x in w)
    pb.set_option(POLARS_BIO_COORDINATE_SYSTEM_CHECK, True)
    try:
        with pytest.raises(pb.MissingCoordinateSystemError):
            _metadata.validate_coordinate_systems(_df(), _df(True))
    finally:
        pb.set_option(POLARS_BIO_COORDINATE_SYSTEM_CHECK, False)
    pb.set_option(POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED, True)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert _metadata.validate_coordinate_systems(_df(), _df()) is True
    finally:
        pb.set_option(POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED, False)


def test_arrow_table_metadata_roundtrip():
    t = pa.table({"chrom": ["chr1"], "start": [1], "end": [2]})
    assert pb.get_coordinate_system(t) is None
    assert pb.get_coordinate_system(pb.set_coordinate_system(t, True)) is True
    assert pb.get_coordinate_system(pb.set_coordinate_system(t, False)) is False


def test_argument_validation():
    with pytest.raises(AssertionError, match="on_cols"):
        pb.overlap(_df(True), _df(True), on_cols=["x"], output_type="pandas.DataFrame")
    with pytest.raises(AssertionError):
        pb.overlap(_df(True), _df(True), output_type="numpy")
    with pytest.raises(ValueError, match="overlap_output"):
        pb.overlap(_df(True), _df(True), overlap_output="semi", output_type="pandas.DataFrame")


def test_c_options_mapping():
    ro = pb.RangeOptions(range_op=pb.RangeOp.Overlap, filter_op=pb.FilterOp.Strict, suffixes=("_x", "_y"),
                         columns_1=["a", "b", "c"], columns_2=["d", "e", "f"], overlap_output=pb.OverlapOutputMode.Left,
                         distinct_output=True, overlap_low_memory=True)
    o = range_op_io._c_opts(ro, 0, 7, pb.ctx)
    assert (o.range_op, o.filter_op, o.output_mode, o.limit) == (0, 1, 2, 7)   # (Left, True) -> LeftDistinct
    assert o.cols1[0] == b"a" and o.cols2[2] == b"f" and o.suffixes[1] == b"_y"
    assert o.max_batch_rows == 8192                                             # low_memory caps output batches
    ro2 = pb.RangeOptions(range_op=pb.RangeOp.Nearest, filter_op=pb.FilterOp.Weak)
    o2 = range_op_io._c_opts(ro2, 1, None, pb.ctx)
    assert (o2.nearest_k, o2.include_overlaps, o2.compute_distance, o2.emit) == (1, 1, 1, 1)  # operation.rs:111-113 defaults
    assert o2.max_batch_rows == 1 << 20


def test_input_adapters(tmp_path):
    df = _df(True)
    r = range_op_io._df_to_reader(df)
    assert r.schema.names == ["chrom", "start", "end"]
    p = tmp_path / "x.csv"
    df.to_csv(p, index=False)
    assert range_op_io._df_to_reader(str(p)).read_all().num_rows == 1
    import pyarrow.parquet as pq

    q = tmp_path / "x.parquet"
    pq.write_table(pa.Table.from_pandas(df), q)
    assert range_op_io._df_to_reader(str(q)).read_all().num_rows == 1
    with pytest.raises(TypeError):
        range_op_io._df_to_reader(42)


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb._native.PbgpuError):  # no silent CPU fallback
        pb.overlap(_df(True), _df(True), output_type="pandas.DataFrame")


def test_bench_roofline_names_the_longest_single_launch():
    """bench.stage_roofline: with partitioned probes count_overlaps / pass 1 are four launches each, so the dominant KERNEL is
    pass 2 even when a multi-launch stage takes longer; without a partition every stage is one kernel and the longest wins."""
    import bench

    n, m, pairs = 100_000_000, 90_000_000, 438_606_732
    km = {"partition_sort_ns": 6.7, "count_ns": 2.5, "scan_ns": 0.02, "emit_ns": 2.3, "count_overlaps_ns": 2.9, "bin_ns": 1.4, "unbin_ns": 0.4}
    r = bench.stage_roofline(n, m, pairs, km, 6534.8, "measured")
    assert "emit" in r["kernel"] and abs(r["frac"] - (12.0 * (n + m) + 8.0 * pairs) / 2.3e-3 / 1e9 / 6534.8) < 1e-9
    assert set(r["all_stages"]) >= {"count_overlaps (all kernels of the call)", "overlap pass 2 (emit)"}
    km2 = dict(km, bin_ns=0.0, unbin_ns=0.0, count_overlaps_ns=0.066, count_ns=0.071, emit_ns=0.058)
    r2 = bench.stage_roofline(10_000_000, 1_000_000, 6_024_485, km2, 6534.8, "measured")
    assert "pass 1" in r2["kernel"]
