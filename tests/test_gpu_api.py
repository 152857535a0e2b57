"""API-level parity: the reference-facing Python surface (pb.overlap / nearest / count_overlaps / coverage)
through the Arrow-level C ABI (pbgpu_range_op) with host frames, modelled on the reference's own tests
(tests/test_native.py, test_pandas.py, test_overlap_output_mode.py, test_suffix_handling.py,
test_wide_dataframes.py, test_coordinate_system_metadata.py).  Expected frames are the committed goldens or
the CPU oracle's answer on the same inputs; order is normalised by sorting on every column, as the reference does."""
import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

import oracle  # noqa: E402
import polars_bio_b200 as pb  # noqa: E402
from tests._golden import exons_fbrain_frames, fixtures, sort_all  # noqa: E402

FX = fixtures()
COLS = ("contig", "pos_start", "pos_end")


def _pdf(d, zero_based):
    df = pd.DataFrame(d)
    df.attrs["coordinate_system_zero_based"] = zero_based
    return df


def test_overlap_golden_frame():
    fx = FX["overlap"]  # reference tests/test_native.py:32-52, tests/test_pandas.py:33-50
    res = pb.overlap(_pdf(fx["df1"], False), _pdf(fx["df2"], False), cols1=COLS, cols2=COLS, output_type="pandas.DataFrame")
    assert len(res) == 16
    assert list(res.columns) == ["contig_1", "pos_start_1", "pos_end_1", "contig_2", "pos_start_2", "pos_end_2"]
    pd.testing.assert_frame_equal(sort_all(res), sort_all(pd.DataFrame(fx["expected"])), check_dtype=True)
    assert res.attrs["coordinate_system_zero_based"] is False


def test_nearest_golden_frame():
    fx = FX["nearest"]  # tests/test_native.py:55-75
    res = pb.nearest(_pdf(fx["df1"], False), _pdf(fx["df2"], False), cols1=COLS, cols2=COLS, output_type="pandas.DataFrame")
    exp = pd.DataFrame(fx["expected"])
    assert len(res) == len(exp)
    pd.testing.assert_frame_equal(sort_all(res), sort_all(exp), check_dtype=True)


def test_nearest_variants():
    fx = FX["nearest"]  # tests/test_native.py:78-179
    d1, d2 = _pdf(fx["df1"], False), _pdf(fx["df2"], False)
    k2 = pb.nearest(d1, d2, cols1=COLS, cols2=COLS, k=2, output_type="pandas.DataFrame")
    assert len(k2) >= 11 and k2.groupby(["contig_1", "pos_start_1", "pos_end_1"]).size().max() <= 2
    assert set(k2.columns) == {"contig_1", "pos_start_1", "pos_end_1", "contig_2", "pos_start_2", "pos_end_2", "distance"}
    no = pb.nearest(d1, d2, cols1=COLS, cols2=COLS, overlap=False, output_type="pandas.DataFrame")
    valid = no.dropna(subset=["distance"])
    assert len(valid) > 0 and (valid["distance"] > 0).all()
    nd = pb.nearest(d1, d2, cols1=COLS, cols2=COLS, distance=False, output_type="pandas.DataFrame")
    assert "distance" not in nd.columns and len(nd) == 11


def test_count_overlaps_golden_frame_both_algorithms():
    fx = FX["count_overlaps"]  # tests/test_native.py:182-203, tests/test_pandas.py:73-105
    exp = sort_all(pd.DataFrame(fx["expected"]))
    for naive in (True, False):
        res = pb.count_overlaps(_pdf(fx["df1"], False), _pdf(fx["df2"], False), cols1=COLS, cols2=COLS,
                                output_type="pandas.DataFrame", naive_query=naive)
        assert list(res.columns) == ["contig", "pos_start", "pos_end", "count"]
        pd.testing.assert_frame_equal(sort_all(res), exp, check_dtype=True)


@pytest.mark.parametrize("kat", FX["overlap_kats"])
def test_boundary_kats_api(kat):
    a = _pdf({"chrom": ["chr1"], "start": [kat["a"][0]], "end": [kat["a"][1]]}, kat["zero_based"])
    b = _pdf({"chrom": ["chr1"], "start": [kat["b"][0]], "end": [kat["b"][1]]}, kat["zero_based"])
    assert len(pb.overlap(a, b, output_type="pandas.DataFrame")) == kat["rows"]
    assert pb.count_overlaps(a, b, output_type="pandas.DataFrame")["count"].tolist() == [kat["rows"]]


@pytest.mark.parametrize("kat", FX["coverage_kats"])
def test_coverage_kats_api_uint32(kat):
    # tests/test_coordinate_system_metadata.py:1577-1623 (UInt32 positions)
    a = pa.table({"chrom": ["chr1"], "start": pa.array([kat["a"][0]], pa.uint32()), "end": pa.array([kat["a"][1]], pa.uint32())})
    b = pa.table({"chrom": ["chr1"], "start": pa.array([kat["b"][0]], pa.uint32()), "end": pa.array([kat["b"][1]], pa.uint32())})
    a = pb.set_coordinate_system(a, kat["zero_based"]); b = pb.set_coordinate_system(b, kat["zero_based"])
    res = pb.coverage(a, b, output_type="pyarrow.Table")
    assert res.num_rows == 1 and res["coverage"].to_pylist() == [kat["coverage"]]
    assert res.schema.field("start").type == pa.uint32()  # coordinate dtypes pass through unchanged


def test_output_modes():
    fx = FX["output_mode"]  # tests/test_overlap_output_mode.py:99-152
    left, right = _pdf(fx["left"], True), _pdf(fx["right"], True)
    by = ["chrom", "start", "end", "name"]
    mult = pb.overlap(left, right, overlap_output="left", output_type="pandas.DataFrame")
    assert list(mult.columns) == by
    pd.testing.assert_frame_equal(mult.sort_values(by).reset_index(drop=True),
                                  pd.DataFrame(fx["expected_left_multiplicity"]).sort_values(by).reset_index(drop=True))
    dist = pb.overlap(left, right, overlap_output="left", distinct_output=True, output_type="pandas.DataFrame")
    pd.testing.assert_frame_equal(dist.sort_values(by).reset_index(drop=True),
                                  pd.DataFrame(fx["expected_left_distinct"]).sort_values(by).reset_index(drop=True))
    join = pb.overlap(left, right, output_type="pandas.DataFrame")
    assert {"chrom_1", "chrom_2", "score_2", "name_1"} <= set(join.columns)
    with pytest.raises(ValueError, match="overlap_output"):
        pb.overlap(left, right, overlap_output="semi", output_type="pandas.DataFrame")


def test_suffixes_and_wide_payload():
    # every column gets the suffix; payload columns of both sides survive with dtypes (tests/test_wide_dataframes.py)
    rng = np.random.default_rng(0)
    n, m = 400, 300
    d1 = pd.DataFrame({"chrom": rng.choice(["chr1", "chr2", "chrX"], n), "start": rng.integers(0, 5000, n)})
    d1["end"] = d1["start"] + rng.integers(1, 200, n)
    d1["name"] = [f"r{i}" for i in range(n)]
    d1["score"] = rng.random(n)
    d1["flag"] = rng.random(n) < 0.5
    d1["maybe"] = pd.array([None if i % 7 == 0 else i for i in range(n)], dtype="Int64")
    d2 = pd.DataFrame({"chrom": rng.choice(["chr1", "chr2", "chr3"], m), "start": rng.integers(0, 5000, m)})
    d2["end"] = d2["start"] + rng.integers(1, 400, m)
    d2["gene"] = [None if i % 11 == 0 else f"g{i % 13}" for i in range(m)]
    d2["weight"] = rng.integers(0, 100, m).astype(np.int16)
    d1.attrs["coordinate_system_zero_based"] = True; d2.attrs["coordinate_system_zero_based"] = True
    res = pb.overlap(d1, d2, suffixes=("_a", "_b"), output_type="pandas.DataFrame")
    assert list(res.columns) == [c + "_a" for c in d1.columns] + [c + "_b" for c in d2.columns]
    c1, c2, names = oracle.encode_contigs(d1["chrom"], d2["chrom"])
    a, b = oracle.Index(c2, d2["start"].to_numpy(np.int32), d2["end"].to_numpy(np.int32), len(names)).overlap_pairs(
        c1, d1["start"].to_numpy(np.int32), d1["end"].to_numpy(np.int32), True)
    exp = pd.concat([d1.iloc[a].reset_index(drop=True).add_suffix("_a"), d2.iloc[b].reset_index(drop=True).add_suffix("_b")], axis=1)
    assert len(res) == len(exp) > 0
    key = ["name_a", "start_b", "end_b", "weight_b"]
    r, e = res.sort_values(key).reset_index(drop=True), exp.sort_values(key).reset_index(drop=True)
    for col in res.columns:
        assert r[col].isna().tolist() == e[col].isna().tolist(), col
        assert r[col].dropna().tolist() == e[col].dropna().tolist(), col
    assert res["weight_b"].dtype == np.int16 and res["score_a"].dtype == np.float64 and res["flag_a"].dtype == bool


def test_position_dtypes_and_null_keys():
    # Int64 / UInt64 / mixed position dtypes are accepted (test_coordinate_system_metadata.py:1436-1575);
    # null contig / position rows never match and count 0
    a = pa.table({"chrom": pa.array(["chr1", "chr1", None, "chr1"]), "start": pa.array([100, 200, 100, None], pa.int64()),
                  "end": pa.array([150, 250, 150, 300], pa.uint64())})
    b = pa.table({"chrom": pa.array(["chr1"], pa.large_string()), "start": pa.array([125], pa.uint64()), "end": pa.array([175], pa.int32())})
    a = pb.set_coordinate_system(a, True); b = pb.set_coordinate_system(b, True)
    cnt = pb.count_overlaps(a, b, output_type="pyarrow.Table")
    assert cnt.num_rows == 4 and cnt["count"].to_pylist() == [1, 0, 0, 0]
    assert pb.overlap(a, b, output_type="pyarrow.Table").num_rows == 1
    big = pa.table({"chrom": ["chr1"], "start": pa.array([2**31], pa.int64()), "end": pa.array([2**31 + 5], pa.int64())})
    with pytest.raises(pb._native.PbgpuError, match="int32"):
        pb.overlap(pb.set_coordinate_system(big, True), b, output_type="pyarrow.Table")


def test_dictionary_and_chunked_inputs():
    fx = FX["overlap"]
    d1, d2 = pd.DataFrame(fx["df1"]), pd.DataFrame(fx["df2"])
    d1["contig"] = d1["contig"].astype("category")  # -> Arrow dictionary
    t1 = pa.Table.from_pandas(d1, preserve_index=False)
    t2 = pa.concat_tables([pa.Table.from_pandas(d2.iloc[:3], preserve_index=False),
                           pa.Table.from_pandas(d2.iloc[3:], preserve_index=False)])  # 2 chunks
    assert t2.column(0).num_chunks == 2
    t1 = pb.set_coordinate_system(t1, False); t2 = pb.set_coordinate_system(t2, False)
    res = pb.overlap(t1, t2, cols1=COLS, cols2=COLS, output_type="pandas.DataFrame")
    exp = pd.DataFrame(fx["expected"])
    res["contig_1"] = res["contig_1"].astype(str)
    pd.testing.assert_frame_equal(sort_all(res), sort_all(exp), check_dtype=False)


def test_exons_fbrain_api_both_directions():
    # tests/test_bioframe.py:128-143 shape: 0-based, full frames; expected from the oracle (bioframe absent), 54,246 rows
    ex, fb = exons_fbrain_frames()
    ex.attrs["coordinate_system_zero_based"] = True; fb.attrs["coordinate_system_zero_based"] = True
    res = pb.overlap(ex, fb, cols1=COLS, cols2=COLS, output_type="pandas.DataFrame")
    assert len(res) == FX["exons_fbrain_pairs"]["strict"]
    assert res["pos_start_1"].dtype == np.int32  # Int32 in -> Int32 out (docs/supplement.md:328-333)
    c1, c2, names = oracle.encode_contigs(ex["contig"], fb["contig"])
    a, b = oracle.Index(c2, fb["pos_start"], fb["pos_end"], len(names)).overlap_pairs(c1, ex["pos_start"], ex["pos_end"], True)
    exp = pd.concat([ex.iloc[a].reset_index(drop=True).add_suffix("_1"), fb.iloc[b].reset_index(drop=True).add_suffix("_2")], axis=1)
    pd.testing.assert_frame_equal(sort_all(res), sort_all(exp), check_dtype=True)
    cnt = pb.count_overlaps(ex, fb, cols1=COLS, cols2=COLS, output_type="pandas.DataFrame")
    assert int(cnt["count"].sum()) == 54246 and len(cnt) == len(ex)
    near = pb.nearest(ex, fb, cols1=COLS, cols2=COLS, output_type="pandas.DataFrame")
    assert len(near) == len(ex)
    ob, od = oracle.Index(c2, fb["pos_start"], fb["pos_end"], len(names)).nearest(c1, ex["pos_start"], ex["pos_end"], True)
    assert np.array_equal(near["distance"].to_numpy(), od[:, 0])  # rows come back in df1 order
    cov = pb.coverage(ex, fb, cols1=COLS, cols2=COLS, output_type="pandas.DataFrame")
    ocov = oracle.Index(c2, fb["pos_start"], fb["pos_end"], len(names)).coverage(c1, ex["pos_start"], ex["pos_end"], True)
    assert np.array_equal(cov["coverage"].to_numpy(), ocov)


def test_low_memory_batches_and_streaming_result():
    ex, fb = exons_fbrain_frames()
    ex.attrs["coordinate_system_zero_based"] = True; fb.attrs["coordinate_system_zero_based"] = True
    from polars_bio_b200 import RangeOp, RangeOptions, FilterOp, range_operation_frame

    opts = RangeOptions(range_op=RangeOp.Overlap, filter_op=FilterOp.Strict, columns_1=list(COLS), columns_2=list(COLS),
                        overlap_low_memory=True)
    res = range_operation_frame(pb.ctx, ex, fb, opts)
    assert res.schema().names[0] == "contig_1"
    sizes = [b.num_rows for b in res.select(["contig_1", "pos_start_2"]).execute_stream()]
    assert sum(sizes) == 54246 and max(sizes) <= 8192 and len(sizes) >= 6
    lim = range_operation_frame(pb.ctx, ex, fb, opts, limit=100)
    assert lim.count() == 100
    pairs = range_operation_frame(pb.ctx, ex, fb, opts, emit=1).to_arrow()
    assert pairs.schema.names == ["left_row", "right_row"] and pairs.num_rows == 54246


def test_errors_do_not_abort():
    fx = FX["overlap"]
    with pytest.raises(pb._native.PbgpuError, match="not found"):
        pb.overlap(_pdf(fx["df1"], False), _pdf(fx["df2"], False), output_type="pandas.DataFrame")  # default cols absent
    bad = pa.table({"chrom": [1, 2], "start": [1, 2], "end": [3, 4]})
    with pytest.raises(pb._native.PbgpuError, match="unsupported type"):
        pb.overlap(pb.set_coordinate_system(bad, True), pb.set_coordinate_system(bad, True), output_type="pyarrow.Table")


def test_string_view_and_binary_payload_and_nearest_pairs():
    # utf8_view contig + payload come back as large_utf8; emit=1 for nearest yields (left_row, right_row?, distance?)
    # (string_view arrays are built directly: pyarrow 24 segfaults exporting the result of .cast(pa.string_view()))
    left = pa.table({"chrom": pa.array(["chr1", "chr1", "chr2", "chr9"], type=pa.string_view()),
                     "start": pa.array([100, 500, 100, 5], pa.int32()), "end": pa.array([200, 600, 200, 9], pa.int32()),
                     "tag": pa.array(["a-long-tag-beyond-12-bytes", "b", None, "d"], type=pa.string_view()),
                     "blob": pa.array([b"\x00\x01", b"", b"xyz", None], pa.binary())})
    right = pa.table({"chrom": ["chr1", "chr1", "chr2"], "start": pa.array([150, 900, 50], pa.int64()),
                      "end": pa.array([160, 950, 120], pa.int64()), "w": pa.array([1.5, 2.5, None], pa.float32())})
    left = pb.set_coordinate_system(left, True); right = pb.set_coordinate_system(right, True)
    res = pb.overlap(left, right, output_type="pyarrow.Table")
    assert res.schema.field("chrom_1").type == pa.large_string() and res.schema.field("tag_1").type == pa.large_string()
    assert res.schema.field("blob_1").type == pa.binary() and res.schema.field("start_2").type == pa.int64()
    got = sorted(zip(res["chrom_1"].to_pylist(), res["start_1"].to_pylist(), res["tag_1"].to_pylist(), res["blob_1"].to_pylist(),
                     res["start_2"].to_pylist(), res["w_2"].to_pylist()), key=lambda r: (r[0], r[1]))
    assert got == [("chr1", 100, "a-long-tag-beyond-12-bytes", b"\x00\x01", 150, 1.5), ("chr2", 100, None, b"xyz", 50, None)]
    near = pb.nearest(left, right, output_type="pyarrow.Table")
    assert near.num_rows == 4
    d = dict(zip(near["start_1"].to_pylist(), near["distance"].to_pylist()))
    assert d == {100: 0, 500: 300, 5: None}  # chr1/chr2 probes at start 100 both overlap; chr9 has no indexed rows -> null
    from polars_bio_b200 import FilterOp, RangeOp, RangeOptions, range_operation_frame

    ro = RangeOptions(range_op=RangeOp.Nearest, filter_op=FilterOp.Strict, nearest_k=2)
    pairs = range_operation_frame(pb.ctx, left, right, ro, emit=1).to_arrow()
    assert pairs.schema.names == ["left_row", "right_row", "distance"]
    assert pairs["left_row"].to_pylist() == [0, 0, 1, 1, 2, 3]
    assert pairs["right_row"].to_pylist() == [0, 1, 1, 0, 2, None]     # row 1: downstream variant (300) before upstream (340)
    assert pairs["distance"].to_pylist() == [0, 700, 300, 340, 0, None]


def test_count_overlaps_passthrough_multibatch_low_memory():
    # the iterated table's columns are re-exported zero-copy: input batch boundaries and sliced views must line up
    fx = FX["count_overlaps"]
    d1 = pd.DataFrame(fx["df1"]); d1["payload"] = [f"row{i}" for i in range(len(d1))]
    t1 = pa.concat_tables([pa.Table.from_pandas(d1.iloc[:4], preserve_index=False), pa.Table.from_pandas(d1.iloc[4:9], preserve_index=False),
                           pa.Table.from_pandas(d1.iloc[9:], preserve_index=False)])
    t2 = pa.Table.from_pandas(pd.DataFrame(fx["df2"]), preserve_index=False)
    t1 = pb.set_coordinate_system(t1, False); t2 = pb.set_coordinate_system(t2, False)
    from polars_bio_b200 import FilterOp, RangeOp, RangeOptions, range_operation_frame

    ro = RangeOptions(range_op=RangeOp.CountOverlapsNaive, filter_op=FilterOp.Weak, columns_1=list(COLS), columns_2=list(COLS))
    pb.set_option("datafusion.execution.batch_size", 3)
    try:
        ro.overlap_low_memory = True
        batches = list(range_operation_frame(pb.ctx, t2, t1, ro).execute_stream())   # engine order: (indexed, iterated)
    finally:
        pb.set_option("datafusion.execution.batch_size", 8192)
    assert [b.num_rows for b in batches] == [4, 5, 2]  # cap 3 is rounded up to 8 rows: slices follow the input batches
    got = pa.Table.from_batches(batches).to_pandas()
    assert got["payload"].tolist() == d1["payload"].tolist()
    assert got["count"].tolist() == fx["expected"]["count"]
    lim = range_operation_frame(pb.ctx, t2, t1, ro, limit=6).to_arrow()
    assert lim.num_rows == 6 and lim["payload"].to_pylist() == d1["payload"].tolist()[:6]


@pytest.mark.parametrize("cap", ["1", "1000", "20000"])
def test_streaming_sink_matches_single_shot(cap, monkeypatch):
    # SURVEY.md 7 step 6 / BASELINE config 5: results larger than one ring slot are emitted chunk by chunk from
    # get_next with the device state kept alive; rows, order, limit and every output flavour must not change.
    ex, fb = exons_fbrain_frames()
    ex.attrs["coordinate_system_zero_based"] = True; fb.attrs["coordinate_system_zero_based"] = True
    ex = ex.assign(tag=np.arange(len(ex)).astype(str))  # a payload column: row ids must travel too
    from polars_bio_b200 import RangeOp, RangeOptions, FilterOp, OverlapOutputMode, range_operation_frame

    def run(**kw):
        opts = RangeOptions(range_op=RangeOp.Overlap, filter_op=FilterOp.Strict, columns_1=list(COLS), columns_2=list(COLS), **kw)
        return opts

    monkeypatch.delenv("PBGPU_SINK_PAIRS", raising=False)
    whole = range_operation_frame(pb.ctx, ex, fb, run()).to_arrow()
    whole_pairs = range_operation_frame(pb.ctx, ex, fb, run(), emit=1).to_arrow()
    whole_left = range_operation_frame(pb.ctx, ex, fb, run(overlap_output=OverlapOutputMode.Left)).to_arrow()
    monkeypatch.setenv("PBGPU_SINK_PAIRS", cap)
    chunked = range_operation_frame(pb.ctx, ex, fb, run()).to_arrow()
    assert chunked.num_rows == 54246 and chunked.equals(whole)  # same rows in the same order
    assert range_operation_frame(pb.ctx, ex, fb, run(), emit=1).to_arrow().equals(whole_pairs)
    assert range_operation_frame(pb.ctx, ex, fb, run(overlap_output=OverlapOutputMode.Left)).to_arrow().equals(whole_left)
    lim = range_operation_frame(pb.ctx, ex, fb, run(), limit=12345).to_arrow()
    assert lim.equals(whole.slice(0, 12345))
    # dropping a half-consumed stream must release the device state cleanly
    it = range_operation_frame(pb.ctx, ex, fb, run(overlap_low_memory=True)).execute_stream()
    first = next(iter(it))
    assert first.num_rows > 0
    del it, first
    res = pb.overlap(ex, fb, cols1=COLS, cols2=COLS, output_type="pandas.DataFrame")
    assert len(res) == 54246
