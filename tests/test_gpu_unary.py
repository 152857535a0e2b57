"""GPU parity of the unary sweeps (pbgpu_merge / pbgpu_cluster / pbgpu_subtract through polars_bio_b200.engine and the
public pb.merge / cluster / complement / subtract) against oracle/unary_np.py and the reference's golden vectors.
Bit-exact: integer work."""
import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import unary_np as U  # noqa: E402
from tests._golden import exons_fbrain, fixtures, sort_all  # noqa: E402

FX = fixtures()


def _eng():
    from polars_bio_b200 import engine

    return engine


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to("cuda:0")


def _np(t):
    return t.cpu().numpy()


def _check_all(lc, ls, le, rc, rs, re, nc, strict, min_dists=(0, 7), sub_oracle=U.subtract):
    eng = _eng()
    fo = eng.FILTER_STRICT if strict else eng.FILTER_WEAK
    L, R = [_dev(x) for x in (lc, ls, le)], [_dev(x) for x in (rc, rs, re)]
    for md in min_dists:
        mc, ms, me, mn = (_np(x) for x in eng.merge_intervals(*R, nc, fo, md))
        oc, os_, oe, on = U.merge(rc, rs, re, nc, strict, md)
        assert np.array_equal(mc, oc) and np.array_equal(ms, os_) and np.array_equal(me, oe) and np.array_equal(mn, on), ("merge", strict, md)
        cid, cs, ce, k = eng.cluster_intervals(*R, nc, fo, md)
        ocid, ocs, oce = U.cluster(rc, rs, re, nc, strict, md)
        assert k == len(oc)
        assert np.array_equal(_np(cid), ocid) and np.array_equal(_np(cs), ocs) and np.array_equal(_np(ce), oce), ("cluster", strict, md)
    row, fs, fe = (_np(x) for x in eng.subtract_intervals(*L, *R, nc, fo))
    orow, ofs, ofe = sub_oracle(lc, ls, le, rc, rs, re, nc, strict)
    assert np.array_equal(row.view(np.uint32), orow) and np.array_equal(fs, ofs) and np.array_equal(fe, ofe), ("subtract", strict)


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_small_ragged_with_degenerate_rows(strict, seed):
    rng = np.random.default_rng(seed)
    for _ in range(8):
        nc = int(rng.integers(1, 5))
        n, m = int(rng.integers(0, 300)), int(rng.integers(0, 300))
        lc = rng.integers(-1, nc + 1, n); ls = rng.integers(-50, 400, n); le = ls + rng.integers(-3, 60, n)
        rc = rng.integers(-1, nc + 1, m); rs = rng.integers(-50, 400, m); re = rs + rng.integers(-3, 40, m)
        _check_all(lc, ls, le, rc, rs, re, nc, strict)


@pytest.mark.parametrize("strict", [True, False])
def test_medium_nested_many_contigs(strict):
    rng = np.random.default_rng(5)
    nc, n, m = 24, 120_000, 200_000
    lc = rng.integers(0, nc, n); ls = rng.integers(0, 3_000_000, n); le = ls + rng.integers(1, 3000, n)
    rc = rng.integers(0, nc, m); rs = rng.integers(0, 3_000_000, m)
    re = rs + np.where(rng.random(m) < 0.02, rng.integers(1, 20_000, m), rng.integers(1, 40, m))  # a few long ones: nesting
    _check_all(lc, ls, le, rc, rs, re, nc, strict, min_dists=(0, 25), sub_oracle=U.subtract_ranks)


def test_sparse_runs_take_the_rank_directory_and_dense_clusters():
    # few, far-apart runs (directory mostly empty buckets) and one dense pile (a single run of 50k rows)
    rng = np.random.default_rng(9)
    rs = np.concatenate([rng.integers(0, 200_000_000, 3000), 77_000_000 + rng.integers(0, 5000, 50_000)])
    re = rs + rng.integers(1, 100, len(rs))
    rc = np.zeros(len(rs), np.int64)
    n = 60_000
    ls = np.concatenate([rng.integers(0, 200_000_000, n // 2), 77_000_000 + rng.integers(-3000, 9000, n // 2)])
    le = ls + rng.integers(1, 4000, n)
    lc = np.zeros(n, np.int64)
    _check_all(lc, ls, le, rc, rs, re, 1, True, min_dists=(0,), sub_oracle=U.subtract_ranks)


def test_span_beyond_32_bits_and_int32_extremes():
    # three contigs spanning ~2^31 each: no global axis, subtract falls back to the bounded searches
    rng = np.random.default_rng(3)
    m, n = 3000, 4000
    rc = rng.integers(0, 3, m); rs = rng.integers(-2**31, 2**31 - 5000, m); re = rs + rng.integers(1, 3000, m)
    lc = rng.integers(0, 3, n); ls = rng.integers(-2**31, 2**31 - 5000, n)
    le = ls + np.minimum(rng.integers(1, 4_000_000, n), 2**31 - 1 - ls)
    for strict in (True, False):
        _check_all(lc, ls, le, rc, rs, re, 3, strict, min_dists=(0, 1000))
    lo, hi = -2**31, 2**31 - 1
    rc = np.zeros(4, np.int64); rs = np.array([lo, hi - 10, 0, hi]); re = np.array([lo + 5, hi, 10, hi])
    lc = np.zeros(3, np.int64); ls = np.array([lo, hi - 20, -5]); le = np.array([hi, hi, 20])
    for strict in (True, False):
        _check_all(lc, ls, le, rc, rs, re, 1, strict, min_dists=(0, 2**31 - 1))


def test_empty_tables():
    e = np.zeros(0, np.int32)
    lc, ls, le = np.zeros(5, np.int32), np.arange(5) * 10, np.arange(5) * 10 + 5
    _check_all(lc, ls, le, e, e, e, 2, True)
    _check_all(e, e, e, lc, ls, le, 2, False)
    _check_all(e, e, e, e, e, e, 1, True)


def test_exons_fbrain_fixtures():
    z = exons_fbrain()
    ex = (z["exons_chrom"].astype(np.int32), z["exons_start"], z["exons_end"])
    fb = (z["fbrain_chrom"].astype(np.int32), z["fbrain_start"], z["fbrain_end"])
    _check_all(*ex, *fb, 24, True, min_dists=(0,), sub_oracle=U.subtract_ranks)
    eng = _eng()
    mc, ms, me, mn = (_np(x) for x in eng.merge_intervals(*[_dev(x) for x in ex], 24, eng.FILTER_STRICT))
    assert int(mn.sum()) == len(ex[0])  # every row lands in exactly one merged interval
    again = [_np(x) for x in eng.merge_intervals(_dev(mc), _dev(ms), _dev(me), 24, eng.FILTER_STRICT)]
    assert np.array_equal(again[1], ms) and np.array_equal(again[2], me) and (again[3] == 1).all()  # idempotent


def test_full_size_properties_merge_and_subtract():
    """10M reads (150 bp) on one contig: merge conserves rows, its output is sorted and disjoint; subtracting the
    merged intervals from the reads leaves nothing; subtracting 1M SNVs leaves pieces that avoid every SNV."""
    eng = _eng()
    from bench import make_config2

    (pc, ps, pe), (bc, bs, be), nc = make_config2(10_000_000, 1_000_000)
    dp, db = [_dev(x) for x in (pc, ps, pe)], [_dev(x) for x in (bc, bs, be)]
    mc, ms, me, mn = eng.merge_intervals(*dp, nc, eng.FILTER_STRICT)
    assert int(mn.sum()) == len(pc)
    assert bool((ms[1:] >= me[:-1]).all()) and bool((me > ms).all())
    row, fs, fe = eng.subtract_intervals(*dp, mc, ms, me, nc, eng.FILTER_STRICT)
    assert row.numel() == 0
    row, fs, fe = eng.subtract_intervals(*dp, *db, nc, eng.FILTER_STRICT)
    cnt = eng.DeviceIndex(*db, nc).count_overlaps(dp[0].new_zeros(fs.numel()), fs, fe, eng.FILTER_STRICT)
    assert int(cnt.sum()) == 0 and bool((fe > fs).all())
    covered = int((dp[2].long() - dp[1].long()).sum()) - int((fe.long() - fs.long()).sum())  # positions removed = SNVs hit, per read
    assert covered == int(eng.DeviceIndex(*db, nc).coverage(*dp, eng.FILTER_STRICT).sum())
    sl = slice(2_000_000, 2_050_000)
    orow, ofs, ofe = U.subtract_ranks(pc[sl], ps[sl], pe[sl], bc, bs, be, nc, True)
    sel = (row >= sl.start) & (row < sl.stop)
    assert np.array_equal(_np(row[sel]) - sl.start, orow) and np.array_equal(_np(fs[sel]), ofs) and np.array_equal(_np(fe[sel]), ofe)


# ---- public API on the device -----------------------------------------------------------------------------------
def _frame(d, zero_based=True):
    df = pd.DataFrame(d)
    df.attrs["coordinate_system_zero_based"] = zero_based
    return df


def test_api_goldens():
    import polars_bio_b200 as pb

    g = FX["merge"]
    out = pb.merge(_frame(g["df"], g["zero_based"]), cols=("contig", "pos_start", "pos_end"), output_type="pandas.DataFrame")
    want = pd.DataFrame(g["expected"]).astype({"pos_start": "int64", "pos_end": "int64", "n_intervals": "int64"})
    pd.testing.assert_frame_equal(sort_all(out), sort_all(want))  # tests/test_native.py:206-223
    for kat in FX["merge_kats"]:  # tests/test_coordinate_system_metadata.py:1032-1054
        df = _frame({"chrom": ["chr1", "chr1"], "start": kat["start"], "end": kat["end"]}, kat["zero_based"])
        assert len(pb.merge(df, output_type="pandas.DataFrame")) == kat["rows"]
    k = FX["unary_kats"]  # tests/test_partitioned_range_operation_regressions.py:24-59
    cols = ["contig", "pos_start", "pos_end"]
    left, right, view = _frame(k["left"]), _frame(k["right"]), _frame(k["view"])
    i64 = lambda d: pd.DataFrame(d).astype({c: "int64" for c in d if c != "contig"})
    pd.testing.assert_frame_equal(sort_all(pb.merge(left, cols=cols, output_type="pandas.DataFrame")), sort_all(i64(k["merge"])))
    pd.testing.assert_frame_equal(sort_all(pb.subtract(left, right, cols1=cols, cols2=cols, output_type="pandas.DataFrame")),
                                  sort_all(i64(k["subtract"])))
    got = pb.complement(left, view_df=view, cols=cols, view_cols=["chrom", "start", "end"], output_type="pandas.DataFrame")
    pd.testing.assert_frame_equal(sort_all(got), sort_all(i64(k["complement"])))
    pd.testing.assert_frame_equal(sort_all(pb.cluster(left, cols=cols, output_type="pandas.DataFrame")), sort_all(i64(k["cluster"])),
                                  check_dtype=False)
    no_view = pb.complement(left, cols=cols, output_type="pandas.DataFrame")
    assert no_view["pos_start"].tolist() == [30] and no_view["pos_end"].tolist() == [U.I64_MAX]


def test_api_parquet_fixtures_against_oracle():
    """The frames of tests/test_bioframe.py:112-126, 374-386, 455-480, 517-529 (exons / fBrain, 0-based), checked
    against the oracle instead of bioframe (absent here): full frame equality after sorting."""
    import polars_bio_b200 as pb
    from tests._golden import exons_fbrain_frames

    ex, fb = exons_fbrain_frames()
    ex.attrs["coordinate_system_zero_based"] = True
    fb.attrs["coordinate_system_zero_based"] = True
    cols = ("contig", "pos_start", "pos_end")
    z = exons_fbrain()
    names = [str(x) for x in z["contigs"]]  # sorted: codes are lexicographic ranks
    exc = z["exons_chrom"].astype(np.int64)
    mc, ms, me, mn = U.merge(exc, z["exons_start"], z["exons_end"], 24, True)
    want = pd.DataFrame({"contig": np.array(names)[mc], "pos_start": ms, "pos_end": me, "n_intervals": mn})
    got = pb.merge(ex, cols=cols, output_type="pandas.DataFrame")
    pd.testing.assert_frame_equal(sort_all(got), sort_all(want))
    cid, cs, ce = U.cluster(exc, z["exons_start"], z["exons_end"], 24, True)
    got = pb.cluster(ex, cols=cols, output_type="pandas.DataFrame")
    assert np.array_equal(got["cluster"].to_numpy(), cid) and np.array_equal(got["cluster_start"].to_numpy(), cs)
    assert np.array_equal(got["cluster_end"].to_numpy(), ce) and got["pos_start"].dtype == np.int32
    row, fs, fe = U.subtract_ranks(exc, z["exons_start"], z["exons_end"], z["fbrain_chrom"].astype(np.int64), z["fbrain_start"], z["fbrain_end"], 24, True)
    want = pd.DataFrame({"contig": np.array(names)[exc[row]], "pos_start": fs, "pos_end": fe})
    got = pb.subtract(ex, fb, cols1=cols, cols2=cols, output_type="pandas.DataFrame")
    pd.testing.assert_frame_equal(sort_all(got), sort_all(want))
    view = ex.groupby("contig").agg({"pos_start": "min", "pos_end": "max"}).reset_index().rename(
        columns={"contig": "chrom", "pos_start": "start", "pos_end": "end"})
    view.attrs["coordinate_system_zero_based"] = True
    vlut = {n: i for i, n in enumerate(names)}
    vcodes = view["chrom"].map(vlut).to_numpy(np.int64)
    vc, fs, fe = U.complement(exc, z["exons_start"], z["exons_end"], 24, True, view=(vcodes, view["start"].to_numpy(), view["end"].to_numpy()))
    want = pd.DataFrame({"contig": np.array(names)[vc], "pos_start": fs, "pos_end": fe})
    got = pb.complement(ex, view_df=view, cols=cols, view_cols=("chrom", "start", "end"), output_type="pandas.DataFrame")
    pd.testing.assert_frame_equal(sort_all(got), sort_all(want))


@pytest.mark.parametrize("strict", [True, False])
def test_arrow_level_entry_equals_the_device_level_host_layer(strict):
    """pb.merge / cluster / complement / subtract run through pbgpu_range_op (csrc/arrow_bridge.cpp run_unary); the same
    tables through polars_bio_b200/unary_op.py (pyarrow plumbing over pbgpu_merge / _cluster / _subtract) must give
    identical tables: contig order by name, dropped null keys, payload columns, Int64 positions."""
    import polars_bio_b200 as pb
    from polars_bio_b200 import FilterOp, unary_op

    rng = np.random.default_rng(5 if strict else 6)
    names = np.array(["chr2", "chr10", "chrX", "chr1", "GL000219.1", "alt_7"])
    fo = FilterOp.Strict if strict else FilterOp.Weak

    def table(n, null_rate, ctype):
        c = names[rng.integers(0, len(names), n)].tolist()
        s = rng.integers(0, 200_000, n)
        e = s + rng.integers(1, 3_000, n)
        chrom = pa.array([None if rng.random() < null_rate else x for x in c], type=ctype)
        start = pa.array([None if rng.random() < null_rate else int(v) for v in s], pa.int64())
        return pb.set_coordinate_system(pa.table({"chrom": chrom, "start": start, "end": pa.array(e.astype(np.int32)),
                                                  "score": pa.array(rng.random(n)), "tag": pa.array([f"row{i}" for i in range(n)])}), strict)

    for n, m, ctype in ((5_000, 2_000, pa.string()), (40_000, 30_000, pa.large_string()), (0, 10, pa.string()), (300, 0, pa.string())):
        t, r = table(n, 0.02, ctype), table(m, 0.02, pa.string())
        cols = ["chrom", "start", "end"]
        plain = lambda x: pa.Table.from_arrays([c.combine_chunks() for c in x.columns], names=x.column_names)  # values and types only: no field nullability / metadata
        key = lambda x: plain(x).sort_by([(c, "ascending") for c in x.column_names if c != "score"])
        for md in (0, 50):
            assert plain(pb.merge(t, min_dist=md, output_type="pyarrow.Table")).equals(plain(unary_op.merge_table(t, cols, fo, md)))
            assert key(pb.cluster(t, min_dist=md, output_type="pyarrow.Table")).equals(key(unary_op.cluster_table(t, cols, fo, md)))
        assert key(pb.subtract(t, r, output_type="pyarrow.Table")).equals(key(unary_op.subtract_table(t, r, cols, cols, fo)))
        view = pb.set_coordinate_system(pa.table({"chrom": names.tolist(), "start": [0] * len(names), "end": [250_000] * len(names)}), strict)
        assert key(pb.complement(t, view_df=view, output_type="pyarrow.Table")).equals(key(unary_op.complement_table(t, cols, fo, view, cols)))
        assert key(pb.complement(t, output_type="pyarrow.Table")).equals(key(unary_op.complement_table(t, cols, fo)))
