"""CPU tests of the unary sweeps: the oracle (oracle/unary_np.py) against the reference's golden vectors and against
its own independent twins, and the Python host layer (polars_bio_b200/unary_op.py + the facade) with the device
calls replaced by the oracle -- schemas, dtypes, contig ordering, null handling.  No GPU compute here."""
import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

from oracle import unary_np as U
from tests._golden import exons_fbrain, fixtures, sort_all

FX = fixtures()


def _codes(names_col):
    names = sorted(set(names_col))
    lut = {n: i for i, n in enumerate(names)}
    return np.array([lut[x] for x in names_col], np.int32), names


# ---- the oracle is pinned by the reference's fixtures -----------------------------------------------------------
def test_oracle_merge_golden_frame():
    g = FX["merge"]
    c, names = _codes(g["df"]["contig"])
    mc, ms, me, mn = U.merge(c, g["df"]["pos_start"], g["df"]["pos_end"], len(names), g["zero_based"])
    got = pd.DataFrame({"contig": [names[i] for i in mc], "pos_start": ms, "pos_end": me, "n_intervals": mn})
    pd.testing.assert_frame_equal(sort_all(got), sort_all(pd.DataFrame(g["expected"])))


@pytest.mark.parametrize("kat", FX["merge_kats"])
def test_oracle_merge_adjacency_kats(kat):
    assert len(U.merge([0, 0], kat["start"], kat["end"], 1, kat["zero_based"])[0]) == kat["rows"]


def test_oracle_partition_regression_kats():
    k = FX["unary_kats"]
    L, R, V = k["left"], k["right"], k["view"]
    z3, z2 = np.zeros(3, np.int32), np.zeros(2, np.int32)
    mc, ms, me, mn = U.merge(z3, L["pos_start"], L["pos_end"], 1, True)
    assert (ms.tolist(), me.tolist(), mn.tolist()) == (k["merge"]["pos_start"], k["merge"]["pos_end"], k["merge"]["n_intervals"])
    cid, cs, ce = U.cluster(z3, L["pos_start"], L["pos_end"], 1, True)
    got = sort_all(pd.DataFrame({"pos_start": L["pos_start"], "pos_end": L["pos_end"], "cluster": cid, "cluster_start": cs, "cluster_end": ce}))
    want = sort_all(pd.DataFrame({x: k["cluster"][x] for x in ("pos_start", "pos_end", "cluster", "cluster_start", "cluster_end")}))
    pd.testing.assert_frame_equal(got, want)
    for fn in (U.subtract, U.subtract_ranks, U.subtract_bruteforce):
        row, fs, fe = fn(z3, L["pos_start"], L["pos_end"], z2, R["pos_start"], R["pos_end"], 1, True)
        assert sorted(zip(fs.tolist(), fe.tolist())) == sorted(zip(k["subtract"]["pos_start"], k["subtract"]["pos_end"])), fn.__name__
    vc, fs, fe = U.complement(z3, L["pos_start"], L["pos_end"], 1, True, view=([0], V["start"], V["end"]))
    assert (fs.tolist(), fe.tolist()) == (k["complement"]["pos_start"], k["complement"]["pos_end"])
    vc, fs, fe = U.complement(z3, L["pos_start"], L["pos_end"], 1, True)  # no view: [0, i64::MAX) per contig
    assert (fs.tolist(), fe.tolist()) == ([30], [U.I64_MAX])


@pytest.mark.parametrize("seed", range(6))
def test_oracle_twins_agree_on_random_ragged_inputs(seed):
    rng = np.random.default_rng(seed)
    for _ in range(60):
        nc = 3
        n, m = int(rng.integers(0, 25)), int(rng.integers(0, 25))
        lc = rng.integers(-1, nc + 1, n); ls = rng.integers(0, 60, n); le = ls + rng.integers(-3, 15, n)
        rc = rng.integers(-1, nc + 1, m); rs = rng.integers(0, 60, m); re = rs + rng.integers(-3, 12, m)
        for strict in (True, False):
            a = U.subtract(lc, ls, le, rc, rs, re, nc, strict)
            b = U.subtract_ranks(lc, ls, le, rc, rs, re, nc, strict)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        # the brute-force twins (pairwise touch graph, position bitmaps) need proper intervals
        le2, re2 = ls + rng.integers(1, 15, n), rs + rng.integers(1, 12, m)
        ok_l, ok_r = np.clip(lc, 0, nc - 1), np.clip(rc, 0, nc - 1)
        for strict in (True, False):
            a = U.subtract(ok_l, ls, le2, ok_r, rs, re2, nc, strict)
            b = U.subtract_bruteforce(ok_l, ls, le2, ok_r, rs, re2, nc, strict)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
            for md in (0, 4):
                x = U.merge(rc, rs, re2, nc, strict, md)
                y = U.merge_bruteforce(rc, rs, re2, nc, strict, md)
                assert all(np.array_equal(p, q) for p, q in zip(x, y))


def test_oracle_twins_agree_on_parquet_fixture_sample():
    z = exons_fbrain()
    sl = slice(1000, 1400)
    a = U.subtract(z["exons_chrom"][sl], z["exons_start"][sl], z["exons_end"][sl], z["fbrain_chrom"], z["fbrain_start"], z["fbrain_end"], 24, True)
    b = U.subtract_ranks(z["exons_chrom"][sl], z["exons_start"][sl], z["exons_end"][sl], z["fbrain_chrom"], z["fbrain_start"], z["fbrain_end"], 24, True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[0]) > 300
    # merge is idempotent and conserves the row count
    mc, ms, me, mn = U.merge(z["exons_chrom"], z["exons_start"], z["exons_end"], 24, True)
    assert int(mn.sum()) == len(z["exons_chrom"])
    mc2, ms2, me2, mn2 = U.merge(mc, ms, me, 24, True)
    assert np.array_equal(ms, ms2) and np.array_equal(me, me2) and (mn2 == 1).all()


# ---- the public calls end to end on the CPU, two routes ----------------------------------------------------------
#   "c"  : the product route -- facade -> pbgpu_range_op of the harness build of csrc/arrow_bridge.cpp (tests/_harness.py),
#          whose unary device calls are plain CPU doubles: the C++ glue (run_unary) is what runs
#   "py" : polars_bio_b200/unary_op.py (the same host logic over the device-level calls) with the engine calls replaced
#          by the oracle
@pytest.fixture(params=["c", "py"])
def fake_device(request, monkeypatch):
    from polars_bio_b200 import _native, range_op_io

    if request.param == "c":
        from tests import _harness

        H = _harness.build()
        H.dbg_streams_ok(1)
        monkeypatch.setattr(_native, "lib", lambda: H)
        yield "c"
        H.dbg_streams_ok(0)
        return
    torch = pytest.importorskip("torch")
    from polars_bio_b200 import engine, unary_op
    from polars_bio_b200.options import RangeOp
    from polars_bio_b200.range_op_io import RangeResult, _df_to_reader

    t = lambda a, dt=np.int32: torch.from_numpy(np.ascontiguousarray(a, dtype=dt))
    monkeypatch.setattr(unary_op, "_to_device", lambda *cols: [t(c) for c in cols])

    def merge_intervals(c, s, e, nc, fo, md=0):
        mc, ms, me, mn = U.merge(c.numpy(), s.numpy(), e.numpy(), nc, fo == 1, md)
        return t(mc), t(ms), t(me), t(mn, np.int64)

    def cluster_intervals(c, s, e, nc, fo, md=0):
        cid, cs, ce = U.cluster(c.numpy(), s.numpy(), e.numpy(), nc, fo == 1, md)
        return t(cid, np.int64), t(cs), t(ce), int(cid.max()) + 1 if len(cid) else 0

    def subtract_intervals(lc, ls, le, rc, rs, re, nc, fo):
        row, fs, fe = U.subtract(lc.numpy(), ls.numpy(), le.numpy(), rc.numpy(), rs.numpy(), re.numpy(), nc, fo == 1)
        return t(row), t(fs), t(fe)

    monkeypatch.setattr(engine, "merge_intervals", merge_intervals)
    monkeypatch.setattr(engine, "cluster_intervals", cluster_intervals)
    monkeypatch.setattr(engine, "subtract_intervals", subtract_intervals)

    def unary_py(ctx, df1, df2, ro):
        t1 = _df_to_reader(df1).read_all()
        t2 = None if df2 is None else _df_to_reader(df2).read_all()
        if ro.range_op == RangeOp.Merge:
            out = unary_op.merge_table(t1, ro.columns_1, ro.filter_op, int(ro.min_dist or 0))
        elif ro.range_op == RangeOp.Cluster:
            out = unary_op.cluster_table(t1, ro.columns_1, ro.filter_op, int(ro.min_dist or 0))
        elif ro.range_op == RangeOp.Complement:
            out = unary_op.complement_table(t1, ro.columns_1, ro.filter_op, t2, ro.columns_2)
        else:
            out = unary_op.subtract_table(t1, t2, ro.columns_1, ro.columns_2, ro.filter_op)
        return RangeResult(out.to_reader())

    monkeypatch.setattr(range_op_io, "range_operation_unary", unary_py)
    yield "py"


def _frame(d, zero_based=True):
    df = pd.DataFrame(d)
    df.attrs["coordinate_system_zero_based"] = zero_based
    return df


def test_host_merge_golden_and_schema(fake_device):
    import polars_bio_b200 as pb

    g = FX["merge"]
    out = pb.merge(_frame(g["df"], g["zero_based"]), cols=("contig", "pos_start", "pos_end"), output_type="pandas.DataFrame")
    want = pd.DataFrame(g["expected"]).astype({"pos_start": "int64", "pos_end": "int64", "n_intervals": "int64"})
    pd.testing.assert_frame_equal(sort_all(out), sort_all(want))  # tests/test_native.py:206-223
    for kat in FX["merge_kats"]:
        df = _frame({"chrom": ["chr1", "chr1"], "start": kat["start"], "end": kat["end"]}, kat["zero_based"])
        assert len(pb.merge(df, output_type="pandas.DataFrame")) == kat["rows"]


def test_host_partition_regression_frames(fake_device):
    import polars_bio_b200 as pb

    k = FX["unary_kats"]
    cols = ["contig", "pos_start", "pos_end"]
    left, right, view = _frame(k["left"]), _frame(k["right"]), _frame(k["view"])
    i64 = lambda d: pd.DataFrame(d).astype({c: "int64" for c in d if c != "contig"})
    pd.testing.assert_frame_equal(sort_all(pb.merge(left, cols=cols, output_type="pandas.DataFrame")), sort_all(i64(k["merge"])))
    pd.testing.assert_frame_equal(sort_all(pb.subtract(left, right, cols1=cols, cols2=cols, output_type="pandas.DataFrame")),
                                  sort_all(i64(k["subtract"])))
    got = pb.complement(left, view_df=view, cols=cols, view_cols=["chrom", "start", "end"], output_type="pandas.DataFrame")
    pd.testing.assert_frame_equal(sort_all(got), sort_all(i64(k["complement"])))
    got = pb.cluster(left, cols=cols, output_type="pandas.DataFrame")
    pd.testing.assert_frame_equal(sort_all(got), sort_all(i64(k["cluster"])), check_dtype=False)
    assert got["cluster"].dtype == np.int64 and got["cluster_start"].dtype == np.int64
    no_view = pb.complement(left, cols=cols, output_type="pandas.DataFrame")
    assert no_view["pos_end"].tolist() == [U.I64_MAX] and no_view["pos_start"].tolist() == [30]


def test_host_contig_order_payload_nulls_and_types(fake_device):
    import polars_bio_b200 as pb

    # contigs appear out of lexicographic order; payload column, a null contig, a null start; large_string + int64 input
    t = pa.table({"chrom": pa.array(["chr2", "chr10", None, "chr2", "chr1", "chr10"], type=pa.large_string()),
                  "start": pa.array([5, 1, 3, 7, None, 100], type=pa.int64()),
                  "end": pa.array([10, 4, 9, 12, 8, 200], type=pa.int64()),
                  "name": ["a", "b", "c", "d", "e", "f"]})
    t = pb.set_coordinate_system(t, True)
    m = pb.merge(t, output_type="pyarrow.Table")
    assert m.column("chrom").to_pylist() == ["chr10", "chr10", "chr2"]  # lexicographic contig order, then start
    assert m.schema.field("chrom").type == pa.large_string() and m.schema.field("start").type == pa.int64()
    assert m.column("n_intervals").to_pylist() == [1, 1, 2]
    c = pb.cluster(t, output_type="pyarrow.Table")
    assert c.column_names == ["chrom", "start", "end", "name", "cluster", "cluster_start", "cluster_end"]
    assert c.column("name").to_pylist() == ["a", "b", "d", "f"]  # null-keyed rows take no part
    assert c.column("cluster").to_pylist() == [2, 0, 2, 1]      # ids: chr10 first (two clusters), then chr2
    other = pb.set_coordinate_system(pa.table({"chrom": ["chr2", "chr3"], "start": [6, 0], "end": [8, 50]}), True)
    s = pb.subtract(t, other, output_type="pyarrow.Table")
    assert s.column_names == ["chrom", "start", "end", "name"]
    rows = sorted(zip(s.column("name").to_pylist(), s.column("start").to_pylist(), s.column("end").to_pylist()))
    assert rows == [("a", 5, 6), ("a", 8, 10), ("b", 1, 4), ("d", 8, 12), ("f", 100, 200)]
    with pytest.raises(Exception, match="int32"):
        big = pb.set_coordinate_system(pa.table({"chrom": ["c"], "start": pa.array([0], pa.int64()), "end": pa.array([2**40], pa.int64())}), True)
        pb.merge(big, output_type="pyarrow.Table")
    with pytest.raises(AssertionError):
        pb.merge(t, on_cols=["name"], output_type="pyarrow.Table")  # on_cols unsupported, like the reference
