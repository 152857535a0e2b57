"""CPU-only tests of the Arrow-level host code (polars_bio_b200/csrc/arrow_bridge.cpp): the product source is compiled
with a stub CUDA header plus a small harness (tests/tools/bridge_harness/) so that stream draining, contig dictionary
encoding, position narrowing and the materialisation code (two-phase gather of fixed-width / bool / utf8 / large_utf8 /
binary / utf8_view / dictionary columns, nulls, NO_PARTNER rows, chunked inputs) run without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pyarrow as pa
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "tools", "bridge_harness")


class _CS(ctypes.Structure):
    _fields_ = [("get_schema", ctypes.c_void_p), ("get_next", ctypes.c_void_p), ("get_last_error", ctypes.c_void_p),
                ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p)]


@pytest.fixture(scope="module")
def lib():
    from tests import _harness

    return _harness.build()


def _roundtrip(lib, table, rows, cols=("chrom", "start", "end")):
    s_in, s_out = _CS(), _CS()
    table.to_reader()._export_to_c(ctypes.addressof(s_in))
    rows = np.asarray(rows, dtype=np.uint32)
    keys = np.zeros((table.num_rows, 3), dtype=np.int32)
    rc = lib.dbg_roundtrip(ctypes.addressof(s_in), *(c.encode() for c in cols), rows.ctypes.data, len(rows), keys.ctypes.data,
                           ctypes.addressof(s_out))
    assert rc == 0, rc
    return pa.RecordBatchReader._import_from_c(ctypes.addressof(s_out)).read_all(), keys


def test_gather_all_supported_types_with_nulls_and_missing_partner(lib):
    t = pa.table({
        "chrom": pa.array(["chr1", "chr1", "chr2", None, "chr9"], type=pa.string_view()),
        "start": pa.array([100, 500, 100, 7, 5], pa.int64()), "end": pa.array([200, 600, None, 9, 9], pa.uint32()),
        "tag": pa.array(["a-long-tag-beyond-12-bytes", "b", None, "", "d"], type=pa.string_view()),
        "blob": pa.array([b"\x00\x01", b"", b"xyz", None, b"q"], pa.binary()),
        "big": pa.array(["L0", None, "L2", "L3", "L4"], pa.large_string()),
        "f": pa.array([1.5, None, 3.5, 4.5, 5.5], pa.float64()), "b": pa.array([True, False, None, True, False]),
        "i8": pa.array([1, 2, 3, None, 5], pa.int8()), "ts": pa.array([1, 2, 3, 4, None], pa.timestamp("us")),
        "cat": pa.array(["x", "y", "x", None, "y"]).dictionary_encode(),
    })
    rows = [4, 0, 0xFFFFFFFF, 2, 3, 1]
    out, keys = _roundtrip(lib, t, rows)
    # keys: shared dictionary codes in order of first appearance; null contig / null position -> code -1
    assert keys[:, 0].tolist() == [0, 0, -1, -1, 2] or keys[:, 0].tolist() == [0, 0, -1, -1, 1]
    assert keys[0].tolist()[1:] == [100, 200] and keys[1].tolist()[1:] == [500, 600]
    exp = t.to_pylist()
    want = [exp[r] if r != 0xFFFFFFFF else {k: None for k in t.column_names} for r in rows]
    assert out.to_pylist() == want
    assert out.schema.field("chrom").type == pa.large_string() and out.schema.field("tag").type == pa.large_string()
    assert out.schema.field("cat").type == pa.large_string()            # dictionary<string> is decoded
    assert out.schema.field("blob").type == pa.binary() and out.schema.field("ts").type == pa.timestamp("us")
    assert out.schema.field("start").type == pa.int64() and out.schema.field("end").type == pa.uint32()


def test_chunked_input_and_many_rows(lib):
    rng = np.random.default_rng(0)
    n = 50_000
    parts = []
    for k in range(5):
        m = n // 5
        parts.append(pa.table({"chrom": pa.array(rng.choice(["chr1", "chr2", "chrX"], m)), "start": pa.array(rng.integers(0, 10**6, m), pa.int32()),
                               "end": pa.array(rng.integers(0, 10**6, m), pa.int32()),
                               "name": pa.array([f"r{k}_{i}" if i % 17 else None for i in range(m)])}))
    t = pa.concat_tables(parts)
    assert t.column("chrom").num_chunks == 5
    rows = rng.integers(0, n, 200_000).astype(np.uint32)
    rows[::1000] = 0xFFFFFFFF
    out, keys = _roundtrip(lib, t, rows)
    names = t.column("chrom").to_pylist()
    codes = {}
    for nm, c in zip(names, keys[:, 0]):
        assert codes.setdefault(nm, int(c)) == int(c)                    # one code per contig string
    assert len(set(codes.values())) == 3
    assert np.array_equal(keys[:, 1], t.column("start").to_numpy()) and np.array_equal(keys[:, 2], t.column("end").to_numpy())
    ref = t.to_pandas()
    got = out.to_pandas()
    ok = rows != 0xFFFFFFFF
    assert got["name"][ok].tolist() == ref["name"].to_numpy()[rows[ok]].tolist()
    assert got["start"][ok].astype("int64").tolist() == ref["start"].to_numpy()[rows[ok]].tolist()
    assert got["start"][~ok].isna().all() and got["chrom"][~ok].isna().all()


def test_count_passthrough_views_follow_input_batches(lib):
    # count_overlaps / coverage re-export the iterated table zero-copy: output batches = slices (<= 8 rows here) of the
    # input batches, sliced offsets must address the right rows for every type, incl. nested ones the gather cannot do
    parts = [pa.table({"chrom": pa.array([f"c{i % 3}" for i in range(lo, hi)]), "start": pa.array(range(lo, hi), pa.int32()),
                       "end": pa.array(range(lo + 5, hi + 5), pa.int32()), "name": pa.array([None if i % 4 == 0 else f"n{i}" for i in range(lo, hi)]),
                       "lst": pa.array([[i, i + 1] for i in range(lo, hi)], pa.list_(pa.int32()))})
             for lo, hi in ((0, 5), (5, 25), (25, 31))]
    t = pa.concat_tables(parts)
    s_in, s_out = _CS(), _CS()
    t.to_reader()._export_to_c(ctypes.addressof(s_in))
    rc = lib.dbg_roundtrip(ctypes.addressof(s_in), b"chrom", b"start", b"end", None, -1, None, ctypes.addressof(s_out))
    assert rc == 0
    batches = list(pa.RecordBatchReader._import_from_c(ctypes.addressof(s_out)))
    assert [b.num_rows for b in batches] == [5, 8, 8, 4, 6]
    out = pa.Table.from_batches(batches)
    assert out.column_names == t.column_names + ["count"]
    assert out.drop(["count"]).to_pylist() == t.to_pylist()
    assert out["count"].to_pylist() == [10 * i for i in range(31)]
    s_in, s_out = _CS(), _CS()
    t.to_reader()._export_to_c(ctypes.addressof(s_in))
    assert lib.dbg_roundtrip(ctypes.addressof(s_in), b"chrom", b"start", b"end", None, 7, None, ctypes.addressof(s_out)) == 0   # limit = 7
    lim = pa.RecordBatchReader._import_from_c(ctypes.addressof(s_out)).read_all()
    assert lim.num_rows == 7 and lim["start"].to_pylist() == list(range(7))


def test_int32_domain_check(lib):
    t = pa.table({"chrom": ["chr1"], "start": pa.array([2**31], pa.int64()), "end": pa.array([2**31 + 5], pa.int64())})
    s_in, s_out = _CS(), _CS()
    t.to_reader()._export_to_c(ctypes.addressof(s_in))
    rows = np.zeros(1, np.uint32)
    rc = lib.dbg_roundtrip(ctypes.addressof(s_in), b"chrom", b"start", b"end", rows.ctypes.data, 1, None, ctypes.addressof(s_out))
    assert rc == 4  # PBGPU_ERANGE


def test_key_encoding_streaming_stores_all_alignments(lib):
    """encode_keys writes the staging buffers with non-temporal stores (aligned 16-byte body + plain head / tail):
    long contig runs (fill path), ragged contigs (row-by-row blocks), int32 fast copies and narrowed int64 positions,
    over batches cut at odd offsets so heads and tails of every length occur."""
    rng = np.random.default_rng(7)
    n = 20_011
    contig = np.where(np.arange(n) < 9_000, "chr1", np.where(np.arange(n) < 15_000, "chrX", "chr7"))
    ragged = rng.integers(0, 3, n)
    contig = np.where((np.arange(n) > 12_000) & (np.arange(n) < 13_500), np.array(["chr1", "chrX", "chrUn_x"])[ragged], contig)
    start = rng.integers(0, 2**31 - 1000, n).astype(np.int32)
    end64 = (start.astype(np.int64) + rng.integers(0, 900, n))
    full = pa.table({"chrom": pa.array(contig), "start": pa.array(start), "end": pa.array(end64, pa.int64())})
    cuts = [0, 1, 4, 1031, 1033, 5000, 5003, 12_345, 16_384, 16_387, n]
    batches = [b for lo, hi in zip(cuts[:-1], cuts[1:]) for b in full.slice(lo, hi - lo).to_batches()]
    t = pa.Table.from_batches(batches)
    _, keys = _roundtrip(lib, t, [0])
    names = {}
    for nm, c in zip(contig, keys[:, 0]):
        assert names.setdefault(nm, c) == c
    assert len(set(names.values())) == len(names) == 4
    assert np.array_equal(keys[:, 1], start) and np.array_equal(keys[:, 2], end64.astype(np.int32))


def _random_column(rng, n, kind):
    null = rng.random(n) < rng.choice([0.0, 0.1, 0.6])
    def mask(vals):
        return [None if z else v for v, z in zip(vals, null)]
    if kind == "i8":
        return pa.array(mask(rng.integers(-128, 127, n).tolist()), pa.int8())
    if kind == "u16":
        return pa.array(mask(rng.integers(0, 65535, n).tolist()), pa.uint16())
    if kind == "i64":
        return pa.array(mask(rng.integers(-2**62, 2**62, n).tolist()), pa.int64())
    if kind == "f32":
        return pa.array(mask(rng.random(n).astype(np.float32).tolist()), pa.float32())
    if kind == "bool":
        return pa.array(mask((rng.random(n) < 0.5).tolist()), pa.bool_())
    if kind == "date":
        return pa.array(mask(rng.integers(0, 20000, n).tolist()), pa.date32())
    words = ["", "a", "chrUn_KI270742v1", "x" * 40, "naïve-ütf8", "t\t\n", "0123456789ab", "0123456789abc"]
    vals = mask([words[i] for i in rng.integers(0, len(words), n)])
    if kind == "utf8":
        return pa.array(vals, pa.string())
    if kind == "large_utf8":
        return pa.array(vals, pa.large_string())
    if kind == "view":
        return pa.array(vals, pa.string_view())
    if kind == "binary":
        return pa.array([None if v is None else v.encode() for v in vals], pa.binary())
    if kind == "dict":
        return pa.array(vals, pa.string()).dictionary_encode()
    raise AssertionError(kind)


@pytest.mark.parametrize("seed", range(12))
def test_gather_random_tables_equal_pyarrow_take(lib, seed):
    """Materialisation of arbitrary result rows (any supported payload type, any null density, chunked and sliced
    inputs, empty tables, missing partners) == pyarrow's own take; utf8_view and dictionary columns come back as large_utf8."""
    rng = np.random.default_rng(1000 + seed)
    kinds = ["i8", "u16", "i64", "f32", "bool", "date", "utf8", "large_utf8", "view", "binary", "dict"]
    n = int(rng.choice([0, 1, 7, 8, 9, 63, 64, 65, 1000, 20_000]))
    chosen = list(rng.choice(kinds, size=int(rng.integers(1, 6)), replace=False))
    cols = {"chrom": pa.array([None if z else f"chr{c}" for c, z in zip(rng.integers(0, 4, n), rng.random(n) < 0.05)], pa.string()),
            "start": pa.array(rng.integers(0, 10**6, n), pa.int32()), "end": pa.array(rng.integers(0, 10**6, n), pa.int64())}
    for j, k in enumerate(chosen):
        cols[f"p{j}_{k}"] = _random_column(rng, n, k)
    t = pa.table(cols)
    if n > 4:  # uneven chunks, the first one sliced off a larger array (non-zero offsets)
        cuts = sorted(set([0, n] + rng.integers(1, n, 3).tolist()))
        pad = pa.concat_tables([t.slice(0, 3), t])
        t = pa.concat_tables([pad.slice(3 + lo, hi - lo) for lo, hi in zip(cuts[:-1], cuts[1:])])
    m = int(rng.choice([0, 1, 5, 4097, 40_000]))
    rows = rng.integers(0, max(n, 1), m).astype(np.uint32) if n else np.full(m, 0xFFFFFFFF, np.uint32)
    if m:
        rows[rng.random(m) < 0.1] = 0xFFFFFFFF
    out, _ = _roundtrip(lib, t, rows)
    assert out.num_rows == m and out.column_names == t.column_names
    idx = pa.array([None if r == 0xFFFFFFFF else int(r) for r in rows], pa.int64())
    want = t.combine_chunks().take(idx) if n else pa.table({c: pa.nulls(m, t.schema.field(c).type) for c in t.column_names})
    for name in t.column_names:
        got_c, want_c = out[name].combine_chunks(), want[name].combine_chunks()
        if pa.types.is_dictionary(want_c.type) or pa.types.is_string_view(want_c.type):  # decoded / flattened to large_utf8
            if pa.types.is_dictionary(want_c.type):
                want_c = want_c.dictionary_decode()
            assert got_c.type == pa.large_string(), (name, got_c.type)
            assert got_c.to_pylist() == want_c.to_pylist(), name
        else:
            assert got_c.type == want_c.type, (name, got_c.type, want_c.type)
            assert got_c.to_pylist() == want_c.to_pylist(), name


# ---- unary sweeps through pbgpu_range_op (run_unary) against the oracle, device calls = the harness's CPU doubles ----
def _unary(lib, op, left, right=None, cols1=("chrom", "start", "end"), cols2=("chrom", "start", "end"), strict=True, min_dist=0):
    from polars_bio_b200 import _native

    lib.dbg_streams_ok(1)
    try:
        s1, s2, so = _CS(), _CS(), _CS()
        (left if isinstance(left, pa.RecordBatchReader) else left.to_reader())._export_to_c(ctypes.addressof(s1))
        if right is not None:
            right.to_reader()._export_to_c(ctypes.addressof(s2))
        o = _native.PbRangeOptions()
        o.range_op, o.filter_op, o.min_dist, o.device = op, 1 if strict else 0, min_dist, -1
        for i in range(3):
            o.cols1[i] = cols1[i].encode()
            o.cols2[i] = cols2[i].encode()
        rc = lib.pbgpu_range_op(ctypes.addressof(s1), ctypes.addressof(s2) if right is not None else None, ctypes.addressof(o), ctypes.addressof(so))
        assert rc == 0, (rc, lib.pbgpu_last_error())
        return pa.RecordBatchReader._import_from_c(ctypes.addressof(so)).read_all()
    finally:
        lib.dbg_streams_ok(0)


def _random_table(rng, n, contig_type, names, null_rate=0.1):
    c = rng.integers(0, len(names), n)
    s = rng.integers(0, 300, n)
    e = s + rng.integers(1, 40, n)
    chrom = [None if rng.random() < null_rate else names[k] for k in c]
    start = [None if rng.random() < null_rate / 2 else int(v) for v in s]
    arr = pa.array(chrom, type=pa.large_string())
    if contig_type == "dict":
        arr = pa.array(chrom, type=pa.string()).dictionary_encode()
    elif contig_type == "utf8":
        arr = arr.cast(pa.string())
    elif contig_type == "view":  # built directly: pyarrow's exporter crashes on sliced views cast from large_string
        arr = pa.array(chrom, type=pa.string_view())
    return pa.table({"chrom": arr, "start": pa.array(start, pa.int64()), "end": pa.array(e.astype(np.int32)), "tag": pa.array([f"r{i}" for i in range(n)])})


def _keys_np(t, names_sorted):
    lut = {n: i for i, n in enumerate(names_sorted)}
    chrom, start, end = t.column("chrom").to_pylist(), t.column("start").to_pylist(), t.column("end").to_pylist()
    ok = [c is not None and s is not None and e is not None for c, s, e in zip(chrom, start, end)]
    c = np.array([lut[x] if k else -1 for x, k in zip(chrom, ok)], np.int32)
    return c, np.array([s if k else 0 for s, k in zip(start, ok)], np.int32), np.array([e if k else 0 for e, k in zip(end, ok)], np.int32)


@pytest.mark.parametrize("contig_type", ["utf8", "large", "dict", "view"])
@pytest.mark.parametrize("strict", [True, False])
def test_unary_sweeps_through_the_c_entry_match_the_oracle(lib, contig_type, strict):
    from oracle import unary_np as U

    rng = np.random.default_rng(11 + (7 if strict else 0))
    names = ["chr2", "chr10", "chrX", "chr1", "alt_7"]  # first-seen order differs from name order
    for trial in range(6):
        t = _random_table(rng, int(rng.integers(0, 60)), contig_type, names)
        r = _random_table(rng, int(rng.integers(0, 40)), "utf8", names + ["only_right"])
        batches = t.to_batches(max_chunksize=int(rng.integers(3, 20)))
        reader = lambda: pa.RecordBatchReader.from_batches(t.schema, batches)
        present = sorted({x for x in t.column("chrom").to_pylist() + r.column("chrom").to_pylist() if x is not None})
        md = int(rng.integers(0, 3)) * 5
        # merge: names of the merged table come from the left table only -> same relative (name) order
        sorted_l = sorted({x for x in t.column("chrom").to_pylist() if x is not None})
        c, s, e = _keys_np(t, sorted_l)
        mc, ms, me, mn = U.merge(c, s, e, len(sorted_l), strict, md)
        got = _unary(lib, 7, reader(), strict=strict, min_dist=md)
        assert got.column_names == ["chrom", "start", "end", "n_intervals"]
        assert got.column("chrom").to_pylist() == [sorted_l[k] for k in mc]
        assert got.column("start").to_pylist() == ms.tolist() and got.column("end").to_pylist() == me.tolist() and got.column("n_intervals").to_pylist() == mn.tolist()
        assert got.schema.field("start").type == pa.int64() and got.schema.field("chrom").type == (pa.string() if contig_type == "utf8" else pa.large_string())
        # cluster: every input column passes through untouched (any type), null-keyed rows dropped
        cid, cs, ce = U.cluster(c, s, e, len(sorted_l), strict, md)
        keep = cid >= 0
        got = _unary(lib, 2, reader(), strict=strict, min_dist=md)
        assert got.column_names == ["chrom", "start", "end", "tag", "cluster", "cluster_start", "cluster_end"]
        assert got.schema.field("chrom").type == t.schema.field("chrom").type
        assert got.column("tag").to_pylist() == [x for x, k in zip(t.column("tag").to_pylist(), keep) if k]
        assert got.column("cluster").to_pylist() == cid[keep].tolist()
        assert got.column("cluster_start").to_pylist() == cs[keep].tolist() and got.column("cluster_end").to_pylist() == ce[keep].tolist()
        # subtract: shared dictionary over both tables
        lc, ls, le = _keys_np(t, present)
        rc_, rs, re_ = _keys_np(r, present)
        row, fs, fe = U.subtract(lc, ls, le, rc_, rs, re_, len(present), strict)
        got = _unary(lib, 5, reader(), r, strict=strict)
        assert got.column_names == ["chrom", "start", "end", "tag"]
        want = sorted(zip([f"r{k}" for k in row], fs.tolist(), fe.tolist()))
        assert sorted(zip(got.column("tag").to_pylist(), got.column("start").to_pylist(), got.column("end").to_pylist())) == want
        tl = t.column("chrom").to_pylist()
        assert got.column("chrom").to_pylist() == [tl[int(x[1:])] for x in got.column("tag").to_pylist()]
        # complement inside a view table, and with the default view
        view = pa.table({"chrom": pa.array(present), "start": pa.array([0] * len(present), pa.int64()), "end": pa.array([500] * len(present), pa.int64())})
        vc = np.arange(len(present), dtype=np.int32)
        kc, fs, fe = U.complement(lc, ls, le, len(present), strict, view=(vc, np.zeros(len(present), np.int32), np.full(len(present), 500, np.int32)))
        got = _unary(lib, 1, reader(), view, strict=strict)
        assert got.column_names == ["chrom", "start", "end"]
        assert sorted(zip(got.column("chrom").to_pylist(), got.column("start").to_pylist(), got.column("end").to_pylist())) == \
            sorted(zip([present[k] for k in kc], fs.tolist(), fe.tolist()))
        kc, fs, fe = U.complement(c, s, e, len(sorted_l), strict)
        got = _unary(lib, 1, reader(), strict=strict)
        assert sorted(zip(got.column("chrom").to_pylist(), got.column("start").to_pylist(), got.column("end").to_pylist())) == \
            sorted(zip([sorted_l[k] for k in kc], fs.tolist(), fe.tolist()))


def test_unary_entry_errors(lib):
    from polars_bio_b200 import _native

    t = pa.table({"chrom": ["a"], "start": [1], "end": [5]})
    lib.dbg_streams_ok(1)
    try:
        for op, right, cols, code in ((5, None, ("chrom", "start", "end"), 1), (7, None, ("chrom", "nope", "end"), 5)):
            s1, so = _CS(), _CS()
            t.to_reader()._export_to_c(ctypes.addressof(s1))
            o = _native.PbRangeOptions()
            o.range_op, o.filter_op, o.device = op, 1, -1
            for i in range(3):
                o.cols1[i] = cols[i].encode()
                o.cols2[i] = cols[i].encode()
            assert lib.pbgpu_range_op(ctypes.addressof(s1), None, ctypes.addressof(o), ctypes.addressof(so)) == code
            assert s1.release is None  # the input was moved (released) although the call failed
    finally:
        lib.dbg_streams_ok(0)


@pytest.mark.parametrize("n_names", [3, 24, 56, 57, 400])
def test_encoder_codes_are_consistent_for_many_short_and_long_names(lib, n_names):
    """The short-name table of the key encoder (collision-free multiplicative hash, at most 56 names; others take the slow
    path): every row of a name gets the same code, different names get different codes, whatever the number of names,
    their lengths (1 .. 12 bytes) and their order."""
    rng = np.random.default_rng(n_names)
    names = sorted({("c" * int(rng.integers(0, 5))) + str(int(rng.integers(0, 10**int(rng.integers(1, 8))))) for _ in range(3 * n_names)})[:n_names]
    n = 40_000
    pick = rng.integers(0, len(names), n)
    for ctype in (pa.string(), pa.large_string()):
        t = pa.table({"chrom": pa.array([names[k] for k in pick], type=ctype), "start": pa.array(np.arange(n, dtype=np.int32)),
                      "end": pa.array(np.arange(n, dtype=np.int32) + 5)})
        _, keys = _roundtrip(lib, t, [0])
        codes = keys[:, 0]
        assert codes.min() >= 0
        first = {}
        for k, c in zip(pick.tolist(), codes.tolist()):
            assert first.setdefault(k, c) == c
        assert len(set(first.values())) == len(first)
        assert np.array_equal(keys[:, 1], np.arange(n)) and np.array_equal(keys[:, 2], np.arange(n) + 5)


def test_count_overlaps_through_the_c_entry_with_more_than_255_indexed_contigs(lib):
    """The indexed side's contig codes are staged as bytes; a table with more than 255 contigs is encoded again with 32-bit
    codes.  300 contigs on the indexed side, 320 on the iterated one (20 of them unknown to the index)."""
    from polars_bio_b200 import _native

    rng = np.random.default_rng(4)
    names = np.array([f"ctg{i}" for i in range(320)])
    m, n = 900, 20_000
    xc, xs = rng.integers(0, 300, m), rng.integers(0, 5_000, m).astype(np.int32)
    xc[:300] = np.arange(300)  # every one of the 300 occurs
    xe = (xs + rng.integers(1, 900, m)).astype(np.int32)
    ic, ps = rng.integers(0, 320, n), rng.integers(0, 5_000, n).astype(np.int32)
    pe = (ps + rng.integers(1, 200, n)).astype(np.int32)
    want = np.zeros(n, np.int64)
    for j in range(m):
        want += (ic == xc[j]) & (ps < xe[j]) & (pe > xs[j])
    lib.dbg_streams_ok(1)
    try:
        s1, s2, so = _CS(), _CS(), _CS()
        pa.table({"chrom": names[xc].tolist(), "start": xs, "end": xe}).to_reader()._export_to_c(ctypes.addressof(s1))
        pa.table({"chrom": names[ic].tolist(), "start": ps, "end": pe}).to_reader()._export_to_c(ctypes.addressof(s2))
        o = _native.PbRangeOptions()
        o.range_op, o.filter_op, o.device = 6, 1, -1
        for i, c in enumerate(("chrom", "start", "end")):
            o.cols1[i] = c.encode()
            o.cols2[i] = c.encode()
        rc = lib.pbgpu_range_op(ctypes.addressof(s1), ctypes.addressof(s2), ctypes.addressof(o), ctypes.addressof(so))
        assert rc == 0, (rc, lib.pbgpu_last_error())
        out = pa.RecordBatchReader._import_from_c(ctypes.addressof(so)).read_all()
    finally:
        lib.dbg_streams_ok(0)
    assert np.array_equal(out.column("count").to_numpy(), want) and want.sum() > 0


@pytest.mark.parametrize("strict", [True, False])
def test_count_overlaps_through_the_c_entry_with_the_helper_thread_build(lib, strict):
    """pbgpu_range_op(count_overlaps) end to end on the harness: the indexed side's upload + build run on a helper thread
    (the double sleeps 3 ms) while this thread encodes the iterated table in three slices; slice kernels are started when
    the index is ready; counts are widened and the iterated table's columns re-exported zero-copy.  Device calls are the
    harness's brute-force doubles; expectation by numpy."""
    from polars_bio_b200 import _native

    rng = np.random.default_rng(21)
    names = np.array(["chr1", "chr2", "chr3", "chrUn_long_name"])
    n, m = 2_500_000, 40
    ic, ps = rng.integers(0, 4, n), rng.integers(0, 10_000, n).astype(np.int32)
    pe = (ps + rng.integers(1, 300, n)).astype(np.int32)
    xc, xs = rng.integers(0, 3, m), rng.integers(0, 10_000, m).astype(np.int32)  # chrUn...: only on the iterated side
    xe = (xs + rng.integers(1, 2_000, m)).astype(np.int32)
    indexed = pa.table({"chrom": names[xc].tolist(), "start": xs, "end": xe})
    iterated = pa.table({"chrom": pa.array(names[ic].tolist(), pa.large_string()), "start": ps, "end": pe, "tag": pa.array(np.arange(n, dtype=np.int64))})
    want = np.zeros(n, np.int64)
    for j in range(m):
        hit = (ic == xc[j]) & ((ps < xe[j]) & (pe > xs[j]) if strict else (ps <= xe[j]) & (pe >= xs[j]))
        want += hit
    lib.dbg_streams_ok(1)
    try:
        s1, s2, so = _CS(), _CS(), _CS()
        indexed.to_reader()._export_to_c(ctypes.addressof(s1))                        # count_overlaps: `left` is the indexed table,
        pa.RecordBatchReader.from_batches(iterated.schema, iterated.to_batches(max_chunksize=700_001))._export_to_c(ctypes.addressof(s2))  # `right` the iterated one
        o = _native.PbRangeOptions()
        o.range_op, o.filter_op, o.device = 6, 1 if strict else 0, -1
        for i, c in enumerate(("chrom", "start", "end")):
            o.cols1[i] = c.encode()
            o.cols2[i] = c.encode()
        rc = lib.pbgpu_range_op(ctypes.addressof(s1), ctypes.addressof(s2), ctypes.addressof(o), ctypes.addressof(so))
        assert rc == 0, (rc, lib.pbgpu_last_error())
        out = pa.RecordBatchReader._import_from_c(ctypes.addressof(so)).read_all()
    finally:
        lib.dbg_streams_ok(0)
    assert out.column_names == ["chrom", "start", "end", "tag", "count"] and out.num_rows == n
    assert np.array_equal(out.column("count").to_numpy(), want)
    assert np.array_equal(out.column("tag").to_numpy(), np.arange(n)) and out.schema.field("chrom").type == pa.large_string()
