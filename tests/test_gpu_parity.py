"""GPU parity tests proper: the CUDA path (through the C ABI of include/pbgpu.h, driven via
polars_bio_b200.engine) against the CPU oracle on the same seeded inputs, against the committed
golden fixtures, and -- at BASELINE.json sizes -- through size-independent properties.
Bit-exact: integer / index work only."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from tests._golden import exons_fbrain, fixtures, synth  # noqa: E402

FX = fixtures()


def _engine():
    from polars_bio_b200 import engine

    return engine


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to("cuda:0")


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


def _run_all(pc, ps, pe, bc, bs, be, nc, strict, ks=((1, True), (3, True), (1, False), (4, False))):
    eng = _engine()
    fo = eng.FILTER_STRICT if strict else eng.FILTER_WEAK
    ix = eng.DeviceIndex(_dev(bc), _dev(bs), _dev(be), nc)
    oix = oracle.Index(bc, bs, be, nc)
    dpc, dps, dpe = _dev(pc), _dev(ps), _dev(pe)
    cnt = ix.count_overlaps(dpc, dps, dpe, fo).cpu().numpy()
    assert np.array_equal(cnt, oix.count_overlaps(pc, ps, pe, strict))
    a, b = ix.overlap_pairs(dpc, dps, dpe, fo)
    oa, ob = oix.overlap_pairs(pc, ps, pe, strict)
    assert len(a) == len(oa)
    assert np.array_equal(_u32(a), oa) and np.array_equal(_u32(b), ob)  # same order too: (probe, start, row)
    # streaming sink: the same pairs in the same order through a two-slot ring of small chunks
    for cap in (1, 257, 1 << 14):
        chunks = [(_u32(x).copy(), _u32(y).copy()) for x, y in ix.overlap_pairs_stream(dpc, dps, dpe, fo, max_pairs=cap)]
        sa = np.concatenate([x for x, _ in chunks]) if chunks else np.zeros(0, np.uint32)
        sb = np.concatenate([y for _, y in chunks]) if chunks else np.zeros(0, np.uint32)
        assert np.array_equal(sa, oa) and np.array_equal(sb, ob), cap
    cov = ix.coverage(dpc, dps, dpe, fo).cpu().numpy()
    assert np.array_equal(cov, oix.coverage(pc, ps, pe, strict))
    for k, inc in ks:
        p, d = ix.nearest(dpc, dps, dpe, fo, k=k, include_overlaps=inc)
        op, od = oix.nearest(pc, ps, pe, strict, k=k, include_overlaps=inc)
        assert np.array_equal(d.cpu().numpy(), od), (k, inc)
        assert np.array_equal(_u32(p), op), (k, inc)
    ix.close()
    return cnt


@pytest.mark.parametrize("strict", [True, False])
def test_golden_csv_fixtures(strict):
    # tests/data/overlap/{reads,targets}.csv (reference tests/test_native.py:32-52): 16 pairs in Weak mode
    fx = FX["overlap"]
    lc, rc, names = oracle.encode_contigs(fx["df1"]["contig"], fx["df2"]["contig"])
    cnt = _run_all(lc, np.array(fx["df1"]["pos_start"], np.int32), np.array(fx["df1"]["pos_end"], np.int32),
                   rc, np.array(fx["df2"]["pos_start"], np.int32), np.array(fx["df2"]["pos_end"], np.int32),
                   len(names), strict)
    if not strict:
        assert cnt.sum() == 16


def test_golden_count_and_nearest_values():
    eng = _engine()
    fx = FX["count_overlaps"]
    lc, rc, names = oracle.encode_contigs(fx["df1"]["contig"], fx["df2"]["contig"])
    ix = eng.DeviceIndex(_dev(rc), _dev(fx["df2"]["pos_start"]), _dev(fx["df2"]["pos_end"]), len(names))
    cnt = ix.count_overlaps(_dev(lc), _dev(fx["df1"]["pos_start"]), _dev(fx["df1"]["pos_end"]), eng.FILTER_WEAK)
    # tests/_expected.py:183-202 (rows are in input order there too)
    assert cnt.cpu().tolist() == [2, 2, 2, 1, 1, 2, 2, 2, 1, 1, 0]
    p, d = ix.nearest(_dev(lc), _dev(fx["df1"]["pos_start"]), _dev(fx["df1"]["pos_end"]), eng.FILTER_WEAK)
    assert sorted(d[:, 0].cpu().tolist()) == sorted(FX["nearest"]["expected"]["distance"])  # incl. the 34


@pytest.mark.parametrize("kat", FX["overlap_kats"])
def test_boundary_kats(kat):
    # tests/test_coordinate_system_metadata.py:735-818
    eng = _engine()
    z = np.zeros(1, np.int32)
    fo = eng.FILTER_STRICT if kat["zero_based"] else eng.FILTER_WEAK
    ix = eng.DeviceIndex(_dev(z), _dev([kat["b"][0]]), _dev([kat["b"][1]]), 1)
    a, _ = ix.overlap_pairs(_dev(z), _dev([kat["a"][0]]), _dev([kat["a"][1]]), fo)
    assert a.numel() == kat["rows"]
    assert int(ix.count_overlaps(_dev(z), _dev([kat["a"][0]]), _dev([kat["a"][1]]), fo)[0]) == kat["rows"]


@pytest.mark.parametrize("kat", FX["coverage_kats"])
def test_coverage_kats(kat):
    eng = _engine()
    z = np.zeros(1, np.int32)
    fo = eng.FILTER_STRICT if kat["zero_based"] else eng.FILTER_WEAK
    ix = eng.DeviceIndex(_dev(z), _dev([kat["b"][0]]), _dev([kat["b"][1]]), 1)
    assert int(ix.coverage(_dev(z), _dev([kat["a"][0]]), _dev([kat["a"][1]]), fo)[0]) == kat["coverage"]


@pytest.mark.parametrize("strict", [True, False])
def test_exons_fbrain_54246(strict):
    # docs/supplement.md:108,149 published pair count; both join directions; full pair parity
    z = exons_fbrain()
    nc = len(z["contigs"])
    ec, es, ee = z["exons_chrom"].astype(np.int32), z["exons_start"], z["exons_end"]
    fc, fs, fe = z["fbrain_chrom"].astype(np.int32), z["fbrain_start"], z["fbrain_end"]
    want = FX["exons_fbrain_pairs"]["strict" if strict else "weak"]
    cnt = _run_all(ec, es, ee, fc, fs, fe, nc, strict, ks=((1, True), (2, False)))
    assert cnt.sum() == want
    cnt2 = _run_all(fc, fs, fe, ec, es, ee, nc, strict, ks=((1, True),))
    assert cnt2.sum() == want


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("seed", [0, 1])
def test_random_ragged(strict, seed):
    # zero-length intervals, contigs on one side only, null keys (-1 / out of range), long shadowing intervals
    pc, ps, pe = synth(5000, 6, 50_000, 200, seed, zero_len_frac=0.05)
    bc, bs, be = synth(3000, 5, 50_000, 5000, 10 + seed, zero_len_frac=0.05)
    pc[::97] = -1
    bc[::89] = 7
    _run_all(pc, ps, pe, bc, bs, be, 6, strict)


@pytest.mark.parametrize("strict", [True, False])
def test_inverted_and_negative_coordinates(strict):
    # start > end rows are evaluated by the bare predicate (rank identity must switch itself off);
    # negative coordinates exercise the biased radix keys
    rng = np.random.default_rng(5)
    pc, ps, pe = synth(2000, 2, 4000, 100, 3)
    bc, bs, be = synth(1500, 2, 4000, 300, 4)
    ps -= 2000; pe -= 2000; bs -= 2000; be -= 2000
    inv = rng.random(1500) < 0.1
    bs2, be2 = np.where(inv, be, bs), np.where(inv, bs, be)
    inv = rng.random(2000) < 0.1
    ps2, pe2 = np.where(inv, pe, ps), np.where(inv, ps, pe)
    _run_all(pc, ps2.astype(np.int32), pe2.astype(np.int32), bc, bs2.astype(np.int32), be2.astype(np.int32), 2, strict,
             ks=((1, True),))


@pytest.mark.parametrize("strict", [True, False])
def test_rank_directory_crowded_buckets_and_contig_edges(strict):
    """Fast path specifics: many indexed rows per directory bucket (>7 -> in-bucket search), probes hanging over
    the first/last indexed coordinate of a contig (clamping onto the global axis), neighbouring contigs whose
    slices touch, and empty / inverted probe intervals (generic path inside the fast kernels)."""
    rng = np.random.default_rng(21)
    m = 6000
    bc = rng.integers(0, 4, m).astype(np.int32)
    centers = rng.choice([1000, 1010, 50_000, 50_003, 900_000], m)          # heavy clustering -> crowded buckets
    bs = (centers + rng.integers(0, 6, m)).astype(np.int32)
    be = (bs + rng.integers(1, 40, m)).astype(np.int32)
    n = 8000
    pc = rng.integers(0, 5, n).astype(np.int32)                              # contig 4 has no indexed rows
    ps = rng.choice([0, 900, 995, 1040, 49_990, 899_990, 900_050, 2_000_000], n).astype(np.int32) + rng.integers(0, 30, n).astype(np.int32)
    pe = (ps + rng.integers(0, 60, n)).astype(np.int32)                      # includes empty probes
    pe[::13] = ps[::13] - 5                                                  # inverted probes
    cnt = _run_all(pc, ps, pe, bc, bs, be, 5, strict, ks=((1, True), (2, False)))
    assert cnt.max() > 64


@pytest.mark.parametrize("strict", [True, False])
def test_flat_emit_on_index_without_nested_intervals(strict):
    """Pass 2 as a pure expansion of (count, start rank) -- taken when no indexed interval contains a later-starting
    one (fixed-length rows: ends ascend with starts): several contigs, duplicates, probes with 0 / few / >32 / >1000
    hits in the same warp, empty + inverted + null-keyed probes, streaming block ranges (inside _run_all)."""
    rng = np.random.default_rng(31)
    m = 40_000
    bc = rng.integers(0, 4, m).astype(np.int32)
    bs = rng.integers(0, 200_000, m).astype(np.int32)
    bs[:2000] = rng.choice([77, 5000, 5001], 2000)                            # duplicates / crowded buckets
    be = (bs + 25).astype(np.int32)                                          # fixed length: nothing nests
    n = 9000
    pc = rng.integers(0, 5, n).astype(np.int32)                              # contig 4 has no indexed rows
    ps = rng.integers(-50, 200_100, n).astype(np.int32)
    ln = rng.choice([0, 1, 10, 150, 700, 30_000], n, p=[0.05, 0.2, 0.3, 0.3, 0.1, 0.05])
    pe = (ps + ln).astype(np.int32)
    pe[::17] = ps[::17] - 3                                                  # inverted probes (generic path)
    pc[::41] = -1
    cnt = _run_all(pc, ps, pe, bc, bs, be, 5, strict, ks=((1, True), (3, False)))
    assert cnt.max() > 1000 and (cnt == 0).sum() > 100


def test_sort_over_many_tiles_and_directory_gaps():
    """Index build at a size where the radix sort spans many tiles (decoupled look-back across tiles), with equal
    keys in different tiles (stability: row order among equal starts), 24 contigs and two far-apart clusters per
    contig (long runs of empty directory buckets filled by whole warps)."""
    rng = np.random.default_rng(41)
    m = 300_000
    bc = rng.integers(0, 24, m).astype(np.int32)
    near = rng.random(m) < 0.5
    bs = np.where(near, rng.integers(0, 4000, m), 90_000_000 + rng.integers(0, 4000, m)).astype(np.int32)
    be = (bs + 1).astype(np.int32)
    n = 50_000
    pc = rng.integers(0, 24, n).astype(np.int32)
    ps = np.where(rng.random(n) < 0.5, rng.integers(0, 4100, n), 90_000_000 + rng.integers(-100, 4100, n)).astype(np.int32)
    pe = (ps + rng.integers(1, 6, n)).astype(np.int32)
    _run_all(pc, ps, pe, bc, bs, be, 24, True, ks=((1, True),))


def test_span_beyond_32_bits_takes_generic_path():
    # three contigs spanning ~2^31 each: the global axis does not fit uint32, the index must fall back
    rng = np.random.default_rng(3)
    m, n = 3000, 4000
    bc = rng.integers(0, 3, m).astype(np.int32)
    bs = rng.integers(0, 2**31 - 5000, m).astype(np.int32)
    bs[:3] = 0; bs[3:6] = 2**31 - 4000
    be = (bs + rng.integers(1, 3000, m)).astype(np.int32)
    pc = rng.integers(0, 3, n).astype(np.int32)
    ps = rng.integers(0, 2**31 - 5000, n).astype(np.int32)
    pe = (ps + rng.integers(1, 4_000_000, n).clip(max=2**31 - 1 - ps.astype(np.int64))).astype(np.int32)
    _run_all(pc, ps, pe, bc, bs, be, 3, True, ks=((1, True),))


def test_int32_extremes():
    # coordinates at the edge of the int32 domain: no end+1 / start-1 arithmetic may overflow
    lo, hi = -2**31, 2**31 - 1
    bc = np.zeros(6, np.int32)
    bs = np.array([lo, lo, hi - 10, hi - 1, 0, hi], np.int32)
    be = np.array([lo + 5, hi, hi, hi, 10, hi], np.int32)
    pc = np.zeros(7, np.int32)
    ps = np.array([lo, lo, hi - 5, hi, -5, hi - 1, lo + 5], np.int32)
    pe = np.array([lo, lo + 1, hi, hi, 5, hi, lo + 6], np.int32)
    for strict in (True, False):
        _run_all(pc, ps, pe, bc, bs, be, 1, strict, ks=((1, True), (3, False)))


def test_empty_inputs():
    eng = _engine()
    e = np.zeros(0, np.int32)
    ix = eng.DeviceIndex(_dev(e), _dev(e), _dev(e), 3)
    c, s, en = synth(100, 3, 1000, 50, 1)
    assert ix.count_overlaps(_dev(c), _dev(s), _dev(en), eng.FILTER_STRICT).sum().item() == 0
    a, b = ix.overlap_pairs(_dev(c), _dev(s), _dev(en), eng.FILTER_STRICT)
    assert a.numel() == 0 and b.numel() == 0
    p, d = ix.nearest(_dev(c), _dev(s), _dev(en), eng.FILTER_STRICT)
    assert (_u32(p) == 0xFFFFFFFF).all() and (d.cpu().numpy() == -1).all()
    ix2 = eng.DeviceIndex(_dev(c), _dev(s), _dev(en), 3)
    a, b = ix2.overlap_pairs(_dev(e), _dev(e), _dev(e), eng.FILTER_WEAK)
    assert a.numel() == 0
    assert ix2.count_overlaps(_dev(e), _dev(e), _dev(e), eng.FILTER_WEAK).numel() == 0


def test_heavy_windows_skewed_output():
    # few long indexed intervals x many probes inside them: the warp-cooperative emit path (window >= 64)
    rng = np.random.default_rng(11)
    m = 4000
    bc = np.zeros(m, np.int32); bs = rng.integers(0, 10_000, m).astype(np.int32)
    be = (bs + rng.integers(5_000, 20_000, m)).astype(np.int32)
    n = 3000
    pc = np.zeros(n, np.int32); ps = rng.integers(0, 30_000, n).astype(np.int32)
    pe = (ps + rng.integers(1, 150, n)).astype(np.int32)
    cnt = _run_all(pc, ps, pe, bc, bs, be, 1, True, ks=((1, True), (5, True)))
    assert cnt.max() >= 64


def test_full_size_properties_config2():
    """BASELINE config 2 (10M reads x 1M variants, one contig) through size-independent properties:
    sum(count_overlaps) == emitted pairs; every emitted pair satisfies the predicate; pairs sorted by probe
    row; per-probe multiplicity == count; oracle agreement on a 200k-probe slice."""
    eng = _engine()
    from bench import make_config2

    (pc, ps, pe), (bc, bs, be), nc = make_config2(10_000_000, 1_000_000)
    ix = eng.DeviceIndex(_dev(bc), _dev(bs), _dev(be), nc)
    dpc, dps, dpe = _dev(pc), _dev(ps), _dev(pe)
    cnt = ix.count_overlaps(dpc, dps, dpe, eng.FILTER_STRICT)
    a, b = ix.overlap_pairs(dpc, dps, dpe, eng.FILTER_STRICT)
    assert int(cnt.sum()) == a.numel() > 5_000_000
    al, bl = a.long(), b.long()
    dbs, dbe = _dev(bs), _dev(be)
    assert bool(((dps[al] < dbe[bl]) & (dpe[al] > dbs[bl])).all())
    assert bool((al[1:] >= al[:-1]).all())
    assert torch.equal(torch.bincount(al, minlength=len(pc)), cnt)
    sl = slice(4_000_000, 4_200_000)
    oc = oracle.Index(bc, bs, be, nc).count_overlaps(pc[sl], ps[sl], pe[sl], True, threads=4)
    assert np.array_equal(cnt[sl].cpu().numpy(), oc)


def test_block_cache_reuse_across_sizes_and_streams():
    """The device block cache (pbgpu.cu) parks freed blocks and hands them to the next request of the same size
    class -- in stream order on the freeing stream, behind an event on any other.  Alternate two table sizes on the
    default stream and two side streams, freeing every index right after use: every result must still equal the
    oracle's (a block handed out too early would be overwritten while the previous job still reads it)."""
    eng = _engine()
    rng = np.random.default_rng(77)
    jobs = []
    for m, n in ((120_000, 400_000), (40_000, 150_000)):
        bc, bs, be = synth(m, 3, 5_000_000, 400, int(rng.integers(1 << 30)))
        pc, ps, pe = synth(n, 3, 5_000_000, 200, int(rng.integers(1 << 30)))
        oix = oracle.Index(bc, bs, be, 3)
        jobs.append(((pc, ps, pe), (bc, bs, be), oix.count_overlaps(pc, ps, pe, True), oix.overlap_pairs(pc, ps, pe, True)))
    streams = [torch.cuda.current_stream(), torch.cuda.Stream(), torch.cuda.Stream()]
    for it in range(12):
        (pc, ps, pe), (bc, bs, be), ocnt, (oa, ob) = jobs[it % 2]
        st = streams[it % 3]
        with torch.cuda.stream(st):
            dp, db = [_dev(x) for x in (pc, ps, pe)], [_dev(x) for x in (bc, bs, be)]
            st.wait_stream(torch.cuda.default_stream())  # the H2D copies above ran on `st` already; keep it explicit
            ix = eng.DeviceIndex(*db, 3)
            cnt = ix.count_overlaps(*dp, eng.FILTER_STRICT)
            a, b = ix.overlap_pairs(*dp, eng.FILTER_STRICT)
            st.synchronize()  # non-blocking stream: synchronise before the index goes back (pbgpu.h)
            ix.close()
        assert np.array_equal(cnt.cpu().numpy(), ocnt), it
        assert np.array_equal(_u32(a), oa) and np.array_equal(_u32(b), ob), it
