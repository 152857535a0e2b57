"""Pins the CPU oracle (oracle/) against every fixture the reference's own tests hold for the
hot path.  CPU only.  If these fail the oracle is wrong and no GPU parity claim
built on it means anything."""
import numpy as np
import pandas as pd
import pytest

import oracle
from oracle import oracle_np
from tests._golden import exons_fbrain, fixtures, sort_all, synth

FX = fixtures()
COLS = ["contig", "pos_start", "pos_end"]


def _enc(df1, df2, cols1=COLS, cols2=COLS):
    c1, c2, names = oracle.encode_contigs(df1[cols1[0]], df2[cols2[0]])
    return (c1, np.asarray(df1[cols1[1]], np.int32), np.asarray(df1[cols1[2]], np.int32),
            c2, np.asarray(df2[cols2[1]], np.int32), np.asarray(df2[cols2[2]], np.int32), len(names))


def test_overlap_golden_16_rows():
    # tests/_expected.py:10-128 via tests/test_native.py:32-52 (Weak: probe = df1, index = df2)
    fx = FX["overlap"]
    lc, ls, le, rc, rs, re, nc = _enc(fx["df1"], fx["df2"])
    a, b = oracle.Index(rc, rs, re, nc).overlap_pairs(lc, ls, le, strict=fx["zero_based"])
    assert len(a) == 16
    d1, d2 = pd.DataFrame(fx["df1"]), pd.DataFrame(fx["df2"])
    got = pd.concat([d1.iloc[a].reset_index(drop=True).add_suffix("_1"),
                     d2.iloc[b].reset_index(drop=True).add_suffix("_2")], axis=1)
    pd.testing.assert_frame_equal(sort_all(got), sort_all(pd.DataFrame(fx["expected"])))
    # numpy twin and brute force agree pair-for-pair
    a2, b2 = oracle_np.overlap_pairs(lc, ls, le, rc, rs, re, strict=False)
    a3, b3 = oracle.brute_pairs(lc, ls, le, rc, rs, re, strict=False)
    key = lambda x, y: sorted(zip(x.tolist(), y.tolist()))
    assert key(a, b) == key(a2, b2) == key(a3, b3)


def test_count_overlaps_golden():
    # tests/_expected.py:183-202: counts [2,2,2,1,1,2,2,2,1,1,0]; iterated = df1, indexed = df2
    fx = FX["count_overlaps"]
    lc, ls, le, rc, rs, re, nc = _enc(fx["df1"], fx["df2"])
    cnt = oracle.Index(rc, rs, re, nc).count_overlaps(lc, ls, le, strict=False)
    got = pd.DataFrame(fx["df1"]).assign(count=cnt)
    pd.testing.assert_frame_equal(sort_all(got), sort_all(pd.DataFrame(fx["expected"])))
    assert np.array_equal(cnt, oracle_np.count_overlaps(lc, ls, le, rc, rs, re, strict=False))


def test_nearest_golden_distance_34():
    # tests/_expected.py:130-172; the oracle's tie-break reproduces the partner columns too
    fx = FX["nearest"]
    lc, ls, le, rc, rs, re, nc = _enc(fx["df1"], fx["df2"])
    ob, od = oracle.Index(rc, rs, re, nc).nearest(lc, ls, le, strict=False, k=1)
    assert (ob[:, 0] != 0xFFFFFFFF).all()
    d1, d2 = pd.DataFrame(fx["df1"]), pd.DataFrame(fx["df2"])
    got = pd.concat([d1.add_suffix("_1"), d2.iloc[ob[:, 0]].reset_index(drop=True).add_suffix("_2")], axis=1)
    got["distance"] = od[:, 0]
    pd.testing.assert_frame_equal(sort_all(got), sort_all(pd.DataFrame(fx["expected"])))
    assert 34 in od[:, 0]
    ob2, od2 = oracle_np.brute_nearest(lc, ls, le, rc, rs, re, strict=False, k=1)
    assert np.array_equal(ob, ob2) and np.array_equal(od, od2)


@pytest.mark.parametrize("kat", FX["overlap_kats"])
def test_overlap_boundary_kats(kat):
    # tests/test_coordinate_system_metadata.py:735-818
    z = np.zeros(1, np.int32)
    a, _ = oracle.Index(z, [kat["b"][0]], [kat["b"][1]], 1).overlap_pairs(
        z, [kat["a"][0]], [kat["a"][1]], strict=kat["zero_based"])
    assert len(a) == kat["rows"]


@pytest.mark.parametrize("kat", FX["count_kats"])
def test_count_boundary_kats(kat):
    # tests/test_coordinate_system_metadata.py:1172-1190; pb.count_overlaps(df1, df2) counts df2 rows per df1 row
    z = np.zeros(1, np.int32)
    cnt = oracle.Index(z, [kat["b"][0]], [kat["b"][1]], 1).count_overlaps(
        z, [kat["a"][0]], [kat["a"][1]], strict=kat["zero_based"])
    assert cnt[0] == kat["count"]


@pytest.mark.parametrize("kat", FX["coverage_kats"])
def test_coverage_boundary_kats(kat):
    # tests/test_coordinate_system_metadata.py:1577-1623
    z = np.zeros(1, np.int32)
    cov = oracle.Index(z, [kat["b"][0]], [kat["b"][1]], 1).coverage(
        z, [kat["a"][0]], [kat["a"][1]], strict=kat["zero_based"])
    assert cov[0] == kat["coverage"]


def test_output_mode_frames():
    # tests/test_overlap_output_mode.py:99-152: Left keeps multiplicity, LeftDistinct dedups by row identity
    fx = FX["output_mode"]
    cols = ["chrom", "start", "end"]
    lc, ls, le, rc, rs, re, nc = _enc(fx["left"], fx["right"], cols, cols)
    a, _ = oracle.Index(rc, rs, re, nc).overlap_pairs(lc, ls, le, strict=True)
    left = pd.DataFrame(fx["left"])
    by = ["chrom", "start", "end", "name"]
    mult = left.iloc[a].sort_values(by).reset_index(drop=True)
    pd.testing.assert_frame_equal(mult, pd.DataFrame(fx["expected_left_multiplicity"]).sort_values(by).reset_index(drop=True))
    dist = left.iloc[np.unique(a)].sort_values(by).reset_index(drop=True)
    pd.testing.assert_frame_equal(dist, pd.DataFrame(fx["expected_left_distinct"]).sort_values(by).reset_index(drop=True))


def test_exons_fbrain_published_pair_count():
    # docs/supplement.md:108,149: 54,246 pairs (0-based); 54,343 in Weak mode
    z = exons_fbrain()
    nc = len(z["contigs"])
    ix = oracle.Index(z["fbrain_chrom"], z["fbrain_start"], z["fbrain_end"], nc)
    ec, es, ee = z["exons_chrom"].astype(np.int32), z["exons_start"], z["exons_end"]
    assert ix.overlap_total(ec, es, ee, strict=True) == FX["exons_fbrain_pairs"]["strict"]
    assert ix.overlap_total(ec, es, ee, strict=False) == FX["exons_fbrain_pairs"]["weak"]
    cnt = ix.count_overlaps(ec, es, ee, strict=True)
    assert cnt.sum() == 54246 and cnt.max() == 7 and np.count_nonzero(cnt) == 51432
    assert np.array_equal(cnt, oracle_np.count_overlaps(ec, es, ee, z["fbrain_chrom"], z["fbrain_start"],
                                                        z["fbrain_end"], strict=True))
    # symmetric direction: fBrain rows with >=1 hit, max 64
    ix2 = oracle.Index(ec, es, ee, nc)
    cnt2 = ix2.count_overlaps(z["fbrain_chrom"].astype(np.int32), z["fbrain_start"], z["fbrain_end"], strict=True)
    assert cnt2.sum() == 54246 and cnt2.max() == 64 and np.count_nonzero(cnt2) == 24189
    # full pair sets agree between the two restatements
    a, b = ix.overlap_pairs(ec, es, ee, strict=True)
    a2, b2 = oracle_np.overlap_pairs(ec, es, ee, z["fbrain_chrom"], z["fbrain_start"], z["fbrain_end"], strict=True)
    o = np.lexsort((b, a))
    assert np.array_equal(a[o], a2) and np.array_equal(b[o], b2)


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tree_vs_brute_random(strict, seed):
    # includes zero-length intervals and contigs present on one side only
    lc, ls, le = synth(300, 5, 2000, 120, seed, zero_len_frac=0.1)
    rc, rs, re = synth(200, 4, 2000, 400, 100 + seed, zero_len_frac=0.1)
    ix = oracle.Index(rc, rs, re, 5)
    a, b = ix.overlap_pairs(lc, ls, le, strict)
    a3, b3 = oracle.brute_pairs(lc, ls, le, rc, rs, re, strict)
    assert sorted(zip(a.tolist(), b.tolist())) == sorted(zip(a3.tolist(), b3.tolist()))
    a2, b2 = oracle_np.overlap_pairs(lc, ls, le, rc, rs, re, strict)
    assert sorted(zip(a.tolist(), b.tolist())) == list(zip(a2.tolist(), b2.tolist()))
    assert np.array_equal(ix.count_overlaps(lc, ls, le, strict), np.bincount(a3, minlength=300))
    assert np.array_equal(ix.count_overlaps(lc, ls, le, strict), oracle_np.count_overlaps(lc, ls, le, rc, rs, re, strict))
    assert np.array_equal(ix.coverage(lc, ls, le, strict), oracle_np.brute_coverage(lc, ls, le, rc, rs, re, strict))
    for k, inc in [(1, True), (3, True), (1, False), (4, False)]:
        ob, od = ix.nearest(lc, ls, le, strict, k=k, include_overlaps=inc)
        ob2, od2 = oracle_np.brute_nearest(lc, ls, le, rc, rs, re, strict, k=k, include_overlaps=inc)
        assert np.array_equal(od, od2), (k, inc)
        assert np.array_equal(ob, ob2), (k, inc)


def test_threads_do_not_change_results():
    lc, ls, le = synth(20000, 3, 100000, 300, 7)
    rc, rs, re = synth(5000, 3, 100000, 1000, 8)
    ix = oracle.Index(rc, rs, re, 3)
    a1, b1 = ix.overlap_pairs(lc, ls, le, True, threads=1)
    a4, b4 = ix.overlap_pairs(lc, ls, le, True, threads=4)
    assert np.array_equal(a1, a4) and np.array_equal(b1, b4)
