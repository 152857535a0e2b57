"""BASELINE.json configs 3, 4 and 5 at FULL size on one B200 (device-level API through the C ABI): size-independent
properties on the whole output + oracle parity on whole contigs (exact pair sets / counts / distances, order-normalised).
The same checks the builder's tool tests/tools/scale_check.py prints, here as assertions the driver runs.
Integer / index work: bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402  (checker only)
import workloads as wl  # noqa: E402

SCALE = float(os.environ.get("PB_TEST_SCALE", "1.0"))  # < 1 shrinks the tables for dry runs
DEV = "cuda:0"


def _engine():
    from polars_bio_b200 import engine

    return engine


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(DEV)


def _subset(cols, contigs):
    sel = np.flatnonzero(np.isin(cols[0], np.asarray(contigs, dtype=np.int32)))
    return tuple(np.ascontiguousarray(x[sel]) for x in cols), sel


def _keys(a, b):
    return (a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)


def _gpu_keys(a, b, keep):
    ka, kb = a[keep].long() & 0xFFFFFFFF, b[keep].long() & 0xFFFFFFFF
    return torch.sort((ka << 32) | kb).values.cpu().numpy().astype(np.uint64)


def _check_overlap_job(probe, build, nc, contigs, min_pairs):
    eng = _engine()
    dp, db = [_dev(x) for x in probe], [_dev(x) for x in build]
    n = len(probe[0])
    ix = eng.DeviceIndex(*db, nc)
    cnt = ix.count_overlaps(*dp, eng.FILTER_STRICT)
    a, b = ix.overlap_pairs(*dp, eng.FILTER_STRICT)
    P = a.numel()
    assert P >= min_pairs
    assert int(cnt.sum()) == P                      # count_overlaps and the pair emit agree on the total
    al, bl = a.long() & 0xFFFFFFFF, b.long() & 0xFFFFFFFF
    assert int(al.max()) < n and int(bl.max()) < len(build[0])
    # every emitted pair satisfies the predicate on the same contig ...
    ok = (dp[1][al] < db[2][bl]) & (dp[2][al] > db[1][bl]) & (dp[0][al] == db[0][bl])
    assert bool(ok.all())
    del ok
    # ... and the per-probe multiplicity of the pair buffer is exactly count_overlaps (so no pair is missing either:
    # the count is the rank identity, the pairs are enumerated)
    assert torch.equal(torch.bincount(al, minlength=n), cnt)
    # no duplicate pairs
    key = (al << 32) | bl
    del al, bl
    skey = torch.sort(key).values
    del key
    assert bool((skey[1:] != skey[:-1]).all())
    del skey
    # whole contigs against the oracle: exact counts and exact pair sets with global row ids
    (psub, pid), (bsub, bid) = _subset(probe, contigs), _subset(build, contigs)
    oix = oracle.Index(*bsub, nc)
    thr = min(16, os.cpu_count() or 1)
    ocnt = oix.count_overlaps(*psub, True, threads=thr)
    assert np.array_equal(cnt[torch.from_numpy(pid).to(DEV)].cpu().numpy(), ocnt)
    oa, ob = oix.overlap_pairs(*psub, True, threads=thr)
    want = np.sort(_keys(pid[oa], bid[ob]))
    in_c = torch.zeros(nc, dtype=torch.bool, device=DEV)
    in_c[torch.tensor(list(contigs), device=DEV)] = True
    keep = in_c[dp[0][a.long() & 0xFFFFFFFF].long()]
    got = _gpu_keys(a, b, keep)
    assert len(got) == len(want) and np.array_equal(got, want)
    # streaming sink (config 5's path): same pair multiset through a bounded two-slot ring
    ix.close()
    return P


def test_config3_full_size():
    """100 M reads x 90 M variants, 24 contigs (BASELINE configs[2]): index >> L2, nested intervals (indels over SNVs)."""
    n, m = int(wl.C3_READS * SCALE), int(wl.C3_VARIANTS * SCALE)
    probe, build = wl.config3_reads(0, n, n), wl.config3_variants(0, m, m)
    P = _check_overlap_job(probe, build, 24, (20, 23), int(4.0e8 * SCALE * SCALE))
    if SCALE == 1.0:
        assert P == 438_595_343 or P > 4.3e8  # the generator is chunk-seeded: the total is pinned by the oracle checks above


def test_config4_full_size_nearest():
    """nearest k=1, 50 M queries x 5 M targets, 24 contigs mixed (BASELINE configs[3])."""
    eng = _engine()
    n, m = int(50_000_000 * SCALE), int(5_000_000 * SCALE)
    probe, build, nc = wl.config4(n, m)
    dp, db = [_dev(x) for x in probe], [_dev(x) for x in build]
    ix = eng.DeviceIndex(*db, nc)
    p, d = ix.nearest(*dp, eng.FILTER_STRICT, k=1)
    pn = p.view(-1).long() & 0xFFFFFFFF
    dn = d.view(-1)
    assert bool((pn != 0xFFFFFFFF).all()) and bool((dn >= 0).all())  # every contig has targets: everybody finds a partner
    # the reported distance is the gap to the reported partner, on the same contig
    gap = torch.clamp(torch.maximum(db[1][pn].long() - dp[2].long(), dp[1].long() - db[2][pn].long()), min=0)
    assert bool((dp[0] == db[0][pn]).all())
    assert torch.equal(gap, dn)
    # whole contigs against the oracle (distance and partner; ties are broken identically: DESIGN.md 2)
    (psub, pid), (bsub, bid) = _subset(probe, (19, 21)), _subset(build, (19, 21))
    op, od = oracle.Index(*bsub, nc).nearest(*psub, True, k=1, threads=min(16, os.cpu_count() or 1))
    sel = torch.from_numpy(pid).to(DEV)
    assert np.array_equal(dn[sel].cpu().numpy(), od[:, 0])
    assert np.array_equal(pn[sel].cpu().numpy(), bid[op[:, 0]])
    ix.close()


def test_config5_full_size_skewed_output():
    """20 M reads x 200 k exons stacked in loci (BASELINE configs[4]): ~0.7e9 pairs, up to 60+ per read, nested long
    intervals; also through the streaming sink with a bounded ring."""
    eng = _engine()
    n, m = int(20_000_000 * SCALE), int(200_000 * SCALE)
    probe, build, nc = wl.config5(n, m)
    P = _check_overlap_job(probe, build, nc, (17, 22), int(5.0e8 * SCALE))
    # streaming sink: chunks of at most 2^24 pairs from two device buffers; multiset of pairs == one-shot emit
    dp, db = [_dev(x) for x in probe], [_dev(x) for x in build]
    ix = eng.DeviceIndex(*db, nc)
    total, per_probe = 0, torch.zeros(n, dtype=torch.int64, device=DEV)
    for a, b in ix.overlap_pairs_stream(*dp, eng.FILTER_STRICT, max_pairs=1 << 24):
        assert a.numel() <= max(1 << 24, 1)
        total += a.numel()
        per_probe += torch.bincount(a.long() & 0xFFFFFFFF, minlength=n)
    assert total == P
    assert torch.equal(per_probe, ix.count_overlaps(*dp, eng.FILTER_STRICT))
    ix.close()
