"""Parity of the probe-partition path (csrc/bins.cuh) on small seeded inputs: run with PBGPU_BIN=1 so that every
fast-path call partitions its probes (by default only indexes far beyond the L2 do).  Counts must equal the oracle's row
for row; pairs as a set (bin order replaces probe order); the streaming sink must deliver the same multiset.
Prints BINS_CHECK_OK.  Driven by tests/test_gpu_bins.py (subprocess: the switch is read once per process)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402 (checker only)
from polars_bio_b200 import engine  # noqa: E402
from tests._golden import exons_fbrain, synth  # noqa: E402

assert os.environ.get("PBGPU_BIN") == "1", "run with PBGPU_BIN=1"
dev = torch.device("cuda:0")


def d(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


def keys(a, b):
    return np.sort((a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64))


def check(pc, ps, pe, bc, bs, be, nc, tag):
    ix = engine.DeviceIndex(d(bc), d(bs), d(be), nc)
    oix = oracle.Index(bc, bs, be, nc)
    dp = [d(x) for x in (pc, ps, pe)]
    for strict in (True, False):
        fo = engine.FILTER_STRICT if strict else engine.FILTER_WEAK
        cnt = ix.count_overlaps(*dp, fo).cpu().numpy()
        ocnt = oix.count_overlaps(pc, ps, pe, strict)
        assert np.array_equal(cnt, ocnt), (tag, strict, "count")
        a, b = ix.overlap_pairs(*dp, fo)
        oa, ob = oix.overlap_pairs(pc, ps, pe, strict)
        ga, gb = a.cpu().numpy().view(np.uint32), b.cpu().numpy().view(np.uint32)
        assert len(ga) == len(oa), (tag, strict, len(ga), len(oa))
        assert np.array_equal(keys(ga, gb), keys(oa, ob)), (tag, strict, "pairs")
        # within one probe the partners keep (start, row) order
        if len(ga):
            o = np.argsort(ga, kind="stable")
            assert np.array_equal(gb[o], ob[np.argsort(oa, kind="stable")]), (tag, strict, "partner order")
        for cap in (1, 1000, 1 << 16):
            parts = [(x.cpu().numpy().view(np.uint32).copy(), y.cpu().numpy().view(np.uint32).copy())
                     for x, y in ix.overlap_pairs_stream(*dp, fo, max_pairs=cap)]
            sa = np.concatenate([x for x, _ in parts]) if parts else np.zeros(0, np.uint32)
            sb = np.concatenate([y for _, y in parts]) if parts else np.zeros(0, np.uint32)
            assert np.array_equal(sa, ga) and np.array_equal(sb, gb), (tag, strict, "stream", cap)
        cov = ix.coverage(*dp, fo).cpu().numpy()
        assert np.array_equal(cov, oix.coverage(pc, ps, pe, strict)), (tag, strict, "coverage")
        p, dist = ix.nearest(*dp, fo, k=2)
        op, od = oix.nearest(pc, ps, pe, strict, k=2)
        assert np.array_equal(dist.cpu().numpy(), od) and np.array_equal(p.cpu().numpy().view(np.uint32), op), (tag, strict, "nearest")
    ix.close()
    print("ok", tag, len(pc), len(bc), flush=True)


rng = np.random.default_rng(2024)
# nested intervals, several contigs, probes of every kind
for n, m, nc, span, blen, plen, zf in ((1, 5, 1, 1000, 50, 40, 0.0), (31, 200, 2, 5000, 300, 100, 0.0), (4097, 3000, 3, 200_000, 5000, 300, 0.1),
                                       (200_003, 150_000, 5, 3_000_000, 2000, 400, 0.05), (50_000, 40, 2, 100_000, 60_000, 200, 0.0),
                                       (300_000, 300_000, 24, 50_000_000, 60, 150, 0.01)):
    bc, bs, be = synth(m, nc, span, blen, int(rng.integers(1 << 30)))
    pc, ps, pe = synth(n, nc + 1, span, plen, int(rng.integers(1 << 30)), zero_len_frac=zf)  # contig nc: no indexed rows
    pc[rng.random(n) < 0.02] = -1  # null keys
    check(pc, ps, pe, bc, bs, be, nc + 1, f"synth n={n} m={m}")
# no nesting at all (SNV-like): flat structure, staged emit still exact
bc = np.zeros(100_000, np.int32); bs = np.sort(rng.integers(0, 10_000_000, 100_000)).astype(np.int32); be = bs + 1
pc, ps, pe = synth(250_000, 1, 10_000_000, 150, 7)
check(pc, ps, pe, bc, bs, be, 1, "snv")
# many equal ends in one bucket: crowded records with unsorted ends (linear count) and the sorted fallback (> 64)
for dup in (40, 500):
    bs = np.concatenate([rng.integers(0, 1_000_000, 5000), np.full(dup, 500_000) - rng.integers(1, 2000, dup)]).astype(np.int32)
    be = np.concatenate([bs[:5000] + rng.integers(1, 3000, 5000), np.full(dup, 500_010)]).astype(np.int32)
    bc = np.zeros(len(bs), np.int32)
    pc, ps, pe = synth(60_000, 1, 1_000_000, 500, 11 + dup)
    ps[:2000] = 500_000 + rng.integers(-30, 30, 2000); pe[:2000] = ps[:2000] + rng.integers(1, 60, 2000)
    check(pc, ps, pe, bc, bs, be, 1, f"equal ends x{dup}")
z = exons_fbrain()
check(z["fbrain_chrom"], z["fbrain_start"], z["fbrain_end"], z["exons_chrom"], z["exons_start"], z["exons_end"], len(z["contigs"]), "fbrain x exons")
check(z["exons_chrom"], z["exons_start"], z["exons_end"], z["fbrain_chrom"], z["fbrain_start"], z["fbrain_end"], len(z["contigs"]), "exons x fbrain")
print("BINS_CHECK_OK")
