"""Explicit id columns (pbgpu_index_build_ids / pbgpu_overlap_count_ids): pairs and nearest partners must come out as
ids[row] on every pass-2 path -- flat expansion, staged walk, generic kernels (inverted rows), partitioned probes
(run once more with PBGPU_BIN=1).  Prints IDS_CHECK_OK."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402 (checker only)
from polars_bio_b200 import engine  # noqa: E402
from tests._golden import synth  # noqa: E402

dev = torch.device("cuda:0")
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rng = np.random.default_rng(5)


def keys(a, b):
    return np.sort((a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64))


def case(tag, bc, bs, be, pc, ps, pe, nc):
    bid = rng.permutation(len(bc)).astype(np.uint32) + np.uint32(3_000_000_000)   # ids beyond 2^31: unsigned all the way
    pid = rng.permutation(len(pc)).astype(np.uint32) + np.uint32(1_000_000)
    ix = engine.DeviceIndex(d(bc), d(bs), d(be), nc, row_ids=d(bid.view(np.int32)))
    oix = oracle.Index(bc, bs, be, nc)
    for strict in (True, False):
        fo = engine.FILTER_STRICT if strict else engine.FILTER_WEAK
        a, b = ix.overlap_pairs(d(pc), d(ps), d(pe), fo, probe_ids=d(pid.view(np.int32)))
        oa, ob = oix.overlap_pairs(pc, ps, pe, strict)
        assert np.array_equal(keys(a.cpu().numpy().view(np.uint32), b.cpu().numpy().view(np.uint32)), keys(pid[oa], bid[ob])), (tag, strict)
        parts = [(x.cpu().numpy().view(np.uint32).copy(), y.cpu().numpy().view(np.uint32).copy())
                 for x, y in ix.overlap_pairs_stream(d(pc), d(ps), d(pe), fo, max_pairs=5000)]
        if parts:  # the streaming sink has no id column for the probes: rows there, ids for the indexed side
            sa, sb = np.concatenate([x for x, _ in parts]), np.concatenate([y for _, y in parts])
            assert np.array_equal(keys(sa, sb), keys(oa, bid[ob])), (tag, strict, "stream")
        p, _ = ix.nearest(d(pc), d(ps), d(pe), fo, k=1)
        op, _ = oix.nearest(pc, ps, pe, strict, k=1)
        want = np.where(op[:, 0] == 0xFFFFFFFF, np.uint32(0xFFFFFFFF), bid[np.minimum(op[:, 0], len(bid) - 1)])
        assert np.array_equal(p.cpu().numpy().view(np.uint32)[:, 0], want), (tag, strict, "nearest")
    ix.close()
    print("ok", tag, flush=True)


bc, bs, be = synth(40_000, 3, 2_000_000, 3000, 1)          # nested: staged walk
pc, ps, pe = synth(90_000, 4, 2_000_000, 300, 2, zero_len_frac=0.05)
pc[::97] = -1
case("nested", bc, bs, be, pc, ps, pe, 4)
bs2 = np.sort(rng.integers(0, 5_000_000, 50_000)).astype(np.int32)  # no nesting: flat expansion
case("flat", np.zeros(50_000, np.int32), bs2, bs2 + 1, np.zeros(80_000, np.int32), *synth(80_000, 1, 5_000_000, 150, 3)[1:], 1)
bs3, be3 = bs.copy(), be.copy()
bs3[::50], be3[::50] = be[::50], bs[::50]                      # inverted rows: generic kernels
case("inverted", bc, bs3, be3, pc, ps, pe, 4)
print("IDS_CHECK_OK")
