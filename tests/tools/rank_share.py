"""One rank's share of BASELINE config 3 at world size W on ONE GPU: the rows of the contigs rank R owns (LPT owner
table over the global per-contig histogram, exactly dist.owner_table), joined with the single-GPU calls.  What the
N-GPU step spends after the exchange, measurable (and profilable under ncu) without N GPUs.

  PB_WORLD=8 PB_RANK=0 [PB_REPS=6] [PBGPU_TRACE_BUILD=1] python tests/tools/rank_share.py

Prints one JSON line: rows, pairs, CUDA-event ms of build / count_overlaps / overlap (median over the repetitions).
Not a benchmark line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import workloads as wl  # noqa: E402
from polars_bio_b200 import _native, dist, engine  # noqa: E402

world, rank = int(os.environ.get("PB_WORLD", "8")), int(os.environ.get("PB_RANK", "0"))
reps = int(os.environ.get("PB_REPS", "6"))
scale = float(os.environ.get("PB_SCALE", "1.0"))
n, m = int(wl.C3_READS * scale), int(wl.C3_VARIANTS * scale)
reads, variants = wl.config3_reads(0, n, n), wl.config3_variants(0, m, m)
hist = np.bincount(reads[0], minlength=24) + np.bincount(variants[0], minlength=24)
owner = dist.owner_table(torch.from_numpy(hist), world).numpy()
mine = owner == rank
keep_r, keep_v = mine[reads[0]], mine[variants[0]]
dev = torch.device("cuda:0")
dp = [torch.from_numpy(np.ascontiguousarray(x[keep_r])).to(dev) for x in reads]
db = [torch.from_numpy(np.ascontiguousarray(x[keep_v])).to(dev) for x in variants]
del reads, variants
torch.cuda.synchronize()


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    b.synchronize()
    return out, a.elapsed_time(b)


rows = []
pairs = 0
for it in range(reps):
    ix, t_build = timed(lambda: engine.DeviceIndex(*db, 24))
    st_b = _native.stage_times()
    cnt, t_count = timed(lambda: ix.count_overlaps(*dp, engine.FILTER_STRICT))
    (a, b), t_overlap = timed(lambda: ix.overlap_pairs(*dp, engine.FILTER_STRICT))
    st_o = _native.stage_times()
    pairs = a.numel()
    assert int(cnt.sum()) == pairs
    del cnt, a, b
    ix.close()
    if it:  # the first repetition warms the block cache
        rows.append((t_build, t_count, t_overlap))
med = np.median(np.array(rows), axis=0)
print(json.dumps({"world": world, "rank": rank, "contigs": [int(c) for c in np.nonzero(mine)[0]], "probe_rows": int(dp[0].numel()),
                  "indexed_rows": int(db[0].numel()), "pairs": int(pairs), "build_ms": round(float(med[0]), 4),
                  "count_overlaps_ms": round(float(med[1]), 4), "overlap_ms": round(float(med[2]), 4),
                  "step_ms": round(float(med.sum()), 4), "stage_ns_last": st_o}))
