"""One index build + one count_overlaps + one two-pass overlap on BASELINE config 3, nothing else: the command to put
under ncu (launch list or `--set full -k regex:...`).  PB_SCALE shrinks the tables.  Not a benchmark: numbers printed
under a profiler are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import workloads as wl  # noqa: E402
from polars_bio_b200 import engine  # noqa: E402

scale = float(os.environ.get("PB_SCALE", "1.0"))
n, m = int(wl.C3_READS * scale), int(wl.C3_VARIANTS * scale)
dev = torch.device("cuda:0")
dp = [torch.from_numpy(x).to(dev) for x in wl.config3_reads(0, n, n)]
db = [torch.from_numpy(x).to(dev) for x in wl.config3_variants(0, m, m)]
torch.cuda.synchronize()
reps = int(os.environ.get("PB_REPS", "1"))
for _ in range(reps):
    ix = engine.DeviceIndex(*db, 24)
    cnt = ix.count_overlaps(*dp, engine.FILTER_STRICT)
    a, b = ix.overlap_pairs(*dp, engine.FILTER_STRICT)
    torch.cuda.synchronize()
    print("pairs", a.numel(), "count sum", int(cnt.sum()))
    del cnt, a, b
    ix.close()
