"""Tuning aid (round 2): how fast are the EXISTING provider kernels on config 3 when the probes arrive
(a) in random order (BASELINE), (b) fully sorted by (contig, start), (c) grouped into coarse coordinate bins with the
rows inside a bin still in random order (what a one-pass probe partition would deliver)?  Uses torch.sort only to
produce the orders; nothing here is product code."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from polars_bio_b200 import _native, engine  # noqa: E402

GRCH38 = np.array([248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                   133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                   58617616, 64444167, 46709983, 50818468, 156040895, 57227415], dtype=np.int64)
dev = torch.device("cuda:0")
SCALE = float(os.environ.get("PB_SCALE", "1.0"))


def gen():
    rng_r, rng_v = np.random.default_rng(3), np.random.default_rng(4)
    n, m = int(100e6 * SCALE), int(90e6 * SCALE)
    p = GRCH38 / GRCH38.sum()
    pc = rng_r.choice(24, size=n, p=p).astype(np.int32)
    ps = (rng_r.random(n) * (GRCH38[pc] - 150)).astype(np.int64).astype(np.int32)
    pe = (ps + 150).astype(np.int32)
    bc = rng_v.choice(24, size=m, p=p).astype(np.int32)
    bs = (rng_v.random(m) * (GRCH38[bc] - 200)).astype(np.int64).astype(np.int32)
    ln = np.where(rng_v.random(m) < 0.9, 1, rng_v.geometric(0.2, m) + 1).astype(np.int32)
    be = (bs + ln).astype(np.int32)
    return (pc, ps, pe), (bc, bs, be)


def timed(fn, reps=3):
    out = fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        del out
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    return out, float(np.median(ms))


def main():
    probe, build = gen()
    dp = [torch.from_numpy(x).to(dev) for x in probe]
    db = [torch.from_numpy(x).to(dev) for x in build]
    ix, build_ms = timed(lambda: engine.DeviceIndex(*db, 24))
    print(json.dumps({"build_ms": build_ms, "index_bytes": ix.nbytes}), flush=True)
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(GRCH38)[:-1]])).to(dev)
    g = off[dp[0].long()] + dp[1].long()  # global coordinate of every probe start
    orders = {"random (baseline)": None, "sorted": torch.argsort(g)}
    for bits in (6, 8, 10, 12):
        shift = 32 - bits
        orders[f"binned {1 << bits} (random inside a bin)"] = torch.argsort(g >> shift, stable=True)
    FO = engine.FILTER_STRICT
    for name, perm in orders.items():
        cols = dp if perm is None else [x[perm].contiguous() for x in dp]
        cnt, count_ms = timed(lambda: ix.count_overlaps(*cols, FO))
        total = int(cnt.sum())
        del cnt
        (a, b), ovl_ms = timed(lambda: ix.overlap_pairs(*cols, FO))
        st = _native.stage_times()
        assert a.numel() == total
        del a, b
        print(json.dumps({"order": name, "count_overlaps_ms": count_ms, "overlap_two_pass_ms": ovl_ms, "pairs": total,
                          "pass1_ms": st["count_ns"] * 1e-6, "scan_ms": st["scan_ns"] * 1e-6, "emit_ms": st["emit_ns"] * 1e-6}), flush=True)
        del cols
    ix.close()


if __name__ == "__main__":
    main()
