"""Full-size runs of BASELINE.json configs 3, 4 and 5 on ONE B200 (device-level API), with size-independent
property checks and oracle parity on a probe sample.  Prints one JSON line per config.

  config 3: 100M reads (150 bp) x 90M variants (90% SNV, 10% indels len~Geom(0.2)+1), 24 contigs ~ GRCh38 lengths
  config 4: nearest k=1, 50M queries (150 bp) x 5M targets (len ~ LogNormal(5.5,1) clipped [1,100k]), 24 contigs
  config 5: 20M reads concentrated on 200k exons (empirical-like lengths: median ~130, heavy tail to 91k) -> ~1e9 pairs
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402 (checker only)
import workloads as wl  # noqa: E402
from polars_bio_b200 import _native, engine  # noqa: E402

GRCH38 = np.array([248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                   133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                   58617616, 64444167, 46709983, 50818468, 156040895, 57227415], dtype=np.int64)
dev = torch.device("cuda:0")
SCALE = float(os.environ.get("PB_SCALE", "1.0"))  # shrink for dry runs


def contigs_by_length(rng, n):
    return rng.choice(24, size=n, p=GRCH38 / GRCH38.sum()).astype(np.int32)


def uniform_on(rng, c, width):
    return (rng.random(len(c)) * (GRCH38[c] - width)).astype(np.int64).astype(np.int32)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def timed(fn, reps=int(os.environ.get("PB_REPS", "3"))):
    out = fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for _ in range(reps):
        del out
        ev[0].record(); out = fn(); ev[1].record(); ev[1].synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
    return out, float(np.median(ms))


def sample_parity(name, pc, ps, pe, bc, bs, be, strict, counts=None, pairs=None, near=None, sample=100_000):
    rng = np.random.default_rng(99)
    idx = np.sort(rng.choice(len(pc), size=min(sample, len(pc)), replace=False))
    t0 = time.time()
    oix = oracle.Index(bc, bs, be, 24)
    thr = os.cpu_count() or 1
    res = {"oracle_index_s": round(time.time() - t0, 1)}
    if counts is not None:
        oc = oix.count_overlaps(pc[idx], ps[idx], pe[idx], strict, threads=thr)
        res["count_sample_equal"] = bool(np.array_equal(counts[idx], oc))
    if pairs is not None:  # pairs of the sampled probes (selected on the device: the pair buffers are GBs), compared as sets
        a, b = pairs  # device int32 tensors; pairs of one probe are contiguous, probes in bin order when the index is beyond the L2
        order = torch.argsort(a, stable=True)
        a, b = a[order], b[order]
        del order
        idx_d = torch.from_numpy(idx.astype(np.int32)).to(a.device)
        lo_ = torch.searchsorted(a, idx_d, right=False); hi_ = torch.searchsorted(a, idx_d, right=True)
        w = hi_ - lo_
        pos = torch.repeat_interleave(lo_ - (torch.cumsum(w, 0) - w), w) + torch.arange(int(w.sum()), device=a.device)
        ga, gb = a[pos].cpu().numpy().astype(np.int64), b[pos].cpu().numpy().astype(np.int64)
        oa, ob = oix.overlap_pairs(pc[idx], ps[idx], pe[idx], strict, threads=thr)
        got = np.sort(ga * len(bc) + gb)
        want = np.sort(idx[oa].astype(np.int64) * len(bc) + ob)
        res["pairs_sample_equal"] = bool(len(got) == len(want) and np.array_equal(got, want))
        res["pairs_in_sample"] = int(len(want))
    if near is not None:
        p, d = near
        op, od = oix.nearest(pc[idx], ps[idx], pe[idx], strict, k=1, threads=thr)
        res["nearest_distance_sample_equal"] = bool(np.array_equal(d[idx], od[:, 0]))
        res["nearest_partner_sample_equal"] = bool(np.array_equal(p[idx], op[:, 0]))
    return res


def config3():
    n, m = int(100e6 * SCALE), int(90e6 * SCALE)
    pc, ps, pe = wl.config3_reads(0, n, n)
    bc, bs, be = wl.config3_variants(0, m, m)
    dp, db = [t(x) for x in (pc, ps, pe)], [t(x) for x in (bc, bs, be)]
    ix, build_ms = timed(lambda: engine.DeviceIndex(*db, 24))
    cnt, count_ms = timed(lambda: ix.count_overlaps(*dp, engine.FILTER_STRICT))
    (a, b), ovl_ms = timed(lambda: ix.overlap_pairs(*dp, engine.FILTER_STRICT))
    P = a.numel()
    props = {"sum_count_eq_pairs": int(cnt.sum()) == P, "sorted_by_probe": bool((a[1:].long() >= a[:-1].long()).all())}
    al, bl = a.long(), b.long()
    props["predicate_holds"] = bool(((dp[1][al] < db[2][bl]) & (dp[2][al] > db[1][bl]) & (dp[0][al] == db[0][bl])).all())
    del al, bl
    par = sample_parity("c3", pc, ps, pe, bc, bs, be, True, counts=cnt.cpu().numpy(),
                        pairs=(a, b))
    step_ms = build_ms + count_ms + ovl_ms
    algo = 12.0 * (n + m) + 8.0 * P
    return {"config": "3: 100M reads x 90M variants, 24 contigs, 1 GPU", "n": n, "m": m, "pairs": P, "index_bytes": ix.nbytes,
            "build_ms": build_ms, "count_overlaps_ms": count_ms, "overlap_two_pass_ms": ovl_ms,
            "pairs_per_s_step": P / (step_ms * 1e-3), "overlap_GBps_algorithmic": algo / (ovl_ms * 1e-3) / 1e9, **props, **par}


def config4():
    rng_q, rng_t = np.random.default_rng(5), np.random.default_rng(6)
    n, m = int(50e6 * SCALE), int(5e6 * SCALE)
    pc = contigs_by_length(rng_q, n); ps = uniform_on(rng_q, pc, 150); pe = (ps + 150).astype(np.int32)
    bc = contigs_by_length(rng_t, m)
    ln = np.clip(rng_t.lognormal(5.5, 1.0, m), 1, 100_000).astype(np.int32)
    bs = uniform_on(rng_t, bc, 100_001); be = (bs + ln).astype(np.int32)
    dp, db = [t(x) for x in (pc, ps, pe)], [t(x) for x in (bc, bs, be)]
    ix, build_ms = timed(lambda: engine.DeviceIndex(*db, 24))
    (p, d), near_ms = timed(lambda: ix.nearest(*dp, engine.FILTER_STRICT, k=1))
    pn, dn = p.cpu().numpy().view(np.uint32)[:, 0], d.cpu().numpy()[:, 0]
    props = {"all_have_partner": bool((pn != 0xFFFFFFFF).all()), "distance_nonneg": bool((dn >= 0).all())}
    par = sample_parity("c4", pc, ps, pe, bc, bs, be, True, near=(pn, dn))
    algo = 12.0 * (n + m) + 12.0 * n
    return {"config": "4: nearest k=1, 50M queries x 5M targets, 24 contigs, 1 GPU", "n": n, "m": m, "build_ms": build_ms,
            "nearest_ms": near_ms, "queries_per_s": n / ((build_ms + near_ms) * 1e-3),
            "nearest_GBps_algorithmic": algo / (near_ms * 1e-3) / 1e9, **props, **par}


def config5():
    rng_r, rng_e = np.random.default_rng(7), np.random.default_rng(8)
    n, m = int(20e6 * SCALE), int(200e3 * SCALE)
    # exons stacked in gene-like loci (about 50 isoform exons per locus, starts jittered by < 120 bp) so that a read
    # on a locus overlaps most of them: ~1e9 pairs in total, windows of 50+ candidates, nested (long) intervals
    n_loci = max(1, m // 50)
    lc = contigs_by_length(rng_e, n_loci)
    lpos = uniform_on(rng_e, lc, 200_000)
    which = np.arange(m) % n_loci
    bc = lc[which]
    ln = np.clip(rng_e.lognormal(4.9, 1.1, m), 10, 91_671).astype(np.int32)  # median ~134, tail to 91,671 (tests/data/exons-like)
    bs = (lpos[which] + rng_e.integers(0, 120, m)).astype(np.int32)
    be = (bs + ln).astype(np.int32)
    on = rng_r.random(n) < 0.8
    ex = rng_r.integers(0, m, n)
    pc = np.where(on, bc[ex], contigs_by_length(rng_r, n)).astype(np.int32)
    ps = np.where(on, bs[ex] + rng_r.integers(-75, 76, n), uniform_on(rng_r, pc, 150)).astype(np.int32)
    ps = np.maximum(ps, 0); pe = (ps + 150).astype(np.int32)
    dp, db = [t(x) for x in (pc, ps, pe)], [t(x) for x in (bc, bs, be)]
    ix, build_ms = timed(lambda: engine.DeviceIndex(*db, 24))
    cnt, count_ms = timed(lambda: ix.count_overlaps(*dp, engine.FILTER_STRICT))
    (a, b), ovl_ms = timed(lambda: ix.overlap_pairs(*dp, engine.FILTER_STRICT), reps=2)
    P = a.numel()
    props = {"sum_count_eq_pairs": int(cnt.sum()) == P, "max_pairs_per_read": int(cnt.max()),
             "peak_hbm_bytes": int(torch.cuda.max_memory_allocated())}
    par = sample_parity("c5", pc, ps, pe, bc, bs, be, True, counts=cnt.cpu().numpy(),
                        pairs=(a, b), sample=20_000)
    algo = 12.0 * (n + m) + 8.0 * P
    return {"config": "5: skewed output, 20M reads x 200k exons, 1 GPU, pairs materialised whole", "n": n, "m": m, "pairs": P,
            "build_ms": build_ms, "count_overlaps_ms": count_ms, "overlap_two_pass_ms": ovl_ms,
            "pairs_per_s_step": P / ((build_ms + count_ms + ovl_ms) * 1e-3),
            "overlap_GBps_algorithmic": algo / (ovl_ms * 1e-3) / 1e9, **props, **par}


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "4", "5"]
    for w in which:
        torch.cuda.reset_peak_memory_stats()
        r = {"3": config3, "4": config4, "5": config5}[w]()
        r["launches_total"] = _native.launch_count()
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
