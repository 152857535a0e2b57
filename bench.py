#!/usr/bin/env python
"""bench.py -- overlap-pairs/sec of the interval-join hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload: BASELINE.json configs[2] (the north-star target, SURVEY.md 8d "Config 3") -- 100 M WGS reads (150 bp) x 90 M
gnomAD-like variants (90 % SNV, 10 % indels), 24 contigs with GRCh38 lengths, rows of all contigs mixed, Strict
(0-based half-open).  It fits one B200, so N = 1 runs the whole job; N > 1 runs the SAME global job strong-scaled:
every rank starts with a contiguous 1/N block of the rows of both tables (arbitrary contigs), one exchange step moves
every row to the rank that owns its contig (LPT over the 24 skewed contigs), then every rank joins its contigs.
One step = one pass of the whole hot path over that batch:
    [N > 1: contig exchange]  ->  index build  ->  count_overlaps (int64 per read)
    ->  overlap pass 1 (count + offsets)  ->  overlap pass 2 (exact-sized (read, variant) pair buffer of GLOBAL row ids)
`value`   : pairs emitted per second, int32 columns already resident in HBM (CUDA events; inputs are 2.3 GB >> L2).
`e2e`     : the same job through the public host-facing call with HOST buffers: N = 1 pb.count_overlaps + pb.overlap on
            host Arrow tables (utf8 contig) with the output frames materialised batch by batch; N > 1 pinned host
            columns -> H2D -> exchange -> join -> D2H of counts and global pair ids.
`roofline`: the dominant provider stage's algorithmic bytes / its CUDA-event duration vs the measured HBM copy peak.
`cpu_baseline` / `--impl reference`: oracle/ (port of the reference's per-contig interval-tree algorithm) on the host
            cores, on a bounded sample of the same workload: the whole sub-job of a few contigs (contigs are
            independent and rows are proportional to contig length, so pairs/s of the sub-job is the job's).
`parity_check`: untimed -- the pair set and counts of one whole contig (global row ids) against the oracle.
`secondary`: BASELINE configs[1] (10 M x 1 M, one contig), device-timed, as a nested object (N = 1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import workloads as wl  # noqa: E402

METRIC = "overlap-pairs/sec"
NC = 24
CPU_SAMPLE_CONTIGS = (18, 19, 20, 21)  # chr19..chr22: 7.1 % of the genome
PARITY_CONTIG = 20                     # chr21


def make_config2(n_reads: int = 10_000_000, n_variants: int = 1_000_000, seed_shift: int = 0):
    """BASELINE configs[1] (kept under this name for the tests)."""
    return wl.config2(n_reads, n_variants, seed_shift)


def workload_string(n: int, m: int) -> str:
    """Identical in both arms (the driver compares the strings)."""
    return (f"config3: {n} reads (150 bp) x {m} variants (90% SNV, 10% indel), 24 contigs ~ GRCh38 lengths, rows in "
            "arbitrary order, Strict; index build + count_overlaps + two-pass pair emit")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel from the newest committed
    `ncu --set full` capture of THIS workload (profiles/*config3*_traffic.json, written by scripts/ncu_summary.py)."""
    import glob

    def captured(path):  # newest capture first: the time recorded in the file (scripts/ncu_summary.py), else the name
        try:
            return (json.load(open(path)).get("_meta", {}).get("captured_unix", 0), os.path.basename(path))
        except Exception:
            return (0, os.path.basename(path))

    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "*config3*_traffic.json")), key=captured, reverse=True):
        try:
            t = json.load(open(p))
            for name, v in t.items():
                if kernel_key in name:
                    return v["traffic_bytes"], "profiles/" + os.path.basename(p)
        except Exception:
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.mark = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def begin_window(self):
        self.mark = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[self.mark:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": "last warm-up step + timed steps (+ up to 0.4 s of the same step)"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample (whole contigs)
# ------------------------------------------------------------------------------------------------
def contig_subset(cols, contigs):
    """Rows of `cols` = (contig, start, end) on the given contigs; returns (columns, global row ids)."""
    c = cols[0]
    sel = np.flatnonzero(np.isin(c, np.asarray(contigs, dtype=c.dtype)))
    return tuple(np.ascontiguousarray(x[sel]) for x in cols), sel


def cpu_sample_tables(n: int, m: int):
    """The sub-job of CPU_SAMPLE_CONTIGS, generated without materialising the whole tables."""
    pr, bu = [], []
    for lo in range(0, n, 16 * wl.CHUNK):
        pr.append(contig_subset(wl.config3_reads(lo, min(n, lo + 16 * wl.CHUNK), n), CPU_SAMPLE_CONTIGS)[0])
    for lo in range(0, m, 16 * wl.CHUNK):
        bu.append(contig_subset(wl.config3_variants(lo, min(m, lo + 16 * wl.CHUNK), m), CPU_SAMPLE_CONTIGS)[0])
    probe = tuple(np.concatenate([p[k] for p in pr]) for k in range(3))
    build = tuple(np.concatenate([p[k] for p in bu]) for k in range(3))
    return probe, build


def cpu_reference_run(probe, build, nc, threads: int, steps: int = 1, warmup: int = 0):
    """Times the oracle (per-contig interval-tree build + per-row query, the reference's algorithm) on host cores.
    One step = index build over the sample's variants + count_overlaps + full pair emit over the sample's reads."""
    import oracle

    times, pairs = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ix = oracle.Index(*build, nc)
        cnt = ix.count_overlaps(*probe, True, threads=threads)
        a, b = ix.overlap_pairs(*probe, True, threads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        pairs = len(a)
        assert int(cnt.sum()) == pairs
        del ix, a, b, cnt
    return pairs, float(np.mean(times))


def pick_threads(probe, build, nc):
    """All host threads the port can use: try every logical CPU and every physical core (half), keep the faster."""
    ncpu = os.cpu_count() or 1
    best, best_t = None, None
    small = tuple(x[:500_000] for x in probe)
    for thr in sorted({ncpu, max(1, ncpu // 2)}, reverse=True):
        _, sec = cpu_reference_run(small, build, nc, thr, steps=1, warmup=1)
        if best is None or sec < best:
            best, best_t = sec, thr
    return best_t


def cpu_sample_description(probe, build, n, m):
    names = ",".join(wl.CONTIG_NAMES[c] for c in CPU_SAMPLE_CONTIGS)
    return (f"whole sub-job of contigs {names} ({len(probe[0])} of {n} reads x {len(build[0])} of {m} variants): "
            "interval-tree build + count_overlaps + pair emit; pairs/s of the sub-job")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    n, m = args.reads, args.variants
    probe, build = cpu_sample_tables(n, m)
    threads = pick_threads(probe, build, NC)
    pairs, sec = cpu_reference_run(probe, build, NC, threads, steps=args.steps, warmup=args.warmup)
    v = pairs / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_string(n, m), "l2": "n/a (CPU)", "pairs_per_step": pairs,
                   "note": "the CPU arm emits index pairs only; the GPU arm's e2e also materialises the output frames"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": "per step: " + cpu_sample_description(probe, build, n, m)},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# parity check of one whole contig (untimed)
# ------------------------------------------------------------------------------------------------
def oracle_contig(probe_sub, probe_ids, build_sub, build_ids):
    """Oracle counts and pair keys (probe_global * 2^32 + build_global, sorted) of one contig's rows."""
    import oracle

    thr = max(1, min(16, os.cpu_count() or 1))
    oix = oracle.Index(*build_sub, NC)
    cnt = oix.count_overlaps(*probe_sub, True, threads=thr)
    a, b = oix.overlap_pairs(*probe_sub, True, threads=thr)
    keys = np.sort((probe_ids[a].astype(np.uint64) << np.uint64(32)) | build_ids[b].astype(np.uint64))
    return cnt, keys


def gpu_pair_keys(a, b, keep):
    """Sorted keys of the device pair buffers (int32 storage of uint32 ids) selected by the boolean mask `keep`."""
    import torch

    ka = (a[keep].long() & 0xFFFFFFFF)
    kb = (b[keep].long() & 0xFFFFFFFF)
    return torch.sort((ka << 32) | kb).values.cpu().numpy().astype(np.uint64)


# ------------------------------------------------------------------------------------------------
def measure_e2e_api(args, probe, build, expect_pairs, dev):
    """N = 1 e2e: the public, reference-facing API (pb.count_overlaps + pb.overlap -> pbgpu_range_op) on HOST Arrow
    tables: contig strings are dictionary-encoded, columns staged to pinned memory, copied H2D, joined, results copied
    D2H and the reference's output frames (df1 rows + count; all df1/df2 columns suffixed) materialised on the host
    batch by batch and consumed as a stream (the reference's own benchmarks consume with a count)."""
    import pyarrow as pa
    import pyarrow.compute as pc
    import torch

    import polars_bio_b200 as pb

    n, m = len(probe[0]), len(build[0])
    names = pa.array(wl.CONTIG_NAMES)

    def table(cols):
        c, s_, e_ = cols
        t = pa.table({"contig": pc.take(names, pa.array(c)), "pos_start": pa.array(s_), "pos_end": pa.array(e_)})
        return pb.set_coordinate_system(t, True)

    reads_t, vars_t = table(probe), table(build)
    cols = ("contig", "pos_start", "pos_end")
    split = []

    def consume(res):
        rows = 0
        for b in res.execute_stream():
            rows += b.num_rows
        return rows

    def step_api():
        t_a = time.perf_counter()
        rc = consume(pb.count_overlaps(reads_t, vars_t, cols1=cols, cols2=cols, output_type="datafusion.DataFrame"))
        t_b = time.perf_counter()
        ro = consume(pb.overlap(reads_t, vars_t, cols1=cols, cols2=cols, output_type="datafusion.DataFrame"))
        t_c = time.perf_counter()
        split.append((t_b - t_a, t_c - t_b))
        return rc, ro

    step_api()
    torch.cuda.synchronize()
    e2e_steps = max(2, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rows_c, rows_o = step_api()
    torch.cuda.synchronize()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    assert rows_c == n and rows_o == expect_pairs, (rows_c, rows_o, expect_pairs)
    # the same overlap with two payload columns on each side (materialised on the device: fixed-width + utf8)
    with_payload = None
    if not args.skip_payload_e2e:
        reads_p = reads_t.append_column("read_id", pa.array(np.arange(n, dtype=np.int64))).append_column(
            "mapq", pa.array((probe[1] % 60).astype(np.uint8)))
        vars_p = vars_t.append_column("variant_id", pa.array(np.arange(m, dtype=np.int64))).append_column(
            "ref", pc.take(pa.array(["A", "C", "G", "T"]), pa.array((build[1] & 3).astype(np.int32))))
        reads_p, vars_p = pb.set_coordinate_system(reads_p, True), pb.set_coordinate_system(vars_p, True)
        tp = []
        for it in range(3):
            t_a = time.perf_counter()
            ro = consume(pb.overlap(reads_p, vars_p, cols1=cols, cols2=cols, output_type="datafusion.DataFrame"))
            tp.append(time.perf_counter() - t_a)
            assert ro == expect_pairs
        sec = float(np.mean(tp[1:]))
        with_payload = {"ms_per_call": sec * 1e3, "pairs_per_s": expect_pairs / sec,
                        "columns": "reads: read_id int64, mapq uint8; variants: variant_id int64, ref utf8 -> 10 output columns",
                        "h2d_bytes": int(9 * m + 9 * n + 9 * n + 8 * m + 5 * m + 1), "d2h_bytes": int((17 + 8 + 1 + 8 + 4 + 1) * expect_pairs),
                        "note": "pb.overlap only, 2 timed calls after 1 warm-up; payload columns gathered on the device"}
        del reads_p, vars_p
    # bytes on the bus per step, counted from what the bridge copies: both calls upload the variants (3 x int32) and
    # the reads (contig code as uint8 + 2 x int32); count_overlaps brings back uint32 counts, overlap the key columns of
    # the result rows (contig code uint8 + 4 x int32 positions; no payload columns -> no row ids)
    return {"value": expect_pairs / e2e_sec, "unit": "pairs/s", "h2d_bytes_per_step": 2 * (9 * m + 9 * n),  # both tables: contig code as a byte + two int32 positions per row
            "d2h_bytes_per_step": int(4 * n + 17 * expect_pairs), "ms_per_step": e2e_sec * 1e3,
            "api": "pb.count_overlaps + pb.overlap on host pyarrow Tables (utf8 contig); output frames materialised per batch "
                   "and consumed as a stream",
            "split_ms": dict(zip(("count_overlaps", "overlap"),
                                 (float(x) * 1e3 for x in np.mean(np.array(split[-e2e_steps:]), axis=0)))),
            "overlap_with_payload": with_payload}


def secondary_config2(dev, steps: int = 10):
    """BASELINE configs[1] device-timed (the round-1 headline), nested into the line for continuity."""
    import torch

    from polars_bio_b200 import _native, engine

    probe, build, nc = wl.config2()
    n, m = len(probe[0]), len(build[0])
    dp = [torch.from_numpy(x).to(dev) for x in probe]
    db = [torch.from_numpy(x).to(dev) for x in build]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    FO = engine.FILTER_STRICT

    def step():
        ix = engine.DeviceIndex(*db, nc)
        cnt = ix.count_overlaps(*dp, FO)
        a, b = ix.overlap_pairs(*dp, FO)
        ix.close()
        return cnt, a, b

    for _ in range(3):
        cnt, a, b = step()
    pairs = a.numel()
    assert int(cnt.sum()) == pairs
    ms, kern = [], []
    for _ in range(steps):
        del cnt, a, b
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cnt, a, b = step(); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
        kern.append(_native.stage_times())
    km = {k: float(np.mean([d[k] for d in kern])) * 1e-6 for k in kern[0]}
    peak, _ = measured_peak_gbs()
    b_count = 12.0 * (n + m) + 8.0 * n
    return {"workload": f"config2: {n} reads x {m} variants, chr1, Strict; index build + count_overlaps + two-pass pair emit",
            "ms_per_step": float(np.mean(ms)), "value": pairs / (float(np.mean(ms)) * 1e-3), "unit": "pairs/s", "pairs_per_step": pairs,
            "l2": "flushed between timed steps (256 MiB write)", "steps": steps,
            "index_build_ms": km["partition_sort_ns"], "count_overlaps_ms": km["count_overlaps_ns"], "pass1_ms": km["count_ns"],
            "pass2_ms": km["emit_ns"],
            "count_overlaps_frac_of_hbm_peak": b_count / (km["count_overlaps_ns"] * 1e-3) / 1e9 / peak if km["count_overlaps_ns"] else None}


def stage_roofline(n, m, pairs, km, peak, peak_src):
    """Per-stage algorithmic bytes over the library's CUDA-event durations; the dominant provider stage
    is the line's `roofline`.  km: stage name -> ms."""
    b_count = 12.0 * (n + m) + 8.0 * n
    b_overlap = 12.0 * (n + m) + 8.0 * pairs
    cands = {
        "count_overlaps (all kernels of the call)": (b_count, km["count_overlaps_ns"]),
        "overlap pass 1 (count; incl. the probe partition when the index exceeds L2)": (b_count, km["count_ns"]),
        "overlap pass 2 (emit)": (b_overlap, km["emit_ns"]),
    }
    cands = {k: v for k, v in cands.items() if v[1]}
    # `roofline` is a KERNEL's: the longest single provider launch.  When the probes were partitioned (index beyond the
    # L2) count_overlaps and pass 1 are four launches each (histogram, partition, count, un-binning; none longer than
    # pass 2's one kernel): they stay in all_stages with their whole-call fractions, and pass 2 is the dominant kernel
    partitioned = km.get("bin_ns", 0.0) > 0.0
    single = {k: v for k, v in cands.items() if not partitioned or "emit" in k} or cands
    dom = max(single, key=lambda k: single[k][1])
    ach = cands[dom][0] / (cands[dom][1] * 1e-3) / 1e9
    traffic, traffic_file = ncu_traffic("emit" if "emit" in dom else "count")
    two_pass_ms = km["count_ns"] + km["scan_ns"] + km["emit_ns"]
    return {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "traffic_source": (f"{traffic_file} (ncu --set full, per launch)" if traffic_file else None),
            "peak_source": peak_src, "algorithmic_bytes": cands[dom][0], "kernel_ms": cands[dom][1],
            "all_stages": {k: {"ms": v[1], "algorithmic_bytes": v[0], "GBps": v[0] / (v[1] * 1e-3) / 1e9,
                               "frac": v[0] / (v[1] * 1e-3) / 1e9 / peak} for k, v in cands.items()},
            "overlap_two_pass": {"ms": two_pass_ms, "algorithmic_bytes": b_overlap,
                                 "frac": b_overlap / (two_pass_ms * 1e-3) / 1e9 / peak if two_pass_ms else None},
            "index_build_ms": km["partition_sort_ns"], "offset_scan_ms": km["scan_ns"],
            "bin_ms": km.get("bin_ns", 0.0), "unbin_ms": km.get("unbin_ns", 0.0),
            "bin_note": "probe partition (histogram + one radix pass) of the step's LAST partitioned call; it is inside that call's stage time",
            "kernel_rule": "longest single provider launch; multi-launch stages (partitioned count_overlaps / pass 1) are listed in all_stages"}


def run_single(args, dev, local):
    import torch

    from polars_bio_b200 import _native, engine

    n, m = args.reads, args.variants
    t_gen = time.perf_counter()
    probe = wl.config3_reads(0, n, n)
    build = wl.config3_variants(0, m, m)
    t_gen = time.perf_counter() - t_gen
    h = [torch.from_numpy(x).pin_memory() for x in (*probe, *build)]
    dpc, dps, dpe, dbc, dbs, dbe = (x.to(dev, non_blocking=True) for x in h)
    torch.cuda.synchronize()
    del h
    FO = engine.FILTER_STRICT

    def step_device(ev=None):
        ix = engine.DeviceIndex(dbc, dbs, dbe, NC)
        if ev: ev[1].record()
        cnt = ix.count_overlaps(dpc, dps, dpe, FO)
        if ev: ev[2].record()
        a, b = ix.overlap_pairs(dpc, dps, dpe, FO)
        if ev: ev[3].record()
        ix.close()
        return cnt, a, b

    sampler = ClockSampler(local)
    sampler.start()
    warm = max(args.warmup, 3)
    for i in range(warm):
        if i == warm - 1:
            sampler.begin_window()
        cnt, a, b = step_device()
    pairs = a.numel()
    assert int(cnt.sum()) == pairs

    # ---- untimed parity check: one whole contig against the oracle (global row ids) ----
    parity = {"contig": wl.CONTIG_NAMES[PARITY_CONTIG]}
    if not args.skip_parity:
        (psub, pid), (bsub, bid) = contig_subset(probe, [PARITY_CONTIG]), contig_subset(build, [PARITY_CONTIG])
        ocnt, okeys = oracle_contig(psub, pid, bsub, bid)
        gcnt = cnt[torch.from_numpy(pid).to(dev)].cpu().numpy()
        keep = dpc[a.long()] == PARITY_CONTIG
        gkeys = gpu_pair_keys(a, b, keep)
        parity.update({"probe_rows": int(len(pid)), "indexed_rows": int(len(bid)), "pairs": int(len(okeys)),
                       "count_overlaps": "ok" if np.array_equal(gcnt, ocnt) else "MISMATCH",
                       "overlap_pairs": "ok" if (len(gkeys) == len(okeys) and np.array_equal(gkeys, okeys)) else "MISMATCH"})
        parity["status"] = "ok" if parity["count_overlaps"] == "ok" and parity["overlap_pairs"] == "ok" else "MISMATCH"
        del keep, gkeys, okeys
    del cnt, a, b

    torch.cuda.synchronize()
    launches0 = _native.launch_count()
    step_ms, stage_ms, kern_ns = [], [], []
    for _ in range(args.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        cnt, a, b = step_device(ev)
        ev[4].record()
        ev[4].synchronize()
        step_ms.append(ev[0].elapsed_time(ev[4]))
        stage_ms.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
        kern_ns.append(_native.stage_times())  # the library's own events around each stage of this step
        del cnt, a, b
    launches = _native.launch_count() - launches0
    t_end = time.perf_counter() + 0.4
    while len(sampler.lines) - sampler.mark < 3 and time.perf_counter() < t_end:  # keep the GPU under the same load until sampled
        step_device()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_per_step = float(np.sum(step_ms)) / args.steps
    value = pairs / (ms_per_step * 1e-3)

    stage = np.mean(np.array(stage_ms), axis=0)  # build, count_overlaps, overlap(count+scan+sync+emit)
    km = {k: float(np.mean([d[k] for d in kern_ns])) * 1e-6 for k in kern_ns[0]}  # ms
    peak, peak_src = measured_peak_gbs()
    roof = stage_roofline(n, m, pairs, km, peak, peak_src)
    roof["step_stage_ms"] = {"index_build": float(stage[0]), "count_overlaps": float(stage[1]), "overlap_two_pass": float(stage[2])}

    e2e = None
    if not args.skip_e2e:
        e2e = measure_e2e_api(args, probe, build, pairs, dev)
    secondary = None
    if not args.skip_secondary:
        del dpc, dps, dpe, dbc, dbs, dbe
        torch.cuda.empty_cache()
        secondary = secondary_config2(dev)

    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": 1, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_string(n, m), "pairs_per_step": float(pairs),
                   "l2": "inputs (2.3 GB of int32 columns) are larger than L2; no flush needed",
                   "parallelism": "single GPU", "data_generation_s": round(t_gen, 1)},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roof,
        "parity_check": parity.get("status", "skipped"),
        "parity_detail": parity,
        "secondary": secondary,
    }
    if not args.no_cpu_baseline:
        cprobe, cbuild = cpu_sample_tables(n, m)
        thr = pick_threads(cprobe, cbuild, NC)
        cp, csec = cpu_reference_run(cprobe, cbuild, NC, thr, steps=1, warmup=1)
        line["cpu_baseline"] = {"value": cp / csec, "unit": "pairs/s", "cores": thr, "kind": "port",
                                "sample": cpu_sample_description(cprobe, cbuild, n, m) + f", 1 run after 1 warm-up ({csec:.2f} s)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_sharded(args, world, rank, dev, local):
    """N > 1: the SAME global job, strong-scaled over contigs WITH the exchange step (see the module docstring)."""
    import torch
    import torch.distributed as dist

    from polars_bio_b200 import _native, dist as pbd, engine

    n, m = args.reads, args.variants
    plo, phi = wl.rank_slice(n, rank, world)
    blo, bhi = wl.rank_slice(m, rank, world)
    probe = wl.config3_reads(plo, phi, n)
    build = wl.config3_variants(blo, bhi, m)
    hp = [torch.from_numpy(x).pin_memory() for x in probe]
    hb = [torch.from_numpy(x).pin_memory() for x in build]
    dp = [x.to(dev, non_blocking=True) for x in hp]
    db = [x.to(dev, non_blocking=True) for x in hb]
    torch.cuda.synchronize()
    FO = engine.FILTER_STRICT
    overlap = os.environ.get("PBGPU_BENCH_OVERLAP", "1") != "0"

    def step(ev=None, trace=None, tables=None):
        tb, tp = tables if tables is not None else (db, dp)
        # the indexed table first: its index build overlaps the transfer of the reads (one stream per table)
        ready = [] if overlap else None
        (x, q), owner = pbd.shard_tables([tuple(tb), tuple(tp)], NC, trace=trace, ready=ready)
        qc, qs, qe, qrow = q
        xc, xs, xe, xrow = x
        main = torch.cuda.current_stream()
        if ready: main.wait_event(ready[0])
        if ev and not ready: ev[1].record()
        ix = engine.DeviceIndex(xc, xs, xe, NC, row_ids=xrow)  # global row ids travel with the rows: pairs come out global
        if ready: main.wait_event(ready[1])
        if ev and ready: ev[1].record()
        cnt = ix.count_overlaps(qc, qs, qe, FO)
        a, b = ix.overlap_pairs(qc, qs, qe, FO, probe_ids=qrow)
        ix.close()
        return cnt, a, b, qrow, qc, xc.numel()

    sampler = ClockSampler(local)
    sampler.start()
    # The first step decides, on all ranks together, whether the peer-memory exchange works on this box (a rank that
    # cannot reach a peer's flags gets an error after the bounded spin, not a hang); if it fails anywhere, every rank
    # switches to the NCCL all-to-all and the line below says so ("exchange": "nccl").
    err = None
    try:
        out = step()
        torch.cuda.synchronize()
        for ex in list(pbd._peer_cache.values()):  # signal/wait time-outs only raise a status word: look at it now
            if ex is not None:
                ex.check()
    except Exception as e:  # noqa: BLE001 -- re-raised below unless the NCCL exchange can take over
        err = e
    ok = torch.tensor([0.0 if err is not None else 1.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) == 0.0:
        if pbd.exchange_kind() != "peer":
            raise err if err is not None else RuntimeError("the first step failed on another rank")
        if rank == 0:
            sys.stderr.write(f"bench: peer-memory exchange failed in the first step ({err!r}); falling back to the NCCL all-to-all\n")
        os.environ["PBGPU_EXCHANGE"] = "nccl"
        pbd.abandon_peer_exchanges()
    warm = max(args.warmup, 3)
    for i in range(warm):
        if i == warm - 1:
            sampler.begin_window()
        cnt, a, b, qrow, qc, m_own = step()
    pairs = a.numel()
    n_own = cnt.numel()
    assert int(cnt.sum()) == pairs

    # ---- untimed parity check: one whole contig's counts and pair set (GLOBAL row ids) against the oracle -------
    parity = {"contig": wl.CONTIG_NAMES[PARITY_CONTIG]}
    if not args.skip_parity:
        (psub, pid), (bsub, bid) = contig_subset(probe, [PARITY_CONTIG]), contig_subset(build, [PARITY_CONTIG])
        mine = {"p": psub, "pid": pid + plo, "b": bsub, "bid": bid + blo}
        # this rank's share of the contig's GPU results (only the owner has any)
        keep_rows = qc == PARITY_CONTIG
        g_rows = (qrow[keep_rows].long() & 0xFFFFFFFF).cpu().numpy()
        g_cnt = cnt[keep_rows].cpu().numpy()
        keep = None
        if a.numel():
            # probe global id -> is it a row of the contig?  (rows of one contig live on one rank: test membership there)
            flag = torch.zeros(n, dtype=torch.bool, device=dev)
            flag[qrow[keep_rows].long() & 0xFFFFFFFF] = True
            keep = flag[a.long() & 0xFFFFFFFF]
            del flag
        mine["g_rows"], mine["g_cnt"] = g_rows, g_cnt
        mine["g_keys"] = gpu_pair_keys(a, b, keep) if keep is not None else np.zeros(0, np.uint64)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        if rank == 0:
            psub = tuple(np.concatenate([g["p"][k] for g in gathered]) for k in range(3))
            bsub = tuple(np.concatenate([g["b"][k] for g in gathered]) for k in range(3))
            pid = np.concatenate([g["pid"] for g in gathered]); bid = np.concatenate([g["bid"] for g in gathered])
            ocnt, okeys = oracle_contig(psub, pid, bsub, bid)
            g_rows = np.concatenate([g["g_rows"] for g in gathered]); g_cnt = np.concatenate([g["g_cnt"] for g in gathered])
            g_keys = np.sort(np.concatenate([g["g_keys"] for g in gathered]))
            o = np.argsort(g_rows)
            o2 = np.argsort(pid)
            parity.update({"probe_rows": int(len(pid)), "indexed_rows": int(len(bid)), "pairs": int(len(okeys)),
                           "count_overlaps": "ok" if (len(g_rows) == len(pid) and np.array_equal(g_rows[o], pid[o2]) and np.array_equal(g_cnt[o], ocnt[o2])) else "MISMATCH",
                           "overlap_pairs": "ok" if (len(g_keys) == len(okeys) and np.array_equal(g_keys, okeys)) else "MISMATCH"})
        del keep
        # the one-call distributed join, both strategies: pair totals must equal the step's, and the contig's pair set too
        for strat in ("shard", "replicate"):
            if strat == "replicate" and args.skip_replicate:
                continue
            a2, b2, used = pbd.overlap(tuple(dp), tuple(db), NC, FO, strategy=strat)
            tot = torch.tensor([float(a2.numel())], device=dev, dtype=torch.float64)
            dist.all_reduce(tot)
            # membership of the probe id in the contig: by the probe's contig code, looked up from this rank's slice when
            # the pair's probe row is local (replicate) or through the id flags (shard: pairs live on the owner)
            if strat == "replicate":
                loc = (a2.long() & 0xFFFFFFFF) - plo
                keep2 = dp[0][loc] == PARITY_CONTIG
            else:
                flag = torch.zeros(n, dtype=torch.bool, device=dev)
                flag[qrow[keep_rows].long() & 0xFFFFFFFF] = True
                keep2 = flag[a2.long() & 0xFFFFFFFF]
                del flag
            keys2 = gpu_pair_keys(a2, b2, keep2) if a2.numel() else np.zeros(0, np.uint64)
            gk = [None] * world if rank == 0 else None
            dist.gather_object(keys2, gk, dst=0)
            if rank == 0:
                k2 = np.sort(np.concatenate(gk))
                parity[f"dist.overlap[{strat}]"] = "ok" if (len(k2) == len(okeys) and np.array_equal(k2, okeys)) else "MISMATCH"
                parity[f"dist.overlap[{strat}] pairs"] = float(tot.item())
            del a2, b2, keep2
        if rank == 0:
            parity["status"] = "ok" if all(v == "ok" for k, v in parity.items() if k in ("count_overlaps", "overlap_pairs") or k.endswith("]")) else "MISMATCH"
    del cnt, a, b, qrow, qc

    dist.barrier(); torch.cuda.synchronize()
    launches0 = _native.launch_count()
    step_ms, xchg_ms, kern_ns = [], [], []
    for _ in range(args.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        out = step(ev)
        ev[2].record(); ev[2].synchronize()
        step_ms.append(ev[0].elapsed_time(ev[2])); xchg_ms.append(ev[0].elapsed_time(ev[1]))
        kern_ns.append(_native.stage_times())
        del out
    launches = _native.launch_count() - launches0
    xtrace = []
    step(trace=xtrace)  # one extra, untimed step with host-side laps of the exchange
    t_end = time.perf_counter() + 0.4
    while True:  # keep the GPUs under the same load until sampled; the step is collective, so the ranks decide together
        more = torch.tensor([1.0 if (len(sampler.lines) - sampler.mark < 3 and time.perf_counter() < t_end) else 0.0], device=dev)
        dist.all_reduce(more, op=dist.ReduceOp.MAX)
        if float(more.item()) == 0.0:
            break
        step()
    dist.barrier(); torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([float(np.sum(step_ms)), float(np.sum(xchg_ms))], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pt = torch.tensor([float(pairs), float(dp[0].numel()), float(db[0].numel())], device=dev, dtype=torch.float64)
    dist.all_reduce(pt, op=dist.ReduceOp.SUM)
    ms_per_step = float(t[0].item()) / args.steps
    pairs_all = float(pt[0].item())
    peak, peak_src = measured_peak_gbs()
    km = {k: float(np.mean([d[k] for d in kern_ns])) * 1e-6 for k in kern_ns[0]}
    roof = stage_roofline(n_own, m_own, pairs, km, peak, peak_src)  # rank 0's share of the job (its contigs)
    roof["note"] = f"rank 0's share of the job: {n_own} reads x {m_own} variants -> {pairs} pairs; stage times of its timed steps"

    # ---- e2e: pinned host columns -> H2D -> exchange -> join -> D2H of counts and global pair ids ----------------
    e2e = None
    if not args.skip_e2e:
        cap = int(pairs * 1.25) + 1024
        h_a = torch.empty(cap, dtype=torch.int32).pin_memory()
        h_b = torch.empty(cap, dtype=torch.int32).pin_memory()
        h_cnt = torch.empty(int(n // world * 1.5) + 4096, dtype=torch.int64).pin_memory()

        def step_e2e():
            tp = [x.to(dev, non_blocking=True) for x in hp]
            tb = [x.to(dev, non_blocking=True) for x in hb]
            cnt, a, b, qrow, qc, _ = step(tables=(tb, tp))
            h_cnt[: cnt.numel()].copy_(cnt, non_blocking=True)
            h_a[: a.numel()].copy_(a, non_blocking=True)
            h_b[: b.numel()].copy_(b, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return cnt.numel(), a.numel()

        for _ in range(2):
            step_e2e()
        dist.barrier(); torch.cuda.synchronize()
        e2e_steps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            rows_c, rows_o = step_e2e()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([sec], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": pairs_all / float(te.item()), "unit": "pairs/s", "h2d_bytes_per_step": int(12 * (n + m)),
               "d2h_bytes_per_step": int(8 * n + 8 * pairs_all), "ms_per_step": float(te.item()) * 1e3,
               "api": "per rank: pinned host int32 columns of its 1/N row block -> H2D -> dist.shard_tables (exchange) -> "
                      "DeviceIndex.count_overlaps + overlap_pairs -> global ids -> D2H into pinned buffers; max over ranks"}

    line = {
        "metric": METRIC, "value": pairs_all / (ms_per_step * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_string(n, m), "pairs_per_step": pairs_all,
                   "l2": "inputs (2.3 GB of int32 columns over all ranks) are larger than L2; no flush needed",
                   "parallelism": (f"strong scaling x{world}: every rank starts with a contiguous 1/{world} block of the rows of both tables; "
                                   "contig owners by LPT over the 24 contigs; "
                                   + ("rows stored straight into their owner's columns over NVLink peer memory (CUDA IPC arenas; owner "
                                      "table + region layout planned on the device; "
                                      + ("histograms and completion flags also travel through peer memory: no NCCL call in a step)"
                                         if pbd.exchange_sync() == "flags" else "NCCL for a histogram all_gather and the closing all_reduce)")
                                      if pbd.exchange_kind() == "peer" else "NCCL all-to-all of 16-byte records")),
                   "exchange": pbd.exchange_kind(), "exchange_sync": pbd.exchange_sync(),
                   "exchange_overlap": ("index build overlaps the reads' transfer (one stream per table; exchange_ms_per_step then "
                                        "includes the index build)" if overlap and pbd.exchange_kind() == "peer" else "none"),
                   "exchange_ms_per_step": float(t[1].item()) / args.steps,
                   "exchange_host_laps_ms": {k: round(v * 1e3, 3) for k, v in xtrace},
                   "exchange_bytes_per_gpu": 16 * (n + m) // world},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roof,
        "parity_check": parity.get("status", "skipped"),
        "parity_detail": parity,
    }
    if rank == 0:
        print(json.dumps(line))
    pbd.close_peer_exchanges()
    dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=wl.C3_READS)
    ap.add_argument("--variants", type=int, default=wl.C3_VARIANTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--skip-secondary", action="store_true")
    ap.add_argument("--skip-payload-e2e", action="store_true", help="N = 1: skip the e2e overlap variant with payload columns")
    ap.add_argument("--skip-replicate", action="store_true", help="N > 1 parity: skip the replicate-strategy leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the Arrow bridge's host pool (key encoding, gather) defaults to every logical CPU: share the box between ranks
        os.environ.setdefault("PBGPU_HOST_THREADS", str(max(4, (os.cpu_count() or 8) // world)))
        dist.init_process_group("nccl", device_id=dev)
        return run_sharded(args, world, rank, dev, local)
    return run_single(args, dev, local)


if __name__ == "__main__":
    main()
