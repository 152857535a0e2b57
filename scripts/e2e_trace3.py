"""Host-path tuning aid: PBGPU_TRACE stage laps of pb.count_overlaps / pb.overlap on BASELINE config 3 (PB_SCALE shrinks it)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PBGPU_TRACE", "1")
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

import polars_bio_b200 as pb
import workloads as wl

scale = float(os.environ.get("PB_SCALE", "1.0"))
n, m = int(wl.C3_READS * scale), int(wl.C3_VARIANTS * scale)
names = pa.array(wl.CONTIG_NAMES)


def table(cols):
    c, s_, e_ = cols
    return pb.set_coordinate_system(pa.table({"contig": pc.take(names, pa.array(c)), "pos_start": pa.array(s_), "pos_end": pa.array(e_)}), True)


reads_t, vars_t = table(wl.config3_reads(0, n, n)), table(wl.config3_variants(0, m, m))
cols = ("contig", "pos_start", "pos_end")
for it in range(3):
    print("---- iteration", it, file=sys.stderr)
    t0 = time.perf_counter()
    rc = sum(b.num_rows for b in pb.count_overlaps(reads_t, vars_t, cols1=cols, cols2=cols, output_type="datafusion.DataFrame").execute_stream())
    t1 = time.perf_counter()
    print(f"== count_overlaps {1e3 * (t1 - t0):.1f} ms rows {rc}", file=sys.stderr)
    os.environ["PBGPU_TRACE"] = "0" if it < 2 else "1"
    ro = 0
    t1 = time.perf_counter()
    res = pb.overlap(reads_t, vars_t, cols1=cols, cols2=cols, output_type="datafusion.DataFrame")
    t2 = time.perf_counter()
    nb = 0
    for b in res.execute_stream():
        ro += b.num_rows
        nb += 1
    t3 = time.perf_counter()
    os.environ["PBGPU_TRACE"] = "1"
    print(f"== overlap call {1e3 * (t2 - t1):.1f} ms + consume {1e3 * (t3 - t2):.1f} ms rows {ro} batches {nb}", file=sys.stderr)
