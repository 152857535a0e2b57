"""Same-box A/B of the host path: wall time of pb.count_overlaps / pb.overlap on BASELINE config 3 (PB_SCALE shrinks it)
under the environment it was started with.  Prints one JSON line (median of PB_ITERS iterations after one warm-up)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

import polars_bio_b200 as pb
import workloads as wl

scale = float(os.environ.get("PB_SCALE", "1.0"))
iters = int(os.environ.get("PB_ITERS", "4"))
n, m = int(wl.C3_READS * scale), int(wl.C3_VARIANTS * scale)
names = pa.array(wl.CONTIG_NAMES)


def table(cols):
    c, s_, e_ = cols
    return pb.set_coordinate_system(pa.table({"contig": pc.take(names, pa.array(c)), "pos_start": pa.array(s_), "pos_end": pa.array(e_)}), True)


reads_t, vars_t = table(wl.config3_reads(0, n, n)), table(wl.config3_variants(0, m, m))
cols = ("contig", "pos_start", "pos_end")
rows = []
for it in range(iters + 1):
    t0 = time.perf_counter()
    rc = sum(b.num_rows for b in pb.count_overlaps(reads_t, vars_t, cols1=cols, cols2=cols, output_type="datafusion.DataFrame").execute_stream())
    t1 = time.perf_counter()
    res = pb.overlap(reads_t, vars_t, cols1=cols, cols2=cols, output_type="datafusion.DataFrame")
    t2 = time.perf_counter()
    ro = sum(b.num_rows for b in res.execute_stream())
    t3 = time.perf_counter()
    if it:
        rows.append((t1 - t0, t2 - t1, t3 - t2))
med = np.median(np.array(rows), axis=0) * 1e3
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PBGPU_")}, "count_overlaps_ms": round(float(med[0]), 1),
                  "overlap_call_ms": round(float(med[1]), 1), "overlap_consume_ms": round(float(med[2]), 1), "total_ms": round(float(med.sum()), 1),
                  "rows": [rc, ro]}))
