import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PBGPU_TRACE"] = "1"
import numpy as np, pyarrow as pa
import polars_bio_b200 as pb
from bench import make_config2
probe, build, nc = make_config2()
def table(cols):
    c, s_, e_ = cols
    return pb.set_coordinate_system(pa.table({"contig": pa.array(np.full(len(c), "chr1")), "pos_start": pa.array(s_), "pos_end": pa.array(e_)}), True)
reads_t, vars_t = table(probe), table(build)
cols = ("contig", "pos_start", "pos_end")
for it in range(3):
    print("---- iteration", it, file=sys.stderr)
    t0 = time.perf_counter(); c = pb.count_overlaps(reads_t, vars_t, cols1=cols, cols2=cols, output_type="pyarrow.Table"); t1 = time.perf_counter()
    o = pb.overlap(reads_t, vars_t, cols1=cols, cols2=cols, output_type="pyarrow.Table"); t2 = time.perf_counter()
    print(f"count_overlaps {1e3*(t1-t0):.1f} ms  overlap {1e3*(t2-t1):.1f} ms rows {c.num_rows} {o.num_rows}", file=sys.stderr)
