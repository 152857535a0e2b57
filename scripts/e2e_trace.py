"""Host-path tuning aid: wall time of pb.count_overlaps / pb.overlap on config 2 with PBGPU_TRACE stage laps from
the C++ side and a Python-side split (export / engine call / import+read_all)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PBGPU_TRACE", "1")
import ctypes

import numpy as np
import pyarrow as pa

import polars_bio_b200 as pb
from bench import make_config2
from polars_bio_b200 import _native, range_op_io
from polars_bio_b200.options import FilterOp, RangeOp, RangeOptions

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
    t3 = time.perf_counter(); del c; t4 = time.perf_counter(); del o; t5 = time.perf_counter()
    print(f"free count result {1e3*(t4-t3):.1f} ms  free overlap result {1e3*(t5-t4):.1f} ms", file=sys.stderr)

# Python-side split of one count_overlaps call
os.environ["PBGPU_TRACE"] = "0"
ro = RangeOptions(range_op=RangeOp.CountOverlapsNaive, filter_op=FilterOp.Strict, columns_1=list(cols), columns_2=list(cols))
for it in range(2):
    t = [time.perf_counter()]
    r1, r2 = vars_t.to_reader(), reads_t.to_reader(); t.append(time.perf_counter())
    s1, s2, so = range_op_io._CStream(), range_op_io._CStream(), range_op_io._CStream()
    r1._export_to_c(ctypes.addressof(s1)); r2._export_to_c(ctypes.addressof(s2)); t.append(time.perf_counter())
    opts = range_op_io._c_opts(ro, 0, None, pb.ctx)
    rc = _native.lib().pbgpu_range_op(ctypes.addressof(s1), ctypes.addressof(s2), ctypes.byref(opts), ctypes.addressof(so)); t.append(time.perf_counter())
    rd = pa.RecordBatchReader._import_from_c(ctypes.addressof(so)); t.append(time.perf_counter())
    tab = rd.read_all(); t.append(time.perf_counter())
    del rd; t.append(time.perf_counter())
    names = ["to_reader", "export_to_c", "pbgpu_range_op", "import_from_c", "read_all", "del reader"]
    print("python split: " + "  ".join(f"{n} {1e3*(b-a):.2f}" for n, a, b in zip(names, t[:-1], t[1:])), file=sys.stderr)
    del tab
