"""Drives the polars-only branches of polars_bio_b200 with the stand-in of tests/fake_polars (TEST INFRASTRUCTURE; run in a
subprocess so that `import polars` resolves to the stand-in before the package is imported):

  python tests/tools/polars_standin_check.py cpu   # LazyFrame.pb / DataFrame.pb namespaces + polars in / out through the
                                                   # unary sweeps, device calls = the CPU doubles of the bridge harness
  python tests/tools/polars_standin_check.py gpu   # + the IO-plugin source of range_lazy_scan (projection / predicate /
                                                   # row-limit pushdown, re-execution) and polars inputs of the binary calls
Prints POLARS_STANDIN_OK <mode> on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "fake_polars"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import polars as pl  # noqa: E402  (the stand-in)
import pyarrow as pa  # noqa: E402

assert pl.__version__.endswith("standin")
import polars_bio_b200 as pb  # noqa: E402
from polars_bio_b200 import _native, polars_ext  # noqa: E402,F401  (registers the `pb` namespaces)

mode = sys.argv[1] if len(sys.argv) > 1 else "cpu"


def frame(d, zero_based=True):
    return pb.set_coordinate_system(pl.DataFrame(pa.table(d)), zero_based)


left = frame({"chrom": ["chr2", "chr10", "chr2", "chr10", "chr2"], "start": [5, 1, 7, 100, 30], "end": [10, 4, 12, 200, 40],
              "name": ["a", "b", "c", "d", "e"]})
other = frame({"chrom": ["chr2", "chr3"], "start": [6, 0], "end": [8, 50]})

if mode == "cpu":  # unary sweeps run end to end on the CPU harness build of the Arrow level
    from tests import _harness

    H = _harness.build()
    H.dbg_streams_ok(1)
    _native.lib = lambda: H

# ---- namespaces + polars in / polars out (eager and lazy) through the unary sweeps --------------------------------
m = left.pb.merge()                       # DataFrame.pb namespace; default output type is a polars LazyFrame
assert isinstance(m, pl.LazyFrame)
mt = m.collect().to_arrow()
assert mt.column_names == ["chrom", "start", "end", "n_intervals"]
assert mt.column("chrom").to_pylist() == ["chr10", "chr10", "chr2", "chr2"] and mt.column("n_intervals").to_pylist() == [1, 1, 2, 1]
c = pb.set_coordinate_system(left.lazy(), True).pb.cluster().collect()    # LazyFrame.pb namespace, LazyFrame input (collected by _df_to_reader)
assert isinstance(c, pl.DataFrame) and c.columns == ["chrom", "start", "end", "name", "cluster", "cluster_start", "cluster_end"]
assert c.to_arrow().column("cluster").to_pylist() == [2, 0, 2, 1, 3]
s = pb.subtract(left, other, output_type="polars.DataFrame")
assert isinstance(s, pl.DataFrame)
rows = sorted(zip(*(s.to_arrow().column(k).to_pylist() for k in ("name", "start", "end"))))
assert rows == [("a", 5, 6), ("a", 8, 10), ("b", 1, 4), ("c", 8, 12), ("d", 100, 200), ("e", 30, 40)], rows
assert pb.get_coordinate_system(s) is True  # the result is tagged like its inputs
g = left.pb.complement().collect().to_arrow()
assert g.column("end").to_pylist()[-1] == np.iinfo(np.int64).max

if mode == "gpu":
    rng = np.random.default_rng(3)
    n, k = 20_000, 3_000
    names = np.array(["chr1", "chr2", "chrX"])

    def table(rows, width, seed):
        r = np.random.default_rng(seed)
        s_ = r.integers(0, 1_000_000, rows).astype(np.int32)
        return {"chrom": names[r.integers(0, 3, rows)].tolist(), "start": s_, "end": (s_ + r.integers(1, width, rows)).astype(np.int32),
                "score": r.random(rows)}

    a, b = table(n, 300, 1), table(k, 2_000, 2)
    fa, fb = frame(a), frame(b)
    want = pb.overlap(pb.set_coordinate_system(pa.table(a), True), pb.set_coordinate_system(pa.table(b), True), output_type="pyarrow.Table")
    key = lambda t: t.sort_by([(c_, "ascending") for c_ in t.column_names])
    lf = pb.set_coordinate_system(fa.lazy(), True).pb.overlap(fb)  # LazyFrame in, LazyFrame out: the IO-plugin source, nothing has run yet
    assert isinstance(lf, pl.LazyFrame) and list(lf.collect_schema()) == want.column_names
    got = lf.collect().to_arrow()
    assert key(got).equals(key(pa.Table.from_arrays(want.columns, names=want.column_names)).cast(got.schema)), "lazy != eager"
    assert lf.collect().height == want.num_rows                                  # re-execution from fresh streams
    proj = lf.select("start_1", "chrom_2").collect()                             # projection pushdown
    assert proj.columns == ["start_1", "chrom_2"] and proj.height == want.num_rows
    assert lf.head(17).collect().height == 17                                    # row-limit pushdown
    flt = lf.filter(pl.col("start_1") > 500_000).collect().to_arrow()            # predicate applied per batch by the source
    assert flt.num_rows == sum(1 for v in want.column("start_1").to_pylist() if v > 500_000) and flt.num_rows > 0
    cnt = pb.count_overlaps(fa, fb, output_type="polars.DataFrame")              # polars in, polars DataFrame out
    assert isinstance(cnt, pl.DataFrame) and cnt.columns[-1] == "count" and cnt.height == n
    assert int(np.sum(cnt.to_arrow().column("count").to_numpy())) == want.num_rows
    near = fa.pb.nearest(fb).collect()
    assert near.height == n and "distance" in near.columns

print("POLARS_STANDIN_OK", mode)
