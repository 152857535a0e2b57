"""Payload materialisation on the device (reference: src/operation.rs:272-303 -- the output of overlap IS
the joined rows).  Wide frames in the shape of the reference's tests/test_wide_dataframes.py:80-257: every extra column of
both inputs must come back, suffixed, with the values of the matched rows.  The device-gathered columns (fixed-width and
utf8 / large_utf8 / binary, nulls included) are compared cell for cell with pyarrow.take over the index pairs of the same
join, and with the host gather path (PBGPU_DEV_GATHER=0, a subprocess: the switch is read once)."""
import os
import subprocess
import sys

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import polars_bio_b200 as pb  # noqa: E402
from polars_bio_b200 import FilterOp, RangeOp, RangeOptions, range_op_io  # noqa: E402
from tests._golden import synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["contig", "pos_start", "pos_end"]
NAMES = np.array(["chr1", "chr2", "chrX"])


def _wide(n, seed, tag, batches=1):
    rng = np.random.default_rng(seed)
    c, s, e = synth(n, 3, 2_000_000, 2500 if tag == "b" else 300, seed)
    null = lambda a, frac, typ=None: pa.array([None if rng.random() < frac else v for v in a], type=typ)
    cols = {
        "contig": pa.array(NAMES[c]), "pos_start": pa.array(s), "pos_end": pa.array(e),
        f"{tag}_gene_id": pa.array([f"GENE{int(x):05d}" for x in rng.integers(0, 50_000, n)]),                # utf8
        f"{tag}_score": null(rng.random(n), 0.1, pa.float64()),                                                # float64 with nulls
        f"{tag}_strand": pa.array(np.array(["+", "-"])[rng.integers(0, 2, n)]),                                # short utf8
        f"{tag}_i8": pa.array(rng.integers(-100, 100, n).astype(np.int8)),
        f"{tag}_u16": null(rng.integers(0, 60_000, n).astype(np.uint16), 0.05, pa.uint16()),
        f"{tag}_i64": pa.array(rng.integers(-2**62, 2**62, n)),
        f"{tag}_ts": pa.array(rng.integers(0, 2**40, n), pa.timestamp("us")),
        f"{tag}_dec": pa.array([None if i % 11 == 0 else int(v) for i, v in enumerate(rng.integers(0, 10**9, n))], pa.int64()).cast(pa.decimal128(24, 3)),
        f"{tag}_note": pa.array([None if i % 7 == 0 else ("" if i % 5 == 0 else "note-" + "x" * int(i % 23)) for i in range(n)], pa.large_string()),
        f"{tag}_blob": pa.array([None if i % 13 == 0 else bytes([i % 251]) * (i % 9) for i in range(n)], pa.binary()),
        f"{tag}_flag": pa.array(rng.random(n) < 0.5),                                                          # bool: host gather path
        f"{tag}_cat": pa.array(np.array(["exon", "intron", "utr"])[rng.integers(0, 3, n)]).dictionary_encode(),  # dictionary: host path
    }
    t = pa.table(cols)
    if batches > 1:  # several record batches with slices that do not start at buffer offset 0
        step = max(1, n // batches)
        t = pa.Table.from_batches([b for off in range(0, n, step) for b in t.slice(off, step).to_batches()])
    return pb.set_coordinate_system(t, True)


def _opts():
    return RangeOptions(range_op=RangeOp.Overlap, filter_op=FilterOp.Strict, suffixes=("_1", "_2"), columns_1=COLS, columns_2=COLS)


def _expected(left, right, pairs):
    l = left.combine_chunks().take(pairs.column(0))
    r = right.combine_chunks().take(pairs.column(1))
    cols, names = [], []
    for t, sfx in ((l, "_1"), (r, "_2")):
        for name in t.column_names:
            col = t.column(name)
            if pa.types.is_dictionary(col.type):
                col = col.cast(pa.large_string())  # dictionary<string> comes back decoded
            cols.append(col)
            names.append(name + sfx)
    return pa.table(cols, names=names)


@pytest.mark.parametrize("batches,sink_pairs", [(1, 0), (7, 0), (5, 4096)])
def test_wide_overlap_equals_take_over_index_pairs(batches, sink_pairs, monkeypatch):
    left, right = _wide(40_000, 1, "a", batches), _wide(9_000, 2, "b", batches)
    if sink_pairs:
        monkeypatch.setenv("PBGPU_SINK_PAIRS", str(sink_pairs))  # many chunks through the streaming sink
    pairs = range_op_io.range_operation_frame(pb.ctx, left, right, _opts(), emit=1).to_arrow()
    got = range_op_io.range_operation_frame(pb.ctx, left, right, _opts()).to_arrow()
    want = _expected(left, right, pairs)
    assert got.num_rows == want.num_rows > 10_000
    assert got.column_names == want.column_names
    for name in want.column_names:
        g, w = got.column(name).combine_chunks(), want.column(name).combine_chunks()
        if pa.types.is_large_string(g.type) and pa.types.is_string(w.type):
            w = w.cast(pa.large_string())  # the contig key column is rebuilt as large_utf8 when needed
        assert g.type == w.type, (name, g.type, w.type)
        assert g.equals(w), name


def test_extra_columns_present_and_values_come_from_their_frame():
    """tests/test_wide_dataframes.py:98-121."""
    left, right = _wide(5_000, 3, "a"), _wide(2_000, 4, "b")
    out = pb.overlap(left, right, cols1=COLS, cols2=COLS, output_type="pandas.DataFrame", suffixes=("_1", "_2"))
    assert len(out) > 0
    for c in ("a_gene_id_1", "a_score_1", "a_strand_1", "b_gene_id_2", "b_i64_2", "b_note_2"):
        assert c in out.columns
    assert set(out["a_gene_id_1"].unique()).issubset(set(left.column("a_gene_id").to_pylist()))
    assert set(out["b_gene_id_2"].unique()).issubset(set(right.column("b_gene_id").to_pylist()))


def test_device_gather_equals_host_gather():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import pyarrow as pa, polars_bio_b200 as pb\n"
            "from tests.test_gpu_payload import _wide, _opts\n"
            "from polars_bio_b200 import range_op_io\n"
            "t = range_op_io.range_operation_frame(pb.ctx, _wide(20000, 1, 'a', 3), _wide(6000, 2, 'b', 2), _opts()).to_arrow()\n"
            "t = t.sort_by([(n, 'ascending') for n in ('pos_start_1', 'pos_end_1', 'pos_start_2', 'pos_end_2', 'a_i64_1', 'b_i64_2')])\n"
            "import hashlib; h = hashlib.sha256()\n"
            "for c in t.columns:\n"
            "    h.update(repr(c.to_pylist()).encode())\n"
            "print('DIGEST', t.num_rows, h.hexdigest())\n") % ROOT
    outs = []
    for flag in ("1", "0"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PBGPU_DEV_GATHER=flag), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][0])
    assert outs[0] == outs[1]
