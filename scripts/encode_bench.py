"""Host key-encoder timing on the CPU harness build (no GPU): ns per row of encode_keys for the config-3 table shape
(utf8 contig names of all 24 contigs mixed, int32 positions).  PB_ROWS rows, PBGPU_HOST_THREADS threads."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa, pyarrow.compute as pc
import workloads as wl
from tests import _harness
L = _harness.build()
L.dbg_encode_ns.restype = ctypes.c_int64
L.dbg_encode_ns.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
n = int(os.environ.get("PB_ROWS", "20000000"))
c, s, e = wl.config3_reads(0, n, n)
t = pa.table({"contig": pc.take(pa.array(wl.CONTIG_NAMES), pa.array(c)), "pos_start": pa.array(s), "pos_end": pa.array(e)})
class CS(ctypes.Structure):
    _fields_ = [("a", ctypes.c_void_p)] * 0 + [("get_schema", ctypes.c_void_p), ("get_next", ctypes.c_void_p), ("get_last_error", ctypes.c_void_p), ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p)]
for code8 in (1, 0):
    st = CS(); t.to_reader()._export_to_c(ctypes.addressof(st))
    ns = L.dbg_encode_ns(ctypes.addressof(st), b"contig", b"pos_start", b"pos_end", 5, code8)
    print(f"rows {n} code8 {code8}: {ns/1e6:.1f} ms = {ns/n:.2f} ns/row (all threads)")
