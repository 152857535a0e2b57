"""Builds the CPU harness of the Arrow-level host code once per test session (TEST INFRASTRUCTURE): the product source
polars_bio_b200/csrc/arrow_bridge.cpp compiled against tests/tools/bridge_harness/stub_cuda.h with harness_tail.inc
appended.  Most device calls of the binary operations are stubs that fail; the unary sweeps and the index build +
count_overlaps have plain CPU doubles, so pbgpu_range_op runs end to end for merge / cluster / complement / subtract and
count_overlaps (after dbg_streams_ok(1))."""
import ctypes
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "tools", "bridge_harness")
_lib = None


def build():
    global _lib
    if _lib is not None:
        return _lib
    d = tempfile.mkdtemp(prefix="pb_bridge_")
    src = open(os.path.join(ROOT, "polars_bio_b200", "csrc", "arrow_bridge.cpp")).read()
    assert "#include <cuda_runtime.h>" in src
    src = src.replace("#include <cuda_runtime.h>", '#include "stub_cuda.h"') + open(os.path.join(HARNESS, "harness_tail.inc")).read()
    cpp = os.path.join(d, "bridge_host.cpp")
    open(cpp, "w").write(src)
    so = os.path.join(d, "libbridge_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    extra = os.environ.get("PB_BRIDGE_CXXFLAGS", "").split()  # e.g. -fsanitize=address,undefined (scripts/bridge_sanitize.sh)
    r = subprocess.run([cxx, "-std=c++17", "-O1", *extra, "-fPIC", "-shared", "-I", HARNESS, "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "polars_bio_b200", "csrc"), "-o", so, cpp, "-lpthread"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    os.environ.setdefault("PBGPU_ASYNC_MIN_ROWS", "0")  # the helper-thread index build also for the small tables of these tests
    L = ctypes.CDLL(so)
    L.dbg_roundtrip.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64,
                                ctypes.c_void_p, ctypes.c_void_p]
    L.dbg_streams_ok.argtypes = [ctypes.c_int]
    L.dbg_streams_ok.restype = None
    L.pbgpu_last_error.restype = ctypes.c_char_p
    L.pbgpu_range_op.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    _lib = L
    return L
