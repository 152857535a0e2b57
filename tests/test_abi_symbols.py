"""The C-ABI library loads and exports every symbol include/pbgpu.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pbgpu.h")).read()
    return sorted(set(re.findall(r"PBGPU_API[^;(]*?\b(pbgpu_\w+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = declared_symbols()
    for must in ("pbgpu_index_build", "pbgpu_count_overlaps", "pbgpu_overlap_count", "pbgpu_overlap_emit",
                 "pbgpu_nearest", "pbgpu_coverage", "pbgpu_range_op", "pbgpu_pack_by_owner", "pbgpu_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "polars_bio_b200", "libpbgpu.so"))
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    lib.pbgpu_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pbgpu_version()
    lib.pbgpu_last_error.restype = ctypes.c_char_p
    assert lib.pbgpu_last_error() is not None


def test_sass_is_sm100a_only():
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "polars_bio_b200", "libpbgpu.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "polars_bio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, re.M), f
                assert "liboracle" not in txt, f


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct that crosses the ABI, as gcc lays it out from include/pbgpu.h, against the
    ctypes mirrors in polars_bio_b200/_native.py (a drifted mirror would corrupt arguments silently)."""
    import subprocess

    from polars_bio_b200 import _native

    structs = {"PbRangeOptions": _native.PbRangeOptions, "pbgpu_peer_step": _native.PbPeerStep, "pbgpu_stage_times": _native.StageTimes}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pbgpu.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, value = line.split()
        cls = structs[cname]
        want = ctypes.sizeof(cls) if field == "sizeof" else getattr(cls, field).offset
        assert int(value) == want, (cname, field, int(value), want)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())
