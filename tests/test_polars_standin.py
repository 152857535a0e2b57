"""The polars-only branches of the package (LazyFrame.pb / DataFrame.pb namespaces, polars inputs and outputs, the
IO-plugin source of range_lazy_scan) driven with the STAND-IN of tests/fake_polars -- polars itself is not installable in
the build image.  Each check runs in a subprocess so that `import polars` resolves to the stand-in before the package is
imported (tests/tools/polars_standin_check.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "tools", "polars_standin_check.py")


def _run(mode):
    r = subprocess.run([sys.executable, SCRIPT, mode], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"POLARS_STANDIN_OK {mode}" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def test_namespaces_and_polars_frames_through_the_unary_sweeps_on_the_cpu_harness():
    _run("cpu")


@pytest.mark.gpu
def test_io_plugin_source_and_polars_frames_through_the_binary_operations():
    _run("gpu")
