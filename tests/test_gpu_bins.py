"""The probe-partition path (csrc/bins.cuh) forced on small inputs: PBGPU_BIN=1 is read once per process, so the
checks run in a subprocess (tests/tools/bins_check.py: counts row for row, pair sets, partner order, streaming sink,
coverage and nearest on the lazily built end order -- all against the oracle)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitioned_probes_match_the_oracle():
    env = dict(os.environ, PBGPU_BIN="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "bins_check.py")], env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "BINS_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
