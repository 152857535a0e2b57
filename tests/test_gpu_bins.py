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


@pytest.mark.parametrize("bins", ["0", "1"])
def test_explicit_id_columns_on_every_pass2_path(bins):
    """pbgpu_index_build_ids / pbgpu_overlap_count_ids (global row ids of a sharded join): tests/tools/ids_check.py with the
    probe partition off and forced."""
    env = dict(os.environ, PBGPU_BIN=bins)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "ids_check.py")], env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "IDS_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
