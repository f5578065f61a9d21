"""GPU suite: the self-checking build (-DSPH_BOUNDS_CHECK) verifies every data-dependent index
of the hot kernels on the device (run bounds, table indices, gather indices, mask words,
emigrant slots).  compute-sanitizer is closed on the GPU pool, so these own checks on small
cases -- together with the comparisons against the CPU oracle -- are the memory-safety evidence."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_no_bounds_violation_in_checked_build():
    from cudafluidsimulator_b200 import build
    lib = build.build_library(checked=True)
    env = dict(os.environ, SPH_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, str(Path(__file__).with_name("checked_build_worker.py"))],
                       capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "CHECKED_BUILD_FLAGS 0" in r.stdout, r.stdout[-3000:]
