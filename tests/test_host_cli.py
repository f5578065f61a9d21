"""CPU suite: the C++ host side -- `./sph` flag handling (ref: src/main.cpp:20-55) and the
`-m time` table (ref: src/times.h:12-36), byte for byte."""
import os
import subprocess
from pathlib import Path

import pytest

from conftest import GOLDEN, ROOT, has_gpu

SPH = ROOT / "cudafluidsimulator_b200" / "sph"
REF = Path("/root/reference/src")


@pytest.fixture(scope="module", autouse=True)
def _cli_built():
    from cudafluidsimulator_b200 import build
    build.build_library()
    build.build_cli()


def run(*args):
    return subprocess.run([str(SPH), *args], capture_output=True, text=True, timeout=120)


def test_usage_and_invalid_values_exit_1():
    # ref: main.cpp:28-53 -- message on stdout, usage, exit status 1
    r = run("-i", "bogus")
    assert r.returncode == 1 and r.stdout.startswith("Invalid argument for option -i: bogus\nProgram Options:")
    r = run("-m", "fast")
    assert r.returncode == 1 and "Invalid argument for option -m: fast" in r.stdout
    r = run("-?")
    assert r.returncode == 1 and "-n  <NUM_PARTICLES>" in r.stdout and "-m  <free/time>" in r.stdout


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_gpu_is_a_loud_failure_not_a_cpu_path():
    r = run("-n", "100")
    assert r.returncode == 2 and "sph_setup failed" in r.stderr


TIMES_PROGRAM = r"""
#include "times.h"
int main() {
    Times t; t.buildGrid = 0.0123456; t.sphUpdate = 1.5; t.memcpy = 0.00042; t.iters = 100;
    displayTimes(&t);
    Times z; displayTimes(&z);
    return 0;
}
"""


def _table(include_dir, tmp_path, name):
    src = tmp_path / f"{name}.cpp"
    src.write_text(TIMES_PROGRAM)
    exe = tmp_path / name
    subprocess.run(["g++", "-O1", "-I", str(include_dir), "-o", str(exe), str(src)], check=True)
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout


def test_display_times_matches_golden_text(tmp_path):
    ours = _table(ROOT / "include", tmp_path, "ours")
    golden = (GOLDEN / "display_times.txt").read_text()
    assert ours == golden


@pytest.mark.skipif(not (REF / "times.h").exists(), reason="/root/reference not mounted")
def test_display_times_matches_reference_header(tmp_path):
    assert _table(ROOT / "include", tmp_path, "ours") == _table(REF, tmp_path, "ref")


@pytest.mark.gpu
def test_time_mode_prints_reference_table():
    r = run("-n", "10000", "-i", "grid", "-m", "time", "-s", "10")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].split() == ["Operation", "Per", "frame", "Total"]
    assert lines[1] == "-" * 45
    assert lines[2].startswith("Grid construction") and lines[3].startswith("SPH update") and \
        lines[4].startswith("Data transfer")


@pytest.mark.gpu
def test_cli_extension_flags():
    r = run("-n", "200000", "-i", "random", "-m", "time", "-s", "5", "-b", "25.6", "-c", "256", "-k", "morton")
    assert r.returncode == 0, r.stderr
    r = run("-n", "2000000", "-i", "grid", "-m", "time", "-s", "2")   # > 109^3 in the reference box
    assert r.returncode == 2 and "lattice" in r.stderr


@pytest.mark.gpu
def test_state_dump_and_load_round_trip(tmp_path):
    """-d after 10 steps, then -l + 10 more steps == 20 steps straight (deterministic sort)."""
    a, b, c = tmp_path / "a.bin", tmp_path / "b.bin", tmp_path / "c.bin"
    assert run("-n", "5000", "-s", "10", "-d", str(a)).returncode == 0
    assert run("-n", "5000", "-s", "10", "-l", str(a), "-d", str(b)).returncode == 0
    assert run("-n", "5000", "-s", "20", "-d", str(c)).returncode == 0
    import numpy as np
    fb, fc = np.fromfile(b, np.uint8), np.fromfile(c, np.uint8)
    assert fb[:8].tobytes() == b"SPHB200\x00" and len(fb) == 12 + 5000 * 24
    pb, pc = fb[12:].view(np.float32), fc[12:].view(np.float32)
    np.testing.assert_allclose(pb, pc, rtol=1e-5, atol=1e-5)
    assert run("-n", "4000", "-l", str(a)).returncode == 2     # particle count mismatch is an error


@pytest.mark.gpu
def test_headless_free_mode_writes_frames(tmp_path):
    prefix = tmp_path / "frame"
    r = run("-n", "20000", "-m", "free", "-f", "3", "-o", str(prefix))
    assert r.returncode == 0 and "3 frames" in r.stdout, r.stderr
    data = (tmp_path / "frame_0002.ppm").read_bytes()
    assert data.startswith(b"P6\n480 480\n255\n") and len(data) == 15 + 480 * 480 * 3
    assert data.count(bytes([40, 90, 255])) > 50      # particles were drawn
