"""In-tree build of the native pieces (no pip, no JIT cache: the .so files must
travel with the repo snapshot to the GPU box).

    python -m cudafluidsimulator_b200.build            # library + CLI
    python -m cudafluidsimulator_b200.build --oracle   # also the test checkers

Targets
    cudafluidsimulator_b200/libsph_b200.so   CUDA kernels + C ABI (include/sph_b200.h), sm_100a only
    cudafluidsimulator_b200/sph              drop-in `./sph -n -i -m` CLI (host C++ over the C ABI)
    oracle/ (via its Makefile)               test infrastructure only
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HOST = PKG / "host"
INCLUDE = ROOT / "include"
LIB = PKG / "libsph_b200.so"
CLI = PKG / "sph"

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only, no PTX fallback for other archs
    "-lineinfo", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-shared",
]
CU_SOURCES = ["sph_api.cu", "sph_kernels.cu", "sph_sort.cu", "sph_cluster.cu"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libsph_b200.so")
    return exe


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, **kw)


LIB_CHECKED = PKG / "libsph_b200_checked.so"


def build_library(force: bool = False, verbose_ptxas: bool = False, checked: bool = False) -> Path:
    """checked=True: the self-checking variant (-DSPH_BOUNDS_CHECK), same ABI, loaded through
    SPH_B200_LIB; used by tests/test_gpu_checked_build.py in place of compute-sanitizer."""
    deps = [CSRC / f for f in CU_SOURCES] + list(CSRC.glob("*.cuh")) + [INCLUDE / "sph_b200.h"]
    target = LIB_CHECKED if checked else LIB
    if force or _stale(target, deps):
        flags = list(NVCC_FLAGS) + (["-Xptxas", "-v"] if verbose_ptxas else [])
        if checked:
            flags.append("-DSPH_BOUNDS_CHECK")
        # libdl: NCCL is resolved with dlopen when a multi-process cluster is created (sph_cluster.cu)
        _run([_nvcc(), *flags, "-I", INCLUDE, "-o", target, *[CSRC / f for f in CU_SOURCES], "-ldl", "-lpthread"])
    return target


def build_cli(force: bool = False) -> Path:
    srcs = [HOST / "main.cpp", HOST / "simulator.cpp"]
    deps = srcs + [INCLUDE / "simulator.h", INCLUDE / "times.h", INCLUDE / "sph_b200.h", LIB]
    if force or _stale(CLI, deps):
        cuda_inc = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda")) / "include"
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-I", INCLUDE, "-I", cuda_inc, "-o", CLI, *srcs,
              "-L", PKG, "-lsph_b200", "-Wl,-rpath,$ORIGIN"])
    return CLI


def build_oracle() -> None:
    """Test infrastructure: the C restatement and (when /root/reference is
    mounted) the unmodified reference CUDA build."""
    _run(["make", "-C", ROOT / "oracle", "--no-print-directory"])


def build_all(force: bool = False, oracle: bool = False) -> None:
    build_library(force)
    build_cli(force)
    if oracle:
        build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, oracle="--oracle" in sys.argv)
    if "--checked" in sys.argv:
        build_library(checked=True)
