"""Builds lucid_b200/_lucid_b200.so in-tree: the sm_100a kernels, the C ABI and the C++ host code.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box; the .so then travels to
the B200 box with the repository snapshot.  -fmad=false is part of the floating-point contract
(DESIGN.md): no operation is contracted into an FMA, so coverage and fragment counts are
reproducible bit for bit against the CPU checker.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO_PATH = os.path.join(HERE, "_lucid_b200.so")
# the host-side input preparation (include/lucid_host.h) on its own, without any CUDA code: what a
# process that only prepares inputs loads (bench.py's CPU reference arm, the CPU tests)
HOST_SO_PATH = os.path.join(HERE, "_lucid_host.so")

CUDA_SOURCES = ["csrc/setup.cu", "csrc/binning.cu", "csrc/raster_bins.cu", "csrc/raster_sort.cu", "csrc/raster_shade.cu",
                "csrc/sync.cu", "csrc/capi.cu", "csrc/quadgen.cu", "csrc/comparators.cu"]
HOST_SOURCES = ["host/lucid_host.cpp", "host/lucid_renderer.cpp"]
HEADERS = ["csrc/common.cuh", "csrc/raster_common.cuh", "csrc/quadgen_rules.h", "../include/lucid_quadgen.h", "../include/lucid_colour_tables.h", "../include/lucid_abi.h", "../include/lucid_b200.h", "../include/lucid_host.h",
           "../include/lucid_renderer.hpp"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(SO_PATH) or not os.path.exists(HOST_SO_PATH):
        return True
    t = min(os.path.getmtime(SO_PATH), os.path.getmtime(HOST_SO_PATH))
    for rel in CUDA_SOURCES + HOST_SOURCES + HEADERS + ["build.py"]:
        path = os.path.join(HERE, rel)
        if os.path.exists(path) and os.path.getmtime(path) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """defines / out: experiment builds (tools/variants.py); the product build uses neither."""
    if not force and not needs_build() and out is None:
        return SO_PATH
    srcs = [os.path.join(HERE, s) for s in CUDA_SOURCES + HOST_SOURCES if os.path.exists(os.path.join(HERE, s))]
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
        "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall,-Wno-unused-function", "-shared",
        "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
        "-I", os.path.join(ROOT, "include"), "-o", out or SO_PATH,
    ]
    cmd += [f"-D{d}" for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    if out is None:
        build_host()
    return out or SO_PATH


def build_host() -> str:
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-shared", "-I", os.path.join(ROOT, "include"),
           "-o", HOST_SO_PATH, os.path.join(HERE, "host/lucid_host.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed (lucid_host)")
    return HOST_SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
