"""Builds libosfm_match.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libosfm_match.so")
SOURCES = ["osfm_match.cu"]
HEADERS = ["ptx.cuh", "common.cuh", "scan_kernel.cuh", "post_kernels.cuh", "float_kernels.cuh",
           "tracks_kernels.cuh", "io_formats.cuh", "ransac_math.cuh", "ransac_kernels.cuh",
           "gpu_exhaustive_matching.h", "../../include/osfm_match.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
] + os.environ.get("OSFM_NVCC_EXTRA", "").split() + [
]


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        build_tools()
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SOURCES
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a gcc wrapper that cannot link OpenMP; use the system one
    env.pop("CXX", None)
    subprocess.check_call(cmd, cwd=HERE, env=env)
    build_tools()
    return SO


PLUGIN_BENCH = os.path.join(HERE, "plugin_loop_bench")


def build_tools() -> str:
    """plugin_loop_bench: the reference's plugin call pattern against the C ABI, in C++ (bench.py's
    e2e.plugin figures)."""
    src = os.path.join(HERE, "plugin_loop_bench.cc")
    if os.path.exists(PLUGIN_BENCH) and os.path.getmtime(PLUGIN_BENCH) >= max(os.path.getmtime(src), os.path.getmtime(SO)):
        return PLUGIN_BENCH
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", src, "-o", PLUGIN_BENCH, "-L" + HERE, "-losfm_match",
                           "-Wl,-rpath,$ORIGIN"], cwd=HERE)
    return PLUGIN_BENCH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
