"""Builds libmvsmpl.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mvsmplfitting_b200.build [--force]
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libmvsmpl.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(_HERE, "..", "include", "*.h"))
    return os.path.getmtime(OUT) < max(os.path.getmtime(p) for p in deps)


def build_debug() -> str:
    """libmvsmpl_dbg.so with -DMVS_PHASE_DBG (per-phase clock stamps in frame_step_kernel); profiling scripts only"""
    out = os.path.join(_HERE, "libmvsmpl_dbg.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-DMVS_PHASE_DBG", "-o", out] + sources() + ["-lcuda"])
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + sources() + ["-lcuda"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    if "--debug" in sys.argv:
        print(build_debug())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
