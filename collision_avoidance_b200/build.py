"""Build recipe of the CUDA library (in-tree, sm_100a only).

``python -m collision_avoidance_b200.build`` compiles csrc/orca_api.cu into
``collision_avoidance_b200/liborca_b200.so``.  The flags are part of the numerical
contract: ``-fmad=false`` (no FMA contraction) keeps every float op single-rounded so the
kernels agree bit-for-bit with a scalar x86 evaluation ("Numerics").
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "liborca_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def _sources_mtime() -> float:
    newest = 0.0
    for root in (CSRC, os.path.join(PKG_DIR, "..", "include")):
        for name in os.listdir(root):
            if name.endswith((".cu", ".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def nvcc_path() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return nvcc


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _sources_mtime():
        return LIB_PATH
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra_flags, "-o", LIB_PATH, os.path.join(CSRC, "orca_api.cu")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a != "-v"]
    print(build(force=True, verbose=True, extra_flags=flags))
