"""Builds ``lib/libqpmpc_b200.so`` (the C-ABI CUDA library) in-tree with nvcc.

``python -m qpmpc_b200.build [--force]``.  sm_100a only; nvcc cross-compiles
without a GPU.  The library links cudart statically and nothing else, so it
loads (and exports its symbols) on a CPU-only box too.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libqpmpc_b200.so")
SOURCES = [os.path.join(CSRC, "qpmpc_b200.cu")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libqpmpc_b200.so")


def _deps():
    out = list(SOURCES)
    for name in os.listdir(CSRC):
        if name.endswith((".cuh", ".h")):
            out.append(os.path.join(CSRC, name))
    out.append(os.path.join(os.path.dirname(HERE), "include", "qpmpc_b200.h"))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in _deps())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
