"""Builds ``lib/libqpmpc_b200.so`` (the C-ABI CUDA library) in-tree with nvcc.

``python -m qpmpc_b200.build [--force] [-v]``.  sm_100a only; nvcc
cross-compiles without a GPU.  Every ``csrc/*.cu`` is compiled to an object
file in parallel (the warp kernels are instantiated in one small translation
unit per dtype and lane-group width) and linked into one shared library that
depends on nothing but the statically linked CUDA runtime, so it loads (and
exports its symbols) on a CPU-only box too.
"""

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(os.path.dirname(HERE), ".scratch", "obj")  # git- and gpurun-ignored
LIB_PATH = os.path.join(LIB_DIR, "libqpmpc_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17", "-Xcompiler", "-fPIC",
]
# tuning experiments: QPMPC_MINB=<resident CTAs/SM> overrides the kernels' default
for _macro in ("QPMPC_MINB", "QPMPC_MINB_F32", "QPMPC_MINB_PAIRED", "QPMPC_THREADS_PAIRED", "QPMPC_SYNC_TAIL", "QPMPC_LR_MINB", "QPMPC_LR_MINB32", "QPMPC_LR_MINB16", "QPMPC_SEG_REDUX", "QPMPC_MINB_PRE"):
    if os.environ.get(_macro):
        NVCC_FLAGS.append(f"-D{_macro}=" + os.environ[_macro])


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libqpmpc_b200.so")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "qpmpc_b200.h"))
    return out


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    built = os.path.getmtime(target)
    return any(os.path.getmtime(d) > built for d in deps)


def is_stale() -> bool:
    return _stale(LIB_PATH, sources() + _headers())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc, hdrs = _nvcc(), _headers()
    todo = [s for s in sources() if force or _stale(_obj(s), [s] + hdrs)]

    def compile_one(src):
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", _obj(src), src]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    workers = max(1, min(len(todo), os.cpu_count() or 1))
    if todo:
        with ThreadPoolExecutor(workers) as pool:
            for src, r in pool.map(compile_one, todo):
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"== {os.path.basename(src)}\n{r.stdout}{r.stderr}")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH,
                    *[_obj(s) for s in sources()]], check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
