"""Build ``libfemder_b200.so`` in-tree with nvcc for sm_100a (no other architecture, no JIT cache).

    python -m femder_b200.build [--force] [-v]

Every ``.cu`` file is compiled to an object file on its own (in parallel, only when it or a header is newer than its
object), then linked.  cuBLAS / cuSOLVER are linked for the modal set-up only (dense FP64 Gram products and the small
dense eigenproblem of the Rayleigh-Ritz step, ``csrc/modal.cu``); no kernel of a solver iteration is a library kernel.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# FDB_BUILD_TUNING=1: the A/B build (-DFDB_TUNING: the solver's and the modal set-up's tuning knobs become environment variables),
# next to the shipped library, never instead of it: load it with FDB_LIB_PATH
TUNING = os.environ.get("FDB_BUILD_TUNING") == "1"
OBJ = os.path.join(HERE, "_ab", "build") if TUNING else os.path.join(HERE, "build")
LIB = os.path.join(HERE, "_ab", "libfemder_b200_tuning.so") if TUNING else os.path.join(HERE, "libfemder_b200.so")
SOURCES = ["pattern.cu", "assemble.cu", "reorder.cu", "sweep.cu", "modal.cu", "comm.cu", "capi.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc.cuh"), os.path.join(CSRC, "solver.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "femder_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--fmad=true"] + (["-DFDB_TUNING"] if TUNING else [])


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _mtime(p):
    return os.path.getmtime(p) if os.path.isfile(p) else 0.0


def _obj(src):
    return os.path.join(OBJ, os.path.splitext(src)[0] + ".o")


def _stale(src):
    t = _mtime(_obj(src))
    return t == 0.0 or any(_mtime(d) > t for d in [os.path.join(CSRC, src)] + HEADERS)


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(_mtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = nvcc_path()

    def compile_one(src):
        cmd = [nvcc] + ARCH + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", _obj(src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        return r.stderr

    todo = [s for s in SOURCES if force or _stale(s)]
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        logs = list(ex.map(compile_one, todo))
    if verbose:
        for s, log in zip(todo, logs):
            print(f"==== {s}\n{log}")
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + [_obj(s) for s in SOURCES] + ["-lcublas", "-lcusolver", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
