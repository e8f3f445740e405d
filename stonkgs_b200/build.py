"""Build libstk.so (the sm_100a kernel library) in-tree with nvcc.

    python -m stonkgs_b200.build [--force] [--debug]

nvcc cross-compiles for sm_100a without a GPU.  Objects go to ``stonkgs_b200/csrc/_build`` and the
shared library to ``stonkgs_b200/libstk.so`` (git-ignored, but shipped to the GPU box by gpurun).
Only translation units whose sources (or headers) changed are recompiled.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libstk.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """``debug``: compile the bring-up instrumentation of the GEMM in (STK_GEMM_DEBUG timelines for tools/gemm_dbg.py);
    the default build carries none of it."""
    os.makedirs(BUILD, exist_ok=True)
    if debug and "-DSTK_GEMM_DEBUG_BUILD=1" not in NVCC_FLAGS:
        NVCC_FLAGS.append("-DSTK_GEMM_DEBUG_BUILD=1")
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "stk.h"))
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in sources:
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([path] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((path, obj, stamp, dig))

    def compile_one(job):
        path, obj, stamp, dig = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {path}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        with open(stamp, "w") as f:
            f.write(dig)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB) or force:
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
