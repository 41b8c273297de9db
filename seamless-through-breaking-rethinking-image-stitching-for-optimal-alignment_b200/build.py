"""Build libstitchb200.so (sm_100a only) from csrc/*.cu with nvcc, in-tree.

The shared object lands in `<package>/lib/libstitchb200.so`; it is git-ignored
but travels to the GPU box with the gpurun snapshot. Incremental: a .cu is
recompiled only when it (or any header) is newer than its object file.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(LIB_DIR, "libstitchb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                 # one rounding per fp32 op unless fmaf is written out
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libstitchb200 cannot be built (no CPU fallback exists)")


def _newest_header_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(PKG_DIR, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_m = _newest_header_mtime()
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = []
    objs = []
    for src in sources:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(op)
        if force or not os.path.exists(op) or os.path.getmtime(op) < max(os.path.getmtime(sp), hdr_m):
            jobs.append((sp, op))

    def compile_one(job):
        sp, op = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", sp, "-o", op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {sp}:\n{r.stdout}\n{r.stderr}")
        log = os.path.join(OBJ_DIR, os.path.basename(sp)[:-3] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return op

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
