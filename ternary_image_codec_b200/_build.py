"""Builds ``ternary_image_codec_b200/libt3c.so`` in-tree with nvcc for sm_100a.

The CUDA sources are compiled ahead of time (no JIT cache): the .so travels with the tree.
``python -m ternary_image_codec_b200._build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "_obj")
LIB = os.path.join(PKG, "libt3c.so")
CU = ["k_general.cu", "k_fast.cu", "k_formats.cu", "api.cu"]
CPP = ["tables.cpp"]
HDRS = ["dev.cuh", "k_super.cuh", "k_fast5.cuh", "launch.h", "t3c_internal.h", os.path.join("..", "..", "include", "t3c.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"] + os.environ.get("T3C_NVCC_EXTRA", "").split()


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _stale(target, deps):
    return not os.path.exists(target) or os.path.getmtime(target) < _newest(deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HDRS] + [os.path.abspath(__file__)]
    jobs = []
    for src in CU + CPP:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        if force or _stale(o, [s] + hdrs):
            extra = ["-Xptxas", "-v"] if verbose and src.endswith(".cu") else []
            jobs.append((o, [NVCC] + NVCC_FLAGS + extra + ["-c", s, "-o", o]))
    if jobs:
        def run(job):
            r = subprocess.run(job[1], capture_output=True, text=True)
            return job, r
        with ThreadPoolExecutor(max_workers=4) as ex:
            for job, r in ex.map(run, jobs):
                if verbose or r.returncode:
                    sys.stderr.write(r.stdout + r.stderr)
                if r.returncode:
                    raise RuntimeError("nvcc failed: " + " ".join(job[1]))
    objs = [os.path.join(OBJ, os.path.splitext(s)[0] + ".o") for s in CU + CPP]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
