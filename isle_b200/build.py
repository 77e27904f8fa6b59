"""Builds isle_b200/libisle_cuda.so in-tree with nvcc for sm_100a (no JIT cache, no torch
extension machinery: the product is a plain C-ABI shared library).

    python -m isle_b200.build [--force] [--no-nccl]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libisle_cuda.so")
SOURCES = ["capi.cu", "threshold.cu", "spmm.cu", "spmm_head.cu", "spmm_head_i8.cu", "blockks.cu", "panel_tc.cu", "kmeans.cu", "dist_tc.cu", "lloyd_full.cu", "catchwords.cu", "topic_model.cu", "coll.cu", "ingest.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _newer(src_files, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_files)


def build(force: bool = False, with_nccl: bool = True, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "isle_cuda.h"))
    flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
             "-Xptxas", "-v" if verbose else "-O3"] + ARCH
    if with_nccl:
        flags.append("-DISLE_WITH_NCCL")
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            jobs.append([_nvcc(), *flags, "-c", src, "-o", obj])
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
                if res.returncode != 0:
                    raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if jobs or force or _newer(objs, LIB):
        link = [_nvcc(), "-shared", *ARCH, "-o", LIB, *objs, "-lcublas", "-lcusolver"]
        if with_nccl:
            # soname libnccl.so.2: inside a torch process the already-loaded (bundled) NCCL
            # satisfies it; stand-alone the system library does.
            link += ["-lnccl"]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, with_nccl="--no-nccl" not in sys.argv, verbose="-v" in sys.argv))
