"""Build libomvg_b200.so (sm_100a only) in-tree with nvcc.  No JIT cache, no torch dependency."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libomvg_b200.so")
SOURCES = ["match.cu", "ba.cu", "io.cu", "geom.cu"]
# geom.cu mirrors a CPU control flow whose comparisons must see the same roundings: no FMA contraction there
EXTRA_FLAGS = {"geom.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libomvg_b200.so cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(PKG), "include", "omvg_b200.h"))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(CSRC, src[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + EXTRA_FLAGS.get(src, []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            jobs.append(cmd)
        objs.append(o)
    if jobs:                                                # the translation units compile side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for _ in ex.map(subprocess.check_call, jobs):
                pass
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lgomp"]   # cudart is linked statically (nvcc default)
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
