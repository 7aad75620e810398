"""Build libf184.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

    python -m final184_b200.csrc.build [--force] [--verbose]

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only; no PTX fallback for other parts, no multi-arch
  -lineinfo                                 ncu source page maps to these files
  -fmad=false (FAITHFUL units)              the reference-faithful kernels are bit-exact against the CPU
                                            oracle, which needs +,-,* to round separately (f184_detmath.h)
  default fmad (FAST units)                 the north-star kernels are tolerance-checked and keep FMA
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libf184.so")
OBJ = os.path.join(HERE, "_obj")

FAITHFUL = ["f184_api.cu", "mode_r_voxelize.cu", "mode_r_trace.cu", "gtao.cu", "blur.cu", "debug_hooks.cu",
            "mode_n_voxelize.cu", "mode_n_inject.cu", "mode_n_mips.cu", "lighting.cu", "composite.cu"]
FAST = [f for f in sorted(os.listdir(HERE)) if f.endswith(".cu") and f not in FAITHFUL]
HEADERS = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))] + \
          [os.path.join(HERE, "..", "..", "include", "f184.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin", "/usr/bin/g++",
          "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, extra, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(HERE, src)
    if not _stale(obj, [path, os.path.abspath(__file__)] + HEADERS):
        return obj, ""
    cmd = [NVCC, *COMMON, *extra, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr if verbose else ""


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    jobs = [(s, ["-fmad=false"]) for s in FAITHFUL] + [(s, []) for s in FAST]
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        res = list(ex.map(lambda j: _compile(j[0], j[1], verbose), jobs))
    objs = [o for o, _ in res]
    for _, log in res:
        if log:
            print(log)
    if force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", *objs, "-o", OUT]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
