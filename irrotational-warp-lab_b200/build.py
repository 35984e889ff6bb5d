"""In-tree build of libiwb200.so (hand-written sm_100a CUDA behind a C ABI).

    python irrotational-warp-lab_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  Each translation unit is compiled separately (in
parallel) with ``-gencode arch=compute_100a,code=sm_100a -lineinfo``; the exact-mode
kernels get ``-fmad=false`` so that no multiply-add is contracted (IEEE replay of the
reference's NumPy operation order).  The library links only against the static CUDA
runtime: no torch, no cuBLAS.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# IWB_VARIANT=name + IWB_EXTRA_FLAGS="-DX=1 ..." build libiwb200_name.so beside the product library (A/B timing of
# kernel variants on one GPU box, selected at run time with IWB_LIB; development aid)
VARIANT = os.environ.get("IWB_VARIANT", "")
OBJ = os.path.join(HERE, "_obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libiwb200" + ("_" + VARIANT if VARIANT else "") + ".so")
EXTRA = os.environ.get("IWB_EXTRA_FLAGS", "").split()
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
          "-I", os.path.join(os.path.dirname(HERE), "include")]

# translation unit -> extra flags
UNITS = {
    "core.cu": [],
    "energy3d.cu": ["-fmad=false"],
    "energy3d_fast.cu": ["-fmad=true"],
    "planar.cu": ["-fmad=false"],
    "einstein.cu": ["-fmad=false"],
}


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", "iwb200.h"), os.path.abspath(__file__)]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(unit, flags, verbose):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
    cmd = [NVCC, *ARCH, *COMMON, *flags, *EXTRA, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    units = {u: f for u, f in UNITS.items() if os.path.exists(os.path.join(CSRC, u))}
    deps = _deps()
    objs = [os.path.join(OBJ, u.replace(".cu", ".o")) for u in units]
    todo = [u for u, o in zip(units, objs) if force or _stale(o, deps)]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=len(todo)) as ex:
            list(ex.map(lambda u: _compile(u, units[u], verbose), todo))
    if todo or force or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
