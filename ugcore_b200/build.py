"""In-tree build of the native libraries (no JIT cache: the .so files travel with gpurun).

  lib/libug4b200.so       CUDA kernels + C ABI (include/ug4b200.h), nvcc sm_100a
  lib/libug4b200_host.so  host-side mirror of ugcore's operator API + descriptor C ABI
                          (include/ug4b200_solver.h), g++ linked against libug4b200.so
  lib/libug4synth.so      synthetic problem generator (host, OpenMP)
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIB = os.environ.get("UG4B200_LIBDIR", os.path.join(ROOT, "lib"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CUDA_SOURCES = ["ctx.cu", "comm.cu", "batch.cu", "kernels/blas1.cu", "kernels/spmv.cu", "kernels/smoothers.cu"]
# -fmad=false: no FMA contraction, so SpMV / Jacobi / AXPY are bit-identical to ugcore's CPU
# algebra (reference release flags have no -march / -ffast-math, cmake/ug/debug.cmake:76-88)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp,-O2", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-6000:] + r.stderr[-6000:])
        raise RuntimeError("build failed: " + " ".join(cmd[:3]) + " ...")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout[-3000:], r.stderr[-3000:])


def _all_files(d, exts):
    out = []
    for base, _, files in os.walk(d):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


def build(force: bool = False, verbose: bool = False) -> None:
    os.makedirs(LIB, exist_ok=True)
    headers = _all_files(CSRC, (".h", ".cuh")) + _all_files(os.path.join(ROOT, "..", "include"), (".h",))

    synth = os.path.join(LIB, "libug4synth.so")
    ssrc = [os.path.join(CSRC, "synth", "synth.cpp")]
    if force or _newer(synth, ssrc + headers):
        _run(["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-w", *ssrc, "-o", synth], verbose)

    dev = os.path.join(LIB, "libug4b200.so")
    dsrc = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    if force or _newer(dev, dsrc + headers):
        # one object per translation unit, compiled concurrently, then one device link
        from concurrent.futures import ThreadPoolExecutor
        extra = os.environ.get("UG4B200_EXTRA_NVCC", "").split()  # kernel-tuning experiments only
        objdir = os.path.join(LIB, "obj")
        os.makedirs(objdir, exist_ok=True)
        cflags = [f for f in NVCC_FLAGS if f != "-shared"]
        objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in dsrc]
        todo = [(s, o) for s, o in zip(dsrc, objs) if force or extra or _newer(o, [s] + headers)]
        with ThreadPoolExecutor(max_workers=max(1, len(todo))) as ex:
            list(ex.map(lambda so: _run([NVCC, *cflags, *extra, "-c", so[0], "-o", so[1]], verbose), todo))
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", dev, "-ldl", "-lgomp"], verbose)

    host = os.path.join(LIB, "libug4b200_host.so")
    hsrc = [os.path.join(CSRC, "solver_capi.cpp")]
    if force or _newer(host, hsrc + headers + [dev]):
        _run(["g++", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-Wall",
              "-I" + os.path.join(os.path.dirname(os.path.dirname(NVCC)), "include"),   # nvtx3 (header-only)
              *hsrc, "-o", host,
              "-L" + LIB, "-lug4b200", "-Wl,-rpath,$ORIGIN"], verbose)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
