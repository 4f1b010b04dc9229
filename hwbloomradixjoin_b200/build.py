"""Compile the CUDA library for sm_100a in-tree (libhwbrj_cuda.so next to this file) and the C host driver."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libhwbrj_cuda.so")
SRC = [os.path.join(HERE, "csrc", f) for f in ("hwbrj.cu", "kernels.cuh", "hash.cuh")]
HDR = os.path.join(ROOT, "include", "hwbrj.h")
DRIVER_SRC = os.path.join(ROOT, "host", "mchashjoins_gpu.c")
DRIVER = os.path.join(ROOT, "build", "mchashjoins_gpu")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: cannot build libhwbrj_cuda.so")
    return p


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, SRC + [HDR]):
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB, SRC[0], "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(" ".join(cmd))
    return LIB


def build_driver(force: bool = False) -> str | None:
    """C host driver (mchashjoins-compatible CLI) linked against the C ABI."""
    if not os.path.exists(DRIVER_SRC):
        return None
    os.makedirs(os.path.dirname(DRIVER), exist_ok=True)
    if force or _stale(DRIVER, [DRIVER_SRC, HDR, LIB]):
        cmd = ["gcc", "-O2", "-std=c11", "-D_GNU_SOURCE", "-I", os.path.join(ROOT, "include"), "-o", DRIVER, DRIVER_SRC,
               "-L", HERE, "-lhwbrj_cuda", "-Wl,-rpath," + HERE, "-lm", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc (host driver) failed:\n" + r.stdout + r.stderr)
    return DRIVER


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_driver(force=True))
