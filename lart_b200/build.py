"""Builds the two shared libraries of the package, in-tree.

  liblart_gpu.so   nvcc -gencode arch=compute_100a,code=sm_100a  (CUDA engine + C ABI)
  liblart_host.so  g++                                            (C++ mini-host)

nvcc cross-compiles without a GPU, so this runs in the CPU container; the built
.so files travel to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA engine cannot be built (no CPU fallback exists)")
    return p


def build_gpu(force=False, verbose=False):
    out = os.path.join(HERE, "liblart_gpu.so")
    srcs = [os.path.join(CSRC, f) for f in ("lart_engine.cu", "lart_device.cuh", "voigt_tables.cuh")] + \
           [os.path.join(INCLUDE, "lart_gpu.h")]
    if force or _stale(out, srcs):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", out, os.path.join(CSRC, "lart_engine.cu")]
        subprocess.check_call(cmd)
    return out


def build_host(force=False):
    out = os.path.join(HERE, "liblart_host.so")
    srcs = [os.path.join(CSRC, "lart_host.cpp"), os.path.join(INCLUDE, "lart_host.h"), os.path.join(INCLUDE, "lart_gpu.h")]
    if force or _stale(out, srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", out,
                               os.path.join(CSRC, "lart_host.cpp")])
    return out


def build_all(force=False):
    return build_host(force), build_gpu(force)
