"""Builds the engine in-tree: pipsort_b200/lib/libpipsort_b200.so (nvcc, sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.environ.get("PIPSORT_B200_LIB") or os.path.join(LIBDIR, "libpipsort_b200.so")
HOST_BIN = os.path.join(LIBDIR, "PIPSORT")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    out.append(os.path.join(os.path.dirname(PKG), "include", "pipsort_b200.h"))
    return out


def build_engine(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    if force or _stale(LIB, sources()):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, "engine.cu"), "-o", LIB]
        subprocess.check_call(cmd)
    return LIB


def build_host(force: bool = False) -> str | None:
    """C++ host (PostCal / Model mirror + the PIPSORT command line) linked against the engine."""
    hdir = os.path.join(PKG, "host")
    if not os.path.isdir(hdir):
        return None
    srcs = [os.path.join(hdir, f) for f in sorted(os.listdir(hdir)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".h")] + [LIB]
    if srcs and (force or _stale(HOST_BIN, deps)):
        cmd = ["g++", "-O2", "-std=c++17", "-I", os.path.join(os.path.dirname(PKG), "include")] + srcs + \
              ["-o", HOST_BIN, "-L", LIBDIR, "-lpipsort_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return HOST_BIN if srcs else None


def build_all(force: bool = False):
    build_engine(force)
    build_host(force)


if __name__ == "__main__":
    import sys
    build_engine(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
    print(LIB)
