"""Builds the engine in-tree: pipsort_b200/lib/libpipsort_b200.so (nvcc, sm_100a only)."""
from __future__ import annotations

import hashlib
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


def _digest(sources, extra=""):
    """Content hash of the sources (mtimes do not survive a snapshot copy to the GPU box; contents do)."""
    h = hashlib.sha256(extra.encode())
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, sources, extra=""):
    """True when `target` is missing or was built from other sources (hash recorded next to it at build time)."""
    if not os.path.exists(target):
        return True
    try:
        with open(target + ".srchash") as f:
            return f.read().strip() != _digest(sources, extra)
    except OSError:
        return True


def _stamp(target, sources, extra=""):
    with open(target + ".srchash", "w") as f:
        f.write(_digest(sources, extra))


def sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    out.append(os.path.join(os.path.dirname(PKG), "include", "pipsort_b200.h"))
    return out


def build_engine(force: bool = False, verbose: bool = False) -> str:
    if os.environ.get("PIPSORT_B200_LIB") and not force:
        # an explicitly named build (A/B of kernel variants, scripts/ab_kernel.py) is loaded as it is
        if not os.path.exists(LIB):
            raise FileNotFoundError(LIB)
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    flags = " ".join(NVCC_FLAGS)
    if force or _stale(LIB, sources(), flags):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, "engine.cu"), "-o", LIB + ".tmp"]
        subprocess.check_call(cmd)
        os.replace(LIB + ".tmp", LIB)          # atomic: several ranks may load the library while one rebuilds it
        _stamp(LIB, sources(), flags)
    return LIB


def build_host(force: bool = False) -> str | None:
    """C++ host (PostCal / Model mirror + the PIPSORT command line) linked against the engine."""
    hdir = os.path.join(PKG, "host")
    if not os.path.isdir(hdir):
        return None
    srcs = [os.path.join(hdir, f) for f in sorted(os.listdir(hdir)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".h")] + \
        [os.path.join(os.path.dirname(PKG), "include", "pipsort_b200.h")]
    if srcs and (force or _stale(HOST_BIN, deps)):
        cmd = ["g++", "-O2", "-std=c++17", "-I", os.path.join(os.path.dirname(PKG), "include")] + srcs + \
              ["-o", HOST_BIN, "-L", LIBDIR, "-lpipsort_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
        _stamp(HOST_BIN, deps)
    return HOST_BIN if srcs else None


def build_all(force: bool = False):
    build_engine(force)
    build_host(force)


if __name__ == "__main__":
    import sys
    build_engine(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
    print(LIB)
