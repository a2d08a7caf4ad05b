"""Builds a variant of the engine library for A/B runs (scripts/ab_kernel.py):
python scripts/build_variant.py NAME [-DMACRO=VALUE ...] [--csrc DIR]  ->  pipsort_b200/lib/var_NAME.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pipsort_b200 import build as B
name = sys.argv[1]
defs = [a for a in sys.argv[2:] if a.startswith("-D")]
csrc = sys.argv[sys.argv.index("--csrc") + 1] if "--csrc" in sys.argv else B.CSRC
out = os.path.join(B.LIBDIR, "var_%s.so" % name)
subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + defs + ["-I", os.path.join(ROOT, "include"), os.path.join(csrc, "engine.cu"), "-o", out])
print(out)
