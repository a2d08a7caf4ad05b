#!/bin/bash
# A/B of the exhaustive kernel builds first (fast), then the whole GPU suite on the new build
mkdir -p gpurun_out
timeout 600 python scripts/ab_kernel.py pipsort_b200/lib/var_base.so pipsort_b200/lib/libpipsort_b200.so 2>&1 | tee gpurun_out/r2i_ab.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2i_tests.log
