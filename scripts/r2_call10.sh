#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/ab_kernel.py pipsort_b200/lib/var_v8.so pipsort_b200/lib/libpipsort_b200.so 2>&1 | tee gpurun_out/r2m_ab.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2m_tests.log
