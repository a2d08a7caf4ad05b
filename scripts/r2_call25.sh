#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plans.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2y_tests.log
timeout 600 python scripts/ab_kernel.py pipsort_b200/lib/var_nst1.so pipsort_b200/lib/libpipsort_b200.so 2>&1 | tee gpurun_out/r2y_ab.log
