#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/ab_kernel.py pipsort_b200/lib/var_v1.so pipsort_b200/lib/libpipsort_b200.so pipsort_b200/lib/var_mb4.so pipsort_b200/lib/var_w5b2.so 2>&1 | tee gpurun_out/r2l_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plans.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2l_tests.log
PIPSORT_B200_LIB=pipsort_b200/lib/var_mb4.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plans.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2l_tests_mb4.log
