#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/ab_kernel.py pipsort_b200/lib/var_v1.so pipsort_b200/lib/var_v4.so pipsort_b200/lib/var_v4_nodmma.so 2>&1 | tee gpurun_out/r2j_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plans.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2j_tests.log
