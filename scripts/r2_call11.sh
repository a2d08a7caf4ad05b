#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/sweep_costs.py 2>&1 | tee gpurun_out/r2n_costs.log
timeout 600 python scripts/ab_kernel.py pipsort_b200/lib/libpipsort_b200.so 2>&1 | tee gpurun_out/r2n_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plans.py tests/test_gpu_p2p.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2n_tests.log
