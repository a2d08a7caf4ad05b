#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plans.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2v_tests.log
timeout 600 python scripts/ab_kernel.py pipsort_b200/lib/var_prev.so pipsort_b200/lib/libpipsort_b200.so 2>&1 | tee gpurun_out/r2v_ab.log
timeout 300 python scripts/trace_chunks.py 150 2>&1 | tee gpurun_out/r2v_trace150.log
timeout 300 python scripts/trace_chunks.py 150 8 3 2>&1 | tee gpurun_out/r2v_trace150_shard.log
(timeout 300 python scripts/shard_times.py 150 8) 2>&1 | tail -3 | tee gpurun_out/r2v_shards.log
