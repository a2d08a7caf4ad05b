#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/sweep_costs.py 2>&1 | tee gpurun_out/r2q_costs.log
timeout 600 python scripts/ab_kernel.py pipsort_b200/lib/libpipsort_b200.so pipsort_b200/lib/var_skew300.so pipsort_b200/lib/var_skew600.so 2>&1 | tee gpurun_out/r2q_ab.log
(timeout 300 python scripts/shard_times.py 1500 8; timeout 300 python scripts/shard_times.py 150 8; timeout 300 python scripts/shard_times.py 150 2) 2>&1 | tail -12 | tee gpurun_out/r2q_shards.log
