#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/batch_threads.py 80 2>&1 | tail -4 | tee gpurun_out/r2x_batch80.log
PIPSORT_BENCH_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-side --no-sat 2>&1 >/dev/null | grep -E "batch call" | tee gpurun_out/r2x_bench_dbg.log
