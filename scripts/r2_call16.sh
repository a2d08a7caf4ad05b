#!/bin/bash
mkdir -p gpurun_out
for t in 1 2 3 4; do PIPSORT_BATCH_THREADS=$t timeout 200 python scripts/batch_threads.py 2>&1 | tail -4; done | tee gpurun_out/r2s_batch.log
PIPSORT_TRACE=1 PIPSORT_TRACE_CREATE=1 timeout 100 python scripts/one_call_trace.py 2>&1 | tail -6 | tee gpurun_out/r2s_one.log
