#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_chunks.py 150 2>&1 | tee gpurun_out/r2p_trace150.log
timeout 300 python scripts/trace_chunks.py 150 8 3 2>&1 | tee gpurun_out/r2p_trace150_shard.log
