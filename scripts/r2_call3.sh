#!/bin/bash
# lane kernel (pair split) + chunked exhaustive planner: parity first, then timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
tail -12 gpurun_out/r2c_tests.log
timeout 100 python scripts/prof_sss.py 5000 2>&1 | tail -2
PIPSORT_B200_LIB=$PWD/pipsort_b200/lib/var_lane4.so timeout 100 python scripts/prof_sss.py 5000 2>&1 | tail -2
PIPSORT_EXH_DEBUG=1 timeout 100 python scripts/prof_one.py 150 6 2>&1 | tail -3
timeout 100 python scripts/prof_one.py 1500 3 2>&1 | tail -1
timeout 300 python scripts/sweep_chunks.py 150 300 1500 2>&1 | tail -4
