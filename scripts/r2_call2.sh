#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
tail -15 gpurun_out/r2b_tests.log
python scripts/prof_sss.py 5000 > gpurun_out/r2b_prof_sss.log 2>&1; tail -4 gpurun_out/r2b_prof_sss.log
PIPSORT_SCORE_WARP=1 python scripts/prof_sss.py 5000 > gpurun_out/r2b_prof_sss_warp.log 2>&1; tail -3 gpurun_out/r2b_prof_sss_warp.log
PIPSORT_TRACE_PREP=1 timeout 600 python scripts/sss_d.py 5000 30 5 raw > gpurun_out/r2b_sss_d.log 2>&1; tail -20 gpurun_out/r2b_sss_d.log
