#!/bin/bash
# ncu captures of the dominant kernels of the current build (one launch each, after the plain run has exited 0)
mkdir -p gpurun_out
tag=${1:-r2_v1}
timeout 120 python scripts/prof_one.py 1500 3 > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:exhaustive_all --launch-skip 2 -c 1 -o gpurun_out/${tag}_b1500 python scripts/prof_one.py 1500 3 > gpurun_out/${tag}_b1500.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:exhaustive_all --launch-skip 4 -c 1 -o gpurun_out/${tag}_b150 python scripts/prof_one.py 150 6 > gpurun_out/${tag}_b150.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
