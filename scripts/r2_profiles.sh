#!/bin/bash
# evidence for profiles/: plain bench line, launch list, ncu --set full captures of the dominant kernels (current build)
mkdir -p gpurun_out
tag=${1:-r2c}
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/${tag}_bench_reference.json 2> /dev/null; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_B150c3.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-sat --no-side > /dev/null 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:exhaustive_all --launch-skip 2 -c 1 -o gpurun_out/${tag}_b1500 python scripts/prof_one.py 1500 3 > gpurun_out/${tag}_b1500.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:exhaustive_all --launch-skip 4 -c 1 -o gpurun_out/${tag}_b150 python scripts/prof_one.py 150 6 > gpurun_out/${tag}_b150.log 2>&1
PROF_NO_FLUSH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_lane --launch-skip 4 -c 1 -o gpurun_out/${tag}_lane python scripts/prof_sss.py 5000 > gpurun_out/${tag}_lane.log 2>&1
ls -la gpurun_out/${tag}_*
