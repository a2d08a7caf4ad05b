#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_cli.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2w_tests.log
for t in 2 3; do PIPSORT_BATCH_THREADS=$t timeout 200 python scripts/batch_threads.py 2>&1 | tail -3; done | tee gpurun_out/r2w_batch.log
PIPSORT_TRACE=1 PIPSORT_TRACE_CREATE=1 timeout 100 python scripts/one_call_trace.py 2>&1 | tail -4 | tee gpurun_out/r2w_one.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-side --no-sat > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2w_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2w_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("e2e batch ms", d["e2e"]["ms_per_step"], "single", d["e2e"]["single_locus_call"]["ms_per_step"])
PY
