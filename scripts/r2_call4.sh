#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_sss.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
tail -15 gpurun_out/r2e_tests.log
timeout 700 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
print("e2e batch ms", d["e2e"]["ms_per_step"], "single", d["e2e"]["single_locus_call"]["ms_per_step"])
for k in ("saturating","A300c2_p0.25","A300c2_p0.75"):
    print(k, d[k]["ms_per_step"], d[k]["roofline"]["frac"], d[k]["roofline"]["kernel_ms"])
s=d["D5000c5_sss"]; print(s["ms_per_neighbourhood"], s["value"], s["roofline"]["frac"], s["ld_gather"]["gbs"], s["search"], s["create_s"])
print(d["cpu_baseline"]["value"], d["clocks"])
PY
