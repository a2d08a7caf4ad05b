#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/ab_kernel.py pipsort_b200/lib/var_dirtyflush.so pipsort_b200/lib/libpipsort_b200.so 2>&1 | tee gpurun_out/r2r_ab.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2r_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("e2e batch ms", d["e2e"]["ms_per_step"], "single", d["e2e"]["single_locus_call"]["ms_per_step"])
for k in ("saturating","A300c2_p0.25","A300c2_p0.75"):
    print(k, d[k]["ms_per_step"], d[k]["roofline"]["kernel_ms"])
s=d["D5000c5_sss"]; print(s["ms_per_neighbourhood"], s["value"])
PY
