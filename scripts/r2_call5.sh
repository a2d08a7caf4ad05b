#!/bin/bash
# 2 GPUs: p2p tests (one rank per GPU), NCCL test, bench at N=2, and N=1 for the same build
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_p2p.py -m gpu -x -q 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo rc=$?; tail -2 gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench_n2.json"))
print("N=2 value", d["value"], "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
print("e2e batch ms", d["e2e"]["ms_per_step"], "single", d["e2e"]["single_locus_call"]["ms_per_step"])
for k in ("saturating","A300c2_p0.25","A300c2_p0.75"):
    print(k, d[k]["ms_per_step"], d[k]["roofline"]["kernel_ms"])
s=d["D5000c5_sss"]; print(s["ms_per_neighbourhood"], s["value"])
PY
