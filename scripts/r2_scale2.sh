#!/bin/bash
# the headline pass at N GPUs without the side workloads: bash scripts/r2_scale2.sh N tag
mkdir -p gpurun_out
n=$1; tag=$2
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu --no-side > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err; echo "N=$n rc=$?"; tail -2 gpurun_out/${tag}_n$n.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_n$n.json"))
print("N=$n value", d["value"], "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
print("e2e batch ms", d["e2e"]["ms_per_step"], "single", d["e2e"]["single_locus_call"]["ms_per_step"])
print("saturating", d["saturating"]["ms_per_step"], d["saturating"]["roofline"]["kernel_ms"])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n scripts/mgpu_phases.py 2>&1 | grep -E "rank [0-7]:" | sort | awk 'NR%6==1' | tee gpurun_out/${tag}_phases_n$n.log
