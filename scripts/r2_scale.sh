#!/bin/bash
# bench at N = 8 and N = 4 on one box (torchrun, one rank per GPU); JSON lines -> gpurun_out/
mkdir -p gpurun_out
for n in 8 4; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > gpurun_out/r2h_bench_n$n.json 2> gpurun_out/r2h_bench_n$n.err; echo "N=$n rc=$?"; tail -2 gpurun_out/r2h_bench_n$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2h_bench_n$n.json"))
print("N=$n value", d["value"], "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
print("e2e batch ms", d["e2e"]["ms_per_step"], "single", d["e2e"]["single_locus_call"]["ms_per_step"])
for k in ("saturating","A300c2_p0.25","A300c2_p0.75"):
    print(k, d[k]["ms_per_step"], d[k]["roofline"]["kernel_ms"])
s=d["D5000c5_sss"]; print(s["ms_per_neighbourhood"], s["value"], s["ms_with_nccl_allgather_of_the_values"], s["search"])
PY
done
