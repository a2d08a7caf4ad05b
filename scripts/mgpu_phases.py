"""Per-phase device timings of the sharded pass (run under torchrun with 2+ ranks): where does the step time go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import pipsort_b200 as P
from pipsort_b200 import synth, distributed as D

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
L = synth.make_locus(150)
c = 3
e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param, max_causal=c, device=local)
e.set_stream(st.cuda_stream)
ok = D.connect_p2p(e)
b = e.shard_ranks(c, world)
def ev(): return torch.cuda.Event(enable_timing=True)
def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
# warm
for _ in range(3):
    e.reset(); e.run_exhaustive(c, b[rank], b[rank + 1]); e.p2p_reduce_to_root(); e.finalize()
sync()
rows = []
for it in range(8):
    e.flush_l2()
    t = [ev() for _ in range(5)]
    t[0].record(st); e.reset(); t[1].record(st)
    e.run_exhaustive(c, b[rank], b[rank + 1]); t[2].record(st)
    e.p2p_reduce_to_root(); t[3].record(st)
    e.finalize(); t[4].record(st)
    sync()
    rows.append([t[i].elapsed_time(t[i + 1]) * 1e3 for i in range(4)])
print(f"rank {rank} p2p={ok} shard={b} phases us (reset, exhaustive, combine, finalize):", flush=True)
for r in rows[2:]:
    print(f"  rank {rank}: " + " ".join(f"{x:8.1f}" for x in r), flush=True)
# pure combine latency: nothing else on the stream
sync()
lat = []
for it in range(10):
    a, z = ev(), ev()
    a.record(st); e.p2p_reduce_to_root(); z.record(st)
    sync()
    lat.append(a.elapsed_time(z) * 1e3)
print(f"rank {rank} combine alone us: " + " ".join(f"{x:.1f}" for x in lat), flush=True)
# NCCL all-reduce of the store alone
acc = e.accumulator_tensor()
for _ in range(3): dist.all_reduce(acc)
sync()
lat = []
for it in range(10):
    a, z = ev(), ev()
    a.record(st); dist.all_reduce(acc); z.record(st)
    sync()
    lat.append(a.elapsed_time(z) * 1e3)
print(f"rank {rank} nccl all-reduce alone us: " + " ".join(f"{x:.1f}" for x in lat), flush=True)
# graph replay timing without L2 flush
e.graph_begin(); e.reset(); e.run_exhaustive(c, b[rank], b[rank + 1]); e.p2p_reduce_to_root(); e.finalize(); gid = e.graph_end()
for flush in (False, True):
    for _ in range(3): e.graph_launch(gid)
    sync()
    tt = []
    for it in range(10):
        if flush: e.flush_l2()
        a, z = ev(), ev()
        a.record(st); e.graph_launch(gid); z.record(st)
        tt.append((a, z))
    sync()
    print(f"rank {rank} graph step us (flush={flush}): " + " ".join(f"{a.elapsed_time(z) * 1e3:.1f}" for a, z in tt), flush=True)
e.close()
dist.destroy_process_group()
