#!/usr/bin/env python
"""bench.py -- causal configurations / second of the exhaustive posterior calculation (BASELINE.json metric).

A "step" is one pass of the hot path (PostCal::computeTotalLikelihood, postcal.cpp:716-1092) over one
synthetic locus: reset accumulators, enumerate + score every union subset of size <= c in this rank's
shard of the rank space, combine the accumulator stores of all ranks with ONE NCCL all-reduce(sum),
finalize (bins -> log-space results).

  value     whole-job configurations/s with the locus already resident in HBM, device-timed (CUDA events on
            the stream everything is launched on), max over ranks
  e2e       the same metric through the public API with HOST buffers: engine creation (H2D of LD / z / maps),
            the pass, and the D2H read of the result arrays inside the timed region
  roofline  FP64 vector-pipe roofline of the dominant kernel (the size-c class launch): algorithmic flops
            (SURVEY.md 8d formula, counted exactly by class) / its device time / the DFMA peak measured here
  cpu_baseline   the reference's own OpenMP CPU build (oracle/_ref/PIPSORT, unmodified sources) on this host,
            on a bounded sample of the same workload (the size <= 2 prefix of the rank space, `-c 2`);
            falls back to the OpenMP port of the oracle when the reference binary is not present

python bench.py --impl reference ... times only the reference CPU implementation and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (SNPs per study, overlap, c)   -- BASELINE.json configs[3] / configs[2] / the saturating B10 of SURVEY 8d
    "B150c3": (150, 0.8, 3),
    "A300c2": (300, 0.8, 2),
    "B1500c3": (1500, 0.8, 3),
}
# dram__bytes_read.sum + dram__bytes_write.sum of one exhaustive_all_kernel launch, from the `ncu --set full` captures
# summarised in profiles/r1_v6_exhaustive_all_*.txt (the loci are L2 resident: LD and the pair tables are read from HBM once)
TRAFFIC_BYTES = {"B150c3": 790528 + 0, "B1500c3": 39675904 + 940800}
METRIC = "causal configurations/sec (exhaustive, c=3, synthetic 150-SNP/study two-ancestry locus)"
UNIT = "configs/s"


def class_flops(snp_map, c):
    """(algorithmic flops, configurations) of the size-c class and of the whole run."""
    from pipsort_b200 import synth
    tot, ntot = synth.flops_per_config_total(snp_map, c)
    low, nlow = synth.flops_per_config_total(snp_map, c - 1) if c > 0 else (0.0, 0)
    return (tot - low, ntot - nlow), (tot, ntot)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    t = [x.strip() for x in line.split(",")]
                    if len(t) < 9:
                        continue
                    try:
                        sm.append(float(t[1])); mx.append(float(t[2]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# CPU arms (test infrastructure: the only place bench.py executes anything under oracle/)
# ---------------------------------------------------------------------------------------------------------
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "PIPSORT")


def reference_cpu_run(L, c_sample, workdir):
    """One run of the UNMODIFIED reference CLI (exhaustive, `-c c_sample`) on the locus; returns
    (configurations, seconds of its own 'Time to eval all' timer, wall seconds)."""
    from pipsort_b200 import synth
    ld, z, mp, ns = synth.write_files(L, os.path.join(workdir, "in"))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    t = time.time()
    p = subprocess.run([REF_BIN, "-l", ld, "-z", z, "-m", mp, "-n", ns, "-o", os.path.join(workdir, "out"), "-c",
                        str(c_sample), "-g", repr(L.gamma), "-p", repr(L.sharing_param)], capture_output=True,
                       text=True, env=env, cwd=workdir)
    wall = time.time() - t
    if p.returncode != 0:
        raise RuntimeError("reference CLI failed: " + p.stderr[-500:])
    m = re.search(r"Time to eval all=\s*(\d+)\[", p.stdout)
    secs = int(m.group(1)) * 1e-6 if m else wall
    return synth.count_configs(L.snp_map, c_sample), secs, wall


def port_cpu_run(L, c, seconds_target=10.0):
    """OpenMP port of the oracle on a rank sub-range sized to ~seconds_target; returns (configs, seconds, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import synth_as_oracle_locus
    from oracle import oracle as O
    OL = synth_as_oracle_locus(L)
    cores = os.cpu_count() or 1
    tot = O.total_union_subsets(OL.U, c)
    n = min(tot, 20000)
    lo = tot - n                                  # the size-c class dominates; sample from its end
    t = time.time(); _, ne = O.exhaustive_omp(OL, c, lo, tot, cores); dt = time.time() - t
    if dt < seconds_target / 4 and n < tot:       # scale the sample up once
        n = int(min(tot, n * seconds_target / max(dt, 1e-3)))
        lo = tot - n
        t = time.time(); _, ne = O.exhaustive_omp(OL, c, lo, tot, cores); dt = time.time() - t
    return ne, dt, cores, f"union-subset ranks [{lo},{tot}) of {tot}"


def cpu_baseline(L, c):
    cores = os.cpu_count() or 1
    if os.path.exists(REF_BIN):
        try:
            with tempfile.TemporaryDirectory() as tmp:
                n, secs, wall = reference_cpu_run(L, c - 1, tmp)
            return {"value": n / secs, "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": f"unmodified reference CLI `-c {c - 1}` on the same locus = the size<={c - 1} prefix of the "
                              f"rank space ({n} configurations); its own 'Time to eval all' {secs:.2f} s (wall {wall:.2f} s); "
                              f"the reference forces 64 OpenMP threads (postcal.cpp:748) on {cores} host cores"}
        except Exception as ex:  # fall through to the port
            err = str(ex)[:200]
    else:
        err = "oracle/_ref/PIPSORT not built"
    ne, dt, cores, what = port_cpu_run(L, c)
    return {"value": ne / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"OpenMP port of the oracle (closed-form O(k^3) likelihood), {what}, {ne} configurations in {dt:.2f} s ({err})"}


def run_reference_arm(args, L, c, wl):
    """bench.py --impl reference: rank 0 alone times the reference's own CPU implementation."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget = 240.0
    vals, walls = [], []
    kind = "reference" if os.path.exists(REF_BIN) else "port"
    sample = ""
    t_all = time.time()
    steps_done = 0
    for i in range(args.warmup + args.steps):
        timed = i >= args.warmup
        if kind == "reference":
            with tempfile.TemporaryDirectory() as tmp:
                n, secs, wall = reference_cpu_run(L, c - 1, tmp)
            sample = (f"unmodified reference CLI `-c {c - 1}` = size<={c - 1} prefix of the rank space, {n} configurations per step; "
                      f"64 OpenMP threads forced by the reference on {cores} cores; timed with its own 'Time to eval all'")
        else:
            n, secs, cores, what = port_cpu_run(L, c)
            sample = f"oracle OpenMP port, {what}"
        if timed:
            vals.append(n / secs); walls.append(secs); steps_done += 1
        elapsed = time.time() - t_all
        per = elapsed / (i + 1)
        if elapsed + per > budget:          # keep the whole arm within a few minutes
            if not timed:
                continue_from = args.warmup  # skip remaining warm-ups (a fresh process has nothing to warm)
                if i + 1 < continue_from:
                    args.warmup = i + 1
            elif steps_done >= 1:
                break
    if not vals:
        n, secs, wall = reference_cpu_run(L, c - 1, tempfile.mkdtemp()) if kind == "reference" else port_cpu_run(L, c)[:3]
        vals, walls, steps_done = [n / secs], [secs], 1
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_desc(wl, L, c), "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def wl_desc(wl, L, c):
    sh, o0, o1 = L.n_types()
    return (f"{wl}: synthetic two-ancestry locus, {int(L.num_snps[0])}+{int(L.num_snps[1])} SNPs, U={L.U} "
            f"({sh} shared, {o0}+{o1} study-specific), exhaustive c={c}, p={L.sharing_param}, gamma={L.gamma}, seed 20261018")


# ---------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """ONE JSON line on stdout is the contract, but libraries print there too (NCCL's version banner at communicator
    creation, for instance): point fd 1 at stderr for the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="B150c3", choices=sorted(WORKLOADS))
    ap.add_argument("--collective", default="p2p", choices=["p2p", "allreduce"],
                    help="multi-GPU combine step: the engine's peer-memory kernels (default) or one NCCL all-reduce")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sat", action="store_true", help="skip the saturating B1500c3 side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from pipsort_b200 import synth
    n_per, overlap, c = WORKLOADS[args.workload]
    L = synth.make_locus(n_per, overlap=overlap)
    if args.impl == "reference":
        run_reference_arm(args, L, c, args.workload)
        return

    import torch
    import torch.distributed as dist
    import pipsort_b200 as P
    from pipsort_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    # everything (engine kernels, NCCL all-reduce, timing events) is enqueued on ONE non-default stream
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    collective_used = []
    graph_used = []

    def measure(L, c, steps, warmup, clocks=False):
        """Device-timed steps on the resident locus; returns dict of timings."""
        total_configs = synth.count_configs(L.snp_map, c)
        e = P.Engine(L.num_snps, L.sigma, L.z, L.d, L.K, L.snp_map, gamma=L.gamma, sharing_param=L.sharing_param,
                     max_causal=c, device=local)
        stream = torch.cuda.current_stream()
        e.set_stream(stream.cuda_stream)
        b = e.shard_ranks(c, world)
        # combine step: the engine's own peer-memory kernels over NVLink (CUDA IPC mailboxes) when the ranks can map each
        # other, else ONE NCCL all-reduce(sum) of the accumulator store
        coll = "allreduce"
        if world > 1 and args.collective != "allreduce":
            coll = "p2p" if D.connect_p2p(e) else "allreduce"
        collective_used.append(coll)

        def step():
            D.run_exhaustive_sharded(e, c, bounds=b, collective=coll)   # reset + this rank's launch + the combine step
            e.finalize()

        e.flush_l2(); step()                  # first pass: builds the pair tables, allocates the scratch buffers
        l0 = e.launch_count()
        e.flush_l2(); step()
        launches_per_step = e.launch_count() - l0
        # the pass is a fixed sequence of the engine's launches: record it once into a CUDA graph and replay it with one
        # launch per step (the host cost of issuing ~6 launches is comparable to this 0.1 ms pass).  The NCCL fallback is
        # issued by torch and stays un-captured.
        run = step
        graphed = False
        if coll != "allreduce" or world == 1:
            e.graph_begin(); step(); gid = e.graph_end()
            run = lambda: e.graph_launch(gid)   # noqa: E731
            graphed = True
        graph_used.append(graphed)
        for _ in range(warmup):
            e.flush_l2(); run()
        sampler = ClockSampler(local)
        if clocks and rank == 0:
            sampler.start()                   # before the barrier: spawning nvidia-smi must not delay rank 0's first steps
        barrier()
        evs, kms = [], []
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_wall = time.perf_counter()
        for a, z in pairs:
            e.flush_l2()                      # L2 flushed between timed iterations (outside the event pair)
            a.record(stream); run(); z.record(stream)
            evs.append((a, z))
            kms.append(None)
        barrier()
        t_wall = time.perf_counter() - t_wall
        if clocks:
            # the timed region lasts a few milliseconds (20 steps of 0.07 ms): keep the SAME step running back to back for
            # another ~0.3 s, untimed, so that nvidia-smi (20 ms period) sees the clocks / throttle reasons under this load
            t_load = time.perf_counter()
            while time.perf_counter() - t_load < 0.3:
                for _ in range(50):
                    run()
                e.sync()
            barrier()
        clk = sampler.stop() if clocks and rank == 0 else None
        if clk is not None:
            clk["window"] = "timed region + 0.3 s of the same step repeated back to back (untimed), nvidia-smi every 20 ms"
        launches = launches_per_step * steps
        ms = [a.elapsed_time(z) for a, z in evs]
        if os.environ.get("PIPSORT_BENCH_DEBUG"):
            print(f"[rank {rank}] per-step ms: " + " ".join(f"{x:.4f}" for x in ms), file=sys.stderr, flush=True)
        # dominant-kernel duration: a few extra steps timed on the launch itself
        for _ in range(min(steps, 10)):
            e.flush_l2(); step(); kms.append(e.last_kernel_ms())
        kms = [k for k in kms if k is not None]
        tot_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        k_ms = torch.tensor([float(np.mean(kms))], dtype=torch.float64, device=dev)
        cnt = e.config_count()            # the count rides in the combined store: the whole job's (on the root at least)
        if world > 1:
            dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(k_ms, op=dist.ReduceOp.MAX)
        res = e.read() if rank == 0 else None
        if world > 1:
            dist.barrier()
        e.close()
        if rank == 0:
            assert cnt == total_configs, (cnt, total_configs)
        return dict(total_configs=total_configs, ms_per_step=float(tot_ms.item()) / steps, kernel_ms=float(k_ms.item()),
                    launches=launches, wall_ms_per_step=1e3 * t_wall / steps, clocks=clk, result=res)

    def measure_e2e(L, c, steps, warmup):
        """Public API with HOST (pinned) buffers: create (H2D) + pass + read (D2H) + destroy, wall clock between syncs."""
        total_configs = synth.count_configs(L.snp_map, c)
        sig = torch.from_numpy(np.concatenate([s.ravel() for s in L.sigma])).pin_memory()
        z = torch.from_numpy(np.concatenate(L.z)).pin_memory()
        h2d = sig.numel() * 8 + z.numel() * 8 + L.snp_map.size * 4 * 3 + 2 * 16
        d2h = 8 * (1 + 2 + L.N + 3 * L.U)

        sig_np, z_np = sig.numpy(), z.numpy()

        def once():
            if world == 1:      # one locus, one C-ABI call: create (H2D) + exhaustive pass + read (D2H) + destroy
                return P.posterior_exhaustive(L.num_snps, sig_np, z_np, L.d, L.K, L.snp_map, c, gamma=L.gamma,
                                              sharing_param=L.sharing_param, device=local)
            e = P.Engine(L.num_snps, sig_np, z_np, L.d, L.K, L.snp_map, gamma=L.gamma,
                         sharing_param=L.sharing_param, max_causal=c, device=local)
            if world > 1:
                D.bind_engine_to_current_stream(e)
            r = D.compute_total_likelihood_sharded(e, c)
            e.close()
            return r

        for _ in range(warmup):
            once()
        barrier()
        t = time.perf_counter()
        for _ in range(steps):
            r = once()
        barrier()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        serial_ms = 1e3 * float(dt.item()) / steps
        # The call a user with many loci makes -- a LIST of loci in one C-ABI call (pipsort_posterior_exhaustive_batch).
        # Every step (= locus) still has its own pinned-host -> device copy of the LD matrices / z / maps, its own kernels
        # and its own device -> host read of the result arrays inside the timed region; the engine overlaps the uploads and
        # preparation of locus i+1 and the read-back of locus i-1 with the evaluation of locus i (three streams).  With
        # several GPUs the LOCI are dealt out to the ranks (independent units, no collective): every rank runs the same
        # call on its own list; a single 0.1 ms locus is not worth sharding (single_locus_call_ms shows that path).
        nb = max(steps, 8) * 4
        locus = dict(num_snps=L.num_snps, sigma=sig_np, z=z_np, d=L.d, K=L.K, snp_map=L.snp_map, gamma=L.gamma,
                     sharing_param=L.sharing_param)
        P.posterior_exhaustive_batch([locus] * 8, c, device=local)
        barrier()
        t = time.perf_counter()
        rs = P.posterior_exhaustive_batch([locus] * nb, c, device=local)
        barrier()
        dtb = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dtb, op=dist.ReduceOp.MAX)
        dtb = float(dtb.item())
        assert all(x.n_configs == total_configs for x in rs)
        assert abs(rs[-1].total - r.total) <= 1e-9 * abs(r.total)
        return {"value": total_configs * nb * world / dtb, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * dtb / (nb * world), "loci_per_call": nb, "loci_total": nb * world,
                "single_locus_call_ms": serial_ms,
                "timing": "host wall clock (max over ranks) around ONE call per rank that evaluates loci_per_call loci from pinned "
                          "host buffers (per locus: H2D of LD / z / maps, preparation + exhaustive + finalize kernels, D2H of the "
                          "result arrays; three loci in flight on three streams per GPU; loci dealt out to the ranks, no "
                          "collective); single_locus_call_ms = one locus per call"
                          + (" with its rank space sharded over the GPUs and one NCCL all-reduce" if world > 1 else "")}

    peak = P.measure_fp64_peak(local)                      # FLOP/s, DFMA chains on every SM
    main_m = measure(L, c, args.steps, args.warmup, clocks=True)
    e2e = measure_e2e(L, c, max(3, min(args.steps, 10)), 2)
    (cf, cn), (tf, tn) = class_flops(L.snp_map, c)
    value = main_m["total_configs"] / (main_m["ms_per_step"] * 1e-3)
    # the dominant launch on rank r covers 1/world of the class (work-weighted shard)
    achieved = cf / world / (main_m["kernel_ms"] * 1e-3) / 1e12
    roof = {"bound": "fp64", "achieved": achieved, "peak": peak / 1e12, "unit": "TFLOP/s", "frac": achieved / (peak / 1e12),
            "traffic": TRAFFIC_BYTES.get(args.workload) if world == 1 else None,
            "note": ("FP64 vector pipe (DFMA), no tensor cores / not HBM bound; achieved = ALGORITHMIC flops of the size-c class "
                     f"({cf:.4g} flop, {cf / max(cn, 1):.1f}/configuration, SURVEY.md 8d) per launch / its CUDA-event duration "
                     f"({main_m['kernel_ms']:.4f} ms); peak = DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 "
                     "figure; nominal 37 TFLOP/s). The kernel shares Cholesky work between the expansions of a union subset, so it "
                     "executes fewer flops than the algorithmic count: frac can exceed 1; executed FP64 pipe utilisation is in profiles/.")}
    sat = None
    if not args.no_sat and args.workload == "B150c3":
        Ls = synth.make_locus(1500, overlap=0.8)
        sm = measure(Ls, 3, 3, 1)
        (scf, scn), _ = class_flops(Ls.snp_map, 3)
        sach = scf / world / (sm["kernel_ms"] * 1e-3) / 1e12
        sat = {"workload": wl_desc("B1500c3", Ls, 3), "configs": sm["total_configs"], "ms_per_step": sm["ms_per_step"],
               "value": sm["total_configs"] / (sm["ms_per_step"] * 1e-3), "unit": UNIT,
               "roofline": {"bound": "fp64", "achieved": sach, "peak": peak / 1e12, "unit": "TFLOP/s", "frac": sach / (peak / 1e12)}}
    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            cpu = cpu_baseline(L, c)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": main_m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl_desc(args.workload, L, c), "configs_per_step": main_m["total_configs"],
                           "l2": "flushed between timed steps (256 MiB memset outside the per-step event pair)",
                           "launch": ("one CUDA-graph launch per step (reset + exhaustive kernel + combine + finalize recorded once)"
                                      if graph_used and graph_used[0] else "individual launches"),
                           "sharding": (f"rank space split into {world} work-weighted contiguous ranges; combine step per step: "
                                        + ("non-root ranks add their non-zero accumulator bins into the root's memory over NVLink "
                                           "(engine kernels, CUDA IPC peer memory, device-side arrival words)"
                                           if collective_used and collective_used[0] == "p2p" else
                                           "one NCCL all-reduce(sum) of the accumulator store")) if world > 1 else "single GPU"},
                "clocks": main_m["clocks"], "e2e": e2e, "gpu_launches": main_m["launches"], "roofline": roof,
                "locus_wall_ms": e2e["ms_per_step"], "wall_ms_per_step": main_m["wall_ms_per_step"]}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if sat is not None:
            line["saturating"] = sat
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
