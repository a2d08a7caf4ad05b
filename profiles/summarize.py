#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into the text summary committed under profiles/.

usage: python profiles/summarize.py gpurun_out/prof.ncu-rep regex:exhaustive_reg  > profiles/<name>.txt
"""
import collections
import csv
import subprocess
import sys

import json
import os

if sys.argv[1] == "--traffic":
    # python profiles/summarize.py --traffic WORKLOAD gpurun_out/x.ncu-rep kernel_substring profiles/<summary>.txt
    #   -> records dram__bytes_read.sum + dram__bytes_write.sum of the first matching launch in profiles/traffic.json
    _, _, wl, rep, ksub, src = sys.argv[:6]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in data:
        if ksub in r[ik]:
            b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
            try:
                with open(path) as f:
                    tab = json.load(f)
            except Exception:
                tab = {}
            tab[wl] = {"bytes": int(round(b)), "source": src, "kernel": r[ik].split("(")[0]}
            with open(path, "w") as f:
                json.dump(tab, f, indent=1, sort_keys=True)
            print(wl, tab[wl])
            break
    sys.exit(0)

rep, kern = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_red.sum", "smsp__inst_executed_op_global_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__waves_per_multiprocessor", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__inst_executed_pipe_fp64.sum", "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
print(f"# ncu --set full --clock-control none summary of {rep}")
for r in data:
    print("-" * 100)
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"{w:70s} {r[i]:>24s} {units[i]}")
    print("warp stall reasons (per issue-active cycle):")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = r[i]
            try:
                if float(v) >= 0.02:
                    print(f"    {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {float(v):.3f}")
            except ValueError:
                pass
if kern:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if hidx:
        h = rows[hidx[0]]
        end = hidx[1] - 1 if len(hidx) > 1 else len(rows)
        ci, si, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        ops, samp, tot = collections.Counter(), collections.Counter(), 0
        for r in rows[hidx[0] + 1:end]:
            try:
                c = int(r[ci])
            except (ValueError, IndexError):
                continue
            parts = r[si].split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            op = op.split(".")[0]
            ops[op] += c; samp[op] += int(r[ss]); tot += c
        print("-" * 100)
        print(f"SASS opcode mix of the first captured launch matching {kern} (warp-level instructions executed): total {tot}")
        for op, c in ops.most_common(30):
            print(f"    {op:12s} {c:>16d}  {100.0 * c / tot:6.2f} %   stall samples {samp[op]}")
