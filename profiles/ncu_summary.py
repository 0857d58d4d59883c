#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics of the first kernel + per-source-line instruction
and stall-sample shares.  Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [minpct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "inst_executed", "sm__inst_executed.avg.per_cycle_active", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]
print("== raw metrics (first captured launch) ==")
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:80s} {units[i]:12s} {r[i]}")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
        try:
            if float(r[i]) >= 2.0:
                print(f"{h:80s} {units[i]:12s} {r[i]}")
        except ValueError:
            pass

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur_file = None
agg = {}
hdr = None
for row in rows:
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        iexec = hdr.index("Instructions Executed")
        isamp = hdr.index("# Samples")
        continue
    if hdr is None or len(row) < len(hdr) or row[2] != "-":
        continue
    try:
        ln = int(row[0])
        e, s = int(row[iexec]), int(row[isamp])
    except ValueError:
        continue
    key = (cur_file, ln)
    pe, ps_, _ = agg.get(key, (0, 0, ""))
    agg[key] = (pe + e, ps_ + s, row[1][:100])
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f"== per source line (>= {minpct}% of instructions or samples); total inst {tot}, samples {tots} ==")
for (f, ln), (e, s, text) in sorted(agg.items()):
    if 100 * e / tot >= minpct or 100 * s / tots >= minpct:
        print(f"{f}:{ln:4d} inst {100*e/tot:5.1f}%  samp {100*s/tots:5.1f}%  {text}")
