#!/usr/bin/env python
"""Quick reader for an .ncu-rep: key raw metrics + the SASS lines with the most stall samples.
    python profiles/ncu_top.py <file.ncu-rep> [n_lines]"""
import csv, subprocess, sys
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h, u, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum"]
for k in want:
    if k in h:
        i = h.index(k); print(f"{k:75s} {u[i]:14s} {r[i]}")
for i, k in enumerate(h):
    if "tensor" in k and k not in want:
        print(f"{k:75s} {u[i]:14s} {r[i]}")
for i, k in enumerate(h):
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        try:
            v = float(r[i].replace(",", ""))
            if v > 0.3: print(f"{k:75s} {r[i]}")
        except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, x in enumerate(rows) if "Source" in x and "# Samples" in x)
h = rows[hi]; data = rows[hi + 1:]
si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
tot = sum(int(x[si]) for x in data if len(x) > si and x[si].isdigit())
print("total samples", tot, "instructions", len(data))
top = sorted([(int(x[si]), i) for i, x in enumerate(data) if len(x) > si and x[si].isdigit()], reverse=True)[:nl]
for s, i in sorted(top, key=lambda t: t[1]):
    print(f"{i:6d} {s:7d} {100*s/tot:5.1f}% {data[i][ie]:>10s}  {data[i][so][:110]}")
