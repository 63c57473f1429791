#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launches.csv into the tracked summaries under profiles/ (run in the build container).

    python profiles/summarize.py <round> <prof.ncu-rep> [launches.csv] [--tag stripe] [--cmd "..."]
Writes profiles/r<round>_stripe_ncu.md, profiles/r<round>_launches.md and profiles/traffic.json
(dram bytes per launch of the dominant kernel, read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_tensor.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__mem_tensor_writes_op_utcmma.sum", "smsp__mem_tensor_reads_op_ldt.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    tag, cmd = "stripe", "python bench.py ..."
    frames = samples = 0
    if "--frames" in sys.argv:
        i = sys.argv.index("--frames"); frames = int(sys.argv[i + 1]); del sys.argv[i:i + 2]
    if "--samples" in sys.argv:
        i = sys.argv.index("--samples"); samples = int(sys.argv[i + 1]); del sys.argv[i:i + 2]
    if "--tag" in sys.argv:
        i = sys.argv.index("--tag"); tag = sys.argv[i + 1]; del sys.argv[i:i + 2]
    if "--cmd" in sys.argv:
        i = sys.argv.index("--cmd"); cmd = sys.argv[i + 1]; del sys.argv[i:i + 2]
    rnd, rep = sys.argv[1], sys.argv[2]
    hdr, units, rows = raw(rep)
    lines = [f"# ncu --set full summary, round {rnd}: `{Path(rep).name}`", "",
             f"Command: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 {cmd}`",
             "(per-launch values; ncu replays the kernel ~40x with cold caches: compare SHARES and byte counts, not absolute times)", ""]
    traffic = {}
    for r in rows:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        lines.append(f"## {d['Kernel Name']}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in WANT:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {u.get(k, '')} |")
        def to_bytes(k):
            v = float(d[k]); un = u.get(k, "")
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(un, 1.0)
        if "metric_stripe" in d["Kernel Name"]:
            tb = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            traffic["metric_stripe_kernel_bytes_per_launch"] = tb
            traffic["dram_read_bytes"] = to_bytes("dram__bytes_read.sum")
            traffic["dram_write_bytes"] = to_bytes("dram__bytes_write.sum")
            traffic["grid"] = d.get("launch__grid_size")
            traffic["source"] = Path(rep).name
        # stall reasons
        st = []
        for k in hdr:
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    st.append((float(d[k]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        lines.append("")
        lines.append("Warp-state samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for v, k in sorted(st, reverse=True)[:8]))
        lines.append("")
    # instruction mix from the source page
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    hdr2, data, k = None, [], 0
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            k += 1
            if k == 2:
                break
            continue
        if r and r[0] == "Address":
            hdr2 = r
            continue
        if hdr2 and len(r) == len(hdr2):
            data.append(dict(zip(hdr2, r)))
    if data:
        ops = collections.Counter()
        tot = 0
        for d in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
            op = m.group(2).split(".")[0] if m else "?"
            c = int(d["Instructions Executed"]); ops[op] += c; tot += c
        lines.append("### Executed warp-instruction mix (source page)")
        lines.append("")
        lines.append(", ".join(f"{op} {100 * c / tot:.1f}%" for op, c in ops.most_common(16)))
        lines.append("")
        tma = sorted({re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"]).group(2) for d in data
                      if re.search(r"UBLKCP|UTMALDG|UTMASTG|SYNCS|REDUX|UTC|LDTM|STTM|FFMA2|FMUL2|FADD2|F2FP", d["Source"])})
        lines.append("TMA / mbarrier / tcgen05 (UTC*MMA, LDTM) / REDUX / packed-FP32 SASS seen in the kernel: " + ", ".join(tma))
        lines.append("")
    (HERE / f"r{rnd}_{tag}_ncu.md").write_text("\n".join(lines))
    if traffic:
        # the capture's geometry (bench.py scales the figure to its own frame count): --frames / --samples, else what the file held
        old = json.loads((HERE / "traffic.json").read_text()) if (HERE / "traffic.json").exists() else {}
        traffic["frames"] = frames or old.get("frames", 1024)
        traffic["samples_per_frame"] = samples or old.get("samples_per_frame", 262144)
        (HERE / "traffic.json").write_text(json.dumps(traffic, indent=1))
    if len(sys.argv) > 3:
        rows = [r for r in csv.reader(open(sys.argv[3])) if r]
        h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hd = rows[h]; ki = hd.index("Kernel Name"); vi = hd.index("Metric Value")
        agg = collections.OrderedDict()
        for r in rows[h + 1:]:
            if len(r) <= vi:
                continue
            name = re.sub(r"\(.*", "", r[ki])
            agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
        tot = sum(sum(v) for v in agg.values())
        out = [f"# ncu launch list, round {rnd} (`--metrics gpu__time_duration.sum --clock-control none`)", "",
               "Serialised, cold-cache per-launch times: the SHARE of each kernel is what matters.", "",
               "| kernel | launches | total us | share |", "|---|---|---|---|"]
        for name, v in agg.items():
            out.append(f"| `{name}` | {len(v)} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
        (HERE / (f"r{rnd}_launches.md" if tag == "stripe" else f"r{rnd}_{tag}_launches.md")).write_text("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
