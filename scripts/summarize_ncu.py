#!/usr/bin/env python
"""Turn an .ncu-rep (read with `ncu -i`) into the small tracked summaries under profiles/:
   python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r02_ncu_scan  [note] [workload tag, e.g. cfg3]"""
import csv, json, subprocess, sys, io

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
workload = sys.argv[4] if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
summ = []
md = [f"# ncu --set full summary ({rep.split('/')[-1]})", "", note, ""]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    d = {"kernel": name, "workload": workload}
    md += [f"## `{name[:90]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for w in WANT:
        if w in idx:
            d[w] = r[idx[w]]; md.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    def mb(x, u): return float(x.replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}[u]
    try:
        tr = mb(d["dram__bytes_read.sum"], units[idx["dram__bytes_read.sum"]]) + mb(d["dram__bytes_write.sum"], units[idx["dram__bytes_write.sum"]])
        d["dram_traffic_bytes_per_launch"] = tr; md.append(f"| dram traffic (read+write) per launch | {tr:.0f} | byte |")
    except Exception:
        pass
    md.append("")
    summ.append(d)
open(out + ".md", "w").write("\n".join(md))
json.dump(summ, open(out + ".json", "w"), indent=1)
print("wrote", out + ".md", out + ".json")
