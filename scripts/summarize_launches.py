#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table:
   python scripts/summarize_launches.py gpurun_out/launches_r01.csv profiles/r01_ncu_launches  "title" """
import csv, collections, re, sys, shutil
src, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else ""
lines = [l for l in open(src) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("mp::", "")[:60]
    tot[name] += ms; cnt[name] += 1
total = sum(tot.values())
md = [f"# ncu launch list summary ({title}; gpu__time_duration.sum, --clock-control none)", "",
      "Per-launch times under ncu are cold-cache and serialised: the SHARE of the step is what compares with bench.py's CUDA-event split.", "",
      "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, v in tot.most_common(): md.append(f"| `{k}` | {cnt[k]} | {v:.3f} | {100*v/total:.1f} % |")
open(out + ".md", "w").write("\n".join(md) + "\n")
shutil.copy(src, out + ".csv")
print("\n".join(md[:14]))
