#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table
(per kernel and grid: launches, total, share, average).

    python tools/launch_list.py gpurun_out/launches.csv profiles/launches_r01.md "python bench.py"
"""
import csv, re, sys
from collections import defaultdict

src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "python bench.py")
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    name = re.sub(r"^void\s+", "", r["Kernel Name"])
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"(agnf::)?(<unnamed>|\(anonymous namespace\))::", "", name).replace("agnf::", "")
    name = name.replace("(bool)", "").replace("(int)", "")
    key = "%s  grid %s" % (name, r["Grid Size"])
    agg[key][0] += 1
    agg[key][1] += ns
tot = sum(v[1] for v in agg.values())
with open(out, "w") as f:
    f.write("# Launch list of `%s` — gpu__time_duration.sum per kernel and grid\n\n" % cmd)
    f.write("Source: `%s` (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).\n" % src.replace("gpurun_out/", "profiles/"))
    f.write("Per-launch times under ncu are cold-cache and serialised; compare shares, not absolutes.\n\n")
    f.write("| kernel, grid | launches | total ms | share | avg us |\n|---|---|---|---|---|\n")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| %s | %d | %.3f | %.1f%% | %.1f |\n" % (k, n, ns / 1e6, 100 * ns / tot, ns / n / 1e3))
print("kernels: %d, total %.1f ms" % (len(agg), tot / 1e6))
