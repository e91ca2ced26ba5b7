#!/usr/bin/env python
"""Summarise an ncu --set full report into profiles/: a markdown table of the headline counters and
(optionally) the per-launch DRAM traffic record bench.py reports as roofline.traffic.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/k1_r01_coop.md [--traffic KERNEL ROWS]
"""
import csv, io, json, os, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
lines = ["# ncu --set full summary of `%s`" % os.path.basename(rep), ""]
traffic = []
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    lines += ["## %s" % name, "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in h:
            lines.append("| %s | %s | %s |" % (k, r[h.index(k)], units[h.index(k)]))
    rd = float(r[h.index("dram__bytes_read.sum")]) * UNIT[units[h.index("dram__bytes_read.sum")]]
    wr = float(r[h.index("dram__bytes_write.sum")]) * UNIT[units[h.index("dram__bytes_write.sum")]]
    lines += ["", "DRAM traffic of this launch: %.4g bytes (read %.4g + write %.4g)" % (rd + wr, rd, wr), ""]
    traffic.append(rd + wr)
open(out, "w").write("\n".join(lines) + "\n")
if "--traffic" in sys.argv:
    i = sys.argv.index("--traffic")
    kernel, nrows = sys.argv[i + 1], int(sys.argv[i + 2])
    p = os.path.join(os.path.dirname(out), "k1_traffic.json")
    recs = json.load(open(p)) if os.path.exists(p) else []
    recs = [x for x in recs if not (x["kernel"] == kernel and x["rows"] == nrows)]
    recs.append({"kernel": kernel, "rows": nrows, "dram_bytes": sum(traffic) / len(traffic),
                 "source": os.path.basename(rep), "launches_averaged": len(traffic)})
    json.dump(recs, open(p, "w"), indent=1)
print(open(out).read())
