#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction counters with source lines (via nvdisasm -gi of the
in-tree library) and print instructions per warp and stall-sample share per source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k1_coop_kernel 'ILb0ELi1ELb0E' [warps]
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kern, inst = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.abspath(os.environ.get("AGNFIT_B200_LIB") or os.path.join(ROOT, "agnfitter_b200", "csrc", "libagnfit_b200.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
seq = None
for f in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
    start = None
    for i, l in enumerate(dis):
        if l.strip().startswith(".section") and ".text." in l and kern in l and inst in l:
            start = i
            break
    if start is None:
        continue
    seq, cur, fn = [], None, None
    for l in dis[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            fn, cur = m.group(1).split("/")[-1], int(m.group(2))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            seq.append((fn, cur, m.group(2)))
    break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data, st = [], defaultdict(int)
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Address" or r[0].startswith("Kernel"):
        break
    data.append((int(r[iI]), int(r[iS])))
    for i in stall:
        st[hdr[i]] += int(r[i] or 0)
assert seq is not None and len(seq) == len(data), (seq and len(seq), len(data))
warps = float(sys.argv[4]) if len(sys.argv) > 4 else max(d[0] for d in data[:40])
agg = defaultdict(lambda: [0.0, 0])
for (fn, ln, _), (e, sm) in zip(seq, data):
    agg[(fn, ln)][0] += e / warps
    agg[(fn, ln)][1] += sm
tot = sum(v[1] for v in agg.values()) or 1
print("instructions per warp: %.0f   (warps = %.0f)" % (sum(d[0] for d in data) / warps, warps))
src = {}
for (fn, ln), v in sorted(agg.items(), key=lambda x: (x[0][0] or "", x[0][1] or 0)):
    if v[0] >= 6 or v[1] > 0.004 * tot:
        if fn not in src:
            p = os.path.join(ROOT, "agnfitter_b200", "csrc", fn or "")
            src[fn] = open(p).read().split("\n") if os.path.isfile(p) else []
        t = src[fn][ln - 1].strip()[:80] if ln and ln <= len(src[fn]) else ""
        print("%-14s %4s  instr/warp %7.1f  samples %5.1f%%  %s" % (fn, ln, v[0], 100.0 * v[1] / tot, t))
s = sum(st.values()) or 1
print("stalls: " + "  ".join("%s %.1f%%" % (k[6:], 100.0 * v / s) for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
