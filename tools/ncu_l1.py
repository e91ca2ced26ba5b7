#!/usr/bin/env python
"""Per-source-line L1 traffic of a kernel from an ncu --set full --import-source report:
shared-memory wavefronts, global tag requests and L2 sectors, joined to source lines through nvdisasm.

    python tools/ncu_l1.py gpurun_out/x.ncu-rep k1_coop_kernel 'ILb0ELi1ELb0ELi9ELb0' [rows_per_launch]
"""
import csv, io, os, subprocess, sys, tempfile
from collections import defaultdict

rep, kern, inst = sys.argv[1], sys.argv[2], sys.argv[3]
rows = float(sys.argv[4]) if len(sys.argv) > 4 else None
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
    seq, cur = [], None
    for l in dis[start + 1:]:
        t = l.strip()
        if t.startswith(".section"):
            break
        if t.startswith("//## File"):
            parts = t.split('"')
            fn = os.path.basename(parts[1])
            ln = int(t.split("line")[1].split()[0])
            cur = (fn, ln)
        elif t.startswith("/*") and "*/" in t and ";" in t:
            seq.append(cur)
    break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(raw)))
hdr = allrows[1]
data = []
for r in allrows[2:]:
    if len(r) < len(hdr) or r[0] == "Address" or r[0].startswith("Kernel"):
        break
    data.append(dict(zip(hdr, r)))
assert seq is not None and len(seq) == len(data), (seq and len(seq), len(data))
agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
for loc, d in zip(seq, data):
    g = lambda k: float(d.get(k) or 0)
    a = agg[loc]
    a[0] += g("L1 Wavefronts Shared")
    a[1] += g("L1 Tag Requests Global")
    a[2] += g("L2 Theoretical Sectors Global")
    a[3] += g("L1 Wavefronts Shared Excessive")
tot = [sum(v[i] for v in agg.values()) for i in range(4)]
print("totals: shared wavefronts %.3g (excessive %.3g), global tag requests %.3g, L2 sectors %.3g" % (
    tot[0], tot[3], tot[1], tot[2]))
if rows:
    print("per row: shared wavefronts %.2f, global requests %.2f, sectors %.2f" % (
        tot[0] / rows * 32, tot[1] / rows * 32, tot[2] / rows))
src_cache = {}
for loc, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
    if v[0] + v[1] < 0.005 * (tot[0] + tot[1]):
        continue
    fn, ln = loc if loc else ("?", 0)
    path = os.path.join(ROOT, "agnfitter_b200", "csrc", fn)
    if path not in src_cache:
        src_cache[path] = open(path).read().split("\n") if os.path.exists(path) else []
    text = src_cache[path][ln - 1].strip()[:90] if 0 < ln <= len(src_cache[path]) else ""
    print("%-16s %4d  shared %5.1f%%  global %5.1f%%  sectors %5.1f%%  %s" % (
        fn, ln, 100 * v[0] / max(tot[0], 1), 100 * v[1] / max(tot[1], 1), 100 * v[2] / max(tot[2], 1), text))
