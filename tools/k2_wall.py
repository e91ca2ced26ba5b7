"""Wall time of the galaxy-family dictionary build at C3 scale, with the library's own host-side time stamps:
python tools/k2_wall.py [bands]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from agnfitter_b200 import synth
from agnfitter_b200.engine import DictBuilder

bands = int(sys.argv[1]) if len(sys.argv) > 1 else 18
fd = synth.synthetic_filterdict(which="example49" if bands == 49 else None)
tpl = synth.synthetic_templates()
builder = DictBuilder(fd, device=0)
z = np.round(np.arange(0.01, 6.0, 0.001), 3)
names = ["enqueue allocs+uploads", "uploads+permute+ranges (sync)", "layout + out alloc", "finite scan", "weights", "contract (sync)", "free scratch"]
for fam in ("GA", "SB", "TO", "BB"):
    F = tpl[fam]["F"]
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = builder.build_family(F, z, on_device=True)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        st = [builder.get_option("k2_prof%d" % k) / 1e3 for k in range(7)]
        n_entries = out.shape[0] * out.shape[1] * builder.n_b
        del out
    d = [st[0]] + [st[k] - st[k - 1] for k in range(1, 7)]
    print("%s T=%d: wall %.2f ms (C call %.2f ms, contract kernel %.2f ms) %.3g entries/s | %s" % (
        fam, len(F), wall, st[6], builder.contract_ms, n_entries / wall * 1e3,
        ", ".join("%s %.2f" % (n, v) for n, v in zip(names, d))))
