"""K2 contraction of the synthetic galaxy family with the 49-band example filter set: python tools/k2_time49.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from agnfitter_b200 import synth
from agnfitter_b200.engine import DictBuilder

fd = synth.synthetic_filterdict(which="example49")
builder = DictBuilder(fd, device=0)
tpl = synth.synthetic_templates()
z = np.round(np.arange(0.0, 6.0, 0.002), 3)
F = tpl["GA"]["F"]
for mma, sw in ((0, 0), (0, 1), (1, 1), (0, 0), (0, 1), (1, 1)):
    builder.set_option("k2_mma", mma)
    builder.set_option("k2_shared_window", sw)
    t = time.perf_counter()
    out = builder.build_family(F, z, on_device=True)
    dt = time.perf_counter() - t
    print("GA T=%d n_z=%d bands=%d mma=%d shared_window=%d kernel=%d contract %.3f ms wall %.1f ms  %.3g entries/s" % (
        len(F), len(z), builder.n_b, mma, sw, builder.get_option("k2_last_kernel"), builder.contract_ms, dt * 1e3,
        out.shape[0] * out.shape[1] * builder.n_b / (builder.contract_ms * 1e-3)), flush=True)
    del out
