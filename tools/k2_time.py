"""Time the K2 contraction of the synthetic galaxy family (C3 shape) with both kernels: python tools/k2_time.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from agnfitter_b200 import synth
from agnfitter_b200.engine import DictBuilder

fd = synth.synthetic_filterdict()
builder = DictBuilder(fd, device=0)
tpl = synth.synthetic_templates()
z = np.round(np.arange(0.0, 6.0, 0.001), 3)
for fam in ("GA", "SB", "TO", "BB"):
    F = tpl[fam]["F"]
    for mma, sw in ((0, 0), (0, 1), (1, 1), (0, 0), (0, 1), (1, 1)):
        builder.set_option("k2_mma", mma)
        builder.set_option("k2_shared_window", sw)
        t = time.perf_counter()
        out = builder.build_family(F, z, on_device=True)
        dt = time.perf_counter() - t
        print("%s T=%d mma=%d shared_window=%d kernel=%d contract %.3f ms wall %.1f ms  %.3g entries/s" % (
            fam, len(F), mma, sw, builder.get_option("k2_last_kernel"), builder.contract_ms, dt * 1e3,
            out.shape[0] * out.shape[1] * builder.n_b / (builder.contract_ms * 1e-3)), flush=True)
        del out
