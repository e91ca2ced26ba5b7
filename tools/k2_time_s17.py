"""K2 contraction of an S17-shaped starburst family (9 600 templates x 721 samples) with the 49-band
example filter set: python tools/k2_time_s17.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from agnfitter_b200 import synth
from agnfitter_b200.engine import DictBuilder

fd = synth.synthetic_filterdict(which="example49")
builder = DictBuilder(fd, device=0)
rng = np.random.default_rng(0)
lognu = np.linspace(11.0, 14.9, 721)
F = {}
for i in range(9600):
    F[(str(20 + 0.1 * (i // 64)), str((i % 64) / 64.0))] = (lognu, 1e-28 * rng.lognormal(0, 1, size=721))
z = np.round(np.arange(0.0, 6.0, 0.006), 3)
for mma, sw in ((0, 0), (0, 1), (1, 1), (0, 0), (0, 1), (1, 1)):
    builder.set_option("k2_mma", mma)
    builder.set_option("k2_shared_window", sw)
    t = time.perf_counter()
    out = builder.build_family(F, z, on_device=True)
    dt = time.perf_counter() - t
    print("S17-like T=%d n_lam=721 n_z=%d bands=%d mma=%d shared_window=%d kernel=%d contract %.3f ms wall %.1f ms  %.3g entries/s" % (
        len(F), len(z), builder.n_b, mma, sw, builder.get_option("k2_last_kernel"), builder.contract_ms, dt * 1e3,
        out.shape[0] * out.shape[1] * builder.n_b / (builder.contract_ms * 1e-3)), flush=True)
    del out
