"""Where the wall time of DictBuilder.build_family goes (host side): python tools/k2_host_profile.py"""
import sys, os, time, cProfile, pstats
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
out = builder.build_family(F, z, on_device=True)
del out
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    out = builder.build_family(F, z, on_device=True)
    del out
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
