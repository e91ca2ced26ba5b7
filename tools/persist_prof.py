"""Phase clocks of the persistent sampler for one source (library built with EXTRA=-DKC_PROF)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import load_k1, reference_objects
from agnfitter_b200 import dropin
from agnfitter_b200.ensemble import EnsembleSampler
ms, g = load_k1("classic18")
data, models, P = reference_objects(ms, g, 0)
post = dropin.Posterior(data, models, P)
if len(sys.argv) > 1:
    post.engine.set_option("ens_persist_rows", int(sys.argv[1]))       # walkers per warp (0 = automatic)
ens = EnsembleSampler(post.engine, 100, seed=1, n_bins=0)
ens.set_ball(0.5 * (post.src.pmin + post.src.pmax), 1e-5, False, 1)
ens.run(500)
base = [post.engine.get_option("kc_prof%d" % k) for k in range(10)]
t0 = time.perf_counter()
ens.run(5000)
dt = time.perf_counter() - t0
now = [post.engine.get_option("kc_prof%d" % k) for k in range(10)]
names = ["proposal", "params ready..box", "norms+records", "band phase", "phase 3 + accept", "syncthreads", "snap",
         "rows + galaxy copies", "box test", "table wait"]
print("rows per warp %s: " % (sys.argv[1] if len(sys.argv) > 1 else "auto") + "%.1f us per step; cycles per pass (warp 0): " % (dt / 5000 * 1e6) + ", ".join(
    "%s %.0f" % (n, (b - a) / 10000.0) for n, a, b in zip(names, base, now)))
