"""Where do the ~50 us of a 50-row agnf_lnprob_batch call go?  (tuning aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import load_k1
from agnfitter_b200 import sources
from agnfitter_b200.engine import Engine
from agnfitter_b200._lib import ptr
ms, g = load_k1("classic18")
eng = Engine(ms, device=0)
src = sources.make_sources(ms, g["src_z"], g["src_dlum"], g["src_nus"], g["src_fluxes"], g["src_fluxerrs"])
eng.set_sources(src)
rng = np.random.default_rng(0)
n = 50
pars = src.pmin[0] + rng.uniform(size=(n, ms.n_par)) * (src.pmax[0] - src.pmin[0])
sid = np.zeros(n, dtype=np.int32); out = np.empty(n)
def timeit(f, reps=2000):
    for _ in range(50): f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e6
print("Engine.lnprob(pars, sid, out)      %.1f us" % timeit(lambda: eng.lnprob(pars, sid, out=out)))
print("Engine.lnprob(pars)                %.1f us" % timeit(lambda: eng.lnprob(pars)))
lib, h = eng.lib, eng.ctx.h
pp, ps, po = ptr(pars), ptr(sid), ptr(out)
print("raw ctypes agnf_lnprob_batch       %.1f us" % timeit(lambda: lib.agnf_lnprob_batch(h, pp, ps, n, po, None, None, None, None)))
print("raw ctypes agnf_synchronize        %.1f us" % timeit(lambda: lib.agnf_synchronize(h)))
for zc in (0, 256):
    eng.set_option("zero_copy_rows", zc)
    for m in (50, 200):
        pm = np.ascontiguousarray(np.tile(pars, (4, 1))[:m]); sm = np.zeros(m, dtype=np.int32); om = np.empty(m)
        a, b, c = ptr(pm), ptr(sm), ptr(om)
        print("raw, zero_copy_rows=%d, %d rows      %.1f us" % (zc, m, timeit(lambda: lib.agnf_lnprob_batch(h, a, b, m, c, None, None, None, None))))
eng.set_option("small_rows", 0)
print("raw, small_rows=0 (pageable copies) %.1f us" % timeit(lambda: lib.agnf_lnprob_batch(h, pp, ps, n, po, None, None, None, None)))
eng.set_option("small_rows", 4096)
eng.set_option("k1_variant", 0)
print("raw, per-thread kernel             %.1f us" % timeit(lambda: lib.agnf_lnprob_batch(h, pp, ps, n, po, None, None, None, None)))
