"""One source through dropin.main with the example settings' schedule (2 x 12 000 burn-in + 25 000 production
steps, 100 walkers): what a RUN_AGNfitter_multi.py user sees per source.  python tools/single_source_time.py"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import load_k1, reference_objects
from agnfitter_b200 import dropin

ms, g = load_k1("classic18")
data, models, P = reference_objects(ms, g, 0)
tmp = tempfile.mkdtemp()
data.output_folder, data.name = tmp + "/", "src0"
mc = dict(Nwalkers=100, Nburnsets=2, Nburn=12000, Nmcmc=25000, Nthin=10, iprint=1000, sampling_algorithm="emcee",
          chain_pickle_protocol=int(os.environ.get("CHAIN_PROTOCOL", "2")))
dropin.main(data, models, P, dict(mc, Nburn=10, Nmcmc=10), seed=1)      # warm-up: context, module load
t0 = time.perf_counter()
out = dropin.main(data, models, P, mc, seed=1)
dt = time.perf_counter() - t0
steps = 2 * 12000 + 25000
print("dropin.main: %d steps x 100 walkers in %.2f s = %.1f us per step, %.3g evals/s; acceptance %.3f; files %s" % (
    steps, dt, dt / steps * 1e6, steps * 100 / dt, float(out["acceptance"][0]),
    sorted(os.listdir(os.path.join(tmp, "src0")))))

# where the time goes
from agnfitter_b200.ensemble import EnsembleSampler
post = dropin.Posterior(data, models, P)
ens = EnsembleSampler(post.engine, 100, seed=1, n_bins=1024)
ens.set_ball(0.5 * (post.src.pmin + post.src.pmax), 1e-5, False, 1)
for label, kw in (("bare steps", {}), ("+ best tracking", dict(track_best=True)),
                  ("+ histograms every 10th", dict(track_best=True, hist_every=10))):
    ens.run(200, **kw)
    t0 = time.perf_counter()
    ens.run(5000, **kw)
    print("%-28s %.1f us per step" % (label, (time.perf_counter() - t0) / 5000 * 1e6))
ens.record(5000)
t0 = time.perf_counter()
ens.run(5000, track_best=True, hist_every=10)
print("%-28s %.1f us per step" % ("+ chain recording", (time.perf_counter() - t0) / 5000 * 1e6))
t0 = time.perf_counter()
ens.save_chains(os.path.join(tmp, "x.sav"), 0)
print("save_chains (5000 steps)     %.2f s" % (time.perf_counter() - t0))
post.engine.set_option("ens_fused", 0)
ens.record(0)
ens.run(200)
t0 = time.perf_counter()
ens.run(5000)
print("%-28s %.1f us per step" % ("three-kernel path, bare", (time.perf_counter() - t0) / 5000 * 1e6))
