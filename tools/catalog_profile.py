"""Host-side cost of catalog.fit_catalog on a large synthetic catalogue with a short schedule (so that the GPU
time is small and the per-source host work shows): python tools/catalog_profile.py [S]"""
import sys, os, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mgpu_catalog_check as M
from agnfitter_b200.catalog import fit_catalog
from agnfitter_b200.realizations import ranges_to_hz

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
t0 = time.perf_counter()
templates, fd, settings, z, dlum, nus, fl, fe = M.make_inputs(S=S)
print("inputs %.2f s" % (time.perf_counter() - t0))
mc = dict(Nwalkers=100, Nburnsets=2, Nburn=20, Nmcmc=40, Nthin=2)
if len(sys.argv) > 2 and sys.argv[2] == "example":       # example/SETTINGS_AGNfitter.py:265-268, out['Nthinning'] = 10
    mc = dict(Nwalkers=100, Nburnsets=2, Nburn=12000, Nmcmc=25000, Nthin=10)
import torch  # noqa: E402  (keep its import time out of the measurement)
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
table = fit_catalog(templates, fd, settings, z, dlum, nus, fl, fe, mc, seed=1, n_bins=256,
                    intlum_models=M.INTLUM_MODELS, intlum_ranges_hz=ranges_to_hz(M.INTLUM_RANGES_UM))
dt = time.perf_counter() - t0
pr.disable()
steps = mc['Nburnsets'] * mc['Nburn'] + mc['Nmcmc']
print("fit_catalog: %d sources, %d steps, %.2f s  (%.1f us per source, %.1f sources/s, %.3g evals/s)" % (
    S, steps, dt, dt / S * 1e6, S / dt, S * 100.0 * steps / dt))
cols = table.shape[1]
print("table", table.shape, "median acceptance %.3f" % np.median(table[:, 2 + 50 + 10 + 1]))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
