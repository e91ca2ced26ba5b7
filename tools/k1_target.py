"""Runs only the K1 launch of interest on bench.py's C3 workload — the program `ncu` profiles and the quickest timing
loop while tuning.

    python tools/k1_target.py plain|ens [--sources N] [--steps K] [--bands 49] [--ga-free] [--c4] [--positions ball]
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("mode", choices=["plain", "ens"])
ap.add_argument("--sources", type=int, default=100000)
ap.add_argument("--walkers", type=int, default=100)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--bands", type=int, default=18)
ap.add_argument("--ga-free", action="store_true")
ap.add_argument("--c4", action="store_true")
ap.add_argument("--positions", default="uniform")
ap.add_argument("--variant", type=int, default=1)
ap.add_argument("--opt", action="append", default=[], help="name=value engine options")
a = ap.parse_args()
a.scaling, a.no_sort_z = "weak", False
eng, ms, dev_tables, raw, pars, sid, info = bench.make_workload(a, 0, 0)
for o in a.opt:
    k, v = o.split("=")
    eng.set_option(k, int(v))
S, W, n = info["S"], a.walkers, len(pars)
print("snap tables: %d bytes, %d cells (%d crowded), %d tasks without a table" % (
    eng.get_option("k1_table_bytes"), eng.get_option("k1_table_cells"), eng.get_option("k1_table_crowded_cells"),
    eng.get_option("k1_tasks_without_table")))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
if a.mode == "plain":
    d_p = torch.from_numpy(pars).to(dev)
    d_s = torch.from_numpy(sid).to(dev)
    d_l = torch.empty(n, dtype=torch.float64, device=dev)
    step = lambda: eng.lnprob_device(d_p, d_s, d_l, stream_ptr=stream.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / a.steps
    print("plain: %.4f ms per %d rows = %.4g evals/s (kernel %d)" % (t, n, n / t * 1e3, eng.get_option("k1_last_kernel")))
    prof = [eng.get_option("g4_prof%d" % k) for k in range(8)]
    if sum(prof):
        tiles = (n + 31) // 32 * (a.steps + 3)
        names = ["params wait", "snap", "box+norm reads / scalars", "rows wait", "band loop", "priors+out", "issue A", "fence+issue B"]
        print("per tile (cycles): " + ", ".join("%s %.0f" % (nm, p / tiles) for nm, p in zip(names, prof)))
else:
    from agnfitter_b200.ensemble import EnsembleSampler
    ens = EnsembleSampler(eng, W, seed=20260101, n_bins=256)
    ens.set_positions(pars.reshape(S, W, ms.n_par))
    ens.run(5, hist_every=10, track_best='thinned')
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ens.run(a.steps, hist_every=10, track_best='thinned')
    t = (time.perf_counter() - t0) / a.steps * 1e3
    print("ens: %.4f ms per step of %d evals = %.4g evals/s" % (t, n, n / t * 1e3))
