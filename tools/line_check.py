"""A/B of the two band-phase mappings of the cooperative K1 kernel (option coop_line: 0 pair form, 1 line form where it
is the default, 2 line form forced) on bench.py's workloads: bit equality of lnp on the same rows, then launch times
(plain and, with --ens, the fused ensemble step).

    python tools/line_check.py [--sources N] [--bands 49] [--c4] [--ga-free]
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--sources", type=int, default=100000)
ap.add_argument("--walkers", type=int, default=100)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--bands", type=int, default=18)
ap.add_argument("--ga-free", action="store_true")
ap.add_argument("--c4", action="store_true")
ap.add_argument("--positions", default="uniform")
ap.add_argument("--variant", type=int, default=1)
ap.add_argument("--ens", action="store_true")
a = ap.parse_args()
a.scaling, a.no_sort_z = "weak", False
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
ref = None
for quads in ((0, 1, 2),):
    eng, ms, dev_tables, raw, pars, sid, info = bench.make_workload(a, 0, 0)
    n = len(pars)
    d_p = torch.from_numpy(pars).to(dev)
    d_s = torch.from_numpy(sid).to(dev)
    d_l = torch.empty(n, dtype=torch.float64, device=dev)
    for q in quads:
        eng.set_option("coop_line", q)
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
        out = d_l.cpu().numpy()
        if ref is None:
            ref = out
        same = np.array_equal(ref, out, equal_nan=True)
        msg = "coop_line %d: plain %.4f ms per %d rows = %.4g evals/s  identical to the pair form: %s" % (
            q, t, n, n / t * 1e3, same)
        if a.ens:
            from agnfitter_b200.ensemble import EnsembleSampler
            ens = EnsembleSampler(eng, a.walkers, seed=20260101, n_bins=256)
            ens.set_positions(pars.reshape(info["S"], a.walkers, ms.n_par))
            ens.run(5, hist_every=10, track_best='thinned')
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ens.run(100, hist_every=10, track_best='thinned')
            te = (time.perf_counter() - t0) / 100 * 1e3
            pos = ens.positions() if hasattr(ens, "positions") else None
            msg += "  ens %.4f ms per step" % te
            del ens
        print(msg, flush=True)
    del eng, ms, dev_tables, d_p, d_s, d_l
    torch.cuda.empty_cache()
