"""Sweep the host-entry-point pipeline chunk size on the bench workload (tuning aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import argparse
args = argparse.Namespace(sources=100000, walkers=100, positions="uniform", variant=1, no_sort_z=False, cpu_sources=16)
eng, ms, dev_tables, raw, pars, sid, info = bench.make_workload(args, 0, 0)
h_pars = torch.from_numpy(pars).pin_memory(); h_sid = torch.from_numpy(sid).pin_memory()
h_lnp = torch.empty(len(pars), dtype=torch.float64).pin_memory()
np_pars, np_sid, np_lnp = h_pars.numpy(), h_sid.numpy(), h_lnp.numpy()
d = torch.empty_like(h_pars, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); d.copy_(h_pars, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("raw H2D of params: %.2f ms = %.1f GB/s" % (dt * 1e3, h_pars.numel() * 8 / dt / 1e9))
for rows in (0, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 5_000_000):
    eng.set_option("pipeline_rows", rows)
    for _ in range(2): eng.lnprob(np_pars, np_sid, out=np_lnp)
    t0 = time.perf_counter()
    for _ in range(5): eng.lnprob(np_pars, np_sid, out=np_lnp)
    dt = (time.perf_counter() - t0) / 5
    print("pipeline_rows %8d: %.2f ms/step  %.3g evals/s  (%.1f GB/s moved)" % (rows, dt * 1e3, len(pars) / dt, 920e6 / dt / 1e9))
