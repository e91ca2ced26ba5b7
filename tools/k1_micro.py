"""Time K1 on a golden case's dictionary with device-resident batches (tuning aid): evals/s per case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import load_k1
from agnfitter_b200 import sources
from agnfitter_b200.engine import Engine

cases = sys.argv[1:] or ["classic18", "cat49_catalog_example", "radio_xray49", "freeebv_xray49"]
n = 2_000_000
for case in cases:
    ms, g = load_k1(case)
    eng = Engine(ms, device=0)
    src = sources.make_sources(ms, g["src_z"], g["src_dlum"], g["src_nus"], g["src_fluxes"], g["src_fluxerrs"])
    eng.set_sources(src)
    rng = np.random.default_rng(0)
    sid = np.sort(rng.integers(src.S, size=n)).astype(np.int32)
    pars = src.pmin[sid] + rng.uniform(size=(n, ms.n_par)) * (src.pmax[sid] - src.pmin[sid])
    d_p = torch.from_numpy(pars).cuda(); d_s = torch.from_numpy(sid).cuda()
    d_l = torch.empty(n, dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    for variant in (1, 0):
        eng.set_option("k1_variant", variant)
        for _ in range(3): eng.lnprob_device(d_p, d_s, d_l, stream_ptr=st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10): eng.lnprob_device(d_p, d_s, d_l, stream_ptr=st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / 10
        print("%-24s n_b=%2d n_par=%2d kernel=%d  %.3f ms  %.3g evals/s" % (case, ms.n_b, ms.n_par, eng.get_option("k1_last_kernel"), ms_, n / ms_ * 1e3))
