"""Runs agnfitter_b200.catalog.fit_catalog on a small synthetic catalogue at the world size given by the
environment (torchrun) and writes rank 0's summary table to argv[1].  Used by tests/test_catalog_gpu.py:
the table must not depend on the number of ranks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


INTLUM_MODELS = ['sb', 'bbb', 'bbbdered', 'gal', 'AGNfrac_opt', 'tor', 'tor+bbb', 'AGNfrac_IR', 'tor']
INTLUM_RANGES_UM = [[8., 1000.], [0.1, 1.], [0.1, 1.], [0.1, 1.], [0.1, 1.], [1., 30.], [0.1, 30], [8., 1000.], [6, 6]]


def make_inputs(S=24, seed=5):
    from agnfitter_b200 import synth, sources
    from agnfitter_b200.engine import Engine, DictBuilder
    fd = synth.synthetic_filterdict()
    tpl = synth.synthetic_templates()
    templates = {long: tpl[short] for short, long in (("GA", "GALAXY"), ("SB", "STARBURST"), ("TO", "TORUS"),
                                                      ("BB", "BBB"))}
    rng = np.random.default_rng(seed)
    z = np.round(rng.uniform(0.05, 3.0, size=S), 2)          # repeated redshifts on purpose
    dlum = synth.z2Dlum(z)
    nus = np.asarray(fd[0])
    # fluxes = ymodel(truth) with noise, evaluated through the product itself on one device
    builder = DictBuilder(fd, device=0)
    ms, _ = synth.build_modelset(tpl, fd, np.unique(z), builder, on_device=False)
    eng = Engine(ms, device=0, ctx=builder.ctx)
    dummy = sources.make_sources(ms, z, dlum, nus, np.ones((S, ms.n_b)), np.ones((S, ms.n_b)))
    eng.set_sources(dummy)
    truth = dummy.pmin + rng.uniform(0.3, 0.7, size=dummy.pmin.shape) * (dummy.pmax - dummy.pmin)
    _, _, ym, _, _ = eng.lnprob(truth, np.arange(S, dtype=np.int32), detail=True)
    fl, fe = synth.catalog_fluxes(rng, ym, nus, z)
    return templates, fd, dict(synth.CLASSIC_SETTINGS), z, dlum, nus, fl, fe


def main(out_path):
    import torch
    import torch.distributed as dist
    from agnfitter_b200.catalog import fit_catalog
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_dev = torch.cuda.device_count()
    device = int(os.environ.get("LOCAL_RANK", "0")) % max(n_dev, 1)
    torch.cuda.set_device(device)
    if world > 1:
        # nccl needs one GPU per rank; with fewer GPUs than ranks the (tiny) gather goes over gloo
        dist.init_process_group("nccl" if n_dev >= world else "gloo")
    templates, fd, settings, z, dlum, nus, fl, fe = make_inputs()
    mc = dict(Nwalkers=16, Nburnsets=2, Nburn=15, Nmcmc=40, Nthin=2)
    from agnfitter_b200.realizations import ranges_to_hz
    chunk = int(os.environ.get("CATALOG_CHUNK", "0")) or None
    exact = os.environ.get("CATALOG_EXACT", "0") == "1"
    table = fit_catalog(templates, fd, settings, z, dlum, nus, fl, fe, mc, seed=123, n_bins=512,
                        rank=rank, world=world, device=device, intlum_models=INTLUM_MODELS,
                        intlum_ranges_hz=ranges_to_hz(INTLUM_RANGES_UM), exact_percentiles=exact,
                        max_sources_per_chunk=chunk)
    if rank == 0:
        np.save(out_path, table)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
