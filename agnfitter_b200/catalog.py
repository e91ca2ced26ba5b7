"""
Whole-catalogue fitting: the multi-source, multi-GPU counterpart of RUN_AGNfitter_multiprocessing
(RUN_AGNfitter_multi.py:246-270).  The reference maps `RUN_AGNfitter_onesource_independent` over catalogue
rows with a process pool; every worker builds its source's dictionary (DICTIONARIES_AGNfitter.py:139-154),
runs the sampler (MCMC_AGNfitter.py:98-135) and writes its own files.  Here every rank (one process per GPU)

  1. takes a contiguous block of the z-sorted catalogue (shard.shard_catalog — a dictionary z-slice is never
     split between ranks),
  2. builds the dictionary for the distinct redshifts of its block on its GPU (K2) and keeps it in HBM,
  3. runs the reference's burn-in / production schedule for all its sources in lock-step (K3 + K1),
  4. reduces the thinned production chain to the reference's five percentiles on the device,

and the only exchange is one gather of the fixed-size per-source summaries to rank 0 (NCCL on the GPUs, gloo
in the CPU tests).  No collective touches the evaluation path.
"""
import numpy as np

from . import shard as _shard
from . import sources as _sources
from .engine import Engine, DictBuilder
from .ensemble import EnsembleSampler, REFERENCE_PERCENTILES
from .flatten import ModelSet, family_from_keys_float

FAMILY_OF = (("GA", "GALAXY"), ("SB", "STARBURST"), ("TO", "TORUS"), ("BB", "BBB"), ("RAD", "AGN_RAD"))


def build_modelset_on_device(templates, filterdict, z_grid, settings, builder):
    """templates: {'GALAXY': dict(F={key: (log nu, Fnu)}, names=[...], types=[...], aux={key: value} | None), ...}
    — the reference's `*Fdict_4plot` dicts with their parameter names / types (the loaders' return values,
    MODEL_AGNfitter.py:69-1147) and, where the priors need them, GALAXYatt_dict / STARBURST_LIRdict.
    Returns (ModelSet without host tables, {family: CUDA tensor [n_z, T, pitch]})."""
    ms = ModelSet(settings=dict(settings), bands=np.asarray(filterdict[0], dtype=np.float64),
                  z_grid=np.asarray(z_grid, dtype=np.float64))
    dev = {}
    for short, long_name in FAMILY_OF:
        if long_name not in templates:
            continue
        fam = templates[long_name]
        keys = list(fam["F"].keys())
        kf = np.array(keys).T.astype(float)
        if kf.ndim == 1:
            kf = kf[None, :]
        fitted = not (short == "RAD" and keys == ['-99.9'])
        tab = builder.build_family(fam["F"], z_grid, on_device=True, reuse_staging=True)
        aux = None if fam.get("aux") is None else np.array([float(fam["aux"][k]) for k in keys])
        if short == "GA" and 'free' in fam["types"]:
            aux = None
        ms.families.append(family_from_keys_float(short, fam["names"], fam["types"], kf, fitted, None, aux))
        dev[short] = tab
    return ms, dev


INTLUM_PERCENTILES = (16.0, 50.0, 84.0)     # PLOTandWRITE_AGNfitter.py:152-173 writes median and +-1 sigma


def summary_columns(n_par, n_q=len(REFERENCE_PERCENTILES), n_intlum=0):
    """Layout of one summary row."""
    return ["source_index", "z"] + ["p%d_q%d" % (p, q) for p in range(n_par) for q in range(n_q)] + \
           ["best_%d" % p for p in range(n_par)] + ["best_lnp", "acceptance"] + \
           ["L%d_q%d" % (m, int(q)) for m in range(n_intlum) for q in INTLUM_PERCENTILES]


def _chunk_bounds(zs, per_source_bytes, per_slice_bytes, budget_bytes, max_sources=None):
    """Cut the rank's z-sorted block into chunks whose dictionary slices + sampler state fit `budget_bytes`; a chunk
    never splits a run of equal redshifts unless the run alone exceeds the budget."""
    n = len(zs)
    bounds, start = [0], 0
    while start < n:
        used, end, n_z, last = 0.0, start, 0, None
        while end < n:
            new_z = zs[end] != last
            need = per_source_bytes + (per_slice_bytes if new_z else 0.0)
            if end > start and (used + need > budget_bytes or (max_sources and end - start >= max_sources)):
                break
            used += need
            last = zs[end]
            end += 1
        bounds.append(end)
        start = end
    return bounds


def fit_catalog(templates, filterdict, settings, z, dlum, nus, fluxes, fluxerrs, mc, seed=0, n_bins=1024,
                rank=0, world=1, device=0, intlum_models=None, intlum_ranges_hz=None, exact_percentiles=False,
                memory_budget_bytes=None, max_sources_per_chunk=None):
    """Fit every source of the catalogue this rank owns and gather the summaries.

    z, dlum [S]; nus [n_b] or [S, n_b]; fluxes, fluxerrs [S, n_b] in the reference's DATA conventions
    (DATA_AGNfitter.py:145-193 already applied).  mc: MCMC_settings() dict (Nwalkers, Nburnsets, Nburn, Nmcmc,
    Nthin).  intlum_models / intlum_ranges_hz (out['intlum_models'], out['intlum_freqranges'] in Hz): also run
    the integrated-luminosity stage (K4) on the final walker positions of every source and append the 16/50/84
    percentiles of log10 of each column (the reference draws realizations2int rows from the thinned chain,
    PLOTandWRITE_AGNfitter.py:608; here the W walkers of the last step are those rows).
    exact_percentiles: keep the thinned second half of the production chain on the device and report
    numpy.percentile-exact values (what PLOTandWRITE_AGNfitter.py:162 computes) instead of histogram quantiles of
    resolution (P.max - P.min) / n_bins.
    The rank's block is processed in chunks along z so that the dictionary slices of a chunk's redshifts, the walker
    state and (exact_percentiles) the kept chain fit `memory_budget_bytes` (default: 60 % of the free device
    memory) — a catalogue of unique redshifts needs one slice per source.  Random streams are keyed by the source's
    position in the z-sorted catalogue and a dictionary slice does not depend on its neighbours, so the result is
    the same for any chunking (tests/test_catalog_gpu.py).
    Returns on rank 0 an array [S_total, len(summary_columns)] ordered by z (None on other ranks)."""
    import torch
    z = np.asarray(z, dtype=np.float64)
    order, lo, hi = _shard.shard_catalog(z, rank, world)
    mine = order[lo:hi]
    nus = np.asarray(nus, dtype=np.float64)
    dlum = np.asarray(dlum, dtype=np.float64)
    rows = None
    if len(mine):
        zs_all = z[mine]
        W = int(mc['Nwalkers'])
        n_b = len(filterdict[0])
        pitch = (n_b + 1) & ~1
        T_sum = sum(len(templates[long_name]["F"]) for _, long_name in FAMILY_OF if long_name in templates)
        n_par_guess = 24
        n_keep = len(range(int(int(mc['Nmcmc']) / 2), int(mc['Nmcmc']), max(int(mc.get('Nthin', 1)), 1)))
        per_source = W * n_par_guess * 8 * 6 + n_bins * n_par_guess * 4 + 16 * pitch * 8
        if exact_percentiles:
            per_source += n_keep * W * (n_par_guess + 1) * 8
        per_slice = T_sum * pitch * 8 * 1.05
        if memory_budget_bytes is None:
            free, _ = torch.cuda.mem_get_info(device)
            memory_budget_bytes = 0.6 * free
        bounds = _chunk_bounds(zs_all, per_source, per_slice, float(memory_budget_bytes), max_sources_per_chunk)
        chunks = []
        staging = {}      # the grouped, pinned template blocks, shared by the chunks' builders (DictBuilder.build_family)
        for c0, c1 in zip(bounds[:-1], bounds[1:]):
            sel = mine[c0:c1]
            zs = z[sel]
            builder = DictBuilder(filterdict, device=device)
            builder._tpl_cache = staging
            ms, dev = build_modelset_on_device(templates, filterdict, np.unique(zs), settings, builder)
            eng = Engine(ms, device=device, ctx=builder.ctx, device_tables=dev)
            src = _sources.make_sources(ms, zs, dlum[sel], nus if nus.ndim == 1 else nus[sel],
                                        np.asarray(fluxes, dtype=np.float64)[sel],
                                        np.asarray(fluxerrs, dtype=np.float64)[sel])
            eng.set_sources(src)
            # random numbers are keyed by the source's position in the z-sorted catalogue, not by the shard or the
            # chunk: the fit of a source does not depend on the number of GPUs or on the memory budget
            ens = EnsembleSampler(eng, W, seed=int(seed), n_bins=0 if exact_percentiles else n_bins,
                                  source_offset=lo + c0)
            out = ens.run_reference_schedule(mc, src.pmin, src.pmax, None,              # balls drawn on the device
                                             exact_percentiles=exact_percentiles)
            blocks = [sel[:, None].astype(np.float64), zs[:, None], out["percentiles"].reshape(len(sel), -1),
                      out["best_pos"], out["best_lnp"][:, None], out["acceptance"][:, None]]
            if intlum_models is not None:
                from .plotset import plotset_from_templates
                from .realizations import SedEngine
                ps = plotset_from_templates({short: templates[long_name] for short, long_name in FAMILY_OF
                                             if long_name in templates}, ms)
                sed = SedEngine(eng, ps, zs, dlum[sel])
                sed.set_bands(intlum_models, intlum_ranges_hz)
                with np.errstate(all="ignore"):
                    L = np.log10(sed.int_lums_ensemble(W))                           # [S, W, n_models]
                q = np.percentile(L, INTLUM_PERCENTILES, axis=1)                       # [3, S, n_models]
                blocks.append(np.transpose(q, (1, 2, 0)).reshape(len(sel), -1))
                del sed
            chunks.append(np.concatenate(blocks, axis=1))
            del ens, eng, dev, builder
            torch.cuda.empty_cache()
        rows = np.concatenate(chunks, axis=0)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dev_t = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
        # ranks with no sources still take part in the gather; they learn the row width from the others
        width = torch.tensor([0 if rows is None else rows.shape[1]], dtype=torch.int64, device=dev_t)
        dist.all_reduce(width, op=dist.ReduceOp.MAX)
        w = int(width.item())
        local = torch.zeros((0, w), dtype=torch.float64, device=dev_t) if rows is None else \
            torch.from_numpy(np.ascontiguousarray(rows)).to(dev_t)
        got = _shard.gather_summaries(local)
        return None if got is None else got.cpu().numpy()
    return rows
