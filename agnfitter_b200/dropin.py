"""
Reference-facing mirror of the hot path: the callables AGNfitter's samplers and driver bind to,
with the reference's names, argument order and return conventions, backed by the sm_100a kernels.

    reference callable (AGNfitter/functions)                              here
    -------------------------------------------------------------------   -----------------------------
    PARAMETERSPACE_AGNfitter.ln_probab(pars, data, models, P)   :153-173   Posterior.ln_probab
    ... called per walker by emcee.EnsembleSampler         MCMC:114        Posterior.ln_probab_batch (vectorize=True)
    my_posterior(params[n, ndim])                          MCMC:56-62      Posterior.my_posterior
    my_params_transform(cube[n, ndim])                     MCMC:64-72      Posterior.my_params_transform
    PARAMETERSPACE_AGNfitter.ymodel(nus, z, dlum, models, P, *par) :180-243  Posterior.ymodel
    DICTIONARIES_AGNfitter.filtering_models(nus, F, filterdict, z) :426-478  engine.DictBuilder.filtering_models
    MODELSDICT.construct_dictionaryarray_filtered / build  DIC:78-154      construct_dictionaryarray_filtered / build
    MCMC_AGNfitter.main(data, models, P, mc), emcee leg     MCMC:33-149     main (sampler on the GPU, same chain files)
    MCMC_AGNfitter.main(...), ultranest leg                 MCMC:74-96      main_ultranest (ultranest on the CPU calling
                                                                            the GPU posterior; many sources at once
                                                                            through multiplex.PosteriorMultiplexer)

`data`, `models`, `P` are the reference's own objects (DATA_AGNfitter.DATA, MODEL_AGNfitter.MODELS,
PARAMETERSPACE_AGNfitter.Pdict): nothing in RUN_AGNfitter_multi.py or the settings files changes.
There is no CPU fallback: constructing a Posterior without an sm_100 device raises.
"""
import numpy as np

from . import sources as _sources
from .engine import Engine, DictBuilder
from .flatten import modelset_from_modelsdicts, ModelSet, save_modelset, load_modelset


def params_transform(cube, pmin, pmax):
    """my_params_transform (MCMC_AGNfitter.py:64-72) without the Python loops:
    params[i, k] = (P.max[k] - P.min[k]) * cube[i, k] + P.min[k], same operation order."""
    cube = np.asarray(cube, dtype=np.float64)
    pmin = np.asarray(pmin, dtype=np.float64)
    pmax = np.asarray(pmax, dtype=np.float64)
    return (pmax - pmin) * cube + pmin


def check_P(ms: ModelSet, P, pmin, pmax):
    """The parameter table the caller built with the reference's Pdict must describe the same
    parameter vector the kernel unpacks (names, order, bounds)."""
    if list(P['names']) != list(ms.par_names):
        raise ValueError("P['names'] = %s but the dictionary describes %s" % (list(P['names']), ms.par_names))
    if not (np.array_equal(np.asarray(P['min'], dtype=np.float64), pmin)
            and np.array_equal(np.asarray(P['max'], dtype=np.float64), pmax)):
        raise ValueError("P['min']/P['max'] differ from the bounds Pdict gives for this dictionary and redshift")


# ---- dense on-disk dictionary beside the legacy pickle -----------------------------------------
# RUN_AGNfitter_multi.py:203-213 caches a source's dictionary as a pickle of Python dicts (`zdict.filename`), and
# every later run of that source unpickles it and walks the dicts again.  The dense form -- the ModelSet the kernels
# consume: key tables, snapping grids and one [T, n_b] fp64 table per family, `flatten.save_modelset` -- is written
# next to it as `<zdict.filename>.b200.npz`; `Posterior` loads it instead of flattening the dicts when it is there and
# describes the same dictionary.  The pickle stays: PLOTandWRITE reads the Python dicts.
def cache_path(md):
    fn = getattr(md, 'filename', None)
    if fn is None:
        return None
    if isinstance(fn, bytes):
        fn = fn.decode()
    return str(fn) + '.b200.npz'


def save_dictionary_cache(md, settings=None):
    """Write the dense dictionary of the MODELSDICT `md` (built by the reference or by `build`) beside its pickle.
    `settings`: the models settings (default `md.modelsettings`).  Returns the path written."""
    path = cache_path(md)
    if path is None:
        raise ValueError("the MODELSDICT has no filename (DICTIONARIES_AGNfitter.py:60)")
    ms = modelset_from_modelsdicts([md], md.modelsettings if settings is None else settings)
    save_modelset(path, ms)
    return path


def load_dictionary_cache(md, settings):
    """The ModelSet stored by save_dictionary_cache if it exists and describes `md` (same redshift, bands, families
    and template counts, same settings of the prior branches); None otherwise (the caller flattens the dicts)."""
    import os
    path = cache_path(md)
    if path is None or not os.path.lexists(path):
        return None
    try:
        ms = load_modelset(path)
    except Exception:
        return None
    counts = {"GA": "GALAXYFdict", "SB": "STARBURSTFdict", "TO": "TORUSFdict", "BB": "BBBFdict", "RAD": "AGN_RADFdict"}
    ok = ms.n_z == 1 and float(ms.z_grid[0]) == float(md.z)
    for f in ms.families:
        d = getattr(md, counts[f.name], None)
        ok = ok and d is not None and len(d) == f.T
    if ok:
        first = next(iter(md.GALAXYFdict.values()))
        ok = np.array_equal(np.asarray(first[0], dtype=np.float64), ms.bands)
    import json
    ok = ok and json.loads(json.dumps(dict(settings), default=str)) == ms.settings
    return ms if ok else None


class Posterior:
    """ln_probab for ONE source, bound to the reference's (data, models, P) triple.

        post = Posterior(data, models, P)
        sampler = emcee.EnsembleSampler(nwalkers, npar, post.ln_probab_batch, vectorize=True)
        # or, unchanged call sites:  emcee.EnsembleSampler(nwalkers, npar, post.ln_probab, args=[data, models, P])
    """

    def __init__(self, data, models, P, device=0, ctx=None):
        md = models.dict_modelfluxes                     # the MODELSDICT instance (MODEL_AGNfitter.py:1483-1490)
        self.ms = load_dictionary_cache(md, models.settings)      # dense .npz beside the pickle, when present
        self.from_cache = self.ms is not None
        if self.ms is None:
            self.ms = modelset_from_modelsdicts([md], models.settings)
        if float(md.z) != float(data.z):
            raise ValueError("the dictionary was built for z=%r, the source is at z=%r" % (md.z, data.z))
        self.engine = Engine(self.ms, device=device, ctx=ctx)
        nus = np.asarray(data.nus, dtype=np.float64)
        self.src = _sources.make_sources(self.ms, [float(data.z)], [float(data.dlum)], nus[None, :],
                                         np.asarray(data.fluxes, dtype=np.float64)[None, :],
                                         np.asarray(data.fluxerrs, dtype=np.float64)[None, :])
        check_P(self.ms, P, self.src.pmin[0], self.src.pmax[0])
        self.engine.set_sources(self.src)
        self.P = P
        self.n_par = self.ms.n_par
        self.bands = self.ms.bands

    # ---- emcee ---------------------------------------------------------------------------------
    def ln_probab(self, pars, data=None, models=None, P=None):
        """Same signature and return type as the reference: one walker -> Python float."""
        return float(self.engine.lnprob(np.asarray(pars, dtype=np.float64)[None, :])[0])

    def ln_probab_batch(self, params):
        """[n, n_par] -> [n]: the log_prob_fn of emcee.EnsembleSampler(..., vectorize=True)."""
        return self.engine.lnprob(params)

    # ---- ultranest -----------------------------------------------------------------------------
    def my_posterior(self, params):
        return self.engine.lnprob(params)

    def my_params_transform(self, cube):
        return params_transform(cube, self.P['min'], self.P['max'])

    # ---- plots / post-processing -----------------------------------------------------------------
    def ymodel(self, data_nus=None, z=None, dlum=None, models=None, P=None, *par):
        """-> (lum[n_b], bands) like PARAMETERSPACE_AGNfitter.ymodel."""
        _, _, ym, _, _ = self.engine.lnprob(np.asarray(par, dtype=np.float64)[None, :], detail=True)
        return ym[0], self.bands


def main(data, models, P, mc, seed=None, device=0):
    """Mirror of the emcee leg of MCMC_AGNfitter.main(data, models, P, mc) (MCMC_AGNfitter.py:33-149) with the
    whole sampler on the GPU: initial ball, `Nburnsets` burn-in sets each restarted around the best sample,
    production run, and the two chain files the rest of the pipeline reads —
    `<data.output_folder><data.name>/samples_burn1-2-3.sav` and `samples_mcmc.sav` in the format of
    MCMC_AGNfitter.save_chains (:192-206).  The move is emcee's stretch move (a = 2) with its acceptance rule; the
    random numbers come from a counter-based generator keyed by `seed` (the reference's run is unseeded, so its
    chains are not reproducible either way).  Returns the summary dict of EnsembleSampler.run_reference_schedule."""
    import os
    from .ensemble import EnsembleSampler
    algo = mc.get('sampling_algorithm', 'emcee')
    if algo == 'ultranest':
        return main_ultranest(data, models, P, mc, device=device)
    if algo != 'emcee':
        raise ValueError("sampling_algorithm must be 'emcee' or 'ultranest' (MCMC_AGNfitter.py:74,98), got %r" % (algo,))
    post = Posterior(data, models, P, device=device)
    seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
    ens = EnsembleSampler(post.engine, int(mc['Nwalkers']), seed=seed, n_bins=1024)
    rng = np.random.default_rng(seed)
    out = ens.run_reference_schedule(mc, post.src.pmin, post.src.pmax, rng, save_dir=str(data.output_folder),
                                     names=[str(data.name)])
    out["seed"] = seed
    return out


def main_ultranest(data, models, P, mc, device=0, posterior=None):
    """The ultranest leg of MCMC_AGNfitter.main (MCMC_AGNfitter.py:74-96), statement for statement, with the two
    closures replaced by the GPU-backed callables: ultranest itself stays on the CPU and keeps calling a Python
    log-posterior (`vectorized=True`: [n, ndim] -> [n]).  `posterior`: an object with `my_posterior` /
    `my_params_transform` — a Posterior (default) or a multiplex.MuxClient when many sources are fitted concurrently.
    ultranest is imported here: a machine without it gets ImportError from this function only."""
    import os
    import ultranest
    import ultranest.stepsampler
    post = posterior if posterior is not None else Posterior(data, models, P, device=device)
    if not os.path.lexists(data.output_folder + str(data.name)):
        os.mkdir(data.output_folder + str(data.name))
    sampler = ultranest.ReactiveNestedSampler(P['names'], post.my_posterior, post.my_params_transform, resume=True,
                                              log_dir=data.output_folder + str(data.name) + '/ultranest',
                                              vectorized=True)
    if mc['direction_generation'] == 'de-mix':
        direction_stepsampler = ultranest.stepsampler.generate_mixture_random_direction
    elif mc['direction_generation'] == 'region-slice':
        direction_stepsampler = ultranest.stepsampler.generate_region_oriented_direction
    elif mc['direction_generation'] == 'cube-ortho-harm':
        direction_stepsampler = ultranest.stepsampler.OrthogonalDirectionGenerator(
            ultranest.stepsampler.generate_random_direction)
    else:
        raise ValueError("direction_generation must be 'de-mix', 'region-slice' or 'cube-ortho-harm'")
    sampler.stepsampler = ultranest.stepsampler.SliceSampler(nsteps=20, generate_direction=direction_stepsampler)
    sampler.run(min_num_live_points=mc['live_points'], min_ess=mc['min_ess'],
                max_num_improvement_loops=mc['num_loops'])
    sampler.plot_run()
    sampler.plot_trace()
    sampler.plot_corner()
    return sampler


class CatalogPosterior:
    """ln_probab for a whole catalogue shard: rows of `params` carry a source index.  This is the
    entry point a batched multi-source sampler uses (SURVEY.md §8b `ln_probab_batch`)."""

    def __init__(self, ms: ModelSet, z, dlum, nus, fluxes, fluxerrs, device=0, ctx=None, device_tables=None):
        self.ms = ms
        self.engine = Engine(ms, device=device, ctx=ctx, device_tables=device_tables)
        self.src = _sources.make_sources(ms, z, dlum, nus, fluxes, fluxerrs)
        self.engine.set_sources(self.src)

    @property
    def pmin(self):
        return self.src.pmin

    @property
    def pmax(self):
        return self.src.pmax

    def ln_probab_batch(self, params, source_id):
        return self.engine.lnprob(params, source_id)

    def my_params_transform(self, cube, source_id):
        sid = np.asarray(source_id)
        return params_transform(cube, self.src.pmin[sid], self.src.pmax[sid])


# ---------------------------------------------------------------------------------------------
_FAMILY_ATTR = (("GALAXY", "GALAXYFdict"), ("STARBURST", "STARBURSTFdict"), ("BBB", "BBBFdict"),
                ("TORUS", "TORUSFdict"), ("AGN_RAD", "AGN_RADFdict"))


def fill_modelsdict(md, builder: DictBuilder, z, templates_4plot):
    """The filtering loop of MODELSDICT.construct_dictionaryarray_filtered
    (DICTIONARIES_AGNfitter.py:92-128) on the GPU: for every family in `templates_4plot`
    ({'GALAXY': {key: (log nu, Fnu)}, 'STARBURST': ..., 'BBB': ..., 'TORUS': ..., ['AGN_RAD': ...]},
    i.e. the reference's `*Fdict_4plot` dicts) set `md.<FAMILY>Fdict = {key: (bands, Fnu[n_b])}`
    with the reference's key order, and keep the `_4plot` dicts on `md` as the reference does."""
    for fam, attr in _FAMILY_ATTR:
        if fam not in templates_4plot:
            continue
        tpl = templates_4plot[fam]
        tab = builder.build_family(tpl, [z])[0]
        setattr(md, attr, {k: (builder.bands, tab[i]) for i, k in enumerate(tpl.keys())})
        setattr(md, attr + "_4plot", tpl)
    md.z = z
    return md


def construct_dictionaryarray_filtered(md, z, filterdict, model=None, builder=None):
    """MODELSDICT.construct_dictionaryarray_filtered(z, filterdict) (DICTIONARIES_AGNfitter.py:78-136) with the
    filtering loops on the GPU: the reference's own loaders (MODEL_AGNfitter.GALAXY / STARBURST / BBB / TORUS /
    AGN_RAD) produce the `_4plot` dictionaries, K2 filters every template of a family in one call, and `md` ends up
    with exactly the attributes the reference method sets (same names, same key order, `(bands, Fnu[n_b])` values).
    `model`: the reference's MODEL_AGNfitter module (default: the one DICTIONARIES_AGNfitter imported)."""
    import sys
    if model is None:
        model = sys.modules[type(md).__module__].model
    if builder is None:
        builder = DictBuilder(filterdict)
    md.GALAXYFdict, md.STARBURSTFdict, md.BBBFdict, md.TORUSFdict = dict(), dict(), dict(), dict()
    md.z = z
    md.filterdict = filterdict
    (md.GALAXYFdict_4plot, md.GALAXY_SFRdict, md.GALAXYatt_dict, galaxy_parnames, galaxy_partypes,
     md.GALAXYfunctions) = model.GALAXY(md.path, md.modelsettings)
    (md.STARBURSTFdict_4plot, md.STARBURST_LIRdict, starburst_parnames, starburst_partypes,
     md.STARBURSTfunctions) = model.STARBURST(md.path, md.modelsettings)
    md.BBBFdict_4plot, bbb_parnames, bbb_partypes, md.BBBfunctions = model.BBB(md.path, md.modelsettings, md.nXRaysdata)
    md.TORUSFdict_4plot, torus_parnames, torus_partypes, md.TORUSfunctions = model.TORUS(md.path, md.modelsettings)
    tpl = {'GALAXY': md.GALAXYFdict_4plot, 'STARBURST': md.STARBURSTFdict_4plot, 'BBB': md.BBBFdict_4plot,
           'TORUS': md.TORUSFdict_4plot}
    norm_parnames = ['GA', 'SB', 'BB', 'TO']
    norm_partypes = ['free', 'free', 'free', 'free']
    md.all_parnames = [galaxy_parnames, starburst_parnames, torus_parnames, bbb_parnames]
    md.all_partypes = [galaxy_partypes, starburst_partypes, torus_partypes, bbb_partypes]
    if md.modelsettings['RADIO'] == True:
        md.AGN_RADFdict = dict()
        md.AGN_RADFdict_4plot, agnrad_parnames, agnrad_partypes, md.AGN_RADfunctions = \
            model.AGN_RAD(md.path, md.modelsettings, md.nRADdata)
        tpl['AGN_RAD'] = md.AGN_RADFdict_4plot
        norm_parnames.append('RAD')
        norm_partypes.append('free')
        md.all_parnames.extend((agnrad_parnames, norm_parnames))
        md.all_partypes.extend((agnrad_partypes, norm_partypes))
    else:
        md.all_parnames.append(norm_parnames)
        md.all_partypes.append(norm_partypes)
    fill_modelsdict(md, builder, z, tpl)


def build(md, model=None):
    """MODELSDICT.build() (DICTIONARIES_AGNfitter.py:139-154): one dictionary per redshift of `md.z_array` (always
    `[data.z]` in the live path, RUN_AGNfitter_multi.py:176), attributes left on `md`, `md.MD[str(z)] = None` as in
    the reference (its construct method returns nothing)."""
    md.MD = dict()
    builder = None
    for z in md.z_array:
        filterdict = [md.fo.central_nu_array, md.fo.lambdas_dict, md.fo.factors_dict]
        if builder is None:
            builder = DictBuilder(filterdict)
        construct_dictionaryarray_filtered(md, z, filterdict, model=model, builder=builder)
        md.MD[str(z)] = None
