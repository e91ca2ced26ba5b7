"""
TEST INFRASTRUCTURE — not part of the product path.

Drives the UNMODIFIED reference functions (imported through oracle/ref_shim.py)
on real template/filter files to produce ground truth for tests/golden/.
Only runs where /root/reference is mounted (this container, not the GPU box).

What comes from the reference verbatim (called, not restated):
    DICTIONARIES_AGNfitter.filtering_models / dictkey_arrays
    PARAMETERSPACE_AGNfitter.Pdict / ymodel / ln_prior / ln_likelihood / ln_probab
    PRIORS_AGNfitter.PRIORS (+ every prior it dispatches to)
    MODEL_AGNfitter.STARBURST('DH02_CE01') / BBB('R06'|'SN12') / TORUS('S04'|'NK0') / AGN_RAD
    MODEL_AGNfitter.BBBred_Prevot / GALAXYred_Calzetti
What is restated (oracle/loaders.py) because astropy is missing: filter-set assembly,
the BC03 galaxy loader, the S17 FITS loader, catalogue row preparation, z2Dlum.
"""
import numpy as np

from . import ref_shim, loaders


DEFAULT_SETTINGS = dict(
    path='models/', modelset='modelsv1',
    GALAXY='BC03', STARBURST='DH02_CE01', BBB='R06', TORUS='S04',
    XRAYS=False, RADIO=False, Blazar='None',
    PRIOR_energy_balance='Flexible', PRIOR_AGNfraction=True,
    PRIOR_midIR_UV=False, PRIOR_galaxy_only=False, turn_on_AGN=True,
)


class Bag(object):
    pass


def build_reference_dictionary(z, filterdict, settings, nRADdata=0, nXRaysdata=0,
                               galaxy_ebv="grid", galaxy_kw=None, starburst_kw=None, root=None):
    """Equivalent of MODELSDICT.construct_dictionaryarray_filtered (DICTIONARIES:78-136) with the
    restated loaders swapped in where the reference's cannot run.  Every template goes through the
    reference's own filtering_models."""
    mods = ref_shim.install()
    MOD, DIC = mods["MOD"], mods["DIC"]
    root = (root or ref_shim.REF_ROOT).rstrip("/") + "/"
    MD = Bag()
    MD.z = z
    MD.modelsettings = settings
    MD.filterdict = filterdict

    def filt(d4):
        out = dict()
        for c in d4.keys():
            nu, Fnu = d4[c]
            bands, F = DIC.filtering_models(nu, Fnu, filterdict, z)
            out[c] = bands, F.flatten()
        return out

    calz = MOD.GALAXYfunctions()[0]
    gkw = dict(galaxy_kw or {})
    (MD.GALAXYFdict_4plot, MD.GALAXY_SFRdict, MD.GALAXYatt_dict, gpn, gpt,
     MD.GALAXYfunctions) = loaders.galaxy_BC03(root, ebv_type=galaxy_ebv, reddening_fn=calz, **gkw)
    MD.GALAXYFdict = filt(MD.GALAXYFdict_4plot)

    if settings['STARBURST'] == 'S17':
        (MD.STARBURSTFdict_4plot, MD.STARBURST_LIRdict, spn, spt,
         MD.STARBURSTfunctions) = loaders.starburst_S17(root, **(starburst_kw or {}))
    else:
        (MD.STARBURSTFdict_4plot, MD.STARBURST_LIRdict, spn, spt,
         MD.STARBURSTfunctions) = MOD.STARBURST(root, settings)
    MD.STARBURSTFdict = filt(MD.STARBURSTFdict_4plot)

    MD.BBBFdict_4plot, bpn, bpt, MD.BBBfunctions = MOD.BBB(root, settings, nXRaysdata)
    MD.BBBFdict = filt(MD.BBBFdict_4plot)

    MD.TORUSFdict_4plot, tpn, tpt, MD.TORUSfunctions = MOD.TORUS(root, settings)
    MD.TORUSFdict = filt(MD.TORUSFdict_4plot)

    norm_parnames = ['GA', 'SB', 'BB', 'TO']
    norm_partypes = ['free'] * 4
    MD.all_parnames = [gpn, spn, tpn, bpn]
    MD.all_partypes = [gpt, spt, tpt, bpt]
    if settings['RADIO'] == True:
        MD.AGN_RADFdict_4plot, rpn, rpt, MD.AGN_RADfunctions = MOD.AGN_RAD(root, settings, nRADdata)
        MD.AGN_RADFdict = filt(MD.AGN_RADFdict_4plot)
        norm_parnames.append('RAD')
        norm_partypes.append('free')
        MD.all_parnames.extend((rpn, norm_parnames))
        MD.all_partypes.extend((rpt, norm_partypes))
    else:
        MD.all_parnames.append(norm_parnames)
        MD.all_partypes.append(norm_partypes)
    return MD


def make_models(MD, settings):
    """MODEL_AGNfitter.MODELS.DICTS (MOD:1483-1490) minus the _4plot twin."""
    mods = ref_shim.install()
    models = Bag()
    models.settings = settings
    models.z = MD.z
    models.dict_modelfluxes = MD
    models.dictkey_arrays = mods["DIC"].dictkey_arrays(MD)
    return models


def make_data(z, nus, fluxes, fluxerrs, dlum=None, name="src"):
    """DATA_AGNfitter.DATA (DAT:357-369) fields read on the hot path."""
    d = Bag()
    d.z = float(z)
    d.nus = np.asarray(nus, dtype=np.float64)
    d.fluxes = np.asarray(fluxes, dtype=np.float64)
    d.fluxerrs = np.asarray(fluxerrs, dtype=np.float64)
    d.dlum = float(loaders.z2Dlum(z) if dlum is None else dlum)
    d.lumfactor = 4. * np.pi * d.dlum ** 2.
    d.name = name
    return d


def reference_eval(pars, data, models, P):
    """Run the reference per walker; returns (lnprob, ymodel[n_b], matched row keys, lnprior, lnlike)."""
    mods = ref_shim.install()
    PAR = mods["PAR"]
    gal, sb, tor, bbb, rad = models.dictkey_arrays
    n = len(pars)
    lnp = np.empty(n)
    lnprior = np.empty(n)
    lnlike = np.empty(n)
    ym = None
    keys = []
    for i in range(n):
        p = tuple(float(x) for x in pars[i])
        y, bands = PAR.ymodel(data.nus, data.z, data.dlum, models, P, *p)
        if ym is None:
            ym = np.empty((n, len(y)))
        ym[i] = y
        k = [gal.matched_parkeys_grid, sb.matched_parkeys_grid, tor.matched_parkeys_grid,
             bbb.matched_parkeys_grid]
        if not isinstance(rad, str):
            if (rad.pars_modelkeys != ['-99.9']).all():
                k.append(rad.matched_parkeys_grid)
            else:
                k.append('-99.9')
        keys.append([tuple(x) if isinstance(x, (list, tuple)) else str(x) for x in k])
        with np.errstate(all="ignore"):
            lnprior[i] = PAR.ln_prior(data, models, P, *p)
            lnlike[i] = PAR.ln_likelihood(data.nus, data.fluxes, data.fluxerrs, data.z, y, models)
            lnp[i] = PAR.ln_probab(p, data, models, P)
    return lnp, ym, keys, lnprior, lnlike


# ------------------------------------------------------------------------------------------------
# SED realisations + integrated luminosities (PLOTandWRITE_AGNfitter.FLUXES_ARRAYS)
# ------------------------------------------------------------------------------------------------
C_ASTROPY = 299792458.0          # astropy.constants.c [m/s]: what u.spectral() converts wavelengths with


class _FreqScalar(object):
    """Stands in for an astropy Quantity element: `.value` in Hz and equality on it (PLT:859,873,956)."""

    def __init__(self, v):
        self.value = float(v)

    def __eq__(self, other):
        return self.value == other.value


class _FreqRanges(object):
    """out['intlum_freqranges'] for the unmodified integrated_luminosities: `ranges * unit` and
    `.to(u.Hz, equivalencies=u.spectral())` (PLT:839) are astropy operations; here the conversion
    nu = c / lambda is done once, up front, and both operations return the converted table."""

    def __init__(self, ranges, unit_m):
        self.hz = C_ASTROPY / (np.asarray(ranges, dtype=np.float64) * unit_m)

    def __mul__(self, other):
        return self

    def to(self, *a, **k):
        return self

    def __getitem__(self, m):
        return [_FreqScalar(self.hz[m][0]), _FreqScalar(self.hz[m][1])]


def intlum_ranges_hz(ranges, unit_m=1e-6):
    """The Hz table the reference's astropy conversion yields for out['intlum_freqranges'] (microns by default)."""
    return C_ASTROPY / (np.asarray(ranges, dtype=np.float64) * unit_m)


class _Chain(object):
    """CHAIN after props() (PLT:438-492): the thinned flat chain and the best-fit row."""

    def __init__(self, flatchain, best):
        self.flatchain = np.array(flatchain, dtype=np.float64)
        self.best_fit_pars = np.array(best, dtype=np.float64)

    def props(self):
        pass


def make_models_4plot(MD, settings):
    mods = ref_shim.install()
    models = make_models(MD, settings)
    models.dictkey_arrays_4plot = mods["DIC"].dictkey_arrays_4plot(MD)
    return models


def reference_fluxes(data, models, P, flatchain, best, intlum_models, intlum_ranges_um, n_real, seed,
                     output_type='int_lums'):
    """Run the UNMODIFIED FLUXES_ARRAYS.fluxes (PLT:566-799).  The rows it draws with np.random.choice are
    reproduced by seeding NumPy's legacy generator; returns (rows used, all_nus_rest, nuLnus tuple, int_lums)."""
    PLT = ref_shim.install_plot()
    mods = ref_shim.install()
    mods["MOD"].z2Dlum = loaders.z2Dlum            # astropy cosmology is absent: restated FlatLambdaCDM distance
    out = {'realizations2int': n_real, 'realizations2plot': n_real, 'intlum_models': list(intlum_models),
           'intlum_freqranges': _FreqRanges(intlum_ranges_um, 1e-6), 'intlum_freqranges_unit': None}
    mc = {'sampling_algorithm': 'emcee'}
    fa = PLT.FLUXES_ARRAYS(_Chain(flatchain, best), P, out, output_type, models.settings, mc)
    np.random.seed(seed)
    pick = np.random.choice(len(flatchain), n_real)
    rows = np.array(flatchain, dtype=np.float64)[pick, :]
    rows[-1] = best
    np.random.seed(seed)
    with np.errstate(all="ignore"):
        fa.fluxes(data, models)
    return rows, fa.all_nus_rest, fa.nuLnus4plotting, (fa.int_lums if output_type == 'int_lums' else None)
