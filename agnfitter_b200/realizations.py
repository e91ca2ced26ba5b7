"""
SED realisations and integrated luminosities on the GPU (K4): host side.

Reference (paths relative to AGNfitter/functions):
    FLUXES_ARRAYS.fluxes              PLOTandWRITE_AGNfitter.py:566-799   -> FluxesArrays.fluxes / SedEngine
    `_4plot` get_model.get_fluxes     DICTIONARIES_AGNfitter.py:350-401   -> template modes 0-3 (K4)
    FLUXES2nuLnu_4plotting            PLOTandWRITE_AGNfitter.py:800-821   -> SedEngine.spectra (nu L_nu)
    integrated_luminosities           PLOTandWRITE_AGNfitter.py:824-964   -> SedEngine.int_lums

Everything that depends only on (template grids, all_nus_rest, output settings) is prepared here once with NumPy,
in the reference's operation order: the sorted template grids (interp1d's mergesort), their reddening curves
(hostmath.prevot_k / calzetti_k on the full-resolution grid), log10 of the templates that take no free-parameter
law (numpy.log10, as the reference), the bracket table of every grid point (numpy.interp's rule), the index
ranges of the integration bands (PLT:859,873,939,943).  Per-row arithmetic runs in the kernel.
There is no CPU fallback: without the CUDA library / an sm_100 device this module raises.
"""
import ctypes as C

import numpy as np

from . import hostmath
from ._lib import as_f64, as_i32, ptr
from .plotset import PlotSet, plotset_from_modelsdict

C_ASTROPY = 299792458.0          # astropy.constants.c: u.spectral() turns out['intlum_freqranges'] into Hz with it

COMPONENT_BIT = {'sb': 0, 'bb': 1, 'ga': 2, 'to': 3, 'rad': 4, 'bbdered': 5}
_SINGLE = {'sb': 'sb', 'bbb': 'bb', 'Lx0.5-2keV': 'bb', 'Lx2-8keV': 'bb', 'bbbdered': 'bbdered', 'gal': 'ga',
           'tor': 'to', 'agn_rad': 'rad'}
_FAM_INDEX = {"GA": 0, "SB": 1, "TO": 2, "BB": 3, "RAD": 4}


class agnf_sed_family(C.Structure):
    _fields_ = [("fam", C.c_int32), ("mode", C.c_int32), ("e_col", C.c_int32), ("a_col", C.c_int32),
                ("n_groups", C.c_int32), ("reserved", C.c_int32), ("T", C.c_int64),
                ("group", C.c_void_p), ("local", C.c_void_p), ("g_nlam", C.c_void_p), ("x", C.c_void_p),
                ("k", C.c_void_p), ("val", C.c_void_p), ("val2", C.c_void_p), ("jtab", C.c_void_p),
                ("wtab", C.c_void_p)]


def all_nus_rest(settings):
    """PLT:619-633."""
    lognu_max = 19 if (settings['BBB'] in ('D12_S', 'D12_K') or settings['XRAYS'] != False) else 16.5
    lognu_min = 9 if settings['RADIO'] == True else 10.5
    return np.arange(lognu_min, lognu_max, 0.001)


def ranges_to_hz(ranges, unit_m=1e-6):
    """(out['intlum_freqranges'] * out['intlum_freqranges_unit']).to(u.Hz, equivalencies=u.spectral()) (PLT:839)
    for a wavelength unit of `unit_m` metres (micron by default)."""
    return C_ASTROPY / (np.asarray(ranges, dtype=np.float64) * unit_m)


def bracket_table(xs, grid, split):
    """numpy.interp's bracket for every grid point inside the sorted template grid xs: j = (number of samples
    <= point) - 1; -1 where interp1d fills (outside [xs_0, xs_last]).  split: the reference interpolates the
    parts below / from 10^16.685 Hz separately (PLT:697-701) — points are bracketed inside their own part."""
    def one(x, g):
        if len(x) == 0:
            return np.full(len(g), -1, dtype=np.int64)
        j = np.searchsorted(x, g, side='right') - 1
        j[(g < x[0]) | (g > x[-1])] = -1
        return j
    if not split:
        return one(xs, grid).astype(np.int32)
    n_a = int(np.count_nonzero(xs < 16.685))
    lo = grid < 16.685
    out = np.empty(len(grid), dtype=np.int64)
    out[lo] = one(xs[:n_a], grid[lo])
    jb = one(xs[n_a:], grid[~lo])
    out[~lo] = np.where(jb >= 0, jb + n_a, -1)
    return out.astype(np.int32)


def bracket_weights(xs, grid, jt):
    """(grid - x_j) / (x_j+1 - x_j) for the brackets of bracket_table; 0 on a template sample and outside."""
    w = np.zeros(len(grid))
    inside = (jt >= 0) & (jt + 1 < len(xs))
    j = jt[inside]
    with np.errstate(all="ignore"):
        w[inside] = (grid[inside] - xs[j]) / (xs[j + 1] - xs[j])
    w[inside & (grid == xs[np.maximum(jt, 0)])] = 0.0
    return w


def family_mode(fam, par_offset):
    """(mode, e_col, a_col) of a family: which branch of the `_4plot` get_fluxes applies (DICTIONARIES:350-401)."""
    if not fam.fitted or 'free' not in fam.par_types:
        return 0, 0, 0
    names, types, n = fam.par_names, fam.par_types, fam.n_dim
    if names[-1] == 'EBVgal':
        return 1, par_offset + n - 1, 0
    if names[-1] == 'EBVbbb':
        return 2, par_offset + n - 1, 0
    if types[-2:] == ['free', 'free'] and names[-2:] == ['EBVbbb', 'alphaScat']:
        return 3, par_offset + n - 2, par_offset + n - 1
    if types[-3:] == ['free', 'free', 'grid'] and names[-3:] == ['EBVbbb', 'alphaScat', 'Gamma']:
        return 3, par_offset + n - 3, par_offset + n - 2
    raise ValueError("unsupported free-parameter layout %s / %s" % (names, types))


def band_plan(intlum_models, ranges_hz, grid, radio):
    """Translate out['intlum_models'] / out['intlum_freqranges'] into K4's bands and output columns, with the
    reference's index rules (PLT:859,873 / 939,943) and its per-name formulas (PLT:866-934)."""
    ranges_hz = np.asarray(ranges_hz, dtype=np.float64)
    bands, index_of = [], {}
    cols = []

    def band_for(f0, f1):
        if f0 == f1:
            key = (int((np.abs(grid - np.log10(f0))).argmin()), -1)
        else:
            sel = np.nonzero((grid >= np.log10(f1)) & (grid <= np.log10(f0)))[0]
            key = (int(sel[0]), int(sel[-1]) + 1) if len(sel) else (0, 0)
        if key not in index_of:
            index_of[key] = len(bands)
            bands.append([key[0], key[1], 0])
        return index_of[key]

    def bits(*names):
        m = 0
        for nm in names:
            if nm == 'rad' and not radio:
                raise ValueError("a radio luminosity was requested from a model set without AGN_RAD")
            m |= 1 << COMPONENT_BIT[nm]
        return m

    for m, name in enumerate(intlum_models):
        f0, f1 = ranges_hz[m]
        q = band_for(f0, f1)
        mono = bands[q][1] < 0
        combo = name in ('AGNfrac_IR', 'tor+bbb', 'AGNfracTO', 'AGNfracGA', 'AGNfracSB', 'AGNfrac_opt', 'gal+bbb',
                         'AGNfrac_rad') or name.find('AGNfrac') == 0
        if not combo:
            if name not in _SINGLE:
                raise ValueError("unknown entry %r in out['intlum_models']" % name)
            kind, num, den = 0, bits(_SINGLE[name]), 0
        elif mono:
            if name == 'AGNfrac':
                kind, num, den = 1, bits('to'), bits('to', 'sb', 'ga')
            elif name == 'tor+bbb':
                kind, num, den = 0, bits('to', 'bb'), 0
            elif name == 'AGNfrac_rad':
                kind, num, den = 1, bits('rad'), bits('to', 'sb', 'rad')
            else:
                raise ValueError("the reference appends nothing for a monochromatic %r (PLT:866-876)" % name)
        else:
            table = {
                'AGNfrac_IR': (1, ('to',), ('to', 'sb')),
                'AGNfracTO': (1, ('to',), ('to', 'sb', 'ga')),
                'AGNfracGA': (1, ('ga',), ('to', 'sb', 'ga')),
                'AGNfracSB': (1, ('sb',), ('to', 'sb', 'ga')),
                'AGNfrac_opt': (1, ('bb',), ('ga', 'bb')),
                'AGNfrac_rad': (1, ('rad',), ('to', 'sb', 'rad')),
                'gal+bbb': (0, ('ga', 'bb'), ()),
                'tor+bbb': (0, ('to', 'bb'), ()),
                'AGNfrac1-10GHz': (1, ('rad',), ('sb', 'rad')),
                'AGNfrac10-30GHz': (1, ('rad',), ('sb', 'rad')),
                'AGNfrac50-200GHz': (1, ('rad',), ('sb', 'rad')),
                'AGNfrac75-20mu': (1, ('to',), ('sb', 'to', 'ga', 'bb')),
                'AGNfrac20-8mu': (1, ('to',), ('sb', 'to', 'ga', 'bb')),
                'AGNfrac8-4mu': (1, ('to',), ('sb', 'to', 'ga', 'bb')),
                'AGNfrac4-2mu': (1, ('to', 'bb'), ('sb', 'to', 'ga', 'bb')),
                'AGNfrac2-1mu': (1, ('to', 'bb'), ('sb', 'to', 'ga', 'bb')),
                'AGNfrac1-0.5mu': (1, ('to', 'bb'), ('sb', 'to', 'ga', 'bb')),
                'AGNfrac0.5-0.3mu': (1, ('bb',), ('to', 'ga', 'bb')),
                'AGNfrac0.3-0.06mu': (1, ('bb',), ('to', 'ga', 'bb')),
            }
            if not radio:      # Lnurad_int = 0 (PLT:917-918, 922-923); with RADIO the reference reads an unbound name
                table['AGNfrac400-1600GHz'] = (1, ('to',), ('sb', 'to'))
                table['AGNfrac187-75mu'] = (1, ('to',), ('sb', 'to', 'ga'))
            if name not in table:
                raise ValueError("%r over a band reads an unbound variable in the reference (PLT:880-934)" % name)
            kind, num_n, den_n = table[name]
            num, den = bits(*num_n), bits(*den_n)
        bands[q][2] |= num | den
        cols.append((kind, q, num, den))
    return np.array(bands, dtype=np.int32).reshape(-1, 3), np.array(cols, dtype=np.int32).reshape(-1, 4)


class SedEngine:
    """K4 bound to an Engine (model set + sources already uploaded) and the matching PlotSet."""

    def __init__(self, engine, ps: PlotSet, src_z, src_dlum):
        self.engine, self.ctx, self.lib = engine, engine.ctx, engine.lib
        ms = engine.ms
        self.ms, self.ps = ms, ps
        S = ms.settings
        self.grid = all_nus_rest(S)
        G = len(self.grid)
        xg = as_f64(10 ** self.grid)
        keep = np.nonzero(~(self.grid > 16) & ~(self.grid < 11.7))[0]          # PLT:733-734
        to_lo, to_hi = (int(keep[0]), int(keep[-1]) + 1) if len(keep) else (0, 0)
        rad_hi = int(np.count_nonzero(~(self.grid >= 17.5)))                  # PLT:693
        self.n_comp = 7 if ms.radio else 6
        self.component_names = ['sb', 'bb', 'ga', 'to'] + (['rad'] if ms.radio else []) + ['total', 'bbdered']
        dered_row = as_i32(ps.dered_row)
        bands_jt = bands_wt = None
        if ps.dered_mode == 3:
            bands = as_f64(ms.bands)
            if np.any(np.diff(bands) <= 0):
                raise ValueError("dictionary band centres must be strictly ascending")
            bands_jt = bracket_table(bands, self.grid, False)
            bands_wt = as_f64(bracket_weights(bands, self.grid, bands_jt))
        grid_a = as_f64(self.grid)
        self.ctx.check(self.lib.agnf_sed_config(self.ctx.h, G, ptr(grid_a), ptr(xg), to_lo, to_hi, rad_hi,
                                                int(ps.dered_mode), ptr(dered_row), ptr(bands_jt), ptr(bands_wt)))
        split = (S['BBB'] != 'KD18') and (S['XRAYS'] == True)                  # PLT:697
        idxs = ms.idxs
        k_fit = 0
        for fam in ms.families:
            pf = ps.fam(fam.name)
            if pf.T != fam.T:
                raise ValueError("%s: %d full-resolution templates for %d dictionary rows" % (fam.name, pf.T, fam.T))
            mode, e_col, a_col = family_mode(fam, idxs[k_fit] if fam.fitted else 0)
            if fam.fitted:
                k_fit += 1
            self._upload_family(fam.name, pf, mode, e_col, a_col, split and fam.name == "BB")
        z = as_f64(np.atleast_1d(src_z))
        dl = as_f64(np.atleast_1d(src_dlum))
        if len(z) != len(dl):
            raise ValueError("z and dlum differ in length")
        self.ctx.check(self.lib.agnf_sed_sources(self.ctx.h, len(z), ptr(z), ptr(dl)))
        self.S = len(z)
        self.n_models = 0

    def _upload_family(self, name, pf, mode, e_col, a_col, split):
        need2 = name == "BB" and self.ps.dered_mode in (1, 2)
        xs_all, k_all, val_all, val2_all, jt_all, wt_all, nlam = [], [], [], [], [], [], []
        with np.errstate(all="ignore"):
            for g, x in enumerate(pf.lognu):
                x_used = np.log10(10 ** x) if mode == 1 else x          # Calzetti hands back log10(10**nu), DIC:366
                order = np.argsort(x_used, kind="mergesort")            # interp1d sorts x this way
                xs = x_used[order]
                F = pf.Fnu[g][:, order]
                if mode == 1:
                    k_all.append(hostmath.calzetti_k(10 ** x)[order])
                elif mode in (2, 3):
                    k_all.append(hostmath.prevot_k(x)[order])
                val_all.append((F if mode else np.log10(F)).reshape(-1))
                if need2:
                    val2_all.append(np.log10(F).reshape(-1))
                xs_all.append(xs)
                nlam.append(len(xs))
                jt_all.append(bracket_table(xs, self.grid, split))
                wt_all.append(bracket_weights(xs, self.grid, jt_all[-1]))
        d = agnf_sed_family()
        d.fam, d.mode, d.e_col, d.a_col = _FAM_INDEX[name], mode, e_col, a_col
        d.n_groups, d.T = len(pf.lognu), pf.T
        keep = [as_i32(pf.group), as_i32(pf.local), as_i32(nlam), as_f64(np.concatenate(xs_all)),
                as_f64(np.concatenate(k_all)) if mode else None, as_f64(np.concatenate(val_all)),
                as_f64(np.concatenate(val2_all)) if need2 else None, as_i32(np.concatenate(jt_all)),
                as_f64(np.concatenate(wt_all))]
        for field, arr in zip(("group", "local", "g_nlam", "x", "k", "val", "val2", "jtab", "wtab"), keep):
            setattr(d, field, None if arr is None else arr.ctypes.data)
        self.ctx.check(self.lib.agnf_sed_family_upload(self.ctx.h, C.byref(d)))

    def set_bands(self, intlum_models, ranges_hz):
        """out['intlum_models'] and the matching frequency ranges in Hz (ranges_to_hz for wavelength settings)."""
        bands, cols = band_plan(list(intlum_models), ranges_hz, self.grid, self.ms.radio)
        b = [as_i32(bands[:, i]) for i in range(3)]
        c = [as_i32(cols[:, i]) for i in range(4)]
        self.ctx.check(self.lib.agnf_sed_bands(self.ctx.h, len(bands), ptr(b[0]), ptr(b[1]), ptr(b[2]), len(cols),
                                               ptr(c[0]), ptr(c[1]), ptr(c[2]), ptr(c[3])))
        self.n_models = len(cols)

    def _rows(self, params, source_id):
        params = as_f64(np.atleast_2d(params))
        if params.shape[1] != self.ms.n_par:
            raise ValueError("params has %d columns, the model set %d parameters" % (params.shape[1], self.ms.n_par))
        sid = as_i32(np.zeros(len(params)) if source_id is None else source_id)
        if len(sid) != len(params):
            raise ValueError("source_id and params differ in length")
        return params, sid

    def int_lums(self, params, source_id=None):
        """[n, n_models] = integrated_luminosities(...).T, linear (the reference keeps log10 of it, PLT:797)."""
        if not self.n_models:
            raise RuntimeError("call set_bands first")
        params, sid = self._rows(params, source_id)
        out = np.empty((len(params), self.n_models), dtype=np.float64)
        self.ctx.check(self.lib.agnf_sed_intlum(self.ctx.h, ptr(params), ptr(sid), len(params), ptr(out), 0))
        return out

    def spectra(self, params, source_id=None):
        """[n, n_comp, G] nu L_nu in the order of nuLnus4plotting (PLT:782-786): see component_names."""
        params, sid = self._rows(params, source_id)
        out = np.empty((len(params), self.n_comp, len(self.grid)), dtype=np.float64)
        self.ctx.check(self.lib.agnf_sed_spectra(self.ctx.h, ptr(params), ptr(sid), len(params), ptr(out), 0))
        return out

    def int_lums_ensemble(self, walkers):
        """[S, W, n_models] for the walker state of the EnsembleSampler bound to the same context: the W final
        positions of each source are its posterior rows.  Positions never leave HBM."""
        if not self.n_models:
            raise RuntimeError("call set_bands first")
        out = np.empty((self.S, int(walkers), self.n_models), dtype=np.float64)
        self.ctx.check(self.lib.agnf_sed_intlum_ens(self.ctx.h, ptr(out)))
        return out

    def int_lums_device(self, params_t, sid_t, out_t):
        """Device-resident variant: torch CUDA tensors [n, n_par] f64, [n] i32, [n, n_models] f64."""
        self.ctx.check(self.lib.agnf_sed_intlum(self.ctx.h, params_t.data_ptr(), sid_t.data_ptr(),
                                                params_t.shape[0], out_t.data_ptr(), 1))


class FluxesArrays:
    """Mirror of PLOTandWRITE_AGNfitter.FLUXES_ARRAYS (PLT:541-799): same constructor, same `fluxes(data, models)`
    entry, same result attributes (`all_nus_rest`, `nuLnus4plotting`, `int_lums`, `int_lums_best`,
    `filtered_modelpoints_nuLnu`), computed by K4 / K1.

    `posterior` is the dropin.Posterior of the source (it owns the uploaded model set).  `out['intlum_freqranges']`
    is taken in the unit `out['intlum_freqranges_unit_m']` (metres per unit; default micron) when
    `out['intlum_freqranges_unit']` is not a plain number — astropy is not required."""

    def __init__(self, chain_obj, P, out, output_type, models_settings, mc_settings, posterior=None, rng=None):
        self.chain_obj, self.P, self.out, self.output_type = chain_obj, P, out, output_type
        self.models_settings, self.mc = models_settings, mc_settings
        self.posterior = posterior
        self.rng = rng if rng is not None else np.random

    def fluxes(self, data, models):
        from .dropin import Posterior
        self.chain_obj.props()
        post = self.posterior if self.posterior is not None else Posterior(data, models, self.P)
        md = models.dict_modelfluxes
        ps = plotset_from_modelsdict(md, post.ms)
        sed = SedEngine(post.engine, ps, [float(data.z)], [float(data.dlum)])
        flat = np.asarray(self.chain_obj.flatchain, dtype=np.float64)
        if self.output_type == 'best_fit':
            raise ValueError("FLUXES_ARRAYS.fluxes never defines realization_nr for 'best_fit' (PLT:612-613,635)")
        n_real = self.out['realizations2plot'] if self.output_type == 'plot' else self.out['realizations2int']
        par = flat[self.rng.choice(len(flat), n_real), :]                          # PLT:596,608
        par[-1] = self.chain_obj.best_fit_pars                                     # PLT:599,610
        self.all_nus_rest = sed.grid
        spec = sed.spectra(par)
        self.nuLnus4plotting = tuple(spec[:, c, :] for c in range(sed.n_comp))
        if self.output_type == 'plot':
            ym = post.engine.lnprob(par, np.zeros(len(par), dtype=np.int32), detail=True)[2]
            lumfactor = 4. * np.pi * float(data.dlum) ** 2.
            self.filtered_modelpoints_nuLnu = ym * lumfactor * 10 ** (np.asarray(data.nus))     # PLT:790-794
        if self.output_type == 'int_lums':
            unit = self.out.get('intlum_freqranges_unit_m', 1e-6)
            sed.set_bands(self.out['intlum_models'], ranges_to_hz(self.out['intlum_freqranges'], unit))
            with np.errstate(all="ignore"):
                self.int_lums = np.log10(sed.int_lums(par).T)                     # PLT:797
            self.int_lums_best = self.int_lums[:, -1]
        self.sed = sed
        return self
