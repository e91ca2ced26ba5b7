"""
TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the UNMODIFIED reference
functions (through oracle/ref_shim.py + oracle/ref_harness.py) on the real template and
filter files under /root/reference.  Run in the build container only:

    python -m oracle.make_golden            # writes tests/golden/k1_*.npz, k2_*.npz
    python -m oracle.make_golden sed        # writes tests/golden/sed_*.npz (SED realisations)
    python -m oracle.make_golden chain      # writes tests/golden/chain_props.npz (CHAIN.props + percentiles, N2)

Each K1 file holds: the flattened ModelSet (agnfitter_b200.flatten arrays), the raw
per-source rows (nus, fluxes, fluxerrs, z, dlum), the reference's P['min']/P['max'],
walker positions, and the reference's outputs for them:
    ref_lnp       PARAMETERSPACE.ln_probab
    ref_ymodel    PARAMETERSPACE.ymodel
    ref_rows      row index (key insertion order) of the keys pick_nD matched, per family
    ref_lnprior   PARAMETERSPACE.ln_prior
    ref_lnlike    PARAMETERSPACE.ln_likelihood
    ref_mask      the x_valid index set of ln_likelihood as a bool vector
Each K2 file holds templates (log nu, Fnu), filter curves, z values and the reference's
DICTIONARIES.filtering_models band fluxes.

Substitution stated everywhere: the galaxy libraries named by the example
settings are absent from the mount; models/GALAXY/BC03_275seds.pickle is used through the
'BC03' code path (grid E(B-V)) and through the free-E(B-V) variant of the same loader.
"""
import os
import sys
import warnings

import numpy as np

from . import ref_shim, loaders, ref_harness as H
from agnfitter_b200 import flatten, sources

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
warnings.filterwarnings("ignore")


def walker_positions(rng, ms, pmin, pmax, n_uniform, n_ball, n_out, n_edge):
    """Mixture used by every golden/parity test: uniform in the box, emcee-like
    ball around mid-box (PARAMETERSPACE:263), out-of-box rows, and snapping edge cases
    (exact grid values, midpoints between neighbouring grid values, exact bounds)."""
    n_par = len(pmin)
    blocks = [rng.uniform(pmin, pmax, size=(n_uniform, n_par))]
    mid = 0.5 * (pmin + pmax)
    blocks.append(mid + (2 * rng.uniform(size=(n_ball, n_par)) - 1) * 0.00001)
    out = rng.uniform(pmin, pmax, size=(n_out, n_par))
    for r in range(n_out):
        j = rng.integers(n_par)
        out[r, j] = pmax[j] + rng.uniform(1e-9, 3.0) if rng.uniform() < 0.5 else pmin[j] - rng.uniform(1e-9, 3.0)
    blocks.append(out)
    edge = rng.uniform(pmin, pmax, size=(n_edge, n_par))
    col = 0
    for f in ms.families:
        if not f.fitted:
            continue
        for d in range(f.n_dim):
            if f.par_types[d] == 'grid':
                g = np.sort(f.grid[d])
                for r in range(n_edge):
                    kind = rng.integers(4)
                    i = rng.integers(len(g))
                    if kind == 0:
                        edge[r, col] = g[i]
                    elif kind == 1 and len(g) > 1:
                        i = min(i, len(g) - 2)
                        edge[r, col] = (g[i] + g[i + 1]) / 2
                    elif kind == 2 and len(g) > 1:
                        i = min(i, len(g) - 2)
                        edge[r, col] = g[i] / 2 + g[i + 1] / 2
                    else:
                        edge[r, col] = pmin[col] if rng.uniform() < 0.5 else pmax[col]
            col += 1
    blocks.append(edge)
    return np.concatenate(blocks, axis=0)


def rows_from_keys(ms, keys):
    out = np.zeros((len(keys), len(ms.families)), dtype=np.int32)
    for fi, f in enumerate(ms.families):
        pos = {k: t for t, k in enumerate(f.keys)}
        for r, k in enumerate(keys):
            kk = k[fi]
            out[r, fi] = pos[tuple(kk) if isinstance(kk, tuple) else kk]
    return out


def synth_fluxes(rng, mods, models, data_nus, z, dlum, P, frac_ul=0.1, frac_missing=0.05):
    """y = ymodel(truth)*(1+0.1 N), sigma = 0.1|y|; some bands -> upper limit / missing.
    Returned in catalogue units of mJy so that the DATA transformation applies."""
    truth = rng.uniform(P['min'], P['max'])
    # keep the truth physically tame so that fluxes are finite and positive
    y, _ = mods["PAR"].ymodel(data_nus, z, dlum, models, P, *tuple(truth))
    y = np.abs(y) * (1 + 0.1 * rng.normal(size=len(y)))
    y = np.abs(y) / 1e-26
    err = 0.1 * np.abs(y)
    u = rng.uniform(size=len(y))
    err[u < frac_ul] = -99.
    # bands the truth model does not reach (flux 0 or denormal-small) carry no datum
    miss = ((u >= frac_ul) & (u < frac_ul + frac_missing)) | ~(y > 1e-12)
    y[miss] = -99.
    err[miss] = -99.
    return y, err


def run_case(name, settings, filterset, zs, catalog=None, galaxy_ebv="grid", galaxy_kw=None,
             starburst_kw=None, n_pos=(400, 60, 40, 100), seed=1, nRADdata=None, nXRaysdata=None,
             force_uv_detection=True, force_xray_detection=False, write=True, eval_rows=None):
    """write=False, eval_rows=N: regenerate the case in memory and run the reference on the first N positions of
    every source only (tests/test_golden_live.py compares them bit for bit with the stored file)."""
    mods = ref_shim.install()
    rng = np.random.default_rng(seed)
    allf = loaders.load_all_filters()
    fd, fnames = loaders.make_filterdict(filterset, allf)
    data_nus = loaders.data_nus_for(filterset)
    n_b = len(fd[0])

    if catalog is not None:
        cat_names, cat_z, cat_fl, cat_fe = catalog
        zs = list(cat_z)
    S = len(zs)
    dl = [float(loaders.z2Dlum(z)) for z in zs]

    # raw catalogue rows -> DATA arrays
    if catalog is not None:
        # catalogue columns are in the filters' column order (SETTINGS:106-154 -> index 0 = GALEX_1500 ...);
        # central wavelengths come from ALL_FILTERS_info.dat and are then sorted (DATA:89-96,180-183)
        info_names = np.loadtxt(os.path.join(ref_shim.REF_ROOT, "models/FILTERS/ALL_FILTERS_info.dat"),
                                delimiter="|", usecols=[1], skiprows=1, dtype=str)
        info_nu = np.loadtxt(os.path.join(ref_shim.REF_ROOT, "models/FILTERS/ALL_FILTERS_info.dat"),
                             delimiter="|", usecols=[3], skiprows=1)
        col_order = list(reversed(loaders.FILTERSET_EXAMPLE49))   # column 0 = GALEX_1500
        nus0 = np.array([info_nu[list(info_names).index(nm)] for nm in col_order])
        rows = loaders.catalog_rows(cat_fl, cat_fe, nus0, np.array(zs), flux_unit_cgs=1e-26)
    else:
        rows = None

    MDs, models_l, P_l, data_l = [], [], [], []
    for s, z in enumerate(zs):
        if rows is not None:
            nR = int(rows["nRADdata"][s]) if nRADdata is None else nRADdata
            nX = int(rows["nXRaysdata"][s]) if nXRaysdata is None else nXRaysdata
        else:
            nR = 0 if nRADdata is None else nRADdata
            nX = 0 if nXRaysdata is None else nXRaysdata
        MD = H.build_reference_dictionary(z, fd, settings, nRADdata=nR, nXRaysdata=nX,
                                          galaxy_ebv=galaxy_ebv, galaxy_kw=galaxy_kw, starburst_kw=starburst_kw)
        MDs.append(MD)
        models_l.append(H.make_models(MD, settings))
    # all sources of one case must share key structure (same nRAD/nXR class)
    ms = flatten.modelset_from_modelsdicts(MDs, settings)

    src_nus = np.empty((S, n_b))
    src_fl = np.empty((S, n_b))
    src_fe = np.empty((S, n_b))
    for s, z in enumerate(zs):
        if rows is not None:
            src_nus[s], src_fl[s], src_fe[s] = rows["nus"][s], rows["fluxes"][s], rows["fluxerrs"][s]
        else:
            d0 = H.make_data(z, data_nus, np.ones(n_b), np.ones(n_b), dlum=dl[s])
            P0 = mods["PAR"].Pdict(d0, models_l[s])
            y, err = synth_fluxes(rng, mods, models_l[s], data_nus, z, dl[s], P0)
            if force_uv_detection:
                # prior_AGNfraction needs one positive flux with 14.3 < nu_rest < 15.3 (PRIORS:163-164)
                r = data_nus + np.log10(1 + z)
                win = np.nonzero((14.3 < r) & (r < 15.3))[0]
                if len(win) and not np.any((y[win] > 0)):
                    y[win[-1]] = abs(rng.normal(1.0, 0.1)) * 1e-3
                    err[win[-1]] = 0.1 * y[win[-1]]
            if force_xray_detection:
                # prior_IR_XRays needs a datum with 17.685 < nu_rest < 18.384 (PRIORS:316-317); a disc without the
                # X-ray extension predicts no flux there, so the synthetic catalogue would mark it missing
                r = data_nus + np.log10(1 + z)
                for j in np.nonzero((17.685 < r) & (r < 18.384))[0]:
                    y[j] = abs(rng.normal(1.0, 0.3)) * 1e-6
                    err[j] = 0.2 * y[j]
            rr = loaders.catalog_rows(y[None, :], err[None, :], data_nus, np.array([z]), flux_unit_cgs=1e-26)
            src_nus[s], src_fl[s], src_fe[s] = rr["nus"][0], rr["fluxes"][0], rr["fluxerrs"][0]
        data_l.append(H.make_data(z, src_nus[s], src_fl[s], src_fe[s], dlum=dl[s]))
        P_l.append(mods["PAR"].Pdict(data_l[s], models_l[s]))

    n_par = len(P_l[0]['names'])
    assert P_l[0]['names'] == ms.par_names, (P_l[0]['names'], ms.par_names)
    pars_all, sid_all = [], []
    ref = dict(lnp=[], ym=[], rows=[], lnprior=[], lnlike=[])
    pmin = np.array([P['min'] for P in P_l], dtype=np.float64)
    pmax = np.array([P['max'] for P in P_l], dtype=np.float64)
    ref_mask = np.zeros((S, n_b), dtype=bool)
    for s in range(S):
        pars = walker_positions(rng, ms, pmin[s], pmax[s], *n_pos)
        if eval_rows is not None:            # an even sample over the four kinds of positions
            pars = pars[np.unique(np.linspace(0, len(pars) - 1, eval_rows).astype(int))]
        lnp, ym, keys, lpr, lll = H.reference_eval(pars, data_l[s], models_l[s], P_l[s])
        pars_all.append(pars)
        sid_all.append(np.full(len(pars), s, dtype=np.int32))
        ref["lnp"].append(lnp)
        ref["ym"].append(ym)
        ref["rows"].append(rows_from_keys(ms, keys))
        ref["lnprior"].append(lpr)
        ref["lnlike"].append(lll)
        # x_valid exactly as the reference evaluates it (PARAMETERSPACE:144-147)
        x, y, z = data_l[s].nus, data_l[s].fluxes, data_l[s].z
        if settings['XRAYS'] == 'Prior':
            sel = (x < np.log10(10 ** (15.38) / (1 + z))) & (y > -99.e-23)
        else:
            sel = (x < np.log10(10 ** (15.38) / (1 + z))) | (x > np.log10(10 ** (16.685) / (1 + z))) & (y > -99.e-23)
        ref_mask[s] = sel

    arrs = flatten.modelset_to_arrays(ms, prefix="ms_")
    arrs.update(dict(
        src_z=np.array(zs, dtype=np.float64), src_dlum=np.array(dl), src_nus=src_nus, src_fluxes=src_fl,
        src_fluxerrs=src_fe, ref_pmin=pmin, ref_pmax=pmax,
        pars=np.concatenate(pars_all), source_id=np.concatenate(sid_all),
        ref_lnp=np.concatenate(ref["lnp"]), ref_ymodel=np.concatenate(ref["ym"]),
        ref_rows=np.concatenate(ref["rows"]), ref_lnprior=np.concatenate(ref["lnprior"]),
        ref_lnlike=np.concatenate(ref["lnlike"]), ref_mask=ref_mask,
        filter_names=np.array(fnames),
    ))
    if not write:
        return arrs
    path = os.path.join(OUT, "k1_%s.npz" % name)
    np.savez_compressed(path, **arrs)
    fin = np.isfinite(np.concatenate(ref["lnp"]))
    print("%-22s S=%d n_b=%d n_par=%d rows=%d finite=%d  T=%s  -> %.2f MB" % (
        name, S, n_b, n_par, len(arrs["pars"]), fin.sum(), [f.T for f in ms.families],
        os.path.getsize(path) / 1e6))
    return ms


def k2_case(name, filterset, zs, seed=3, write=True):
    """Reference filtering_models on a sample of every loadable family."""
    mods = ref_shim.install()
    MOD, DIC = mods["MOD"], mods["DIC"]
    rng = np.random.default_rng(seed)
    root = ref_shim.REF_ROOT + "/"
    allf = loaders.load_all_filters()
    fd, fnames = loaders.make_filterdict(filterset, allf)
    bands = fd[0]
    S = dict(H.DEFAULT_SETTINGS)
    fams = {}
    g4, *_ = loaders.galaxy_BC03(root, reddening_fn=MOD.GALAXYfunctions()[0],
                                 age_slice=[0, 7, 13, 19, 24], tau_slice=[0, 5, 10])
    fams["GA"] = [g4[k] for k in list(g4.keys())[::15]]
    s4, *_ = MOD.STARBURST(root, S)
    fams["SB_DH02"] = [s4[k] for k in list(s4.keys())[::9]]
    s17, *_ = loaders.starburst_S17(root, tdust_slice=[0, 70, 149], fracpah_slice=[0, 30, 63])
    fams["SB_S17"] = list(s17.values())
    t4, *_ = MOD.TORUS(root, S)
    fams["TO_S04"] = [t4[k] for k in list(t4.keys())[::6]]
    b4, *_ = MOD.BBB(root, S, 0)
    fams["BB_R06"] = [b4['0.0']]
    Sx = dict(S, XRAYS=True)
    bx, *_ = MOD.BBB(root, Sx, 2)
    fams["BB_R06X"] = [bx[k] for k in list(bx.keys())[::20]]
    Sr = dict(S, RADIO=True)
    r4, *_ = MOD.AGN_RAD(root, Sr, 4)
    fams["RAD4"] = [r4[k] for k in list(r4.keys())[::400]]
    arrs = dict(bands=bands, z=np.array(zs, dtype=np.float64), filter_names=np.array(fnames),
                family_names=np.array(list(fams.keys())))
    lam_all = np.concatenate([np.asarray(fd[1][b][0]) for b in bands])
    fac_all = np.concatenate([np.asarray(fd[2][b][0]) for b in bands])
    off = np.cumsum([0] + [len(fd[1][b][0]) for b in bands])
    arrs.update(filt_lambda=lam_all, filt_factor=fac_all, filt_offset=off.astype(np.int64))
    n_entries = 0
    for fname, tpl in fams.items():
        nmax = max(len(t[0]) for t in tpl)
        lognu = np.full((len(tpl), nmax), np.nan)
        Fnu = np.full((len(tpl), nmax), np.nan)
        nlen = np.zeros(len(tpl), dtype=np.int64)
        out = np.empty((len(zs), len(tpl), len(bands)))
        for i, (nu, F) in enumerate(tpl):
            nu = np.asarray(nu, dtype=np.float64)
            F = np.asarray(F, dtype=np.float64).reshape(-1)
            nlen[i] = len(nu)
            lognu[i, :len(nu)] = nu
            Fnu[i, :len(nu)] = F
            for zi, z in enumerate(zs):
                _, f = DIC.filtering_models(nu, F, fd, z)
                out[zi, i] = np.asarray(f).flatten()
        arrs["%s_lognu" % fname] = lognu
        arrs["%s_Fnu" % fname] = Fnu
        arrs["%s_len" % fname] = nlen
        arrs["%s_ref" % fname] = out
        n_entries += out.size
    if not write:
        return arrs
    path = os.path.join(OUT, "k2_%s.npz" % name)
    np.savez_compressed(path, **arrs)
    print("%-22s bands=%d z=%d entries=%d -> %.2f MB" % (name, len(bands), len(zs), n_entries,
                                                        os.path.getsize(path) / 1e6))


def k2_write(name, fd, fnames, zs, fams, DIC, write=True):
    """Run the reference's filtering_models over `fams` ({name: [(log nu, Fnu), ...]}) and store inputs + outputs."""
    bands = fd[0]
    arrs = dict(bands=bands, z=np.array(zs, dtype=np.float64), filter_names=np.array(fnames),
                family_names=np.array(list(fams.keys())))
    lam_all = np.concatenate([np.asarray(fd[1][b][0]) for b in bands])
    fac_all = np.concatenate([np.asarray(fd[2][b][0]) for b in bands])
    off = np.cumsum([0] + [len(fd[1][b][0]) for b in bands])
    arrs.update(filt_lambda=lam_all, filt_factor=fac_all, filt_offset=off.astype(np.int64))
    n_entries = 0
    for fname, tpl in fams.items():
        nmax = max(len(t[0]) for t in tpl)
        lognu = np.full((len(tpl), nmax), np.nan)
        Fnu = np.full((len(tpl), nmax), np.nan)
        nlen = np.zeros(len(tpl), dtype=np.int64)
        out = np.empty((len(zs), len(tpl), len(bands)))
        for i, (nu, F) in enumerate(tpl):
            nu = np.asarray(nu, dtype=np.float64)
            F = np.asarray(F, dtype=np.float64).reshape(-1)
            nlen[i] = len(nu)
            lognu[i, :len(nu)] = nu
            Fnu[i, :len(nu)] = F
            for zi, z in enumerate(zs):
                _, f = DIC.filtering_models(nu, F, fd, z)
                out[zi, i] = np.asarray(f).flatten()
        arrs["%s_lognu" % fname] = lognu
        arrs["%s_Fnu" % fname] = Fnu
        arrs["%s_len" % fname] = nlen
        arrs["%s_ref" % fname] = out
        n_entries += out.size
    if not write:
        return arrs
    path = os.path.join(OUT, "k2_%s.npz" % name)
    np.savez_compressed(path, **arrs)
    print("%-22s bands=%d z=%d entries=%d -> %.2f MB" % (name, len(bands), len(zs), n_entries,
                                                        os.path.getsize(path) / 1e6))


def k2_families_case(name, filterset, zs, write=True):
    """The remaining template families of SURVEY §8a-T, loaded by the reference's OWN loaders
    (MODEL_AGNfitter.TORUS / BBB under the import shims): every torus option and the SN12 / KD18 discs."""
    mods = ref_shim.install()
    MOD, DIC = mods["MOD"], mods["DIC"]
    root = ref_shim.REF_ROOT + "/"
    fd, fnames = loaders.make_filterdict(filterset, loaders.load_all_filters())
    S = dict(H.DEFAULT_SETTINGS)
    fams = {}
    for opt, n_keep in (("NK0", 5), ("NK0_2P", 6), ("NK0_3P", 8), ("SKIRTOR_1P", 5), ("SKIRTOR_2P", 6),
                        ("SKIRTOR_3P", 8), ("CAT3D_3P", 8)):
        d4, *_ = MOD.TORUS(root, dict(S, TORUS=opt))
        keys = list(d4.keys())
        fams["TO_" + opt] = [d4[k] for k in keys[::max(1, len(keys) // n_keep)][:n_keep]]
    for opt, n_keep in (("SN12", 8), ("KD18", 8)):
        d4, *_ = MOD.BBB(root, dict(S, BBB=opt), 0)
        keys = list(d4.keys())
        fams["BB_" + opt] = [d4[k] for k in keys[::max(1, len(keys) // n_keep)][:n_keep]]
    return k2_write(name, fd, fnames, zs, fams, DIC, write=write)


SED_MODELS = ['sb', 'bbb', 'bbbdered', 'gal', 'AGNfrac_opt', 'tor', 'tor+bbb', 'AGNfrac_IR', 'gal', 'bbb',
              'AGNfrac_opt', 'tor', 'sb', 'tor+bbb', 'AGNfracTO', 'AGNfrac75-20mu', 'AGNfrac0.5-0.3mu']
# first 13 rows: example/SETTINGS_AGNfitter.py:316-321 (microns)
SED_RANGES_UM = [[8., 1000.], [0.1, 1.], [0.1, 1.], [0.1, 1.], [0.1, 1.], [1., 30.], [0.1, 30], [8., 1000.],
                 [0.4, 0.5], [0.4, 0.5], [0.4, 0.5], [6, 6], [1., 30.], [12., 12.], [1., 100.], [20., 75.], [0.3, 0.5]]


def sed_case(name, settings, filterset, zs, nR=0, nX=0, galaxy_ebv="grid", extra=(), n_real=10, seed=41, stride=7,
             write=True, galaxy_kw=None):
    """SED realisations + integrated luminosities: the reference's own FLUXES_ARRAYS.fluxes
    (PLOTandWRITE_AGNfitter.py:566-799) on rows drawn from a synthetic flat chain, for every source of the case.
    Stores the model set, the full-resolution templates, the rows, nuLnu on every `stride`-th frequency and the
    integrated luminosities (log10, as the reference keeps them)."""
    from agnfitter_b200 import plotset
    mods = ref_shim.install()
    rng = np.random.default_rng(seed)
    fd, fnames = loaders.make_filterdict(filterset, loaders.load_all_filters())
    data_nus = loaders.data_nus_for(filterset)
    n_b = len(fd[0])
    gk = dict(dict(age_slice=[0, 12, 24], tau_slice=[0, 10]), **(galaxy_kw or {}))
    sk = dict(tdust_slice=range(0, 150, 30), fracpah_slice=range(0, 64, 16))
    im = SED_MODELS + [e[0] for e in extra]
    ir = np.array(SED_RANGES_UM + [e[1] for e in extra], dtype=np.float64)
    MDs, pars_all, sid_all, nl_all, il_all, pmin, pmax, dl = [], [], [], [], [], [], [], []
    for s, z in enumerate(zs):
        MD = H.build_reference_dictionary(z, fd, settings, nRADdata=nR, nXRaysdata=nX, galaxy_ebv=galaxy_ebv,
                                          galaxy_kw=gk, starburst_kw=sk)
        MDs.append(MD)
        models = H.make_models_4plot(MD, settings)
        d = H.make_data(z, data_nus, np.ones(n_b) * 1e-27, np.ones(n_b) * 1e-28)
        P = mods["PAR"].Pdict(d, models)
        chain = rng.uniform(P['min'], P['max'], size=(200, len(P['min'])))
        rows, nus, nuLnus, il = H.reference_fluxes(d, models, P, chain, chain[17], im, ir, n_real, seed + s)
        pars_all.append(rows)
        sid_all.append(np.full(len(rows), s, dtype=np.int32))
        nl_all.append(np.stack([a[:, ::stride] for a in nuLnus], axis=1))       # [R, n_comp, G/stride]
        il_all.append(il.T)                                                      # [R, n_models]
        pmin.append(P['min'])
        pmax.append(P['max'])
        dl.append(d.dlum)
    ms = flatten.modelset_from_modelsdicts(MDs, settings)
    ps = plotset.plotset_from_modelsdict(MDs[0], ms)
    arrs = flatten.modelset_to_arrays(ms, prefix="ms_")
    arrs.update(plotset.plotset_to_arrays(ps))
    arrs.update(dict(
        src_z=np.array(zs, dtype=np.float64), src_dlum=np.array(dl), src_nus=np.tile(data_nus, (len(zs), 1)),
        src_fluxes=np.full((len(zs), n_b), 1e-27), src_fluxerrs=np.full((len(zs), n_b), 1e-28),
        ref_pmin=np.array(pmin, dtype=np.float64), ref_pmax=np.array(pmax, dtype=np.float64),
        pars=np.concatenate(pars_all), source_id=np.concatenate(sid_all),
        intlum_models=np.array(im), intlum_ranges_um=ir, intlum_ranges_hz=H.intlum_ranges_hz(ir),
        all_nus_rest=nus, stride=np.array(stride),
        ref_nuLnu=np.concatenate(nl_all), ref_int_lums=np.concatenate(il_all)))
    if not write:
        return arrs
    path = os.path.join(OUT, "sed_%s.npz" % name)
    np.savez_compressed(path, **arrs)
    print("%-22s S=%d rows=%d G=%d comps=%d models=%d T=%s -> %.2f MB" % (
        name, len(zs), len(arrs["pars"]), len(nus), arrs["ref_nuLnu"].shape[1], len(im),
        [f.T for f in ms.families], os.path.getsize(path) / 1e6))


def sed_main():
    os.makedirs(OUT, exist_ok=True)
    base = dict(H.DEFAULT_SETTINGS)
    # classic model set: R06 disc with free E(B-V) (dereddened curve = template '0.0'), ragged DH02_CE01 grids
    sed_case("classic18", base, loaders.FILTERSET_DEFAULT, [0.283, 1.58])
    # free E(B-V)_gal (Calzetti on the full template), X-ray extended disc interpolated in two parts, S17 starburst
    sed_case("freeebv_xray49", dict(base, STARBURST='S17', XRAYS=True), loaders.FILTERSET_EXAMPLE49, [0.129],
             nX=1, galaxy_ebv="free", seed=43)
    # X-ray disc with the Gamma grid dimension
    sed_case("xray_gamma49", dict(base, XRAYS=True), loaders.FILTERSET_EXAMPLE49, [0.031024], nX=3, seed=44)
    # five components: fitted synchrotron family, two-parameter SKIRTOR torus, radio-side luminosity fractions
    sed_case("radio_r06_49", dict(base, RADIO=True, TORUS='SKIRTOR_2P'), loaders.FILTERSET_EXAMPLE49, [1.234],
             nR=3, seed=45, extra=[('agn_rad', [1e4, 3e5]), ('AGNfrac_rad', [1e4, 3e5]),
                                   ('AGNfrac1-10GHz', [3e4, 3e5]), ('AGNfrac_rad', [2e5, 2e5])])
    # fixed radio template (one radio datum), KD18 disc (no BB norm: 4 pi dlum^2 path, dereddened = reddened)
    sed_case("radio1_kd18_49", dict(base, RADIO=True, BBB='KD18', TORUS='NK0'), loaders.FILTERSET_EXAMPLE49, [0.5],
             nR=1, seed=46)
    # SN12 disc without X-rays: the dereddened curve is interpolated over the band centres of the FILTERED
    # dictionary (PLOTandWRITE_AGNfitter.py:714-717)
    sed_case("radio2_sn12_49", dict(base, RADIO=True, BBB='SN12'), loaders.FILTERSET_EXAMPLE49, [0.129, 0.9],
             nR=2, seed=47)


K2_SPECS = {
    "default18": (k2_case, (loaders.FILTERSET_DEFAULT, [0.0, 0.283, 1.58, 5.999])),
    "example49": (k2_case, (loaders.FILTERSET_EXAMPLE49, [0.031024, 0.129, 2.2])),
    "families18": (k2_families_case, (loaders.FILTERSET_DEFAULT, [0.05, 1.0, 3.3])),
}


def replay_k2(name):
    fn, args = K2_SPECS[name]
    return fn(name, *args, write=False)


def sed_case_specs():
    """{case name: (args, kwargs) of sed_case} read off sed_main()."""
    global sed_case
    rec, real = _Recorder(), sed_case
    sed_case = rec
    try:
        sed_main()
    finally:
        sed_case = real
    return rec.k1


def replay_sed(name):
    a, k = sed_case_specs()[name]
    return sed_case(name, *a, **dict(k, write=False))


class _Recorder:
    """Collects the arguments main() / r2_main() pass to run_case, so that tests can replay any case."""

    def __init__(self):
        self.k1 = {}

    def __call__(self, name, *a, **k):
        self.k1[name] = (a, k)


def k1_case_specs():
    """{case name: (args, kwargs) of run_case} for every K1 golden file, read off main() and r2_main()."""
    global run_case
    rec, real = _Recorder(), run_case
    k2a, k2b = globals()["k2_case"], globals()["k2_families_case"]
    run_case = rec
    globals()["k2_case"] = globals()["k2_families_case"] = lambda *a, **k: None
    try:
        argv, sys.argv = sys.argv, [sys.argv[0]]
        main()
        r2_main()
    finally:
        sys.argv = argv
        run_case = real
        globals()["k2_case"], globals()["k2_families_case"] = k2a, k2b
    return rec.k1


def replay_k1(name, eval_rows=100):
    """Regenerate golden case `name` in memory with the reference evaluated on a sample of its positions."""
    a, k = k1_case_specs()[name]
    return run_case(name, *a, **dict(k, write=False, eval_rows=eval_rows))


def r2_main(only=None):
    """Cases added in round 2 (python -m oracle.make_golden r2 [name])."""
    os.makedirs(OUT, exist_ok=True)
    base = dict(H.DEFAULT_SETTINGS)
    gk = dict(age_slice=[0, 6, 12, 18, 24], tau_slice=[0, 5, 10])
    sk = dict(tdust_slice=range(0, 150, 15), fracpah_slice=range(0, 64, 8))
    want = lambda nm: only is None or only == nm
    # the example settings' own model set (example/SETTINGS_AGNfitter.py:193-236): BC03_metal-shaped galaxy keys with
    # free E(B-V) (the 275-SED library stands in for the absent blob), S17, THB21 through the reference's own loader
    # (raw-unpickled frame, ref_shim), SKIRTOR_1P, Flexible energy balance, AGN-fraction prior; both catalogue sources
    if want("example_thb21_skirtor1p_49"):
        cat = loaders.read_ascii_catalog(os.path.join(ref_shim.REF_ROOT, "data", "catalog_example.txt"))
        run_case("example_thb21_skirtor1p_49",
                 dict(base, GALAXY='BC03_metal', STARBURST='S17', BBB='THB21', TORUS='SKIRTOR_1P'),
                 loaders.FILTERSET_EXAMPLE49, None, catalog=cat, galaxy_ebv="free", galaxy_kw=dict(gk, metal=0.02),
                 starburst_kw=sk, n_pos=(300, 40, 30, 80), seed=51, nRADdata=0, nXRaysdata=0)
    # Restrictive energy balance on grid-E(B-V) galaxies (PRIORS_AGNfitter.py:101-104 through GALAXYatt_dict, :80-81)
    if want("restrictive18"):
        run_case("restrictive18", dict(base, PRIOR_energy_balance='Restrictive'), loaders.FILTERSET_DEFAULT,
                 [0.283, 2.4], galaxy_kw=gk, n_pos=(400, 50, 30, 80), seed=52)


    # prior_midIR_UV (PRIORS_AGNfitter.py:327-357): R06 branch (BBBFdict['0.0']) on the default bands ...
    if want("midir_uv18"):
        run_case("midir_uv18", dict(base, PRIOR_midIR_UV=True), loaders.FILTERSET_DEFAULT, [0.283, 1.1],
                 galaxy_kw=gk, n_pos=(400, 50, 30, 80), seed=53)
    # ... and the branch that rewrites the E(B-V) key of a grid disc (SN12, no BB norm: BB == 0 by value), 49 bands
    if want("midir_uv_sn12_49"):
        run_case("midir_uv_sn12_49", dict(base, BBB='SN12', TORUS='SKIRTOR_1P', PRIOR_midIR_UV=True),
                 loaders.FILTERSET_EXAMPLE49, [0.129], galaxy_kw=gk, n_pos=(300, 40, 30, 80), seed=54)
    # prior_IR_XRays (PRIORS_AGNfitter.py:302-324, XRAYS='Prior_midIR'): THB21 disc, NK0 torus, X-ray bands in chi^2
    if want("xray_midir49"):
        run_case("xray_midir49", dict(base, BBB='THB21', TORUS='NK0', XRAYS='Prior_midIR', PRIOR_midIR_UV=True),
                 loaders.FILTERSET_EXAMPLE49, [0.129, 0.8], galaxy_kw=gk, n_pos=(300, 40, 30, 80), seed=55,
                 force_xray_detection=True)


def chain_props_case(write=True):
    """The reference's own chain reduction as the checker of N2: a synthetic samples_mcmc.sav (emcee leg format,
    MCMC_AGNfitter.py:192-206) goes through the UNMODIFIED CHAIN.props (PLOTandWRITE_AGNfitter.py:438-492: thinning
    [:, int(n/2):n:Nthin], flat chain, best-fit row) and the percentile statement of write_parameters_outputvalues
    (:162: np.percentile(..., [2.5, 16, 50, 84, 97.5], axis=0)).  The chain is a random walk with repeated samples
    (rejected proposals repeat a walker's position), ties and a -inf log-posterior entry."""
    import pickle, tempfile
    PLT = ref_shim.install_plot()
    rng = np.random.default_rng(77)
    W, n, n_par, Nthin = 12, 61, 10, 3          # n_par of the classic model set (the GPU test's ensemble)
    step = rng.normal(size=(W, n, n_par)) * (rng.uniform(size=(W, n, 1)) < 0.35)      # ~65 % repeated positions
    scale = np.array([1.0, 1e-3, 50.0, 1e-9, 7.0, 1.0, 0.1, 3.0, 1e5, 1e-300])
    shift = np.array([0.0, 5.0, -300.0, 1e-6, 2.0, -1.0, 0.5, 10.0, -1e7, 0.0])
    chain = np.cumsum(step, axis=1) * scale + shift
    chain[:, :, 4] = np.round(chain[:, :, 4])                                           # many exact ties
    lnprob = -0.5 * ((chain - chain.mean(axis=(0, 1))) ** 2).sum(axis=2)
    lnprob[3, 40] = -np.inf
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "samples_mcmc.sav")
        with open(fn, "wb") as f:
            pickle.dump(dict(chain=chain, accept=np.full(W, 0.3), lnprob=lnprob, final_pos=chain[:, -1], state=None,
                             acor=np.ones(n_par)), f, protocol=2)
        ch = PLT.CHAIN(fn, None, dict(Nsample=1000, Nthinning=Nthin), dict(sampling_algorithm='emcee'))
        ch.props()
    pct = np.percentile(ch.flatchain_sorted, [2.5, 16, 50, 84, 97.5], axis=0)          # PLT:162 (chain_pars columns)
    lnp_pct = np.percentile(ch.lnprob_flat, [2.5, 16, 50, 84, 97.5], axis=0)
    out = dict(chain=chain, lnprob=lnprob, Nthin=np.int64(Nthin), ref_flatchain=np.asarray(ch.flatchain),
               ref_lnprob_flat=np.asarray(ch.lnprob_flat), ref_best_fit_pars=np.asarray(ch.best_fit_pars),
               ref_lnprob_max=np.asarray(ch.lnprob_max), ref_percentiles=pct, ref_lnp_percentiles=lnp_pct)
    if write:
        np.savez_compressed(os.path.join(OUT, "chain_props.npz"), **out)
        print("wrote chain_props.npz", pct.shape)
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "chain":
        chain_props_case()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sed":
        sed_main()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "r2":
        r2_main(sys.argv[2] if len(sys.argv) > 2 else None)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "families":
        os.makedirs(OUT, exist_ok=True)
        k2_families_case("families18", loaders.FILTERSET_DEFAULT, [0.05, 1.0, 3.3])
        return
    os.makedirs(OUT, exist_ok=True)
    base = dict(H.DEFAULT_SETTINGS)
    # A: the classic 10-parameter model set, default 18 bands, the two z of BASELINE configs[0]
    run_case("classic18", base, loaders.FILTERSET_DEFAULT, [0.283, 1.58], n_pos=(900, 100, 60, 240), seed=11)
    # B: example / brown2018 catalogues (49 bands), S17 starburst subset, galaxy subset
    gk = dict(age_slice=[0, 6, 12, 18, 24], tau_slice=[0, 5, 10])
    sk = dict(tdust_slice=range(0, 150, 15), fracpah_slice=range(0, 64, 8))
    for cname in ("catalog_example", "brown2018_test"):
        cat = loaders.read_ascii_catalog(os.path.join(ref_shim.REF_ROOT, "data", cname + ".txt"))
        run_case("cat49_" + cname, dict(base, STARBURST='S17', TORUS='NK0'), loaders.FILTERSET_EXAMPLE49, None,
                 catalog=cat, galaxy_kw=gk, starburst_kw=sk, n_pos=(300, 40, 30, 80), seed=12,
                 nRADdata=0, nXRaysdata=0)
    # C: free E(B-V)_gal + Restrictive energy balance + X-ray extended R06 (EBVbbb, alphaScat free)
    run_case("freeebv_xray49", dict(base, STARBURST='S17', XRAYS=True, PRIOR_energy_balance='Restrictive'),
             loaders.FILTERSET_EXAMPLE49, [0.129, 2.2], galaxy_ebv="free", galaxy_kw=gk, starburst_kw=sk,
             n_pos=(400, 50, 30, 80), seed=13, nXRaysdata=1)
    run_case("xray_gamma49", dict(base, XRAYS=True, PRIOR_energy_balance='None', turn_on_AGN=True),
             loaders.FILTERSET_EXAMPLE49, [0.031024], galaxy_kw=gk, n_pos=(300, 40, 30, 80), seed=14, nXRaysdata=3)
    # D: five components with radio (SN12 disk: no BB norm, /4 pi dlum^2 path) + IR/synchrotron prior
    for nrad in (1, 2, 4):
        run_case("radio%d_sn12_49" % nrad, dict(base, RADIO=True, BBB='SN12'), loaders.FILTERSET_EXAMPLE49,
                 [0.129], galaxy_kw=gk, n_pos=(300, 40, 30, 80), seed=15 + nrad, nRADdata=nrad)
    # C4 of BASELINE.json: radio-to-X-ray full model — five components, X-ray extended R06 (EBVbbb and alphaScat
    # free, Gamma grid), 3 000-key synchrotron family, 18 parameters
    run_case("radio_xray49", dict(base, RADIO=True, XRAYS=True, STARBURST='S17', TORUS='SKIRTOR_2P'),
             loaders.FILTERSET_EXAMPLE49, [0.129, 1.234], galaxy_kw=gk, starburst_kw=sk,
             n_pos=(300, 40, 30, 80), seed=31, nRADdata=4, nXRaysdata=3)
    # AGN switched off (norm caps at -9.8), no AGN-fraction prior
    run_case("agnoff18", dict(base, turn_on_AGN=False, PRIOR_AGNfraction=False), loaders.FILTERSET_DEFAULT,
             [3.005], galaxy_kw=gk, n_pos=(300, 40, 30, 80), seed=21)
    # K2
    k2_case("default18", loaders.FILTERSET_DEFAULT, [0.0, 0.283, 1.58, 5.999])
    k2_case("example49", loaders.FILTERSET_EXAMPLE49, [0.031024, 0.129, 2.2])
    k2_families_case("families18", loaders.FILTERSET_DEFAULT, [0.05, 1.0, 3.3])


if __name__ == "__main__":
    main()
