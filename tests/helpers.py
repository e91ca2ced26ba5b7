"""Shared helpers for the parity tests (tests/ only)."""
import os
import types

import numpy as np

from agnfitter_b200 import flatten

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: ln-probability and band fluxes within 1e-10 relative in fp64
RTOL = 1e-10


def load_k1(name):
    with np.load(os.path.join(GOLDEN, "k1_%s.npz" % name), allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    ms = flatten.modelset_from_arrays(g, prefix="ms_")
    return ms, g


def load_sed(name):
    """(ModelSet, PlotSet, arrays) of a SED-realisation golden case (tests/golden/sed_*.npz)."""
    from agnfitter_b200 import plotset
    with np.load(os.path.join(GOLDEN, "sed_%s.npz" % name), allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    return flatten.modelset_from_arrays(g, prefix="ms_"), plotset.plotset_from_arrays(g), g


def sed_component_order(ms):
    """Order of the arrays in FLUXES4plotting / nuLnus4plotting (PLOTandWRITE_AGNfitter.py:778-782)."""
    return ['sb', 'bb', 'ga', 'to'] + (['rad'] if ms.radio else []) + ['total', 'bbdered']


def load_k2(name):
    with np.load(os.path.join(GOLDEN, "k2_%s.npz" % name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def assert_close_with_specials(got, want, rtol=RTOL, what=""):
    """Equal up to rtol where finite; identical placement of -inf/+inf/NaN."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    nan_w, nan_g = np.isnan(want), np.isnan(got)
    assert np.array_equal(nan_w, nan_g), "%s: NaN placement differs at %s" % (what, np.nonzero(nan_w != nan_g)[0][:10])
    inf_w = np.isinf(want)
    assert np.array_equal(inf_w, np.isinf(got)), "%s: inf placement differs at %s" % (
        what, np.nonzero(inf_w != np.isinf(got))[0][:10])
    assert np.array_equal(np.sign(want[inf_w]), np.sign(got[inf_w])), "%s: inf sign differs" % what
    fin = ~(nan_w | inf_w)
    denom = np.maximum(np.abs(want[fin]), np.finfo(np.float64).tiny)
    rel = np.abs(got[fin] - want[fin]) / denom
    rel[want[fin] == got[fin]] = 0.0
    worst = rel.max() if rel.size else 0.0
    assert worst <= rtol, "%s: max relative error %.3e > %.1e (at %s)" % (what, worst, rtol, np.argmax(rel))
    return worst


def oracle_src(ms, g, s):
    zpos = {float(v): i for i, v in enumerate(ms.z_grid)}
    return dict(z=float(g["src_z"][s]), dlum=float(g["src_dlum"][s]), nus=g["src_nus"][s],
                fluxes=g["src_fluxes"][s], fluxerrs=g["src_fluxerrs"][s],
                pmin=g["ref_pmin"][s], pmax=g["ref_pmax"][s], zi=zpos[float(g["src_z"][s])])


def reference_objects(ms, g, s, ps=None):
    """Stand-ins with exactly the attributes the reference's objects expose on this path:
    MODELSDICT (DICTIONARIES_AGNfitter.py:86-136), MODELS (.settings, .dict_modelfluxes), DATA."""
    zi = int(np.nonzero(ms.z_grid == float(g["src_z"][s]))[0][0])
    md = types.SimpleNamespace(z=float(ms.z_grid[zi]))

    def fdict(f):
        if not f.fitted:
            return {'-99.9': (ms.bands, f.table[zi, 0])}
        out = {}
        for t in range(f.T):
            key = tuple(repr(float(v)) for v in f.keys_float[:, t])
            out[key if f.keys_float.shape[0] > 1 else key[0]] = (ms.bands, f.table[zi, t])
        return out

    fams = {f.name: f for f in ms.families}
    md.GALAXYFdict, md.STARBURSTFdict = fdict(fams["GA"]), fdict(fams["SB"])
    md.TORUSFdict, md.BBBFdict = fdict(fams["TO"]), fdict(fams["BB"])
    names = [fams[k].par_names for k in ("GA", "SB", "TO", "BB")]
    types_ = [fams[k].par_types for k in ("GA", "SB", "TO", "BB")]
    if "RAD" in fams:
        md.AGN_RADFdict = fdict(fams["RAD"])
        names.append(fams["RAD"].par_names)
        types_.append(fams["RAD"].par_types)
    md.all_parnames = names + [ms.norm_names]
    md.all_partypes = types_ + [['free'] * len(ms.norm_names)]
    if fams["GA"].aux is not None:
        md.GALAXYatt_dict = dict(zip(md.GALAXYFdict.keys(), fams["GA"].aux))
    md.STARBURST_LIRdict = dict(zip(md.STARBURSTFdict.keys(), fams["SB"].aux)) if fams["SB"].aux is not None else {}
    if ps is not None:         # the full-resolution twins (DICTIONARIES_AGNfitter.py:92-131), same key order
        for short, attr in (("GA", "GALAXY"), ("SB", "STARBURST"), ("TO", "TORUS"), ("BB", "BBB"), ("RAD", "AGN_RAD")):
            if short in fams:
                pf = ps.fam(short)
                keys = list(getattr(md, attr + "Fdict").keys())
                setattr(md, attr + "Fdict_4plot", {k: pf.template(t) for t, k in enumerate(keys)})
    models = types.SimpleNamespace(settings=ms.settings, dict_modelfluxes=md)
    data = types.SimpleNamespace(z=float(g["src_z"][s]), dlum=float(g["src_dlum"][s]), nus=g["src_nus"][s],
                                 fluxes=g["src_fluxes"][s], fluxerrs=g["src_fluxerrs"][s])
    P = {'names': list(ms.par_names), 'min': list(g["ref_pmin"][s]), 'max': list(g["ref_pmax"][s])}
    return data, models, P
