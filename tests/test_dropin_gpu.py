"""GPU: the reference-facing callables (same names / signatures as PARAMETERSPACE_AGNfitter and the
closures of MCMC_AGNfitter.main) give the reference's numbers."""
import types

import numpy as np
import pytest

from helpers import load_k1, load_k2, assert_close_with_specials
from agnfitter_b200 import dropin
from agnfitter_b200.engine import DictBuilder

pytestmark = pytest.mark.gpu


def _reference_objects(ms, g, s):
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
    models = types.SimpleNamespace(settings=ms.settings, dict_modelfluxes=md)
    data = types.SimpleNamespace(z=float(g["src_z"][s]), dlum=float(g["src_dlum"][s]), nus=g["src_nus"][s],
                                 fluxes=g["src_fluxes"][s], fluxerrs=g["src_fluxerrs"][s])
    P = {'names': list(ms.par_names), 'min': list(g["ref_pmin"][s]), 'max': list(g["ref_pmax"][s])}
    return data, models, P


@pytest.mark.parametrize("case", ["classic18", "cat49_catalog_example", "radio2_sn12_49", "freeebv_xray49"])
def test_posterior_matches_reference(case):
    ms, g = load_k1(case)
    for s in range(len(g["src_z"])):
        data, models, P = _reference_objects(ms, g, s)
        post = dropin.Posterior(data, models, P)
        sel = np.nonzero(g["source_id"] == s)[0]
        pars = g["pars"][sel]
        # emcee vectorize=True / ultranest vectorized=True: whole blocks
        assert_close_with_specials(post.ln_probab_batch(pars), g["ref_lnp"][sel], what="ln_probab_batch")
        assert_close_with_specials(post.my_posterior(pars), g["ref_lnp"][sel], what="my_posterior")
        # the reference's own call shape: a tuple of floats + (data, models, P), returns a float
        for k in range(0, len(sel), max(1, len(sel) // 25)):
            v = post.ln_probab(tuple(pars[k]), data, models, P)
            assert isinstance(v, float)
            assert_close_with_specials(np.array([v]), g["ref_lnp"][sel[k:k + 1]], what="ln_probab")
        lum, bands = post.ymodel(data.nus, data.z, data.dlum, models, P, *pars[0])
        assert np.array_equal(bands, ms.bands)
        assert_close_with_specials(lum, g["ref_ymodel"][sel[0]], what="ymodel")
        cube = np.random.default_rng(s).uniform(size=(64, ms.n_par))
        t = post.my_params_transform(cube)
        assert np.all(t >= np.asarray(P['min'])) and np.all(t <= np.asarray(P['max']))


def test_fill_modelsdict_matches_reference_dictionary():
    g = load_k2("default18")
    bands, off = g["bands"], g["filt_offset"]
    fd = [bands, {b: [g["filt_lambda"][off[i]:off[i + 1]]] for i, b in enumerate(bands)},
          {b: [g["filt_factor"][off[i]:off[i + 1]]] for i, b in enumerate(bands)}]
    builder = DictBuilder(fd, device=0)
    fam = "TO_S04" if "TO_S04" in [str(x) for x in g["family_names"]] else [str(x) for x in g["family_names"]][0]
    ref = g[fam + "_ref"]
    tpl = {}
    for i in range(ref.shape[1]):
        n = int(g[fam + "_len"][i])
        tpl["k%03d" % i] = (g[fam + "_lognu"][i, :n], g[fam + "_Fnu"][i, :n])
    z = float(g["z"][1])
    md = dropin.fill_modelsdict(types.SimpleNamespace(), builder, z, {"TORUS": tpl})
    assert list(md.TORUSFdict.keys()) == list(tpl.keys()) and md.z == z
    for i, k in enumerate(tpl):
        b, F = md.TORUSFdict[k]
        assert np.array_equal(b, bands)
        assert_close_with_specials(F, ref[1, i], what="TORUSFdict[%s]" % k)
    assert md.TORUSFdict_4plot is tpl
