"""GPU: the reference-facing callables (same names / signatures as PARAMETERSPACE_AGNfitter and the
closures of MCMC_AGNfitter.main) give the reference's numbers."""
import types

import numpy as np
import pytest

from helpers import load_k1, load_k2, assert_close_with_specials, reference_objects as _reference_objects
from agnfitter_b200 import dropin
from agnfitter_b200.engine import DictBuilder

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("case", ["classic18", "radio2_sn12_49"])
def test_posterior_from_the_dense_dictionary_cache(case, tmp_path):
    """with `<zdict.filename>.b200.npz` beside the pickle (dropin.save_dictionary_cache, written by the
    patched RUN_AGNfitter_multi.py) the Posterior loads the dense arrays instead of flattening the Python dicts, and
    evaluates the same bits."""
    ms, g = load_k1(case)
    data, models, P = _reference_objects(ms, g, 0)
    md = models.dict_modelfluxes
    md.modelsettings = models.settings
    sel = np.nonzero(g["source_id"] == 0)[0]
    plain = dropin.Posterior(data, models, P)
    assert not plain.from_cache
    want = plain.ln_probab_batch(g["pars"][sel])
    md.filename = str(tmp_path / "MODELSDICT_z")
    dropin.save_dictionary_cache(md)
    cached = dropin.Posterior(data, models, P)
    assert cached.from_cache
    assert np.array_equal(cached.ln_probab_batch(g["pars"][sel]), want, equal_nan=True)
    assert_close_with_specials(want, g["ref_lnp"][sel], what="ln_probab_batch")


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


def test_main_writes_the_reference_chain_files(tmp_path):
    """dropin.main = MCMC_AGNfitter.main's emcee leg: same signature, same output files."""
    import pickle
    ms, g = load_k1("classic18")
    data, models, P = _reference_objects(ms, g, 0)
    data.output_folder, data.name = str(tmp_path) + "/", "obj7"
    mc = dict(Nwalkers=20, Nburnsets=2, Nburn=12, Nmcmc=24, Nthin=2, iprint=1000, sampling_algorithm='emcee')
    out = dropin.main(data, models, P, mc, seed=3)
    again = dropin.main(data, models, P, mc, seed=3)
    assert np.array_equal(out["best_pos"], again["best_pos"]) and out["best_lnp"][0] == again["best_lnp"][0]
    with open(tmp_path / "obj7" / "samples_mcmc.sav", "rb") as f:
        smp = pickle.load(f)
    assert smp["chain"].shape == (20, 24, ms.n_par) and smp["lnprob"].shape == (20, 24)
    # every recorded sample carries its own log-posterior (checked against the batch entry point)
    post = dropin.Posterior(data, models, P)
    flat = smp["chain"].reshape(-1, ms.n_par)
    assert_close_with_specials(post.ln_probab_batch(flat), smp["lnprob"].ravel(), what="chain lnprob")
    assert 0.0 < smp["accept"].mean() < 1.0
    with pytest.raises(ValueError):
        dropin.main(data, models, P, dict(mc, sampling_algorithm='metropolis'))


def test_main_ultranest_leg_wiring(tmp_path, monkeypatch):
    """sampling_algorithm='ultranest' (the example settings' choice, SETTINGS_AGNfitter.py:262): dropin.main hands
    ultranest the GPU-backed callables the way MCMC_AGNfitter.main hands it its closures (:79-96).  ultranest is not
    installed here, so a stand-in module records the calls and drives the two callables like the vectorised sampler
    does."""
    import sys
    ms, g = load_k1("classic18")
    data, models, P = _reference_objects(ms, g, 0)
    data.output_folder, data.name = str(tmp_path) + "/", "obj9"
    seen = {}

    class FakeSampler:
        def __init__(self, names, loglike, transform, resume, log_dir, vectorized):
            seen.update(names=names, log_dir=log_dir, vectorized=vectorized, resume=resume)
            self.loglike, self.transform = loglike, transform

        def run(self, **kw):
            seen["run"] = kw
            cube = np.random.default_rng(0).uniform(size=(300, len(seen["names"])))
            p = self.transform(cube)
            seen["params"], seen["lnp"] = p, self.loglike(p)

        def plot_run(self): seen["plots"] = seen.get("plots", 0) + 1
        def plot_trace(self): seen["plots"] = seen.get("plots", 0) + 1
        def plot_corner(self): seen["plots"] = seen.get("plots", 0) + 1

    step = types.ModuleType("ultranest.stepsampler")
    step.generate_mixture_random_direction = "de-mix"
    step.generate_region_oriented_direction = "region"
    step.generate_random_direction = "random"
    step.OrthogonalDirectionGenerator = lambda f: ("ortho", f)
    step.SliceSampler = lambda nsteps, generate_direction: ("slice", nsteps, generate_direction)
    un = types.ModuleType("ultranest")
    un.ReactiveNestedSampler, un.stepsampler = FakeSampler, step
    monkeypatch.setitem(sys.modules, "ultranest", un)
    monkeypatch.setitem(sys.modules, "ultranest.stepsampler", step)
    mc = dict(sampling_algorithm='ultranest', direction_generation='de-mix', live_points=400, min_ess=400, num_loops=10)
    sampler = dropin.main(data, models, P, mc)
    assert seen["names"] == P['names'] and seen["vectorized"] is True and seen["resume"] is True
    assert seen["log_dir"] == data.output_folder + "obj9/ultranest" and (tmp_path / "obj9").is_dir()
    assert sampler.stepsampler == ("slice", 20, "de-mix") and seen["plots"] == 3
    assert seen["run"] == dict(min_num_live_points=400, min_ess=400, max_num_improvement_loops=10)
    want = dropin.Posterior(data, models, P).ln_probab_batch(seen["params"])
    assert np.array_equal(seen["lnp"], want, equal_nan=True)
