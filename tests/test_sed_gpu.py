"""GPU: K4 (SED realisations + integrated luminosities) through the C ABI against the golden outputs
of the reference's FLUXES_ARRAYS.fluxes and against the CPU oracle.  fp64 within 1e-10 relative; exact zeros, NaN
and the -100 fill in the same places."""
import numpy as np
import pytest

from conftest import SED_CASES
from helpers import load_sed, sed_component_order, assert_close_with_specials
from agnfitter_b200 import sources
from agnfitter_b200.engine import Engine
from agnfitter_b200.realizations import SedEngine
from oracle import sed_oracle as SO

pytestmark = pytest.mark.gpu


def _engine(ms, g):
    eng = Engine(ms, device=0)
    src = sources.make_sources(ms, g["src_z"], g["src_dlum"], g["src_nus"], g["src_fluxes"], g["src_fluxerrs"])
    eng.set_sources(src)
    return eng


@pytest.mark.parametrize("case", SED_CASES)
def test_k4_matches_reference_golden(case):
    ms, ps, g = load_sed(case)
    eng = _engine(ms, g)
    sed = SedEngine(eng, ps, g["src_z"], g["src_dlum"])
    assert np.array_equal(sed.grid, g["all_nus_rest"])
    assert sed.component_names == sed_component_order(ms)
    stride = int(g["stride"])
    spec = sed.spectra(g["pars"], g["source_id"])
    want = g["ref_nuLnu"]
    got = spec[:, :, ::stride]
    assert np.array_equal(got == 0.0, want == 0.0), case
    assert_close_with_specials(got, want, what=case + " nuLnu")
    sed.set_bands([str(m) for m in g["intlum_models"]], g["intlum_ranges_hz"])
    with np.errstate(all="ignore"):
        il = np.log10(sed.int_lums(g["pars"], g["source_id"]))
    assert_close_with_specials(il, g["ref_int_lums"], what=case + " int_lums")


@pytest.mark.parametrize("case", ["classic18", "freeebv_xray49", "radio2_sn12_49"])
def test_k4_full_grid_matches_oracle(case):
    """Every grid point (the golden files keep every 7th) and rows the fixtures do not hold: box-uniform rows,
    rows on grid values and rows outside the box (snapped to the edge key, like the reference)."""
    ms, ps, g = load_sed(case)
    eng = _engine(ms, g)
    sed = SedEngine(eng, ps, g["src_z"], g["src_dlum"])
    rng = np.random.default_rng(3)
    order = sed_component_order(ms)
    for s in range(len(g["src_z"])):
        lo, hi = g["ref_pmin"][s], g["ref_pmax"][s]
        pars = rng.uniform(lo - 0.05 * (hi - lo), hi + 0.05 * (hi - lo), size=(24, ms.n_par))
        pars[:4] = rng.uniform(lo, hi, size=(4, ms.n_par))
        sid = np.full(len(pars), s, dtype=np.int32)
        z, dlum = float(g["src_z"][s]), float(g["src_dlum"][s])
        grid, fl = SO.realization_fluxes(ms, ps, pars, z, dlum)
        nl = SO.fluxes_to_nuLnu(grid, fl, z, dlum)
        spec = sed.spectra(pars, sid)
        for ci, c in enumerate(order):
            assert np.array_equal(spec[:, ci] == 0.0, nl[c] == 0.0), (case, c)
            assert_close_with_specials(spec[:, ci], nl[c], what="%s %s" % (case, c))
        models = [str(m) for m in g["intlum_models"]]
        sed.set_bands(models, g["intlum_ranges_hz"])
        want = SO.integrated_luminosities(models, g["intlum_ranges_hz"], grid, nl, ms.radio).T
        assert_close_with_specials(sed.int_lums(pars, sid), want, what=case + " int_lums")


def test_k4_batch_properties():
    """Size-independent properties on a large batch: rows are independent (a row gives the same bits alone, in a
    batch, and in a permuted batch), a repeat launch is bit-identical, and the total is the sum of its parts."""
    ms, ps, g = load_sed("classic18")
    eng = _engine(ms, g)
    sed = SedEngine(eng, ps, g["src_z"], g["src_dlum"])
    sed.set_bands([str(m) for m in g["intlum_models"]], g["intlum_ranges_hz"])
    rng = np.random.default_rng(5)
    n = 3000
    sid = rng.integers(0, len(g["src_z"]), size=n).astype(np.int32)
    pars = rng.uniform(g["ref_pmin"][sid], g["ref_pmax"][sid])
    a = sed.int_lums(pars, sid)
    assert np.array_equal(a, sed.int_lums(pars, sid), equal_nan=True)
    perm = rng.permutation(n)
    assert np.array_equal(a[perm], sed.int_lums(pars[perm], sid[perm]), equal_nan=True)
    for i in (0, 17, n - 1):
        assert np.array_equal(a[i], sed.int_lums(pars[i:i + 1], sid[i:i + 1])[0], equal_nan=True)
    spec = sed.spectra(pars[:64], sid[:64])
    names = sed.component_names
    F = {c: spec[:, names.index(c)] for c in names}
    fin = np.isfinite(F['total'])
    tot = F['sb'] + F['bb'] + F['ga'] + F['to']
    assert np.allclose(tot[fin], F['total'][fin], rtol=1e-13, atol=0)


def test_k4_api_errors():
    ms, ps, g = load_sed("classic18")
    eng = _engine(ms, g)
    sed = SedEngine(eng, ps, g["src_z"], g["src_dlum"])
    with pytest.raises(RuntimeError):
        sed.int_lums(g["pars"], g["source_id"])                 # no bands yet
    with pytest.raises(ValueError):
        sed.set_bands(['agn_rad'], [[1e9, 1e10]])               # no radio component in this model set
    with pytest.raises(ValueError):
        sed.set_bands(['AGNfracTO'], [[1e13, 1e13]])            # the reference appends nothing for this
    with pytest.raises(RuntimeError):
        sed.spectra(g["pars"], np.full(len(g["pars"]), 99, dtype=np.int32))


@pytest.mark.parametrize("case,output_type", [("classic18", "int_lums"), ("radio2_sn12_49", "plot")])
def test_fluxes_arrays_mirror(case, output_type):
    """FluxesArrays = FLUXES_ARRAYS with the reference's constructor / entry point / result attributes
    (PLOTandWRITE_AGNfitter.py:541-799), fed the reference's kinds of objects (MODELSDICT with `_4plot` dicts)."""
    import types
    from helpers import reference_objects
    from agnfitter_b200.realizations import FluxesArrays, ranges_to_hz
    from oracle import agn_oracle as O
    ms, ps, g = load_sed(case)
    s = 0
    data, models, P = reference_objects(ms, g, s, ps)
    rng = np.random.default_rng(9)
    flat = rng.uniform(g["ref_pmin"][s], g["ref_pmax"][s], size=(300, ms.n_par))
    chain = types.SimpleNamespace(flatchain=flat, best_fit_pars=flat[5], props=lambda: None)
    out = {'realizations2int': 12, 'realizations2plot': 7, 'intlum_models': [str(m) for m in g["intlum_models"]],
           'intlum_freqranges': g["intlum_ranges_um"]}
    fa = FluxesArrays(chain, P, out, output_type, ms.settings, {'sampling_algorithm': 'emcee'},
                      rng=np.random.RandomState(4))
    fa.fluxes(data, models)
    n_real = 12 if output_type == 'int_lums' else 7
    pick = np.random.RandomState(4).choice(len(flat), n_real)
    par = flat[pick]
    par[-1] = flat[5]
    z, dlum = float(data.z), float(data.dlum)
    grid, fl = SO.realization_fluxes(ms, ps, par, z, dlum)
    nl = SO.fluxes_to_nuLnu(grid, fl, z, dlum)
    assert np.array_equal(fa.all_nus_rest, grid)
    assert len(fa.nuLnus4plotting) == len(sed_component_order(ms))
    for got, c in zip(fa.nuLnus4plotting, sed_component_order(ms)):
        assert_close_with_specials(got, nl[c], what=c)
    if output_type == 'int_lums':
        with np.errstate(all="ignore"):
            want = np.log10(SO.integrated_luminosities(out['intlum_models'], ranges_to_hz(out['intlum_freqranges']),
                                                       grid, nl, ms.radio))
        assert_close_with_specials(fa.int_lums, want, what="int_lums")
        assert np.array_equal(fa.int_lums_best, fa.int_lums[:, -1], equal_nan=True)
    else:
        src = dict(z=z, dlum=dlum, nus=data.nus, fluxes=data.fluxes, fluxerrs=data.fluxerrs, pmin=g["ref_pmin"][s],
                   pmax=g["ref_pmax"][s], zi=0)
        ym = O.ln_probab(ms, src, par, return_parts=True)[1]
        want = ym * (4. * np.pi * dlum ** 2.) * 10 ** (np.asarray(data.nus))
        assert_close_with_specials(fa.filtered_modelpoints_nuLnu, want, what="filtered model points")
