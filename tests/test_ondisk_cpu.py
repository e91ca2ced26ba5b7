"""CPU: the dense on-disk dictionary format — `.npz` files of plain arrays instead of the pickle of
Python dicts (RUN_AGNfitter_multi.py:203-207).  A saved and re-loaded model set is the same model set: identical
arrays, identical snapping structure, identical oracle results; no pickled objects inside."""
import numpy as np

from helpers import load_k1, load_sed
from agnfitter_b200 import flatten, plotset
from oracle import agn_oracle as O, sed_oracle as SO


def test_modelset_npz_roundtrip(tmp_path):
    ms, g = load_k1("radio2_sn12_49")
    path = str(tmp_path / "dict.npz")
    flatten.save_modelset(path, ms)
    with np.load(path, allow_pickle=False) as z:              # plain arrays only
        assert all(z[k].dtype != object for k in z.files)
    back = flatten.load_modelset(path)
    assert back.settings == ms.settings and back.par_names == ms.par_names
    assert np.array_equal(back.bands, ms.bands) and np.array_equal(back.z_grid, ms.z_grid)
    for a, b in zip(ms.families, back.families):
        assert a.name == b.name and a.par_names == b.par_names and a.par_types == b.par_types and a.fitted == b.fitted
        assert np.array_equal(a.keys_float, b.keys_float) and np.array_equal(a.stride, b.stride)
        assert all(np.array_equal(x, y) for x, y in zip(a.grid, b.grid))
        assert np.array_equal(a.table, b.table)
        assert (a.aux is None) == (b.aux is None) and (a.aux is None or np.array_equal(a.aux, b.aux))
    s = 0
    sel = np.nonzero(g["source_id"] == s)[0][:200]
    src = dict(z=float(g["src_z"][s]), dlum=float(g["src_dlum"][s]), nus=g["src_nus"][s], fluxes=g["src_fluxes"][s],
               fluxerrs=g["src_fluxerrs"][s], pmin=g["ref_pmin"][s], pmax=g["ref_pmax"][s], zi=0)
    assert np.array_equal(O.ln_probab(back, src, g["pars"][sel]), O.ln_probab(ms, src, g["pars"][sel]), equal_nan=True)


def test_plotset_npz_roundtrip(tmp_path):
    ms, ps, g = load_sed("classic18")
    path = str(tmp_path / "plot.npz")
    plotset.save_plotset(path, ps)
    with np.load(path, allow_pickle=False) as z:
        assert all(z[k].dtype != object for k in z.files)
    back = plotset.load_plotset(path)
    assert back.dered_mode == ps.dered_mode and np.array_equal(back.dered_row, ps.dered_row)
    z0, d0 = float(g["src_z"][0]), float(g["src_dlum"][0])
    rows = g["pars"][g["source_id"] == 0][:3]
    _, a = SO.realization_fluxes(ms, ps, rows, z0, d0)
    _, b = SO.realization_fluxes(ms, back, rows, z0, d0)
    assert all(np.array_equal(a[c], b[c], equal_nan=True) for c in a)


def test_dictionary_cache_beside_the_pickle(tmp_path):
    """dropin.save_dictionary_cache / load_dictionary_cache: the dense `.npz` written next to the reference's dictionary
    pickle (RUN_AGNfitter_multi.py:203-213) gives back the flattened dictionary bit for bit, and is ignored when it
    describes another dictionary (other redshift, other template count, other settings) or is unreadable."""
    from helpers import reference_objects
    from agnfitter_b200 import dropin
    ms, g = load_k1("classic18")
    data, models, P = reference_objects(ms, g, 0)
    md = models.dict_modelfluxes
    md.modelsettings = models.settings
    assert dropin.cache_path(md) is None and dropin.load_dictionary_cache(md, models.settings) is None
    md.filename = str(tmp_path / "dict_z.pickle").encode()         # RUN_AGNfitter_multi.py:191 passes bytes around
    assert dropin.load_dictionary_cache(md, models.settings) is None
    path = dropin.save_dictionary_cache(md)
    assert path == str(tmp_path / "dict_z.pickle") + ".b200.npz"
    back = dropin.load_dictionary_cache(md, models.settings)
    flat = flatten.modelset_from_modelsdicts([md], models.settings)
    assert back is not None and back.par_names == flat.par_names
    for a, b in zip(flat.families, back.families):
        assert np.array_equal(a.table, b.table) and np.array_equal(a.keys_float, b.keys_float)
        assert np.array_equal(a.stride, b.stride) and (a.aux is None or np.array_equal(a.aux, b.aux))
    other = dict(models.settings)
    other['PRIOR_AGNfraction'] = not (other.get('PRIOR_AGNfraction') == True)
    assert dropin.load_dictionary_cache(md, other) is None          # other prior settings
    z_true = md.z
    md.z = z_true + 0.5
    assert dropin.load_dictionary_cache(md, models.settings) is None  # other redshift
    md.z = z_true
    k = next(iter(md.TORUSFdict))
    dropped = md.TORUSFdict.pop(k)
    assert dropin.load_dictionary_cache(md, models.settings) is None  # other template count
    md.TORUSFdict[k] = dropped
    open(path, "wb").write(b"not an npz")
    assert dropin.load_dictionary_cache(md, models.settings) is None  # unreadable: fall back to the dicts
