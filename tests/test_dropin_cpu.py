"""CPU: host-side logic of the reference-facing mirror (agnfitter_b200/dropin.py)."""
import numpy as np
import pytest

from helpers import load_k1
from agnfitter_b200 import dropin, sources


def _reference_params_transform(cube, P):
    """MCMC_AGNfitter.py:64-72, restated verbatim (loops and all)."""
    params = cube.copy()
    for i in range(params.shape[0]):
        normalization = []
        for k in range(params.shape[1]):
            normalization.append((P['max'][k] - P['min'][k]) * cube[i][k] + P['min'][k])
        params[i] = normalization
    return params


def test_params_transform_matches_reference_loop():
    ms, g = load_k1("classic18")
    P = {'min': list(g["ref_pmin"][0]), 'max': list(g["ref_pmax"][0])}
    cube = np.random.default_rng(3).uniform(size=(257, ms.n_par))
    assert np.array_equal(dropin.params_transform(cube, P['min'], P['max']), _reference_params_transform(cube, P))


def test_check_P_accepts_reference_table_and_rejects_others():
    ms, g = load_k1("classic18")
    pmin, pmax = sources.pdict_bounds(ms, float(g["src_z"][0]))
    P = {'names': list(ms.par_names), 'min': list(pmin), 'max': list(pmax)}
    dropin.check_P(ms, P, pmin, pmax)
    bad = dict(P, names=list(reversed(P['names'])))
    with pytest.raises(ValueError):
        dropin.check_P(ms, bad, pmin, pmax)
    bad = dict(P, max=[v + 1 for v in P['max']])
    with pytest.raises(ValueError):
        dropin.check_P(ms, bad, pmin, pmax)
