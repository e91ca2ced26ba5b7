"""CPU: the oracle (oracle/agn_oracle.py) and the product's host-side logic against the golden
vectors produced by the reference's own functions (oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import K1_CASES, K2_CASES
from helpers import load_k1, load_k2, assert_close_with_specials, oracle_src, RTOL
from agnfitter_b200 import hostmath, sources
from oracle import agn_oracle as O


@pytest.mark.parametrize("case", K1_CASES)
def test_oracle_lnprob_matches_reference(case):
    ms, g = load_k1(case)
    for s in range(len(g["src_z"])):
        sel = g["source_id"] == s
        lnp, ym, rows, lpr, lll = O.ln_probab(ms, oracle_src(ms, g, s), g["pars"][sel], return_parts=True)
        assert np.array_equal(rows, g["ref_rows"][sel]), "template indices must be bit-exact"
        assert_close_with_specials(ym, g["ref_ymodel"][sel], what="ymodel")
        assert_close_with_specials(lll, g["ref_lnlike"][sel], what="ln_likelihood")
        assert_close_with_specials(lpr, g["ref_lnprior"][sel], what="ln_prior")
        assert_close_with_specials(lnp, g["ref_lnp"][sel], what="ln_probab")


@pytest.mark.parametrize("case", K1_CASES)
def test_host_masks_and_bounds_are_bit_exact(case):
    ms, g = load_k1(case)
    for s in range(len(g["src_z"])):
        z = float(g["src_z"][s])
        m = hostmath.likelihood_mask(g["src_nus"][s], g["src_fluxes"][s], z, ms.settings['XRAYS'])
        assert np.array_equal(m, g["ref_mask"][s])
        v = O.valid_bands(ms, g["src_nus"][s], g["src_fluxes"][s], z)
        assert np.array_equal(np.nonzero(g["ref_mask"][s])[0], v)
        pmin, pmax = sources.pdict_bounds(ms, z)
        assert np.array_equal(pmin, g["ref_pmin"][s]) and np.array_equal(pmax, g["ref_pmax"][s])
    assert ms.n_par == g["pars"].shape[1]


@pytest.mark.parametrize("case", K2_CASES)
def test_oracle_filtering_models_matches_reference(case):
    g = load_k2(case)
    bands, off = g["bands"], g["filt_offset"]
    lam = [g["filt_lambda"][off[b]:off[b + 1]] for b in range(len(bands))]
    fac = [g["filt_factor"][off[b]:off[b + 1]] for b in range(len(bands))]
    for fam in [str(x) for x in g["family_names"]]:
        ref = g[fam + "_ref"]
        for i in range(ref.shape[1]):
            n = int(g[fam + "_len"][i])
            for zi, z in enumerate(g["z"]):
                got = O.filtering_models(g[fam + "_lognu"][i, :n], g[fam + "_Fnu"][i, :n], bands, lam, fac, float(z))
                assert_close_with_specials(got, ref[zi, i], rtol=1e-13, what="%s[%d] z=%g" % (fam, i, z))
