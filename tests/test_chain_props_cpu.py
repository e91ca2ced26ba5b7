"""CPU: the chain reduction the GPU implements restated in NumPy against the golden file produced by
the reference's own CHAIN.props + np.percentile (oracle/make_golden.py chain); with /root/reference mounted the golden
is regenerated and compared bit for bit."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_props.npz")


def thinned(chain, lnprob, nthin):
    """CHAIN.props, emcee leg (PLOTandWRITE_AGNfitter.py:471-487)."""
    n = chain.shape[1]
    return (chain[:, int(n / 2):n:nthin, :].reshape(-1, chain.shape[2]), lnprob[:, int(n / 2):n:nthin].ravel())


def test_restated_reduction_equals_the_references():
    g = np.load(GOLDEN)
    flat, lnflat = thinned(g["chain"], g["lnprob"], int(g["Nthin"]))
    assert np.array_equal(flat, g["ref_flatchain"]) and np.array_equal(lnflat, g["ref_lnprob_flat"])
    assert np.array_equal(flat[(-lnflat).argsort()[0]], g["ref_best_fit_pars"])
    assert np.array_equal(np.percentile(flat, [2.5, 16, 50, 84, 97.5], axis=0), g["ref_percentiles"])
    assert np.array_equal(np.percentile(lnflat, [2.5, 16, 50, 84, 97.5], axis=0), g["ref_lnp_percentiles"],
                          equal_nan=True)


@pytest.mark.needs_reference
def test_golden_is_what_the_reference_produces_today():
    from oracle import make_golden
    fresh = make_golden.chain_props_case(write=False)
    g = np.load(GOLDEN)
    for k in g.files:
        assert np.array_equal(np.asarray(fresh[k]), g[k], equal_nan=True), k
