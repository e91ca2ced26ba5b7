"""GPU: many concurrently running samplers share K1 launches through PosteriorMultiplexer; every value equals the
value a per-source call returns, bit for bit."""
import threading

import numpy as np
import pytest

from helpers import load_k1
from agnfitter_b200.dropin import CatalogPosterior
from agnfitter_b200.multiplex import PosteriorMultiplexer

pytestmark = pytest.mark.gpu


def test_multiplexed_values_equal_per_source_calls():
    ms, g = load_k1("cat49_catalog_example")
    reps = 40                                   # the catalogue's sources, repeated: 80 concurrent clients
    z = np.tile(g["src_z"], reps)
    S = len(z)
    cat = CatalogPosterior(ms, z, np.tile(g["src_dlum"], reps), np.tile(g["src_nus"], (reps, 1)),
                           np.tile(g["src_fluxes"], (reps, 1)), np.tile(g["src_fluxerrs"], (reps, 1)))
    asked = {}
    with PosteriorMultiplexer(cat, max_wait_s=0.02) as mux:
        def work(s):
            rng = np.random.default_rng(100 + s)
            mine = []
            with mux.client(s) as c:
                for _ in range(6):              # what ultranest sends: a few hundred points of the unit cube
                    cube = rng.uniform(-0.01, 1.01, size=(int(rng.integers(50, 400)), ms.n_par))
                    p = c.my_params_transform(cube)
                    mine.append((p, c.my_posterior(p)))
            asked[s] = mine
        th = [threading.Thread(target=work, args=(s,)) for s in range(S)]
        [t.start() for t in th]
        [t.join(timeout=120) for t in th]
        assert len(asked) == S and mux.requests == 6 * S
        assert mux.launches <= mux.requests // 4, "requests must have been merged (%d launches for %d requests)" % (
            mux.launches, mux.requests)
    for s, mine in asked.items():
        for p, l in mine:
            direct = cat.ln_probab_batch(p, np.full(len(p), s, dtype=np.int32))
            assert np.array_equal(l, direct, equal_nan=True)
            assert np.isfinite(l).any() and np.isneginf(l).any()        # in-box and out-of-box rows both occur
