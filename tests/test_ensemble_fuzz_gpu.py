"""GPU: the ensemble step on randomised model sets (every galaxy / disc / radio layout, 1-64 bands, every prior
combination): fused launch and three-kernel path against the CPU oracle step for step — proposals bit for bit,
identical accept decisions."""
import numpy as np
import pytest

from test_k1_fuzz_gpu import _random_case
from agnfitter_b200 import sources
from agnfitter_b200.engine import Engine
from agnfitter_b200.ensemble import EnsembleSampler
from oracle import agn_oracle as O
from oracle import ensemble_oracle as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(10))
def test_ensemble_on_random_modelsets(seed):
    ms, z, dlum, nus, fl, fe = _random_case(4000 + seed)
    src = sources.make_sources(ms, z, dlum, nus, fl, fe)
    osrc = [dict(z=float(z[s]), dlum=float(dlum[s]), nus=nus, fluxes=fl[s], fluxerrs=fe[s], pmin=src.pmin[s],
                 pmax=src.pmax[s], zi=int(np.nonzero(ms.z_grid == z[s])[0][0])) for s in range(src.S)]

    def lnprob(q):
        return np.stack([O.ln_probab(ms, osrc[s], q[s]) for s in range(src.S)])
    rng = np.random.default_rng(seed)
    W = int(rng.choice([8, 12, 20]))
    lo, hi = np.minimum(src.pmin, src.pmax), np.maximum(src.pmin, src.pmax)
    pos0 = lo[:, None, :] + rng.uniform(size=(src.S, W, ms.n_par)) * (hi - lo)[:, None, :]
    want_lnp0 = lnprob(pos0)
    for fused in (1, 3, 0):          # cooperative fused launch, gather4 fused launch, three kernels
        eng = Engine(ms, device=0)
        eng.set_sources(src)
        eng.set_option("ens_fused", 1 if fused else 0)
        if fused == 3:
            eng.set_option("k1_variant", 3)
        ens = EnsembleSampler(eng, W, seed=77 + seed, n_bins=0)
        ens.set_positions(pos0)
        opos, olnp, oacc = pos0.copy(), want_lnp0.copy(), np.zeros((src.S, W), dtype=np.int64)
        for it in range(6):
            ens.run(1)
            E.step(opos, olnp, oacc, lnprob, it, 77 + seed)
            gpos, glnp, gacc = ens.state()
            tag = "seed %d fused %d step %d n_b %d ga_free %s bb_mode %d radio %s" % (
                seed, fused, it, ms.n_b, ms.ga_free_ebv, ms.bb_mode, ms.radio)
            assert np.array_equal(gacc, oacc), "accept decisions differ: " + tag
            assert np.array_equal(gpos, opos), "positions differ: " + tag
            fin = np.isfinite(olnp)
            assert np.array_equal(np.isfinite(glnp), fin), tag
            assert np.allclose(glnp[fin], olnp[fin], rtol=1e-10, atol=0), tag
