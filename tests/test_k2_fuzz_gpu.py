"""GPU: K2 on randomised filter sets and template grids — 1-30 bands of 5-400 samples (stored ascending or
descending), template grids of 12-600 samples in ascending / descending order, 1-70 templates, dense and sparse
z grids whose length is not a multiple of the batch sizes: the tensor-core, shared-window and per-slice
contractions against each other (1e-13, same zero pattern) and against the CPU oracle (1e-10)."""
import numpy as np
import pytest

from helpers import assert_close_with_specials
from agnfitter_b200 import synth
from agnfitter_b200.engine import DictBuilder
from oracle import agn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(10))
def test_k2_random_filter_sets(seed):
    rng = np.random.default_rng(3000 + seed)
    n_b = int(rng.integers(1, 31))
    nus = np.sort(rng.uniform(11.6, 15.2, size=n_b)).round(6)
    while len(np.unique(nus)) != n_b:
        nus = np.sort(rng.uniform(11.6, 15.2, size=n_b)).round(6)
    nks = [int(v) for v in rng.choice([5, 23, 128, 400], size=n_b)]
    spans = [float(v) for v in rng.uniform(0.05, 0.6, size=n_b)]
    dec = tuple(int(i) for i in np.nonzero(rng.uniform(size=n_b) < 0.25)[0])
    fd = synth.synthetic_filterdict(nus=list(nus), nks=nks, spans=spans, decreasing=dec)
    builder = DictBuilder(fd, device=0)
    n_lam = int(rng.choice([12, 57, 200, 407, 600]))
    lognu = np.sort(rng.uniform(10.8, 16.2, size=n_lam))
    if rng.integers(2):
        lognu = lognu[::-1].copy()
    T = int(rng.integers(1, 71))
    F = {str(t): (lognu, 1e-28 * rng.lognormal(0, 1.5, size=n_lam)) for t in range(T)}
    if T > 2:                                            # a template that vanishes over part of the spectrum
        f2 = F["1"][1].copy()
        f2[: n_lam // 3] = 0.0
        F["1"] = (lognu, f2)
    nz = int(rng.integers(1, 40))
    if rng.integers(2):
        z = np.round(rng.uniform(0, 3) + 0.001 * np.arange(nz), 3)
    else:
        z = np.unique(np.round(rng.uniform(0, 5.9, size=nz), 3))
    got = {}
    for name, mma, sw in (("mma", 1, 1), ("shared", 0, 1), ("slice", 0, 0)):
        builder.set_option("k2_mma", mma)
        builder.set_option("k2_shared_window", sw)
        got[name] = builder.build_family(F, z)
        dev = builder.build_family(F, z, on_device=True).cpu().numpy()
        assert np.array_equal(dev[:, :, :n_b], got[name]) and not dev[:, :, n_b:].any(), (seed, name)
    tag = "seed %d n_b %d n_lam %d T %d n_z %d" % (seed, n_b, n_lam, T, len(z))
    for name in ("mma", "shared"):
        assert np.array_equal(got[name] == 0.0, got["slice"] == 0.0), (tag, name)
        assert_close_with_specials(got[name], got["slice"], rtol=1e-13, what="%s %s" % (tag, name))
    bands = fd[0]
    lam = [np.asarray(fd[1][b][0]) for b in bands]
    fac = [np.asarray(fd[2][b][0]) for b in bands]
    for _ in range(6):
        zi, ti = int(rng.integers(len(z))), int(rng.integers(T))
        want = O.filtering_models(F[str(ti)][0], F[str(ti)][1], bands, lam, fac, float(z[zi]))
        assert np.array_equal(got["mma"][zi, ti] == 0.0, want == 0.0), tag
        assert_close_with_specials(got["mma"][zi, ti], want, what=tag)
