"""GPU: K2 (dictionary build) through the C ABI against the reference's filtering_models outputs
(golden) and against the CPU oracle on synthetic templates.  Band fluxes within 1e-10 relative;
entries the reference computes as exactly 0 (no spectral coverage) must be exactly 0."""
import numpy as np
import pytest

from conftest import K2_CASES
from helpers import load_k2, assert_close_with_specials
from agnfitter_b200.engine import DictBuilder
from agnfitter_b200 import synth
from oracle import agn_oracle as O

pytestmark = pytest.mark.gpu


def _filterdict(g):
    bands, off = g["bands"], g["filt_offset"]
    ld = {b: [g["filt_lambda"][off[i]:off[i + 1]]] for i, b in enumerate(bands)}
    fd = {b: [g["filt_factor"][off[i]:off[i + 1]]] for i, b in enumerate(bands)}
    return [bands, ld, fd]


@pytest.mark.parametrize("case", K2_CASES)
def test_k2_matches_reference_golden(case):
    g = load_k2(case)
    builder = DictBuilder(_filterdict(g), device=0)
    for fam in [str(x) for x in g["family_names"]]:
        ref = g[fam + "_ref"]                       # [n_z, T, n_b]
        tpl = {}
        for i in range(ref.shape[1]):
            n = int(g[fam + "_len"][i])
            tpl[i] = (g[fam + "_lognu"][i, :n], g[fam + "_Fnu"][i, :n])
        got = builder.build_family(tpl, g["z"])
        assert got.shape == ref.shape
        assert np.array_equal(got == 0.0, ref == 0.0), "%s: zero pattern (coverage) differs" % fam
        assert_close_with_specials(got, ref, what=fam)
        if fam in ("GA", "TO_S04"):
            dev = builder.build_family(tpl, g["z"], on_device=True).cpu().numpy()
            n_b = got.shape[2]
            assert dev.shape[2] == (n_b + 1) // 2 * 2, "device tables carry an even row pitch"
            assert np.array_equal(dev[..., :n_b], got)
            assert not dev[..., n_b:].any(), "pad column must be zero"


def test_k2_filtering_models_signature():
    g = load_k2("default18")
    builder = DictBuilder(_filterdict(g), device=0)
    bands, F = builder.filtering_models(g["BB_R06_lognu"][0], g["BB_R06_Fnu"][0], 0.283)
    assert F.shape == (len(g["bands"]), 1) and np.array_equal(bands, g["bands"])
    assert_close_with_specials(F[:, 0], g["BB_R06_ref"][1, 0], what="filtering_models")


def test_k2_matches_oracle_on_synthetic_families():
    fd = synth.synthetic_filterdict()
    builder = DictBuilder(fd, device=0)
    tpl = synth.synthetic_templates()
    bands = fd[0]
    lam = [np.asarray(fd[1][b][0]) for b in bands]
    fac = [np.asarray(fd[2][b][0]) for b in bands]
    z = np.array([0.0, 0.011, 0.5, 2.345, 5.999])
    rng = np.random.default_rng(1)
    for name in ("GA", "SB", "TO", "BB"):
        keys = list(tpl[name]["F"].keys())
        pick = [keys[i] for i in sorted(rng.choice(len(keys), size=min(len(keys), 24), replace=False))]
        sub = {k: tpl[name]["F"][k] for k in pick}
        got = builder.build_family(sub, z)
        for zi, zz in enumerate(z):
            for ti, k in enumerate(pick):
                want = O.filtering_models(sub[k][0], sub[k][1], bands, lam, fac, float(zz))
                assert np.array_equal(got[zi, ti] == 0.0, want == 0.0)
                assert_close_with_specials(got[zi, ti], want, what="%s z=%g" % (name, zz))


def test_k2_dense_zgrid_linearity_and_row_map():
    """Size-independent properties on a dense z grid: linearity in the template flux, and ragged
    families (two frequency grids) scatter into the right rows."""
    fd = synth.synthetic_filterdict()
    builder = DictBuilder(fd, device=0)
    tpl = synth.synthetic_templates()
    z = np.round(np.arange(0.0, 6.0, 0.01), 3)
    to = tpl["TO"]["F"]
    a = builder.build_family(to, z)
    scaled = {k: (v[0], 3.0 * v[1]) for k, v in to.items()}
    b = builder.build_family(scaled, z)
    assert_close_with_specials(b, 3.0 * a, rtol=1e-14, what="linearity")
    # ragged: every other template on a truncated grid
    rag = {}
    for i, (k, v) in enumerate(to.items()):
        rag[k] = (v[0][:-5], v[1][:-5]) if i % 2 else v
    c = builder.build_family(rag, z[:7])
    full_sub = builder.build_family({k: v for i, (k, v) in enumerate(to.items()) if i % 2 == 0}, z[:7])
    assert np.array_equal(c[:, 0::2], full_sub)


def test_k2_contraction_kernels_agree():
    """The tensor-core contraction (FP64 mma, 8 slices per batch), the shared-window contraction (4 slices per read
    of the template window) and the per-slice kernels: same zero pattern, values within 1e-13 (different summation
    grouping), on slice counts that are not multiples of the batch sizes, host and device outputs; a family with
    an infinite sample falls back to the per-slice kernels."""
    fd = synth.synthetic_filterdict()
    builder = DictBuilder(fd, device=0)
    tpl = synth.synthetic_templates()
    # sparse grids (slices far apart: the union window of 8 slices may not fit -> 4-slice kernel) and dense ones
    for fam, nz, dense in (("TO", 37, False), ("SB", 8, False), ("BB", 3, False), ("GA", 6, False), ("GA", 21, True),
                           ("GA", 8, True)):
        F = tpl[fam]["F"]
        F = {k: F[k] for k in list(F.keys())[:211]}
        z = np.round(0.731 + 0.001 * np.arange(nz), 3) if dense else np.round(np.linspace(0.0, 5.9, nz), 3)
        got = {}
        for name, mma, sw in (("mma", 1, 1), ("shared", 0, 1), ("slice", 0, 0)):
            builder.set_option("k2_mma", mma)
            builder.set_option("k2_shared_window", sw)
            got[name] = builder.build_family(F, z)
            kern = builder.get_option("k2_last_kernel")
            # the 1 366-sample starburst grid does not leave room for a 32-template tile
            want = 1 if (fam == "SB" or name == "slice") else (3 if name == "mma" else 2)
            if name == "mma" and fam != "SB" and not (dense or fam == "TO"):
                assert kern in (2, 3), (fam, name, kern)     # the 8-slice weight block may not fit next to the tile
            else:
                assert kern == want, (fam, name, kern)
            if name != "slice":
                dev = builder.build_family(F, z, on_device=True).cpu().numpy()[:, :, :builder.n_b]
                assert np.array_equal(got[name], dev), (fam, name)
        builder.set_option("k2_mma", 1)
        builder.set_option("k2_shared_window", 1)
        for name in ("mma", "shared"):
            assert np.array_equal(got[name] == 0.0, got["slice"] == 0.0), (fam, name)
            assert_close_with_specials(got[name], got["slice"], rtol=1e-13, what="%s vs per-slice %s" % (name, fam))
    to = dict(list(tpl["TO"]["F"].items())[:40])
    k0 = next(iter(to))
    bad = np.array(to[k0][1], dtype=np.float64).copy()
    bad[len(bad) // 2] = np.inf
    to[k0] = (to[k0][0], bad)
    got = builder.build_family(to, np.array([0.5, 0.501, 0.502, 0.503]))
    assert builder.get_option("k2_last_kernel") == 1
    assert np.isfinite(got[:, 1:]).all()


@pytest.mark.parametrize("n_b", [1, 5, 17])
def test_k2_odd_band_counts(n_b):
    """Band counts that are odd (host outputs have an odd row pitch: the result tile is stored by the store loop
    instead of bulk stores) or tiny, through all three contraction paths and against the oracle."""
    nus, nks, spans = synth.DEFAULT18_NU[:n_b], synth.DEFAULT18_NK[:n_b], synth.DEFAULT18_SPAN[:n_b]
    fd = synth.synthetic_filterdict(nus=nus, nks=nks, spans=spans, decreasing=(0,))
    builder = DictBuilder(fd, device=0)
    tpl = synth.synthetic_templates()
    F = dict(list(tpl["GA"]["F"].items())[:75])
    z = np.round(1.204 + 0.001 * np.arange(19), 3)
    got = {}
    for name, mma, sw in (("mma", 1, 1), ("shared", 0, 1), ("slice", 0, 0)):
        builder.set_option("k2_mma", mma)
        builder.set_option("k2_shared_window", sw)
        got[name] = builder.build_family(F, z)
        assert builder.get_option("k2_last_kernel") == {"mma": 3, "shared": 2, "slice": 1}[name]
        dev = builder.build_family(F, z, on_device=True).cpu().numpy()
        assert np.array_equal(dev[:, :, :n_b], got[name]) and not dev[:, :, n_b:].any()      # pad column stays zero
    for name in ("mma", "shared"):
        assert np.array_equal(got[name] == 0.0, got["slice"] == 0.0)
        assert_close_with_specials(got[name], got["slice"], rtol=1e-13, what=name)
    bands = fd[0]
    lam = [np.asarray(fd[1][b][0]) for b in bands]
    fac = [np.asarray(fd[2][b][0]) for b in bands]
    keys = list(F.keys())
    for zi, ti in ((0, 0), (7, 33), (18, 74)):
        want = O.filtering_models(F[keys[ti]][0], F[keys[ti]][1], bands, lam, fac, float(z[zi]))
        assert_close_with_specials(got["mma"][zi, ti], want, what="oracle z=%g t=%d" % (z[zi], ti))


def test_k2_full_dense_zgrid_is_slice_independent():
    """BASELINE config C5: z = arange(0, 6, 0.001) (6 000 slices).  Size-independent properties at full size:
    every slice of the dense build equals, bit for bit, the same slice built on its own (the batching of
    z-slices inside the tiled kernel must not leak between slices), and sampled entries match the oracle."""
    fd = synth.synthetic_filterdict()
    builder = DictBuilder(fd, device=0)
    tpl = synth.synthetic_templates()
    keys = list(tpl["GA"]["F"].keys())[::10]                     # 1 100 galaxy templates
    sub = {k: tpl["GA"]["F"][k] for k in keys}
    z = np.round(np.arange(0.0, 6.0, 0.001), 3)
    dense = builder.build_family(sub, z, on_device=True)          # [6000, 1100, 18] stays in HBM
    assert tuple(dense.shape) == (6000, len(keys), 18)
    rng = np.random.default_rng(2)
    for zi in sorted(rng.choice(len(z), size=12, replace=False)) + [0, 1, 2, 3, 5997, 5998, 5999]:
        alone = builder.build_family(sub, z[zi:zi + 1])[0]
        assert np.array_equal(dense[zi].cpu().numpy(), alone), "slice %d depends on its neighbours" % zi
    bands = fd[0]
    lam = [np.asarray(fd[1][b][0]) for b in bands]
    fac = [np.asarray(fd[2][b][0]) for b in bands]
    for zi, ti in zip(rng.choice(len(z), size=25), rng.choice(len(keys), size=25)):
        want = O.filtering_models(sub[keys[ti]][0], sub[keys[ti]][1], bands, lam, fac, float(z[zi]))
        got = dense[zi, ti].cpu().numpy()
        assert np.array_equal(got == 0.0, want == 0.0)
        assert_close_with_specials(got, want, what="z=%g t=%d" % (z[zi], ti))


def test_k2_staged_templates_are_reused_and_revalidated():
    """build_family(reuse_staging=True) keeps the grouped, pinned template blocks of a family and reuses them when the
    same dict comes again (catalogue chunks, rebuilds for other redshifts): same tables bit for bit, for ragged
    families too; a dict whose values changed is staged afresh."""
    g = load_k2("default18")
    builder = DictBuilder(_filterdict(g), device=0)
    fams = [str(x) for x in g["family_names"]]
    ragged = {}
    for fam in fams[:3]:                                  # three families merged: several frequency grids in one dict
        for i in range(g[fam + "_ref"].shape[1]):
            n = int(g[fam + "_len"][i])
            ragged[(fam, i)] = (g[fam + "_lognu"][i, :n], g[fam + "_Fnu"][i, :n])
    for tpl in (ragged, {k: v for k, v in list(ragged.items())[:7]}):
        want = builder.build_family(tpl, g["z"])
        first = builder.build_family(tpl, g["z"], reuse_staging=True)
        staged = builder._tpl_cache[id(tpl)][2]
        again = builder.build_family(tpl, g["z"][::-1].copy(), reuse_staging=True)[::-1]
        assert builder._tpl_cache[id(tpl)][2] is staged, "the second call must reuse the staged blocks"
        assert np.array_equal(first, want, equal_nan=True) and np.array_equal(again, want, equal_nan=True)
        k0 = next(iter(tpl))
        tpl[k0] = (tpl[k0][0], tpl[k0][1] * 2.0)            # a new value tuple: the cache entry no longer matches
        doubled = builder.build_family(tpl, g["z"], reuse_staging=True)
        assert builder._tpl_cache[id(tpl)][2] is not staged
        assert np.array_equal(doubled, builder.build_family(tpl, g["z"]), equal_nan=True)
        assert not np.array_equal(doubled[:, 0], want[:, 0])
