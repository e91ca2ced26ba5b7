"""GPU: K1 (batched log-posterior) through the C ABI against (a) the golden vectors produced by the
reference's own functions and (b) the CPU oracle on fresh seeded inputs.
Bars (north_star): template row indices bit-exact; ymodel / lnprior / lnlike / lnp within 1e-10
relative in fp64, with identical placement of -inf / NaN."""
import numpy as np
import pytest

from conftest import K1_CASES
from helpers import load_k1, assert_close_with_specials, oracle_src, RTOL
from agnfitter_b200 import sources
from agnfitter_b200.engine import Engine
from oracle import agn_oracle as O

pytestmark = pytest.mark.gpu


def _engine(ms, g, variant=None):
    eng = Engine(ms, device=0)
    if variant is not None:
        eng.set_option("k1_variant", variant)
    src = sources.make_sources(ms, g["src_z"], g["src_dlum"], g["src_nus"], g["src_fluxes"], g["src_fluxerrs"])
    eng.set_sources(src)
    return eng, src


KERNEL_OF_VARIANT = {0: 0, 2: 1, 3: 2}


@pytest.mark.parametrize("variant", [0, 2, 3])
@pytest.mark.parametrize("case", K1_CASES)
def test_k1_matches_reference_golden(case, variant):
    """variant 3 = gather4 (TMA) kernel, 2 = cooperative kernel (forced; what the default, 1, picks when profitable),
    0 = per-thread kernel (fallback / cross-check)."""
    ms, g = load_k1(case)
    eng, _ = _engine(ms, g, variant)
    lnp, rows, ym, lpr, lll = eng.lnprob(g["pars"], g["source_id"], detail=True)
    assert eng.get_option("k1_last_kernel") == KERNEL_OF_VARIANT[variant], "the requested kernel must be the one that ran"
    assert np.array_equal(rows, g["ref_rows"]), "template indices must be bit-exact"
    assert_close_with_specials(ym, g["ref_ymodel"], what="ymodel")
    assert_close_with_specials(lll, g["ref_lnlike"], what="ln_likelihood")
    assert_close_with_specials(lpr, g["ref_lnprior"], what="ln_prior")
    assert_close_with_specials(lnp, g["ref_lnp"], what="ln_probab")
    # the production path (no detail outputs, early-out for out-of-box rows) gives the same lnp
    fast = eng.lnprob(g["pars"], g["source_id"])
    assert np.array_equal(np.isneginf(fast), np.isneginf(lnp))
    ok = np.isfinite(lnp)
    assert np.array_equal(fast[ok], lnp[ok])


@pytest.mark.parametrize("case", K1_CASES)
def test_k1_line_form_equals_pair_form(case):
    """The cooperative kernel's two band-phase mappings (`coop_line` 2: eight lanes per 128 bytes of a row, per-source
    operands cached in registers -- forced here, the default takes it for the 18-band sets only; 0: round 1's
    lane-per-chunk pair form) evaluate the same arithmetic per (row, chunk):
    every output is the same bit for bit, in batches that mix sources inside a warp and end in a ragged tile."""
    ms, g = load_k1(case)
    eng, src = _engine(ms, g, 2)
    rng = np.random.default_rng(4242)
    n = 7013
    sid = rng.integers(src.S, size=n).astype(np.int32)          # unsorted: the cached operands change row by row
    pars = src.pmin[sid] + rng.uniform(-0.01, 1.01, size=(n, ms.n_par)) * (src.pmax[sid] - src.pmin[sid])
    pars = np.concatenate([pars, g["pars"]])
    sid = np.concatenate([sid, g["source_id"].astype(np.int32)])
    got = {}
    for line in (2, 0):
        eng.set_option("coop_line", line)
        got[line] = eng.lnprob(pars, sid, detail=True) + (eng.lnprob(pars, sid),)
        assert eng.get_option("k1_last_kernel") == 1
    eng.set_option("coop_line", 1)
    for a, b in zip(got[2], got[0]):
        assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("case", ["classic18", "freeebv_xray49", "radio4_sn12_49"])
def test_k1_matches_oracle_on_fresh_positions(case):
    ms, g = load_k1(case)
    eng, src = _engine(ms, g)
    rng = np.random.default_rng(777)
    n_per = 20000 if case == "classic18" else 4000
    for s in range(src.S):
        pmin, pmax = src.pmin[s], src.pmax[s]
        pars = rng.uniform(pmin - 0.02 * (pmax - pmin), pmax + 0.02 * (pmax - pmin), size=(n_per, ms.n_par))
        pars[: n_per // 10] = 0.5 * (pmin + pmax) + (2 * rng.uniform(size=(n_per // 10, ms.n_par)) - 1) * 1e-5
        sid = np.full(n_per, s, dtype=np.int32)
        lnp, rows, ym, lpr, lll = eng.lnprob(pars, sid, detail=True)
        o_lnp, o_ym, o_rows, o_lpr, o_lll = O.ln_probab(ms, oracle_src(ms, g, s), pars, return_parts=True)
        assert np.array_equal(rows, o_rows)
        assert_close_with_specials(ym, o_ym, what="ymodel")
        assert_close_with_specials(lll, o_lll, what="lnlike")
        assert_close_with_specials(lpr, o_lpr, what="lnprior")
        assert_close_with_specials(lnp, o_lnp, what="lnp")


def test_k1_snap_ties_and_specials():
    """Exact midpoints between grid values, exact grid values, +-inf and NaN coordinates."""
    ms, g = load_k1("classic18")
    eng, src = _engine(ms, g)
    rng = np.random.default_rng(5)
    pmin, pmax = src.pmin[0], src.pmax[0]
    n = 6000
    pars = rng.uniform(pmin, pmax, size=(n, ms.n_par))
    col = 0
    for f in ms.families:
        for d in range(f.n_dim):
            if f.par_types[d] == 'grid':
                gs = np.sort(f.grid[d])
                i = rng.integers(len(gs) - 1, size=n)
                kind = rng.integers(5, size=n)
                mid = np.where(kind == 0, (gs[i] + gs[i + 1]) / 2, pars[:, col])
                mid = np.where(kind == 1, gs[i], mid)
                mid = np.where(kind == 2, np.nextafter((gs[i] + gs[i + 1]) / 2, np.inf), mid)
                mid = np.where(kind == 3, np.nextafter((gs[i] + gs[i + 1]) / 2, -np.inf), mid)
                pars[:, col] = mid
            col += 1
    pars[0, 0] = np.nan
    pars[1, 1] = np.inf
    pars[2, 3] = -np.inf
    pars[3, 4] = 1e300
    sid = np.zeros(n, dtype=np.int32)
    lnp, rows, ym, lpr, lll = eng.lnprob(pars, sid, detail=True)
    o_lnp, o_ym, o_rows, o_lpr, o_lll = O.ln_probab(ms, oracle_src(ms, g, 0), pars, return_parts=True)
    assert np.array_equal(rows, o_rows)
    assert_close_with_specials(lnp, o_lnp, what="lnp")


def test_k1_full_size_properties():
    """At BASELINE size (1e7 rows is the bench; here 2e6): results do not depend on batch order or
    batch split, and equal the small-batch results bit for bit."""
    ms, g = load_k1("classic18")
    eng, src = _engine(ms, g)
    rng = np.random.default_rng(99)
    n = 2_000_000
    sid = rng.integers(src.S, size=n).astype(np.int32)
    pars = src.pmin[sid] + rng.uniform(size=(n, ms.n_par)) * (src.pmax[sid] - src.pmin[sid])
    a = eng.lnprob(pars, sid)
    perm = rng.permutation(n)
    b = eng.lnprob(pars[perm], sid[perm])
    assert np.array_equal(a[perm], b, equal_nan=True)
    c = np.concatenate([eng.lnprob(pars[i:i + 333_333], sid[i:i + 333_333]) for i in range(0, n, 333_333)])
    assert np.array_equal(a, c, equal_nan=True)
    k = rng.choice(n, size=3000, replace=False)
    for s in range(src.S):
        kk = k[sid[k] == s]
        o = O.ln_probab(ms, oracle_src(ms, g, s), pars[kk])
        assert_close_with_specials(a[kk], o, what="lnp sample")


def test_k1_empty_and_errors():
    ms, g = load_k1("agnoff18")
    eng, src = _engine(ms, g)
    assert eng.lnprob(np.empty((0, ms.n_par))).shape == (0,)
    with pytest.raises(ValueError):
        eng.lnprob(np.zeros((3, ms.n_par + 1)))
    with pytest.raises(IndexError):
        eng.lnprob(np.zeros((3, ms.n_par)), np.array([0, 1, 7], dtype=np.int32))


def test_k1_variants_agree_and_tail_handling():
    """All kernels on batches whose size is not a multiple of the warp / CTA size; the gather4 and the cooperative
    kernel do the same arithmetic in the same order: bit-identical."""
    ms, g = load_k1("classic18")
    rng = np.random.default_rng(3)
    outs = {}
    for variant in (0, 2, 3):
        eng, src = _engine(ms, g, variant)
        res = []
        for n in (1, 31, 33, 127, 129, 255, 257, 1000, 4097):
            r2 = np.random.default_rng(n)
            sid = r2.integers(src.S, size=n).astype(np.int32)
            pars = src.pmin[sid] + r2.uniform(size=(n, ms.n_par)) * (src.pmax[sid] - src.pmin[sid])
            res.append(eng.lnprob(pars, sid, detail=True))
        outs[variant] = res
    for a, b, c in zip(outs[0], outs[2], outs[3]):
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[1], c[1])      # rows
        for k in (0, 2, 3, 4):
            assert_close_with_specials(b[k], a[k], rtol=1e-12, what="variant output %d" % k)
            assert np.array_equal(b[k], c[k], equal_nan=True), "gather4 vs cooperative, output %d" % k


def test_exp10_fast_accuracy():
    """The cooperative kernel's own 10**x (2^(x D log2 10 / D): D-entry table + polynomial) against NumPy, incl.
    special values.  The argument scaling costs |x| * 3.3 ulp of relative accuracy; inside the prior box
    (|x| <= 10) that is 4e-15, three orders below what 1e-10 on ln_probab needs."""
    ms, g = load_k1("agnoff18")
    eng, _ = _engine(ms, g)
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-40, 40, 100000), rng.uniform(-299, 299, 50000),
                        rng.uniform(-1e-3, 1e-3, 1000),
                        np.array([0.0, -0.0, 1.0, -1.0, 10.0, -10.0, 299.999, -299.999, 300.0, -300.0, 308.0,
                                  -320.0, 400.0, -400.0, np.inf, -np.inf, np.nan])])
    y = eng.debug_exp10(x)
    with np.errstate(all="ignore"):
        want = 10.0 ** x
    assert np.array_equal(np.isnan(y), np.isnan(want)) and np.array_equal(np.isinf(y), np.isinf(want))
    fin = np.isfinite(want) & (want > 1e-300)
    rel = np.abs(y[fin] - want[fin]) / want[fin]
    bound = 4e-16 * (1.0 + 3.4 * np.abs(x[fin]))
    assert np.all(rel <= bound), (rel / bound).max()
    assert np.all(rel[np.abs(x[fin]) <= 10] <= 1.5e-14)
