"""GPU: randomised model sets.  Band counts from 1 to 64 (odd and even: padded pitch, runtime chunk count),
unsorted non-uniform grids with 1-3 dimensions, free E(B-V)_gal on/off, all four BBB parameter layouts, radio
on/off (fitted or fixed template), every prior combination — both kernels against the CPU oracle:
rows bit-exact, ymodel / ln_prior / ln_likelihood / ln_probab within 1e-10."""
import itertools

import numpy as np
import pytest

from helpers import assert_close_with_specials
from agnfitter_b200 import sources
from agnfitter_b200.engine import Engine
from agnfitter_b200.flatten import ModelSet, family_from_keys_float
from oracle import agn_oracle as O

pytestmark = pytest.mark.gpu


def _grid(rng, n, lo, hi):
    """n distinct values, 4 decimals, in random (first-appearance) order, non-uniform spacing."""
    while True:
        v = np.unique(np.round(lo + (hi - lo) * np.sort(rng.uniform(size=n)) ** rng.uniform(0.6, 1.6), 4))
        if len(v) == n:
            return rng.permutation(v)


def _family(rng, name, names, types, sizes, ranges, n_z, n_b, with_aux, fitted=True, scale=1e-28):
    vals = []
    for t, n, (lo, hi) in zip(types, sizes, ranges):
        vals.append(np.array([0.0, 1.0]) if t == 'free' else _grid(rng, n, lo, hi))
    keys = np.array(list(itertools.product(*vals))).T            # [n_dim, T], first dimension slowest
    T = keys.shape[1]
    table = scale * rng.lognormal(0.0, 1.5, size=(n_z, T, n_b))
    aux = scale * 1e70 * rng.lognormal(0.0, 2.0, size=T) if with_aux else None
    return family_from_keys_float(name, names, types, keys, fitted, table, aux)


def _random_case(seed):
    rng = np.random.default_rng(seed)
    n_b = int(rng.choice([1, 2, 3, 7, 9, 17, 18, 25, 32, 33, 49, 50, 63, 64]))
    ga_free = bool(rng.integers(2))
    bb_mode = int(rng.integers(4))
    radio = bool(rng.integers(2))
    bbb_norm = bool(rng.integers(2))
    eb = str(rng.choice(['Flexible', 'Restrictive', 'None']))
    agnfrac = bool(rng.integers(2))
    xrays = bool(bb_mode >= 2)
    n_z = int(rng.integers(1, 4))
    z_grid = np.sort(np.round(rng.uniform(0.0, 0.5, size=n_z), 3))
    while len(np.unique(z_grid)) != n_z:
        z_grid = np.sort(np.round(rng.uniform(0.0, 0.5, size=n_z), 3))
    # bands: always one in the 1500 A window, plus radio / far-IR / X-ray bands when there is room
    must = [14.62]
    if n_b >= 4:
        must += [9.21, 12.47, 17.35]
    bands = np.sort(np.concatenate([must[:n_b], rng.uniform(11.0, 15.2, size=max(0, n_b - len(must)))]))[:n_b]
    settings = dict(BBB='R06' if bbb_norm else 'SN12', RADIO=radio, XRAYS=xrays, turn_on_AGN=True,
                    PRIOR_energy_balance=eb, PRIOR_AGNfraction=agnfrac, PRIOR_midIR_UV=False,
                    PRIOR_galaxy_only=False, GALAXY='BC03', STARBURST='DH02_CE01', TORUS='S04', Blazar='None')
    ms = ModelSet(settings=settings, bands=bands, z_grid=z_grid)
    if ga_free:
        ms.families.append(_family(rng, "GA", ['tau', 'age', 'EBVgal'], ['grid', 'grid', 'free'],
                                   [int(rng.integers(2, 7)), int(rng.integers(2, 9)), 2], [(0.1, 15), (5, 10), (0, 1)],
                                   n_z, n_b, False, scale=1e-30))
    else:
        ms.families.append(_family(rng, "GA", ['tau', 'age', 'EBVgal'], ['grid'] * 3,
                                   [int(rng.integers(1, 6)), int(rng.integers(2, 8)), int(rng.integers(2, 12))],
                                   [(0.1, 15), (5, 10), (0.06, 1)], n_z, n_b, True, scale=1e-30))
    ms.families.append(_family(rng, "SB", ['irlum'], ['grid'], [int(rng.integers(2, 60))], [(8, 14)], n_z, n_b, True))
    nd = int(rng.integers(1, 4))
    ms.families.append(_family(rng, "TO", ['oa', 'incl', 'tv'][:nd], ['grid'] * nd,
                               [int(rng.integers(2, 8)) for _ in range(nd)], [(10, 80), (0, 90), (3, 11)][:nd],
                               n_z, n_b, False, scale=1e-29))
    bb = {0: (['logBHmass', 'logEddra'], ['grid', 'grid'], [4, 5], [(6, 10), (-3, 0)]),
          1: (['EBVbbb'], ['free'], [2], [(0, 1)]),
          2: (['logBHmass', 'EBVbbb', 'alphaScat'], ['grid', 'free', 'free'], [5, 2, 2], [(6, 10), (0, 1), (-0.4, 0.4)]),
          3: (['EBVbbb', 'alphaScat', 'Gamma'], ['free', 'free', 'grid'], [2, 2, 4], [(0, 1), (-0.4, 0.4), (1.6, 2.4)])}[bb_mode]
    ms.families.append(_family(rng, "BB", *bb, n_z, n_b, False, scale=1e-31 if bbb_norm else 1e25))
    if radio:
        if rng.integers(2):
            ms.families.append(_family(rng, "RAD", ['curv', 'Ecut'], ['grid', 'grid'], [3, 6], [(-1, 0), (10, 14)],
                                       n_z, n_b, False, scale=1e-27))
        else:
            f = _family(rng, "RAD", ['fix'], ['grid'], [1], [(-100, -99)], n_z, n_b, False, fitted=False, scale=1e-27)
            ms.families.append(f)
    # sources
    S = 3
    z = rng.choice(z_grid, size=S)
    dlum = rng.uniform(1e26, 3e28, size=S)
    nus = np.round(bands, 4)
    fl = 1e-27 * rng.lognormal(0.0, 1.0, size=(S, n_b))
    fe = 0.1 * fl * rng.uniform(0.5, 2.0, size=(S, n_b))
    u = rng.uniform(size=(S, n_b))
    missing = (u < 0.1) & (np.abs(bands - 14.62) > 1e-9)[None, :]
    fl[missing] = -99e-26
    fe[missing] = -99e-26
    return ms, z, dlum, nus, fl, fe


def _positions(rng, ms, pmin, pmax, n):
    # (a random grid can fall outside Pdict's overrides, e.g. EBVgal >= 0.05: the box is then empty, pmax < pmin,
    # and every row is out of box — a legitimate case, so the generator must not choke on it)
    span = np.abs(pmax - pmin)
    lo, hi = np.minimum(pmin, pmax), np.maximum(pmin, pmax)
    p = rng.uniform(lo - 0.02 * span, hi + 0.02 * span, size=(n, ms.n_par))
    p[: n // 8] = 0.5 * (pmin + pmax) + (2 * rng.uniform(size=(n // 8, ms.n_par)) - 1) * 1e-5
    col = 0
    for f in ms.families:
        if not f.fitted:
            continue
        for d in range(f.n_dim):
            if f.par_types[d] == 'grid' and len(f.grid[d]) > 1:
                g = np.sort(f.grid[d])
                k = rng.integers(len(g) - 1, size=n)
                kind = rng.integers(8, size=n)
                p[:, col] = np.where(kind == 0, (g[k] + g[k + 1]) / 2, p[:, col])          # exact midpoints (ties)
                p[:, col] = np.where(kind == 1, g[k], p[:, col])                            # exact grid values
                p[:, col] = np.where(kind == 2, np.nextafter((g[k] + g[k + 1]) / 2, np.inf), p[:, col])
            col += 1
    return p


@pytest.mark.parametrize("seed", range(16))
def test_random_modelsets_match_the_oracle(seed):
    ms, z, dlum, nus, fl, fe = _random_case(1000 + seed)
    src = sources.make_sources(ms, z, dlum, nus, fl, fe)
    rng = np.random.default_rng(seed)
    n_per = 1500
    # (kernel, band-phase mapping of the cooperative kernel): gather4, cooperative with the line form forced wherever it
    # can run (>= 16 bands; the default takes it for the 18-band sets only), cooperative pair form, per-thread
    for variant, line in ((3, 1), (2, 2), (2, 0), (0, 1)):
        eng = Engine(ms, device=0)
        eng.set_option("k1_variant", variant)
        eng.set_option("coop_line", line)
        eng.set_sources(src)
        for s in range(src.S):
            pars = _positions(np.random.default_rng(seed * 10 + s), ms, src.pmin[s], src.pmax[s], n_per)
            sid = np.full(n_per, s, dtype=np.int32)
            lnp, rows, ym, lpr, lll = eng.lnprob(pars, sid, detail=True)
            zi = int(np.nonzero(ms.z_grid == z[s])[0][0])
            osrc = dict(z=float(z[s]), dlum=float(dlum[s]), nus=nus, fluxes=fl[s], fluxerrs=fe[s], pmin=src.pmin[s],
                        pmax=src.pmax[s], zi=zi)
            o_lnp, o_ym, o_rows, o_lpr, o_lll = O.ln_probab(ms, osrc, pars, return_parts=True)
            tag = "seed %d variant %d line %d n_b %d" % (seed, variant, line, ms.n_b)
            assert np.array_equal(rows, o_rows), tag
            assert_close_with_specials(ym, o_ym, what="ymodel " + tag)
            assert_close_with_specials(lll, o_lll, what="lnlike " + tag)
            assert_close_with_specials(lpr, o_lpr, what="lnprior " + tag)
            assert_close_with_specials(lnp, o_lnp, what="lnp " + tag)
            assert np.array_equal(eng.lnprob(pars, sid)[np.isfinite(lnp)], lnp[np.isfinite(lnp)])
        # the cooperative kernel may only refuse when its per-warp tiles cannot fit shared memory: three
        # partial-sum planes (free E(B-V) + energy balance) on more than ~56 bands
        too_big = ms.ga_free_ebv and ms.settings['PRIOR_energy_balance'] != 'None' and ms.n_b > 56
        if variant == 2 and not too_big:
            why = "n_b=%d ga_free=%s bb_mode=%d radio=%s" % (ms.n_b, ms.ga_free_ebv, ms.bb_mode, ms.radio)
            assert eng.get_option("k1_last_kernel") == 1, "cooperative kernel refused: " + why
        if variant == 3 and ms.n_par <= 20:
            assert eng.get_option("k1_last_kernel") == 2, "gather4 kernel refused: n_b=%d n_par=%d" % (ms.n_b, ms.n_par)
