"""CPU: the host-side preparation of K4 (agnfitter_b200/realizations.py) — bracket tables and weights against
numpy.interp on the golden templates (incl. the two-part interpolation of X-ray discs and the -100 fill), the
translation of out['intlum_models'] / out['intlum_freqranges'] into bands and output columns against the oracle's
direct evaluation, and the branch selection of the `_4plot` get_fluxes."""
import numpy as np
import pytest

from conftest import SED_CASES
from helpers import load_sed
from agnfitter_b200 import realizations as R
from oracle import sed_oracle as SO


def _numpy_eval(xs, ys, grid, jt, wt):
    out = np.full(len(grid), -100.0)
    on = jt >= 0
    j = jt[on]
    jn = np.minimum(j + 1, len(xs) - 1)
    with np.errstate(all="ignore"):
        out[on] = np.where(wt[on] == 0.0, ys[j], ys[j] + wt[on] * (ys[jn] - ys[j]))
    return out


@pytest.mark.parametrize("case", SED_CASES)
def test_bracket_tables_reproduce_numpy_interp(case):
    ms, ps, g = load_sed(case)
    grid = R.all_nus_rest(ms.settings)
    assert np.array_equal(grid, g["all_nus_rest"])
    split_bb = (ms.settings['BBB'] != 'KD18') and (ms.settings['XRAYS'] == True)
    rng = np.random.default_rng(0)
    for pf in ps.families:
        split = split_bb and pf.name == "BB"
        for gi, x in enumerate(pf.lognu):
            order = np.argsort(x, kind="mergesort")
            xs = x[order]
            jt = R.bracket_table(xs, grid, split)
            wt = R.bracket_weights(xs, grid, jt)
            assert jt.dtype == np.int32 and jt.min() >= -1 and jt.max() < len(xs)
            assert np.all((wt >= 0) & (wt < 1)) and np.all(wt[jt < 0] == 0)
            inside = jt >= 0
            assert np.all(xs[jt[inside]] <= grid[inside])
            strict = inside & (wt > 0)
            assert np.all(grid[strict] < xs[jt[strict] + 1])
            ys = rng.uniform(-40, -20, size=len(xs))                 # finite log10 F: compare with the oracle's rule
            if split:
                lo, hi = xs < 16.685, xs >= 16.685
                glo = grid < 16.685
                want = np.concatenate((SO.loglin_interp(xs[lo], ys[lo], grid[glo]),
                                       SO.loglin_interp(xs[hi], ys[hi], grid[~glo])))
            else:
                want = SO.loglin_interp(xs, ys, grid)
            got = _numpy_eval(xs, ys, grid, jt, wt)
            assert np.array_equal(got == -100.0, want == -100.0), (case, pf.name, gi)
            assert np.allclose(got, want, rtol=0, atol=2e-13), (case, pf.name, gi)


@pytest.mark.parametrize("case", ["classic18", "radio_r06_49"])
def test_band_plan_matches_the_oracle(case):
    """Feed the plan synthetic per-component spectra: integrating them over the plan's index ranges and combining
    them by its (kind, numerator, denominator) columns gives integrated_luminosities' numbers."""
    ms, ps, g = load_sed(case)
    grid = g["all_nus_rest"]
    models = [str(m) for m in g["intlum_models"]]
    hz = g["intlum_ranges_hz"]
    bands, cols = R.band_plan(models, hz, grid, ms.radio)
    assert len(cols) == len(models) and len(bands) <= len(models)
    rng = np.random.default_rng(1)
    comps = ['sb', 'bb', 'ga', 'to', 'rad', 'bbdered']
    nl = {c: 10 ** rng.uniform(40, 46, size=(3, len(grid))) for c in comps if c != 'rad' or ms.radio}
    with np.errstate(all="ignore"):
        want = SO.integrated_luminosities(models, hz, grid, nl, ms.radio).T
    x = 10 ** grid
    for s, (kind, q, num, den) in enumerate(cols):
        i0, i1, mask = bands[q]
        assert (num | den) & ~mask == 0

        def value(c):
            a = nl[comps[c]]
            if i1 < 0:
                return a[:, i0]
            sel = slice(i0, i1)
            return np.trapezoid(a[:, sel] / x[sel], x=x[sel])
        tot = lambda m: sum(value(c) for c in (3, 0, 2, 1, 4, 5) if (m >> c) & 1)
        got = 10 ** (tot(num) / tot(den)) if kind else tot(num)
        assert np.allclose(got, want[:, s], rtol=1e-12, atol=0), (case, models[s])


def test_band_plan_rejects_what_the_reference_cannot_evaluate():
    grid = np.arange(10.5, 16.5, 0.001)
    with pytest.raises(ValueError):
        R.band_plan(['agn_rad'], [[1e9, 1e10]], grid, radio=False)
    with pytest.raises(ValueError):
        R.band_plan(['AGNfracTO'], [[1e13, 1e13]], grid, radio=False)          # monochromatic: nothing appended
    with pytest.raises(ValueError):
        R.band_plan(['AGNfrac400-1600GHz'], [[1.6e12, 4e11]], grid, radio=True)   # unbound Lnurad_int with RADIO
    with pytest.raises(ValueError):
        R.band_plan(['nonsense'], [[1e13, 1e12]], grid, radio=False)
    bands, cols = R.band_plan(['AGNfrac400-1600GHz', 'sb'], [[1.6e12, 4e11], [1.6e12, 4e11]], grid, radio=False)
    assert len(bands) == 1 and cols[0][0] == 1 and cols[1][0] == 0


def test_family_mode_follows_get_fluxes_branches():
    for case, want in (("classic18", {"GA": 0, "SB": 0, "TO": 0, "BB": 2}),
                       ("freeebv_xray49", {"GA": 1, "BB": 3}), ("xray_gamma49", {"BB": 3}),
                       ("radio1_kd18_49", {"BB": 2, "RAD": 0}), ("radio2_sn12_49", {"BB": 2})):
        ms, ps, g = load_sed(case)
        k = 0
        for fam in ms.families:
            mode, e_col, a_col = R.family_mode(fam, ms.idxs[k] if fam.fitted else 0)
            if fam.fitted:
                k += 1
            if fam.name in want:
                assert mode == want[fam.name], (case, fam.name, mode)
            if mode:
                assert ms.par_names[e_col] in ('EBVgal', 'EBVbbb')
            if mode == 3:
                assert ms.par_names[a_col] == 'alphaScat'
