"""CPU: catalogue ingest (agnfitter_b200/ingest.py) against the rows the reference path produced for its
two catalogues (stored in the K1 golden files) and against a row-by-row restatement of DATA_all.PROPS."""
import os

import numpy as np
import pytest

from helpers import load_k1
from agnfitter_b200 import ingest

REF = "/root/reference"


def _props_rowwise(flux, fluxerr, nus0, z, unit, err10):
    """DATA_AGNfitter.py:141-193, one catalogue row at a time as the reference does it."""
    out = []
    for j in range(flux.shape[0]):
        fluxes0, fluxerrs0 = flux[j] * unit, fluxerr[j] * unit
        nd = np.ones(fluxes0.shape)
        nd[fluxerr[j] <= -99] = 0.
        sel = (fluxerr[j] <= -99) & (flux[j] > -99)
        fluxes0[sel] = fluxes0[sel] * 0.5
        fluxerrs0[sel] = fluxes0[sel]
        if err10:
            fluxerrs0[nd != 0] = np.sqrt(fluxerrs0[nd != 0] ** 2 + (fluxes0[nd != 0] * 0.1) ** 2)
        o = nus0.argsort()
        nus = nus0[o]
        rad = fluxes0[o][(nus < (10.5 - np.log10(1 + z[j]))) & (nd[o] > 0)]
        xr = fluxes0[o][(nus > (16.685 - np.log10(1 + z[j]))) & (nd[o] > 0)]
        out.append((nus, fluxes0[o], fluxerrs0[o], nd[o], len(rad), len(xr)))
    return out


@pytest.mark.parametrize("err10", [False, True])
def test_props_matches_rowwise_restatement(err10):
    rng = np.random.default_rng(4)
    S, n_b = 40, 23
    nus0 = rng.permutation(np.linspace(9.0, 18.2, n_b))
    z = np.round(rng.uniform(0.01, 5, S), 3)
    flux = rng.lognormal(0, 2, size=(S, n_b))
    err = 0.1 * flux
    u = rng.uniform(size=(S, n_b))
    err[u < 0.15] = -99.                       # upper limits
    flux[u < 0.05] = -99.                      # ... of which missing data
    cat = ingest.props(flux, err, nus0, z, 1e-26, err10=err10, dlum=np.ones(S))
    for j, (nus, f, e, nd, nr, nx) in enumerate(_props_rowwise(flux.copy(), err.copy(), nus0, z, 1e-26, err10)):
        assert np.array_equal(cat.nus[j], nus) and np.array_equal(cat.fluxes[j], f)
        assert np.array_equal(cat.fluxerrs[j], e) and np.array_equal(cat.ndflag[j], nd)
        assert cat.nRADdata[j] == nr and cat.nXRaysdata[j] == nx


def test_read_ascii_roundtrip(tmp_path):
    p = tmp_path / "cat.txt"
    p.write_text("#name z w1 f1 e1 w2 f2 e2 w3 f3 e3\n"
                 "A 0.5 0.15 1.0 0.1 0.55 2.0 -99 24.0 -99 -99\n"
                 "B 1.25 0.15 3.0 0.3 0.55 4.0 0.4 24.0 5.0 0.5\n")
    cat = dict(name=0, redshift=1, flux_list=[3, 6, 9], fluxerr_list=[4, 7, 10], flux_unit_cgs=1e-26)
    cat.update({'freq/wl_list': [2, 5, 8], 'freq/wl_format': 'wavelength', 'freq/wl_unit_si': 1e-6})
    c = ingest.read_ascii(str(p), cat)
    assert c.names == ["A", "B"] and np.array_equal(c.z, [0.5, 1.25]) and c.nus.shape == (2, 3)
    assert np.all(np.diff(c.nus, axis=1) > 0)
    # row A: band 0.55 um is an upper limit (flux 2 -> 1, err 1), band 24 um is missing (-99 stays)
    assert np.allclose(c.fluxes[0], [-99e-26, 1e-26, 1e-26]) and np.allclose(c.fluxerrs[0], [-99e-26, 1e-26, 0.1e-26])
    assert np.array_equal(c.ndflag[0], [0, 0, 1]) and c.dlum.shape == (2,) and np.all(c.dlum > 1e27)


@pytest.mark.needs_reference
@pytest.mark.parametrize("name", ["catalog_example", "brown2018_test"])
def test_ingest_reproduces_the_reference_rows(name):
    """The golden K1 files hold the DATA arrays the reference path produced for its own catalogues."""
    ms, g = load_k1("cat49_" + name)
    n_b = ms.n_b
    cat = dict(name=0, redshift=1, flux_list=list(range(3, 3 * n_b + 1, 3)), fluxerr_list=list(range(4, 3 * n_b + 2, 3)),
               flux_unit_cgs=1e-26)
    # central frequencies in catalogue column order, as DATA_all.PROPS finds them (DAT:59-96): looked up by filter
    # name in ALL_FILTERS_info.dat; column 0 of the catalogue is the last filter of the example set's list
    from oracle import loaders
    info = os.path.join(REF, "models/FILTERS/ALL_FILTERS_info.dat")
    info_names = list(np.loadtxt(info, delimiter="|", usecols=[1], skiprows=1, dtype=str))
    info_nu = np.loadtxt(info, delimiter="|", usecols=[3], skiprows=1)
    central = np.array([info_nu[info_names.index(nm)] for nm in reversed(loaders.FILTERSET_EXAMPLE49)])
    c = ingest.read_ascii(os.path.join(REF, "data", name + ".txt"), cat, central_nus=central)
    assert np.array_equal(c.z, g["src_z"])
    assert np.array_equal(c.nus, g["src_nus"])
    assert np.array_equal(c.fluxes, g["src_fluxes"]) and np.array_equal(c.fluxerrs, g["src_fluxerrs"])
    assert np.allclose(c.dlum, g["src_dlum"], rtol=1e-12)
