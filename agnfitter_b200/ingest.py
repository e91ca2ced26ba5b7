"""
Catalogue ingest for whole-catalogue runs: the arithmetic of DATA_all.PROPS
(DATA_AGNfitter.py:47-200, ASCII branch) for all rows at once, producing exactly the per-source arrays the
log-posterior path consumes (DATA_AGNfitter.py:357-369: nus, fluxes, fluxerrs, ndflag, z, dlum, nRADdata,
nXRaysdata).  The reference loops over catalogue rows with pandas / astropy objects; here every step is one
NumPy expression over [S, n_b] — same operations, same order, so the arrays are bit-identical
(tests/test_ingest_cpu.py compares them with the rows the reference path produced for its two catalogues).

What each step mirrors:
  unit scaling                 flux * flux_unit -> erg s^-1 cm^-2 Hz^-1           DAT:103-105,141-142
  upper limits                 err <= -99 & flux > -99: flux *= 0.5; err = flux   DAT:161-170 (and 147-156)
  ndflag                       0 where err <= -99                                 DAT:151,165
  optional +10 % error         err = sqrt(err^2 + (0.1 flux)^2) where ndflag != 0 DAT:174-176
  sort by frequency            nus0.argsort() per row                             DAT:180-183
  nRADdata / nXRaysdata        valid data below 10.5 / above 16.685 - log10(1+z)  DAT:186-193
  luminosity distance          FlatLambdaCDM(H0=67.4, Om0=0.315, Tcmb0=2.725)     MODEL_AGNfitter.py:1285-1290
"""
from dataclasses import dataclass

import numpy as np

from .synth import z2Dlum

C_LIGHT = 299792458.0      # astropy.constants.c, used by u.spectral() (DAT:139)


@dataclass
class Catalog:
    names: list
    z: np.ndarray          # [S]
    dlum: np.ndarray       # [S] cm
    nus: np.ndarray        # [S, n_b] log10 Hz, ascending per row
    fluxes: np.ndarray     # [S, n_b] erg s^-1 cm^-2 Hz^-1
    fluxerrs: np.ndarray   # [S, n_b]
    ndflag: np.ndarray     # [S, n_b]
    nRADdata: np.ndarray   # [S]
    nXRaysdata: np.ndarray  # [S]

    @property
    def S(self):
        return len(self.z)


def props(flux, fluxerr, nus0, z, flux_unit_cgs, ndflag=None, err10=False, names=None, dlum=None):
    """flux, fluxerr [S, n_b] in catalogue units; nus0 [n_b] or [S, n_b] log10 Hz in column order; z [S].
    flux_unit_cgs: value of the catalogue's flux unit in erg s^-1 cm^-2 Hz^-1 (mJy -> 1e-26).
    ndflag [S, n_b] when the catalogue has flag columns (cat['ndflag_bool']), else None."""
    flux = np.atleast_2d(np.asarray(flux, dtype=np.float64))
    fluxerr = np.atleast_2d(np.asarray(fluxerr, dtype=np.float64))
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))
    S, n_b = flux.shape
    nus0 = np.asarray(nus0, dtype=np.float64)
    if nus0.ndim == 1:
        nus0 = np.broadcast_to(nus0, (S, n_b))
    fluxes0 = flux * flux_unit_cgs
    fluxerrs0 = fluxerr * flux_unit_cgs
    nd = np.ones((S, n_b)) if ndflag is None else np.array(ndflag, dtype=np.float64)
    no_err = fluxerr <= -99
    ul = no_err & (flux > -99)
    nd[no_err] = 0.
    fluxes0[ul] = fluxes0[ul] * 0.5
    fluxerrs0[ul] = fluxes0[ul]
    if err10:
        m = nd != 0
        fluxerrs0[m] = np.sqrt(fluxerrs0[m] ** 2 + (fluxes0[m] * 0.1) ** 2)
    order = np.argsort(nus0, axis=1)                      # ndarray.argsort() default (quicksort) per row, DAT:180
    rows = np.arange(S)[:, None]
    nus = nus0[rows, order]
    fluxes0, fluxerrs0, nd = fluxes0[rows, order], fluxerrs0[rows, order], nd[rows, order]
    lz = np.log10(1 + z)[:, None]
    nrad = ((nus < (10.5 - lz)) & (nd > 0)).sum(axis=1)
    nxr = ((nus > (16.685 - lz)) & (nd > 0)).sum(axis=1)
    if dlum is None:
        dlum = z2Dlum(z)
    return Catalog(names=list(names) if names is not None else ["%d" % i for i in range(S)], z=z,
                   dlum=np.atleast_1d(np.asarray(dlum, dtype=np.float64)), nus=np.ascontiguousarray(nus),
                   fluxes=np.ascontiguousarray(fluxes0), fluxerrs=np.ascontiguousarray(fluxerrs0), ndflag=nd,
                   nRADdata=nrad, nXRaysdata=nxr)


def read_ascii(path, cat, central_nus=None):
    """Read a whitespace-separated catalogue with the keys of CATALOG_settings() (example/SETTINGS_AGNfitter.py:
    17-89): 'name', 'redshift', 'flux_list', 'fluxerr_list', optional 'ndflag_bool'/'ndflag_list',
    'err+10%flux_moreflex', and either
      * central_nus [n_b]: log10 Hz of the chosen filters in catalogue column order
        (cat['use_central_wavelength'] = True: the reference takes them from ALL_FILTERS_info.dat, DAT:59-96), or
      * 'freq/wl_list' + 'freq/wl_format' ('wavelength' | 'frequency') + 'freq/wl_unit_si' (metres or Hz per unit).
    'flux_unit_cgs' is the flux unit in erg s^-1 cm^-2 Hz^-1 (the reference's astropy unit u.Jy*1e-3 -> 1e-26)."""
    names, rows = [], []
    with open(path) as f:
        for line in f:
            if not line.strip() or line.lstrip().startswith("#"):
                continue
            tok = line.split()
            names.append(tok[cat['name']])
            rows.append(tok)
    z = np.array([float(r[cat['redshift']]) for r in rows])
    col = lambda idx: np.array([[float(r[c]) for c in idx] for r in rows])
    flux, err = col(cat['flux_list']), col(cat['fluxerr_list'])
    nd = col(cat['ndflag_list']) if cat.get('ndflag_bool') else None
    if central_nus is not None:
        nus0 = np.asarray(central_nus, dtype=np.float64)
    else:
        x = col(cat['freq/wl_list']) * cat['freq/wl_unit_si']
        nus0 = np.log10(x) if cat['freq/wl_format'] == 'frequency' else np.log10(C_LIGHT / x)
    return props(flux, err, nus0, z, cat['flux_unit_cgs'], ndflag=nd, err10=bool(cat.get('err+10%flux_moreflex')),
                 names=names)
