"""
TEST INFRASTRUCTURE — not part of the product path.

Restated loaders for the reference's input files that cannot be read through
the reference's own code in this container (astropy is absent):

  * models/FILTERS/ALL_FILTERS  -> filter curves + FILTER_SET arrays
        follows FILTERS_AGNfitter.py:82-122 (selection, sorted central_nu array,
        lambdas_dict/factors_dict keyed by central_nu, one-element lists)
  * models/GALAXY/BC03_275seds.pickle through the 'BC03' branch
        follows MODEL_AGNfitter.py:71-142 (grid EBVgal) and the 'free' EBVgal
        variant MODEL_AGNfitter.py:116-140 / 193-219
  * models/STARBURST/s17_*.fits (hand-parsed FITS bintables)
        follows MODEL_AGNfitter.py:333-372
  * catalogue rows  -> per-source data arrays
        follows DATA_AGNfitter.py:54-200 (ASCII branch) and MODEL:1285-1290 (z2Dlum)

The remaining families are loaded by calling the reference's own loaders under
oracle/ref_shim.py (R06, S04, DH02_CE01, AGN_RAD, SN12, NK0 are numpy-only).
Everything returns the same container shapes the reference produces:
    {key-tuple-or-str: (log10 nu [n], Fnu [n])}.
"""
import os
import pickle
import itertools
from collections import defaultdict

import numpy as np

from . import ref_shim

REF_ROOT = ref_shim.REF_ROOT


# --------------------------------------------------------------------------
# tolerant unpickling of astropy-flavoured pickles
# --------------------------------------------------------------------------
class _QuantityStub(np.ndarray):
    """ndarray subclass standing in for astropy.units.quantity.Quantity.
    Its pickled state is (ndarray_state, {'_unit': ...}); the raw buffer arrives
    as a latin-1 str under encoding='latin1'."""

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            nd_state, extra = state
        else:
            nd_state, extra = state, {}
        nd_state = list(nd_state)
        if isinstance(nd_state[-1], str):
            nd_state[-1] = nd_state[-1].encode("latin1")
        np.ndarray.__setstate__(self, tuple(nd_state))
        self._unit_stub = extra.get("_unit")

    @property
    def value(self):
        return np.asarray(self)


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self._state = state

    def __call__(self, *a, **k):
        return _Anything()


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("astropy"):
            if name == "Quantity":
                return _QuantityStub
            if name[0].islower() or name.startswith("_"):
                return lambda *a, **k: _Anything()
            return type(name, (_Anything,), {})
        if module == "functions.FILTERS_AGNfitter" or name == "FILTER":
            return type("FILTER", (), {})
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        return super().find_class(module, name)


def tolerant_load(path):
    with open(path, "rb") as f:
        return _TolerantUnpickler(f, encoding="latin1").load()


# --------------------------------------------------------------------------
# filters
# --------------------------------------------------------------------------
FILTERSET_DEFAULT = ['SPIRE500', 'SPIRE350', 'SPIRE250', 'MIPS24', 'IRAC4',
                     'IRAC3', 'IRAC2', 'IRAC1', 'Ks_2mass', 'H_2mass', 'J_2mass',
                     'Y_VISTA', 'z_SUBARU', 'i_SUBARU', 'r_SUBARU', 'B_SUBARU', 'u_CHFT',
                     'GALEX_2500']     # FILTERS_AGNfitter.py:86-89

# names enabled in example/SETTINGS_AGNfitter.py:106-154 (49 bands)
FILTERSET_EXAMPLE49 = ['4XMM_EP5', '4XMM_EP4', '4XMM_EP3', '4XMM_EP2', 'RAD_1_4GHz', 'RAD_5GHz',
                       'RAD_15GHz', 'SPIRE500', 'SPIRE350', 'SPIRE250', 'PACS160', 'PACS100', 'PACS70',
                       'M1', 'WISE4', 'WISE3', 'IRAC4', 'IRAC3', 'WISE2', 'IRAC2', 'IRAC1', 'WISE1',
                       'Ks_2mass', 'H_2mass', 'J_2mass', 'y_PS1_good', 'SKY_z', 'z', 'z_PS1', 'SKY_i',
                       'i_PS1', 'i', 'SKY_r', 'r_PS1', 'r', 'SWIFT_V', 'SKY_g', 'g_PS1', 'g', 'SWIFT_B',
                       'SKY_v', 'u', 'SKY_u', 'SWIFT_U', 'SWIFT_W1', 'GALEX_2500', 'SWIFT_M2', 'SWIFT_W2',
                       'GALEX_1500']


def load_all_filters(root=REF_ROOT):
    """{filtername: object with .lambdas [A], .factors, .central_nu (log Hz)} in file order."""
    return tolerant_load(os.path.join(root, "models/FILTERS/ALL_FILTERS"))


def make_filterdict(names, all_filters=None, root=REF_ROOT):
    """Rebuild `[central_nu_array, lambdas_dict, factors_dict]` the way FILTER_SET does
    (FILTERS_AGNfitter.py:109-122): objects chosen in ALL_FILTERS iteration order,
    dict values are one-element lists so np.array(...) has shape (1, n_k) (DICTIONARIES:462-463)."""
    if all_filters is None:
        all_filters = load_all_filters(root)
    chosen = [o for o in all_filters.values() if o.filtername in names]
    lambdas_dict = defaultdict(list)
    factors_dict = defaultdict(list)
    central = []
    for o in chosen:
        lambdas_dict[o.central_nu].append(np.asarray(o.lambdas, dtype=np.float64))
        factors_dict[o.central_nu].append(np.asarray(o.factors, dtype=np.float64))
        central.append(o.central_nu)
    return [np.array(sorted(central)), lambdas_dict, factors_dict], [o.filtername for o in chosen]


def data_nus_for(names, root=REF_ROOT):
    """Catalogue-side band frequencies: ALL_FILTERS_info.dat column 3 (4 decimals),
    DATA_AGNfitter.py:65-96 with use_central_wavelength=True; returned sorted ascending."""
    info = os.path.join(root, "models/FILTERS/ALL_FILTERS_info.dat")
    fnames = np.loadtxt(info, delimiter="|", usecols=[1], skiprows=1, dtype=str)
    cnu = np.loadtxt(info, delimiter="|", usecols=[3], skiprows=1)
    out = [cnu[list(fnames).index(n)] for n in names]
    return np.sort(np.array(out))


# --------------------------------------------------------------------------
# GALAXY from BC03_275seds through the 'BC03' code path
# --------------------------------------------------------------------------
_C_ASTROPY = 299792458.0  # m/s, astropy.constants.c used by u.spectral()


def galaxy_BC03(root=REF_ROOT, ebv_type="grid", reddening_fn=None, ebv_free_pair=(0.0, 1.0),
                age_slice=None, tau_slice=None, metal=None):
    """MODEL_AGNfitter.py:71-142 restated for BC03_275seds.pickle (same dict schema as the
    missing BC03_840seds).  `reddening_fn(nu, Fnu, ebv) -> (nu, Fnu_red)` must be the
    reference's GALAXYred_Calzetti (passed in by the caller, from the shimmed MOD module).
    Returns the same six objects GALAXY() returns.
    metal: a metallicity value -> the key layout of the 'BC03_metal' branch the example settings name
    (MODEL_AGNfitter.py:145-232: keys (metal, tau, age, EBVgal), names ['metal','tau','age','EBVgal'], the
    free-E(B-V) pair (0., 0.8)); its library blob is absent from the mount, so the 275-SED library stands in
    with a metallicity axis of size 1."""
    from scipy.integrate import trapezoid
    d = tolerant_load(os.path.join(root, "models/GALAXY/BC03_275seds.pickle"))
    tau_vals = np.asarray(d["tau-values"])
    age_vals = np.asarray(d["age-values"])
    sed = np.asarray(d["SED"])          # (1, n_age, n_tau, 1, 1, n_wl)
    sfr = np.asarray(d["SFR"])
    wl = np.asarray(d["wavelength"])     # Angstrom, increasing
    _, n_age, n_tau, _, _, _ = sed.shape
    ages = range(n_age) if age_slice is None else age_slice
    taus = range(n_tau) if tau_slice is None else tau_slice

    if ebv_type == "grid":
        ebv = np.array(np.arange(0., 100., 2.5) / 100)
    else:
        ebv = np.array(ebv_free_pair if metal is None else (0., 0.8))
    F4plot, SFRd, attd = dict(), dict(), dict()
    # u.spectral(): nu = c / lambda  (lambda in Angstrom -> m)
    nus_all = (_C_ASTROPY / (wl * 1e-10))[::-1]
    for agei, taui, ebvi in itertools.product(ages, taus, range(len(ebv))):
        Fwl = sed[:, agei, taui, :, :, :].squeeze()
        Fnu = (Fwl * 3.34e-19 * wl ** 2.)[::-1]
        nus3, Fnu3 = nus_all[0:len(nus_all):3], Fnu[0:len(nus_all):3]
        e_apply = ebv[ebvi] if ebv_type == "grid" else ebv[0]
        gal_nu, Fnu_red = reddening_fn(nus3, Fnu3, e_apply)
        key = (str(tau_vals[taui]), str(np.log10(age_vals[agei])), str(ebv[ebvi]))
        if metal is not None:
            key = (str(metal),) + key
        F4plot[key] = (np.log10(gal_nu), Fnu_red / 1e18)           # renorm_template('GA'), MOD:1458-1460
        SFRd[key[:-1]] = sfr[:, agei, taui, :, :].squeeze()
        attd[key] = trapezoid(Fnu3, x=nus3) - trapezoid(Fnu_red, x=gal_nu)
    names = ['tau', 'age', 'EBVgal']
    types = ['grid', 'grid', ebv_type]
    if metal is not None:
        names, types = ['metal'] + names, ['grid'] + types
    return F4plot, SFRd, attd, names, types, [0]


# --------------------------------------------------------------------------
# STARBURST S17 from FITS bintables (no astropy)
# --------------------------------------------------------------------------
def _read_fits_bintable_row(path):
    """Minimal FITS BINTABLE reader for the single-row, big-endian float64 vector-column
    tables shipped in models/STARBURST (TFORMn = '<count>D', TDIMn gives the 2-D shape)."""
    with open(path, "rb") as f:
        raw = f.read()

    def parse_header(off):
        cards = {}
        while True:
            block = raw[off:off + 2880]
            off += 2880
            done = False
            for i in range(0, 2880, 80):
                card = block[i:i + 80].decode("ascii")
                key = card[:8].strip()
                if key == "END":
                    done = True
                    break
                if card[8:10] == "= ":
                    val = card[10:].split("/")[0].strip()
                    cards[key] = val.strip("'").strip() if val.startswith("'") else val
            if done:
                return cards, off

    hdr0, off = parse_header(0)
    naxis0 = int(hdr0.get("NAXIS", 0))
    size0 = 0
    if naxis0:
        size0 = abs(int(hdr0["BITPIX"])) // 8
        for i in range(naxis0):
            size0 *= int(hdr0["NAXIS%d" % (i + 1)])
    off += ((size0 + 2879) // 2880) * 2880
    hdr, off = parse_header(off)
    ncol = int(hdr["TFIELDS"])
    out = {}
    pos = off
    for c in range(1, ncol + 1):
        form = hdr["TFORM%d" % c]
        name = hdr["TTYPE%d" % c]
        code = form[-1]
        count = int(form[:-1]) if form[:-1] else 1
        dt = {"D": ">f8", "E": ">f4", "J": ">i4", "K": ">i8", "I": ">i2"}[code]
        arr = np.frombuffer(raw, dtype=dt, count=count, offset=pos).astype(np.float64)
        pos += count * np.dtype(dt).itemsize
        tdim = hdr.get("TDIM%d" % c)
        if tdim:
            dims = [int(x) for x in tdim.strip("()").split(",")]
            arr = arr.reshape(dims[::-1])
        out[name] = arr
    return out


def starburst_S17(root=REF_ROOT, tdust_slice=None, fracpah_slice=None):
    """MODEL_AGNfitter.py:333-372 restated (astropy Table.read replaced by a raw FITS parse;
    `(lam*u.micron).to(u.Hz, u.spectral())` = c/lambda with astropy's c)."""
    dust = _read_fits_bintable_row(os.path.join(root, "models/STARBURST/s17_lowvsg_dust.fits"))
    pah = _read_fits_bintable_row(os.path.join(root, "models/STARBURST/s17_lowvsg_pah.fits"))
    Dwl, DnuLnu = dust["LAM"], dust["SED"]
    Pwl, PnuLnu = pah["LAM"], pah["SED"]
    Tdust = np.array(dust["TDUST"])
    Lir = np.array(dust["LIR"]) * 3.826e33
    fracPAH = np.concatenate(((np.arange(0.0, 0.1, 0.01) / 100.), (np.arange(0.1, 5.5, 0.1) / 100.)))
    Dnu = _C_ASTROPY / (Dwl * 1e-6)
    Pnu = _C_ASTROPY / (Pwl * 1e-6)
    DLnu = DnuLnu / Dnu
    PLnu = PnuLnu / Pnu
    F4plot, LIRd = dict(), dict()
    ts = range(len(Tdust)) if tdust_slice is None else tdust_slice
    fps = range(len(fracPAH)) if fracpah_slice is None else fracpah_slice
    for t, fp in itertools.product(ts, fps):
        nu0 = np.array(Dnu[t, :])[::-1]
        F0 = np.array((1 - fracPAH[fp]) * DLnu[t, :] + (fracPAH[fp]) * PLnu[t, :])[::-1]
        key = (str(Tdust[t]), str(fracPAH[fp]))
        F4plot[key] = (np.log10(nu0), F0 / 1e20)                     # renorm_template('SB')
        LIRd[key] = Lir[t]
    return F4plot, LIRd, ['Tdust', 'fracPAH'], ['grid', 'grid'], []


# --------------------------------------------------------------------------
# cosmology + catalogue rows
# --------------------------------------------------------------------------
def z2Dlum(z):
    """MODEL_AGNfitter.py:1285-1290 restated: astropy FlatLambdaCDM(H0=67.4, Tcmb0=2.725, Om0=0.315)
    luminosity distance in cm.  Photons + 3.04 massless neutrino species are included the way
    astropy does (Ogamma0 from Tcmb0; Onu0 = 0.2271*Neff*Ogamma0; Ode0 = 1 - Om0 - Ogamma0 - Onu0)."""
    from scipy.integrate import quad
    H0 = 67.4
    Om0 = 0.315
    Tcmb0 = 2.725
    c_kms = 299792.458
    h = H0 / 100.0
    # critical density (g/cm^3) and photon density as in astropy.cosmology
    G = 6.6743e-8          # cm^3 g^-1 s^-2 (CODATA 2018)
    Mpc_cm = 3.0856775814913673e24
    H0_s = H0 * 1e5 / Mpc_cm
    rho_crit = 3.0 * H0_s ** 2 / (8.0 * np.pi * G)
    a_B_c2 = 4.0 * 5.670374419e-5 / (2.99792458e10) ** 3     # 4 sigma / c^3  (g cm^-3 K^-4)
    Ogamma0 = a_B_c2 * Tcmb0 ** 4 / rho_crit
    Onu0 = 0.22710731766 * 3.04 * Ogamma0
    Or0 = Ogamma0 + Onu0
    Ode0 = 1.0 - Om0 - Or0

    def inv_E(zz):
        zp = 1.0 + zz
        return 1.0 / np.sqrt(zp ** 3 * (Or0 * zp + Om0) + Ode0)

    zs = np.atleast_1d(np.asarray(z, dtype=np.float64))
    out = np.empty_like(zs)
    for i, zz in enumerate(zs):
        integ, _ = quad(inv_E, 0.0, zz, epsrel=1e-12, epsabs=0)
        out[i] = (1.0 + zz) * (c_kms / H0) * integ * Mpc_cm
    return out if np.ndim(z) else float(out[0])


def catalog_rows(flux, fluxerr, nus0, z, flux_unit_cgs=1e-26, err10=False):
    """DATA_AGNfitter.py:141-193 restated for the ASCII/no-ndflag-columns case
    (ndflag_bool False), vectorised over sources.

    flux, fluxerr : [S, n_b] in catalogue units (e.g. mJy -> flux_unit_cgs = 1e-26)
    nus0          : [n_b] log10 Hz in catalogue column order
    Returns dict(nus, fluxes, fluxerrs, ndflag, nRADdata, nXRaysdata) with bands sorted by nu.
    """
    flux = np.asarray(flux, dtype=np.float64)
    fluxerr = np.asarray(fluxerr, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    fluxes0 = flux * flux_unit_cgs
    fluxerrs0 = fluxerr * flux_unit_cgs
    ndflag = np.ones(fluxes0.shape)
    ul = (fluxerr <= -99) & (flux > -99)
    ndflag[fluxerr <= -99] = 0.
    fluxes0[ul] = fluxes0[ul] * 0.5                 # DAT:167-168
    fluxerrs0[ul] = fluxes0[ul]                     # DAT:169-170 (already halved)
    if err10:
        m = ndflag != 0
        fluxerrs0[m] = np.sqrt(fluxerrs0[m] ** 2 + (fluxes0[m] * 0.1) ** 2)
    order = np.asarray(nus0).argsort()
    nus = np.asarray(nus0)[order]
    fluxes0, fluxerrs0, ndflag = fluxes0[:, order], fluxerrs0[:, order], ndflag[:, order]
    lz = np.log10(1 + z)[:, None]
    nRAD = ((nus[None, :] < (10.5 - lz)) & (ndflag > 0)).sum(axis=1)
    nXR = ((nus[None, :] > (16.685 - lz)) & (ndflag > 0)).sum(axis=1)
    return dict(nus=np.broadcast_to(nus, fluxes0.shape).copy(), fluxes=fluxes0, fluxerrs=fluxerrs0,
                ndflag=ndflag, nRADdata=nRAD, nXRaysdata=nXR)


def read_ascii_catalog(path, n_bands=49):
    """Parse data/catalog_example.txt-style files: name, z, n_bands x (wl, flux, err)."""
    names, zs, fl, fe = [], [], [], []
    with open(path) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            tok = line.split()
            names.append(tok[0])
            zs.append(float(tok[1]))
            vals = np.array(tok[2:2 + 3 * n_bands], dtype=np.float64).reshape(n_bands, 3)
            fl.append(vals[:, 1])
            fe.append(vals[:, 2])
    return names, np.array(zs), np.array(fl), np.array(fe)
