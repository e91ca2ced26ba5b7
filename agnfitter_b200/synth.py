"""
Deterministic synthetic inputs of the shapes BASELINE.json names:
filters with the default 18-band set's band centres / sample counts / widths, template
families with the classic model set's grids (BC03-like 25 ages x 11 taus x 40 E(B-V) on 407
wavelengths, DH02_CE01-like 161 x 1366, S04-like 60 x 30, R06-like with free EBVbbb on 438
points), and a catalogue generator.  No file of the reference is read: the GPU box has none.
Shapes and grids are the real ones; spectra are smooth analytic stand-ins ("data": "synthetic").
"""
import time

import numpy as np

from . import hostmath
from .flatten import Family, ModelSet, family_from_keys_float

# default 18-band set (FILTERS_AGNfitter.py:86-89): log10 central nu, samples per curve,
# log10(lambda_max / lambda_min) of each curve — measured from models/FILTERS/ALL_FILTERS
DEFAULT18_NU = [11.774163, 11.932734, 12.079841, 13.108272, 13.585464, 13.720665, 13.825887, 13.927985,
                14.141412, 14.260066, 14.379951, 14.467483, 14.520668, 14.593551, 14.677685, 14.827378,
                14.894249, 15.114285]
DEFAULT18_NK = [512, 512, 512, 128, 421, 370, 427, 505, 200, 169, 5205, 116, 198, 218, 189, 190, 273, 23]
DEFAULT18_SPAN = [0.778, 0.778, 0.778, 0.253, 0.264, 0.173, 0.158, 0.148, 0.121, 0.134, 0.173, 0.097, 0.132,
                  0.136, 0.127, 0.182, 0.155, 0.251]

# the 49-band set of example/SETTINGS_AGNfitter.py (radio to X-ray), same three quantities per curve
EXAMPLE49_NU = [9.145694, 9.698535, 10.175657, 11.774163, 11.932734, 12.079841, 12.246828, 12.460938, 12.629178,
                13.105391, 13.129294, 13.385592, 13.585464, 13.720665, 13.811172, 13.825887, 13.927985, 13.947874,
                14.141412, 14.260066, 14.379951, 14.49291, 14.511204, 14.524307, 14.538192, 14.59011, 14.599052,
                14.601643, 14.677783, 14.683271, 14.685333, 14.74192, 14.783841, 14.789474, 14.802849, 14.837293,
                14.894014, 14.925019, 14.935101, 14.935668, 15.051964, 15.114285, 15.123116, 15.158924, 15.295975,
                17.264134, 17.562297, 17.891176, 18.220277]
EXAMPLE49_NK = [1000, 1000, 1000, 512, 512, 512, 3560, 1428, 405, 128, 2559, 2559, 421, 370, 548, 427, 505, 398, 200,
                169, 5205, 201, 24, 141, 153, 22, 167, 89, 22, 178, 75, 1282, 26, 172, 89, 1282, 31, 47, 35, 1282, 1282,
                23, 1282, 1282, 14, 1001, 1001, 1001, 1001]
EXAMPLE49_SPAN = [0.269, 0.269, 0.269, 0.778, 0.778, 0.778, 0.421, 0.292, 0.342, 0.253, 1.052, 1.052, 0.264, 0.173,
                  0.5, 0.158, 0.148, 0.41, 0.121, 0.134, 0.173, 0.087, 0.107, 0.162, 0.076, 0.118, 0.096, 0.128, 0.147,
                  0.125, 0.128, 0.7, 0.22, 0.157, 0.206, 0.7, 0.085, 0.142, 0.108, 0.7, 0.7, 0.251, 0.7, 0.7, 0.131,
                  0.301, 0.301, 0.352, 0.426]

CLASSIC_SETTINGS = dict(
    path='models/', modelset='modelsv1', GALAXY='BC03', STARBURST='DH02_CE01', BBB='R06', TORUS='S04',
    XRAYS=False, RADIO=False, Blazar='None', PRIOR_energy_balance='Flexible', PRIOR_AGNfraction=True,
    PRIOR_midIR_UV=False, PRIOR_galaxy_only=False, turn_on_AGN=True)

_C = 2.997e8


def synthetic_filterdict(nus=DEFAULT18_NU, nks=DEFAULT18_NK, spans=DEFAULT18_SPAN, decreasing=(0, 1, 2), which=None):
    """[central_nu_array, lambdas_dict, factors_dict] in the reference's container shapes
    (FILTERS_AGNfitter.py:111-122).  Curves listed in `decreasing` store lambda descending, like the
    SPIRE curves of ALL_FILTERS do.  which='example49' selects the 49-band example set's geometry."""
    if which == "example49":
        nus, nks, spans, decreasing = EXAMPLE49_NU, EXAMPLE49_NK, EXAMPLE49_SPAN, (3, 4, 5)
    lambdas_dict, factors_dict = {}, {}
    for i, (nu, nk, span) in enumerate(zip(nus, nks, spans)):
        lam_c = _C / 10 ** nu * 1e10
        lo, hi = lam_c * 10 ** (-span / 2), lam_c * 10 ** (span / 2)
        lam = np.linspace(lo, hi, nk)
        x = (lam - lam_c) / (0.25 * (hi - lo))
        T = np.exp(-0.5 * x ** 4) * (0.6 + 0.4 * np.cos(3.0 * x) ** 2)
        if i in decreasing:
            lam, T = lam[::-1].copy(), T[::-1].copy()
        lambdas_dict[nu] = [lam]
        factors_dict[nu] = [T]
    return [np.array(sorted(nus)), lambdas_dict, factors_dict]


def _bump(lognu, centre, width):
    return np.exp(-0.5 * ((lognu - centre) / width) ** 2)


C4_SETTINGS = dict(CLASSIC_SETTINGS, RADIO=True, XRAYS=True)


def synthetic_templates(seed=20260101, ga_free=False, c4=False):
    """The four classic families as `*_4plot`-style dicts {key: (log10 nu, Fnu)} with the real key
    strings' structure, plus GALAXYatt / STARBURST_LIR stand-ins.  ga_free: the galaxy family in the shape of the
    reference's default `BC03_metal` library instead — keys (metal, tau, age, EBVgal) with E(B-V)_gal a FREE
    parameter applied on the fly (two key values per template, MODEL_AGNfitter.py:145-301), 4 x 11 x 25 x 2 = 2 200
    unreddened templates, no GALAXYatt table (the energy-balance prior integrates on the fly, PRIORS:83-93).
    c4: the radio-to-X-ray model of BASELINE configs[3] — the disc family in the shape of R06 + X-ray extension
    (EBVbbb, alphaScat free, Gamma grid: 2 x 2 x 21 keys, 1 438 points, MODEL_AGNfitter.py:600-645) and a fitted
    synchrotron family in the shape AGN_RAD builds for > 3 radio data (alpha1 x alpha2 x nu_t = 10 x 10 x 30 =
    3 000 keys on 400 points, MODEL_AGNfitter.py:498-527): five components, 18 parameters."""
    rng = np.random.default_rng(seed)
    out = {}
    # GALAXY: keys (tau, age, EBVgal) enumerated age-major like MODEL_AGNfitter.py:95-108
    tau = np.array([0.1, 0.25, 0.5, 0.75, 1., 2.5, 5., 7.5, 10., 12.5, 15.])
    age = np.linspace(5.0, 10.0, 25)
    ebv = np.array(np.arange(0., 100., 2.5) / 100)
    lam = np.geomspace(91.0, 1.6e6, 407)                       # Angstrom
    nu = (2.99792458e18 / lam)[::-1]
    lognu = np.log10(nu)
    kcal = hostmath.calzetti_k(nu)
    ga, att = {}, {}
    for a in age:
        for t in tau:
            base = 1e-2 * (_bump(lognu, 14.2 + 0.08 * (10 - a), 0.45) + 0.3 * _bump(lognu, 14.9 - 0.03 * t, 0.25)
                           + 1e-3)
            for e in ebv:
                red = base * 10 ** (-0.4 * kcal * e)
                key = (str(t), str(a), str(e))
                ga[key] = (lognu, red * 1e-30)
                att[key] = float(np.trapezoid(base - red, x=nu)) * 1e-12
    out["GA"] = dict(F=ga, names=['tau', 'age', 'EBVgal'], types=['grid', 'grid', 'grid'], aux=att)
    if ga_free:
        ga = {}
        for m in (0.2, 0.4, 1.0, 2.5):
            for a in age:
                for t in tau:
                    base = 1e-2 * (_bump(lognu, 14.2 + 0.08 * (10 - a) + 0.02 * m, 0.45)
                                   + 0.3 * _bump(lognu, 14.9 - 0.03 * t, 0.25) + 1e-3)
                    for e in (0.0, 0.8):
                        ga[(str(m), str(t), str(a), str(e))] = (lognu, base * 1e-30)
        out["GA"] = dict(F=ga, names=['metal', 'tau', 'age', 'EBVgal'], types=['grid', 'grid', 'grid', 'free'], aux=None)
    # STARBURST: 161 keys, descending-then-ascending irlum like DH02_CE01's key order
    irl = np.round(np.concatenate([np.linspace(14.3404, 8.3091, 100), np.linspace(8.3613, 13.5477, 61)]), 4)
    assert len(set(irl.tolist())) == 161
    lognu_sb = np.linspace(11.435, 14.92, 1366)
    sb, lir = {}, {}
    for v in irl:
        sb[str(v)] = (lognu_sb, 1e-28 * 10 ** (0.3 * (v - 11)) * (_bump(lognu_sb, 12.4 + 0.03 * (v - 11), 0.35) + 0.02 * _bump(lognu_sb, 13.6, 0.3)))
        lir[str(v)] = float(pow(10, v) * 3.826e33 * 1e-6)
    out["SB"] = dict(F=sb, names=['irlum'], types=['grid'], aux=lir)
    # TORUS: Nh 21.5 .. 24.45 step 0.05
    nh = np.round(np.arange(21.5, 24.46, 0.05), 2)[:60]
    lognu_to = np.linspace(11.5, 16.3, 30)
    to = {str(v): (lognu_to, 1e-29 * _bump(lognu_to, 13.5 - 0.1 * (v - 23), 0.5) * 10 ** (-(v - 21.5) * _bump(lognu_to, 15., 0.6)))
          for v in nh}
    out["TO"] = dict(F=to, names=['Nh'], types=['grid'], aux=None)
    # BBB R06-like: two keys '0.0','1.0' holding the same unreddened template (MODEL:569-575)
    lognu_bb = 10.0 + 0.016 * np.arange(438)
    f = 1e-31 * 10 ** (0.3 * (lognu_bb - 10)) * np.exp(-10 ** (lognu_bb - 16.2))
    out["BB"] = dict(F={'0.0': (lognu_bb, f), '1.0': (lognu_bb, f)}, names=['EBVbbb'], types=['free'], aux=None)
    if c4:
        lognu_x = np.concatenate([lognu_bb, np.linspace(17.0, 19.0, 1000)])
        bb = {}
        for e in (0.0, 1.0):
            for a in (-0.4, 0.4):
                for g in np.arange(0.0, 3.15, 0.15):
                    fx = 1e-31 * 10 ** (0.3 * (lognu_x - 10)) * np.exp(-10 ** (lognu_x - 16.2))
                    hard = lognu_x >= 17.0
                    fx[hard] = 1e-33 * 10 ** ((1.0 - g) * (lognu_x[hard] - 17.0))
                    bb[(str(e), str(a), str(g))] = (lognu_x, fx)
        out["BB"] = dict(F=bb, names=['EBVbbb', 'alphaScat', 'Gamma'], types=['free', 'free', 'grid'], aux=None)
        lognu_r = np.arange(7, 15, 0.02)
        rad = {}
        for a1 in np.arange(-1.0, 1.0, 0.2):
            for a2 in np.arange(-1.0, 0, 0.1):
                for nt in np.arange(7, 13, 0.2):
                    x = lognu_r - nt
                    rad[(str(a1), str(a2), str(nt))] = (lognu_r, 1e-27 * 10 ** np.where(x < 0, a1 * x, a2 * x)
                                                        * np.exp(-10 ** (lognu_r - 13.0)))
        out["RAD"] = dict(F=rad, names=['alpha1', 'alpha2', 'nut'], types=['grid', 'grid', 'grid'], aux=None)
    return out


def z2Dlum(z):
    """Luminosity distance [cm] of FlatLambdaCDM(H0=67.4, Om0=0.315, Tcmb0=2.725) — what
    MODEL_AGNfitter.z2Dlum asks astropy for (MODEL_AGNfitter.py:1285-1290); photons and 3.04
    massless neutrino species included.  Host-side catalogue ingest, not on the GPU path."""
    from scipy.integrate import quad
    H0, Om0, Tcmb0 = 67.4, 0.315, 2.725
    Mpc_cm = 3.0856775814913673e24
    H0_s = H0 * 1e5 / Mpc_cm
    rho_crit = 3.0 * H0_s ** 2 / (8.0 * np.pi * 6.6743e-8)
    Ogamma0 = 4.0 * 5.670374419e-5 / (2.99792458e10) ** 3 * Tcmb0 ** 4 / rho_crit
    Or0 = Ogamma0 * (1.0 + 0.22710731766 * 3.04)
    Ode0 = 1.0 - Om0 - Or0
    zs = np.atleast_1d(np.asarray(z, dtype=np.float64))
    uz, inv = np.unique(zs, return_inverse=True)
    d = np.empty_like(uz)
    for i, zz in enumerate(uz):
        integ, _ = quad(lambda x: 1.0 / np.sqrt((1 + x) ** 3 * (Or0 * (1 + x) + Om0) + Ode0), 0.0, zz,
                        epsrel=1e-12, epsabs=0)
        d[i] = (1.0 + zz) * (299792.458 / H0) * integ * Mpc_cm
    out = d[inv]
    return out if np.ndim(z) else float(out[0])


def build_modelset(templates, filterdict, z_grid, builder, settings=CLASSIC_SETTINGS, on_device=True):
    """Run K2 (DictBuilder) over every family and assemble a ModelSet.  With on_device=True the
    band-flux tables stay in HBM: returns (ModelSet with table=None, {name: CUDA tensor})."""
    ms = ModelSet(settings=dict(settings), bands=np.asarray(filterdict[0], dtype=np.float64),
                  z_grid=np.asarray(z_grid, dtype=np.float64))
    dev = {}
    for name in ("GA", "SB", "TO", "BB", "RAD"):
        if name not in templates:
            continue
        fam = templates[name]
        keys = list(fam["F"].keys())
        arr = np.array(keys)
        kf = arr.T.astype(float)
        t0 = time.perf_counter()
        tab = builder.build_family(fam["F"], z_grid, on_device=on_device,      # synchronous (the ctx stream is drained)
                                   reuse_staging=True)
        builder.k2_wall_s = getattr(builder, "k2_wall_s", 0.0) + time.perf_counter() - t0
        builder.k2_contract_ms = getattr(builder, "k2_contract_ms", 0.0) + builder.contract_ms
        aux = None if fam["aux"] is None else np.array([fam["aux"][k] for k in keys])
        if name == "GA" and 'free' in fam["types"]:
            aux = None
        f = family_from_keys_float(name, fam["names"], fam["types"], kf, True,
                                   None if on_device else tab, aux)
        if on_device:
            dev[name] = tab
        ms.families.append(f)
    return ms, dev


def synthetic_catalog(ms, S, seed=20260101, z_lo=0.01, z_hi=6.0):
    """z ~ U(z_lo, z_hi) rounded to 3 decimals, truth ~ U(P.min, P.max).  Returns (z, dlum, truth);
    fluxes are drawn by `catalog_fluxes` once ymodel(truth) is known."""
    rng = np.random.default_rng(seed)
    z = np.round(rng.uniform(z_lo, z_hi, size=S), 3)
    return z, z2Dlum(z), rng


def catalog_fluxes(rng, ymodel_truth, nus, z, frac_ul=0.10, frac_missing=0.05):
    """y = ymodel(truth) (1 + 0.1 N), sigma = 0.1 |y|; 10 % of bands -> upper limit (sigma = -99),
    5 % -> missing (-99/-99); then the DATA transformation UL -> (UL/2, UL/2)
    (DATA_AGNfitter.py:152-170), in cgs.  One detection is kept in the 1500-A window that
    prior_AGNfraction reads (PRIORS_AGNfitter.py:163-164)."""
    S, n_b = ymodel_truth.shape
    y = np.abs(ymodel_truth) * (1 + 0.1 * rng.normal(size=(S, n_b)))
    y = np.abs(y)
    y = np.where(y > 1e-300, y, 1e-30)
    err = 0.1 * y
    u = rng.uniform(size=(S, n_b))
    ul = u < frac_ul
    miss = (u >= frac_ul) & (u < frac_ul + frac_missing)
    r = nus[None, :] + np.log10(1 + z)[:, None]
    win = (14.3 < r) & (r < 15.3)
    last = n_b - 1 - np.argmax(win[:, ::-1], axis=1)
    keep = np.zeros((S, n_b), dtype=bool)
    keep[np.arange(S), last] = win.any(axis=1)
    ul &= ~keep
    miss &= ~keep
    fl = y.copy()
    fe = err.copy()
    fl[ul] = y[ul] * 0.5
    fe[ul] = fl[ul]
    fl[miss] = -99. * 1e-26
    fe[miss] = -99. * 1e-26
    return fl, fe


def walker_positions(rng, pmin, pmax, W, mode="uniform", frac_out=0.01):
    """[S*W, n_par] positions, source-major.  mode 'uniform': U(P.min, P.max) (worst-case gather
    locality); 'ball': mid-box +- 1e-5 (PARAMETERSPACE:263).  frac_out of the rows are pushed out
    of the box on one coordinate."""
    S, n_par = pmin.shape
    if mode == "ball":
        mid = 0.5 * (pmin + pmax)
        p = mid[:, None, :] + (2 * rng.uniform(size=(S, W, n_par)) - 1) * 0.00001
    else:
        p = pmin[:, None, :] + rng.uniform(size=(S, W, n_par)) * (pmax - pmin)[:, None, :]
    p = p.reshape(S * W, n_par)
    n_out = int(frac_out * S * W)
    if n_out:
        rows = rng.choice(S * W, size=n_out, replace=False)
        cols = rng.integers(n_par, size=n_out)
        p[rows, cols] = pmax[rows // W, cols] + rng.uniform(0.01, 1.0, size=n_out)
    sid = np.repeat(np.arange(S, dtype=np.int32), W)
    return p, sid
