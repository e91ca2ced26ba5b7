"""
Host-side, walker-independent arithmetic of the log-posterior path.

Everything here depends only on (filter set, z, catalogue row) and is evaluated
once per source / z-slice with NumPy on the host, in the reference's operation
order, so that masks and band selectors are bit-identical by construction
(SURVEY.md §7 step 2, Appendix A "per-source constants").  The kernels receive
the results as plain arrays.

Reference functions mirrored (file:line relative to the AGNfitter tree):
  prevot_k           MODEL_AGNfitter.py:1178-1214  (BBBred_Prevot, the k(lambda) part)
  calzetti_k         MODEL_AGNfitter.py:1217-1255  (GALAXYred_Calzetti, the k(lambda) part)
  likelihood_mask    PARAMETERSPACE_AGNfitter.py:144-147 (precedence quirk kept)
  band_1500          PRIORS_AGNfitter.py:166-167
  agnfrac_branch     PRIORS_AGNfitter.py:163-164,169-172,182
  irsyn_constants    PRIORS_AGNfitter.py:224-252
  trapz_weights      PRIORS_AGNfitter.py:89-91 (scipy trapz over band centres)
  maximal_age        PRIORS_AGNfitter.py:480-509
  gaussian_norm      PRIORS_AGNfitter.py:428-430
"""
from math import pi, sqrt

import numpy as np


def prevot_k(rest_lognu):
    """k(lambda) of the SMC law at rest-frame log10(nu); 0 at/above 10^16.685 Hz."""
    x = np.asarray(rest_lognu, dtype=np.float64)
    redd_x = 2.998 * 1e10 / (10 ** (x) * 1e-8)
    redd_x = redd_x[::-1]
    wl_200eV = 2.998 * 1e10 / (10 ** (16.685) * 1e-8)
    k = np.zeros(len(redd_x))
    w0 = redd_x > wl_200eV
    with np.errstate(all="ignore"):
        k[w0] = 1.39 * pow((pow(10., -4.) * redd_x[w0]), -1.2) - 0.38
    return k[::-1]


def calzetti_k(rest_nu):
    """k(lambda) of the Calzetti law evaluated on an array of rest-frame nu [Hz]
    (ascending).  The lambda <= 0.12 um branch extrapolates from the array's own
    points nearest to 0.12 / 0.125 um and adds R_V a second time, as the
    reference does."""
    RV = 4.05
    c = 2.998 * 1e8
    nu = np.asarray(rest_nu, dtype=np.float64)
    with np.errstate(all="ignore"):
        wl = (c / nu * 1e6)[::-1]
        k = np.zeros(len(wl))
        w0 = wl <= 0.12
        w1 = wl < 0.63
        w2 = wl >= 0.63
        x1 = np.argmin(np.abs(wl - 0.12))
        x2 = np.argmin(np.abs(wl - 0.125))
        k[w2] = 2.659 * (-1.857 + 1.040 / wl[w2]) + RV
        k[w1] = 2.659 * (-2.156 + (1.509 / wl[w1]) - (0.198 / wl[w1] ** 2) + (0.011 / wl[w1] ** 3)) + RV
        if (wl[x1] - wl[x2]) != 0:
            k[w0] = k[x1] + ((wl[w0] - 0.12) * (k[x1] - k[x2]) / (wl[x1] - wl[x2])) + RV
        else:
            k[w0] = 0
    return k[::-1]


def likelihood_mask(nus, fluxes, z, xrays_setting):
    """Boolean validity mask of ln_likelihood.  NB the second expression parses as
    a | (b & c): the flux test only guards the X-ray side."""
    x = np.asarray(nus)
    y = np.asarray(fluxes)
    if xrays_setting == 'Prior':
        return (x < np.log10(10 ** (15.38) / (1 + z))) & (y > -99.e-23)
    return (x < np.log10(10 ** (15.38) / (1 + z))) | (x > np.log10(10 ** (16.685) / (1 + z))) & (y > -99.e-23)


def band_1500(bands, z):
    """Index of the LAST dictionary band with 14.3 < bands+log10(1+z) < 15.3, or -1."""
    r = np.asarray(bands) + np.log10(1 + z)
    sel = np.nonzero((14.3 < r) & (r < 15.3))[0]
    return int(sel[-1]) if len(sel) else -1


def agnfrac_branch(nus, fluxes, z, lumfactor):
    """0 -> 'faint' branch Gaussian(-2,2); 1 -> 'bright' branch {-9999 | Gaussian(2,2)}.
    Raises IndexError exactly where the reference would (no positive flux in the window)."""
    nus = np.asarray(nus)
    fluxes = np.asarray(fluxes)
    r = nus + np.log10(1 + z)
    d = fluxes[(14.3 < r) & (r < 15.3)]
    d = d[d > 0][-1]
    abs_mag_data = 51.6 - 2.5 * np.log10(lumfactor * d)
    characteristic_mag = -35.4 * (1 + z) ** 0.524 / (1 + (1 + z) ** 0.678)
    return 0 if abs_mag_data > (characteristic_mag - 1.) else 1


def irsyn_constants(nus, fluxes, bands, z):
    """(mode, j_IR): mode 0 -> prior contributes 0; +1 -> Gaussian(+2,2); -1 -> Gaussian(-2,2).
    mode 2 marks the reference's unbound-variable case (ratio == 2 exactly)."""
    nus = np.asarray(nus)
    fluxes = np.asarray(fluxes)
    lz = np.log10(1 + z)
    rad = fluxes[((8 - lz) < nus) & (nus < (10 - lz))]
    if True not in (rad > 0):
        return 0, 0
    rad = rad[rad > 0][-1]
    nu_rad = nus[fluxes == rad][0]
    ir = fluxes[((12.2 - lz) < nus) & (nus < (12.8 - lz))]
    if True not in (ir > 0):
        return 0, 0
    ir = max(ir[ir > 0])
    nu_ir = nus[fluxes == ir][0]
    syn_exp = rad * ((10 ** nu_ir / 10 ** nu_rad) ** (-0.75))
    j = int(np.argmin(np.abs(np.asarray(bands) - nu_ir)))
    ratio = ir / syn_exp
    if ratio < 2:
        return 1, j
    if ratio > 2:
        return -1, j
    return 2, j


def trapz_weights(x):
    """w with sum(w*f) == trapz(f, x) up to summation order."""
    x = np.asarray(x, dtype=np.float64)
    w = np.zeros_like(x)
    if len(x) > 1:
        d = np.diff(x)
        w[:-1] += d / 2.0
        w[1:] += d / 2.0
    return w


def maximal_age(z):
    """Age-of-the-Universe bound.  The integrand is constant in the integration
    variable (closure over the outer a, E) — kept as is."""
    import scipy.integrate
    z = np.double(z)
    O_m = 0.266
    O_r = 0.
    O_k = 0.
    O_L = 1. - O_m
    H_0 = 70
    H_sec = H_0 / 3.0857e19
    secondsinyear = 31556926
    a = 1 / (1 + z)
    E = O_m * (1 + z) ** 3 + O_r * (1 + z) ** 4 + O_k * (1 + z) + O_L
    integral, _ = scipy.integrate.quad(lambda zz: a / sqrt(E), z, 1089)
    return (integral * (1 / H_sec) / secondsinyear) - 1e9


def gaussian_norm(sigma):
    """log(1/(sqrt(2 pi) sigma)) — the walker-independent part of Gaussian_prior."""
    return float(np.log(1.0 / (np.sqrt(2 * np.pi) * sigma)))


def lumfactor(dlum):
    return 4. * pi * dlum ** 2.
