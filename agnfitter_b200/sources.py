"""
Per-source and per-z-slice operands of the batched log-posterior kernel.

The reference hands `ln_probab` a DATA object per source (DATA_AGNfitter.py:357-369:
nus, fluxes, fluxerrs, z, dlum, lumfactor) and a parameter table P from
PARAMETERSPACE_AGNfitter.Pdict (PARAMETERSPACE:31-102).  Here a whole catalogue becomes
struct-of-arrays blocks, with every walker-independent quantity evaluated once on the host
by agnfitter_b200.hostmath (reference arithmetic, reference operation order).
"""
from dataclasses import dataclass

import numpy as np

from . import hostmath
from .flatten import ModelSet

# consts block layout, per source (float64)
C_INV4PIDL2 = 0     # 1 / (4 pi dlum^2)                       PARAMETERSPACE:232, PRIORS:178
C_LUMFACTOR = 1     # 4. * pi * dlum**2.                       PRIORS:307,331
C_LOGF_XRAY = 2     # log10(F_2-10keV,data * lumfactor)       PRIORS:317-318 (prior_IR_XRays only)
C_LUMFACTOR_BB = 3  # 4*pi*(dlum)**2                           PRIORS:346
C_NCONST = 4


@dataclass
class ZSlices:
    cgal: np.ndarray      # [n_z, n_b]  -0.4 * k_Calzetti(band centre, rest frame)
    cbbb: np.ndarray      # [n_z, n_b]  -0.4 * k_Prevot(band centre, rest frame)
    xray: np.ndarray      # [n_z] uint64 bit b set: bands_b + log10(1+z) >= 16.685   DICTIONARIES:254-256
    j1500: np.ndarray     # [n_z] int32 band read by prior_AGNfraction                PRIORS:166-167
    omega: np.ndarray     # [n_z, n_b] trapezoid weights over observed band nu        PRIORS:89-91


def make_zslices(ms: ModelSet) -> ZSlices:
    n_z, n_b = ms.n_z, ms.n_b
    if n_b > 64:
        raise ValueError("at most 64 bands are supported (validity masks are 64-bit)")
    cgal = np.zeros((n_z, n_b))
    cbbb = np.zeros((n_z, n_b))
    xray = np.zeros(n_z, dtype=np.uint64)
    j1500 = np.zeros(n_z, dtype=np.int32)
    omega = np.zeros((n_z, n_b))
    for i, z in enumerate(ms.z_grid):
        rest = ms.bands + np.log10((1 + z))
        cbbb[i] = -0.4 * hostmath.prevot_k(rest)
        cgal[i] = -0.4 * hostmath.calzetti_k(10 ** rest)
        bits = 0
        for b in np.nonzero(rest >= 16.685)[0]:
            bits |= (1 << int(b))
        xray[i] = np.uint64(bits)
        j1500[i] = hostmath.band_1500(ms.bands, z)
        omega[i] = hostmath.trapz_weights((10 ** rest) / (1 + z))
    return ZSlices(cgal=cgal, cbbb=cbbb, xray=xray, j1500=j1500, omega=omega)


def extra_prior_plan(ms: ModelSet):
    """Which of prior_midIR_UV (PRIORS_AGNfitter.py:327-357, dispatch :50-54) and prior_IR_XRays (:302-324,
    dispatch :68-71) the settings enable, and their walker-independent operands: the band read for the 6 um
    torus flux (first band with 13.59897 < log nu < 13.79897 — dictionary band centres, observed frame), the
    band read for the 2500 A disc flux (first with 14.8 < log nu < 15.15) and the row of the unreddened twin of
    every disc template.  Raises what the reference raises when an operand does not exist."""
    midir_uv = ms.settings.get('PRIOR_midIR_UV') == True
    ir_xray = ms.settings.get('XRAYS') == 'Prior_midIR'
    plan = dict(midir_uv=midir_uv, ir_xray=ir_xray, j_6um=-1, j_2500=-1, bb_dered_row=None)
    if not (midir_uv or ir_xray):
        return plan
    to = ms.fam("TO")
    if to.n_dim != 1:
        raise TypeError("unhashable type: 'list' — TORUSFdict[tor_obj.matched_parkeys_grid] (PRIORS_AGNfitter.py:"
                        "304,329) only works for one-parameter torus families; this one has %d" % to.n_dim)
    b = ms.bands
    hit = np.nonzero((13.59897 < b) & (b < 13.79897))[0]
    if len(hit) == 0:
        raise IndexError("no band with 13.59897 < log nu < 13.79897 (PRIORS_AGNfitter.py:306,331)")
    plan["j_6um"] = int(hit[0])
    if midir_uv:
        hit = np.nonzero((14.8 < b) & (b < 15.15))[0]
        if len(hit) == 0:
            raise IndexError("no band with 14.8 < log nu < 15.15 (PRIORS_AGNfitter.py:347)")
        plan["j_2500"] = int(hit[0])
        bb = ms.fam("BB")
        rows = np.arange(bb.T, dtype=np.int32)
        if ms.settings['BBB'] in ('R06', 'THB21') and ms.settings['XRAYS'] != True:
            if bb.n_dim != 1:
                raise KeyError('0.0')
            zero = np.nonzero(bb.keys_float[0] == 0.0)[0]         # BBBFdict['0.0'], PRIORS:336
            if len(zero) == 0:
                raise KeyError('0.0')
            rows[:] = zero[0]
        else:                                                       # params[EBVbbb_pos] = str(0.0), PRIORS:339-342
            if bb.n_dim == 1:
                raise TypeError("'str' object does not support item assignment (PRIORS_AGNfitter.py:341)")
            d = bb.par_names.index('EBVbbb')
            lut = {tuple(bb.keys_float[:, t]): t for t in range(bb.T)}
            for t in range(bb.T):
                k = list(bb.keys_float[:, t])
                for dd in range(bb.n_dim):          # matched_parkeys_grid holds the FIRST key's value on free dims
                    if bb.par_types[dd] == 'free':
                        k[dd] = bb.keys_float[dd, 0]
                k[d] = 0.0
                if tuple(k) not in lut:
                    raise KeyError(tuple(k))
                rows[t] = lut[tuple(k)]
        plan["bb_dered_row"] = rows
    return plan


def pdict_bounds(ms: ModelSet, z):
    """P['min'], P['max'] for one source (PARAMETERSPACE:42-100)."""
    mins, maxs = [], []
    for f in ms.families:
        if not f.fitted:
            continue
        for d in range(f.n_dim):
            mins.append(float(np.min(f.keys_float[d])))
            maxs.append(float(np.max(f.keys_float[d])))
    agn_on = ms.settings['turn_on_AGN'] == True
    mins += [-10., -10., -10.]
    maxs += [10., 10., 10. if agn_on else -9.8]
    if ms.bbb_has_norm:
        mins.append(-10.)
        maxs.append(10. if agn_on else -9.8)
    if ms.radio:
        mins.append(-10.)
        maxs.append(10. if agn_on else -9.8)
    agemax = None
    for i, nm in enumerate(ms.par_names):
        if nm in ('age', 'tau'):
            if agemax is None:
                agemax = float(np.log10(hostmath.maximal_age(z)))
            maxs[i] = agemax
        if nm == 'Tdust':
            maxs[i] = 42.
        if nm == 'EBVgal':
            mins[i] = 0.05
    return np.array(mins, dtype=np.float64), np.array(maxs, dtype=np.float64)


@dataclass
class SourceSet:
    zi: np.ndarray          # [S] int32 z-slice of each source
    y: np.ndarray           # [S, n_b] fluxes
    isig: np.ndarray        # [S, n_b] 1/fluxerr
    mask: np.ndarray        # [S] uint64 likelihood validity bits
    pmin: np.ndarray        # [S, n_par]
    pmax: np.ndarray        # [S, n_par]
    consts: np.ndarray      # [S, C_NCONST]
    agn_branch: np.ndarray  # [S] int32   0 faint / 1 bright                 PRIORS:182
    irsyn_mode: np.ndarray  # [S] int32   0 off, +1, -1                        PRIORS:244-252
    irsyn_j: np.ndarray     # [S] int32

    @property
    def S(self):
        return len(self.zi)


def _bits(boolmat):
    """[S, n_b] bool -> [S] uint64 with bit b = column b."""
    n_b = boolmat.shape[1]
    w = (np.uint64(1) << np.arange(n_b, dtype=np.uint64))
    return (boolmat.astype(np.uint64) * w[None, :]).sum(axis=1, dtype=np.uint64)


def make_sources(ms: ModelSet, z, dlum, nus, fluxes, fluxerrs, bounds_cache=None) -> SourceSet:
    """z, dlum: [S]; nus/fluxes/fluxerrs: [S, n_b] (sorted by nu as DATA_all.PROPS leaves them).
    Every source's z must be present in ms.z_grid (exact float match).  Vectorised over sources;
    each expression is the per-source reference expression broadcast along axis 0."""
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))
    dlum = np.atleast_1d(np.asarray(dlum, dtype=np.float64))
    nus = np.atleast_2d(np.asarray(nus, dtype=np.float64))
    fluxes = np.atleast_2d(np.asarray(fluxes, dtype=np.float64))
    fluxerrs = np.atleast_2d(np.asarray(fluxerrs, dtype=np.float64))
    S, n_b = fluxes.shape
    if n_b != ms.n_b:
        raise ValueError("catalogue has %d bands, dictionary has %d" % (n_b, ms.n_b))
    if nus.shape[0] == 1 and S > 1:
        nus = np.broadcast_to(nus, (S, n_b))
    zpos = {float(v): i for i, v in enumerate(ms.z_grid)}
    try:
        zi = np.array([zpos[float(v)] for v in z], dtype=np.int32)
    except KeyError as e:
        raise ValueError("source redshift %s has no dictionary z-slice" % e)

    with np.errstate(divide="ignore"):
        isig = 1.0 / fluxerrs
    zc = z[:, None]
    # ln_likelihood's x_valid (PARAMETERSPACE:144-147); `a | b & c` == a | (b & c)
    if ms.settings['XRAYS'] == 'Prior':
        valid = (nus < np.log10(10 ** (15.38) / (1 + zc))) & (fluxes > -99.e-23)
    else:
        valid = (nus < np.log10(10 ** (15.38) / (1 + zc))) | (nus > np.log10(10 ** (16.685) / (1 + zc))) & (fluxes > -99.e-23)
    mask = _bits(valid)

    n_par = ms.n_par
    pmin = np.empty((S, n_par))
    pmax = np.empty((S, n_par))
    cache = {} if bounds_cache is None else bounds_cache
    uz, inv = np.unique(z, return_inverse=True)
    bmin = np.empty((len(uz), n_par))
    bmax = np.empty((len(uz), n_par))
    for k, zk in enumerate(uz):
        zk = float(zk)
        if zk not in cache:
            cache[zk] = pdict_bounds(ms, zk)
        bmin[k], bmax[k] = cache[zk]
    pmin[:], pmax[:] = bmin[inv], bmax[inv]

    consts = np.zeros((S, C_NCONST))
    consts[:, C_INV4PIDL2] = 1.0 / (4 * np.pi * dlum ** 2)
    consts[:, C_LUMFACTOR] = 4. * np.pi * dlum ** 2.
    consts[:, C_LUMFACTOR_BB] = 4 * np.pi * (dlum) ** 2
    if ms.settings.get('XRAYS') == 'Prior_midIR':
        # PRIORS:316-318 — first datum with 17.685 < log nu_rest < 18.384
        r = nus + np.log10(1 + zc)
        ok = (17.685 < r) & (r < 18.384)
        if not ok.any(axis=1).all():
            bad = int(np.nonzero(~ok.any(axis=1))[0][0])
            raise IndexError("source %d has no datum with 17.685 < log nu_rest < 18.384; prior_IR_XRays raises "
                             "IndexError in the reference (PRIORS_AGNfitter.py:317)" % bad)
        fx = fluxes[np.arange(S), np.argmax(ok, axis=1)]
        with np.errstate(all="ignore"):
            consts[:, C_LOGF_XRAY] = np.log10(fx * consts[:, C_LUMFACTOR])
    agn_branch = np.zeros(S, dtype=np.int32)
    irsyn_mode = np.zeros(S, dtype=np.int32)
    irsyn_j = np.zeros(S, dtype=np.int32)
    if ms.settings.get('PRIOR_AGNfraction') == True:
        # PRIORS:163-164,169-172,182 — last positive flux with 14.3 < nu_rest < 15.3
        r = nus + np.log10(1 + zc)
        ok = (14.3 < r) & (r < 15.3) & (fluxes > 0)
        if not ok.any(axis=1).all():
            bad = int(np.nonzero(~ok.any(axis=1))[0][0])
            raise IndexError("source %d has no positive flux with 14.3 < log nu_rest < 15.3; prior_AGNfraction "
                             "raises IndexError in the reference (PRIORS_AGNfitter.py:164)" % bad)
        last = n_b - 1 - np.argmax(ok[:, ::-1], axis=1)
        d1500 = fluxes[np.arange(S), last]
        lumf = 4. * np.pi * dlum ** 2.
        abs_mag_data = 51.6 - 2.5 * np.log10(lumf * d1500)
        characteristic_mag = -35.4 * (1 + z) ** 0.524 / (1 + (1 + z) ** 0.678)
        agn_branch[:] = np.where(abs_mag_data > (characteristic_mag - 1.), 0, 1)
    if ms.radio:
        for s in range(S):
            mode, j = hostmath.irsyn_constants(nus[s], fluxes[s], ms.bands, z[s])
            if mode == 2:
                raise ValueError("source %d: F_IR/F_syn == 2 exactly leaves prior_IR_SYNfraction "
                                 "undefined in the reference (PRIORS_AGNfitter.py:244-254)" % s)
            irsyn_mode[s], irsyn_j[s] = mode, j
    return SourceSet(zi=zi, y=np.ascontiguousarray(fluxes), isig=np.ascontiguousarray(isig), mask=mask,
                     pmin=pmin, pmax=pmax, consts=consts, agn_branch=agn_branch,
                     irsyn_mode=irsyn_mode, irsyn_j=irsyn_j)
