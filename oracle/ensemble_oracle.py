"""
TEST INFRASTRUCTURE — CPU oracle of the ensemble step (K3).  NOT part of the product path.

emcee is not installed here and the reference's chains are unpinned by the reference itself (unseeded
RNG, PARAMETERSPACE_AGNfitter.py:260,287), so the sampler parity is pinned on the ALGORITHM emcee
documents for `StretchMove` / `RedBlueMove` (emcee 3: moves/stretch.py, moves/red_blue.py), with
`randomize_split=False`:

    for each half S (the other half is C):
        zz   = ((a - 1) * u1 + 1) ** 2 / a
        q    = C[j] - (C[j] - S) * zz[:, None]              j uniform over the other half
        lnpdiff = (ndim - 1) * log(zz) + lnp(q) - lnp(S);   accept where lnpdiff > log(u3)

The uniforms come from Philox4x32-10 (Salmon, Moraes, Dror & Shaw 2011; Random123's known-answer
vectors are checked in tests/test_ensemble_cpu.py) with counter (step, half, source, walker) and the
64-bit seed as key — the same stream agnfitter_b200/csrc/k3_ensemble.cu draws.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter arrays (uint32); key scalars.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def draws(step, half, S, H, seed, src_offset=0):
    """(u1 [S,H] in [0,1), j [S,H] in [0,H), ln u3 [S,H]) of one half-step; the sources are catalogue rows
    src_offset .. src_offset + S - 1."""
    s = (np.arange(S, dtype=np.int64) + int(src_offset)).astype(np.uint32)[:, None]
    k = np.arange(H, dtype=np.uint32)[None, :]
    seed = int(seed) & (2 ** 64 - 1)
    x0, x1, x2, x3 = philox4x32_10(np.uint32(step), np.uint32(half), s, k, seed & 0xFFFFFFFF, seed >> 32)
    u1 = (((x0.astype(np.uint64) << np.uint64(32)) | x1.astype(np.uint64)) >> np.uint64(11)).astype(np.float64) \
        * 1.1102230246251565e-16
    j = ((x2.astype(np.uint64) * np.uint64(H)) >> np.uint64(32)).astype(np.int64)
    u3 = (x3.astype(np.float64) + 0.5) * 2.3283064365386963e-10
    return u1, j, np.log(u3)


def proposals(pos, step, half, seed, a=2.0, src_offset=0):
    """pos [S, W, n_par] -> (q [S, H, n_par], factor [S, H], ln u3 [S, H]) for one half-step."""
    S, W, n_par = pos.shape
    H = W // 2
    u1, j, lnu = draws(step, half, S, H, seed, src_offset)
    t = (a - 1.0) * u1 + 1.0
    zz = (t * t) / a
    mine = pos[:, H:] if half else pos[:, :H]
    other = pos[:, :H] if half else pos[:, H:]
    c = np.take_along_axis(other, j[:, :, None], axis=1)
    q = c - (c - mine) * zz[:, :, None]
    return q, (n_par - 1) * np.log(zz), lnu


def step(pos, lnp, n_acc, lnprob_fn, step_idx, seed, a=2.0, src_offset=0):
    """One full stretch-move step in place.  lnprob_fn(q [S, H, n_par]) -> [S, H]."""
    S, W, n_par = pos.shape
    H = W // 2
    for half in (0, 1):
        q, factor, lnu = proposals(pos, step_idx, half, seed, a, src_offset)
        new = lnprob_fn(q)
        sl = slice(H, W) if half else slice(0, H)
        with np.errstate(invalid="ignore"):
            accept = (factor + new - lnp[:, sl]) > lnu
        pos[:, sl][accept] = q[accept]
        lnp[:, sl][accept] = new[accept]
        n_acc[:, sl] += accept
    return pos, lnp, n_acc


def ball(centre, W, scale, normal, stream_id, seed, src_offset=0):
    """Positions [S, W, n_par] of agnf_ens_set_ball: centre + scale * d with d uniform in (-1, 1) (the initial ball,
    PARAMETERSPACE_AGNfitter.py:260-263) or standard normal by Box-Muller (the restart ball, :287-292); Philox
    counter (stream_id, 2 + p, source, walker)."""
    centre = np.atleast_2d(np.asarray(centre, dtype=np.float64))
    S, n_par = centre.shape
    s = (np.arange(S, dtype=np.int64) + int(src_offset)).astype(np.uint32)[:, None, None]
    w = np.arange(W, dtype=np.uint32)[None, :, None]
    p = (np.arange(n_par, dtype=np.uint32) + np.uint32(2))[None, None, :]
    seed = int(seed) & (2 ** 64 - 1)
    x0, x1, x2, x3 = philox4x32_10(np.uint32(stream_id), p, s, w, seed & 0xFFFFFFFF, seed >> 32)
    to53 = lambda hi, lo: (((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(11)) \
        .astype(np.float64) * 1.1102230246251565e-16
    u1, u2 = to53(x0, x1), to53(x2, x3)
    d = np.sqrt(-2.0 * np.log1p(-u1)) * np.cos(2.0 * np.pi * u2) if normal else 2.0 * u1 - 1.0
    return centre[:, None, :] + d * scale
