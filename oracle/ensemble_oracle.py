"""
TEST INFRASTRUCTURE — CPU oracle of the ensemble step (K3).  NOT part of the product path.

emcee is not installed here and the reference's chains are unpinned by the reference itself (unseeded
RNG, PARAMETERSPACE_AGNfitter.py:260,287), so the sampler parity is pinned on the ALGORITHM emcee
documents for `StretchMove` / `RedBlueMove` (emcee 3: moves/stretch.py, moves/red_blue.py; the reference names no version — its
`integrated_time(x, c=5, tol=50, quiet=False)` patch, MCMC_AGNfitter.py:107, is the emcee >= 3 signature — and builds
the sampler with the default move, :114, i.e. `randomize_split=True`):

    inds = arange(W) % 2; random.shuffle(inds)              a fresh equal split of the walkers at every step
    for each set S (the other set is C):
        zz   = ((a - 1) * u1 + 1) ** 2 / a
        q    = C[j] - (C[j] - S) * zz[:, None]              j uniform over the other half
        lnpdiff = (ndim - 1) * log(zz) + lnp(q) - lnp(S);   accept where lnpdiff > log(u3)

The split is restated as: order the walkers of a source by a 32-bit Philox key of (step, source, walker), ties by
walker index; the first W/2 are set 0 (`split_perm`; any uniformly random equal partition is equivalent in
distribution, because partners are drawn uniformly inside the complement).  `split=None` keeps the fixed
first-half / second-half sets (`randomize_split=False`).

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


SPLIT_TAG = 0xFFFFFFFF


def split_perm(step, S, W, seed, src_offset=0):
    """[S, W] walker indices: slots [0, W/2) hold set 0 of this step, slots [W/2, W) set 1."""
    s = (np.arange(S, dtype=np.int64) + int(src_offset)).astype(np.uint32)[:, None]
    q = np.arange((W + 3) // 4, dtype=np.uint32)[None, :]
    seed = int(seed) & (2 ** 64 - 1)
    x = philox4x32_10(np.uint32(int(step) & 0xFFFFFFFF), np.uint32(SPLIT_TAG), s, q, seed & 0xFFFFFFFF, seed >> 32)
    keys = np.stack(x, axis=-1).reshape(S, -1)[:, :W]
    return np.argsort(keys, axis=1, kind="stable")


def proposals(pos, step, half, seed, a=2.0, src_offset=0, perm=None):
    """pos [S, W, n_par] -> (q [S, H, n_par], factor [S, H], ln u3 [S, H], moved walkers [S, H]) for one half-step."""
    S, W, n_par = pos.shape
    H = W // 2
    if perm is None:
        perm = np.broadcast_to(np.arange(W), (S, W))
    u1, j, lnu = draws(step, half, S, H, seed, src_offset)
    t = (a - 1.0) * u1 + 1.0
    zz = (t * t) / a
    mine_w = perm[:, H:] if half else perm[:, :H]
    other_w = perm[:, :H] if half else perm[:, H:]
    mine = np.take_along_axis(pos, mine_w[:, :, None], axis=1)
    partner = np.take_along_axis(other_w, j, axis=1)
    c = np.take_along_axis(pos, partner[:, :, None], axis=1)
    q = c - (c - mine) * zz[:, :, None]
    return q, (n_par - 1) * np.log(zz), lnu, mine_w


def step(pos, lnp, n_acc, lnprob_fn, step_idx, seed, a=2.0, src_offset=0, split="random"):
    """One full stretch-move step in place.  lnprob_fn(q [S, H, n_par]) -> [S, H].  split: "random" (emcee 3's
    default) or "fixed"."""
    S, W, n_par = pos.shape
    perm = split_perm(step_idx, S, W, seed, src_offset) if split == "random" else None
    rows = np.arange(S)[:, None]
    for half in (0, 1):
        q, factor, lnu, mine_w = proposals(pos, step_idx, half, seed, a, src_offset, perm)
        new = lnprob_fn(q)
        with np.errstate(invalid="ignore"):
            accept = (factor + new - lnp[rows, mine_w]) > lnu
        si, ki = np.nonzero(accept)
        wi = mine_w[si, ki]
        pos[si, wi] = q[si, ki]
        lnp[si, wi] = new[si, ki]
        n_acc[si, wi] += 1
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
