"""CPU: the ensemble oracle's building blocks — Philox known answers and the stretch-move algebra."""
import numpy as np

from oracle import ensemble_oracle as E


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors: philox4x32 10 rounds."""
    def run(c, k):
        return [int(v) for v in E.philox4x32_10(*[np.uint32(x) for x in c], k[0], k[1])]
    assert run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_ranges_and_independence():
    u1, j, lnu = E.draws(step=3, half=1, S=50, H=40, seed=1234567890123)
    assert u1.shape == (50, 40) and np.all((u1 >= 0) & (u1 < 1))
    assert np.all((j >= 0) & (j < 40)) and np.all(np.isfinite(lnu)) and np.all(lnu < 0)
    u1b, jb, _ = E.draws(step=4, half=1, S=50, H=40, seed=1234567890123)
    assert not np.array_equal(u1, u1b) and abs(np.corrcoef(u1.ravel(), u1b.ravel())[0, 1]) < 0.1
    assert abs(u1.mean() - 0.5) < 0.03 and abs(j.mean() - 19.5) < 1.5


def test_split_perm_is_an_equal_random_partition():
    """emcee 3's RedBlueMove: every step divides the walkers afresh into two sets of W/2 (moves/red_blue.py)."""
    S, W = 200, 30
    p = E.split_perm(7, S, W, seed=5)
    assert p.shape == (S, W) and np.array_equal(np.sort(p, axis=1), np.broadcast_to(np.arange(W), (S, W)))
    p2 = E.split_perm(8, S, W, seed=5)
    assert not np.array_equal(p, p2)
    in0 = np.zeros((S, W))
    np.put_along_axis(in0, p[:, : W // 2], 1.0, axis=1)
    assert np.all(in0.sum(axis=1) == W // 2)
    assert np.all(np.abs(in0.mean(axis=0) - 0.5) < 0.15)          # every walker lands in set 0 about half the time
    # a shard starting at catalogue row 3 draws the split of rows 3.. of the full run
    assert np.array_equal(E.split_perm(7, S - 3, W, seed=5, src_offset=3), p[3:])


import pytest


@pytest.mark.parametrize("split", ["random", "fixed"])
def test_stretch_move_samples_a_gaussian(split):
    """Detailed balance in practice: the oracle step samples N(mu, sigma^2) in 3-D — with emcee 3's per-step random
    split (the default) and with the fixed halves alike."""
    rng = np.random.default_rng(0)
    S, W, n = 4, 40, 3
    mu, sig = np.array([1.0, -2.0, 0.5]), np.array([0.5, 2.0, 1.0])
    lnprob = lambda q: -0.5 * (((q - mu) / sig) ** 2).sum(axis=-1)
    pos = rng.normal(size=(S, W, n))
    lnp = lnprob(pos)
    acc = np.zeros((S, W), dtype=np.int64)
    keep = []
    for it in range(1500):
        E.step(pos, lnp, acc, lnprob, it, seed=42, split=split)
        if it >= 300 and it % 5 == 0:
            keep.append(pos.copy())
    x = np.concatenate(keep, axis=1).reshape(-1, n)
    assert np.allclose(x.mean(axis=0), mu, atol=0.08 * sig.max())
    assert np.allclose(x.std(axis=0), sig, rtol=0.06)
    frac = acc.mean() / 1500
    assert 0.3 < frac < 0.8


def test_integrated_time_estimator():
    """AR(1) chains with known autocorrelation time tau = (1 + rho) / (1 - rho)."""
    from agnfitter_b200.ensemble import integrated_time
    rng = np.random.default_rng(0)
    rho, n, W = 0.8, 20000, 8
    x = np.zeros((W, n, 2))
    e = rng.normal(size=(W, n, 2))
    for t in range(1, n):
        x[:, t, 0] = rho * x[:, t - 1, 0] + e[:, t, 0]
    x[:, :, 1] = e[:, :, 1]
    tau = integrated_time(x)
    assert abs(tau[0] - 9.0) < 1.0 and abs(tau[1] - 1.0) < 0.15


def test_shard_stream_is_independent_of_the_sharding():
    """Source g draws the same numbers whether its shard starts at row 0 or at row g, and successive calls differ."""
    from agnfitter_b200.ensemble import ShardStream
    S, W, n_par = 11, 6, 5                       # W * n_par = 30: not a multiple of Philox's 4-double blocks
    whole = ShardStream(9, 0)
    u, n = whole.uniform_block(S, W, n_par), whole.normal_block(S, W, n_par)
    for first in (0, 3, 10):
        part = ShardStream(9, first)
        assert np.array_equal(part.uniform_block(S - first, W, n_par), u[first:])
        assert np.array_equal(part.normal_block(S - first, W, n_par), n[first:])
    assert not np.array_equal(whole.uniform_block(S, W, n_par), u)
    assert 0.0 <= u.min() and u.max() < 1.0 and np.isfinite(n).all() and abs(n.mean()) < 0.3 and 0.7 < n.std() < 1.3
