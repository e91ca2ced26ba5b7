"""GPU: the ensemble sampler (K3 + K1) against the CPU oracle step for step — same Philox stream, same
proposals bit for bit, same accept decisions — and its on-device percentile reduction; then the
reference's burn-in / production schedule on the example catalogue's two sources."""
import os

import numpy as np
import pytest

from helpers import load_k1, oracle_src
from agnfitter_b200 import sources
from agnfitter_b200.engine import Engine
from agnfitter_b200.ensemble import EnsembleSampler, REFERENCE_PERCENTILES
from oracle import agn_oracle as O
from oracle import ensemble_oracle as E

pytestmark = pytest.mark.gpu


def assert_histogram_quantiles(got, samples, width, q=REFERENCE_PERCENTILES):
    """got [n_par, n_q] are q-quantiles of samples [N, n_par] up to the histogram's resolution: the
    empirical CDF brackets q within one bin either side (np.percentile interpolates across gaps between
    order statistics, a histogram cannot, so values are compared through the CDF, not directly)."""
    n = samples.shape[0]
    for p in range(samples.shape[1]):
        x = np.sort(samples[:, p])
        for k, qq in enumerate(q):
            lo = np.searchsorted(x, got[p, k] - 1.01 * width[p], side="left") / n     # P(x < got - w)
            hi = np.searchsorted(x, got[p, k] + 1.01 * width[p], side="right") / n    # P(x <= got + w)
            assert lo <= qq / 100 + 1.0 / n and hi >= qq / 100 - 1.0 / n, (p, qq, got[p, k], lo, hi)


def _setup(case):
    ms, g = load_k1(case)
    eng = Engine(ms, device=0)
    src = sources.make_sources(ms, g["src_z"], g["src_dlum"], g["src_nus"], g["src_fluxes"], g["src_fluxerrs"])
    eng.set_sources(src)
    osrc = [oracle_src(ms, g, s) for s in range(src.S)]

    def lnprob(q):                                   # q [S, H, n_par] -> [S, H], CPU oracle
        return np.stack([O.ln_probab(ms, osrc[s], q[s]) for s in range(src.S)])
    return ms, eng, src, lnprob


@pytest.mark.parametrize("split", ["random", "fixed"])
@pytest.mark.parametrize("fused", [1, 4, 3, 0])
@pytest.mark.parametrize("case", ["classic18", "cat49_catalog_example"])
def test_ensemble_matches_oracle_step_for_step(case, fused, split):
    """fused = 1: the default — for a shard this small the persistent per-source kernel (one launch per run() call,
    one CTA per source); 4: one cooperative K1 launch per half-step that draws, evaluates and accepts (what large
    shards run); 3: the same through the gather4 kernel; 0: propose / K1 / accept kernels.
    split: emcee 3's per-step random division of the walkers (default) or the fixed halves."""
    ms, eng, src, lnprob = _setup(case)
    eng.set_option("ens_fused", 1 if fused else 0)
    if fused == 3:
        eng.set_option("k1_variant", 3)
    if fused == 4:
        eng.set_option("ens_persist", 0)
    W, seed, n_steps = 24, 20260101, 40
    rng = np.random.default_rng(5)
    pos0 = src.pmin[:, None, :] + rng.uniform(size=(src.S, W, ms.n_par)) * (src.pmax - src.pmin)[:, None, :]
    pos0[:, : W // 3] = EnsembleSampler.initial_positions(src.pmin, src.pmax, W, rng)[:, : W // 3]
    ens = EnsembleSampler(eng, W, seed=seed, n_bins=0, split=split)
    ens.set_positions(pos0)
    gpos, glnp, gacc = ens.state()
    opos, olnp, oacc = pos0.copy(), lnprob(pos0), np.zeros((src.S, W), dtype=np.int64)
    assert np.array_equal(gpos, pos0)
    fin = np.isfinite(olnp)
    assert np.array_equal(np.isfinite(glnp), fin) and np.allclose(glnp[fin], olnp[fin], rtol=1e-10, atol=0)
    for it in range(n_steps):
        ens.run(1)
        E.step(opos, olnp, oacc, lnprob, it, seed, split=split)
        gpos, glnp, gacc = ens.state()
        assert np.array_equal(gacc, oacc), "accept decisions differ at step %d" % it
        assert np.array_equal(gpos, opos), "positions differ at step %d" % it       # proposals are bit-identical
        fin = np.isfinite(olnp)
        assert np.allclose(glnp[fin], olnp[fin], rtol=1e-10, atol=0)
    assert oacc.sum() > 0, "the test must exercise accepted moves"
    if fused:
        assert eng.get_option("k1_last_kernel") == (2 if fused == 3 else 1), "the requested kernel must be the one that ran"
        assert eng.get_option("ens_last_persist") == (1 if fused == 1 else 0)


def test_ensemble_percentiles_and_best():
    ms, eng, src, _ = _setup("classic18")
    W, n_bins = 40, 2048
    rng = np.random.default_rng(11)
    ens = EnsembleSampler(eng, W, seed=7, n_bins=n_bins)
    ens.set_positions(EnsembleSampler.initial_positions(src.pmin, src.pmax, W, rng))
    ens.run(200, track_best=True)
    ens.reset_stats()
    chain, best = [], -np.inf
    for it in range(120):
        ens.run(1, hist_every=1, track_best=True)
        chain.append(ens.state()[0])
    chain = np.stack(chain, axis=1)                              # [S, steps, W, n_par]
    got = ens.percentiles()
    width = (src.pmax - src.pmin) / n_bins
    for s in range(src.S):
        assert_histogram_quantiles(got[s], chain[s].reshape(-1, ms.n_par), width[s])
    bp, bl = ens.best()
    pos, lnp, acc = ens.state()
    assert np.all(bl >= lnp.max(axis=1))
    back = eng.lnprob(bp, np.arange(src.S, dtype=np.int32))
    assert np.array_equal(back, bl), "best position and best lnP must belong together"
    assert acc.sum() > 0


def test_reference_schedule_on_example_catalogue():
    """Burn-in sets restarted from the best sample, reset, production (MCMC_AGNfitter.py:122-135), then the thinned
    second half of the production chain that PLOTandWRITE keeps (:476-487): the GPU run and the oracle run with the
    same seeds give the same percentiles and the same best-fit row."""
    ms, eng, src, lnprob = _setup("cat49_catalog_example")
    mc = dict(Nwalkers=20, Nburnsets=2, Nburn=25, Nmcmc=60, Nthin=2)
    seed, n_bins = 99, 4096
    ens = EnsembleSampler(eng, mc['Nwalkers'], seed=seed, n_bins=n_bins)
    out = ens.run_reference_schedule(mc, src.pmin, src.pmax, np.random.default_rng(3))

    # the same schedule on the CPU oracle
    rng = np.random.default_rng(3)
    W = mc['Nwalkers']
    pos = EnsembleSampler.initial_positions(src.pmin, src.pmax, W, rng)
    step = 0
    best_l = np.full(src.S, -np.inf)
    best_p = np.zeros((src.S, ms.n_par))

    def track(pos, lnp):
        k = lnp.argmax(axis=1)
        v = lnp[np.arange(src.S), k]
        up = v > best_l
        best_l[up] = v[up]
        best_p[up] = pos[np.arange(src.S), k][up]

    # emcee's sampler is not reset between the burn-in sets (MCMC_AGNfitter.py:125-128): get_best_position takes the
    # argmax over everything sampled so far, and the starting positions are not part of the chain
    for _ in range(mc['Nburnsets']):
        lnp = lnprob(pos)
        acc = np.zeros((src.S, W), dtype=np.int64)
        for _ in range(mc['Nburn']):
            E.step(pos, lnp, acc, lnprob, step, seed)
            step += 1
            track(pos, lnp)
        pos = EnsembleSampler.ball_around(best_p, W, rng)
    lnp = lnprob(pos)
    best_l[:] = -np.inf
    acc = np.zeros((src.S, W), dtype=np.int64)
    kept = []
    half = int(mc['Nmcmc'] / 2)
    for k in range(mc['Nmcmc']):                    # state k = the ensemble after production step k + 1
        E.step(pos, lnp, acc, lnprob, step, seed)
        step += 1
        if k >= half and (k - half) % mc['Nthin'] == 0:          # CHAIN.props: chain[:, int(n/2):n:Nthin]
            kept.append(pos.copy())
            track(pos, lnp)
    chain = np.stack(kept, axis=1)
    width = (src.pmax - src.pmin) / n_bins
    for s in range(src.S):
        assert_histogram_quantiles(out['percentiles'][s], chain[s].reshape(-1, ms.n_par), width[s])
    assert np.allclose(out['best_lnp'], best_l, rtol=1e-10)
    assert np.array_equal(out['best_pos'], best_p)
    assert np.allclose(out['acceptance'], acc.mean(axis=1) / mc['Nmcmc'])


def test_chain_recording_and_reference_chain_files(tmp_path):
    """agnf_ens_record keeps the walker state of every step on the device; the recorded chain is exactly the
    sequence of states (step for step); run_reference_schedule(save_dir=...) writes samples_burn1-2-3.sav /
    samples_mcmc.sav that the reference's readers (get_best_position, PARAMETERSPACE:268-294; CHAIN.props,
    PLOTandWRITE:471-492) can consume: same keys, axis order and dtypes as MCMC_AGNfitter.save_chains (:192-206)."""
    import pickle
    ms, eng, src, lnprob = _setup("cat49_catalog_example")
    W = 16
    ens = EnsembleSampler(eng, W, seed=77, n_bins=0)
    rng = np.random.default_rng(1)
    ens.set_positions(EnsembleSampler.initial_positions(src.pmin, src.pmax, W, rng))
    ens.record(12)
    states = []
    for _ in range(12):
        ens.run(1)
        p, l, _ = ens.state()
        states.append((p, l))
    ch, lp = ens.chain()
    assert ch.shape == (src.S, W, 12, ms.n_par) and lp.shape == (src.S, W, 12)
    for k, (p, l) in enumerate(states):
        assert np.array_equal(ch[:, :, k, :], p) and np.array_equal(lp[:, :, k], l, equal_nan=True)
    with pytest.raises(RuntimeError):
        ens.run(1)                                            # buffer full: fails instead of dropping samples
    ens.record(6, every=2)
    ens.run(12)
    assert ens.chain()[0].shape[2] == 6
    ens.record(0)
    # the reference's files
    mc = dict(Nwalkers=W, Nburnsets=2, Nburn=10, Nmcmc=30, Nthin=2)
    ens = EnsembleSampler(eng, W, seed=78, n_bins=256)
    names = ["srcA", "srcB"][:src.S]
    out = ens.run_reference_schedule(mc, src.pmin, src.pmax, np.random.default_rng(2), save_dir=str(tmp_path),
                                     names=names)
    for s, nm in enumerate(names):
        with open(tmp_path / nm / "samples_burn1-2-3.sav", "rb") as f:
            burn = pickle.load(f)
        with open(tmp_path / nm / "samples_mcmc.sav", "rb") as f:
            mcmc = pickle.load(f)
        assert set(mcmc.keys()) == {"chain", "accept", "lnprob", "final_pos", "state", "acor"}
        assert burn["chain"].shape == (W, 20, ms.n_par) and burn["lnprob"].shape == (W, 20)
        assert mcmc["chain"].shape == (W, 30, ms.n_par) and mcmc["lnprob"].shape == (W, 30)
        assert mcmc["accept"].shape == (W,) and mcmc["acor"].shape == (ms.n_par,)
        assert np.all((mcmc["accept"] >= 0) & (mcmc["accept"] <= 1))
        # get_best_position's arithmetic on the burn-in file (PARAMETERSPACE_AGNfitter.py:281-283): the production run
        # started from a ball around exactly this sample — the running best over BOTH burn-in sets
        i = burn["lnprob"].ravel().argmax()
        ml = burn["chain"].reshape(-1, ms.n_par)[i]
        assert np.all(ml >= src.pmin[s]) and np.all(ml <= src.pmax[s])
        assert np.array_equal(ml, out["restart_centres"][-1][s])
        # ... and 'accept' covers all burn-in steps, like emcee's acceptance_fraction without a reset (the jump to the
        # restart ball between the two sets shows up in the chain as one more change that is not an accepted move)
        moved = (np.abs(np.diff(burn["chain"], axis=1)).sum(axis=2) > 0).sum(axis=1)
        assert np.all(burn["accept"] * 20 >= moved - 1 - 1e-9) and np.all(burn["accept"] * 20 <= 20)
        within = np.r_[0:9, 10:19]
        moved_in = (np.abs(np.diff(burn["chain"], axis=1)).sum(axis=2) > 0)[:, within].sum(axis=1)
        assert np.all(burn["accept"] * 20 >= moved_in - 1e-9)
        # CHAIN.props' arithmetic on the production file
        nsamples = mcmc["chain"].shape[1]
        flat = mcmc["chain"][:, int(nsamples / 2):nsamples:2, :].reshape(-1, ms.n_par)
        lnflat = mcmc["lnprob"][:, int(nsamples / 2):nsamples:2].ravel()
        best = flat[(-lnflat).argsort()[0]]
        assert np.isfinite(lnflat).all() and flat.shape == (W * 8, ms.n_par)
        assert lnflat.max() <= out["best_lnp"][s] + 1e-12
        assert np.array_equal(mcmc["final_pos"], ens.state()[0][s])
        assert np.all(best >= src.pmin[s]) and np.all(best <= src.pmax[s])


def test_device_balls_match_the_oracle_and_the_sharding():
    """agnf_ens_set_ball: the uniform ball bit for bit, the normal one to a few ulp of the device's log / cos;
    a shard that starts at catalogue row 1 draws what the full run draws for that row; the schedule with
    rng=None (balls on the device) runs and restarts around the best sample."""
    ms, eng, src, lnprob = _setup("cat49_catalog_example")
    W, seed = 12, 99
    ens = EnsembleSampler(eng, W, seed=seed, n_bins=0)
    mid = 0.5 * (src.pmin + src.pmax)
    ens.set_ball(mid, 1e-5, normal=False, stream_id=1)
    pos, lnp, acc = ens.state()
    assert np.array_equal(pos, E.ball(mid, W, 1e-5, False, 1, seed))
    assert np.all(np.abs(pos - mid[:, None, :]) <= 1e-5) and not acc.any()
    want_lnp = lnprob(pos)
    fin = np.isfinite(want_lnp)
    assert np.array_equal(np.isfinite(lnp), fin) and np.allclose(lnp[fin], want_lnp[fin], rtol=1e-10, atol=0)
    ens.set_ball(mid, 0.5, normal=True, stream_id=2)
    pos_n = ens.state()[0]
    want = E.ball(mid, W, 0.5, True, 2, seed)
    assert np.allclose(pos_n, want, rtol=0, atol=1e-13)
    assert not np.array_equal(pos_n, pos)
    d = (pos_n - mid[:, None, :]) / 0.5
    assert abs(d.mean()) < 0.25 and 0.75 < d.std() < 1.25
    if src.S > 1:                                   # the second source alone, as a shard starting at row 1
        from agnfitter_b200 import sources as _s
        ms1, g1 = load_k1("cat49_catalog_example")
        one = _s.make_sources(ms, g1["src_z"][1:2], g1["src_dlum"][1:2], g1["src_nus"][1:2], g1["src_fluxes"][1:2],
                              g1["src_fluxerrs"][1:2])
        eng1 = Engine(ms, device=0)
        eng1.set_sources(one)
        ens1 = EnsembleSampler(eng1, W, seed=seed, n_bins=0, source_offset=1)
        ens1.set_ball(mid[1:2], 0.5, normal=True, stream_id=2)
        assert np.array_equal(ens1.state()[0][0], pos_n[1])
    mc = dict(Nwalkers=W, Nburnsets=2, Nburn=8, Nmcmc=16, Nthin=2)
    ens = EnsembleSampler(eng, W, seed=seed, n_bins=128)
    out = ens.run_reference_schedule(mc, src.pmin, src.pmax, None)
    assert np.isfinite(out["best_lnp"]).all() and np.all(out["acceptance"] >= 0)


def test_sampler_state_is_invalidated_by_a_new_source_upload():
    """walkers belong to one set of sources; uploading others tears the ensemble down instead of
    leaving it pointing at freed operands."""
    from agnfitter_b200._lib import AgnfError
    ms, eng, src, _ = _setup("classic18")
    ens = EnsembleSampler(eng, 16, seed=1, n_bins=64)
    ens.set_positions(EnsembleSampler.initial_positions(src.pmin, src.pmax, 16, np.random.default_rng(0)))
    ens.run(3, hist_every=1)
    eng.set_sources(src)
    with pytest.raises(AgnfError, match="agnf_ens_create"):
        ens.run(1)
    ens2 = EnsembleSampler(eng, 16, seed=1, n_bins=64, split="fixed")       # the other split is allowed again
    ens2.set_positions(EnsembleSampler.initial_positions(src.pmin, src.pmax, 16, np.random.default_rng(0)))
    ens2.run(3, hist_every=1)
    assert np.isfinite(ens2.percentiles()).all()


def test_chains_agree_in_distribution_with_independent_oracle_chains():
    """Distribution-level check on both example-catalogue sources: medians of R independent GPU ensembles (random
    split, one Philox stream per replica) against independent CPU-oracle ensembles with other seeds.  Same algorithm,
    same starting ball, same number of steps -> the two sets of medians are draws from ONE distribution whether or
    not the chains have converged; their means must agree within the Monte Carlo error estimated from the replicas."""
    ms, g = load_k1("cat49_catalog_example")
    S0 = len(g["src_z"])
    R_gpu, R_cpu, W, n_burn, n_keep = 48, 6, 24, 150, 150
    rep = np.repeat(np.arange(S0), R_gpu)
    eng = Engine(ms, device=0)
    src = sources.make_sources(ms, g["src_z"][rep], g["src_dlum"][rep], g["src_nus"][rep], g["src_fluxes"][rep],
                               g["src_fluxerrs"][rep])
    eng.set_sources(src)
    ens = EnsembleSampler(eng, W, seed=2024, n_bins=0)
    mid = 0.5 * (src.pmin + src.pmax)
    ens.set_ball(mid, 1e-5, normal=False, stream_id=1)
    ens.run(n_burn)
    ens.record(n_keep)
    ens.run(n_keep)
    ch, _ = ens.chain()                                                # [S0 * R_gpu, W, n_keep, n_par]
    med_gpu = np.median(ch.reshape(S0, R_gpu, -1, ms.n_par), axis=2)   # [S0, R_gpu, n_par]
    osrc = [oracle_src(ms, g, s) for s in range(S0)]
    med_cpu = np.empty((S0, R_cpu, ms.n_par))
    for s in range(S0):
        lnprob = lambda q: np.stack([O.ln_probab(ms, osrc[s], q[r]) for r in range(q.shape[0])])
        pos = E.ball(np.repeat(mid[s * R_gpu][None, :], R_cpu, axis=0), W, 1e-5, False, 1, seed=777 + s)
        lnp = lnprob(pos)
        acc = np.zeros((R_cpu, W), dtype=np.int64)
        keep = []
        for it in range(n_burn + n_keep):
            E.step(pos, lnp, acc, lnprob, it, seed=777 + s)
            if it >= n_burn:
                keep.append(pos.copy())
        med_cpu[s] = np.median(np.stack(keep, axis=1).reshape(R_cpu, -1, ms.n_par), axis=1)
    span = (src.pmax[0] - src.pmin[0])
    for s in range(S0):
        sd = med_gpu[s].std(axis=0, ddof=1)
        err = np.sqrt(sd ** 2 / R_gpu + sd ** 2 / R_cpu)
        diff = np.abs(med_gpu[s].mean(axis=0) - med_cpu[s].mean(axis=0))
        assert np.all(diff <= 4.5 * err + 1e-3 * span), (s, diff / np.maximum(err, 1e-300))


def test_exact_percentiles_match_the_references_chain_props():
    """N2 exact: the radix-select percentiles of the recorded chain equal what the reference's own CHAIN.props +
    np.percentile (PLOTandWRITE_AGNfitter.py:471-492, :162) give for the same chain (tests/golden/chain_props.npz,
    generated from the unmodified reference): thinning [:, n/2::Nthin], values to the last bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_props.npz"))
    chain, nthin = g["chain"], int(g["Nthin"])                    # [W, n, n_par]
    W, n, n_par = chain.shape
    ms, g1 = load_k1("classic18")
    assert ms.n_par == n_par
    eng = Engine(ms, device=0)
    # three copies of the source: source 1 gets the chain, its neighbours something else (indexing across sources)
    rep = lambda a: np.repeat(np.asarray(a)[:1], 3, axis=0)
    src = sources.make_sources(ms, rep(g1["src_z"]), rep(g1["src_dlum"]), rep(g1["src_nus"]), rep(g1["src_fluxes"]),
                               rep(g1["src_fluxerrs"]))
    eng.set_sources(src)
    ens = EnsembleSampler(eng, W, seed=1, n_bins=0)
    ens.record(n)
    rng = np.random.default_rng(0)
    other = rng.normal(size=(2, W, n, n_par))
    for k in range(n):
        ens.set_positions(np.stack([other[0, :, k], chain[:, k], other[1, :, k]]))
        ens.record_state()
    got, got_lnp = ens.chain_percentiles(first=int(n / 2), stride=nthin, with_lnp=True)
    assert np.array_equal(got[1].T, g["ref_percentiles"]), np.abs(got[1].T - g["ref_percentiles"]).max()
    for s, o in ((0, other[0]), (2, other[1])):
        flat = o[:, int(n / 2):n:nthin].reshape(-1, n_par)
        assert np.array_equal(got[s].T, np.percentile(flat, REFERENCE_PERCENTILES, axis=0))
    # the log-posterior column (-inf for the out-of-box states of this synthetic chain) and other quantile sets
    ch, lp = ens.chain()
    with np.errstate(invalid="ignore"):
        want = np.percentile(lp[:, :, int(n / 2):n:nthin].reshape(3, -1), REFERENCE_PERCENTILES, axis=1).T
    assert np.array_equal(got_lnp, want, equal_nan=True)
    q = [0.0, 1.0, 33.3, 50.0, 99.9, 100.0]
    full = ens.chain_percentiles(q)
    assert np.array_equal(full[1].T, np.percentile(chain.reshape(-1, n_par), q, axis=0))
    with pytest.raises(RuntimeError):
        ens.chain_percentiles(first=n)


def test_reference_schedule_exact_percentiles():
    """run_reference_schedule(exact_percentiles=True) keeps only the thinned second half of the production chain on
    the device and reports np.percentile-exact values; with chain files requested the same numbers come from the
    full recorded chain."""
    import pickle, tempfile
    ms, eng, src, _ = _setup("cat49_catalog_example")
    mc = dict(Nwalkers=20, Nburnsets=1, Nburn=15, Nmcmc=41, Nthin=3)
    a = EnsembleSampler(eng, 20, seed=5, n_bins=0).run_reference_schedule(mc, src.pmin, src.pmax, None,
                                                                            exact_percentiles=True)
    assert a["percentiles_exact"] and a["percentiles"].shape == (src.S, ms.n_par, 5)
    with tempfile.TemporaryDirectory() as d:
        eng.set_sources(src)
        b = EnsembleSampler(eng, 20, seed=5, n_bins=0).run_reference_schedule(mc, src.pmin, src.pmax, None, save_dir=d,
                                                                                names=["a", "b"][:src.S],
                                                                                exact_percentiles=True)
        assert np.array_equal(a["percentiles"], b["percentiles"]) and np.array_equal(a["best_pos"], b["best_pos"])
        for s, nm in enumerate(["a", "b"][:src.S]):
            with open(os.path.join(d, nm, "samples_mcmc.sav"), "rb") as f:
                smp = pickle.load(f)
            n = smp["chain"].shape[1]
            flat = smp["chain"][:, int(n / 2):n:3, :].reshape(-1, ms.n_par)
            lnflat = smp["lnprob"][:, int(n / 2):n:3].ravel()
            assert np.array_equal(b["percentiles"][s].T, np.percentile(flat, REFERENCE_PERCENTILES, axis=0))
            assert np.array_equal(b["lnp_percentiles"][s], np.percentile(lnflat, REFERENCE_PERCENTILES))
            assert lnflat.max() == b["best_lnp"][s]
            assert np.array_equal(eng.lnprob(b["best_pos"][s][None, :], np.array([s], dtype=np.int32)), [lnflat.max()])


@pytest.mark.parametrize("case", ["classic18", "radio_xray49"])
def test_persistent_sampler_equals_lock_step(case):
    """One launch for a whole run() call (one CTA per source, steps looped inside the kernel) against the lock-step
    path: same walkers, acceptance counters, best sample, histograms and recorded chain, bit for bit — over calls
    with thinned steps, every-step best tracking and chain recording with a stride."""
    ms, eng, src, _ = _setup(case)
    W = 20
    rng = np.random.default_rng(8)
    pos0 = src.pmin[:, None, :] + rng.uniform(size=(src.S, W, ms.n_par)) * (src.pmax - src.pmin)[:, None, :]
    got = {}
    # (persist, walkers per warp): automatic (ceil(H / 8): eight warps share a half-step), 32 (one warp per 32 rows)
    # and an uneven cut (3: the last warp owns a single row)
    for persist, rows in ((1, 0), (1, 32), (1, 3), (0, 0)):
        eng.set_sources(src)
        eng.set_option("ens_persist", persist)
        eng.set_option("ens_persist_rows", rows)
        ens = EnsembleSampler(eng, W, seed=31, n_bins=128)
        ens.set_positions(pos0)
        ens.run(7, track_best=True)
        ens.record(9, every=3)
        ens.run(25, hist_every=5, track_best='thinned')
        assert eng.get_option("ens_last_persist") == persist
        ens.run(1)
        got[(persist, rows)] = (ens.state(), ens.best(), ens.percentiles(), ens.chain())
    eng.set_option("ens_persist_rows", 0)
    b = got[(0, 0)]
    for key in ((1, 0), (1, 32), (1, 3)):
        a = got[key]
        for x, y in zip(a[0], b[0]):
            assert np.array_equal(x, y, equal_nan=True), key
        assert np.array_equal(a[1][0], b[1][0]) and np.array_equal(a[1][1], b[1][1]), key
        assert np.array_equal(a[2], b[2], equal_nan=True), key
        assert a[3][0].shape[2] == 9 and np.array_equal(a[3][0], b[3][0]), key
        assert np.array_equal(a[3][1], b[3][1], equal_nan=True), key
    assert a[0][2].sum() > 0
