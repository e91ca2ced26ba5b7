"""
GPU-resident ensemble sampling for a whole catalogue shard (K3).

Mirrors the emcee leg of MCMC_AGNfitter.main (MCMC_AGNfitter.py:98-135,156-189) for S sources at once:
initial ball (PARAMETERSPACE_AGNfitter.get_initial_positions, :251-265), `Nburnsets` burn-in sets each
restarted from a ball around the best sample seen (get_best_position, :268-294), `sampler.reset()`,
production run; the percentiles PLOTandWRITE reports (PLOTandWRITE_AGNfitter.py:152-173) are reduced on
the device from histograms of the thinned chain.  Walkers never leave HBM.
"""
import ctypes as C

import numpy as np

from ._lib import as_f64, ptr

REFERENCE_PERCENTILES = (2.5, 16.0, 50.0, 84.0, 97.5)


def integrated_time(chain, c=5, max_walkers=32):
    """Integrated autocorrelation time per parameter of a chain [W, n, n_par], the way emcee estimates
    `sampler.acor` (emcee.autocorr.integrated_time with c = 5, no convergence check — the reference patches emcee to
    quiet=True for the same reason, MCMC_AGNfitter.py:100-112): autocorrelation function by FFT, averaged over
    walkers, summed up to the first window M >= c * tau(M).  A diagnostic (the reference prints its mean,
    PLOTandWRITE_AGNfitter.py:491): at most `max_walkers` evenly spaced walkers enter the average."""
    from scipy import fft as sfft
    x = np.asarray(chain, dtype=np.float64)
    W, n, n_par = x.shape
    tau = np.full(n_par, np.nan)
    if n < 2:
        return tau
    if W > max_walkers:
        x = x[np.linspace(0, W - 1, max_walkers).astype(int)]
        W = max_walkers
    y = np.ascontiguousarray(x.transpose(2, 0, 1))                        # [n_par, W, n]: FFTs along contiguous rows
    y -= y.mean(axis=2, keepdims=True)
    size = sfft.next_fast_len(2 * n, real=True)
    f = sfft.rfft(y, n=size, axis=2, workers=-1)
    f *= np.conjugate(f)
    a = sfft.irfft(f, n=size, axis=2, workers=-1)[:, :, :n]
    for p in range(n_par):
        ok = a[p, :, 0] > 0
        if not ok.any():
            continue
        acf = (a[p, ok] / a[p, ok, :1]).sum(axis=0) / ok.sum()      # walkers of zero variance are left out
        taus = 2.0 * np.cumsum(acf) - 1.0
        m = np.arange(n) < c * taus
        win = int(np.argmin(m)) if m.any() else n - 1                # emcee.autocorr.auto_window, edge cases included
        tau[p] = taus[win]
    return tau


class ShardStream:
    """Host-side random numbers for the initial balls of a catalogue shard, vectorised over its sources and
    independent of the sharding: draw k of the source at catalogue row g is always the same number, whichever
    rank holds g.  Every call opens a fresh Philox stream keyed by (seed, call number) and skips the blocks of
    the rows before the shard (Philox.advance moves by whole 4-double blocks, so each source owns a whole number
    of blocks)."""

    def __init__(self, seed, first_source=0):
        self.seed, self.first, self.stage = int(seed), int(first_source), 0

    def _doubles(self, S, per_source):
        blocks = (per_source + 3) // 4
        bg = np.random.Philox(key=[self.seed & (2 ** 64 - 1), self.stage])
        self.stage += 1
        bg.advance(self.first * blocks)
        return np.random.Generator(bg).random((S, 4 * blocks))[:, :per_source]

    def uniform_block(self, S, W, n_par):
        return self._doubles(S, W * n_par).reshape(S, W, n_par)

    def normal_block(self, S, W, n_par):
        """Box-Muller: a fixed number of uniforms per normal (the ziggurat's count varies, which would shift every
        later source's numbers)."""
        u = self._doubles(S, 2 * W * n_par).reshape(S, W, n_par, 2)
        return np.sqrt(-2.0 * np.log1p(-u[..., 0])) * np.cos(2.0 * np.pi * u[..., 1])


class EnsembleSampler:
    def __init__(self, engine, walkers, seed=0, a=2.0, n_bins=1024, source_offset=0, split="random"):
        """engine: an Engine with sources uploaded.  walkers: even.  n_bins: histogram resolution of the
        on-device percentiles ((P.max - P.min) / n_bins per parameter); 0 = no histograms.  source_offset:
        catalogue row of the first uploaded source (enters the random-number counter: a shard draws what a
        single-GPU run draws for the same sources).  split: "random" — the walkers are re-divided into two equal
        sets at every step, what emcee 3's default move does (`RedBlueMove(randomize_split=True)`: the sampler the
        reference builds, MCMC_AGNfitter.py:114) — or "fixed": first half / second half."""
        if split not in ("random", "fixed"):
            raise ValueError("split must be 'random' or 'fixed'")
        self.engine = engine
        self.ctx, self.lib = engine.ctx, engine.lib
        self.S, self.W, self.n_par = engine.S, int(walkers), engine.n_par
        if self.S < 1:
            raise ValueError("upload sources first (Engine.set_sources)")
        self.n_bins = int(n_bins)
        self.ctx.check(self.lib.agnf_set_option(self.ctx.h, b"ens_source_offset", int(source_offset)))
        self.split = split
        self._set_split(split)
        self.ctx.check(self.lib.agnf_ens_create(self.ctx.h, self.W, int(seed) & (2 ** 64 - 1), float(a), self.n_bins))
        self._steps = 0

    def _set_split(self, split):
        # the option may only change while no ensemble exists: re-uploading the sources tears the old one down
        rc = self.lib.agnf_set_option(self.ctx.h, b"ens_split", 1 if split == "random" else 0)
        if rc != 0:
            raise ValueError("an ensemble with the other split already exists on this context; upload the sources "
                             "again (Engine.set_sources) before creating a sampler with split=%r" % split)

    # -- positions --------------------------------------------------------------------------------
    @staticmethod
    def initial_positions(pmin, pmax, walkers, rng):
        """mid-box +- 1e-5 per coordinate (PARAMETERSPACE:260-263), for every source: [S, W, n_par]."""
        pmin, pmax = np.atleast_2d(pmin), np.atleast_2d(pmax)
        u = rng.uniform(size=(pmin.shape[0], walkers, pmin.shape[1]))
        return 0.5 * (pmin + pmax)[:, None, :] + (2 * u - 1) * 0.00001

    @staticmethod
    def ball_around(best_pos, walkers, rng):
        """best + N(0,1) * 1e-5 (PARAMETERSPACE:287-292): [S, W, n_par]."""
        best_pos = np.atleast_2d(best_pos)
        return best_pos[:, None, :] + rng.normal(size=(best_pos.shape[0], walkers, best_pos.shape[1])) * 0.00001

    def set_positions(self, pos):
        pos = as_f64(pos)
        if pos.shape != (self.S, self.W, self.n_par):
            raise ValueError("positions must have shape (S, W, n_par) = %s" % ((self.S, self.W, self.n_par),))
        # acceptance counters and the running best are NOT reset: emcee's sampler accumulates over the burn-in sets
        # (MCMC_AGNfitter.py:125-128); reset_stats() / reset_best() do what sampler.reset() does (:182)
        self.ctx.check(self.lib.agnf_ens_set_positions(self.ctx.h, ptr(pos)))

    def set_ball(self, centre, scale=0.00001, normal=False, stream_id=0):
        """Positions drawn on the device around per-source centres [S, n_par]: centre + scale * uniform(-1, 1)
        (the initial ball) or centre + scale * N(0, 1) (the restart ball).  Nothing but the centres is uploaded."""
        centre = as_f64(centre)
        if centre.shape != (self.S, self.n_par):
            raise ValueError("centres must have shape (S, n_par) = %s" % ((self.S, self.n_par),))
        self.ctx.check(self.lib.agnf_ens_set_ball(self.ctx.h, ptr(centre), float(scale), int(bool(normal)),
                                                  int(stream_id) & 0xFFFFFFFF))

    # -- stepping ---------------------------------------------------------------------------------
    def run(self, n_steps, hist_every=0, track_best=False):
        """hist_every: every hist_every-th step of this call feeds the percentile histograms.  track_best: True
        (every step), 'thinned' (only at those steps) or False."""
        tb = 2 if track_best == 'thinned' else int(bool(track_best))
        self.ctx.check(self.lib.agnf_ens_run(self.ctx.h, int(n_steps), int(hist_every), tb))
        self._steps += int(n_steps)

    def reset_stats(self):
        self.ctx.check(self.lib.agnf_ens_reset_stats(self.ctx.h))
        self._steps = 0

    def reset_best(self):
        self.ctx.check(self.lib.agnf_ens_reset_best(self.ctx.h))

    def state(self):
        """-> pos [S, W, n_par], lnp [S, W], n_accepted [S, W]"""
        pos = np.empty((self.S, self.W, self.n_par))
        lnp = np.empty((self.S, self.W))
        acc = np.empty((self.S, self.W), dtype=np.int32)
        self.ctx.check(self.lib.agnf_ens_get(self.ctx.h, ptr(pos), ptr(lnp), ptr(acc)))
        return pos, lnp, acc

    def best(self):
        bp = np.empty((self.S, self.n_par))
        bl = np.empty(self.S)
        self.ctx.check(self.lib.agnf_ens_best(self.ctx.h, ptr(bp), ptr(bl)))
        return bp, bl

    def percentiles(self, q=REFERENCE_PERCENTILES):
        q = as_f64(q)
        out = np.empty((self.S, self.n_par, len(q)))
        self.ctx.check(self.lib.agnf_ens_percentiles(self.ctx.h, ptr(q), len(q), ptr(out)))
        return out

    def chain_percentiles(self, q=REFERENCE_PERCENTILES, first=0, stride=1, with_lnp=False):
        """EXACT percentiles (numpy.percentile, linear method) of the recorded chain entries first, first + stride,
        ... over all walkers: [S, n_par, n_q] (and [S, n_q] for the log-posterior with with_lnp=True).  On a
        recorded production run, first = int(n / 2), stride = Nthin reproduces what PLOTandWRITE_AGNfitter.py:162
        computes from CHAIN.props' thinned flat chain — on the device, by radix select, without a sort."""
        q = as_f64(q)
        out = np.empty((self.S, self.n_par, len(q)))
        lnp = np.empty((self.S, len(q))) if with_lnp else None
        self.ctx.check(self.lib.agnf_ens_chain_percentiles(self.ctx.h, int(first), int(stride), ptr(q), len(q),
                                                           ptr(out), ptr(lnp) if with_lnp else None))
        return (out, lnp) if with_lnp else out

    def record_state(self):
        """Append the current walker state to the recording buffer without stepping."""
        self.ctx.check(self.lib.agnf_ens_record_state(self.ctx.h))

    # -- chains (the reference's .sav files) ------------------------------------------------------
    def record(self, max_steps, every=1):
        """Keep the walker state of every `every`-th step from now on (device buffer of max_steps entries;
        0 stops recording).  For runs whose chains are wanted on disk — a handful of sources, not a catalogue:
        the buffer takes S * W * (n_par + 1) * 8 bytes per step."""
        self.ctx.check(self.lib.agnf_ens_record(self.ctx.h, int(max_steps), int(every)))

    def chain(self):
        """-> chain [S, W, n, n_par], lnprob [S, W, n] in emcee's axis order (sampler.chain / .lnprobability)."""
        n = C.c_int64()
        self.ctx.check(self.lib.agnf_ens_chain(self.ctx.h, None, None, C.byref(n)))
        n = int(n.value)
        ch = np.empty((n, self.S, self.W, self.n_par))
        lp = np.empty((n, self.S, self.W))
        self.ctx.check(self.lib.agnf_ens_chain(self.ctx.h, ptr(ch), ptr(lp), None))
        return np.ascontiguousarray(ch.transpose(1, 2, 0, 3)), np.ascontiguousarray(lp.transpose(1, 2, 0))

    def save_chains(self, filename, s=0, background=False, protocol=2):
        """Write source s the way MCMC_AGNfitter.save_chains does (MCMC_AGNfitter.py:192-206): a protocol-2
        pickle of dict(chain [W, n, n_par], accept [W], lnprob [W, n], final_pos [W, n_par], state, acor [n_par])
        that PARAMETERSPACE_AGNfitter.get_best_position (:268-294) and PLOTandWRITE's CHAIN.props (:471-492)
        read back.  `state` (emcee's NumPy RNG state) has no counterpart — the walk is driven by a counter-based
        generator — and is stored as None; `acor` is the integrated autocorrelation time (integrated_time).
        background=True: the chain is fetched now, the autocorrelation estimate and the file write run on a
        thread (returned; join it) while the sampler goes on — the library call releases the GIL.
        protocol: the reference writes protocol 2 (:203), which stores a 200 MB chain through a latin-1 string (about
        1 s per file); its readers (`pickle.load(f, encoding='latin1')`) accept any protocol, so 4 is a safe speed-up."""
        ch, lp = self.chain()
        pos, _, acc = self.state()
        payload = dict(chain=ch[s], accept=acc[s] / float(max(self._steps, 1)), lnprob=lp[s], final_pos=pos[s],
                       state=None)

        def write():
            import pickle
            payload["acor"] = integrated_time(payload["chain"])
            with open(filename, 'wb') as f:
                pickle.dump(payload, f, protocol=protocol)
        if not background:
            write()
            return None
        import threading
        t = threading.Thread(target=write)
        t.start()
        return t

    # -- the reference's schedule -----------------------------------------------------------------
    def run_reference_schedule(self, mc, pmin, pmax, rng, save_dir=None, names=None, exact_percentiles=False):
        """mc: the dict of MCMC_settings() (Nwalkers, Nburnsets, Nburn, Nmcmc, Nthin).  rng: a NumPy
        Generator, a list of S Generators (one per source), a ShardStream (vectorised), or None: the balls are
        drawn on the device (set_ball).  With all but the first the balls of a source do not depend on which other
        sources share the shard.  save_dir: also write `<save_dir>/<name>/samples_burn1-2-3.sav`
        and `samples_mcmc.sav` for every source, in the reference's format (MCMC_AGNfitter.py:156-206).
        exact_percentiles: keep the thinned second half of the production chain on the device (S * W * (n_par + 1) *
        8 bytes per kept step) and report numpy.percentile-exact values (chain_percentiles) instead of histogram
        quantiles; the log-posterior's percentiles come with them (`lnp_percentiles`).
        Returns a summary dict: percentiles [S, n_par, 5] and best position / lnP of the thinned second half of the
        production chain (what CHAIN.props keeps), mean acceptance fraction per source over the production run."""
        if int(mc['Nwalkers']) != self.W:
            raise ValueError("mc['Nwalkers'] != walkers")
        proto = int(mc.get('chain_pickle_protocol', 2))       # the reference's protocol; 4 writes a 200 MB chain ~1 s faster
        per_source = isinstance(rng, (list, tuple))
        pmin, pmax = np.atleast_2d(pmin), np.atleast_2d(pmax)

        def initial():
            if isinstance(rng, ShardStream):         # same arithmetic as initial_positions, vectorised over sources
                u = rng.uniform_block(self.S, self.W, self.n_par)
                return 0.5 * (pmin + pmax)[:, None, :] + (2 * u - 1) * 0.00001
            if not per_source:
                return self.initial_positions(pmin, pmax, self.W, rng)
            return np.concatenate([self.initial_positions(pmin[s], pmax[s], self.W, rng[s]) for s in range(self.S)])

        def ball(best_pos):
            if isinstance(rng, ShardStream):
                return best_pos[:, None, :] + rng.normal_block(self.S, self.W, self.n_par) * 0.00001
            if not per_source:
                return self.ball_around(best_pos, self.W, rng)
            return np.concatenate([self.ball_around(best_pos[s], self.W, rng[s]) for s in range(self.S)])

        if rng is None:                              # balls drawn on the device (catalogue scale)
            calls = [0]

            def device_ball(centre, normal):
                calls[0] += 1
                self.set_ball(centre, 0.00001, normal, stream_id=calls[0])
        nb, nburn, n = int(mc['Nburnsets']), int(mc['Nburn']), int(mc['Nmcmc'])
        writers = []
        if save_dir is not None:
            import os
            names = list(names) if names is not None else ["%d" % s for s in range(self.S)]
            for nm in names:
                os.makedirs(os.path.join(save_dir, str(nm)), exist_ok=True)
            self.record(max(nb * nburn, 1))         # emcee's sampler accumulates all burn-in sets (no reset between)
        self.reset_stats()
        self.reset_best()
        centres = []
        if rng is None:
            device_ball(0.5 * (pmin + pmax), False)
        else:
            self.set_positions(initial())
        for k_set in range(nb):
            # the sampler keeps accumulating over the sets (no reset, MCMC_AGNfitter.py:125-128): the restart centre is
            # the best sample of ALL burn-in steps so far — what get_best_position reads back from the burn-in file
            self.run(nburn, track_best=True)
            best_pos, _ = self.best()
            centres.append(best_pos)
            # run_burn_in rewrites the burn-in file after every set (MCMC:171); only the last version survives,
            # and the restart point it is read back for (get_best_position) is tracked on the device
            if save_dir is not None and k_set == nb - 1:     # written while the production run goes on
                writers = [self.save_chains(os.path.join(save_dir, str(nm), 'samples_burn1-2-3.sav'), s_, True, proto)
                           for s_, nm in enumerate(names)]
            if rng is None:
                device_ball(best_pos, True)
            else:
                self.set_positions(ball(best_pos))
        self.reset_stats()
        if save_dir is not None:
            self.record(max(n, 1))                  # sampler.reset() (MCMC:182)
        # PLOTandWRITE's CHAIN.props keeps chain[:, int(n/2):n:Nthin] (PLOTandWRITE_AGNfitter.py:476-487): the second
        # half of the production run, every Nthin-th state, for the percentiles AND for the best-fit row
        half, nt = int(n / 2), max(int(mc.get('Nthin', 1)), 1)
        self.run(half)
        self.reset_best()
        thinned_only = exact_percentiles and save_dir is None and n - half > 0
        if thinned_only:
            self.record(len(range(half, n, nt)), every=nt)      # exactly chain[:, half:n:Nthin]
        if n - half > 0:
            # hist_every marks the thinned steps: they feed the histograms (if any) and, with track_best='thinned',
            # the best-fit row — the reference picks it from the thinned flat chain (PLOTandWRITE:483-487)
            mark = bool(self.n_bins) or bool(exact_percentiles)
            tb = 'thinned' if mark else True
            self.run(1, hist_every=1 if mark else 0, track_best=tb)
            self.run(n - half - 1, hist_every=nt if mark else 0, track_best=tb)
        exact = lnp_pct = None
        if exact_percentiles and n - half > 0:
            exact, lnp_pct = self.chain_percentiles(first=0 if thinned_only else half, stride=1 if thinned_only else nt,
                                                    with_lnp=True)
            if thinned_only:
                self.record(0)
        if save_dir is not None:
            for s_, nm in enumerate(names):
                self.save_chains(os.path.join(save_dir, str(nm), 'samples_mcmc.sav'), s_, protocol=proto)
            self.record(0)
            for t in writers:
                t.join()
        _, _, acc = self.state()
        bp, bl = self.best()
        pct = exact if exact is not None else (self.percentiles() if self.n_bins else None)
        return dict(percentiles=pct, percentiles_exact=exact is not None, lnp_percentiles=lnp_pct, best_pos=bp,
                    best_lnp=bl, acceptance=acc.mean(axis=1) / max(n, 1), restart_centres=centres)
