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


class EnsembleSampler:
    def __init__(self, engine, walkers, seed=0, a=2.0, n_bins=1024, source_offset=0):
        """engine: an Engine with sources uploaded.  walkers: even.  n_bins: histogram resolution of the
        on-device percentiles ((P.max - P.min) / n_bins per parameter); 0 = no histograms.  source_offset:
        catalogue row of the first uploaded source (enters the random-number counter: a shard draws what a
        single-GPU run draws for the same sources)."""
        self.engine = engine
        self.ctx, self.lib = engine.ctx, engine.lib
        self.S, self.W, self.n_par = engine.S, int(walkers), engine.n_par
        if self.S < 1:
            raise ValueError("upload sources first (Engine.set_sources)")
        self.n_bins = int(n_bins)
        self.ctx.check(self.lib.agnf_set_option(self.ctx.h, b"ens_source_offset", int(source_offset)))
        self.ctx.check(self.lib.agnf_ens_create(self.ctx.h, self.W, int(seed) & (2 ** 64 - 1), float(a), self.n_bins))

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
        self.ctx.check(self.lib.agnf_ens_set_positions(self.ctx.h, ptr(pos)))

    # -- stepping ---------------------------------------------------------------------------------
    def run(self, n_steps, hist_every=0, track_best=False):
        self.ctx.check(self.lib.agnf_ens_run(self.ctx.h, int(n_steps), int(hist_every), int(bool(track_best))))

    def reset_stats(self):
        self.ctx.check(self.lib.agnf_ens_reset_stats(self.ctx.h))

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

    # -- the reference's schedule -----------------------------------------------------------------
    def run_reference_schedule(self, mc, pmin, pmax, rng):
        """mc: the dict of MCMC_settings() (Nwalkers, Nburnsets, Nburn, Nmcmc, Nthin).  rng: a NumPy
        Generator, or a list of S Generators (one per source: the initial balls then do not depend on
        which other sources share the shard).  Returns a summary dict: percentiles [S, n_par, 5] of the
        thinned production chain, best position / lnP, mean acceptance fraction per source."""
        if int(mc['Nwalkers']) != self.W:
            raise ValueError("mc['Nwalkers'] != walkers")
        per_source = isinstance(rng, (list, tuple))

        def initial():
            if not per_source:
                return self.initial_positions(pmin, pmax, self.W, rng)
            return np.concatenate([self.initial_positions(pmin[s], pmax[s], self.W, rng[s]) for s in range(self.S)])

        def ball(best_pos):
            if not per_source:
                return self.ball_around(best_pos, self.W, rng)
            return np.concatenate([self.ball_around(best_pos[s], self.W, rng[s]) for s in range(self.S)])

        self.set_positions(initial())
        for _ in range(int(mc['Nburnsets'])):
            self.run(int(mc['Nburn']), track_best=True)
            best_pos, _ = self.best()
            self.set_positions(ball(best_pos))
        self.reset_stats()
        n = int(mc['Nmcmc'])
        self.run(n, hist_every=int(mc.get('Nthin', 1)) if self.n_bins else 0, track_best=True)
        _, _, acc = self.state()
        bp, bl = self.best()
        return dict(percentiles=self.percentiles() if self.n_bins else None, best_pos=bp, best_lnp=bl,
                    acceptance=acc.mean(axis=1) / max(n, 1))
