"""
Coalescing of many samplers' log-posterior calls into one K1 launch per tick.

ultranest (MCMC_AGNfitter.py:74-96) drives ONE source and hands `my_posterior` a few hundred points per call
(`vectorized=True`); a call of that size is pure launch latency on a B200 (bench.py `small_batch`).  A catalogue is
fitted by running one such sampler per source concurrently — threads here; the samplers are Python loops that spend
their time waiting for the posterior — and letting a dispatcher merge whatever requests are pending into a single
`agnf_lnprob_batch` call with the rows' `source_id` filled in (SURVEY.md §8b `ln_probab_batch`):

    cat = CatalogPosterior(ms, z, dlum, nus, fluxes, fluxerrs)          # whole shard resident on the GPU
    with PosteriorMultiplexer(cat) as mux:
        def fit(s):
            with mux.client(s) as c:                                      # one per source / thread
                sampler = ultranest.ReactiveNestedSampler(names, c.my_posterior, c.my_params_transform,
                                                          vectorized=True, ...)
                sampler.run(...)
        threads = [threading.Thread(target=fit, args=(s,)) for s in range(cat.src.S)]

K1's result for a row does not depend on what else is in the batch (tests/test_k1_gpu.py: split / permutation
invariance), so a multiplexed value equals the value of a per-source call bit for bit
(tests/test_multiplex_gpu.py).  The dispatcher launches as soon as every open client has a request pending, or after
`max_wait_s` if some client is busy elsewhere, so a lone straggler is never held up.
"""
import threading
import time

import numpy as np

from .dropin import params_transform


class _Request:
    __slots__ = ("sid", "params", "result", "error", "done")

    def __init__(self, sid, params):
        self.sid = sid
        self.params = params
        self.result = None
        self.error = None
        self.done = threading.Event()


class MuxClient:
    """The callables one sampler binds to (reference names): my_posterior, my_params_transform, ln_probab_batch."""

    def __init__(self, mux, source):
        self._mux = mux
        self.source = int(source)
        self._open = True

    def my_posterior(self, params):
        """[n, n_par] -> [n] like MCMC_AGNfitter.my_posterior (:56-62); blocks until the merged launch returns."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p[None, :]
        if not self._open:
            raise RuntimeError("this client has been closed")
        return self._mux._submit(self.source, p)

    ln_probab_batch = my_posterior

    def my_params_transform(self, cube):
        """MCMC_AGNfitter.my_params_transform (:64-72) for this source's prior box."""
        return params_transform(cube, self._mux.pmin[self.source], self._mux.pmax[self.source])

    def close(self):
        if self._open:
            self._open = False
            self._mux._client_closed()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class PosteriorMultiplexer:
    """`evaluate(params [n, n_par], source_id [n]) -> lnp [n]` is the merged call: CatalogPosterior.ln_probab_batch in
    production, any callable with that signature in tests."""

    def __init__(self, catalog_posterior=None, evaluate=None, pmin=None, pmax=None, max_wait_s=2e-4):
        if catalog_posterior is not None:
            evaluate = catalog_posterior.ln_probab_batch
            pmin, pmax = catalog_posterior.pmin, catalog_posterior.pmax
        if evaluate is None:
            raise ValueError("need a CatalogPosterior or an evaluate callable")
        self._evaluate = evaluate
        self.pmin = None if pmin is None else np.asarray(pmin, dtype=np.float64)
        self.pmax = None if pmax is None else np.asarray(pmax, dtype=np.float64)
        self.max_wait_s = float(max_wait_s)
        self._cond = threading.Condition()
        self._pending = []
        self._n_open = 0
        self._stop = False
        self.launches = 0          # merged calls issued
        self.requests = 0          # client calls served
        self.rows = 0              # rows evaluated
        self._thread = threading.Thread(target=self._run, name="agnf-mux", daemon=True)
        self._thread.start()

    # ---- clients ---------------------------------------------------------------------------------------
    def client(self, source):
        with self._cond:
            if self._stop:
                raise RuntimeError("the multiplexer has been closed")
            self._n_open += 1
        return MuxClient(self, source)

    def _client_closed(self):
        with self._cond:
            self._n_open -= 1
            self._cond.notify_all()

    def _submit(self, sid, params):
        req = _Request(sid, params)
        with self._cond:
            if self._stop:
                raise RuntimeError("the multiplexer has been closed")
            self._pending.append(req)
            self._cond.notify_all()
        req.done.wait()
        if req.error is not None:
            raise req.error
        return req.result

    def evaluate_many(self, requests):
        """Synchronous form for callers that do their own scheduling: [(source, params [n_i, n_par]), ...] ->
        [lnp [n_i], ...] through ONE merged launch."""
        reqs = [_Request(int(s), np.ascontiguousarray(p, dtype=np.float64).reshape(-1, np.shape(p)[-1]))
                for s, p in requests]
        self._serve(reqs)
        for r in reqs:
            if r.error is not None:
                raise r.error
        return [r.result for r in reqs]

    # ---- dispatcher ------------------------------------------------------------------------------------
    def _serve(self, batch):
        try:
            sizes = [len(r.params) for r in batch]
            params = batch[0].params if len(batch) == 1 else np.concatenate([r.params for r in batch], axis=0)
            sid = np.repeat(np.array([r.sid for r in batch], dtype=np.int32), sizes)
            lnp = np.asarray(self._evaluate(params, sid))
            o = 0
            for r, n in zip(batch, sizes):
                r.result = lnp[o:o + n].copy()
                o += n
            self.launches += 1
            self.requests += len(batch)
            self.rows += int(o)
        except BaseException as e:          # hand the failure to every waiting sampler instead of hanging them
            for r in batch:
                r.error = e
        finally:
            for r in batch:
                r.done.set()

    def _run(self):
        while True:
            with self._cond:
                while not self._pending and not self._stop:
                    self._cond.wait()
                if self._stop and not self._pending:
                    return
                # merge: wait (briefly) for the other open clients' requests
                deadline = time.perf_counter() + self.max_wait_s
                while len(self._pending) < self._n_open and not self._stop:
                    left = deadline - time.perf_counter()
                    if left <= 0:
                        break
                    self._cond.wait(left)
                batch, self._pending = self._pending, []
            if batch:
                self._serve(batch)

    def close(self):
        with self._cond:
            self._stop = True
            self._cond.notify_all()
        self._thread.join(timeout=10)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
