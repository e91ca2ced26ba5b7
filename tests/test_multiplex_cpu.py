"""CPU: the request-merging logic of PosteriorMultiplexer (threads, batching, error propagation) against a stand-in
evaluate callable; the GPU twin (tests/test_multiplex_gpu.py) checks bit equality with per-source K1 calls."""
import threading

import numpy as np
import pytest

from agnfitter_b200.multiplex import PosteriorMultiplexer


def _fake_eval(params, sid):
    return params.sum(axis=1) * (1.0 + sid)


def test_merged_calls_return_each_clients_rows():
    S, n_par = 24, 7
    pmin, pmax = np.zeros((S, n_par)), np.ones((S, n_par)) * np.arange(1, S + 1)[:, None]
    seen = []

    def evaluate(params, sid):
        seen.append(len(np.unique(sid)))
        return _fake_eval(params, sid)

    out = {}
    with PosteriorMultiplexer(evaluate=evaluate, pmin=pmin, pmax=pmax, max_wait_s=0.05) as mux:
        def work(s):
            rng = np.random.default_rng(s)
            got = []
            with mux.client(s) as c:
                for k in range(12):
                    cube = rng.uniform(size=(int(rng.integers(1, 40)), n_par))
                    p = c.my_params_transform(cube)
                    assert np.array_equal(p, (pmax[s] - pmin[s]) * cube + pmin[s])
                    got.append((p, c.my_posterior(p)))
            out[s] = got
        th = [threading.Thread(target=work, args=(s,)) for s in range(S)]
        [t.start() for t in th]
        [t.join(timeout=60) for t in th]
        assert mux.requests == S * 12 and mux.launches < mux.requests
        assert mux.rows == sum(len(p) for g in out.values() for p, _ in g)
    for s, got in out.items():
        for p, l in got:
            assert np.array_equal(l, _fake_eval(p, np.full(len(p), s)))
    assert max(seen) > 1, "requests of different sources must have shared a call"


def test_evaluate_many_and_errors():
    mux = PosteriorMultiplexer(evaluate=_fake_eval, pmin=np.zeros((3, 2)), pmax=np.ones((3, 2)))
    a, b = np.ones((4, 2)), np.full((2, 2), 3.0)
    ra, rb = mux.evaluate_many([(0, a), (2, b)])
    assert np.array_equal(ra, _fake_eval(a, np.zeros(4))) and np.array_equal(rb, _fake_eval(b, np.full(2, 2)))
    assert mux.launches == 1
    mux.close()
    with pytest.raises(RuntimeError):
        mux.client(0)

    def boom(params, sid):
        raise ValueError("kernel failed")
    with PosteriorMultiplexer(evaluate=boom, pmin=np.zeros((1, 2)), pmax=np.ones((1, 2))) as m2:
        with m2.client(0) as c:
            with pytest.raises(ValueError, match="kernel failed"):      # the sampler thread sees the failure
                c.my_posterior(np.zeros((3, 2)))


def test_a_lone_client_is_not_held_up():
    """Two clients open, only one calls: the dispatcher launches after max_wait_s instead of waiting for the other."""
    with PosteriorMultiplexer(evaluate=_fake_eval, pmin=np.zeros((2, 2)), pmax=np.ones((2, 2)), max_wait_s=0.01) as mux:
        idle = mux.client(1)
        with mux.client(0) as c:
            assert np.array_equal(c.my_posterior(np.ones((5, 2))), _fake_eval(np.ones((5, 2)), np.zeros(5)))
        idle.close()
