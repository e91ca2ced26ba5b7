"""needs_reference: integration/agnfitter_b200_backend.patch — the change a maintainer makes on the reference side
 — applies cleanly to a copy of the reference, and the patched MCMC_AGNfitter.main routes
through the drop-in exactly where the patch says: `mc['backend'] = 'b200'` -> dropin.main for the emcee leg,
Posterior.my_posterior / my_params_transform / ln_probab_batch (vectorize=True) otherwise; without the key the
reference path is untouched."""
import importlib
import os
import shutil
import subprocess
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.needs_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "agnfitter_b200_backend.patch")
REF = "/root/reference"


@pytest.fixture()
def patched_tree(tmp_path):
    os.makedirs(tmp_path / "functions")
    for rel in ("RUN_AGNfitter_multi.py", "functions/MCMC_AGNfitter.py", "functions/PARAMETERSPACE_AGNfitter.py",
                "functions/PRIORS_AGNfitter.py"):
        shutil.copy(os.path.join(REF, rel), tmp_path / rel)
    chk = subprocess.run(["git", "apply", "--check", "--verbose", PATCH], cwd=tmp_path, capture_output=True, text=True)
    assert chk.returncode == 0, chk.stderr
    subprocess.run(["git", "apply", PATCH], cwd=tmp_path, check=True)
    return tmp_path


def test_patch_applies_and_touches_only_the_two_call_sites(patched_tree):
    for rel, needle in (("functions/MCMC_AGNfitter.py", "mc.get('backend') == 'b200'"),
                        ("RUN_AGNfitter_multi.py", "b200.build(zdict)"),
                        ("RUN_AGNfitter_multi.py", "b200.save_dictionary_cache(zdict)")):
        new = open(patched_tree / rel).read()
        old = open(os.path.join(REF, rel)).read()
        assert needle in new and needle not in old
        removed = [l for l in old.splitlines() if l.strip() and l not in new.splitlines()]
        assert len(removed) <= 2, removed          # one re-indented line per file: nothing of the reference is dropped
    compile(open(patched_tree / "RUN_AGNfitter_multi.py").read(), "RUN_AGNfitter_multi.py", "exec")


def test_patched_main_routes_through_the_dropin(patched_tree, monkeypatch):
    from oracle import ref_shim
    ref_shim.install()
    calls = []

    class StubSampler:
        def __init__(self, nwalkers, npar, fn, **kw):
            calls.append(("EnsembleSampler", fn, kw))
            raise RuntimeError("stop here")

    emcee = types.ModuleType("emcee")
    emcee.__file__ = str(patched_tree / "emcee" / "__init__.py")
    os.makedirs(patched_tree / "emcee")
    open(patched_tree / "emcee" / "autocorr.py", "w").write("def integrated_time(x, c=5, tol=50, quiet=False):\n    pass\n")
    emcee.EnsembleSampler = StubSampler
    emcee.autocorr = types.ModuleType("emcee.autocorr")
    ultra = types.ModuleType("ultranest")
    ultra.ReactiveNestedSampler = ultra.stepsampler = ultra.dychmc = ultra.popstepsampler = object
    for name, mod in (("emcee", emcee), ("emcee.autocorr", emcee.autocorr), ("ultranest", ultra),
                      ("ultranest.plot", types.SimpleNamespace(cornerplot=None, traceplot=None)),
                      ("ultranest.stepsampler", types.SimpleNamespace(OrthogonalDirectionGenerator=None))):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.setattr(importlib, "reload", lambda m: m)
    # the patched module, imported as part of the shimmed `functions` package but from the patched copy
    pkg = sys.modules["functions"]
    monkeypatch.setattr(pkg, "__path__", [str(patched_tree / "functions")] + list(pkg.__path__))
    monkeypatch.delitem(sys.modules, "functions.MCMC_AGNfitter", raising=False)
    M = importlib.import_module("functions.MCMC_AGNfitter")
    assert M.__file__.startswith(str(patched_tree))

    from agnfitter_b200 import dropin

    class StubPosterior:
        def __init__(self, data, models, P):
            calls.append(("Posterior", data, models, P))

        def ln_probab_batch(self, params):
            return np.zeros(len(params))

        def my_posterior(self, params):
            calls.append(("my_posterior", params.shape))
            return np.zeros(len(params))

        def my_params_transform(self, cube):
            return cube

    monkeypatch.setattr(dropin, "Posterior", StubPosterior)
    monkeypatch.setattr(dropin, "main", lambda data, models, P, mc: calls.append(("dropin.main", mc)) or "gpu")
    data = types.SimpleNamespace(output_folder=str(patched_tree) + "/", name="src")
    P = dict(names=["a", "b"], min=[0., 0.], max=[1., 1.])
    mc = dict(sampling_algorithm="emcee", Nwalkers=8, Nburn=0, Nmcmc=0, backend="b200")
    assert M.main(data, "models", P, mc) == "gpu" and calls[-1][0] == "dropin.main"
    calls.clear()
    with pytest.raises(RuntimeError, match="stop here"):
        M.main(data, "models", P, dict(mc, b200_sampler=False))
    assert calls[0][0] == "Posterior" and calls[1][0] == "EnsembleSampler"
    assert calls[1][2] == dict(vectorize=True) and calls[1][1].__name__ == "ln_probab_batch"
    calls.clear()
    with pytest.raises(RuntimeError, match="stop here"):
        M.main(data, "models", P, dict(sampling_algorithm="emcee", Nwalkers=8, Nburn=0, Nmcmc=0))
    assert calls[0][0] == "EnsembleSampler" and calls[0][2] == dict(args=[data, "models", P])   # the reference's own call
