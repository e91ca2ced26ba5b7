"""
TEST INFRASTRUCTURE — not part of the product path.

Import shims that let the UNMODIFIED reference hot-path modules under
/root/reference/functions run in this container (Python 3.12, numpy 2.x,
scipy 1.18, no astropy/emcee/ultranest/matplotlib).  Only used

  * by oracle/make_golden.py to generate tests/golden/*.npz, and
  * by tests marked `needs_reference` (skipped when /root/reference is absent,
    i.e. on the GPU box).

Nothing here is copied from the reference; it only prepares `sys.modules` so
that `functions.PARAMETERSPACE_AGNfitter`, `functions.PRIORS_AGNfitter`,
`functions.DICTIONARIES_AGNfitter` and `functions.MODEL_AGNfitter` import
(SURVEY.md §8c recipe):

  1. an empty package object `functions` with __path__ pointing at the
     reference (bypasses functions/__init__.py:9 `from acor import *`);
  2. removed-API aliases: collections.Iterable (PARAMETERSPACE:18),
     scipy.integrate.trapz (DICTIONARIES:21), np.float (DATA:103);
  3. inert stub modules for astropy.* (MODEL:24-32, FILTERS:22-24).
"""
import os
import sys
import types
import collections
import collections.abc
import importlib

REF_ROOT = os.environ.get("AGNFITTER_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "functions"))


class _Inert:
    """Stands in for astropy unit objects: supports the arithmetic the
    reference performs at import time / in settings files, nothing more."""

    def __init__(self, name="unit"):
        self._name = name

    def __mul__(self, o):
        return self

    __rmul__ = __mul__
    __truediv__ = __mul__
    __rtruediv__ = __mul__
    __pow__ = __mul__

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Inert(self._name + "." + k)

    def __repr__(self):
        return "<inert %s>" % self._name


def _stub_module(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda k, _n=name: _Inert(_n + "." + k)  # PEP 562
    return m


_installed = False


def install():
    """Idempotently install the shims and return the dict of reference modules."""
    global _installed
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REF_ROOT)
    if not _installed:
        import numpy as np
        import scipy.integrate

        if not hasattr(collections, "Iterable"):
            collections.Iterable = collections.abc.Iterable
        if not hasattr(scipy.integrate, "trapz"):
            scipy.integrate.trapz = scipy.integrate.trapezoid
        if not hasattr(np, "float"):
            np.float = float

        pkg = types.ModuleType("functions")
        pkg.__path__ = [os.path.join(REF_ROOT, "functions")]
        sys.modules["functions"] = pkg
        for name in ("astropy", "astropy.units", "astropy.table", "astropy.constants",
                     "astropy.cosmology", "astropy.io", "astropy.io.ascii",
                     "astropy.units.quantity"):
            if name not in sys.modules:
                sys.modules[name] = _stub_module(name)
        sys.modules["astropy"].units = sys.modules["astropy.units"]
        sys.modules["astropy"].constants = sys.modules["astropy.constants"]
        _installed = True

    mods = {}
    for short, name in (("MOD", "functions.MODEL_AGNfitter"),
                        ("DIC", "functions.DICTIONARIES_AGNfitter"),
                        ("PRI", "functions.PRIORS_AGNfitter"),
                        ("PAR", "functions.PARAMETERSPACE_AGNfitter")):
        mods[short] = importlib.import_module(name)
    return mods


def install_plot():
    """Also import PLOTandWRITE_AGNfitter (the SED-realisation / integrated-luminosity stage).
    matplotlib is absent here; the module only needs it at import time and inside plotting methods."""
    install()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.colors", "matplotlib.cm",
                 "matplotlib.patches", "matplotlib.gridspec", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf"):
        if name not in sys.modules:
            sys.modules[name] = _stub_module(name)
    sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    return importlib.import_module("functions.PLOTandWRITE_AGNfitter")
