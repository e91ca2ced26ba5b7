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
  3. inert stub modules for astropy.* (MODEL:24-32, FILTERS:22-24);
  4. `pickle` as MODEL_AGNfitter sees it falls back to a raw unpickler for the one
     pandas pickle pandas 3.0 refuses (models/BBB/THB21.pickle, MODEL:859: BlockPlacement
     TypeError) and hands back its 2 x 1 object block behind the two accessors the
     loader uses (`frame[col].values.item()`, MODEL:860).
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


class _Column:
    def __init__(self, arr):
        self._arr = arr

    @property
    def values(self):
        import numpy as np
        out = np.empty(1, dtype=object)
        out[0] = self._arr
        return out


class _RawFrame(dict):
    """{'nu': _Column, 'SED': _Column}: what THB21dict['nu'].values.item() needs (MODEL:860)."""


class _Cap:
    def __new__(cls, *a, **k):
        o = object.__new__(cls)
        o.a, o.k = a, k
        return o

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.state = s

    def __call__(self, *a, **k):
        return _Cap(*a, **k)


def raw_pandas_frame(fileobj):
    """Columns of a pickled single-block object DataFrame without pandas: the BlockManager's arguments are
    captured by stand-in classes; block values [n_col, n_row] and the column Index data are plain ndarrays."""
    import pickle

    class U(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith("pandas"):
                return type(name, (_Cap,), {})
            if module.startswith("numpy.core"):
                module = module.replace("numpy.core", "numpy._core")
            if module == "__builtin__":
                module = "builtins"
            return super().find_class(module, name)

    d = U(fileobj, encoding="latin1").load()
    mgr = d.state["_mgr"]
    blocks, axes = mgr.a
    cols = [str(c) for c in axes[0].a[1]["data"]]
    if len(blocks) != 1:
        raise ValueError("expected a single-block frame")
    vals = blocks[0].a[0]
    if vals.shape[0] != len(cols) or vals.shape[1] != 1:
        raise ValueError("unexpected block shape %r" % (vals.shape,))
    return _RawFrame({c: _Column(vals[i, 0]) for i, c in enumerate(cols)})


class _PickleProxy:
    """`pickle` for the reference's MODEL_AGNfitter: identical except that a load pandas rejects is redone raw."""

    def __getattr__(self, k):
        import pickle
        return getattr(pickle, k)

    def load(self, f, *a, **k):
        import pickle
        try:
            return pickle.load(f, *a, **k)
        except TypeError:
            f.seek(0)
            return raw_pandas_frame(f)


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
    if not isinstance(mods["MOD"].pickle, _PickleProxy):
        mods["MOD"].pickle = _PickleProxy()
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
