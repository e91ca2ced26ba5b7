"""CPU: the C-ABI library loads and exports every symbol include/agnfit_b200.h declares;
without a GPU the product fails loudly instead of falling back."""
import os
import re

import pytest

from agnfitter_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "agnfit_b200.h")).read()
    return sorted(set(re.findall(r"AGNF_API\s+[\w\s\*]+?\b(agnf_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.SIGNATURES.keys())


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.agnf_version() == 1


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(_lib.AgnfError, match="no CUDA device|CPU path"):
        _lib.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "agnfitter_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_no_exception_can_cross_the_abi():
    """Every int-returning extern "C" entry point runs inside the try/catch guard that turns a C++ exception
    (std::bad_alloc from a host container, ...) into the error code + agnf_last_error convention."""
    src = open(os.path.join(ROOT, "agnfitter_b200", "csrc", "agnfit_capi.cu")).read()
    bodies = re.findall(r"^AGNF_API int (agnf_\w+)\([^{]*\{\n(.*?)^\}", src, flags=re.M | re.S)
    assert len(bodies) >= 30
    for name, body in bodies:
        assert body.lstrip().startswith("AGNF_GUARD_BEGIN") and "AGNF_GUARD_END(" in body.rstrip().splitlines()[-1], name
    int_decls = [n for n in _declared_symbols() if re.search(r"AGNF_API\s+int\s+%s\s*\(" % n,
                                                             open(os.path.join(ROOT, "include", "agnfit_b200.h")).read())]
    guarded = {n for n, _ in bodies} | {"agnf_version"}
    assert set(int_decls) <= guarded, sorted(set(int_decls) - guarded)
