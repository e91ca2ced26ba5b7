"""
ctypes binding of libagnfit_b200.so (include/agnfit_b200.h).  This is the whole FFI layer: plain
pointers and sizes, no torch types.  There is NO CPU fallback — if the shared library is missing
or no sm_100 device is present, loading / context creation raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# AGNFIT_B200_LIB: alternative build of the same library (kernel-tuning experiments only)
LIB_PATH = os.environ.get("AGNFIT_B200_LIB") or os.path.join(_HERE, "csrc", "libagnfit_b200.so")

AGNF_MAX_FAMILIES = 5
AGNF_MAX_DIMS = 4
AGNF_MAX_BANDS = 64
AGNF_MAX_PARS = 32


class AgnfError(RuntimeError):
    pass


class agnf_config(C.Structure):
    _fields_ = [("n_b", C.c_int32), ("n_z", C.c_int32), ("n_par", C.c_int32), ("n_fam", C.c_int32),
                ("bbb_has_norm", C.c_int32), ("radio", C.c_int32), ("eb_mode", C.c_int32),
                ("prior_agnfrac", C.c_int32), ("ga_free_ebv", C.c_int32), ("bb_mode", C.c_int32)]


class agnf_family_desc(C.Structure):
    _fields_ = [("fam", C.c_int32), ("n_dim", C.c_int32), ("par_offset", C.c_int32),
                ("table_on_device", C.c_int32), ("T", C.c_int64),
                ("dim_is_grid", C.c_void_p), ("dim_size", C.c_void_p), ("dim_stride", C.c_void_p),
                ("grid_sorted", C.c_void_p), ("grid_perm", C.c_void_p), ("table", C.c_void_p),
                ("aux", C.c_void_p), ("table_pitch", C.c_int32), ("reserved", C.c_int32)]


class agnf_extra_priors(C.Structure):
    _fields_ = [("midir_uv", C.c_int32), ("ir_xray", C.c_int32), ("j_6um", C.c_int32), ("j_2500", C.c_int32),
                ("bb_dered_row", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/agnfit_b200.h declares
SIGNATURES = {
    "agnf_version": (C.c_int, []),
    "agnf_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "agnf_ctx_destroy": (None, [C.c_void_p]),
    "agnf_last_error": (C.c_char_p, [C.c_void_p]),
    "agnf_synchronize": (C.c_int, [C.c_void_p]),
    "agnf_modelset_config": (C.c_int, [C.c_void_p, C.POINTER(agnf_config)]),
    "agnf_family_upload": (C.c_int, [C.c_void_p, C.POINTER(agnf_family_desc)]),
    "agnf_extra_priors_set": (C.c_int, [C.c_void_p, C.POINTER(agnf_extra_priors)]),
    "agnf_zslices_upload": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "agnf_sources_upload": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 7 + [C.c_int32] + [C.c_void_p] * 3),
    "agnf_lnprob_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 5),
    "agnf_lnprob_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 6),
    "agnf_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "agnf_get_option": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "agnf_debug_exp10": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "agnf_debug_fp64_peak": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "agnf_launch_count": (C.c_int64, [C.c_void_p]),
    "agnf_last_kernel_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "agnf_filters_upload": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnf_dict_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32]),
    "agnf_ens_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_double, C.c_int32]),
    "agnf_ens_set_positions": (C.c_int, [C.c_void_p, C.c_void_p]),
    "agnf_ens_run": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32]),
    "agnf_ens_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnf_ens_best": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnf_ens_reset_stats": (C.c_int, [C.c_void_p]),
    "agnf_ens_set_ball": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_uint32]),
    "agnf_ens_reset_best": (C.c_int, [C.c_void_p]),
    "agnf_ens_record": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32]),
    "agnf_ens_chain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "agnf_ens_percentiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "agnf_ens_chain_percentiles": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                             C.c_void_p]),
    "agnf_ens_record_state": (C.c_int, [C.c_void_p]),
    "agnf_sed_config": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnf_sed_family_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "agnf_sed_sources": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "agnf_sed_bands": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnf_sed_intlum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "agnf_sed_spectra": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "agnf_sed_intlum_ens": (C.c_int, [C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """dlopen the in-tree library (built by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AgnfError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "or `make -C agnfitter_b200/csrc`; there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.agnf_version() != 1:
            raise AgnfError("libagnfit_b200.so ABI version %d, expected 1" % lib.agnf_version())
        _lib = lib
    return _lib


def ptr(a):
    """Host pointer of a C-contiguous NumPy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def as_u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Context:
    """RAII wrapper of agnf_ctx.  One per (process, device); create after fork."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        if self.lib.agnf_ctx_create(int(device), C.byref(h)) != 0:
            raise AgnfError(self.lib.agnf_last_error(None).decode())
        self.h = h
        self.device = int(device)

    def check(self, rc):
        if rc != 0:
            raise AgnfError(self.lib.agnf_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.agnf_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
