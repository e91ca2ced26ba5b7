"""
Host-side driver of the two CUDA kernels behind the C ABI.

`Engine` holds one ModelSet (dictionary for a z grid) and one SourceSet (a catalogue shard) on a
B200 and evaluates ln_probab for arbitrary (source, walker) batches — the batched counterpart of
PARAMETERSPACE_AGNfitter.ln_probab (PARAMETERSPACE_AGNfitter.py:153-173).  `DictBuilder` is the
batched counterpart of MODELSDICT.build / filtering_models (DICTIONARIES_AGNfitter.py:78-154,
426-478).  PyTorch appears only as the owner of optional device buffers; all compute is in
libagnfit_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Context, as_f64, as_i32, as_i64, as_u64, ptr
from .flatten import ModelSet, FAMILY_ORDER
from .sources import SourceSet, ZSlices, make_zslices, extra_prior_plan, C_NCONST

_EB_MODE = {'Flexible': 1, 'Restrictive': 2}


class Engine:
    def __init__(self, ms: ModelSet, device=0, ctx=None, device_tables=None, zslices: ZSlices = None):
        """device_tables: optional {family name: torch CUDA tensor [n_z, T, n_b] float64} — e.g. the
        output of DictBuilder.build_family(..., on_device=True); used in place of ms.fam(name).table."""
        self.ms = ms
        self.ctx = ctx if ctx is not None else Context(device)
        self.lib = self.ctx.lib
        self._keep = []     # device tensors that must outlive the ctx's use of them
        self.n_b, self.n_par, self.n_fam = ms.n_b, ms.n_par, len(ms.families)
        cfg = _lib.agnf_config(
            n_b=ms.n_b, n_z=ms.n_z, n_par=ms.n_par, n_fam=len(ms.families),
            bbb_has_norm=int(ms.bbb_has_norm), radio=int(ms.radio),
            eb_mode=_EB_MODE.get(ms.settings.get('PRIOR_energy_balance'), 0),
            prior_agnfrac=int(ms.settings.get('PRIOR_AGNfraction') == True),
            ga_free_ebv=int(ms.ga_free_ebv), bb_mode=int(ms.bb_mode))
        # the two branches the reference itself cannot evaluate: prior_low_AGNfraction unpacks 4 of
        # 5 values (PRIORS_AGNfitter.py:363), prior_UV_xrays reads an undefined name (:267)
        if ms.settings.get('PRIOR_galaxy_only') == True:
            raise NotImplementedError("PRIOR_galaxy_only raises in the reference (PRIORS_AGNfitter.py:363)")
        if ms.settings.get('XRAYS') == 'Prior_UV':
            raise NotImplementedError("XRAYS='Prior_UV' raises in the reference (PRIORS_AGNfitter.py:267)")
        xp = extra_prior_plan(ms)
        self.ctx.check(self.lib.agnf_modelset_config(self.ctx.h, C.byref(cfg)))
        if [f.name for f in ms.families] != list(FAMILY_ORDER[:len(ms.families)]):
            raise ValueError("families must be ordered %s" % (FAMILY_ORDER,))
        idxs = ms.idxs
        for fi, f in enumerate(ms.families):
            sg = f.sorted_grids() if f.fitted else []
            n_dim = f.n_dim
            is_grid = as_i32([1 if t == 'grid' else 0 for t in f.par_types[:n_dim]])
            size = as_i32([len(g) for g in f.grid[:n_dim]])
            stride = as_i64(f.stride[:n_dim])
            gs = as_f64(np.concatenate([s[0] for s in sg])) if sg else as_f64([])
            gp = as_i32(np.concatenate([s[1] for s in sg])) if sg else as_i32([])
            dev = device_tables.get(f.name) if device_tables else None
            pitch = 0
            if dev is not None:
                if (dev.dim() != 3 or tuple(dev.shape[:2]) != (ms.n_z, f.T) or dev.shape[2] < ms.n_b
                        or not dev.is_contiguous() or str(dev.dtype) != "torch.float64"):
                    raise ValueError("device table for %s must be contiguous float64 [n_z, T, pitch >= n_b]" % f.name)
                self._keep.append(dev)
                table_ptr, on_dev, pitch = dev.data_ptr(), 1, int(dev.shape[2])
            else:
                if f.table is None:
                    raise ValueError("family %s has no host table and no device table was given" % f.name)
                table = as_f64(f.table)
                self._keep.append(table)
                table_ptr, on_dev = ptr(table), 0
            aux = as_f64(f.aux) if f.aux is not None else None
            if f.name == "GA" and cfg.eb_mode and not cfg.ga_free_ebv and aux is None:
                raise ValueError("energy-balance prior needs GALAXYatt for a grid-E(B-V) galaxy family")
            if f.name == "SB" and cfg.eb_mode and aux is None:
                raise ValueError("energy-balance prior needs STARBURST_LIR")
            d = _lib.agnf_family_desc(fam=fi, n_dim=n_dim, par_offset=idxs[fi], table_on_device=on_dev, T=f.T,
                                      dim_is_grid=ptr(is_grid), dim_size=ptr(size), dim_stride=ptr(stride),
                                      grid_sorted=ptr(gs), grid_perm=ptr(gp), table=table_ptr, aux=ptr(aux),
                                      table_pitch=pitch, reserved=0)
            self.ctx.check(self.lib.agnf_family_upload(self.ctx.h, C.byref(d)))
        if xp["midir_uv"] or xp["ir_xray"]:
            dered = as_i32(xp["bb_dered_row"]) if xp["bb_dered_row"] is not None else None
            d = _lib.agnf_extra_priors(midir_uv=int(xp["midir_uv"]), ir_xray=int(xp["ir_xray"]), j_6um=xp["j_6um"],
                                       j_2500=xp["j_2500"], bb_dered_row=ptr(dered))
            self.ctx.check(self.lib.agnf_extra_priors_set(self.ctx.h, C.byref(d)))
        zs = zslices if zslices is not None else make_zslices(ms)
        self.zslices = zs
        cgal, cbbb, xray, j1500, omega = (as_f64(zs.cgal), as_f64(zs.cbbb), as_u64(zs.xray), as_i32(zs.j1500),
                                          as_f64(zs.omega))
        self.ctx.check(self.lib.agnf_zslices_upload(self.ctx.h, ptr(cgal), ptr(cbbb), ptr(xray), ptr(j1500),
                                                    ptr(omega)))
        self.S = 0

    # -- catalogue shard ---------------------------------------------------------------------
    def set_sources(self, src: SourceSet):
        if src.y.shape[1] != self.n_b or src.pmin.shape[1] != self.n_par:
            raise ValueError("SourceSet does not match the ModelSet")
        zi, y, isig, mask = as_i32(src.zi), as_f64(src.y), as_f64(src.isig), as_u64(src.mask)
        pmin, pmax, consts = as_f64(src.pmin), as_f64(src.pmax), as_f64(src.consts)
        br, im, ij = as_i32(src.agn_branch), as_i32(src.irsyn_mode), as_i32(src.irsyn_j)
        self.ctx.check(self.lib.agnf_sources_upload(
            self.ctx.h, src.S, ptr(zi), ptr(y), ptr(isig), ptr(mask), ptr(pmin), ptr(pmax), ptr(consts),
            C_NCONST, ptr(br), ptr(im), ptr(ij)))
        self.S = src.S

    # -- evaluation: host buffers in, host buffers out (the reference-facing path) -------------
    def lnprob(self, params, source_id=None, detail=False, out=None):
        """params [n, n_par] (or [n_par]); source_id [n] int (default: all zeros).
        Returns lnp [n]; with detail=True returns (lnp, rows[n,n_fam], ymodel[n,n_b], lnprior, lnlike)."""
        params = np.atleast_2d(as_f64(params))
        n = params.shape[0]
        if params.shape[1] != self.n_par:
            raise ValueError("params has %d columns, model has %d parameters" % (params.shape[1], self.n_par))
        if source_id is None:
            sid = np.zeros(n, dtype=np.int32)
        else:
            sid = as_i32(source_id)
            if sid.shape != (n,):
                raise ValueError("source_id must have shape (n,)")
            if n and (sid.min() < 0 or sid.max() >= self.S):
                raise IndexError("source_id out of range [0, %d)" % self.S)
        lnp = out if out is not None else np.empty(n, dtype=np.float64)
        if detail:
            rows = np.empty((n, self.n_fam), dtype=np.int32)
            ym = np.empty((n, self.n_b), dtype=np.float64)
            lpr = np.empty(n, dtype=np.float64)
            lll = np.empty(n, dtype=np.float64)
            self.ctx.check(self.lib.agnf_lnprob_batch(self.ctx.h, ptr(params), ptr(sid), n, ptr(lnp), ptr(rows),
                                                      ptr(ym), ptr(lpr), ptr(lll)))
            return lnp, rows, ym, lpr, lll
        self.ctx.check(self.lib.agnf_lnprob_batch(self.ctx.h, ptr(params), ptr(sid), n, ptr(lnp), None, None,
                                                  None, None))
        return lnp

    # -- evaluation on device-resident buffers (torch tensors are only pointer carriers) -------
    def lnprob_device(self, params_t, sid_t, lnp_t, stream_ptr=None, rows_t=None, ymodel_t=None):
        n = params_t.shape[0]
        self.ctx.check(self.lib.agnf_lnprob_batch_device(
            self.ctx.h, params_t.data_ptr(), sid_t.data_ptr(), n, lnp_t.data_ptr(),
            rows_t.data_ptr() if rows_t is not None else None,
            ymodel_t.data_ptr() if ymodel_t is not None else None, None, None, stream_ptr))

    def set_option(self, name, value):
        self.ctx.check(self.lib.agnf_set_option(self.ctx.h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int64()
        self.ctx.check(self.lib.agnf_get_option(self.ctx.h, name.encode(), C.byref(v)))
        return int(v.value)

    def debug_exp10(self, x):
        x = as_f64(x)
        y = np.empty_like(x)
        self.ctx.check(self.lib.agnf_debug_exp10(self.ctx.h, ptr(x), x.size, ptr(y)))
        return y

    def synchronize(self):
        self.ctx.check(self.lib.agnf_synchronize(self.ctx.h))

    def last_kernel_ms(self, which=1):
        ms = C.c_float()
        self.ctx.check(self.lib.agnf_last_kernel_ms(self.ctx.h, which, C.byref(ms)))
        return float(ms.value)

    @property
    def launch_count(self):
        return int(self.lib.agnf_launch_count(self.ctx.h))


# ---------------------------------------------------------------------------------------------
_C = 2.997e8         # DICTIONARIES_AGNfitter.py:483
_ANG = 1.e10


class DictBuilder:
    """filtering_models for whole template families over a z grid (K2)."""

    def __init__(self, filterdict, device=0, ctx=None):
        """filterdict = [central_nu_array, lambdas_dict, factors_dict] exactly as MODELSDICT.build
        passes it (DICTIONARIES_AGNfitter.py:148); dict values are lists holding one array."""
        self.ctx = ctx if ctx is not None else Context(device)
        self.lib = self.ctx.lib
        bands, lambdas_dict, factors_dict = filterdict
        self.bands = as_f64(bands)
        lam, fac, off = [], [], [0]
        for b in self.bands:
            L = np.array(lambdas_dict[b], dtype=np.float64)
            F = np.array(factors_dict[b], dtype=np.float64)
            if L.ndim != 2 or L.shape[0] != 1:
                raise ValueError("band %r maps to %d filter curves; the reference breaks on this too "
                                 "(SURVEY §8a-Q11)" % (b, L.shape[0] if L.ndim == 2 else -1))
            lam.append(L[0])
            fac.append(F[0])
            off.append(off[-1] + L.shape[1])
        self.n_b = len(self.bands)
        # NB every array handed to the C ABI is bound to a local first: ptr() of a temporary dangles
        band_lam_c = as_f64(_C / (10 ** self.bands) * _ANG)    # nu2lambda_angstrom(iband), DICTIONARIES:464
        off_a = as_i64(off)
        lam_a = as_f64(np.concatenate(lam))
        fac_a = as_f64(np.concatenate(fac))
        self.ctx.check(self.lib.agnf_filters_upload(
            self.ctx.h, self.n_b, ptr(band_lam_c), ptr(off_a), ptr(lam_a), ptr(fac_a)))
        self.contract_ms = 0.0

    def set_option(self, name, value):
        self.ctx.check(self.lib.agnf_set_option(self.ctx.h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int64()
        self.ctx.check(self.lib.agnf_get_option(self.ctx.h, name.encode(), C.byref(v)))
        return int(v.value)

    def build_family(self, templates, z_array, on_device=False, reuse_staging=False):
        """templates: the reference's `*Fdict_4plot` ({key: (log10 nu, Fnu)}, insertion-ordered).
        Returns the band-flux table in key order: a NumPy array [n_z, T, n_b], or with on_device=True
        a torch CUDA tensor [n_z, T, pitch] (pitch = n_b rounded up to even, pad column zero) that
        never leaves HBM.
        reuse_staging: keep this family's grouped, pinned template blocks and reuse them when the SAME dict object
        (same length, the same value tuples) comes again -- a catalogue fitted in z-chunks, or a dictionary rebuilt
        for other redshifts, then skips the grouping and the concatenation of ~10^4 small arrays (about a third of the
        wall time of a build).  Arrays modified in place between calls are the caller's business."""
        keys_n = len(templates)
        T_total = keys_n
        z = as_f64(np.atleast_1d(z_array))
        n_z = len(z)
        pitch = self.n_b
        if on_device:
            import torch
            pitch = (self.n_b + 1) & ~1          # even pitch: rows stay 16-byte aligned for the cooperative K1
            out_t = torch.empty((n_z, T_total, pitch), dtype=torch.float64,
                                device="cuda:%d" % self.ctx.device)
            out_ptr = out_t.data_ptr()
        else:
            out = np.zeros((n_z, T_total, self.n_b), dtype=np.float64)
            out_ptr = ptr(out)
        staged = None
        cache = getattr(self, "_tpl_cache", None)
        if reuse_staging:
            if cache is None:
                cache = self._tpl_cache = {}
            hit = cache.get(id(templates))
            if hit is not None and hit[0] is templates and len(hit[1]) == keys_n and \
                    all(a is b for a, b in zip(templates.values(), hit[1])):
                staged = hit[2]
        if staged is None:
            staged = self._stage_templates(templates, private=reuse_staging)      # a generator
            if reuse_staging:
                staged = list(staged)
                cache[id(templates)] = (templates, list(templates.values()), staged)
        self.contract_ms = 0.0
        for lam_sorted, perm, F, members, n_lam in staged:      # not cached: each group is built before the next is staged
            identity = (len(members) == T_total)
            row_map = None if identity else as_i64(members)
            self.ctx.check(self.lib.agnf_dict_build(
                self.ctx.h, ptr(lam_sorted), ptr(perm), ptr(F), len(members), n_lam, ptr(z), n_z,
                ptr(row_map), T_total, out_ptr, 1 if on_device else 0, pitch))
            ms = C.c_float()
            self.ctx.check(self.lib.agnf_last_kernel_ms(self.ctx.h, 2, C.byref(ms)))
            self.contract_ms += float(ms.value)
        return out_t if on_device else out

    def _stage_templates(self, templates, private=False):
        """Group the templates by frequency grid and assemble each group's [members, n_lam] flux block in pinned host
        memory.  Yields (lam_sorted, perm, F, members, n_lam) group by group.  private: blocks of their own (kept by
        the staging cache) instead of the builder's shared grow-only buffer, which the next group overwrites."""
        T_total = len(templates)
        # group templates by identical frequency grid (ragged families, e.g. DH02_CE01: 193 / 1366 points).
        # Loaders normally hand every template of a family the SAME array object: group by identity first
        # and compare contents only between the distinct objects.
        vals = list(templates.values())          # positional access: hashing tuple-of-string keys 10^4 times is slow
        ids = np.fromiter(map(id, (v[0] for v in vals)), dtype=np.int64, count=T_total)
        uniq, first, inverse = np.unique(ids, return_index=True, return_inverse=True)
        groups = {}
        for u in range(len(uniq)):
            nu = as_f64(vals[int(first[u])][0]).reshape(-1)
            members = np.nonzero(inverse == u)[0].tolist() if len(uniq) > 1 else list(range(T_total))
            groups.setdefault(nu.tobytes(), (nu, []))[1].extend(members)
        for _, members in groups.values():
            members.sort()
        for nu, members in groups.values():
            lam = (_C / (10 ** nu) * _ANG)[::-1]                       # DICTIONARIES:450-451 without (1+z)
            order = np.argsort(lam, kind="mergesort")                   # scipy interp1d sorts x this way
            lam_sorted = as_f64(lam[order])
            n_lam = len(nu)
            perm = as_i32((n_lam - 1) - order)                          # column in the stored (unreversed) Fnu
            # the template block is assembled directly in a pinned staging buffer: the upload inside agnf_dict_build
            # is then a true asynchronous DMA instead of a staged pageable copy
            parts = [vals[t][1] for t in members]
            F = self._pinned(len(members) * n_lam) if private else self._staging(len(members) * n_lam)
            try:
                np.concatenate(parts, axis=None, out=F, casting="unsafe")
            except ValueError:          # total size != len(members) * n_lam
                raise ValueError("a template's flux array does not match its frequency grid")
            yield lam_sorted, perm, F, members, n_lam

    def _pinned(self, n):
        """A pinned host block of n doubles of its own (the returned array keeps the tensor alive)."""
        import torch
        return torch.empty(max(n, 1), dtype=torch.float64).pin_memory().numpy()[:n]

    def _staging(self, n):
        """Pinned host buffer of n doubles (grow-only; reused by every build_family call of this builder)."""
        import torch
        buf = getattr(self, "_pin", None)
        if buf is None or buf.numel() < n:
            cap = max(n, 1 << 16)
            self._pin = buf = torch.empty(cap, dtype=torch.float64).pin_memory()
        return buf.numpy()[:n]

    def filtering_models(self, model_nus, model_fluxes, z):
        """Drop-in for DICTIONARIES_AGNfitter.filtering_models (one template, one z):
        returns (bands, Fnu[n_b, 1]) like the reference (np.array of shape-(1,) rows)."""
        tab = self.build_family({"t": (model_nus, model_fluxes)}, [z])
        return self.bands, tab[0, 0].reshape(-1, 1)
