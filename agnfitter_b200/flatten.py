"""
Flattening layer: the reference's dict-of-tuples-of-strings model dictionary ->
dense arrays the sm_100a kernels index directly.

The reference keeps, per template family, `{key: (bands, Fnu[n_b])}` with keys that
are strings or tuples of strings (DICTIONARIES_AGNfitter.py:86-128) and snaps a walker
to a key by per-dimension `argmin |float(key_d) - p_d|` over the FULL key list
(DICTIONARIES_AGNfitter.py:184-214).  Here each family becomes

    keys_float [n_dim, T]   the reference's pars_modelkeys_float (kept for the oracle)
    grid[d]    [n_d]        distinct values of dimension d in first-appearance order
                            (== what the full-list argmin can ever return)
    sorted[d], perm[d]      ascending copy + original index, for the kernel's binary search
    stride[d]               row = sum_d idx_d * stride[d]   (product grids, any key order)
    table      [n_z, T, n_b] band fluxes, row order == key insertion order
    aux        [T]           GALAXYatt (GA, grid E(B-V)) or STARBURST_LIR (SB)

`free` dimensions contribute index 0 (the first key's value, DICTIONARIES:197,210).
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

FAMILY_ORDER = ("GA", "SB", "TO", "BB", "RAD")     # order of parameter blocks in P (PARAMETERSPACE:78-87)

# bb_mode values (which on-the-fly law get_fluxes applies to the BBB row, DICTIONARIES:232-270)
BB_ALL_GRID = 0
BB_EBV_FREE = 1            # ['EBVbbb'] free
BB_EBV_ALPHA_FREE = 2      # ['EBVbbb','alphaScat'] free,free
BB_EBV_ALPHA_GAMMA = 3     # ['EBVbbb','alphaScat','Gamma'] free,free,grid


@dataclass
class Family:
    name: str
    par_names: List[str]
    par_types: List[str]
    keys: list
    keys_float: np.ndarray           # [n_dim, T]
    fitted: bool                     # False: single fixed template, no entry in P (AGN_RAD '-99.9')
    grid: List[np.ndarray]           # per dim, first-appearance distinct values ('free' dims: 1 value)
    stride: np.ndarray               # [n_dim] int64
    table: Optional[np.ndarray]      # [n_z, T, n_b] float64; None when the table lives only in HBM
    aux: Optional[np.ndarray] = None  # [T]

    @property
    def n_dim(self):
        return len(self.par_names) if self.fitted else 0

    @property
    def T(self):
        return self.keys_float.shape[1]

    def sorted_grids(self):
        """(values ascending, original index) per dimension for the device binary search."""
        out = []
        for g in self.grid:
            order = np.argsort(g, kind="mergesort")
            out.append((np.ascontiguousarray(g[order]), order.astype(np.int32)))
        return out


def _first_appearance_unique(col):
    seen = {}
    for v in col:
        if v not in seen:
            seen[v] = len(seen)
    vals = np.array(list(seen.keys()), dtype=np.float64)
    inv = np.array([seen[v] for v in col], dtype=np.int64)
    return vals, inv


def flatten_family(name, fdicts, par_names, par_types, aux_dict=None, fixed_key=None):
    """fdicts: list (one per z-slice) of the reference's `{key: (bands, Fnu)}` dicts with
    identical key order.  Returns a Family.  `fixed_key` marks the no-parameter case."""
    keys = list(fdicts[0].keys())
    for fd in fdicts[1:]:
        if list(fd.keys()) != keys:
            raise ValueError("%s: z-slices disagree on key order" % name)
    T = len(keys)
    n_b = len(np.asarray(fdicts[0][keys[0]][1]).reshape(-1))
    table = np.empty((len(fdicts), T, n_b), dtype=np.float64)
    for zi, fd in enumerate(fdicts):
        for t, k in enumerate(keys):
            table[zi, t] = np.asarray(fd[k][1], dtype=np.float64).reshape(-1)

    arr = np.array(keys)
    keys_float = arr.T.astype(float)             # DICTIONARIES:175-176
    if keys_float.ndim == 1:
        keys_float = keys_float[None, :]
    fitted = fixed_key is None
    n_dim = keys_float.shape[0]
    if fitted and n_dim != len(par_names):
        raise ValueError("%s: %d key columns but %d parameter names" % (name, n_dim, len(par_names)))

    grid, invs = [], []
    for d in range(n_dim):
        vals, inv = _first_appearance_unique(keys_float[d])
        strs = arr[:, d] if arr.ndim == 2 else arr
        if len(set(strs.tolist())) != len(vals):
            raise ValueError("%s dim %d: key strings and their float values are not one-to-one" % (name, d))
        if fitted and par_types[d] == 'free':
            vals, inv_used = vals[:1], np.zeros_like(inv)   # always the first key's value
        else:
            inv_used = inv
        grid.append(vals)
        invs.append(inv)

    # row = sum idx_d * stride_d must reproduce the dict lookup for every combination
    stride = np.zeros(n_dim, dtype=np.int64)
    if fitted:
        lookup = {tuple(int(invs[d][t]) for d in range(n_dim)): t for t in range(T)}
        origin = tuple(0 for _ in range(n_dim))
        if origin not in lookup:
            raise ValueError("%s: first key is not the grid origin" % name)
        for d in range(n_dim):
            if par_types[d] == 'free' or len(grid[d]) == 1:
                continue
            step = list(origin)
            step[d] = 1
            if tuple(step) not in lookup:
                raise ValueError("%s: keys are not a product grid (dim %d)" % (name, d))
            stride[d] = lookup[tuple(step)] - lookup[origin]
        sizes = [len(g) for g in grid]
        n_combo = int(np.prod(sizes))
        idx = np.indices(sizes).reshape(n_dim, n_combo)
        rows = lookup[origin] + (idx * stride[:, None]).sum(axis=0)
        for c in range(n_combo):
            key = tuple(int(i) for i in idx[:, c])
            if lookup.get(key, -1) != rows[c]:
                raise ValueError("%s: keys are not a strided product grid; combination %s has no row "
                                 "(the reference would raise KeyError there)" % (name, key))
        if lookup[origin] != 0:
            raise ValueError("%s: origin row is not 0" % name)

    aux = None
    if aux_dict is not None:
        aux = np.array([float(aux_dict[k]) for k in keys], dtype=np.float64)
    return Family(name=name, par_names=list(par_names), par_types=list(par_types), keys=keys,
                  keys_float=keys_float, fitted=fitted, grid=grid, stride=stride, table=table, aux=aux)


@dataclass
class ModelSet:
    """Everything walker-independent and source-independent for one model configuration
    (one choice of GALAXY/STARBURST/TORUS/BBB[/AGN_RAD] libraries, one filter set, a z grid)."""
    settings: dict
    bands: np.ndarray                  # [n_b] dictionary band centres, log10 Hz (full precision)
    z_grid: np.ndarray                 # [n_z]
    families: List[Family] = field(default_factory=list)

    @property
    def n_b(self):
        return len(self.bands)

    @property
    def n_z(self):
        return len(self.z_grid)

    def fam(self, name):
        for f in self.families:
            if f.name == name:
                return f
        return None

    @property
    def bbb_has_norm(self):
        return self.settings['BBB'] in ('R06', 'THB21')        # PARAMETERSPACE:60,202,231

    @property
    def radio(self):
        return self.settings['RADIO'] == True

    @property
    def norm_names(self):
        n = ['GA', 'SB', 'TO']
        if self.bbb_has_norm:
            n.append('BB')
        if self.radio:
            n.append('RAD')
        return n

    @property
    def par_names(self):
        names = []
        for f in self.families:
            if f.fitted:
                names.extend(f.par_names)
        return names + self.norm_names

    @property
    def n_par(self):
        return len(self.par_names)

    @property
    def idxs(self):
        """Start offset of each fitted family's block, then the norm block (P['idxs'])."""
        out, o = [], 0
        for f in self.families:
            out.append(o)
            o += f.n_dim
        out.append(o)
        return out

    @property
    def ga_free_ebv(self):
        ga = self.fam("GA")
        return 'free' in ga.par_types and ga.par_names[-1] == 'EBVgal'

    @property
    def bb_mode(self):
        bb = self.fam("BB")
        if 'free' not in bb.par_types:
            return BB_ALL_GRID
        if bb.par_names[-1] == 'EBVbbb':
            return BB_EBV_FREE
        if bb.par_types[-2:] == ['free', 'free'] and bb.par_names[-2:] == ['EBVbbb', 'alphaScat']:
            return BB_EBV_ALPHA_FREE
        if bb.par_types[-3:] == ['free', 'free', 'grid'] and bb.par_names[-3:] == ['EBVbbb', 'alphaScat', 'Gamma']:
            return BB_EBV_ALPHA_GAMMA
        raise ValueError("unsupported BBB parameter layout %s / %s" % (bb.par_names, bb.par_types))


def modelset_from_modelsdicts(mds, settings):
    """Build a ModelSet from a list of reference MODELSDICT-like objects, one per z
    (attributes per DICTIONARIES_AGNfitter.py:86-136), all built for the same filter set."""
    md0 = mds[0]
    radio = settings['RADIO'] == True
    if radio:
        gpn, spn, tpn, bpn, rpn, _ = md0.all_parnames
        gpt, spt, tpt, bpt, rpt, _ = md0.all_partypes
    else:
        gpn, spn, tpn, bpn, _ = md0.all_parnames
        gpt, spt, tpt, bpt, _ = md0.all_partypes
    bands = np.asarray(next(iter(md0.GALAXYFdict.values()))[0], dtype=np.float64)
    ms = ModelSet(settings=dict(settings), bands=bands, z_grid=np.array([float(m.z) for m in mds]))
    ga_free = 'free' in gpt
    ms.families.append(flatten_family("GA", [m.GALAXYFdict for m in mds], gpn, gpt,
                                      aux_dict=None if ga_free else md0.GALAXYatt_dict))
    ms.families.append(flatten_family("SB", [m.STARBURSTFdict for m in mds], spn, spt,
                                      aux_dict=md0.STARBURST_LIRdict))
    ms.families.append(flatten_family("TO", [m.TORUSFdict for m in mds], tpn, tpt))
    ms.families.append(flatten_family("BB", [m.BBBFdict for m in mds], bpn, bpt))
    if radio:
        keys = list(md0.AGN_RADFdict.keys())
        fixed = keys == ['-99.9']                        # PARAMETERSPACE:41,206
        ms.families.append(flatten_family("RAD", [m.AGN_RADFdict for m in mds], rpn, rpt,
                                          fixed_key='-99.9' if fixed else None))
    return ms


# ----------------------------------------------------------------------------
# dense on-disk form (.npz) — replaces the pickle-of-dicts cache for this path
# ----------------------------------------------------------------------------
def modelset_to_arrays(ms: ModelSet, prefix=""):
    import json
    out = {prefix + "settings_json": np.array(json.dumps(ms.settings, default=str)),      # odd values as their str()
           prefix + "bands": ms.bands, prefix + "z_grid": ms.z_grid,
           prefix + "family_names": np.array([f.name for f in ms.families])}
    for f in ms.families:
        p = prefix + f.name + "_"
        out[p + "par_names"] = np.array(f.par_names)
        out[p + "par_types"] = np.array(f.par_types)
        out[p + "keys_float"] = f.keys_float
        out[p + "fitted"] = np.array(f.fitted)
        out[p + "table"] = f.table
        if f.aux is not None:
            out[p + "aux"] = f.aux
    return out


def family_from_keys_float(name, par_names, par_types, keys_float, fitted, table, aux=None):
    """Rebuild a Family from dense arrays (inverse of modelset_to_arrays for one family)."""
    keys_float = np.atleast_2d(np.asarray(keys_float, dtype=np.float64))
    n_dim, T = keys_float.shape
    grid, invs = [], []
    for d in range(n_dim):
        vals, inv = _first_appearance_unique(keys_float[d].tolist())
        if fitted and par_types[d] == 'free':
            vals = vals[:1]
        grid.append(vals)
        invs.append(inv)
    stride = np.zeros(n_dim, dtype=np.int64)
    if fitted:
        lookup = {tuple(int(invs[d][t]) for d in range(n_dim)): t for t in range(T)}
        origin = tuple(0 for _ in range(n_dim))
        for d in range(n_dim):
            if par_types[d] == 'free' or len(grid[d]) == 1:
                continue
            step = list(origin)
            step[d] = 1
            stride[d] = lookup[tuple(step)] - lookup[origin]
        sizes = [len(g) for g in grid]
        idx = np.indices(sizes).reshape(n_dim, -1)
        rows = (idx * stride[:, None]).sum(axis=0)
        for c in range(idx.shape[1]):
            if lookup.get(tuple(int(i) for i in idx[:, c]), -1) != rows[c]:
                raise ValueError("%s: keys are not a strided product grid" % name)
    keys = [tuple(keys_float[:, t]) for t in range(T)]
    return Family(name=name, par_names=list(par_names), par_types=list(par_types), keys=keys,
                  keys_float=keys_float, fitted=bool(fitted), grid=grid, stride=stride,
                  table=None if table is None else np.ascontiguousarray(table, dtype=np.float64),
                  aux=None if aux is None else np.asarray(aux, dtype=np.float64))


def modelset_from_arrays(arrs, prefix=""):
    import json
    ms = ModelSet(settings=json.loads(str(arrs[prefix + "settings_json"])),
                  bands=np.asarray(arrs[prefix + "bands"], dtype=np.float64),
                  z_grid=np.asarray(arrs[prefix + "z_grid"], dtype=np.float64))
    for name in [str(x) for x in arrs[prefix + "family_names"]]:
        p = prefix + name + "_"
        aux = arrs[p + "aux"] if (p + "aux") in arrs else None
        ms.families.append(family_from_keys_float(
            name, [str(x) for x in arrs[p + "par_names"]], [str(x) for x in arrs[p + "par_types"]],
            arrs[p + "keys_float"], bool(arrs[p + "fitted"]), arrs[p + "table"], aux))
    return ms


def save_modelset(path, ms: ModelSet):
    np.savez_compressed(path, **modelset_to_arrays(ms))


def load_modelset(path):
    with np.load(path, allow_pickle=False) as z:
        return modelset_from_arrays({k: z[k] for k in z.files})
