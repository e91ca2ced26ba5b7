"""
Dense form of the reference's full-resolution template dictionaries (`*Fdict_4plot`,
DICTIONARIES_AGNfitter.py:92-131), the operand of the SED-realisation / integrated-luminosity stage
(PLOTandWRITE_AGNfitter.py:566-964).

A `PlotSet` lines up with a `flatten.ModelSet`: same families, same key (row) order, so the row indices the
log-posterior kernel snaps to address the full-resolution templates too.  Templates of a family that share a
frequency grid form a group (DH02_CE01 has two: 193 and 1 366 samples); a group stores its grid once and its
fluxes as one [T_g, n_lam] block.

What `FLUXES_ARRAYS.fluxes` does with the accretion-disc template for the "dereddened" curve depends on the
settings (PLT:709-730); `dered_mode` records the branch:
    0  the reddened curve is reused                                    (PLT:729-730)
    1  one fixed template, key '0.0'            R06 / THB21 without X-rays  (PLT:709-712)
    2  the matched grid key with EBVbbb := 0.0, no alpha_ox scaling    XRAYS == True  (PLT:719-727)
    3  SN12 without X-rays: the matched grid key with EBVbbb := 0.0, but read from the FILTERED dictionary and
       interpolated over the band centres (PLT:714-717)
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

FAMILY_ATTR = (("GA", "GALAXYFdict_4plot"), ("SB", "STARBURSTFdict_4plot"), ("TO", "TORUSFdict_4plot"),
               ("BB", "BBBFdict_4plot"), ("RAD", "AGN_RADFdict_4plot"))


@dataclass
class PlotFamily:
    name: str
    lognu: List[np.ndarray]            # per group: log10 nu as stored in the dictionary [n_lam]
    Fnu: List[np.ndarray]              # per group: [T_g, n_lam]
    group: np.ndarray                  # [T] group of each template
    local: np.ndarray                  # [T] row of each template inside its group's block

    @property
    def T(self):
        return len(self.group)

    def template(self, t):
        g = int(self.group[t])
        return self.lognu[g], self.Fnu[g][int(self.local[t])]


@dataclass
class PlotSet:
    families: List[PlotFamily] = field(default_factory=list)
    dered_mode: int = 0
    dered_row: np.ndarray = None       # [T_BB] row of the template the dereddened curve is read from

    def fam(self, name):
        for f in self.families:
            if f.name == name:
                return f
        raise KeyError(name)


def plot_family(name, templates):
    """templates: list of (log nu, Fnu) in key order."""
    by_grid, lognu, members = {}, [], []
    group = np.zeros(len(templates), dtype=np.int32)
    local = np.zeros(len(templates), dtype=np.int32)
    for t, (nu, F) in enumerate(templates):
        nu = np.ascontiguousarray(nu, dtype=np.float64).reshape(-1)
        F = np.asarray(F, dtype=np.float64).reshape(-1)
        if len(F) != len(nu):
            raise ValueError("%s template %d: %d fluxes on %d frequencies" % (name, t, len(F), len(nu)))
        g = by_grid.setdefault(nu.tobytes(), len(lognu))
        if g == len(lognu):
            lognu.append(nu)
            members.append([])
        group[t] = g
        local[t] = len(members[g])
        members[g].append(F)
    return PlotFamily(name=name, lognu=lognu, Fnu=[np.ascontiguousarray(np.stack(m)) for m in members],
                      group=group, local=local)


def dered_branch(settings):
    if settings['XRAYS'] == True:
        return 2
    if settings['BBB'] in ('R06', 'THB21'):
        return 1
    if settings['BBB'] == 'SN12':
        return 3
    return 0


def plotset_from_modelsdict(MD, ms):
    """MD: a reference MODELSDICT-like object holding the `*_4plot` dicts; ms: the ModelSet flattened from its
    filtered twins (flatten.modelset_from_modelsdicts) — the key order of the two must agree."""
    ps = PlotSet(dered_mode=dered_branch(ms.settings))
    for name, attr in FAMILY_ATTR:
        if name == "RAD" and not ms.radio:
            continue
        d4 = getattr(MD, attr)
        fam = ms.fam(name)
        keys = list(d4.keys())
        if len(keys) != fam.T:
            raise ValueError("%s: %d full-resolution templates for %d dictionary rows" % (name, len(keys), fam.T))
        ps.families.append(plot_family(name, [d4[k] for k in keys]))
    ps.dered_row = dered_rows(ms, ps.dered_mode)
    return ps


def plotset_from_templates(templates, ms):
    """templates: {'GA': {'F': {key: (log nu, Fnu)}, ...}, ...} (synth.synthetic_templates / the loaders' `*_4plot`
    dicts) in the key order the ModelSet was built from."""
    ps = PlotSet(dered_mode=dered_branch(ms.settings))
    for fam in ms.families:
        F = templates[fam.name]["F"]
        if len(F) != fam.T:
            raise ValueError("%s: %d full-resolution templates for %d dictionary rows" % (fam.name, len(F), fam.T))
        ps.families.append(plot_family(fam.name, list(F.values())))
    ps.dered_row = dered_rows(ms, ps.dered_mode)
    return ps


def dered_rows(ms, mode):
    bb = ms.fam("BB")
    T = bb.T
    rows = np.arange(T, dtype=np.int32)
    if mode == 1:
        hit = np.nonzero(bb.keys_float[0] == 0.0)[0]
        if len(hit) == 0:
            raise ValueError("no accretion-disc template with key '0.0'")
        rows[:] = hit[0]
    elif mode in (2, 3):
        d = bb.par_names.index('EBVbbb')
        lut = {tuple(bb.keys_float[:, t]): t for t in range(T)}
        for t in range(T):
            k = list(bb.keys_float[:, t])
            k[d] = 0.0
            rows[t] = lut[tuple(k)]
    return rows


def plotset_to_arrays(ps, prefix="ps_"):
    out = {prefix + "family_names": np.array([f.name for f in ps.families]),
           prefix + "dered_mode": np.array(ps.dered_mode), prefix + "dered_row": ps.dered_row}
    for f in ps.families:
        p = prefix + f.name + "_"
        out[p + "group"] = f.group
        out[p + "local"] = f.local
        out[p + "n_groups"] = np.array(len(f.lognu))
        for g in range(len(f.lognu)):
            out[p + "lognu%d" % g] = f.lognu[g]
            out[p + "Fnu%d" % g] = f.Fnu[g]
    return out


def plotset_from_arrays(arrs, prefix="ps_"):
    ps = PlotSet(dered_mode=int(arrs[prefix + "dered_mode"]),
                 dered_row=np.asarray(arrs[prefix + "dered_row"], dtype=np.int32))
    for name in [str(x) for x in arrs[prefix + "family_names"]]:
        p = prefix + name + "_"
        ng = int(arrs[p + "n_groups"])
        ps.families.append(PlotFamily(
            name=name, lognu=[np.asarray(arrs[p + "lognu%d" % g], dtype=np.float64) for g in range(ng)],
            Fnu=[np.asarray(arrs[p + "Fnu%d" % g], dtype=np.float64) for g in range(ng)],
            group=np.asarray(arrs[p + "group"], dtype=np.int32), local=np.asarray(arrs[p + "local"], dtype=np.int32)))
    return ps


def save_plotset(path, ps):
    """Dense `.npz` beside flatten.save_modelset: the full-resolution templates of a model set."""
    np.savez_compressed(path, **plotset_to_arrays(ps))


def load_plotset(path):
    with np.load(path, allow_pickle=False) as z:
        return plotset_from_arrays({k: z[k] for k in z.files})
