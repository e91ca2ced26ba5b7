#!/usr/bin/env python
"""
bench.py — BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload ("config.workload"): C3 of SURVEY.md §8d — a synthetic catalogue of 100 000 sources x 100
walkers per GPU, default 18-band set, classic 10-parameter model set (11 000 + 161 + 60 + 2
templates), z ~ U(0.01, 6) rounded to 3 decimals; the dictionary for every distinct z of the shard
is built on the GPU by K2 and stays in HBM; one step = one pass of the batched log-posterior (K1)
over all 10^7 (source, walker) rows.  Sources are independent, so ranks hold independent shards
(weak scaling, no data-path collective); the only collectives are the timing reductions.

One JSON line on stdout (rank 0).  `value` = evals/s with inputs resident in HBM over EXACTLY --steps
passes; `sustained` = the same loop kept up for >= 2 s with the clocks sampled over it; `e2e` = the same
metric through the host-buffer entry point (pinned host params -> H2D -> K1 -> D2H lnp) every step.
Further keys: `roofline` (K1), `roofline_k2` (dictionary build: HBM-write, FP64 and DMMA fractions against
peaks measured in this run), `ensemble` (the GPU-resident sampler loop, >= 2 s), `e2e_fit` (catalogue in host
memory -> fit_catalog -> summary table on the host, final gather inside the timed region), `configs` (K1 on
the other BASELINE shapes: 49 bands, free-E(B-V) galaxies, the 18-parameter radio-to-X-ray model), `strong`
(N > 1: the 100k-source catalogue divided over the ranks), `cpu_baseline` (N = 1).

`--impl reference` times the CPU restatement of the reference (oracle/agn_oracle.py, pinned to the reference by
tests/golden) on all host cores on a bounded sample of THE SAME sources and walker positions; that process
never loads libagnfit_b200.so (its dictionary slices come from oracle.filtering_models).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "walker-source log-posterior evals/s"
UNIT = "evals/s"
EVALS_PER_FIT = 100 * (2 * 12000 + 25000)     # example/SETTINGS_AGNfitter.py:265-268


K1_KERNEL_NAMES = {0: "k1_lnprob_kernel", 1: "k1_coop_kernel", 2: "k1_g4_kernel"}


def k1_bytes_per_eval(n_comp, n_b, n_par, W):
    """gathered rows + parameters + result, plus per-source vectors amortised over W."""
    return 8 * (n_comp * n_b + n_par + 1) + 8 * 5 * n_b / W


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons DURING the timed region, through NVML (sub-millisecond per
    sample; the timed region of a default run is tens of milliseconds), nvidia-smi as a fallback."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = False
        self.how = "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except ValueError:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            self.nv = None
            self.how = "nvidia-smi"

    def _nvml_sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for nm, bit in (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                        ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                        ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                        ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(nm)

    def _smi_sample(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        s = [x.strip() for x in out.split(",")]
        if len(s) >= 8:
            self.sm.append(float(s[1]))
            self.mx.append(float(s[2]))
            for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[4:8]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self._nvml_sample()
                else:
                    self._smi_sample()
            except Exception:
                pass
            time.sleep(0.002 if self.nv is not None else 0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}


def ncu_traffic(kernel, rows):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` over `rows` rows, from the
    committed ncu --set full capture (profiles/k1_traffic.json, written by tools/ncu_summary.py);
    None when no capture of this launch shape is on file."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if not os.path.exists(p):
        return None
    for rec in json.load(open(p)):
        if rec.get("kernel") == kernel and int(rec.get("rows", -1)) == int(rows):
            return float(rec["dram_bytes"])
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


_CPU_WORK = None     # (ModelSet, oracle sources, walker blocks): inherited by fork, never pickled


def oracle_eval_sources(idx):
    """Worker: the oracle port on the sources `idx` (each with its W walkers)."""
    from oracle import agn_oracle as O
    ms, srcs, pars_list = _CPU_WORK
    n = 0
    for i in idx:
        O.ln_probab(ms, srcs[i], pars_list[i])
        n += len(pars_list[i])
    return n


def cpu_port_throughput(ms_small, srcs, pars_list, cores, repeats=1):
    """evals/s of the oracle port (vectorised over the walkers of a source), Pool over sources."""
    import multiprocessing as mp
    global _CPU_WORK
    _CPU_WORK = (ms_small, srcs, pars_list)
    chunks = [list(range(k, len(srcs), cores)) for k in range(cores)]
    chunks = [c for c in chunks if c]
    best = 0.0
    if cores == 1:
        for _ in range(repeats):
            t0 = time.perf_counter()
            n = oracle_eval_sources(chunks[0])
            best = max(best, n / (time.perf_counter() - t0))
        return best
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        pool.map(oracle_eval_sources, [c[:1] for c in chunks])          # warm the workers
        for _ in range(repeats):
            t0 = time.perf_counter()
            n = sum(pool.map(oracle_eval_sources, chunks, chunksize=1))
            best = max(best, n / (time.perf_counter() - t0))
    return best


def cpu_dictionary_throughput(n_templates=400, bands=18):
    """Dictionary entries/s of the oracle port of filtering_models (one core) on a sample of the same synthetic
    templates / filters: the CPU figure beside K2's."""
    from agnfitter_b200 import synth
    from oracle import agn_oracle as O
    fd = synth.synthetic_filterdict(which="example49" if bands == 49 else None)
    tpl = synth.synthetic_templates()["GA"]["F"]
    b = fd[0]
    lam = [np.asarray(fd[1][x][0]) for x in b]
    fac = [np.asarray(fd[2][x][0]) for x in b]
    keys = list(tpl.keys())[:: max(1, len(tpl) // n_templates)][:n_templates]
    t0 = time.perf_counter()
    for i, k in enumerate(keys):
        O.filtering_models(tpl[k][0], tpl[k][1], b, lam, fac, 0.5 + 0.001 * i)
    dt = time.perf_counter() - t0
    return len(keys) * len(b) / dt, len(keys)


# example/SETTINGS_AGNfitter.py:316-321 (the reference's default output settings)
EXAMPLE_INTLUM_MODELS = ['sb', 'bbb', 'bbbdered', 'gal', 'AGNfrac_opt', 'tor', 'tor+bbb', 'AGNfrac_IR', 'gal', 'bbb',
                         'AGNfrac_opt', 'tor', 'sb']
EXAMPLE_INTLUM_RANGES_UM = [[8., 1000.], [0.1, 1.], [0.1, 1.], [0.1, 1.], [0.1, 1.], [1., 30.], [0.1, 30],
                            [8., 1000.], [0.4, 0.5], [0.4, 0.5], [0.4, 0.5], [6, 6], [1., 30.]]


def cpu_sed_rate(ms, ps, src_raw, pars, W, n_rows=24):
    """Rows/s of the oracle port of FLUXES_ARRAYS.fluxes + integrated_luminosities (one core) on rows of the
    same workload: the CPU figure beside K4's."""
    from agnfitter_b200.realizations import ranges_to_hz
    from oracle import sed_oracle as SO
    hz = ranges_to_hz(EXAMPLE_INTLUM_RANGES_UM)
    t0 = time.perf_counter()
    z, dl = float(src_raw["z"][0]), float(src_raw["dlum"][0])
    with np.errstate(all="ignore"):
        grid, fl = SO.realization_fluxes(ms, ps, pars[:n_rows], z, dl)
        nl = SO.fluxes_to_nuLnu(grid, fl, z, dl)
        SO.integrated_luminosities(EXAMPLE_INTLUM_MODELS, hz, grid, nl, ms.radio)
    return n_rows / (time.perf_counter() - t0)


def build_host_sample(ms, dev_tables, src_raw, sample_ids, pars, W):
    """Small host-side ModelSet (only the sampled sources' z-slices) + oracle inputs."""
    from agnfitter_b200.flatten import ModelSet, family_from_keys_float
    zi = src_raw["zi"][sample_ids]
    uz, inv = np.unique(zi, return_inverse=True)
    small = ModelSet(settings=dict(ms.settings), bands=ms.bands, z_grid=ms.z_grid[uz])
    import torch
    idx = torch.as_tensor(uz.astype(np.int64), device=next(iter(dev_tables.values())).device)
    for f in ms.families:
        tab = np.ascontiguousarray(dev_tables[f.name].index_select(0, idx)[..., :ms.n_b].cpu().numpy())
        small.families.append(family_from_keys_float(f.name, f.par_names, f.par_types, f.keys_float, f.fitted,
                                                     tab, f.aux))
    srcs, pl = [], []
    for k, s in enumerate(sample_ids):
        srcs.append(dict(z=float(src_raw["z"][s]), dlum=float(src_raw["dlum"][s]), nus=src_raw["nus"],
                         fluxes=src_raw["fluxes"][s], fluxerrs=src_raw["fluxerrs"][s], pmin=src_raw["pmin"][s],
                         pmax=src_raw["pmax"][s], zi=int(inv[k])))
        pl.append(np.ascontiguousarray(pars[s * W:(s + 1) * W]))
    return small, srcs, pl


def make_workload(args, rank, device, need_gpu=True):
    """Synthetic shard for this rank: dictionary (K2, on device), sources, walker positions."""
    import torch
    from agnfitter_b200 import synth, sources
    from agnfitter_b200.engine import Engine, DictBuilder
    t_setup = time.perf_counter()
    fd = synth.synthetic_filterdict(which="example49" if args.bands == 49 else None)
    tpl = synth.synthetic_templates(ga_free=args.ga_free, c4=args.c4)
    settings = synth.C4_SETTINGS if args.c4 else synth.CLASSIC_SETTINGS
    nus = np.asarray(fd[0])
    S, W = args.sources, args.walkers
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.scaling == "strong" and world > 1:
        # ONE catalogue of S sources (rank 0's) divided over the ranks along z (shard.shard_catalog)
        from agnfitter_b200 import shard as _shard
        z, dlum, rng = synth.synthetic_catalog(None, S, seed=20260101)
        order, lo, hi = _shard.shard_catalog(z, rank, world)
        z, dlum = z[order[lo:hi]], dlum[order[lo:hi]]
        rng = np.random.default_rng(20260101 + 7919 * (rank + 1))
        S = len(z)
    else:
        z, dlum, rng = synth.synthetic_catalog(None, S, seed=20260101 + rank)
        if not args.no_sort_z:
            order = np.argsort(z, kind="stable")       # host-side processing order of the shard
            z, dlum = z[order], dlum[order]
    z_grid = np.unique(z)
    builder = DictBuilder(fd, device=device)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    builder.k2_wall_s = builder.k2_contract_ms = 0.0
    ms, dev_tables = synth.build_modelset(tpl, fd, z_grid, builder, settings=settings, on_device=True)
    torch.cuda.synchronize(device)
    k2_cold = builder.k2_wall_s           # first build of the process: pinned staging, scratch pool, output allocation
    del ms, dev_tables
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    builder.k2_wall_s = builder.k2_contract_ms = 0.0
    ms, dev_tables = synth.build_modelset(tpl, fd, z_grid, builder, settings=settings, on_device=True)
    torch.cuda.synchronize(device)
    modelset_wall = time.perf_counter() - t0
    k2_wall = builder.k2_wall_s           # DictBuilder.build_family calls only (templates in NumPy -> tables in HBM)
    n_entries = len(z_grid) * sum(f.T for f in ms.families) * ms.n_b
    eng = Engine(ms, device=device, ctx=builder.ctx, device_tables=dev_tables)
    eng.set_option("k1_variant", args.variant)
    # fluxes from ymodel(truth) of every source, evaluated by K1 itself
    dummy = sources.make_sources(ms, z, dlum, nus, np.ones((S, ms.n_b)), np.ones((S, ms.n_b)))
    eng.set_sources(dummy)
    truth = dummy.pmin + rng.uniform(size=dummy.pmin.shape) * (dummy.pmax - dummy.pmin)
    _, _, ym, _, _ = eng.lnprob(truth, np.arange(S, dtype=np.int32), detail=True)
    fl, fe = synth.catalog_fluxes(rng, ym, nus, z)
    src = sources.make_sources(ms, z, dlum, nus, fl, fe)
    eng.set_sources(src)
    pars, sid = synth.walker_positions(rng, src.pmin, src.pmax, W, mode=args.positions, frac_out=0.01)
    raw = dict(z=z, dlum=dlum, nus=nus, fluxes=fl, fluxerrs=fe, pmin=src.pmin, pmax=src.pmax, zi=src.zi)
    info = dict(k2_wall_s=k2_wall, k2_cold_wall_s=k2_cold, k2_entries=n_entries, modelset_wall_s=modelset_wall,
                k2_contract_ms=builder.k2_contract_ms, n_z=len(z_grid), setup_s=time.perf_counter() - t_setup,
                builder=builder, templates=tpl, filterdict=fd, z_grid=z_grid, S=S)
    return eng, ms, dev_tables, raw, pars, sid, info


def workload_config(args, world, n_z, n_b, n_par, fam_T):
    """`config` of the JSON line: a function of the workload only, so both arms print the same object."""
    S, W = args.sources, args.walkers
    n = S * W
    dict_gb = n_z * sum(fam_T) * n_b * 8 / 1e9
    if args.c4:
        model = "radio-to-X-ray model (5 components, %d parameters, %s templates)" % (n_par, "+".join(map(str, fam_T)))
    elif args.ga_free:
        model = "free-E(B-V) galaxy variant (%d parameters, %s templates)" % (n_par, "+".join(map(str, fam_T)))
    else:
        model = "classic 10-parameter model (11000+161+60+2 templates)"
    return {"workload": "C3: synthetic %d sources x %d walkers per GPU, %s, %s, dictionary of %d z-slices in HBM"
                        % (S, W, "18-band default set" if n_b == 18 else "%d-band example set" % n_b, model, n_z),
            "positions": args.positions + " in prior box + 1% out-of-box rows",
            "source_order": "random z" if args.no_sort_z else "sorted by z",
            "rows_per_step_per_gpu": n, "n_bands": n_b, "n_par": n_par,
            "l2_policy": "inputs larger than L2 (params %.0f MB + dictionary %.1f GB per GPU)" % (
                n * n_par * 8 / 1e6, dict_gb),
            "parallelism": "sources sharded over %d GPU(s), no data-path collective" % world,
            "scaling_mode": args.scaling}


def shape_only_modelset(tpl, fd, z_grid, settings):
    """ModelSet with the families' key grids but no tables (bounds, parameter layout)."""
    from agnfitter_b200.flatten import ModelSet, family_from_keys_float
    ms = ModelSet(settings=dict(settings), bands=np.asarray(fd[0], dtype=np.float64),
                  z_grid=np.asarray(z_grid, dtype=np.float64))
    for name in ("GA", "SB", "TO", "BB", "RAD"):
        if name not in tpl:
            continue
        fam = tpl[name]
        keys = list(fam["F"].keys())
        kf = np.array(keys).T.astype(float)
        aux = None if fam["aux"] is None else np.array([fam["aux"][k] for k in keys])
        if name == "GA" and 'free' in fam["types"]:
            aux = None
        ms.families.append(family_from_keys_float(name, fam["names"], fam["types"], kf, True, None, aux))
    return ms


def _oracle_family_slices(job):
    """Worker: band fluxes [T, n_b] of one family at one z through oracle.filtering_models."""
    from oracle import agn_oracle as O
    name, z = job
    tpl, bands, lam, fac = _CPU_WORK
    F = tpl[name]["F"]
    out = np.empty((len(F), len(bands)))
    for t, (nu, f) in enumerate(F.values()):
        out[t] = O.filtering_models(nu, f, bands, lam, fac, z)
    return name, z, out


def reference_workload(args, rank=0):
    """The GPU arm's workload rebuilt WITHOUT the product library, for a bounded sample of its sources: the same
    catalogue (seed), the same truth parameters, fluxes and walker positions (the random streams are consumed in
    the GPU arm's order and shapes), dictionary slices of the sampled redshifts from oracle.filtering_models."""
    import multiprocessing as mp
    from agnfitter_b200 import synth, sources
    from agnfitter_b200.flatten import ModelSet, family_from_keys_float
    from oracle import agn_oracle as O
    global _CPU_WORK
    fd = synth.synthetic_filterdict(which="example49" if args.bands == 49 else None)
    tpl = synth.synthetic_templates(ga_free=args.ga_free, c4=args.c4)
    settings = synth.C4_SETTINGS if args.c4 else synth.CLASSIC_SETTINGS
    nus = np.asarray(fd[0])
    S, W = args.sources, args.walkers
    z, dlum, rng = synth.synthetic_catalog(None, S, seed=20260101 + rank)
    if not args.no_sort_z:
        order = np.argsort(z, kind="stable")
        z, dlum = z[order], dlum[order]
    z_grid = np.unique(z)
    shape = shape_only_modelset(tpl, fd, z_grid, settings)
    n_b, n_par = shape.n_b, shape.n_par
    # sample: whole z-slices, evenly spread over the grid, until args.cpu_sources sources
    per_slice = max(1, int(round(S / len(z_grid))))
    n_slices = max(1, min(len(z_grid), args.cpu_sources // per_slice))
    zs_sample = z_grid[np.unique(np.linspace(0, len(z_grid) - 1, n_slices).astype(int))]
    sample_ids = np.nonzero(np.isin(z, zs_sample))[0][:args.cpu_sources]
    zs_sample = np.unique(z[sample_ids])
    cores = os.cpu_count() or 1
    bands = fd[0]
    lam = [np.asarray(fd[1][b][0]) for b in bands]
    fac = [np.asarray(fd[2][b][0]) for b in bands]
    _CPU_WORK = (tpl, bands, lam, fac)
    jobs = [(name, float(zz)) for zz in zs_sample for name in tpl]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(min(cores, len(jobs))) as pool:
        done = pool.map(_oracle_family_slices, jobs, chunksize=1)
    dict_s = time.perf_counter() - t0
    small = ModelSet(settings=dict(settings), bands=shape.bands, z_grid=zs_sample)
    for f in shape.families:
        tab = np.stack([next(o for (nm, zz, o) in done if nm == f.name and zz == float(zq)) for zq in zs_sample])
        small.families.append(family_from_keys_float(f.name, f.par_names, f.par_types, f.keys_float, f.fitted, tab,
                                                     f.aux))
    # the GPU arm's random stream (make_workload): truth, fluxes, positions — full shapes, sample rows used
    bounds = {}
    pmin = np.empty((S, n_par))
    pmax = np.empty((S, n_par))
    for k, zk in enumerate(z_grid):
        bounds[float(zk)] = sources.pdict_bounds(shape, float(zk))
    for i in range(S):
        pmin[i], pmax[i] = bounds[float(z[i])]
    truth = pmin + rng.uniform(size=pmin.shape) * (pmax - pmin)
    ym = np.zeros((S, n_b))
    zpos = {float(v): i for i, v in enumerate(zs_sample)}
    for i in sample_ids:
        ym[i] = O.ymodel(small, zpos[float(z[i])], float(z[i]), float(dlum[i]), truth[i][None, :])[0][0]
    fl, fe = synth.catalog_fluxes(rng, ym, nus, z)
    pars, sid = synth.walker_positions(rng, pmin, pmax, W, mode=args.positions, frac_out=0.01)
    srcs, pl = [], []
    for i in sample_ids:
        srcs.append(dict(z=float(z[i]), dlum=float(dlum[i]), nus=nus, fluxes=fl[i], fluxerrs=fe[i], pmin=pmin[i],
                         pmax=pmax[i], zi=zpos[float(z[i])]))
        pl.append(np.ascontiguousarray(pars[i * W:(i + 1) * W]))
    info = dict(n_z=len(z_grid), n_b=n_b, n_par=n_par, fam_T=[f.T for f in shape.families],
                dict_entries=len(zs_sample) * sum(f.T for f in shape.families) * n_b, dict_s=dict_s,
                sample_ids=sample_ids)
    return small, srcs, pl, info


def product_library_mapped():
    try:
        return any("libagnfit_b200" in line for line in open("/proc/self/maps"))
    except OSError:
        return None


_ORIGINAL_AFFINITY = []


def unbind_host_threads():
    """The CPU-baseline legs use every host core again."""
    if _ORIGINAL_AFFINITY:
        os.sched_setaffinity(0, _ORIGINAL_AFFINITY[0])


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and, by first touch, its pinned staging buffers) to the NUMA node its GPU
    hangs off, so that the host<->device copies of the e2e leg do not cross the socket interconnect.
    Returns a short description for the JSON line; does nothing where sysfs does not say."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return "numa node %d has no allowed cpus" % node
        _ORIGINAL_AFFINITY.append(allowed)
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:          # containers without sysfs topology: keep the default placement
        return "unbound (%s)" % type(e).__name__


def timed_passes(eng, step, stream, n_steps, torch):
    """CUDA events on the launching stream around n_steps calls of `step`; returns milliseconds."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n_steps):
        step()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1)


def k1_leg(args, rank, local, overrides, steps, torch, seconds=0.5):
    """K1 on another BASELINE shape: builds its own shard (dictionary + sources), times `steps` passes and a
    `seconds`-long run; everything is released afterwards."""
    a2 = argparse.Namespace(**dict(vars(args), **overrides))
    eng, ms, dev_tables, raw, pars, sid, info = make_workload(a2, rank, local)
    dev = torch.device("cuda", local)
    d_pars = torch.from_numpy(pars).to(dev)
    d_sid = torch.from_numpy(sid).to(dev)
    d_lnp = torch.empty(len(pars), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(dev)
    sp = stream.cuda_stream
    step = lambda: eng.lnprob_device(d_pars, d_sid, d_lnp, stream_ptr=sp)
    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    ms_short = timed_passes(eng, step, stream, steps, torch) / steps
    n_long = max(steps, int(seconds * 1e3 / max(ms_short, 1e-3)))
    ms_long = timed_passes(eng, step, stream, n_long, torch) / n_long
    n, n_comp = len(pars), len(ms.families)
    bpe = k1_bytes_per_eval(n_comp, ms.n_b, ms.n_par, a2.walkers)
    peak, _ = measured_peak()
    kname = K1_KERNEL_NAMES[eng.get_option("k1_last_kernel")]
    out = dict(sources=a2.sources, walkers=a2.walkers, n_bands=ms.n_b, n_par=ms.n_par, components=n_comp,
               templates="+".join(str(f.T) for f in ms.families), n_z=info["n_z"], rows_per_step=n,
               ms_per_step=ms_short, evals_per_s=n / (ms_short * 1e-3), sustained_ms_per_step=ms_long,
               sustained_evals_per_s=n / (ms_long * 1e-3), sustained_steps=n_long, bytes_per_eval=bpe,
               roofline_frac=bpe * n / (ms_short * 1e-3) / 1e9 / peak, kernel=kname,
               dictionary_entries=int(info["k2_entries"]), dictionary_wall_s=info["k2_wall_s"],
               dictionary_entries_per_s=info["k2_entries"] / info["k2_wall_s"])
    del eng, dev_tables, d_pars, d_sid, d_lnp, info
    torch.cuda.empty_cache()
    return out


def fit_leg(args, rank, world, local, torch, dist):
    """Product-level end to end: catalogue arrays in host memory -> catalog.fit_catalog (K2 for the shard's
    redshifts, the reference's full burn-in / production schedule on the GPU-resident ensemble, percentiles, best
    fit) -> summary table on rank 0's host, the final gather inside the timed region."""
    from agnfitter_b200 import synth, catalog
    S_tot = args.fit_sources * world
    fd = synth.synthetic_filterdict()
    tpl = synth.synthetic_templates()
    long_names = dict(GA="GALAXY", SB="STARBURST", TO="TORUS", BB="BBB")
    templates = {long_names[k]: v for k, v in tpl.items()}
    rng = np.random.default_rng(77)
    z = np.round(rng.uniform(0.01, 6.0, size=S_tot), 3)
    dlum = synth.z2Dlum(z)
    nus = np.asarray(fd[0])
    y = 1e-27 * 10 ** (-0.5 * (nus[None, :] - 14.0)) * (1 + 0.1 * rng.normal(size=(S_tot, len(nus))))
    y = np.abs(y)
    mc = dict(Nwalkers=args.walkers, Nburnsets=2, Nburn=args.fit_nburn, Nmcmc=args.fit_nmcmc, Nthin=10)
    torch.cuda.synchronize(local)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    table = catalog.fit_catalog(templates, fd, synth.CLASSIC_SETTINGS, z, dlum, nus, y, 0.1 * y, mc, seed=5,
                                n_bins=256, rank=rank, world=world, device=local)
    torch.cuda.synchronize(local)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt[0])
    evals = S_tot * args.walkers * (2 * args.fit_nburn + args.fit_nmcmc)
    out = dict(what="catalog.fit_catalog: catalogue arrays on the host -> K2 dictionary for the shard's redshifts -> "
                    "2 burn-in sets + production run of the GPU-resident ensemble -> percentiles / best fit -> one "
                    "gather of the per-source summaries to rank 0 (inside the timed region)",
               sources_total=S_tot, sources_per_gpu=args.fit_sources, walkers=args.walkers,
               schedule=dict(Nburnsets=2, Nburn=args.fit_nburn, Nmcmc=args.fit_nmcmc, Nthin=10),
               seconds=dt, sources_per_s=S_tot / dt, evals_per_s=evals / dt,
               summary_bytes=None if table is None else int(table.nbytes))
    if table is not None:
        assert table.shape[0] == S_tot and np.isfinite(table[:, 2:12]).all()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: default placement"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng, ms, dev_tables, raw, pars, sid, info = make_workload(args, rank, local)
    S, W, n = info["S"], args.walkers, len(pars)
    n_b, n_par, n_comp = ms.n_b, ms.n_par, len(ms.families)
    dev = torch.device("cuda", local)

    # ---- device-resident operands for `value` ------------------------------------------------
    h_pars = torch.from_numpy(pars).pin_memory()
    h_sid = torch.from_numpy(sid).pin_memory()
    d_pars = h_pars.to(dev, non_blocking=True)
    d_sid = h_sid.to(dev, non_blocking=True)
    d_lnp = torch.empty(n, dtype=torch.float64, device=dev)
    h_lnp = torch.empty(n, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize(dev)
    stream = torch.cuda.Stream(dev)          # a real stream: handle 0 would mean "the ctx's own stream" to the ABI
    sp = stream.cuda_stream
    assert sp != 0

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        eng.lnprob_device(d_pars, d_sid, d_lnp, stream_ptr=sp)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # ... and keep warming until the clocks / power management have seen ~0.25 s of this load: at N = 8 the first tens
    # of milliseconds after an idle phase run under a transient power cap (the sustained leg below shows the plateau)
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.25:
        for _ in range(10):
            step_resident()
        torch.cuda.synchronize(dev)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    per_launch = []
    ms_total = timed_passes(eng, step_resident, stream, args.steps, torch)
    barrier()
    launches = eng.launch_count - l0
    # per-launch durations of the dominant kernel (events recorded by the library on the launching stream)
    for _ in range(min(args.steps, 5)):
        step_resident()
        torch.cuda.synchronize(dev)
        per_launch.append(eng.last_kernel_ms(1))
    k1_ms = float(np.mean(per_launch))

    # ---- the same loop kept up for >= args.sustain_s seconds (steady-state clocks) ----------------------------
    sustained = None
    if args.sustain_s > 0:
        n_long = max(args.steps, int(args.sustain_s * 1e3 / max(ms_total / args.steps, 1e-3)) + 1)
        cs = ClockSampler(local)
        cs.start()
        barrier()
        ms_long = timed_passes(eng, step_resident, stream, n_long, torch)
        barrier()
        cs.stop_flag = True
        cs.join(timeout=2)
        launches += n_long
        sustained = dict(steps=n_long, ms=ms_long, clocks=cs.summary())

    # ---- N > 1: the catalogue of `sources` sources divided over the ranks (strong scaling of the same pass) -------
    strong_ms = 0.0
    if world > 1 and args.scaling == "weak":
        n_strong = (S // world) * W
        d_ps, d_ss, d_ls = d_pars[:n_strong], d_sid[:n_strong], d_lnp[:n_strong]
        step_s = lambda: eng.lnprob_device(d_ps, d_ss, d_ls, stream_ptr=sp)
        for _ in range(3):
            step_s()
        barrier()
        strong_ms = timed_passes(eng, step_s, stream, 20 * args.steps, torch) / (20 * args.steps)
        barrier()
        launches += 20 * args.steps

    # ---- the sampler loop the evaluations are for: S sources x W walkers, stretch moves on the GPU -----
    ens_info = None
    if args.ensemble_steps > 0:
        from agnfitter_b200.ensemble import EnsembleSampler
        # as the second half of the production leg of the reference's schedule runs it
        # (ensemble.run_reference_schedule): every 10th step (out['Nthinning']) feeds the percentile histograms and
        # the best-fit tracker
        ens = EnsembleSampler(eng, W, seed=20260101 + rank, n_bins=256)
        ens.set_positions(pars.reshape(S, W, n_par))
        ens.run(10, hist_every=10, track_best='thinned')       # warm-up
        barrier()
        cs = ClockSampler(local)
        cs.start()
        l_e = eng.launch_count
        t0 = time.perf_counter()
        ens.run(args.ensemble_steps, hist_every=10, track_best='thinned')   # synchronises the ctx stream before returning
        ens_s = time.perf_counter() - t0
        cs.stop_flag = True
        cs.join(timeout=2)
        ens_launches = eng.launch_count - l_e
        _, _, acc = ens.state()
        ens_info = dict(steps=args.ensemble_steps, s_per_step=ens_s / args.ensemble_steps, seconds=ens_s,
                        evals_per_s=S * W * args.ensemble_steps / ens_s, launches=int(ens_launches),
                        acceptance=float(acc.mean() / (args.ensemble_steps + 10)), clocks=cs.summary())
        del ens

    # ---- K2 on its own: the galaxy family of this shard once more, kernel and wall time; FP64 peaks -------------
    k2_info = None
    if rank == 0 and not args.no_k2_leg:
        builder, tpl = info["builder"], info["templates"]
        z_grid = info["z_grid"]
        torch.cuda.synchronize(dev)
        walls, kerns = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            tab = builder.build_family(tpl["GA"]["F"], z_grid, on_device=True)
            torch.cuda.synchronize(dev)
            walls.append(time.perf_counter() - t0)
            kerns.append(builder.contract_ms)
            del tab
        peaks = {}
        for kind, nm in ((0, "dfma"), (1, "dmma")):
            v = ctypes.c_double()
            eng.ctx.check(eng.lib.agnf_debug_fp64_peak(eng.ctx.h, kind, ctypes.byref(v)))
            peaks[nm] = float(v.value)
        T_ga = len(tpl["GA"]["F"])
        entries = len(z_grid) * T_ga * n_b
        fd = info["filterdict"]
        sum_nk = int(sum(len(fd[1][b][0]) for b in fd[0]))
        k_ms = float(min(kerns))
        peak_hbm, _ = measured_peak()
        alg_flops = 2.0 * len(z_grid) * T_ga * sum_nk
        exe_flops = builder.get_option("k2_last_mflops") * 1e6
        on_tc = builder.get_option("k2_last_kernel") == 3
        k2_info = {"bound": "hbm", "family": "GA (%d templates x %d samples)" % (T_ga, len(next(iter(tpl["GA"]["F"].values()))[0])),
                   "kernel": {3: "k2_contract_mma", 2: "k2_contract_sw", 1: "k2_contract_tile"}.get(builder.get_option("k2_last_kernel"), "?"),
                   "entries": int(entries), "kernel_ms": k_ms, "wall_ms": float(min(walls)) * 1e3,
                   "wall_over_kernel": float(min(walls)) * 1e3 / k_ms,
                   "entries_per_s_kernel": entries / (k_ms * 1e-3), "entries_per_s_wall": entries / float(min(walls)),
                   "achieved": entries * 8 / (k_ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s",
                   "frac": entries * 8 / (k_ms * 1e-3) / 1e9 / peak_hbm,
                   "algorithmic_bytes_per_launch": entries * 8,
                   "fp64_peak_tflops": peaks,
                   "executed_flops_per_launch": exe_flops,
                   "executed_tflops": exe_flops / (k_ms * 1e-3) / 1e12,
                   "frac_of_fp64_peak": exe_flops / (k_ms * 1e-3) / 1e12 / max(peaks["dmma" if on_tc else "dfma"], 1e-9),
                   "binding": "fp64 (the contraction's multiply-adds over the 8-slice union windows against the "
                              "measured %s peak); the HBM write of the result (`frac`) overlaps it" % (
                                  "DMMA" if on_tc else "DFMA"),
                   "algorithmic_flops_per_launch": alg_flops,
                   "algorithmic_tflops": alg_flops / (k_ms * 1e-3) / 1e12,
                   "note": "algorithmic flops = 2 n_k per entry (SURVEY 8d: one multiply-add per filter sample); the "
                           "kernel contracts the filter curves to <= ~40 template samples per band first, so it "
                           "executes far fewer — the output write is what bounds it"}

    # ---- output stage: integrated luminosities of 100 chain rows + best fit per source --------
    sed_info = None
    if args.sed_sources > 0 and not args.c4:
        from agnfitter_b200 import synth as _synth
        from agnfitter_b200.plotset import plotset_from_templates
        from agnfitter_b200.realizations import SedEngine, ranges_to_hz
        ps = plotset_from_templates(_synth.synthetic_templates(ga_free=args.ga_free), ms)
        sed = SedEngine(eng, ps, raw["z"], raw["dlum"])
        sed.set_bands(EXAMPLE_INTLUM_MODELS, ranges_to_hz(EXAMPLE_INTLUM_RANGES_UM))
        n_src = min(args.sed_sources, S)
        rows_per = min(W, 101)
        idx = (np.arange(n_src)[:, None] * W + np.arange(rows_per)[None, :]).reshape(-1)
        d_sp = d_pars[torch.from_numpy(idx).to(dev)].contiguous()
        d_ss = d_sid[torch.from_numpy(idx).to(dev)].contiguous()
        d_so = torch.empty((len(idx), sed.n_models), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        for _ in range(2):
            sed.int_lums_device(d_sp, d_ss, d_so)
        barrier()
        l_s = eng.launch_count
        t_k = []
        for _ in range(3):
            sed.int_lums_device(d_sp, d_ss, d_so)
            t_k.append(eng.last_kernel_ms(1))
        sed_launches = eng.launch_count - l_s
        k4_ms = float(np.mean(t_k))
        sed_info = dict(rows=int(len(idx)), rows_per_source=int(rows_per), grid_points=int(len(sed.grid)),
                        columns=int(sed.n_models), kernel_ms=round(k4_ms, 3),
                        rows_per_s=len(idx) / (k4_ms * 1e-3), sources_per_s=n_src / (k4_ms * 1e-3),
                        launches=int(sed_launches))
        if rank == 0 and world == 1:
            sed_info["cpu_rows_per_s_one_core"] = cpu_sed_rate(ms, ps, raw, pars, W)
        del sed, d_sp, d_ss, d_so

    # ---- the drop-in's own call shape: emcee hands ONE source's half-ensemble (W/2 rows) per call ----------
    # ---- ultranest's call shape at catalogue scale: one sampler thread per source, each asking for a few hundred
    #      points per call (MCMC_AGNfitter.py:56-62, vectorized=True); the multiplexer merges the pending requests of
    #      all threads into one launch per tick (agnfitter_b200/multiplex.py) ---------------------------------------
    mux_info = None
    if args.mux_sources > 0 and rank == 0:
        import threading
        from agnfitter_b200.multiplex import PosteriorMultiplexer
        n_cl, n_calls, n_pts = min(args.mux_sources, S), 12, 400
        rng_m = np.random.default_rng(7)
        cubes = rng_m.uniform(size=(n_cl, n_pts, n_par))
        reqs = [raw["pmin"][s] + cubes[s] * (raw["pmax"][s] - raw["pmin"][s]) for s in range(n_cl)]
        sids = [np.full(n_pts, s, dtype=np.int32) for s in range(n_cl)]
        for s in range(4):
            eng.lnprob(reqs[s], sids[s])
        t0 = time.perf_counter()
        for k in range(2):                                   # the same requests, one call per source (no merging)
            for s in range(n_cl):
                eng.lnprob(reqs[s], sids[s])
        per_source_s = (time.perf_counter() - t0) / 2
        with PosteriorMultiplexer(evaluate=eng.lnprob, pmin=raw["pmin"], pmax=raw["pmax"], max_wait_s=5e-4) as mux:
            def client(s):
                with mux.client(s) as c:
                    for k in range(n_calls):
                        c.my_posterior(reqs[s])
            th = [threading.Thread(target=client, args=(s,)) for s in range(n_cl)]
            t0 = time.perf_counter()
            [t.start() for t in th]
            [t.join() for t in th]
            mux_s = time.perf_counter() - t0
            mux_info = dict(concurrent_sources=n_cl, calls_per_source=n_calls, points_per_call=n_pts,
                            evals_per_s=n_cl * n_calls * n_pts / mux_s, merged_launches=int(mux.launches),
                            requests=int(mux.requests),
                            per_source_calls_evals_per_s=n_cl * n_pts / per_source_s,
                            what="one Python thread per source calling my_posterior(400 points) like ultranest's "
                                 "vectorised sampler; PosteriorMultiplexer merges the pending requests into one "
                                 "agnf_lnprob_batch per tick (values bit-identical to per-source calls: "
                                 "tests/test_multiplex_gpu.py); beside it the same requests issued one call per source")
        launches += mux_info["merged_launches"] + 2 * n_cl + 4

    small_calls = {}
    for rows_per_call in (W // 2, 16 * W):
        sp_, ss_ = np.ascontiguousarray(pars[:rows_per_call]), np.ascontiguousarray(sid[:rows_per_call])
        out_ = np.empty(rows_per_call)
        for _ in range(20):
            eng.lnprob(sp_, ss_, out=out_)
        t0 = time.perf_counter()
        for _ in range(200):
            eng.lnprob(sp_, ss_, out=out_)
        small_calls["%d_rows" % rows_per_call] = {"us_per_call": (time.perf_counter() - t0) / 200 * 1e6}
        small_calls["%d_rows" % rows_per_call]["evals_per_s"] = rows_per_call / small_calls["%d_rows" % rows_per_call]["us_per_call"] * 1e6

    # ---- end to end through the host-buffer entry point ----------------------------------------
    np_pars, np_sid, np_lnp = h_pars.numpy(), h_sid.numpy(), h_lnp.numpy()
    for _ in range(2):
        eng.lnprob(np_pars, np_sid, out=np_lnp)
    barrier()
    l1 = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.lnprob(np_pars, np_sid, out=np_lnp)        # synchronous: returns with lnp in host memory
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    launches += eng.launch_count - l1
    barrier()
    # a plain pinned host -> device copy of the same params block on every rank at once: what this box's host
    # memory / PCIe fabric gives the e2e leg to work with
    t0 = time.perf_counter()
    for _ in range(3):
        d_pars.copy_(h_pars, non_blocking=True)
    torch.cuda.synchronize(dev)
    h2d_s = (time.perf_counter() - t0) / 3
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    check = float(np.nansum(np.where(np.isfinite(np_lnp), 1.0, 0.0)))

    kname = K1_KERNEL_NAMES[eng.get_option("k1_last_kernel")]
    pipeline_rows = eng.get_option("pipeline_rows")

    # ---- product-level end to end: fit_catalog ---------------------------------------------------------------
    del d_lnp, h_lnp
    fit_info = None
    if args.fit_sources > 0 and not (args.c4 or args.ga_free or args.bands != 18):
        fit_info = fit_leg(args, rank, world, local, torch, dist)

    # ---- the other BASELINE shapes (N = 1): K1 at catalogue scale on each --------------------------------------
    legs = {}
    if world == 1 and not args.no_config_legs and not (args.c4 or args.ga_free or args.bands != 18):
        del d_pars, d_sid, eng, dev_tables
        torch.cuda.empty_cache()
        S2 = min(args.sources, args.leg_sources)
        legs["C1_C2_shape_49_bands"] = k1_leg(args, rank, local, dict(bands=49, sources=S2), args.steps, torch)
        legs["free_ebv_galaxy_18_bands"] = k1_leg(args, rank, local, dict(ga_free=True, sources=S2), args.steps, torch)
        legs["free_ebv_galaxy_49_bands"] = k1_leg(args, rank, local, dict(ga_free=True, bands=49, sources=S2),
                                                  args.steps, torch)
        legs["C4_radio_to_xray_49_bands"] = k1_leg(args, rank, local, dict(c4=True, bands=49, sources=S2),
                                                   args.steps, torch)

    ens_ms = (ens_info["s_per_step"] * 1e3) if ens_info else 0.0
    sus_ms = (sustained["ms"] / sustained["steps"]) if sustained else 0.0
    t_ms = torch.tensor([ms_total, e2e_s * 1e3, ens_ms, sus_ms, strong_ms, h2d_s * 1e3], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(n)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total_max, e2e_ms_max, ens_ms_max, sus_ms_max, strong_ms_max, h2d_ms_max = [float(x) for x in t_ms]
    total_rows = float(cnt[0])
    value = total_rows * args.steps / (ms_total_max * 1e-3)
    e2e_value = total_rows * args.steps / (e2e_ms_max * 1e-3)

    if rank == 0:
        bpe = k1_bytes_per_eval(n_comp, n_b, n_par, W)
        peak, peak_src = measured_peak()
        achieved = bpe * n / (k1_ms * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        cpu_block = None
        k2_cpu = k2_cpu_n = None
        if world == 1:
            # CPU baseline: the oracle port on a bounded sample of the same workload (rank 0, N = 1 only)
            unbind_host_threads()
            small, srcs, pl, rinfo = reference_workload(args, rank)
            ns = len(srcs)
            k2_cpu, k2_cpu_n = cpu_dictionary_throughput(bands=n_b)
            cpu_all = cpu_port_throughput(small, srcs, pl, cores)
            cpu_one = cpu_port_throughput(small, srcs[:max(1, ns // cores)], pl[:max(1, ns // cores)], 1)
            cpu_block = {"value": cpu_all, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "oracle/agn_oracle.ln_probab on %d of the %d sources x %d walkers (same sources "
                                   "and positions), vectorised over a source's walkers, fork Pool over sources"
                                   % (ns, S, W),
                         "one_core": cpu_one}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world, info["n_z"], n_b, n_par, [f.T for f in ms.families]),
            "run": {"k1_variant": kname,
                    "host_placement": numa},
            "sustained": None if not sustained else {
                "what": "the same pass repeated for >= %.1f s, CUDA events on the launch stream, max over ranks"
                        % args.sustain_s,
                "steps": sustained["steps"], "ms_per_step": sus_ms_max, "seconds": sus_ms_max * sustained["steps"] * 1e-3,
                "value": total_rows / (sus_ms_max * 1e-3), "unit": UNIT,
                "roofline_frac": bpe * n / (sus_ms_max * 1e-3) / 1e9 / peak, "clocks": sustained["clocks"]},
            # measured through the sampler loop when it ran (key "ensemble"), else the K1-bound figure
            "sources_fitted_per_s": (total_rows / (ens_ms_max * 1e-3) / EVALS_PER_FIT) if ens_info
                                    else value / EVALS_PER_FIT,
            "sources_fitted_per_s_k1_bound": value / EVALS_PER_FIT,
            "evals_per_fit": EVALS_PER_FIT,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * n_par * 8 + n * 4),
                    "d2h_bytes_per_step": int(n * 8), "ms_per_step": e2e_ms_max / args.steps,
                    "api": "agnf_lnprob_batch (host buffers; %d-row chunks pipelined over 3 streams)"
                           % pipeline_rows,
                    "pinned_h2d_alone": {"what": "plain pinned H2D copy of the params block, all ranks at once",
                                         "ms": h2d_ms_max, "gb_per_s_per_gpu": n * n_par * 8 / (h2d_ms_max * 1e-3) / 1e9,
                                         "gb_per_s_total": world * n * n_par * 8 / (h2d_ms_max * 1e-3) / 1e9}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(kname, n), "peak_source": peak_src, "kernel": kname,
                         "algorithmic_bytes_per_launch": bpe * n,
                         "bytes_per_eval": bpe, "kernel_ms": k1_ms},
            "roofline_k2": k2_info,
            "cpu_baseline": cpu_block,
            "ultranest_multiplex": mux_info,
            "small_batch": dict(small_calls, what="agnf_lnprob_batch (host buffers, pinned staging block), one call per batch: the "
                                            "shape emcee (vectorize=True, one source) and ultranest produce"),
            "ensemble": None if not ens_info else {
                "what": "agnf_ens_run: lock-step stretch moves for every source of the shard, walkers resident in "
                        "HBM; one fused launch per half-step (propose -> ln_probab -> accept inside k1_coop_kernel), "
                        "percentile histograms and best-fit tracker fed every 10th step",
                "steps": ens_info["steps"], "ms_per_step": ens_ms_max, "seconds": ens_ms_max * ens_info["steps"] * 1e-3,
                "evals_per_s": total_rows / (ens_ms_max * 1e-3),
                "sources_fitted_per_s": total_rows / (ens_ms_max * 1e-3) / EVALS_PER_FIT,
                "mean_acceptance": ens_info["acceptance"], "launches": ens_info["launches"],
                "clocks": ens_info["clocks"]},
            "strong": None if not strong_ms_max else {
                "what": "the same pass over %d sources in total (%d per GPU): the catalogue of BASELINE configs[2] "
                        "divided over the ranks" % ((S // world) * world, S // world),
                "ms_per_step": strong_ms_max, "evals_per_s": (S // world) * world * W / (strong_ms_max * 1e-3)},
            "e2e_fit": fit_info,
            "configs": legs or None,
            "sed_realizations": None if not sed_info else dict(
                sed_info, what="agnf_sed_intlum (K4): full-resolution SEDs of 5 components on all_nus_rest and the "
                               "example settings' 13 integrated luminosities / AGN fractions per chain row, "
                               "rows resident in HBM; cpu = oracle port of FLUXES_ARRAYS.fluxes, one core"),
            "dictionary_build": {"entries": int(info["k2_entries"]), "wall_s": info["k2_wall_s"],
                                 "entries_per_s": info["k2_entries"] / info["k2_wall_s"],
                                 "contract_kernels_ms": info["k2_contract_ms"],
                                 "first_build_wall_s": info["k2_cold_wall_s"],
                                 "modelset_wall_s": info["modelset_wall_s"],
                                 "what": "DictBuilder.build_family over every family and every distinct z of the shard: "
                                         "template dictionaries in NumPy -> band-flux tables in HBM; wall time incl. "
                                         "host preparation, uploads, helper kernels, second build of the process (the grouped, pinned "
                                         "template blocks of the first are reused, as fit_catalog does from chunk to chunk) "
                                         "(first_build_wall_s: the first, which also allocates the pinned staging "
                                         "block, the scratch pool and the output; modelset_wall_s adds the parsing "
                                         "of the key strings into snapping grids)",
                                 "cpu_entries_per_s_one_core": k2_cpu,
                                 "cpu_sample": None if k2_cpu is None else
                                 "oracle filtering_models on %d galaxy templates x %d bands" % (k2_cpu_n, n_b)},
            "clocks": sampler.summary(),
            "finite_rows_last_step": check,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU implementation of the path.  The reference is Python and is not present on the GPU box,
    so this arm times its restatement (oracle/agn_oracle.py, pinned to the reference by tests/golden) on all host
    cores, each step a bounded sample of the GPU arm's own sources and walker positions.  Nothing here touches the
    product library: the sample's dictionary slices come from oracle.filtering_models."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    small, srcs, pl, rinfo = reference_workload(args, 0)
    cores = os.cpu_count() or 1
    ns = len(srcs)
    for _ in range(max(args.warmup, 1)):
        cpu_port_throughput(small, srcs[:cores], pl[:cores], cores)
    vals = [cpu_port_throughput(small, srcs, pl, cores) for _ in range(args.steps)]
    v = float(np.mean(vals))
    n_step = ns * args.walkers
    mapped = product_library_mapped()
    assert not mapped, "the reference arm must not load the product library"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": n_step / v * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world, rinfo["n_z"], rinfo["n_b"], rinfo["n_par"], rinfo["fam_T"]),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "each step: oracle port (NumPy, vectorised over a source's %d walkers) on %d of the "
                                   "workload's sources with the GPU arm's walker positions, fork Pool over sources on "
                                   "%d cores; their %d dictionary entries from oracle.filtering_models (%.1f s, "
                                   "outside the timed region)"
                                   % (args.walkers, ns, cores, rinfo["dict_entries"], rinfo["dict_s"])},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "product_library_mapped": mapped,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sources", type=int, default=100000, help="sources per GPU (weak) / in total (strong)")
    ap.add_argument("--walkers", type=int, default=100)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --sources per GPU (the driver's scaling run); strong: one catalogue of --sources "
                         "sources divided over the ranks along z")
    ap.add_argument("--positions", default="uniform", choices=["uniform", "ball"])
    ap.add_argument("--variant", type=int, default=1, help="K1 kernel: 1 automatic (default: gather4 when supported), 0 per-thread, 2 cooperative, 3 gather4")
    ap.add_argument("--no-sort-z", action="store_true",
                    help="keep the catalogue in random z order (default: the shard is processed in z order so "
                         "that sources sharing a dictionary z-slice are adjacent)")
    ap.add_argument("--cpu-sources", type=int, default=384, help="sources in the CPU-baseline sample")
    ap.add_argument("--bands", type=int, default=18, choices=[18, 49],
                    help="18: the default band set (the workload BASELINE.json names); 49: the example settings' "
                         "radio-to-X-ray set (C1/C2 shape at catalogue scale)")
    ap.add_argument("--ga-free", action="store_true",
                    help="galaxy family in the shape of the reference's default BC03_metal library: E(B-V)_gal free, "
                         "reddening and the energy-balance integral on the fly (not the BASELINE workload)")
    ap.add_argument("--c4", action="store_true",
                    help="the radio-to-X-ray model of BASELINE configs[3]: 5 components, 18 parameters")
    ap.add_argument("--sed-sources", type=int, default=2000,
                    help="sources whose 100 chain rows go through the SED / integrated-luminosity stage (0: skip)")
    ap.add_argument("--ensemble-steps", type=int, default=1000,
                    help="lock-step stretch-move steps of the GPU-resident sampler to time (>= 2 s at C3; 0: skip)")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained K1 run in seconds (0: skip)")
    ap.add_argument("--fit-sources", type=int, default=10000,
                    help="sources per GPU of the fit_catalog end-to-end leg (0: skip)")
    ap.add_argument("--fit-nburn", type=int, default=12000)
    ap.add_argument("--fit-nmcmc", type=int, default=25000)
    ap.add_argument("--leg-sources", type=int, default=20000, help="sources of each per-config K1 leg")
    ap.add_argument("--mux-sources", type=int, default=256,
                    help="concurrent sampler threads of the ultranest-multiplexer leg (0: skip)")
    ap.add_argument("--no-config-legs", action="store_true")
    ap.add_argument("--no-k2-leg", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
