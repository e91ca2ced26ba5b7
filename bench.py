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

One JSON line on stdout (rank 0).  `value` = evals/s with inputs resident in HBM; `e2e` = the same
through the host-buffer entry point (pinned host params -> H2D -> K1 -> D2H lnp) every step.
"""
import argparse
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
    tpl = synth.synthetic_templates(ga_free=args.ga_free)
    nus = np.asarray(fd[0])
    S, W = args.sources, args.walkers
    z, dlum, rng = synth.synthetic_catalog(None, S, seed=20260101 + rank)
    if not args.no_sort_z:
        order = np.argsort(z, kind="stable")       # host-side processing order of the shard
        z, dlum = z[order], dlum[order]
    z_grid = np.unique(z)
    builder = DictBuilder(fd, device=device)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    ms, dev_tables = synth.build_modelset(tpl, fd, z_grid, builder, on_device=True)
    torch.cuda.synchronize(device)
    k2_wall = time.perf_counter() - t0
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
    info = dict(k2_wall_s=k2_wall, k2_entries=n_entries, n_z=len(z_grid), setup_s=time.perf_counter() - t_setup)
    return eng, ms, dev_tables, raw, pars, sid, info


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
    S, W, n = args.sources, args.walkers, len(pars)
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
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    per_launch = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count - l0
    # per-launch durations of the dominant kernel (events recorded by the library on the launching stream)
    for _ in range(min(args.steps, 5)):
        step_resident()
        torch.cuda.synchronize(dev)
        per_launch.append(eng.last_kernel_ms(1))
    k1_ms = float(np.mean(per_launch))

    # ---- the sampler loop the evaluations are for: S sources x W walkers, stretch moves on the GPU -----
    ens_info = None
    if args.ensemble_steps > 0:
        from agnfitter_b200.ensemble import EnsembleSampler
        # as the second half of the production leg of the reference's schedule runs it
        # (ensemble.run_reference_schedule): every 10th step (out['Nthinning']) feeds the percentile histograms and
        # the best-fit tracker
        ens = EnsembleSampler(eng, W, seed=20260101 + rank, n_bins=256)
        ens.set_positions(pars.reshape(S, W, n_par))
        ens.run(2, hist_every=10, track_best='thinned')       # warm-up
        barrier()
        l_e = eng.launch_count
        t0 = time.perf_counter()
        ens.run(args.ensemble_steps, hist_every=10, track_best='thinned')   # synchronises the ctx stream before returning
        ens_s = time.perf_counter() - t0
        ens_launches = eng.launch_count - l_e
        _, _, acc = ens.state()
        ens_info = dict(steps=args.ensemble_steps, s_per_step=ens_s / args.ensemble_steps,
                        evals_per_s=S * W * args.ensemble_steps / ens_s, launches=int(ens_launches),
                        acceptance=float(acc.mean() / (args.ensemble_steps + 2)))
        del ens

    # ---- output stage: integrated luminosities of 100 chain rows + best fit per source --------
    sed_info = None
    if args.sed_sources > 0:
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
        if rank == 0:
            sed_info["cpu_rows_per_s_one_core"] = cpu_sed_rate(ms, ps, raw, pars, W)
        del sed, d_sp, d_ss, d_so

    # ---- the drop-in's own call shape: emcee hands ONE source's half-ensemble (W/2 rows) per call ----------
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
    sampler.stop_flag = True
    sampler.join(timeout=2)
    check = float(np.nansum(np.where(np.isfinite(np_lnp), 1.0, 0.0)))

    ens_ms = (ens_info["s_per_step"] * 1e3) if ens_info else 0.0
    t_ms = torch.tensor([ms_total, e2e_s * 1e3, ens_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(n)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total_max, e2e_ms_max, ens_ms_max = float(t_ms[0]), float(t_ms[1]), float(t_ms[2])
    total_rows = float(cnt[0])
    value = total_rows * args.steps / (ms_total_max * 1e-3)
    e2e_value = total_rows * args.steps / (e2e_ms_max * 1e-3)

    if rank == 0:
        bpe = k1_bytes_per_eval(n_comp, n_b, n_par, W)
        peak, peak_src = measured_peak()
        kname = "k1_coop_kernel" if eng.get_option("k1_last_kernel") == 1 else "k1_lnprob_kernel"
        achieved = bpe * n / (k1_ms * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        # CPU baseline: the oracle port on a bounded sample of the same workload
        unbind_host_threads()
        ns = min(args.cpu_sources, S)
        sample_ids = np.linspace(0, S - 1, ns).astype(np.int64)
        small, srcs, pl = build_host_sample(ms, dev_tables, raw, sample_ids, pars, W)
        k2_cpu, k2_cpu_n = cpu_dictionary_throughput(bands=n_b)
        cpu_all = cpu_port_throughput(small, srcs, pl, cores)
        cpu_one = cpu_port_throughput(small, srcs[:max(1, ns // cores)], pl[:max(1, ns // cores)], 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3: synthetic %d sources x %d walkers per GPU, %s, %s, dictionary of %d "
                                   "z-slices in HBM"
                                   % (S, W, "18-band default set" if n_b == 18 else "%d-band example set" % n_b,
                                      "free-E(B-V) galaxy variant (%d parameters, %s templates)" % (
                                          n_par, "+".join(str(f.T) for f in ms.families)) if args.ga_free else
                                      "classic 10-parameter model (11000+161+60+2 templates)", info["n_z"]),
                       "positions": args.positions + " in prior box + 1% out-of-box rows",
                       "source_order": "random z" if args.no_sort_z else "sorted by z",
                       "k1_variant": "cooperative (k1_coop_kernel)" if eng.get_option("k1_last_kernel") == 1 else "per-thread (k1_lnprob_kernel)",
                       "rows_per_step_per_gpu": n, "n_bands": n_b, "n_par": n_par,
                       "l2_policy": "inputs larger than L2 (params %.0f MB + dictionary %.1f GB per GPU)" % (
                           n * n_par * 8 / 1e6, info["k2_entries"] * 8 / 1e9),
                       "parallelism": "sources sharded over %d GPU(s), no data-path collective" % world,
                       "host_placement": numa},
            # measured through the sampler loop when it ran (key "ensemble"), else the K1-bound figure
            "sources_fitted_per_s": (total_rows / (ens_ms_max * 1e-3) / EVALS_PER_FIT) if ens_info
                                    else value / EVALS_PER_FIT,
            "sources_fitted_per_s_k1_bound": value / EVALS_PER_FIT,
            "evals_per_fit": EVALS_PER_FIT,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * n_par * 8 + n * 4),
                    "d2h_bytes_per_step": int(n * 8), "ms_per_step": e2e_ms_max / args.steps,
                    "api": "agnf_lnprob_batch (host buffers; %d-row chunks pipelined over 3 streams)"
                           % eng.get_option("pipeline_rows")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(kname, n), "peak_source": peak_src, "kernel": kname,
                         "algorithmic_bytes_per_launch": bpe * n,
                         "bytes_per_eval": bpe, "kernel_ms": k1_ms},
            "cpu_baseline": {"value": cpu_all, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "oracle/agn_oracle.ln_probab on %d of the %d sources x %d walkers, vectorised "
                                       "over a source's walkers, fork Pool over sources" % (ns, S, W),
                             "one_core": cpu_one},
            "small_batch": dict(small_calls, what="agnf_lnprob_batch (host buffers, pinned staging block), one call per batch: the "
                                            "shape emcee (vectorize=True, one source) and ultranest produce"),
            "ensemble": None if not ens_info else {
                "what": "agnf_ens_run: lock-step stretch moves for every source of the shard, walkers resident in "
                        "HBM; one fused launch per half-step (propose -> ln_probab -> accept inside k1_coop_kernel), "
                        "percentile histograms and best-fit tracker fed every 10th step",
                "steps": ens_info["steps"], "ms_per_step": ens_ms_max,
                "evals_per_s": total_rows / (ens_ms_max * 1e-3),
                "sources_fitted_per_s": total_rows / (ens_ms_max * 1e-3) / EVALS_PER_FIT,
                "mean_acceptance": ens_info["acceptance"], "launches": ens_info["launches"]},
            "sed_realizations": None if not sed_info else dict(
                sed_info, what="agnf_sed_intlum (K4): full-resolution SEDs of 5 components on all_nus_rest and the "
                               "example settings' 13 integrated luminosities / AGN fractions per chain row, "
                               "rows resident in HBM; cpu = oracle port of FLUXES_ARRAYS.fluxes, one core"),
            "dictionary_build": {"entries": int(info["k2_entries"]), "wall_s": info["k2_wall_s"],
                                 "entries_per_s": info["k2_entries"] / info["k2_wall_s"],
                                 "what": "K2 over every family and every distinct z of the shard, tables left in HBM; "
                                         "wall time incl. host preparation and uploads",
                                 "cpu_entries_per_s_one_core": k2_cpu,
                                 "cpu_sample": "oracle filtering_models on %d galaxy templates x %d bands" % (k2_cpu_n, n_b)},
            "clocks": sampler.summary(),
            "finite_rows_last_step": check,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU implementation of the path.  The reference is Python and is not present
    on the GPU box, so this arm times the oracle port (oracle/agn_oracle.py, pinned to the
    reference by tests/golden) on all host cores, on bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("the reference arm builds its synthetic dictionary sample with K2; needs the GPU box")
    a2 = argparse.Namespace(**vars(args))
    a2.sources = min(args.sources, 4000)         # the sample's dictionary; evaluated subset below
    eng, ms, dev_tables, raw, pars, sid, info = make_workload(a2, 0, 0)
    cores = os.cpu_count() or 1
    ns = min(args.cpu_sources, a2.sources)
    sample_ids = np.linspace(0, a2.sources - 1, ns).astype(np.int64)
    small, srcs, pl = build_host_sample(ms, dev_tables, raw, sample_ids, pars, args.walkers)
    for _ in range(max(args.warmup, 1)):
        cpu_port_throughput(small, srcs[:cores], pl[:cores], cores)
    vals = [cpu_port_throughput(small, srcs, pl, cores) for _ in range(args.steps)]
    v = float(np.mean(vals))
    n_step = ns * args.walkers
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": n_step / v * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: synthetic %d sources x %d walkers per GPU, 18-band default set, classic "
                               "10-parameter model (11000+161+60+2 templates)" % (args.sources, args.walkers),
                   "positions": args.positions + " in prior box + 1% out-of-box rows",
                   "sample": "each step evaluates %d of the sources x %d walkers (bounded sample of the workload; "
                             "the sample's dictionary z-slices were built with K2, outside the timed region)"
                             % (ns, args.walkers),
                   "n_bands": ms.n_b, "n_par": ms.n_par,
                   "parallelism": "host cores of rank 0 only (%d); other ranks idle" % cores},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "oracle port (NumPy, vectorised over a source's %d walkers), fork Pool over %d "
                                   "sources on %d cores" % (args.walkers, ns, cores)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sources", type=int, default=100000, help="sources per GPU")
    ap.add_argument("--walkers", type=int, default=100)
    ap.add_argument("--positions", default="uniform", choices=["uniform", "ball"])
    ap.add_argument("--variant", type=int, default=1, help="K1 kernel: 1 cooperative (default), 0 per-thread")
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
    ap.add_argument("--sed-sources", type=int, default=2000,
                    help="sources whose 100 chain rows go through the SED / integrated-luminosity stage (0: skip)")
    ap.add_argument("--ensemble-steps", type=int, default=10,
                    help="also time this many lock-step stretch-move steps of the GPU-resident sampler (0: skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
