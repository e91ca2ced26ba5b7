/*
 * agnfit_b200 — C ABI of the B200-native AGNfitter hot path.
 *
 * The reference (GabrielaCR/AGNfitter) is pure Python and has no FFI; its plugin surface for
 * this path is a set of Python callables.  Each entry point below names the
 * reference callable it replaces (file:line relative to the AGNfitter tree); the ctypes binding
 * a maintainer would add on the reference side is shown in INTEGRATION.md and implemented in
 * agnfitter_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success, non-zero on error,
 *     and never throws; agnf_last_error() returns a message for the last failure on that ctx
 *     (or, for ctx == NULL, of the last failed agnf_ctx_create on this thread).
 *   - the caller owns every buffer it passes; the library owns only what *_upload / *_build
 *     allocate on the device and frees it in agnf_ctx_destroy.  No global state.
 *   - one ctx per (process, device); create it AFTER fork (RUN_AGNfitter_multi.py:258 forks).
 *   - "host" pointers are ordinary (ideally pinned) host memory; "_device" entry points take
 *     device pointers and an optional cudaStream_t (as void*, NULL = the ctx's own stream).
 *   - numeric results follow the reference: -inf outside the prior box or for a non-finite
 *     prior (PARAMETERSPACE_AGNfitter.py:119-121,170-173), -9999 soft rejections
 *     (PRIORS_AGNfitter.py:98,192), NaN propagation otherwise.
 */
#ifndef AGNFIT_B200_H
#define AGNFIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AGNF_API __attribute__((visibility("default")))
#else
#define AGNF_API
#endif

#define AGNF_VERSION 1
#define AGNF_MAX_FAMILIES 5   /* GA, SB, TO, BB, RAD — PARAMETERSPACE_AGNfitter.py:199 */
#define AGNF_MAX_DIMS 4       /* most dimensions any template family has (MODEL_AGNfitter.py:163,498-510) */
#define AGNF_MAX_BANDS 64     /* validity masks are 64-bit */
#define AGNF_MAX_PARS 32

typedef struct agnf_ctx agnf_ctx;

AGNF_API int agnf_version(void);
AGNF_API int agnf_ctx_create(int device, agnf_ctx** out);
AGNF_API void agnf_ctx_destroy(agnf_ctx* ctx);
AGNF_API const char* agnf_last_error(const agnf_ctx* ctx);
/* blocks until everything queued on the ctx's stream has finished */
AGNF_API int agnf_synchronize(agnf_ctx* ctx);

/* ---------------------------------------------------------------------------------------------
 * K1 — batched log-posterior.  Replaces, for whole (source x walker) batches,
 *   PARAMETERSPACE_AGNfitter.ln_probab   (PARAMETERSPACE_AGNfitter.py:153-173)
 *     = ymodel (180-243) + ln_prior (109-124) -> PRIORS_AGNfitter.PRIORS (PRIORS_AGNfitter.py:7-75)
 *       + ln_likelihood (127-150), with get_model.pick_nD / get_fluxes
 *       (DICTIONARIES_AGNfitter.py:184-272) and the on-the-fly reddening laws
 *       (MODEL_AGNfitter.py:1178-1255) inside.
 * ------------------------------------------------------------------------------------------- */

/* models.settings keys read on the path (PARAMETERSPACE:41-77,201-239; PRIORS:13-68) */
typedef struct {
  int32_t n_b;           /* bands                                                          */
  int32_t n_z;           /* z-slices of the dictionary                                      */
  int32_t n_par;         /* len(P['names'])                                                 */
  int32_t n_fam;         /* 4, or 5 with RADIO                                              */
  int32_t bbb_has_norm;  /* BBB in {R06, THB21}: BB is a fitted norm; else BB=0, /4 pi dL^2 */
  int32_t radio;         /* settings['RADIO']                                               */
  int32_t eb_mode;       /* PRIOR_energy_balance: 0 off, 1 'Flexible', 2 'Restrictive'      */
  int32_t prior_agnfrac; /* PRIOR_AGNfraction (+ prior_stellar_mass)                        */
  int32_t ga_free_ebv;   /* last GALAXY parameter is a free EBVgal (Calzetti on the fly)    */
  int32_t bb_mode;       /* 0 all grid; 1 EBVbbb free; 2 EBVbbb,alphaScat free; 3 +Gamma grid */
} agnf_config;

AGNF_API int agnf_modelset_config(agnf_ctx* ctx, const agnf_config* cfg);

/* One template family = one get_model object (DICTIONARIES_AGNfitter.py:172-182).
 * Row order of `table` is the key insertion order of the reference's *Fdict. */
typedef struct {
  int32_t fam;                 /* 0 GA, 1 SB, 2 TO, 3 BB, 4 RAD                              */
  int32_t n_dim;               /* fitted parameters of this family; 0 = one fixed template   */
  int32_t par_offset;          /* first column of the family's block in a params row (P['idxs']) */
  int32_t table_on_device;     /* `table` is a device pointer (e.g. written by agnf_dict_build) */
  int64_t T;                   /* rows (distinct keys)                                       */
  const int32_t* dim_is_grid;  /* [n_dim] 1 'grid' / 0 'free'                                */
  const int32_t* dim_size;     /* [n_dim] distinct values ('free': 1)                        */
  const int64_t* dim_stride;   /* [n_dim] row = sum idx_d * stride_d                         */
  const double* grid_sorted;   /* concatenated over dims: distinct values, ascending         */
  const int32_t* grid_perm;    /* same shape: first-appearance index of each sorted value    */
  const double* table;         /* [n_z, T, n_b] host, or [n_z, T, table_pitch] device          */
  const double* aux;           /* [T] GALAXYatt_dict (GA, grid EBV) / STARBURST_LIRdict (SB); or NULL */
  int32_t table_pitch;         /* device tables: row pitch in doubles (0 = n_b).  An even pitch with  */
  int32_t reserved;            /* zeroed pad columns (agnf_dict_build writes them) keeps rows 16-byte  */
                               /* aligned, which the cooperative K1 kernel needs.                      */
} agnf_family_desc;

AGNF_API int agnf_family_upload(agnf_ctx* ctx, const agnf_family_desc* fam);

/* The two torus-anchored priors of PRIORS_AGNfitter.PRIORS (dispatch PRIORS_AGNfitter.py:50-54, 68-71):
 *   prior_midIR_UV  (PRIORS_AGNfitter.py:327-357, settings['PRIOR_midIR_UV'] == True)
 *   prior_IR_XRays  (PRIORS_AGNfitter.py:302-324, settings['XRAYS'] == 'Prior_midIR')
 * Both read the snapped torus row at the first band with 13.59897 < log nu < 13.79897 (j_6um); the first also
 * reads the dereddened disc row (bb_dered_row[snapped BBB row]) at the first band with 14.8 < log nu < 15.15
 * (j_2500).  Per source they need consts[1] = 4. * pi * dlum**2. (lumfactor), consts[3] = 4*pi*(dlum)**2 (the
 * factor :346 applies when BB != 0) and consts[2] = log10(F_2-10keV,data * lumfactor) (:317-318); the host
 * evaluates them with the reference's expressions (agnfitter_b200/sources.py).  Call after agnf_family_upload of
 * the disc family; disabled by default and by a later agnf_modelset_config. */
typedef struct {
  int32_t midir_uv;             /* 1: add prior_midIR_UV                                      */
  int32_t ir_xray;              /* 1: add prior_IR_XRays                                      */
  int32_t j_6um;                /* band index, see above                                      */
  int32_t j_2500;               /* band index (midir_uv only)                                 */
  const int32_t* bb_dered_row;  /* [T_BB] row of the unreddened twin of every disc template (midir_uv only) */
} agnf_extra_priors;

AGNF_API int agnf_extra_priors_set(agnf_ctx* ctx, const agnf_extra_priors* xp);

/* per-z-slice constants, all [n_z, n_b] / [n_z] host arrays (agnfitter_b200/sources.py:make_zslices) */
AGNF_API int agnf_zslices_upload(agnf_ctx* ctx, const double* cgal, const double* cbbb, const uint64_t* xray_bits,
                        const int32_t* j1500, const double* omega);

/* per-source operands = the DATA object's fields (DATA_AGNfitter.py:357-369) after the
 * walker-independent parts of ln_likelihood / Pdict / PRIORS have been evaluated on the host */
AGNF_API int agnf_sources_upload(agnf_ctx* ctx, int64_t S, const int32_t* zi, const double* y, const double* inv_sigma,
                        const uint64_t* valid_bits, const double* pmin, const double* pmax,
                        const double* consts, int32_t n_const, const int32_t* agn_branch,
                        const int32_t* irsyn_mode, const int32_t* irsyn_j);

/* ln_probab for n rows.  params [n, n_par] row-major, source_id [n] with 0 <= source_id < S (the caller's
 * contract; agnfitter_b200.engine checks it).  Rows of one source should be adjacent and sources ordered by
 * z-slice: the kernels are correct for any order, fastest for that one.  Outputs (any may be NULL
 * except lnp): lnp [n]; rows [n, n_fam] snapped template row per family (pick_nD);
 * ymodel [n, n_b] (PARAMETERSPACE.ymodel); lnprior [n]; lnlike [n].
 * Host variant: copies in/out on the ctx stream and returns after the results are in the
 * caller's buffers. */
AGNF_API int agnf_lnprob_batch(agnf_ctx* ctx, const double* params, const int32_t* source_id, int64_t n, double* lnp,
                      int32_t* rows, double* ymodel, double* lnprior, double* lnlike);

/* Device variant: all pointers are device pointers; asynchronous on `stream`. */
AGNF_API int agnf_lnprob_batch_device(agnf_ctx* ctx, const double* d_params, const int32_t* d_source_id, int64_t n,
                             double* d_lnp, int32_t* d_rows, double* d_ymodel, double* d_lnprior,
                             double* d_lnlike, void* stream);

/* tuning knobs.  "k1_variant": 1 (default) = cooperative K1 kernel when its preconditions hold (16-byte
 * aligned params and tables, even table pitch) and two CTAs fit an SM, else the per-thread kernel;
 * 2 = cooperative whenever the preconditions hold; 0 = per-thread K1 kernel (no preconditions).
 * agnf_get_option also answers "k1_last_kernel" (which kernel the last K1 launch ran: 1 / 0, -1 none)
 * and "band_pitch" (n_b rounded up to even: the pitch to ask agnf_dict_build for).
 * "coop_line" (default 1): band phase of the cooperative kernel in line form (eight lanes per 128 bytes of one
 * evaluation's row, per-source operands cached in registers) where that is the faster mapping (18-band sets);
 * 2 = wherever it can run (>= 16 bands); 0 = round 1's lane-per-chunk pair form.  Same results bit for bit.
 * "pipeline_rows" (default 2^20): agnf_lnprob_batch with only lnp requested cuts batches larger than
 * this into chunks whose upload / kernel / download overlap on three streams; 0 = one shot.
 * "small_rows" (default 4096): batches of at most this many rows (what emcee / ultranest send per call) go
 * through one pinned staging block — one upload, one launch, one download; 0 = off.
 * "zero_copy_rows" (default 256): of those, batches of at most this many rows (a half-ensemble) are read by the
 * kernel straight from the pinned block through its device mapping and written back the same way — one launch,
 * one synchronisation, no timing events (agnf_last_kernel_ms keeps the previous timed launch); 0 = off.
 * "k2_mma" (default 1): agnf_dict_build contracts on the FP64 tensor cores (8 z-slices per batch) when the family
 * allows it.  "k2_shared_window" (default 1): else it evaluates 4 z-slices per read of the template window when
 * possible; 0 (with k2_mma 0) = the per-slice kernels.  agnf_get_option("k2_last_kernel"): 3 tensor-core,
 * 2 shared-window, 1 per-slice.
 * "ens_fused" (default 1): agnf_ens_run issues one fused launch per half-step (the cooperative K1 kernel
 * draws the proposal, evaluates it and applies the acceptance rule in place); 0 = separate propose / K1 /
 * accept kernels (same chain, bit for bit).
 * "ens_persist" (default -1): one launch per agnf_ens_run call, one CTA per source, for shards of <= 1184 sources
 * (1: whenever possible, 0: never); "ens_persist_rows" (default 0 = automatic): walkers of a half-ensemble per warp of
 * that kernel (1..32) -- fewer rows per warp shorten a step (one source: 25 -> 16 us), more CTAs per SM serve larger
 * shards.  Same chain bit for bit whatever the cut.
 * "ens_source_offset" (default 0): catalogue row of the first uploaded source; it is added to the source index
 * in the sampler's Philox counter, so a shard of a catalogue draws what a single-GPU run draws for the same
 * sources (results do not depend on how the catalogue is split over GPUs). */
AGNF_API int agnf_set_option(agnf_ctx* ctx, const char* name, int64_t value);
AGNF_API int agnf_get_option(agnf_ctx* ctx, const char* name, int64_t* value);
/* measurement hook for the K2 roofline (BASELINE.md §3 asks for a MEASURED FP64 peak): peak TFLOP/s of a DFMA loop
 * (kind 0, the FP64 pipe) or of an mma.sync.m8n8k4.f64 loop (kind 1, the FP64 tensor-core path k2_contract_mma
 * uses) on this device, best of 3 launches timed with CUDA events. */
AGNF_API int agnf_debug_fp64_peak(agnf_ctx* ctx, int32_t kind, double* tflops);

/* y[i] = 10^x[i] through the kernels' own exp10 routine (host pointers; unit-test hook) */
AGNF_API int agnf_debug_exp10(agnf_ctx* ctx, const double* x, int64_t n, double* y);

/* number of kernel launches this ctx has issued (bench.py's gpu_launches claim) */
AGNF_API int64_t agnf_launch_count(const agnf_ctx* ctx);
/* elapsed device time of the most recent K1 / K2 main kernel in ms (CUDA events on the launching stream) */
AGNF_API int agnf_last_kernel_ms(agnf_ctx* ctx, int which /*1 = K1, 2 = K2 contraction*/, float* ms);

/* ---------------------------------------------------------------------------------------------
 * K2 — dictionary build.  Replaces DICTIONARIES_AGNfitter.filtering_models
 * (DICTIONARIES_AGNfitter.py:426-503) applied to every template of a family and every z
 * (MODELSDICT.construct_dictionaryarray_filtered / build, DICTIONARIES_AGNfitter.py:78-154).
 * ------------------------------------------------------------------------------------------- */

/* filterdict = [central_nu_array, lambdas_dict, factors_dict] (FILTERS_AGNfitter.py:111-122),
 * concatenated: band b (bands sorted by central nu) owns samples [offset[b], offset[b+1]) in the
 * order stored in the filter file.  band_lambda_c[b] = nu2lambda_angstrom(central_nu_array[b])
 * (DICTIONARIES_AGNfitter.py:464,498-503), evaluated by the caller with the same NumPy expression
 * the reference uses so that no transcendental is re-derived on this side. */
AGNF_API int agnf_filters_upload(agnf_ctx* ctx, int32_t n_b, const double* band_lambda_c, const int64_t* offset,
                        const double* lambda_angstrom, const double* factor);

/* filtering_models for T templates that share one frequency grid, at n_z redshifts.
 *   lam_sorted [n_lam]  rest-frame wavelengths nu2lambda_angstrom(model_nus), ascending
 *                       (caller: NumPy expression of DICTIONARIES:450,502 + stable argsort, as
 *                       scipy interp1d does)
 *   perm       [n_lam]  column of Fnu that holds the flux of lam_sorted[j]
 *   Fnu        [T, n_lam] template F_nu in the template's stored order (host pointer)
 *   z          [n_z]
 *   out        [n_z, T_total, out_pitch]; template t writes row row_map[t] (NULL: identity, T_total = T);
 *              out_pitch 0 = n_b; columns n_b..out_pitch-1 are written as 0.
 * out is a host pointer (dense, results copied back) or, with out_on_device = 1, a device pointer that
 * can be handed to agnf_family_upload(table_on_device = 1, table_pitch = out_pitch) without leaving HBM. */
AGNF_API int agnf_dict_build(agnf_ctx* ctx, const double* lam_sorted, const int32_t* perm, const double* Fnu, int64_t T,
                    int32_t n_lam, const double* z, int32_t n_z, const int64_t* row_map, int64_t T_total,
                    double* out, int32_t out_on_device, int32_t out_pitch);

/* ---------------------------------------------------------------------------------------------
 * K3 — GPU-resident ensemble sampler for every uploaded source in lock-step.
 * Replaces, for whole catalogue shards, the emcee loop of MCMC_AGNfitter.main
 * (MCMC_AGNfitter.py:114-135,156-189): sampler.sample() -> stretch move -> ln_probab per walker.
 * Move = emcee's StretchMove (a = 2 by default) on the fixed first/second half-ensembles; random
 * numbers are Philox4x32-10 with counter (step, half, source, walker) and key `seed`, so a CPU
 * implementation reproduces the chain (oracle/ensemble_oracle.py).  Positions, log-probabilities and
 * the percentile histograms stay in HBM; only summaries come back.
 * ------------------------------------------------------------------------------------------- */

/* allocate the state for the S sources currently uploaded, `walkers` (even) walkers each.
 * n_bins > 0 also allocates per-(source, parameter) histograms over [P.min, P.max] for agnf_ens_percentiles. */
AGNF_API int agnf_ens_create(agnf_ctx* ctx, int32_t walkers, uint64_t seed, double a, int32_t n_bins);
/* pos [S, walkers, n_par] (host): initial positions (PARAMETERSPACE_AGNfitter.get_initial_positions /
 * get_best_position, :251-294); evaluates their ln_probab; clears the acceptance counters and the running best. */
AGNF_API int agnf_ens_set_positions(agnf_ctx* ctx, const double* pos);
/* The same with the positions drawn on the device: pos[s, w, :] = centre[s, :] + scale * d, d uniform in (-1, 1)
 * (normal = 0: get_initial_positions' ball around mid-box, :260-263) or standard normal (normal = 1:
 * get_best_position's ball around the best sample, :287-292).  centre [S, n_par] (host).  Draws come from the
 * sampler's Philox generator with `stream_id` in the counter (use a different one per call); like the move's
 * draws they depend on the source's catalogue row, not on the shard. */
AGNF_API int agnf_ens_set_ball(agnf_ctx* ctx, const double* centre, double scale, int32_t normal, uint32_t stream_id);
/* n_steps full stretch-move steps (2 half-steps each = 2 K1 launches).  hist_every > 0: every hist_every-th
 * step of this call adds all walker positions to the histograms (the reference thins its chain the same way,
 * PLOTandWRITE_AGNfitter.py:471-492).  track_best: 1 = keep the best (position, lnP) seen per source after every
 * step (what get_best_position reads back from the burn-in file); 2 = only at the thinned steps (the best-fit row
 * the reference picks from its thinned flat chain, :483-487); 0 = off. */
AGNF_API int agnf_ens_run(agnf_ctx* ctx, int64_t n_steps, int32_t hist_every, int32_t track_best);
/* Chain recording for the reference's chain files (samples_burn1-2-3.sav / samples_mcmc.sav,
 * MCMC_AGNfitter.py:192-206): from now on the next step of agnf_ens_run and every `every`-th one after it append the
 * walker state to a device buffer of max_steps entries (0: stop recording and free it; every = Nthin started at
 * step int(n/2) records exactly the thinned flat chain of CHAIN.props).  agnf_ens_run fails when the buffer is full.
 * agnf_ens_chain copies what was recorded: chain [n, S, W, n_par], lnp [n, S, W] (either may be NULL). */
AGNF_API int agnf_ens_record(agnf_ctx* ctx, int64_t max_steps, int32_t every);
AGNF_API int agnf_ens_chain(agnf_ctx* ctx, double* chain, double* lnp, int64_t* n_recorded);
/* current state to host buffers (any may be NULL): pos [S, W, n_par], lnp [S, W], n_acc [S, W] */
AGNF_API int agnf_ens_get(agnf_ctx* ctx, double* pos, double* lnp, int32_t* n_acc);
AGNF_API int agnf_ens_best(agnf_ctx* ctx, double* best_pos, double* best_lnp);   /* [S, n_par], [S] */
/* zero the acceptance counters and histograms (after burn-in: sampler.reset(), MCMC_AGNfitter.py:178) */
AGNF_API int agnf_ens_reset_stats(agnf_ctx* ctx);
/* forget the best sample seen so far (the next tracked step starts a new search) */
AGNF_API int agnf_ens_reset_best(agnf_ctx* ctx);
/* q [n_q] ascending percentiles in [0, 100] (the reference reports 2.5/16/50/84/97.5,
 * PLOTandWRITE_AGNfitter.py:152-173); out [S, n_par, n_q]; resolution = (P.max - P.min) / n_bins. */
AGNF_API int agnf_ens_percentiles(agnf_ctx* ctx, const double* q, int32_t n_q, double* out);
/* EXACT percentiles of the recorded chain: numpy.percentile (method 'linear') over the samples of the recorded
 * entries first, first + stride, ... of every walker — with first = int(n / 2), stride = Nthin on a recorded
 * production run this is what PLOTandWRITE_AGNfitter.py:162-191 computes from CHAIN.props' flat chain
 * (chain[:, int(n/2):n:Nthin, :], :476-487).  q [n_q <= 16] ascending in [0, 100]; out_pos [S, n_par, n_q];
 * out_lnp [S, n_q] (the log-posterior column, may be NULL).  Radix select on the device: no sort, no host copy. */
AGNF_API int agnf_ens_chain_percentiles(agnf_ctx* ctx, int64_t first, int64_t stride, const double* q, int32_t n_q,
                                        double* out_pos, double* out_lnp);
/* append the CURRENT walker state to the recording buffer (agnf_ens_record) without stepping */
AGNF_API int agnf_ens_record_state(agnf_ctx* ctx);

/* ---- K4: SED realisations + integrated luminosities ---------------------------------------
 * Replaces the per-realisation loop of FLUXES_ARRAYS.fluxes (functions/PLOTandWRITE_AGNfitter.py:566-799), the
 * `_4plot` get_fluxes it calls (functions/DICTIONARIES_AGNfitter.py:293-422), FLUXES2nuLnu_4plotting (:800-821) and
 * integrated_luminosities (:824-964) for a batch of (source, parameter row) pairs.  The model set, its families and
 * the sources must have been uploaded for K1 (rows are snapped with K1's grids).  Host-side preparation of the
 * arrays below: agnfitter_b200/realizations.py. */
typedef struct {
  int32_t fam;              /* 0 GA, 1 SB, 2 TO, 3 BB, 4 RAD                                                      */
  int32_t mode;             /* what `_4plot` get_fluxes does to the template (DICTIONARIES_AGNfitter.py:350-401):
                               0 nothing; 1 Calzetti with E = params[e_col]; 2 Prevot with E = params[e_col];
                               3 Prevot below 10^16.685 Hz, F * 10^(params[a_col] / 0.3838) at and above            */
  int32_t e_col, a_col;
  int32_t n_groups;         /* templates sharing a frequency grid form a group                                   */
  int32_t reserved;
  int64_t T;
  const int32_t* group;     /* [T] group of template t (row order = key order of the dictionary)                   */
  const int32_t* local;     /* [T] row of template t inside its group's block                                      */
  const int32_t* g_nlam;    /* [n_groups] samples per template                                                     */
  const double* x;          /* grids of all groups, each ascending (interp1d's mergesort order), concatenated      */
  const double* k;          /* reddening curve k(lambda) on the same layout (modes 1-3) or NULL                    */
  const double* val;        /* per group a [T_g, n_lam] block in sorted-x order, concatenated: F_nu (modes 1-3)
                               or log10 F_nu (mode 0)                                                              */
  const double* val2;       /* BB with dered_mode 1 or 2: log10 F_nu of the raw templates, same layout; else NULL  */
  const int32_t* jtab;      /* [n_groups, G] j with x[j] <= grid[i] < x[j+1] (numpy.interp's bracket), -1: the point
                               lies outside the (part of the) template that interp1d sees -> fill value -100       */
  const double* wtab;       /* [n_groups, G] (grid[i] - x[j]) / (x[j+1] - x[j]); 0 where grid[i] == x[j] or j < 0   */
} agnf_sed_family;

/* grid = all_nus_rest (PLOTandWRITE_AGNfitter.py:633), xg = 10**grid; the torus is kept on grid points
 * [to_lo, to_hi) (:733-734), the radio component on [0, rad_hi) (:693).  dered_mode / dered_row [T_BB]: which
 * template the dereddened disc curve comes from (:709-730; agnfitter_b200/plotset.py).  bands_jt / bands_wt [G]:
 * bracket table (as agnf_sed_family.jtab / wtab) of the dictionary band centres, needed for dered_mode 3 only. */
AGNF_API int agnf_sed_config(agnf_ctx* ctx, int32_t G, const double* grid, const double* xg, int32_t to_lo, int32_t to_hi,
                    int32_t rad_hi, int32_t dered_mode, const int32_t* dered_row, const int32_t* bands_jt,
                    const double* bands_wt);
AGNF_API int agnf_sed_family_upload(agnf_ctx* ctx, const agnf_sed_family* fam);
/* data.z and data.dlum of every source (same order as agnf_sources_upload) */
AGNF_API int agnf_sed_sources(agnf_ctx* ctx, int64_t S, const double* z, const double* dlum);
/* Integration bands (out['intlum_freqranges'], :839,859,873,939,943): band q covers grid points [i0, i1); i1 = -1
 * marks a monochromatic luminosity at point i0; mask = components needed (bit 0 sb, 1 bbb, 2 gal, 3 tor,
 * 4 agn_rad, 5 bbbdered).  Output column s (out['intlum_models'][s]) = sum of the `num` components of band
 * `range[s]` (kind 0) or 10**(sum num / sum den) (kind 1, the AGN fractions). */
AGNF_API int agnf_sed_bands(agnf_ctx* ctx, int32_t n_ranges, const int32_t* i0, const int32_t* i1, const int32_t* mask,
                   int32_t n_specs, const int32_t* kind, const int32_t* range, const int32_t* num,
                   const int32_t* den);
/* out [n, n_specs]: integrated_luminosities(...) transposed, linear (the reference then takes log10, :797).
 * on_device: params / source_id / out are device pointers. */
AGNF_API int agnf_sed_intlum(agnf_ctx* ctx, const double* params, const int32_t* source_id, int64_t n, double* out,
                    int32_t on_device);
/* out [n, n_comp, G]: nuLnus4plotting (:786) per row, n_comp = 6 (SB, BB, GA, TO, TOTAL, BBdered) or 7 with the
 * radio component (SB, BB, GA, TO, RAD, TOTAL, BBdered). */
AGNF_API int agnf_sed_spectra(agnf_ctx* ctx, const double* params, const int32_t* source_id, int64_t n, double* out,
                     int32_t on_device);

/* K4 over the ensemble's resident walker state (agnf_ens_*): every walker of every source is one row, the
 * W final positions of a source standing in for the rows FLUXES_ARRAYS.fluxes draws from the thinned chain
 * (:596-610).  out [S, W, n_specs] (host).  Nothing but the result leaves HBM. */
AGNF_API int agnf_sed_intlum_ens(agnf_ctx* ctx, double* out);

#ifdef __cplusplus
}
#endif
#endif /* AGNFIT_B200_H */
