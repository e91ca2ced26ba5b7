// K1 — batched (source x walker) log-posterior for sm_100a.
//
// Two kernels produce the same results (rows bit-identical, floating point to ~1e-15):
//   k1_coop.cu    the production kernel.  A warp owns 32 consecutive rows of `params`.  The
//                 per-evaluation scalar work (box test, snapping, normalisations, priors) runs
//                 thread-per-row; the band work (gather, reddening, model sum, chi^2) runs
//                 lane-per-16-byte-chunk over the warp's 32 x n_b tile so that every template
//                 row is fetched by consecutive lanes with 128-bit loads (2 L1 wavefronts per
//                 row instead of 9); the two layouts meet in per-warp shared memory.
//   k1_g4.cu      thread per row, template rows through the TMA unit's gather4 mode (the default when its
//                 preconditions hold; see the file header).
//   k1_lnprob.cu  one thread per row end to end; no alignment or size requirements.  Used when
//                 the cooperative kernel's preconditions do not hold, and as a cross-check.
// All arithmetic is fp64.
//
// What each step mirrors in the reference (paths relative to AGNfitter/functions):
//   box test            PARAMETERSPACE_AGNfitter.py:119-121
//   snap_dim            DICTIONARIES_AGNfitter.py:184-214 (first argmin of |key - p| over the key list)
//   row gather          DICTIONARIES_AGNfitter.py:216-219
//   r_GA / r_BB         DICTIONARIES_AGNfitter.py:222-270 + MODEL_AGNfitter.py:1178-1255
//   model sum           PARAMETERSPACE_AGNfitter.py:231-240
//   chi^2               PARAMETERSPACE_AGNfitter.py:144-150
//   priors              PRIORS_AGNfitter.py:78-106, 109-208, 212-254, 428-430
//   ln_probab           PARAMETERSPACE_AGNfitter.py:167-173
#pragma once
#include <cstdint>
#include <cuda.h>            // CUtensorMap (types only: the library does not link libcuda)
#include <cuda_runtime.h>
#include "../../include/agnfit_b200.h"

// The cooperative kernel evaluates 10^(c E) as 2^(c' E / D) with c' = c * D log2(10) pre-scaled on the host and a
// D-entry table of 2^(j/D).  D = 16: the table is 128 bytes = one shared-memory row, so divergent lookups
// never bank-conflict (degree-7 polynomial); D = 64: degree 5, conflicts ~2x the ideal wavefronts.
#ifndef AGNF_POW2_DENOM
#define AGNF_POW2_DENOM 16
#endif

namespace agnf {

struct DimDesc {
  int32_t is_grid;
  int32_t size;
  int32_t smem_off;   // offset (in values) of this dimension inside grid_sorted / grid_perm
  int32_t ext_off;    // offset of this dimension's -inf sentinel inside the sentinel-extended arrays
  int32_t lut_off;    // offset (in entries) of this dimension's cell table
  int32_t lut_m;      // cells; 0: no table for this dimension (binary search)
  int64_t stride;
  double lut_lo;      // left edge of cell 0
  double lut_inv;     // cells per unit
  double lut_big;     // the table path answers only for |p| <= lut_big (k1_common.cuh::snap_dim_lut)
};

// one snapped dimension of a family, as the cooperative kernel wants it (grid dims with > 1 value only)
struct SnapTask {
  int32_t col;        // column of the parameter in a params row
  int32_t n;          // grid values
  int32_t ext_off;    // see DimDesc
  int32_t lut_off;
  int32_t lut_mmax;   // cells - 1; -1: no table (binary search)
  int32_t stride;     // rows per grid step
  double lo, inv, big;
  // decision table of the gather4 kernel (agnfit_capi.cu::build_decision_table): cell c = int((p - lo) * dinv)
  // answers with the snapped index itself, or names the ONE switch point the cell straddles
  int32_t dlut_off;   // offset (entries) of this dimension's uint16 cell codes
  int32_t dmmax;      // cells - 1; -1: no decision table
  int32_t brec_off;   // offset (doubles) of this dimension's switch points
  int32_t smem_off;   // offset of this dimension in grid_sorted / grid_perm (binary-search fallback)
  double dinv;
  int32_t fam;        // family the task belongs to (flat task list)
  int32_t perm_off;   // offset (uint16 entries) of the first-appearance indices in sorted order; -1: the sorted
                      // order IS the first-appearance order (index = sorted position)
};

// Decision tables.  Between two neighbouring grid values sv[j] < sv[j+1] the reference's argmin switches from one to the
// other at exactly one double, thr[j] (the smallest p that snaps to the upper value, ties included).  A cell code is
//   0xffff                      : not answerable here (binary search)
//   bit 15 clear                : the snapped first-appearance index itself
//   bit 15 set, low 14 bits = j : the cell straddles thr[j] (and, with bit 14, thr[j+1] too): sorted position
//                                 j + (p >= thr[j]) [+ (p >= thr[j+1])], mapped through perm unless it is the identity

struct FamDesc {
  const double* table;  // [n_z, T, pitch]; columns n_b..pitch-1 are zero
  const double* aux;    // [T] or nullptr
  int64_t T;
  int32_t n_dim;
  int32_t par_offset;
  int32_t pitch;        // row pitch of `table` in doubles (>= n_b)
  int32_t n_snap;
  DimDesc dim[AGNF_MAX_DIMS];
  SnapTask snap[AGNF_MAX_DIMS];
};

// what one evaluation needs to know about its source (gather4 kernel): 48 bytes, three 16-byte loads
struct K1SrcRec {
  int32_t zi, agn_branch, irsyn_mode, irsyn_j;
  uint64_t valid, xray;
  int32_t j1500, pad;
  double inv4pidl2;
};

struct K1Params {
  agnf_config cfg;
  FamDesc fam[AGNF_MAX_FAMILIES];
  int32_t n_grid_vals;        // total grid values
  int32_t norm_offset;        // column of GA in a params row
  int32_t nbp;                // pitch of every library-owned per-band array: n_b rounded up to even
  int32_t coop_line;          // cooperative kernel: band phase in line form (k1_coop_tile.inc) -- 1: where it is the
                              // faster one (k1_coop.cu::coop_line_ok), 2: wherever it runs, 0: never (pair form)
  const double* grid_sorted;  // [n_grid_vals]  (per-thread kernel)
  const int32_t* grid_perm;   // [n_grid_vals]
  // cooperative kernel: one blob a CTA pulls into shared memory with a single bulk copy:
  //   sentinel-extended grid doubles | their first-appearance indices int32 | cell tables uint16,
  //   each padded to 16 bytes
  const unsigned char* stage;
  int32_t stage_bytes, stage_off_grid, stage_off_perm, stage_off_lut;
  // gather4 kernel: every snap task of every family in one list (groups of four run as one straight-line block)
  int32_t n_tasks, pad3;
  SnapTask tasks[AGNF_MAX_FAMILIES * AGNF_MAX_DIMS];
  // cell codes uint16 | switch points double | first-appearance indices uint16, each padded to 16 bytes
  const unsigned char* stage2;
  int32_t stage2_bytes, stage2_off_dlut, stage2_off_brec, stage2_off_perm;
  // Gaussian_prior constants, evaluated on the host: log(1/(sqrt(2 pi) sigma)) for sigma = 2, 1.5, 0.1
  double lognorm_s2, lognorm_s15, lognorm_s01;
  // z-slices (pad columns are zero)
  const double* cgal;         // [n_z, nbp]  -0.4 k_Calzetti
  const double* cbbb;         // [n_z, nbp]  -0.4 k_Prevot
  const double* cgal64;       // cgal * D log2(10), D = AGNF_POW2_DENOM (see above)
  const double* cbbb64;       // cbbb * D log2(10)
  const uint64_t* xray;       // [n_z]
  const int32_t* j1500;       // [n_z]
  const double* omega;        // [n_z, nbp]
  // free E(B-V)_gal + energy balance: sum_b omega[z, b] * (F_GA[z, t, b] * 3.826e33) of every galaxy row, i.e. the
  // unreddened integral of PRIORS_AGNfitter.py:89, which does not depend on the walker ([n_z, T_GA]; else nullptr)
  const double* eb0;
  // sources (pad columns are zero)
  const int32_t* zi;          // [S]
  const double* y;            // [S, nbp]
  const double* isig;         // [S, nbp]
  const uint64_t* valid;      // [S]
  const double* pmin;         // [S, n_par]
  const double* pmax;         // [S, n_par]
  const double2* pbox;        // [S, n_par] (min, max) pairs
  const double* consts;       // [S, n_const]
  int32_t n_const;
  int32_t pad;
  const int32_t* agn_branch;  // [S]
  const int32_t* irsyn_mode;  // [S]
  const int32_t* irsyn_j;     // [S]
  int64_t S;
  const K1SrcRec* srcrec;     // [S] (gather4 kernel; rebuilt after every source / z-slice upload)
  // prior_midIR_UV / prior_IR_XRays (agnf_extra_priors_set)
  int32_t xp_midir_uv, xp_ir_xray, xp_j6, xp_j2500;
  const int32_t* xp_bb_dered; // [T_BB]
  double xp_c6um;             // 10**13.69897 (host libm)
  double lognorm_s06, lognorm_s05;
};

// Ensemble mode of the cooperative kernel: row i = (source i / H, walker i % H of the active half);
// the kernel draws the stretch-move proposal from the walker state itself, evaluates it and applies the
// acceptance rule in place (see k3_ensemble.cuh).  `pos == nullptr`: ordinary batch.
struct K1Ensemble {
  double* pos = nullptr;      // [S, W, n_par]
  double* lnp = nullptr;      // [S, W]
  int32_t* n_acc = nullptr;   // [S, W]
  int32_t W = 0, half = 0;
  int64_t step = 0;
  uint64_t seed = 0;
  double a = 2.0;
  int64_t src_offset = 0;     // see K3State
  const uint16_t* split = nullptr;   // [S, W] this step's split (k3_ensemble.cuh), or nullptr = fixed halves
};

// Persistent per-source sampler (k1_coop_kernel<..., PERSIST>): one CTA per source runs n_steps whole steps in one
// launch, with everything agnf_ens_run does per step done in the kernel (see agnfit_capi.cu::agnf_ens_run).
struct K1Persist {
  int64_t step0 = 0, n_steps = 0;   // stretch-move steps step0 .. step0 + n_steps - 1
  int32_t hist_every = 0;           // every hist_every-th step of this launch is a thinned step
  int32_t track_best = 0;           // 0 off, 1 every step, 2 thinned steps
  uint32_t* hist = nullptr;         // [S, n_par, n_bins] or nullptr
  const double2* pbox = nullptr;    // [S, n_par]
  int32_t n_bins = 0;
  int32_t rows_per_warp = 32;       // walkers of a half-ensemble per warp (1..32): fewer rows, more warps, shorter steps
  double* best_pos = nullptr;       // [S, n_par]
  double* best_lnp = nullptr;       // [S]
  double* chain = nullptr;          // recording buffer [cap, S, W, n_par] or nullptr
  double* chain_lnp = nullptr;      // [cap, S, W]
  int64_t chain_n0 = 0;             // entries recorded before this launch
  int64_t chain_origin = 0;         // step index of the first recorded step
  int32_t chain_every = 1, pad2 = 0;
  int64_t S = 0;
};

struct K1Batch {
  const double* params;       // [n, n_par]  (ignored in ensemble mode)
  const int32_t* source_id;   // [n]         (ignored in ensemble mode)
  int64_t n;
  double* lnp;                // [n]
  int32_t* rows;              // [n, n_fam] or nullptr
  double* ymodel;             // [n, n_b]   or nullptr
  double* lnprior;            // [n]        or nullptr
  double* lnlike;             // [n]        or nullptr
  K1Ensemble ens;
};

cudaError_t launch_k1(const K1Params& P, const K1Batch& B, cudaStream_t stream);

// gather4 kernel (k1_g4.cu): one TMA tensor map per family table
struct K1TensorMaps {
  CUtensorMap m[AGNF_MAX_FAMILIES];
  int32_t valid[AGNF_MAX_FAMILIES];
};
int k1_g4_make_map(K1TensorMaps& M, int f, const double* table, int64_t rows, int pitch);   // 0 = ok
bool k1_g4_supported(const K1Params& P, const K1TensorMaps& M, const K1Batch& B, const char** why);
cudaError_t launch_k1_srcrec(const K1Params& P, K1SrcRec* out, cudaStream_t stream);
unsigned long long k1_g4_prof_read(int i, bool reset);   // -DG4_PROF builds only (else 0)
cudaError_t launch_k1_g4(const K1Params& P, const K1TensorMaps& M, const K1Batch& B, cudaStream_t stream);

// cooperative kernel (k1_coop.cu); `why` receives a static string when unsupported
bool k1_coop_supported(const K1Params& P, const K1Batch& B, const char** why);
bool k1_coop_profitable(const K1Params& P, const K1Batch& B);   // false: the per-thread kernel is expected to win
cudaError_t launch_k1_coop(const K1Params& P, const K1Batch& B, cudaStream_t stream);
// persistent per-source sampler; B.ens describes the ensemble, B.n is ignored
unsigned long long k1_coop_prof_read(int i);     // -DKC_PROF builds only (else 0)
bool k1_coop_persist_supported(const K1Params& P, const K1Batch& B, int rows_per_warp = 32);
cudaError_t launch_k1_coop_persist(const K1Params& P, const K1Batch& B, const K1Persist& A, cudaStream_t stream);
// eb0[z, t] for the galaxy family (see K1Params::eb0), chunk sums in the order the kernel used to add them
cudaError_t launch_k1_eb0(const K1Params& P, double* eb0, cudaStream_t stream);
cudaError_t launch_exp10_fast(const double* x, double* y, int64_t n, cudaStream_t stream);

}  // namespace agnf
