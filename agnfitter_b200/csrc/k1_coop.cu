// K1, cooperative kernel (see k1_lnprob.cuh for the mapping to the reference).
//
// A warp owns 32 consecutive rows of `params` (= 32 evaluations).
//
//   phase 1  thread-per-evaluation.  One 1-D bulk copy (TMA unit, mbarrier completion) per CTA brings
//            the snap decision tables into shared memory, one per warp brings that warp's
//            contiguous params rows.  Each thread box-tests its row, snaps every family (decision
//            table, k1_dev.cuh::snap_fast), evaluates the normalisations 10^x, leaves a 10- or
//            14-double record in shared memory and prefetches its galaxy-template row into L2.
//   phase 2  band phase, lanes over the 16-byte chunks (2 bands) of the warp's 32 x nbp tile.  A lane
//            fetches its evaluation's hand-off record, gathers one double2 per family plus y, 1/sigma
//            and the reddening coefficients, forms the model flux of its two bands and their masked
//            chi^2 terms, and stores the partial sum to shared memory.  Two lane mappings, same
//            arithmetic per (row, chunk), same results bit for bit (k1_coop_tile.inc):
//            line form (default for the 18-band sets): eight lanes take eight consecutive chunks of
//            ONE row, four rows per pass -- a quarter-warp's gather stays inside one row (1-2 lines) and
//            the lane's chunk is fixed, so the per-source operands stay in registers;
//            pair form: chunk c = e * nbp/2 + k goes to lane c % 32 of iteration c / 32.
//   phase 3  thread-per-evaluation again: sums its nbp/2 partials in band order (deterministic),
//            re-reads the two band values prior_AGNfraction needs (L1 hits) and -- line form -- its own
//            record (normalisations, E(B-V), row offsets: cheaper than carrying them in registers
//            through phase 2), evaluates the priors and writes lnP.
//
// Only __syncwarp() separates the phases; the CTA synchronises once, to publish the mbarriers.
#include "k1_lnprob.cuh"
#include "k1_common.cuh"
#include "k1_dev.cuh"
#include "k3_common.cuh"

namespace agnf {

namespace {

#ifndef KC_THREADS_DEF
#define KC_THREADS_DEF 256
#endif
#ifndef KC_MINBLOCKS
#define KC_MINBLOCKS 3
#endif
#ifndef KC_LINE_MINB
#define KC_LINE_MINB 4      // CTAs per SM of the plain line-form launch
#endif
// snaps evaluated together as one straight-line block in the throughput kernels (the line form at 64 registers
// runs ~1 % faster with 1 than with round 1's 2: C3 1.182 -> 1.169 ms, fused step 1.870 -> 1.845 ms)
#ifndef KC_SNAP_GROUP
#define KC_SNAP_GROUP 1
#endif
constexpr int KC_THREADS = KC_THREADS_DEF;
constexpr int KC_WARPS = KC_THREADS / 32;

// hand-off record of one evaluation, in doubles (q = chunk index = element offset / 2):
//   0: {int32 qGA, qSB}   1: {int32 qTO, qBB}      q = (z-slice * T + snapped row) * pitch / 2
//   2: nGA   3: nSB   4: nTO   5: nBB * cBB
//   6: EBVbbb   7: 10^(alphaScat/0.3838)
//   8: {int32 s * nbp/2, zi * nbp/2}   9: valid bits
//   10: EBVgal   11: nRAD   12: {int32 qRAD, -}   13: X-ray band bits     (EXT records only)
// EXT = GA_FREE || RADIO || BB_MODE >= 2
template <bool EXT>
struct Rec {
  static constexpr int STRIDE = EXT ? 14 : 10;
};

// shared-memory plan, computed on the host
struct CoopPlan {
  int32_t stage_pad;      // bytes reserved for the staged blob (16-byte multiple)
  int32_t cpr, cprp;      // chunks per row; partial-sum pitch (odd)
  int32_t per_warp;       // doubles per warp: 32 records + max(params block, partial planes) [+ proposals]
  int32_t q_off;          // doubles from the partial-sum area to the proposal rows (ensemble mode), else 0
  int32_t smem_bytes;
  int32_t walkers_off;    // persistent sampler: bytes from the start of shared memory to the source's walkers
};

// The line form of the band phase (k1_coop_tile.inc) wants at least one full group of eight chunks per row.  It is the
// default where it is the faster one -- the 18-band sets with their unrolled instantiation (C3: 1.36 -> 1.19 ms per
// 1e7 rows, free-E(B-V) galaxies 0.449 -> 0.424 ms per 2e6); on the 49-band sets (three groups + 1 chunk, runtime
// trip counts) the pair form stays 1-6 % ahead.  coop_line = 2 forces it wherever it can run (tests).
inline bool coop_line_ok(const K1Params& P) {
  return P.nbp >= 16 && (P.coop_line == 2 || (P.coop_line == 1 && P.nbp / 2 == 9));
}

template <bool EXT>
CoopPlan make_plan(const K1Params& P, bool ens, int n_warps = KC_WARPS) {
  CoopPlan c;
  const bool eb_free = P.cfg.ga_free_ebv != 0 && P.cfg.eb_mode != 0;
  c.stage_pad = (P.stage2_bytes + 15) & ~15;
  c.cpr = P.nbp / 2;
  c.cprp = c.cpr | 1;
  const int a = 32 * c.cprp * (eb_free ? 2 : 1), b = 32 * P.cfg.n_par;
  // plain batches land their params block in the area the partial sums take over later; ensemble mode keeps its
  // proposal rows in an area of their own until the acceptance test
  const int part = (((ens || a > b) ? a : b) + 1) & ~1;
  c.q_off = ens ? part : 0;
  c.per_warp = 32 * Rec<EXT>::STRIDE + part + (ens ? ((32 * P.cfg.n_par + 1) & ~1) : 0);
  c.smem_bytes = 64 * 8 /*exp2 table*/ + 16 * 8 /*mbarriers*/ + c.stage_pad + n_warps * c.per_warp * 8;
  c.walkers_off = c.smem_bytes;
  return c;
}

// shared-memory copy of the snap tasks (persistent sampler)
constexpr uint32_t KC_TASK_BYTES = (AGNF_MAX_FAMILIES * AGNF_MAX_DIMS * sizeof(SnapTask) + 15) & ~15u;

#ifdef KC_PROF
__device__ unsigned long long kc_prof[12];
#endif
// ---- per-step work of the persistent sampler, each the per-source form of a K3 kernel (k3_ensemble.cu) --------------
// k3_best_kernel: the first walker with the largest lnp, if it beats the stored best
__device__ __forceinline__ void persist_best(const double* pos, const double* lnp, int W, int n_par, double* best_pos,
                                             double* best_lnp, int lane) {
  double bv = -INFINITY;
  int bk = -1;
  for (int k = lane; k < W; k += 32) {
    const double v = lnp[k];
    if (v > bv) {
      bv = v;
      bk = k;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
    if (ok >= 0 && (bk < 0 || ov > bv || (ov == bv && ok < bk))) {
      bv = ov;
      bk = ok;
    }
  }
  if (bk >= 0 && bv > *best_lnp) {          // every lane holds the same (bv, bk)
    for (int c = lane; c < n_par; c += 32) best_pos[c] = pos[(int64_t)bk * n_par + c];
    __syncwarp();
    if (lane == 0) *best_lnp = bv;
  }
}

// k3_hist_kernel for one source
__device__ __forceinline__ void persist_hist(const double* pos, int W, int n_par, uint32_t* hist, const double2* pbox,
                                             int n_bins, int64_t src) {
  for (int e = threadIdx.x; e < W * n_par; e += blockDim.x) {
    const int c = e % n_par;
    const double2 mm = pbox[src * n_par + c];
    const double v = pos[e];
    const double f = (v - mm.x) / (mm.y - mm.x) * (double)n_bins;
    int b = (int)f;
    b = max(0, min(n_bins - 1, b));
    if (v == v) atomicAdd(hist + (src * n_par + c) * n_bins + b, 1u);
  }
}

// CPR_T: chunks per row as a compile-time constant (fully unrolled band phase), or 0 = runtime
// ENS: fused stretch-move launch (propose -> ln_probab -> accept), see K1Ensemble
// MINB: CTAs per SM the register allocation is sized for (3: 80 registers; 2: 128, for shapes whose per-warp tiles
// leave room for two CTAs anyway)
// LINE: the line form of the band phase (eight lanes per 128 bytes of a row, k1_coop_tile.inc)
template <bool GA_FREE, int BB_MODE, bool RADIO, int CPR_T, bool ENS, int MINB, bool LINE>
__global__ void __launch_bounds__(KC_THREADS, MINB)
k1_coop_kernel(const __grid_constant__ K1Params P, const K1Batch B, const CoopPlan C) {
  constexpr bool PERSIST = false;
  constexpr bool EXT = GA_FREE || RADIO || BB_MODE >= 2;
  constexpr int STRIDE = Rec<EXT>::STRIDE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const agnf_config& cfg = P.cfg;
  const int n_par = cfg.n_par, n_b = cfg.n_b, nbp = P.nbp;
  const bool eb_free = GA_FREE && cfg.eb_mode != 0;
  const int CPR = CPR_T ? CPR_T : C.cpr;
  const int CPRP = CPR | 1;
  // shared-window addresses
  const uint32_t tab_s = smem_u32(smem_raw);                 // 64 doubles
  const uint32_t bar_s = tab_s + 64 * 8;                     // [0] blob, [1 + w] warp w
  const uint32_t stage_s = bar_s + 16 * 8;
  const uint32_t dlut_s = stage_s + P.stage2_off_dlut, brec_s = stage_s + P.stage2_off_brec,
                 perm_s = stage_s + P.stage2_off_perm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t scal_s = stage_s + C.stage_pad + (unsigned)warp * (unsigned)C.per_warp * 8u;
  const uint32_t part_s = scal_s + 32 * STRIDE * 8;
  const uint32_t q_s = part_s + 8u * (unsigned)C.q_off;        // params / proposal rows of the warp

  const int64_t row0 = ((int64_t)blockIdx.x * KC_WARPS + warp) * 32;
  const int rows_here = (int)max((int64_t)0, min((int64_t)32, B.n - row0));
  const uint32_t par_bytes = ENS ? 0u : (uint32_t)((rows_here * n_par) >> 1) * 16u;

  if (threadIdx.x == 0) {
    for (int w = 0; w <= KC_WARPS; ++w) mbar_init(bar_s + 8u * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < AGNF_POW2_DENOM)
    sts_f64(tab_s + 8u * threadIdx.x, c_exp2_tab[threadIdx.x * (64 / AGNF_POW2_DENOM)]);
  __syncthreads();
  // one bulk copy brings the snap decision tables; one per warp brings that warp's params rows
  if (threadIdx.x == 0 && P.stage2_bytes > 0) {
    mbar_expect_tx(bar_s, (uint32_t)P.stage2_bytes);
    bulk_g2s(stage_s, P.stage2, (uint32_t)P.stage2_bytes, bar_s);
  }
  if (lane == 0 && par_bytes > 0) {
    mbar_expect_tx(bar_s + 8u + 8u * warp, par_bytes);
    bulk_g2s(q_s, B.params + row0 * n_par, par_bytes, bar_s + 8u + 8u * warp);   // 256-byte aligned source
  }
  double* const w_pos = B.ens.pos;      // walker state (ensemble mode): global memory
  double* const w_lnp = B.ens.lnp;
  const uint16_t* const split_tab = B.ens.split;
  const int64_t e_step = B.ens.step;
  const int e_half = B.ens.half;
#define TILE_EXIT return
#include "k1_coop_tile.inc"
#undef TILE_EXIT
}

// The persistent per-source sampler (ensemble mode): one CTA per source runs `A.n_steps` whole stretch-move steps — both
// half-steps, the per-step split, best tracking, histograms and chain recording — inside ONE launch (see K1Persist).
// The source's walkers and log-posteriors live in shared memory for the whole launch.
template <bool GA_FREE, int BB_MODE, bool RADIO, int CPR_T, bool LINE>
__global__ void __launch_bounds__(KC_THREADS, 2)
k1_coop_persist_kernel(const __grid_constant__ K1Params P, const K1Batch B, const CoopPlan C, const K1Persist A) {
  constexpr bool ENS = true, PERSIST = true;
  constexpr bool EXT = GA_FREE || RADIO || BB_MODE >= 2;
  constexpr int STRIDE = Rec<EXT>::STRIDE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const agnf_config& cfg = P.cfg;
  const int n_par = cfg.n_par, n_b = cfg.n_b, nbp = P.nbp;
  const bool eb_free = GA_FREE && cfg.eb_mode != 0;
  const int CPR = CPR_T ? CPR_T : C.cpr;
  const int CPRP = CPR | 1;
  // shared-window addresses
  const uint32_t tab_s = smem_u32(smem_raw);                 // 64 doubles
  const uint32_t bar_s = tab_s + 64 * 8;                     // [0] blob, [1 + w] warp w
  const uint32_t stage_s = bar_s + 16 * 8;
  const uint32_t dlut_s = stage_s + P.stage2_off_dlut, brec_s = stage_s + P.stage2_off_brec,
                 perm_s = stage_s + P.stage2_off_perm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t scal_s = stage_s + C.stage_pad + (unsigned)warp * (unsigned)C.per_warp * 8u;
  const uint32_t part_s = scal_s + 32 * STRIDE * 8;
  const uint32_t q_s = part_s + 8u * (unsigned)C.q_off;        // params / proposal rows of the warp

  const int n_warps = (int)(blockDim.x >> 5);
  if (threadIdx.x == 0) {
    for (int w = 0; w <= n_warps; ++w) mbar_init(bar_s + 8u * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < AGNF_POW2_DENOM)
    sts_f64(tab_s + 8u * threadIdx.x, c_exp2_tab[threadIdx.x * (64 / AGNF_POW2_DENOM)]);
  __syncthreads();
  if (threadIdx.x == 0 && P.stage2_bytes > 0) {
    mbar_expect_tx(bar_s, (uint32_t)P.stage2_bytes);
    bulk_g2s(stage_s, P.stage2, (uint32_t)P.stage2_bytes, bar_s);
  }
  const int64_t p_src = blockIdx.x;
  const int p_W = B.ens.W, p_H = p_W >> 1;
  // global walker indices address the shared-memory copy
  double* const w_pos = reinterpret_cast<double*>(smem_raw + C.walkers_off) - p_src * p_W * n_par;
  double* const w_lnp = reinterpret_cast<double*>(smem_raw + C.walkers_off) + (int64_t)p_W * n_par - p_src * p_W;
  double* const my_pos = w_pos + p_src * p_W * n_par;
  double* const my_lnp = w_lnp + p_src * p_W;
  int64_t n_rec = 0;
  for (int e = threadIdx.x; e < p_W * n_par; e += blockDim.x) my_pos[e] = B.ens.pos[p_src * p_W * n_par + e];
  for (int e = threadIdx.x; e < p_W; e += blockDim.x) my_lnp[e] = B.ens.lnp[p_src * p_W + e];
  __syncthreads();
  const int64_t row0 = p_src * p_H + warp * A.rows_per_warp;
  const int rows_here = max(0, min(A.rows_per_warp, p_H - warp * A.rows_per_warp));
  constexpr uint32_t par_bytes = 0;      // ensemble mode reads no params block
  // per-warp landing tile of the galaxy rows (after the walkers), its mbarrier = the warp's params barrier
  // this step's split of the source's walkers: shared memory too (the table in global memory is not used)
  const uint32_t split_off = (uint32_t)C.walkers_off + (uint32_t)(((p_W * (n_par + 1) * 8) + 15) & ~15);
  uint16_t* const my_split = reinterpret_cast<uint16_t*>(smem_raw + split_off);
  const uint16_t* const split_tab = B.ens.split ? my_split - p_src * p_W : nullptr;
  // the snap tasks, copied out of the parameter block (k1_coop_tile.inc: TILE_TASKS)
  const uint32_t tasks_off = split_off + (uint32_t)((p_W * 2 + 15) & ~15);
  SnapTask* const task_tab = reinterpret_cast<SnapTask*>(smem_raw + tasks_off);
  for (int e = threadIdx.x; e < P.n_tasks * (int)(sizeof(SnapTask) / 4); e += blockDim.x)
    reinterpret_cast<int32_t*>(task_tab)[e] = reinterpret_cast<const int32_t*>(P.tasks)[e];
  __syncthreads();
  const uint32_t ga_tile_s = tab_s + tasks_off + KC_TASK_BYTES +
                             (uint32_t)warp * 32u * (uint32_t)P.fam[0].pitch * 8u;
  uint32_t ga_parity = 0;
  for (int64_t pass = 0; pass < 2 * A.n_steps; ++pass) {
    const int64_t e_step = A.step0 + (pass >> 1);
    const int e_half = (int)(pass & 1);
    if (e_half == 0 && B.ens.split) {   // this step's split of the source's walkers (k3_split_kernel's ranking)
      uint32_t* scratch = reinterpret_cast<uint32_t*>(smem_raw + (part_s - tab_s));     // warp 0's partial-sum area
      if (warp == 0)
        split_rank_warp(my_split, scratch, p_W, e_step,
                        p_src + B.ens.src_offset, B.ens.seed, lane);
      __syncthreads();
    }
#ifdef KC_PROF
    long long tick__ = clock64();
#endif
    do {
#define TILE_EXIT break
#define TILE_PINNED
#define TILE_PERSIST
#include "k1_coop_tile.inc"
#undef TILE_PERSIST
#undef TILE_PINNED
#undef TILE_EXIT
    } while (0);
#ifdef KC_PROF
    {
      const long long now__ = clock64();
      if (lane == 0 && warp == 0) atomicAdd(&kc_prof[4], (unsigned long long)(now__ - tick__));
      tick__ = now__;
    }
#endif
    __syncthreads();                     // the other half (and the per-step work) reads what this pass wrote
#ifdef KC_PROF
    {
      const long long now__ = clock64();
      if (lane == 0 && warp == 0) atomicAdd(&kc_prof[5], (unsigned long long)(now__ - tick__));
      tick__ = now__;
    }
#endif
    if (e_half == 1) {
      const int64_t k = pass >> 1;
      const bool thinned = A.hist_every > 0 && ((k + 1) % A.hist_every) == 0;
      if (A.track_best == 1 || (A.track_best == 2 && thinned)) {
        if (warp == 0) persist_best(my_pos, my_lnp, p_W, n_par, A.best_pos + p_src * n_par, A.best_lnp + p_src, lane);
      }
      if (thinned && A.hist) persist_hist(my_pos, p_W, n_par, A.hist, A.pbox, A.n_bins, p_src);
      if (A.chain && ((e_step - A.chain_origin) % A.chain_every) == 0) {
        const int64_t entry = A.chain_n0 + n_rec++;
        const int64_t nw = A.S * p_W;
        double* dst = A.chain + (entry * nw + p_src * p_W) * n_par;
        for (int e = threadIdx.x; e < p_W * n_par; e += blockDim.x) dst[e] = my_pos[e];
        for (int e = threadIdx.x; e < p_W; e += blockDim.x) A.chain_lnp[entry * nw + p_src * p_W + e] = my_lnp[e];
      }
      __syncthreads();
    }
  }
  for (int e = threadIdx.x; e < p_W * n_par; e += blockDim.x) B.ens.pos[p_src * p_W * n_par + e] = my_pos[e];
  for (int e = threadIdx.x; e < p_W; e += blockDim.x) B.ens.lnp[p_src * p_W + e] = my_lnp[e];
}

template <bool GA_FREE, int BB_MODE, bool RADIO, int CPR_T, bool ENS, int MINB, bool LINE>
cudaError_t launch_coop_ens(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  auto kern = k1_coop_kernel<GA_FREE, BB_MODE, RADIO, CPR_T, ENS, MINB, LINE>;
  const CoopPlan C = make_plan<(GA_FREE || RADIO || BB_MODE >= 2)>(P, B.ens.pos != nullptr);
  static int attr_bytes[64] = {0};   // per instantiation and device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || C.smem_bytes > attr_bytes[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C.smem_bytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_bytes[dev] = C.smem_bytes;
  }
  const int64_t blocks = (B.n + KC_THREADS - 1) / KC_THREADS;
  kern<<<(unsigned)blocks, KC_THREADS, (size_t)C.smem_bytes, stream>>>(P, B, C);
  return cudaGetLastError();
}

// warps of a persistent CTA: the half-ensemble cut into groups of rows_per_warp walkers
inline int persist_warps(int W, int rows_per_warp) { return ((W >> 1) + rows_per_warp - 1) / rows_per_warp; }

// the persistent per-source sampler: one CTA of ceil(W/2 / rows_per_warp) warps per source
template <bool GA_FREE, int BB_MODE, bool RADIO, int CPR_T, bool LINE>
cudaError_t launch_coop_persist(const K1Params& P, const K1Batch& B, const K1Persist& A, cudaStream_t stream) {
  auto kern = k1_coop_persist_kernel<GA_FREE, BB_MODE, RADIO, CPR_T, LINE>;
  const int n_warps = persist_warps(B.ens.W, A.rows_per_warp);
  CoopPlan C = make_plan<(GA_FREE || RADIO || BB_MODE >= 2)>(P, true, n_warps);
  C.smem_bytes += (B.ens.W * (P.cfg.n_par + 1) * 8 + 15) & ~15;          // the source's walkers + log-posteriors
  C.smem_bytes += (B.ens.W * 2 + 15) & ~15;                              // this step's split
  C.smem_bytes += KC_TASK_BYTES;                                         // the snap tasks
  C.smem_bytes += n_warps * 32 * P.fam[0].pitch * 8;                     // landing tiles of the galaxy rows
  static int attr_bytes[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || C.smem_bytes > attr_bytes[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C.smem_bytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_bytes[dev] = C.smem_bytes;
  }
  kern<<<(unsigned)A.S, 32 * n_warps, (size_t)C.smem_bytes, stream>>>(P, B, C, A);
  return cudaGetLastError();
}

// This file is compiled twice: k1_coop.o holds the ordinary instantiations, k1_coop_ens.cu (which
// defines K1_COOP_ENS_TU and includes this file) the fused stretch-move ones, so the halves build in parallel.
#ifdef K1_COOP_ENS_TU
constexpr bool TU_ENS = true;
#else
constexpr bool TU_ENS = false;
#endif

template <bool GA_FREE, int BB_MODE, bool RADIO, int CPR_T, bool LINE>
cudaError_t launch_coop_cpr(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  const size_t smem = (size_t)make_plan<(GA_FREE || RADIO || BB_MODE >= 2)>(P, B.ens.pos != nullptr).smem_bytes;
  if constexpr (LINE && !TU_ENS && CPR_T != 0) {
    // The line form leaves the L1 data pipe unsaturated, so resident warps are what it runs on: four CTAs of 64
    // registers (32 warps per SM) beat three of 80 when they fit the 196 KB carve-out (C3: 1.28 -> 1.19 ms per 1e7
    // rows).  The fused stretch-move instantiations carry their proposal rows and stay at three.
    if (KC_LINE_MINB * (smem + 1024) <= 196 * 1024)
      return launch_coop_ens<GA_FREE, BB_MODE, RADIO, CPR_T, TU_ENS, KC_LINE_MINB, LINE>(P, B, stream);
  }
  if constexpr (CPR_T == 0) {
    // Three CTAs of this shape on an SM only if they fit the 196 KB shared-memory carve-out: at the next step (228 KB)
    // the L1 keeps 28 KB and the gathers of the small families miss (measured on the 49-band set: 0.81 ms per 2e6
    // rows with three CTAs of 80 registers, 0.70 ms with two of 114)
#ifndef KC_FORCE_MINB3
    if (3 * (smem + 1024) > 196 * 1024) return launch_coop_ens<GA_FREE, BB_MODE, RADIO, CPR_T, TU_ENS, 2, LINE>(P, B, stream);
#endif
  }
  return launch_coop_ens<GA_FREE, BB_MODE, RADIO, CPR_T, TU_ENS, KC_MINBLOCKS, LINE>(P, B, stream);
}

// the default 18-band set has fully unrolled instantiations; every other band count runs a runtime-trip-count one
template <bool GA_FREE, int BB_MODE, bool RADIO>
cudaError_t launch_coop(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  const bool line = coop_line_ok(P);
  switch (P.nbp / 2) {
    case 9:
      if (line) return launch_coop_cpr<GA_FREE, BB_MODE, RADIO, 9, true>(P, B, stream);
#ifdef KC_KEEP_PAIR9      // the unrolled pair form of round 1 (A/B builds; the runtime-trip-count one serves otherwise)
      return launch_coop_cpr<GA_FREE, BB_MODE, RADIO, 9, false>(P, B, stream);
#else
      return launch_coop_cpr<GA_FREE, BB_MODE, RADIO, 0, false>(P, B, stream);
#endif
    default:
      if (line) return launch_coop_cpr<GA_FREE, BB_MODE, RADIO, 0, true>(P, B, stream);
      return launch_coop_cpr<GA_FREE, BB_MODE, RADIO, 0, false>(P, B, stream);
  }
}

template <bool GA_FREE, bool RADIO>
cudaError_t launch_coop_bb(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
#ifdef KC_DEV          // quick development builds (tools/dev_build.sh): the classic model set only
  if (GA_FREE || RADIO || P.cfg.bb_mode != 1) return cudaErrorNotSupported;
  return launch_coop<false, 1, false>(P, B, stream);
#else
  switch (P.cfg.bb_mode) {
    case 0: return launch_coop<GA_FREE, 0, RADIO>(P, B, stream);
    case 1: return launch_coop<GA_FREE, 1, RADIO>(P, B, stream);
    case 2: return launch_coop<GA_FREE, 2, RADIO>(P, B, stream);
    default: return launch_coop<GA_FREE, 3, RADIO>(P, B, stream);
  }
#endif
}

#ifndef K1_COOP_ENS_TU
size_t coop_smem_bytes(const K1Params& P, bool ens) {
  const bool ext = P.cfg.ga_free_ebv != 0 || P.cfg.radio != 0 || P.cfg.bb_mode >= 2;
  return (size_t)(ext ? make_plan<true>(P, ens) : make_plan<false>(P, ens)).smem_bytes;
}

#endif

}  // namespace

#ifndef K1_COOP_ENS_TU
bool k1_coop_profitable(const K1Params& P, const K1Batch& B) {
  // The kernel hides its latencies with resident warps; when the per-warp tiles are so large that only one
  // CTA fits an SM (free E(B-V) galaxies with the energy-balance prior on ~50 bands: three partial-sum planes),
  // the per-thread kernel is the faster one (measured 2x on the freeebv_xray49 case).
  return 2 * coop_smem_bytes(P, B.ens.pos != nullptr) <= 220 * 1024;
}

// everything but the shared-memory footprint of the 8-warp CTA
static const char* coop_basic_checks(const K1Params& P, const K1Batch& B) {
  const char* w = nullptr;
  if (P.nbp & 1) w = "band pitch is odd";
  else if (!B.ens.pos && (reinterpret_cast<uintptr_t>(B.params) & 15)) w = "params is not 16-byte aligned";
  else if (B.ens.pos && P.S * (int64_t)B.ens.W >= (int64_t)INT32_MAX) w = "too many walkers for 32-bit row indices";
  else if (P.S * (int64_t)(P.nbp / 2) >= (int64_t)INT32_MAX) w = "too many sources for 32-bit chunk offsets";
  else if (P.cfg.ga_free_ebv && P.cfg.eb_mode != 0 && !P.eb0) w = "the unreddened galaxy integrals have not been built";
  else {
    for (int f = 0; f < P.cfg.n_fam && !w; ++f) {
      if (P.fam[f].pitch & 1) w = "a family table has an odd row pitch";
      else if (reinterpret_cast<uintptr_t>(P.fam[f].table) & 15) w = "a family table is not 16-byte aligned";
      else if ((int64_t)P.cfg.n_z * P.fam[f].T * (P.fam[f].pitch / 2) >= (int64_t)INT32_MAX)
        w = "a family table has 2^31 16-byte chunks or more";
    }
  }
  return w;
}

bool k1_coop_persist_supported(const K1Params& P, const K1Batch& B, int rows_per_warp) {
  if (rows_per_warp < 1 || rows_per_warp > 32) return false;
  const int n_warps = ((B.ens.W >> 1) + rows_per_warp - 1) / rows_per_warp;
  if (!B.ens.pos || B.ens.W < 2 || n_warps * 32 > KC_THREADS || coop_basic_checks(P, B)) return false;
  const bool ext = P.cfg.ga_free_ebv != 0 || P.cfg.radio != 0 || P.cfg.bb_mode >= 2;
  const CoopPlan C = ext ? make_plan<true>(P, true, n_warps) : make_plan<false>(P, true, n_warps);
  // warp 0's partial-sum area doubles as the key buffer of the per-step split
  const int64_t key_bytes = 4 * (int64_t)split_scratch_words(B.ens.W);
  return (size_t)C.smem_bytes + (size_t)B.ens.W * (P.cfg.n_par + 1) * 8 + 32 + (size_t)B.ens.W * 2 + KC_TASK_BYTES +
                 (size_t)n_warps * 32 * P.fam[0].pitch * 8 <=
             200 * 1024 && key_bytes <= (int64_t)C.q_off * 8;
}

bool k1_coop_supported(const K1Params& P, const K1Batch& B, const char** why) {
  const char* w = coop_basic_checks(P, B);
  if (!w && coop_smem_bytes(P, B.ens.pos != nullptr) > 200 * 1024) w = "grids + per-warp tiles exceed shared memory";
  if (why) *why = w;
  return w == nullptr;
}

cudaError_t launch_k1_coop_ens(const K1Params& P, const K1Batch& B, cudaStream_t stream);   // k1_coop_ens.o

cudaError_t launch_k1_coop(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  if (B.n <= 0) return cudaSuccess;
  if (B.ens.pos) return launch_k1_coop_ens(P, B, stream);
  const bool gf = P.cfg.ga_free_ebv != 0, rad = P.cfg.radio != 0;
  if (gf) return rad ? launch_coop_bb<true, true>(P, B, stream) : launch_coop_bb<true, false>(P, B, stream);
  return rad ? launch_coop_bb<false, true>(P, B, stream) : launch_coop_bb<false, false>(P, B, stream);
}

// One thread per (z, galaxy row): the chunk sums the band phase used to write to its first energy-balance plane,
// added in the order phase 3 added them -- bit-identical to evaluating them per walker.
__global__ void k1_eb0_kernel(K1Params P, double* __restrict__ eb0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const FamDesc& F = P.fam[0];
  if (i >= (int64_t)P.cfg.n_z * F.T) return;
  const int64_t zi = i / F.T;
  const double2* __restrict__ row = reinterpret_cast<const double2*>(F.table + i * F.pitch);
  const double2* __restrict__ om = reinterpret_cast<const double2*>(P.omega + zi * P.nbp);
  double acc = 0.0;
  for (int k = 0; k < P.nbp / 2; ++k) {
    const double2 d = row[k], o = om[k];
    acc += sum2_pinned(o.x, d.x * 3.826e33, o.y, d.y * 3.826e33);
  }
  eb0[i] = acc;
}

cudaError_t launch_k1_eb0(const K1Params& P, double* eb0, cudaStream_t stream) {
  const int64_t n = (int64_t)P.cfg.n_z * P.fam[0].T;
  if (n == 0) return cudaSuccess;
  k1_eb0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P, eb0);
  return cudaGetLastError();
}

// 10^x through the kernel's own routine (debug / unit test hook)
__global__ void exp10_fast_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n) {
  __shared__ __align__(16) double tab[AGNF_POW2_DENOM];
  if (threadIdx.x < AGNF_POW2_DENOM) tab[threadIdx.x] = c_exp2_tab[threadIdx.x * (64 / AGNF_POW2_DENOM)];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = pow10_via2(x[i], smem_u32(tab));
}

cudaError_t launch_exp10_fast(const double* x, double* y, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  exp10_fast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, y, n);
  return cudaGetLastError();
}

#else   // K1_COOP_ENS_TU

template <bool GA_FREE, bool RADIO>
cudaError_t launch_persist_bb(const K1Params& P, const K1Batch& B, const K1Persist& A, cudaStream_t stream) {
  const bool line = coop_line_ok(P);
  const bool c9 = P.nbp / 2 == 9;
#define PERSIST_GO(BB) \
  (line ? (c9 ? launch_coop_persist<GA_FREE, BB, RADIO, 9, true>(P, B, A, stream) : launch_coop_persist<GA_FREE, BB, RADIO, 0, true>(P, B, A, stream)) \
        : launch_coop_persist<GA_FREE, BB, RADIO, 0, false>(P, B, A, stream))
#ifdef KC_DEV
  if (GA_FREE || RADIO || P.cfg.bb_mode != 1) return cudaErrorNotSupported;
  return line ? (c9 ? launch_coop_persist<false, 1, false, 9, true>(P, B, A, stream) : launch_coop_persist<false, 1, false, 0, true>(P, B, A, stream))
              : launch_coop_persist<false, 1, false, 0, false>(P, B, A, stream);
#else
  switch (P.cfg.bb_mode) {
    case 0: return PERSIST_GO(0);
    case 1: return PERSIST_GO(1);
    case 2: return PERSIST_GO(2);
    default: return PERSIST_GO(3);
  }
#endif
#undef PERSIST_GO
}

unsigned long long k1_coop_prof_read(int i) {
#ifdef KC_PROF
  unsigned long long h[12] = {0};
  cudaMemcpyFromSymbol(h, kc_prof, sizeof(h));
  return (i >= 0 && i < 12) ? h[i] : 0;
#else
  (void)i;
  return 0;
#endif
}

cudaError_t launch_k1_coop_persist(const K1Params& P, const K1Batch& B, const K1Persist& A, cudaStream_t stream) {
  if (A.S <= 0 || A.n_steps <= 0) return cudaSuccess;
  const bool gf = P.cfg.ga_free_ebv != 0, rad = P.cfg.radio != 0;
  if (gf) return rad ? launch_persist_bb<true, true>(P, B, A, stream) : launch_persist_bb<true, false>(P, B, A, stream);
  return rad ? launch_persist_bb<false, true>(P, B, A, stream) : launch_persist_bb<false, false>(P, B, A, stream);
}


cudaError_t launch_k1_coop_ens(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  const bool gf = P.cfg.ga_free_ebv != 0, rad = P.cfg.radio != 0;
  if (gf) return rad ? launch_coop_bb<true, true>(P, B, stream) : launch_coop_bb<true, false>(P, B, stream);
  return rad ? launch_coop_bb<false, true>(P, B, stream) : launch_coop_bb<false, false>(P, B, stream);
}

#endif  // K1_COOP_ENS_TU

}  // namespace agnf
