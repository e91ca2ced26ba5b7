// K1, cooperative kernel (see k1_lnprob.cuh for the mapping to the reference).
//
// A warp owns 32 consecutive rows of `params` (= 32 evaluations).
//
//   phase 1  thread-per-evaluation.  One 1-D bulk copy (TMA unit, mbarrier completion) per CTA brings
//            the snap decision tables into shared memory, one per warp brings that warp's
//            contiguous params rows.  Each thread box-tests its row, snaps every family (decision
//            table, k1_dev.cuh::snap_fast), evaluates the normalisations 10^x, leaves a 10- or
//            14-double record in shared memory and prefetches its galaxy-template row into L2.
//   phase 2  lane-per-chunk.  The 32 x nbp tile of band values is cut into 16-byte chunks
//            (2 bands); chunk c = e * nbp/2 + k goes to lane c % 32 of iteration c / 32, so the
//            nbp/2 iterations keep all 32 lanes busy and consecutive lanes read consecutive
//            chunks of the same template row: each 144-byte row costs 2 L1 wavefronts (the
//            per-thread kernel pays 9).  A lane fetches its evaluation's record (LDS.128,
//            conflict-free: record strides 10 and 14 doubles map 8 consecutive evaluations to
//            disjoint banks), gathers one double2 per family plus y, 1/sigma and the reddening
//            coefficients, forms the model flux of its two bands and their masked chi^2 terms,
//            and stores the partial sum to shared memory.
//   phase 3  thread-per-evaluation again: sums its nbp/2 partials in band order (deterministic),
//            re-reads the two band values prior_AGNfraction needs (L1 hits), evaluates the
//            priors and writes lnP.
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
#ifndef KC_SNAP_GROUP
#define KC_SNAP_GROUP 2
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
};

template <bool EXT>
CoopPlan make_plan(const K1Params& P, bool ens) {
  CoopPlan c;
  const bool eb_free = P.cfg.ga_free_ebv != 0 && P.cfg.eb_mode != 0;
#ifdef KC_OLD_SNAP
  c.stage_pad = (P.stage_bytes + 15) & ~15;
#else
  c.stage_pad = (P.stage2_bytes + 15) & ~15;
#endif
  c.cpr = P.nbp / 2;
  c.cprp = c.cpr | 1;
  const int a = 32 * c.cprp * (eb_free ? 2 : 1), b = 32 * P.cfg.n_par;
  // plain batches land their params block in the area the partial sums take over later; ensemble mode keeps its
  // proposal rows in an area of their own until the acceptance test
  const int part = (((ens || a > b) ? a : b) + 1) & ~1;
  c.q_off = ens ? part : 0;
  c.per_warp = 32 * Rec<EXT>::STRIDE + part + (ens ? ((32 * P.cfg.n_par + 1) & ~1) : 0);
  c.smem_bytes = 64 * 8 /*exp2 table*/ + 16 * 8 /*mbarriers*/ + c.stage_pad + KC_WARPS * c.per_warp * 8;
  return c;
}

// CPR_T: chunks per row as a compile-time constant (fully unrolled band phase), or 0 = runtime
// ENS: fused stretch-move launch (propose -> ln_probab -> accept), see K1Ensemble
// MINB: CTAs per SM the register allocation is sized for (3: 80 registers; 2: 128, for shapes whose per-warp tiles
// leave room for two CTAs anyway)
template <bool GA_FREE, int BB_MODE, bool RADIO, int CPR_T, bool ENS, int MINB>
__global__ void __launch_bounds__(KC_THREADS, MINB)
k1_coop_kernel(const __grid_constant__ K1Params P, const K1Batch B, const CoopPlan C) {
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
#ifdef KC_OLD_SNAP
  const uint32_t ext_s = stage_s + P.stage_off_grid, eperm_s = stage_s + P.stage_off_perm,
                 lut_s = stage_s + P.stage_off_lut;
  const unsigned char* stage_generic = smem_raw + 64 * 8 + 16 * 8;
#else
  const uint32_t dlut_s = stage_s + P.stage2_off_dlut, brec_s = stage_s + P.stage2_off_brec,
                 perm_s = stage_s + P.stage2_off_perm;
#endif
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t scal_s = stage_s + C.stage_pad + (unsigned)warp * (unsigned)C.per_warp * 8u;
  const uint32_t part_s = scal_s + 32 * STRIDE * 8;

  const int64_t row0 = ((int64_t)blockIdx.x * KC_WARPS + warp) * 32;
  const int rows_here = (int)max((int64_t)0, min((int64_t)32, B.n - row0));
  const uint32_t q_s = part_s + 8u * (unsigned)C.q_off;        // params / proposal rows of the warp
  const int n_dbl = rows_here * n_par;
  const uint32_t par_bytes = ENS ? 0u : (uint32_t)(n_dbl >> 1) * 16u;

  if (threadIdx.x == 0) {
    for (int w = 0; w <= KC_WARPS; ++w) mbar_init(bar_s + 8u * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < AGNF_POW2_DENOM)
    sts_f64(tab_s + 8u * threadIdx.x, c_exp2_tab[threadIdx.x * (64 / AGNF_POW2_DENOM)]);
  __syncthreads();
  // one bulk copy brings the grids + cell tables; one per warp brings that warp's params rows
#ifdef KC_OLD_SNAP
  if (threadIdx.x == 0 && P.stage_bytes > 0) {
    mbar_expect_tx(bar_s, (uint32_t)P.stage_bytes);
    bulk_g2s(stage_s, P.stage, (uint32_t)P.stage_bytes, bar_s);
  }
#else
  if (threadIdx.x == 0 && P.stage2_bytes > 0) {
    mbar_expect_tx(bar_s, (uint32_t)P.stage2_bytes);
    bulk_g2s(stage_s, P.stage2, (uint32_t)P.stage2_bytes, bar_s);
  }
#endif
  if (lane == 0 && par_bytes > 0) {
    mbar_expect_tx(bar_s + 8u + 8u * warp, par_bytes);
    bulk_g2s(q_s, B.params + row0 * n_par, par_bytes, bar_s + 8u + 8u * warp);   // 256-byte aligned source
  }
  if (rows_here == 0) return;
  const bool live = lane < rows_here;
  const int my = live ? lane : rows_here - 1;      // tail lanes shadow the last row
  const int64_t i = row0 + my;

  // ================================ phase 1: thread per evaluation =================================
  int s;
  int64_t wrow = 0;                 // ensemble mode: this walker's row in pos / lnp / n_acc
  double ens_factor = 0.0, ens_lnu = 0.0, ens_lnp_old = 0.0;
  if (ENS) {
    // the proposal is drawn here instead of being read: q = c_j - (c_j - x_k) zz (k3_common.cuh)
    const int H = B.ens.W >> 1;
    const int64_t src = i / H;
    const int kw = (int)(i - src * H);
    const StretchDraw d = stretch_draw(B.ens.seed, B.ens.a, B.ens.step, B.ens.half, src + B.ens.src_offset, kw, H);
    s = (int)src;
    wrow = src * B.ens.W + split_walker(B.ens.split, src, B.ens.W, (B.ens.half ? H : 0) + kw);
    const int my_w = (int)wrow;
    const int my_c = (int)(src * B.ens.W + split_walker(B.ens.split, src, B.ens.W, (B.ens.half ? 0 : H) + d.j));
    ens_lnp_old = B.ens.lnp[wrow];          // needed only by the acceptance test: in flight during the evaluation
    {
      // Both walker rows are fetched by the whole warp: element e of the flat [32][n_par] block goes to lane
      // e % 32, so consecutive lanes read consecutive coordinates of one walker (a per-thread row read would
      // touch 32 different lines per load).  The proposal q = c - (c - x) zz is formed right here, element by
      // element (the row's stretch factor comes by shuffle), and only q is stored.  Four block rows per round:
      // their eight loads are issued back to back before the stores (a store to shared memory orders the loads
      // before it, so load-store pairs would serialise 2 n_par memory round trips).
      const int d_row = 32 / n_par, d_col = 32 - d_row * n_par;
      int row = lane / n_par, col = lane - row * n_par;
      for (int r0 = 0; r0 < n_par; r0 += 4) {
        double xv[4], cv[4], zv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          xv[u] = 0.0;
          cv[u] = 0.0;
          zv[u] = 0.0;
          if (r0 + u < n_par) {
            const int wr = __shfl_sync(0xffffffffu, my_w, row), cr = __shfl_sync(0xffffffffu, my_c, row);
            zv[u] = __shfl_sync(0xffffffffu, d.zz, row);
            xv[u] = B.ens.pos[(int64_t)wr * n_par + col];
            cv[u] = B.ens.pos[(int64_t)cr * n_par + col];
            row += d_row;
            col += d_col;
            if (col >= n_par) {
              col -= n_par;
              ++row;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (r0 + u < n_par) sts_f64(q_s + 8u * (unsigned)((r0 + u) * 32 + lane), stretch_coord(cv[u], xv[u], zv[u]));
        }
      }
    }
    __syncwarp();
    ens_factor = (double)(n_par - 1) * log(d.zz);
    ens_lnu = d.lnu;
  } else {
    s = __ldg(B.source_id + i);
    if ((n_dbl & 1) && lane == 0) sts_f64(q_s + 8u * (n_dbl - 1), __ldg(B.params + row0 * n_par + n_dbl - 1));
    if (par_bytes > 0) mbar_wait(bar_s + 8u + 8u * warp, 0);
  }
  const int zi = __ldg(P.zi + s);
  const uint64_t valid = __ldg(P.valid + s);
  const double inv4pidl2 = __ldg(P.consts + (int64_t)s * P.n_const);
  __syncwarp();
  const uint32_t p_s = q_s + 8u * (unsigned)(my * n_par);

  bool inbox = true;
  {
    const double2* __restrict__ box = P.pbox + (int64_t)s * n_par;
    int c = 0;
    if ((n_par & 1) == 0) {          // even row length: rows are 16-byte aligned -> conflict-free 128-bit reads
      for (; c < n_par; c += 2) {
        const double2 v = lds_v2f64(p_s + 8u * c);
        const double2 m0 = __ldg(box + c), m1 = __ldg(box + c + 1);
        inbox = inbox && (m0.x <= v.x) && (v.x <= m0.y) && (m1.x <= v.y) && (v.y <= m1.y);
      }
    }
    for (; c < n_par; ++c) {
      const double v = lds_f64(p_s + 8u * c);
      const double2 mm = __ldg(box + c);
      inbox = inbox && (mm.x <= v) && (v <= mm.y);
    }
  }

#ifdef KC_OLD_SNAP
  if (P.stage_bytes > 0) mbar_wait(bar_s, 0);
  const double* ext_g = reinterpret_cast<const double*>(stage_generic + P.stage_off_grid);
  const int32_t* eperm_g = reinterpret_cast<const int32_t*>(stage_generic + P.stage_off_perm);
  int row[AGNF_MAX_FAMILIES], grow[AGNF_MAX_FAMILIES];
#pragma unroll
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
    row[f] = 0;
    grow[f] = 0;
    if (f < (RADIO ? 5 : 4)) {
      const FamDesc& F = P.fam[f];
      int r = 0;
      for (int t = 0; t < F.n_snap; ++t) {
        const SnapTask& T = F.snap[t];
        r += snap_task(T, lds_f64(p_s + 8u * T.col), ext_s, eperm_s, lut_s, ext_g, eperm_g) * T.stride;
      }
      row[f] = r;
      grow[f] = zi * (int)F.T + r;
    }
  }
#else
  if (P.stage2_bytes > 0) mbar_wait(bar_s, 0);
  // the snaps of all families through the decision tables (k1_dev.cuh::snap_fast: one 16-bit read, plus one 16-byte
  // read when the cell straddles a switch point), four at a time as one straight-line block
  int row[AGNF_MAX_FAMILIES], grow[AGNF_MAX_FAMILIES];
#pragma unroll
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) row[f] = 0;
  bool slow = false;
  for (int t0 = 0; t0 < P.n_tasks; t0 += KC_SNAP_GROUP) {
    int idx[KC_SNAP_GROUP];
    bool sl[KC_SNAP_GROUP];
#pragma unroll
    for (int u = 0; u < KC_SNAP_GROUP; ++u) {
      const SnapTask& T = P.tasks[min(t0 + u, P.n_tasks - 1)];
      sl[u] = false;
      idx[u] = snap_fast(T, lds_f64(p_s + 8u * T.col), dlut_s, brec_s, perm_s, sl[u]);
    }
#pragma unroll
    for (int u = 0; u < KC_SNAP_GROUP; ++u) {
      const SnapTask& T = P.tasks[min(t0 + u, P.n_tasks - 1)];
      const bool on = t0 + u < P.n_tasks;
      const int add = on ? idx[u] * T.stride : 0;
      slow = slow || (on && sl[u]);
#pragma unroll
      for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) row[f] += (T.fam == f) ? add : 0;
    }
  }
  if (slow) {        // NaN, huge values, crowded cells: the binary search over the full grids (global memory)
#pragma unroll
    for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
      if (f < (RADIO ? 5 : 4)) row[f] = snap_family_slow(P, f, p_s);
    }
  }
#pragma unroll
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) grow[f] = f < (RADIO ? 5 : 4) ? zi * (int)P.fam[f].T + row[f] : 0;
#endif
  // start the only DRAM-bound gather early: this evaluation's galaxy row -> L2
  {
    const char* ga = reinterpret_cast<const char*>(P.fam[0].table + (int64_t)grow[0] * P.fam[0].pitch);
    prefetch_l2(ga);
    prefetch_l2(ga + 8 * n_b - 8);
  }

  const int no = P.norm_offset;
  const double GA = lds_f64(p_s + 8u * no), SB = lds_f64(p_s + 8u * no + 8u), TO = lds_f64(p_s + 8u * no + 16u);
  int kk = 3;
  double BB = 0.0, RAD = 0.0;
  if (cfg.bbb_has_norm) BB = lds_f64(p_s + 8u * (no + kk++));
  if (RADIO) RAD = lds_f64(p_s + 8u * (no + kk++));
  double ebv_gal = 0.0, ebv_bbb = 0.0, alpha = 0.0;
  if (GA_FREE) ebv_gal = lds_f64(p_s + 8u * (P.fam[0].par_offset + P.fam[0].n_dim - 1));
  if (BB_MODE == 1) ebv_bbb = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 1));
  if (BB_MODE == 2) {
    ebv_bbb = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 2));
    alpha = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 1));
  }
  if (BB_MODE == 3) {
    ebv_bbb = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 3));
    alpha = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 2));
  }
  const double nGA = pow10_via2(GA, tab_s), nSB = pow10_via2(SB, tab_s), nTO = pow10_via2(TO, tab_s);
  const double nBB = cfg.bbb_has_norm ? pow10_via2(BB, tab_s) : 1.0;
  const double nRAD = RADIO ? pow10_via2(RAD, tab_s) : 0.0;
  const double cBB = cfg.bbb_has_norm ? 1.0 : inv4pidl2;
  const double xfac = (BB_MODE >= 2) ? pow10_via2(alpha / 0.3838, tab_s) : 1.0;
  {
    const uint32_t rec = scal_s + (unsigned)lane * (STRIDE * 8);
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};\n" ::"r"(rec), "r"(grow[0] * (P.fam[0].pitch >> 1)),
                 "r"(grow[1] * (P.fam[1].pitch >> 1)), "r"(grow[2] * (P.fam[2].pitch >> 1)),
                 "r"(grow[3] * (P.fam[3].pitch >> 1))
                 : "memory");
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(rec + 16), "d"(nGA), "d"(nSB) : "memory");
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(rec + 32), "d"(nTO), "d"(nBB * cBB) : "memory");
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(rec + 48), "d"(ebv_bbb), "d"(xfac) : "memory");
    asm volatile("st.shared.v2.u64 [%0], {%1, %2};\n" ::"r"(rec + 64),
                 "l"((unsigned long long)(unsigned)(s * CPR) | ((unsigned long long)(unsigned)(zi * CPR) << 32)),
                 "l"((unsigned long long)valid)
                 : "memory");
    if (EXT) {
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(rec + 80), "d"(ebv_gal), "d"(nRAD) : "memory");
      asm volatile("st.shared.v2.u64 [%0], {%1, %2};\n" ::"r"(rec + 96),
                   "l"((unsigned long long)(unsigned)(RADIO ? grow[4] * (P.fam[4].pitch >> 1) : 0)),
                   "l"((unsigned long long)(BB_MODE >= 2 ? __ldg(P.xray + zi) : 0ull))
                   : "memory");
    }
  }
  if (B.rows && live) {
    for (int f = 0; f < cfg.n_fam; ++f) B.rows[i * cfg.n_fam + f] = row[f];
  }
  __syncwarp();      // records visible; the params block is dead and becomes the partial-sum area

  // ================================== phase 2: lane per chunk ======================================
  {
    const double2* __restrict__ tGA = reinterpret_cast<const double2*>(P.fam[0].table);
    const double2* __restrict__ tSB = reinterpret_cast<const double2*>(P.fam[1].table);
    const double2* __restrict__ tTO = reinterpret_cast<const double2*>(P.fam[2].table);
    const double2* __restrict__ tBB = reinterpret_cast<const double2*>(P.fam[3].table);
    const double2* __restrict__ tRAD = reinterpret_cast<const double2*>(P.fam[RADIO ? 4 : 0].table);
    const double2* __restrict__ y2 = reinterpret_cast<const double2*>(P.y);
    const double2* __restrict__ i2 = reinterpret_cast<const double2*>(P.isig);
    const double2* __restrict__ cb2 = reinterpret_cast<const double2*>(P.cbbb64);
    const double2* __restrict__ cg2 = reinterpret_cast<const double2*>(P.cgal64);
    const double2* __restrict__ om2 = reinterpret_cast<const double2*>(P.omega);
    int e = lane / CPR;
    int k = lane - e * CPR;
    const int de = 32 / CPR, dk = 32 - de * CPR;
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
      const uint32_t rec = scal_s + (unsigned)e * (STRIDE * 8);
      const int4 g = lds_v4s32(rec);
      const double2 n01 = lds_v2f64(rec + 16);
      const double2 n23 = lds_v2f64(rec + 32);
      const int4 sv = lds_v4s32(rec + 64);                  // {s*CPR, zi*CPR, valid lo, valid hi}
      const double2 dga = __ldg(tGA + (unsigned)(g.x + k));
      const double2 dsb = __ldg(tSB + (unsigned)(g.y + k));
      const double2 dto = __ldg(tTO + (unsigned)(g.z + k));
      const double2 dbb = __ldg(tBB + (unsigned)(g.w + k));
      const double2 yv = __ldg(y2 + (unsigned)(sv.x + k));
      const double2 iv = __ldg(i2 + (unsigned)(sv.x + k));
      double2 drad = make_double2(0.0, 0.0);
      double c_ebvgal = 0.0, c_nrad = 0.0;
      if (GA_FREE || RADIO) {
        const double2 x = lds_v2f64(rec + 80);
        c_ebvgal = x.x;
        c_nrad = x.y;
      }
      if (RADIO) drad = __ldg(tRAD + (unsigned)(lds_s32(rec + 96) + k));
      double bb0 = dbb.x, bb1 = dbb.y;
      if (BB_MODE != 0) {
        const double2 eb = lds_v2f64(rec + 48);
        const double2 cb = __ldg(cb2 + (unsigned)(sv.y + k));
        const double w0 = dbb.x * pow2_64(cb.x * eb.x, tab_s);
        const double w1 = dbb.y * pow2_64(cb.y * eb.x, tab_s);
        bb0 = (w0 != w0) ? dbb.x : w0;                        // NaN -> unreddened (MODEL:1211)
        bb1 = (w1 != w1) ? dbb.y : w1;
        if (BB_MODE >= 2) {
          const unsigned xr = (unsigned)(lds_u64(rec + 104) >> (2 * k));
          if (xr & 1u) bb0 = dbb.x * eb.y;
          if (xr & 2u) bb1 = dbb.y * eb.y;
        }
      }
      double ga0 = dga.x, ga1 = dga.y;
      if (GA_FREE) {
        const double2 cg = __ldg(cg2 + (unsigned)(sv.y + k));
        const double rg0 = pow2_64(cg.x * c_ebvgal, tab_s), rg1 = pow2_64(cg.y * c_ebvgal, tab_s);
        if (eb_free) {
          const double2 om = __ldg(om2 + (unsigned)(sv.y + k));
          sts_f64(part_s + 8u * (unsigned)(32 * CPRP + e * CPRP + k),
                  om.x * (((dga.x * 1e18) * rg0) * 3.826e33) + om.y * (((dga.y * 1e18) * rg1) * 3.826e33));
        }
        ga0 = dga.x * rg0;
        ga1 = dga.y * rg1;
      }
      double m0 = n01.y * dsb.x + n23.y * bb0 + n01.x * ga0 + n23.x * dto.x;
      double m1 = n01.y * dsb.y + n23.y * bb1 + n01.x * ga1 + n23.x * dto.y;
      if (RADIO) {
        m0 += c_nrad * drad.x;
        m1 += c_nrad * drad.y;
      }
      const unsigned long long vmask = ((unsigned long long)(unsigned)sv.w << 32) | (unsigned)sv.z;
      const unsigned vb = (unsigned)(vmask >> (2 * k));
      const double r0 = (yv.x - m0) * iv.x, r1 = (yv.y - m1) * iv.y;
      const double c0 = (vb & 1u) ? r0 * r0 : 0.0;
      sts_f64(part_s + 8u * (unsigned)(e * CPRP + k), (vb & 2u) ? fma(r1, r1, c0) : c0);
      if (B.ymodel && row0 + e < B.n) {
        double* ym = B.ymodel + (row0 + e) * n_b + 2 * k;
        ym[0] = m0;
        if (2 * k + 1 < n_b) ym[1] = m1;
      }
      k += dk;
      e += de;
      if (k >= CPR) {
        k -= CPR;
        ++e;
      }
    }
  }
  __syncwarp();

  // ================================ phase 3: thread per evaluation =================================
  if (!live && !ENS) return;          // ensemble mode keeps tail lanes for the cooperative write-back
  double chi = 0.0, ebA = 0.0, ebB = 0.0;
  {
    const uint32_t mine = part_s + 8u * (unsigned)(lane * CPRP);
#pragma unroll
    for (int k = 0; k < CPR; ++k) chi += lds_f64(mine + 8u * k);
    if (eb_free) {
      for (int k = 0; k < CPR; ++k) ebB += lds_f64(mine + 8u * (32 * CPRP + k));
    }
  }
  const double lnlike = -0.5 * chi;

  double prior = 0.0;
  if (cfg.eb_mode != 0) {
    double L_att;
    if (GA_FREE) {
      ebA = __ldg(P.eb0 + (int64_t)grow[0]);      // the unreddened integral of this galaxy row (walker-independent)
      L_att = fabs((ebA - ebB) * nGA);
    }
    else L_att = __ldg(P.fam[0].aux + row[0]) * nGA;
    const double L_sb = __ldg(P.fam[1].aux + row[1]) * nSB;
    double pr;
    if (L_sb < L_att) pr = -9999.0;
    else if (L_sb >= L_att)
      pr = (cfg.eb_mode == 1) ? 0.0 : gaussian_prior(P.lognorm_s01, 100.0, 0.0, log10(L_sb / L_att));
    else pr = NAN;   // unordered: the reference has no branch for it
    prior += pr;
  }
  if (cfg.prior_agnfrac) {
    const int j = __ldg(P.j1500 + zi);
    double ga_r = __ldg(P.fam[0].table + (int64_t)grow[0] * P.fam[0].pitch + j);
    if (GA_FREE) ga_r = ga_r * pow2_64(__ldg(P.cgal64 + (int64_t)zi * nbp + j) * ebv_gal, tab_s);
    const double dbb = __ldg(P.fam[3].table + (int64_t)grow[3] * P.fam[3].pitch + j);
    double bb_r = dbb;
    if (BB_MODE != 0) {
      if (BB_MODE >= 2 && ((__ldg(P.xray + zi) >> j) & 1ull)) {
        bb_r = dbb * xfac;
      } else {
        const double w = dbb * pow2_64(__ldg(P.cbbb64 + (int64_t)zi * nbp + j) * ebv_bbb, tab_s);
        bb_r = (w != w) ? dbb : w;
      }
    }
    double bbj = bb_r * nBB;
    if (BB == 0.0) bbj = bbj * inv4pidl2;          // decided by VALUE (PRIORS:177)
    const double frac = log10(bbj / (ga_r * nGA));
    double pr;
    if (__ldg(P.agn_branch + s) == 0) pr = gaussian_prior(P.lognorm_s2, 0.25, -2.0, frac);
    else pr = (frac < 0.0) ? -9999.0 : gaussian_prior(P.lognorm_s2, 0.25, 2.0, frac);
    prior += pr + gaussian_prior(P.lognorm_s15, 1.0 / 2.25, 4.5, GA);
  }
  double tor6 = 0.0, lumf = 0.0;
  if (P.xp_midir_uv | P.xp_ir_xray) {
    tor6 = __ldg(P.fam[2].table + (int64_t)grow[2] * P.fam[2].pitch + P.xp_j6) * nTO;
    lumf = __ldg(P.consts + (int64_t)s * P.n_const + 1);
  }
  if (P.xp_midir_uv) {
    const int64_t dr = (int64_t)zi * P.fam[3].T + __ldg(P.xp_bb_dered + row[3]);
    const double bb2500 = __ldg(P.fam[3].table + dr * P.fam[3].pitch + P.xp_j2500) * nBB;
    prior += prior_midir_uv(P, tor6, bb2500, BB, lumf, __ldg(P.consts + (int64_t)s * P.n_const + 3));
  }
  if (RADIO) {
    const int mode = __ldg(P.irsyn_mode + s);
    if (mode != 0) {
      const int j = __ldg(P.irsyn_j + s);
      const double sb_ir = __ldg(P.fam[1].table + (int64_t)grow[1] * P.fam[1].pitch + j) * nSB;
      const double syn_ir = __ldg(P.fam[4].table + (int64_t)grow[4] * P.fam[4].pitch + j) * nRAD;
      prior += gaussian_prior(P.lognorm_s2, 0.25, mode > 0 ? 2.0 : -2.0, log10(syn_ir / sb_ir));
    }
  }
  if (P.xp_ir_xray) prior += prior_ir_xrays(P, tor6, lumf, __ldg(P.consts + (int64_t)s * P.n_const + 2));
  const double lnprior = inbox ? prior : -INFINITY;
  const bool fin = (lnprior - lnprior) == 0.0;     // isfinite
  const double lnp_new = fin ? lnprior + lnlike : -INFINITY;
  if (ENS) {
    // emcee's acceptance rule, in place: only this thread touches walker `wrow` during the launch, and the
    // partners it read belong to the other (frozen) half
    const double lnpdiff = ens_factor + lnp_new - ens_lnp_old;
    const bool accept = live && (lnpdiff > ens_lnu);          // false for NaN and for -inf proposals
    if (accept) {
      B.ens.lnp[wrow] = lnp_new;
      atomicAdd(B.ens.n_acc + wrow, 1);                       // result unused: a fire-and-forget reduction
    }
    // accepted rows go back to `pos` the way they came: consecutive lanes store consecutive coordinates
    const unsigned acc_mask = __ballot_sync(0xffffffffu, accept);
    if (acc_mask) {
      const int my_w = (int)wrow;
      const int d_row = 32 / n_par, d_col = 32 - d_row * n_par;
      int row = lane / n_par, col = lane - row * n_par;
      for (int r = 0; r < n_par; ++r) {
        const int wr = __shfl_sync(0xffffffffu, my_w, row);
        if ((acc_mask >> row) & 1u) B.ens.pos[(int64_t)wr * n_par + col] = lds_f64(q_s + 8u * (unsigned)(r * 32 + lane));
        row += d_row;
        col += d_col;
        if (col >= n_par) {
          col -= n_par;
          ++row;
        }
      }
    }
    return;
  }
  B.lnp[i] = lnp_new;
  if (B.lnprior) B.lnprior[i] = lnprior;
  if (B.lnlike) B.lnlike[i] = lnlike;
}

template <bool GA_FREE, int BB_MODE, bool RADIO, int CPR_T, bool ENS, int MINB>
cudaError_t launch_coop_ens(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  auto kern = k1_coop_kernel<GA_FREE, BB_MODE, RADIO, CPR_T, ENS, MINB>;
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

// This file is compiled twice: k1_coop.o holds the ordinary instantiations, k1_coop_ens.cu (which
// defines K1_COOP_ENS_TU and includes this file) the fused stretch-move ones, so the halves build in parallel.
#ifdef K1_COOP_ENS_TU
constexpr bool TU_ENS = true;
#else
constexpr bool TU_ENS = false;
#endif

template <bool GA_FREE, int BB_MODE, bool RADIO, int CPR_T>
cudaError_t launch_coop_cpr(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  if (CPR_T == 0) {
    // Three CTAs of this shape on an SM only if they fit the 196 KB shared-memory carve-out: at the next step (228 KB)
    // the L1 keeps 28 KB and the gathers of the small families miss (measured on the 49-band set: 0.81 ms per 2e6
    // rows with three CTAs of 80 registers, 0.70 ms with two of 114)
    const size_t smem = (size_t)make_plan<(GA_FREE || RADIO || BB_MODE >= 2)>(P, B.ens.pos != nullptr).smem_bytes;
#ifndef KC_FORCE_MINB3
    if (3 * (smem + 1024) > 196 * 1024) return launch_coop_ens<GA_FREE, BB_MODE, RADIO, CPR_T, TU_ENS, 2>(P, B, stream);
#endif
  }
  return launch_coop_ens<GA_FREE, BB_MODE, RADIO, CPR_T, TU_ENS, KC_MINBLOCKS>(P, B, stream);
}

// the default 18-band set has a fully unrolled instantiation; every other band count runs the
// runtime-trip-count instantiation
template <bool GA_FREE, int BB_MODE, bool RADIO>
cudaError_t launch_coop(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  switch (P.nbp / 2) {
    case 9: return launch_coop_cpr<GA_FREE, BB_MODE, RADIO, 9>(P, B, stream);
    default: return launch_coop_cpr<GA_FREE, BB_MODE, RADIO, 0>(P, B, stream);
  }
}

template <bool GA_FREE, bool RADIO>
cudaError_t launch_coop_bb(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  switch (P.cfg.bb_mode) {
    case 0: return launch_coop<GA_FREE, 0, RADIO>(P, B, stream);
    case 1: return launch_coop<GA_FREE, 1, RADIO>(P, B, stream);
    case 2: return launch_coop<GA_FREE, 2, RADIO>(P, B, stream);
    default: return launch_coop<GA_FREE, 3, RADIO>(P, B, stream);
  }
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

bool k1_coop_supported(const K1Params& P, const K1Batch& B, const char** why) {
  const char* w = nullptr;
  if (P.nbp & 1) w = "band pitch is odd";
  else if (!B.ens.pos && (reinterpret_cast<uintptr_t>(B.params) & 15)) w = "params is not 16-byte aligned";
  else if (coop_smem_bytes(P, B.ens.pos != nullptr) > 200 * 1024) w = "grids + per-warp tiles exceed shared memory";
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
    acc += o.x * (d.x * 3.826e33) + o.y * (d.y * 3.826e33);
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

cudaError_t launch_k1_coop_ens(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  const bool gf = P.cfg.ga_free_ebv != 0, rad = P.cfg.radio != 0;
  if (gf) return rad ? launch_coop_bb<true, true>(P, B, stream) : launch_coop_bb<true, false>(P, B, stream);
  return rad ? launch_coop_bb<false, true>(P, B, stream) : launch_coop_bb<false, false>(P, B, stream);
}

#endif  // K1_COOP_ENS_TU

}  // namespace agnf
