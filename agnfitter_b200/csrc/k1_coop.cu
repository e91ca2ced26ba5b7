// K1, cooperative kernel (see k1_lnprob.cuh for the mapping to the reference).
//
// A warp owns 32 consecutive rows of `params` (= 32 evaluations).
//
//   phase 1  thread-per-evaluation.  The warp's params block is contiguous in global memory and is
//            staged with 16-byte cp.async; each thread then box-tests its row, snaps every family
//            (cell table + guarded walk, k1_common.cuh), evaluates the normalisations 10^x and
//            leaves a 10- or 14-double record in shared memory.
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
// Only __syncwarp() separates the phases; the CTA synchronises once, after staging the grids.
#include "k1_lnprob.cuh"
#include "k1_common.cuh"

namespace agnf {

namespace {

constexpr int KC_THREADS = 256;
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
  int32_t per_warp;       // doubles per warp: 32 records + max(params block, partial planes)
  int32_t smem_bytes;
};

template <bool EXT>
CoopPlan make_plan(const K1Params& P) {
  CoopPlan c;
  const bool eb_free = P.cfg.ga_free_ebv != 0 && P.cfg.eb_mode != 0;
  c.stage_pad = (P.stage_bytes + 15) & ~15;
  c.cpr = P.nbp / 2;
  c.cprp = c.cpr | 1;
  const int a = 32 * c.cprp * (eb_free ? 3 : 1), b = 32 * P.cfg.n_par;
  const int part = ((a > b ? a : b) + 1) & ~1;
  c.per_warp = 32 * Rec<EXT>::STRIDE + part;
  c.smem_bytes = 64 * 8 /*exp2 table*/ + 16 * 8 /*mbarriers*/ + c.stage_pad + KC_WARPS * c.per_warp * 8;
  return c;
}

// ---- mbarrier + bulk copy (TMA unit, 1-D) -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spin > (1 << 24)) __trap();      // a lost copy must not hang the device
  }
}

template <bool GA_FREE, int BB_MODE, bool RADIO>
__global__ void __launch_bounds__(KC_THREADS, 3)
k1_coop_kernel(const __grid_constant__ K1Params P, const K1Batch B, const CoopPlan C) {
  constexpr bool EXT = GA_FREE || RADIO || BB_MODE >= 2;
  constexpr int STRIDE = Rec<EXT>::STRIDE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const agnf_config& cfg = P.cfg;
  const int n_par = cfg.n_par, n_b = cfg.n_b, nbp = P.nbp;
  const bool eb_free = GA_FREE && cfg.eb_mode != 0;
  double* s_tab = reinterpret_cast<double*>(smem_raw);                        // 64
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tab + 64);                  // [0] blob, [1 + w] warp w
  unsigned char* s_stage = reinterpret_cast<unsigned char*>(s_bar + 16);
  const double* s_ext = reinterpret_cast<const double*>(s_stage + P.stage_off_grid);
  const int32_t* s_eperm = reinterpret_cast<const int32_t*>(s_stage + P.stage_off_perm);
  const uint16_t* s_lut = reinterpret_cast<const uint16_t*>(s_stage + P.stage_off_lut);
  double* s_warp = reinterpret_cast<double*>(s_stage + C.stage_pad);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* scal = s_warp + (size_t)warp * C.per_warp;
  double* part = scal + 32 * STRIDE;
  const int64_t row0 = ((int64_t)blockIdx.x * KC_WARPS + warp) * 32;
  const int rows_here = (int)max((int64_t)0, min((int64_t)32, B.n - row0));
  const int n_dbl = rows_here * n_par;
  const uint32_t par_bytes = (uint32_t)(n_dbl >> 1) * 16u;

  if (threadIdx.x == 0) {
    for (int w = 0; w <= KC_WARPS; ++w) mbar_init(s_bar + w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < 64) s_tab[threadIdx.x] = c_exp2_tab[threadIdx.x];
  __syncthreads();
  // one bulk copy brings the grids + cell tables; one per warp brings that warp's params rows
  if (threadIdx.x == 0) {
    if (P.stage_bytes > 0) {
      mbar_expect_tx(s_bar, (uint32_t)P.stage_bytes);
      bulk_g2s(s_stage, P.stage, (uint32_t)P.stage_bytes, s_bar);
    }
  }
  if (lane == 0 && par_bytes > 0) {
    mbar_expect_tx(s_bar + 1 + warp, par_bytes);
    bulk_g2s(part, B.params + row0 * n_par, par_bytes, s_bar + 1 + warp);   // row0*n_par*8 is a multiple of 256
  }
  if (rows_here == 0) return;
  const bool live = lane < rows_here;
  const int my = live ? lane : rows_here - 1;      // tail lanes shadow the last row
  const int64_t i = row0 + my;

  // ================================ phase 1: thread per evaluation =================================
  const int s = __ldg(B.source_id + i);
  const int zi = __ldg(P.zi + s);
  const uint64_t valid = __ldg(P.valid + s);
  const double inv4pidl2 = __ldg(P.consts + (int64_t)s * P.n_const);
  if ((n_dbl & 1) && lane == 0) part[n_dbl - 1] = __ldg(B.params + row0 * n_par + n_dbl - 1);
  if (par_bytes > 0) mbar_wait(s_bar + 1 + warp, 0);
  __syncwarp();
  const double* __restrict__ p = part + my * n_par;

  bool inbox = true;
  {
    const double2* __restrict__ box = P.pbox + (int64_t)s * n_par;
    for (int c = 0; c < n_par; ++c) {
      const double v = p[c];
      const double2 mm = __ldg(box + c);
      inbox = inbox && (mm.x <= v) && (v <= mm.y);
    }
  }

  if (P.stage_bytes > 0) mbar_wait(s_bar, 0);
  int row[AGNF_MAX_FAMILIES], grow[AGNF_MAX_FAMILIES];
#pragma unroll
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
    row[f] = 0;
    grow[f] = 0;
    if (f < (RADIO ? 5 : 4)) {
      const FamDesc& F = P.fam[f];
      int64_t r = 0;
      for (int d = 0; d < F.n_dim; ++d) {
        const DimDesc& D = F.dim[d];
        if (D.is_grid && D.size > 1) {
          const double pv = p[F.par_offset + d];
          int idx;
          if (D.lut_m > 0)
            idx = snap_dim_lut(s_ext + D.ext_off, s_eperm + D.ext_off, D.size, pv, s_lut + D.lut_off, D.lut_m,
                               D.lut_lo, D.lut_inv, D.lut_big);
          else
            idx = snap_dim(s_ext + D.ext_off + 1, s_eperm + D.ext_off + 1, D.size, pv);
          r += (int64_t)idx * D.stride;
        }
      }
      row[f] = (int)r;
      grow[f] = (int)((int64_t)zi * F.T + r);
    }
  }

  const int no = P.norm_offset;
  const double GA = p[no], SB = p[no + 1], TO = p[no + 2];
  int kk = 3;
  double BB = 0.0, RAD = 0.0;
  if (cfg.bbb_has_norm) BB = p[no + kk++];
  if (RADIO) RAD = p[no + kk++];
  double ebv_gal = 0.0, ebv_bbb = 0.0, alpha = 0.0;
  if (GA_FREE) ebv_gal = p[P.fam[0].par_offset + P.fam[0].n_dim - 1];
  if (BB_MODE == 1) ebv_bbb = p[P.fam[3].par_offset + P.fam[3].n_dim - 1];
  if (BB_MODE == 2) {
    ebv_bbb = p[P.fam[3].par_offset + P.fam[3].n_dim - 2];
    alpha = p[P.fam[3].par_offset + P.fam[3].n_dim - 1];
  }
  if (BB_MODE == 3) {
    ebv_bbb = p[P.fam[3].par_offset + P.fam[3].n_dim - 3];
    alpha = p[P.fam[3].par_offset + P.fam[3].n_dim - 2];
  }
  const double nGA = exp10_fast(GA, s_tab), nSB = exp10_fast(SB, s_tab), nTO = exp10_fast(TO, s_tab);
  const double nBB = cfg.bbb_has_norm ? exp10_fast(BB, s_tab) : 1.0;
  const double nRAD = RADIO ? exp10_fast(RAD, s_tab) : 0.0;
  const double cBB = cfg.bbb_has_norm ? 1.0 : inv4pidl2;
  const double xfac = (BB_MODE >= 2) ? exp10_fast(alpha / 0.3838, s_tab) : 1.0;
  const int CPR = C.cpr, CPRP = C.cprp;
  {
    double* rec = scal + lane * STRIDE;
    *reinterpret_cast<int4*>(rec) =
        make_int4(grow[0] * (P.fam[0].pitch >> 1), grow[1] * (P.fam[1].pitch >> 1),
                  grow[2] * (P.fam[2].pitch >> 1), grow[3] * (P.fam[3].pitch >> 1));
    *reinterpret_cast<double2*>(rec + 2) = make_double2(nGA, nSB);
    *reinterpret_cast<double2*>(rec + 4) = make_double2(nTO, nBB * cBB);
    *reinterpret_cast<double2*>(rec + 6) = make_double2(ebv_bbb, xfac);
    *reinterpret_cast<ulonglong2*>(rec + 8) = make_ulonglong2(
        (unsigned long long)(unsigned)(s * CPR) | ((unsigned long long)(unsigned)(zi * CPR) << 32), valid);
    if (EXT) {
      *reinterpret_cast<double2*>(rec + 10) = make_double2(ebv_gal, nRAD);
      *reinterpret_cast<ulonglong2*>(rec + 12) = make_ulonglong2(
          (unsigned long long)(unsigned)(RADIO ? grow[4] * (P.fam[4].pitch >> 1) : 0),
          BB_MODE >= 2 ? __ldg(P.xray + zi) : 0ull);
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
    const double2* __restrict__ cb2 = reinterpret_cast<const double2*>(P.cbbb);
    const double2* __restrict__ cg2 = reinterpret_cast<const double2*>(P.cgal);
    const double2* __restrict__ om2 = reinterpret_cast<const double2*>(P.omega);
    int e = lane / CPR;
    int k = lane - e * CPR;
    const int de = 32 / CPR, dk = 32 - de * CPR;
#pragma unroll 3
    for (int it = 0; it < CPR; ++it) {
      const double* rec = scal + e * STRIDE;
      const int4 g = *reinterpret_cast<const int4*>(rec);
      const double2 n01 = *reinterpret_cast<const double2*>(rec + 2);
      const double2 n23 = *reinterpret_cast<const double2*>(rec + 4);
      const ulonglong2 sv = *reinterpret_cast<const ulonglong2*>(rec + 8);
      const unsigned qs = (unsigned)(sv.x & 0xffffffffull), qz = (unsigned)(sv.x >> 32);
      const double2 dga = __ldg(tGA + (unsigned)(g.x + k));
      const double2 dsb = __ldg(tSB + (unsigned)(g.y + k));
      const double2 dto = __ldg(tTO + (unsigned)(g.z + k));
      const double2 dbb = __ldg(tBB + (unsigned)(g.w + k));
      const double2 yv = __ldg(y2 + (qs + k));
      const double2 iv = __ldg(i2 + (qs + k));
      double2 drad = make_double2(0.0, 0.0);
      double c_ebvgal = 0.0, c_nrad = 0.0;
      if (GA_FREE || RADIO) {
        const double2 x = *reinterpret_cast<const double2*>(rec + 10);
        c_ebvgal = x.x;
        c_nrad = x.y;
      }
      if (RADIO) {
        const int gr = *reinterpret_cast<const int*>(rec + 12);
        drad = __ldg(tRAD + (unsigned)(gr + k));
      }
      double bb0 = dbb.x, bb1 = dbb.y;
      if (BB_MODE != 0) {
        const double2 eb = *reinterpret_cast<const double2*>(rec + 6);
        const double2 cb = __ldg(cb2 + (qz + k));
        const double w0 = dbb.x * exp10_fast(cb.x * eb.x, s_tab);
        const double w1 = dbb.y * exp10_fast(cb.y * eb.x, s_tab);
        bb0 = (w0 != w0) ? dbb.x : w0;                        // NaN -> unreddened (MODEL:1211)
        bb1 = (w1 != w1) ? dbb.y : w1;
        if (BB_MODE >= 2) {
          const unsigned xr = (unsigned)(*reinterpret_cast<const unsigned long long*>(rec + 13) >> (2 * k));
          if (xr & 1u) bb0 = dbb.x * eb.y;
          if (xr & 2u) bb1 = dbb.y * eb.y;
        }
      }
      double ga0 = dga.x, ga1 = dga.y;
      if (GA_FREE) {
        const double2 cg = __ldg(cg2 + (qz + k));
        const double rg0 = exp10_fast(cg.x * c_ebvgal, s_tab), rg1 = exp10_fast(cg.y * c_ebvgal, s_tab);
        if (eb_free) {
          const double2 om = __ldg(om2 + (qz + k));
          part[32 * CPRP + e * CPRP + k] = om.x * (dga.x * 3.826e33) + om.y * (dga.y * 3.826e33);
          part[64 * CPRP + e * CPRP + k] =
              om.x * (((dga.x * 1e18) * rg0) * 3.826e33) + om.y * (((dga.y * 1e18) * rg1) * 3.826e33);
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
      const unsigned vb = (unsigned)(sv.y >> (2 * k));
      const double r0 = (yv.x - m0) * iv.x, r1 = (yv.y - m1) * iv.y;
      const double c0 = (vb & 1u) ? r0 * r0 : 0.0;
      part[e * CPRP + k] = (vb & 2u) ? fma(r1, r1, c0) : c0;
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
  if (!live) return;
  double chi = 0.0, ebA = 0.0, ebB = 0.0;
  {
    const double* mine = part + lane * CPRP;
    for (int k = 0; k < CPR; ++k) chi += mine[k];
    if (eb_free) {
      for (int k = 0; k < CPR; ++k) ebA += mine[32 * CPRP + k];
      for (int k = 0; k < CPR; ++k) ebB += mine[64 * CPRP + k];
    }
  }
  const double lnlike = -0.5 * chi;

  double prior = 0.0;
  if (cfg.eb_mode != 0) {
    double L_att;
    if (GA_FREE) L_att = fabs((ebA - ebB) * nGA);
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
    if (GA_FREE) ga_r = ga_r * exp10_fast(__ldg(P.cgal + (int64_t)zi * nbp + j) * ebv_gal, s_tab);
    const double dbb = __ldg(P.fam[3].table + (int64_t)grow[3] * P.fam[3].pitch + j);
    double bb_r = dbb;
    if (BB_MODE != 0) {
      if (BB_MODE >= 2 && ((__ldg(P.xray + zi) >> j) & 1ull)) {
        bb_r = dbb * xfac;
      } else {
        const double w = dbb * exp10_fast(__ldg(P.cbbb + (int64_t)zi * nbp + j) * ebv_bbb, s_tab);
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
  if (RADIO) {
    const int mode = __ldg(P.irsyn_mode + s);
    if (mode != 0) {
      const int j = __ldg(P.irsyn_j + s);
      const double sb_ir = __ldg(P.fam[1].table + (int64_t)grow[1] * P.fam[1].pitch + j) * nSB;
      const double syn_ir = __ldg(P.fam[4].table + (int64_t)grow[4] * P.fam[4].pitch + j) * nRAD;
      prior += gaussian_prior(P.lognorm_s2, 0.25, mode > 0 ? 2.0 : -2.0, log10(syn_ir / sb_ir));
    }
  }
  const double lnprior = inbox ? prior : -INFINITY;
  const bool fin = (lnprior - lnprior) == 0.0;     // isfinite
  B.lnp[i] = fin ? lnprior + lnlike : -INFINITY;
  if (B.lnprior) B.lnprior[i] = lnprior;
  if (B.lnlike) B.lnlike[i] = lnlike;
}

template <bool GA_FREE, int BB_MODE, bool RADIO>
cudaError_t launch_coop(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  auto kern = k1_coop_kernel<GA_FREE, BB_MODE, RADIO>;
  const CoopPlan C = make_plan<(GA_FREE || RADIO || BB_MODE >= 2)>(P);
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

template <bool GA_FREE, bool RADIO>
cudaError_t launch_coop_bb(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  switch (P.cfg.bb_mode) {
    case 0: return launch_coop<GA_FREE, 0, RADIO>(P, B, stream);
    case 1: return launch_coop<GA_FREE, 1, RADIO>(P, B, stream);
    case 2: return launch_coop<GA_FREE, 2, RADIO>(P, B, stream);
    default: return launch_coop<GA_FREE, 3, RADIO>(P, B, stream);
  }
}

size_t coop_smem_bytes(const K1Params& P) {
  const bool ext = P.cfg.ga_free_ebv != 0 || P.cfg.radio != 0 || P.cfg.bb_mode >= 2;
  return (size_t)(ext ? make_plan<true>(P) : make_plan<false>(P)).smem_bytes;
}

}  // namespace

bool k1_coop_supported(const K1Params& P, const K1Batch& B, const char** why) {
  const char* w = nullptr;
  if (P.nbp & 1) w = "band pitch is odd";
  else if (reinterpret_cast<uintptr_t>(B.params) & 15) w = "params is not 16-byte aligned";
  else if (coop_smem_bytes(P) > 200 * 1024) w = "grids + per-warp tiles exceed shared memory";
  else if (P.S * (int64_t)(P.nbp / 2) >= (int64_t)INT32_MAX) w = "too many sources for 32-bit chunk offsets";
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

cudaError_t launch_k1_coop(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  if (B.n <= 0) return cudaSuccess;
  const bool gf = P.cfg.ga_free_ebv != 0, rad = P.cfg.radio != 0;
  if (gf) return rad ? launch_coop_bb<true, true>(P, B, stream) : launch_coop_bb<true, false>(P, B, stream);
  return rad ? launch_coop_bb<false, true>(P, B, stream) : launch_coop_bb<false, false>(P, B, stream);
}

// 10^x through the kernels' own exp10_fast (debug / unit test hook)
__global__ void exp10_fast_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n) {
  __shared__ double tab[64];
  if (threadIdx.x < 64) tab[threadIdx.x] = c_exp2_tab[threadIdx.x];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = exp10_fast(x[i], tab);
}

cudaError_t launch_exp10_fast(const double* x, double* y, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  exp10_fast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, y, n);
  return cudaGetLastError();
}

}  // namespace agnf
