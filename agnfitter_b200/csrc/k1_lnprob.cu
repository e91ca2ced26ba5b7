// K1 — batched log-posterior kernel (see k1_lnprob.cuh for the mapping to the reference).
#include "k1_lnprob.cuh"
#include "k1_common.cuh"

#include <cfloat>

namespace agnf {

namespace {

constexpr int K1_THREADS = 128;

struct BandAcc {
  double chi = 0.0;
  double ga_j = 0.0, bb_j = 0.0;      // band j1500, after reddening and norm
  double sb_ir = 0.0, syn_ir = 0.0;   // band j_IR
  double ebA = 0.0, ebB = 0.0;        // energy-balance integrals (free E(B-V)_gal)
};

struct WalkerCtx {
  double nGA, nSB, nTO, nBB, nRAD;
  double cBB;        // 1 or 1/(4 pi dL^2) applied to the BBB row in the model sum
  double ebv_gal, ebv_bbb, xfac;
  uint64_t valid, xray;
  int j1500, jir;
};

template <bool GA_FREE, int BB_MODE, bool RADIO>
__device__ __forceinline__ double band_step(const agnf_config& cfg, const WalkerCtx& w, BandAcc& a, int b,
                                            double dga, double dsb, double dto, double dbb, double drad,
                                            double yb, double isb, double cg, double cb, double om,
                                            bool eb_free) {
  double rga = 1.0;
  if (GA_FREE) rga = exp10(cg * w.ebv_gal);
  const double ga_r = GA_FREE ? dga * rga : dga;
  double bb_r = dbb;
  if (BB_MODE != 0) {
    if (BB_MODE >= 2 && ((w.xray >> b) & 1ull)) {
      bb_r = dbb * w.xfac;
    } else {
      const double v = dbb * exp10(cb * w.ebv_bbb);
      bb_r = (v != v) ? dbb : v;      // NaN -> unreddened (MODEL:1211)
    }
  }
  double m = w.nSB * dsb + w.nBB * (bb_r * w.cBB) + w.nGA * ga_r + w.nTO * dto;
  if (RADIO) m += w.nRAD * drad;
  if ((w.valid >> b) & 1ull) {
    const double r = (yb - m) * isb;
    a.chi += r * r;
  }
  if (b == w.j1500) {
    a.ga_j = ga_r * w.nGA;
    a.bb_j = bb_r * w.nBB;
  }
  if (RADIO && b == w.jir) {
    a.sb_ir = dsb * w.nSB;
    a.syn_ir = drad * w.nRAD;
  }
  if (GA_FREE && eb_free) {
    a.ebA += om * (dga * 3.826e33);
    a.ebB += om * (((dga * 1e18) * rga) * 3.826e33);
  }
  return m;
}

template <bool GA_FREE, int BB_MODE, bool RADIO>
__global__ void __launch_bounds__(K1_THREADS)
k1_lnprob_kernel(const __grid_constant__ K1Params P, const K1Batch B) {
  extern __shared__ double s_raw[];
  double* s_grid = s_raw;
  int32_t* s_perm = reinterpret_cast<int32_t*>(s_raw + P.n_grid_vals);
  for (int i = threadIdx.x; i < P.n_grid_vals; i += blockDim.x) {
    s_grid[i] = P.grid_sorted[i];
    s_perm[i] = P.grid_perm[i];
  }
  __syncthreads();

  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n) return;
  const agnf_config& cfg = P.cfg;
  const int n_b = cfg.n_b, n_par = cfg.n_par;
  const int s = __ldg(B.source_id + i);
  const double* __restrict__ p = B.params + i * n_par;

  // ---- box prior -------------------------------------------------------------------------
  bool inbox = true;
  {
    const double* __restrict__ pmin = P.pmin + (int64_t)s * n_par;
    const double* __restrict__ pmax = P.pmax + (int64_t)s * n_par;
    for (int c = 0; c < n_par; ++c) {
      const double v = __ldg(p + c);
      inbox = inbox && (__ldg(pmin + c) <= v) && (v <= __ldg(pmax + c));
    }
  }
  const bool detail = (B.rows != nullptr) || (B.ymodel != nullptr) || (B.lnprior != nullptr) ||
                      (B.lnlike != nullptr);
  if (!inbox && !detail) {   // ymodel's value is discarded by the reference in this case
    B.lnp[i] = -INFINITY;
    return;
  }

  // ---- snap every family to its dictionary row ------------------------------------------------
  const int zi = __ldg(P.zi + s);
  const double* trow[AGNF_MAX_FAMILIES];
  int64_t row[AGNF_MAX_FAMILIES];
#pragma unroll
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
    row[f] = 0;
    trow[f] = nullptr;
    if (f < cfg.n_fam) {
      const FamDesc& F = P.fam[f];
      int64_t r = 0;
      for (int d = 0; d < F.n_dim; ++d) {
        if (F.dim[d].is_grid) {
          const int off = F.dim[d].smem_off;
          const int idx = snap_dim(s_grid + off, s_perm + off, F.dim[d].size, __ldg(p + F.par_offset + d));
          r += (int64_t)idx * F.dim[d].stride;
        }
      }
      row[f] = r;
      trow[f] = F.table + ((int64_t)zi * F.T + r) * F.pitch;
    }
  }
  if (B.rows) {
    for (int f = 0; f < cfg.n_fam; ++f) B.rows[i * cfg.n_fam + f] = (int32_t)row[f];
  }

  // ---- normalisations and free parameters ----------------------------------------------------
  WalkerCtx w;
  const int no = P.norm_offset;
  const double GA = __ldg(p + no), SB = __ldg(p + no + 1), TO = __ldg(p + no + 2);
  int k = 3;
  double BB = 0.0, RAD = 0.0;
  if (cfg.bbb_has_norm) BB = __ldg(p + no + k++);
  if (RADIO) RAD = __ldg(p + no + k++);
  const double inv4pidl2 = __ldg(P.consts + (int64_t)s * P.n_const + 0);
  w.nGA = exp10(GA);
  w.nSB = exp10(SB);
  w.nTO = exp10(TO);
  w.nBB = cfg.bbb_has_norm ? exp10(BB) : 1.0;
  w.nRAD = RADIO ? exp10(RAD) : 0.0;
  w.cBB = cfg.bbb_has_norm ? 1.0 : inv4pidl2;
  w.ebv_gal = 0.0;
  w.ebv_bbb = 0.0;
  w.xfac = 1.0;
  if (GA_FREE) w.ebv_gal = __ldg(p + P.fam[0].par_offset + P.fam[0].n_dim - 1);
  if (BB_MODE == 1) w.ebv_bbb = __ldg(p + P.fam[3].par_offset + P.fam[3].n_dim - 1);
  if (BB_MODE == 2) {
    w.ebv_bbb = __ldg(p + P.fam[3].par_offset + P.fam[3].n_dim - 2);
    w.xfac = exp10(__ldg(p + P.fam[3].par_offset + P.fam[3].n_dim - 1) / 0.3838);
  }
  if (BB_MODE == 3) {
    w.ebv_bbb = __ldg(p + P.fam[3].par_offset + P.fam[3].n_dim - 3);
    w.xfac = exp10(__ldg(p + P.fam[3].par_offset + P.fam[3].n_dim - 2) / 0.3838);
  }
  w.valid = __ldg(P.valid + s);
  w.xray = __ldg(P.xray + zi);
  w.j1500 = __ldg(P.j1500 + zi);
  w.jir = RADIO ? __ldg(P.irsyn_j + s) : -1;
  const bool eb_free = GA_FREE && cfg.eb_mode != 0;

  // ---- band loop: gather, model, chi^2 -------------------------------------------------------
  const double* __restrict__ tGA = trow[0];
  const double* __restrict__ tSB = trow[1];
  const double* __restrict__ tTO = trow[2];
  const double* __restrict__ tBB = trow[3];
  const double* __restrict__ tRAD = RADIO ? trow[4] : trow[0];
  const double* __restrict__ ys = P.y + (int64_t)s * P.nbp;
  const double* __restrict__ is = P.isig + (int64_t)s * P.nbp;
  const double* __restrict__ cg = P.cgal + (int64_t)zi * P.nbp;
  const double* __restrict__ cb = P.cbbb + (int64_t)zi * P.nbp;
  const double* __restrict__ om = P.omega + (int64_t)zi * P.nbp;
  double* ym = B.ymodel ? B.ymodel + i * n_b : nullptr;
  BandAcc a;
  {
    for (int b = 0; b < n_b; ++b) {
      const double drad = RADIO ? __ldg(tRAD + b) : 0.0;
      const double cgb = GA_FREE ? __ldg(cg + b) : 0.0;
      const double omb = GA_FREE ? __ldg(om + b) : 0.0;
      const double cbb = (BB_MODE != 0) ? __ldg(cb + b) : 0.0;
      const double m = band_step<GA_FREE, BB_MODE, RADIO>(cfg, w, a, b, __ldg(tGA + b), __ldg(tSB + b),
                                                         __ldg(tTO + b), __ldg(tBB + b), drad, __ldg(ys + b),
                                                         __ldg(is + b), cgb, cbb, omb, eb_free);
      if (ym) ym[b] = m;
    }
  }
  const double lnlike = -0.5 * a.chi;

  // ---- physical priors -------------------------------------------------------------------------
  double prior = 0.0;
  if (cfg.eb_mode != 0) {
    double L_att;
    if (GA_FREE) L_att = fabs((a.ebA - a.ebB) * w.nGA);
    else L_att = __ldg(P.fam[0].aux + row[0]) * w.nGA;
    const double L_sb = __ldg(P.fam[1].aux + row[1]) * w.nSB;
    double pr;
    if (L_sb < L_att) pr = -9999.0;
    else if (L_sb >= L_att) pr = (cfg.eb_mode == 1) ? 0.0 : gaussian_prior(P.lognorm_s01, 100.0, 0.0, log10(L_sb / L_att));
    else pr = NAN;   // unordered: the reference has no branch for it
    prior += pr;
  }
  if (cfg.prior_agnfrac) {
    double bbj = a.bb_j;
    if (BB == 0.0) bbj = bbj * inv4pidl2;          // decided by VALUE (PRIORS:177)
    const double frac = log10(bbj / a.ga_j);
    double pr;
    if (__ldg(P.agn_branch + s) == 0) pr = gaussian_prior(P.lognorm_s2, 0.25, -2.0, frac);
    else pr = (frac < 0.0) ? -9999.0 : gaussian_prior(P.lognorm_s2, 0.25, 2.0, frac);
    prior += pr + gaussian_prior(P.lognorm_s15, 1.0 / 2.25, 4.5, GA);
  }
  double tor6 = 0.0, lumf = 0.0;
  if (P.xp_midir_uv | P.xp_ir_xray) {
    tor6 = __ldg(tTO + P.xp_j6) * w.nTO;
    lumf = __ldg(P.consts + (int64_t)s * P.n_const + 1);
  }
  if (P.xp_midir_uv) {
    const int64_t dr = (int64_t)zi * P.fam[3].T + __ldg(P.xp_bb_dered + row[3]);
    const double bb2500 = __ldg(P.fam[3].table + dr * P.fam[3].pitch + P.xp_j2500) * w.nBB;
    prior += prior_midir_uv(P, tor6, bb2500, BB, lumf, __ldg(P.consts + (int64_t)s * P.n_const + 3));
  }
  if (RADIO) {
    const int mode = __ldg(P.irsyn_mode + s);
    if (mode != 0) {
      const double f = log10(a.syn_ir / a.sb_ir);
      prior += gaussian_prior(P.lognorm_s2, 0.25, mode > 0 ? 2.0 : -2.0, f);
    }
  }
  if (P.xp_ir_xray) prior += prior_ir_xrays(P, tor6, lumf, __ldg(P.consts + (int64_t)s * P.n_const + 2));
  const double lnprior = inbox ? prior : -INFINITY;
  const bool fin = (lnprior - lnprior) == 0.0;     // isfinite
  B.lnp[i] = fin ? lnprior + lnlike : -INFINITY;
  if (B.lnprior) B.lnprior[i] = lnprior;
  if (B.lnlike) B.lnlike[i] = lnlike;
}

template <bool GA_FREE, int BB_MODE, bool RADIO>
cudaError_t launch_variant(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  const int64_t blocks = (B.n + K1_THREADS - 1) / K1_THREADS;
  const size_t smem = (size_t)P.n_grid_vals * (sizeof(double) + sizeof(int32_t));
  auto kern = k1_lnprob_kernel<GA_FREE, BB_MODE, RADIO>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)blocks, K1_THREADS, smem, stream>>>(P, B);
  return cudaGetLastError();
}

template <bool GA_FREE, bool RADIO>
cudaError_t launch_bb(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  switch (P.cfg.bb_mode) {
    case 0: return launch_variant<GA_FREE, 0, RADIO>(P, B, stream);
    case 1: return launch_variant<GA_FREE, 1, RADIO>(P, B, stream);
    case 2: return launch_variant<GA_FREE, 2, RADIO>(P, B, stream);
    default: return launch_variant<GA_FREE, 3, RADIO>(P, B, stream);
  }
}

}  // namespace

cudaError_t launch_k1(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  if (B.n <= 0) return cudaSuccess;
  const bool gf = P.cfg.ga_free_ebv != 0, rad = P.cfg.radio != 0;
  if (gf) return rad ? launch_bb<true, true>(P, B, stream) : launch_bb<true, false>(P, B, stream);
  return rad ? launch_bb<false, true>(P, B, stream) : launch_bb<false, false>(P, B, stream);
}

}  // namespace agnf
