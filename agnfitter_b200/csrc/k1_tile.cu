// K1, tiled variant: thread-per-row arithmetic, warp-cooperative memory.
//
// The first K1 (k1_lnprob.cu) lets every thread gather its own 8*n_b-byte template rows with
// 128-bit loads.  ncu on the C3 workload showed that version bound by the L1/TEX wavefront rate
// and by exposed load latency (profiles/k1_r01_summary.md): a warp-wide LDG.128 whose 32 lanes
// point into 32 different rows costs 32 tag lookups, and each thread needs 9 of them per family.
//
// Here a warp moves the rows of its 32 evaluations cooperatively instead: the 32*NB/2 16-byte
// chunks of one family are spread over the lanes so that consecutive lanes copy consecutive
// chunks of the same row (cp.async.cg 16 B, global -> shared, no registers, L1 bypassed), i.e.
// every request covers whole rows.  Each thread then reads its own row back from shared memory
// (row pitch 8*NB bytes: the four 8-lane phases of an LDS.128 hit disjoint banks for NB = 18).
// Three per-warp buffers keep three families in flight: SB and BB are consumed while the galaxy
// rows (the only ones that normally come from DRAM) are still arriving.  The params rows of a
// warp are contiguous in global memory and are staged the same way.
//
// The arithmetic, its order, and every reference citation are those of k1_lnprob.cu.
#include "k1_lnprob.cuh"

#include <climits>
#include <cmath>

namespace agnf {

namespace {

constexpr int KT_THREADS = 128;
constexpr int KT_WARPS = KT_THREADS / 32;

__constant__ double c_exp2_tab[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484,
    1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,
    1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429, 1.718619298122478,
    1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,
    1.9571441241754002};

// 10^x for |x| < 300 to ~3e-16 relative: x = k*log10(2)/32 + r, |r| <= log10(2)/64;
// 10^x = 2^(k>>5) * 2^((k&31)/32) * exp(r ln 10), the last factor by a degree-6 polynomial
// (next term < 4e-18).  Anything else (NaN, inf, huge) goes to the library exp10.
__device__ __forceinline__ double exp10_fast(double x, const double* __restrict__ tab) {
  if (!(fabs(x) < 300.0)) return exp10(x);
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52: round-to-nearest-int trick
  const double t = fma(x, 106.30169903639559, magic);      // 32 * log2(10)
  const int k = __double2loint(t);
  const double kf = t - magic;
  double r = fma(kf, -0.009407187360920943, x);            // Cody-Waite: log10(2)/32 = C1 + C2
  r = fma(kf, -3.5784690306318245e-12, r);
  double p = 0.2069958486968681;
  p = fma(p, r, 0.5393829291955814);
  p = fma(p, r, 1.171255148912267);
  p = fma(p, r, 2.034678592293476);
  p = fma(p, r, 2.650949055239199);
  p = fma(p, r, 2.302585092994046);
  p = fma(p, r, 1.0);
  const double v = tab[k & 31] * p;
  return __hiloint2double(__double2hiint(v) + ((k >> 5) << 20), __double2loint(v));
}

__device__ __forceinline__ int snap_dim(const double* __restrict__ sv, const int32_t* __restrict__ perm,
                                        int n, double p) {
  // identical to k1_lnprob.cu::snap_dim (first argmin of |g - p|, DICTIONARIES_AGNfitter.py:204)
  if (p != p) return 0;
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (sv[mid] <= p) lo = mid + 1; else hi = mid;
  }
  const int l = lo - 1, r = lo;
  double best;
  if (l < 0) best = fabs(sv[r] - p);
  else if (r >= n) best = fabs(sv[l] - p);
  else best = fmin(fabs(sv[l] - p), fabs(sv[r] - p));
  int idx = INT_MAX;
  for (int k = l; k >= 0; --k) {
    if (fabs(sv[k] - p) == best) idx = min(idx, perm[k]); else break;
  }
  for (int k = r; k < n; ++k) {
    if (fabs(sv[k] - p) == best) idx = min(idx, perm[k]); else break;
  }
  return idx;
}

__device__ __forceinline__ double gaussian_prior(double mu, double sigma, double x) {
  const double lognorm = log(1.0 / (sqrt(2.0 * 3.141592653589793) * sigma));
  const double d = x - mu;
  return lognorm - 0.5 * (d * d) / (sigma * sigma);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Warp-cooperative gather of one family's rows: evaluation e (lane e) owns element offset
// `rowoff` into `table`; chunk c = e * NB/2 + k (16 bytes) lands at buf + 2c doubles.
template <int NB>
__device__ __forceinline__ void stage_rows(double* __restrict__ buf, const double* __restrict__ table,
                                           long long rowoff, int lane) {
  constexpr int CPR = NB / 2;                 // 16-byte chunks per row
#pragma unroll
  for (int it = 0; it < CPR; ++it) {          // 32 * CPR chunks over 32 lanes
    const int c = it * 32 + lane;
    const int e = c / CPR;
    const int k = c - e * CPR;
    const long long off = __shfl_sync(0xffffffffu, rowoff, e);
    cp_async16(buf + 2 * c, table + off + 2 * k);
  }
  cp_async_commit();
}

template <int NB, bool GA_FREE, int BB_MODE, bool RADIO>
__global__ void __launch_bounds__(KT_THREADS, 4)
k1_tile_kernel(const __grid_constant__ K1Params P, const K1Batch B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = P.n_grid_vals;
  double* s_tab = reinterpret_cast<double*>(smem_raw);                  // 32
  double* s_grid = s_tab + 32;                                          // G (+ pad to 16 B)
  const int grid_pad = (G + 1) & ~1;
  int32_t* s_perm = reinterpret_cast<int32_t*>(s_grid + grid_pad);      // G (+ pad to 16 B)
  const int perm_pad = (G + 3) & ~3;
  double* s_bufs = reinterpret_cast<double*>(s_perm + perm_pad);        // KT_WARPS * 3 * 32 * NB
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    s_grid[i] = P.grid_sorted[i];
    s_perm[i] = P.grid_perm[i];
  }
  if (threadIdx.x < 32) s_tab[threadIdx.x] = c_exp2_tab[threadIdx.x];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* bufA = s_bufs + (size_t)warp * 3 * 32 * NB;
  double* bufB = bufA + 32 * NB;
  double* bufG = bufB + 32 * NB;

  const agnf_config& cfg = P.cfg;
  const int n_par = cfg.n_par;
  const int64_t row0 = (int64_t)blockIdx.x * KT_THREADS + warp * 32;    // first row of this warp
  if (row0 >= B.n) return;                                              // whole warp past the end
  const int64_t i = row0 + lane;
  const bool live = i < B.n;
  const int64_t ii = live ? i : B.n - 1;                                // tail lanes shadow the last row

  // ---- params of the warp: one contiguous block, staged through bufA/bufB --------------------
  double* s_par = bufA;                                                 // [32][n_par], n_par <= 32 fits A+B
  {
    const int64_t rows_here = min((int64_t)32, B.n - row0);
    const int n_chunks = (int)(rows_here * n_par) / 2;                  // 16-byte chunks (n_par*rows even?)
    const double* src = B.params + row0 * n_par;
    // row0 * n_par * 8 bytes is a multiple of 256: the block is 16-byte aligned
    for (int c = lane; c < n_chunks; c += 32) cp_async16(s_par + 2 * c, src + 2 * c);
    if (((rows_here * n_par) & 1) && lane == 0) s_par[rows_here * n_par - 1] = __ldg(src + rows_here * n_par - 1);
    cp_async_commit();
  }
  const int s = __ldg(B.source_id + ii);
  const int zi = __ldg(P.zi + s);
  cp_async_wait<0>();
  __syncwarp();
  const double* __restrict__ p = s_par + (live ? lane : (int)(B.n - 1 - row0)) * n_par;

  // ---- box prior ---------------------------------------------------------------------------
  bool inbox = true;
  {
    const double* __restrict__ pmin = P.pmin + (int64_t)s * n_par;
    const double* __restrict__ pmax = P.pmax + (int64_t)s * n_par;
    for (int c = 0; c < n_par; ++c) {
      const double v = p[c];
      inbox = inbox && (__ldg(pmin + c) <= v) && (v <= __ldg(pmax + c));
    }
  }

  // ---- snap ----------------------------------------------------------------------------------
  long long rowoff[AGNF_MAX_FAMILIES];
  int row[AGNF_MAX_FAMILIES];
#pragma unroll
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
    row[f] = 0;
    rowoff[f] = 0;
    if (f < cfg.n_fam) {
      const FamDesc& F = P.fam[f];
      int64_t r = 0;
      for (int d = 0; d < F.n_dim; ++d) {
        if (F.dim[d].is_grid) {
          const int off = F.dim[d].smem_off;
          const int idx = snap_dim(s_grid + off, s_perm + off, F.dim[d].size, p[F.par_offset + d]);
          r += (int64_t)idx * F.dim[d].stride;
        }
      }
      row[f] = (int)r;
      rowoff[f] = ((long long)zi * F.T + r) * NB;
    }
  }

  // ---- normalisations and free parameters (read before the params buffer is recycled) ---------
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
  __syncwarp();                                   // everyone is done with s_par

  // ---- start the gathers: consumption order SB, BB, GA, TO [, RAD] ------------------------------
  stage_rows<NB>(bufA, P.fam[1].table, rowoff[1], lane);      // group 0: SB
  stage_rows<NB>(bufB, P.fam[3].table, rowoff[3], lane);      // group 1: BB
  stage_rows<NB>(bufG, P.fam[0].table, rowoff[0], lane);      // group 2: GA

  if (B.rows && live) {
    for (int f = 0; f < cfg.n_fam; ++f) B.rows[i * cfg.n_fam + f] = row[f];
  }
  const double inv4pidl2 = __ldg(P.consts + (int64_t)s * P.n_const + 0);
  const double nGA = exp10_fast(GA, s_tab), nSB = exp10_fast(SB, s_tab), nTO = exp10_fast(TO, s_tab);
  const double nBB = cfg.bbb_has_norm ? exp10_fast(BB, s_tab) : 1.0;
  const double nRAD = RADIO ? exp10_fast(RAD, s_tab) : 0.0;
  const double cBB = cfg.bbb_has_norm ? 1.0 : inv4pidl2;
  const double xfac = (BB_MODE >= 2) ? exp10_fast(alpha / 0.3838, s_tab) : 1.0;
  const uint64_t valid = __ldg(P.valid + s);
  const uint64_t xray = __ldg(P.xray + zi);
  const int j1500 = __ldg(P.j1500 + zi);
  const int jir = RADIO ? __ldg(P.irsyn_j + s) : -1;
  const bool eb_free = GA_FREE && cfg.eb_mode != 0;

  double m[NB];
  double ga_j = 0.0, bb_j = 0.0, sb_ir = 0.0, syn_ir = 0.0, ebA = 0.0, ebB = 0.0;

  // SB
  cp_async_wait<2>();
  __syncwarp();
  {
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(bufA + lane * NB);
#pragma unroll
    for (int h = 0; h < NB / 2; ++h) {
      const double2 v = r2[h];
      m[2 * h] = nSB * v.x;
      m[2 * h + 1] = nSB * v.y;
      if (RADIO) {
        if (2 * h == jir) sb_ir = v.x * nSB;
        if (2 * h + 1 == jir) sb_ir = v.y * nSB;
      }
    }
  }
  __syncwarp();
  stage_rows<NB>(bufA, P.fam[2].table, rowoff[2], lane);      // group 3: TO (bufA is free again)

  // BB
  cp_async_wait<2>();
  __syncwarp();
  {
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(bufB + lane * NB);
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(P.cbbb + (int64_t)zi * NB);
#pragma unroll
    for (int h = 0; h < NB / 2; ++h) {
      const double2 v = r2[h];
      double e[2] = {v.x, v.y};
      if (BB_MODE != 0) {
        const double2 cb = __ldg(c2 + h);
        const double cbv[2] = {cb.x, cb.y};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int b = 2 * h + q;
          if (BB_MODE >= 2 && ((xray >> b) & 1ull)) {
            e[q] = e[q] * xfac;
          } else {
            const double w = e[q] * exp10_fast(cbv[q] * ebv_bbb, s_tab);
            e[q] = (w != w) ? e[q] : w;                       // NaN -> unreddened (MODEL:1211)
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int b = 2 * h + q;
        m[b] = m[b] + nBB * (e[q] * cBB);
        if (b == j1500) bb_j = e[q] * nBB;
      }
    }
  }
  if (RADIO) {
    __syncwarp();
    stage_rows<NB>(bufB, P.fam[4].table, rowoff[4], lane);    // group 4: RAD
  }

  // GA
  if (RADIO) cp_async_wait<2>(); else cp_async_wait<1>();
  __syncwarp();
  {
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(bufG + lane * NB);
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(P.cgal + (int64_t)zi * NB);
    const double2* __restrict__ o2 = reinterpret_cast<const double2*>(P.omega + (int64_t)zi * NB);
#pragma unroll
    for (int h = 0; h < NB / 2; ++h) {
      const double2 v = r2[h];
      const double d[2] = {v.x, v.y};
      double cgv[2] = {0.0, 0.0}, omv[2] = {0.0, 0.0};
      if (GA_FREE) {
        const double2 cg = __ldg(c2 + h);
        const double2 om = __ldg(o2 + h);
        cgv[0] = cg.x; cgv[1] = cg.y; omv[0] = om.x; omv[1] = om.y;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int b = 2 * h + q;
        double rga = 1.0;
        if (GA_FREE) rga = exp10_fast(cgv[q] * ebv_gal, s_tab);
        const double ga_r = GA_FREE ? d[q] * rga : d[q];
        m[b] = m[b] + nGA * ga_r;
        if (b == j1500) ga_j = ga_r * nGA;
        if (GA_FREE && eb_free) {
          ebA += omv[q] * (d[q] * 3.826e33);
          ebB += omv[q] * (((d[q] * 1e18) * rga) * 3.826e33);
        }
      }
    }
  }

  // TO
  if (RADIO) cp_async_wait<1>(); else cp_async_wait<0>();
  __syncwarp();
  {
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(bufA + lane * NB);
#pragma unroll
    for (int h = 0; h < NB / 2; ++h) {
      const double2 v = r2[h];
      m[2 * h] = m[2 * h] + nTO * v.x;
      m[2 * h + 1] = m[2 * h + 1] + nTO * v.y;
    }
  }
  if (RADIO) {
    cp_async_wait<0>();
    __syncwarp();
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(bufB + lane * NB);
#pragma unroll
    for (int h = 0; h < NB / 2; ++h) {
      const double2 v = r2[h];
      m[2 * h] += nRAD * v.x;
      m[2 * h + 1] += nRAD * v.y;
      if (2 * h == jir) syn_ir = v.x * nRAD;
      if (2 * h + 1 == jir) syn_ir = v.y * nRAD;
    }
  }

  // ---- chi^2 -----------------------------------------------------------------------------------
  double chi = 0.0;
  {
    const double2* __restrict__ y2 = reinterpret_cast<const double2*>(P.y + (int64_t)s * NB);
    const double2* __restrict__ i2 = reinterpret_cast<const double2*>(P.isig + (int64_t)s * NB);
#pragma unroll
    for (int h = 0; h < NB / 2; ++h) {
      const double2 yb = __ldg(y2 + h);
      const double2 ib = __ldg(i2 + h);
      if ((valid >> (2 * h)) & 1ull) {
        const double r = (yb.x - m[2 * h]) * ib.x;
        chi += r * r;
      }
      if ((valid >> (2 * h + 1)) & 1ull) {
        const double r = (yb.y - m[2 * h + 1]) * ib.y;
        chi += r * r;
      }
    }
  }
  if (!live) return;
  if (B.ymodel) {
    double* ym = B.ymodel + i * NB;
#pragma unroll
    for (int b = 0; b < NB; ++b) ym[b] = m[b];
  }
  const double lnlike = -0.5 * chi;

  // ---- physical priors (as in k1_lnprob.cu) -----------------------------------------------------
  double prior = 0.0;
  if (cfg.eb_mode != 0) {
    double L_att;
    if (GA_FREE) L_att = fabs((ebA - ebB) * nGA);
    else L_att = __ldg(P.fam[0].aux + row[0]) * nGA;
    const double L_sb = __ldg(P.fam[1].aux + row[1]) * nSB;
    double pr;
    if (L_sb < L_att) pr = -9999.0;
    else if (L_sb >= L_att) pr = (cfg.eb_mode == 1) ? 0.0 : gaussian_prior(0.0, 0.1, log10(L_sb / L_att));
    else pr = NAN;
    prior += pr;
  }
  if (cfg.prior_agnfrac) {
    double bbj = bb_j;
    if (BB == 0.0) bbj = bbj * inv4pidl2;
    const double frac = log10(bbj / ga_j);
    double pr;
    if (__ldg(P.agn_branch + s) == 0) pr = gaussian_prior(-2.0, 2.0, frac);
    else pr = (frac < 0.0) ? -9999.0 : gaussian_prior(2.0, 2.0, frac);
    prior += pr + gaussian_prior(4.5, 1.5, GA);
  }
  if (RADIO) {
    const int mode = __ldg(P.irsyn_mode + s);
    if (mode != 0) {
      const double f = log10(syn_ir / sb_ir);
      prior += gaussian_prior(mode > 0 ? 2.0 : -2.0, 2.0, f);
    }
  }
  const double lnprior = inbox ? prior : -INFINITY;
  const bool fin = (lnprior - lnprior) == 0.0;
  B.lnp[i] = fin ? lnprior + lnlike : -INFINITY;
  if (B.lnprior) B.lnprior[i] = lnprior;
  if (B.lnlike) B.lnlike[i] = lnlike;
}

template <int NB, bool GA_FREE, int BB_MODE, bool RADIO>
cudaError_t launch_tile(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  auto kern = k1_tile_kernel<NB, GA_FREE, BB_MODE, RADIO>;
  const int G = P.n_grid_vals;
  const size_t smem = 32 * 8 + (size_t)((G + 1) & ~1) * 8 + (size_t)((G + 3) & ~3) * 4 +
                      (size_t)KT_WARPS * 3 * 32 * NB * 8;
  static bool attr_done = false;   // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const int64_t blocks = (B.n + KT_THREADS - 1) / KT_THREADS;
  kern<<<(unsigned)blocks, KT_THREADS, smem, stream>>>(P, B);
  return cudaGetLastError();
}

template <int NB, bool GA_FREE, bool RADIO>
cudaError_t launch_tile_bb(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  switch (P.cfg.bb_mode) {
    case 0: return launch_tile<NB, GA_FREE, 0, RADIO>(P, B, stream);
    case 1: return launch_tile<NB, GA_FREE, 1, RADIO>(P, B, stream);
    case 2: return launch_tile<NB, GA_FREE, 2, RADIO>(P, B, stream);
    default: return launch_tile<NB, GA_FREE, 3, RADIO>(P, B, stream);
  }
}

template <int NB>
cudaError_t launch_tile_nb(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  const bool gf = P.cfg.ga_free_ebv != 0, rad = P.cfg.radio != 0;
  if (gf) return rad ? launch_tile_bb<NB, true, true>(P, B, stream) : launch_tile_bb<NB, true, false>(P, B, stream);
  return rad ? launch_tile_bb<NB, false, true>(P, B, stream) : launch_tile_bb<NB, false, false>(P, B, stream);
}

}  // namespace

bool k1_tile_supported(const K1Params& P) {
  // table offsets of a row must stay 16-byte aligned (NB even) and the staged grids must leave
  // room for the per-warp buffers; band counts are compiled in
  return P.cfg.n_b == 18 && P.n_grid_vals <= 2048;
}

cudaError_t launch_k1_tile(const K1Params& P, const K1Batch& B, cudaStream_t stream) {
  if (B.n <= 0) return cudaSuccess;
  switch (P.cfg.n_b) {
    case 18: return launch_tile_nb<18>(P, B, stream);
    default: return cudaErrorInvalidValue;
  }
}

// 10^x through the kernel's own exp10_fast (debug / unit test hook)
__global__ void exp10_fast_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n) {
  __shared__ double tab[32];
  if (threadIdx.x < 32) tab[threadIdx.x] = c_exp2_tab[threadIdx.x];
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
