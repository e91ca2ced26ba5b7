// Device helpers shared by the K1 kernels (k1_lnprob.cu, k1_coop.cu).
#pragma once
#include <climits>
#include <cmath>
#include <cstdint>

#include "k1_lnprob.cuh"   // AGNF_POW2_DENOM

namespace agnf {

// shared memory through 32-bit shared-window addresses (no generic -> shared conversions)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
  return v;
}
// Read of shared data that is written once before the kernel's first barrier and never again (tables,
// grids): not volatile, so the compiler may schedule it freely; its address always depends on values
// obtained after that barrier, which keeps it behind it.
__device__ __forceinline__ double lds_f64_ro(uint32_t a) {
  double v;
  asm("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
  return v;
}

// 2^(j/64), j = 0..63 (the cooperative kernel stages every (64/D)-th entry, see AGNF_POW2_DENOM)
__constant__ double c_exp2_tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, 1.0442737824274138, 1.0556451783605572,
    1.0671404006768237, 1.0787607977571199, 1.0905077326652577, 1.102382583307841, 1.1143867425958924,
    1.1265216186082418, 1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687, 1.241857812073484,
    1.255380757024691, 1.2690509571917332, 1.2828700160787783, 1.2968395546510096, 1.3109612115247644,
    1.3252366431597413, 1.339667524053303, 1.3542555469368927, 1.3690024229745905, 1.383909881963832,
    1.3989796725383112, 1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384, 1.5422108254079407,
    1.559004400237837, 1.5759808451078865, 1.593142151342267, 1.6104903319492543, 1.6280274218573478,
    1.645755478153965, 1.6636765803267364, 1.681792830507429, 1.7001063537185235, 1.718619298122478,
    1.7373338352737062, 1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656, 1.9152065613971474,
    1.9360617934922943, 1.9571441241754002, 1.978456026387951};

// ---- 2^(t/D): the factors 10^(c E) with c pre-scaled by D log2(10) on the host (D = AGNF_POW2_DENOM) ----
// t = k + r, k integer, |r| <= 1/2: 2^(t/D) = 2^(k/D) * exp(r ln2/D); 2^(k/D) = 2^(k div D) * tab[k mod D].
// Polynomial coefficients (ln 2 / D)^n / n! sit in constant memory so the DFMAs take them as operands;
// the first neglected term is < 4e-17 (D = 64, degree 5) / 2e-18 (D = 16, degree 7).
#if AGNF_POW2_DENOM == 64
constexpr int P2_DEG = 5, P2_SHIFT = 6;
__constant__ double c_p2[5] = {0.010830424696249145, 5.86490495505617e-05, 2.1173137155464776e-07,
                               5.732851688640402e-10, 1.2417843701716925e-12};
#elif AGNF_POW2_DENOM == 16
constexpr int P2_DEG = 7, P2_SHIFT = 4;
__constant__ double c_p2[7] = {0.04332169878499658, 0.0009383847928089872, 1.3550807779497457e-05,
                               1.467610032291943e-07, 1.2715871950558131e-09, 9.181219573844438e-12,
                               5.682086126528621e-14};
#else
#error "AGNF_POW2_DENOM must be 16 or 64"
#endif
constexpr double P2_LOG2_10_D = (double)AGNF_POW2_DENOM * 3.321928094887362;   // D log2(10)

// |t| >= 1000 D (or NaN) goes to the library
static __device__ __noinline__ double pow2_64_slow(double t) { return exp2(t * (1.0 / AGNF_POW2_DENOM)); }

__device__ __forceinline__ double pow2_64(double t, uint32_t tab_s) {
  if (!(fabs(t) < 1000.0 * AGNF_POW2_DENOM)) return pow2_64_slow(t);
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52
  const double km = t + magic;
  const int k = __double2loint(km);
  const double r = t - (km - magic);
  double p = c_p2[P2_DEG - 1];
#pragma unroll
  for (int n = P2_DEG - 2; n >= 0; --n) p = fma(p, r, c_p2[n]);
  p = fma(p, r, 1.0);
  const double v = lds_f64_ro(tab_s + ((k << 3) & (8 * AGNF_POW2_DENOM - 8))) * p;
  return __hiloint2double(__double2hiint(v) + ((k >> P2_SHIFT) << 20), __double2loint(v));
}

// ---- pinned forms -------------------------------------------------------------------------------------------------
// The cooperative kernel's arithmetic leaves the multiply-add contraction of a few expressions to the compiler; what
// it compiles to (read off its SASS) is spelled out here for the kernels that must reproduce it bit for bit in other
// instantiations (persistent sampler, gather4 kernel), where the compiler's own choice can differ by an ulp:
//   a*b + c*d [+ e*f + g*h]   ->  fma(.., fma(e, f, fma(a, b, c*d)))      (the SECOND product is the plain multiply)
//   2^(a b / D), band loop    ->  t = a*b rounded, integer part from t + magic (not fused)
//   10^x, normalisations      ->  t = x C rounded for the remainder, integer part from fma(x, C, magic)
__device__ __forceinline__ double sum2_pinned(double a, double b, double c, double d) {
  return __fma_rn(a, b, __dmul_rn(c, d));
}
__device__ __forceinline__ double sum4_pinned(double a, double b, double c, double d, double e, double f, double g,
                                              double h) {
  return __fma_rn(g, h, __fma_rn(e, f, __fma_rn(a, b, __dmul_rn(c, d))));
}
template <bool NB>
__device__ __forceinline__ double pow10_pinned(double x, uint32_t tab_s, bool& bad) {
  const double t = __dmul_rn(x, P2_LOG2_10_D);
  if (NB) bad = bad || !(fabs(t) < 1000.0 * AGNF_POW2_DENOM);
  else if (!(fabs(t) < 1000.0 * AGNF_POW2_DENOM)) return pow2_64_slow(t);
  const double magic = 6755399441055744.0;
  const double km = __fma_rn(x, P2_LOG2_10_D, magic);
  const int k = __double2loint(km);
  const double r = t - (km - magic);
  double p = c_p2[P2_DEG - 1];
#pragma unroll
  for (int n = P2_DEG - 2; n >= 0; --n) p = fma(p, r, c_p2[n]);
  p = fma(p, r, 1.0);
  const double v = lds_f64_ro(tab_s + ((k << 3) & (8 * AGNF_POW2_DENOM - 8))) * p;
  return __hiloint2double(__double2hiint(v) + ((k >> P2_SHIFT) << 20), __double2loint(v));
}

__device__ __forceinline__ double pow10_pinned_tile(double x, uint32_t tab_s) {
  bool unused = false;
  return pow10_pinned<false>(x, tab_s, unused);
}

// The same arithmetic without the range branch: out-of-range (or NaN) arguments raise `bad` and return garbage; the
// caller repeats the evaluation through pow2_64 (straight-line band loops: k1_g4.cu)
__device__ __forceinline__ double pow2_64_nb(double t, uint32_t tab_s, bool& bad) {
  bad = bad || !(fabs(t) < 1000.0 * AGNF_POW2_DENOM);
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52
  const double km = t + magic;
  const int k = __double2loint(km);
  const double r = t - (km - magic);
  double p = c_p2[P2_DEG - 1];
#pragma unroll
  for (int n = P2_DEG - 2; n >= 0; --n) p = fma(p, r, c_p2[n]);
  p = fma(p, r, 1.0);
  const double v = lds_f64_ro(tab_s + ((k << 3) & (8 * AGNF_POW2_DENOM - 8))) * p;
  return __hiloint2double(__double2hiint(v) + ((k >> P2_SHIFT) << 20), __double2loint(v));
}

// 10^x for the normalisations through the same routine (|x| <= 10 in the prior box: relative error
// <= |x| D log2(10) 2^-53 ln2 / D ~ 2.6e-15)
__device__ __forceinline__ double pow10_via2(double x, uint32_t tab_s) { return pow2_64(x * P2_LOG2_10_D, tab_s); }

// Given l = last index with sv[l] <= p (or -1), return the first-appearance index of the minimum
// of |g - p| over the distinct grid values of one dimension == np.abs(keys_float - p).argmin()
// (DICTIONARIES_AGNfitter.py:204) mapped to the distinct-value list:
//   * the minimisers of the rounded |g - p| form a contiguous run around p in sorted order
//     (fp subtraction is monotone), so the two neighbours + a tie walk find them all;
//   * among minimisers numpy returns the earliest key, i.e. the smallest first-appearance index.
__device__ __forceinline__ int snap_resolve(const double* __restrict__ sv, const int32_t* __restrict__ perm,
                                            int n, double p, int l) {
  const int r = l + 1;
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

// Binary-search snap (all-NaN -> index 0, like np.argmin of an all-NaN array).
__device__ __forceinline__ int snap_dim(const double* __restrict__ sv, const int32_t* __restrict__ perm,
                                        int n, double p) {
  if (p != p) return 0;
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (sv[mid] <= p) lo = mid + 1; else hi = mid;
  }
  return snap_resolve(sv, perm, n, p, lo - 1);
}

// Same result through a cell table.  `ext` / `eperm` are the sorted values and their first-appearance
// indices with a sentinel on either side: ext[0] = -inf, ext[1..n] = sv[0..n-1], ext[n+1] = +inf.
// lut[c] is a guess for l (ext indexing) at the left edge of cell c of a uniform partition of
// [sv[0], sv[n-1]]; three predicated steps correct it and an explicit bracket test
// ext[l] <= p < ext[l+1] decides whether the guess path may answer.  It may only when the minimum can
// tie between the two bracketing values and nowhere else: the host sets `big` so that for
// |p| <= big the spacing of the grid exceeds the rounding granularity of |g - p| (agnfit_capi.cu).
// Everything else (NaN, +-inf, huge |p|, a failed bracket) takes the binary search.
__device__ __forceinline__ int snap_dim_lut(const double* __restrict__ ext, const int32_t* __restrict__ eperm,
                                            int n, double p, const uint16_t* __restrict__ lut, int m,
                                            double lo, double inv, double big) {
  if (fabs(p) <= big) {
    const double t = fmin(fmax((p - lo) * inv, 0.0), (double)(m - 1));
    int l = lut[__double2int_rz(t)];
    if (ext[l] > p) --l;
    if (ext[l + 1] <= p) ++l;
    if (ext[l + 1] <= p) ++l;
    const double a = ext[l], b = ext[l + 1];
    if (a <= p && p < b) {
      const double da = fabs(a - p), db = fabs(b - p);
      const int ia = eperm[l], ib = eperm[l + 1];
      return da < db ? ia : (db < da ? ib : min(ia, ib));
    }
  }
  return snap_dim(ext + 1, eperm + 1, n, p);
}

// Gaussian_prior(mu, sigma, x) = log(1/(sqrt(2 pi) sigma)) - 0.5 (x-mu)^2 / sigma^2 (PRIORS_AGNfitter.py:428-430);
// lognorm is evaluated once on the host (libm), inv_var = 1/sigma^2.
__device__ __forceinline__ double gaussian_prior(double lognorm, double inv_var, double mu, double x) {
  // every operation rounded on its own, no fused multiply-add: all kernels (and NumPy) round alike
  const double d = __dsub_rn(x, mu);
  return __dsub_rn(lognorm, __dmul_rn(__dmul_rn(0.5, __dmul_rn(d, d)), inv_var));
}

// prior_midIR_UV (PRIORS_AGNfitter.py:327-357).  tor6 = torus band flux at 6 um times 10^TO; bb2500 = band flux of
// the unreddened disc template at 2500 A times 10^BB; lumf = 4. pi dlum^2. (:331); lumf_bb = 4 pi (dlum)^2 (:346)
__device__ __forceinline__ double prior_midir_uv(const K1Params& P, double tor6, double bb2500, double BB, double lumf,
                                                 double lumf_bb) {
  const double x = __dsub_rn(log10(__dmul_rn(tor6, lumf)), 27.30103);
  const double to_model = __dsub_rn(__dadd_rn(16.2530786, __dmul_rn(1.024, x)), __dmul_rn(0.047, __dmul_rn(x, x))) / 0.643;
  if (BB != 0.0) bb2500 = bb2500 * lumf_bb;                   // decided by VALUE (:345)
  return gaussian_prior(P.lognorm_s06, 1.0 / (0.6 * 0.6), 0.0, __dsub_rn(log10(bb2500), to_model));
}

// prior_IR_XRays (PRIORS_AGNfitter.py:302-324).  logf_data = log10(F_2-10keV,data * lumfactor) (:317-318, host)
__device__ __forceinline__ double prior_ir_xrays(const K1Params& P, double tor6, double lumf, double logf_data) {
  const double nuLnu = (P.xp_c6um * tor6) * lumf;
  const double x = log10(nuLnu / 1e41);
  const double logf_model = __dsub_rn(__dadd_rn(22.9494264, __dmul_rn(1.024, x)), __dmul_rn(0.047, __dmul_rn(x, x)));
  return gaussian_prior(P.lognorm_s05, 1.0 / (0.5 * 0.5), 0.0, __dsub_rn(logf_data, logf_model));
}

}  // namespace agnf
