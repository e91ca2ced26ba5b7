// K2 — dictionary build (filter convolution of redshifted templates) for sm_100a.
//
// Reference: DICTIONARIES_AGNfitter.filtering_models (DICTIONARIES_AGNfitter.py:426-478) +
// unit helpers (483-503), looped over templates and z by construct_dictionaryarray_filtered
// (78-136).  Per (template, z, band) the reference
//   1. redshifts the template wavelengths  lam_j = c/10^nu_j * 1e10 * (1+z)       (450)
//   2. converts F_nu -> F_lambda = F_nu / lam^2 * c * 1e10                          (453, 492-497)
//   3. nearest-neighbour resamples onto the n_k filter wavelengths                 (454, 467)
//      scipy interp1d('nearest'): midpoints x[j]/2 + x[j+1]/2, searchsorted side='left',
//      clip, 0 outside [x_0, x_last]
//   4. trapz(F*T, lam_f) / trapz(T, lam_f)                                         (470-472)
//   5. converts back with the band's central wavelength: * lam_c^2 / c / 1e10      (475, 486-491)
//
// Step 3 makes every filter sample pick exactly one template sample, and all templates of a
// family share the wavelength grid, so steps 2-5 collapse, per (z, band), into a short weight
// vector W[z][b][r] over the template samples j_lo..j_lo+cnt-1 the band can see:
//     out[z][t][b] = sum_r W[z][b][r] * F_nu[t][j_lo[z][b] + r]
// i.e. the dense contraction (templates x lambda) . (lambda x bands) of the north star, stored
// banded.  Three kernels:
//   k2_ranges    j_lo / cnt per (z, band) and the per-band max cnt       (integer work, bit-exact)
//   k2_weights   W[z][woff[b] + r]: two binary searches + a short sequential sum per entry; band b owns
//                rmax_b = max_z cnt[z][b] slots per z-slice (compact: no padding to the widest band)
//   k2_contract  the banded contraction; writes the [n_z, T, pitch] table K1 gathers from.  The tiled
//                kernel keeps 32 templates in shared memory across a chunk of z-slices; the per-slice
//                kernel is the fallback when a slice of weights does not fit (X-ray power-law tails)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace agnf {

struct K2Filters {
  int32_t n_b = 0;
  int64_t n_samples = 0;
  const double* lam = nullptr;      // [n_samples] per band ascending
  const double* wn = nullptr;       // [n_samples] trapezoid weight * throughput, same order
  const int64_t* off = nullptr;     // [n_b + 1]
  const double* scale = nullptr;    // [n_b] lam_c^2 / c / 1e10 / trapz(T, lam_f)
};

struct K2Problem {
  K2Filters filt;
  const double* lam0 = nullptr;     // [n_lam] rest-frame template wavelengths, ascending
  const double* Fs = nullptr;       // [T, ldF] template F_nu in lam0 order
  int64_t T = 0;
  int32_t n_lam = 0;
  int32_t ldF = 0;
  const double* z = nullptr;        // [n_z]
  int32_t n_z = 0;
  int32_t wstride = 0;              // weights per z-slice (even): sum of the per-band slot counts
  int32_t* jlo = nullptr;           // [n_z, n_b]
  int32_t* cnt = nullptr;           // [n_z, n_b]
  int32_t* d_rmaxb = nullptr;       // [n_b] max over z of cnt (written by k2_ranges)
  const int32_t* woff = nullptr;    // [n_b + 1] first slot of each band inside a z-slice of W
  const int32_t* band_of = nullptr; // [wstride] band that owns a slot (-1: padding)
  const int32_t* band_order = nullptr; // [n_b] bands by descending slot count (work balancing in the tiled kernel)
  double* W = nullptr;              // [n_z, wstride]
  int32_t tile_ok = 0;              // Fs is zero-padded to a multiple of 32 rows: the tiled contraction may be used
  const int64_t* row_map = nullptr; // [T] or nullptr
  int64_t T_total = 0;
  int32_t out_pitch = 0;            // row pitch of out in doubles (>= n_b); pad columns are written as 0
  double* out = nullptr;            // [n_z, T_total, out_pitch]
};

cudaError_t launch_k2_permute(const double* Fnu, const int32_t* perm, double* Fs, int64_t T, int32_t n_lam,
                              int32_t ldF, cudaStream_t stream);
cudaError_t launch_k2_ranges(const K2Problem& p, cudaStream_t stream);
cudaError_t launch_k2_weights(const K2Problem& p, cudaStream_t stream);
cudaError_t launch_k2_contract(const K2Problem& p, cudaStream_t stream);

}  // namespace agnf
