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
//   k2_contract  the banded contraction; writes the [n_z, T, pitch] table K1 gathers from.  Four kernels:
//                tensor-core (FP64 mma.m8n8k4: 32 templates x 8 consecutive z-slices per band, the k-blocks of 4
//                samples aligned to absolute sample indices so that a slice's value does not depend on its batch),
//                shared-window (32 templates resident in shared memory; 4 consecutive z-slices share every
//                read of the template window, their weights re-aligned to the batch's common first sample),
//                tiled (same tile, one slice at a time: used when a template holds a non-finite sample,
//                which must not leak into a neighbouring slice), per-slice (fallback when a slice of
//                weights does not fit shared memory: X-ray power-law tails)
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
  // shared-window path: 4 consecutive z-slices (a batch) are evaluated from ONE read of the template window
  int32_t sw_ok = 0;                // every template sample is finite and the plan below was built
  int32_t n_zb = 0;                 // batches = ceil(n_z / 4)
  int32_t wstrideA = 0;             // slots per batch (each slot holds the 4 slices' weights)
  int32_t* jb = nullptr;            // [n_zb, n_b] first template sample any slice of the batch sees in the band
  int32_t* cb = nullptr;            // [n_zb, n_b] samples from jb to the last one any slice sees
  int32_t* d_rmaxbA = nullptr;      // [n_b] max over batches of cb
  const int32_t* woffA = nullptr;   // [n_b + 1]
  const int32_t* band_ofA = nullptr; // [wstrideA]
  int32_t hdrA = 0;                 // doubles of header in front of every batch's weights: jb[n_b] | cb[n_b] (int32)
  int64_t blkA = 0;                 // doubles per batch block = hdrA + wstrideA * 4
  double* WA = nullptr;             // [n_zb, blkA]
  // tensor-core path (FP64 mma.m8n8k4): batches of 8 slices, windows aligned to absolute multiples of 4 samples
  int32_t mma_ok = 0;
  int32_t n_zb8 = 0;                // batches = ceil(n_z / 8)
  int32_t wstride8 = 0;             // slots per batch (each slot holds the 8 slices' weights of one sample)
  int32_t ldF8 = 0;                 // row pitch of Fs8: = 4 (mod 8), >= n_lam + 3
  const double* Fs8 = nullptr;      // [T_pad, ldF8] templates in sorted-wavelength order, zero padded
  int32_t* jb8 = nullptr;           // [n_zb8, n_b] first sample of the batch window, rounded down to a multiple of 4
  int32_t* cb8 = nullptr;           // [n_zb8, n_b] window length, a multiple of 4
  int32_t* d_rmaxb8 = nullptr;      // [n_b]
  const int32_t* woff8 = nullptr;   // [n_b + 1]
  const int32_t* band_of8 = nullptr; // [wstride8]
  int32_t hdr8 = 0;                 // doubles of header in front of every batch's weights: jb8[n_b] | cb8[n_b] (int32)
  int64_t blk8 = 0;                 // doubles per batch block = hdr8 + wstride8 * 8
  double* WA8 = nullptr;            // [n_zb8, blk8]
  const int64_t* row_map = nullptr; // [T] or nullptr
  int64_t T_total = 0;
  int32_t out_pitch = 0;            // row pitch of out in doubles (>= n_b); pad columns are written as 0
  double* out = nullptr;            // [n_z, T_total, out_pitch]
};

cudaError_t launch_k2_permute(const double* Fnu, const int32_t* perm, double* Fs, int64_t T, int32_t n_lam,
                              int32_t ldF, int32_t* nonfinite, cudaStream_t stream);
cudaError_t launch_k2_ranges(const K2Problem& p, cudaStream_t stream);
cudaError_t launch_k2_weights(const K2Problem& p, cudaStream_t stream);
cudaError_t launch_k2_batch_ranges(const K2Problem& p, cudaStream_t stream);
cudaError_t launch_k2_weights_aligned(const K2Problem& p, cudaStream_t stream);
bool k2_sw_fits(const K2Problem& p);     // the shared-window kernel's tiles fit shared memory
cudaError_t launch_k2_batch_ranges8(const K2Problem& p, cudaStream_t stream);
cudaError_t launch_k2_weights_aligned8(const K2Problem& p, cudaStream_t stream);
bool k2_mma_fits(const K2Problem& p);    // the tensor-core kernel's tiles fit shared memory
cudaError_t launch_k2_contract(const K2Problem& p, cudaStream_t stream);

}  // namespace agnf
