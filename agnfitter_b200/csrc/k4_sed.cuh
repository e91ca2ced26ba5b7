// K4 — full-resolution SED realisations and integrated luminosities for a batch of (source, parameter row)
// pairs.
//
// The reference evaluates, per source, ~100 rows drawn from the chain (PLOTandWRITE_AGNfitter.py:566-799): per
// row and component it looks up the full-resolution template (`*_4plot`, DICTIONARIES_AGNfitter.py:293-422),
// applies the free-parameter reddening law to the whole template, interpolates log10 F_nu linearly onto
// all_nus_rest = arange(lognu_min, lognu_max, 0.001) (6 000 - 10 000 points, scipy interp1d -> numpy.interp),
// scales by the normalisations, converts to nu L_nu (PLT:800-821) and integrates over the requested bands
// (PLT:824-964).  Here one CTA owns one row:
//   1. snap the row's parameters to template rows (same routine as K1),
//   2. build the log10 F_nu rows of its 4-5 templates (+ the dereddened disc) in shared memory, reddening applied
//      per template sample exactly where the reference applies it,
//   3. integrals mode: every requested band is swept once, a warp taking 31 trapezoid intervals per pass
//      (32 evaluations, neighbours exchanged by shuffle); all components the band needs are evaluated at the
//      same point; fixed-order reductions, so results are reproducible run to run;
//      spectra mode: every grid point once, all components + their total written as nu L_nu.
// The bracketing interval and interpolation weight of every grid point inside every template grid are a
// host-built table (the grids are shared by all rows), so the inner loop is: one 16-byte table read, two
// shared-memory reads, one 10^x.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "k1_lnprob.cuh"

namespace agnf {

constexpr int K4_MAX_RANGES = 32;
constexpr int K4_MAX_SPECS = 64;
constexpr int K4_COMPS = 6;          // sb, bb, ga, to, rad, bbdered (FLUXES4plotting order without the total)

// Where grid point i falls inside a template grid: x[j] <= grid_i < x[j+1] (numpy.interp's bracket; j = -1:
// outside the template, interp1d's fill value) and w = (grid_i - x[j]) / (x[j+1] - x[j]), 0 on a template sample.
struct __align__(16) K4Bracket {
  double w;
  int32_t j;
  int32_t pad;
};

struct K4Family {
  const double* x = nullptr;         // sorted template grids of all groups, concatenated
  const double* k = nullptr;         // reddening curve on the same layout (modes 1-3) or nullptr
  const double* val = nullptr;       // template samples in sorted-x order: F_nu (modes 1-3) or log10 F_nu (mode 0)
  const double* val2 = nullptr;      // accretion disc only: log10 F_nu of the raw templates (dereddened curve)
  const int64_t* t_off = nullptr;    // [T] offset of template t inside val
  const int32_t* t_grp = nullptr;    // [T] grid group of template t
  const int32_t* g_xoff = nullptr;   // [n_groups] offset of the group's grid inside x / k
  const int32_t* g_nlam = nullptr;   // [n_groups]
  const K4Bracket* btab = nullptr;   // [n_groups, G] bracket of every grid point inside the group's grid
  int64_t T = 0;
  int32_t mode = 0;                  // 0 plain; 1 Calzetti; 2 Prevot; 3 Prevot below 10^16.685 Hz, alpha_ox scaling above
  int32_t e_col = 0, a_col = 0;      // params columns of E(B-V) and alphaScat
  int32_t smem_off = 0;              // offset (doubles) of this family's row in shared memory
  int32_t max_nlam = 0;
  int32_t present = 0;
};

struct K4Params {
  K4Family fam[AGNF_MAX_FAMILIES];
  int32_t G = 0;
  int32_t to_lo = 0, to_hi = 0;      // the torus is kept on grid points [to_lo, to_hi) (PLT:733-734)
  int32_t rad_hi = 0;                // the radio component on [0, rad_hi) (PLT:693)
  int32_t dered_mode = 0;            // see agnfitter_b200/plotset.py
  int32_t dered_off = 0;             // shared-memory offset of the dereddened disc row
  int32_t smem_doubles = 0;
  const int32_t* dered_row = nullptr;  // [T_BB]
  // dered_mode 3 (SN12 without X-rays, PLT:714-717): the dereddened curve is interpolated over the BAND CENTRES
  // from the filtered dictionary row (K1's table), not from a full-resolution template
  const K4Bracket* bands_bt = nullptr; // [G] brackets inside the grid of dictionary band centres
  const double* grid = nullptr;      // [G] all_nus_rest
  const double* xg = nullptr;        // [G] 10**all_nus_rest
  const double* src_z = nullptr;     // [S]
  const double* src_dlum = nullptr;  // [S]
  // integration bands
  int32_t n_ranges = 0, n_specs = 0;
  int32_t r_i0[K4_MAX_RANGES];       // first grid point of the band; monochromatic: the point
  int32_t r_i1[K4_MAX_RANGES];       // one past the last point; -1: monochromatic
  int32_t r_mask[K4_MAX_RANGES];     // components needed (bit c)
  int32_t s_kind[K4_MAX_SPECS];      // 0: sum of the numerator components; 1: 10**(numerator / denominator)
  int32_t s_range[K4_MAX_SPECS];
  int32_t s_num[K4_MAX_SPECS];
  int32_t s_den[K4_MAX_SPECS];
};

struct K4Batch {
  const double* params;              // [n, n_par]
  const int32_t* source_id;          // [n]
  int64_t n;
  double* intlum;                    // [n, n_specs] (integrals mode)
  double* spectra;                   // [n, n_comp, G] nu L_nu (spectra mode), n_comp = 6 or 7 with radio
};

constexpr int K4_THREADS = 256;

cudaError_t launch_k4_intlum(const K1Params& P, const K4Params* Qdev, const K4Params& Qhost, const K4Batch& B,
                             cudaStream_t stream);
cudaError_t launch_k4_spectra(const K1Params& P, const K4Params* Qdev, const K4Params& Qhost, const K4Batch& B,
                              cudaStream_t stream);

}  // namespace agnf
