// K1 — batched (source x walker) log-posterior for sm_100a.
//
// One thread evaluates one row of `params` end to end (snap -> gather -> normalise -> chi^2 ->
// priors).  Rows of one source are contiguous in a batch, so the per-source vectors (y, 1/sigma,
// reddening coefficients, bounds) are warp-uniform broadcast loads and the z-slice of the
// dictionary a warp touches stays hot in L1/L2; the per-walker template rows are 8*n_b-byte
// contiguous gathers.  All arithmetic is fp64.
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
#include <cuda_runtime.h>
#include "../../include/agnfit_b200.h"

namespace agnf {

struct DimDesc {
  int32_t is_grid;
  int32_t size;
  int32_t smem_off;   // offset (in values) of this dimension inside the staged grid arrays
  int32_t pad;
  int64_t stride;
};

struct FamDesc {
  const double* table;  // [n_z, T, n_b]
  const double* aux;    // [T] or nullptr
  int64_t T;
  int32_t n_dim;
  int32_t par_offset;
  DimDesc dim[AGNF_MAX_DIMS];
};

struct K1Params {
  agnf_config cfg;
  FamDesc fam[AGNF_MAX_FAMILIES];
  int32_t n_grid_vals;        // total staged grid values
  int32_t norm_offset;        // column of GA in a params row
  const double* grid_sorted;  // [n_grid_vals]
  const int32_t* grid_perm;   // [n_grid_vals]
  // z-slices
  const double* cgal;         // [n_z, n_b]  -0.4 k_Calzetti
  const double* cbbb;         // [n_z, n_b]  -0.4 k_Prevot
  const uint64_t* xray;       // [n_z]
  const int32_t* j1500;       // [n_z]
  const double* omega;        // [n_z, n_b]
  // sources
  const int32_t* zi;          // [S]
  const double* y;            // [S, n_b]
  const double* isig;         // [S, n_b]
  const uint64_t* valid;      // [S]
  const double* pmin;         // [S, n_par]
  const double* pmax;         // [S, n_par]
  const double* consts;       // [S, n_const]
  int32_t n_const;
  int32_t pad;
  const int32_t* agn_branch;  // [S]
  const int32_t* irsyn_mode;  // [S]
  const int32_t* irsyn_j;     // [S]
  int64_t S;
};

struct K1Batch {
  const double* params;       // [n, n_par]
  const int32_t* source_id;   // [n]
  int64_t n;
  double* lnp;                // [n]
  int32_t* rows;              // [n, n_fam] or nullptr
  double* ymodel;             // [n, n_b]   or nullptr
  double* lnprior;            // [n]        or nullptr
  double* lnlike;             // [n]        or nullptr
};

cudaError_t launch_k1(const K1Params& P, const K1Batch& B, cudaStream_t stream);

// tiled variant (k1_tile.cu): same results, warp-cooperative cp.async staging of the template rows
bool k1_tile_supported(const K1Params& P);
cudaError_t launch_k1_tile(const K1Params& P, const K1Batch& B, cudaStream_t stream);
cudaError_t launch_exp10_fast(const double* x, double* y, int64_t n, cudaStream_t stream);

}  // namespace agnf
