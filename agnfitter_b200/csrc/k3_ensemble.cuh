// K3 — GPU-resident affine-invariant ensemble step for many sources in lock-step.
//
// The reference hands one source at a time to emcee (MCMC_AGNfitter.py:114-135): a stretch move offers
// only W/2 independent evaluations per step, far too few for a B200.  Here S sources advance together:
// per half-step, k3_propose draws the stretch factors and partners of all S * W/2 walkers and writes
// their proposals as one params batch, K1 evaluates it, k3_accept applies emcee's acceptance rule.
// Walker positions and log-probabilities never leave HBM.
//
// Move (Goodman & Weare 2010 as implemented by emcee's StretchMove/RedBlueMove with a = 2):
//     zz = ((a - 1) u1 + 1)^2 / a            j = floor(u2 * W/2)          (partner in the other half)
//     q  = c_j - (c_j - x_k) * zz            accept iff (n_par - 1) ln zz + lnp(q) - lnp(x_k) > ln u3
// The two sets are drawn afresh at every step, as emcee 3's RedBlueMove does by default (`randomize_split=True`:
// the reference builds EnsembleSampler with the default move, MCMC_AGNfitter.py:114): k3_split_kernel orders the W
// walkers of a source by a 32-bit Philox key (counter (step, 0xFFFFFFFF, source, walker / 4)), the first W/2 form
// set 0.  `split == nullptr` keeps the fixed first/second half (`randomize_split=False`).
// u1, u2, u3 come from Philox4x32-10 keyed by (seed) with counter (step, half, source, slot), so the
// stream is reproducible on the CPU (oracle/ensemble_oracle.py) bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace agnf {

struct K3State {
  double* pos;        // [S, W, n_par]
  double* lnp;        // [S, W]
  int32_t* n_acc;     // [S, W] accepted proposals
  int64_t S;
  int32_t W, n_par;
  uint64_t seed;
  double a;           // stretch scale (emcee default 2.0)
  int64_t src_offset; // added to the local source index in the Philox counter: a shard that starts at catalogue
                      // row `src_offset` draws exactly what a single-GPU run draws for those sources
  uint16_t* split;    // [S, W] this step's split (see above), or nullptr
};

struct K3Scratch {
  double* prop;       // [S * W/2, n_par] proposals of the current half-step
  double* prop_lnp;   // [S * W/2]
  double* factor;     // [S * W/2] (n_par - 1) ln zz
  double* lnu;        // [S * W/2] ln u3
  int32_t* sid;       // [S * W/2] source of each proposal row
};

// histogram accumulation for on-device percentiles: hist [S, n_par, n_bins] uint32 over [pmin, pmax]
struct K3Hist {
  uint32_t* hist;
  const double2* pbox;   // [S, n_par] (min, max)
  int32_t n_bins;
};

// walker positions around per-source centres [S, n_par] (device): centre + scale * (uniform(-1, 1) | normal)
cudaError_t launch_k3_ball(const K3State& st, const double* centre, double scale, int normal, uint32_t stream_id,
                           cudaStream_t stream);
cudaError_t launch_k3_fill_sid(int32_t* sid, int64_t S, int32_t half, cudaStream_t stream);
// split[s, :] = walkers of source s ordered by their Philox key of `step` (ties by walker index)
cudaError_t launch_k3_split(const K3State& st, int64_t step, cudaStream_t stream);
cudaError_t launch_k3_propose(const K3State& st, const K3Scratch& sc, int64_t step, int half, cudaStream_t stream);
cudaError_t launch_k3_accept(const K3State& st, const K3Scratch& sc, int half, cudaStream_t stream);
cudaError_t launch_k3_hist(const K3State& st, const K3Hist& h, cudaStream_t stream);
// percentiles (linear interpolation inside the bin) of every (source, parameter): out [S, n_par, n_q]
cudaError_t launch_k3_percentiles(const K3Hist& h, int64_t S, int32_t n_par, const double* q, int32_t n_q,
                                  double* out, cudaStream_t stream);
// EXACT percentiles of the recorded chain (numpy.percentile, method 'linear' — what PLOTandWRITE_AGNfitter.py:162-191
// applies to CHAIN.props' thinned flat chain): one CTA per (source, column) finds the two order statistics every
// percentile needs by an 8-bit radix select over the order-preserving 64-bit keys of its N = n_used * W samples
// (8 passes, one 256-bin shared-memory histogram per distinct key prefix), then interpolates like numpy's _lerp.
// The samples of (source s, column c): data[((first + r * stride) * S + s) * W * E + w * E + c], r < n_used, w < W
// (E = n_par for the positions, 1 for the log-posterior).  out [S, E, n_q].
struct K3Select {
  const double* data;
  int64_t S;
  int32_t W, E;
  int64_t first, stride, n_used;
  const double* q;      // [n_q] ascending, in [0, 100] (device)
  int32_t n_q;
  double* out;
};
constexpr int K3_SELECT_MAX_Q = 16;
cudaError_t launch_k3_select(const K3Select& a, cudaStream_t stream);
// best (max lnp) walker of every source: best_pos [S, n_par], best_lnp [S]
cudaError_t launch_k3_best(const K3State& st, double* best_pos, double* best_lnp, cudaStream_t stream);

}  // namespace agnf
