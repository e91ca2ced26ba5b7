// K3 — ensemble step kernels (see k3_ensemble.cuh).
#include "k3_ensemble.cuh"
#include "k3_common.cuh"

#include <cmath>

namespace agnf {

namespace {

constexpr int K3_THREADS = 256;

__global__ void k3_fill_sid_kernel(int32_t* __restrict__ sid, int64_t S, int32_t half) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S * half) sid[i] = (int32_t)(i / half);
}

// one thread per (source, walker of the active half)
__global__ void __launch_bounds__(K3_THREADS) k3_propose_kernel(K3State st, K3Scratch sc, int64_t step, int half) {
  const int H = st.W / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.S * H) return;
  const int64_t s = i / H;
  const int k = (int)(i - s * H);
  const StretchDraw d = stretch_draw(st.seed, st.a, step, half, s + st.src_offset, k, H);
  const double zz = d.zz;
  const int j = d.j;
  const int np = st.n_par;
  const double* __restrict__ xk = st.pos + ((s * st.W) + split_walker(st.split, s, st.W, (half ? H : 0) + k)) * np;  // moved
  const double* __restrict__ cj = st.pos + ((s * st.W) + split_walker(st.split, s, st.W, (half ? 0 : H) + j)) * np;  // partner
  double* __restrict__ q = sc.prop + i * np;
  for (int c = 0; c < np; ++c) q[c] = stretch_coord(cj[c], xk[c], zz);
  sc.factor[i] = (double)(np - 1) * log(zz);
  sc.lnu[i] = d.lnu;
}

__global__ void __launch_bounds__(K3_THREADS) k3_accept_kernel(K3State st, K3Scratch sc, int half) {
  const int H = st.W / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.S * H) return;
  const int64_t s = i / H;
  const int k = (int)(i - s * H);
  const int64_t w = s * st.W + split_walker(st.split, s, st.W, (half ? H : 0) + k);
  const double newl = sc.prop_lnp[i];
  const double lnpdiff = sc.factor[i] + newl - st.lnp[w];
  if (lnpdiff > sc.lnu[i]) {          // false for NaN and for -inf proposals
    const int np = st.n_par;
    const double* __restrict__ q = sc.prop + i * np;
    double* __restrict__ x = st.pos + w * np;
    for (int c = 0; c < np; ++c) x[c] = q[c];
    st.lnp[w] = newl;
    st.n_acc[w] += 1;
  }
}

// one warp per source: rank every walker's key among the source's W keys (ties by walker index), k3_common.cuh
__global__ void __launch_bounds__(K3_THREADS) k3_split_kernel(K3State st, int64_t step) {
  extern __shared__ uint32_t s_split[];                   // [warps per CTA][split_scratch_words(W)]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * (K3_THREADS / 32) + warp;
  if (s >= st.S) return;
  split_rank_warp(st.split + s * st.W, s_split + warp * split_scratch_words(st.W), st.W, step, s + st.src_offset,
                  st.seed, lane);
}

// one thread per (source, walker, parameter)
__global__ void __launch_bounds__(K3_THREADS) k3_hist_kernel(K3State st, K3Hist h) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = st.S * st.W * st.n_par;
  if (i >= total) return;
  const int c = (int)(i % st.n_par);
  const int64_t s = i / ((int64_t)st.W * st.n_par);
  const double2 mm = h.pbox[s * st.n_par + c];
  const double v = st.pos[i];
  const double f = (v - mm.x) / (mm.y - mm.x) * (double)h.n_bins;
  int b = (int)f;
  b = max(0, min(h.n_bins - 1, b));
  if (v == v) atomicAdd(h.hist + (s * st.n_par + c) * h.n_bins + b, 1u);
}

// one thread per (source, parameter)
__global__ void __launch_bounds__(K3_THREADS) k3_percentiles_kernel(K3Hist h, int64_t S, int32_t n_par,
                                                                     const double* __restrict__ q, int32_t n_q,
                                                                     double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * n_par) return;
  const uint32_t* __restrict__ hh = h.hist + i * h.n_bins;
  const double2 mm = h.pbox[i];
  uint64_t total = 0;
  for (int b = 0; b < h.n_bins; ++b) total += hh[b];
  const double width = (mm.y - mm.x) / (double)h.n_bins;
  int b = 0;
  uint64_t cum = 0;              // samples in bins < b
  for (int k = 0; k < n_q; ++k) {    // q ascending
    double val = NAN;
    if (total > 0) {
      const double target = q[k] * 0.01 * (double)total;
      while (b < h.n_bins - 1 && (double)(cum + hh[b]) < target) {
        cum += hh[b];
        ++b;
      }
      const double inbin = hh[b] > 0 ? (target - (double)cum) / (double)hh[b] : 0.5;
      val = mm.x + ((double)b + fmin(fmax(inbin, 0.0), 1.0)) * width;
    }
    out[i * n_q + k] = val;
  }
}

// one thread per source: walker with the largest lnp; keeps the stored best (initialised to -inf) if larger
__global__ void __launch_bounds__(K3_THREADS) k3_best_kernel(K3State st, double* __restrict__ best_pos,
                                                              double* __restrict__ best_lnp) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.S) return;
  double best = best_lnp[s];
  int arg = -1;
  for (int k = 0; k < st.W; ++k) {
    const double v = st.lnp[s * st.W + k];
    if (v > best) {
      best = v;
      arg = k;
    }
  }
  if (arg >= 0) {
    best_lnp[s] = best;
    for (int c = 0; c < st.n_par; ++c) best_pos[s * st.n_par + c] = st.pos[(s * st.W + arg) * st.n_par + c];
  }
}

// ---- exact percentiles by radix select ------------------------------------------------------------------------
constexpr int K3_SEL_R = 2 * K3_SELECT_MAX_Q;            // order statistics per (source, column)

__device__ __forceinline__ uint64_t sel_key(double v) {   // order-preserving: -inf < ... < -0 < +0 < ... < +inf < NaN
  const uint64_t b = (uint64_t)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double sel_val(uint64_t k) {
  const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(K3_THREADS) k3_select_kernel(K3Select a) {
  extern __shared__ uint32_t s_hist[];                     // [slots][256]
  __shared__ uint64_t s_prefix[K3_SEL_R];                  // key bits fixed so far, per order statistic
  __shared__ uint64_t s_rank[K3_SEL_R];                    // rank among the samples that share the prefix
  __shared__ int32_t s_slot[K3_SEL_R];                     // histogram of the statistic's prefix
  __shared__ uint64_t s_slot_prefix[K3_SEL_R];
  __shared__ int32_t s_nslots;
  const int64_t s = blockIdx.x / a.E;
  const int c = (int)(blockIdx.x - s * a.E);
  const int64_t N = a.n_used * a.W;
  const int R = 2 * a.n_q;
  if (N <= 0) {
    if (threadIdx.x < a.n_q) a.out[((int64_t)s * a.E + c) * a.n_q + threadIdx.x] = NAN;
    return;
  }
  if (threadIdx.x < R) {
    // numpy: virtual index (N - 1) * (q / 100), lower neighbour floor(), upper min(lower + 1, N - 1)
    const double h = (double)(N - 1) * (a.q[threadIdx.x >> 1] / 100.0);
    const int64_t lo = (int64_t)floor(h);
    s_rank[threadIdx.x] = (uint64_t)((threadIdx.x & 1) ? min(lo + 1, N - 1) : lo);
    s_prefix[threadIdx.x] = 0;
  }
  const int64_t row_stride = (int64_t)a.S * a.W * a.E;
  const double* __restrict__ base = a.data + (a.first * a.S + s) * (int64_t)a.W * a.E + c;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    __syncthreads();
    if (threadIdx.x == 0) {                                // distinct prefixes -> histogram slots
      int n = 0;
      for (int r = 0; r < R; ++r) {
        int f = -1;
        for (int k = 0; k < n; ++k)
          if (s_slot_prefix[k] == s_prefix[r]) f = k;
        if (f < 0) {
          f = n++;
          s_slot_prefix[f] = s_prefix[r];
        }
        s_slot[r] = f;
      }
      s_nslots = n;
    }
    __syncthreads();
    const int ns = s_nslots;
    for (int i = threadIdx.x; i < ns * 256; i += K3_THREADS) s_hist[i] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < N; i += K3_THREADS) {
      const int64_t r = i / a.W;
      const int w = (int)(i - r * a.W);
      const uint64_t k = sel_key(__ldg(base + r * a.stride * row_stride + (int64_t)w * a.E));
      const uint64_t hi = pass == 0 ? 0 : (k >> (shift + 8));
      const unsigned d = (unsigned)(k >> shift) & 255u;
      for (int sl = 0; sl < ns; ++sl)
        if (hi == s_slot_prefix[sl]) atomicAdd(&s_hist[sl * 256 + d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < R) {
      const uint32_t* hh = s_hist + s_slot[threadIdx.x] * 256;
      uint64_t rank = s_rank[threadIdx.x], cum = 0;
      int d = 0;
      for (; d < 255; ++d) {
        if (cum + hh[d] > rank) break;
        cum += hh[d];
      }
      s_rank[threadIdx.x] = rank - cum;
      s_prefix[threadIdx.x] = (s_prefix[threadIdx.x] << 8) | (uint64_t)d;
    }
  }
  __syncthreads();
  if (threadIdx.x < a.n_q) {
    const double lo = sel_val(s_prefix[2 * threadIdx.x]), hi = sel_val(s_prefix[2 * threadIdx.x + 1]);
    const double h = (double)(N - 1) * (a.q[threadIdx.x] / 100.0);
    const double t = h - floor(h);
    // numpy's _lerp: a + (b - a) t, and b - (b - a) (1 - t) where t >= 0.5
    const double diff = hi - lo;
    double v = __dadd_rn(lo, __dmul_rn(diff, t));
    if (t >= 0.5) v = __dsub_rn(hi, __dmul_rn(diff, 1.0 - t));
    a.out[((int64_t)s * a.E + c) * a.n_q + threadIdx.x] = v;
  }
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + K3_THREADS - 1) / K3_THREADS); }

}  // namespace

// pos[s, w, p] = centre[s, p] + scale * d, d uniform in (-1, 1) or standard normal (Box-Muller), drawn from
// Philox with counter (stream, 2 + p, source + offset, walker): disjoint from the stretch move's counters, whose
// second word is the half-ensemble index 0 / 1
__global__ void k3_ball_kernel(K3State st, const double* __restrict__ centre, double scale, int normal,
                               uint32_t stream_id) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = st.S * st.W * st.n_par;
  if (i >= n) return;
  const int p = (int)(i % st.n_par);
  const int64_t sw = i / st.n_par;
  const int w = (int)(sw % st.W);
  const int64_t s = sw / st.W;
  uint32_t x[4];
  philox4x32_10(stream_id, 2u + (uint32_t)p, (uint32_t)(s + st.src_offset), (uint32_t)w, (uint32_t)st.seed,
                (uint32_t)(st.seed >> 32), x);
  const double u1 = (double)((((uint64_t)x[0] << 32) | x[1]) >> 11) * 1.1102230246251565e-16;
  const double u2 = (double)((((uint64_t)x[2] << 32) | x[3]) >> 11) * 1.1102230246251565e-16;
  double d;
  if (normal) d = sqrt(-2.0 * log1p(-u1)) * cospi(2.0 * u2);
  else d = __dsub_rn(__dmul_rn(2.0, u1), 1.0);
  st.pos[i] = __dadd_rn(centre[s * st.n_par + p], __dmul_rn(d, scale));
}

cudaError_t launch_k3_ball(const K3State& st, const double* centre, double scale, int normal, uint32_t stream_id,
                           cudaStream_t stream) {
  const int64_t n = st.S * st.W * st.n_par;
  if (n == 0) return cudaSuccess;
  k3_ball_kernel<<<blocks_for(n), K3_THREADS, 0, stream>>>(st, centre, scale, normal, stream_id);
  return cudaGetLastError();
}

cudaError_t launch_k3_fill_sid(int32_t* sid, int64_t S, int32_t half, cudaStream_t stream) {
  if (S * half <= 0) return cudaSuccess;
  k3_fill_sid_kernel<<<blocks_for(S * half), K3_THREADS, 0, stream>>>(sid, S, half);
  return cudaGetLastError();
}

cudaError_t launch_k3_split(const K3State& st, int64_t step, cudaStream_t stream) {
  if (!st.split || st.S == 0) return cudaSuccess;
  const int wpc = K3_THREADS / 32;
  const size_t smem = (size_t)wpc * split_scratch_words(st.W) * sizeof(uint32_t);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k3_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k3_split_kernel<<<(unsigned)((st.S + wpc - 1) / wpc), K3_THREADS, smem, stream>>>(st, step);
  return cudaGetLastError();
}

cudaError_t launch_k3_propose(const K3State& st, const K3Scratch& sc, int64_t step, int half, cudaStream_t stream) {
  k3_propose_kernel<<<blocks_for(st.S * (st.W / 2)), K3_THREADS, 0, stream>>>(st, sc, step, half);
  return cudaGetLastError();
}

cudaError_t launch_k3_accept(const K3State& st, const K3Scratch& sc, int half, cudaStream_t stream) {
  k3_accept_kernel<<<blocks_for(st.S * (st.W / 2)), K3_THREADS, 0, stream>>>(st, sc, half);
  return cudaGetLastError();
}

cudaError_t launch_k3_hist(const K3State& st, const K3Hist& h, cudaStream_t stream) {
  k3_hist_kernel<<<blocks_for(st.S * st.W * st.n_par), K3_THREADS, 0, stream>>>(st, h);
  return cudaGetLastError();
}

cudaError_t launch_k3_percentiles(const K3Hist& h, int64_t S, int32_t n_par, const double* q, int32_t n_q,
                                  double* out, cudaStream_t stream) {
  k3_percentiles_kernel<<<blocks_for(S * n_par), K3_THREADS, 0, stream>>>(h, S, n_par, q, n_q, out);
  return cudaGetLastError();
}

cudaError_t launch_k3_select(const K3Select& a, cudaStream_t stream) {
  if (a.S <= 0 || a.E <= 0 || a.n_q <= 0) return cudaSuccess;
  if (a.n_q > K3_SELECT_MAX_Q) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(2 * a.n_q) * 256 * sizeof(uint32_t);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k3_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         2 * K3_SELECT_MAX_Q * 256 * (int)sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k3_select_kernel<<<(unsigned)(a.S * a.E), K3_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_k3_best(const K3State& st, double* best_pos, double* best_lnp, cudaStream_t stream) {
  k3_best_kernel<<<blocks_for(st.S), K3_THREADS, 0, stream>>>(st, best_pos, best_lnp);
  return cudaGetLastError();
}

}  // namespace agnf
