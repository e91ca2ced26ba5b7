// K2 — dictionary build kernels (see k2_dictbuild.cuh for the mapping to the reference).
#include "k2_dictbuild.cuh"

namespace agnf {

namespace {

constexpr double C_LIGHT = 2.997e8;   // DICTIONARIES_AGNfitter.py:483 (sic)
constexpr double ANGSTROM = 1.e10;    // DICTIONARIES_AGNfitter.py:484
constexpr int K2_THREADS = 256;
constexpr int K2_TILE_T = 64;         // templates per CTA in the contraction

// Redshifted template wavelength and interp1d's midpoints, with explicitly rounded operations so
// that no FMA contraction can change a comparison: x_j = lam0_j * (1+z);  mid_j = x_j/2 + x_{j+1}/2.
__device__ __forceinline__ double x_at(const double* __restrict__ lam0, int j, double zp1) {
  return __dmul_rn(__ldg(lam0 + j), zp1);
}
__device__ __forceinline__ double mid_at(const double* __restrict__ lam0, int j, double zp1) {
  return __dadd_rn(__dmul_rn(x_at(lam0, j, zp1), 0.5), __dmul_rn(x_at(lam0, j + 1, zp1), 0.5));
}

// np.searchsorted(mid, v, side='left').clip(0, n-1): number of midpoints strictly below v.
__device__ __forceinline__ int nearest_index(const double* __restrict__ lam0, int n, double zp1, double v) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int m = (lo + hi) >> 1;
    if (mid_at(lam0, m, zp1) < v) lo = m + 1; else hi = m;
  }
  return lo;
}

__global__ void k2_permute_kernel(const double* __restrict__ Fnu, const int32_t* __restrict__ perm,
                                  double* __restrict__ Fs, int64_t T, int32_t n_lam, int32_t ldF) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * (int64_t)ldF) return;
  const int64_t t = idx / ldF;
  const int j = (int)(idx - t * ldF);
  Fs[idx] = (j < n_lam) ? __ldg(Fnu + t * n_lam + __ldg(perm + j)) : 0.0;
}

__global__ void k2_ranges_kernel(K2Problem p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  if (idx >= p.n_z * n_b) return;
  const int zi = idx / n_b, b = idx - zi * n_b;
  const double zp1 = __dadd_rn(1.0, __ldg(p.z + zi));
  const int64_t o0 = __ldg(p.filt.off + b), o1 = __ldg(p.filt.off + b + 1);
  int jl = 0, c = 0;
  if (o1 > o0) {
    jl = nearest_index(p.lam0, p.n_lam, zp1, __ldg(p.filt.lam + o0));
    const int jh = nearest_index(p.lam0, p.n_lam, zp1, __ldg(p.filt.lam + o1 - 1));
    c = jh - jl + 1;
  }
  p.jlo[idx] = jl;
  p.cnt[idx] = c;
  atomicMax(p.d_rmax, c);
}

// One thread per (z, band, r): the weight of template sample j = j_lo + r in band b at redshift z.
__global__ void k2_weights_kernel(K2Problem p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  const int64_t total = (int64_t)p.n_z * n_b * p.rmax;
  if (idx >= total) return;
  const int r = (int)(idx % p.rmax);
  const int zb = (int)(idx / p.rmax);
  const int zi = zb / n_b, b = zb - zi * n_b;
  double w = 0.0;
  if (r < __ldg(p.cnt + zb)) {
    const int j = __ldg(p.jlo + zb) + r;
    const int n = p.n_lam;
    const double zp1 = __dadd_rn(1.0, __ldg(p.z + zi));
    const double* __restrict__ fl = p.filt.lam;
    const int64_t o0 = __ldg(p.filt.off + b), o1 = __ldg(p.filt.off + b + 1);
    // samples that resample to j: searchsorted(mid, lam, 'left') == j  <=>  mid[j-1] < lam <= mid[j],
    // and inside [x_0, x_last] (everything else is filled with 0).
    int64_t lo = o0, hi = o1;
    if (j == 0) {
      const double x0 = x_at(p.lam0, 0, zp1);
      while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (__ldg(fl + m) < x0) lo = m + 1; else hi = m; }
    } else {
      const double thr = mid_at(p.lam0, j - 1, zp1);
      while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (__ldg(fl + m) <= thr) lo = m + 1; else hi = m; }
    }
    const int64_t i0 = lo;
    const double up = (j == n - 1) ? x_at(p.lam0, n - 1, zp1) : mid_at(p.lam0, j, zp1);
    lo = i0;
    hi = o1;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (__ldg(fl + m) <= up) lo = m + 1; else hi = m; }
    const int64_t i1 = lo;
    double s = 0.0;
    for (int64_t i = i0; i < i1; ++i) s += __ldg(p.filt.wn + i);
    const double xj = x_at(p.lam0, j, zp1);
    const double q = 1.0 / (xj * xj) * C_LIGHT * ANGSTROM;     // F_nu -> F_lambda at lam_j
    w = (q * s) * __ldg(p.filt.scale + b);
  }
  p.W[idx] = w;
}

// Banded contraction.  grid = (n_z, ceil(T / K2_TILE_T)); the z-slice's weights are staged in
// shared memory and reused by the tile's templates; outputs are written band-fastest so a warp
// stores contiguous rows of the [n_z, T, n_b] table.
// W_SMEM = false: the weight rows are too long for shared memory (e.g. X-ray bands over the
// 1000-point power-law tail of MODEL_AGNfitter.XRAYS) and are read through L1/L2 instead.
template <bool W_SMEM>
__global__ void __launch_bounds__(K2_THREADS) k2_contract_kernel(K2Problem p) {
  extern __shared__ double s_w[];
  const int n_b = p.filt.n_b;
  const int rmax = p.rmax;
  const int rp = W_SMEM ? (rmax | 1) : 0;         // odd row pitch: spreads bands over banks
  int32_t* s_jlo = reinterpret_cast<int32_t*>(s_w + (size_t)n_b * rp);
  int32_t* s_cnt = s_jlo + n_b;
  const int zi = blockIdx.x;
  const double* __restrict__ Wz = p.W + (int64_t)zi * n_b * rmax;
  if (W_SMEM) {
    for (int i = threadIdx.x; i < n_b * rmax; i += blockDim.x) {
      const int b = i / rmax, r = i - b * rmax;
      s_w[b * rp + r] = __ldg(Wz + i);
    }
  }
  for (int b = threadIdx.x; b < n_b; b += blockDim.x) {
    s_jlo[b] = __ldg(p.jlo + zi * n_b + b);
    s_cnt[b] = __ldg(p.cnt + zi * n_b + b);
  }
  __syncthreads();
  const int64_t t0 = (int64_t)blockIdx.y * K2_TILE_T;
  const int pitch = p.out_pitch;
  const int n_out = K2_TILE_T * pitch;
  for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
    const int tt = o / pitch, b = o - tt * pitch;
    const int64_t t = t0 + tt;
    if (t >= p.T) break;
    if (b >= n_b) {                 // pad column
      const int64_t prow = p.row_map ? __ldg(p.row_map + t) : t;
      p.out[((int64_t)zi * p.T_total + prow) * pitch + b] = 0.0;
      continue;
    }
    const double* __restrict__ F = p.Fs + t * p.ldF + s_jlo[b];
    const double* w = W_SMEM ? (s_w + b * rp) : (Wz + b * rmax);
    const int c = s_cnt[b];
    double acc = 0.0;
    for (int r = 0; r < c; ++r) acc = fma(w[r], __ldg(F + r), acc);
    const int64_t row = p.row_map ? __ldg(p.row_map + t) : t;
    p.out[((int64_t)zi * p.T_total + row) * pitch + b] = acc;
  }
}

}  // namespace

cudaError_t launch_k2_permute(const double* Fnu, const int32_t* perm, double* Fs, int64_t T, int32_t n_lam,
                              int32_t ldF, cudaStream_t stream) {
  const int64_t total = T * (int64_t)ldF;
  if (total == 0) return cudaSuccess;
  k2_permute_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(Fnu, perm, Fs, T, n_lam, ldF);
  return cudaGetLastError();
}

cudaError_t launch_k2_ranges(const K2Problem& p, cudaStream_t stream) {
  const int total = p.n_z * p.filt.n_b;
  if (total == 0) return cudaSuccess;
  k2_ranges_kernel<<<(total + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_k2_weights(const K2Problem& p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.n_z * p.filt.n_b * p.rmax;
  if (total == 0) return cudaSuccess;
  k2_weights_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_k2_contract(const K2Problem& p, cudaStream_t stream) {
  if (p.n_z == 0 || p.T == 0) return cudaSuccess;
  const int rp = p.rmax | 1;
  const size_t idx_bytes = 2 * (size_t)p.filt.n_b * sizeof(int32_t);
  const size_t smem = (size_t)p.filt.n_b * rp * sizeof(double) + idx_bytes;
  dim3 grid((unsigned)p.n_z, (unsigned)((p.T + K2_TILE_T - 1) / K2_TILE_T));
  if (smem <= 96 * 1024) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k2_contract_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
      if (e != cudaSuccess) return e;
    }
    k2_contract_kernel<true><<<grid, K2_THREADS, smem, stream>>>(p);
  } else {
    k2_contract_kernel<false><<<grid, K2_THREADS, idx_bytes, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace agnf
