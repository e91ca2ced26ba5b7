// K2 — dictionary build kernels (see k2_dictbuild.cuh for the mapping to the reference).
#include "k2_dictbuild.cuh"

#include <climits>

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

// also raises *nonfinite when a template holds a NaN / inf sample (such families take the per-slice kernels)
__global__ void k2_permute_kernel(const double* __restrict__ Fnu, const int32_t* __restrict__ perm,
                                  double* __restrict__ Fs, int64_t T, int32_t n_lam, int32_t ldF,
                                  int32_t* __restrict__ nonfinite) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * (int64_t)ldF) return;
  const int64_t t = idx / ldF;
  const int j = (int)(idx - t * ldF);
  const double v = (j < n_lam) ? __ldg(Fnu + t * n_lam + __ldg(perm + j)) : 0.0;
  Fs[idx] = v;
  if (nonfinite && (v - v) != 0.0) *nonfinite = 1;      // benign race: every writer stores 1
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
  atomicMax(p.d_rmaxb + b, c);
}

// weight of template sample j in band b at redshift index zi (j inside the band's window at that z)
__device__ __forceinline__ double band_weight(const K2Problem& p, int zi, int b, int j) {
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
  return (q * s) * __ldg(p.filt.scale + b);
}

// One thread per (z, slot): the weight of template sample j = j_lo + r in band b at redshift z.
__global__ void k2_weights_kernel(K2Problem p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  const int64_t total = (int64_t)p.n_z * p.wstride;
  if (idx >= total) return;
  const int slot = (int)(idx % p.wstride);
  const int zi = (int)(idx / p.wstride);
  const int b = __ldg(p.band_of + slot);
  double w = 0.0;
  const int r = b >= 0 ? slot - __ldg(p.woff + b) : 0;
  const int zb = zi * n_b + (b >= 0 ? b : 0);
  if (b >= 0 && r < __ldg(p.cnt + zb)) w = band_weight(p, zi, b, __ldg(p.jlo + zb) + r);
  p.W[idx] = w;
}

// Shared-window path.  One thread per (batch, band): the window all 4 slices of the batch see together.
__global__ void k2_batch_ranges_kernel(K2Problem p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  if (idx >= p.n_zb * n_b) return;
  const int zb = idx / n_b, b = idx - zb * n_b;
  int lo = INT_MAX, hi = 0;
  for (int zl = 0; zl < 4; ++zl) {
    const int z = 4 * zb + zl;
    if (z >= p.n_z) break;
    const int c = __ldg(p.cnt + z * n_b + b);
    if (c > 0) {
      const int j = __ldg(p.jlo + z * n_b + b);
      lo = min(lo, j);
      hi = max(hi, j + c);
    }
  }
  const int c = hi > lo ? hi - lo : 0;
  p.jb[idx] = c ? lo : 0;
  p.cb[idx] = c;
  atomicMax(p.d_rmaxbA + b, c);
}

// One thread per (batch, slot, slice): weights re-aligned to the batch window, 4 slices interleaved.
__global__ void k2_weights_aligned_kernel(K2Problem p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  const int64_t total = (int64_t)p.n_zb * p.wstrideA * 4;
  if (idx >= total) return;
  const int zl = (int)(idx & 3);
  const int64_t q = idx >> 2;
  const int slot = (int)(q % p.wstrideA);
  const int zb = (int)(q / p.wstrideA);
  const int z = 4 * zb + zl;
  const int b = __ldg(p.band_ofA + slot);
  double w = 0.0;
  if (b >= 0 && z < p.n_z) {
    const int j = __ldg(p.jb + zb * n_b + b) + (slot - __ldg(p.woffA + b));
    const int r = j - __ldg(p.jlo + z * n_b + b);
    if (r >= 0 && r < __ldg(p.cnt + z * n_b + b) && slot - __ldg(p.woffA + b) < __ldg(p.cb + zb * n_b + b))
      w = band_weight(p, z, b, j);
  }
  p.WA[(int64_t)zb * p.blkA + p.hdrA + (int64_t)slot * 4 + zl] = w;
}

// window table of every batch (jb[n_b] | cb[n_b], int32) in front of its weights, see k2_headers8_kernel
__global__ void k2_headersA_kernel(K2Problem p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  if (idx >= p.n_zb * n_b) return;
  const int zb = idx / n_b, b = idx - zb * n_b;
  int32_t* h = reinterpret_cast<int32_t*>(p.WA + (int64_t)zb * p.blkA);
  h[b] = p.jb[idx];
  h[n_b + b] = p.cb[idx];
}

// Tensor-core path.  One thread per (batch of 8 slices, band): the window all 8 slices see together, its start
// rounded down and its length rounded up to multiples of 4 samples (the k extent of one mma).
__global__ void k2_batch_ranges8_kernel(K2Problem p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  if (idx >= p.n_zb8 * n_b) return;
  const int zb = idx / n_b, b = idx - zb * n_b;
  int lo = INT_MAX, hi = 0;
  for (int zl = 0; zl < 8; ++zl) {
    const int z = 8 * zb + zl;
    if (z >= p.n_z) break;
    const int c = __ldg(p.cnt + z * n_b + b);
    if (c > 0) {
      const int j = __ldg(p.jlo + z * n_b + b);
      lo = min(lo, j);
      hi = max(hi, j + c);
    }
  }
  int j0 = 0, c = 0;
  if (hi > lo) {
    j0 = lo & ~3;
    c = (hi - j0 + 3) & ~3;
  }
  p.jb8[idx] = j0;
  p.cb8[idx] = c;
  atomicMax(p.d_rmaxb8 + b, c);
}

// One thread per (batch, slot, slice): weights re-aligned to the batch window, 8 slices interleaved.
__global__ void k2_weights_aligned8_kernel(K2Problem p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  const int64_t total = (int64_t)p.n_zb8 * p.wstride8 * 8;
  if (idx >= total) return;
  const int zl = (int)(idx & 7);
  const int64_t q = idx >> 3;
  const int slot = (int)(q % p.wstride8);
  const int zb = (int)(q / p.wstride8);
  const int z = 8 * zb + zl;
  const int b = __ldg(p.band_of8 + slot);
  double w = 0.0;
  if (b >= 0 && z < p.n_z) {
    const int j = __ldg(p.jb8 + zb * n_b + b) + (slot - __ldg(p.woff8 + b));
    const int r = j - __ldg(p.jlo + z * n_b + b);
    if (r >= 0 && r < __ldg(p.cnt + z * n_b + b)) w = band_weight(p, z, b, j);
  }
  p.WA8[(int64_t)zb * p.blk8 + p.hdr8 + (int64_t)slot * 8 + zl] = w;
}

// Every batch's weight block starts with the window table of the batch (jb8[n_b] | cb8[n_b], int32): it reaches
// shared memory with the same bulk copy as the weights, so a work unit starts without a global-memory round trip.
__global__ void k2_headers8_kernel(K2Problem p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_b = p.filt.n_b;
  if (idx >= p.n_zb8 * n_b) return;
  const int zb = idx / n_b, b = idx - zb * n_b;
  int32_t* h = reinterpret_cast<int32_t*>(p.WA8 + (int64_t)zb * p.blk8);
  h[b] = p.jb8[idx];
  h[n_b + b] = p.cb8[idx];
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
  const int ws = p.wstride;
  int32_t* s_jlo = reinterpret_cast<int32_t*>(s_w + (W_SMEM ? ws : 0));
  int32_t* s_cnt = s_jlo + n_b;
  const int zi = blockIdx.x;
  const double* __restrict__ Wz = p.W + (int64_t)zi * ws;
  if (W_SMEM) {
    for (int i = threadIdx.x; i < ws; i += blockDim.x) s_w[i] = __ldg(Wz + i);
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
    const double* w = (W_SMEM ? s_w : Wz) + __ldg(p.woff + b);
    const int c = s_cnt[b];
    double acc = 0.0;
    for (int r = 0; r < c; ++r) acc = fma(w[r], __ldg(F + r), acc);
    const int64_t row = p.row_map ? __ldg(p.row_map + t) : t;
    p.out[((int64_t)zi * p.T_total + row) * pitch + b] = acc;
  }
}


// ---- tiled contraction --------------------------------------------------------------------------
// k2_contract_kernel re-reads the family's whole template block from L2 once per z-slice (C3: 36 MB x
// 5 991 slices).  Here a CTA keeps a tile of TT templates in shared memory (one bulk copy through the TMA
// unit) and walks a chunk of z-slices: per slice the weight rows arrive by a second, double-buffered
// bulk copy, every thread forms its (template, band) dot products from shared memory, the tile of
// results is transposed through shared memory and leaves with coalesced stores.  Template traffic
// drops by the chunk length, weight traffic by TT.
constexpr int K2T_THREADS = 512;

__device__ __forceinline__ uint32_t k2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k2_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void k2_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k2_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void k2_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1 << 24)) __trap();
  }
}

struct K2TilePlan {
  int32_t tt;          // templates per tile (power of two, <= 32)
  int32_t zb;          // z-slices per batch (one weight copy, one synchronisation pair)
  int32_t zc;          // z-slices per CTA (multiple of zb)
  int32_t opitch;      // pitch of a result tile in shared memory: out_pitch (bulk store) or out_pitch + 1
  int32_t bulk_store;  // rows are not remapped and out_pitch is even: slices leave by bulk stores
  int32_t off_w, off_o, off_bar, smem_bytes;   // byte offsets inside dynamic shared memory
};

template <int TT>
__global__ void __launch_bounds__(K2T_THREADS, 1) k2_contract_tile_kernel(K2Problem p, K2TilePlan c) {
  extern __shared__ __align__(16) unsigned char k2_smem[];
  double* Fs_s = reinterpret_cast<double*>(k2_smem);                        // [tt][ldF]
  double* W_s = reinterpret_cast<double*>(k2_smem + c.off_w);               // [2][zb * wstride]
  double* O_s = reinterpret_cast<double*>(k2_smem + c.off_o);               // [zb][tt][opitch]
  const uint32_t bar = k2_smem_u32(k2_smem + c.off_bar);                    // [0] template tile, [1], [2] weights
  const int n_b = p.filt.n_b, ws = p.wstride, ldF = p.ldF, pitch = p.out_pitch, ZB = c.zb;
  constexpr int TASKS_PER_PASS = K2T_THREADS / TT;       // (slice, band) pairs the CTA covers per pass
  const int tt_of_thread = threadIdx.x % TT;
  const int64_t t0 = (int64_t)blockIdx.y * TT;
  const int nt = (int)min((int64_t)TT, p.T - t0);
  const int z0 = blockIdx.x * c.zc;
  const int nz = min(c.zc, p.n_z - z0);
  const int n_batches = (nz + ZB - 1) / ZB;
  if (threadIdx.x == 0) {
    k2_mbar_init(bar, 1);
    k2_mbar_init(bar + 8, 1);
    k2_mbar_init(bar + 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = threadIdx.x; i < ZB * TT * c.opitch; i += K2T_THREADS) O_s[i] = 0.0;   // pad columns stay zero
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t fbytes = (uint32_t)TT * (uint32_t)ldF * 8u;        // Fs is padded to whole tiles with zero rows
    k2_mbar_expect_tx(bar, fbytes);
    k2_bulk_g2s(k2_smem_u32(Fs_s), p.Fs + t0 * ldF, fbytes, bar);
    const uint32_t wb = (uint32_t)min(ZB, nz) * (uint32_t)ws * 8u;
    k2_mbar_expect_tx(bar + 8, wb);
    k2_bulk_g2s(k2_smem_u32(W_s), p.W + (int64_t)z0 * ws, wb, bar + 8);
  }
  k2_mbar_wait(bar, 0);
  for (int ib = 0; ib < n_batches; ++ib) {
    const int buf = ib & 1;
    const int zfirst = z0 + ib * ZB;
    const int nzb = min(ZB, z0 + nz - zfirst);
    if (threadIdx.x == 0 && ib + 1 < n_batches) {   // the other buffer was last read before the previous __syncthreads
      const int znext = zfirst + ZB;
      const uint32_t wb = (uint32_t)min(ZB, z0 + nz - znext) * (uint32_t)ws * 8u;
      k2_mbar_expect_tx(bar + 8 + 8 * (buf ^ 1), wb);
      k2_bulk_g2s(k2_smem_u32(W_s + (size_t)(buf ^ 1) * ZB * ws), p.W + (int64_t)znext * ws, wb,
                  bar + 8 + 8 * (buf ^ 1));
    }
    k2_mbar_wait(bar + 8 + 8 * buf, (ib >> 1) & 1);
    const double* __restrict__ Wb = W_s + (size_t)buf * ZB * ws;
    // one (slice, band) pair per warp pass (TT = 32): the weight reads are broadcasts, the template
    // reads are conflict-free across the 32 templates (odd ldF); four accumulators break the FMA chain
    // Work units = (band, slice) pairs, listed longest band first (band_order) with the slices of a band
    // adjacent, and dealt to the warps in snake order (pass 0: warp 0..15, pass 1: warp 15..0, ...): the
    // three far-infrared bands see ~75 template samples, most optical bands ~15, and a plain round-robin
    // leaves most warps idle behind the two or three that drew the long bands.
    {
      const int n_task = nzb * n_b;
      const int tt = tt_of_thread;
      const int g = threadIdx.x / TT;                       // this thread's unit slot inside a pass
      for (int pass = 0, base = 0; base < n_task; ++pass, base += TASKS_PER_PASS) {
        const int task = base + ((pass & 1) ? (TASKS_PER_PASS - 1 - g) : g);
        if (task >= n_task) continue;
        const int rank = task / nzb, zl = task - rank * nzb;
        const int b = __ldg(p.band_order + rank);
        const int zb_idx = (zfirst + zl) * n_b + b;
        const double* __restrict__ F = Fs_s + tt * ldF + __ldg(p.jlo + zb_idx);
        const double* __restrict__ w = Wb + zl * ws + __ldg(p.woff + b);
        const int cnum = __ldg(p.cnt + zb_idx);
        // weights: 128-bit broadcast reads (woff and wstride are even: the row is 16-byte aligned); slots
        // beyond cnt hold 0 by construction, but the template samples beyond the band's window are NOT
        // read (a non-finite sample there must not leak into this band): the tail is predicated
        const double2* __restrict__ w2 = reinterpret_cast<const double2*>(w);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int r = 0;
        for (; r + 4 <= cnum; r += 4) {
          const double2 wa = w2[r >> 1], wb = w2[(r >> 1) + 1];
          a0 = fma(wa.x, F[r], a0);
          a1 = fma(wa.y, F[r + 1], a1);
          a2 = fma(wb.x, F[r + 2], a2);
          a3 = fma(wb.y, F[r + 3], a3);
        }
        if (r < cnum) a0 = fma(w[r], F[r], a0);
        if (r + 1 < cnum) a1 = fma(w[r + 1], F[r + 1], a1);
        if (r + 2 < cnum) a2 = fma(w[r + 2], F[r + 2], a2);
        O_s[(zl * TT + tt) * c.opitch + b] = (a0 + a1) + (a2 + a3);
      }
    }
    __syncthreads();
    if (c.bulk_store) {
      // the nt x pitch results of a slice are one contiguous block of `out`: with an unpadded tile
      // (c.opitch == pitch) each slice leaves with a single bulk store (shared -> global through the TMA unit)
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy writes -> async proxy
        for (int zl = 0; zl < nzb; ++zl) {
          double* dst = p.out + ((int64_t)(zfirst + zl) * p.T_total + t0) * pitch;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst),
                       "r"(k2_smem_u32(O_s + (size_t)zl * TT * c.opitch)), "r"((uint32_t)(nt * pitch * 8))
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");    // the tile may be overwritten
      }
    } else {
      // remapped rows (ragged families) or an odd pitch (dense host output): element-wise coalesced stores
      const int per_z = nt * pitch;
      const int d_tt = K2T_THREADS / pitch, d_col = K2T_THREADS - d_tt * pitch;
      for (int zl = 0; zl < nzb; ++zl) {
        int tt = threadIdx.x / pitch, col = threadIdx.x - tt * pitch;
        double* __restrict__ dst = p.out + ((int64_t)(zfirst + zl) * p.T_total) * pitch;
        for (int i = threadIdx.x; i < per_z; i += K2T_THREADS) {
          const int64_t row = p.row_map ? __ldg(p.row_map + t0 + tt) : t0 + tt;
          dst[row * pitch + col] = O_s[(zl * TT + tt) * c.opitch + col];
          tt += d_tt;
          col += d_col;
          if (col >= pitch) {
            col -= pitch;
            ++tt;
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---- shared-window contraction --------------------------------------------------------------------
// The tiled kernel above spends 3 shared-memory wavefronts per warp-FMA (2 for the template sample, 1 for the
// broadcast weight).  Consecutive z-slices see almost the same template samples (dz = 0.001 moves a band by a
// fraction of a sample), so here a batch of 4 slices shares every read: per sample of the batch window one
// LDS.64 of the template value and two broadcast LDS.128 of the 4 interleaved weights feed 4 FMAs
// (1 wavefront per warp-FMA, 1.75 instructions).  Every slice keeps ONE accumulator that runs over the window
// in ascending sample order; samples outside the slice's own window carry weight 0 and leave the sum
// unchanged, so a slice's result does not depend on which slices share its batch (bit for bit).
// A round is nb batches (4 nb slices): its (band, batch) units are dealt to the 16 warps longest band first
// in snake order; weights of the next round arrive by a bulk copy while this one is computed.
constexpr int K2_SW_MAX_ROUNDS = 32;
struct K2SwPlan {
  int32_t nb;          // batches per round: 1, 2 or 4
  int32_t nb_shift;
  int32_t nbuf;        // weight buffers: 2 = the next round's weights arrive during this one; 1 = wide filter sets
                       // whose double-buffered weights would not fit (49 bands): loaded at the start of the round
  int32_t zbc;         // batches per CTA (a multiple of nb)
  int32_t opitch;      // pitch of a result tile (out_pitch when slices leave by bulk stores, else + 1)
  int32_t bulk_store;
  int32_t off_w, off_o, off_bar, smem_bytes;
};

template <int TT>
__global__ void __launch_bounds__(K2T_THREADS, 1) k2_contract_sw_kernel(K2Problem p, K2SwPlan c) {
  extern __shared__ __align__(16) unsigned char k2_smem[];
  double* Fs_s = reinterpret_cast<double*>(k2_smem);                        // [TT][ldF]
  double* W_s = reinterpret_cast<double*>(k2_smem + c.off_w);               // [2][nb][wstrideA][4]
  double* O_s = reinterpret_cast<double*>(k2_smem + c.off_o);               // [nb][4][TT][opitch]
  // barriers: +0 template tile, +8/+16 weight buffers, +24 "result tile drained"; then one unit counter per round
  const uint32_t bar = k2_smem_u32(k2_smem + c.off_bar);
  int* next_unit = reinterpret_cast<int*>(k2_smem + c.off_bar + 32);        // [K2_SW_MAX_ROUNDS]
  int* order_s = next_unit + K2_SW_MAX_ROUNDS;                              // [n_b] bands, longest first
  int* woff_s = order_s + p.filt.n_b;                                       // [n_b] first slot of each band
  const int n_b = p.filt.n_b, ldF = p.ldF, pitch = p.out_pitch;
  const int64_t blkA = p.blkA;
  const int tt = threadIdx.x % TT;
  const int64_t t0 = (int64_t)blockIdx.y * TT;
  const int nt = (int)min((int64_t)TT, p.T - t0);
  const int zb0 = blockIdx.x * c.zbc;
  const int zb1 = min(zb0 + c.zbc, p.n_zb);
  const int NB = c.nb;
  const int n_units = n_b * NB;
  const size_t wround = (size_t)NB * blkA;                                  // doubles per weight buffer
  if (threadIdx.x == 0) {
    k2_mbar_init(bar, 1);
    k2_mbar_init(bar + 8, 1);
    k2_mbar_init(bar + 16, 1);
    k2_mbar_init(bar + 24, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < K2_SW_MAX_ROUNDS) next_unit[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < n_b; i += K2T_THREADS) {
    order_s[i] = __ldg(p.band_order + i);
    woff_s[i] = __ldg(p.woffA + i);
  }
  for (int i = threadIdx.x; i < NB * 4 * TT * c.opitch; i += K2T_THREADS) O_s[i] = 0.0;   // pad columns stay zero
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t fbytes = (uint32_t)TT * (uint32_t)ldF * 8u;
    k2_mbar_expect_tx(bar, fbytes);
    k2_bulk_g2s(k2_smem_u32(Fs_s), p.Fs + t0 * ldF, fbytes, bar);
    const uint32_t wbytes = (uint32_t)min(NB, zb1 - zb0) * (uint32_t)blkA * 8u;
    k2_mbar_expect_tx(bar + 8, wbytes);
    k2_bulk_g2s(k2_smem_u32(W_s), p.WA + (int64_t)zb0 * blkA, wbytes, bar + 8);
  }
  k2_mbar_wait(bar, 0);
  int ir = 0;
  for (int zbr = zb0; zbr < zb1; zbr += NB, ++ir) {
    const int buf = (c.nbuf == 2) ? (ir & 1) : 0;
    const int nbr = min(NB, zb1 - zbr);
    if (threadIdx.x == 0) {
      if (c.nbuf == 2 && zbr + NB < zb1) {
        const uint32_t wbytes = (uint32_t)min(NB, zb1 - zbr - NB) * (uint32_t)blkA * 8u;
        k2_mbar_expect_tx(bar + 8 + 8 * (buf ^ 1), wbytes);
        k2_bulk_g2s(k2_smem_u32(W_s + (size_t)(buf ^ 1) * wround), p.WA + (int64_t)(zbr + NB) * blkA, wbytes,
                    bar + 8 + 8 * (buf ^ 1));
      }
      if (c.nbuf == 1 && ir > 0) {       // every warp left the previous round (barrier below): reuse the buffer
        const uint32_t wbytes = (uint32_t)nbr * (uint32_t)blkA * 8u;
        k2_mbar_expect_tx(bar + 8, wbytes);
        k2_bulk_g2s(k2_smem_u32(W_s), p.WA + (int64_t)zbr * blkA, wbytes, bar + 8);
      }
      if (ir > 0 && c.bulk_store) {      // the previous round's result tile has left shared memory: release it
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar + 24) : "memory");
      }
    }
    k2_mbar_wait(bar + 8 + 8 * buf, (c.nbuf == 2 ? (ir >> 1) : ir) & 1);
    const double* __restrict__ Wb = W_s + (size_t)buf * wround;
    // units (band, batch of the round), longest band first, handed out through a shared counter
    for (;;) {
      int u = 0;
      if (tt == 0) u = atomicAdd(next_unit + ir, 1);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= n_units) break;
      const int bi = u & (NB - 1);
      if (bi >= nbr) continue;
      const int b = order_s[u >> c.nb_shift];
      const int* __restrict__ hdr = reinterpret_cast<const int*>(Wb + (size_t)bi * blkA);
      const int jbv = hdr[b];
      const int rend = hdr[n_b + b];
      const double* __restrict__ F = Fs_s + tt * ldF + jbv;
      const double2* __restrict__ w2 =
          reinterpret_cast<const double2*>(Wb + (size_t)bi * blkA + p.hdrA + (size_t)woff_s[b] * 4);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
      for (int r = 0; r < rend; ++r) {
        const double f = F[r];
        const double2 wa = w2[2 * r], wb = w2[2 * r + 1];
        a0 = fma(wa.x, f, a0);
        a1 = fma(wa.y, f, a1);
        a2 = fma(wb.x, f, a2);
        a3 = fma(wb.y, f, a3);
      }
      if (ir > 0 && c.bulk_store) k2_mbar_wait(bar + 24, (ir - 1) & 1);
      double* o = O_s + ((size_t)(bi * 4) * TT + tt) * c.opitch + b;
      o[0] = a0;
      o[(size_t)TT * c.opitch] = a1;
      o[(size_t)2 * TT * c.opitch] = a2;
      o[(size_t)3 * TT * c.opitch] = a3;
    }
    __syncthreads();
    const int nz = min(4 * nbr, p.n_z - 4 * zbr);          // slices of this round
    if (c.bulk_store) {
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        for (int zl = 0; zl < nz; ++zl) {
          double* dst = p.out + ((int64_t)(4 * zbr + zl) * p.T_total + t0) * pitch;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst),
                       "r"(k2_smem_u32(O_s + (size_t)zl * TT * c.opitch)), "r"((uint32_t)(nt * pitch * 8))
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    } else {
      const int per_z = nt * pitch;
      const int d_tt = K2T_THREADS / pitch, d_col = K2T_THREADS - d_tt * pitch;
      for (int zl = 0; zl < nz; ++zl) {
        int t2 = threadIdx.x / pitch, col = threadIdx.x - t2 * pitch;
        double* __restrict__ dst = p.out + ((int64_t)(4 * zbr + zl) * p.T_total) * pitch;
        for (int i = threadIdx.x; i < per_z; i += K2T_THREADS) {
          const int64_t row = p.row_map ? __ldg(p.row_map + t0 + t2) : t0 + t2;
          dst[row * pitch + col] = O_s[(zl * TT + t2) * c.opitch + col];
          t2 += d_tt;
          col += d_col;
          if (col >= pitch) {
            col -= pitch;
            ++t2;
          }
        }
      }
      __syncthreads();
    }
  }
  if (c.bulk_store && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}

bool k2_sw_plan(const K2Problem& p, K2SwPlan* c) {
  constexpr int TT = 32;
  if (!p.sw_ok || p.wstrideA < 1) return false;
  const bool bulk = !p.row_map && (p.out_pitch % 2 == 0) && (reinterpret_cast<uintptr_t>(p.out) % 16 == 0);
  const int opitch = bulk ? p.out_pitch : p.out_pitch + 1;
  const size_t f = (size_t)TT * p.ldF * 8;
  const size_t off_w = (f + 15) & ~(size_t)15;
  for (int pass = 0; pass < 4; ++pass) {
    const int nb = pass == 0 ? 4 : (pass == 1 ? 2 : 1), sh = pass == 0 ? 2 : (pass == 1 ? 1 : 0);
    const int nbuf = pass == 3 ? 1 : 2;
    const size_t off_o = off_w + (size_t)nbuf * nb * p.blkA * 8;
    const size_t off_bar = (off_o + (size_t)nb * 4 * TT * opitch * 8 + 15) & ~(size_t)15;
    const size_t total = off_bar + 32 + K2_SW_MAX_ROUNDS * 4 + 2 * (size_t)p.filt.n_b * 4;
    if (total > 225 * 1024) continue;
    c->nbuf = nbuf;
    c->nb = nb;
    c->nb_shift = sh;
    c->opitch = opitch;
    c->bulk_store = bulk ? 1 : 0;
    c->off_w = (int32_t)off_w;
    c->off_o = (int32_t)off_o;
    c->off_bar = (int32_t)off_bar;
    c->smem_bytes = (int32_t)total;
    const int64_t tiles = (p.T + TT - 1) / TT;
    int64_t zbc = ((int64_t)p.n_zb * tiles) / (8 * 148);
    if (zbc > K2_SW_MAX_ROUNDS) zbc = K2_SW_MAX_ROUNDS;      // rounds per CTA <= zbc / nb <= K2_SW_MAX_ROUNDS
    zbc = zbc / nb * nb;
    if (zbc < nb) zbc = nb;
    c->zbc = (int32_t)zbc;
    return true;
  }
  return false;
}

// ---- tensor-core contraction -----------------------------------------------------------------------
// D[32 templates x 8 slices] += A[32 x 4 samples] * B[4 samples x 8 slices] with FP64 mma.sync.m8n8k4: per k-block
// one 256-byte fragment of weights and one per 8 templates feed 256 multiply-adds each — 0.31-0.375 shared-memory
// wavefronts per 32 multiply-adds, against 1.0 in the shared-window kernel and 3.0 in the tiled one.  A slice's
// weights are zero outside its own window and the k-blocks sit at absolute multiples of 4 samples, so which
// products share an mma depends on the sample index only: a slice's value is the same bit for bit whichever
// slices share its batch.  Row pitch = 4 (mod 8) makes the 8 x 4 template fragments conflict-free.
// Work units = (band, half of the 32 templates), longest band first through a shared counter; weights of the next
// batch arrive by a bulk copy during this one; the 8 result slices leave by bulk stores that drain during the next.
struct K2MmaPlan {
  int32_t nbuf;        // weight buffers (2, or 1 when two would not fit)
  int32_t zbc;         // batches per CTA
  int32_t opitch, bulk_store;
  int32_t off_w, off_o, off_bar, smem_bytes;
};

__device__ __forceinline__ void k2_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(K2T_THREADS, 1) k2_contract_mma_kernel(K2Problem p, K2MmaPlan c) {
  constexpr int TT = 32;
  extern __shared__ __align__(16) unsigned char k2_smem[];
  double* Fs_s = reinterpret_cast<double*>(k2_smem);                        // [TT][ldF8]
  double* W_s = reinterpret_cast<double*>(k2_smem + c.off_w);               // [nbuf][header | wstride8 x 8]
  double* O_s = reinterpret_cast<double*>(k2_smem + c.off_o);               // [8][TT][opitch]
  const uint32_t bar = k2_smem_u32(k2_smem + c.off_bar);
  int* next_unit = reinterpret_cast<int*>(k2_smem + c.off_bar + 32);        // [K2_SW_MAX_ROUNDS]
  int* order_s = next_unit + K2_SW_MAX_ROUNDS;                              // [n_b] bands, longest first
  int* woff_s = order_s + p.filt.n_b;                                       // [n_b] first slot of each band
  const int n_b = p.filt.n_b, ldF = p.ldF8, pitch = p.out_pitch;
  const int64_t blk = p.blk8;
  const int lane = threadIdx.x & 31, grp = lane >> 2, tig = lane & 3;
  const int64_t t0 = (int64_t)blockIdx.y * TT;
  const int nt = (int)min((int64_t)TT, p.T - t0);
  const int zb0 = blockIdx.x * c.zbc;
  const int zb1 = min(zb0 + c.zbc, p.n_zb8);
  const int n_units = 2 * n_b;
  const uint32_t wbytes = (uint32_t)blk * 8u;
  if (threadIdx.x == 0) {
    k2_mbar_init(bar, 1);
    k2_mbar_init(bar + 8, 1);
    k2_mbar_init(bar + 16, 1);
    k2_mbar_init(bar + 24, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < K2_SW_MAX_ROUNDS) next_unit[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < n_b; i += K2T_THREADS) {
    order_s[i] = __ldg(p.band_order + i);
    woff_s[i] = __ldg(p.woff8 + i);
  }
  for (int i = threadIdx.x; i < 8 * TT * c.opitch; i += K2T_THREADS) O_s[i] = 0.0;   // pad columns stay zero
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t fbytes = (uint32_t)TT * (uint32_t)ldF * 8u;
    k2_mbar_expect_tx(bar, fbytes);
    k2_bulk_g2s(k2_smem_u32(Fs_s), p.Fs8 + t0 * ldF, fbytes, bar);
    k2_mbar_expect_tx(bar + 8, wbytes);
    k2_bulk_g2s(k2_smem_u32(W_s), p.WA8 + (int64_t)zb0 * blk, wbytes, bar + 8);
  }
  k2_mbar_wait(bar, 0);
  int ir = 0;
  for (int zb = zb0; zb < zb1; ++zb, ++ir) {
    const int buf = (c.nbuf == 2) ? (ir & 1) : 0;
    if (threadIdx.x == 0) {
      if (c.nbuf == 2 && zb + 1 < zb1) {
        k2_mbar_expect_tx(bar + 8 + 8 * (buf ^ 1), wbytes);
        k2_bulk_g2s(k2_smem_u32(W_s + (size_t)(buf ^ 1) * blk), p.WA8 + (int64_t)(zb + 1) * blk, wbytes,
                    bar + 8 + 8 * (buf ^ 1));
      }
      if (c.nbuf == 1 && ir > 0) {
        k2_mbar_expect_tx(bar + 8, wbytes);
        k2_bulk_g2s(k2_smem_u32(W_s), p.WA8 + (int64_t)zb * blk, wbytes, bar + 8);
      }
      if (ir > 0 && c.bulk_store) {      // the previous batch's result tile has left shared memory: release it
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar + 24) : "memory");
      }
    }
    k2_mbar_wait(bar + 8 + 8 * buf, (c.nbuf == 2 ? (ir >> 1) : ir) & 1);
    const int* __restrict__ hdr = reinterpret_cast<const int*>(W_s + (size_t)buf * blk);
    const double* __restrict__ Wb = W_s + (size_t)buf * blk + p.hdr8;
    for (;;) {
      int u = 0;
      if (lane == 0) u = atomicAdd(next_unit + ir, 1);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= n_units) break;
      const int b = order_s[u >> 1];
      const int half = u & 1;
      const int jbv = hdr[b];
      const int nkb = hdr[n_b + b] >> 2;
      // A fragments: template row 16 half + grp (+ 8), samples jbv + 4 kb + tig; B fragment: sample tig, slice grp
      const double* __restrict__ A0 = Fs_s + (16 * half + grp) * ldF + jbv + tig;
      const double* __restrict__ A1 = A0 + 8 * ldF;
      const double* __restrict__ Bp = Wb + (size_t)woff_s[b] * 8 + tig * 8 + grp;
      double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 2
      for (int kb = 0; kb < nkb; ++kb) {
        const double a0 = A0[4 * kb], a1 = A1[4 * kb], bw = Bp[32 * kb];
        k2_dmma(c00, c01, a0, bw);
        k2_dmma(c10, c11, a1, bw);
      }
      if (ir > 0 && c.bulk_store) k2_mbar_wait(bar + 24, (ir - 1) & 1);
      // D fragment: template row grp (+ 8), slices 2 tig and 2 tig + 1
      double* o = O_s + ((size_t)(2 * tig) * TT + 16 * half + grp) * c.opitch + b;
      o[0] = c00;
      o[(size_t)TT * c.opitch] = c01;
      o[(size_t)8 * c.opitch] = c10;
      o[(size_t)(TT + 8) * c.opitch] = c11;
    }
    __syncthreads();
    const int nz = min(8, p.n_z - 8 * zb);
    if (c.bulk_store) {
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        for (int zl = 0; zl < nz; ++zl) {
          double* dst = p.out + ((int64_t)(8 * zb + zl) * p.T_total + t0) * pitch;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst),
                       "r"(k2_smem_u32(O_s + (size_t)zl * TT * c.opitch)), "r"((uint32_t)(nt * pitch * 8))
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    } else {
      const int per_z = nt * pitch;
      const int d_tt = K2T_THREADS / pitch, d_col = K2T_THREADS - d_tt * pitch;
      for (int zl = 0; zl < nz; ++zl) {
        int t2 = threadIdx.x / pitch, col = threadIdx.x - t2 * pitch;
        double* __restrict__ dst = p.out + ((int64_t)(8 * zb + zl) * p.T_total) * pitch;
        for (int i = threadIdx.x; i < per_z; i += K2T_THREADS) {
          const int64_t row = p.row_map ? __ldg(p.row_map + t0 + t2) : t0 + t2;
          dst[row * pitch + col] = O_s[(zl * TT + t2) * c.opitch + col];
          t2 += d_tt;
          col += d_col;
          if (col >= pitch) {
            col -= pitch;
            ++t2;
          }
        }
      }
      __syncthreads();
    }
  }
  if (c.bulk_store && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}

bool k2_mma_plan(const K2Problem& p, K2MmaPlan* c) {
  constexpr int TT = 32;
  if (!p.mma_ok || p.wstride8 < 4 || (p.ldF8 & 7) != 4) return false;
  const bool bulk = !p.row_map && (p.out_pitch % 2 == 0) && (reinterpret_cast<uintptr_t>(p.out) % 16 == 0);
  const int opitch = bulk ? p.out_pitch : p.out_pitch + 1;
  const size_t f = (size_t)TT * p.ldF8 * 8;
  const size_t off_w = (f + 15) & ~(size_t)15;
  for (int nbuf = 2; nbuf >= 1; --nbuf) {
    const size_t off_o = off_w + (size_t)nbuf * p.blk8 * 8;
    const size_t off_bar = (off_o + (size_t)8 * TT * opitch * 8 + 15) & ~(size_t)15;
    const size_t total = off_bar + 32 + K2_SW_MAX_ROUNDS * 4 + 2 * (size_t)p.filt.n_b * 4;
    if (total > 225 * 1024) continue;
    c->nbuf = nbuf;
    c->opitch = opitch;
    c->bulk_store = bulk ? 1 : 0;
    c->off_w = (int32_t)off_w;
    c->off_o = (int32_t)off_o;
    c->off_bar = (int32_t)off_bar;
    c->smem_bytes = (int32_t)total;
    const int64_t tiles = (p.T + TT - 1) / TT;
    int64_t zbc = ((int64_t)p.n_zb8 * tiles) / (8 * 148);
    if (zbc > K2_SW_MAX_ROUNDS) zbc = K2_SW_MAX_ROUNDS;
    if (zbc < 1) zbc = 1;
    c->zbc = (int32_t)zbc;
    return true;
  }
  return false;
}

bool k2_tile_plan(const K2Problem& p, K2TilePlan* c) {
  if ((p.wstride & 1) || p.wstride < 2) return false;              // bulk copies move 16-byte multiples
  // a dense tile for the bulk store (two-way bank conflicts on the 18 x 32 result writes are cheaper than a
  // store loop); the padded, conflict-free pitch only when rows are remapped
  const bool bulk = !p.row_map && (p.out_pitch % 2 == 0) && (reinterpret_cast<uintptr_t>(p.out) % 16 == 0);
  const int opitch = bulk ? p.out_pitch : p.out_pitch + 1;
  for (int tt = 32; tt >= 2; tt >>= 1) {
    for (int zb = 4; zb >= 1; zb >>= 1) {
      const size_t f = (size_t)tt * p.ldF * 8;
      const size_t off_w = (f + 15) & ~(size_t)15;
      const size_t off_o = off_w + 2 * (size_t)zb * p.wstride * 8;
      const size_t off_bar = (off_o + (size_t)zb * tt * opitch * 8 + 15) & ~(size_t)15;
      const size_t total = off_bar + 32;
      if (total > 200 * 1024) continue;
      c->tt = tt;
      c->zb = zb;
      c->bulk_store = bulk ? 1 : 0;
      c->opitch = opitch;
      c->off_w = (int32_t)off_w;
      c->off_o = (int32_t)off_o;
      c->off_bar = (int32_t)off_bar;
      c->smem_bytes = (int32_t)total;
      const int64_t tiles = (p.T + tt - 1) / tt;
      // enough CTAs for ~4 waves of 148 SMs, at most 64 slices per CTA, whole batches
      int64_t zc = (p.n_z * tiles) / (4 * 148);
      if (zc > 64) zc = 64;
      zc = (zc / zb) * zb;
      if (zc < zb) zc = zb;
      c->zc = (int32_t)zc;
      return true;
    }
  }
  return false;
}

}  // namespace

cudaError_t launch_k2_permute(const double* Fnu, const int32_t* perm, double* Fs, int64_t T, int32_t n_lam,
                              int32_t ldF, int32_t* nonfinite, cudaStream_t stream) {
  const int64_t total = T * (int64_t)ldF;
  if (total == 0) return cudaSuccess;
  k2_permute_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(Fnu, perm, Fs, T, n_lam, ldF, nonfinite);
  return cudaGetLastError();
}

cudaError_t launch_k2_ranges(const K2Problem& p, cudaStream_t stream) {
  const int total = p.n_z * p.filt.n_b;
  if (total == 0) return cudaSuccess;
  k2_ranges_kernel<<<(total + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_k2_weights(const K2Problem& p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.n_z * p.wstride;
  if (total == 0) return cudaSuccess;
  k2_weights_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_k2_batch_ranges(const K2Problem& p, cudaStream_t stream) {
  const int total = p.n_zb * p.filt.n_b;
  if (total == 0) return cudaSuccess;
  k2_batch_ranges_kernel<<<(total + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_k2_weights_aligned(const K2Problem& p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.n_zb * p.wstrideA * 4;
  if (total == 0) return cudaSuccess;
  k2_weights_aligned_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(p);
  const int nh = p.n_zb * p.filt.n_b;
  k2_headersA_kernel<<<(nh + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_k2_batch_ranges8(const K2Problem& p, cudaStream_t stream) {
  const int total = p.n_zb8 * p.filt.n_b;
  if (total == 0) return cudaSuccess;
  k2_batch_ranges8_kernel<<<(total + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_k2_weights_aligned8(const K2Problem& p, cudaStream_t stream) {
  const int64_t total = (int64_t)p.n_zb8 * p.wstride8 * 8;
  if (total == 0) return cudaSuccess;
  k2_weights_aligned8_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(p);
  const int nh = p.n_zb8 * p.filt.n_b;
  k2_headers8_kernel<<<(nh + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

bool k2_mma_fits(const K2Problem& p) {
  K2MmaPlan c;
  return k2_mma_plan(p, &c);
}

bool k2_sw_fits(const K2Problem& p) {
  K2SwPlan c;
  return k2_sw_plan(p, &c);
}

cudaError_t launch_k2_contract(const K2Problem& p, cudaStream_t stream) {
  if (p.n_z == 0 || p.T == 0) return cudaSuccess;
  K2MmaPlan mc;
  if (p.tile_ok && k2_mma_plan(p, &mc)) {
    cudaError_t e = cudaFuncSetAttribute(k2_contract_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         mc.smem_bytes);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((p.n_zb8 + mc.zbc - 1) / mc.zbc), (unsigned)((p.T + 31) / 32));
    k2_contract_mma_kernel<<<grid, K2T_THREADS, (size_t)mc.smem_bytes, stream>>>(p, mc);
    return cudaGetLastError();
  }
  K2SwPlan sc;
  if (p.tile_ok && k2_sw_plan(p, &sc)) {
    cudaError_t e = cudaFuncSetAttribute(k2_contract_sw_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         sc.smem_bytes);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((p.n_zb + sc.zbc - 1) / sc.zbc), (unsigned)((p.T + 31) / 32));
    k2_contract_sw_kernel<32><<<grid, K2T_THREADS, (size_t)sc.smem_bytes, stream>>>(p, sc);
    return cudaGetLastError();
  }
  K2TilePlan c;
  if (p.tile_ok && k2_tile_plan(p, &c)) {
    dim3 grid((unsigned)((p.n_z + c.zc - 1) / c.zc), (unsigned)((p.T + c.tt - 1) / c.tt));
    auto go = [&](auto kern) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes);
      if (e != cudaSuccess) return e;
      kern<<<grid, K2T_THREADS, (size_t)c.smem_bytes, stream>>>(p, c);
      return cudaGetLastError();
    };
    switch (c.tt) {
      case 32: return go(k2_contract_tile_kernel<32>);
      case 16: return go(k2_contract_tile_kernel<16>);
      case 8: return go(k2_contract_tile_kernel<8>);
      case 4: return go(k2_contract_tile_kernel<4>);
      default: return go(k2_contract_tile_kernel<2>);
    }
  }
  const size_t idx_bytes = 2 * (size_t)p.filt.n_b * sizeof(int32_t);
  const size_t smem = (size_t)p.wstride * sizeof(double) + idx_bytes;
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
