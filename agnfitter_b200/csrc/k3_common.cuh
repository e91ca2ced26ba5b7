// Stretch-move arithmetic shared by k3_ensemble.cu (stand-alone propose kernel) and k1_coop.cu (the
// fused propose -> ln_probab -> accept launch).  See k3_ensemble.cuh for the move and its RNG.
#pragma once
#include <cstdint>

namespace agnf {

// Philox4x32-10 (Salmon et al. 2011), counter (c0..c3), key (k0, k1)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

struct StretchDraw {
  double zz;       // stretch factor
  int j;           // partner index inside the other half
  double lnu;      // ln u3 of the acceptance test
};

// draws of walker k (of the active half) of source s at (step, half); H = walkers per half
__device__ __forceinline__ StretchDraw stretch_draw(uint64_t seed, double a, int64_t step, int half, int64_t s, int k,
                                                    int H) {
  uint32_t x[4];
  philox4x32_10((uint32_t)step, (uint32_t)half, (uint32_t)s, (uint32_t)k, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  const double u1 = (double)((((uint64_t)x[0] << 32) | x[1]) >> 11) * 1.1102230246251565e-16;   // 2^-53
  const double u3 = ((double)x[3] + 0.5) * 2.3283064365386963e-10;                            // 2^-32
  // zz = ((a - 1) u1 + 1)^2 / a, every operation rounded on its own (bit-identical to NumPy)
  const double t = __dadd_rn(__dmul_rn(a - 1.0, u1), 1.0);
  StretchDraw d;
  d.zz = __ddiv_rn(__dmul_rn(t, t), a);
  d.j = (int)(((uint64_t)x[2] * (uint64_t)H) >> 32);
  d.lnu = log(u3);
  return d;
}

// emcee 3's RedBlueMove shuffles the walkers into two equal sets at every step (`randomize_split=True`, the default):
// `split` [S, W] lists, per source, the walkers of set 0 in slots [0, W/2) and of set 1 in slots [W/2, W)
// (k3_split_kernel); nullptr = the fixed first-half / second-half split.
constexpr uint32_t K3_SPLIT_TAG = 0xFFFFFFFFu;     // second counter word of the split's Philox stream
// COHERENT: the table is rewritten inside the running kernel (persistent sampler): no read-only-cache loads
template <bool COHERENT = false>
__device__ __forceinline__ int split_walker(const uint16_t* split, int64_t s_local, int W, int slot) {
  if (!split) return slot;
  if (COHERENT) return (int)*(reinterpret_cast<const volatile uint16_t*>(split) + s_local * W + slot);
  return (int)__ldg(split + s_local * W + slot);
}

// The split ranking of one source by one warp: out[rank(w)] = w, rank = position of walker w's 32-bit Philox key among
// the source's W keys, ties by walker index.  Keys are uniform, so they are dropped into 128 buckets by their top seven
// bits (shared-memory histogram, warp scan, scatter) and a key is compared only against its own bucket (W/128
// candidates instead of W): the result is the same ranking the all-pairs count gives, for any key values.
// `scratch`: split_scratch_words(W) 32-bit words of shared memory owned by the warp.  OUT16 = uint16_t (may be global).
__host__ __device__ constexpr int split_scratch_words(int W) { return ((W + 3) & ~3) + 256 + W + (W + 1) / 2; }
__device__ __forceinline__ void split_rank_warp(uint16_t* out, uint32_t* scratch, int W, int64_t step,
                                                int64_t src_global, uint64_t seed, int lane) {
  const int W4 = (W + 3) & ~3;
  uint32_t* keys = scratch;                 // [W4]
  uint32_t* cnt = scratch + W4;             // [128] entries per bucket
  uint32_t* fill = cnt + 128;               // [128] bucket start, advanced to the bucket end by the scatter
  uint32_t* skey = fill + 128;              // [W] keys in bucket order
  uint16_t* sidx = reinterpret_cast<uint16_t*>(skey + W);      // [W] their walkers
  for (int q = lane; q < W4 / 4; q += 32) {
    uint32_t x[4];
    philox4x32_10((uint32_t)step, K3_SPLIT_TAG, (uint32_t)src_global, (uint32_t)q, (uint32_t)seed, (uint32_t)(seed >> 32), x);
    keys[4 * q] = x[0];
    keys[4 * q + 1] = x[1];
    keys[4 * q + 2] = x[2];
    keys[4 * q + 3] = x[3];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) cnt[lane + 32 * u] = 0u;
  __syncwarp();
  for (int w = lane; w < W; w += 32) atomicAdd(cnt + (keys[w] >> 25), 1u);
  __syncwarp();
  {
    // exclusive scan of the 128 counts: lane owns buckets 4 lane .. 4 lane + 3
    const uint32_t c0 = cnt[4 * lane], c1 = cnt[4 * lane + 1], c2 = cnt[4 * lane + 2], c3 = cnt[4 * lane + 3];
    uint32_t incl = c0 + c1 + c2 + c3;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t base = incl - (c0 + c1 + c2 + c3);
    fill[4 * lane] = base;
    fill[4 * lane + 1] = base + c0;
    fill[4 * lane + 2] = base + c0 + c1;
    fill[4 * lane + 3] = base + c0 + c1 + c2;
  }
  __syncwarp();
  for (int w = lane; w < W; w += 32) {
    const uint32_t k = keys[w];
    const uint32_t at = atomicAdd(fill + (k >> 25), 1u);
    skey[at] = k;
    sidx[at] = (uint16_t)w;
  }
  __syncwarp();
  for (int w = lane; w < W; w += 32) {
    const uint32_t k = keys[w], b = k >> 25;
    const uint32_t end = fill[b], n = cnt[b];
    uint32_t r = end - n;
    for (uint32_t e = end - n; e < end; ++e) {
      const uint32_t ke = skey[e];
      r += (ke < k) || (ke == k && (int)sidx[e] < w);
    }
    out[r] = (uint16_t)w;
  }
  __syncwarp();
}

// q = c - (c - x) * zz, no contraction
__device__ __forceinline__ double stretch_coord(double c, double x, double zz) {
  return __dsub_rn(c, __dmul_rn(__dsub_rn(c, x), zz));
}

}  // namespace agnf
