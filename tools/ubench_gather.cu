// Microbenchmark behind the K1 design: how fast can one SM gather random 144-byte template rows
//   A  per-thread cp.async.bulk (TMA unit) into a private shared-memory row, read back thread-per-row (LDS.128)
//   B  cooperative LDG.128, lane per 16-byte chunk of the warp's 32-row tile (what k1_coop_kernel does today)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_gather ubench_gather.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ double2 lds_v2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

constexpr int NB = 18, ROWB = NB * 8;     // 144-byte rows
constexpr int CPR = NB / 2;

struct Tabs {
  const double* ga; const double* sb; const double* to; const double* bb;
  uint32_t n_ga, n_sb, n_to, n_bb, n_z;
};

__device__ __forceinline__ void pick_rows(const Tabs& T, uint32_t i, uint32_t rows_per_z, const double*& a, const double*& b,
                                          const double*& c, const double*& d) {
  const uint32_t z = min(i / rows_per_z, T.n_z - 1);
  const uint32_t h = hash(i);
  a = T.ga + ((size_t)z * T.n_ga + h % T.n_ga) * NB;
  b = T.sb + ((size_t)z * T.n_sb + (h >> 8) % T.n_sb) * NB;
  c = T.to + ((size_t)z * T.n_to + (h >> 16) % T.n_to) * NB;
  d = T.bb + ((size_t)z * T.n_bb + (h >> 24) % T.n_bb) * NB;
}

// ---- A: per-thread bulk copies, NT threads per CTA, STAGES tiles in flight per warp -------------------------------
template <int NT, int NFAM, int ST>
__global__ void __launch_bounds__(NT) gatherA(Tabs T, uint32_t n, uint32_t rows_per_z, double* out, int tiles_per_warp) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar0 = smem_u32(smem);                         // 2 barriers per warp
  const uint32_t tile_bytes = 32 * NFAM * ROWB;
  const uint32_t buf0 = smem_u32(smem) + 256 + warp * ST * tile_bytes;
  if (threadIdx.x < NW * 2) mbar_init(bar0 + 8 * threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  const uint32_t gw = blockIdx.x * NW + warp;                   // global warp id
  auto issue = [&](int t, int st) {
    const uint32_t i = (gw * tiles_per_warp + t) * 32 + lane;
    const uint32_t bar = bar0 + 8 * (warp * 2 + st);
    if (lane == 0) mbar_expect_tx(bar, tile_bytes);
    __syncwarp();
    const double *a, *b, *c, *d;
    pick_rows(T, i < n ? i : 0, rows_per_z, a, b, c, d);
    const uint32_t dst = buf0 + st * tile_bytes + lane * (NFAM * ROWB);
    bulk_g2s(dst, a, ROWB, bar);
    if (NFAM > 1) bulk_g2s(dst + ROWB, b, ROWB, bar);
    if (NFAM > 2) bulk_g2s(dst + 2 * ROWB, c, ROWB, bar);
    if (NFAM > 3) bulk_g2s(dst + 3 * ROWB, d, ROWB, bar);
  };
  issue(0, 0);
  for (int t = 0; t < tiles_per_warp; ++t) {
    const int st = ST == 2 ? (t & 1) : 0;
    if (ST == 2 && t + 1 < tiles_per_warp) issue(t + 1, st ^ 1);
    mbar_wait(bar0 + 8 * (warp * 2 + st), ST == 2 ? ((t >> 1) & 1) : (t & 1));
    const uint32_t src = buf0 + st * tile_bytes + lane * (NFAM * ROWB);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
      double m0 = 0.0, m1 = 0.0;
#pragma unroll
      for (int f = 0; f < NFAM; ++f) {
        const double2 v = lds_v2(src + f * ROWB + k * 16);
        m0 = fma(v.x, 1.0 + f, m0);
        m1 = fma(v.y, 1.0 + f, m1);
      }
      acc += m0 * m0 + m1 * m1;
    }
    const uint32_t i = (gw * tiles_per_warp + t) * 32 + lane;
    if (i < n) out[i] = acc;
    __syncwarp();
    if (ST == 1 && t + 1 < tiles_per_warp) issue(t + 1, 0);
  }
}

// ---- B: cooperative LDG.128 ------------------------------------------------------------------------------------
template <int NFAM>
__global__ void __launch_bounds__(256, 3) gatherB(Tabs T, uint32_t n, uint32_t rows_per_z, double* out) {
  __shared__ uint32_t offs[8][32][4];
  __shared__ double part[8][32][CPR + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t row0 = (blockIdx.x * 8 + warp) * 32;
  {
    const uint32_t i = row0 + lane;
    const double *a, *b, *c, *d;
    pick_rows(T, i < n ? i : 0, rows_per_z, a, b, c, d);
    offs[warp][lane][0] = (uint32_t)((a - T.ga) / 2);
    offs[warp][lane][1] = (uint32_t)((b - T.sb) / 2);
    offs[warp][lane][2] = (uint32_t)((c - T.to) / 2);
    offs[warp][lane][3] = (uint32_t)((d - T.bb) / 2);
  }
  __syncwarp();
  const double2* tg = reinterpret_cast<const double2*>(T.ga);
  const double2* ts = reinterpret_cast<const double2*>(T.sb);
  const double2* tt = reinterpret_cast<const double2*>(T.to);
  const double2* tb = reinterpret_cast<const double2*>(T.bb);
  int e = lane / CPR, k = lane - e * CPR;
  const int de = 32 / CPR, dk = 32 - de * CPR;
#pragma unroll
  for (int it = 0; it < CPR; ++it) {
    const uint4 g = *reinterpret_cast<const uint4*>(offs[warp][e]);
    double m0 = 0.0, m1 = 0.0;
    { const double2 v = __ldg(tg + g.x + k); m0 = fma(v.x, 1.0, m0); m1 = fma(v.y, 1.0, m1); }
    if (NFAM > 1) { const double2 v = __ldg(ts + g.y + k); m0 = fma(v.x, 2.0, m0); m1 = fma(v.y, 2.0, m1); }
    if (NFAM > 2) { const double2 v = __ldg(tt + g.z + k); m0 = fma(v.x, 3.0, m0); m1 = fma(v.y, 3.0, m1); }
    if (NFAM > 3) { const double2 v = __ldg(tb + g.w + k); m0 = fma(v.x, 4.0, m0); m1 = fma(v.y, 4.0, m1); }
    part[warp][e][k] = m0 * m0 + m1 * m1;
    k += dk; e += de;
    if (k >= CPR) { k -= CPR; ++e; }
  }
  __syncwarp();
  double acc = 0.0;
#pragma unroll
  for (int kk = 0; kk < CPR; ++kk) acc += part[warp][lane][kk];
  if (row0 + lane < n) out[row0 + lane] = acc;
}

// ---- C: what a thread-per-row K1 would do: GA/SB/TO rows by per-thread bulk copies into padded slots
//         (slot pitch 3*144+16 bytes: quarter-warps read 8 distinct 16-byte bank groups), BB row by LDG.128 (2 rows
//         per z: warp-uniform addresses), thread-per-row readback, NT threads per CTA, one tile per warp per round
template <int NT, int NTMA>
__global__ void __launch_bounds__(NT) gatherC(Tabs T, uint32_t n, uint32_t rows_per_z, double* out, int tiles_per_warp) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NW = NT / 32;
  constexpr uint32_t SLOT = NTMA * ROWB + 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar0 = smem_u32(smem);
  const uint32_t buf0 = smem_u32(smem) + 256 + warp * 32 * SLOT;
  if (threadIdx.x < NW) mbar_init(bar0 + 8 * threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  const uint32_t gw = blockIdx.x * NW + warp;
  const uint32_t bar = bar0 + 8 * warp;
  for (int t = 0; t < tiles_per_warp; ++t) {
    const uint32_t i = (gw * tiles_per_warp + t) * 32 + lane;
    if (lane == 0) mbar_expect_tx(bar, 32 * NTMA * ROWB);
    __syncwarp();
    const double *a, *b, *c, *d;
    pick_rows(T, i < n ? i : 0, rows_per_z, a, b, c, d);
    const uint32_t dst = buf0 + lane * SLOT;
    bulk_g2s(dst, a, ROWB, bar);
    if (NTMA > 1) bulk_g2s(dst + ROWB, b, ROWB, bar);
    if (NTMA > 2) bulk_g2s(dst + 2 * ROWB, c, ROWB, bar);
    const double2* d2 = reinterpret_cast<const double2*>(d);
    const double2* c2 = reinterpret_cast<const double2*>(c);
    const double2* b2 = reinterpret_cast<const double2*>(b);
    mbar_wait(bar, t & 1);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
      double m0 = 0.0, m1 = 0.0;
      { const double2 v = lds_v2(dst + k * 16); m0 = fma(v.x, 1.0, m0); m1 = fma(v.y, 1.0, m1); }
      { const double2 v = NTMA > 1 ? lds_v2(dst + ROWB + k * 16) : __ldg(b2 + k); m0 = fma(v.x, 2.0, m0); m1 = fma(v.y, 2.0, m1); }
      { const double2 v = NTMA > 2 ? lds_v2(dst + 2 * ROWB + k * 16) : __ldg(c2 + k); m0 = fma(v.x, 3.0, m0); m1 = fma(v.y, 3.0, m1); }
      { const double2 v = __ldg(d2 + k); m0 = fma(v.x, 4.0, m0); m1 = fma(v.y, 4.0, m1); }
      acc += m0 * m0 + m1 * m1;
    }
    if (i < n) out[i] = acc;
    __syncwarp();
  }
}

int main(int argc, char** argv) {
  const uint32_t n = 10000000u, n_z = argc > 1 ? atoi(argv[1]) : 600;
  Tabs T{};
  T.n_ga = 11000; T.n_sb = 161; T.n_to = 60; T.n_bb = 2; T.n_z = n_z;
  const uint32_t rows_per_z = (n + n_z - 1) / n_z;
  double *ga, *sb, *to, *bb, *out;
  CK(cudaMalloc(&ga, (size_t)n_z * T.n_ga * ROWB));
  CK(cudaMalloc(&sb, (size_t)n_z * T.n_sb * ROWB));
  CK(cudaMalloc(&to, (size_t)n_z * T.n_to * ROWB));
  CK(cudaMalloc(&bb, (size_t)n_z * T.n_bb * ROWB));
  CK(cudaMalloc(&out, (size_t)n * 8));
  CK(cudaMemset(ga, 0, (size_t)n_z * T.n_ga * ROWB));
  CK(cudaMemset(sb, 0, (size_t)n_z * T.n_sb * ROWB));
  CK(cudaMemset(to, 0, (size_t)n_z * T.n_to * ROWB));
  CK(cudaMemset(bb, 0, (size_t)n_z * T.n_bb * ROWB));
  T.ga = ga; T.sb = sb; T.to = to; T.bb = bb;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto launch) {
    for (int w = 0; w < 3; ++w) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 10; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-34s %8.3f ms per 1e7 rows   %.3e rows/s\n", name, ms / 10, n / (ms / 10 * 1e-3));
  };
  auto runA = [&](auto kern, int NT, int NFAM, int ST, int tiles, const char* name) {
    const int NW = NT / 32;
    const size_t smem = 256 + (size_t)NW * ST * 32 * NFAM * ROWB;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t warps = (n + 32 * tiles - 1) / (32 * tiles);
    const uint32_t blocks = (warps + NW - 1) / NW;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    printf("[%d warps/SM] ", occ * NW);
    timeit(name, [&]() { kern<<<blocks, NT, smem>>>(T, n, rows_per_z, out, tiles); });
  };
  runA(gatherA<64, 4, 1>, 64, 4, 1, 1, "A tma 4fam NT64 st1 tiles1");
  runA(gatherA<64, 4, 1>, 64, 4, 1, 8, "A tma 4fam NT64 st1 tiles8");
  runA(gatherA<32, 4, 1>, 32, 4, 1, 8, "A tma 4fam NT32 st1 tiles8");
  runA(gatherA<32, 4, 2>, 32, 4, 2, 8, "A tma 4fam NT32 st2 tiles8");
  runA(gatherA<64, 2, 1>, 64, 2, 1, 8, "A tma 2fam NT64 st1 tiles8");
  runA(gatherA<128, 1, 2>, 128, 1, 2, 8, "A tma 1fam(GA) NT128 st2 tiles8");
  runA(gatherA<128, 1, 1>, 128, 1, 1, 8, "A tma 1fam(GA) NT128 st1 tiles8");
  {
    Tabs S = T;                       // every family small: L2-resident, isolates the copy engine's issue rate
    S.ga = T.sb; S.n_ga = T.n_sb;
    Tabs keep = T; T = S;
    runA(gatherA<128, 1, 2>, 128, 1, 2, 8, "A tma 1fam(L2) NT128 st2 tiles8");
    runA(gatherA<64, 4, 1>, 64, 4, 1, 8, "A tma 4fam(L2) NT64 st1 tiles8");
    timeit("B coop LDG.128 4fam(L2)", [&]() { gatherB<4><<<(n + 255) / 256, 256>>>(T, n, rows_per_z, out); });
    T = keep;
  }
  timeit("B coop LDG.128 4fam", [&]() { gatherB<4><<<(n + 255) / 256, 256>>>(T, n, rows_per_z, out); });
  timeit("B coop LDG.128 1fam(GA)", [&]() { gatherB<1><<<(n + 255) / 256, 256>>>(T, n, rows_per_z, out); });
  auto runC = [&](auto kern, int NT, int NTMA, int tiles, const char* name) {
    const int NW = NT / 32;
    const size_t smem = 256 + (size_t)NW * 32 * (NTMA * ROWB + 16);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t warps = (n + 32 * tiles - 1) / (32 * tiles);
    const uint32_t blocks = (warps + NW - 1) / NW;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    printf("[%d warps/SM] ", occ * NW);
    timeit(name, [&]() { kern<<<blocks, NT, smem>>>(T, n, rows_per_z, out, tiles); });
  };
  runC(gatherC<128, 3>, 128, 3, 1, "C tma GA,SB,TO + ldg BB NT128 t1");
  runC(gatherC<128, 3>, 128, 3, 4, "C tma GA,SB,TO + ldg BB NT128 t4");
  runC(gatherC<64, 3>, 64, 3, 4, "C tma GA,SB,TO + ldg BB NT64 t4");
  runC(gatherC<256, 3>, 256, 3, 1, "C tma GA,SB,TO + ldg BB NT256 t1");
  runC(gatherC<128, 1>, 128, 1, 4, "C tma GA + ldg SB,TO,BB NT128 t4");
  runC(gatherC<256, 1>, 256, 1, 4, "C tma GA + ldg SB,TO,BB NT256 t4");
  runC(gatherC<128, 2>, 128, 2, 4, "C tma GA,SB + ldg TO,BB NT128 t4");
  return 0;
}
