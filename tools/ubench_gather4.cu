// Microbenchmark behind the K1 design: random 144-byte template rows through the TMA unit's
// tile::gather4 mode (UTMALDG.2D.GATHER4: one operation fetches four rows of a 2-D tensor by row index),
// thread-per-row readback.  Checks the gathered values against plain loads.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_gather4 ubench_gather4.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spin = 0; !ok; ++spin) {
    if (spin > (1 << 22)) __trap();
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int r0, int r1, int r2,
                                        int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];\n" ::"r"(dst), "l"(tm), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
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

constexpr int NB = 18, ROWB = NB * 8, CPR = NB / 2;
constexpr int GROUP = 640;                 // 4 rows of 144 bytes, padded to the 128-byte destination alignment

struct Tabs {
  const double* t[4];
  uint32_t n[4];
  uint32_t n_z;
};
struct Maps {
  CUtensorMap m[4];
};

__device__ __forceinline__ void pick_rows(const Tabs& T, uint32_t i, uint32_t rows_per_z, int r[4]) {
  const uint32_t z = min(i / rows_per_z, T.n_z - 1);
  const uint32_t h = hash(i);
  r[0] = (int)(z * T.n[0] + h % T.n[0]);
  r[1] = (int)(z * T.n[1] + (h >> 8) % T.n[1]);
  r[2] = (int)(z * T.n[2] + (h >> 16) % T.n[2]);
  r[3] = (int)(z * T.n[3] + (h >> 24) % T.n[3]);
}

// NTMA families through gather4, the others by per-thread LDG.128; ISSUE = 0: lane 4g issues group g,
// 1: lane 0 issues all eight groups of a family
template <int NT, int NTMA, int ISSUE, int SHIFT = 0>
__global__ void __launch_bounds__(NT) gatherG(const __grid_constant__ Maps M, Tabs T, uint32_t n, uint32_t rows_per_z,
                                              double* out, int tiles_per_warp) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NW = NT / 32;
  constexpr uint32_t TILE = 8 * GROUP;                         // one family, 32 rows
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
  const uint32_t bar = base + 8 * warp;
  const uint32_t buf = base + 128 + warp * NTMA * TILE;
  if (threadIdx.x < NW) mbar_init(base + 8 * threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  const uint32_t gw = blockIdx.x * NW + warp;
  for (int t = 0; t < tiles_per_warp; ++t) {
    const uint32_t i = (gw * tiles_per_warp + t) * 32 + lane;
    int r[4];
    pick_rows(T, i < n ? i : 0, rows_per_z, r);
    if (lane == 0) mbar_expect_tx(bar, SHIFT ? 8 * NTMA * GROUP : 32 * NTMA * ROWB);
    __syncwarp();
#pragma unroll
    for (int f = 0; f < NTMA; ++f) {
      if (ISSUE == 0) {
        const int r1 = __shfl_down_sync(0xffffffffu, r[f], 1), r2 = __shfl_down_sync(0xffffffffu, r[f], 2),
                  r3 = __shfl_down_sync(0xffffffffu, r[f], 3);
        if ((lane & 3) == 0) gather4(buf + f * TILE + (lane >> 2) * GROUP, &M.m[f], bar, SHIFT && ((lane >> 2) & 1) ? -2 : 0, r[f], r1, r2, r3);
      } else {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int r0 = __shfl_sync(0xffffffffu, r[f], 4 * g), r1 = __shfl_sync(0xffffffffu, r[f], 4 * g + 1),
                    r2 = __shfl_sync(0xffffffffu, r[f], 4 * g + 2), r3 = __shfl_sync(0xffffffffu, r[f], 4 * g + 3);
          if (lane == 0) gather4(buf + f * TILE + g * GROUP, &M.m[f], bar, 0, r0, r1, r2, r3);
        }
      }
    }
    const double2* g2[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) g2[f] = reinterpret_cast<const double2*>(T.t[f] + (size_t)r[f] * NB);
    const uint32_t mine = SHIFT ? buf + (lane >> 2) * GROUP + (lane & 3) * 160 + ((lane >> 2) & 1) * 16
                                : buf + (lane >> 2) * GROUP + (lane & 3) * ROWB;
    mbar_wait(bar, t & 1);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
      double m0 = 0.0, m1 = 0.0;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const double2 v = f < NTMA ? lds_v2(mine + f * TILE + k * 16) : __ldg(g2[f] + k);
        m0 = fma(v.x, 1.0 + f, m0);
        m1 = fma(v.y, 1.0 + f, m1);
      }
      acc += m0 * m0 + m1 * m1;
    }
    if (i < n) out[i] = acc;
    __syncwarp();
  }
}

// reference: plain per-thread loads
__global__ void gatherR(Tabs T, uint32_t n, uint32_t rows_per_z, double* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r[4];
  pick_rows(T, i, rows_per_z, r);
  double acc = 0.0;
  for (int k = 0; k < CPR; ++k) {
    double m0 = 0.0, m1 = 0.0;
    for (int f = 0; f < 4; ++f) {
      const double2 v = reinterpret_cast<const double2*>(T.t[f] + (size_t)r[f] * NB)[k];
      m0 = fma(v.x, 1.0 + f, m0);
      m1 = fma(v.y, 1.0 + f, m1);
    }
    acc += m0 * m0 + m1 * m1;
  }
  out[i] = acc;
}

__global__ void fill(double* p, size_t n, uint32_t salt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (double)(hash((uint32_t)i ^ salt) & 0xffff) * (1.0 / 65536.0);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const uint32_t n = 10000000u, n_z = argc > 1 ? atoi(argv[1]) : 5991;
  const int box_rows = argc > 2 ? atoi(argv[2]) : 1;
  Tabs T{};
  T.n[0] = 11000; T.n[1] = 161; T.n[2] = 60; T.n[3] = 2; T.n_z = n_z;
  const uint32_t rows_per_z = (n + n_z - 1) / n_z;
  double* tab[4];
  double *out, *ref;
  for (int f = 0; f < 4; ++f) {
    const size_t cnt = (size_t)n_z * T.n[f] * NB;
    CK(cudaMalloc(&tab[f], cnt * 8));
    fill<<<(unsigned)((cnt + 255) / 256), 256>>>(tab[f], cnt, 0x1234u * (f + 1));
    T.t[f] = tab[f];
  }
  CK(cudaMalloc(&out, (size_t)n * 8));
  CK(cudaMalloc(&ref, (size_t)n * 8));
  CK(cudaDeviceSynchronize());
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode || qres != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  Maps M, M20;
  for (int f = 0; f < 4; ++f) {
    cuuint64_t dims[2] = {(cuuint64_t)NB, (cuuint64_t)n_z * T.n[f]};
    cuuint64_t strides[1] = {(cuuint64_t)ROWB};
    cuuint32_t box[2] = {(cuuint32_t)NB, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&M.m[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, tab[f], dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled family %d: error %d\n", f, (int)r); return 1; }
    cuuint32_t box20[2] = {(cuuint32_t)NB + 2, 1};
    r = encode(&M20.m[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, tab[f], dims, strides, box20, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled (box 20) family %d: error %d\n", f, (int)r); return 1; }
  }
  gatherR<<<(n + 255) / 256, 256>>>(T, n, rows_per_z, ref);
  CK(cudaDeviceSynchronize());
  double* h_ref = (double*)malloc((size_t)n * 8);
  double* h_out = (double*)malloc((size_t)n * 8);
  CK(cudaMemcpy(h_ref, ref, (size_t)n * 8, cudaMemcpyDeviceToHost));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto launch) {
    CK(cudaMemset(out, 0, (size_t)n * 8));
    for (int w = 0; w < 3; ++w) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 10; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(h_out, out, (size_t)n * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) bad += h_out[i] != h_ref[i];
    printf("%-40s %8.3f ms per 1e7 rows   %.3e rows/s   mismatches %zu\n", name, ms / 10, n / (ms / 10 * 1e-3), bad);
  };
  auto runG = [&](auto kern, int NT, int NTMA, int tiles, const char* name, bool shift = false) {
    const int NW = NT / 32;
    const size_t smem = 128 + 128 + (size_t)NW * NTMA * 8 * GROUP;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t warps = (n + 32 * tiles - 1) / (32 * tiles);
    const uint32_t blocks = (warps + NW - 1) / NW;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    printf("[%2d warps/SM] ", occ * NW);
    timeit(name, [&]() { kern<<<blocks, NT, smem>>>(shift ? M20 : M, T, n, rows_per_z, out, tiles); });
  };
  printf("box rows %d\n", box_rows);
  runG(gatherG<128, 1, 0>, 128, 1, 4, "G gather4 GA | ldg SB,TO,BB NT128 t4");
  runG(gatherG<128, 2, 0>, 128, 2, 4, "G gather4 GA,SB | ldg TO,BB NT128 t4");
  runG(gatherG<128, 3, 0>, 128, 3, 1, "G gather4 GA,SB,TO | ldg BB NT128 t1");
  runG(gatherG<128, 3, 0>, 128, 3, 4, "G gather4 GA,SB,TO | ldg BB NT128 t4");
  runG(gatherG<64, 3, 0>, 64, 3, 4, "G gather4 GA,SB,TO | ldg BB NT64 t4");
  runG(gatherG<256, 3, 0>, 256, 3, 1, "G gather4 GA,SB,TO | ldg BB NT256 t1");
  runG(gatherG<128, 3, 1>, 128, 3, 4, "G gather4 GA,SB,TO lane0 issues NT128 t4");
  runG(gatherG<128, 4, 0>, 128, 4, 4, "G gather4 all four NT128 t4");
  runG(gatherG<64, 4, 0>, 64, 4, 4, "G gather4 all four NT64 t4");
  runG(gatherG<256, 4, 0>, 256, 4, 1, "G gather4 all four NT256 t1");
  runG(gatherG<64, 3, 0, 1>, 64, 3, 4, "S shifted odd groups GA,SB,TO NT64 t4", true);
  runG(gatherG<128, 3, 0, 1>, 128, 3, 4, "S shifted odd groups GA,SB,TO NT128 t4", true);
  runG(gatherG<128, 3, 0, 1>, 128, 3, 16, "S shifted odd groups GA,SB,TO NT128 t16", true);
  runG(gatherG<64, 4, 0, 1>, 64, 4, 4, "S shifted odd groups all four NT64 t4", true);
  return 0;
}
