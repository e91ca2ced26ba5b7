// PTX wrappers (shared-window loads/stores, mbarrier, bulk copies) and the cell-table snap shared by the
// cooperative kernel (k1_coop.cu) and the gather4 kernel (k1_g4.cu).
#pragma once
#include "k1_lnprob.cuh"
#include "k1_common.cuh"

namespace agnf {

// ---- shared memory through 32-bit shared-window addresses (no generic -> shared conversions) ------
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_v4s32(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int2 lds_v2s32(uint32_t a) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32_ro(uint32_t a) {          // see lds_f64_ro
  int v;
  asm("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u16_ro(uint32_t a) {
  unsigned short v;
  asm("ld.shared.u16 %0, [%1];\n" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory");
}
// Store without the compiler-level memory barrier: for values nobody reads before the next __syncwarp() (the band
// phase's partial sums), so that the loads of the following pass may be scheduled above it.
__device__ __forceinline__ void sts_f64_nb(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v));
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}

// ---- mbarrier + bulk copy (TMA unit, 1-D) -------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1 << 24)) __trap();      // a lost copy must not hang the device
  }
}

// snap through the cell table; see k1_common.cuh::snap_dim_lut for the argument (same steps, shared
// window addresses, task constants from the kernel parameter bank)
__device__ __forceinline__ int snap_task(const SnapTask& T, double p, uint32_t ext_s, uint32_t eperm_s, uint32_t lut_s,
                                         const double* ext_g, const int32_t* eperm_g) {
  if (T.lut_mmax >= 0) {
    const int c = min(max(__double2int_rz((p - T.lo) * T.inv), 0), T.lut_mmax);
    int l = (int)lds_u16_ro(lut_s + 2u * (unsigned)(T.lut_off + c));
    const uint32_t base = ext_s + 8u * (unsigned)T.ext_off;
    if (lds_f64_ro(base + 8u * l) > p) --l;
    if (lds_f64_ro(base + 8u * l + 8u) <= p) ++l;
    if (lds_f64_ro(base + 8u * l + 8u) <= p) ++l;
    const double a = lds_f64_ro(base + 8u * l), b = lds_f64_ro(base + 8u * l + 8u);
    if (a <= p && p < b && fabs(p) <= T.big) {
      const uint32_t pb = eperm_s + 4u * (unsigned)(T.ext_off + l);
      const int ia = lds_s32_ro(pb), ib = lds_s32_ro(pb + 4u);
      const double da = fabs(a - p), db = fabs(b - p);
      return da < db ? ia : (db < da ? ib : min(ia, ib));
    }
  }
  return snap_dim(ext_g + T.ext_off + 1, eperm_g + T.ext_off + 1, T.n, p);
}

// The binary search over the full grids in global memory for one family (the rare evaluations the decision tables
// cannot answer); out of line so that it costs the main path no registers.  p_s = shared-window address of the row.
static __device__ __noinline__ int snap_family_slow(const K1Params& P, int f, uint32_t p_s) {
  const FamDesc& F = P.fam[f];
  int r = 0;
  for (int t = 0; t < F.n_snap; ++t) {
    const SnapTask& T = F.snap[t];
    r += snap_dim(P.grid_sorted + T.smem_off, P.grid_perm + T.smem_off, T.n, lds_f64(p_s + 8u * T.col)) * T.stride;
  }
  return r;
}

// Snap through the decision table (k1_lnprob.cuh): one 16-bit read, plus one or two 8-byte reads when the cell
// straddles switch points; branch-free.  `slow` is raised when the table cannot answer (NaN, |p| > big, a crowded
// cell, no table): the caller then takes the binary search for this evaluation.
__device__ __forceinline__ int snap_fast(const SnapTask& T, double p, uint32_t dlut_s, uint32_t brec_s, uint32_t perm_s,
                                         bool& slow) {
  const int c = min(max(__double2int_rz((p - T.lo) * T.dinv), 0), max(T.dmmax, 0));
  const unsigned code = lds_u16_ro(dlut_s + 2u * (unsigned)(T.dlut_off + c));
  int idx = (int)(code & 0x3fffu);
  if (code & 0x8000u) {
    const uint32_t r = brec_s + 8u * (unsigned)(T.brec_off + idx);
    idx += (p >= lds_f64_ro(r)) ? 1 : 0;
    if (code & 0x4000u) idx += (p >= lds_f64_ro(r + 8u)) ? 1 : 0;
    if (T.perm_off >= 0) idx = (int)lds_u16_ro(perm_s + 2u * (unsigned)(T.perm_off + idx));
  } else {
    idx = (int)(code & 0x7fffu);
  }
  slow = slow || !(fabs(p) <= T.big) || code == 0xffffu || T.dmmax < 0;
  return idx;
}

}  // namespace agnf
