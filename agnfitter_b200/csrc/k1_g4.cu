// K1, gather4 kernel (see k1_lnprob.cuh for the mapping to the reference).
//
// One thread owns one evaluation from the first instruction to the last; a warp owns a tile of 32 consecutive
// rows of `params`; ONE persistent CTA per SM (as many warps as shared memory holds, <= 12) walks the tiles with a
// grid stride.  The template rows an evaluation needs are private to it (every walker snaps to its own galaxy /
// starburst / torus row), so neither L1 nor coalescing can help; what bounds such a gather is the number of L1
// wavefronts per byte.  Here the rows never pass through the LSU on their way in: the TMA unit's tile::gather4 mode
// (UTMALDG.2D.GATHER4: one operation fetches FOUR rows of a 2-D tensor by row index) lands them in shared memory,
// and each thread then reads its own rows with 128-bit loads that are bank-conflict free by construction.
//
// A warp runs a two-stage software pipeline over its tiles (few warps fit: latencies are hidden inside the warp):
//   front(t+1)  params rows of tile t+1 (bulk copy started one stage earlier; ensemble mode forms the proposal from
//               the walker state), per-source record and prior box (loads issued first), the snaps of all
//               families (decision table, k1_dev.cuh::snap_fast: one 16-bit read, one 16-byte read when the cell
//               straddles a switch point), the scalar gathers the priors need (issued here, consumed in back),
//               box test, normalisations 10^x — while the rows of tile t are still in flight;
//   back(t)     waits for the rows, band loop (9 chunks of 2 bands per segment as ONE straight-line block; wider
//               band sets take several segments through the same landing area): LDS.128 of the thread's own rows,
//               y / 1/sigma / reddening coefficients by loads whose addresses are warp-uniform, model sum in the
//               reference's order, masked chi^2 in band order — the same arithmetic in the same order as the
//               cooperative kernel: results are bit-identical; priors, ln_probab, acceptance rule (ensemble);
//   issue(t+1)  lanes 0, 4, 8, .. issue one gather4 per family for the rows of lanes l .. l+3.
// Landing layout per family: 8 groups of 640 bytes = 4 row slots of 160 bytes; a slot holds the 18 doubles of a band
// segment; the tensor map's box is 20 doubles wide and odd groups start at column -2 (out-of-bounds columns are
// zero-filled, not fetched), which moves their data 16 bytes up: lane l reads 16-byte unit
// 40 (l >> 2) + 10 (l & 3) + ((l >> 2) & 1) + k, distinct mod 8 within every quarter-warp.
#include "k1_lnprob.cuh"
#include "k1_common.cuh"
#include "k1_dev.cuh"
#include "k3_common.cuh"

#include <type_traits>

namespace agnf {

namespace {

constexpr int G4_MAX_WARPS = 12;
constexpr int G4_MAX_THREADS = 32 * G4_MAX_WARPS;
constexpr int G4_CPS = 9;                       // chunks (2 bands) per segment
constexpr int G4_SEGW = 2 * G4_CPS;             // bands per segment
constexpr int G4_BOXW = G4_SEGW + 2;            // box width in doubles
constexpr int G4_SLOT = G4_BOXW * 8;            // 160 bytes
constexpr int G4_GROUP = 4 * G4_SLOT;           // 640 bytes = 5 x 128
constexpr int G4_TILE = 8 * G4_GROUP;           // one family, 32 rows: 5120 bytes
constexpr int G4_HEAD = 512;                    // 2^(j/16) table (128 bytes) + mbarriers
constexpr int G4_MAXP = 20;                     // parameters per row the unrolled box test covers
static_assert(AGNF_POW2_DENOM == 16, "the 2^(j/D) table takes the first 128 bytes of shared memory");

// -DG4_PROF: per-phase clock sums (lane 0 of every warp), read back through agnf_get_option("g4_prof<i>")
#ifdef G4_PROF
__device__ unsigned long long g4_prof[8];
#define G4_TICK(i)                                                      \
  do {                                                                  \
    const long long now__ = clock64();                                  \
    if (lane == 0) atomicAdd(&g4_prof[i], (unsigned long long)(now__ - tick__)); \
    tick__ = now__;                                                     \
  } while (0)
#else
#define G4_TICK(i)
#endif

struct G4Plan {
  int32_t stage_pad;      // bytes reserved for the staged decision tables (128-byte multiple)
  int32_t per_warp;       // bytes per warp: landing tiles + params rows
  int32_t par_off;        // bytes from the warp's area to its params rows
  int32_t n_tma;          // families fetched by gather4
  int32_t slot[AGNF_MAX_FAMILIES];   // landing tile of family f, or -1: read by ordinary loads
  int32_t n_seg;          // band segments
  int32_t cpr;            // chunks per row of the per-band arrays (nbp / 2)
  int32_t warps;          // warps per CTA
  int32_t smem_bytes;
  int64_t n_tiles;
};

__device__ __forceinline__ double2 ld_v2f64(const double2* p) {          // generic address space
  double2 v;
  asm volatile("ld.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int r0, int r1, int r2,
                                        int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];\n" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

// what front() hands to issue() and back(): one evaluation's scalars, all in registers
struct G4Front {
  int64_t i;                // row of `params` / of the half-ensemble
  int s, zi;
  bool live, inbox;
  int row[AGNF_MAX_FAMILIES], grow[AGNF_MAX_FAMILIES];
  double GA, BB, nGA, nSB, nTO, nBB, nBBc, nRAD, xfac, ebv_gal, ebv_bbb, inv4pidl2;
  uint64_t valid, xray;
  int j15, jir, ir_mode, agn_branch;
  double aux_ga, aux_sb;    // GALAXYatt_dict / eb0 and STARBURST_LIRdict entries of the snapped rows
  // ensemble mode
  int64_t wrow;
  int ens_c;
  double ens_zz, ens_factor, ens_lnu, ens_lnp_old;
};

template <bool GA_FREE, int BB_MODE, bool RADIO, bool ENS, bool DETAIL>
__global__ void __launch_bounds__(G4_MAX_THREADS, 1)
k1_g4_kernel(const __grid_constant__ K1Params P, const __grid_constant__ K1TensorMaps M, const K1Batch B, const G4Plan C) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const agnf_config& cfg = P.cfg;
  const int n_par = cfg.n_par, n_b = cfg.n_b, nbp = P.nbp;
  const bool eb_free = GA_FREE && cfg.eb_mode != 0;
  const int CPR = C.cpr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  // shared-window addresses
  const uint32_t tab_s = smem_u32(smem_raw);                 // 16 doubles
  const uint32_t bar_s = tab_s + 128;                        // [0] tables, [1 + 2 w] params of warp w, [2 + 2 w] rows
  const uint32_t stage_s = tab_s + G4_HEAD;
  const uint32_t dlut_s = stage_s + P.stage2_off_dlut, brec_s = stage_s + P.stage2_off_brec,
                 perm_s = stage_s + P.stage2_off_perm;
  const uint32_t land_s = stage_s + C.stage_pad + (unsigned)warp * (unsigned)C.per_warp;
  const uint32_t q_s = land_s + (unsigned)C.par_off;         // params / proposal rows of the warp
  const uint32_t pbar = bar_s + 8u + 16u * warp, rbar = pbar + 8u;

  if (threadIdx.x == 0) {
    for (int w = 0; w <= 2 * n_warps; ++w) mbar_init(bar_s + 8u * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < AGNF_POW2_DENOM)
    sts_f64(tab_s + 8u * threadIdx.x, c_exp2_tab[threadIdx.x * (64 / AGNF_POW2_DENOM)]);
  __syncthreads();
  if (threadIdx.x == 0 && P.stage2_bytes > 0) {
    mbar_expect_tx(bar_s, (uint32_t)P.stage2_bytes);
    bulk_g2s(stage_s, P.stage2, (uint32_t)P.stage2_bytes, bar_s);
  }
  if (P.stage2_bytes > 0) mbar_wait(bar_s, 0);

  // landing addresses of this lane's rows (see the header comment)
  const uint32_t mine = land_s + (unsigned)(lane >> 2) * G4_GROUP + (unsigned)(lane & 3) * G4_SLOT +
                        (unsigned)((lane >> 2) & 1) * 16u;
  const uint32_t grp_s = land_s + (unsigned)(lane >> 2) * G4_GROUP;
  const int col_shift = ((lane >> 2) & 1) ? -2 : 0;
  const bool issuer = (lane & 3) == 0;
  uint32_t ppar = 0, rpar = 0;                  // mbarrier phase parities
  const int64_t stride = (int64_t)gridDim.x * n_warps;
#ifdef G4_PROF
  long long tick__ = clock64();
#endif

  // the params rows of a tile -> shared memory (plain batches; ensemble mode forms its proposals in front())
  auto start_params = [&](int64_t tile) {
    if (ENS) return;
    const int64_t row0 = tile * 32;
    const int n_dbl = (int)min((int64_t)32, B.n - row0) * n_par;
    const uint32_t par_bytes = (uint32_t)(n_dbl >> 1) * 16u;
    if (lane == 0 && par_bytes > 0) {
      mbar_expect_tx(pbar, par_bytes);
      bulk_g2s(q_s, B.params + row0 * n_par, par_bytes, pbar);   // 256-byte aligned source
    }
  };

  // ------------------------------------ front: everything up to the gathers ---------------------------
  auto front = [&](int64_t tile) {
    G4Front F;
    const int64_t row0 = tile * 32;
    const int rows_here = (int)min((int64_t)32, B.n - row0);
    F.live = lane < rows_here;
    const int my = F.live ? lane : rows_here - 1;      // tail lanes shadow the last row
    F.i = row0 + my;
    const int n_dbl = rows_here * n_par;
    F.wrow = 0;
    F.ens_c = 0;
    F.ens_zz = F.ens_factor = F.ens_lnu = F.ens_lnp_old = 0.0;
    int my_w = 0;
    double zz = 0.0;
    if (ENS) {
      // the proposal is drawn here instead of being read: q = c_j - (c_j - x_k) zz (k3_common.cuh)
      const int H = B.ens.W >> 1;
      const int64_t src = F.i / H;
      const int kw = (int)(F.i - src * H);
      const StretchDraw d = stretch_draw(B.ens.seed, B.ens.a, B.ens.step, B.ens.half, src + B.ens.src_offset, kw, H);
      F.s = (int)src;
      F.wrow = src * B.ens.W + split_walker(B.ens.split, src, B.ens.W, (B.ens.half ? H : 0) + kw);
      my_w = (int)F.wrow;
      F.ens_c = (int)(src * B.ens.W + split_walker(B.ens.split, src, B.ens.W, (B.ens.half ? 0 : H) + d.j));
      F.ens_lnp_old = B.ens.lnp[F.wrow];      // needed only by the acceptance test
      zz = d.zz;
      F.ens_factor = (double)(n_par - 1) * log(d.zz);
      F.ens_lnu = d.lnu;
      F.ens_zz = d.zz;
    } else {
      F.s = __ldg(B.source_id + F.i);
    }
    // per-source record and prior box: issued first, consumed after the snaps
    const int4* __restrict__ sr = reinterpret_cast<const int4*>(P.srcrec + F.s);
    const int4 sr0 = __ldg(sr), sr1 = __ldg(sr + 1), sr2 = __ldg(sr + 2);
    const double2* __restrict__ box = P.pbox + (int64_t)F.s * n_par;
    double2 bx[G4_MAXP];
#pragma unroll
    for (int c = 0; c < G4_MAXP; ++c) bx[c] = __ldg(box + min(c, n_par - 1));
    if (ENS) {
      // Both walker rows are fetched by the whole warp: element e of the flat [32][n_par] block goes to lane
      // e % 32, so consecutive lanes read consecutive coordinates of one walker.  The proposal is formed right
      // here, element by element (the row's stretch factor comes by shuffle), and only q is stored.
      const int d_row = 32 / n_par, d_col = 32 - d_row * n_par;
      int row = lane / n_par, col = lane - row * n_par;
      for (int r0 = 0; r0 < n_par; r0 += 4) {
        double xv[4], cv[4], zv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          xv[u] = 0.0;
          cv[u] = 0.0;
          zv[u] = 0.0;
          if (r0 + u < n_par) {
            const int wr = __shfl_sync(0xffffffffu, my_w, row), cr = __shfl_sync(0xffffffffu, F.ens_c, row);
            zv[u] = __shfl_sync(0xffffffffu, zz, row);
            xv[u] = B.ens.pos[(int64_t)wr * n_par + col];
            cv[u] = B.ens.pos[(int64_t)cr * n_par + col];
            row += d_row;
            col += d_col;
            if (col >= n_par) {
              col -= n_par;
              ++row;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (r0 + u < n_par) sts_f64(q_s + 8u * (unsigned)((r0 + u) * 32 + lane), stretch_coord(cv[u], xv[u], zv[u]));
        }
      }
      __syncwarp();
    } else {
      if ((n_dbl & 1) && lane == 0) sts_f64(q_s + 8u * (n_dbl - 1), __ldg(B.params + row0 * n_par + n_dbl - 1));
      if (((uint32_t)(n_dbl >> 1) * 16u) > 0) {
        mbar_wait(pbar, ppar);
        ppar ^= 1u;
      }
      __syncwarp();
    }
    G4_TICK(0);                          // params rows in shared memory
    const uint32_t p_s = q_s + 8u * (unsigned)(my * n_par);

    // the snaps of all families, four at a time as one straight-line block (their table reads are independent)
#pragma unroll
    for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) F.row[f] = 0;
    bool slow = false;
    for (int t0 = 0; t0 < P.n_tasks; t0 += 4) {
      int idx[4];
      bool sl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const SnapTask& T = P.tasks[min(t0 + u, P.n_tasks - 1)];
        sl[u] = false;
        idx[u] = snap_fast(T, lds_f64(p_s + 8u * T.col), dlut_s, brec_s, perm_s, sl[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const SnapTask& T = P.tasks[min(t0 + u, P.n_tasks - 1)];
        const bool on = t0 + u < P.n_tasks;
        const int add = on ? idx[u] * T.stride : 0;
        slow = slow || (on && sl[u]);
#pragma unroll
        for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) F.row[f] += (T.fam == f) ? add : 0;
      }
    }
    if (slow) {        // NaN, huge values, crowded cells: the binary search over the full grids (global memory)
#pragma unroll
      for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
        if (f < (RADIO ? 5 : 4)) F.row[f] = snap_family_slow(P, f, p_s);
      }
    }
    // the scalar gathers the priors need: issued now, consumed at the end of back()
    F.aux_ga = F.aux_sb = 0.0;
    F.zi = sr0.x;
#pragma unroll
    for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) F.grow[f] = f < (RADIO ? 5 : 4) ? F.zi * (int)P.fam[f].T + F.row[f] : 0;
    if (cfg.eb_mode != 0) {
      F.aux_ga = GA_FREE ? __ldg(P.eb0 + (int64_t)F.grow[0]) : __ldg(P.fam[0].aux + F.row[0]);
      F.aux_sb = __ldg(P.fam[1].aux + F.row[1]);
    }
    G4_TICK(1);                          // snapped

    // box test (NaN fails), normalisations, free parameters
    F.inbox = true;
    {
      double2 pv[G4_MAXP / 2];
      if ((n_par & 1) == 0) {          // even row length: rows are 16-byte aligned -> conflict-free 128-bit reads
#pragma unroll
        for (int c = 0; c < G4_MAXP / 2; ++c) pv[c] = lds_v2f64(p_s + 16u * (unsigned)min(c, (n_par >> 1) - 1));
      } else {
#pragma unroll
        for (int c = 0; c < G4_MAXP / 2; ++c) {
          pv[c].x = lds_f64(p_s + 8u * (unsigned)min(2 * c, n_par - 1));
          pv[c].y = lds_f64(p_s + 8u * (unsigned)min(2 * c + 1, n_par - 1));
        }
      }
#pragma unroll
      for (int c = 0; c < G4_MAXP; ++c) {
        const double v = (c & 1) ? pv[c >> 1].y : pv[c >> 1].x;
        const bool ok = (bx[c].x <= v) && (v <= bx[c].y);
        F.inbox = F.inbox && (ok || c >= n_par);
      }
    }
    const int no = P.norm_offset;
    const double GA = lds_f64(p_s + 8u * no), SB = lds_f64(p_s + 8u * no + 8u), TO = lds_f64(p_s + 8u * no + 16u);
    int kk = 3;
    double BB = 0.0, RAD = 0.0;
    if (cfg.bbb_has_norm) BB = lds_f64(p_s + 8u * (no + kk++));
    if (RADIO) RAD = lds_f64(p_s + 8u * (no + kk++));
    double alpha = 0.0;
    F.ebv_gal = F.ebv_bbb = 0.0;
    if (GA_FREE) F.ebv_gal = lds_f64(p_s + 8u * (P.fam[0].par_offset + P.fam[0].n_dim - 1));
    if (BB_MODE == 1) F.ebv_bbb = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 1));
    if (BB_MODE == 2) {
      F.ebv_bbb = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 2));
      alpha = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 1));
    }
    if (BB_MODE == 3) {
      F.ebv_bbb = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 3));
      alpha = lds_f64(p_s + 8u * (P.fam[3].par_offset + P.fam[3].n_dim - 2));
    }
    F.GA = GA;
    F.BB = BB;
    F.agn_branch = sr0.y;
    F.ir_mode = sr0.z;
    F.jir = (RADIO && sr0.z != 0) ? sr0.w : -1;
    F.valid = ((uint64_t)(uint32_t)sr1.y << 32) | (uint32_t)sr1.x;
    F.xray = ((uint64_t)(uint32_t)sr1.w << 32) | (uint32_t)sr1.z;
    F.j15 = cfg.prior_agnfrac ? sr2.x : -1;
    F.inv4pidl2 = __hiloint2double(sr2.w, sr2.z);
    // 10^x of the normalisations: branch-free first, repeated through the range-checked routine if any argument
    // was out of range (results are identical whenever the fast path applies)
    bool bad_n = false;
    F.nGA = pow10_pinned<true>(GA, tab_s, bad_n);
    F.nSB = pow10_pinned<true>(SB, tab_s, bad_n);
    F.nTO = pow10_pinned<true>(TO, tab_s, bad_n);
    F.nBB = cfg.bbb_has_norm ? pow10_pinned<true>(BB, tab_s, bad_n) : 1.0;
    F.nRAD = RADIO ? pow10_pinned<true>(RAD, tab_s, bad_n) : 0.0;
    F.xfac = (BB_MODE >= 2) ? pow10_pinned<true>((alpha / 0.3838), tab_s, bad_n) : 1.0;
    if (bad_n) {
      F.nGA = pow10_pinned_tile(GA, tab_s);
      F.nSB = pow10_pinned_tile(SB, tab_s);
      F.nTO = pow10_pinned_tile(TO, tab_s);
      if (cfg.bbb_has_norm) F.nBB = pow10_pinned_tile(BB, tab_s);
      if (RADIO) F.nRAD = pow10_pinned_tile(RAD, tab_s);
      if (BB_MODE >= 2) F.xfac = pow10_pinned_tile(alpha / 0.3838, tab_s);
    }
    F.nBBc = F.nBB * (cfg.bbb_has_norm ? 1.0 : F.inv4pidl2);
    if (B.rows && F.live) {
#pragma unroll
      for (int f = 0; f < AGNF_MAX_FAMILIES; ++f)
        if (f < cfg.n_fam) B.rows[F.i * cfg.n_fam + f] = F.row[f];
    }
    __syncwarp();                        // every lane is done with the params rows
    G4_TICK(2);                          // box test, scalars
    return F;
  };

  // ------------------------------------ issue: one gather4 per family and four rows ---------------------
  const uint32_t seg_bytes = (uint32_t)C.n_tma * G4_TILE;
  auto issue = [&](const G4Front& F, int seg) {
    if (lane == 0) mbar_expect_tx(rbar, seg_bytes);
    __syncwarp();
#pragma unroll
    for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
      if (f < (RADIO ? 5 : 4) && C.slot[f] >= 0) {
        const int r1 = __shfl_down_sync(0xffffffffu, F.grow[f], 1), r2 = __shfl_down_sync(0xffffffffu, F.grow[f], 2),
                  r3 = __shfl_down_sync(0xffffffffu, F.grow[f], 3);
        if (issuer)
          gather4(grp_s + (unsigned)C.slot[f] * G4_TILE, &M.m[f], rbar, seg * G4_SEGW + col_shift, F.grow[f], r1, r2, r3);
      }
    }
  };

  // ------------------------------------ back: band loop, priors, result -------------------------------
  auto back = [&](const G4Front& F) {
    const int s = F.s, zi = F.zi;
    const int64_t i = F.i;
    const bool live = F.live;
    const double nGA = F.nGA, nSB = F.nSB, nTO = F.nTO, nBBc = F.nBBc, nRAD = F.nRAD, xfac = F.xfac;
    const double ebv_gal = F.ebv_gal, ebv_bbb = F.ebv_bbb;
    const uint64_t valid = F.valid, xray = F.xray;
    const int j15 = F.j15, jir = F.jir;
    const int j6 = (P.xp_midir_uv | P.xp_ir_xray) ? P.xp_j6 : -1;
    double cap_ga = 0.0, cap_bb = 0.0, cap_to = 0.0, cap_sb = 0.0, cap_rad = 0.0;
    // per-source / per-slice vectors: chunk kg of this thread's source
    const double2* __restrict__ y2 = reinterpret_cast<const double2*>(P.y) + (int64_t)s * CPR;
    const double2* __restrict__ i2 = reinterpret_cast<const double2*>(P.isig) + (int64_t)s * CPR;
    const double2* __restrict__ cb2 = reinterpret_cast<const double2*>(P.cbbb64) + (int64_t)zi * CPR;
    const double2* __restrict__ cg2 = reinterpret_cast<const double2*>(P.cgal64) + (int64_t)zi * CPR;
    const double2* __restrict__ om2 = reinterpret_cast<const double2*>(P.omega) + (int64_t)zi * CPR;
    // families that are not gathered (a handful of rows per z-slice: the loads are nearly warp-uniform)
    const double2* __restrict__ gBB =
        reinterpret_cast<const double2*>(P.fam[3].table + (int64_t)F.grow[3] * P.fam[3].pitch);
    const bool bb_ldg = C.slot[3] < 0;
    const uint32_t aGA = mine + (unsigned)C.slot[0] * G4_TILE, aSB = mine + (unsigned)C.slot[1] * G4_TILE,
                   aTO = mine + (unsigned)C.slot[2] * G4_TILE, aBB = mine + (unsigned)max(C.slot[3], 0) * G4_TILE,
                   aRAD = mine + (unsigned)max(C.slot[RADIO ? 4 : 0], 0) * G4_TILE;

    double chi = 0.0, ebB = 0.0;
    for (int seg = 0; seg < C.n_seg; ++seg) {
      mbar_wait(rbar, rpar);
      rpar ^= 1u;
      G4_TICK(3);                        // rows landed
      // the disc row through a generic address: shared memory when gathered, else the table itself (no branch)
      const double2* bbp = bb_ldg ? gBB + seg * G4_CPS : reinterpret_cast<const double2*>(smem_raw + (aBB - tab_s));
      bool bad = false;
      const double chi_in = chi, ebB_in = ebB;
      auto chunk = [&](auto fast, const int k, const int kg) {
        auto pw2 = [&](double t) {
          if constexpr (decltype(fast)::value) return pow2_64_nb(t, tab_s, bad);
          else return pow2_64(t, tab_s);
        };
        const double2 dga = lds_v2f64(aGA + 16u * k);
        const double2 dsb = lds_v2f64(aSB + 16u * k);
        const double2 dto = lds_v2f64(aTO + 16u * k);
        const double2 dbb = ld_v2f64(bbp + k);
        const double2 yv = __ldg(y2 + kg);
        const double2 iv = __ldg(i2 + kg);
        double2 drad = make_double2(0.0, 0.0);
        if (RADIO) drad = lds_v2f64(aRAD + 16u * k);
        if (2 * kg == (j15 & ~1)) {
          cap_ga = (j15 & 1) ? dga.y : dga.x;
          cap_bb = (j15 & 1) ? dbb.y : dbb.x;
        }
        if (2 * kg == (j6 & ~1)) cap_to = (j6 & 1) ? dto.y : dto.x;
        if (RADIO && 2 * kg == (jir & ~1)) {
          cap_sb = (jir & 1) ? dsb.y : dsb.x;
          cap_rad = (jir & 1) ? drad.y : drad.x;
        }
        double bb0 = dbb.x, bb1 = dbb.y;
        if (BB_MODE != 0) {
          const double2 cb = __ldg(cb2 + kg);
          const double w0 = dbb.x * pw2(__dmul_rn(cb.x, ebv_bbb));
          const double w1 = dbb.y * pw2(__dmul_rn(cb.y, ebv_bbb));
          bb0 = (w0 != w0) ? dbb.x : w0;                        // NaN -> unreddened (MODEL:1211)
          bb1 = (w1 != w1) ? dbb.y : w1;
          if (BB_MODE >= 2) {
            const unsigned xr = (unsigned)(xray >> (2 * kg));
            if (xr & 1u) bb0 = dbb.x * xfac;
            if (xr & 2u) bb1 = dbb.y * xfac;
          }
        }
        double ga0 = dga.x, ga1 = dga.y;
        if (GA_FREE) {
          const double2 cg = __ldg(cg2 + kg);
          const double rg0 = pw2(__dmul_rn(cg.x, ebv_gal)), rg1 = pw2(__dmul_rn(cg.y, ebv_gal));
          if (eb_free) {
            const double2 om = __ldg(om2 + kg);
            ebB += sum2_pinned(om.x, ((dga.x * 1e18) * rg0) * 3.826e33, om.y, ((dga.y * 1e18) * rg1) * 3.826e33);
          }
          ga0 = dga.x * rg0;
          ga1 = dga.y * rg1;
        }
        // the cooperative kernel's contraction, spelled out (k1_common.cuh: pinned forms)
        double m0 = sum4_pinned(nSB, dsb.x, nBBc, bb0, nGA, ga0, nTO, dto.x);
        double m1 = sum4_pinned(nSB, dsb.y, nBBc, bb1, nGA, ga1, nTO, dto.y);
        if (RADIO) {
          m0 = __fma_rn(nRAD, drad.x, m0);
          m1 = __fma_rn(nRAD, drad.y, m1);
        }
        const unsigned vb = (unsigned)(valid >> (2 * kg));
        const double r0 = (yv.x - m0) * iv.x, r1 = (yv.y - m1) * iv.y;
        const double c0 = (vb & 1u) ? r0 * r0 : 0.0;
        chi += (vb & 2u) ? fma(r1, r1, c0) : c0;
        if (DETAIL && B.ymodel && live) {
          double* ym = B.ymodel + i * n_b + 2 * kg;
          ym[0] = m0;
          if (2 * kg + 1 < n_b) ym[1] = m1;
        }
      };
      if (seg * G4_CPS + G4_CPS <= CPR) {      // a full segment: one straight-line block, loads of all chunks up front
#pragma unroll
        for (int k = 0; k < G4_CPS; ++k) chunk(std::true_type{}, k, seg * G4_CPS + k);
      } else {
#pragma unroll
        for (int k = 0; k < G4_CPS; ++k)
          if (seg * G4_CPS + k < CPR) chunk(std::true_type{}, k, seg * G4_CPS + k);
      }
      if (bad) {       // a reddening exponent out of the fast routine's range: this segment again, range-checked
        chi = chi_in;
        ebB = ebB_in;
#pragma unroll 1
        for (int k = 0; k < G4_CPS; ++k)
          if (seg * G4_CPS + k < CPR) chunk(std::false_type{}, k, seg * G4_CPS + k);
      }
      if (seg + 1 < C.n_seg) {
        __syncwarp();                     // every lane has read this segment: the landing area is free again
        issue(F, seg + 1);
      }
    }
    const double lnlike = -0.5 * chi;
    G4_TICK(4);                          // band loop

    double prior = 0.0;
    if (cfg.eb_mode != 0) {
      const double L_att = GA_FREE ? fabs((F.aux_ga - ebB) * nGA) : F.aux_ga * nGA;
      const double L_sb = F.aux_sb * nSB;
      double pr;
      if (L_sb < L_att) pr = -9999.0;
      else if (L_sb >= L_att)
        pr = (cfg.eb_mode == 1) ? 0.0 : gaussian_prior(P.lognorm_s01, 100.0, 0.0, log10(L_sb / L_att));
      else pr = NAN;   // unordered: the reference has no branch for it
      prior += pr;
    }
    if (cfg.prior_agnfrac) {
      const int j = j15;
      double ga_r = cap_ga;
      if (GA_FREE) ga_r = ga_r * pow2_64(__dmul_rn(__ldg(P.cgal64 + (int64_t)zi * nbp + j), ebv_gal), tab_s);
      const double dbb = cap_bb;
      double bb_r = dbb;
      if (BB_MODE != 0) {
        if (BB_MODE >= 2 && ((xray >> j) & 1ull)) {
          bb_r = dbb * xfac;
        } else {
          const double w = dbb * pow2_64(__dmul_rn(__ldg(P.cbbb64 + (int64_t)zi * nbp + j), ebv_bbb), tab_s);
          bb_r = (w != w) ? dbb : w;
        }
      }
      double bbj = bb_r * F.nBB;
      if (F.BB == 0.0) bbj = bbj * F.inv4pidl2;      // decided by VALUE (PRIORS:177)
      const double frac = log10(bbj / (ga_r * nGA));
      double pr;
      if (F.agn_branch == 0) pr = gaussian_prior(P.lognorm_s2, 0.25, -2.0, frac);
      else pr = (frac < 0.0) ? -9999.0 : gaussian_prior(P.lognorm_s2, 0.25, 2.0, frac);
      prior += pr + gaussian_prior(P.lognorm_s15, 1.0 / 2.25, 4.5, F.GA);
    }
    double tor6 = 0.0, lumf = 0.0;
    if (P.xp_midir_uv | P.xp_ir_xray) {
      tor6 = cap_to * nTO;
      lumf = __ldg(P.consts + (int64_t)s * P.n_const + 1);
    }
    if (P.xp_midir_uv) {
      const int64_t dr = (int64_t)zi * P.fam[3].T + __ldg(P.xp_bb_dered + F.row[3]);
      const double bb2500 = __ldg(P.fam[3].table + dr * P.fam[3].pitch + P.xp_j2500) * F.nBB;
      prior += prior_midir_uv(P, tor6, bb2500, F.BB, lumf, __ldg(P.consts + (int64_t)s * P.n_const + 3));
    }
    if (RADIO) {
      if (F.ir_mode != 0) {
        const double sb_ir = cap_sb * nSB;
        const double syn_ir = cap_rad * nRAD;
        prior += gaussian_prior(P.lognorm_s2, 0.25, F.ir_mode > 0 ? 2.0 : -2.0, log10(syn_ir / sb_ir));
      }
    }
    if (P.xp_ir_xray) prior += prior_ir_xrays(P, tor6, lumf, __ldg(P.consts + (int64_t)s * P.n_const + 2));
    const double lnprior = F.inbox ? prior : -INFINITY;
    const bool fin = (lnprior - lnprior) == 0.0;     // isfinite
    const double lnp_new = fin ? lnprior + lnlike : -INFINITY;
    if (ENS) {
      // emcee's acceptance rule, in place: only this thread touches walker `wrow` during the launch, and the
      // partners it read belong to the other (frozen) half
      const double lnpdiff = F.ens_factor + lnp_new - F.ens_lnp_old;
      const bool accept = live && (lnpdiff > F.ens_lnu);        // false for NaN and for -inf proposals
      if (accept) {
        B.ens.lnp[F.wrow] = lnp_new;
        atomicAdd(B.ens.n_acc + F.wrow, 1);                     // result unused: a fire-and-forget reduction
      }
      // accepted rows go back to `pos` the way they came, consecutive lanes storing consecutive coordinates; the
      // proposal rows may already hold the next tile's, so an accepted proposal is formed again from the two
      // walker rows (both unchanged so far: same operands, same roundings, same value)
      const unsigned acc_mask = __ballot_sync(0xffffffffu, accept);
      if (acc_mask) {
        const int my_w = (int)F.wrow;
        const int d_row = 32 / n_par, d_col = 32 - d_row * n_par;
        int rr = lane / n_par, col = lane - rr * n_par;
        for (int r = 0; r < n_par; ++r) {
          const int wr = __shfl_sync(0xffffffffu, my_w, rr), cr = __shfl_sync(0xffffffffu, F.ens_c, rr);
          const double zr = __shfl_sync(0xffffffffu, F.ens_zz, rr);
          if ((acc_mask >> rr) & 1u) {
            const double xv = B.ens.pos[(int64_t)wr * n_par + col], cv = B.ens.pos[(int64_t)cr * n_par + col];
            B.ens.pos[(int64_t)wr * n_par + col] = stretch_coord(cv, xv, zr);
          }
          rr += d_row;
          col += d_col;
          if (col >= n_par) {
            col -= n_par;
            ++rr;
          }
        }
      }
    } else if (live) {
      B.lnp[i] = lnp_new;
      if (B.lnprior) B.lnprior[i] = lnprior;
      if (B.lnlike) B.lnlike[i] = lnlike;
    }
    G4_TICK(5);                          // priors, output
  };

  // ------------------------------------ the pipeline ---------------------------------------------------
  int64_t tile = (int64_t)blockIdx.x * n_warps + warp;
  if (tile >= C.n_tiles) return;
  start_params(tile);
  G4Front cur = front(tile);
  if (tile + stride < C.n_tiles) start_params(tile + stride);
  issue(cur, 0);
  G4_TICK(6);
  for (;;) {
    const int64_t next = tile + stride;
    const bool has_next = next < C.n_tiles;
    G4Front nxt;
    if (has_next) {
      nxt = front(next);                 // overlaps the flight of `cur`'s rows
      if (next + stride < C.n_tiles) start_params(next + stride);
    }
    back(cur);
    if (!has_next) break;
    __syncwarp();                        // every lane has read the landing area
    issue(nxt, 0);
    G4_TICK(6);                          // gathers issued
    cur = nxt;
    tile = next;
  }
}

G4Plan make_plan(const K1Params& P, const K1Batch& B) {
  G4Plan c{};
  c.stage_pad = (P.stage2_bytes + 127) & ~127;
  c.cpr = P.nbp / 2;
  c.n_seg = (c.cpr + G4_CPS - 1) / G4_CPS;
  c.n_tma = 0;
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
    c.slot[f] = -1;
    if (f >= P.cfg.n_fam) continue;
    // a disc family of one or two rows per z-slice (R06 / THB21 with E(B-V)_bbb free) is read directly: every lane of a
    // source reads one of two addresses
    if (f == 3 && P.fam[f].T <= 2) continue;
    c.slot[f] = c.n_tma++;
  }
  c.par_off = c.n_tma * G4_TILE;
  c.per_warp = c.par_off + ((32 * P.cfg.n_par * 8 + 127) & ~127);
  c.n_tiles = (B.n + 31) / 32;
  // one CTA per SM: as many warps as the 227 KB hold (the register file holds 12 of them)
  const int avail = 227 * 1024 - G4_HEAD - c.stage_pad;
  c.warps = std::max(0, std::min(G4_MAX_WARPS, avail / c.per_warp));
  c.smem_bytes = G4_HEAD + c.stage_pad + c.warps * c.per_warp;
  return c;
}

template <bool GA_FREE, int BB_MODE, bool RADIO, bool ENS, bool DETAIL>
cudaError_t launch_g4_inst(const K1Params& P, const K1TensorMaps& M, const K1Batch& B, cudaStream_t stream) {
  auto kern = k1_g4_kernel<GA_FREE, BB_MODE, RADIO, ENS, DETAIL>;
  const G4Plan C = make_plan(P, B);
  static int attr_bytes[64] = {0};   // per instantiation and device
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (C.smem_bytes > attr_bytes[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C.smem_bytes);
    if (e != cudaSuccess) return e;
    attr_bytes[dev] = C.smem_bytes;
  }
  if (!sms[dev]) {
    cudaError_t e = cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  // small batches: fewer warps per CTA so that every SM gets tiles
  int warps = C.warps;
  while (warps > 1 && (C.n_tiles + warps - 1) / warps < sms[dev]) --warps;
  const int64_t want = (C.n_tiles + warps - 1) / warps;
  const unsigned blocks = (unsigned)(want < sms[dev] ? want : sms[dev]);
  kern<<<blocks, 32 * warps, (size_t)C.smem_bytes, stream>>>(P, M, B, C);
  return cudaGetLastError();
}

template <bool GA_FREE, int BB_MODE, bool RADIO>
cudaError_t launch_g4_ens(const K1Params& P, const K1TensorMaps& M, const K1Batch& B, cudaStream_t stream) {
  if (B.ens.pos) return launch_g4_inst<GA_FREE, BB_MODE, RADIO, true, false>(P, M, B, stream);
  if (B.ymodel) return launch_g4_inst<GA_FREE, BB_MODE, RADIO, false, true>(P, M, B, stream);
  return launch_g4_inst<GA_FREE, BB_MODE, RADIO, false, false>(P, M, B, stream);
}

template <bool GA_FREE, bool RADIO>
cudaError_t launch_g4_bb(const K1Params& P, const K1TensorMaps& M, const K1Batch& B, cudaStream_t stream) {
  switch (P.cfg.bb_mode) {
    case 0: return launch_g4_ens<GA_FREE, 0, RADIO>(P, M, B, stream);
    case 1: return launch_g4_ens<GA_FREE, 1, RADIO>(P, M, B, stream);
    case 2: return launch_g4_ens<GA_FREE, 2, RADIO>(P, M, B, stream);
    default: return launch_g4_ens<GA_FREE, 3, RADIO>(P, M, B, stream);
  }
}

}  // namespace

bool k1_g4_supported(const K1Params& P, const K1TensorMaps& M, const K1Batch& B, const char** why) {
  const char* w = nullptr;
  if (P.nbp & 1) w = "band pitch is odd";
  else if (!B.ens.pos && (reinterpret_cast<uintptr_t>(B.params) & 15)) w = "params is not 16-byte aligned";
  else if (B.ens.pos && P.S * (int64_t)B.ens.W >= (int64_t)INT32_MAX) w = "too many walkers for 32-bit row indices";
  else if (P.cfg.ga_free_ebv && P.cfg.eb_mode != 0 && !P.eb0) w = "the unreddened galaxy integrals have not been built";
  else if (P.cfg.n_par > G4_MAXP) w = "more parameters than the unrolled box test covers";
  else if (!P.srcrec) w = "the per-source records have not been built";
  else if (P.cfg.n_fam < 4) w = "fewer than four families";
  else {
    for (int f = 0; f < P.cfg.n_fam && !w; ++f) {
      if (P.fam[f].pitch & 1) w = "a family table has an odd row pitch";
      else if (reinterpret_cast<uintptr_t>(P.fam[f].table) & 15) w = "a family table is not 16-byte aligned";
      else if ((int64_t)P.cfg.n_z * P.fam[f].T >= (int64_t)INT32_MAX) w = "a family table has 2^31 rows or more";
      else if (!M.valid[f]) w = "no tensor map for a family table";
    }
    if (!w && make_plan(P, B).warps < 4) w = "decision tables + landing tiles leave room for fewer than four warps";
  }
  if (why) *why = w;
  return w == nullptr;
}

cudaError_t launch_k1_g4(const K1Params& P, const K1TensorMaps& M, const K1Batch& B, cudaStream_t stream) {
  if (B.n <= 0) return cudaSuccess;
  const bool gf = P.cfg.ga_free_ebv != 0, rad = P.cfg.radio != 0;
  if (gf) return rad ? launch_g4_bb<true, true>(P, M, B, stream) : launch_g4_bb<true, false>(P, M, B, stream);
  return rad ? launch_g4_bb<false, true>(P, M, B, stream) : launch_g4_bb<false, false>(P, M, B, stream);
}

// per-source record of the gather4 kernel: what an evaluation needs to know about its source, in three 16-byte loads
__global__ void k1_srcrec_kernel(K1Params P, K1SrcRec* __restrict__ out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.S) return;
  K1SrcRec r;
  r.zi = P.zi[s];
  r.agn_branch = P.agn_branch[s];
  r.irsyn_mode = P.irsyn_mode[s];
  r.irsyn_j = P.irsyn_j[s];
  r.valid = P.valid[s];
  r.xray = P.xray[r.zi];
  r.j1500 = P.j1500[r.zi];
  r.pad = 0;
  r.inv4pidl2 = P.consts[s * P.n_const];
  out[s] = r;
}

cudaError_t launch_k1_srcrec(const K1Params& P, K1SrcRec* out, cudaStream_t stream) {
  if (P.S <= 0) return cudaSuccess;
  k1_srcrec_kernel<<<(unsigned)((P.S + 255) / 256), 256, 0, stream>>>(P, out);
  return cudaGetLastError();
}

unsigned long long k1_g4_prof_read(int i, bool reset) {
#ifdef G4_PROF
  unsigned long long h[8] = {0};
  cudaMemcpyFromSymbol(h, g4_prof, sizeof(h));
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g4_prof, z, sizeof(z));
  }
  return (i >= 0 && i < 8) ? h[i] : 0;
#else
  (void)i;
  (void)reset;
  return 0;
#endif
}

// Tensor map of one family table for the gather4 kernel: 2-D fp64 [n_z * T rows, pitch columns], box 20 x 1.
// cuTensorMapEncodeTiled comes through the runtime's driver entry point lookup: the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int k1_g4_make_map(K1TensorMaps& M, int f, const double* table, int64_t rows, int pitch) {
  M.valid[f] = 0;
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
      return 1;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if ((pitch & 1) || (reinterpret_cast<uintptr_t>(table) & 15) || rows < 1 || rows >= (int64_t)INT32_MAX) return 1;
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 8};
  cuuint32_t box[2] = {(cuuint32_t)G4_BOXW, 1};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = encode(&M.m[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(table), dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 1;
  M.valid[f] = 1;
  return 0;
}

}  // namespace agnf
