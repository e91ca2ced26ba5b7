// K4 — SED realisations / integrated luminosities, see k4_sed.cuh.
#include "k4_sed.cuh"

#include "k1_common.cuh"

namespace agnf {
namespace {

constexpr int K4_WARPS = K4_THREADS / 32;
constexpr double K4_FILL = -100.0;                 // interp1d(..., fill_value=-100), PLT:675-732

// component -> family (GA 0, SB 1, TO 2, BB 3, RAD 4); component 5 is the dereddened disc
__constant__ int c_comp_fam[K4_COMPS] = {1, 3, 0, 2, 4, 3};
// order in which the reference adds components in its fractions (to, sb, ga, bb, rad), then the dereddened disc
__constant__ int c_sum_order[K4_COMPS] = {3, 0, 2, 1, 4, 5};

struct RowCtx {
  const K4Bracket* bt[K4_COMPS];    // bracket table of the component's template grid
  int32_t yoff[K4_COMPS];           // shared-memory offset of its log10 F_nu row
  int32_t active[K4_COMPS];
  double scale[K4_COMPS];           // 10**norm
  int32_t by_area[K4_COMPS];        // disc without a normalisation parameter: F / (4 pi dlum^2) first (PLT:749-750)
  int64_t row[AGNF_MAX_FAMILIES];
  double area, lumfactor, zp1;
};

// numpy.interp for one point (compiled_base.c: arr_interp): y_j on a template sample; else the linear
// interpolant, written with the precomputed weight w = (x - x_j) / (x_j+1 - x_j) instead of slope * (x - x_j)
// (same value to ~1e-16 |y|); the reference's NaN fall-backs — evaluate from the other end, then y_j if both ends
// are equal — keep 10**-inf = 0 and NaN exactly where numpy puts them.
__device__ __forceinline__ double interp_y(const double* y, const K4Bracket b) {
  if (b.j < 0) return K4_FILL;
  const double yj = y[b.j];
  if (b.w == 0.0) return yj;
  const double yn = y[b.j + 1];
  const double d = yn - yj;
  double r = fma(b.w, d, yj);
  if (r != r) {
    r = fma(b.w - 1.0, d, yn);
    if (r != r && yj == yn) r = yj;
  }
  return r;
}

struct Eval {
  const K4Params* Q;
  const RowCtx* C;
  const double* rows;     // shared-memory template rows
  uint32_t tab_s;
  // F_nu of component c at grid point i, normalised (PLT:744-756)
  __device__ __forceinline__ double flux(int c, int i) const {
    if (c == 3 && (i < Q->to_lo || i >= Q->to_hi)) return 0.0;
    if (c == 4 && i >= Q->rad_hi) return 0.0;
    const int4 raw = __ldg(reinterpret_cast<const int4*>(C->bt[c] + i));
    K4Bracket b;
    b.w = __hiloint2double(raw.y, raw.x);
    b.j = raw.z;
    const double yv = interp_y(rows + C->yoff[c], b);
    double f = pow10_via2(yv, tab_s);
    if (C->by_area[c]) f = f / C->area;
    return f * C->scale[c];
  }
};

template <bool SPECTRA>
__global__ void __launch_bounds__(K4_THREADS) k4_sed_kernel(const __grid_constant__ K1Params P,
                                                            const K4Params* __restrict__ Q, K4Batch B) {
  extern __shared__ __align__(16) double k4_smem[];
  __shared__ RowCtx C;
  double* tab = k4_smem;                         // 2^(j/D)
  double* rows = k4_smem + AGNF_POW2_DENOM;      // template rows
  double* wpart = rows + Q->smem_doubles;        // [n_ranges][K4_COMPS][K4_WARPS]
  double* I_s = wpart + Q->n_ranges * K4_COMPS * K4_WARPS;     // [n_ranges][K4_COMPS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r = blockIdx.x;
  const double* __restrict__ p = B.params + r * P.cfg.n_par;
  const int n_fam = P.cfg.n_fam;

  if (tid < AGNF_POW2_DENOM) tab[tid] = c_exp2_tab[tid * (64 / AGNF_POW2_DENOM)];
  if (tid < n_fam) {                              // snap, as K1 does (DICTIONARIES_AGNfitter.py:184-214)
    const FamDesc& F = P.fam[tid];
    int64_t row = 0;
    for (int d = 0; d < F.n_dim; ++d) {
      if (F.dim[d].is_grid) {
        const int off = F.dim[d].smem_off;
        row += (int64_t)snap_dim(P.grid_sorted + off, P.grid_perm + off, F.dim[d].size, __ldg(p + F.par_offset + d)) *
               F.dim[d].stride;
      }
    }
    C.row[tid] = row;
  }
  __syncthreads();
  if (tid < K4_COMPS) {
    const int c = tid;
    const int f = c_comp_fam[c];
    const bool on = f < n_fam && Q->fam[f].present && !(c == 4 && !P.cfg.radio);
    C.active[c] = on ? 1 : 0;
    if (on) {
      const K4Family& F = Q->fam[f];
      int64_t t = C.row[f];
      int yoff = F.smem_off;
      if (c == 5 && (Q->dered_mode == 1 || Q->dered_mode == 2)) {
        t = __ldg(Q->dered_row + t);
        yoff = Q->dered_off;
      }
      const int g = __ldg(F.t_grp + t);
      C.bt[c] = F.btab + (int64_t)g * Q->G;
      if (c == 5 && Q->dered_mode == 3) {
        C.bt[c] = Q->bands_bt;
        yoff = Q->dered_off;
      }
      C.yoff[c] = yoff;
      // normalisations: GA, SB, TO, [BB], [RAD] after the family blocks (PARAMETERSPACE_AGNfitter.py:201-215)
      const int no = P.norm_offset;
      double norm = 0.0;
      if (c == 2) norm = __ldg(p + no);
      else if (c == 0) norm = __ldg(p + no + 1);
      else if (c == 3) norm = __ldg(p + no + 2);
      else if (c == 1 || c == 5) norm = P.cfg.bbb_has_norm ? __ldg(p + no + 3) : 0.0;
      else norm = __ldg(p + no + 3 + (P.cfg.bbb_has_norm ? 1 : 0));
      C.scale[c] = exp10(norm);
      C.by_area[c] = ((c == 1 || c == 5) && !P.cfg.bbb_has_norm) ? 1 : 0;
    }
  }
  if (tid == 0) {
    const int s = __ldg(B.source_id + r);
    const double dl = __ldg(Q->src_dlum + s);
    C.area = 4.0 * 3.141592653589793 * (dl * dl);        // 4*math.pi*data.dlum**2 (PLT:749)
    C.lumfactor = 4.0 * 3.141592653589793 * (dl * dl);   // 4. * math.pi * distance**2. (PLT:814)
    C.zp1 = 1.0 + __ldg(Q->src_z + s);
  }
  __syncthreads();
  const uint32_t tab_s = smem_u32(tab);

  // ---- log10 F_nu rows of this realisation's templates (get_fluxes of the _4plot objects, DICTIONARIES:350-401)
  for (int f = 0; f < n_fam; ++f) {
    const K4Family& F = Q->fam[f];
    if (!F.present) continue;
    const int64_t t = C.row[f];
    const int g = __ldg(F.t_grp + t);
    const int n = __ldg(F.g_nlam + g);
    const double* __restrict__ v = F.val + __ldg(F.t_off + t);
    double* y = rows + F.smem_off;
    if (F.mode == 0) {
      for (int j = tid; j < n; j += K4_THREADS) y[j] = __ldg(v + j);
    } else {
      const int xo = __ldg(F.g_xoff + g);
      const double E = __ldg(p + F.e_col);
      const double a = F.mode == 3 ? __ldg(p + F.a_col) : 0.0;
      for (int j = tid; j < n; j += K4_THREADS) {
        const double Fj = __ldg(v + j);
        double Fr;
        if (F.mode == 3 && !(__ldg(F.x + xo + j) < 16.685)) {
          Fr = Fj * exp10(a / 0.3838);                                   // alpha_ox scatter, DICTIONARIES:385,396
        } else {
          Fr = Fj * pow10_via2((-0.4 * __ldg(F.k + xo + j)) * E, tab_s);  // MODEL_AGNfitter.py:1209,1252
          if (F.mode != 1 && Fr != Fr) Fr = Fj;                          // Prevot: NaN -> unreddened, :1210
        }
        y[j] = log10(Fr);
      }
    }
    if (f == 3 && Q->dered_mode == 3) {
      const FamDesc& K = P.fam[3];
      const int zi = __ldg(P.zi + __ldg(B.source_id + r));
      const double* __restrict__ tr = K.table + ((int64_t)zi * K.T + __ldg(Q->dered_row + t)) * K.pitch;
      double* yd = rows + Q->dered_off;
      for (int j = tid; j < P.cfg.n_b; j += K4_THREADS) yd[j] = log10(__ldg(tr + j));
    }
    if (f == 3 && (Q->dered_mode == 1 || Q->dered_mode == 2)) {
      const int64_t td = __ldg(Q->dered_row + t);
      const int nd = __ldg(F.g_nlam + __ldg(F.t_grp + td));
      const double* __restrict__ v2 = F.val2 + __ldg(F.t_off + td);
      double* yd = rows + Q->dered_off;
      for (int j = tid; j < nd; j += K4_THREADS) yd[j] = __ldg(v2 + j);
    }
  }
  __syncthreads();
  Eval ev{Q, &C, rows, tab_s};

  if (SPECTRA) {
    // nu L_nu = F * lumfactor * (10**nu_rest / (1 + z))  (PLT:812-819); order sb, bb, ga, to, [rad], total, bbdered
    const bool radio = P.cfg.radio != 0;
    const int n_comp = radio ? 7 : 6;
    const int G = Q->G;
    double* __restrict__ out = B.spectra + r * (int64_t)n_comp * G;
    for (int i = tid; i < G; i += K4_THREADS) {
      const double nuobs = __ldg(Q->xg + i) / C.zp1;
      const double sb = ev.flux(0, i), bb = ev.flux(1, i), ga = ev.flux(2, i), to = ev.flux(3, i);
      const double bd = ev.flux(5, i);
      double tot = ((sb + bb) + ga) + to;
      int o = 4;
      if (radio) {
        const double rad = ev.flux(4, i);
        tot += rad;
        out[(int64_t)4 * G + i] = (rad * C.lumfactor) * nuobs;
        o = 5;
      }
      out[i] = (sb * C.lumfactor) * nuobs;
      out[(int64_t)G + i] = (bb * C.lumfactor) * nuobs;
      out[(int64_t)2 * G + i] = (ga * C.lumfactor) * nuobs;
      out[(int64_t)3 * G + i] = (to * C.lumfactor) * nuobs;
      out[(int64_t)o * G + i] = (tot * C.lumfactor) * nuobs;
      out[(int64_t)(o + 1) * G + i] = (bd * C.lumfactor) * nuobs;
    }
    return;
  }

  // ---- integrated luminosities: trapz of L_nu = nuLnu / nu over each band (PLT:943-947) -----------------
  const int n_ranges = Q->n_ranges;
  const double lfz = C.lumfactor / C.zp1;
  for (int q = 0; q < n_ranges; ++q) {
    const int i0 = Q->r_i0[q], i1 = Q->r_i1[q], mask = Q->r_mask[q];
    if (i1 < 0) continue;                         // monochromatic: below
    double acc[K4_COMPS];
#pragma unroll
    for (int c = 0; c < K4_COMPS; ++c) acc[c] = 0.0;
    const int n_iv = i1 - i0 - 1;                 // trapezoid intervals
    for (int base = warp * 31; base < n_iv; base += K4_WARPS * 31) {
      const int pt = i0 + base + lane;
      const bool have = pt < i1;
      const double xq = have ? __ldg(Q->xg + pt) : 0.0;
      const double dx = __shfl_down_sync(0xffffffffu, xq, 1) - xq;
      const bool iv = lane < 31 && pt + 1 < i1;
#pragma unroll
      for (int c = 0; c < K4_COMPS; ++c) {
        if (!((mask >> c) & 1)) continue;
        const double v = have ? ev.flux(c, pt) * lfz : 0.0;
        const double vn = __shfl_down_sync(0xffffffffu, v, 1);
        if (iv) acc[c] += (dx * (vn + v)) / 2.0;
      }
    }
#pragma unroll
    for (int c = 0; c < K4_COMPS; ++c) {
      if (!((mask >> c) & 1)) continue;
      double a = acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
      if (lane == 0) wpart[(q * K4_COMPS + c) * K4_WARPS + warp] = a;
    }
  }
  __syncthreads();
  for (int t = tid; t < n_ranges * K4_COMPS; t += K4_THREADS) {
    const int q = t / K4_COMPS, c = t - q * K4_COMPS;
    double v = 0.0;
    if ((Q->r_mask[q] >> c) & 1) {
      if (Q->r_i1[q] < 0) {                       // nuLnu at one frequency (PLT:939-941)
        const int i = Q->r_i0[q];
        v = (ev.flux(c, i) * C.lumfactor) * (__ldg(Q->xg + i) / C.zp1);
      } else {
        for (int w = 0; w < K4_WARPS; ++w) v += wpart[t * K4_WARPS + w];
      }
    }
    I_s[t] = v;
  }
  __syncthreads();
  for (int s = tid; s < Q->n_specs; s += K4_THREADS) {
    const double* I = I_s + Q->s_range[s] * K4_COMPS;
    const int nm = Q->s_num[s], dm = Q->s_den[s];
    double num = 0.0, den = 0.0;
    bool first_n = true, first_d = true;
#pragma unroll
    for (int k = 0; k < K4_COMPS; ++k) {
      const int c = c_sum_order[k];
      if ((nm >> c) & 1) {
        num = first_n ? I[c] : num + I[c];
        first_n = false;
      }
      if ((dm >> c) & 1) {
        den = first_d ? I[c] : den + I[c];
        first_d = false;
      }
    }
    B.intlum[r * Q->n_specs + s] = Q->s_kind[s] ? exp10(num / den) : num;
  }
}

template <bool SPECTRA>
cudaError_t launch(const K1Params& P, const K4Params* Qdev, const K4Params& Qhost, const K4Batch& B,
                   cudaStream_t stream) {
  if (B.n == 0) return cudaSuccess;
  const size_t smem = (size_t)(AGNF_POW2_DENOM + Qhost.smem_doubles + Qhost.n_ranges * K4_COMPS * (K4_WARPS + 1)) * 8;
  cudaError_t e = cudaFuncSetAttribute(k4_sed_kernel<SPECTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  k4_sed_kernel<SPECTRA><<<(unsigned)B.n, K4_THREADS, smem, stream>>>(P, Qdev, B);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_k4_intlum(const K1Params& P, const K4Params* Qdev, const K4Params& Qhost, const K4Batch& B,
                             cudaStream_t stream) {
  return launch<false>(P, Qdev, Qhost, B, stream);
}

cudaError_t launch_k4_spectra(const K1Params& P, const K4Params* Qdev, const K4Params& Qhost, const K4Batch& B,
                              cudaStream_t stream) {
  return launch<true>(P, Qdev, Qhost, B, stream);
}

}  // namespace agnf
