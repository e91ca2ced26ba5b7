// C ABI of libagnfit_b200.so (declared in include/agnfit_b200.h).  Owns the device-side copies of
// the dictionary / per-source operands and launches K1 and K2.  No torch types anywhere.
#include <algorithm>
#include <climits>
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/agnfit_b200.h"
#include "k1_lnprob.cuh"
#include "k2_dictbuild.cuh"
#include "k3_ensemble.cuh"
#include "k4_sed.cuh"

using namespace agnf;

namespace agnf {
cudaError_t launch_fp64_peak(int kind, double* scratch, int blocks, int iters, double* flops, cudaStream_t stream);
}

struct ScratchBlock {
  void* ptr;
  size_t bytes;
  bool in_use;
};

struct agnf_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};   // host-variant pipeline: H2D / K1 / D2H of successive chunks overlap
  cudaEvent_t pipe_done[3] = {nullptr, nullptr, nullptr};
  int64_t pipeline_rows = 1 << 20;              // chunk size of the host-variant pipeline (0: one shot)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;     // bracket the last K1 kernel
  cudaEvent_t ev2 = nullptr, ev3 = nullptr;     // bracket the last K2 contraction
  bool k1_timed = false, k2_timed = false;
  bool untimed_once = false;                   // the next agnf_lnprob_batch_device call records no timing events
  std::string err;
  std::vector<void*> owned;                    // device allocations freed in destroy
  int64_t launches = 0;
  K1TensorMaps maps{};                         // gather4 kernel: one tensor map per family table
  int k1_variant = 1;                          // 1: automatic (the cooperative kernel when profitable, else
                                               // per-thread); 0: per-thread kernel; 2: cooperative whenever
                                               // supported; 3: gather4 (TMA) kernel whenever supported
  int coop_line = 1;                           // cooperative kernel: line form of the band phase (0: pair form; 2: forced)
  int last_k1_kernel = -1;                     // which kernel the last K1 launch used (0 / 1 cooperative / 2 gather4)

  bool have_cfg = false;
  K1Params k1{};
  bool fam_set[AGNF_MAX_FAMILIES] = {false, false, false, false, false};
  std::vector<double> h_grid;                  // staged grid values (host mirror, rebuilt per upload)
  std::vector<int32_t> h_perm;
  void* d_stage = nullptr;
  std::vector<ScratchBlock> pool;              // scratch blocks of agnf_dict_build (scratch_alloc)
  double k2_prof[10] = {0};                    // host-side time stamps of the last agnf_dict_build (us since entry)
  int64_t snap_crowded = 0, snap_cells = 0, snap_no_table = 0;   // decision-table statistics (debug options)
  bool srcrec_dirty = true;                    // per-source records of the gather4 kernel
  void* d_stage2 = nullptr;                    // decision tables of the gather4 kernel
  std::vector<int32_t> fam_grid_len = std::vector<int32_t>(AGNF_MAX_FAMILIES, 0);
  std::vector<std::vector<double>> fam_grid = std::vector<std::vector<double>>(AGNF_MAX_FAMILIES);
  std::vector<std::vector<int32_t>> fam_perm = std::vector<std::vector<int32_t>>(AGNF_MAX_FAMILIES);
  bool grids_dirty = true;
  void* d_grid = nullptr;
  void* d_perm = nullptr;
  bool have_z = false, have_src = false;
  bool eb0_dirty = true;                        // K1Params::eb0 must be rebuilt (galaxy table or z-slices changed)

  // small-batch path of the host variant: one pinned block [params | source_id | lnp] so that a call is one
  // upload, one launch, one download — all truly asynchronous — instead of three pageable copies
  void* pin = nullptr;                          // cudaHostAlloc
  void* pin_dev = nullptr;                      // device mirror of the block
  void* pin_map = nullptr;                      // device-side address of the pinned block itself (zero copy)
  int64_t zero_copy_rows = 256;                 // batches up to this many rows are read in place over PCIe
  int64_t small_rows = 4096;                    // batches up to this many rows take the path (0: off)
  // host-variant staging
  void* st_params = nullptr; size_t st_params_cap = 0;
  void* st_sid = nullptr;    size_t st_sid_cap = 0;
  void* st_out = nullptr;    size_t st_out_cap = 0;

  // K2
  K2Filters filt{};
  bool have_filt = false;

  // K3 ensemble (allocated by agnf_ens_create)
  bool have_ens = false;
  K3State ens{};
  K3Scratch ens_sc{};
  K3Hist ens_hist{};
  double* ens_best_pos = nullptr;
  double* ens_best_lnp = nullptr;
  int64_t ens_step = 0;              // stretch-move steps taken since agnf_ens_create (Philox counter)
  int ens_persist_rows = 0;          // persistent sampler: walkers per warp (1..32); 0 automatic (see agnf_ens_run)
  int ens_persist = -1;              // persistent per-source sampler: -1 automatic (small shards), 0 never, 1 whenever possible
  int ens_last_persist = 0;          // whether the last agnf_ens_run took the persistent kernel
  int ens_fused = 1;                 // 1: one fused K1 launch per half-step; 0: propose / K1 / accept kernels
  int ens_split = 1;                 // 1: emcee 3's random split of the walkers at every step; 0: fixed halves
  int k2_shared_window = 1;          // 1: 4 z-slices share each template read when the family allows it
  int k2_mma = 1;                    // 1: FP64 tensor-core contraction (8-slice batches) when the family allows it
  double k2_last_flops = 0.0;        // multiply-adds x 2 the last contraction executed (windows as laid out)
  int k2_last_kernel = 0;            // contraction of the last agnf_dict_build: 1 per-slice kernels, 2 shared-window
  int64_t ens_src_offset = 0;        // catalogue row of the shard's first source (Philox counter)
  int64_t ens_hist_samples = 0;
  // chain recording (agnf_ens_record): step-major copies of the walker state, for the reference's chain files
  double* ens_chain = nullptr;       // [cap, S * W * n_par]
  double* ens_chain_lnp = nullptr;   // [cap, S * W]
  int64_t ens_chain_cap = 0, ens_chain_n = 0;
  int32_t ens_chain_every = 1;
  int64_t ens_chain_origin = 0;      // ens_step when recording was switched on

  // K4 SED realisations
  K4Params k4{};
  void* d_k4 = nullptr;              // device copy of k4
  bool k4_cfg = false, k4_src = false, k4_bands = false, k4_dirty = true;
  bool k4_fam[AGNF_MAX_FAMILIES] = {false, false, false, false, false};
  int64_t k4_S = 0;
};

static thread_local std::string g_create_err;

namespace {

int fail(agnf_ctx* ctx, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_err = buf;
  return 1;
}

// for the catch handlers of the entry points: must not throw itself
int fail_noexcept(agnf_ctx* ctx, const char* what) noexcept {
  try {
    std::string& dst = ctx ? ctx->err : g_create_err;
    dst.assign("internal error: ");
    dst.append(what ? what : "?");
  } catch (...) {
  }
  return 1;
}

#define CUDA_TRY(ctx, expr)                                                                  \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return fail(ctx, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

int dev_alloc(agnf_ctx* ctx, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 8;
  CUDA_TRY(ctx, cudaMalloc(p, bytes));
  ctx->owned.push_back(*p);
  return 0;
}

void dev_free(agnf_ctx* ctx, void* p) {
  if (!p) return;
  auto it = std::find(ctx->owned.begin(), ctx->owned.end(), p);
  if (it != ctx->owned.end()) ctx->owned.erase(it);
  cudaFree(p);
}

// Scratch memory of agnf_dict_build: blocks go back to a per-context pool instead of cudaFree (which synchronises
// the device) and are handed out again to later calls.  Everything that touches them runs on ctx->stream, so reuse
// is ordered.  Blocks are released in agnf_ctx_destroy (they sit in `owned`) or by agnf_set_option("release_scratch").
int scratch_alloc(agnf_ctx* ctx, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 8;
  size_t best = SIZE_MAX;
  int arg = -1;
  for (size_t i = 0; i < ctx->pool.size(); ++i) {
    const auto& b = ctx->pool[i];
    if (!b.in_use && b.bytes >= bytes && b.bytes <= 2 * bytes + 4096 && b.bytes < best) {
      best = b.bytes;
      arg = (int)i;
    }
  }
  if (arg >= 0) {
    ctx->pool[(size_t)arg].in_use = true;
    *p = ctx->pool[(size_t)arg].ptr;
    return 0;
  }
  const size_t cap = bytes < 65536 ? ((bytes + 511) & ~(size_t)511) : ((bytes + 65535) & ~(size_t)65535);
  if (dev_alloc(ctx, p, cap)) return 1;
  ctx->pool.push_back({*p, cap, true});
  return 0;
}

void scratch_free(agnf_ctx* ctx, void* p) {
  if (!p) return;
  for (auto& b : ctx->pool)
    if (b.ptr == p) {
      b.in_use = false;
      return;
    }
  dev_free(ctx, p);
}

void scratch_release(agnf_ctx* ctx) {
  for (auto& b : ctx->pool)
    if (!b.in_use) dev_free(ctx, b.ptr);
  ctx->pool.erase(std::remove_if(ctx->pool.begin(), ctx->pool.end(), [](const ScratchBlock& b) { return !b.in_use; }),
                  ctx->pool.end());
}

template <typename T>
int upload(agnf_ctx* ctx, const T** dst, const T* src, size_t count) {
  if (*dst) dev_free(ctx, const_cast<T*>(*dst));
  void* p = nullptr;
  if (dev_alloc(ctx, &p, count * sizeof(T))) return 1;
  if (count) CUDA_TRY(ctx, cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *dst = static_cast<const T*>(p);
  return 0;
}

// [rows, n_b] host array -> [rows, nbp] device array, pad columns zero
int upload_padded(agnf_ctx* ctx, const double** dst, const double* src, size_t rows, int n_b, int nbp) {
  if (*dst) dev_free(ctx, const_cast<double*>(*dst));
  void* p = nullptr;
  if (dev_alloc(ctx, &p, rows * (size_t)nbp * 8)) return 1;
  if (rows) {
    if (nbp != n_b) CUDA_TRY(ctx, cudaMemsetAsync(p, 0, rows * (size_t)nbp * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(p, (size_t)nbp * 8, src, (size_t)n_b * 8, (size_t)n_b * 8, rows,
                                    cudaMemcpyHostToDevice, ctx->stream));
  }
  *dst = static_cast<const double*>(p);
  return 0;
}

int ensure(agnf_ctx* ctx, void** p, size_t* cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  if (*p) dev_free(ctx, *p);
  *cap = 0;
  if (dev_alloc(ctx, p, bytes)) return 1;
  *cap = bytes;
  return 0;
}

// The ensemble (walkers, scratch, histograms, chain buffer) is sized for one set of sources and one parameter
// layout: a new model-set configuration or source upload invalidates it (agnf_ens_create must follow).
void ens_teardown(agnf_ctx* ctx) {
  if (!ctx->have_ens) return;
  cudaStreamSynchronize(ctx->stream);
  for (void* q : {(void*)ctx->ens.pos, (void*)ctx->ens.lnp, (void*)ctx->ens.n_acc, (void*)ctx->ens.split,
                  (void*)ctx->ens_sc.prop, (void*)ctx->ens_sc.prop_lnp, (void*)ctx->ens_sc.factor,
                  (void*)ctx->ens_sc.lnu, (void*)ctx->ens_sc.sid, (void*)ctx->ens_hist.hist, (void*)ctx->ens_best_pos,
                  (void*)ctx->ens_best_lnp, (void*)ctx->ens_chain, (void*)ctx->ens_chain_lnp})
    dev_free(ctx, q);
  ctx->ens = K3State{};
  ctx->ens_sc = K3Scratch{};
  ctx->ens_hist = K3Hist{};
  ctx->ens_best_pos = ctx->ens_best_lnp = nullptr;
  ctx->ens_chain = ctx->ens_chain_lnp = nullptr;
  ctx->ens_chain_cap = ctx->ens_chain_n = 0;
  ctx->have_ens = false;
}

// ---- decision tables for the gather4 kernel's snap (k1_dev.cuh::snap_fast) -------------------------------------
// Doubles in lattice order <-> int64 keys (monotone), for bisection over every representable value.
inline int64_t dkey(double x) {
  int64_t k;
  memcpy(&k, &x, 8);
  return k < 0 ? (int64_t)0x8000000000000000ull - k : k;
}
inline double dval(int64_t k) {
  if (k < 0) k = (int64_t)0x8000000000000000ull - k;
  double x;
  memcpy(&x, &k, 8);
  return x;
}
// the device's cell function, operation for operation (no contraction possible: subtract, multiply, truncate, clamp)
inline int dcell(double p, double lo, double inv, int mmax) {
  volatile double d = p - lo;
  volatile double t = d * inv;
  int c;
  if (!(t == t)) c = 0;
  else if (t >= 2147483647.0) c = INT32_MAX;
  else if (t <= -2147483648.0) c = INT32_MIN;
  else c = (int)t;                                   // truncation, like cvt.rzi
  return std::min(std::max(c, 0), mmax);
}
// which of the two bracketing grid values a <= p <= b the reference's argmin picks (first-appearance indices ia, ib)
inline bool picks_upper(double a, double b, int ia, int ib, double p) {
  volatile double da = std::fabs(a - p), db = std::fabs(b - p);
  return ia < ib ? (db < da) : (db <= da);
}

// One dimension: sorted values sv[n], first-appearance indices pv[n].  Appends cell codes to `dlut`, n - 1 switch
// points to `thr` and (unless pv is the identity) n indices to `perm16`.  Returns the number of cells (0: no table).
int build_decision_table(const double* sv, const int32_t* pv, int n, double big, std::vector<uint16_t>& dlut,
                         std::vector<double>& thr, std::vector<uint16_t>& perm16, double* inv_out, int* perm_off) {
  if (n < 2 || n >= 16000) return 0;
  const double lo = sv[0], span = sv[n - 1] - sv[0];
  if (!(span > 0.0) || !std::isfinite(span) || !(big > 0.0)) return 0;
  // switch points: the smallest p in (sv[j], sv[j+1]] for which the argmin moves to the upper value
  std::vector<double> sw((size_t)n - 1);
  for (int j = 0; j + 1 < n; ++j) {
    int64_t a = dkey(sv[j]), b = dkey(sv[j + 1]);     // picks_upper(a) false, picks_upper(b) true
    while (b - a > 1) {
      const int64_t mid = a + (b - a) / 2;
      if (picks_upper(sv[j], sv[j + 1], pv[j], pv[j + 1], dval(mid))) b = mid; else a = mid;
    }
    sw[(size_t)j] = dval(b);
  }
  bool identity = true;
  for (int k = 0; k < n; ++k) identity = identity && pv[k] == k;
  int best_m = 0;
  std::vector<uint16_t> codes;
  int m0 = 32;
  while (m0 < 2 * n) m0 *= 2;
  for (int m = m0; m <= 8 * m0; m *= 2) {
    const double inv = (double)m / span;
    const int mmax = m - 1;
    std::vector<int> first_str((size_t)m, -1), n_str((size_t)m, 0), first_after((size_t)m + 1, 0);
    bool usable = true;
    for (int j = 0; j + 1 < n && usable; ++j) {
      const double t = sw[(size_t)j];
      const int cs = dcell(t, lo, inv, mmax), cp = dcell(dval(dkey(t) - 1), lo, inv, mmax);
      if (cs - cp > 1 || cp > cs) usable = false;     // a cell narrower than one ulp: no table at this size
      if (cs == cp) {
        if (n_str[(size_t)cs] == 0) first_str[(size_t)cs] = j;
        n_str[(size_t)cs] += 1;
        first_after[(size_t)cs + 1] += 1;             // cells after cs start past switch j
      } else {
        first_after[(size_t)cs] += 1;                 // the switch sits on the edge cp | cs
      }
    }
    if (!usable) continue;
    int crowded = 0;
    std::vector<uint16_t> cand((size_t)m);
    int s0 = 0;                                        // switches entirely below the cell = sorted position at its start
    for (int c = 0; c < m; ++c) {
      s0 += first_after[(size_t)c];
      const int ns = n_str[(size_t)c];
      if (ns == 0) cand[(size_t)c] = (uint16_t)pv[std::min(s0, n - 1)];
      else if (ns <= 2 && first_str[(size_t)c] == s0)
        cand[(size_t)c] = (uint16_t)(0x8000u | (ns == 2 ? 0x4000u : 0u) | (unsigned)s0);
      else {
        cand[(size_t)c] = 0xffffu;
        ++crowded;
      }
    }
    if (best_m == 0 || crowded == 0) {
      best_m = m;
      codes = cand;
      *inv_out = inv;
    }
    if (crowded == 0) break;
  }
  if (best_m == 0) return 0;
  dlut.insert(dlut.end(), codes.begin(), codes.end());
  thr.insert(thr.end(), sw.begin(), sw.end());
  *perm_off = -1;
  if (!identity) {
    *perm_off = (int)perm16.size();
    for (int k = 0; k < n; ++k) perm16.push_back((uint16_t)pv[k]);
  }
  return best_m;
}

int rebuild_grids(agnf_ctx* ctx) {
  if (!ctx->grids_dirty) return 0;
  ctx->h_grid.clear();
  ctx->h_perm.clear();
  std::vector<double> ext;          // sentinel-extended sorted grids
  std::vector<int32_t> eperm;
  std::vector<uint16_t> lut;
  std::vector<uint16_t> dlut;       // decision tables (gather4 kernel)
  std::vector<double> brec;         // switch points
  std::vector<uint16_t> perm16;
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
    if (!ctx->fam_set[f]) continue;
    int off = (int)ctx->h_grid.size();
    int local = 0;
    FamDesc& F = ctx->k1.fam[f];
    for (int d = 0; d < F.n_dim; ++d) {
      DimDesc& D = F.dim[d];
      D.smem_off = off + local;
      const double* sv = ctx->fam_grid[f].data() + local;
      const int32_t* pv = ctx->fam_perm[f].data() + local;
      const int n = D.size;
      D.ext_off = (int)ext.size();
      ext.push_back(-INFINITY);
      eperm.push_back(INT32_MAX);
      for (int k = 0; k < n; ++k) {
        ext.push_back(sv[k]);
        eperm.push_back(pv[k]);
      }
      ext.push_back(INFINITY);
      eperm.push_back(INT32_MAX);
      // Cell table for the snap (k1_common.cuh::snap_dim_lut).  The table path is exact only if, for
      // |p| <= big, two grid values can never give the same rounded |g - p| unless they bracket p:
      // with M = max |g| and big = 1e6 M, |g - p| < 2e6 M, whose ulp (< 5e-10 M) must stay below the
      // smallest spacing; we ask for spacing >= 1e-6 M.  Each cell may hold at most 2 grid values
      // (the kernel makes 2 up-steps); otherwise the cell count doubles, up to 4096.
      D.lut_off = (int)lut.size();
      D.lut_m = 0;
      D.lut_lo = D.lut_inv = D.lut_big = 0.0;
      if (D.is_grid && n > 1 && n < 60000 && std::isfinite(sv[0]) && std::isfinite(sv[n - 1])) {
        const double M = std::max(std::fabs(sv[0]), std::fabs(sv[n - 1]));
        double min_gap = INFINITY;
        for (int k = 1; k < n; ++k) min_gap = std::min(min_gap, sv[k] - sv[k - 1]);
        const double span = sv[n - 1] - sv[0];
        if (min_gap >= 1e-6 * M && std::isfinite(span) && span > 0.0) {
          int m = 16;
          while (m < 2 * n && m < 4096) m *= 2;
          for (; m <= 4096; m *= 2) {
            std::vector<uint16_t> cand((size_t)m);
            bool ok = true;
            int l = 0;                               // sv index of the last value <= left edge
            for (int c = 0; c < m && ok; ++c) {
              const double left = sv[0] + span * ((double)c / (double)m);
              const double right = sv[0] + span * ((double)(c + 1) / (double)m);
              while (l + 1 < n && sv[l + 1] <= left) ++l;
              int le = l;
              while (le + 1 < n && sv[le + 1] <= right) ++le;
              if (le - l > 2) ok = false;
              cand[(size_t)c] = (uint16_t)(l + 1);   // ext indexing
            }
            if (ok) {
              D.lut_m = m;
              D.lut_lo = sv[0];
              D.lut_inv = (double)m / span;
              D.lut_big = 1e6 * M;
              lut.insert(lut.end(), cand.begin(), cand.end());
              break;
            }
          }
        }
      }
      local += D.size;
    }
    F.n_snap = 0;
    for (int d = 0; d < F.n_dim; ++d) {
      const DimDesc& D = F.dim[d];
      if (!D.is_grid || D.size < 2) continue;
      SnapTask& T = F.snap[F.n_snap++];
      T.col = F.par_offset + d;
      T.n = D.size;
      T.ext_off = D.ext_off;
      T.lut_off = D.lut_off;
      T.lut_mmax = D.lut_m > 0 ? D.lut_m - 1 : -1;
      T.stride = (int32_t)D.stride;
      T.lo = D.lut_lo;
      T.inv = D.lut_inv;
      T.big = D.lut_big;
      // decision table: needs the same guarantee as the cell table (lut_big > 0: spacing resolvable for |p| <= big)
      T.smem_off = D.smem_off;
      T.dlut_off = (int32_t)dlut.size();
      T.brec_off = (int32_t)brec.size();
      T.dmmax = -1;
      T.dinv = 0.0;
      T.perm_off = -1;
      if (D.lut_m > 0) {
        const int local_off = D.smem_off - off;
        double dinv = 0.0;
        int perm_off = -1;
        const int m = build_decision_table(ctx->fam_grid[f].data() + local_off, ctx->fam_perm[f].data() + local_off,
                                           D.size, D.lut_big, dlut, brec, perm16, &dinv, &perm_off);
        if (m > 0) {
          T.dmmax = m - 1;
          T.dinv = dinv;
          T.perm_off = perm_off;
        }
      }
    }
    ctx->h_grid.insert(ctx->h_grid.end(), ctx->fam_grid[f].begin(), ctx->fam_grid[f].end());
    ctx->h_perm.insert(ctx->h_perm.end(), ctx->fam_perm[f].begin(), ctx->fam_perm[f].end());
  }
  // blob for the cooperative kernel
  auto pad16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
  const size_t off_grid = 0, off_perm = off_grid + pad16(ext.size() * 8),
               off_lut = off_perm + pad16(eperm.size() * 4), total = off_lut + pad16(lut.size() * 2);
  std::vector<unsigned char> blob(total + 16, 0);
  if (!ext.empty()) memcpy(blob.data() + off_grid, ext.data(), ext.size() * 8);
  if (!eperm.empty()) memcpy(blob.data() + off_perm, eperm.data(), eperm.size() * 4);
  if (!lut.empty()) memcpy(blob.data() + off_lut, lut.data(), lut.size() * 2);

  ctx->k1.n_tasks = 0;
  ctx->snap_no_table = 0;
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {
    if (!ctx->fam_set[f]) continue;
    for (int t = 0; t < ctx->k1.fam[f].n_snap; ++t) {
      SnapTask T = ctx->k1.fam[f].snap[t];
      T.fam = f;
      ctx->k1.tasks[ctx->k1.n_tasks++] = T;
      if (T.dmmax < 0) ctx->snap_no_table += 1;
    }
  }
  ctx->snap_cells = (int64_t)dlut.size();
  ctx->snap_crowded = 0;
  for (uint16_t c : dlut) ctx->snap_crowded += c == 0xffffu;
  // blob of the gather4 kernel
  const size_t off_dlut = 0, off_brec = pad16(dlut.size() * 2), off_p16 = off_brec + pad16(brec.size() * 8 + 8),
               total2 = off_p16 + pad16(perm16.size() * 2);
  std::vector<unsigned char> blob2(total2 + 16, 0);
  if (!dlut.empty()) memcpy(blob2.data() + off_dlut, dlut.data(), dlut.size() * 2);
  if (!brec.empty()) memcpy(blob2.data() + off_brec, brec.data(), brec.size() * 8);
  if (!perm16.empty()) memcpy(blob2.data() + off_p16, perm16.data(), perm16.size() * 2);
  if (ctx->d_stage2) dev_free(ctx, ctx->d_stage2);
  ctx->d_stage2 = nullptr;
  if (dev_alloc(ctx, &ctx->d_stage2, total2 + 16)) return 1;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage2, blob2.data(), total2 + 16, cudaMemcpyHostToDevice, ctx->stream));
  ctx->k1.stage2 = static_cast<const unsigned char*>(ctx->d_stage2);
  ctx->k1.stage2_bytes = (int32_t)total2;
  ctx->k1.stage2_off_dlut = (int32_t)off_dlut;
  ctx->k1.stage2_off_brec = (int32_t)off_brec;
  ctx->k1.stage2_off_perm = (int32_t)off_p16;

  if (ctx->d_grid) dev_free(ctx, ctx->d_grid);
  if (ctx->d_perm) dev_free(ctx, ctx->d_perm);
  if (ctx->d_stage) dev_free(ctx, ctx->d_stage);
  if (dev_alloc(ctx, &ctx->d_grid, ctx->h_grid.size() * sizeof(double))) return 1;
  if (dev_alloc(ctx, &ctx->d_perm, ctx->h_perm.size() * sizeof(int32_t))) return 1;
  if (dev_alloc(ctx, &ctx->d_stage, total)) return 1;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage, blob.data(), total, cudaMemcpyHostToDevice, ctx->stream));
  if (!ctx->h_grid.empty()) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_grid, ctx->h_grid.data(), ctx->h_grid.size() * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_perm, ctx->h_perm.data(), ctx->h_perm.size() * sizeof(int32_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));     // host vectors go out of scope / may be rebuilt
  ctx->k1.grid_sorted = static_cast<const double*>(ctx->d_grid);
  ctx->k1.grid_perm = static_cast<const int32_t*>(ctx->d_perm);
  ctx->k1.n_grid_vals = (int32_t)ctx->h_grid.size();
  ctx->k1.stage = static_cast<const unsigned char*>(ctx->d_stage);
  ctx->k1.stage_bytes = (int32_t)total;
  ctx->k1.stage_off_grid = (int32_t)off_grid;
  ctx->k1.stage_off_perm = (int32_t)off_perm;
  ctx->k1.stage_off_lut = (int32_t)off_lut;
  ctx->grids_dirty = false;
  return 0;
}

int check_ready(agnf_ctx* ctx) {
  if (!ctx->have_cfg) return fail(ctx, "agnf_modelset_config has not been called");
  for (int f = 0; f < ctx->k1.cfg.n_fam; ++f)
    if (!ctx->fam_set[f]) return fail(ctx, "family %d has not been uploaded", f);
  if (!ctx->have_z) return fail(ctx, "agnf_zslices_upload has not been called");
  if (!ctx->have_src) return fail(ctx, "agnf_sources_upload has not been called");
  if (rebuild_grids(ctx)) return 1;
  // free E(B-V)_gal + energy balance: the unreddened integral of every galaxy row, once per dictionary / z-slice
  // upload (needs an even, 16-byte aligned table: exactly the cooperative kernel's own preconditions)
  K1Params& P = ctx->k1;
  const bool want = P.cfg.ga_free_ebv && P.cfg.eb_mode != 0 && !(P.fam[0].pitch & 1) && !(P.nbp & 1) &&
                    !(reinterpret_cast<uintptr_t>(P.fam[0].table) & 15);
  if (ctx->srcrec_dirty) {
    if (P.srcrec) dev_free(ctx, const_cast<K1SrcRec*>(P.srcrec));
    P.srcrec = nullptr;
    void* d = nullptr;
    if (dev_alloc(ctx, &d, (size_t)std::max<int64_t>(P.S, 1) * sizeof(K1SrcRec))) return 1;
    CUDA_TRY(ctx, launch_k1_srcrec(P, static_cast<K1SrcRec*>(d), ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    P.srcrec = static_cast<const K1SrcRec*>(d);
    ctx->srcrec_dirty = false;
    ctx->launches += 1;
  }
  if (want && ctx->eb0_dirty) {
    if (P.eb0) dev_free(ctx, const_cast<double*>(P.eb0));
    P.eb0 = nullptr;
    void* d = nullptr;
    if (dev_alloc(ctx, &d, (size_t)P.cfg.n_z * (size_t)P.fam[0].T * 8)) return 1;
    CUDA_TRY(ctx, launch_k1_eb0(P, static_cast<double*>(d), ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    P.eb0 = static_cast<const double*>(d);
    ctx->eb0_dirty = false;
    ctx->launches += 1;
  }
  return 0;
}

// append the current walker state to the recording buffer (two device-to-device copies on the ctx stream)
int ens_append_state(agnf_ctx* ctx) {
  const K3State& st = ctx->ens;
  if (ctx->ens_chain_n >= ctx->ens_chain_cap)
    return fail(ctx, "chain buffer full after %lld recorded steps (agnf_ens_record)", (long long)ctx->ens_chain_n);
  const size_t nw = (size_t)st.S * st.W;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ens_chain + (size_t)ctx->ens_chain_n * nw * st.n_par, st.pos, nw * st.n_par * 8,
                                cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ens_chain_lnp + (size_t)ctx->ens_chain_n * nw, st.lnp, nw * 8,
                                cudaMemcpyDeviceToDevice, ctx->stream));
  ++ctx->ens_chain_n;
  return 0;
}

// K1 launch through the kernel the options and the preconditions select; returns the kernel id in *which
cudaError_t k1_dispatch(agnf_ctx* ctx, const K1Batch& B, cudaStream_t st, int* which) {
  const int v = ctx->k1_variant;
  if (v == 3 && k1_g4_supported(ctx->k1, ctx->maps, B, nullptr)) {
    *which = 2;
    return launch_k1_g4(ctx->k1, ctx->maps, B, st);
  }
  if (v != 0 && k1_coop_supported(ctx->k1, B, nullptr) && (v >= 2 || k1_coop_profitable(ctx->k1, B))) {
    *which = 1;
    return launch_k1_coop(ctx->k1, B, st);
  }
  *which = 0;
  return B.ens.pos ? cudaErrorNotSupported : launch_k1(ctx->k1, B, st);
}

// fused stretch-move launches need the gather4 or the cooperative kernel
bool k1_fused_available(agnf_ctx* ctx, const K1Batch& B) {
  const int v = ctx->k1_variant;
  if (v == 3 && k1_g4_supported(ctx->k1, ctx->maps, B, nullptr)) return true;
  return v != 0 && k1_coop_supported(ctx->k1, B, nullptr) && (v >= 2 || k1_coop_profitable(ctx->k1, B));
}

}  // namespace

// No exception leaves the library (include/agnfit_b200.h: "never throws"): host containers inside the entry points can
// throw std::bad_alloc / std::length_error; every int-returning entry point turns that into an error code.
#define AGNF_GUARD_BEGIN try {
#define AGNF_GUARD_END(ctx)                                                                   \
  }                                                                                           \
  catch (const std::bad_alloc&) { return fail_noexcept(ctx, "out of host memory"); }          \
  catch (const std::exception& e) { return fail_noexcept(ctx, e.what()); }                    \
  catch (...) { return fail_noexcept(ctx, "unknown internal error"); }

extern "C" {

AGNF_API int agnf_version(void) { return AGNF_VERSION; }

AGNF_API const char* agnf_last_error(const agnf_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

AGNF_API int agnf_ctx_create(int device, agnf_ctx** out) {
  AGNF_GUARD_BEGIN
  if (!out) return fail(nullptr, "agnf_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, "no CUDA device available (%s); this library has no CPU path",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, "device %d out of range (have %d)", device, n);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, "device %d is sm_%d%d; this build only contains sm_100a code", device, prop.major,
                prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, "cudaSetDevice: %s", cudaGetErrorString(e));
  agnf_ctx* ctx = new agnf_ctx();
  ctx->device = device;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < 3; ++i) {
    cudaStreamCreateWithFlags(&ctx->pipe[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->pipe_done[i], cudaEventDisableTiming);
  }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  cudaEventCreate(&ctx->ev2);
  cudaEventCreate(&ctx->ev3);
  *out = ctx;
  return 0;
  AGNF_GUARD_END(nullptr)
}

AGNF_API void agnf_ctx_destroy(agnf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (void* p : ctx->owned) cudaFree(p);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  for (int i = 0; i < 3; ++i) {
    if (ctx->pipe[i]) cudaStreamSynchronize(ctx->pipe[i]);
    if (ctx->pipe_done[i]) cudaEventDestroy(ctx->pipe_done[i]);
    if (ctx->pipe[i]) cudaStreamDestroy(ctx->pipe[i]);
  }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev2) cudaEventDestroy(ctx->ev2);
  if (ctx->ev3) cudaEventDestroy(ctx->ev3);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

AGNF_API int agnf_synchronize(agnf_ctx* ctx) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int64_t agnf_launch_count(const agnf_ctx* ctx) { return ctx ? ctx->launches : 0; }

AGNF_API int agnf_last_kernel_ms(agnf_ctx* ctx, int which, float* ms) {
  AGNF_GUARD_BEGIN
  if (!ctx || !ms) return 1;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (which == 1) {
    if (!ctx->k1_timed) return fail(ctx, "no K1 launch has been timed yet");
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return 0;
  }
  if (which == 2) {
    if (!ctx->k2_timed) return fail(ctx, "no K2 launch has been timed yet");
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev3));
    CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->ev2, ctx->ev3));
    return 0;
  }
  return fail(ctx, "agnf_last_kernel_ms: which must be 1 or 2");
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_set_option(agnf_ctx* ctx, const char* name, int64_t value) {
  AGNF_GUARD_BEGIN
  if (!ctx || !name) return 1;
  if (strcmp(name, "k1_variant") == 0) {
    if (value < 0 || value > 3)
      return fail(ctx, "k1_variant must be 0 (per-thread), 1 (automatic), 2 (cooperative) or 3 (gather4)");
    ctx->k1_variant = (int)value;
    return 0;
  }
  if (strcmp(name, "coop_line") == 0) {
    if (value < 0 || value > 2)
      return fail(ctx, "coop_line must be 0 (pair form), 1 (line form where it is the faster one) or 2 (line form wherever it runs)");
    ctx->coop_line = (int)value;
    ctx->k1.coop_line = ctx->coop_line;
    return 0;
  }
  if (strcmp(name, "ens_source_offset") == 0) {
    if (value < 0) return fail(ctx, "ens_source_offset must be >= 0");
    ctx->ens_src_offset = value;
    if (ctx->have_ens) ctx->ens.src_offset = value;
    return 0;
  }
  if (strcmp(name, "release_scratch") == 0) {
    cudaStreamSynchronize(ctx->stream);
    scratch_release(ctx);
    return 0;
  }
  if (strcmp(name, "k2_mma") == 0) {
    if (value != 0 && value != 1) return fail(ctx, "k2_mma must be 0 or 1");
    ctx->k2_mma = (int)value;
    return 0;
  }
  if (strcmp(name, "k2_shared_window") == 0) {
    if (value != 0 && value != 1) return fail(ctx, "k2_shared_window must be 0 or 1");
    ctx->k2_shared_window = (int)value;
    return 0;
  }
  if (strcmp(name, "ens_persist_rows") == 0) {
    if (value < 0 || value > 32) return fail(ctx, "ens_persist_rows must be 0 (automatic) or 1..32 walkers per warp");
    ctx->ens_persist_rows = (int)value;
    return 0;
  }
  if (strcmp(name, "ens_persist") == 0) {
    if (value < -1 || value > 1) return fail(ctx, "ens_persist must be -1 (automatic), 0 or 1");
    ctx->ens_persist = (int)value;
    return 0;
  }
  if (strcmp(name, "ens_fused") == 0) {
    if (value != 0 && value != 1) return fail(ctx, "ens_fused must be 0 or 1");
    ctx->ens_fused = (int)value;
    return 0;
  }
  if (strcmp(name, "ens_split") == 0) {
    if (value != 0 && value != 1) return fail(ctx, "ens_split must be 0 (fixed halves) or 1 (random split per step)");
    if (ctx->have_ens && value != ctx->ens_split)
      return fail(ctx, "ens_split must be chosen before agnf_ens_create");
    ctx->ens_split = (int)value;
    return 0;
  }
  if (strcmp(name, "zero_copy_rows") == 0) {
    if (value < 0 || value > (1 << 20)) return fail(ctx, "zero_copy_rows must be in [0, 2^20]");
    ctx->zero_copy_rows = value;
    return 0;
  }
  if (strcmp(name, "small_rows") == 0) {
    if (value < 0 || value > (1 << 20)) return fail(ctx, "small_rows must be in [0, 2^20]");
    if (ctx->pin && value > ctx->small_rows) {          // the pinned block was sized for the old limit
      cudaFreeHost(ctx->pin);
      dev_free(ctx, ctx->pin_dev);
      ctx->pin = ctx->pin_dev = nullptr;
    }
    ctx->small_rows = value;
    return 0;
  }
  if (strcmp(name, "pipeline_rows") == 0) {
    if (value < 0) return fail(ctx, "pipeline_rows must be >= 0");
    ctx->pipeline_rows = value;
    return 0;
  }
  return fail(ctx, "unknown option '%s'", name);
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_get_option(agnf_ctx* ctx, const char* name, int64_t* value) {
  AGNF_GUARD_BEGIN
  if (!ctx || !name || !value) return 1;
  if (strcmp(name, "k1_variant") == 0) *value = ctx->k1_variant;
  else if (strcmp(name, "k1_last_kernel") == 0) *value = ctx->last_k1_kernel;
  else if (strcmp(name, "coop_line") == 0) *value = ctx->coop_line;
  else if (strcmp(name, "k1_table_bytes") == 0) *value = ctx->k1.stage2_bytes;
  else if (strcmp(name, "k1_table_cells") == 0) *value = ctx->snap_cells;
  else if (strcmp(name, "k1_table_crowded_cells") == 0) *value = ctx->snap_crowded;
  else if (strcmp(name, "k1_tasks_without_table") == 0) *value = ctx->snap_no_table;
  else if (strncmp(name, "k2_prof", 7) == 0 && name[7] >= '0' && name[7] <= '9') *value = (int64_t)ctx->k2_prof[name[7] - '0'];
  else if (strncmp(name, "kc_prof", 7) == 0) {
    cudaStreamSynchronize(ctx->stream);
    *value = (int64_t)k1_coop_prof_read(name[7] - '0');
  }
  else if (strncmp(name, "g4_prof", 7) == 0) {
    cudaStreamSynchronize(ctx->stream);
    *value = (int64_t)k1_g4_prof_read(name[7] - '0', name[8] == 'r');
  }
  else if (strcmp(name, "band_pitch") == 0) *value = ctx->have_cfg ? ctx->k1.nbp : 0;
  else if (strcmp(name, "pipeline_rows") == 0) *value = ctx->pipeline_rows;
  else if (strcmp(name, "small_rows") == 0) *value = ctx->small_rows;
  else if (strcmp(name, "zero_copy_rows") == 0) *value = ctx->zero_copy_rows;
  else if (strcmp(name, "ens_fused") == 0) *value = ctx->ens_fused;
  else if (strcmp(name, "ens_persist") == 0) *value = ctx->ens_persist;
  else if (strcmp(name, "ens_persist_rows") == 0) *value = ctx->ens_persist_rows;
  else if (strcmp(name, "ens_last_persist") == 0) *value = ctx->ens_last_persist;
  else if (strcmp(name, "ens_split") == 0) *value = ctx->ens_split;
  else if (strcmp(name, "k2_shared_window") == 0) *value = ctx->k2_shared_window;
  else if (strcmp(name, "k2_mma") == 0) *value = ctx->k2_mma;
  else if (strcmp(name, "k2_last_kernel") == 0) *value = ctx->k2_last_kernel;
  else if (strcmp(name, "k2_last_mflops") == 0) *value = (int64_t)(ctx->k2_last_flops * 1e-6);
  else if (strcmp(name, "ens_source_offset") == 0) *value = ctx->ens_src_offset;
  else return fail(ctx, "unknown option '%s'", name);
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_debug_exp10(agnf_ctx* ctx, const double* x, int64_t n, double* y) {
  AGNF_GUARD_BEGIN
  if (!ctx || !x || !y || n < 0) return 1;
  if (n == 0) return 0;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  void *dx = nullptr, *dy = nullptr;
  if (dev_alloc(ctx, &dx, (size_t)n * 8) || dev_alloc(ctx, &dy, (size_t)n * 8)) return 1;
  cudaError_t e = cudaMemcpyAsync(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = launch_exp10_fast(static_cast<const double*>(dx), static_cast<double*>(dy), n, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(y, dy, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, dx);
  dev_free(ctx, dy);
  if (e != cudaSuccess) return fail(ctx, "agnf_debug_exp10: %s", cudaGetErrorString(e));
  ctx->launches += 1;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_debug_fp64_peak(agnf_ctx* ctx, int32_t kind, double* tflops) {
  AGNF_GUARD_BEGIN
  if (!ctx || !tflops) return 1;
  if (kind != 0 && kind != 1) return fail(ctx, "agnf_debug_fp64_peak: kind must be 0 (DFMA) or 1 (DMMA)");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  const int blocks = prop.multiProcessorCount * 8;
  void* scratch = nullptr;
  if (dev_alloc(ctx, &scratch, (size_t)blocks * 256 * 8)) return 1;
  double flops = 0.0, best = 0.0;
  cudaError_t e = cudaSuccess;
  for (int rep = 0; rep < 4 && e == cudaSuccess; ++rep) {            // first repetition warms up
    e = cudaEventRecord(ctx->ev2, ctx->stream);
    if (e == cudaSuccess) e = launch_fp64_peak(kind, static_cast<double*>(scratch), blocks, 20000, &flops, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev3, ctx->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(ctx->ev3);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
    if (e == cudaSuccess && rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  dev_free(ctx, scratch);
  if (e != cudaSuccess) return fail(ctx, "agnf_debug_fp64_peak: %s", cudaGetErrorString(e));
  ctx->launches += 4;
  *tflops = best;
  return 0;
  AGNF_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------- K1

AGNF_API int agnf_modelset_config(agnf_ctx* ctx, const agnf_config* cfg) {
  AGNF_GUARD_BEGIN
  if (!ctx || !cfg) return 1;
  if (cfg->n_b < 1 || cfg->n_b > AGNF_MAX_BANDS) return fail(ctx, "n_b=%d outside [1,%d]", cfg->n_b, AGNF_MAX_BANDS);
  if (cfg->n_par < 3 || cfg->n_par > AGNF_MAX_PARS) return fail(ctx, "n_par=%d outside [3,%d]", cfg->n_par, AGNF_MAX_PARS);
  if (cfg->n_fam != 4 && cfg->n_fam != 5) return fail(ctx, "n_fam must be 4 or 5");
  if ((cfg->n_fam == 5) != (cfg->radio != 0)) return fail(ctx, "n_fam == 5 iff radio");
  if (cfg->n_z < 1) return fail(ctx, "n_z must be >= 1");
  if (cfg->eb_mode < 0 || cfg->eb_mode > 2) return fail(ctx, "eb_mode must be 0, 1 or 2");
  if (cfg->bb_mode < 0 || cfg->bb_mode > 3) return fail(ctx, "bb_mode must be 0..3");
  // a new configuration invalidates everything uploaded for the previous one
  ens_teardown(ctx);
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) ctx->fam_set[f] = false;
  ctx->have_z = ctx->have_src = false;
  ctx->grids_dirty = true;
  ctx->k4_cfg = ctx->k4_src = ctx->k4_bands = false;
  ctx->k1.cfg = *cfg;
  const int n_norm = 3 + (cfg->bbb_has_norm ? 1 : 0) + (cfg->radio ? 1 : 0);
  ctx->k1.norm_offset = cfg->n_par - n_norm;
  ctx->k1.nbp = (cfg->n_b + 1) & ~1;
  ctx->k1.coop_line = ctx->coop_line;
  // log(1 / (sqrt(2 pi) sigma)) as PRIORS_AGNfitter.py:429 writes it, in host libm
  ctx->k1.lognorm_s2 = std::log(1.0 / (std::sqrt(2.0 * 3.141592653589793) * 2.0));
  ctx->k1.lognorm_s15 = std::log(1.0 / (std::sqrt(2.0 * 3.141592653589793) * 1.5));
  ctx->k1.lognorm_s01 = std::log(1.0 / (std::sqrt(2.0 * 3.141592653589793) * 0.1));
  ctx->k1.lognorm_s06 = std::log(1.0 / (std::sqrt(2.0 * 3.141592653589793) * 0.6));
  ctx->k1.lognorm_s05 = std::log(1.0 / (std::sqrt(2.0 * 3.141592653589793) * 0.5));
  ctx->k1.xp_c6um = std::pow(10.0, 13.69897);       // (10**13.69897), PRIORS_AGNfitter.py:308
  ctx->k1.xp_midir_uv = ctx->k1.xp_ir_xray = 0;
  ctx->have_cfg = true;
  ctx->eb0_dirty = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_family_upload(agnf_ctx* ctx, const agnf_family_desc* fd) {
  AGNF_GUARD_BEGIN
  if (!ctx || !fd) return 1;
  if (!ctx->have_cfg) return fail(ctx, "call agnf_modelset_config first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int f = fd->fam;
  if (f < 0 || f >= ctx->k1.cfg.n_fam) return fail(ctx, "family id %d out of range", f);
  if (fd->n_dim < 0 || fd->n_dim > AGNF_MAX_DIMS) return fail(ctx, "family %d: n_dim=%d", f, fd->n_dim);
  if (fd->T < 1 || !fd->table) return fail(ctx, "family %d: empty table", f);
  FamDesc& F = ctx->k1.fam[f];
  const agnf_config& cfg = ctx->k1.cfg;
  F.T = fd->T;
  F.n_dim = fd->n_dim;
  F.par_offset = fd->par_offset;
  if (fd->par_offset < 0 || fd->par_offset + fd->n_dim > ctx->k1.norm_offset)
    return fail(ctx, "family %d: parameter block [%d,%d) overlaps the normalisations at %d", f, fd->par_offset,
                fd->par_offset + fd->n_dim, ctx->k1.norm_offset);
  ctx->fam_grid[f].clear();
  ctx->fam_perm[f].clear();
  int64_t max_row = 0;
  int cursor = 0;
  for (int d = 0; d < fd->n_dim; ++d) {
    F.dim[d].is_grid = fd->dim_is_grid[d];
    F.dim[d].size = fd->dim_size[d];
    F.dim[d].stride = fd->dim_stride[d];
    if (fd->dim_size[d] < 1) return fail(ctx, "family %d dim %d: size < 1", f, d);
    for (int k = 0; k < fd->dim_size[d]; ++k) {
      const double v = fd->grid_sorted[cursor + k];
      if (k > 0 && !(v > fd->grid_sorted[cursor + k - 1]))
        return fail(ctx, "family %d dim %d: grid values must be strictly ascending", f, d);
      const int32_t pi = fd->grid_perm[cursor + k];
      if (pi < 0 || pi >= fd->dim_size[d]) return fail(ctx, "family %d dim %d: bad permutation", f, d);
      ctx->fam_grid[f].push_back(v);
      ctx->fam_perm[f].push_back(pi);
    }
    if (fd->dim_is_grid[d]) max_row += (int64_t)(fd->dim_size[d] - 1) * fd->dim_stride[d];
    cursor += fd->dim_size[d];
  }
  if (max_row >= fd->T) return fail(ctx, "family %d: strides address row %lld but T=%lld", f, (long long)max_row,
                                    (long long)fd->T);
  if (fd->table_on_device) {
    const int pitch = fd->table_pitch > 0 ? fd->table_pitch : cfg.n_b;
    if (pitch < cfg.n_b) return fail(ctx, "family %d: table_pitch %d < n_b %d", f, pitch, cfg.n_b);
    F.table = fd->table;
    F.pitch = pitch;
  } else {
    if (fd->table_pitch > 0 && fd->table_pitch != cfg.n_b)
      return fail(ctx, "family %d: host tables must be dense (table_pitch 0 or n_b)", f);
    const double* old = nullptr;   // previous upload of this family (if any) stays owned until destroy
    if (upload_padded(ctx, &old, fd->table, (size_t)cfg.n_z * (size_t)fd->T, cfg.n_b, ctx->k1.nbp)) return 1;
    F.table = old;
    F.pitch = ctx->k1.nbp;
  }
  F.aux = nullptr;
  if (fd->aux) {
    const double* a = nullptr;
    if (upload(ctx, &a, fd->aux, (size_t)fd->T)) return 1;
    F.aux = a;
  }
  // gather4 kernel: a tensor map over the table (failure only disables that kernel)
  k1_g4_make_map(ctx->maps, f, F.table, (int64_t)cfg.n_z * F.T, F.pitch);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->fam_set[f] = true;
  ctx->grids_dirty = true;
  if (f == 0) ctx->eb0_dirty = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_extra_priors_set(agnf_ctx* ctx, const agnf_extra_priors* xp) {
  AGNF_GUARD_BEGIN
  if (!ctx || !xp) return 1;
  if (!ctx->have_cfg) return fail(ctx, "call agnf_modelset_config first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K1Params& P = ctx->k1;
  const bool any = xp->midir_uv || xp->ir_xray;
  if (any && (xp->j_6um < 0 || xp->j_6um >= P.cfg.n_b))
    return fail(ctx, "j_6um=%d outside the band set: the reference raises IndexError without a band at 6 um "
                     "(PRIORS_AGNfitter.py:306,331)", xp->j_6um);
  if (any && P.n_const < 4 && ctx->have_src) return fail(ctx, "sources were uploaded with n_const < 4");
  if (xp->midir_uv) {
    if (xp->j_2500 < 0 || xp->j_2500 >= P.cfg.n_b)
      return fail(ctx, "j_2500=%d outside the band set (PRIORS_AGNfitter.py:347)", xp->j_2500);
    if (!ctx->fam_set[3] || !ctx->fam_set[2]) return fail(ctx, "upload the torus and disc families first");
    if (!xp->bb_dered_row) return fail(ctx, "prior_midIR_UV needs bb_dered_row");
    for (int64_t t = 0; t < P.fam[3].T; ++t)
      if (xp->bb_dered_row[t] < 0 || xp->bb_dered_row[t] >= P.fam[3].T)
        return fail(ctx, "bb_dered_row[%lld] out of range", (long long)t);
    if (upload(ctx, &P.xp_bb_dered, xp->bb_dered_row, (size_t)P.fam[3].T)) return 1;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  P.xp_midir_uv = xp->midir_uv ? 1 : 0;
  P.xp_ir_xray = xp->ir_xray ? 1 : 0;
  P.xp_j6 = xp->j_6um;
  P.xp_j2500 = xp->j_2500;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_zslices_upload(agnf_ctx* ctx, const double* cgal, const double* cbbb, const uint64_t* xray_bits,
                        const int32_t* j1500, const double* omega) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_cfg) return fail(ctx, "call agnf_modelset_config first");
  if (!cgal || !cbbb || !xray_bits || !j1500 || !omega) return fail(ctx, "agnf_zslices_upload: NULL argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t nz = ctx->k1.cfg.n_z, nb = ctx->k1.cfg.n_b;
  for (size_t i = 0; i < nz; ++i)
    if (ctx->k1.cfg.prior_agnfrac && (j1500[i] < 0 || j1500[i] >= (int)nb))
      return fail(ctx, "z-slice %zu: no band with 14.3 < log nu_rest < 15.3; prior_AGNfraction would raise "
                       "IndexError in the reference (PRIORS_AGNfitter.py:166)", i);
  if (upload_padded(ctx, &ctx->k1.cgal, cgal, nz, (int)nb, ctx->k1.nbp)) return 1;
  if (upload_padded(ctx, &ctx->k1.cbbb, cbbb, nz, (int)nb, ctx->k1.nbp)) return 1;
  {
    const double k64 = (double)AGNF_POW2_DENOM * 3.321928094887362;      // D log2(10)
    std::vector<double> a(nz * nb), b(nz * nb);
    for (size_t i = 0; i < nz * nb; ++i) {
      a[i] = cgal[i] * k64;
      b[i] = cbbb[i] * k64;
    }
    if (upload_padded(ctx, &ctx->k1.cgal64, a.data(), nz, (int)nb, ctx->k1.nbp)) return 1;
    if (upload_padded(ctx, &ctx->k1.cbbb64, b.data(), nz, (int)nb, ctx->k1.nbp)) return 1;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (upload(ctx, &ctx->k1.xray, xray_bits, nz)) return 1;
  if (upload(ctx, &ctx->k1.j1500, j1500, nz)) return 1;
  if (upload_padded(ctx, &ctx->k1.omega, omega, nz, (int)nb, ctx->k1.nbp)) return 1;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->have_z = true;
  ctx->eb0_dirty = true;
  ctx->srcrec_dirty = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_sources_upload(agnf_ctx* ctx, int64_t S, const int32_t* zi, const double* y, const double* inv_sigma,
                        const uint64_t* valid_bits, const double* pmin, const double* pmax,
                        const double* consts, int32_t n_const, const int32_t* agn_branch,
                        const int32_t* irsyn_mode, const int32_t* irsyn_j) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_cfg) return fail(ctx, "call agnf_modelset_config first");
  if (S < 1) return fail(ctx, "S must be >= 1");
  if (!zi || !y || !inv_sigma || !valid_bits || !pmin || !pmax || !consts || !agn_branch || !irsyn_mode || !irsyn_j)
    return fail(ctx, "agnf_sources_upload: NULL argument");
  if (n_const < 1) return fail(ctx, "n_const must be >= 1");
  if ((ctx->k1.xp_midir_uv || ctx->k1.xp_ir_xray) && n_const < 4)
    return fail(ctx, "prior_midIR_UV / prior_IR_XRays need consts[1..3] (n_const >= 4)");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ens_teardown(ctx);          // walkers belong to the previous set of sources
  const size_t nb = ctx->k1.cfg.n_b, np = ctx->k1.cfg.n_par, s = (size_t)S;
  for (size_t i = 0; i < s; ++i)
    if (zi[i] < 0 || zi[i] >= ctx->k1.cfg.n_z) return fail(ctx, "source %zu: z-slice %d out of range", i, zi[i]);
  if (upload(ctx, &ctx->k1.zi, zi, s)) return 1;
  if (upload_padded(ctx, &ctx->k1.y, y, s, (int)nb, ctx->k1.nbp)) return 1;
  if (upload_padded(ctx, &ctx->k1.isig, inv_sigma, s, (int)nb, ctx->k1.nbp)) return 1;
  if (upload(ctx, &ctx->k1.valid, valid_bits, s)) return 1;
  if (upload(ctx, &ctx->k1.pmin, pmin, s * np)) return 1;
  if (upload(ctx, &ctx->k1.pmax, pmax, s * np)) return 1;
  {
    std::vector<double2> box(s * np);
    for (size_t i = 0; i < s * np; ++i) box[i] = make_double2(pmin[i], pmax[i]);
    if (upload(ctx, &ctx->k1.pbox, box.data(), s * np)) return 1;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // `box` dies at the end of this scope
  }
  if (upload(ctx, &ctx->k1.consts, consts, s * (size_t)n_const)) return 1;
  if (upload(ctx, &ctx->k1.agn_branch, agn_branch, s)) return 1;
  if (upload(ctx, &ctx->k1.irsyn_mode, irsyn_mode, s)) return 1;
  if (upload(ctx, &ctx->k1.irsyn_j, irsyn_j, s)) return 1;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->k1.n_const = n_const;
  ctx->k1.S = S;
  ctx->have_src = true;
  ctx->srcrec_dirty = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_lnprob_batch_device(agnf_ctx* ctx, const double* d_params, const int32_t* d_source_id, int64_t n,
                             double* d_lnp, int32_t* d_rows, double* d_ymodel, double* d_lnprior,
                             double* d_lnlike, void* stream) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (n < 0) return fail(ctx, "n < 0");
  if (n == 0) return 0;
  if (!d_params || !d_source_id || !d_lnp) return fail(ctx, "agnf_lnprob_batch_device: NULL argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (check_ready(ctx)) return 1;
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
  K1Batch B{d_params, d_source_id, n, d_lnp, d_rows, d_ymodel, d_lnprior, d_lnlike};
  const bool timed = !ctx->untimed_once;          // latency-bound callers skip the two event records (~3 us each)
  ctx->untimed_once = false;
  if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, st));
  CUDA_TRY(ctx, k1_dispatch(ctx, B, st, &ctx->last_k1_kernel));
  if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, st));
  if (timed) ctx->k1_timed = true;
  ctx->launches += 1;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_lnprob_batch(agnf_ctx* ctx, const double* params, const int32_t* source_id, int64_t n, double* lnp,
                      int32_t* rows, double* ymodel, double* lnprior, double* lnlike) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (n < 0) return fail(ctx, "n < 0");
  if (n == 0) return 0;
  if (!params || !source_id || !lnp) return fail(ctx, "agnf_lnprob_batch: NULL argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (check_ready(ctx)) return 1;
  const size_t nb = ctx->k1.cfg.n_b, np = ctx->k1.cfg.n_par, nf = ctx->k1.cfg.n_fam, N = (size_t)n;
  if (!rows && !ymodel && !lnprior && !lnlike && ctx->small_rows > 0 && n <= ctx->small_rows) {
    // what a sampler sends per call (emcee: one half-ensemble; ultranest: a few hundred points)
    const size_t cap = (size_t)ctx->small_rows;
    const size_t off_sid = cap * AGNF_MAX_PARS * 8, off_out = off_sid + cap * 4, bytes = off_out + cap * 8;
    if (!ctx->pin) {
      CUDA_TRY(ctx, cudaHostAlloc(&ctx->pin, bytes, cudaHostAllocMapped));
      CUDA_TRY(ctx, cudaHostGetDevicePointer(&ctx->pin_map, ctx->pin, 0));
      if (dev_alloc(ctx, &ctx->pin_dev, bytes)) return 1;
    }
    char* h = static_cast<char*>(ctx->pin);
    char* d = static_cast<char*>(ctx->pin_dev);
    // pack [params | source_id] contiguously so that ONE copy carries both (lnp comes back from the tail)
    const size_t pbytes = N * np * 8, pb16 = (pbytes + 15) & ~(size_t)15;
    memcpy(h, params, pbytes);
    memcpy(h + pb16, source_id, N * 4);
    cudaStream_t s1 = ctx->stream;
    if (n <= ctx->zero_copy_rows) {
      // a sampler's half-ensemble: the kernel reads the pinned block through its device mapping and writes the
      // results back into it -- one launch and one synchronisation instead of copy, launch, copy, synchronise
      char* m = static_cast<char*>(ctx->pin_map);
      ctx->untimed_once = true;
      if (agnf_lnprob_batch_device(ctx, reinterpret_cast<const double*>(m), reinterpret_cast<const int32_t*>(m + pb16),
                                   n, reinterpret_cast<double*>(m + off_out), nullptr, nullptr, nullptr, nullptr, s1))
        return 1;
      CUDA_TRY(ctx, cudaStreamSynchronize(s1));
      memcpy(lnp, h + off_out, N * 8);
      return 0;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(d, h, pb16 + N * 4, cudaMemcpyHostToDevice, s1));
    if (agnf_lnprob_batch_device(ctx, reinterpret_cast<const double*>(d), reinterpret_cast<const int32_t*>(d + pb16), n,
                                 reinterpret_cast<double*>(d + off_out), nullptr, nullptr, nullptr, nullptr, s1))
      return 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(h + off_out, d + off_out, N * 8, cudaMemcpyDeviceToHost, s1));
    CUDA_TRY(ctx, cudaStreamSynchronize(s1));
    memcpy(lnp, h + off_out, N * 8);
    return 0;
  }
  if (ensure(ctx, &ctx->st_params, &ctx->st_params_cap, N * np * sizeof(double))) return 1;
  if (ensure(ctx, &ctx->st_sid, &ctx->st_sid_cap, N * sizeof(int32_t))) return 1;
  // output staging: lnp | lnprior | lnlike | ymodel | rows
  size_t off_lnp = 0, off_pr = N * 8, off_ll = 2 * N * 8, off_ym = 3 * N * 8;
  size_t off_rows = off_ym + (ymodel ? N * nb * 8 : 0);
  size_t total = off_rows + (rows ? N * nf * 4 : 0);
  if (ensure(ctx, &ctx->st_out, &ctx->st_out_cap, total)) return 1;
  char* o = static_cast<char*>(ctx->st_out);
  cudaStream_t st = ctx->stream;
  const bool plain = !rows && !ymodel && !lnprior && !lnlike;
  if (plain && ctx->pipeline_rows > 0 && n > ctx->pipeline_rows) {
    // Production shape (lnp only): cut the batch into chunks and rotate them over three streams, so
    // that chunk c's upload, chunk c-1's kernel and chunk c-2's download run at the same time.  Chunk
    // boundaries are multiples of 256 rows: every chunk's params block stays 16-byte aligned and CTA
    // tiles never straddle chunks, so results are bit-identical to the one-shot path.
    CUDA_TRY(ctx, cudaStreamSynchronize(st));             // uploads queued on the ctx stream are visible
    const int64_t chunk = (ctx->pipeline_rows + 255) & ~(int64_t)255;
    int k = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk, ++k) {
      const int64_t m = std::min(chunk, n - r0);
      cudaStream_t ps = ctx->pipe[k % 3];
      double* dp = static_cast<double*>(ctx->st_params) + r0 * np;
      int32_t* ds = static_cast<int32_t*>(ctx->st_sid) + r0;
      double* dl = reinterpret_cast<double*>(o + off_lnp) + r0;
      CUDA_TRY(ctx, cudaMemcpyAsync(dp, params + r0 * np, (size_t)m * np * 8, cudaMemcpyHostToDevice, ps));
      CUDA_TRY(ctx, cudaMemcpyAsync(ds, source_id + r0, (size_t)m * 4, cudaMemcpyHostToDevice, ps));
      if (agnf_lnprob_batch_device(ctx, dp, ds, m, dl, nullptr, nullptr, nullptr, nullptr, ps)) return 1;
      CUDA_TRY(ctx, cudaMemcpyAsync(lnp + r0, dl, (size_t)m * 8, cudaMemcpyDeviceToHost, ps));
    }
    for (int i = 0; i < 3; ++i) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->pipe[i]));
    return 0;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_params, params, N * np * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_sid, source_id, N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (agnf_lnprob_batch_device(ctx, static_cast<const double*>(ctx->st_params),
                               static_cast<const int32_t*>(ctx->st_sid), n,
                               reinterpret_cast<double*>(o + off_lnp),
                               rows ? reinterpret_cast<int32_t*>(o + off_rows) : nullptr,
                               ymodel ? reinterpret_cast<double*>(o + off_ym) : nullptr,
                               lnprior ? reinterpret_cast<double*>(o + off_pr) : nullptr,
                               lnlike ? reinterpret_cast<double*>(o + off_ll) : nullptr, st))
    return 1;
  CUDA_TRY(ctx, cudaMemcpyAsync(lnp, o + off_lnp, N * 8, cudaMemcpyDeviceToHost, st));
  if (lnprior) CUDA_TRY(ctx, cudaMemcpyAsync(lnprior, o + off_pr, N * 8, cudaMemcpyDeviceToHost, st));
  if (lnlike) CUDA_TRY(ctx, cudaMemcpyAsync(lnlike, o + off_ll, N * 8, cudaMemcpyDeviceToHost, st));
  if (ymodel) CUDA_TRY(ctx, cudaMemcpyAsync(ymodel, o + off_ym, N * nb * 8, cudaMemcpyDeviceToHost, st));
  if (rows) CUDA_TRY(ctx, cudaMemcpyAsync(rows, o + off_rows, N * nf * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  return 0;
  AGNF_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------- K2

AGNF_API int agnf_filters_upload(agnf_ctx* ctx, int32_t n_b, const double* band_lambda_c, const int64_t* offset,
                        const double* lambda_angstrom, const double* factor) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (n_b < 1 || n_b > AGNF_MAX_BANDS) return fail(ctx, "n_b=%d outside [1,%d]", n_b, AGNF_MAX_BANDS);
  if (!band_lambda_c || !offset || !lambda_angstrom || !factor) return fail(ctx, "agnf_filters_upload: NULL argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t N = offset[n_b];
  std::vector<double> lam((size_t)N), wn((size_t)N), scale((size_t)n_b);
  for (int b = 0; b < n_b; ++b) {
    const int64_t o0 = offset[b], o1 = offset[b + 1];
    const int64_t nk = o1 - o0;
    if (nk < 2) return fail(ctx, "band %d has %lld samples; need >= 2 for a trapezoid", b, (long long)nk);
    const double* x = lambda_angstrom + o0;
    const double* T = factor + o0;
    // trapz(T, x) = sum d_i (T_i + T_{i+1}) / 2 ; node weights w_i = (d_{i-1} + d_i) / 2
    double den = 0.0;
    std::vector<double> w((size_t)nk, 0.0);
    for (int64_t i = 0; i + 1 < nk; ++i) {
      const double d = x[i + 1] - x[i];
      den += d * (T[i + 1] + T[i]) / 2.0;
      w[(size_t)i] += d / 2.0;
      w[(size_t)i + 1] += d / 2.0;
    }
    std::vector<int64_t> order((size_t)nk);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t c) { return x[a] < x[c]; });
    for (int64_t i = 0; i < nk; ++i) {
      lam[(size_t)(o0 + i)] = x[order[(size_t)i]];
      wn[(size_t)(o0 + i)] = w[(size_t)order[(size_t)i]] * T[order[(size_t)i]];
    }
    const double lc = band_lambda_c[b];
    scale[(size_t)b] = (lc * lc) / 2.997e8 / 1.e10 / den;   // fluxlambda_2_fluxnu, DICTIONARIES:486-491
  }
  K2Filters& F = ctx->filt;
  if (upload(ctx, &F.lam, lam.data(), (size_t)N)) return 1;
  if (upload(ctx, &F.wn, wn.data(), (size_t)N)) return 1;
  if (upload(ctx, &F.off, offset, (size_t)n_b + 1)) return 1;
  if (upload(ctx, &F.scale, scale.data(), (size_t)n_b)) return 1;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  F.n_b = n_b;
  F.n_samples = N;
  ctx->have_filt = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_dict_build(agnf_ctx* ctx, const double* lam_sorted, const int32_t* perm, const double* Fnu, int64_t T,
                    int32_t n_lam, const double* z, int32_t n_z, const int64_t* row_map, int64_t T_total,
                    double* out, int32_t out_on_device, int32_t out_pitch) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_filt) return fail(ctx, "call agnf_filters_upload first");
  if (!lam_sorted || !perm || !Fnu || !z || !out) return fail(ctx, "agnf_dict_build: NULL argument");
  if (T < 1 || n_lam < 1 || n_z < 1) return fail(ctx, "agnf_dict_build: empty problem");
  if (!row_map) T_total = T;
  if (T_total < T) return fail(ctx, "T_total < T");
  if (out_pitch == 0) out_pitch = ctx->filt.n_b;
  if (out_pitch < ctx->filt.n_b) return fail(ctx, "out_pitch %d < n_b %d", out_pitch, ctx->filt.n_b);
  if (!out_on_device && out_pitch != ctx->filt.n_b) return fail(ctx, "host outputs must be dense (out_pitch 0 or n_b)");
  for (int j = 1; j < n_lam; ++j)
    if (!(lam_sorted[j] >= lam_sorted[j - 1])) return fail(ctx, "lam_sorted must be ascending");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const auto t_entry = std::chrono::steady_clock::now();
  auto stamp = [&](int i) {
    ctx->k2_prof[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_entry).count();
  };
  for (double& v : ctx->k2_prof) v = 0.0;
  const int n_b = ctx->filt.n_b;
  K2Problem p;
  p.filt = ctx->filt;
  p.T = T;
  p.n_lam = n_lam;
  p.ldF = n_lam | 1;
  p.n_z = n_z;
  p.T_total = T_total;
  p.out_pitch = out_pitch;

  void *d_lam = nullptr, *d_perm = nullptr, *d_F = nullptr, *d_Fs = nullptr, *d_z = nullptr, *d_jlo = nullptr,
       *d_cnt = nullptr, *d_rmax = nullptr, *d_W = nullptr, *d_rowmap = nullptr, *d_out = nullptr, *d_woff = nullptr,
       *d_bandof = nullptr, *d_border = nullptr, *d_jb = nullptr, *d_cb = nullptr, *d_rmaxA = nullptr,
       *d_woffA = nullptr, *d_bandofA = nullptr, *d_WA = nullptr, *d_jb8 = nullptr, *d_cb8 = nullptr,
       *d_rmax8 = nullptr, *d_Fs8 = nullptr, *d_woff8 = nullptr, *d_bandof8 = nullptr, *d_WA8 = nullptr, *d_flag = nullptr;
  auto cleanup = [&]() {
    for (void* q : {d_lam, d_perm, d_F, d_Fs, d_z, d_jlo, d_cnt, d_rmax, d_W, d_rowmap, d_woff, d_bandof, d_border,
                    d_jb, d_cb, d_rmaxA, d_woffA, d_bandofA, d_WA, d_jb8, d_cb8, d_rmax8, d_Fs8, d_woff8, d_bandof8,
                    d_WA8, d_flag})
      scratch_free(ctx, q);
    if (!out_on_device) scratch_free(ctx, d_out);
  };
#define K2_TRY(expr)                                                                                    \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      cleanup();                                                                                        \
      return fail(ctx, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);    \
    }                                                                                                   \
  } while (0)
#define K2_ALLOC(ptr, bytes)                 \
  do {                                       \
    if (scratch_alloc(ctx, &(ptr), (bytes))) {\
      cleanup();                             \
      return 1;                              \
    }                                        \
  } while (0)

  K2_ALLOC(d_lam, (size_t)n_lam * 8);
  K2_ALLOC(d_perm, (size_t)n_lam * 4);
  K2_ALLOC(d_F, (size_t)T * n_lam * 8);
  const int64_t T_pad = (T + 31) & ~(int64_t)31;          // whole 32-template tiles for the tiled contraction
  K2_ALLOC(d_Fs, (size_t)T_pad * p.ldF * 8);
  K2_TRY(cudaMemsetAsync(d_Fs, 0, (size_t)T_pad * p.ldF * 8, st));
  K2_ALLOC(d_z, (size_t)n_z * 8);
  K2_ALLOC(d_jlo, (size_t)n_z * n_b * 4);
  K2_ALLOC(d_cnt, (size_t)n_z * n_b * 4);
  K2_ALLOC(d_rmax, (size_t)n_b * 4);
  K2_TRY(cudaMemcpyAsync(d_lam, lam_sorted, (size_t)n_lam * 8, cudaMemcpyHostToDevice, st));
  K2_TRY(cudaMemcpyAsync(d_perm, perm, (size_t)n_lam * 4, cudaMemcpyHostToDevice, st));
  K2_TRY(cudaMemcpyAsync(d_F, Fnu, (size_t)T * n_lam * 8, cudaMemcpyHostToDevice, st));
  K2_TRY(cudaMemcpyAsync(d_z, z, (size_t)n_z * 8, cudaMemcpyHostToDevice, st));
  K2_TRY(cudaMemsetAsync(d_rmax, 0, (size_t)n_b * 4, st));
  K2_ALLOC(d_flag, 8);
  K2_TRY(cudaMemsetAsync(d_flag, 0, 8, st));
  if (row_map) {
    for (int64_t t = 0; t < T; ++t)
      if (row_map[t] < 0 || row_map[t] >= T_total) {
        cleanup();
        return fail(ctx, "row_map[%lld] out of range", (long long)t);
      }
    K2_ALLOC(d_rowmap, (size_t)T * 8);
    K2_TRY(cudaMemcpyAsync(d_rowmap, row_map, (size_t)T * 8, cudaMemcpyHostToDevice, st));
  }
  p.lam0 = static_cast<const double*>(d_lam);
  p.Fs = static_cast<const double*>(d_Fs);
  p.z = static_cast<const double*>(d_z);
  p.jlo = static_cast<int32_t*>(d_jlo);
  p.cnt = static_cast<int32_t*>(d_cnt);
  p.d_rmaxb = static_cast<int32_t*>(d_rmax);
  p.row_map = static_cast<const int64_t*>(d_rowmap);

  stamp(0);          // allocations + uploads enqueued
  K2_TRY(launch_k2_permute(static_cast<const double*>(d_F), static_cast<const int32_t*>(d_perm),
                           static_cast<double*>(d_Fs), T, n_lam, p.ldF, static_cast<int32_t*>(d_flag), st));
  K2_TRY(launch_k2_ranges(p, st));
  // compact weight layout: band b owns max_z cnt[z][b] slots of every z-slice
  std::vector<int32_t> rmaxb((size_t)n_b), woff((size_t)n_b + 1, 0);
  int32_t nonfinite = 0;
  K2_TRY(cudaMemcpyAsync(rmaxb.data(), d_rmax, (size_t)n_b * 4, cudaMemcpyDeviceToHost, st));
  K2_TRY(cudaMemcpyAsync(&nonfinite, d_flag, 4, cudaMemcpyDeviceToHost, st));
  K2_TRY(cudaStreamSynchronize(st));
  stamp(1);          // uploads, permutation and ranges done
  for (int b = 0; b < n_b; ++b)     // even slot counts: every band's weight row starts 16-byte aligned
    woff[(size_t)b + 1] = woff[(size_t)b] + ((std::max(rmaxb[(size_t)b], 1) + 1) & ~1);
  const int32_t wstride = (woff[(size_t)n_b] + 1) & ~1;       // even: a z-slice of weights is a 16-byte multiple
  std::vector<int32_t> band_of((size_t)wstride, -1);
  for (int b = 0; b < n_b; ++b)
    for (int32_t k = woff[(size_t)b]; k < woff[(size_t)b + 1]; ++k) band_of[(size_t)k] = b;
  K2_ALLOC(d_woff, ((size_t)n_b + 1) * 4);
  K2_ALLOC(d_bandof, (size_t)wstride * 4);
  K2_TRY(cudaMemcpyAsync(d_woff, woff.data(), ((size_t)n_b + 1) * 4, cudaMemcpyHostToDevice, st));
  K2_TRY(cudaMemcpyAsync(d_bandof, band_of.data(), (size_t)wstride * 4, cudaMemcpyHostToDevice, st));
  std::vector<int32_t> border((size_t)n_b);
  std::iota(border.begin(), border.end(), 0);
  std::stable_sort(border.begin(), border.end(),
                   [&](int32_t x, int32_t y) { return rmaxb[(size_t)x] > rmaxb[(size_t)y]; });
  K2_ALLOC(d_border, (size_t)n_b * 4);
  K2_TRY(cudaMemcpyAsync(d_border, border.data(), (size_t)n_b * 4, cudaMemcpyHostToDevice, st));
  p.band_order = static_cast<const int32_t*>(d_border);
  p.wstride = wstride;
  p.woff = static_cast<const int32_t*>(d_woff);
  p.band_of = static_cast<const int32_t*>(d_bandof);
  p.tile_ok = 1;
  const size_t out_count = (size_t)n_z * (size_t)T_total * (size_t)n_b;
  if (out_on_device) {
    d_out = out;
  } else {
    K2_ALLOC(d_out, out_count * 8);
    if (row_map) K2_TRY(cudaMemcpyAsync(d_out, out, out_count * 8, cudaMemcpyHostToDevice, st));  // keep other rows
  }
  p.out = static_cast<double*>(d_out);
  // shared-window plan: 4 consecutive slices share each read of the template window.  A non-finite template
  // sample would leak into the neighbouring slices through 0 * inf, so such families take the per-slice kernels.
  stamp(2);
  const bool finite = ctx->k2_shared_window != 0 && !nonfinite;      // scanned on the device by the permutation kernel
  stamp(3);
  if (finite && ctx->k2_mma) {
    // tensor-core plan: batches of 8 slices, windows aligned to multiples of 4 samples, templates re-laid out
    // with a row pitch = 4 (mod 8) (conflict-free 8 x 4 fragments) and >= n_lam + 3 (windows are rounded up)
    p.n_zb8 = (n_z + 7) / 8;
    K2_ALLOC(d_jb8, (size_t)p.n_zb8 * n_b * 4);
    K2_ALLOC(d_cb8, (size_t)p.n_zb8 * n_b * 4);
    K2_ALLOC(d_rmax8, (size_t)n_b * 4);
    K2_TRY(cudaMemsetAsync(d_rmax8, 0, (size_t)n_b * 4, st));
    p.jb8 = static_cast<int32_t*>(d_jb8);
    p.cb8 = static_cast<int32_t*>(d_cb8);
    p.d_rmaxb8 = static_cast<int32_t*>(d_rmax8);
    K2_TRY(launch_k2_batch_ranges8(p, st));
    std::vector<int32_t> rmax8((size_t)n_b), woff8((size_t)n_b + 1, 0);
    K2_TRY(cudaMemcpyAsync(rmax8.data(), d_rmax8, (size_t)n_b * 4, cudaMemcpyDeviceToHost, st));
    K2_TRY(cudaStreamSynchronize(st));
    for (int b = 0; b < n_b; ++b) woff8[(size_t)b + 1] = woff8[(size_t)b] + std::max(rmax8[(size_t)b], 4);
    p.wstride8 = woff8[(size_t)n_b];
    p.hdr8 = ((2 * n_b * 4 + 15) / 16 * 16) / 8;
    p.blk8 = (int64_t)p.hdr8 + (int64_t)p.wstride8 * 8;
    int32_t L = n_lam + 3;
    L += ((4 - (L & 7)) + 8) & 7;
    p.ldF8 = L;
    p.mma_ok = 1;
    if (k2_mma_fits(p)) {
      K2_ALLOC(d_Fs8, (size_t)T_pad * p.ldF8 * 8);
      K2_TRY(cudaMemsetAsync(d_Fs8, 0, (size_t)T_pad * p.ldF8 * 8, st));
      K2_TRY(launch_k2_permute(static_cast<const double*>(d_F), static_cast<const int32_t*>(d_perm),
                               static_cast<double*>(d_Fs8), T, n_lam, p.ldF8, nullptr, st));
      p.Fs8 = static_cast<const double*>(d_Fs8);
      std::vector<int32_t> band_of8((size_t)p.wstride8, -1);
      for (int b = 0; b < n_b; ++b)
        for (int32_t k = woff8[(size_t)b]; k < woff8[(size_t)b + 1]; ++k) band_of8[(size_t)k] = b;
      std::stable_sort(border.begin(), border.end(),
                       [&](int32_t x, int32_t y) { return rmax8[(size_t)x] > rmax8[(size_t)y]; });
      K2_TRY(cudaMemcpyAsync(d_border, border.data(), (size_t)n_b * 4, cudaMemcpyHostToDevice, st));
      K2_ALLOC(d_woff8, ((size_t)n_b + 1) * 4);
      K2_ALLOC(d_bandof8, (size_t)p.wstride8 * 4);
      K2_TRY(cudaMemcpyAsync(d_woff8, woff8.data(), ((size_t)n_b + 1) * 4, cudaMemcpyHostToDevice, st));
      K2_TRY(cudaMemcpyAsync(d_bandof8, band_of8.data(), (size_t)p.wstride8 * 4, cudaMemcpyHostToDevice, st));
      p.woff8 = static_cast<const int32_t*>(d_woff8);
      p.band_of8 = static_cast<const int32_t*>(d_bandof8);
      K2_ALLOC(d_WA8, (size_t)p.n_zb8 * (size_t)p.blk8 * 8);
      p.WA8 = static_cast<double*>(d_WA8);
      K2_TRY(launch_k2_weights_aligned8(p, st));
      K2_TRY(cudaStreamSynchronize(st));      // the host vectors above feed asynchronous copies
    } else {
      p.mma_ok = 0;
    }
  }
  if (finite && !p.mma_ok) {
    p.n_zb = (n_z + 3) / 4;
    K2_ALLOC(d_jb, (size_t)p.n_zb * n_b * 4);
    K2_ALLOC(d_cb, (size_t)p.n_zb * n_b * 4);
    K2_ALLOC(d_rmaxA, (size_t)n_b * 4);
    K2_TRY(cudaMemsetAsync(d_rmaxA, 0, (size_t)n_b * 4, st));
    p.jb = static_cast<int32_t*>(d_jb);
    p.cb = static_cast<int32_t*>(d_cb);
    p.d_rmaxbA = static_cast<int32_t*>(d_rmaxA);
    K2_TRY(launch_k2_batch_ranges(p, st));
    std::vector<int32_t> rmaxA((size_t)n_b), woffA((size_t)n_b + 1, 0);
    K2_TRY(cudaMemcpyAsync(rmaxA.data(), d_rmaxA, (size_t)n_b * 4, cudaMemcpyDeviceToHost, st));
    K2_TRY(cudaStreamSynchronize(st));
    for (int b = 0; b < n_b; ++b) woffA[(size_t)b + 1] = woffA[(size_t)b] + std::max(rmaxA[(size_t)b], 1);
    p.wstrideA = woffA[(size_t)n_b];
    p.hdrA = ((2 * n_b * 4 + 15) / 16 * 16) / 8;
    p.blkA = (int64_t)p.hdrA + (int64_t)p.wstrideA * 4;
    p.sw_ok = 1;
    if (k2_sw_fits(p)) {
      std::vector<int32_t> band_ofA((size_t)p.wstrideA, -1);
      for (int b = 0; b < n_b; ++b)
        for (int32_t k = woffA[(size_t)b]; k < woffA[(size_t)b + 1]; ++k) band_ofA[(size_t)k] = b;
      std::stable_sort(border.begin(), border.end(),
                       [&](int32_t x, int32_t y) { return rmaxA[(size_t)x] > rmaxA[(size_t)y]; });
      K2_TRY(cudaMemcpyAsync(d_border, border.data(), (size_t)n_b * 4, cudaMemcpyHostToDevice, st));
      K2_ALLOC(d_woffA, ((size_t)n_b + 1) * 4);
      K2_ALLOC(d_bandofA, (size_t)p.wstrideA * 4);
      K2_TRY(cudaMemcpyAsync(d_woffA, woffA.data(), ((size_t)n_b + 1) * 4, cudaMemcpyHostToDevice, st));
      K2_TRY(cudaMemcpyAsync(d_bandofA, band_ofA.data(), (size_t)p.wstrideA * 4, cudaMemcpyHostToDevice, st));
      p.woffA = static_cast<const int32_t*>(d_woffA);
      p.band_ofA = static_cast<const int32_t*>(d_bandofA);
      K2_ALLOC(d_WA, (size_t)p.n_zb * (size_t)p.blkA * 8);
      p.WA = static_cast<double*>(d_WA);
      K2_TRY(launch_k2_weights_aligned(p, st));
      K2_TRY(cudaStreamSynchronize(st));      // the host vectors above feed asynchronous copies
    } else {
      p.sw_ok = 0;
    }
  }
  if (!p.sw_ok && !p.mma_ok) {
    K2_ALLOC(d_W, (size_t)n_z * wstride * 8);
    p.W = static_cast<double*>(d_W);
    K2_TRY(launch_k2_weights(p, st));
  }
  ctx->k2_last_kernel = p.mma_ok ? 3 : (p.sw_ok ? 2 : 1);
  // flops the contraction executes: every band's (padded) window x templates x slices of a batch
  if (p.mma_ok) ctx->k2_last_flops = 2.0 * (double)p.n_zb8 * 8.0 * (double)p.wstride8 * (double)T_pad;
  else if (p.sw_ok) ctx->k2_last_flops = 2.0 * (double)p.n_zb * 4.0 * (double)p.wstrideA * (double)T_pad;
  else ctx->k2_last_flops = 2.0 * (double)n_z * (double)wstride * (double)T;
  stamp(4);          // weights built
  K2_TRY(cudaEventRecord(ctx->ev2, st));
  K2_TRY(launch_k2_contract(p, st));
  K2_TRY(cudaEventRecord(ctx->ev3, st));
  ctx->k2_timed = true;
  ctx->launches += 4;
  if (!out_on_device) K2_TRY(cudaMemcpyAsync(out, d_out, out_count * 8, cudaMemcpyDeviceToHost, st));
  K2_TRY(cudaStreamSynchronize(st));
  stamp(5);          // contraction done
  cleanup();
  stamp(6);          // scratch released
  return 0;
#undef K2_TRY
#undef K2_ALLOC
  AGNF_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------- K3

AGNF_API int agnf_ens_create(agnf_ctx* ctx, int32_t walkers, uint64_t seed, double a, int32_t n_bins) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (check_ready(ctx)) return 1;
  if (walkers < 2 || (walkers & 1)) return fail(ctx, "walkers must be even and >= 2 (two half-ensembles)");
  if (!(a > 1.0)) return fail(ctx, "stretch scale a must be > 1");
  if (n_bins < 0) return fail(ctx, "n_bins < 0");
  ens_teardown(ctx);
  const int64_t S = ctx->k1.S;
  const int np = ctx->k1.cfg.n_par, H = walkers / 2;
  K3State& st = ctx->ens;
  st.S = S;
  st.W = walkers;
  st.n_par = np;
  st.seed = seed;
  st.a = a;
  st.src_offset = ctx->ens_src_offset;
  void* p = nullptr;
#define ENS_ALLOC(dst, type, count)                             \
  do {                                                          \
    if (dev_alloc(ctx, &p, (size_t)(count) * sizeof(type))) return 1; \
    dst = static_cast<type*>(p);                                \
  } while (0)
  ENS_ALLOC(st.pos, double, S * walkers * np);
  ENS_ALLOC(st.lnp, double, S * walkers);
  ENS_ALLOC(st.n_acc, int32_t, S * walkers);
  ENS_ALLOC(ctx->ens_sc.prop, double, S * H * np);
  ENS_ALLOC(ctx->ens_sc.prop_lnp, double, S * H);
  ENS_ALLOC(ctx->ens_sc.factor, double, S * H);
  ENS_ALLOC(ctx->ens_sc.lnu, double, S * H);
  ENS_ALLOC(ctx->ens_sc.sid, int32_t, S * walkers);       // [S*W] for the initial evaluation, [S*H] prefix per half-step
  ENS_ALLOC(ctx->ens_best_pos, double, S * np);
  ENS_ALLOC(ctx->ens_best_lnp, double, S);
  st.split = nullptr;
  if (ctx->ens_split) {
    if (walkers > 65535) return fail(ctx, "the random split indexes walkers with 16 bits");
    ENS_ALLOC(st.split, uint16_t, S * walkers);
  }
  ctx->ens_hist.hist = nullptr;
  ctx->ens_hist.n_bins = n_bins;
  ctx->ens_hist.pbox = nullptr;            // read from ctx->k1 at launch time
  if (n_bins > 0) {
    ENS_ALLOC(ctx->ens_hist.hist, uint32_t, S * np * n_bins);
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->ens_hist.hist, 0, (size_t)S * np * n_bins * 4, ctx->stream));
  }
#undef ENS_ALLOC
  CUDA_TRY(ctx, cudaMemsetAsync(st.n_acc, 0, (size_t)S * walkers * 4, ctx->stream));
  {
    std::vector<double> ninf((size_t)S, -INFINITY);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ens_best_lnp, ninf.data(), (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->ens_step = 0;
  ctx->ens_hist_samples = 0;
  ctx->have_ens = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

namespace {

// The walkers have new positions: evaluate them.  The acceptance counters and the running best are NOT touched:
// emcee's sampler keeps accumulating over the burn-in sets (the reference never resets it between them,
// MCMC_AGNfitter.py:125-128), get_best_position takes the argmax over everything sampled so far
// (PARAMETERSPACE_AGNfitter.py:281-283), and the starting positions themselves are not part of emcee's chain.
int ens_positions_changed(agnf_ctx* ctx) {
  K3State& st = ctx->ens;
  cudaStream_t s = ctx->stream;
  const int64_t n = st.S * st.W;
  CUDA_TRY(ctx, launch_k3_fill_sid(ctx->ens_sc.sid, st.S, st.W, s));
  if (agnf_lnprob_batch_device(ctx, st.pos, ctx->ens_sc.sid, n, st.lnp, nullptr, nullptr, nullptr, nullptr, s)) return 1;
  CUDA_TRY(ctx, launch_k3_fill_sid(ctx->ens_sc.sid, st.S, st.W / 2, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  ctx->launches += 3;
  return 0;
}

}  // namespace

AGNF_API int agnf_ens_set_positions(agnf_ctx* ctx, const double* pos) {
  AGNF_GUARD_BEGIN
  if (!ctx || !pos) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K3State& st = ctx->ens;
  CUDA_TRY(ctx, cudaMemcpyAsync(st.pos, pos, (size_t)st.S * st.W * st.n_par * 8, cudaMemcpyHostToDevice, ctx->stream));
  return ens_positions_changed(ctx);
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_set_ball(agnf_ctx* ctx, const double* centre, double scale, int32_t normal, uint32_t stream_id) {
  AGNF_GUARD_BEGIN
  if (!ctx || !centre) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K3State& st = ctx->ens;
  void* d_c = nullptr;
  if (dev_alloc(ctx, &d_c, (size_t)st.S * st.n_par * 8)) return 1;
  cudaError_t e = cudaMemcpyAsync(d_c, centre, (size_t)st.S * st.n_par * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess)
    e = launch_k3_ball(st, static_cast<const double*>(d_c), scale, normal ? 1 : 0, stream_id, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, d_c);
  if (e != cudaSuccess) return fail(ctx, "agnf_ens_set_ball failed: %s", cudaGetErrorString(e));
  ctx->launches += 1;
  return ens_positions_changed(ctx);
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_run(agnf_ctx* ctx, int64_t n_steps, int32_t hist_every, int32_t track_best) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  if (n_steps < 0) return fail(ctx, "n_steps < 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K3State& st = ctx->ens;
  cudaStream_t s = ctx->stream;
  const int64_t nh = st.S * (st.W / 2);
  // fused launch (propose -> ln_probab -> accept inside the cooperative K1 kernel) when its preconditions hold
  K1Batch FB{nullptr, nullptr, nh, nullptr, nullptr, nullptr, nullptr, nullptr};
  FB.ens.pos = st.pos;
  FB.ens.lnp = st.lnp;
  FB.ens.n_acc = st.n_acc;
  FB.ens.W = st.W;
  FB.ens.seed = st.seed;
  FB.ens.a = st.a;
  FB.ens.src_offset = st.src_offset;
  FB.ens.split = st.split;
  ctx->ens_hist.pbox = ctx->k1.pbox;
  const bool fused = ctx->ens_fused && k1_fused_available(ctx, FB);
  // Small shards: ONE launch for all n_steps, one CTA per source (k1_coop_kernel<..., PERSIST>): the lock-step path
  // pays ~5 launches per step, which is all a shard of a few hundred sources costs.  Same Philox counters, same
  // arithmetic: the chain is identical bit for bit (tests/test_ensemble_gpu.py).
  const int v = ctx->k1_variant;
  // A half-step of one source is a chain of dependent memory round trips, so its latency falls with the rows a warp
  // owns: up to two CTAs of eight warps fit an SM, which serves shards of <= 296 sources; larger ones keep 32 rows per
  // warp (more CTAs per SM, better throughput).
  int rpw = ctx->ens_persist_rows;
  if (rpw <= 0) {
    rpw = st.S <= 2 * 148 ? std::max(1, ((int)(st.W >> 1) + 7) / 8) : 32;
    if (rpw != 32 && !k1_coop_persist_supported(ctx->k1, FB, rpw)) rpw = 32;
  }
  const bool persist = ctx->ens_fused && v != 0 && v != 3 && n_steps > 0 && ctx->ens_persist != 0 &&
                       k1_coop_persist_supported(ctx->k1, FB, rpw) && (ctx->ens_persist == 1 || st.S <= 1184);
  ctx->ens_last_persist = persist ? 1 : 0;
  if (persist) {
    K1Persist A;
    A.step0 = ctx->ens_step;
    A.n_steps = n_steps;
    A.hist_every = hist_every;
    A.track_best = track_best;
    A.hist = ctx->ens_hist.hist;
    A.pbox = ctx->k1.pbox;
    A.n_bins = ctx->ens_hist.n_bins;
    A.best_pos = ctx->ens_best_pos;
    A.best_lnp = ctx->ens_best_lnp;
    A.S = st.S;
    A.rows_per_warp = rpw;
    int64_t n_rec = 0, n_thin = 0;
    if (hist_every > 0) n_thin = n_steps / hist_every;
    if (ctx->ens_chain) {
      for (int64_t k = 0; k < n_steps; ++k)
        n_rec += ((ctx->ens_step + k - ctx->ens_chain_origin) % ctx->ens_chain_every) == 0;
      if (ctx->ens_chain_n + n_rec > ctx->ens_chain_cap)
        return fail(ctx, "chain buffer full after %lld recorded steps (agnf_ens_record)", (long long)ctx->ens_chain_cap);
      A.chain = ctx->ens_chain;
      A.chain_lnp = ctx->ens_chain_lnp;
      A.chain_n0 = ctx->ens_chain_n;
      A.chain_origin = ctx->ens_chain_origin;
      A.chain_every = ctx->ens_chain_every;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, s));
    CUDA_TRY(ctx, launch_k1_coop_persist(ctx->k1, FB, A, s));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, s));
    ctx->k1_timed = true;
    ctx->last_k1_kernel = 1;
    ctx->launches += 1;
    ctx->ens_step += n_steps;
    ctx->ens_chain_n += n_rec;
    if (ctx->ens_hist.hist) ctx->ens_hist_samples += n_thin * st.W;
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return 0;
  }
  for (int64_t it = 0; it < n_steps; ++it) {
    if (st.split) {
      CUDA_TRY(ctx, launch_k3_split(st, ctx->ens_step, s));
      ctx->launches += 1;
    }
    for (int half = 0; half < 2; ++half) {
      if (fused) {
        FB.ens.half = half;
        FB.ens.step = ctx->ens_step;
        const bool timed = it + 1 == n_steps;           // events around the last step's launches only
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, s));
        CUDA_TRY(ctx, k1_dispatch(ctx, FB, s, &ctx->last_k1_kernel));
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, s));
        ctx->k1_timed = true;
        ctx->launches += 1;
        continue;
      }
      CUDA_TRY(ctx, launch_k3_propose(st, ctx->ens_sc, ctx->ens_step, half, s));
      if (agnf_lnprob_batch_device(ctx, ctx->ens_sc.prop, ctx->ens_sc.sid, nh, ctx->ens_sc.prop_lnp, nullptr, nullptr,
                                   nullptr, nullptr, s))
        return 1;
      CUDA_TRY(ctx, launch_k3_accept(st, ctx->ens_sc, half, s));
      ctx->launches += 2;
    }
    ++ctx->ens_step;
    const bool thinned = hist_every > 0 && ((it + 1) % hist_every) == 0;     // counted from the start of this call
    if (track_best == 1 || (track_best == 2 && thinned)) {
      CUDA_TRY(ctx, launch_k3_best(st, ctx->ens_best_pos, ctx->ens_best_lnp, s));
      ctx->launches += 1;
    }
    if (thinned && ctx->ens_hist.hist) {      // without histograms `hist_every` only marks the thinned steps (track_best 2)
      CUDA_TRY(ctx, launch_k3_hist(st, ctx->ens_hist, s));
      ctx->ens_hist_samples += st.W;
      ctx->launches += 1;
    }
    if (ctx->ens_chain && ((ctx->ens_step - ctx->ens_chain_origin - 1) % ctx->ens_chain_every) == 0) {
      if (ens_append_state(ctx)) {
        cudaStreamSynchronize(s);
        return 1;
      }
    }
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_record_state(agnf_ctx* ctx) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens || !ctx->ens_chain) return fail(ctx, "no recording buffer (agnf_ens_record)");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (ens_append_state(ctx)) return 1;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_chain_percentiles(agnf_ctx* ctx, int64_t first, int64_t stride, const double* q, int32_t n_q,
                                        double* out_pos, double* out_lnp) {
  AGNF_GUARD_BEGIN
  if (!ctx || !q || !out_pos) return 1;
  if (!ctx->have_ens || !ctx->ens_chain) return fail(ctx, "no recorded chain (agnf_ens_record)");
  if (n_q < 1 || n_q > K3_SELECT_MAX_Q) return fail(ctx, "n_q outside [1, %d]", K3_SELECT_MAX_Q);
  for (int k = 0; k < n_q; ++k)
    if (!(q[k] >= 0.0 && q[k] <= 100.0) || (k > 0 && q[k] < q[k - 1]))
      return fail(ctx, "percentiles must be ascending values in [0, 100]");
  if (first < 0 || stride < 1 || first >= ctx->ens_chain_n)
    return fail(ctx, "agnf_ens_chain_percentiles: first = %lld, stride = %lld with %lld recorded entries",
                (long long)first, (long long)stride, (long long)ctx->ens_chain_n);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const K3State& st = ctx->ens;
  void *dq = nullptr, *dout = nullptr, *dout2 = nullptr;
  const size_t n_pos = (size_t)st.S * st.n_par * n_q, n_lnp = (size_t)st.S * n_q;
  if (scratch_alloc(ctx, &dq, (size_t)n_q * 8) || scratch_alloc(ctx, &dout, n_pos * 8) ||
      scratch_alloc(ctx, &dout2, n_lnp * 8)) {
    scratch_free(ctx, dq);
    scratch_free(ctx, dout);
    scratch_free(ctx, dout2);
    return 1;
  }
  cudaStream_t s = ctx->stream;
  cudaError_t e = cudaMemcpyAsync(dq, q, (size_t)n_q * 8, cudaMemcpyHostToDevice, s);
  K3Select a;
  a.S = st.S;
  a.W = st.W;
  a.first = first;
  a.stride = stride;
  a.n_used = (ctx->ens_chain_n - first + stride - 1) / stride;
  a.q = static_cast<const double*>(dq);
  a.n_q = n_q;
  a.data = ctx->ens_chain;
  a.E = st.n_par;
  a.out = static_cast<double*>(dout);
  if (e == cudaSuccess) e = launch_k3_select(a, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_pos, dout, n_pos * 8, cudaMemcpyDeviceToHost, s);
  if (out_lnp) {
    a.data = ctx->ens_chain_lnp;
    a.E = 1;
    a.out = static_cast<double*>(dout2);
    if (e == cudaSuccess) e = launch_k3_select(a, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_lnp, dout2, n_lnp * 8, cudaMemcpyDeviceToHost, s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  scratch_free(ctx, dq);
  scratch_free(ctx, dout);
  scratch_free(ctx, dout2);
  if (e != cudaSuccess) return fail(ctx, "agnf_ens_chain_percentiles: %s", cudaGetErrorString(e));
  ctx->launches += out_lnp ? 2 : 1;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_record(agnf_ctx* ctx, int64_t max_steps, int32_t every) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  if (max_steps < 0 || every < 1) return fail(ctx, "agnf_ens_record: max_steps >= 0 and every >= 1");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  dev_free(ctx, ctx->ens_chain);
  dev_free(ctx, ctx->ens_chain_lnp);
  ctx->ens_chain = nullptr;
  ctx->ens_chain_lnp = nullptr;
  ctx->ens_chain_cap = ctx->ens_chain_n = 0;
  ctx->ens_chain_every = every;
  ctx->ens_chain_origin = ctx->ens_step;
  if (max_steps == 0) return 0;
  const K3State& st = ctx->ens;
  const size_t nw = (size_t)st.S * st.W;
  void *a = nullptr, *b = nullptr;
  if (dev_alloc(ctx, &a, (size_t)max_steps * nw * st.n_par * 8)) return 1;
  if (dev_alloc(ctx, &b, (size_t)max_steps * nw * 8)) {
    dev_free(ctx, a);
    return 1;
  }
  ctx->ens_chain = static_cast<double*>(a);
  ctx->ens_chain_lnp = static_cast<double*>(b);
  ctx->ens_chain_cap = max_steps;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_chain(agnf_ctx* ctx, double* chain, double* lnp, int64_t* n_recorded) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const K3State& st = ctx->ens;
  const size_t nw = (size_t)st.S * st.W, n = (size_t)ctx->ens_chain_n;
  if (n_recorded) *n_recorded = ctx->ens_chain_n;
  if (chain && n)
    CUDA_TRY(ctx, cudaMemcpyAsync(chain, ctx->ens_chain, n * nw * st.n_par * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (lnp && n) CUDA_TRY(ctx, cudaMemcpyAsync(lnp, ctx->ens_chain_lnp, n * nw * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_get(agnf_ctx* ctx, double* pos, double* lnp, int32_t* n_acc) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K3State& st = ctx->ens;
  const size_t n = (size_t)st.S * st.W;
  if (pos) CUDA_TRY(ctx, cudaMemcpyAsync(pos, st.pos, n * st.n_par * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (lnp) CUDA_TRY(ctx, cudaMemcpyAsync(lnp, st.lnp, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_acc) CUDA_TRY(ctx, cudaMemcpyAsync(n_acc, st.n_acc, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_best(agnf_ctx* ctx, double* best_pos, double* best_lnp) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K3State& st = ctx->ens;
  if (best_pos)
    CUDA_TRY(ctx, cudaMemcpyAsync(best_pos, ctx->ens_best_pos, (size_t)st.S * st.n_par * 8, cudaMemcpyDeviceToHost,
                                  ctx->stream));
  if (best_lnp)
    CUDA_TRY(ctx, cudaMemcpyAsync(best_lnp, ctx->ens_best_lnp, (size_t)st.S * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_reset_best(agnf_ctx* ctx) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  std::vector<double> ninf((size_t)ctx->ens.S, -INFINITY);
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ens_best_lnp, ninf.data(), (size_t)ctx->ens.S * 8, cudaMemcpyHostToDevice,
                                ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_reset_stats(agnf_ctx* ctx) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K3State& st = ctx->ens;
  CUDA_TRY(ctx, cudaMemsetAsync(st.n_acc, 0, (size_t)st.S * st.W * 4, ctx->stream));
  if (ctx->ens_hist.hist)
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->ens_hist.hist, 0, (size_t)st.S * st.n_par * ctx->ens_hist.n_bins * 4, ctx->stream));
  ctx->ens_hist_samples = 0;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_ens_percentiles(agnf_ctx* ctx, const double* q, int32_t n_q, double* out) {
  AGNF_GUARD_BEGIN
  if (!ctx || !q || !out) return 1;
  if (!ctx->have_ens || !ctx->ens_hist.hist) return fail(ctx, "the ensemble has no histograms");
  if (n_q < 1 || n_q > 64) return fail(ctx, "n_q outside [1, 64]");
  for (int k = 0; k < n_q; ++k)
    if (!(q[k] >= 0.0 && q[k] <= 100.0) || (k > 0 && q[k] < q[k - 1]))
      return fail(ctx, "percentiles must be ascending values in [0, 100]");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K3State& st = ctx->ens;
  void *dq = nullptr, *dout = nullptr;
  const size_t n_out = (size_t)st.S * st.n_par * n_q;
  if (dev_alloc(ctx, &dq, (size_t)n_q * 8) || dev_alloc(ctx, &dout, n_out * 8)) return 1;
  cudaError_t e = cudaMemcpyAsync(dq, q, (size_t)n_q * 8, cudaMemcpyHostToDevice, ctx->stream);
  ctx->ens_hist.pbox = ctx->k1.pbox;
  if (e == cudaSuccess)
    e = launch_k3_percentiles(ctx->ens_hist, st.S, st.n_par, static_cast<const double*>(dq), n_q,
                              static_cast<double*>(dout), ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, n_out * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, dq);
  dev_free(ctx, dout);
  if (e != cudaSuccess) return fail(ctx, "agnf_ens_percentiles: %s", cudaGetErrorString(e));
  ctx->launches += 1;
  return 0;
  AGNF_GUARD_END(ctx)
}

}  // extern "C"


// ------------------------------------------------------------------------------------------- K4

AGNF_API int agnf_sed_config(agnf_ctx* ctx, int32_t G, const double* grid, const double* xg, int32_t to_lo, int32_t to_hi,
                    int32_t rad_hi, int32_t dered_mode, const int32_t* dered_row, const int32_t* bands_jt,
                    const double* bands_wt) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_cfg) return fail(ctx, "agnf_modelset_config has not been called");
  if (G < 2 || !grid || !xg) return fail(ctx, "agnf_sed_config: empty frequency grid");
  if (to_lo < 0 || to_hi > G || rad_hi < 0 || rad_hi > G) return fail(ctx, "agnf_sed_config: cut-off outside the grid");
  if (dered_mode < 0 || dered_mode > 3) return fail(ctx, "dered_mode must be 0..3");
  if (dered_mode != 0 && !dered_row) return fail(ctx, "dered_mode %d needs dered_row", dered_mode);
  if (dered_mode == 3 && (!bands_jt || !bands_wt)) return fail(ctx, "dered_mode 3 needs the bracket table of the band centres");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  K4Params& Q = ctx->k4;
  Q.G = G;
  Q.to_lo = to_lo;
  Q.to_hi = to_hi;
  Q.rad_hi = rad_hi;
  Q.dered_mode = dered_mode;
  if (upload(ctx, &Q.grid, grid, (size_t)G)) return 1;
  if (upload(ctx, &Q.xg, xg, (size_t)G)) return 1;
  if (dered_mode != 0) {
    if (!ctx->fam_set[3]) return fail(ctx, "upload the accretion-disc family before agnf_sed_config");
    if (upload(ctx, &Q.dered_row, dered_row, (size_t)ctx->k1.fam[3].T)) return 1;
  }
  if (dered_mode == 3) {
    std::vector<K4Bracket> bt((size_t)G);
    for (int i = 0; i < G; ++i) {
      if (bands_jt[i] < -1 || bands_jt[i] >= ctx->k1.cfg.n_b) return fail(ctx, "bands_jt[%d] out of range", i);
      bt[(size_t)i] = K4Bracket{bands_jt[i] < 0 ? 0.0 : bands_wt[i], bands_jt[i], 0};
    }
    if (upload(ctx, &Q.bands_bt, bt.data(), (size_t)G)) return 1;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (int f = 0; f < AGNF_MAX_FAMILIES; ++f) {     // bracket tables depend on the grid: families must follow
    ctx->k4_fam[f] = false;
    Q.fam[f].present = 0;
  }
  ctx->k4_cfg = true;
  ctx->k4_dirty = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_sed_family_upload(agnf_ctx* ctx, const agnf_sed_family* d) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->k4_cfg) return fail(ctx, "call agnf_sed_config first");
  if (!d || d->fam < 0 || d->fam >= ctx->k1.cfg.n_fam) return fail(ctx, "agnf_sed_family_upload: bad family id");
  if (!ctx->fam_set[d->fam]) return fail(ctx, "family %d has no dictionary yet (agnf_family_upload)", d->fam);
  if (d->T != ctx->k1.fam[d->fam].T)
    return fail(ctx, "family %d: %lld full-resolution templates for %lld dictionary rows", d->fam, (long long)d->T,
                (long long)ctx->k1.fam[d->fam].T);
  if (d->mode < 0 || d->mode > 3) return fail(ctx, "mode must be 0..3");
  if (d->n_groups < 1 || !d->group || !d->local || !d->g_nlam || !d->x || !d->val || !d->jtab || !d->wtab)
    return fail(ctx, "agnf_sed_family_upload: NULL argument");
  if (d->mode != 0 && !d->k) return fail(ctx, "mode %d needs the reddening curve", d->mode);
  const int n_par = ctx->k1.cfg.n_par;
  if (d->mode != 0 && (d->e_col < 0 || d->e_col >= n_par)) return fail(ctx, "e_col out of range");
  if (d->mode == 3 && (d->a_col < 0 || d->a_col >= n_par)) return fail(ctx, "a_col out of range");
  const bool need2 = d->fam == 3 && (ctx->k4.dered_mode == 1 || ctx->k4.dered_mode == 2);
  if (need2 && !d->val2) return fail(ctx, "the accretion-disc family needs val2 for dered_mode %d", ctx->k4.dered_mode);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int ng = d->n_groups;
  std::vector<int32_t> xoff((size_t)ng), count((size_t)ng, 0);
  std::vector<int64_t> voff((size_t)ng);
  int64_t nx = 0;
  int32_t max_nlam = 0;
  for (int g = 0; g < ng; ++g) {
    if (d->g_nlam[g] < 1) return fail(ctx, "family %d group %d has no samples", d->fam, g);
    xoff[(size_t)g] = (int32_t)nx;
    nx += d->g_nlam[g];
    max_nlam = std::max(max_nlam, d->g_nlam[g]);
  }
  for (int64_t t = 0; t < d->T; ++t) {
    if (d->group[t] < 0 || d->group[t] >= ng) return fail(ctx, "family %d: group[%lld] out of range", d->fam, (long long)t);
    count[(size_t)d->group[t]] = std::max(count[(size_t)d->group[t]], d->local[t] + 1);
  }
  int64_t nv = 0;
  for (int g = 0; g < ng; ++g) {
    voff[(size_t)g] = nv;
    nv += (int64_t)count[(size_t)g] * d->g_nlam[g];
  }
  std::vector<int64_t> t_off((size_t)d->T);
  for (int64_t t = 0; t < d->T; ++t) {
    if (d->local[t] < 0) return fail(ctx, "family %d: local[%lld] < 0", d->fam, (long long)t);
    t_off[(size_t)t] = voff[(size_t)d->group[t]] + (int64_t)d->local[t] * d->g_nlam[d->group[t]];
  }
  std::vector<K4Bracket> bt((size_t)ng * (size_t)ctx->k4.G);
  for (int64_t i = 0, n = (int64_t)ng * ctx->k4.G; i < n; ++i) {
    const int g = (int)(i / ctx->k4.G);
    const int32_t j = d->jtab[i];
    const double w = j < 0 ? 0.0 : d->wtab[i];
    // a point strictly inside an interval needs the next sample; on a sample (w == 0) the last one is allowed
    if (j < -1 || j >= d->g_nlam[g] || (w != 0.0 && j + 1 >= d->g_nlam[g]) || !(w >= 0.0 && w < 1.0))
      return fail(ctx, "family %d: bracket table entry %lld out of range", d->fam, (long long)i);
    bt[(size_t)i] = K4Bracket{w, j, 0};
  }
  K4Family& F = ctx->k4.fam[d->fam];
  F.T = d->T;
  F.mode = d->mode;
  F.e_col = d->e_col;
  F.a_col = d->a_col;
  F.max_nlam = max_nlam;
  if (upload(ctx, &F.x, d->x, (size_t)nx)) return 1;
  if (d->mode != 0) {
    if (upload(ctx, &F.k, d->k, (size_t)nx)) return 1;
  }
  if (upload(ctx, &F.val, d->val, (size_t)nv)) return 1;
  if (need2) {
    if (upload(ctx, &F.val2, d->val2, (size_t)nv)) return 1;
  }
  if (upload(ctx, &F.t_off, t_off.data(), (size_t)d->T)) return 1;
  if (upload(ctx, &F.t_grp, d->group, (size_t)d->T)) return 1;
  if (upload(ctx, &F.g_xoff, xoff.data(), (size_t)ng)) return 1;
  if (upload(ctx, &F.g_nlam, d->g_nlam, (size_t)ng)) return 1;
  if (upload(ctx, &F.btab, bt.data(), bt.size())) return 1;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));     // the vectors above feed asynchronous copies
  F.present = 1;
  ctx->k4_fam[d->fam] = true;
  ctx->k4_dirty = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_sed_sources(agnf_ctx* ctx, int64_t S, const double* z, const double* dlum) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (S < 1 || !z || !dlum) return fail(ctx, "agnf_sed_sources: empty");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (upload(ctx, &ctx->k4.src_z, z, (size_t)S)) return 1;
  if (upload(ctx, &ctx->k4.src_dlum, dlum, (size_t)S)) return 1;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->k4_S = S;
  ctx->k4_src = true;
  ctx->k4_dirty = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_sed_bands(agnf_ctx* ctx, int32_t n_ranges, const int32_t* i0, const int32_t* i1, const int32_t* mask,
                   int32_t n_specs, const int32_t* kind, const int32_t* range, const int32_t* num,
                   const int32_t* den) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->k4_cfg) return fail(ctx, "call agnf_sed_config first");
  if (n_ranges < 1 || n_ranges > K4_MAX_RANGES) return fail(ctx, "1..%d integration bands", K4_MAX_RANGES);
  if (n_specs < 1 || n_specs > K4_MAX_SPECS) return fail(ctx, "1..%d output columns", K4_MAX_SPECS);
  if (!i0 || !i1 || !mask || !kind || !range || !num || !den) return fail(ctx, "agnf_sed_bands: NULL argument");
  K4Params& Q = ctx->k4;
  const int all = (1 << K4_COMPS) - 1;
  for (int q = 0; q < n_ranges; ++q) {
    if (i0[q] < 0 || i0[q] > Q.G || (i1[q] >= 0 && (i1[q] < i0[q] || i1[q] > Q.G)) || (i1[q] < 0 && i0[q] >= Q.G))
      return fail(ctx, "band %d lies outside the grid", q);
    if (mask[q] & ~all) return fail(ctx, "band %d: unknown component bits", q);
    if ((mask[q] & 16) && !ctx->k1.cfg.radio) return fail(ctx, "band %d asks for the radio component of a model set without one", q);
    Q.r_i0[q] = i0[q];
    Q.r_i1[q] = i1[q];
    Q.r_mask[q] = mask[q];
  }
  for (int s = 0; s < n_specs; ++s) {
    if (range[s] < 0 || range[s] >= n_ranges) return fail(ctx, "column %d: band out of range", s);
    if ((num[s] | den[s]) & ~mask[range[s]]) return fail(ctx, "column %d uses a component its band does not compute", s);
    if (kind[s] != 0 && kind[s] != 1) return fail(ctx, "column %d: kind must be 0 or 1", s);
    Q.s_kind[s] = kind[s];
    Q.s_range[s] = range[s];
    Q.s_num[s] = num[s];
    Q.s_den[s] = den[s];
  }
  Q.n_ranges = n_ranges;
  Q.n_specs = n_specs;
  ctx->k4_bands = true;
  ctx->k4_dirty = true;
  return 0;
  AGNF_GUARD_END(ctx)
}

namespace {

int k4_ready(agnf_ctx* ctx, bool need_bands) {
  if (check_ready(ctx)) return 1;
  if (!ctx->k4_cfg) return fail(ctx, "agnf_sed_config has not been called");
  for (int f = 0; f < ctx->k1.cfg.n_fam; ++f)
    if (!ctx->k4_fam[f]) return fail(ctx, "full-resolution family %d has not been uploaded", f);
  if (!ctx->k4_src) return fail(ctx, "agnf_sed_sources has not been called");
  if (ctx->k4_S != ctx->k1.S) return fail(ctx, "agnf_sed_sources: %lld sources, agnf_sources_upload had %lld",
                                          (long long)ctx->k4_S, (long long)ctx->k1.S);
  if (need_bands && !ctx->k4_bands) return fail(ctx, "agnf_sed_bands has not been called");
  if (ctx->k4_dirty) {
    K4Params& Q = ctx->k4;
    int off = 0;
    for (int f = 0; f < ctx->k1.cfg.n_fam; ++f) {
      Q.fam[f].smem_off = off;
      off += (Q.fam[f].max_nlam + 1) & ~1;
    }
    Q.dered_off = off;
    if (Q.dered_mode == 1 || Q.dered_mode == 2) off += (Q.fam[3].max_nlam + 1) & ~1;
    if (Q.dered_mode == 3) off += (ctx->k1.cfg.n_b + 1) & ~1;
    Q.smem_doubles = off;
    if ((size_t)off * 8 > 160 * 1024) return fail(ctx, "templates of one realisation need %d KB of shared memory", off / 128);
    if (!ctx->d_k4 && dev_alloc(ctx, &ctx->d_k4, sizeof(K4Params))) return 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_k4, &Q, sizeof(K4Params), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->k4_dirty = false;
  }
  return 0;
}

int k4_run(agnf_ctx* ctx, const double* params, const int32_t* source_id, int64_t n, double* out, int32_t on_device,
           bool spectra) {
  if (!ctx) return 1;
  if (n < 0) return fail(ctx, "n < 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (k4_ready(ctx, !spectra)) return 1;
  if (n == 0) return 0;
  if (!params || !source_id || !out) return fail(ctx, "NULL argument");
  if (n > 2147483647LL) return fail(ctx, "at most 2^31 - 1 rows per call");
  const int n_par = ctx->k1.cfg.n_par;
  const int n_comp = ctx->k1.cfg.radio ? 7 : 6;
  const size_t out_per = spectra ? (size_t)n_comp * (size_t)ctx->k4.G : (size_t)ctx->k4.n_specs;
  cudaStream_t st = ctx->stream;
  void *dp = nullptr, *ds = nullptr, *dout = nullptr;
  auto cleanup = [&]() {
    if (!on_device) {
      dev_free(ctx, dp);
      dev_free(ctx, ds);
      dev_free(ctx, dout);
    }
  };
  if (on_device) {
    dp = const_cast<double*>(params);
    ds = const_cast<int32_t*>(source_id);
    dout = out;
  } else {
    for (int64_t i = 0; i < n; ++i)
      if (source_id[i] < 0 || source_id[i] >= ctx->k1.S) return fail(ctx, "source_id[%lld] out of range", (long long)i);
    if (dev_alloc(ctx, &dp, (size_t)n * n_par * 8) || dev_alloc(ctx, &ds, (size_t)n * 4) ||
        dev_alloc(ctx, &dout, (size_t)n * out_per * 8)) {
      cleanup();
      return 1;
    }
    cudaError_t e = cudaMemcpyAsync(dp, params, (size_t)n * n_par * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ds, source_id, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
      cleanup();
      return fail(ctx, "upload failed: %s", cudaGetErrorString(e));
    }
  }
  K4Batch B{static_cast<const double*>(dp), static_cast<const int32_t*>(ds), n,
            spectra ? nullptr : static_cast<double*>(dout), spectra ? static_cast<double*>(dout) : nullptr};
  cudaError_t e = cudaEventRecord(ctx->ev0, st);
  if (e == cudaSuccess)
    e = spectra ? launch_k4_spectra(ctx->k1, static_cast<const K4Params*>(ctx->d_k4), ctx->k4, B, st)
                : launch_k4_intlum(ctx->k1, static_cast<const K4Params*>(ctx->d_k4), ctx->k4, B, st);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->ev1, st);
  if (e == cudaSuccess && !on_device)
    e = cudaMemcpyAsync(out, dout, (size_t)n * out_per * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) return fail(ctx, "K4 failed: %s", cudaGetErrorString(e));
  ctx->k1_timed = true;
  ctx->launches += 1;
  return 0;
}

}  // namespace

AGNF_API int agnf_sed_intlum(agnf_ctx* ctx, const double* params, const int32_t* source_id, int64_t n, double* out,
                    int32_t on_device) {
  AGNF_GUARD_BEGIN
  return k4_run(ctx, params, source_id, n, out, on_device, false);
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_sed_spectra(agnf_ctx* ctx, const double* params, const int32_t* source_id, int64_t n, double* out,
                     int32_t on_device) {
  AGNF_GUARD_BEGIN
  return k4_run(ctx, params, source_id, n, out, on_device, true);
  AGNF_GUARD_END(ctx)
}

AGNF_API int agnf_sed_intlum_ens(agnf_ctx* ctx, double* out) {
  AGNF_GUARD_BEGIN
  if (!ctx) return 1;
  if (!ctx->have_ens) return fail(ctx, "call agnf_ens_create first");
  if (!out) return fail(ctx, "NULL argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (k4_ready(ctx, true)) return 1;
  const K3State& st = ctx->ens;
  const int64_t n = st.S * st.W;
  void *d_sid = nullptr, *d_out = nullptr;
  if (dev_alloc(ctx, &d_sid, (size_t)n * 4)) return 1;
  if (dev_alloc(ctx, &d_out, (size_t)n * ctx->k4.n_specs * 8)) {
    dev_free(ctx, d_sid);
    return 1;
  }
  cudaError_t e = launch_k3_fill_sid(static_cast<int32_t*>(d_sid), st.S, st.W, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  int rc = e == cudaSuccess ? 0 : fail(ctx, "k3_fill_sid failed: %s", cudaGetErrorString(e));
  if (!rc) rc = k4_run(ctx, st.pos, static_cast<const int32_t*>(d_sid), n, static_cast<double*>(d_out), 1, false);
  if (!rc) {
    e = cudaMemcpy(out, d_out, (size_t)n * ctx->k4.n_specs * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(ctx, "download failed: %s", cudaGetErrorString(e));
  }
  dev_free(ctx, d_sid);
  dev_free(ctx, d_out);
  return rc;
  AGNF_GUARD_END(ctx)
}
