// C-ABI of the tn4ml hot path (see include/tn4ml_b200.h for the contract).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <vector>
#include <type_traits>

#include "common.cuh"
#include "mps_simt.cuh"
#include "mps_tc.cuh"
#include "smpo_simt.cuh"

namespace tnb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define TNB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(_e));  \
  } while (0)

// Optional per-launch timing (bench.py's roofline leg): when enabled, every kernel launch
// is bracketed by CUDA events on the launching stream.  Off by default; not thread-safe
// by design (a measurement aid, enabled by one thread around one timed region).
struct ProfRec { const char *name; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static inline void prof_begin(cudaStream_t st) {
  if (!g_prof_on) return;
  ProfRec r{"?", nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
static inline void prof_end(const char *name, cudaStream_t st) {
  if (!g_prof_on || g_prof.empty()) return;
  g_prof.back().name = name;
  cudaEventRecord(g_prof.back().b, st);
}

#define TNB_LAUNCH_CHECK(name)                                                           \
  do {                                                                                   \
    prof_end(name, st);                                                                  \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) return fail("launch %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline size_t esize(const tnb_problem *p) { return p->dtype == TNB_F64 ? 8 : 4; }

static int emb_dim(const tnb_embedding &e) {
  switch (e.kind) {
    case TNB_EMB_TRIG: return 2 * e.k;
    case TNB_EMB_FOURIER: return e.p;
    case TNB_EMB_POLY: return e.degree + (e.include_bias ? 1 : 0);
    case TNB_EMB_RBF: return e.n_centers;
  }
  return -1;
}

static int validate(const tnb_problem *p) {
  if (!p) return fail("null problem");
  if (p->dtype != TNB_F32 && p->dtype != TNB_F64) return fail("dtype must be TNB_F32 or TNB_F64");
  if (p->L < 2) return fail("L must be >= 2 (got %d)", p->L);
  if (p->chi < 1 || p->chi > 512) return fail("chi must be in [1, 512] (got %d)", p->chi);
  if (p->d < 1 || p->d > TNB_MAX_PHYS) return fail("d must be in [1, %d] (got %d)", TNB_MAX_PHYS, p->d);
  if (p->n_out < 1 || p->n_out > TNB_MAX_OUT) return fail("n_out must be in [1, %d]", TNB_MAX_OUT);
  int ed = emb_dim(p->emb);
  if (ed != p->d)
    return fail("physical dimension of the model (%d) does not match the embedding dimension (%d)", p->d, ed);
  if (p->kind == TNB_KIND_MPS) {
    if (p->out_site < 0 || p->out_site >= p->L) return fail("out_site out of range");
    if (p->head > TNB_HEAD_MSE || p->head < 0) return fail("head %d is not an MPS head", p->head);
    if (p->head == TNB_HEAD_NLL && p->n_out != 1) return fail("NLL head needs n_out == 1");
  } else if (p->kind == TNB_KIND_SMPO) {
    if (p->spacing < 1) return fail("spacing must be >= 1");
    if (p->head < TNB_HEAD_TSN || p->head > TNB_HEAD_QUAD) return fail("head %d is not an SMPO head", p->head);
    if (p->chi > 64) return fail("SMPO kernels support chi <= 64 (got %d)", p->chi);
  } else {
    return fail("unknown kind %d", p->kind);
  }
  return 0;
}

template <typename T> static void fill_emb(const tnb_embedding &e, EmbParams<T> &o, int d) {
  o.kind = e.kind; o.k = e.k; o.p = e.p; o.degree = e.degree; o.include_bias = e.include_bias;
  o.n_centers = e.n_centers; o.d = d; o.gamma = (T)e.gamma;
  for (int i = 0; i < kMaxPhys; ++i) o.centers[i] = (T)e.centers[i];
}

// ------------------------------------------------------------------------------------
// MPS workspace plan
// ------------------------------------------------------------------------------------
struct MpsPlan {
  int chip, nslab, TN, NG, NT, BT, KT, WS, dmma_sweep, dmma_class;
  size_t off_Wn, off_Wt, off_F, off_G, off_exps, off_esum, off_y, off_dy, off_wsimt, total;
  size_t smem_sweep, smem_class;
  int RB;  // register-tile rows of the SIMT sweep / class kernels
  // tensor-core sweep (f32, TNB_PREC_TENSOR)
  int tc, tc_N, tc_NH, tc_NMMA, tc_nstage, tc_cols, tc_tiles, tc_cg;
  uint32_t tc_stage_bytes;
  size_t tc_site_img, off_img_n, off_img_t, off_kW, off_cs, off_wq, off_qmax, off_wqc, off_fimg, off_gimg, tc_smem, img_bond_bytes;
};

// Measurement knobs of the tensor path (environment, read once): TNB_TC_NH forces a legal column-chunk count,
// TNB_TC_TILES=1 forces one sample tile per CTA, TNB_TC_PACE_NS sets the image writer's pause between pieces,
// TNB_DA_STAGES lowers the pipeline depth of the weight-gradient GEMM, TNB_TC_CG=2 turns on the (experimental) CTA-pair chain kernel.
struct TcKnobs { int nh = 0, tiles = 0, pace_ns = -1, da_stages = 0, cg = 0; };
static const TcKnobs &tc_knobs() {
  static TcKnobs k;
  static std::once_flag once;
  std::call_once(once, [] {
    if (const char *e = getenv("TNB_TC_NH")) k.nh = atoi(e);
    if (const char *e = getenv("TNB_TC_TILES")) k.tiles = atoi(e);
    if (const char *e = getenv("TNB_TC_PACE_NS")) k.pace_ns = atoi(e);
    if (const char *e = getenv("TNB_DA_STAGES")) k.da_stages = atoi(e);
    if (const char *e = getenv("TNB_TC_CG")) k.cg = atoi(e);
  });
  return k;
}

// The tcgen05 sweep covers: float32, chi a multiple of 16 in [32, 256], D*chi <= 512 TMEM columns.
static bool tc_supported(const tnb_problem *p) {
  if (p->kind != TNB_KIND_MPS || p->dtype != TNB_F32 || p->precision != TNB_PREC_TENSOR) return false;
  if (p->chi % 16 != 0 || p->chi < 32 || p->chi > 256) return false;
  if (p->d * p->chi > 512) return false;
  if (!(p->d == 2 || p->d == 3 || p->d == 4 || p->d == 6 || p->d == 8)) return false;
  return true;
}

// Column chunks per chain step: the smallest NH dividing chi/16 whose chunk (D*chi/NH columns) fits one MMA (N <= 256).
static int tc_pick_nh(const tnb_problem *p) {
  const int groups = p->chi / 16;
  auto legal = [&](int nh) { return nh >= 1 && nh <= kTcMaxNH && groups % nh == 0 && p->d * p->chi / nh <= 256; };
  if (legal(tc_knobs().nh)) return tc_knobs().nh;
  for (int nh = 1; nh <= groups; ++nh)
    if (legal(nh)) return nh;
  return groups;
}

static MpsPlan plan_mps(const tnb_problem *p, long long B, int need_grad) {
  MpsPlan P{};
  const size_t w = esize(p);
  P.chip = (p->chi + 3) / 4 * 4;
  P.nslab = p->L - 1 + p->n_out;
  int ngroups = P.chip / kRN;
  P.TN = 1;
  while (P.TN < ngroups && P.TN < 32) P.TN <<= 1;
  int ng = (ngroups + P.TN - 1) / P.TN;
  P.NG = ng <= 1 ? 1 : (ng <= 2 ? 2 : 4);
  P.NT = 256;
  // register-tile rows per thread: 4; TNB_SIMT_RB=8 selects 8 for float64 when the batch still fills the machine with
  // 256-thread CTAs (twice the FMAs per operand load, but half the resident warps)
  static const int rb_knob = [] { const char *e = getenv("TNB_SIMT_RB"); return e ? atoi(e) : 0; }();
  P.RB = 4;
  if (p->dtype == TNB_F64 && P.NG <= 2 && rb_knob == 8) {  // (measured slower at chi = 128: 204 registers, one CTA per SM)
    const long long bt8 = (long long)(256 / P.TN) * 8;
    if ((B + bt8 - 1) / bt8 >= 148) P.RB = 8;
  }
  static const int nt_knob = [] { const char *e = getenv("TNB_SIMT_NT"); return e ? atoi(e) : 0; }();
  if (nt_knob == 128 || nt_knob == 64) P.NT = nt_knob;
  while (P.NT > 32) {
    long long bt = (long long)(P.NT / P.TN) * P.RB;
    if ((B + bt - 1) / bt >= 148) break;
    P.NT >>= 1;
  }
  if (P.NT < P.TN) P.NT = P.TN < 32 ? 32 : P.TN;
  P.BT = P.NT / P.TN * P.RB;
  size_t row = (size_t)p->d * P.chip * w;
  long kt = (long)(16384 / row);
  static const int kt_knob = [] { const char *e = getenv("TNB_SIMT_KT"); return e ? atoi(e) : 0; }();
  if (kt_knob > 0) kt = kt_knob;
  P.KT = kt < 1 ? 1 : (kt > 16 ? 16 : (int)kt);
  if (P.KT > P.chip) P.KT = P.chip;
  if (P.KT >= 4) P.KT = P.KT / 4 * 4;  // (whole k-steps of the DMMA step GEMM)
  P.WS = p->dtype == TNB_F64 ? P.NG * P.TN * kRN : P.chip;  // row stride of the kernel-layout weights (mps_prep_kernel)
  size_t slab = (size_t)P.chip * p->d * P.WS * w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  P.off_Wn = take(slab * P.nslab);
  P.off_Wt = take(slab * P.nslab);
  const bool tc_mode = tc_supported(p);
  size_t fslots = need_grad ? (size_t)p->L + 1 : 2;
  P.off_F = tc_mode ? 0 : take(fslots * B * P.chip * w);
  P.off_G = (need_grad && !tc_mode) ? take(((size_t)p->L + 1) * B * P.chip * w) : 0;
  P.off_exps = take((size_t)p->L * B * sizeof(int));
  P.off_esum = take((size_t)2 * B * sizeof(int));
  P.off_y = take((size_t)B * p->n_out * w);
  P.off_dy = need_grad ? take((size_t)B * p->n_out * w) : 0;
  P.off_wsimt = (need_grad && !tc_mode) ? take((size_t)p->L * p->d * B * w) : 0;  // per-sample dA weights of the SIMT family
  P.tc = tc_supported(p) ? 1 : 0;
  if (P.tc) {
    P.tc_N = p->d * p->chi;
    P.tc_NH = tc_pick_nh(p);
    P.tc_NMMA = P.tc_N / P.tc_NH;
    P.tc_stage_bytes = 64u * P.tc_NMMA;  // one K-slab (16 rows) of one column chunk, hi and lo planes
    P.tc_site_img = (size_t)P.tc_NH * (p->chi / 16) * P.tc_stage_bytes;  // (stage_bytes: still the whole stage here)
    const size_t a_bytes = (size_t)2 * kTcRows * p->chi * 2;
    // two sample tiles per CTA when one column chunk covers the step (nothing else hides the epilogue); TNB_TC_TILES=1 forces one
    // (pays when the step's MMA time is comparable to its exposed tail: N = 256; measured slower at N <= 128)
    P.tc_tiles = (P.tc_NH == 1 && P.tc_N == 256) ? 2 : 1;
    if (tc_knobs().tiles == 1) P.tc_tiles = 1;
    if (tc_knobs().tiles == 2 && P.tc_NH == 1 && (p->d == 2 || p->d == 4 || p->d == 8)) P.tc_tiles = 2;
    // CTA pair (cta_group::2), opt-in: the wide-bond case (chi > 128, d = 2), where shared-memory bandwidth caps the MMA
    // rate.  Parity-green but measured slower than the one-CTA kernel so far, hence not the default.
    P.tc_cg = (p->chi > 128 && p->d == 2 && P.tc_tiles == 1 && tc_knobs().cg == 2) ? 2 : 1;
    P.tc_stage_bytes /= (uint32_t)P.tc_cg;  // one CTA's half of a stage
    P.tc_cols = 32;
    while (P.tc_cols < P.tc_tiles * P.tc_N) P.tc_cols <<= 1;
    const size_t fixed = (size_t)P.tc_tiles * 4096 /* row exchange */ + 1024 /* tables, barriers */;
    long ns = (long)((232448 - (long)fixed - (long)(P.tc_tiles * a_bytes)) / (long)P.tc_stage_bytes);
    P.tc_nstage = ns > 12 ? 12 : (int)ns;
    P.tc_smem = P.tc_tiles * a_bytes + (size_t)P.tc_nstage * P.tc_stage_bytes + fixed;
    if (P.tc_nstage < 2) P.tc = 0;
  }
  if (P.tc) {
    // the tensor path keeps its environments as fp16 hi/lo bond images instead of F / G
    const size_t span = (size_t)kTcRows * P.tc_cg;  // a CTA pair always writes both tiles
    P.img_bond_bytes = (size_t)((B + span - 1) / span * span) * p->chi * 4;  // whole 128-sample tiles
    P.off_fimg = take((need_grad ? (size_t)p->L + 1 : 1) * P.img_bond_bytes);
    P.off_img_n = take(P.tc_site_img * P.nslab);
    P.off_img_t = take(P.tc_site_img * P.nslab);
    P.off_kW = take((size_t)P.nslab * sizeof(int));
    P.off_cs = take((size_t)2 * P.nslab * sizeof(float));
    if (need_grad) {
      P.off_wq = take((size_t)p->L * p->d * B * sizeof(float));
      P.off_qmax = take(((size_t)p->L + 1) * sizeof(int));  // [L] chain sites + [1] class site
      P.off_wqc = take((size_t)p->d * p->n_out * B * sizeof(float));
      P.off_gimg = take(((size_t)p->L + 1) * P.img_bond_bytes);
    }
  }
  P.total = o;
  size_t base = ((size_t)P.BT * P.chip + (size_t)2 * P.KT * p->d * P.WS + (size_t)P.BT * p->d +
                 (size_t)P.BT * p->n_out) * w;
  P.smem_sweep = base;
  P.smem_class = base + (size_t)P.BT * P.chip * w;
  // float64, d = 2, one column group, full 256-thread CTAs: the step GEMM runs as DMMA warp tiles and needs one more
  // [BT][chip] tile behind everything else (TNB_SIMT_DMMA=0 keeps the FMA loop)
  static const int dmma_knob = [] { const char *e = getenv("TNB_SIMT_DMMA"); return e ? atoi(e) : 1; }();
  P.dmma_sweep = P.dmma_class = 0;
  if (p->dtype == TNB_F64 && p->d == 2 && P.NG <= 2 && P.RB == 4 && P.TN >= 8 && P.NT == 256 && P.KT % 4 == 0 && dmma_knob) {
    const size_t tile = (size_t)P.BT * P.chip * w, cap = 200 * 1024;  // (each kernel only if it still fits)
    if (align_up(P.smem_sweep, 16) + tile <= cap) {
      P.dmma_sweep = (int)align_up(P.smem_sweep, 16);
      P.smem_sweep = P.dmma_sweep + tile;
    }
    if (align_up(P.smem_class, 16) + tile <= cap) {
      P.dmma_class = (int)align_up(P.smem_class, 16);
      P.smem_class = P.dmma_class + tile;
    }
  }
  return P;
}

template <typename T>
static void fill_mps_args(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                          void *ws, MpsArgs<T> &A) {
  char *w = (char *)ws;
  A.L = p->L; A.chi = p->chi; A.chip = P.chip; A.D = p->d; A.C = p->n_out; A.c = p->out_site; A.head = p->head;
  A.B = B;
  fill_emb<T>(p->emb, A.emb, p->d);
  for (int i = 0; i < kMaxOut; ++i) A.cw[i] = (T)p->class_weights[i];
  A.x = (const T *)x; A.targets = (const T *)targets;
  A.Wn = (const T *)(w + P.off_Wn); A.Wt = (const T *)(w + P.off_Wt);
  A.F = (T *)(w + P.off_F); A.G = (T *)(w + P.off_G);
  A.exps = (int *)(w + P.off_exps); A.esum = (int *)(w + P.off_esum);
  A.y = (T *)(w + P.off_y); A.dy = (T *)(w + P.off_dy);
  A.wq = (T *)(w + P.off_wsimt);
  A.loss = nullptr; A.loss_sum = nullptr; A.out = nullptr; A.out_exp = nullptr;
  A.stash = 0; A.KT = P.KT; A.TN = P.TN; A.WS = P.WS; A.dmma_sweep = P.dmma_sweep; A.dmma_class = P.dmma_class; A.loss_scale = T(1); A.seed_only = 0;
}

static void fill_tc_args(const MpsPlan &P, void *ws, const MpsArgs<float> &M, TcArgs &T) {
  char *w = (char *)ws;
  T.M = M;
  T.img_n = (const uint8_t *)(w + P.off_img_n);
  T.img_t = (const uint8_t *)(w + P.off_img_t);
  T.kW = (const int *)(w + P.off_kW);
  T.tile0 = 0;
  T.cg = P.tc_cg;
  { const char *e = getenv("TNB_TC_DBG"); T.dbg = e ? atoi(e) : 0; }
  T.colsum = (const float *)(w + P.off_cs);
  T.pace_ns = tc_knobs().pace_ns >= 0 ? tc_knobs().pace_ns : P.tc_N * 2 / 5;
  T.wq = (float *)(w + P.off_wq);
  T.qmax = (int *)(w + P.off_qmax);
  T.wqc = (float *)(w + P.off_wqc);
  T.fimg = (uint8_t *)(w + P.off_fimg);
  T.gimg = M.stash ? (uint8_t *)(w + P.off_gimg) : nullptr;
  T.img_bond_bytes = P.img_bond_bytes;
  T.img_stash = M.stash;
  T.qmaxc = T.qmax + M.L;
  T.N = P.tc_N; T.NH = P.tc_NH; T.NMMA = P.tc_NMMA; T.nstage = P.tc_nstage; T.tmem_cols = P.tc_cols;
  T.stage_bytes = P.tc_stage_bytes; T.site_img_bytes = P.tc_site_img;
}

template <int D, int SLOTS, int TILES, int CG> static cudaError_t tc_set_attr() {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(tc_chain_kernel<D, SLOTS, TILES, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  });
  return err;
}

template <int D, int SLOTS, int TILES, int CG = 1>
static cudaError_t tc_launch_one(long long ntiles, long long tile0, size_t smem, TcArgs T, int adjoint, cudaStream_t st) {
  cudaError_t e = tc_set_attr<D, SLOTS, TILES, CG>();
  if (e != cudaSuccess || ntiles <= 0) return e;
  T.tile0 = tile0 * kTcRows;
  if (CG == 1) {
    dim3 grid((unsigned)(ntiles / TILES), adjoint ? 2 : 1);
    tc_chain_kernel<D, SLOTS, TILES, CG><<<grid, kTcThreads, smem, st>>>(T, adjoint);
    return cudaSuccess;
  }
  // CTA pair: clusters of two CTAs along x (consecutive sample tiles); an odd tile count gets an all-padding tile
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((ntiles + 1) / 2 * 2), adjoint ? 2 : 1);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, tc_chain_kernel<D, SLOTS, TILES, CG>, T, adjoint);
}

// launch the tcgen05 sweep (forward environments or adjoint chains)
static int tc_launch_sweep(const tnb_problem *p, const MpsPlan &P, const TcArgs &T, long long B, int adjoint,
                           cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  const bool wide = p->chi > 128;   // 16-column groups per epilogue thread: 4 (chi <= 256) or 2 (chi <= 128)
  const long long ntiles = (B + kTcRows - 1) / kTcRows;
  // Two-tile CTAs take whole rounds of 148 SMs (per half of the adjoint grid); the remaining tiles run as one-tile
  // CTAs behind them, so that the last round is not half empty with double-length CTAs.
  long long npair = 0;
  if (P.tc_tiles == 2) {
    const long long per_round = 2LL * 148 / (adjoint ? 2 : 1);
    npair = ntiles / per_round * per_round;
    if (ntiles - npair > per_round / 2) npair = ntiles / 2 * 2;  // the rest would need two more rounds anyway
  }
#define TNB_TC_GO(DD, SS)                                                                                     \
  do {                                                                                                        \
    if (npair > 0) e = tc_launch_one<DD, SS, 2>(npair, 0, P.tc_smem, T, adjoint, st);                         \
    if (e == cudaSuccess && ntiles > npair) e = tc_launch_one<DD, SS, 1>(ntiles - npair, npair, P.tc_smem, T, adjoint, st); \
  } while (0)
#define TNB_TC_GO1(DD, SS) e = tc_launch_one<DD, SS, 1>(ntiles, 0, P.tc_smem, T, adjoint, st)
#define TNB_TC_PAIR(DD, SS) e = tc_launch_one<DD, SS, 1, 2>(ntiles, 0, P.tc_smem, T, adjoint, st)
  prof_begin(st);
  switch (p->d) {
    case 2:
      if (wide && P.tc_cg == 2) TNB_TC_PAIR(2, 4);
      else if (wide) TNB_TC_GO1(2, 4);
      else TNB_TC_GO(2, 2);
      break;
    case 3: if (wide) TNB_TC_GO1(3, 4); else TNB_TC_GO1(3, 2); break;
    case 4: TNB_TC_GO(4, 2); break;
    case 6: TNB_TC_GO1(6, 2); break;
    case 8: TNB_TC_GO(8, 2); break;
    default: return fail("tensor path: unsupported physical dimension %d", p->d);
  }
#undef TNB_TC_GO
#undef TNB_TC_GO1
#undef TNB_TC_PAIR
  if (e != cudaSuccess) return fail("cudaFuncSetAttribute(tc_chain): %s", cudaGetErrorString(e));
  TNB_LAUNCH_CHECK(adjoint ? "tc_chain_kernel(adj)" : "tc_chain_kernel(fwd)");
  return 0;
}

template <typename T, int NG, int RB> static int set_smem_attrs() {
  // 200 KB opt-in ceiling for the dynamic-smem kernels; done once per instantiation.
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    cudaError_t e = cudaFuncSetAttribute(mps_sweep_kernel<T, NG, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mps_class_kernel<T, NG, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    err = e;
  });
  if (err != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(err));
  return 0;
}

template <typename T, int NG, int RB>
static int mps_forward_t(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                         const void *params, void *loss, double *loss_sum, void *out, int32_t *out_exp,
                         int keep_stash, void *ws, cudaStream_t st) {
  if (set_smem_attrs<T, NG, RB>()) return 1;
  if (P.smem_class > 200 * 1024) return fail("shared-memory plan too large (%zu B)", P.smem_class);
  MpsArgs<T> A;
  fill_mps_args<T>(p, P, B, x, targets, ws, A);
  A.stash = keep_stash ? 1 : 0;
  A.loss = (T *)loss; A.loss_sum = loss_sum; A.out = (T *)out; A.out_exp = out_exp;
  if constexpr (std::is_same<T, float>::value) {
    if (P.tc) {
      // tensor path: weight images, then ONE kernel for both chains, the class site and the head
      TcArgs Tc;
      fill_tc_args(P, ws, A, Tc);
      prof_begin(st);
      tc_sitemax_kernel<<<P.nslab, 256, 0, st>>>((const float *)params, (int *)Tc.kW, p->L, p->chi, p->d, p->n_out,
                                                 p->out_site);
      TNB_LAUNCH_CHECK("tc_sitemax_kernel");
      if (p->n_out > 1) {
        prof_begin(st);
        tc_classmin_kernel<<<1, 32, 0, st>>>((int *)Tc.kW, p->out_site, p->n_out);
        TNB_LAUNCH_CHECK("tc_classmin_kernel");
      }
      prof_begin(st);
      tc_colsum_kernel<<<dim3(P.nslab, 2), 256, 0, st>>>((const float *)params, Tc.kW, (float *)Tc.colsum, p->L, p->chi,
                                                         p->d, p->n_out, p->out_site);
      TNB_LAUNCH_CHECK("tc_colsum_kernel");
      size_t items = (size_t)P.tc_N * (p->chi / 8) * P.nslab;
      int ib = (int)((items + 255) / 256);
      if (ib > 148 * 32) ib = 148 * 32;
      prof_begin(st);
      tc_image_kernel<<<ib, 256, 0, st>>>((const float *)params, Tc.kW, (uint8_t *)Tc.img_n, (uint8_t *)Tc.img_t, p->L,
                                          p->chi, p->d, p->n_out, p->out_site, P.tc_NH, P.tc_cg);
      TNB_LAUNCH_CHECK("tc_image_kernel");
      return tc_launch_sweep(p, P, Tc, B, 0, st);
    }
  }
  size_t welems = (size_t)P.chip * p->d * P.chip * P.nslab;
  int pblocks = (int)((welems + 255) / 256);
  if (pblocks > 148 * 16) pblocks = 148 * 16;
  prof_begin(st);
  mps_prep_kernel<T><<<pblocks, 256, 0, st>>>((const T *)params, (T *)A.Wn, (T *)A.Wt, p->L, p->chi, P.chip, p->d,
                                              p->n_out, p->out_site, P.WS, P.TN);
  TNB_LAUNCH_CHECK("mps_prep_kernel");
  int tiles = (int)((B + P.BT - 1) / P.BT);
  prof_begin(st);
  mps_sweep_kernel<T, NG, RB><<<dim3(tiles, 2), P.NT, P.smem_sweep, st>>>(A, 0);
  TNB_LAUNCH_CHECK("mps_sweep_kernel(fwd)");
  prof_begin(st);
  mps_class_kernel<T, NG, RB><<<tiles, P.NT, P.smem_class, st>>>(A);
  TNB_LAUNCH_CHECK("mps_class_kernel");
  return 0;
}

template <typename T, int NG, int RB>
static int mps_backward_t(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                          double loss_scale, void *grads, int accumulate, void *ws, cudaStream_t st) {
  if (set_smem_attrs<T, NG, RB>()) return 1;
  MpsArgs<T> A;
  fill_mps_args<T>(p, P, B, x, targets, ws, A);
  A.stash = 1;
  A.loss_scale = (T)loss_scale;
  if (!accumulate) TNB_CUDA(cudaMemsetAsync(grads, 0, (size_t)tnb_param_count(p) * sizeof(T), st));
  if constexpr (std::is_same<T, float>::value) {
    if (P.tc) {
      TcArgs Tc;
      fill_tc_args(P, ws, A, Tc);
      TNB_CUDA(cudaMemsetAsync(Tc.qmax, 0x80, ((size_t)p->L + 1) * sizeof(int), st));
      if (tc_launch_sweep(p, P, Tc, B, 1, st)) return 1;  // seeds through the class site + adjoint chains
      static std::once_flag once;
      static cudaError_t aerr = cudaSuccess;
      std::call_once(once, [] {
        aerr = cudaFuncSetAttribute(tc_da_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 2048);
        if (aerr == cudaSuccess)
          aerr = cudaFuncSetAttribute(tc_da_kernel<kDaMaxCpg>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 2048);
      });
      if (aerr != cudaSuccess) return fail("cudaFuncSetAttribute(tc_da): %s", cudaGetErrorString(aerr));
      TcDaArgs Da;
      Da.M = A;
      Da.grads = (float *)grads;
      Da.fimg = Tc.fimg;
      Da.gimg = Tc.gimg;
      Da.img_bond_bytes = Tc.img_bond_bytes;
      Da.nstage = (int)((232448 - 2048 - 1024) / kDaStage);
      if (tc_knobs().da_stages >= 2 && tc_knobs().da_stages < Da.nstage) Da.nstage = tc_knobs().da_stages;
      Da.tmem_cols = 32;
      while (Da.tmem_cols < 2 * p->chi) Da.tmem_cols <<= 1;
      const long long maxsp = (B + 2047) / 2048;
      Da.ra = p->chi < 128 ? p->chi : 128;
      Da.cpg = 256 / Da.ra;
      for (int pass = 0; pass < 2; ++pass) {
        // pass 0: every chain site (copies sigma = s);  pass 1: the class site (copies sigma = (s, k))
        Da.Dv = pass == 0 ? p->d : p->d * p->n_out;
        Da.only_site = pass == 0 ? -1 : p->out_site;
        Da.wq = pass == 0 ? Tc.wq : Tc.wqc;
        Da.qmax = pass == 0 ? Tc.qmax : Tc.qmaxc;
        Da.mgroups = ((p->chi + Da.ra - 1) / Da.ra) * ((Da.Dv + Da.cpg - 1) / Da.cpg);
        // split K (the batch) so that the grid fills the 148 SMs about four times over with little wave quantisation
        const long long ctas = (long long)(pass == 0 ? p->L - 1 : 1) * Da.mgroups;
        long long best = 1;
        double best_eff = 0.0;
        for (long long nsp = 1; nsp <= maxsp && nsp <= 64; ++nsp) {
          const long long n = ctas * nsp;
          const double eff = (double)n / (double)((n + 147) / 148 * 148);
          const double score = eff - (n < 148 * 3 ? 1.0 : 0.0) - 0.002 * (double)nsp;  // prefer >= 3 waves, fewer splits
          if (nsp == 1 || score > best_eff) { best = nsp; best_eff = score; }
        }
        Da.nsplit = (int)best;
        Da.bchunk = ((B + Da.nsplit - 1) / Da.nsplit + kDaKC - 1) / kDaKC * kDaKC;
        dim3 gda((unsigned)(Da.mgroups * Da.nsplit), (unsigned)(pass == 0 ? p->L : 1));
        prof_begin(st);
        const size_t da_smem = (size_t)Da.nstage * kDaStage + 1024;
        if (Da.cpg == 2) tc_da_kernel<2><<<gda, kDaThreads, da_smem, st>>>(Da);
        else tc_da_kernel<kDaMaxCpg><<<gda, kDaThreads, da_smem, st>>>(Da);
        TNB_LAUNCH_CHECK(pass == 0 ? "tc_da_kernel" : "tc_da_kernel(class)");
      }
      return 0;
    }
  }
  int tiles = (int)((B + P.BT - 1) / P.BT);
  prof_begin(st);
  mps_sweep_kernel<T, NG, RB><<<dim3(tiles, 2), P.NT, P.smem_sweep, st>>>(A, 1);
  TNB_LAUNCH_CHECK("mps_sweep_kernel(adj)");
  DaArgs<T> Dg;
  Dg.M = A;
  Dg.grads = (T *)grads;
  {
    const long long n = (long long)p->L * B;
    prof_begin(st);
    mps_wq_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A);
    TNB_LAUNCH_CHECK("mps_wq_kernel");
  }
  // micro-tile rows: 8 (a 128 x 64 tile per CTA) when the flattened rows (s, a) fill it, else 4; TNB_DA_RM forces
  static const int rm_knob = [] { const char *e = getenv("TNB_DA_RM"); return e ? atoi(e) : 0; }();
  const int Mrows = p->d * p->chi;
  const int RM = rm_knob == 4 || rm_knob == 8 ? rm_knob : (Mrows > 64 ? 8 : 4);
  Dg.tiles_m = (Mrows + 16 * RM - 1) / (16 * RM);
  Dg.tiles_r = (p->chi + 63) / 64;
  long long work = (long long)Dg.tiles_m * Dg.tiles_r * p->L;
  long long ns = (148LL * 4 + work - 1) / work;
  long long maxsplit = (B + 255) / 256;
  if (ns > maxsplit) ns = maxsplit;
  if (ns < 1) ns = 1;
  Dg.nsplit = (int)ns;
  Dg.bchunk = (B + ns - 1) / ns;
  Dg.bchunk = (Dg.bchunk + 15) / 16 * 16;
  Dg.only_site = -1;
  dim3 grid(Dg.tiles_m * Dg.tiles_r * Dg.nsplit, p->L, 1);
  prof_begin(st);
  if (RM == 8) mps_da_kernel<T, 8><<<grid, 256, 0, st>>>(Dg);
  else mps_da_kernel<T, 4><<<grid, 256, 0, st>>>(Dg);
  TNB_LAUNCH_CHECK("mps_da_kernel");
  // class site: one GEMM per class
  Dg.only_site = p->out_site;
  work = (long long)Dg.tiles_m * Dg.tiles_r * p->n_out;
  ns = (148LL * 2 + work - 1) / work;
  if (ns > maxsplit) ns = maxsplit;
  if (ns < 1) ns = 1;
  Dg.nsplit = (int)ns;
  Dg.bchunk = ((B + ns - 1) / ns + 15) / 16 * 16;
  dim3 gridc(Dg.tiles_m * Dg.tiles_r * Dg.nsplit, 1, p->n_out);
  prof_begin(st);
  if (RM == 8) mps_da_kernel<T, 8><<<gridc, 256, 0, st>>>(Dg);
  else mps_da_kernel<T, 4><<<gridc, 256, 0, st>>>(Dg);
  TNB_LAUNCH_CHECK("mps_da_kernel(class)");
  return 0;
}

#define TNB_DISPATCH_MPS(fn, ...)                                                   \
  do {                                                                              \
    if (p->dtype == TNB_F64) {                                                      \
      if (P.NG == 1 && P.RB == 8) return fn<double, 1, 8>(__VA_ARGS__);             \
      if (P.NG == 2 && P.RB == 8) return fn<double, 2, 8>(__VA_ARGS__);             \
      if (P.NG == 1) return fn<double, 1, 4>(__VA_ARGS__);                          \
      if (P.NG == 2) return fn<double, 2, 4>(__VA_ARGS__);                          \
      return fn<double, 4, 4>(__VA_ARGS__);                                         \
    } else {                                                                        \
      if (P.NG == 1) return fn<float, 1, 4>(__VA_ARGS__);                           \
      if (P.NG == 2) return fn<float, 2, 4>(__VA_ARGS__);                           \
      return fn<float, 4, 4>(__VA_ARGS__);                                          \
    }                                                                               \
  } while (0)

}  // namespace tnb

namespace tnb {
template <typename T>
__global__ void adam_kernel(long long n, T *__restrict__ p, const T *__restrict__ g, T *__restrict__ mu, T *__restrict__ nu,
                            T lr, T b1, T b2, T eps, T eps_root, T c1, T c2, T gs) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T gi = g[i] * gs;
    const T m = b1 * mu[i] + (T(1) - b1) * gi;
    const T v = b2 * nu[i] + (T(1) - b2) * gi * gi;
    mu[i] = m;
    nu[i] = v;
    p[i] -= lr * (m * c1) / (Num<T>::sqrt(v * c2 + eps_root) + eps);
  }
}
}  // namespace tnb

using namespace tnb;

extern "C" {

const char *tnb_version(void) { return "tn4ml_b200 0.1.0 (sm_100a)"; }
const char *tnb_last_error(void) { return g_err; }
int64_t tnb_launch_count(void) { return g_launches.load(); }

int32_t tnb_site_nout(const tnb_problem *p, int32_t site) {
  if (!p || site < 0 || site >= p->L) return 0;
  if (p->kind == TNB_KIND_MPS) return site == p->out_site ? p->n_out : 1;
  return site % p->spacing == 0 ? p->n_out : 1;
}

int64_t tnb_site_offset(const tnb_problem *p, int32_t site) {
  if (!p || site < 0 || site > p->L) return -1;
  int64_t per = (int64_t)p->chi * p->chi * p->d;
  int64_t nwith;
  if (p->kind == TNB_KIND_MPS) nwith = site > p->out_site ? 1 : 0;
  else nwith = (site + p->spacing - 1) / p->spacing;  // output sites among [0, site)
  return per * (site - nwith) + per * p->n_out * nwith;
}

int64_t tnb_param_count(const tnb_problem *p) { return p ? tnb_site_offset(p, p->L) : -1; }

int64_t tnb_workspace_bytes(const tnb_problem *p, int64_t batch, int32_t need_grad) {
  if (validate(p)) return -1;
  if (batch < 1) { fail("batch must be >= 1"); return -1; }
  if (p->kind == TNB_KIND_MPS) return (int64_t)plan_mps(p, batch, need_grad).total;
  return (int64_t)plan_smpo(p, batch, need_grad).total;
}

int tnb_forward(const tnb_problem *p, int64_t batch, const void *x, const void *targets, const void *params,
                void *loss, double *loss_sum, void *out, int32_t *out_exp, int32_t keep_stash, void *workspace,
                int64_t workspace_bytes, void *stream) {
  if (validate(p)) return 1;
  if (batch < 1) return fail("batch must be >= 1");
  if (!x || !params || !workspace) return fail("null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->kind == TNB_KIND_MPS) {
    if (p->head != TNB_HEAD_NLL && !targets) return fail("classifier heads need targets");
    MpsPlan P = plan_mps(p, batch, keep_stash);
    if ((int64_t)P.total > workspace_bytes)
      return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)workspace_bytes);
    TNB_DISPATCH_MPS(mps_forward_t, p, P, batch, x, targets, params, loss, loss_sum, out, out_exp, keep_stash,
                     workspace, st);
  }
  SmpoPlan P = plan_smpo(p, batch, keep_stash);
  if ((int64_t)P.total > workspace_bytes)
    return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)workspace_bytes);
  int rc;
  const char *what = "";
  prof_begin(st);
  if (p->dtype == TNB_F64)
    rc = smpo_forward_t<double>(p, P, batch, x, params, loss, loss_sum, out, out_exp, keep_stash, workspace, st, &what);
  else
    rc = smpo_forward_t<float>(p, P, batch, x, params, loss, loss_sum, out, out_exp, keep_stash, workspace, st, &what);
  prof_end("smpo_fwd_kernel", st);
  g_launches.fetch_add(P.launches_fwd, std::memory_order_relaxed);
  if (rc) return fail("%s", what);
  return 0;
}

int tnb_backward(const tnb_problem *p, int64_t batch, const void *x, const void *targets, const void *params,
                 double loss_scale, void *grads, int32_t accumulate, void *workspace, int64_t workspace_bytes,
                 void *stream) {
  if (validate(p)) return 1;
  if (batch < 1) return fail("batch must be >= 1");
  if (!x || !params || !workspace || !grads) return fail("null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->kind == TNB_KIND_MPS) {
    MpsPlan P = plan_mps(p, batch, 1);
    if ((int64_t)P.total > workspace_bytes)
      return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)workspace_bytes);
    TNB_DISPATCH_MPS(mps_backward_t, p, P, batch, x, targets, loss_scale, grads, accumulate, workspace, st);
  }
  SmpoPlan P = plan_smpo(p, batch, 1);
  if ((int64_t)P.total > workspace_bytes)
    return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)workspace_bytes);
  int rc;
  const char *what = "";
  prof_begin(st);
  if (p->dtype == TNB_F64)
    rc = smpo_backward_t<double>(p, P, batch, x, params, loss_scale, grads, accumulate, workspace, st, &what);
  else
    rc = smpo_backward_t<float>(p, P, batch, x, params, loss_scale, grads, accumulate, workspace, st, &what);
  prof_end("smpo_bwd_kernel", st);
  g_launches.fetch_add(P.launches_bwd, std::memory_order_relaxed);
  if (rc) return fail("%s", what);
  return 0;
}

int tnb_profile(int32_t enable) {
  for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = enable != 0;
  return 0;
}

int32_t tnb_profile_read(char *names, float *ms, int32_t max_records) {
  int32_t n = 0;
  for (auto &r : g_prof) {
    if (n >= max_records) break;
    if (cudaEventSynchronize(r.b) != cudaSuccess) break;
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[n] = t;
    snprintf(names + (size_t)n * 48, 48, "%s", r.name);
    ++n;
  }
  return n;
}

#ifdef TNB_TC_TRACE
/* measurement build only: copy the in-kernel timeline of the chain kernel (mps_tc_chain.cuh) to the host */
int tnb_debug_trace_read(long long *host, int32_t n) {
  if (n > 64 * 32) n = 64 * 32;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host, tnb::g_tc_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? 0 : 1;
}
#endif

int tnb_adam_update(int32_t dtype, int64_t n, void *params, const void *grads, void *mu, void *nu, double lr, double b1,
                    double b2, double eps, double eps_root, int64_t count, double grad_scale, void *stream) {
  if (!params || !grads || !mu || !nu) return fail("null device pointer");
  if (n < 1 || count < 1) return fail("n and count must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  const double c1 = 1.0 / (1.0 - pow(b1, (double)count)), c2 = 1.0 / (1.0 - pow(b2, (double)count));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  prof_begin(st);
  if (dtype == TNB_F64)
    adam_kernel<double><<<blocks, 256, 0, st>>>(n, (double *)params, (const double *)grads, (double *)mu, (double *)nu, lr,
                                                 b1, b2, eps, eps_root, c1, c2, grad_scale);
  else if (dtype == TNB_F32)
    adam_kernel<float><<<blocks, 256, 0, st>>>(n, (float *)params, (const float *)grads, (float *)mu, (float *)nu,
                                                (float)lr, (float)b1, (float)b2, (float)eps, (float)eps_root, (float)c1,
                                                (float)c2, (float)grad_scale);
  else
    return fail("bad dtype");
  TNB_LAUNCH_CHECK("adam_kernel");
  return 0;
}

int tnb_embed(const tnb_embedding *emb, int32_t dtype, int64_t n, const void *x, void *phi, void *stream) {
  if (!emb || !x || !phi) return fail("null pointer");
  int d = emb_dim(*emb);
  if (d < 1 || d > TNB_MAX_PHYS) return fail("embedding dimension %d out of range", d);
  if (n < 1) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((n + 255) / 256);
  if (dtype == TNB_F64) {
    EmbParams<double> e;
    fill_emb<double>(*emb, e, d);
    prof_begin(st);
    embed_kernel<double><<<blocks, 256, 0, st>>>(e, n, (const double *)x, (double *)phi);
  } else if (dtype == TNB_F32) {
    EmbParams<float> e;
    fill_emb<float>(*emb, e, d);
    prof_begin(st);
    embed_kernel<float><<<blocks, 256, 0, st>>>(e, n, (const float *)x, (float *)phi);
  } else {
    return fail("bad dtype");
  }
  TNB_LAUNCH_CHECK("embed_kernel");
  return 0;
}

}  // extern "C"
