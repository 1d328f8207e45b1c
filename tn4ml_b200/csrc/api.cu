// C-ABI of the tn4ml hot path (see include/tn4ml_b200.h for the contract): validation, workspace planning, dispatch to
// the kernel families (mps_simt.cu, mps_tc.cu, smpo.cu, tn_ops.cu), the fused optimizer update and the layout helpers.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host.h"

namespace tnb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Optional per-launch timing (bench.py's roofline leg): when enabled, every kernel launch is bracketed by CUDA events on
// the launching stream.  Off by default (one relaxed atomic load per launch); records are kept under a mutex.
struct ProfRec { const char *name; cudaEvent_t a, b; cudaStream_t st; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
void prof_begin(cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{"?", nullptr, nullptr, st};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}
void prof_end(const char *name, cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (size_t i = g_prof.size(); i-- > 0;)  // the open record of this stream
    if (g_prof[i].st == st && g_prof[i].name[0] == '?') {
      g_prof[i].name = name;
      cudaEventRecord(g_prof[i].b, st);
      break;
    }
}

// Measurement knobs: read once, when the library is loaded (static initialisation) -- no getenv on any call path.
static Knobs read_knobs() {
  Knobs k;
  auto geti = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
  k.tc_nh = geti("TNB_TC_NH", 0);
  k.tc_tiles = geti("TNB_TC_TILES", 0);
  k.tc_pace_ns = geti("TNB_TC_PACE_NS", -1);
  k.tc_ngrp2 = geti("TNB_TC_NGRP2", 1);
  k.da_small = geti("TNB_DA_SMALL", 64);
  k.da_pw4 = geti("TNB_DA_PW4", 1);
  k.da_stages = geti("TNB_DA_STAGES", 0);
  k.tc_cg = geti("TNB_TC_CG", 0);  // 0: the default rule of plan_mps; 1: never the CTA pair; 2: wherever it is legal
  k.simt_rb = geti("TNB_SIMT_RB", 0);
  k.simt_nt = geti("TNB_SIMT_NT", 0);
  k.simt_kt = geti("TNB_SIMT_KT", 0);
  k.simt_dmma = geti("TNB_SIMT_DMMA", 1);
  k.da_rm = geti("TNB_DA_RM", 0);
  k.rr = geti("TNB_RR", 1);
  k.rr_chain = geti("TNB_RR_CHAIN", 1);
  k.rr_da = geti("TNB_RR_DA", 1);
  k.tc_pad = geti("TNB_TC_PAD", 1);
#ifdef TNB_TC_TRACE
  k.tc_dbg = geti("TNB_TC_DBG", 0);
#endif
  return k;
}
static const Knobs g_knobs = read_knobs();
const Knobs &knobs() { return g_knobs; }
static std::atomic<int> g_deterministic{getenv("TNB_DETERMINISTIC") ? atoi(getenv("TNB_DETERMINISTIC")) : 0};
bool deterministic() { return g_deterministic.load(std::memory_order_relaxed) != 0; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: done by tnb_init(device), and guarded here for callers
// that skipped it (one relaxed load per call once the device is ready).
static std::atomic<int> g_dev_ready[64];
static std::mutex g_dev_mu;
int ensure_device_ready() {
  int dev = 0;
  TNB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail("device index %d out of range", dev);
  if (g_dev_ready[dev].load(std::memory_order_acquire)) return 0;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_dev_ready[dev].load(std::memory_order_relaxed)) return 0;
  cudaDeviceProp prop;
  TNB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail("tn4ml_b200 is built for sm_100a; device %d is sm_%d%d", dev, prop.major, prop.minor);
  cudaError_t e = mps_simt_set_attrs();
  if (e == cudaSuccess) e = mps_tc_set_attrs();
  if (e == cudaSuccess) e = smpo_set_attrs();
  if (e == cudaSuccess) e = rr_set_attrs();
  if (e == cudaSuccess) e = mps_rr_set_attrs();
  if (e != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  g_dev_ready[dev].store(1, std::memory_order_release);
  return 0;
}

static int emb_dim(const tnb_embedding &e) {
  switch (e.kind) {
    case TNB_EMB_TRIG: return 2 * e.k;
    case TNB_EMB_FOURIER: return e.p;
    case TNB_EMB_POLY: {
      const int n = e.n_in > 0 ? e.n_in : 1;
      long long total = e.include_bias ? 1 : 0;
      for (int dg = 1; dg <= e.degree; ++dg) {
        if (!e.cross) { total += n; continue; }
        long long c = 1;  // C(n + dg - 1, dg)
        for (int q = 1; q <= dg; ++q) c = c * (n + dg - q) / q;
        total += c;
      }
      return total > TNB_MAX_PHYS ? -1 : (int)total;
    }
    case TNB_EMB_RBF: return e.n_centers;
    case TNB_EMB_LINCOMP: return e.p == 2 || e.p == 3 ? e.p : -1;
    case TNB_EMB_LEGENDRE:
    case TNB_EMB_LAGUERRE:
    case TNB_EMB_HERMITE: return e.degree + 1;
    case TNB_EMB_ARRAYS: return (e.n_in > 0 ? e.n_in : 1) + (e.include_bias ? 1 : 0);
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

// Bond the tensor path runs at.  The tensor kernels take bonds that are multiples of 16 from 32 to 256; any other float32 bond
// from 8 up runs ZERO-PADDED to the next such bond: the weight images are built padded from the caller's [chi][chi][d] slabs, the
// environments' extra entries stay exactly zero, and the gradient kernels write the caller's layout.  Bonds below 32 with
// d = 2 or 3 land on the register-resident mma.sync family (measured against the SIMT fp32 family at L = 196: chi = 10: 0.68 ->
// 0.35 ms per step at B = 64, 9.2 M -> 31.8 M samples/s at B = 65536; chi = 24: 6.0 M -> 31.3 M samples/s).
static int tc_bond(const tnb_problem *p) {
  if (p->chi < 8 || p->chi > 256 || !knobs().tc_pad) return p->chi;
  const int up = (p->chi + 15) / 16 * 16;
  return up < 32 ? 32 : up;
}

// The tcgen05 sweep covers: float32, chi a multiple of 16 in [32, 256], D*chi <= 512 TMEM columns.
static bool tc_supported(const tnb_problem *p) {
  if (p->kind != TNB_KIND_MPS || p->dtype != TNB_F32 || p->precision != TNB_PREC_TENSOR) return false;
  const int chi = tc_bond(p);
  if (chi % 16 != 0 || chi < 32 || chi > 256) return false;
  if (p->d * chi > 512) return false;
  if (!(p->d == 2 || p->d == 3 || p->d == 4 || p->d == 6 || p->d == 8)) return false;
  return true;
}

// Column chunks per chain step: the smallest NH dividing chi/16 whose chunk (D*chi/NH columns) fits one MMA (N <= 256).
static int tc_pick_nh(const tnb_problem *p, int chi) {
  const int groups = chi / 16;
  auto legal = [&](int nh) { return nh >= 1 && nh <= kTcMaxNH && groups % nh == 0 && p->d * chi / nh <= 256; };
  if (legal(knobs().tc_nh)) return knobs().tc_nh;
  for (int nh = 1; nh <= groups; ++nh)
    if (legal(nh)) return nh;
  return groups;
}

MpsPlan plan_mps(const tnb_problem *p, long long B, int need_grad) {
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
  const int rb_knob = knobs().simt_rb;
  P.RB = 4;
  if (p->dtype == TNB_F64 && P.NG <= 2 && rb_knob == 8) {  // (measured slower at chi = 128: 204 registers, one CTA per SM)
    const long long bt8 = (long long)(256 / P.TN) * 8;
    if ((B + bt8 - 1) / bt8 >= 148) P.RB = 8;
  }
  const int nt_knob = knobs().simt_nt;
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
  const int kt_knob = knobs().simt_kt;
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
  // decide the kernel family (incl. the shared-memory check of the tensor path) BEFORE laying out the workspace, so that
  // the SIMT buffers always exist when the SIMT family will run
  P.tc = tc_supported(p) ? 1 : 0;
  P.tc_chi = tc_bond(p);
  if (P.tc) {
    const int tchi = P.tc_chi;
    P.tc_N = p->d * tchi;
    P.tc_NH = tc_pick_nh(p, tchi);
    P.tc_NMMA = P.tc_N / P.tc_NH;
    P.tc_stage_bytes = 64u * P.tc_NMMA;  // one K-slab (16 rows) of one column chunk, hi and lo planes
    P.tc_site_img = (size_t)P.tc_NH * (tchi / 16) * P.tc_stage_bytes;  // (stage_bytes: still the whole stage here)
    const size_t a_bytes = (size_t)2 * kTcRows * tchi * 2;
    // two sample tiles per CTA when one column chunk covers the step (nothing else hides the epilogue); TNB_TC_TILES=1 forces one
    // (pays when the step's MMA time is comparable to its exposed tail: N = 256; measured slower at N <= 128)
    P.tc_tiles = (P.tc_NH == 1 && P.tc_N == 256) ? 2 : 1;
    if (knobs().tc_tiles == 1) P.tc_tiles = 1;
    if (knobs().tc_tiles == 2 && P.tc_NH == 1 && (p->d == 2 || p->d == 4 || p->d == 8)) P.tc_tiles = 2;
    // CTA pair (cta_group::2): legal for the wide-bond case (chi > 128, d = 2); the default at chi = 256, where it was measured
    // (same box, four alternations of 10-step runs: 59.8 -> 58.7 ms per step, e2e 2.18 -> 2.22 M samples/s; chain kernels
    // 19.5 + 21.2 -> 18.8 + 20.2 ms).  TNB_TC_CG=1 turns it off, TNB_TC_CG=2 takes it wherever it is legal.
    {
      const bool legal = tchi > 128 && p->d == 2 && P.tc_tiles == 1;
      const int want = knobs().tc_cg == 0 ? (tchi == 256 ? 2 : 1) : knobs().tc_cg;
      P.tc_cg = (legal && want == 2) ? 2 : 1;
    }
    P.tc_stage_bytes /= (uint32_t)P.tc_cg;  // one CTA's half of a stage
    P.tc_cols = 32;
    while (P.tc_cols < P.tc_tiles * P.tc_N) P.tc_cols <<= 1;
    // small steps (one column chunk of at most 256 columns, bond up to 64, one tile): the 352-thread kernel, two CTAs per SM,
    // each with half of the shared memory and at most 256 TMEM columns
    P.tc_ngrp = (P.tc_NH == 1 && P.tc_N <= 256 && tchi <= 64 && P.tc_tiles == 1 && P.tc_cg == 1 && knobs().tc_ngrp2) ? 2 : 4;
    const size_t fixed = (size_t)P.tc_tiles * 4096 /* row exchange */ + 1024 /* tables, barriers */;
    const long smem_cap = P.tc_ngrp == 2 ? (long)kTcSmemMax / 2 - 1024 : (long)kTcSmemMax;
    long ns = (long)((smem_cap - (long)fixed - (long)(P.tc_tiles * a_bytes)) / (long)P.tc_stage_bytes);
    P.tc_nstage = ns > 12 ? 12 : (int)ns;
    P.tc_smem = P.tc_tiles * a_bytes + (size_t)P.tc_nstage * P.tc_stage_bytes + fixed;
    if (P.tc_nstage < 2) P.tc = 0;
    P.tc_rr = (P.tc && P.tc_NH == 1 && P.tc_cg == 1 && tchi == 32 && mps_rr_supported(p)) ? 1 : 0;
  }
  const bool tc_mode = P.tc != 0;
  size_t fslots = need_grad ? (size_t)p->L + 1 : 2;
  P.off_F = tc_mode ? 0 : take(fslots * B * P.chip * w);
  P.off_G = (need_grad && !tc_mode) ? take(((size_t)p->L + 1) * B * P.chip * w) : 0;
  P.off_exps = take((size_t)p->L * B * sizeof(int));
  P.off_esum = take((size_t)2 * B * sizeof(int));
  P.off_y = take((size_t)B * p->n_out * w);
  P.off_dy = need_grad ? take((size_t)B * p->n_out * w) : 0;
  P.off_wsimt = (need_grad && !tc_mode) ? take((size_t)p->L * p->d * B * w) : 0;  // per-sample dA weights of the SIMT family
  if (P.tc) {
    // the tensor path keeps its environments as fp16 hi/lo bond images instead of F / G
    const size_t span = (size_t)kTcRows * P.tc_cg;  // a CTA pair always writes both tiles
    P.img_bond_bytes = (size_t)((B + span - 1) / span * span) * P.tc_chi * 4;  // whole 128-sample tiles
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
  // float64, d = 2, one or two column groups, full 256-thread CTAs: the step GEMM runs as DMMA warp tiles and needs one more
  // [BT][chip] tile behind everything else (TNB_SIMT_DMMA=0 keeps the FMA loop).  With it the environment tile, the weight
  // segments and the output tile get 4 elements of padding per row (bank conflicts of the fragment loads, mps_simt.cuh).
  const int dmma_knob = knobs().simt_dmma;
  const bool dmma_ok = p->dtype == TNB_F64 && p->d >= 2 && P.NG <= 2 && P.RB == 4 && P.TN >= 8 && P.NT == 256 &&
                       (P.KT % 4 == 0 || (long)P.KT * P.NG * p->d / 2 >= 4) && dmma_knob;
  for (int pad = dmma_ok ? 4 : 0; pad >= 0; pad -= 4) {
    P.vst = P.chip + pad;
    P.wsz = 2 * P.KT * p->d * P.WS + (pad ? 2 * P.KT * P.NG * p->d * 4 : 0);
    size_t base = ((size_t)P.BT * P.vst + (size_t)P.wsz + (size_t)P.BT * p->d + (size_t)P.BT * p->n_out) * w;
    P.smem_sweep = base;
    P.smem_class = base + (size_t)P.BT * P.chip * w;
    P.dmma_sweep = P.dmma_class = 0;
    if (!dmma_ok) break;
    const size_t tile = (size_t)P.BT * P.vst * w, cap = kSimtSmemMax;  // (each kernel only if it still fits)
    if (align_up(P.smem_sweep, 16) + tile <= cap) {
      P.dmma_sweep = (int)align_up(P.smem_sweep, 16);
      P.smem_sweep = P.dmma_sweep + tile;
    }
    if (align_up(P.smem_class, 16) + tile <= cap) {
      P.dmma_class = (int)align_up(P.smem_class, 16);
      P.smem_class = P.dmma_class + tile;
    }
    if (pad == 0 || (P.dmma_sweep && (P.dmma_class || P.chip > 128))) break;  // else: retry without the padding
  }
  return P;
}

}  // namespace tnb

using namespace tnb;

extern "C" {

const char *tnb_version(void) { return "tn4ml_b200 0.1.0 (sm_100a)"; }
const char *tnb_last_error(void) { return g_err; }
int64_t tnb_launch_count(void) { return g_launches.load(); }
int tnb_set_deterministic(int32_t on) { g_deterministic.store(on ? 1 : 0); return 0; }

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

int tnb_init(int32_t device) {
  int prev = 0;
  TNB_CUDA(cudaGetDevice(&prev));
  TNB_CUDA(cudaSetDevice(device));
  const int rc = ensure_device_ready();
  cudaSetDevice(prev);
  return rc;
}

int64_t tnb_workspace_bytes(const tnb_problem *p, int64_t batch, int32_t need_grad) {
  if (validate(p)) return -1;
  if (batch < 1) { fail("batch must be >= 1"); return -1; }
  if (p->kind == TNB_KIND_MPS) return (int64_t)plan_mps(p, batch, need_grad).total;
  return smpo_workspace_bytes(p, batch, need_grad);
}

int tnb_forward(const tnb_problem *p, int64_t batch, const void *x, const void *targets, const void *params,
                void *loss, double *loss_sum, void *out, int32_t *out_exp, int32_t keep_stash, void *workspace,
                int64_t workspace_bytes, void *stream) {
  if (validate(p)) return 1;
  if (batch < 1) return fail("batch must be >= 1");
  if (!x || !params || !workspace) return fail("null device pointer");
  if (ensure_device_ready()) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->kind == TNB_KIND_MPS) {
    if (p->head != TNB_HEAD_NLL && !targets) return fail("classifier heads need targets");
    MpsPlan P = plan_mps(p, batch, keep_stash);
    if ((int64_t)P.total > workspace_bytes)
      return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)workspace_bytes);
    if (P.tc)
      return mps_tc_forward(p, P, batch, x, targets, params, loss, loss_sum, out, out_exp, keep_stash, workspace, st);
    return mps_simt_forward(p, P, batch, x, targets, params, loss, loss_sum, out, out_exp, keep_stash, workspace, st);
  }
  return smpo_forward(p, batch, x, params, loss, loss_sum, out, out_exp, keep_stash, workspace, workspace_bytes, st);
}

static int backward_impl(const tnb_problem *p, int64_t batch, const void *x, const void *targets, const void *params,
                         double loss_scale, void *grads, int32_t accumulate, void *workspace, int64_t workspace_bytes,
                         void *stream, int part, int nparts) {
  if (validate(p)) return 1;
  if (batch < 1) return fail("batch must be >= 1");
  if (!x || !params || !workspace || !grads) return fail("null device pointer");
  if (nparts < 1 || part < 0 || part >= nparts) return fail("part %d of %d is out of range", part, nparts);
  if (ensure_device_ready()) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->kind == TNB_KIND_MPS) {
    MpsPlan P = plan_mps(p, batch, 1);
    if ((int64_t)P.total > workspace_bytes)
      return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)workspace_bytes);
    if (P.tc) return mps_tc_backward(p, P, batch, x, targets, loss_scale, grads, accumulate, workspace, st, part, nparts);
    return mps_simt_backward(p, P, batch, x, targets, loss_scale, grads, accumulate, workspace, st, part, nparts);
  }
  if (part > 0) return 0;  // the double-layer VJP is one kernel over all sites: part 0 computes everything
  // SMPO / MPO heads take no targets: the `targets` argument carries optional per-sample loss weights [batch] instead
  return smpo_backward(p, batch, x, params, loss_scale, targets, grads, accumulate, workspace, workspace_bytes, st);
}

int tnb_backward(const tnb_problem *p, int64_t batch, const void *x, const void *targets, const void *params,
                 double loss_scale, void *grads, int32_t accumulate, void *workspace, int64_t workspace_bytes,
                 void *stream) {
  return backward_impl(p, batch, x, targets, params, loss_scale, grads, accumulate, workspace, workspace_bytes, stream, 0, 1);
}

int tnb_backward_part(const tnb_problem *p, int64_t batch, const void *x, const void *targets, const void *params,
                      double loss_scale, void *grads, int32_t accumulate, void *workspace, int64_t workspace_bytes,
                      void *stream, int32_t part, int32_t nparts) {
  return backward_impl(p, batch, x, targets, params, loss_scale, grads, accumulate, workspace, workspace_bytes, stream,
                       part, nparts);
}

int tnb_pair_grad(const tnb_problem *p, int64_t batch, const void *x, const void *targets, int32_t site, void *dT,
                  void *workspace, int64_t workspace_bytes, void *stream) {
  if (validate(p)) return 1;
  if (batch < 1) return fail("batch must be >= 1");
  if (!x || !dT || !workspace) return fail("null device pointer");
  if (p->kind != TNB_KIND_MPS) return fail("tnb_pair_grad: MPS models only (the reference's Sweeps strategy, strategy.py)");
  if (site < 0 || site + 1 >= p->L) return fail("pair (%d, %d) out of range", site, site + 1);
  if (ensure_device_ready()) return 1;
  MpsPlan P = plan_mps(p, batch, 1);
  if (P.tc) return fail("tnb_pair_grad reads the SIMT family's environment stash: set precision = TNB_PREC_SIMT");
  if ((int64_t)P.total > workspace_bytes)
    return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)workspace_bytes);
  return mps_simt_pair_grad(p, P, batch, x, targets, site, dT, workspace, (cudaStream_t)stream);
}

int32_t tnb_part_sites(const tnb_problem *p, int32_t part, int32_t nparts, int32_t *site_begin, int32_t *site_end) {
  if (!p || !site_begin || !site_end || nparts < 1 || part < 0 || part >= nparts) return fail("bad arguments");
  int s0, s1;
  part_sites(p->L, part, nparts, &s0, &s1);
  if (p->kind != TNB_KIND_MPS) { s0 = part == 0 ? 0 : p->L; s1 = p->L; }  // one part owns every site
  *site_begin = s0;
  *site_end = s1;
  return 0;
}

// Ragged reference-shaped site tensors <-> the zero-padded packed layout: site i is C-contiguous
// (chi_l[i], chi_r[i], d, n_i); its slab is (chi, chi, d, n_i).  One strided device copy per site (row a of the site =
// chi_r*d*n_i contiguous elements at slab pitch chi*d*n_i); stream ordered, capturable.
static int pack_copy(const tnb_problem *p, int to_packed, const void *const *sites, const int32_t *chi_l,
                     const int32_t *chi_r, void *packed, cudaStream_t st) {
  if (!p || !sites || !chi_l || !chi_r || !packed) return fail("null pointer");
  const size_t w = esize(p);
  if (to_packed) TNB_CUDA(cudaMemsetAsync(packed, 0, (size_t)tnb_param_count(p) * w, st));
  for (int i = 0; i < p->L; ++i) {
    if (!sites[i]) return fail("site %d: null pointer", i);
    if (chi_l[i] < 1 || chi_l[i] > p->chi || chi_r[i] < 1 || chi_r[i] > p->chi)
      return fail("site %d: bonds (%d, %d) do not fit chi = %d", i, chi_l[i], chi_r[i], p->chi);
    const size_t inner = (size_t)p->d * tnb_site_nout(p, i) * w;
    char *slab = (char *)packed + (size_t)tnb_site_offset(p, i) * w;
    const size_t width = (size_t)chi_r[i] * inner, pitch = (size_t)p->chi * inner;
    if (to_packed)
      TNB_CUDA(cudaMemcpy2DAsync(slab, pitch, sites[i], width, width, (size_t)chi_l[i], cudaMemcpyDeviceToDevice, st));
    else
      TNB_CUDA(cudaMemcpy2DAsync((void *)sites[i], width, slab, pitch, width, (size_t)chi_l[i], cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int tnb_pack_params(const tnb_problem *p, const void *const *sites, const int32_t *chi_l, const int32_t *chi_r,
                    void *packed, void *stream) {
  return pack_copy(p, 1, sites, chi_l, chi_r, packed, (cudaStream_t)stream);
}

int tnb_unpack_grads(const tnb_problem *p, const void *packed, void *const *sites, const int32_t *chi_l,
                     const int32_t *chi_r, void *stream) {
  return pack_copy(p, 0, (const void *const *)sites, chi_l, chi_r, (void *)packed, (cudaStream_t)stream);
}

int tnb_profile(int32_t enable) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on.store(enable != 0);
  return 0;
}

int32_t tnb_profile_read(char *names, float *ms, int32_t max_records) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int32_t n = 0;
  for (auto &r : g_prof) {
    if (n >= max_records) break;
    if (r.name[0] == '?') continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) break;
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[n] = t;
    snprintf(names + (size_t)n * 48, 48, "%s", r.name);
    ++n;
  }
  return n;
}

int tnb_embed(const tnb_embedding *emb, int32_t dtype, int64_t n, const void *x, void *phi, void *stream) {
  if (!emb || !x || !phi) return fail("null pointer");
  int d = emb_dim(*emb);
  if (d < 1 || d > TNB_MAX_PHYS) return fail("embedding dimension %d out of range", d);
  if (n < 1) return 0;
  return embed_launch(emb, d, dtype, n, x, phi, (cudaStream_t)stream);
}

}  // extern "C"
