// Host-side plumbing shared by the translation units of the library: status/error, launch accounting, optional
// per-launch timing, measurement knobs, the MPS workspace plan and the per-family entry points api.cu dispatches to.
// No kernels here.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/tn4ml_b200.h"
#include "consts.h"

namespace tnb {

// ---- status ------------------------------------------------------------------------------------
int fail(const char *fmt, ...);  // formats into the thread-local message read by tnb_last_error(); returns 1

#define TNB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) return ::tnb::fail("%s failed: %s", #expr, cudaGetErrorString(_e));  \
  } while (0)

// ---- launch accounting / optional per-launch timing (tnb_profile) ---------------------------------
void prof_begin(cudaStream_t st);
void prof_end(const char *name, cudaStream_t st);
void count_launches(int n);

#define TNB_LAUNCH_CHECK(name)                                                           \
  do {                                                                                   \
    ::tnb::prof_end(name, st);                                                           \
    ::tnb::count_launches(1);                                                            \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) return ::tnb::fail("launch %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline size_t esize(const tnb_problem *p) { return p->dtype == TNB_F64 ? 8 : 4; }
// sites [s0, s1) of part `part` of `nparts` (contiguous, sizes differ by at most one)
static inline void part_sites(int L, int part, int nparts, int *s0, int *s1) {
  const int base = L / nparts, rem = L % nparts;
  *s0 = part * base + (part < rem ? part : rem);
  *s1 = *s0 + base + (part < rem ? 1 : 0);
}

// ---- measurement knobs: environment, read ONCE when the library is loaded (never on a call path) ----
//   TNB_TC_NH forces a legal column-chunk count, TNB_TC_TILES=1|2 sample tiles per chain CTA, TNB_TC_PACE_NS the image
//   writer's pause between pieces, TNB_DA_STAGES the pipeline depth of the weight-gradient GEMM, TNB_TC_CG=2 the CTA-pair
//   chain kernel; TNB_SIMT_RB / NT / KT / DMMA and TNB_DA_RM the register tiles of the SIMT family; TNB_RR=0 turns the
//   register-resident small-bond family off.  TNB_TC_DBG exists in -DTNB_TC_TRACE builds only.
struct Knobs {
  int tc_nh = 0, tc_tiles = 0, tc_pace_ns = -1, da_stages = 0, tc_cg = 0;
  int simt_rb = 0, simt_nt = 0, simt_kt = 0, simt_dmma = 1, da_rm = 0;
  int rr = 1;
  int tc_pad = 1;    // TNB_TC_PAD=0: bonds that are not multiples of 16 in [32, 256] stay on the SIMT family (no zero padding)
  int rr_da = 1;     // TNB_RR_DA=0: bond-32 weight gradients stay on tc_da_kernel
  int rr_chain = 1;  // TNB_RR_CHAIN=0: bond-32 MPS chains stay on the tcgen05 kernel (two 352-thread CTAs per SM)
  int tc_ngrp2 = 1;   // TNB_TC_NGRP2=0: never run the 352-thread chain kernel (two CTAs per SM) at small steps
  int da_small = 64;  // TNB_DA_SMALL: largest bond whose dA GEMM runs several small CTAs per SM (0: never)
  int da_pw4 = 1;     // bonds up to 32: three CTAs of 4 producer warps per SM (TNB_DA_PW4=0: two of 8)
  int tc_dbg = 0;
};
const Knobs &knobs();
// tnb_set_deterministic: fixed-order gradient reduction (one K-split per output tile: a single writer per gradient entry)
bool deterministic();

// ---- MPS workspace plan --------------------------------------------------------------------------
struct MpsPlan {
  int chip, nslab, TN, NG, NT, BT, KT, WS, dmma_sweep, dmma_class, vst, wsz;
  size_t off_Wn, off_Wt, off_F, off_G, off_exps, off_esum, off_y, off_dy, off_wsimt, total;
  size_t smem_sweep, smem_class;
  int RB;  // register-tile rows of the SIMT sweep / class kernels
  // tensor-core sweep (f32, TNB_PREC_TENSOR)
  int tc, tc_N, tc_NH, tc_NMMA, tc_nstage, tc_cols, tc_tiles, tc_cg, tc_ngrp;
  int tc_chi;  // bond of the tensor path's images and kernels: p->chi, or 32 when a smaller bond runs zero-padded on the register-resident family
  int tc_rr;  // bond 32: the chains run register resident on mma.sync (mps_rr_chain.cuh); images, stash and dA GEMM as planned
  uint32_t tc_stage_bytes;
  size_t tc_site_img, off_img_n, off_img_t, off_kW, off_cs, off_wq, off_qmax, off_wqc, off_fimg, off_gimg, tc_smem, img_bond_bytes;
};
MpsPlan plan_mps(const tnb_problem *p, long long B, int need_grad);

// ---- per-family entry points (each in its own translation unit) --------------------------------------
// SIMT family (mps_simt.cu): float64 and the float32 fallback
int mps_simt_forward(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                     const void *params, void *loss, double *loss_sum, void *out, int32_t *out_exp, int keep_stash,
                     void *ws, cudaStream_t st);
// part / nparts: the backward pass in site ranges (tnb_backward_part): part 0 also runs the adjoint sweep
int mps_simt_backward(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                      double loss_scale, void *grads, int accumulate, void *ws, cudaStream_t st, int part, int nparts);
// two-site gradient of the Sweeps strategy from the stash of forward + backward part 0 (tnb_pair_grad)
int mps_simt_pair_grad(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets, int site,
                       void *dT, void *ws, cudaStream_t st);
cudaError_t mps_simt_set_attrs();
// tcgen05 family (mps_tc.cu): float32, TNB_PREC_TENSOR
int mps_tc_forward(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                   const void *params, void *loss, double *loss_sum, void *out, int32_t *out_exp, int keep_stash,
                   void *ws, cudaStream_t st);
int mps_tc_backward(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                    double loss_scale, void *grads, int accumulate, void *ws, cudaStream_t st, int part, int nparts);
cudaError_t mps_tc_set_attrs();
// register-resident chain at bond 32 (mps_rr.cu): same images and outputs as the tcgen05 chain kernel
cudaError_t mps_rr_set_attrs();
bool mps_rr_supported(const tnb_problem *p);
// SMPO / MPO double layer (smpo.cu)
int64_t smpo_workspace_bytes(const tnb_problem *p, long long B, int need_grad);
int smpo_forward(const tnb_problem *p, long long B, const void *x, const void *params, void *loss, double *loss_sum,
                 void *out, int32_t *out_exp, int keep_stash, void *ws, int64_t ws_bytes, cudaStream_t st);
int smpo_backward(const tnb_problem *p, long long B, const void *x, const void *params, double loss_scale,
                  const void *sample_w, void *grads, int accumulate, void *ws, int64_t ws_bytes, cudaStream_t st);
cudaError_t smpo_set_attrs();
// register-resident small-bond family (rr.cu): SMPO / MPO, float32, bond <= 32, d = 2
cudaError_t rr_set_attrs();
bool rr_smpo_supported(const tnb_problem *p, int need_grad);
int64_t rr_smpo_workspace_bytes(const tnb_problem *p, long long B, int need_grad);
int rr_smpo_forward(const tnb_problem *p, long long B, const void *x, const void *params, void *loss, double *loss_sum,
                    void *out, int32_t *out_exp, int keep_stash, void *ws, int64_t ws_bytes, cudaStream_t st);
int rr_smpo_backward(const tnb_problem *p, long long B, const void *x, const void *params, double loss_scale,
                     const void *sample_w, void *grads, int accumulate, void *ws, int64_t ws_bytes, cudaStream_t st);
// Embedding.__call__ for n scalars (mps_simt.cu)
int embed_launch(const tnb_embedding *emb, int d, int32_t dtype, int64_t n, const void *x, void *phi, cudaStream_t st);

// Every cudaFuncSetAttribute of the library for the CURRENT device (idempotent, per device): called by tnb_init and,
// as a guard, by the first call on a device that was not initialised explicitly.
int ensure_device_ready();

}  // namespace tnb
