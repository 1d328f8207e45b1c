// Register-resident small-bond family (float32, fp16x3 on mma.sync): host-side launch code.  Kernels: smpo_rr.cuh.
#include "host.h"
#include "smpo_rr.cuh"

namespace tnb {

using rr::RrArgs;

static size_t bwd_smem(const tnb_problem *p) {
  return (size_t)(rr::kBwdWarps + p->n_out + p->spacing - 1) * 2048 * sizeof(float) + (size_t)rr::kRingStages * rr::kImgFloat4 * sizeof(float4);
}

cudaError_t rr_set_attrs() {
  cudaError_t e = cudaFuncSetAttribute(rr::smpo_rr_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rr::smpo_rr_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax);
  return e;
}

bool rr_smpo_supported(const tnb_problem *p, int need_grad) {
  if (!knobs().rr) return false;
  if (p->kind != TNB_KIND_SMPO || p->dtype != TNB_F32) return false;
  if (p->chi > rr::kN || p->d != 2 || p->n_out > 4) return false;
  // the four headline feature maps with one scalar per site (the others go through the shared-memory family: this family's VJP
  // kernel sits at the 255-register limit and an out-of-line feature-map call costs it 15 %)
  if (p->emb.kind > TNB_EMB_RBF || p->emb.n_in > 1) return false;
  if (p->spacing > 32) return false;  // (phi of a window's sites: one lane per site)
  if (need_grad && bwd_smem(p) > (size_t)kTcSmemMax) return false;  // (the window's gradient accumulator lives in shared memory)
  return true;
}

struct RrPlan {
  int nwin, per_cta, grid_bwd, nunit;
  size_t off_E, off_Eexp, off_esum, off_nm, off_dE, off_dEexp, off_Q, off_Wimg, off_Wsh, total;
};

static RrPlan plan_rr(const tnb_problem *p, long long B, int need_grad) {
  RrPlan P{};
  P.nwin = (p->L + p->spacing - 1) / p->spacing;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  P.off_E = take(need_grad ? (size_t)P.nwin * B * 1024 * sizeof(float) : 0);
  P.off_Eexp = take(need_grad ? (size_t)P.nwin * B * sizeof(int) : 0);
  P.off_esum = take((size_t)B * sizeof(int));
  P.off_nm = take((size_t)B * sizeof(float));
  P.nunit = P.nwin * (p->spacing - 1 + p->n_out);
  P.off_Wimg = take((size_t)P.nunit * 2 * rr::kImgFloat4 * sizeof(float4));
  P.off_Wsh = take((size_t)P.nunit * sizeof(int));
  // VJP: samples per CTA so that one wave of 148 CTAs covers the batch (whole rounds of kBwdWarps samples where possible)
  long long per = (B + 147) / 148;
  per = (per + rr::kBwdWarps - 1) / rr::kBwdWarps * rr::kBwdWarps;
  if (per > B) per = (B + rr::kBwdWarps - 1) / rr::kBwdWarps * rr::kBwdWarps;
  P.per_cta = (int)per;
  P.grid_bwd = (int)((B + per - 1) / per);
  P.off_dE = take(need_grad ? (size_t)B * 1024 * sizeof(float) : 0);
  P.off_dEexp = take(need_grad ? (size_t)B * sizeof(int) : 0);
  P.off_Q = take(need_grad ? (size_t)P.grid_bwd * rr::kBwdWarps * p->spacing * 1056 * sizeof(uint32_t) : 0);
  P.total = o;
  return P;
}

int64_t rr_smpo_workspace_bytes(const tnb_problem *p, long long B, int need_grad) { return (int64_t)plan_rr(p, B, need_grad).total; }

static void fill(const tnb_problem *p, const RrPlan &P, long long B, const void *x, const void *params, void *ws, RrArgs &A) {
  char *w = (char *)ws;
  A.L = p->L; A.chi = p->chi; A.D = p->d; A.DO = p->n_out; A.S = p->spacing; A.nwin = P.nwin; A.head = p->head; A.B = B;
  fill_emb<float>(p->emb, A.emb, p->d);
  A.x = (const float *)x; A.params = (const float *)params; A.sw = nullptr;
  A.Estash = (float *)(w + P.off_E); A.Eexp = (int *)(w + P.off_Eexp); A.esum = (int *)(w + P.off_esum);
  A.nm = (float *)(w + P.off_nm);
  A.dEscr = (float *)(w + P.off_dE); A.dEexp = (int *)(w + P.off_dEexp); A.Qscr = (uint32_t *)(w + P.off_Q);
  A.per_cta = P.per_cta;
  A.Wimg = (const float4 *)(w + P.off_Wimg); A.Wsh = (const int *)(w + P.off_Wsh);
  A.loss = nullptr; A.loss_sum = nullptr; A.out = nullptr; A.out_exp = nullptr; A.grads = nullptr; A.loss_scale = 1.f;
  A.stash = 0;
}

int rr_smpo_forward(const tnb_problem *p, long long B, const void *x, const void *params, void *loss, double *loss_sum,
                    void *out, int32_t *out_exp, int keep_stash, void *ws, int64_t ws_bytes, cudaStream_t st) {
  RrPlan P = plan_rr(p, B, keep_stash);
  if ((int64_t)P.total > ws_bytes) return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)ws_bytes);
  RrArgs A;
  fill(p, P, B, x, params, ws, A);
  A.stash = keep_stash ? 1 : 0;
  A.loss = (float *)loss; A.loss_sum = loss_sum; A.out = (float *)out; A.out_exp = out_exp;
  // weight images in fragment order + the a-priori power-of-two scale of every unit (once per call: 16 KB per unit)
  prof_begin(st);
  rr::smpo_rr_unitmax_kernel<<<P.nunit, 256, 0, st>>>((const float *)params, (int *)A.Wsh, p->L, p->chi, p->spacing, p->n_out);
  TNB_LAUNCH_CHECK("smpo_rr_unitmax_kernel");
  {
    const long long total = (long long)P.nunit * 2 * rr::kImgFloat4;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    prof_begin(st);
    rr::smpo_rr_prep_kernel<<<blocks, 256, 0, st>>>((const float *)params, A.Wsh, (float4 *)A.Wimg, P.nunit, p->L, p->chi,
                                                    p->spacing, p->n_out);
    TNB_LAUNCH_CHECK("smpo_rr_prep_kernel");
  }
  const unsigned grid = (unsigned)((B + rr::kFwdWarps - 1) / rr::kFwdWarps);
  prof_begin(st);
  if (p->n_out == 2) rr::smpo_rr_fwd_kernel<2><<<grid, rr::kFwdWarps * 32, 0, st>>>(A);
  else rr::smpo_rr_fwd_kernel<0><<<grid, rr::kFwdWarps * 32, 0, st>>>(A);
  TNB_LAUNCH_CHECK("smpo_rr_fwd_kernel");
  return 0;
}

int rr_smpo_backward(const tnb_problem *p, long long B, const void *x, const void *params, double loss_scale,
                     const void *sample_w, void *grads, int accumulate, void *ws, int64_t ws_bytes, cudaStream_t st) {
  RrPlan P = plan_rr(p, B, 1);
  if ((int64_t)P.total > ws_bytes) return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)ws_bytes);
  RrArgs A;
  fill(p, P, B, x, params, ws, A);
  A.stash = 1;
  A.grads = (float *)grads;
  A.loss_scale = (float)loss_scale;
  A.sw = (const float *)sample_w;
  if (!accumulate) TNB_CUDA(cudaMemsetAsync(grads, 0, (size_t)tnb_param_count(p) * sizeof(float), st));
  const size_t smem = bwd_smem(p);
  prof_begin(st);
  if (p->n_out == 2) rr::smpo_rr_bwd_kernel<2><<<P.grid_bwd, rr::kBwdWarps * 32, smem, st>>>(A);
  else rr::smpo_rr_bwd_kernel<0><<<P.grid_bwd, rr::kBwdWarps * 32, smem, st>>>(A);
  TNB_LAUNCH_CHECK("smpo_rr_bwd_kernel");
  return 0;
}

}  // namespace tnb
