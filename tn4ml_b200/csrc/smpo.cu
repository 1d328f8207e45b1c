// SMPO / MPO double-layer path: host-side launch code.  Kernels: smpo_simt.cuh (per-sample matrices in shared memory,
// float64 and the general float32 case) and smpo_rr.cuh (register-resident fp16x3 family, float32, bond <= 32).
#include "host.h"
#include "smpo_simt.cuh"

namespace tnb {

cudaError_t smpo_set_attrs() {
  cudaError_t e = smpo_set_attrs_t<double>();
  if (e == cudaSuccess) e = smpo_set_attrs_t<float>();
  return e;
}

int64_t smpo_workspace_bytes(const tnb_problem *p, long long B, int need_grad) {
  if (rr_smpo_supported(p, need_grad)) return rr_smpo_workspace_bytes(p, B, need_grad);
  return (int64_t)plan_smpo(p, B, need_grad).total;
}

int smpo_forward(const tnb_problem *p, long long B, const void *x, const void *params, void *loss, double *loss_sum,
                 void *out, int32_t *out_exp, int keep_stash, void *ws, int64_t ws_bytes, cudaStream_t st) {
  if (rr_smpo_supported(p, keep_stash))
    return rr_smpo_forward(p, B, x, params, loss, loss_sum, out, out_exp, keep_stash, ws, ws_bytes, st);
  SmpoPlan P = plan_smpo(p, B, keep_stash);
  if ((int64_t)P.total > ws_bytes) return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)ws_bytes);
  int rc;
  const char *what = "";
  prof_begin(st);
  if (p->dtype == TNB_F64)
    rc = smpo_forward_t<double>(p, P, B, x, params, loss, loss_sum, out, out_exp, keep_stash, ws, st, &what);
  else
    rc = smpo_forward_t<float>(p, P, B, x, params, loss, loss_sum, out, out_exp, keep_stash, ws, st, &what);
  prof_end("smpo_fwd_kernel", st);
  count_launches(P.launches_fwd);
  if (rc) return fail("%s", what);
  return 0;
}

int smpo_backward(const tnb_problem *p, long long B, const void *x, const void *params, double loss_scale,
                  const void *sample_w, void *grads, int accumulate, void *ws, int64_t ws_bytes, cudaStream_t st) {
  if (rr_smpo_supported(p, 1))
    return rr_smpo_backward(p, B, x, params, loss_scale, sample_w, grads, accumulate, ws, ws_bytes, st);
  SmpoPlan P = plan_smpo(p, B, 1);
  if ((int64_t)P.total > ws_bytes) return fail("workspace too small: need %zu bytes, got %lld", P.total, (long long)ws_bytes);
  int rc;
  const char *what = "";
  prof_begin(st);
  if (p->dtype == TNB_F64)
    rc = smpo_backward_t<double>(p, P, B, x, params, loss_scale, sample_w, grads, accumulate, ws, st, &what);
  else
    rc = smpo_backward_t<float>(p, P, B, x, params, loss_scale, sample_w, grads, accumulate, ws, st, &what);
  prof_end("smpo_bwd_kernel", st);
  count_launches(P.launches_bwd);
  if (rc) return fail("%s", what);
  return 0;
}

}  // namespace tnb
