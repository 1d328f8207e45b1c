// SIMT family of the MPS path (float64, and the float32 fallback): host-side launch code.
// Kernels: mps_simt.cuh.  Entry points declared in host.h, dispatched from api.cu.
#include <type_traits>

#include "host.h"
#include "mps_args.h"
#include "mps_simt.cuh"

namespace tnb {

template <typename T, int NG, int RB> static cudaError_t set_smem_attrs() {
  cudaError_t e = cudaFuncSetAttribute(mps_sweep_kernel<T, NG, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSimtSmemMax);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mps_class_kernel<T, NG, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSimtSmemMax);
  return e;
}

// every instantiation the dispatcher below can reach, for the current device
cudaError_t mps_simt_set_attrs() {
  cudaError_t e = set_smem_attrs<double, 1, 8>();
  if (e == cudaSuccess) e = set_smem_attrs<double, 2, 8>();
  if (e == cudaSuccess) e = set_smem_attrs<double, 1, 4>();
  if (e == cudaSuccess) e = set_smem_attrs<double, 2, 4>();
  if (e == cudaSuccess) e = set_smem_attrs<double, 4, 4>();
  if (e == cudaSuccess) e = set_smem_attrs<float, 1, 4>();
  if (e == cudaSuccess) e = set_smem_attrs<float, 2, 4>();
  if (e == cudaSuccess) e = set_smem_attrs<float, 4, 4>();
  return e;
}

template <typename T, int NG, int RB>
static int mps_forward_t(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                         const void *params, void *loss, double *loss_sum, void *out, int32_t *out_exp,
                         int keep_stash, void *ws, cudaStream_t st) {
  if (P.smem_class > (size_t)kSimtSmemMax) return fail("shared-memory plan too large (%zu B)", P.smem_class);
  MpsArgs<T> A;
  fill_mps_args<T>(p, P, B, x, targets, ws, A);
  A.stash = keep_stash ? 1 : 0;
  A.loss = (T *)loss; A.loss_sum = loss_sum; A.out = (T *)out; A.out_exp = out_exp;
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
                          double loss_scale, void *grads, int accumulate, void *ws, cudaStream_t st, int part, int nparts) {
  MpsArgs<T> A;
  fill_mps_args<T>(p, P, B, x, targets, ws, A);
  A.stash = 1;
  A.loss_scale = (T)loss_scale;
  DaArgs<T> Dg;
  Dg.M = A;
  Dg.grads = (T *)grads;
  if (part == 0) {
    if (!accumulate) TNB_CUDA(cudaMemsetAsync(grads, 0, (size_t)tnb_param_count(p) * sizeof(T), st));
    int tiles = (int)((B + P.BT - 1) / P.BT);
    prof_begin(st);
    mps_sweep_kernel<T, NG, RB><<<dim3(tiles, 2), P.NT, P.smem_sweep, st>>>(A, 1);
    TNB_LAUNCH_CHECK("mps_sweep_kernel(adj)");
    const long long n = (long long)p->L * B;
    prof_begin(st);
    mps_wq_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A);
    TNB_LAUNCH_CHECK("mps_wq_kernel");
  }
  int s0, s1;
  part_sites(p->L, part, nparts, &s0, &s1);
  if (s1 <= s0) return 0;
  // micro-tile rows: 8 (a 128 x 64 tile per CTA) when the flattened rows (s, a) fill it, else 4; TNB_DA_RM forces
  const int rm_knob = knobs().da_rm;
  const int Mrows = p->d * p->chi;
  const int RM = rm_knob == 4 || rm_knob == 8 ? rm_knob : (Mrows > 64 ? 8 : 4);
  Dg.tiles_m = (Mrows + 16 * RM - 1) / (16 * RM);
  Dg.tiles_r = (p->chi + 63) / 64;
  long long work = (long long)Dg.tiles_m * Dg.tiles_r * (s1 - s0);
  long long ns = (148LL * 4 + work - 1) / work;
  long long maxsplit = (B + 255) / 256;
  if (ns > maxsplit) ns = maxsplit;
  if (ns < 1 || deterministic()) ns = 1;  // (deterministic: a single writer per gradient entry)
  Dg.nsplit = (int)ns;
  Dg.bchunk = (B + ns - 1) / ns;
  Dg.bchunk = (Dg.bchunk + 15) / 16 * 16;
  Dg.only_site = -1;
  Dg.site0 = s0;
  dim3 grid(Dg.tiles_m * Dg.tiles_r * Dg.nsplit, s1 - s0, 1);
  prof_begin(st);
  if (RM == 8) mps_da_kernel<T, 8><<<grid, 256, 0, st>>>(Dg);
  else mps_da_kernel<T, 4><<<grid, 256, 0, st>>>(Dg);
  TNB_LAUNCH_CHECK("mps_da_kernel");
  // class site: one GEMM per class (in the part that owns it)
  if (p->out_site < s0 || p->out_site >= s1) return 0;
  Dg.only_site = p->out_site;
  work = (long long)Dg.tiles_m * Dg.tiles_r * p->n_out;
  ns = (148LL * 2 + work - 1) / work;
  if (ns > maxsplit) ns = maxsplit;
  if (ns < 1 || deterministic()) ns = 1;
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

int mps_simt_forward(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                     const void *params, void *loss, double *loss_sum, void *out, int32_t *out_exp, int keep_stash,
                     void *ws, cudaStream_t st) {
  TNB_DISPATCH_MPS(mps_forward_t, p, P, B, x, targets, params, loss, loss_sum, out, out_exp, keep_stash, ws, st);
}

int mps_simt_backward(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                      double loss_scale, void *grads, int accumulate, void *ws, cudaStream_t st, int part, int nparts) {
  TNB_DISPATCH_MPS(mps_backward_t, p, P, B, x, targets, loss_scale, grads, accumulate, ws, st, part, nparts);
}

// tnb_pair_grad: the two-site gradient of the Sweeps strategy from the stash of forward + tnb_backward_part(part 0)
template <typename T>
static int mps_pair_t(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets, int site,
                      void *dT, void *ws, cudaStream_t st) {
  MpsArgs<T> A;
  fill_mps_args<T>(p, P, B, x, targets, ws, A);
  A.stash = 1;
  const bool cls = site == p->out_site || site + 1 == p->out_site;
  const int nk = cls ? p->n_out : 1;
  const int tiles = (p->chi + 15) / 16;
  const long long work = (long long)tiles * tiles * p->d * p->d * nk;
  long long ns = (148LL * 2 + work - 1) / work;
  const long long maxsplit = (B + 255) / 256;
  if (ns > maxsplit) ns = maxsplit;
  if (ns < 1 || deterministic()) ns = 1;
  long long bchunk = ((B + ns - 1) / ns + 31) / 32 * 32;
  if (ns > 1) TNB_CUDA(cudaMemsetAsync(dT, 0, (size_t)p->chi * p->chi * p->d * p->d * nk * sizeof(T), st));
  dim3 grid((unsigned)(tiles * tiles * ns), (unsigned)(p->d * p->d), (unsigned)nk);
  prof_begin(st);
  mps_pair_kernel<T><<<grid, 256, 0, st>>>(A, site, (T *)dT, (int)ns, bchunk);
  TNB_LAUNCH_CHECK("mps_pair_kernel");
  return 0;
}

int mps_simt_pair_grad(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets, int site,
                       void *dT, void *ws, cudaStream_t st) {
  if (p->dtype == TNB_F64) return mps_pair_t<double>(p, P, B, x, targets, site, dT, ws, st);
  return mps_pair_t<float>(p, P, B, x, targets, site, dT, ws, st);
}

// Embedding.__call__ for n scalars (tnb_embed)
int embed_launch(const tnb_embedding *emb, int d, int32_t dtype, int64_t n, const void *x, void *phi, cudaStream_t st) {
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

}  // namespace tnb
