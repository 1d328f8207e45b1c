// Update side of train_step (tn4ml/models/model.py:591-614) on the packed buffers: per-site gradient clipping
// (tn4ml/util.py:131-155) and optax.adam as one kernel each, instead of eager optax over an L-leaf pytree.
#include <cmath>

#include "common.cuh"
#include "host.h"

namespace tnb {

// gs_dev (may be null): one more gradient factor read from device memory, so that a factor computed on the device
// (1 / global batch after an all-reduce, a global-norm clip) never goes through the host.
template <typename T>
__global__ void adam_kernel(long long n, T *__restrict__ p, const T *__restrict__ g, T *__restrict__ mu, T *__restrict__ nu,
                            T lr, T b1, T b2, T eps, T eps_root, T c1, T c2, T gs, const T *__restrict__ gs_dev) {
  if (gs_dev) gs *= *gs_dev;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T gi = g[i] * gs;
    const T m = b1 * mu[i] + (T(1) - b1) * gi;
    const T v = b2 * nu[i] + (T(1) - b2) * gi * gi;
    mu[i] = m;
    nu[i] = v;
    p[i] -= lr * (m * c1) / (Num<T>::sqrt(v * c2 + eps_root) + eps);
  }
}

// site_norms[site] += sum of squares of this block's part of the site's slab (padding is zero, so the slab norm is the
// norm of the reference-shaped gradient tensor).  grid = (blocks per site, L)
template <typename T>
__global__ void site_sumsq_kernel(const T *__restrict__ g, const long long *__restrict__ offs, double *__restrict__ norms,
                                  T pre) {
  const int site = blockIdx.y;
  const long long a = offs[site], b = offs[site + 1];
  double acc = 0.0;
  for (long long i = a + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)(g[i] * pre);
    acc += v * v;
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x + 31) / 32; ++w) acc += red[w];
    atomicAdd(&norms[site], acc);
  }
}

// g_site *= pre * min(1, thr / (||pre * g_site|| + 1e-6))        (util.gradient_clip, one factor per site tensor)
template <typename T>
__global__ void site_clip_kernel(T *__restrict__ g, const long long *__restrict__ offs, const double *__restrict__ norms,
                                 double thr, T pre) {
  const int site = blockIdx.y;
  const long long a = offs[site], b = offs[site + 1];
  const double f = fmin(1.0, thr / (sqrt(norms[site]) + 1e-6));
  const T sc = pre * (T)f;
  for (long long i = a + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += (long long)gridDim.x * blockDim.x)
    g[i] *= sc;
}

__global__ void site_offsets_kernel(long long *offs, int L, long long per, int kind, int n_out, int out_site, int spacing) {
  for (int site = threadIdx.x; site <= L; site += blockDim.x) {
    long long nwith;
    if (kind == TNB_KIND_MPS) nwith = site > out_site ? 1 : 0;
    else nwith = (site + spacing - 1) / spacing;
    offs[site] = per * (site - nwith) + per * n_out * nwith;
  }
}

}  // namespace tnb

using namespace tnb;

extern "C" {

static int adam_impl(int32_t dtype, int64_t n, void *params, const void *grads, void *mu, void *nu, double lr, double b1,
                     double b2, double eps, double eps_root, int64_t count, double grad_scale, const void *gs_dev,
                     void *stream) {
  if (!params || !grads || !mu || !nu) return fail("null device pointer");
  if (n < 1 || count < 1) return fail("n and count must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  const double c1 = 1.0 / (1.0 - pow(b1, (double)count)), c2 = 1.0 / (1.0 - pow(b2, (double)count));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  prof_begin(st);
  if (dtype == TNB_F64)
    adam_kernel<double><<<blocks, 256, 0, st>>>(n, (double *)params, (const double *)grads, (double *)mu, (double *)nu, lr,
                                                 b1, b2, eps, eps_root, c1, c2, grad_scale, (const double *)gs_dev);
  else if (dtype == TNB_F32)
    adam_kernel<float><<<blocks, 256, 0, st>>>(n, (float *)params, (const float *)grads, (float *)mu, (float *)nu,
                                                (float)lr, (float)b1, (float)b2, (float)eps, (float)eps_root, (float)c1,
                                                (float)c2, (float)grad_scale, (const float *)gs_dev);
  else
    return fail("bad dtype");
  TNB_LAUNCH_CHECK("adam_kernel");
  return 0;
}

int tnb_adam_update(int32_t dtype, int64_t n, void *params, const void *grads, void *mu, void *nu, double lr, double b1,
                    double b2, double eps, double eps_root, int64_t count, double grad_scale, void *stream) {
  return adam_impl(dtype, n, params, grads, mu, nu, lr, b1, b2, eps, eps_root, count, grad_scale, nullptr, stream);
}

int tnb_adam_update_dev(int32_t dtype, int64_t n, void *params, const void *grads, void *mu, void *nu, double lr, double b1,
                        double b2, double eps, double eps_root, int64_t count, double grad_scale,
                        const void *grad_scale_dev, void *stream) {
  return adam_impl(dtype, n, params, grads, mu, nu, lr, b1, b2, eps, eps_root, count, grad_scale, grad_scale_dev, stream);
}

int tnb_gradient_clip(const tnb_problem *p, void *grads, double threshold, double pre_scale, void *scratch, void *stream) {
  if (!p || !grads || !scratch) return fail("null pointer");
  if (!(threshold > 0)) return fail("Threshold must be positive.");  // util.py:144
  cudaStream_t st = (cudaStream_t)stream;
  const int L = p->L;
  long long *offs = (long long *)scratch;
  double *norms = (double *)((char *)scratch + align_up((size_t)(L + 1) * sizeof(long long), 256));
  TNB_CUDA(cudaMemsetAsync(norms, 0, (size_t)L * sizeof(double), st));
  prof_begin(st);
  site_offsets_kernel<<<1, 256, 0, st>>>(offs, L, (long long)p->chi * p->chi * p->d, p->kind, p->n_out, p->out_site,
                                         p->spacing > 0 ? p->spacing : 1);
  TNB_LAUNCH_CHECK("site_offsets_kernel");
  const long long per = (long long)p->chi * p->chi * p->d * (p->kind == TNB_KIND_MPS ? 1 : p->n_out);
  int bx = (int)((per + 256 * 8 - 1) / (256 * 8));
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  dim3 grid(bx, L);
  prof_begin(st);
  if (p->dtype == TNB_F64) site_sumsq_kernel<double><<<grid, 256, 0, st>>>((const double *)grads, offs, norms, pre_scale);
  else site_sumsq_kernel<float><<<grid, 256, 0, st>>>((const float *)grads, offs, norms, (float)pre_scale);
  TNB_LAUNCH_CHECK("site_sumsq_kernel");
  prof_begin(st);
  if (p->dtype == TNB_F64) site_clip_kernel<double><<<grid, 256, 0, st>>>((double *)grads, offs, norms, threshold, pre_scale);
  else site_clip_kernel<float><<<grid, 256, 0, st>>>((float *)grads, offs, norms, threshold, (float)pre_scale);
  TNB_LAUNCH_CHECK("site_clip_kernel");
  return 0;
}

int64_t tnb_gradient_clip_scratch_bytes(const tnb_problem *p) {
  if (!p) return -1;
  return (int64_t)(align_up((size_t)(p->L + 1) * sizeof(long long), 256) + (size_t)p->L * sizeof(double));
}

}  // extern "C"
