// Model-only operations of the training loop, on the device:
//   <A|A> by a transfer-matrix sweep                       tn4ml/models/tn.py:66-80, smpo.py:234-248
//   value and gradient of the Frobenius-norm regularisers  tn4ml/metrics.py:89-168 (incl. the SMPO form, where the reference's
//                                                          "norm" is Tr(P^dag P), metrics.py:99-103 + smpo.py:406-438)
// in place of L eager einsums + autograd per training step.
//
//   left  environments  E_0 = e0 e0^T,  E_{i+1}[r,r'] = sum_{a,a',j} E_i[a,a'] A_i[a,r,j] A_i[a',r',j]     (j = (s, o))
//   right environments  R_L = e0 e0^T,  R_i[a,a']     = sum_{r,r',j} A_i[a,r,j] R_{i+1}[r,r'] A_i[a',r',j]
//   N2 = <A|A> = E_L[0,0];   d log N2 / dA_i[a,r,j] = (2 / N2) * (E_i A_i^j R_{i+1})[a,r]
// Every environment is carried as (matrix, binary exponent): the matrix is rescaled by the exact power of two of its
// max-abs whenever it is used, so no chain length leaves the floating-point range (the reference's L > 200 models do).
//
// One PERSISTENT cooperative kernel runs both sweeps (blockIdx.y = direction); a site is two tiled GEMM phases
// (Y = E A, then E' = sum_j A^jT Y^j) separated by a barrier among the CTAs of the direction.  A second kernel forms the
// gradients of all sites at once from the stored Y_i and R_{i+1}.
#include <cmath>

#include "common.cuh"
#include "host.h"

namespace tnb {

constexpr int kNT = 64;   // output tile edge
constexpr int kNK = 16;   // k-slab
constexpr int kNThreads = 256;

struct NormArgs {
  int L, chi, d, kind, n_out, out_site, spacing;
  int ntile;              // CTAs per direction
  const void *params;
  void *Y;                // left sweep: Y_i = (2^-e E_i) A_i for every site, packed like the parameters
  void *env;              // [2 directions][L + 1][chi * chi] environments as computed (not yet rescaled)
  float *emax;            // [2][L + 1] max-abs of each environment (float bits, non-negative: atomicMax on int)
  int *cexp;              // [2][L + 1] cumulative exponent: true env = stored env * 2^cexp
  unsigned *bar;          // [2] barrier counters (zeroed before launch)
  double *log2_n2;        // out: log2 <A|A>
};

__host__ __device__ inline int site_nj(const NormArgs &A, int site) {
  if (A.kind == TNB_KIND_MPS) return site == A.out_site ? A.n_out : 1;
  return site % A.spacing == 0 ? A.n_out : 1;
}
__host__ __device__ inline long long site_off(const NormArgs &A, int site) {
  const long long per = (long long)A.chi * A.chi * A.d;
  const long long nwith = A.kind == TNB_KIND_MPS ? (site > A.out_site ? 1 : 0) : (site + A.spacing - 1) / A.spacing;
  return per * (site - nwith) + per * A.n_out * nwith;
}

// C[m, n] (rs_c, cs_c) (+)= alpha * sum_t sum_k A_t[m, k] (rs_a, cs_a) * B_t[k, n] (rs_b, cs_b), A_t = A + t * bs_a, for the
// 64 x 64 tile at (m0, n0).  256 threads, 4 x 4 register tiles; returns the tile's max-abs of C through *tmax (thread 0).
template <typename T>
__device__ void tile_gemm(T *C, long long rs_c, long long cs_c, const T *A, long long rs_a, long long cs_a, const T *B,
                          long long rs_b, long long cs_b, int M, int N, int K, int nsum, long long bs_a, long long bs_b,
                          T alpha, int m0, int n0, T *As, T *Bs, float *tmax) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int t = 0; t < nsum; ++t) {
    const T *At = A + (long long)t * bs_a, *Bt = B + (long long)t * bs_b;
    for (int k0 = 0; k0 < K; k0 += kNK) {
      // stage A[m0 .. m0+63][k0 .. k0+15] as As[k][m] and B[k0 ..][n0 ..] as Bs[k][n]
      for (int e = tid; e < kNT * kNK; e += kNThreads) {
        const int mm = e % kNT, kk = e / kNT;
        const int m = m0 + mm, k = k0 + kk;
        As[kk * (kNT + 4) + mm] = (m < M && k < K) ? At[(long long)m * rs_a + (long long)k * cs_a] : T(0);
        const int n = n0 + mm;
        Bs[kk * (kNT + 4) + mm] = (n < N && k < K) ? Bt[(long long)k * rs_b + (long long)n * cs_b] : T(0);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kNK; ++kk) {
        T a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk * (kNT + 4) + ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk * (kNT + 4) + tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        const T v = alpha * acc[i][j];
        C[(long long)m * rs_c + (long long)n * cs_c] = v;
        mx = fmaxf(mx, fabsf((float)v));
      }
    }
  }
  if (tmax) {
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    __shared__ float red[kNThreads / 32];
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kNThreads / 32; ++w) mx = fmaxf(mx, red[w]);
      atomicMax(reinterpret_cast<int *>(tmax), __float_as_int(mx));
    }
    __syncthreads();
  }
}

// barrier among the `n` CTAs of one direction (co-resident: cooperative launch); counter grows by n per barrier
__device__ __forceinline__ void dir_barrier(unsigned *ctr, unsigned n, unsigned &gen) {
  __syncthreads();
  if (n > 1) {
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      const unsigned target = (gen + 1) * n;
      while (*((volatile unsigned *)ctr) < target) { __nanosleep(64); }
      __threadfence();
    }
    __syncthreads();
  }
  ++gen;
}

// exponent e with max * 2^-e in [0.5, 1)  (0 for a zero matrix)
__device__ __forceinline__ int exp_of(float mx) {
  int e = 0;
  if (mx > 0.f) frexpf(mx, &e);
  return e;
}

template <typename T> __global__ void __launch_bounds__(kNThreads) tn_env_sweep_kernel(NormArgs A) {
  __shared__ T As[kNK * (kNT + 4)], Bs[kNK * (kNT + 4)];
  const int dir = blockIdx.y, L = A.L, chi = A.chi;
  const unsigned ntile = gridDim.x;
  unsigned gen = 0;
  unsigned *bar = A.bar + dir;
  T *env = (T *)A.env + (size_t)dir * (L + 1) * chi * chi;
  float *emax = A.emax + dir * (L + 1);
  int *cexp = A.cexp + dir * (L + 1);
  T *Ybase = (T *)A.Y;                                  // left sweep keeps every Y_i
  T *Yscr = Ybase;                                      // right sweep: one scratch slab behind the packed Y
  const int tmn = (chi + kNT - 1) / kNT;
  // boundary environment e0 e0^T
  const int b0 = dir == 0 ? 0 : L;
  for (int i = blockIdx.x * kNThreads + threadIdx.x; i < chi * chi; i += ntile * kNThreads) env[(size_t)b0 * chi * chi + i] = T(i == 0);
  if (blockIdx.x == 0 && threadIdx.x == 0) { emax[b0] = 1.f; cexp[b0] = 0; }
  dir_barrier(bar, ntile, gen);
  for (int step = 0; step < L; ++step) {
    const int site = dir == 0 ? step : L - 1 - step;
    const int bin = dir == 0 ? site : site + 1, bout = dir == 0 ? site + 1 : site;
    const int nj = site_nj(A, site), J = A.d * nj, NY = chi * J;
    const T *W = (const T *)A.params + site_off(A, site);
    const T *Ein = env + (size_t)bin * chi * chi;
    T *Eout = env + (size_t)bout * chi * chi;
    const int e = exp_of(emax[bin]);
    const T alpha = Num<T>::ldexp(T(1), -e);
    if (blockIdx.x == 0 && threadIdx.x == 0) { cexp[bout] = cexp[bin] + e; emax[bout] = 0.f; }
    T *Y = dir == 0 ? Ybase + site_off(A, site) : Yscr + site_off(A, L);
    // phase 1.  left:  Y[a', (r, j)] = sum_a E[a', a] W[a, (r, j)]            (plain row-major GEMM, N = chi * J)
    //           right: Y[(a, j), r'] laid out [a][r'][j]: Y[a, r', j] = sum_r W[a, r, j] R[r, r']   (one GEMM per j)
    if (dir == 0) {
      const int tn = (NY + kNT - 1) / kNT;
      for (int t = blockIdx.x; t < tmn * tn; t += ntile)
        tile_gemm<T>(Y, NY, 1, Ein, chi, 1, W, NY, 1, chi, NY, chi, 1, 0, 0, alpha, (t / tn) * kNT, (t % tn) * kNT, As, Bs, nullptr);
    } else {
      for (int t = blockIdx.x; t < tmn * tmn * J; t += ntile) {
        const int j = t / (tmn * tmn), tt = t % (tmn * tmn);
        tile_gemm<T>(Y + j, NY, J, W + j, NY, J, Ein, chi, 1, chi, chi, chi, 1, 0, 0, alpha, (tt / tmn) * kNT, (tt % tmn) * kNT, As,
                     Bs, nullptr);
      }
    }
    dir_barrier(bar, ntile, gen);
    // phase 2.  left:  E'[r, r'] = sum_j sum_a W[a, r, j] Y[a, r', j]     (A operand W^jT: rs = J, cs = NY; sum over j)
    //           right: R'[a, a'] = sum_j sum_r' Y[a, r', j] W[a', r', j]  (B operand W^jT)
    for (int t = blockIdx.x; t < tmn * tmn; t += ntile) {
      const int m0 = (t / tmn) * kNT, n0 = (t % tmn) * kNT;
      if (dir == 0)
        tile_gemm<T>(Eout, chi, 1, W, J, NY, Y, NY, J, chi, chi, chi, J, 1, 1, T(1), m0, n0, As, Bs, &emax[bout]);
      else
        tile_gemm<T>(Eout, chi, 1, Y, NY, J, W, J, NY, chi, chi, chi, J, 1, 1, T(1), m0, n0, As, Bs, &emax[bout]);
    }
    dir_barrier(bar, ntile, gen);
  }
  if (dir == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    // N2 = E_L[0,0] * 2^cexp[L]
    const double v = (double)env[(size_t)L * chi * chi];
    *A.log2_n2 = log2(fabs(v)) + (double)cexp[L];
  }
}

struct NormGradArgs {
  NormArgs N;
  void *grads;
  int nreg;
  int reg_kind[8];
  double reg_coef[8];
  double *reg_value;  // += sum_r coef_r * reg_r  (written by block (0,0,0))
};

// value and d/d(ln N2) of the regularisers; l = ln(norm), norm = sqrt(N2), or N2 itself for the *_TRACE kinds (the reference's
// SpacedMatrixProductOperator branch: metrics.py:99-103)
__device__ inline void reg_terms(const NormGradArgs &G, double ln_n2, double *value, double *dvalue_dlnn2) {
  double v = 0.0, dv = 0.0;
  for (int r = 0; r < G.nreg; ++r) {
    const double c = G.reg_coef[r];
    const double kappa = G.reg_kind[r] >= 4 ? 1.0 : 0.5;
    const double l = kappa * ln_n2;
    switch (G.reg_kind[r] & 3) {
      case TNB_REG_LOG_FROB: v += c * l; dv += c * kappa; break;
      case TNB_REG_LOG_POW_FROB: v += c * 2.0 * l; dv += c * 2.0 * kappa; break;
      case TNB_REG_LOG_RELU_FROB: v += c * fmax(l, 0.0); dv += l > 0.0 ? c * kappa : 0.0; break;
      case TNB_REG_QUAD_FROB: v += c * (l - 1.0) * (l - 1.0); dv += c * 2.0 * (l - 1.0) * kappa; break;
    }
  }
  *value = v;
  *dvalue_dlnn2 = dv;
}

// grads_i[a, r, j] += f * (2 / N2) * sum_{r'} Y_i[a, r', j] R_{i+1}[r', r] with every exponent applied once at the end.
// grid = (tiles, J_max, L)
template <typename T> __global__ void __launch_bounds__(kNThreads) tn_norm_grad_kernel(NormGradArgs G) {
  __shared__ T As[kNK * (kNT + 4)], Bs[kNK * (kNT + 4)];
  const NormArgs &A = G.N;
  const int site = blockIdx.z, j = blockIdx.y, chi = A.chi, L = A.L;
  const int nj = site_nj(A, site), J = A.d * nj, NY = chi * J;
  const double ln_n2 = *A.log2_n2 * 0.69314718055994530942;
  double val, dval;
  reg_terms(G, ln_n2, &val, &dval);
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && G.reg_value) *G.reg_value += val;
  if (j >= J || !G.grads) return;
  const int tmn = (chi + kNT - 1) / kNT;
  const T *Y = (const T *)A.Y + site_off(A, site);
  const T *R = (const T *)A.env + (size_t)(L + 1) * chi * chi + (size_t)(site + 1) * chi * chi;
  const int *cl = A.cexp, *cr = A.cexp + (L + 1);
  const float *ml = A.emax, *mr = A.emax + (L + 1);
  // Y_i carries E_i * 2^-(e_i) of the stored E_i (true E_i = stored * 2^cl[i]); R stored, true = stored * 2^cr[i+1]
  const int ex = cl[site] + exp_of(ml[site]) + cr[site + 1];
  const double scale = 2.0 * dval * exp2((double)ex - *A.log2_n2);
  T *g = (T *)G.grads + site_off(A, site);
  (void)mr;
  // C = g (accumulate): computed into registers by tile_gemm with alpha, then added -- reuse tile_gemm on a temp view is
  // not possible in place, so the accumulate is done here: one tile per CTA.x
  for (int t = blockIdx.x; t < tmn * tmn; t += gridDim.x) {
    const int m0 = (t / tmn) * kNT, n0 = (t % tmn) * kNT;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][q] = T(0);
    for (int k0 = 0; k0 < chi; k0 += kNK) {
      for (int e = tid; e < kNT * kNK; e += kNThreads) {
        const int mm = e % kNT, kk = e / kNT;
        const int m = m0 + mm, k = k0 + kk, n = n0 + mm;
        As[kk * (kNT + 4) + mm] = (m < chi && k < chi) ? Y[(long long)m * NY + (long long)k * J + j] : T(0);
        Bs[kk * (kNT + 4) + mm] = (n < chi && k < chi) ? R[(long long)k * chi + n] : T(0);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kNK; ++kk) {
        T a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk * (kNT + 4) + ty * 4 + i];
#pragma unroll
        for (int q = 0; q < 4; ++q) b[q] = Bs[kk * (kNT + 4) + tx * 4 + q];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] = fma(a[i], b[q], acc[i][q]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n = n0 + tx * 4 + q;
        if (m < chi && n < chi) g[(long long)m * NY + (long long)n * J + j] += (T)(scale * (double)acc[i][q]);
      }
    }
  }
}

// params *= exp(mult * ln2 * log2_value)   (normalize(): divide every site by norm^(1/L), tn.py:103-107)
template <typename T> __global__ void tn_scale_kernel(T *p, long long n, const double *log2_value, double mult) {
  const T f = (T)exp2(mult * *log2_value);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] *= f;
}

static void layout(const tnb_problem *p, size_t *off_Y, size_t *off_env, size_t *off_emax, size_t *off_cexp, size_t *off_bar,
                   size_t *total) {
  const size_t w = esize(p);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  const size_t slab_max = (size_t)p->chi * p->chi * p->d * p->n_out;
  *off_Y = take(((size_t)tnb_param_count(p) + slab_max) * w);  // every Y_i + one scratch slab (right sweep)
  *off_env = take((size_t)2 * (p->L + 1) * p->chi * p->chi * w);
  *off_emax = take((size_t)2 * (p->L + 1) * sizeof(float));
  *off_cexp = take((size_t)2 * (p->L + 1) * sizeof(int));
  *off_bar = take(2 * sizeof(unsigned));
  *total = o;
}

}  // namespace tnb

using namespace tnb;

extern "C" {

int64_t tnb_norm_workspace_bytes(const tnb_problem *p) {
  if (!p || p->L < 1 || p->chi < 1) { fail("bad problem"); return -1; }
  size_t a, b, c, d, e, total;
  layout(p, &a, &b, &c, &d, &e, &total);
  return (int64_t)total;
}

int tnb_norm(const tnb_problem *p, const void *params, double *log2_norm2, int32_t n_regs, const int32_t *reg_kinds,
             const double *reg_coefs, void *grads, double *reg_value, void *workspace, int64_t workspace_bytes, void *stream) {
  if (!p || !params || !log2_norm2 || !workspace) return fail("null pointer");
  if (n_regs < 0 || n_regs > 8) return fail("at most 8 regularisers");
  if (n_regs > 0 && (!reg_kinds || !reg_coefs)) return fail("null regulariser tables");
  size_t oY, oE, oM, oC, oB, total;
  layout(p, &oY, &oE, &oM, &oC, &oB, &total);
  if ((int64_t)total > workspace_bytes) return fail("workspace too small: need %zu bytes, got %lld", total, (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  char *w = (char *)workspace;
  NormArgs A;
  A.L = p->L; A.chi = p->chi; A.d = p->d; A.kind = p->kind; A.n_out = p->n_out; A.out_site = p->out_site;
  A.spacing = p->spacing > 0 ? p->spacing : 1;
  A.params = params; A.Y = w + oY; A.env = w + oE; A.emax = (float *)(w + oM); A.cexp = (int *)(w + oC);
  A.bar = (unsigned *)(w + oB); A.log2_n2 = log2_norm2;
  // CTAs per direction: enough tiles for the widest phase, capped so that 2 * ntile CTAs are co-resident
  const int tmn = (p->chi + kNT - 1) / kNT;
  int want = tmn * ((p->chi * p->d + kNT - 1) / kNT);
  int dev = 0, sms = 148, per_sm = 1;
  TNB_CUDA(cudaGetDevice(&dev));
  TNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (p->dtype == TNB_F64) TNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tn_env_sweep_kernel<double>, kNThreads, 0));
  else TNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tn_env_sweep_kernel<float>, kNThreads, 0));
  int cap = sms * (per_sm > 0 ? per_sm : 1) / 2;
  if (cap > 64) cap = 64;
  A.ntile = want < 1 ? 1 : (want > cap ? cap : want);
  TNB_CUDA(cudaMemsetAsync(A.bar, 0, 2 * sizeof(unsigned), st));
  void *kargs[] = {&A};
  prof_begin(st);
  cudaError_t e;
  dim3 grid(A.ntile, 2);
  if (p->dtype == TNB_F64) e = cudaLaunchCooperativeKernel((void *)tn_env_sweep_kernel<double>, grid, dim3(kNThreads), kargs, 0, st);
  else e = cudaLaunchCooperativeKernel((void *)tn_env_sweep_kernel<float>, grid, dim3(kNThreads), kargs, 0, st);
  if (e != cudaSuccess) return fail("cooperative launch of tn_env_sweep_kernel: %s", cudaGetErrorString(e));
  TNB_LAUNCH_CHECK("tn_env_sweep_kernel");
  if (n_regs > 0 && (grads || reg_value)) {
    NormGradArgs G;
    G.N = A; G.grads = grads; G.nreg = n_regs; G.reg_value = reg_value;
    for (int r = 0; r < n_regs; ++r) { G.reg_kind[r] = reg_kinds[r]; G.reg_coef[r] = reg_coefs[r]; }
    int gx = tmn * tmn;
    if (gx > 64) gx = 64;
    dim3 gg(gx, p->d * p->n_out, p->L);
    prof_begin(st);
    if (p->dtype == TNB_F64) tn_norm_grad_kernel<double><<<gg, kNThreads, 0, st>>>(G);
    else tn_norm_grad_kernel<float><<<gg, kNThreads, 0, st>>>(G);
    TNB_LAUNCH_CHECK("tn_norm_grad_kernel");
  }
  return 0;
}

int tnb_scale(int32_t dtype, int64_t n, void *params, const double *log2_value, double mult, void *stream) {
  if (!params || !log2_value || n < 1) return fail("bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  prof_begin(st);
  if (dtype == TNB_F64) tn_scale_kernel<double><<<blocks, 256, 0, st>>>((double *)params, n, log2_value, mult);
  else if (dtype == TNB_F32) tn_scale_kernel<float><<<blocks, 256, 0, st>>>((float *)params, n, log2_value, mult);
  else return fail("bad dtype");
  TNB_LAUNCH_CHECK("tn_scale_kernel");
  return 0;
}

}  // extern "C"
