// Shared device helpers: embeddings (K0), loss heads (K3), small utilities.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/tn4ml_b200.h"

namespace tnb {

constexpr int kMaxPhys = TNB_MAX_PHYS;
constexpr int kMaxOut = TNB_MAX_OUT;

template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float pi() { return 3.14159265358979323846f; }
  static __device__ __forceinline__ float sinpi(float x) { return ::sinpif(x); }
  static __device__ __forceinline__ float cospi(float x) { return ::cospif(x); }
  static __device__ __forceinline__ float exp(float x) { return ::expf(x); }
  static __device__ __forceinline__ float log(float x) { return ::logf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return ::sqrtf(x); }
  static __device__ __forceinline__ float abs(float x) { return ::fabsf(x); }
  static __device__ __forceinline__ float ldexp(float x, int e) { return ::ldexpf(x, e); }
  static __device__ __forceinline__ float frexp(float x, int *e) { return ::frexpf(x, e); }
  static __device__ __forceinline__ float max(float a, float b) { return ::fmaxf(a, b); }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double pi() { return 3.14159265358979323846; }
  static __device__ __forceinline__ double sinpi(double x) { return ::sinpi(x); }
  static __device__ __forceinline__ double cospi(double x) { return ::cospi(x); }
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
  static __device__ __forceinline__ double ldexp(double x, int e) { return ::ldexp(x, e); }
  static __device__ __forceinline__ double frexp(double x, int *e) { return ::frexp(x, e); }
  static __device__ __forceinline__ double max(double a, double b) { return ::fmax(a, b); }
};

// Embedding parameters in kernel-argument form (typed copy of tnb_embedding).
template <typename T> struct EmbParams {
  int kind, k, p, degree, include_bias, n_centers, d;
  T gamma;
  T centers[kMaxPhys];
};

// Raw feature map phi(x) -> out[0..d)   (tn4ml/embeddings.py:340-351, 407-423, 693-717, 601-602)
template <typename T>
__device__ __forceinline__ void feature_map(const EmbParams<T> &e, T x, T *out) {
  using N = Num<T>;
  switch (e.kind) {
    case TNB_EMB_TRIG: {
      // 1/sqrt(k) [cos(pi x / 2^i), i=1..k ; sin(pi x / 2^i), i=1..k]
      T inv = T(1) / N::sqrt(T(e.k));
      T scale = T(0.5);
      for (int i = 0; i < e.k; ++i) {
        out[i] = N::cospi(x * scale) * inv;
        out[e.k + i] = N::sinpi(x * scale) * inv;
        scale *= T(0.5);
      }
      break;
    }
    case TNB_EMB_FOURIER: {
      // (1/p) |sum_k exp(2 pi i k y / p)|,  y = (p-1) x - j  ==  |sin(pi y) / (p sin(pi y / p))|
      T pinv = T(1) / T(e.p);
      for (int j = 0; j < e.p; ++j) {
        T y = T(e.p - 1) * x - T(j);
        T den = N::sinpi(y * pinv);
        T num = N::sinpi(y);
        T v;
        if (N::abs(den) < T(1e-6)) {
          // y/p within 3e-7 of an integer: direct sum of the p phasors (limit value 1)
          T re = 0, im = 0;
          for (int kk = 0; kk < e.p; ++kk) {
            re += N::cospi(T(2 * kk) * y * pinv);
            im += N::sinpi(T(2 * kk) * y * pinv);
          }
          v = N::sqrt(re * re + im * im) * pinv;
        } else {
          v = N::abs(num / den) * pinv;
        }
        out[j] = v;
      }
      break;
    }
    case TNB_EMB_POLY: {
      int o = 0;
      if (e.include_bias) out[o++] = T(1);
      T pw = x;
      for (int dg = 1; dg <= e.degree; ++dg) {
        out[o++] = pw;
        pw *= x;
      }
      break;
    }
    case TNB_EMB_RBF: {
      T n2 = 0;
      for (int j = 0; j < e.n_centers; ++j) {
        T v = N::exp(-e.gamma * (x - e.centers[j]));
        out[j] = v;
        n2 += v * v;
      }
      T inv = T(1) / N::sqrt(n2);
      for (int j = 0; j < e.n_centers; ++j) out[j] *= inv;
      break;
    }
  }
}

// embed(): the product state is normalised (embeddings.py:1374-1387); the global factor
// prod_i ||phi_i||^-1 is applied site by site, which is the same product state.
template <typename T>
__device__ __forceinline__ void embed_site(const EmbParams<T> &e, T x, T *out) {
  feature_map(e, x, out);
  T n2 = 0;
  for (int s = 0; s < e.d; ++s) n2 += out[s] * out[s];
  T inv = T(1) / Num<T>::sqrt(n2);
  for (int s = 0; s < e.d; ++s) out[s] *= inv;
}

// Same map with the physical dimension as a compile-time constant: every loop unrolls and phi stays in
// registers (the tensor-core chain kernel evaluates it once per site and sample inside its epilogue warps).
template <int D, typename T>
__device__ __forceinline__ void embed_site_d(const EmbParams<T> &e, T x, T (&out)[D]) {
  using N = Num<T>;
  if (e.kind == TNB_EMB_TRIG) {
    constexpr int K = D / 2 > 0 ? D / 2 : 1;
    const T inv = T(1) / N::sqrt(T(K));
    T scale = T(0.5);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      out[i] = N::cospi(x * scale) * inv;
      if (K + i < D) out[K + i] = N::sinpi(x * scale) * inv;
      scale *= T(0.5);
    }
  } else if (e.kind == TNB_EMB_FOURIER) {
    const T pinv = T(1) / T(D);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const T y = T(D - 1) * x - T(j);
      const T den = N::sinpi(y * pinv);
      const T num = N::sinpi(y);
      T v;
      if (N::abs(den) < T(1e-6)) {
        T re = 0, im = 0;
        for (int kk = 0; kk < D; ++kk) {
          re += N::cospi(T(2 * kk) * y * pinv);
          im += N::sinpi(T(2 * kk) * y * pinv);
        }
        v = N::sqrt(re * re + im * im) * pinv;
      } else {
        v = N::abs(num / den) * pinv;
      }
      out[j] = v;
    }
  } else if (e.kind == TNB_EMB_POLY) {
    T pw = e.include_bias ? T(1) : x;
#pragma unroll
    for (int s = 0; s < D; ++s) {
      out[s] = pw;
      pw *= x;
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; ++j) out[j] = N::exp(-e.gamma * (x - e.centers[j]));
  }
  T n2 = 0;
#pragma unroll
  for (int s = 0; s < D; ++s) n2 += out[s] * out[s];
  const T inv = T(1) / N::sqrt(n2);
#pragma unroll
  for (int s = 0; s < D; ++s) out[s] *= inv;
}

// ---------------------------------------------------------------------------------------
// Classifier / NLL heads on the rescaled output y' = y * 2^-E (E = esum).  Returns the
// per-sample loss; if dy != nullptr writes dLoss/dy' (unscaled by loss_scale).
// All classifier heads act on y/||y|| and are therefore invariant to the 2^-E factor.
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ inline T mps_head(int head, int C, const T *y, int E, const T *t, const T *cw, T *dy) {
  using N = Num<T>;
  if (head == TNB_HEAD_NLL) {
    // -log(s^2), s = y' 2^E   (metrics.py:53)
    T a = N::abs(y[0]);
    T loss = T(-2) * (N::log(a) + T(E) * T(0.69314718055994530942));
    if (dy) dy[0] = T(-2) / y[0];
    return loss;
  }
  T n2 = 0;
  for (int k = 0; k < C; ++k) n2 += y[k] * y[k];
  T nrm = N::sqrt(n2);
  T inv = T(1) / nrm;
  T dyh[kMaxOut];
  T loss = 0;
  if (head == TNB_HEAD_MSE) {
    for (int k = 0; k < C; ++k) {
      T diff = y[k] * inv - t[k];
      loss += diff * diff;
      dyh[k] = T(2) * diff / T(C);
    }
    loss /= T(C);
  } else {
    T mx = y[0] * inv;
    for (int k = 1; k < C; ++k) mx = N::max(mx, y[k] * inv);
    T se = 0;
    for (int k = 0; k < C; ++k) se += N::exp(y[k] * inv - mx);
    T lse = N::log(se) + mx;
    if (head == TNB_HEAD_XENT_SOFTMAX_ARGMAX) {
      int idx = 0;
      for (int k = 1; k < C; ++k)
        if (t[k] > t[idx]) idx = k;
      loss = lse - y[idx] * inv;
      for (int k = 0; k < C; ++k) dyh[k] = N::exp(y[k] * inv - lse) - (k == idx ? T(1) : T(0));
    } else {
      T tsum = 0, w = 1;
      for (int k = 0; k < C; ++k) {
        tsum += t[k];
        loss -= t[k] * (y[k] * inv - lse);
      }
      if (head == TNB_HEAD_SOFTMAX_XENT_WEIGHTED) {
        w = 0;
        for (int k = 0; k < C; ++k) w += t[k] * cw[k];
      }
      loss *= w;
      for (int k = 0; k < C; ++k) dyh[k] = w * (tsum * N::exp(y[k] * inv - lse) - t[k]);
    }
  }
  if (dy) {
    T dot = 0;
    for (int k = 0; k < C; ++k) dot += dyh[k] * y[k] * inv;
    for (int k = 0; k < C; ++k) dy[k] = (dyh[k] - y[k] * inv * dot) * inv;
  }
  return loss;
}

// SMPO heads on n = nm * 2^E  (metrics.py:81, :187, :206).  Returns loss, writes dLoss/dnm.
template <typename T>
__device__ inline T smpo_head(int head, T nm, int E, T *dnm) {
  using N = Num<T>;
  const T ln2 = T(0.69314718055994530942);
  if (head == TNB_HEAD_LOGQUAD) {
    T u = T(2) * (N::log(nm) + T(E) * ln2) - T(1);
    if (dnm) *dnm = T(4) * u / nm;
    return u * u;
  }
  T n = N::ldexp(nm, E);
  T sc = N::ldexp(T(1), E);
  if (head == TNB_HEAD_TSN) {
    if (dnm) *dnm = T(2) * n * sc;
    return n * n;
  }
  // TNB_HEAD_QUAD
  T u = n * n - T(1);
  if (dnm) *dnm = T(4) * u * n * sc;
  return u * u;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace tnb
