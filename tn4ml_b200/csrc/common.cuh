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

// Embedding parameters in kernel-argument form (typed copy of tnb_embedding).  n_in = number of raw input values per site
// (Embedding.input_dim; 1 for the scalar feature maps): site i of sample b reads x[(b * L + i) * n_in ...].
template <typename T> struct EmbParams {
  int kind, k, p, degree, include_bias, n_centers, d, n_in, cross;
  T gamma;
  T centers[kMaxPhys];
};

// typed copy of the ABI struct (host side, every family's argument builder)
template <typename T> static inline void fill_emb(const tnb_embedding &e, EmbParams<T> &o, int d) {
  o.kind = e.kind; o.k = e.k; o.p = e.p; o.degree = e.degree; o.include_bias = e.include_bias;
  o.n_centers = e.n_centers; o.d = d; o.gamma = (T)e.gamma;
  o.n_in = e.n_in > 0 ? e.n_in : 1; o.cross = e.cross;
  for (int i = 0; i < kMaxPhys; ++i) o.centers[i] = (T)e.centers[i];
}

// Raw feature map phi(xp[0 .. n_in)) -> out[0..d)
// (tn4ml/embeddings.py:340-351, 407-423, 693-717, 601-602, 478-484, 754-768, 805-820, 857-869, 913-928)
template <typename T>
__device__ __forceinline__ void feature_map(const EmbParams<T> &e, const T *xp, T *out) {
  using N = Num<T>;
  const T x = xp[0];
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
      // [1?] then, degree by degree, the monomials over the n_in inputs in itertools.combinations_with_replacement order
      // (non-decreasing index tuples, lexicographic); without cross terms only the pure powers x_j^dg survive.
      int o = 0;
      if (e.include_bias) out[o++] = T(1);
      if (e.n_in == 1) {
        T pw = x;
        for (int dg = 1; dg <= e.degree; ++dg) {
          out[o++] = pw;
          pw *= x;
        }
      } else if (!e.cross) {
        for (int dg = 1; dg <= e.degree; ++dg)
          for (int j = 0; j < e.n_in; ++j) {
            T pw = xp[j];
            for (int q = 1; q < dg; ++q) pw *= xp[j];
            out[o++] = pw;
          }
      } else {
        for (int dg = 1; dg <= e.degree && dg <= 8; ++dg) {
          int idx[8];
          for (int q = 0; q < dg; ++q) idx[q] = 0;
          while (true) {
            T pr = T(1);
            for (int q = 0; q < dg; ++q) pr *= xp[idx[q]];
            if (o < kMaxPhys) out[o++] = pr;
            int q = dg - 1;  // next non-decreasing tuple
            while (q >= 0 && idx[q] == e.n_in - 1) --q;
            if (q < 0) break;
            const int v = idx[q] + 1;
            for (int r = q; r < dg; ++r) idx[r] = v;
          }
        }
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
    case TNB_EMB_LINCOMP: {
      // [x, 1 - x] or [1, x, 1 - x], divided by its 2-norm
      int o = 0;
      if (e.p == 3) out[o++] = T(1);
      out[o++] = x;
      out[o++] = T(1) - x;
      T n2 = 0;
      for (int j = 0; j < o; ++j) n2 += out[j] * out[j];
      T inv = T(1) / N::sqrt(n2);
      for (int j = 0; j < o; ++j) out[j] *= inv;
      break;
    }
    case TNB_EMB_LEGENDRE: {
      // P_0 = 1, P_1 = x, (i + 1) P_{i+1} = (2 i + 1) x P_i - i P_{i-1}        (tn4ml/scipy/special.py:6-28)
      T pm = T(1), pc = x;
      out[0] = pm;
      if (e.degree >= 1) out[1] = pc;
      for (int i = 1; i < e.degree; ++i) {
        T pn = (T(2 * i + 1) * x * pc - T(i) * pm) / T(i + 1);
        pm = pc; pc = pn;
        out[i + 1] = pn;
      }
      break;
    }
    case TNB_EMB_LAGUERRE: {
      // exp(-x / 2) L_k(x):  L_0 = 1, L_1 = 1 - x, i L_i = (2 i - 1 - x) L_{i-1} - (i - 1) L_{i-2}   (special.py:57-79)
      const T w = N::exp(-x * T(0.5));
      T lm = T(1), lc = T(1) - x;
      out[0] = w * lm;
      if (e.degree >= 1) out[1] = w * lc;
      for (int i = 2; i <= e.degree; ++i) {
        T ln = ((T(2 * i - 1) - x) * lc - T(i - 1) * lm) / T(i);
        lm = lc; lc = ln;
        out[i] = w * ln;
      }
      break;
    }
    case TNB_EMB_HERMITE: {
      // exp(-x^2 / 2) H_k(x):  H_0 = 1, H_1 = 2 x, H_i = 2 x H_{i-1} - 2 (i - 1) H_{i-2}             (special.py:108-130)
      const T w = N::exp(-T(0.5) * x * x);
      T hm = T(1), hc = T(2) * x;
      out[0] = w * hm;
      if (e.degree >= 1) out[1] = w * hc;
      for (int i = 2; i <= e.degree; ++i) {
        T hn = T(2) * x * hc - T(2 * (i - 1)) * hm;
        hm = hc; hc = hn;
        out[i] = w * hn;
      }
      break;
    }
    case TNB_EMB_ARRAYS: {
      // pre-embedded inputs: [1?] + x[0 .. n_in)                                                    (embeddings.py:913-928)
      int o = 0;
      if (e.include_bias) out[o++] = T(1);
      for (int j = 0; j < e.n_in; ++j) out[o++] = xp[j];
      break;
    }
  }
}

// embed(): the product state is normalised (embeddings.py:1374-1387); the global factor
// prod_i ||phi_i||^-1 is applied site by site, which is the same product state.
template <typename T>
__device__ __forceinline__ void embed_site(const EmbParams<T> &e, const T *xp, T *out) {
  feature_map(e, xp, out);
  T n2 = 0;
  for (int s = 0; s < e.d; ++s) n2 += out[s] * out[s];
  T inv = T(1) / Num<T>::sqrt(n2);
  for (int s = 0; s < e.d; ++s) out[s] *= inv;
}

template <typename T>
__device__ __noinline__ void feature_map_generic(const EmbParams<T> &e, const T *xp, T *out, int d) {
  T xin[kMaxPhys], tmp[kMaxPhys];
  for (int j = 0; j < e.n_in; ++j) xin[j] = xp ? xp[j] : T(0.5);  // xp == nullptr: padding row
  feature_map(e, xin, tmp);
  for (int s = 0; s < d; ++s) out[s] = tmp[s];
}

// Same map with the physical dimension as a compile-time constant: every loop unrolls and phi stays in
// registers (the tensor-core chain kernel evaluates it once per site and sample inside its epilogue warps).
// xp == nullptr stands for a padding row (any finite input).
// GENERIC = false: the four headline maps only, no out-of-line call (kernels at their register limit: the register-resident
// SMPO family; its dispatcher sends the other maps to the shared-memory family).
template <int D, typename T, bool GENERIC = true>
__device__ __forceinline__ void embed_site_d(const EmbParams<T> &e, const T *xp, T (&out)[D]) {
  using N = Num<T>;
  const T x = xp ? xp[0] : T(0.5);
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
  } else if (e.kind == TNB_EMB_POLY && e.n_in == 1) {
    T pw = e.include_bias ? T(1) : x;
#pragma unroll
    for (int s = 0; s < D; ++s) {
      out[s] = pw;
      pw *= x;
    }
  } else if (e.kind == TNB_EMB_RBF || !GENERIC) {
#pragma unroll
    for (int j = 0; j < D; ++j) out[j] = N::exp(-e.gamma * (x - e.centers[j]));
  } else {
    // the less common maps: the generic routine out of line (its local arrays stay out of the caller's frame)
    T tmp[D];
    feature_map_generic(e, xp, tmp, D);
#pragma unroll
    for (int s = 0; s < D; ++s) out[s] = tmp[s];
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
