// K5 (SIMT family): SMPO / MPO double-layer environments, forward and VJP.
//
//   v = P|phi>  is an MPS over the output legs (tn4ml/models/smpo.py:269-404); n = <v|v> is
//   evaluated window by window (one window = an output site and the leg-less sites after it):
//       P_w   = M_{w0+1} ... M_{w1-1}                (per-sample chi x chi products)
//       T^o   = M^o_{w0} P_w
//       E_{w+1} = 2^-delta_w * sum_o T^o^T E_w T^o   (delta_w: exact power-of-two rescale)
//   which is the cheapest association (cfg 3).  The VJP recomputes each window
//   from its stored entry environment E_w (one chi x chi stash per window and sample).
//
// Thread geometry: 4x4 register micro-tile per thread, (chip/4)^2 threads per sample,
// SPC samples per CTA; every sample-local matrix lives in shared memory (row stride ld).
#pragma once
#include "common.cuh"
#include "consts.h"
#include "mps_simt.cuh"  // lds_vec

namespace tnb {

template <typename T> struct SmpoArgs {
  int L, chi, chip, ld, D, DO, S, nwin, head;
  long long B;
  EmbParams<T> emb;
  const T *x;
  const T *params;
  T *Estash;  // [nwin][B][chip*chip]
  int *dexp;  // [nwin][B]
  int *esum;  // [B]
  T *nm;      // [B] mantissa of n
  T *Qscr;    // [grid][SPC][S-1][chip*chip]
  T *loss;
  double *loss_sum;
  T *out;
  int *out_exp;
  T *grads;
  T loss_scale;
  const T *sw;  // [B] per-sample weight of the loss in the VJP (may be null)
  int stash, SPC, TPS;
};

struct SmpoPlan {
  int chip, ld, TPS, SPC, nwin, grid, NT;
  size_t off_E, off_dexp, off_esum, off_nm, off_Q, total, smem_fwd, smem_bwd;
  int launches_fwd, launches_bwd;
};

__host__ __device__ inline long long smpo_site_off(int chi, int D, int DO, int S, int site) {
  long long per = (long long)chi * chi * D;
  long long nwith = (site + S - 1) / S;
  return per * (site - nwith) + per * DO * nwith;
}

template <typename T> __device__ __forceinline__ void sts_vec4(T *p, const T (&w)[4]);
template <> __device__ __forceinline__ void sts_vec4<float>(float *p, const float (&w)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ void sts_vec4<double>(double *p, const double (&w)[4]) {
  *reinterpret_cast<double2 *>(p) = make_double2(w[0], w[1]);
  *reinterpret_cast<double2 *>(p + 2) = make_double2(w[2], w[3]);
}

// C[i0..i0+3][j0..j0+3] (+)= alpha * op(A) op(B), all chip x chip in shared memory with row stride ld (rows 16-byte
// aligned).  Four k at a time: every operand is fetched with 16-byte loads along its contiguous direction
// (8 vector loads per 64 FMAs instead of 32 scalar ones -- the scalar version was bound by the LSU, not the FMA pipe).
// float64 on the FP64 tensor pipe: the same product with `mma.sync.m8n8k4.f64` warp tiles.  A sample's (chip/4)^2 threads
// are whole warps for chip = 32 (2 warps) and 64 (8 warps); every warp owns 8 of the 8x8 output tiles (chip 32: two tile
// rows x four columns, chip 64: one tile row x eight columns).  Per k-step of 4 a lane loads one element per A / B tile
// (RPW + NTI loads for 8 MMAs = 2048 FMAs of the warp); the FMA version issued 64 FMAs and 8 vector loads per thread for
// the same k-step, i.e. 8x the math instructions.
template <int CHIP, bool TA, bool TB>
__device__ __forceinline__ void mm_dmma(double *Cs, const double *As, const double *Bs, int ld, int lt, double alpha,
                                        bool accum) {
  constexpr int NTI = CHIP / 8, RPW = 8 / NTI;
  const int wi = lt >> 5, lane = lt & 31, g = lane >> 2, t = lane & 3;
  double c[RPW][NTI][2];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int cn = 0; cn < NTI; ++cn) c[r][cn][0] = c[r][cn][1] = 0.0;
#pragma unroll 2
  for (int k0 = 0; k0 < CHIP; k0 += 4) {
    const int k = k0 + t;
    double a[RPW], b[NTI];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = (wi * RPW + r) * 8 + g;
      a[r] = TA ? As[k * ld + i] : As[i * ld + k];
    }
#pragma unroll
    for (int cn = 0; cn < NTI; ++cn) {
      const int j = cn * 8 + g;
      b[cn] = TB ? Bs[j * ld + k] : Bs[k * ld + j];
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int cn = 0; cn < NTI; ++cn)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                     : "+d"(c[r][cn][0]), "+d"(c[r][cn][1])
                     : "d"(a[r]), "d"(b[cn]));
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int i = (wi * RPW + r) * 8 + g;
#pragma unroll
    for (int cn = 0; cn < NTI; ++cn) {
      double2 *dst = reinterpret_cast<double2 *>(&Cs[i * ld + cn * 8 + 2 * t]);
      double2 o;
      if (accum) {
        const double2 old = *dst;
        o.x = fma(alpha, c[r][cn][0], old.x);
        o.y = fma(alpha, c[r][cn][1], old.y);
      } else {
        o.x = alpha * c[r][cn][0];
        o.y = alpha * c[r][cn][1];
      }
      *dst = o;
    }
  }
}

// float32 on the tensor cores with fp32 accuracy: 3xTF32 (x = big + small, both TF32; big*big + big*small + small*big
// accumulated in fp32) on `mma.sync.m16n8k8.tf32` warp tiles.  A warp owns one 16-row tile and four 8-column tiles of the
// sample's output (chip 32: 2 warps per sample, chip 64: 8); 12 MMAs per k-step of 8 replace 128 FMAs per thread.
__device__ __forceinline__ void tf32_split(float x, uint32_t &big, uint32_t &small) {
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(big) : "f"(x));
  const float rest = x - __uint_as_float(big);
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(small) : "f"(rest));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CHIP, bool TA, bool TB>
__device__ __forceinline__ void mm_tf32x3(float *Cs, const float *As, const float *Bs, int ld, int lt, float alpha,
                                          bool accum) {
  constexpr int NW = CHIP / 32;  // warps per 16-row tile
  const int wi = lt >> 5, lane = lt & 31, g = lane >> 2, t = lane & 3;
  const int i_base = (wi / NW) * 16, j_base = (wi % NW) * 32;
  float c[4][4];
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int q = 0; q < 4; ++q) c[n][q] = 0.f;
#pragma unroll 2
  for (int k0 = 0; k0 < CHIP; k0 += 8) {
    uint32_t ab[4], as_[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i_base + g + (q & 1) * 8, k = k0 + t + (q >> 1) * 4;
      tf32_split(TA ? As[k * ld + i] : As[i * ld + k], ab[q], as_[q]);
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      uint32_t bb[2], bs[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int k = k0 + t + q * 4, j = j_base + n * 8 + g;
        tf32_split(TB ? Bs[j * ld + k] : Bs[k * ld + j], bb[q], bs[q]);
      }
      mma_tf32(c[n], as_, bb);
      mma_tf32(c[n], ab, bs);
      mma_tf32(c[n], ab, bb);
    }
  }
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float2 *dst = reinterpret_cast<float2 *>(&Cs[(i_base + g + h * 8) * ld + j_base + n * 8 + 2 * t]);
      float2 o;
      if (accum) {
        const float2 old = *dst;
        o.x = fmaf(alpha, c[n][2 * h], old.x);
        o.y = fmaf(alpha, c[n][2 * h + 1], old.y);
      } else {
        o.x = alpha * c[n][2 * h];
        o.y = alpha * c[n][2 * h + 1];
      }
      *dst = o;
    }
}

template <typename T, bool TA, bool TB> struct MmTensor {
  static __device__ __forceinline__ bool run(T *, const T *, const T *, int, int, int, int, T, bool) { return false; }
};
template <bool TA, bool TB> struct MmTensor<float, TA, TB> {
  static __device__ __forceinline__ bool run(float *Cs, const float *As, const float *Bs, int chip, int ld, int i0, int j0,
                                             float alpha, bool accum) {
    const int lt = (i0 >> 2) * (chip >> 2) + (j0 >> 2);  // thread index inside the sample
    if (chip == 32) { mm_tf32x3<32, TA, TB>(Cs, As, Bs, ld, lt, alpha, accum); return true; }
    if (chip == 64) { mm_tf32x3<64, TA, TB>(Cs, As, Bs, ld, lt, alpha, accum); return true; }
    return false;
  }
};
template <bool TA, bool TB> struct MmTensor<double, TA, TB> {
  static __device__ __forceinline__ bool run(double *Cs, const double *As, const double *Bs, int chip, int ld, int i0,
                                             int j0, double alpha, bool accum) {
    const int lt = (i0 >> 2) * (chip >> 2) + (j0 >> 2);  // thread index inside the sample
    if (chip == 32) { mm_dmma<32, TA, TB>(Cs, As, Bs, ld, lt, alpha, accum); return true; }
    if (chip == 64) { mm_dmma<64, TA, TB>(Cs, As, Bs, ld, lt, alpha, accum); return true; }
    return false;
  }
};

template <typename T, bool TA, bool TB>
__device__ __forceinline__ void mm(T *Cs, const T *As, const T *Bs, int chip, int ld, int i0, int j0, T alpha,
                                   bool accum) {
  if (MmTensor<T, TA, TB>::run(Cs, As, Bs, chip, ld, i0, j0, alpha, accum)) return;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int k0 = 0; k0 < chip; k0 += 4) {
    T a[4][4], b[4][4];  // a[kk][i], b[kk][j]
    if (TA) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) lds_vec<T, 4>(&As[(k0 + kk) * ld + i0], a[kk]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        T r[4];
        lds_vec<T, 4>(&As[(i0 + i) * ld + k0], r);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) a[kk][i] = r[kk];
      }
    }
    if (TB) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        T r[4];
        lds_vec<T, 4>(&Bs[(j0 + j) * ld + k0], r);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) b[kk][j] = r[kk];
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) lds_vec<T, 4>(&Bs[(k0 + kk) * ld + j0], b[kk]);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[kk][i], b[kk][j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    T *dst = &Cs[(i0 + i) * ld + j0];
    T o[4];
    if (accum) {
      T old[4];
      lds_vec<T, 4>(dst, old);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fma(alpha, acc[i][j], old[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = alpha * acc[i][j];
    }
    sts_vec4<T>(dst, o);
  }
}

template <typename T> __device__ __forceinline__ void ldg_pair(const T *p, T &a, T &b);
template <> __device__ __forceinline__ void ldg_pair<float>(const float *p, float &a, float &b) {
  const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
  a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void ldg_pair<double>(const double *p, double &a, double &b) {
  const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
  a = v.x; b = v.y;
}

// M[a][r] = sum_s phi[s] * A_site[a][r][s][o]   (this thread's 4x4 block)
template <typename T>
__device__ __forceinline__ void form_m(const SmpoArgs<T> &A, int site, int o, const T *phi, T *Ms, int i0, int j0) {
  const int nj = (site % A.S == 0) ? A.DO : 1;
  const T *P = A.params + smpo_site_off(A.chi, A.D, A.DO, A.S, site);
  // d = 2, full 4x4 block inside the bond: the block's weights are 4 runs of 8 (leg-less site) or 16 (output site, both
  // output indices interleaved) consecutive values; all of them are requested before the first is used.  (The scalar
  // loop below -- a dependent load per (a, r, s) behind bounds checks -- was 46 % of the backward kernel's stall samples.)
  if (A.D == 2 && nj <= 2 && i0 + 4 <= A.chi && j0 + 4 <= A.chi) {
    T w[4][4][2];  // [i][j][s]
    if (nj == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const T *q = P + ((size_t)(i0 + i) * A.chi + j0) * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) ldg_pair<T>(q + 2 * j, w[i][j][0], w[i][j][1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const T *q = P + ((size_t)(i0 + i) * A.chi + j0) * 4 + o;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w[i][j][0] = __ldg(q + 4 * j);
          w[i][j][1] = __ldg(q + 4 * j + 2);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ms[(i0 + i) * A.ld + j0 + j] = fma(phi[1], w[i][j][1], phi[0] * w[i][j][0]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int a = i0 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = j0 + j;
      T v = T(0);
      if (a < A.chi && r < A.chi) {
        const T *q = P + ((size_t)(a * A.chi + r) * A.D) * nj + o;
        for (int s = 0; s < A.D; ++s) v = fma(phi[s], __ldg(q + (size_t)s * nj), v);
      }
      Ms[a * A.ld + r] = v;
    }
  }
}

template <typename T>
__device__ __forceinline__ void blk_copy_to_global(T *g, const T *s, int chip, int ld, int i0, int j0) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) g[(size_t)(i0 + i) * chip + j0 + j] = s[(i0 + i) * ld + j0 + j];
}
template <typename T>
__device__ __forceinline__ void blk_copy_from_global(T *s, const T *g, int chip, int ld, int i0, int j0) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[(i0 + i) * ld + j0 + j] = g[(size_t)(i0 + i) * chip + j0 + j];
}
template <typename T> __device__ __forceinline__ void blk_zero(T *s, int ld, int i0, int j0) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[(i0 + i) * ld + j0 + j] = T(0);
}

// phi for every site of window [w0, w1) and every sample of the CTA -> phi_s[sl][k][s]
template <typename T>
__device__ __forceinline__ void window_phi(const SmpoArgs<T> &A, long long b0, int w0, int w1, T *phi_s) {
  const int n = (w1 - w0) * A.SPC;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int sl = i / (w1 - w0), k = i - sl * (w1 - w0);
    long long b = b0 + sl;
    T ph[kMaxPhys];
    if (b < A.B) {
      embed_site(A.emb, A.x + ((size_t)b * A.L + w0 + k) * A.emb.n_in, ph);
    } else {
      for (int s = 0; s < A.D; ++s) ph[s] = T(0);
    }
    for (int s = 0; s < A.D; ++s) phi_s[((size_t)sl * A.S + k) * A.D + s] = ph[s];
  }
}

template <typename T> __global__ void __launch_bounds__(256) smpo_fwd_kernel(SmpoArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int chip = A.chip, ld = A.ld, TPS = A.TPS, SPC = A.SPC;
  const size_t slot = (size_t)chip * ld;
  constexpr int NSLOT = 6;
  T *mats = reinterpret_cast<T *>(smem_raw);
  T *phi_s = mats + slot * NSLOT * SPC;
  T *red_s = phi_s + (size_t)SPC * A.S * A.D;
  __shared__ double lsum_s[8];
  const int tid = threadIdx.x;
  const int sl = tid / TPS, lt = tid - sl * TPS;
  const int nb = chip / 4;
  const int i0 = (lt / nb) * 4, j0 = (lt % nb) * 4;
  const long long b0 = (long long)blockIdx.x * SPC;
  const long long b = b0 + sl;
  const bool valid = b < A.B;
  T *base = mats + slot * NSLOT * sl;
  T *e = base, *e2 = base + slot, *p = base + 2 * slot, *m = base + 3 * slot, *t = base + 4 * slot,
    *u = base + 5 * slot;
  blk_zero(e, ld, i0, j0);
  if (i0 == 0 && j0 == 0) e[0] = T(1);
  int esum = 0;
  for (int w = 0; w < A.nwin; ++w) {
    const int w0 = w * A.S, w1 = min(w0 + A.S, A.L);
    window_phi(A, b0, w0, w1, phi_s);
    if (A.stash && valid) blk_copy_to_global(A.Estash + ((size_t)w * A.B + b) * chip * chip, e, chip, ld, i0, j0);
    __syncthreads();
    const T *phw = phi_s + (size_t)sl * A.S * A.D;
    bool havep = false;
    for (int j = w0 + 1; j < w1; ++j) {
      form_m(A, j, 0, phw + (size_t)(j - w0) * A.D, m, i0, j0);
      __syncthreads();
      if (!havep) {
        T *tmp = p; p = m; m = tmp;
        havep = true;
      } else {
        mm<T, false, false>(t, p, m, chip, ld, i0, j0, T(1), false);
        __syncthreads();
        T *tmp = p; p = t; t = tmp;
      }
    }
    blk_zero(e2, ld, i0, j0);
    for (int o = 0; o < A.DO; ++o) {
      form_m(A, w0, o, phw, m, i0, j0);
      __syncthreads();
      const T *Tm = m;
      if (havep) {
        mm<T, false, false>(t, m, p, chip, ld, i0, j0, T(1), false);
        __syncthreads();
        Tm = t;
      }
      mm<T, false, false>(u, e, Tm, chip, ld, i0, j0, T(1), false);
      __syncthreads();
      mm<T, true, false>(e2, Tm, u, chip, ld, i0, j0, T(1), true);
      __syncthreads();
    }
    // exact power-of-two rescale of the new environment
    T mx = T(0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) mx = Num<T>::max(mx, Num<T>::abs(e2[(i0 + i) * ld + j0 + j]));
    red_s[tid] = mx;
    __syncthreads();
    mx = T(0);
    for (int q = 0; q < TPS; ++q) mx = Num<T>::max(mx, red_s[sl * TPS + q]);
    int de = 0;
    if (mx > T(0)) Num<T>::frexp(mx, &de);
    const T sc = Num<T>::ldexp(T(1), -de);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) e2[(i0 + i) * ld + j0 + j] *= sc;
    esum += de;
    if (lt == 0 && valid && A.dexp) A.dexp[(size_t)w * A.B + b] = de;
    __syncthreads();
    T *tmp = e; e = e2; e2 = tmp;
  }
  double l = 0.0;
  if (lt == 0 && valid) {
    T nmv = e[0];
    A.nm[b] = nmv;
    A.esum[b] = esum;
    if (A.out) A.out[b] = nmv;
    if (A.out_exp) A.out_exp[b] = esum;
    T lv = smpo_head<T>(A.head, nmv, esum, (T *)nullptr);
    if (A.loss) A.loss[b] = lv;
    l = (double)lv;
  }
  if (A.loss_sum) {
    for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
    if ((tid & 31) == 0) lsum_s[tid >> 5] = l;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int q = 0; q < (int)(blockDim.x + 31) / 32; ++q) tot += lsum_s[q];
      atomicAdd(A.loss_sum, tot);
    }
  }
}

// sum over the CTA's samples of phi[s] * dM  ->  one atomicAdd per (a, r, s)
template <typename T>
__device__ __forceinline__ void accum_grad(const SmpoArgs<T> &A, int site, int o, const T *mats, size_t sample_stride,
                                           size_t slot_off, const T *phi_s, int k) {
  const int nj = (site % A.S == 0) ? A.DO : 1;
  T *g = A.grads + smpo_site_off(A.chi, A.D, A.DO, A.S, site);
  const int total = A.chi * A.chi * A.D;
  const int CD = A.chi * A.D;
  if ((int)blockDim.x % CD == 0) {  // a thread keeps its (r, s) column and walks the rows: no divisions in the loop
    const int cidx = threadIdx.x % CD, astep = blockDim.x / CD;
    const int r = cidx / A.D, sidx = cidx - r * A.D;
    for (int a = threadIdx.x / CD; a < A.chi; a += astep) {
      T v = T(0);
      for (int q = 0; q < A.SPC; ++q)
        v = fma(phi_s[((size_t)q * A.S + k) * A.D + sidx], mats[sample_stride * q + slot_off + (size_t)a * A.ld + r], v);
      atomicAdd(&g[((size_t)(a * A.chi + r) * A.D + sidx) * nj + o], v);
    }
    return;
  }
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    int s = idx % A.D;
    int ar = idx / A.D;
    int a = ar / A.chi, r = ar - a * A.chi;
    T v = T(0);
    for (int q = 0; q < A.SPC; ++q)
      v = fma(phi_s[((size_t)q * A.S + k) * A.D + s], mats[sample_stride * q + slot_off + (size_t)a * A.ld + r], v);
    atomicAdd(&g[(size_t)idx * nj + o], v);
  }
}

template <typename T> __global__ void __launch_bounds__(256) smpo_bwd_kernel(SmpoArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int chip = A.chip, ld = A.ld, TPS = A.TPS, SPC = A.SPC;
  const size_t slot = (size_t)chip * ld;
  constexpr int NSLOT = 8;
  T *mats = reinterpret_cast<T *>(smem_raw);
  T *phi_s = mats + slot * NSLOT * SPC;
  const int tid = threadIdx.x;
  const int sl = tid / TPS, lt = tid - sl * TPS;
  const int nb = chip / 4;
  const int i0 = (lt / nb) * 4, j0 = (lt % nb) * 4;
  const long long b0 = (long long)blockIdx.x * SPC;
  const long long b = b0 + sl;
  const bool valid = b < A.B;
  const size_t sstride = slot * NSLOT;
  T *base = mats + sstride * sl;
  T *de = base, *de2 = base + slot, *p = base + 2 * slot, *m = base + 3 * slot, *t = base + 4 * slot,
    *u = base + 5 * slot, *dt = base + 6 * slot, *dp = base + 7 * slot;
  T *Q = A.Qscr + ((size_t)blockIdx.x * SPC + sl) * (size_t)(A.S > 1 ? A.S - 1 : 1) * chip * chip;
  blk_zero(de, ld, i0, j0);
  if (i0 == 0 && j0 == 0) {
    T dn = T(0);
    if (valid) {
      smpo_head<T>(A.head, A.nm[b], A.esum[b], &dn);
      dn *= A.loss_scale;
      if (A.sw) dn *= A.sw[b];
    }
    de[0] = dn;
  }
  for (int w = A.nwin - 1; w >= 0; --w) {
    const int w0 = w * A.S, w1 = min(w0 + A.S, A.L);
    const int K = w1 - w0 - 1;
    window_phi(A, b0, w0, w1, phi_s);
    const int dexp = valid ? A.dexp[(size_t)w * A.B + b] : 0;
    const T sc = Num<T>::ldexp(T(1), -dexp);
    const T *Eg = A.Estash + ((size_t)w * A.B + (valid ? b : 0)) * chip * chip;
    __syncthreads();
    const T *phw = phi_s + (size_t)sl * A.S * A.D;
    // recompute the chain, keeping the prefix products Q_k = M_1 .. M_k
    bool havep = false;
    for (int k = 1; k <= K; ++k) {
      form_m(A, w0 + k, 0, phw + (size_t)k * A.D, m, i0, j0);
      __syncthreads();
      if (!havep) {
        T *tmp = p; p = m; m = tmp;
        havep = true;
      } else {
        mm<T, false, false>(t, p, m, chip, ld, i0, j0, T(1), false);
        __syncthreads();
        T *tmp = p; p = t; t = tmp;
      }
      blk_copy_to_global(Q + (size_t)(k - 1) * chip * chip, p, chip, ld, i0, j0);
    }
    blk_zero(de2, ld, i0, j0);
    blk_zero(dp, ld, i0, j0);
    for (int o = 0; o < A.DO; ++o) {
      form_m(A, w0, o, phw, m, i0, j0);
      blk_copy_from_global(dt, Eg, chip, ld, i0, j0);  // E_w into the dt slot
      __syncthreads();
      const T *Tm = m;
      if (havep) {
        mm<T, false, false>(t, m, p, chip, ld, i0, j0, T(1), false);
        __syncthreads();
        Tm = t;
      }
      mm<T, false, false>(u, dt, Tm, chip, ld, i0, j0, T(1), false);  // U = E T
      __syncthreads();
      mm<T, false, false>(dt, u, de, chip, ld, i0, j0, T(2) * sc, false);  // dT = 2 sc U dE'
      __syncthreads();
      mm<T, false, false>(u, Tm, de, chip, ld, i0, j0, T(1), false);  // X = T dE'
      __syncthreads();
      mm<T, false, true>(de2, u, Tm, chip, ld, i0, j0, sc, true);  // dE_w += sc X T^T
      __syncthreads();
      size_t dmo_off;
      if (havep) {
        mm<T, false, true>(u, dt, p, chip, ld, i0, j0, T(1), false);  // dM^o = dT P^T
        __syncthreads();
        dmo_off = (size_t)(u - base);
      } else {
        dmo_off = (size_t)(dt - base);
      }
      accum_grad(A, w0, o, mats, sstride, dmo_off, phi_s, 0);
      if (havep) mm<T, true, false>(dp, m, dt, chip, ld, i0, j0, T(1), true);  // dP += M^o^T dT
      __syncthreads();
    }
    // chain backward: G = dP ; dM_k = Q_{k-1}^T G ; G <- G M_k^T
    for (int k = K; k >= 1; --k) {
      size_t dm_off;
      if (k > 1) {
        blk_copy_from_global(t, Q + (size_t)(k - 2) * chip * chip, chip, ld, i0, j0);
        form_m(A, w0 + k, 0, phw + (size_t)k * A.D, m, i0, j0);
        __syncthreads();
        mm<T, true, false>(u, t, dp, chip, ld, i0, j0, T(1), false);
        __syncthreads();
        dm_off = (size_t)(u - base);
      } else {
        dm_off = (size_t)(dp - base);
      }
      accum_grad(A, w0 + k, 0, mats, sstride, dm_off, phi_s, k);
      if (k > 1) {
        mm<T, false, true>(t, dp, m, chip, ld, i0, j0, T(1), false);
        __syncthreads();
        T *tmp = dp; dp = t; t = tmp;
      } else {
        __syncthreads();
      }
    }
    T *tmp = de; de = de2; de2 = tmp;
  }
}

static inline SmpoPlan plan_smpo(const tnb_problem *p, long long B, int need_grad) {
  SmpoPlan P{};
  const size_t w = p->dtype == TNB_F64 ? 8 : 4;
  P.chip = p->chi <= 16 ? 16 : (p->chi <= 32 ? 32 : 64);
  P.ld = P.chip + (int)(16 / w);
  P.TPS = (P.chip / 4) * (P.chip / 4);
  P.nwin = (p->L + p->spacing - 1) / p->spacing;
  const int nslot = need_grad ? 8 : 6;
  size_t per_sample = (size_t)nslot * P.chip * P.ld * w;
  int spc = 256 / P.TPS;
  if (spc < 1) spc = 1;
  while (spc > 1 && per_sample * spc + 8192 > 200 * 1024) --spc;
  if (per_sample * spc + 8192 > 200 * 1024) spc = 0;
  P.SPC = spc;
  P.NT = spc * P.TPS;
  P.grid = spc ? (int)((B + spc - 1) / spc) : 0;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) / 256 * 256; return r; };
  P.off_E = take(need_grad ? (size_t)P.nwin * B * P.chip * P.chip * w : 0);
  P.off_dexp = take((size_t)P.nwin * B * sizeof(int));
  P.off_esum = take((size_t)B * sizeof(int));
  P.off_nm = take((size_t)B * w);
  P.off_Q = take(need_grad ? (size_t)P.grid * (spc ? spc : 1) * (p->spacing > 1 ? p->spacing - 1 : 1) * P.chip * P.chip * w : 0);
  P.total = o;
  size_t extra = ((size_t)(spc ? spc : 1) * p->spacing * p->d + 256) * w;
  P.smem_fwd = (size_t)6 * P.chip * P.ld * w * (spc ? spc : 1) + extra;
  P.smem_bwd = (size_t)8 * P.chip * P.ld * w * (spc ? spc : 1) + extra;
  P.launches_fwd = 1;
  P.launches_bwd = 1;
  return P;
}

template <typename T>
static void fill_smpo_args(const tnb_problem *p, const SmpoPlan &P, long long B, const void *x, const void *params,
                           void *ws, SmpoArgs<T> &A) {
  char *w = (char *)ws;
  A.L = p->L; A.chi = p->chi; A.chip = P.chip; A.ld = P.ld; A.D = p->d; A.DO = p->n_out; A.S = p->spacing;
  A.nwin = P.nwin; A.head = p->head; A.B = B;
  fill_emb<T>(p->emb, A.emb, p->d);
  A.x = (const T *)x; A.params = (const T *)params;
  A.Estash = (T *)(w + P.off_E); A.dexp = (int *)(w + P.off_dexp); A.esum = (int *)(w + P.off_esum);
  A.nm = (T *)(w + P.off_nm); A.Qscr = (T *)(w + P.off_Q);
  A.loss = nullptr; A.loss_sum = nullptr; A.out = nullptr; A.out_exp = nullptr; A.grads = nullptr;
  A.loss_scale = T(1); A.sw = nullptr; A.stash = 0; A.SPC = P.SPC; A.TPS = P.TPS;
}

template <typename T> static cudaError_t smpo_set_attrs_t() {
  cudaError_t e = cudaFuncSetAttribute(smpo_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmpoSmemMax);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(smpo_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmpoSmemMax);
  return e;
}

template <typename T>
static int smpo_forward_t(const tnb_problem *p, const SmpoPlan &P, long long B, const void *x, const void *params,
                          void *loss, double *loss_sum, void *out, int32_t *out_exp, int keep_stash, void *ws,
                          cudaStream_t st, const char **what) {
  if (P.SPC < 1) { *what = "bond dimension too large for the SMPO shared-memory plan"; return 1; }
  SmpoArgs<T> A;
  fill_smpo_args<T>(p, P, B, x, params, ws, A);
  A.stash = keep_stash ? 1 : 0;
  A.loss = (T *)loss; A.loss_sum = loss_sum; A.out = (T *)out; A.out_exp = out_exp;
  smpo_fwd_kernel<T><<<P.grid, P.NT, P.smem_fwd, st>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *what = cudaGetErrorString(e); return 1; }
  return 0;
}

template <typename T>
static int smpo_backward_t(const tnb_problem *p, const SmpoPlan &P, long long B, const void *x, const void *params,
                           double loss_scale, const void *sample_w, void *grads, int accumulate, void *ws,
                           cudaStream_t st, const char **what) {
  if (P.SPC < 1) { *what = "bond dimension too large for the SMPO shared-memory plan"; return 1; }
  SmpoArgs<T> A;
  fill_smpo_args<T>(p, P, B, x, params, ws, A);
  A.stash = 1;
  A.grads = (T *)grads;
  A.loss_scale = (T)loss_scale;
  A.sw = (const T *)sample_w;
  if (!accumulate) {
    long long n = smpo_site_off(p->chi, p->d, p->n_out, p->spacing, p->L);
    cudaError_t e = cudaMemsetAsync(grads, 0, (size_t)n * sizeof(T), st);
    if (e != cudaSuccess) { *what = cudaGetErrorString(e); return 1; }
  }
  smpo_bwd_kernel<T><<<P.grid, P.NT, P.smem_bwd, st>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *what = cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace tnb
