// K5 (register-resident family): SMPO / MPO double-layer environments for bond <= 32, float32, d = 2.
//
// One WARP owns one sample.  Every per-sample chi x chi matrix of the window recursion
//     P_w = M_{w0+1} ... M_{w1-1},   T^o = M^o_{w0} P_w,   U^o = E_w T^o,   E_{w+1} = sum_o T^o^T U^o
// (tn4ml/models/smpo.py:269-404 + tn4ml/metrics.py:56-81; the association of SURVEY 8d, cfg 3) lives in the warp's registers as
// tensor-core fragments: the fp32 accumulator of one product is re-split in registers into the fp16 hi/lo operand of the next
// (`mma.sync.m16n8k16`), transposed operands come from `movmatrix`, and nothing but the window's entry environment (the
// VJP's stash) ever goes to memory.  No shared memory, no block barrier.
//
// Why mma.sync and not tcgen05 here: these are products of two PER-SAMPLE 32 x 32 matrices.  A tcgen05.mma shares its B
// operand between all 128 rows of the tile, so four samples' products would run block-diagonally at 25 % efficiency with a
// TMEM round trip (tcgen05.ld -> convert -> tcgen05.st) between every two products; the shared-weight reformulation
// (E <- sum M^T E M as chain steps with chi rows per sample) costs 3x the flops.  DESIGN.md has the numbers.
//
// Arithmetic: "fp16x3" as on the tcgen05 family -- x * 2^s = hi + lo (two fp16, 22 bits), hi*hi + lo*hi + hi*lo accumulated in
// fp32.  Tensor-core accumulation truncates, so per output tile the two small terms are issued first and the hi*hi terms
// last (two truncations at full magnitude per product), and sums over the output index o are added in fp32 on the CUDA cores.
// Every matrix is carried as (fp32 fragment, binary exponent): the split rescales by the exact power of two of the warp-wide
// max-abs, so no chain length leaves the fp16 / fp32 range.
#pragma once
#include <cuda_fp16.h>
#include <limits.h>

#include "common.cuh"
#include "consts.h"

namespace tnb {
namespace rr {

constexpr int kN = 32;  // padded bond

struct RrArgs {
  int L, chi, D, DO, S, nwin, head;
  long long B;
  EmbParams<float> emb;
  const float *x;
  const float *params;
  float *Estash;  // [nwin][B][32 regs][32 lanes]  entry environment of every window (C-layout fragment image)
  int *Eexp;      // [nwin][B]
  int *esum;      // [B]
  float *nm;      // [B] mantissa of n = nm * 2^esum
  float *loss;
  double *loss_sum;
  float *out;
  int *out_exp;
  float *grads;
  float loss_scale;
  const float *sw;  // [B] per-sample weight of the loss in the VJP (may be null)
  int stash;
  // VJP
  float *dEscr;   // [B][1024] adjoint of the window environment between two windows (fragment image) ...
  int *dEexp;     // [B]       ... and its exponent
  uint32_t *Qscr; // [grid * warps][S][1056] prefix products Q_k = M_1 .. M_k of the current window (Q-form images + exponent)
  int per_cta;    // samples per CTA (window-major loop: every warp takes per_cta / warps samples one after the other)
  // weight images (smpo_rr_prep_kernel): per unit (a leg-less site, or one output index of an output site) the two slices
  // A[:, :, s = 0 / 1] in FRAGMENT ORDER, pre-scaled by 2^Wsh[unit]:  Wimg[(unit * 2 + form) * 512 + reg * 32 + lane]
  //   form 0 ("P"): (A0, A1 at (row, col), A0, A1 at (row, col + 1)), row = 16 mt + 8 r + g, col = 8 nt + 2 t
  //   form 1 ("Q"): (A0, A1 at (row, col), A0, A1 at (row + 1, col)), row = 16 mt + 8 r + 2 t, col = 8 nt + g
  const float4 *Wimg;
  const int *Wsh;
};

constexpr int kRingStages = 3;              // staging ring of weight images in shared memory (8 KB each)
constexpr int kImgFloat4 = 512;             // float4 per image

__host__ __device__ inline int unit_of(int site, int o, int S, int DO) {
  return site + (DO - 1) * ((site + S - 1) / S) + (site % S == 0 ? o : 0);
}

// fp32 matrix in accumulator (C) layout: tile (mt, nt) = rows 16 mt .. +15, cols 8 nt .. +7; lane (g = lane / 4, t = lane % 4)
// holds [0],[1] = (row g, cols 2t, 2t+1) and [2],[3] = (row g + 8, cols 2t, 2t+1).   true matrix = c * 2^ex
struct Mat {
  float c[2][4][4];
  int ex;
};
// fp16 hi / lo image.  "P" form: reg [mt][nt][r] = the two values (row 16 mt + 8 r + g, cols 8 nt + 2t, +1) -- the row
// fragments of the matrix.  "Q" form: every 8 x 8 block of P transposed (movmatrix): reg [mt][nt][r] = (rows 16 mt + 8 r + 2t,
// +1; col 8 nt + g).  P(X) is the left operand X and the right operand X^T; Q(X) the right operand X and the left operand X^T.
struct Img {
  uint32_t hi[2][4][2], lo[2][4][2];
  int ex;
};

// max over the warp of a NON-NEGATIVE float (its bit pattern orders like the value; a NaN comes out on top): one redux.sync
__device__ __forceinline__ float warp_max(float v) {
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}
// 2^e for e in [-126, 127] without the ldexpf sequence
__device__ __forceinline__ float pow2i(int e) { return __int_as_float((e + 127) << 23); }

// split a matrix into its fp16 hi / lo row-fragment image, rescaled so that the largest entry is in [2^13, 2^14)
__device__ __forceinline__ void split(const Mat &X, Img &P) {
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) m = fmaxf(m, fabsf(X.c[i][j][q]));
  m = warp_max(m);
  // m = f * 2^e with f in [0.5, 1): scale by 2^(14 - e).  (Clamped: products are renormalised at every split, so |14 - e| > 120
  // only occurs for an all-zero or non-finite matrix.)
  const bool ok = m > 0.f && m < 3.0e38f;
  const int e = ((__float_as_int(m) >> 23) & 0xff) - 126;  // frexp exponent of a normal float (denormals clamp below)
  const int sh = ok ? max(-120, min(120, 14 - e)) : 0;
  const float fs = ok ? pow2i(sh) : 0.f;
  P.ex = X.ex - sh;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float a = X.c[i][j][2 * r] * fs, b = X.c[i][j][2 * r + 1] * fs;
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
        P.hi[i][j][r] = *reinterpret_cast<const uint32_t *>(&h);
        P.lo[i][j][r] = *reinterpret_cast<const uint32_t *>(&l);
      }
}

__device__ __forceinline__ uint32_t movm(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
  return d;
}
// Q form of an image in P form (and back: the transposition is an involution)
__device__ __forceinline__ void transpose_blocks(const Img &P, Img &Q) {
  Q.ex = P.ex;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        Q.hi[i][j][r] = movm(P.hi[i][j][r]);
        Q.lo[i][j][r] = movm(P.lo[i][j][r]);
      }
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Z = op(X) * op(Y) for 32 x 32 matrices.
//   LT = false: left operand X from P(X);  LT = true: left operand X^T from Q(X)
//   RT = false: right operand Y from Q(Y); RT = true: right operand Y^T from P(Y)
template <bool LT, bool RT> __device__ __forceinline__ void product(const Img &Lm, const Img &Rm, Mat &Z) {
  Z.ex = Lm.ex + Rm.ex;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) Z.c[mt][nt][q] = 0.f;
  // Term-major order: the eight output tiles take the same term one after the other, so that consecutive MMAs are independent
  // (a tile's six accumulations are a dependent chain of ~21 cycles each).  Small terms first, the hi * hi terms last: the
  // tensor core truncates every accumulation.
#pragma unroll
  for (int term = 0; term < 6; ++term) {
    const int kt = term < 4 ? (term >> 1) : term - 4;       // 0 0 1 1 | 0 1
    const bool a_lo = term < 4 && (term & 1) == 0;           // lo*hi, hi*lo, lo*hi, hi*lo | hi*hi, hi*hi
    const bool b_lo = term < 4 && (term & 1) == 1;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      uint32_t a[4];
      if (!LT) {
        a[0] = a_lo ? Lm.lo[mt][2 * kt][0] : Lm.hi[mt][2 * kt][0];
        a[1] = a_lo ? Lm.lo[mt][2 * kt][1] : Lm.hi[mt][2 * kt][1];
        a[2] = a_lo ? Lm.lo[mt][2 * kt + 1][0] : Lm.hi[mt][2 * kt + 1][0];
        a[3] = a_lo ? Lm.lo[mt][2 * kt + 1][1] : Lm.hi[mt][2 * kt + 1][1];
      } else {
        a[0] = a_lo ? Lm.lo[kt][2 * mt][0] : Lm.hi[kt][2 * mt][0];
        a[1] = a_lo ? Lm.lo[kt][2 * mt + 1][0] : Lm.hi[kt][2 * mt + 1][0];
        a[2] = a_lo ? Lm.lo[kt][2 * mt][1] : Lm.hi[kt][2 * mt][1];
        a[3] = a_lo ? Lm.lo[kt][2 * mt + 1][1] : Lm.hi[kt][2 * mt + 1][1];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bb[2];
        if (!RT) {
          bb[0] = b_lo ? Rm.lo[kt][nt][0] : Rm.hi[kt][nt][0];
          bb[1] = b_lo ? Rm.lo[kt][nt][1] : Rm.hi[kt][nt][1];
        } else {
          bb[0] = b_lo ? Rm.lo[nt >> 1][2 * kt][nt & 1] : Rm.hi[nt >> 1][2 * kt][nt & 1];
          bb[1] = b_lo ? Rm.lo[nt >> 1][2 * kt + 1][nt & 1] : Rm.hi[nt >> 1][2 * kt + 1][nt & 1];
        }
        mma16816(Z.c[mt][nt], a, bb);
      }
    }
  }
}

// Z += X with the exponents aligned (exact powers of two); empty Z (ex == INT_MIN) takes X
__device__ __forceinline__ void add_aligned(Mat &Z, const Mat &X) {
  if (Z.ex == INT_MIN) {
    Z = X;
    return;
  }
  const int e = max(Z.ex, X.ex);
  const int dz = Z.ex - e, dx = X.ex - e;  // one of them is 0, the other <= 0
  const float fz = dz < -126 ? 0.f : pow2i(dz), fx = dx < -126 ? 0.f : pow2i(dx);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) Z.c[i][j][q] = fmaf(Z.c[i][j][q], fz, X.c[i][j][q] * fx);
  Z.ex = e;
}

// ---------------------------------------------------------------------------------------------------------------
// VJP helpers + kernel.  Window-major: a CTA owns `per_cta` samples and walks the windows backwards; inside a window its W warps take the
// samples W at a time.  For one sample and window (E = stashed entry environment, dE' = adjoint of the exit environment):
//     recompute  Q_k = M_1 .. M_k (prefixes to scratch),  P = Q_K,  T^o = M^o P,  X^o = T^o dE'
//     dT^o = 2 E X^o (= 2 U^o dE')      dE += X^o T^o^T      dM^o = dT^o P^T      dP += M^o^T dT^o
//     G = dP;  for k = K .. 1:  dM_k = Q_{k-1}^T G,  G <- G M_k^T
// The per-site gradients dA[a, r, s(, o)] = sum_b phi_s[b] dM[b][a, r] are reduced over the CTA's samples in shared memory
// (every warp stages its phi-weighted fragment, the warps then sum disjoint slices of the W slots into the window's
// accumulator) and go to global memory with one RED per element, window and CTA.
// ---------------------------------------------------------------------------------------------------------------
#ifndef TNB_RR_BWD_WARPS
#define TNB_RR_BWD_WARPS 8
#endif
constexpr int kBwdWarps = TNB_RR_BWD_WARPS;

// four adjacent floats added to global memory with one instruction (16-byte aligned address)
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void store_mat(float *dst, const Mat &X, int lane) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[((i * 4 + j) * 4 + q) * 32 + lane] = X.c[i][j][q];
}
__device__ __forceinline__ void load_mat(const float *src, Mat &X, int lane) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) X.c[i][j][q] = src[((i * 4 + j) * 4 + q) * 32 + lane];
}
__device__ __forceinline__ void store_img(uint32_t *dst, const Img &X, int lane) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        dst[((i * 4 + j) * 2 + r) * 32 + lane] = X.hi[i][j][r];
        dst[(16 + (i * 4 + j) * 2 + r) * 32 + lane] = X.lo[i][j][r];
      }
  if (lane == 0) dst[1024] = (uint32_t)X.ex;
}
__device__ __forceinline__ void load_img(const uint32_t *src, Img &X, int lane) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        X.hi[i][j][r] = src[((i * 4 + j) * 2 + r) * 32 + lane];
        X.lo[i][j][r] = src[(16 + (i * 4 + j) * 2 + r) * 32 + lane];
      }
  X.ex = (int)src[1024];
}

// Stage phi_s * dM of this warp's sample, reduce over the CTA's warps, add into the window accumulator `acc` ([2][1024]:
// s, fragment element).  Called by EVERY warp of the CTA (two block barriers); `on` = this warp has something to add.
// Element order of a staged / accumulated plane: ((mt * 4 + nt) * 32 + lane) * 4 + q  -- a lane's four values of one tile are
// one 16-byte store, and a lane sums four consecutive elements of the eight slots with 16-byte loads.
__device__ __forceinline__ void reduce_grad(float *stage, float *acc, const Mat &dM, bool on, float ph0, float ph1, int warp,
                                            int lane) {
  float *mine = stage + warp * 2048;
  const float f = on ? pow2i(max(-126, min(126, dM.ex))) : 0.f;
  const float f0 = ph0 * f, f1 = ph1 * f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = on ? dM.c[i][j][q] : 0.f;
      float4 *dst = reinterpret_cast<float4 *>(mine + ((i * 4 + j) * 32 + lane) * 4);
      dst[0] = make_float4(v[0] * f0, v[1] * f0, v[2] * f0, v[3] * f0);
      dst[256] = make_float4(v[0] * f1, v[1] * f1, v[2] * f1, v[3] * f1);  // plane s = 1: + 1024 floats
    }
  __syncthreads();
  constexpr int per_warp = 2048 / kBwdWarps;  // elements of the slice a warp sums
  static_assert(per_warp % 128 == 0, "a warp sums whole rows of 32 float4");
#pragma unroll
  for (int i = 0; i < per_warp / 128; ++i) {
    const int e = warp * per_warp + i * 128 + lane * 4;
    float4 v = *reinterpret_cast<const float4 *>(stage + e);
#pragma unroll
    for (int w = 1; w < kBwdWarps; ++w) {
      const float4 u = *reinterpret_cast<const float4 *>(stage + w * 2048 + e);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    float4 *a = reinterpret_cast<float4 *>(acc + e);
    float4 t = *a;
    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    *a = t;
  }
  __syncthreads();
}

// ---- weight images ---------------------------------------------------------------------------------------------------
// One thread per (unit, form, reg, lane).  The power-of-two scale of a unit comes from an a-priori bound: |phi0 A0 + phi1 A1|
// <= sqrt(2) max|A| for a unit-norm phi, so sh = 14 - exponent(sqrt(2) max|A|) puts every site matrix below 2^14 without a
// per-sample maximum.
__global__ void smpo_rr_unitmax_kernel(const float *__restrict__ params, int *__restrict__ Wsh, int L, int chi, int S, int DO) {
  const int unit = blockIdx.x;
  // unit -> (site, o)
  int site = 0, o = 0;
  {
    // units are ordered by site; an output site contributes DO units
    const int per_win = S - 1 + DO;
    const int w = unit / per_win, rem = unit - w * per_win;
    if (rem < DO) { site = w * S; o = rem; }
    else { site = w * S + (rem - DO + 1); o = 0; }
  }
  if (site >= L) { if (threadIdx.x == 0) Wsh[unit] = 0; return; }
  const int nj = site % S == 0 ? DO : 1;
  const long long per = (long long)chi * chi * 2;
  const long long nwith = (site + S - 1) / S;
  const float *W = params + per * (site - nwith) + per * DO * nwith;
  float m = 0.f;
  for (int i = threadIdx.x; i < chi * chi * 2; i += blockDim.x) m = fmaxf(m, fabsf(W[(size_t)i * nj + o]));
  __shared__ float red[32];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x + 31) / 32; ++w) m = fmaxf(m, red[w]);
    int e = 0;
    if (m > 0.f) frexpf(m * 1.41421357f, &e);
    Wsh[unit] = m > 0.f ? max(-100, min(100, 14 - e)) : 0;
  }
}

__global__ void smpo_rr_prep_kernel(const float *__restrict__ params, const int *__restrict__ Wsh, float4 *__restrict__ Wimg,
                                    int nunit, int L, int chi, int S, int DO) {
  const long long total = (long long)nunit * 2 * kImgFloat4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int lane = (int)(i & 31), reg = (int)((i >> 5) & 15), form = (int)((i >> 9) & 1), unit = (int)(i >> 10);
    const int per_win = S - 1 + DO;
    const int w = unit / per_win, rem = unit - w * per_win;
    int site, o;
    if (rem < DO) { site = w * S; o = rem; }
    else { site = w * S + (rem - DO + 1); o = 0; }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (site < L) {
      const int nj = site % S == 0 ? DO : 1;
      const long long per = (long long)chi * chi * 2;
      const long long nwith = (site + S - 1) / S;
      const float *W = params + per * (site - nwith) + per * DO * nwith;
      const float sc = ldexpf(1.f, Wsh[unit]);
      const int g = lane >> 2, t = lane & 3, r = reg & 1, nt = (reg >> 1) & 3, mt = reg >> 3;
      int row0, col0, row1, col1;
      if (form == 0) { row0 = row1 = 16 * mt + 8 * r + g; col0 = 8 * nt + 2 * t; col1 = col0 + 1; }
      else { row0 = 16 * mt + 8 * r + 2 * t; row1 = row0 + 1; col0 = col1 = 8 * nt + g; }
      auto at = [&](int row, int col, int sidx) -> float {
        return (row < chi && col < chi) ? W[(((size_t)row * chi + col) * 2 + sidx) * nj + o] * sc : 0.f;
      };
      v = make_float4(at(row0, col0, 0), at(row0, col0, 1), at(row1, col1, 0), at(row1, col1, 1));
    }
    Wimg[i] = v;
  }
}

// ---- staging ring: the CTA's warps walk the same sequence of weight images; image q of the sequence sits in ring slot
//      q % kRingStages, fetched kRingStages - 1 requests ahead with cp.async (every thread copies its share) ----
struct Request {
  int unit, form;
};

__device__ __forceinline__ void ring_issue(float4 *ring, int q, const RrArgs &A, Request rq, bool valid) {
  if (valid) {
    const float4 *src = A.Wimg + ((size_t)rq.unit * 2 + rq.form) * kImgFloat4;
    float4 *dst = ring + (size_t)(q % kRingStages) * kImgFloat4;
    for (int i = threadIdx.x; i < kImgFloat4; i += blockDim.x) cp_async16(dst + i, src + i);
  }
  cp_async_commit();  // (always: the group count must advance uniformly)
}
// image q is complete and visible to the whole CTA; the slot of image q - 1 may be refilled afterwards
__device__ __forceinline__ const float4 *ring_acquire(float4 *ring, int q) {
  cp_async_wait<kRingStages - 2>();
  __syncthreads();
  return ring + (size_t)(q % kRingStages) * kImgFloat4;
}

// site matrix M = phi0 A0 + phi1 A1 straight into its fp16 hi / lo image (P or Q form, as staged); ex = -Wsh[unit]
__device__ __forceinline__ void form_img(const float4 *buf, int sh, float ph0, float ph1, Img &X, int lane) {
  X.ex = -sh;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float4 v = buf[((mt * 4 + nt) * 2 + r) * 32 + lane];
        const float a = fmaf(ph1, v.y, ph0 * v.x), b = fmaf(ph1, v.w, ph0 * v.z);
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
        X.hi[mt][nt][r] = *reinterpret_cast<const uint32_t *>(&h);
        X.lo[mt][nt][r] = *reinterpret_cast<const uint32_t *>(&l);
      }
}

// phi of the window's sites: lane k evaluates site w0 + k; read with __shfl_sync
__device__ __forceinline__ void window_phi(const RrArgs &A, long long b, int w0, int w1, int lane, float &p0, float &p1) {
  p0 = p1 = 0.f;
  if (w0 + lane < w1) {
    float ph[2];
    embed_site_d<2, float, false>(A.emb, A.x + ((size_t)b * A.L + w0 + lane), ph);  // (n_in == 1: rr_smpo_supported)
    p0 = ph[0];
    p1 = ph[1];
  }
}

// Request j of window w.  Forward: chain sites k = 1 .. K (k = 1 in P form: it starts the product as a left operand; the rest
// in Q form: right operands), then the output site's DO units in P form.  VJP: the same, then for k = K .. 2 the P form of
// site k (G M_k^T), preceded at k = 2 by the Q form of site 1 (Q_1 = M_1).
// (No integer division on this path: every thread evaluates it once per image.  Units of window w start at w * (S - 1 + DO):
//  the DO output units first, then the chain sites.)
__device__ __forceinline__ int win_K(const RrArgs &A, int w) { return min(A.S, A.L - w * A.S) - 1; }
__device__ __forceinline__ int nreq_of(const RrArgs &A, int w, bool bwd) {
  const int K = win_K(A, w);
  return K + A.DO + (bwd ? (K > 1 ? K : 0) : 0);
}
__device__ __forceinline__ Request request_of(const RrArgs &A, int w, int j) {
  const int K = win_K(A, w), u0 = w * (A.S - 1 + A.DO), uc = u0 + A.DO - 1;  // chain site k is unit uc + k
  Request rq;
  if (j < K) { rq.unit = uc + j + 1; rq.form = j == 0 ? 0 : 1; return rq; }
  j -= K;
  if (j < A.DO) { rq.unit = u0 + j; rq.form = 0; return rq; }
  j -= A.DO;  // VJP tail: P(K), P(K-1), ..., P(3), Q(1), P(2)
  if (j < K - 2) { rq.unit = uc + K - j; rq.form = 0; return rq; }
  if (j == K - 2) { rq.unit = uc + 1; rq.form = 1; return rq; }
  rq.unit = uc + 2;
  rq.form = 0;
  return rq;
}
// walks the request sequence of the whole kernel (windows ascending for the forward pass, descending for the VJP, `reps`
// passes over every window: the VJP's sample rounds)
struct Cursor {
  int w, j, rep;
  bool done;
};
__device__ __forceinline__ void cursor_next(const RrArgs &A, Cursor &c, bool bwd, int reps) {
  if (c.done) return;
  if (++c.j < nreq_of(A, c.w, bwd)) return;
  c.j = 0;
  if (++c.rep < reps) return;
  c.rep = 0;
  if (bwd) { if (--c.w < 0) c.done = true; }
  else { if (++c.w >= A.nwin) c.done = true; }
}

#ifndef TNB_RR_FWD_WARPS
#define TNB_RR_FWD_WARPS 16
#endif
constexpr int kFwdWarps = TNB_RR_FWD_WARPS;

template <int DO> __global__ void __launch_bounds__(kFwdWarps * 32, 1) smpo_rr_fwd_kernel(RrArgs A) {
  __shared__ __align__(16) float4 ring[kRingStages * kImgFloat4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * kFwdWarps + warp;
  const bool on = b < A.B;  // (every warp takes part in the staging barriers)
  const long long bs = on ? b : 0;
  Cursor pf{0, 0, 0, false};
  int q_issue = 0, q_use = 0;
  for (; q_issue < kRingStages - 1; ++q_issue) {
    ring_issue(ring, q_issue, A, pf.done ? Request{0, 0} : request_of(A, pf.w, pf.j), !pf.done);
    cursor_next(A, pf, false, 1);
  }
  auto next_image = [&]() -> const float4 * {
    const float4 *buf = ring_acquire(ring, q_use);
    ring_issue(ring, q_issue, A, pf.done ? Request{0, 0} : request_of(A, pf.w, pf.j), !pf.done);
    cursor_next(A, pf, false, 1);
    ++q_issue;
    ++q_use;
    return buf;
  };
  Mat E;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) E.c[i][j][q] = 0.f;
  if (lane == 0) E.c[0][0][0] = 1.f;
  E.ex = 0;
  for (int w = 0; w < A.nwin; ++w) {
    const int w0 = w * A.S, w1 = min(w0 + A.S, A.L), K = w1 - w0 - 1;
    const int uc = w * (A.S - 1 + A.DO) + A.DO - 1;  // chain site k of this window is weight unit uc + k
    if (A.stash && on) {
      store_mat(A.Estash + ((size_t)w * A.B + b) * 1024, E, lane);
      if (lane == 0) A.Eexp[(size_t)w * A.B + b] = E.ex;
    }
    float p0, p1;
    window_phi(A, bs, w0, w1, lane, p0, p1);
    Img qP;
    if (K > 0) {
      Img a;
      form_img(next_image(), A.Wsh[uc + 1], __shfl_sync(0xffffffffu, p0, 1), __shfl_sync(0xffffffffu, p1, 1), a,
               lane);
      for (int k = 2; k <= K; ++k) {
        Img q;
        form_img(next_image(), A.Wsh[uc + k], __shfl_sync(0xffffffffu, p0, k),
                 __shfl_sync(0xffffffffu, p1, k), q, lane);
        Mat Pm;
        product<false, false>(a, q, Pm);
        split(Pm, a);
      }
      transpose_blocks(a, qP);
    }
    Img pE;
    split(E, pE);
    Mat En;
    En.ex = INT_MIN;
    const float q0 = __shfl_sync(0xffffffffu, p0, 0), q1 = __shfl_sync(0xffffffffu, p1, 0);
#pragma unroll 1
    for (int o = 0; o < A.DO; ++o) {
      Img pT, qT;
      form_img(next_image(), A.Wsh[uc + 1 - A.DO + o], q0, q1, pT, lane);
      if (K > 0) {
        Mat T;
        product<false, false>(pT, qP, T);
        split(T, pT);
      }
      transpose_blocks(pT, qT);
      Mat U;
      product<false, false>(pE, qT, U);
      Img pU, qU;
      split(U, pU);
      transpose_blocks(pU, qU);
      Mat C;
      product<true, false>(qT, qU, C);
      add_aligned(En, C);
    }
    E = En;
  }
  // n = E[0][0] * 2^ex
  if (lane == 0 && on) {
    int e2 = 0;
    const float nmv = frexpf(E.c[0][0][0], &e2);
    const int es = E.ex + e2;
    A.nm[b] = nmv;
    A.esum[b] = es;
    if (A.out) A.out[b] = nmv;
    if (A.out_exp) A.out_exp[b] = es;
    const float lval = smpo_head<float>(A.head, nmv, es, (float *)nullptr);
    if (A.loss) A.loss[b] = lval;
    if (A.loss_sum) atomicAdd(A.loss_sum, (double)lval);
  }
  cp_async_wait<0>();
}

template <int DO> __global__ void __launch_bounds__(kBwdWarps * 32, 1) smpo_rr_bwd_kernel(RrArgs A) {
  extern __shared__ __align__(16) float smem[];
  float *stage = smem;                        // [warps][2][1024]
  float *wacc = smem + kBwdWarps * 2048;      // window accumulator: site 0: [DO][2][1024], sites 1 .. S-1: [2][1024]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_floats = (A.DO + A.S - 1) * 2048;
  float4 *ring = reinterpret_cast<float4 *>(wacc + acc_floats);
  const long long base = (long long)blockIdx.x * A.per_cta;
  const int rounds = (A.per_cta + kBwdWarps - 1) / kBwdWarps;
  uint32_t *Qs = A.Qscr + ((size_t)blockIdx.x * kBwdWarps + warp) * (size_t)A.S * 1056;
  for (int i = threadIdx.x; i < acc_floats; i += blockDim.x) wacc[i] = 0.f;
  Cursor pf{A.nwin - 1, 0, 0, false};
  int q_issue = 0, q_use = 0;
  for (; q_issue < kRingStages - 1; ++q_issue) {
    ring_issue(ring, q_issue, A, pf.done ? Request{0, 0} : request_of(A, pf.w, pf.j), !pf.done);
    cursor_next(A, pf, true, rounds);
  }
  auto next_image = [&]() -> const float4 * {
    const float4 *buf = ring_acquire(ring, q_use);
    ring_issue(ring, q_issue, A, pf.done ? Request{0, 0} : request_of(A, pf.w, pf.j), !pf.done);
    cursor_next(A, pf, true, rounds);
    ++q_issue;
    ++q_use;
    return buf;
  };
  __syncthreads();
  for (int w = A.nwin - 1; w >= 0; --w) {
    const int w0 = w * A.S, w1 = min(w0 + A.S, A.L), K = w1 - w0 - 1;
    const int uc = w * (A.S - 1 + A.DO) + A.DO - 1;  // chain site k of this window is weight unit uc + k
    for (int r = 0; r < rounds; ++r) {
      const long long b = base + (long long)r * kBwdWarps + warp;
      const bool on = r * kBwdWarps + warp < A.per_cta && b < A.B;
      const long long bs = on ? b : 0;
      float p0, p1;
      window_phi(A, bs, w0, w1, lane, p0, p1);
      Mat E, dEn;
      load_mat(A.Estash + ((size_t)w * A.B + bs) * 1024, E, lane);
      E.ex = A.Eexp[(size_t)w * A.B + bs];
      if (w == A.nwin - 1) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) dEn.c[i][j][q] = 0.f;
        float dn = 0.f;
        if (lane == 0) {
          smpo_head<float>(A.head, A.nm[bs], A.esum[bs], &dn);
          dn *= A.loss_scale;
          if (A.sw) dn *= A.sw[bs];
        }
        dEn.c[0][0][0] = dn;
        dEn.ex = -A.esum[bs];
      } else {
        load_mat(A.dEscr + (size_t)bs * 1024, dEn, lane);
        dEn.ex = A.dEexp[bs];
      }
      // ---- recompute the chain, prefixes Q_2 .. Q_{K-1} to scratch ----
      Img pP, qP;
      if (K > 0) {
        form_img(next_image(), A.Wsh[uc + 1], __shfl_sync(0xffffffffu, p0, 1),
                 __shfl_sync(0xffffffffu, p1, 1), pP, lane);
        for (int k = 2; k <= K; ++k) {
          Img q;
          form_img(next_image(), A.Wsh[uc + k], __shfl_sync(0xffffffffu, p0, k),
                   __shfl_sync(0xffffffffu, p1, k), q, lane);
          Mat Pm;
          product<false, false>(pP, q, Pm);
          split(Pm, pP);
          if (k < K) {
            Img qq;
            transpose_blocks(pP, qq);
            store_img(Qs + (size_t)k * 1056, qq, lane);
          }
        }
        transpose_blocks(pP, qP);
      }
      Img pE, pdE;
      split(E, pE);
      split(dEn, pdE);
      Mat dE, dP;
      dE.ex = INT_MIN;
      dP.ex = INT_MIN;
      const float q0 = __shfl_sync(0xffffffffu, p0, 0), q1 = __shfl_sync(0xffffffffu, p1, 0);
#pragma unroll 1
      for (int o = 0; o < A.DO; ++o) {
        Img pM, pT;
        form_img(next_image(), A.Wsh[uc + 1 - A.DO + o], q0, q1, pM, lane);
        if (K > 0) {
          Mat T;
          product<false, false>(pM, qP, T);
          split(T, pT);
        } else {
          pT = pM;
        }
        Mat X;
        product<false, true>(pT, pdE, X);   // X = T dE'   (dE' symmetric: the right operand dE'^T from P(dE'))
        Img pX, qX;
        split(X, pX);
        transpose_blocks(pX, qX);
        Mat dT;
        product<false, false>(pE, qX, dT);  // dT = 2 U dE' = 2 E T dE' = 2 E X   (U itself is not needed in the VJP)
        dT.ex += 1;
        Mat dEc;
        product<false, true>(pX, pT, dEc);  // (T dE') T^T
        add_aligned(dE, dEc);
        Img pdT;
        split(dT, pdT);
        Mat dMo;
        if (K > 0) product<false, true>(pdT, pP, dMo);  // dT P^T
        else dMo = dT;
        reduce_grad(stage, wacc + (size_t)o * 2048, dMo, on, q0, q1, warp, lane);
        if (K > 0) {
          Img qM, qdT;
          transpose_blocks(pM, qM);
          transpose_blocks(pdT, qdT);
          Mat dPc;
          product<true, false>(qM, qdT, dPc);  // M^o^T dT
          add_aligned(dP, dPc);
        }
      }
      // ---- chain backward ----
      Mat G = dP;
      for (int k = K; k >= 1; --k) {
        const float pk0 = __shfl_sync(0xffffffffu, p0, k), pk1 = __shfl_sync(0xffffffffu, p1, k);
        float *acc_k = wacc + (size_t)(A.DO + k - 1) * 2048;
        if (k == 1) {
          reduce_grad(stage, acc_k, G, on, pk0, pk1, warp, lane);
          break;
        }
        Img pG, qG, qQ;
        split(G, pG);
        transpose_blocks(pG, qG);
        if (k - 1 >= 2) load_img(Qs + (size_t)(k - 1) * 1056, qQ, lane);
        else  // Q_1 = M_1
          form_img(next_image(), A.Wsh[uc + 1], __shfl_sync(0xffffffffu, p0, 1),
                   __shfl_sync(0xffffffffu, p1, 1), qQ, lane);
        Mat dMk;
        product<true, false>(qQ, qG, dMk);  // Q_{k-1}^T G
        reduce_grad(stage, acc_k, dMk, on, pk0, pk1, warp, lane);
        Img pMk;
        form_img(next_image(), A.Wsh[uc + k], pk0, pk1, pMk, lane);
        product<false, true>(pG, pMk, G);  // G M_k^T
      }
      if (on && w > 0) {
        store_mat(A.dEscr + (size_t)b * 1024, dE, lane);
        if (lane == 0) A.dEexp[b] = dE.ex;
      }
    }
    // ---- flush the window's gradients: one 16-byte RED per four elements, window and CTA.  In the packed layout
    //      [row][col][s][o] a leg-less site's (col, col + 1) x (s = 0, 1) and an output site's (s, o) quadruple (DO = 2) are
    //      contiguous; the accumulator holds them as [o][s][fragment element], element q / q + 1 = columns col / col + 1 ----
    __syncthreads();
    const int nsite = w1 - w0;
    const long long per = (long long)A.chi * A.chi * A.D;
    const bool vec_ok = (A.chi & 1) == 0;
    for (int idx = 0; idx < nsite; ++idx) {
      // site w0 is the window's output site; nwith = output sites before this one
      const long long nwith = w + (idx > 0 ? 1 : 0);
      float *g = A.grads + per * (w0 + idx - nwith) + per * A.DO * nwith;
      if (idx == 0 && DO == 2) {
        float *a0 = wacc, *a1 = wacc + 2048;
        for (int fe = threadIdx.x; fe < 1024; fe += blockDim.x) {
          const int q = fe & 3, ln = (fe >> 2) & 31, nt = (fe >> 7) & 3, mt = fe >> 9, gg = ln >> 2, tt = ln & 3;
          const int row = 16 * mt + 8 * (q >> 1) + gg, col = 8 * nt + 2 * tt + (q & 1);
          const float v0 = a0[fe], v1 = a1[fe], v2 = a0[1024 + fe], v3 = a1[1024 + fe];  // (s0,o0) (s0,o1) (s1,o0) (s1,o1)
          a0[fe] = a1[fe] = a0[1024 + fe] = a1[1024 + fe] = 0.f;
          if (row < A.chi && col < A.chi) red_add_v4(&g[((size_t)row * A.chi + col) * 4], v0, v1, v2, v3);
        }
      } else if (idx > 0 && vec_ok) {
        float *acc = wacc + (size_t)(A.DO + idx - 1) * 2048;
        for (int pi = threadIdx.x; pi < 512; pi += blockDim.x) {
          const int fe = pi * 2;                             // elements q = 2 qh (column col) and q + 1 (column col + 1)
          const int qh = (fe >> 1) & 1, ln = (fe >> 2) & 31, nt = (fe >> 7) & 3, mt = fe >> 9, gg = ln >> 2, tt = ln & 3;
          const int row = 16 * mt + 8 * qh + gg, col = 8 * nt + 2 * tt;
          const float2 p0 = *reinterpret_cast<const float2 *>(acc + fe), p1 = *reinterpret_cast<const float2 *>(acc + 1024 + fe);
          const float v0 = p0.x, v1 = p1.x, v2 = p0.y, v3 = p1.y;
          *reinterpret_cast<float2 *>(acc + fe) = make_float2(0.f, 0.f);
          *reinterpret_cast<float2 *>(acc + 1024 + fe) = make_float2(0.f, 0.f);
          if (row < A.chi && col < A.chi) red_add_v4(&g[((size_t)row * A.chi + col) * 2], v0, v1, v2, v3);
        }
      } else {
        const int nj = idx == 0 ? A.DO : 1;
        for (int o = 0; o < nj; ++o) {
          float *acc = wacc + (size_t)(idx == 0 ? o : A.DO + idx - 1) * 2048;
          for (int e = threadIdx.x; e < 2048; e += blockDim.x) {
            const int sx = e >> 10, fe = e & 1023;
            const int q = fe & 3, ln = (fe >> 2) & 31, nt = (fe >> 7) & 3, mt = fe >> 9, gg = ln >> 2, tt = ln & 3;
            const int row = 16 * mt + 8 * (q >> 1) + gg, col = 8 * nt + 2 * tt + (q & 1);
            const float v = acc[e];
            acc[e] = 0.f;
            if (row < A.chi && col < A.chi && v != 0.f) atomicAdd(&g[(((size_t)row * A.chi + col) * A.D + sx) * nj + o], v);
          }
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

}  // namespace rr
}  // namespace tnb
