// K1/K3/K4 (SIMT family): MPS environment sweeps, class site + heads, VJP.
//
// Generic in bond dimension (<= 512), physical dimension (<= 16) and dtype (f32/f64).
// It is the fp64 path of every config and the fp32 path below the tensor-core threshold.
//
//   recurrence (what XLA runs for metrics.py:47 and :473-476):
//     F[j+1][b,r] = 2^-delta_j[b] * sum_{a,s} F[j][b,a] * phi_j[b,s] * A_j[a,r,s]      (left half)
//     F[j][b,a]   = 2^-delta_j[b] * sum_{r,s} A_j[a,r,s] * phi_j[b,s] * F[j+1][b,r]    (right half)
//   delta_j[b] is the binary exponent of the row maximum: the environment is carried as
//   (mantissa vector, integer exponent) so that no chain length or bond dimension can
//   leave the floating-point range; powers of two are exact, so fp64 results
//   equal the reference's unscaled arithmetic to rounding.
#pragma once
#include "common.cuh"
#include "consts.h"

namespace tnb {

template <typename T> struct MpsArgs {
  int L, chi, chip, D, C, c, head;
  long long B;
  EmbParams<T> emb;
  T cw[kMaxOut];
  const T *x;
  const T *targets;
  const T *Wn;  // kernel-layout weights [slab][k=a][s][n=r]
  const T *Wt;  // transposed           [slab][k=r][s][n=a]
  T *F;         // forward environments  [slot][B][chip]
  T *G;         // adjoint environments  [bond][B][chip]
  int *exps;    // delta_j[b]            [L][B]
  int *esum;    // sum of deltas         [2][B]  (left half, right half)
  T *y;         // rescaled outputs y'   [B][C]
  T *dy;        // dLoss/dy' * loss_scale [B][C]
  T *wq;        // SIMT dA weights [L][D][B]: phi_s[b] * 2^-delta_site[b] (class site: phi_s[b])
  T *loss;      // per-sample loss [B] (may be null)
  double *loss_sum;
  T *out;       // [B][C] copy of y' for the caller (may be null)
  int *out_exp; // [B]
  int stash;    // 1: every bond of F is stored (needed by backward)
  int KT, TN;
  int dmma_sweep, dmma_class;  // byte offsets of the DMMA output tile in the sweep / class kernels' shared memory (0: off)
  int WS;        // shared-memory stride of one (k, s) weight row: chip, or NG*TN*4 for float64 (bank-conflict-free pair layout)
  int vst;       // row stride of the environment tile v_s in shared memory: chip, or chip + 4 when the DMMA step GEMM reads it
                 // (eight rows at a stride of chip * 8 bytes fall on the same banks: an 8-way conflict on every A fragment)
  int wsz;       // elements of the weight staging region W_s (two chunk buffers)
  T loss_scale;
  int seed_only;  // adjoint launch: write the seeds G[c], G[c+1] (and dy) only; the chains run elsewhere
};

template <typename T> __device__ __forceinline__ int mps_slab(const MpsArgs<T> &A, int site) {
  return site <= A.c ? site : site + A.C - 1;
}
template <typename T> __device__ __forceinline__ T *mps_fenv(const MpsArgs<T> &A, int bond) {
  int slot = A.stash ? bond : (bond == A.c ? 0 : (bond == A.c + 1 ? 1 : -1));
  return slot < 0 ? nullptr : A.F + (size_t)slot * A.B * A.chip;
}

template <typename T, int RN> __device__ __forceinline__ void lds_vec(const T *p, T (&w)[RN]);
template <> __device__ __forceinline__ void lds_vec<float, 4>(const float *p, float (&w)[4]) {
  float4 v = *reinterpret_cast<const float4 *>(p);
  w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}
template <> __device__ __forceinline__ void lds_vec<double, 4>(const double *p, double (&w)[4]) {
  double2 v0 = *reinterpret_cast<const double2 *>(p);
  double2 v1 = *reinterpret_cast<const double2 *>(p + 2);
  w[0] = v0.x; w[1] = v0.y; w[2] = v1.x; w[3] = v1.y;
}

// two consecutive elements with one load (16 bytes for double; p is 16-byte aligned)
template <typename T> __device__ __forceinline__ void lds_pair(const T *p, T &a, T &b) { a = p[0]; b = p[1]; }
template <> __device__ __forceinline__ void lds_pair<double>(const double *p, double &a, double &b) {
  const double2 v = *reinterpret_cast<const double2 *>(p);
  a = v.x; b = v.y;
}

// kRN (consts.h): register tile of RB rows (template parameter, 4 or 8) x kRN columns per thread

// float64, one or two column groups, 256 threads: the step GEMM on the FP64 tensor pipe (`mma.sync.m8n8k4.f64`).
// out[BT, D*WS] = v[BT, K] x W[K, D*WS] as 8x8 tiles; a warp owns 4 row tiles x 2 position tiles x TWO physical indices at a time
// (16 MMAs per k-step of 4 for 4 + 4 operand loads; the FMA loop issued 128 FMAs and 32 loads per thread for the same
// work).  D = 2 (D2): one pass.  Other D: one pass per pair of physical indices (each streams only its own two (k, s)
// segments of the weight rows, so the weights are still read once), the phi-weighted sum carried in registers between passes.
// Columns are the POSITIONS of the pair layout (see step_gemm); phi is applied when the tile is written to `out_s` in
// natural column order, from where every thread picks up its (tb, tn) register tile as before.
template <int NG, int RB, bool ZERO, bool D2>
__device__ __forceinline__ void step_gemm_dmma(const double *__restrict__ Wg, int K, int D, int chip, int KT, const double *v_s,
                                               int vst, const double *phi_s, const double *rs_s, int rs_stride, double *W_s,
                                               double *out_s, double (&acc)[NG][RB][kRN], int tb, int tn, int TN, int WS) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int ncg = TN >> 2, wc = warp % ncg, wr = warp / ncg;
  const int R0 = wr * 32, P0 = wc * 16;
  const int GW = TN * 4;          // positions of one column group
  // shared-memory stride of one (k, s) segment: + 4 elements, so that the four k-rows a B fragment reads (2 segments apart)
  // start 64 bytes apart instead of on the same banks; the output tile's rows are padded the same way
  const int GWp = vst > chip ? GW + 4 : GW;
  const int ost = vst;            // row stride of out_s
  const int srow = 2 * GWp;       // one k-row of one group and pass in shared memory: [2 physical indices][GWp]
  const int grow = D * WS;        // one k-row in global memory: [s][NG groups][GW]
  const int nchunks = (K + KT - 1) / KT;
  const int pieces = GW / 2;      // 16-byte pieces of one (k, s) segment (a power of two: TN is)
  const int pshift = __ffs(pieces) - 1;
  // column groups one after the other (chi > 128 has two): each streams its own part of the weight rows
  for (int grp = 0; grp < NG; ++grp) {
    double o[4][2][2];  // phi-weighted sum over the passes (D != 2)
    if constexpr (!D2) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int p = 0; p < 2; ++p) o[r][p][0] = o[r][p][1] = 0.0;
    }
    for (int sp = 0; sp < (D2 ? 2 : D); sp += 2) {
      const bool two = D2 || sp + 1 < D;  // (odd D: the last pass has one physical index; its second segment is not loaded)
      auto issue = [&](int ch) {
        const int k0 = ch * KT;
        const int kt = min(KT, K - k0);
        const int n16 = kt * 2 * pieces;
        const double *src = Wg + (size_t)k0 * grow + (size_t)sp * WS + (size_t)grp * GW;
        double *dst = W_s + (size_t)(ch & 1) * KT * srow;
        for (int i = tid; i < n16; i += NT) {
          const int seg = i >> pshift, q = i & (pieces - 1);  // seg = k * 2 + (s - sp)
          if (!two && (seg & 1)) continue;
          cp_async16(reinterpret_cast<char *>(dst + (size_t)seg * GWp) + q * 16,
                     reinterpret_cast<const char *>(src + (size_t)(seg >> 1) * grow + (size_t)(seg & 1) * WS) + q * 16);
        }
        cp_async_commit();
      };
      double c[4][2][2][2];  // [row tile][position tile][s - sp][2]
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int q = 0; q < 2; ++q) c[r][p][q][0] = c[r][p][q][1] = 0.0;
      issue(0);
      for (int ch = 0; ch < nchunks; ++ch) {
        if (ch + 1 < nchunks) {
          issue(ch + 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();
        const int k0 = ch * KT;
        const int kt = min(KT, K - k0);
        const double *Wb = W_s + (size_t)(ch & 1) * KT * srow + P0 + g;
        const double *va = v_s + (size_t)(R0 + g) * vst + k0 + t;
#pragma unroll 2
        for (int kk = 0; kk < kt; kk += 4) {
          double a[4], b[2][2];
#pragma unroll
          for (int r = 0; r < 4; ++r) a[r] = va[(size_t)(8 * r) * vst + kk];
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int p = 0; p < 2; ++p) b[p][q] = (q == 0 || two) ? Wb[(size_t)((kk + t) * 2 + q) * GWp + 8 * p] : 0.0;
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
              for (int q = 0; q < 2; ++q)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                             : "+d"(c[r][p][q][0]), "+d"(c[r][p][q][1])
                             : "d"(a[r]), "d"(b[p][q]));
        }
        __syncthreads();
      }
      // phi-weighted combination of the pass's physical indices
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = R0 + 8 * r + g;
        const double rsv = rs_s ? rs_s[row * rs_stride] : 1.0;
        const double r0 = rsv * phi_s[row * (D2 ? 2 : D) + sp], r1 = two ? rsv * phi_s[row * (D2 ? 2 : D) + sp + 1] : 0.0;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          if constexpr (D2) {
            o[r][p][0] = fma(r0, c[r][p][0][0], r1 * c[r][p][1][0]);
            o[r][p][1] = fma(r0, c[r][p][0][1], r1 * c[r][p][1][1]);
          } else {
            o[r][p][0] += fma(r0, c[r][p][0][0], r1 * c[r][p][1][0]);
            o[r][p][1] += fma(r0, c[r][p][0][1], r1 * c[r][p][1][1]);
          }
        }
      }
    }
    // -> out_s[row][n]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = R0 + 8 * r + g;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int pos = P0 + 8 * p + 2 * t;  // even: the two values are columns n, n + 1
        const int half = pos / (TN * 2), rem = pos - half * TN * 2;
        const int n = grp * GW + 4 * (rem >> 1) + 2 * half;
        if (n < chip) {
          double2 ov;
          ov.x = o[r][p][0];
          ov.y = o[r][p][1];
          *reinterpret_cast<double2 *>(&out_s[(size_t)row * ost + n]) = ov;
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int grp = 0; grp < NG; ++grp) {
    const int n0 = (grp * TN + tn) * kRN;
    if (n0 < chip) {
#pragma unroll
      for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int rn = 0; rn < kRN; ++rn) {
          const double v = out_s[(size_t)(tb * RB + rb) * ost + n0 + rn];
          acc[grp][rb][rn] = ZERO ? v : acc[grp][rb][rn] + v;
        }
    }
  }
}

template <typename T, int NG, int RB, bool ZERO> struct StepDmma {
  template <typename... Args> static __device__ __forceinline__ void run(Args...) {}
  static constexpr bool kAvail = false;
};
template <int NG, int RB, bool ZERO> struct StepDmma<double, NG, RB, ZERO> {
  static constexpr bool kAvail = NG <= 2;
  static __device__ __forceinline__ void run(const double *Wg, int K, int D, int chip, int KT, const double *v_s, int vst,
                                             const double *phi_s, const double *rs_s, int rs_stride, double *W_s,
                                             double *out_s, double (&acc)[NG][RB][kRN], int tb, int tn, int TN, int WS) {
    if (D == 2)
      step_gemm_dmma<NG, RB, ZERO, true>(Wg, K, 2, chip, KT, v_s, vst, phi_s, rs_s, rs_stride, W_s, out_s, acc, tb, tn, TN, WS);
    else
      step_gemm_dmma<NG, RB, ZERO, false>(Wg, K, D, chip, KT, v_s, vst, phi_s, rs_s, rs_stride, W_s, out_s, acc, tb, tn, TN, WS);
  }
};

// acc[g][rb][rn] += sum_{k<K, s<D} v[row][k] * phi[row][s] * rs[row] * W[k][s][n]
// W streamed from global through a double-buffered cp.async ring of KT k-rows.
// ZERO: acc is known to be zero on entry (lets the d = 2 fast path drop it from the live registers).
template <typename T, int NG, int RB, bool ZERO = false>
__device__ __forceinline__ void step_gemm(const T *__restrict__ Wg, int K, int D, int chip, int KT,
                                          const T *v_s, int vst, const T *phi_s, const T *rs_s, int rs_stride,
                                          T *W_s, T (&acc)[NG][RB][kRN], int tb, int tn, int TN, int WS,
                                          T *out_s = nullptr) {
  if constexpr (StepDmma<T, NG, RB, ZERO>::kAvail) {
    if (out_s != nullptr && D >= 2) {
      // (a group's rows of one pass (two physical indices) are 2 / (NG * D) of the full rows the buffers were sized for: that
      //  many times the k-rows per chunk, in whole k-steps of 4)
      const int ktd = max(4, (KT * NG * D / 2) / 4 * 4);
      StepDmma<T, NG, RB, ZERO>::run(Wg, K, D, chip, ktd, v_s, vst, phi_s, rs_s, rs_stride, W_s, out_s, acc, tb, tn, TN, WS);
      return;
    }
  }
  const int tid = threadIdx.x, NT = blockDim.x;
  const int row_elems = D * WS;  // one k-row, in global and in shared memory
  // float64: a thread's four columns n0..n0+3 are two 16-byte loads; at the natural stride of 32 bytes per thread the
  // lanes of a warp hit only half of the banks.  mps_prep_kernel therefore writes the rows as
  // [group g][pair 0 | pair 1][tn][2] (WS = NG*TN*4 elements): lane tn reads 16 bytes at tn*16 for each pair.
  constexpr bool kPairs = sizeof(T) == 8;
  const int nchunks = (K + KT - 1) / KT;
  auto issue = [&](int ch) {
    int k0 = ch * KT;
    int kt = min(KT, K - k0);
    int n16 = (int)((size_t)kt * row_elems * sizeof(T) / 16);
    const char *src = reinterpret_cast<const char *>(Wg + (size_t)k0 * row_elems);
    char *dst = reinterpret_cast<char *>(W_s + (size_t)(ch & 1) * KT * row_elems);
    for (int i = tid; i < n16; i += NT) cp_async16(dst + (size_t)i * 16, src + (size_t)i * 16);
    cp_async_commit();
  };
  T rs[RB];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) rs[rb] = rs_s ? rs_s[(tb * RB + rb) * rs_stride] : T(1);
  T rp0[RB], rp1[RB];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    rp0[rb] = D == 2 ? rs[rb] * phi_s[(tb * RB + rb) * 2] : T(0);
    rp1[rb] = D == 2 ? rs[rb] * phi_s[(tb * RB + rb) * 2 + 1] : T(0);
  }
  const bool fast = NG == 1 && D == 2;
  T f0[RB][kRN], f1[RB][kRN];  // fast path: per-physical-index accumulators
#pragma unroll
  for (int rb = 0; rb < RB; ++rb)
#pragma unroll
    for (int rn = 0; rn < kRN; ++rn) f0[rb][rn] = f1[rb][rn] = T(0);
  issue(0);
  for (int ch = 0; ch < nchunks; ++ch) {
    if (ch + 1 < nchunks) {
      issue(ch + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int k0 = ch * KT;
    const int kt = min(KT, K - k0);
    const T *Wb = W_s + (size_t)(ch & 1) * KT * row_elems;
    if (fast) {
      // D == 2, one column group: plain v x W products into one accumulator set per physical index (no multiply by phi
      // in the loop; 28 % of the executed instructions were FMAs before), kk unrolled
      if (tn * kRN < chip) {
        const T *vrow = v_s + (size_t)(tb * RB) * vst + k0;
        const T *wp = Wb + (kPairs ? tn * 2 : tn * kRN);
#pragma unroll 2
        for (int kk = 0; kk < kt; ++kk) {
          T vv[RB], w0[kRN], w1[kRN];
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) vv[rb] = vrow[rb * vst + kk];
          const T *p0 = wp + (size_t)(2 * kk) * WS, *p1 = p0 + WS;
          if (kPairs) {
            lds_pair<T>(p0, w0[0], w0[1]);
            lds_pair<T>(p0 + TN * 2, w0[2], w0[3]);
            lds_pair<T>(p1, w1[0], w1[1]);
            lds_pair<T>(p1 + TN * 2, w1[2], w1[3]);
          } else {
            lds_vec<T, kRN>(p0, w0);
            lds_vec<T, kRN>(p1, w1);
          }
#pragma unroll
          for (int rb = 0; rb < RB; ++rb)
#pragma unroll
            for (int rn = 0; rn < kRN; ++rn) {
              f0[rb][rn] = fma(vv[rb], w0[rn], f0[rb][rn]);
              f1[rb][rn] = fma(vv[rb], w1[rn], f1[rb][rn]);
            }
        }
      }
      __syncthreads();
      continue;
    }
    for (int kk = 0; kk < kt; ++kk) {
      T vv[RB];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) vv[rb] = v_s[(tb * RB + rb) * vst + k0 + kk];
      for (int s = 0; s < D; ++s) {
        T a[RB];
        if (D == 2) {  // the common case keeps rs * phi in registers (one shared-memory load less per row and (k, s))
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) a[rb] = vv[rb] * (s == 0 ? rp0[rb] : rp1[rb]);
        } else {
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) a[rb] = vv[rb] * rs[rb] * phi_s[(tb * RB + rb) * D + s];
        }
        const T *wrow = Wb + (size_t)(kk * D + s) * WS;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          int n0 = (g * TN + tn) * kRN;
          if (n0 < chip) {
            T w[kRN];
            if (kPairs) {
              const T *pg = wrow + g * TN * kRN + tn * 2;
              lds_pair<T>(pg, w[0], w[1]);
              lds_pair<T>(pg + TN * 2, w[2], w[3]);
            } else {
              lds_vec<T, kRN>(wrow + n0, w);
            }
#pragma unroll
            for (int rb = 0; rb < RB; ++rb)
#pragma unroll
              for (int rn = 0; rn < kRN; ++rn) acc[g][rb][rn] = fma(a[rb], w[rn], acc[g][rb][rn]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (fast) {
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int rn = 0; rn < kRN; ++rn)
        acc[0][rb][rn] = fma(rp0[rb], f0[rb][rn], ZERO ? rp1[rb] * f1[rb][rn] : fma(rp1[rb], f1[rb][rn], acc[0][rb][rn]));
  }
}

template <typename T, int NG, int RB> __device__ __forceinline__ void zero_acc(T (&acc)[NG][RB][kRN]) {
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int rn = 0; rn < kRN; ++rn) acc[g][rb][rn] = T(0);
}

// phi_s[row][s] for the CTA's rows at one site
template <typename T>
__device__ __forceinline__ void load_phi(const MpsArgs<T> &A, long long b0, int BT, int site, T *phi_s) {
  for (int r = threadIdx.x; r < BT; r += blockDim.x) {
    long long b = b0 + r;
    T ph[kMaxPhys], pad[kMaxPhys];
    if (b >= A.B)
      for (int j = 0; j < A.emb.n_in; ++j) pad[j] = T(0.5);
    embed_site(A.emb, b < A.B ? A.x + ((size_t)b * A.L + site) * A.emb.n_in : pad, ph);
    for (int s = 0; s < A.D; ++s) phi_s[r * A.D + s] = ph[s];
  }
}

// copy rows [b0, b0+BT) of an environment into smem (zero beyond the batch)
template <typename T>
__device__ __forceinline__ void load_env(const T *env, long long b0, long long B, int BT, int chip, T *dst, int dstride) {
  for (int i = threadIdx.x; i < BT * chip; i += blockDim.x) {
    int r = i / chip;
    long long b = b0 + r;
    dst[r * dstride + (i - r * chip)] = b < B ? env[(size_t)b * chip + (i - r * chip)] : T(0);
  }
}

// ---------------------------------------------------------------------------------------
// Environment sweep.  grid = (sample tiles, 2 halves).  adjoint = 0: forward environments
// F (left half: bonds 0..c, right half: bonds L..c+1).  adjoint = 1: seeds G[c] / G[c+1]
// through the class site from dLoss/dy', then the adjoint chains towards the boundaries,
// reusing the forward exponents delta_j.
// ---------------------------------------------------------------------------------------
template <typename T, int NG, int RB>
__global__ void __launch_bounds__(256, (NG == 1 && RB == 4) ? 2 : 1) mps_sweep_kernel(MpsArgs<T> A, int adjoint) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TN = A.TN, TB = blockDim.x / TN, BT = TB * RB;
  const int chip = A.chip, D = A.D, C = A.C, c = A.c, L = A.L;
  const int vst = A.vst;
  T *v_s = reinterpret_cast<T *>(smem_raw);
  T *W_s = v_s + (size_t)BT * vst;
  T *phi_s = W_s + (size_t)A.wsz;
  T *dy_s = phi_s + (size_t)BT * D;
  T *out_s = A.dmma_sweep ? reinterpret_cast<T *>(smem_raw + A.dmma_sweep) : nullptr;  // DMMA output tile (float64)
  const int tid = threadIdx.x;
  const int tn = tid % TN, tb = tid / TN;
  const int half = blockIdx.y;
  const long long b0 = (long long)blockIdx.x * BT;

  int site0, nsites, dir;
  bool transposed;
  if (!adjoint) {
    if (half == 0) { site0 = 0; nsites = c; dir = 1; transposed = false; }
    else { site0 = L - 1; nsites = L - 1 - c; dir = -1; transposed = true; }
  } else {
    if (half == 0) { if (c == 0) return; site0 = c - 1; nsites = c - 1; dir = -1; transposed = true; }
    else { if (c == L - 1) return; site0 = c + 1; nsites = L - 2 - c; dir = 1; transposed = false; }
    if (A.seed_only) nsites = 0;
  }

  T acc[NG][RB][kRN];
  int esum_r[RB];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) esum_r[rb] = 0;

  auto store_env = [&](T *env, const T(&sc)[RB]) {
    // acc * sc -> v_s and (optionally) global env
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      int n0 = (g * TN + tn) * kRN;
      if (n0 < chip) {
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          int r = tb * RB + rb;
          long long b = b0 + r;
#pragma unroll
          for (int rn = 0; rn < kRN; ++rn) {
            T val = acc[g][rb][rn] * sc[rb];
            v_s[r * vst + n0 + rn] = val;
            if (env && b < A.B) env[(size_t)b * chip + n0 + rn] = val;
          }
        }
      }
    }
  };

  if (!adjoint) {
    // boundary vector e_0 (bond dimension 1 zero-padded to chip)
    T *env = mps_fenv(A, half == 0 ? 0 : L);
    for (int i = tid; i < BT * chip; i += blockDim.x) {
      int r = i / chip, n = i - r * chip;
      T val = n == 0 ? T(1) : T(0);
      v_s[r * vst + n] = val;
      long long b = b0 + r;
      if (env && b < A.B) env[(size_t)b * chip + n] = val;
    }
  } else {
    // seed through the class site:  G[c] = sum_k dy_k M_c^k F[c+1]   /   G[c+1] = sum_k dy_k F[c] M_c^k
    for (int r = tid; r < BT; r += blockDim.x) {
      long long b = b0 + r;
      T dyl[kMaxOut];
      if (b < A.B) {
        T yl[kMaxOut], tl[kMaxOut];
        for (int k = 0; k < C; ++k) {
          yl[k] = A.y[(size_t)b * C + k];
          tl[k] = A.targets ? A.targets[(size_t)b * C + k] : T(0);
        }
        int E = A.esum[b] + A.esum[A.B + b];
        mps_head(A.head, C, yl, E, tl, A.cw, dyl);
        for (int k = 0; k < C; ++k) {
          dyl[k] *= A.loss_scale;
          A.dy[(size_t)b * C + k] = dyl[k];
        }
      } else {
        for (int k = 0; k < C; ++k) dyl[k] = T(0);
      }
      for (int k = 0; k < C; ++k) dy_s[r * C + k] = dyl[k];
    }
    load_env(mps_fenv(A, half == 0 ? c + 1 : c), b0, A.B, BT, chip, v_s, vst);
    load_phi(A, b0, BT, c, phi_s);
    __syncthreads();
    zero_acc<T, NG, RB>(acc);
    const T *Wc = (half == 0 ? A.Wt : A.Wn) + (size_t)c * chip * D * A.WS;
    for (int k = 0; k < C; ++k)
      step_gemm<T, NG, RB>(Wc + (size_t)k * chip * D * A.WS, chip, D, chip, A.KT, v_s, vst, phi_s, dy_s + k, C, W_s,
                       acc, tb, tn, TN, A.WS, out_s);
    T one[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) one[rb] = T(1);
    store_env(A.G + (size_t)(half == 0 ? c : c + 1) * A.B * chip, one);
  }
  __syncthreads();

  for (int step = 0; step < nsites; ++step) {
    const int site = site0 + step * dir;
    const int bond_out = transposed ? site : site + 1;
    load_phi(A, b0, BT, site, phi_s);
    __syncthreads();
    zero_acc<T, NG, RB>(acc);
    const T *W = (transposed ? A.Wt : A.Wn) + (size_t)mps_slab(A, site) * chip * D * A.WS;
    step_gemm<T, NG, RB, true>(W, chip, D, chip, A.KT, v_s, vst, phi_s, (const T *)nullptr, 0, W_s, acc, tb, tn, TN, A.WS, out_s);
    T sc[RB];
    if (!adjoint) {
      T m[RB];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        m[rb] = T(0);
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int rn = 0; rn < kRN; ++rn) m[rb] = Num<T>::max(m[rb], Num<T>::abs(acc[g][rb][rn]));
        for (int off = TN >> 1; off > 0; off >>= 1)
          m[rb] = Num<T>::max(m[rb], __shfl_xor_sync(0xffffffffu, m[rb], off));
        int e = 0;
        if (m[rb] > T(0)) Num<T>::frexp(m[rb], &e);
        sc[rb] = Num<T>::ldexp(T(1), -e);
        esum_r[rb] += e;
        long long b = b0 + tb * RB + rb;
        if (tn == 0 && b < A.B) A.exps[(size_t)site * A.B + b] = e;
      }
      store_env(mps_fenv(A, bond_out), sc);
    } else {
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        long long b = b0 + tb * RB + rb;
        int e = b < A.B ? A.exps[(size_t)site * A.B + b] : 0;
        sc[rb] = Num<T>::ldexp(T(1), -e);
      }
      store_env(A.G + (size_t)bond_out * A.B * chip, sc);
    }
    __syncthreads();
  }
  if (!adjoint && tn == 0) {
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      long long b = b0 + tb * RB + rb;
      if (b < A.B) A.esum[(size_t)half * A.B + b] = esum_r[rb];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Class site + head:  y'[b,k] = F[c][b,:] . M_c^k[b] . F[c+1][b,:]   then the loss head.
// ---------------------------------------------------------------------------------------
template <typename T, int NG, int RB>
__global__ void __launch_bounds__(256, (NG == 1 && RB == 4) ? 2 : 1) mps_class_kernel(MpsArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TN = A.TN, TB = blockDim.x / TN, BT = TB * RB;
  const int chip = A.chip, D = A.D, C = A.C, c = A.c;
  const int vst = A.vst;
  T *v_s = reinterpret_cast<T *>(smem_raw);
  T *W_s = v_s + (size_t)BT * vst;
  T *phi_s = W_s + (size_t)A.wsz;
  T *y_s = phi_s + (size_t)BT * D;
  T *fr_s = y_s + (size_t)BT * C;
  T *out_s = A.dmma_class ? reinterpret_cast<T *>(smem_raw + A.dmma_class) : nullptr;
  __shared__ double red_s[8];
  const int tid = threadIdx.x;
  const int tn = tid % TN, tb = tid / TN;
  const long long b0 = (long long)blockIdx.x * BT;

  load_env(mps_fenv(A, c), b0, A.B, BT, chip, v_s, vst);
  load_env(mps_fenv(A, c + 1), b0, A.B, BT, chip, fr_s, chip);
  load_phi(A, b0, BT, c, phi_s);
  __syncthreads();
  T acc[NG][RB][kRN];
  const T *Wc = A.Wn + (size_t)c * chip * D * A.WS;
  for (int k = 0; k < C; ++k) {
    zero_acc<T, NG, RB>(acc);
    step_gemm<T, NG, RB, true>(Wc + (size_t)k * chip * D * A.WS, chip, D, chip, A.KT, v_s, vst, phi_s, (const T *)nullptr, 0,
                     W_s, acc, tb, tn, TN, A.WS, out_s);
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      int r = tb * RB + rb;
      T part = T(0);
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        int n0 = (g * TN + tn) * kRN;
        if (n0 < chip) {
#pragma unroll
          for (int rn = 0; rn < kRN; ++rn) part = fma(acc[g][rb][rn], fr_s[r * chip + n0 + rn], part);
        }
      }
      for (int off = TN >> 1; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if (tn == 0) y_s[r * C + k] = part;
    }
  }
  __syncthreads();
  double lsum = 0.0;
  for (int r = tid; r < BT; r += blockDim.x) {
    long long b = b0 + r;
    if (b < A.B) {
      T yl[kMaxOut], tl[kMaxOut];
      for (int k = 0; k < C; ++k) {
        yl[k] = y_s[r * C + k];
        tl[k] = A.targets ? A.targets[(size_t)b * C + k] : T(0);
        A.y[(size_t)b * C + k] = yl[k];
        if (A.out) A.out[(size_t)b * C + k] = yl[k];
      }
      int E = A.esum[b] + A.esum[A.B + b];
      if (A.out_exp) A.out_exp[b] = E;
      T l = mps_head(A.head, C, yl, E, tl, A.cw, (T *)nullptr);
      if (A.loss) A.loss[b] = l;
      lsum += (double)l;
    }
  }
  if (A.loss_sum) {
    for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
    if ((tid & 31) == 0) red_s[tid >> 5] = lsum;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) tot += red_s[w];
      atomicAdd(A.loss_sum, tot);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Weight gradients (K4):  dA_j[a,r,s(,k)] += sum_b w_s[b] U[b,a] V[b,r],   w_s[b] = phi_j[b,s] * w[b]
//   j < c : U = F[j],  V = G[j+1], w = 2^-delta_j        j > c : U = G[j], V = F[j+1], w = 2^-delta_j
//   j = c : U = F[c],  V = F[c+1], w = dy'[b,k]
// One GEMM per site with the rows m = (s, a) flattened (M = D * chi), K = batch: the V tile is read once for all s and
// no tile row is wasted at small bonds.  grid = (tiles_m * tiles_r * splits, L, n_j); 16x16 threads, RM x 4 micro-tile,
// (16 RM) x 64 output tile.  The next 16 samples are fetched into registers while the current ones are multiplied
// (the loads of the first version were synchronous, and its D copies re-read U and V).
// ---------------------------------------------------------------------------------------
template <typename T> struct DaArgs {
  MpsArgs<T> M;
  T *grads;  // packed layout, chi (unpadded) strides
  int tiles_m, tiles_r, nsplit;
  long long bchunk;  // samples per split
  int only_site;     // >= 0: this site only (the class site); < 0: site = site0 + blockIdx.y, class site skipped
  int site0;         // first site of this launch (the backward pass may run in site ranges)
};

// per-sample weights of the dA GEMM, one thread per (site, sample) (the embedding is evaluated once, not once per tile)
template <typename T> __global__ void mps_wq_kernel(MpsArgs<T> A) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)A.L * A.B) return;
  const int site = (int)(i / A.B);
  const long long b = i - (long long)site * A.B;
  T ph[kMaxPhys];
  embed_site(A.emb, A.x + ((size_t)b * A.L + site) * A.emb.n_in, ph);
  const T f = site == A.c ? T(1) : Num<T>::ldexp(T(1), -A.exps[(size_t)site * A.B + b]);
  for (int s = 0; s < A.D; ++s) A.wq[((size_t)site * A.D + s) * A.B + b] = ph[s] * f;
}

template <typename T, int RM> __global__ void __launch_bounds__(256) mps_da_kernel(DaArgs<T> P) {
  const MpsArgs<T> &A = P.M;
  constexpr int TM = 16 * RM, TR = 64, BK = 16;
  constexpr int UE = BK * TM / 256, VE = BK * TR / 256;  // staged elements per thread and chunk
  constexpr int UROWS = 256 / TM, VROWS = 256 / TR;      // samples covered by one pass of the CTA
  // rows padded by 8 elements: the four k-rows a DMMA fragment load touches (lanes wt = 0 .. 3) would otherwise start on the
  // same banks (row strides of 1024 / 512 bytes): a 4-way conflict on every operand load of the float64 path
  __shared__ __align__(16) T U_s[BK][TM + 8];
  __shared__ __align__(16) T V_s[BK][TR + 8];
  const int site = P.only_site >= 0 ? P.only_site : P.site0 + (int)blockIdx.y;
  const int c = A.c, chip = A.chip, chi = A.chi, D = A.D, C = A.C;
  if (P.only_site < 0 && site == c) return;
  const bool cls = site == c;
  const int nj = cls ? C : 1;
  const int kcls = blockIdx.z;
  int bx = blockIdx.x;
  const int split = bx % P.nsplit;
  bx /= P.nsplit;
  const int tr = bx % P.tiles_r, tm = bx / P.tiles_r;
  const int m0 = tm * TM, r0 = tr * TR, Mrows = D * chi;
  const T *U = site <= c ? mps_fenv(A, site) : A.G + (size_t)site * A.B * chip;
  const T *V = site < c ? A.G + (size_t)(site + 1) * A.B * chip : mps_fenv(A, site + 1);
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  // staging coordinates: a thread always stages the same U row m = (s, a) and the same V column
  const int ucol = tid % TM, ukk = tid / TM, vcol = tid % TR, vkk = tid / TR;
  const int um = m0 + ucol;
  const bool uok = um < Mrows, vok = r0 + vcol < chi;
  const int us = uok ? um / chi : 0, ua = uok ? um - us * chi : 0;
  // float64: a thread's four V columns are two 16-byte loads; stored as [pair 0 | pair 1][tx][2] the 16 lanes read
  // consecutive 16-byte pieces (at the natural 32-byte stride they fall on half of the banks)
  // float64 with the 128 x 64 tile: the product runs as DMMA warp tiles (mma.sync.m8n8k4.f64, a warp owns 4 x 4 of the
  // 8x8 output tiles: 16 MMAs per k-step of 4 for 8 operand loads) on the natural V layout
  constexpr bool kDmma = sizeof(T) == 8 && RM == 8;
  const int vpos = (sizeof(T) == 8 && !kDmma) ? ((vcol & 3) >> 1) * 32 + (vcol >> 2) * 2 + (vcol & 1) : vcol;
  const int lane = tid & 31, wg = lane >> 2, wt = lane & 3;
  const int M0 = ((tid >> 5) & 3) * 32, N0 = (tid >> 7) * 32;  // this warp's 32 x 32 part of the tile
  double cd[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cd[i][j][0] = cd[i][j][1] = 0.0;
  const T *Wq = A.wq + ((size_t)site * D + us) * A.B;
  T acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  const long long bbeg = (long long)split * P.bchunk;
  const long long bend = min(A.B, bbeg + P.bchunk);
  T ureg[UE], wreg[UE], vreg[VE];
  auto fetch = [&](long long bb) {
#pragma unroll
    for (int e = 0; e < UE; ++e) {
      const long long b = bb + e * UROWS + ukk;
      const bool ok = uok && b < bend;
      ureg[e] = ok ? U[(size_t)b * chip + ua] : T(0);
      wreg[e] = ok ? Wq[b] : T(0);
      if (cls && ok) wreg[e] *= A.dy[(size_t)b * C + kcls];
    }
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      const long long b = bb + e * VROWS + vkk;
      vreg[e] = (vok && b < bend) ? V[(size_t)b * chip + r0 + vcol] : T(0);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int e = 0; e < UE; ++e) U_s[e * UROWS + ukk][ucol] = ureg[e] * wreg[e];
#pragma unroll
    for (int e = 0; e < VE; ++e) V_s[e * VROWS + vkk][vpos] = vreg[e];
  };
  if (bbeg < bend) {
    fetch(bbeg);
    stage();
    __syncthreads();
  }
  for (long long bb = bbeg; bb < bend; bb += BK) {
    const bool more = bb + BK < bend;
    if (more) fetch(bb + BK);
    if constexpr (kDmma) {
      // (a warp whose 32 x 32 part of the tile lies outside the matrix -- bonds of 32 and less fill only the first column half,
      //  short rows only some row quarters -- issues no MMAs: it would only take the FP64 tensor pipe away from its neighbours)
      if (m0 + M0 < Mrows && r0 + N0 < chi)
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = (double)U_s[kk + wt][M0 + 8 * i + wg];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = (double)V_s[kk + wt][N0 + 8 * j + wg];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                         : "+d"(cd[i][j][0]), "+d"(cd[i][j][1])
                         : "d"(a[i]), "d"(b[j]));
      }
    } else {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T u[RM], v[4];
#pragma unroll
      for (int i = 0; i < RM / 4; ++i) {
        T t4[4];
        lds_vec<T, 4>(&U_s[kk][ty * RM + 4 * i], t4);
#pragma unroll
        for (int q = 0; q < 4; ++q) u[4 * i + q] = t4[q];
      }
      if (sizeof(T) == 8) {
        lds_pair<T>(&V_s[kk][tx * 2], v[0], v[1]);
        lds_pair<T>(&V_s[kk][32 + tx * 2], v[2], v[3]);
      } else {
        lds_vec<T, 4>(&V_s[kk][tx * 4], v);
      }
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(u[i], v[j], acc[i][j]);
    }
    }
    __syncthreads();
    if (more) stage();
    __syncthreads();
  }
  // scatter into the packed layout [a][r][s][k]
  const size_t site_off = (size_t)site * chi * chi * D + (site > c ? (size_t)chi * chi * D * (C - 1) : 0);
  T *g = P.grads + site_off;
  if constexpr (kDmma) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + M0 + 8 * i + wg;
      if (m >= Mrows) continue;
      const int s = m / chi, a = m - s * chi;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = r0 + N0 + 8 * j + 2 * wt + e;
          if (r < chi) atomicAdd(&g[(((size_t)a * chi + r) * D + s) * nj + kcls], (T)cd[i][j][e]);
        }
    }
  } else {
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int m = m0 + ty * RM + i;
      if (m >= Mrows) continue;
      const int s = m / chi, a = m - s * chi;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + tx * 4 + j;
        if (r >= chi) continue;
        atomicAdd(&g[(((size_t)a * chi + r) * D + s) * nj + kcls], acc[i][j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Two-site gradient of the Sweeps strategy (tn4ml/strategy.py:135-176: the pair (site, site + 1) is contracted into one
// tensor T and the loss is differentiated with respect to T):
//     dT[a][r][s][s'][k] = sum_b U[b, a] * V[b, r] * w_s[b] * w'_{s'}[b] * (dy[b, k] if the pair holds the class site)
// from the stash of a forward + adjoint pass: U = F[site] (left environment) or, right of the class site, the adjoint
// G[site]; V = the adjoint G[site + 2] left of the class site, else the right environment F[site + 2]; w, w' = the per-sample
// weights of the two sites (phi_s * 2^-delta, mps_wq_kernel).  One CTA per 16 x 16 tile of (a, r), per (s, s'), per class
// index and per batch split; split-K partial sums are added atomically (dT is zeroed by the caller).
// ---------------------------------------------------------------------------------------
template <typename T> __global__ void __launch_bounds__(256) mps_pair_kernel(MpsArgs<T> A, int site, T *dT, int nsplit, long long bchunk) {
  constexpr int BK = 32;
  __shared__ T U_s[BK][17];
  __shared__ T V_s[BK][17];
  const int chi = A.chi, chip = A.chip, D = A.D, C = A.C, c = A.c;
  const bool cls = site == c || site + 1 == c;
  const int nk = cls ? C : 1;
  const int tiles = (chi + 15) / 16;
  int bx = blockIdx.x;
  const int split = bx % nsplit;
  bx /= nsplit;
  const int tr = bx % tiles, ta = bx / tiles;
  const int s = blockIdx.y / D, sp = blockIdx.y - s * D, k = blockIdx.z;
  const T *U = site > c ? A.G + (size_t)site * A.B * chip : mps_fenv(A, site);
  const T *V = site + 1 < c ? A.G + (size_t)(site + 2) * A.B * chip : mps_fenv(A, site + 2);
  const T *W0 = A.wq + ((size_t)site * D + s) * A.B;
  const T *W1 = A.wq + ((size_t)(site + 1) * D + sp) * A.B;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const long long bbeg = (long long)split * bchunk, bend = min(A.B, bbeg + bchunk);
  T acc = T(0);
  for (long long bb = bbeg; bb < bend; bb += BK) {
    // 256 threads stage 32 samples x 16 columns of U (weighted) and V, two rows each
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = ty + 16 * e;
      const long long b = bb + kk;
      const bool ok = b < bend;
      const int a = ta * 16 + tx, r = tr * 16 + tx;
      T w = T(0);
      if (ok) {
        w = W0[b] * W1[b];
        if (cls) w *= A.dy[(size_t)b * C + k];
      }
      U_s[kk][tx] = ok && a < chi ? U[(size_t)b * chip + a] * w : T(0);
      V_s[kk][tx] = ok && r < chi ? V[(size_t)b * chip + r] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) acc = fma(U_s[kk][ty], V_s[kk][tx], acc);
    __syncthreads();
  }
  const int a = ta * 16 + ty, r = tr * 16 + tx;
  if (a < chi && r < chi) {
    T *dst = &dT[((((size_t)a * chi + r) * D + s) * D + sp) * nk + k];
    if (nsplit > 1) atomicAdd(dst, acc);
    else *dst = acc;
  }
}

// ---------------------------------------------------------------------------------------
// Weight re-layout (once per step; parameters are tiny next to the batch work):
//   packed A_j[a][r][s][o]  ->  Wn[slab][k=a][s][n=r] and Wt[slab][k=r][s][n=a], zero padded to chip.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void mps_prep_kernel(const T *__restrict__ params, T *__restrict__ Wn, T *__restrict__ Wt, int L,
                                int chi, int chip, int D, int C, int c, int WS, int TN) {
  const int nslab = L - 1 + C;
  const size_t slab_elems = (size_t)chip * D * chip;
  const size_t total = slab_elems * nslab;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int slab = (int)(i / slab_elems);
    size_t rem = i - (size_t)slab * slab_elems;
    int k = (int)(rem / ((size_t)D * chip));
    int s = (int)((rem / chip) % D);
    int n = (int)(rem % chip);
    int site, o, nj;
    if (slab < c) { site = slab; o = 0; nj = 1; }
    else if (slab < c + C) { site = c; o = slab - c; nj = C; }
    else { site = slab - C + 1; o = 0; nj = 1; }
    const size_t site_off = (size_t)site * chi * chi * D + (site > c ? (size_t)chi * chi * D * (C - 1) : 0);
    T vn = T(0), vt = T(0);
    if (k < chi && n < chi) {
      vn = params[site_off + (((size_t)k * chi + n) * D + s) * nj + o];
      vt = params[site_off + (((size_t)n * chi + k) * D + s) * nj + o];
    }
    // position of column n inside its row of WS elements (see step_gemm): natural, or the float64 pair layout
    int pos = n;
    if (sizeof(T) == 8) {
      const int q4 = n >> 2, j = n & 3, g = q4 / TN, tn = q4 - g * TN;
      pos = g * TN * 4 + (j >> 1) * TN * 2 + tn * 2 + (j & 1);
    }
    const size_t dst = (((size_t)slab * chip + k) * D + s) * WS + pos;
    Wn[dst] = vn;
    Wt[dst] = vt;
  }
}

template <typename T>
__global__ void embed_kernel(EmbParams<T> e, long long n, const T *__restrict__ x, T *__restrict__ phi) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T ph[kMaxPhys];
  feature_map(e, x + i * e.n_in, ph);
  for (int s = 0; s < e.d; ++s) phi[i * e.d + s] = ph[s];
}

}  // namespace tnb
