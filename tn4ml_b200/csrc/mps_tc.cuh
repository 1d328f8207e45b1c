// K2 (tensor-core family): MPS environment sweep on tcgen05 / TMEM, fp32 in / fp32 out.
//
// One CTA owns 128 samples (one TMEM lane each) and walks a half-chain site by site:
//     acc[b, s*chi + n] = sum_k v[b,k] * W_site[k, s, n]          tcgen05.mma, M=128, K=chi, N=D*chi
//     v'[b, n]          = 2^-e[b] * sum_s phi_site[b,s] * acc[b, s*chi + n]   (epilogue, TMEM -> regs)
// The environment never leaves the SM between sites: the epilogue writes v' straight back into
// shared memory as the next A operand (and streams a copy to the HBM stash for the VJP).
//
// Arithmetic: "fp16x3".  Every fp32 operand x is split exactly into x = hi + lo (two fp16 after
// a power-of-two pre-scale, 22 significant bits) and hi*hi + lo*hi + hi*lo is accumulated in
// fp32 by three kind::f16 MMAs.  Measured against the fp64 oracle this is as accurate as
// fp32 FMA arithmetic ("precision of the tensor path"), at 1.5x the cost of a
// single TF32 pass, where a single TF32 pass misses the 1e-4 parity bar by two orders.
//
// Weights are pre-split once per step into per-site "stage images": byte-exact copies of the
// shared-memory operand layout (canonical K-major, no swizzle), so one 1-D bulk async copy
// (TMA engine) moves a whole pipeline stage and completes on an mbarrier.
//
// Warp roles of the chain kernel (608 threads; mps_tc_chain.cuh): warps 0-15 epilogue (four threads per
// sample row == TMEM lane), warp 16 TMEM owner + MMA issuer, warp 17 bulk-copy producer (weights in),
// warp 18 bulk-copy image writer (environments out).
#pragma once
#include <cuda_fp16.h>
#include <limits.h>

#include "mps_simt.cuh"
#include "ptx_sm100.cuh"

namespace tnb {

// geometry constants (kTcRows, kTcKA, kTcEpiWarps, kTcMmaWarp, kTcThreads, kTcMaxNH): consts.h

struct TcArgs {
  MpsArgs<float> M;
  const uint8_t *img_n;  // stage images, normal direction      [slab][NH chunks][chi/16][2 planes][NMMA*32 B]
  const uint8_t *img_t;  // stage images, transposed direction
  const int *kW;         // per-slab weight pre-scale exponent
  const float *colsum;   // [2 directions][slab]: max_n sum_{k,s} |W[k,s,n]| * 2^kW  (bound of one chain step, see chain kernel)
  int pace_ns;           // image writer: pause between 4 KB pieces
  int cg;                // CTAs per MMA: 1, or 2 (CTA pair; stage_bytes is then one CTA's half of a stage)
#ifdef TNB_TC_TRACE
  int dbg;               // timing-experiment switches (TNB_TC_DBG): -DTNB_TC_TRACE builds only, absent from the release library
#endif
  long long tile0;       // first sample of this launch (a sweep may be split into a two-tile and a one-tile launch)
  float *wq;             // adjoint: phi_s[b] * 2^q[b] per (site, s, b): per-sample weight of the dA GEMM
  int *qmax;             // adjoint: max_b q[site][b]
  // Environment stash of the tensor path: fp16 hi/lo images (value * 2^14, rows normalised), one per bond, laid out
  // [slot][sample block of 8][plane hi|lo][chi/8 octets][8 samples][8 halves]: the 16-byte unit is 8 consecutive bond
  // indices of one sample.  A tile of 128 samples is byte-identical to the chain kernel's shared-memory A operand
  // (K-major, LBO 128, SBO chi*32); a run of 32 samples is an MN-major operand
  // tile of the dA GEMM (K-block stride chi*32, octet stride 128).  fimg: forward environments, gimg: adjoints.
  uint8_t *fimg, *gimg;
  size_t img_bond_bytes;
  int img_stash;         // 1: every bond has a slot; 0: forward-only, a single slot for bond c+1
  float *wqc;            // adjoint: phi_s[b] * dLoss/dy'_k[b] per ((s,k), b): weights of the class-site dA GEMM
  int *qmaxc;            // adjoint: max exponent of wqc
  int N, NH, NMMA, nstage, tmem_cols;  // N = D*chi accumulator columns = NH chunks of NMMA (n in [h*chi/NH, ..) x all s)
  uint32_t stage_bytes;
  size_t site_img_bytes;
};

// ---- weight images ----------------------------------------------------------------------
// Per-slab statistics of the site tensors, one pass per step (the weights change with every optimizer update):
//   max |W|                        -> kW[slab] = 14 - exponent(max |W|), so that max |W * 2^kW| is in [2^13, 2^14)
//   max_n sum_{k,s} |W[k,s,n]|     -> colsum[dir][slab] (dir 0: n = right bond, dir 1: n = left bond): with |phi_s| <= 1 and an
//                                     operand row of max-abs m, every output of the chain step through this slab is bounded
//                                     by m * colsum, which fixes the step's power-of-two scale BEFORE the step runs (no row
//                                     maximum on the critical path).
// grid (slab, kTcStatGroups): block g takes the rows a of its quarter (row sums: one warp per row, lanes along the contiguous
// (r, s) run) and the columns r of its quarter (column sums: threads along r, phases along a); the three maxima are
// non-negative floats combined with atomicMax on their bit patterns (kW and colsum hold raw bits until tc_finalize_kernel).
constexpr int kTcStatGroups = 4;
static __global__ void __launch_bounds__(256) tc_stats_kernel(const float *__restrict__ params, int *__restrict__ kW,
                                                       int *__restrict__ colsum_bits, int L, int chi, int D, int C, int c) {
  const int slab = blockIdx.x, g = blockIdx.y, nslab = L - 1 + C;
  int site, o, nj;
  if (slab < c) { site = slab; o = 0; nj = 1; }
  else if (slab < c + C) { site = c; o = slab - c; nj = C; }
  else { site = slab - C + 1; o = 0; nj = 1; }
  const float *W = params + (size_t)site * chi * chi * D + (site > c ? (size_t)chi * chi * D * (C - 1) : 0) + o;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = (chi + kTcStatGroups - 1) / kTcStatGroups;
  const int lo = g * q, hi = min(chi, lo + q);
  float wmax = 0.f, rowbest = 0.f, colbest = 0.f;
  // rows a in [lo, hi): sum over (r, s)
  const int run = chi * D;
  for (int a = lo + warp; a < hi; a += 8) {
    const float *row = W + (size_t)a * run * nj;
    float acc = 0.f;
    for (int i = lane; i < run; i += 32) {
      const float v = fabsf(row[(size_t)i * nj]);
      acc += v;
      wmax = fmaxf(wmax, v);
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    rowbest = fmaxf(rowbest, acc);
  }
  // columns r in [lo, hi): sum over (a, s); thread = (column, phase of a)
  __shared__ float part[256];
  const int ncol = hi - lo;
  if (ncol > 0) {
    const int nph = max(1, 256 / ncol);
    for (int c0 = 0; c0 < ncol; c0 += 256) {  // (chi / 4 <= 256 columns: one trip)
      const int col = c0 + tid % min(ncol, 256), ph = tid / min(ncol, 256);
      float acc = 0.f;
      if (col < ncol && ph < nph)
        for (int a = ph; a < chi; a += nph) {
          const float *w = W + (((size_t)a * chi + lo + col) * D) * nj;
          for (int s2 = 0; s2 < D; ++s2) acc += fabsf(w[(size_t)s2 * nj]);
        }
      part[tid] = acc;
      __syncthreads();
      if (ph == 0 && col < ncol) {
        for (int p2 = 1; p2 < nph; ++p2) acc += part[p2 * min(ncol, 256) + (col - c0)];
        colbest = fmaxf(colbest, acc);
      }
      __syncthreads();
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
    rowbest = fmaxf(rowbest, __shfl_xor_sync(0xffffffffu, rowbest, off));
    colbest = fmaxf(colbest, __shfl_xor_sync(0xffffffffu, colbest, off));
  }
  if (lane == 0) {
    atomicMax(&kW[slab], __float_as_int(wmax));
    atomicMax(&colsum_bits[slab], __float_as_int(colbest));          // dir 0: n = right bond
    atomicMax(&colsum_bits[nslab + slab], __float_as_int(rowbest));  // dir 1: n = left bond
  }
}

// raw maxima -> kW (class slabs share one scale: the seed phase accumulates all of them into one TMEM tile, so kW <- the min
// over the C class slabs) and colsum * 2^kW with a margin for the fp16x3 rounding of the products.  One block.
static __global__ void tc_finalize_kernel(int *__restrict__ kW, float *__restrict__ colsum, int L, int C, int c) {
  const int nslab = L - 1 + C;
  __shared__ int kmin;
  if (threadIdx.x == 0) kmin = INT_MAX;
  __syncthreads();
  for (int slab = threadIdx.x; slab < nslab; slab += blockDim.x) {
    const float m = __int_as_float(kW[slab]);
    int e = 0;
    if (m > 0.f) frexpf(m, &e);
    const int k = kTcKA - e;
    kW[slab] = k;
    if (C > 1 && slab >= c && slab < c + C) atomicMin(&kmin, k);
  }
  __syncthreads();
  for (int slab = threadIdx.x; slab < nslab; slab += blockDim.x) {
    int k = kW[slab];
    if (C > 1 && slab >= c && slab < c + C) { k = kmin; kW[slab] = k; }
    for (int dir = 0; dir < 2; ++dir) colsum[dir * nslab + slab] = ldexpf(colsum[dir * nslab + slab], k) * 1.001f;
  }
}

// one thread per (slab, n', k-octet): 8 weights -> 16 B of hi and 16 B of lo, both directions.
// Stage (h, ks) of a slab holds the K-slab ks of column chunk h: columns n'' = s*chi_h + (n - h*chi_h).
// cg = 2 (CTA pair): a stage is [rank][plane hi|lo][NMMA / 2 columns], rank r holding the columns the r-th CTA supplies.
static __global__ void tc_image_kernel(const float *__restrict__ params, const int *__restrict__ kW, uint8_t *__restrict__ img_n,
                                uint8_t *__restrict__ img_t, int L, int chi_w, int chi, int D, int C, int c, int NH, int cg) {
  // chi_w: bond of the caller's [chi_w][chi_w][D] slabs; chi >= chi_w: bond of the images (entries beyond chi_w are zero)
  const int nslab = L - 1 + C;
  const int N = D * chi, KO = chi / 8, KS = chi / 16;
  const int chi_h = chi / NH, NMMA = N / NH;
  const size_t per_slab = (size_t)N * KO;
  const size_t total = per_slab * nslab;
  const uint32_t stage_bytes = 64u * NMMA;
  const size_t site_img_bytes = (size_t)NH * KS * stage_bytes;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int slab = (int)(i / per_slab);
    const size_t rem = i - (size_t)slab * per_slab;
    const int ko = (int)(rem / N);       // k octet: k = ko*8 .. ko*8+7
    const int np = (int)(rem % N);       // n' = s*chi + n
    const int s = np / chi, n = np - s * chi;
    const int h = n / chi_h, npp = s * chi_h + (n - h * chi_h);
    int site, o, nj;
    if (slab < c) { site = slab; o = 0; nj = 1; }
    else if (slab < c + C) { site = c; o = slab - c; nj = C; }
    else { site = slab - C + 1; o = 0; nj = 1; }
    const size_t site_off = (size_t)site * chi_w * chi_w * D + (site > c ? (size_t)chi_w * chi_w * D * (C - 1) : 0);
    const float sc = ldexpf(1.f, kW[slab]);
    const int ks = ko / 2, kc = ko & 1;
    const int ncol = NMMA / cg;  // columns per CTA
    const int rk = npp / ncol, nq = npp - rk * ncol;
    const size_t off = (size_t)slab * site_img_bytes + ((size_t)h * KS + ks) * stage_bytes + (size_t)rk * ncol * 64 +
                       (size_t)(nq / 8) * 256 + kc * 128 + (nq % 8) * 16;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      __half hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = ko * 8 + j;
        const int a = dir == 0 ? k : n, r = dir == 0 ? n : k;
        const float w = (a < chi_w && r < chi_w) ? params[site_off + (((size_t)a * chi_w + r) * D + s) * nj + o] * sc : 0.f;
        hi[j] = __float2half_rn(w);
        lo[j] = __float2half_rn(w - __half2float(hi[j]));
      }
      uint8_t *base = (dir == 0 ? img_n : img_t) + off;
      *reinterpret_cast<uint4 *>(base) = *reinterpret_cast<const uint4 *>(hi);
      *reinterpret_cast<uint4 *>(base + (size_t)ncol * 32) = *reinterpret_cast<const uint4 *>(lo);
    }
  }
}

// byte offset of sample b's 16-byte units inside a bond image (add plane * chi*16 + octet * 128)
__device__ __forceinline__ size_t tc_img_row(long long b, int chi) { return (size_t)(b >> 3) * chi * 32 + (size_t)(b & 7) * 16; }

// pack 8 floats (already scaled) into fp16 hi / lo octets
__device__ __forceinline__ void split8(const float *w, uint4 &hi, uint4 &lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __half h0 = __float2half_rn(w[2 * j]), h1 = __float2half_rn(w[2 * j + 1]);
    __half l0 = __float2half_rn(w[2 * j] - __half2float(h0)), l1 = __float2half_rn(w[2 * j + 1] - __half2float(h1));
    h[j] = __halves2half2(h0, h1);
    l[j] = __halves2half2(l0, l1);
  }
  hi = *reinterpret_cast<uint4 *>(h);
  lo = *reinterpret_cast<uint4 *>(l);
}

}  // namespace tnb
#include "mps_tc_chain.cuh"  // tc_chain_kernel: chains, class site, head and adjoint seeds
namespace tnb {

// ---------------------------------------------------------------------------------------
// K4 (tensor-core family): weight gradients as a batch-reduction GEMM on tcgen05.
//   dA_site[a, r, sigma] = 2^qmax * sum_b (U[b,a] * wq[site][sigma][b] * 2^-qmax) * V[b,r]
//   rows m = sigma*chi + a, columns r (N = chi), K = samples.
// Both operands are the fp16 hi/lo bond images the chain kernels write; a run of 32 samples of an image is an
// MN-major operand tile as it stands, so both arrive by bulk async copies (deep prefetch, no load latency in any warp).
// V is used as it is.  U carries the per-sample weight: one CTA owns `cpg` weighted copies sigma of `ra` bond rows
// (ra = min(chi, 128), cpg*ra <= 256 rows, two M tiles); the raw rows land in the CTA's first copy, and 16 producer
// warps (lane == sample, warp == row octet) rebuild fp32 from hi/lo, apply each copy's weight, re-split and write the
// copies (the first one in place).  One warp issues the MMAs; warps 0-3 drain TMEM once at the end and reduce into the
// packed gradient with fp32 atomics (split-K across CTAs).  With ra = 128 and two copies, tile t is copy t of the same
// 128 rows: both values of a (row, column) go through a per-warp shared-memory transpose and leave as 128-byte runs of
// the packed layout [a][r][sigma].
// ---------------------------------------------------------------------------------------
struct TcDaArgs {
  MpsArgs<float> M;
  const float *wq;
  const int *qmax;
  float *grads;
  int mgroups, nsplit, nstage, tmem_cols;
  long long bchunk;  // samples per split (multiple of 32)
  const uint8_t *fimg, *gimg;
  size_t img_bond_bytes;
  int Dv;            // weighted copies sigma < Dv: D for chain sites, D*C for the class site
  int ra;            // bond rows per copy and CTA: min(chi, 128)
  int cpg;           // copies per CTA: 256 / ra
  uint32_t stage_bytes;  // one pipeline stage: the U region (kDaU) + the V chunk of this bond (chi * 4 * kDaKC bytes), 1 KB aligned
  int only_site;     // >= 0: this site only (the class site; wq/qmax point at its own tables)
  int site0;         // first site of this launch (site = site0 + blockIdx.y): the backward pass may run in site ranges
};

// stage geometry of the dA GEMM (kDaKC, kDaProducers, kDaThreads, kDaMaxCpg, kDaU*, kDaStage): consts.h

#ifdef TNB_TC_TRACE
#define TNB_DA_TRACE(ch_, ev)                                                                              \
  do {                                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 1 && (ch_) >= 200 && (ch_) < 232 && (threadIdx.x & 31) == 0)      \
      g_tc_trace[((ch_)-200) * 64 + 32 + (ev)] = clock64();                                                \
  } while (0)
#else
#define TNB_DA_TRACE(ch_, ev) do { } while (0)
#endif

// CPG: upper bound of the copies per CTA (compile time, so that the producer loop is straight-line code).
// PW: producer warps.  16 (one CTA of 576 threads per SM) for bonds of 128 and more; 8 / 4 for bonds up to 64 / 32, where a
// CTA has at most 8 / 4 row octets per copy to re-weight and a chunk is a handful of small MMAs whose cost is the issue latency
// of the one MMA warp (~120 cycles each): two CTAs of 320 threads (three of 192) per SM, stages sized by the bond, run that many
// pipelines side by side (chi = 32: 3.99 -> 2.13 -> 1.71 ms; chi = 64: 4.22 -> 2.86 ms; chi = 128: no gain, 5.6 ms either way).
template <int CPG, int PW = 16>
__global__ void __launch_bounds__(PW * 32 + 64, PW == 16 ? 1 : (PW == 8 ? 2 : 3)) tc_da_kernel(const __grid_constant__ TcDaArgs A) {
  constexpr int kThreads = PW * 32 + 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  const MpsArgs<float> &M = A.M;
  const int chi = M.chip, Dv = A.Dv, c = M.c;
  const int site = A.only_site >= 0 ? A.only_site : A.site0 + (int)blockIdx.y;
  if (A.only_site < 0 && site == c) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int group = blockIdx.x % A.mgroups, split = blockIdx.x / A.mgroups;
  const int nag = (chi + A.ra - 1) / A.ra;              // row groups
  const int ag = group % nag, cg = group / nag;
  const int a0 = ag * A.ra;                             // first bond row of this CTA
  const int ra = A.ra, ra_eff = min(ra, chi - a0);      // rows per copy (stride) / rows in use
  const int sig0 = cg * A.cpg;                          // first copy of this CTA
  const int ncopy = min(A.cpg, Dv - sig0);              // copies of this CTA (<= CPG)
  const int nrows = ncopy * ra;                         // operand rows spanned (<= 256)
  const int ntile = (nrows + 127) / 128;
  const int noct = ra_eff / 8;                          // row octets of one copy in use
  const long long kbeg = (long long)split * A.bchunk;
  const long long kend = min(M.B, kbeg + A.bchunk);
  if (kbeg >= kend) return;
  const int nchunks = (int)((kend - kbeg + kDaKC - 1) / kDaKC);
  // both operands come from the bond images written by the chain kernels:
  //   U = F[site] (left half, class site) or G[site] (right half): re-weighted per sample in-kernel
  //   V = G[site+1] (left half) or F[site+1] (right half, class site): used as it is
  const size_t chunk0 = (size_t)(kbeg / 8) * chi * 32;
  const uint8_t *Uimg = (site <= c ? A.fimg : A.gimg) + (size_t)site * A.img_bond_bytes + chunk0;
  const uint8_t *Vimg = (site < c ? A.gimg : A.fimg) + (size_t)(site + 1) * A.img_bond_bytes + chunk0;
  const uint32_t b_bytes = (uint32_t)chi * 4 * kDaKC;  // one chunk (16 samples) of an image

  constexpr int kMmaWarp = PW, kCopyWarp = kMmaWarp + 1;
  const uint32_t kStage = A.stage_bytes;
  uint8_t *stages = smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(stages + (size_t)A.nstage * kStage);
  uint64_t *loaded = bars, *ready = bars + A.nstage, *empty = bars + 2 * A.nstage, *done = bars + 3 * A.nstage;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
  if (tid == 0) {
    for (int i = 0; i < A.nstage; ++i) {
      ptx::mbar_init(&loaded[i], 1);                 // the bulk copies' expect_tx arrival
      ptx::mbar_init(&ready[i], PW);                 // one arrival per producer warp
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)A.tmem_cols);
    ptx::tmem_relinquish();
  }
  // operand rows that no copy fills (partial last tile) must read as zero
  for (int st = 0; st < A.nstage; ++st)
    for (uint32_t i = tid * 16u; i < kDaU; i += kThreads * 16u)
      *reinterpret_cast<uint4 *>(stages + (size_t)st * kStage + i) = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int qmax = A.only_site >= 0 ? A.qmax[0] : A.qmax[site];

  if (warp == kMmaWarp) {
    // the whole warp waits; the instructions go out under elect.sync (no per-instruction ELECT / BRA.U.ANY loop)
    {
      const uint32_t idesc = ptx::idesc_f16_f32(128, chi) | (1u << 15) | (1u << 16);  // A and B MN-major
      // V tile = kDaKC samples of a bond image: sample-block (K) stride chi*32, lo plane at chi*16
      const uint32_t vlbo = (uint32_t)chi * 32, vlo = (uint32_t)chi * 16;
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        ptx::mbar_wait(&ready[stage], phase);
        ptx::tc_fence_after();
        TNB_DA_TRACE(ch, 0);
        if (ptx::elect_one()) {
          const uint32_t st = ptx::smem_u32(stages + (size_t)stage * kStage);
          for (int t = 0; t < ntile; ++t) {
#pragma unroll
            for (int ks = 0; ks < kDaKC / 16; ++ks) {
              const uint64_t da_hi = ptx::smem_desc_kmajor(st + t * 16 * 128 + ks * 2 * kDaUBlock, kDaUBlock, 128);
              const uint64_t da_lo = ptx::smem_desc_kmajor(st + kDaUPlane + t * 16 * 128 + ks * 2 * kDaUBlock, kDaUBlock, 128);
              const uint64_t db_hi = ptx::smem_desc_kmajor(st + kDaU + ks * 2 * vlbo, vlbo, 128);
              const uint64_t db_lo = ptx::smem_desc_kmajor(st + kDaU + vlo + ks * 2 * vlbo, vlbo, 128);
              const uint32_t d = tmem_base + (uint32_t)t * chi;
              ptx::mma_f16(d, da_hi, db_hi, idesc, (ch > 0 || ks > 0) ? 1u : 0u);
              ptx::mma_f16(d, da_lo, db_hi, idesc, 1u);
              ptx::mma_f16(d, da_hi, db_lo, idesc, 1u);
            }
          }
          ptx::mma_commit(&empty[stage]);
        }
        __syncwarp();
        TNB_DA_TRACE(ch, 1);
        if (++stage == A.nstage) { stage = 0; phase ^= 1; }
      }
      if (ptx::elect_one()) ptx::mma_commit(done);
    }
  } else if (warp == kCopyWarp) {
    // ===== both operands: bulk async copies of one image chunk per stage (U: per sample block and plane, into the
    //       rows of the CTA's first copy; V: the whole chunk) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int pf = A.nstage + 2;  // L2 prefetch distance (chunks): the stage ring alone does not cover the HBM latency
      for (int ch = 0; ch < pf && ch < nchunks; ++ch) {
        ptx::bulk_prefetch_l2(Uimg + (size_t)ch * b_bytes, b_bytes);
        ptx::bulk_prefetch_l2(Vimg + (size_t)ch * b_bytes, b_bytes);
      }
      for (int ch = 0; ch < nchunks; ++ch) {
        if (ch + pf < nchunks) {
          ptx::bulk_prefetch_l2(Uimg + (size_t)(ch + pf) * b_bytes, b_bytes);
          ptx::bulk_prefetch_l2(Vimg + (size_t)(ch + pf) * b_bytes, b_bytes);
        }
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        TNB_DA_TRACE(ch, 2);
        uint8_t *st = stages + (size_t)stage * kStage;
        const uint32_t u_bytes = (uint32_t)noct * 128;  // the CTA's rows of one (sample block, plane)
        ptx::mbar_expect_tx(&loaded[stage], 2 * (kDaKC / 8) * u_bytes + b_bytes);
        const uint8_t *usrc = Uimg + (size_t)ch * b_bytes + (size_t)(a0 / 8) * 128;
#pragma unroll
        for (int bp = 0; bp < 2 * (kDaKC / 8); ++bp)  // (sample block, plane)
          ptx::bulk_g2s(st + (uint32_t)bp * kDaUPlane, usrc + (size_t)bp * chi * 16, u_bytes, &loaded[stage]);
        ptx::bulk_g2s(st + kDaU, Vimg + (size_t)ch * b_bytes, b_bytes, &loaded[stage]);
        if (++stage == A.nstage) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== U operand producers: kDaKC lanes == the samples of the chunk, (warp, lane / kDaKC) == row octet of a copy =====
    // unit (sample block, plane, octet o, sample % 8) of the raw chunk sits in the first copy's rows
    constexpr int kOctPerWarp = 32 / kDaKC;
    const int smp = lane & (kDaKC - 1);
    const uint32_t lane_off = (uint32_t)(smp >> 3) * kDaUBlock + (uint32_t)(smp & 7) * 16;
    const int o0 = kOctPerWarp * warp + lane / kDaKC;  // my octets: o0, o0 + 16 * kOctPerWarp, ...
    const float qs = ldexpf(1.f, -qmax);          // (the image already carries the 2^14 operand scale)
    const float *wq0 = A.wq + ((size_t)(A.only_site >= 0 ? 0 : site) * Dv + sig0) * M.B;
    // weights of the next chunk, one per copy: loaded one chunk ahead and not touched until then (an instruction
    // that consumes a loaded value would stall this in-order warp for the whole memory latency)
    float wn[CPG];
    auto load_w = [&](int ch) {
      const long long b_ = kbeg + (long long)ch * kDaKC + smp;
#pragma unroll
      for (int s = 0; s < CPG; ++s) {
        wn[s] = 0.f;
        if (s < ncopy && b_ < kend) wn[s] = __ldg(wq0 + (size_t)s * M.B + b_);
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    load_w(0);
    for (int ch = 0; ch < nchunks; ++ch) {
      uint64_t wc[CPG];
#pragma unroll
      for (int s = 0; s < CPG; ++s) wc[s] = ptx::f2_pack(wn[s] * qs, wn[s] * qs);
      if (ch + 1 < nchunks) load_w(ch + 1);
      ptx::mbar_wait(&loaded[stage], phase);
      if (warp == 0) TNB_DA_TRACE(ch, 3);
      uint8_t *st = stages + (size_t)stage * kStage + lane_off;
      for (int o = o0; o < noct; o += kOctPerWarp * PW) {
        const uint4 hv = *reinterpret_cast<const uint4 *>(st + o * 128);
        const uint4 lv = *reinterpret_cast<const uint4 *>(st + kDaUPlane + o * 128);
        const __half2 *h_ = reinterpret_cast<const __half2 *>(&hv);
        const __half2 *l_ = reinterpret_cast<const __half2 *>(&lv);
        uint64_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fh = __half22float2(h_[j]), fl = __half22float2(l_[j]);
          v[j] = ptx::f2_add(ptx::f2_pack(fh.x, fh.y), ptx::f2_pack(fl.x, fl.y));
        }
#pragma unroll
        for (int s = 0; s < CPG; ++s) {
          if (s < ncopy) {
            uint32_t ho[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t w2 = ptx::f2_mul(v[j], wc[s]);
              float wa, wb;
              ptx::f2_unpack(w2, wa, wb);
              const __half2 hh = __floats2half2_rn(wa, wb);
              const float2 hf = __half22float2(hh);
              float la, lb;
              ptx::f2_unpack(ptx::f2_sub(w2, ptx::f2_pack(hf.x, hf.y)), la, lb);
              const __half2 ll = __floats2half2_rn(la, lb);
              ho[j] = *reinterpret_cast<const uint32_t *>(&hh);
              lo[j] = *reinterpret_cast<const uint32_t *>(&ll);
            }
            *reinterpret_cast<uint4 *>(st + (s * (ra / 8) + o) * 128) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
            *reinterpret_cast<uint4 *>(st + kDaUPlane + (s * (ra / 8) + o) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ready[stage]);
      if (warp == 0) TNB_DA_TRACE(ch, 4);
      if (++stage == A.nstage) { stage = 0; phase ^= 1; }
    }
    // ===== epilogue (warps 0-3): TMEM -> packed gradient =====
    if (warp < 4) {
      ptx::mbar_wait(done, 0);
      ptx::tc_fence_after();
      const float sc = ldexpf(1.f, qmax - 2 * kTcKA);
      const int chi_w = M.chi;  // bond of the caller's gradient slabs (< chi when the tensor path runs zero-padded)
      const size_t site_off = (size_t)site * chi_w * chi_w * M.D + (site > c ? (size_t)chi_w * chi_w * M.D * (M.C - 1) : 0);
      float *g = A.grads + site_off;
      if (ra == 128 && ncopy == 2) {
        // tile t == copy t of rows a0 .. a0+127: transpose 32 rows x (16 columns x 2 copies) through shared memory (the
        // pipeline stages are idle now) so that a warp adds one 128-byte run of [a][r][sigma] per instruction
        float *stg = reinterpret_cast<float *>(stages) + warp * (32 * 33);
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int abase = a0 + warp * 32;
        for (int n0 = 0; n0 < chi; n0 += 16) {
          float v0[16], v1[16];
          ptx::tmem_ld16(trow + n0, v0);
          ptx::tmem_ld16(trow + chi + n0, v1);
          ptx::tmem_ld_wait();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            stg[lane * 33 + 2 * j] = v0[j] * sc;
            stg[lane * 33 + 2 * j + 1] = v1[j] * sc;
          }
          __syncwarp();
          for (int rr = 0; rr < 32; ++rr) {
            const int a = abase + rr;
            if (a >= a0 + ra_eff || a >= chi_w) break;
            if (n0 + (lane >> 1) < chi_w) atomicAdd(&g[((size_t)a * chi_w + n0 + (lane >> 1)) * Dv + sig0 + (lane & 1)], stg[rr * 33 + lane]);
          }
        }
      } else {
        for (int t = 0; t < ntile; ++t) {
          if (t * 128 + warp * 32 >= nrows) continue;  // warp-uniform: tcgen05.ld is a warp-collective instruction
          const int mm = t * 128 + tid;
          const int sl = mm / ra, al = mm - sl * ra;
          const bool on = sl < ncopy && al < ra_eff && a0 + al < chi_w;
          const int a = a0 + al, sg = sig0 + sl;
          const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)t * chi;
          for (int n0 = 0; n0 < chi; n0 += 16) {
            float v[16];
            ptx::tmem_ld16(trow + n0, v);
            ptx::tmem_ld_wait();
            if (on) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (n0 + j < chi_w) atomicAdd(&g[((size_t)a * chi_w + n0 + j) * Dv + sg], v[j] * sc);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, (uint32_t)A.tmem_cols);
}

}  // namespace tnb
