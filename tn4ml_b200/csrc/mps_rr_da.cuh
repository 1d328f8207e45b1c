// K4 (small-bond regime of the float32 tensor family): weight gradients of the chain sites at bond 32 on `mma.sync`.
//
//   dA_site[a, r, sigma] = 2^qmax * sum_b U[b,a] * (wq[site][sigma][b] * 2^-qmax * V[b,r])
//
// Same operands, weights, scaling and output layout as tc_da_kernel (mps_tc.cuh), which at this bond is bound by the issue
// cost of its small tcgen05.mma and by the U re-weighting in 4 producer warps (1.70 ms for 6.6 GB of images at cfg-5 / chi = 32).
// Here one WARP owns a chunk of 16 samples (the K of `mma.sync.m16n8k16`) of one site:
//   * both operands are the bond images as they lie in HBM: a chunk of 16 samples of an image is 2 KB, bulk-copied into the
//     warp's own ring; its 16-byte units are 8 x 8 fp16 tiles [sample % 8][bond index % 8], so `ldmatrix.trans` yields the
//     A fragments (rows a, k = sample) of U without any arithmetic, and the raw (k = sample, n = r) B fragments of V;
//   * the per-sample weight goes on V: a B-fragment register holds two consecutive samples of one column, one packed fp32x2
//     multiply per pair and copy sigma, then the fp16 hi/lo re-split; N = (sigma, r): 2 m-tiles x 4*D n-tiles x 3 terms
//     = 48 HMMAs per chunk for D = 2, the accumulator dA_site[32, 32*D] (64 registers per thread) stays in the warp for its
//     whole share of the batch;
//   * at the end the CTA's warps are summed through shared memory in a fixed order and added to the packed gradient
//     (one atomic per entry and CTA; a single writer per entry when nsplit == 1: the deterministic mode).
// The class site (D*C weighted copies) stays on tc_da_kernel.
#pragma once
#include "mps_rr_chain.cuh"

namespace tnb {

constexpr int kRdWarps = 8;
constexpr int kRdStages = 3;             // chunks in flight per warp
constexpr uint32_t kRdChunk = 2048;      // 16 samples of a bond-32 image
constexpr uint32_t kRdStage = 2 * kRdChunk;  // U chunk, V chunk

__device__ __forceinline__ void rrd_ldsm4t(uint32_t saddr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(saddr));
}

template <int D>
__global__ void __launch_bounds__(kRdWarps * 32, D == 2 ? 2 : 1) rr_da_kernel(const __grid_constant__ TcDaArgs A) {
  constexpr int NT = 4 * D;  // n-tiles: tile sigma*4 + j holds columns r = 8j .. 8j+7 of copy sigma
  extern __shared__ __align__(1024) uint8_t smem[];
  const MpsArgs<float> &M = A.M;
  const int c = M.c;
  const int site = A.site0 + (int)blockIdx.y;
  if (site == c) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long kbeg = (long long)blockIdx.x * A.bchunk;
  const long long kend = min(M.B, kbeg + A.bchunk);
  if (kbeg >= kend) return;
  const int nchunks = (int)((kend - kbeg + 15) / 16);
  const int mine = nchunks > warp ? (nchunks - warp + kRdWarps - 1) / kRdWarps : 0;  // chunks warp, warp + 8, ...
  // U = F[site] (left of the class site) or G[site] (right); V = G[site+1] (left) or F[site+1] (right)
  const size_t chunk0 = (size_t)(kbeg / 8) * 1024;
  const uint8_t *Uimg = (site <= c ? A.fimg : A.gimg) + (size_t)site * A.img_bond_bytes + chunk0;
  const uint8_t *Vimg = (site < c ? A.gimg : A.fimg) + (size_t)(site + 1) * A.img_bond_bytes + chunk0;

  uint8_t *ring = smem + (size_t)warp * kRdStages * kRdStage;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)kRdWarps * kRdStages * kRdStage) + warp * kRdStages;
  if (lane == 0) {
    for (int i = 0; i < kRdStages; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  auto issue = [&](int i) {  // this warp's i-th chunk -> ring slot i % kRdStages
    const size_t off = (size_t)(warp + i * kRdWarps) * kRdChunk;
    uint8_t *dst = ring + (size_t)(i % kRdStages) * kRdStage;
    ptx::mbar_expect_tx(&bars[i % kRdStages], kRdStage);
    ptx::bulk_g2s(dst, Uimg + off, kRdChunk, &bars[i % kRdStages]);
    ptx::bulk_g2s(dst + kRdChunk, Vimg + off, kRdChunk, &bars[i % kRdStages]);
  };
  if (lane == 0)
    for (int i = 0; i < kRdStages && i < mine; ++i) issue(i);

  const int qmax = A.qmax[site];
  const float qs = rrc_pow2(-qmax);  // (the images already carry the 2^14 operand scale)
  const float *wq0 = A.wq + (size_t)site * D * M.B;
  // weights of a chunk: samples 2t, 2t+1 (k-octet 0) and 8+2t, 9+2t (k-octet 1) of every copy, fetched one chunk ahead
  float2 wn[D][2];
  auto load_w = [&](int i) {
    const long long b0 = kbeg + (long long)(warp + i * kRdWarps) * 16 + 2 * t;
#pragma unroll
    for (int s = 0; s < D; ++s)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long b = b0 + 8 * h;  // (even: M.B rows of 4-byte weights, 8-byte aligned pairs only if M.B is even)
        wn[s][h].x = b < kend ? __ldg(wq0 + (size_t)s * M.B + b) : 0.f;
        wn[s][h].y = b + 1 < kend ? __ldg(wq0 + (size_t)s * M.B + b + 1) : 0.f;
      }
  };
  float acc[2][NT][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][j][i] = 0.f;

  // ldmatrix.trans addressing inside a chunk: unit (sample block k, plane p, octet o, sample % 8) at k*1024 + p*512 + o*128 + (s%8)*16
  const uint32_t row16 = (uint32_t)(lane & 7) * 16, mi = (uint32_t)lane >> 3;
  // U, m-tile mt, plane p: matrices {(octet 2mt, block 0), (octet 2mt+1, block 0), (octet 2mt, block 1), (octet 2mt+1, block 1)}
  const uint32_t u_off = row16 + (mi & 1) * 128 + (mi >> 1) * 1024;
  // V, octet j: matrices {(block 0, hi), (block 1, hi), (block 0, lo), (block 1, lo)}
  const uint32_t v_off = kRdChunk + row16 + (mi & 1) * 1024 + (mi >> 1) * 512;
  const uint32_t ring_s = ptx::smem_u32(ring);

  if (mine > 0) load_w(0);
  for (int i = 0; i < mine; ++i) {
    uint64_t wc[D][2];
#pragma unroll
    for (int s = 0; s < D; ++s)
#pragma unroll
      for (int h = 0; h < 2; ++h) wc[s][h] = ptx::f2_pack(wn[s][h].x * qs, wn[s][h].y * qs);
    if (i + 1 < mine) load_w(i + 1);
    const int slot = i % kRdStages;
    ptx::mbar_wait(&bars[slot], (uint32_t)(i / kRdStages) & 1u);
    const uint32_t st = ring_s + (uint32_t)slot * kRdStage;
    uint32_t ahi[2][4], alo[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      rrd_ldsm4t(st + u_off + (uint32_t)mt * 256, ahi[mt][0], ahi[mt][1], ahi[mt][2], ahi[mt][3]);
      rrd_ldsm4t(st + u_off + (uint32_t)mt * 256 + 512, alo[mt][0], alo[mt][1], alo[mt][2], alo[mt][3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t h0, h1, l0, l1;
      rrd_ldsm4t(st + v_off + (uint32_t)j * 128, h0, h1, l0, l1);
      const float2 v0 = rrc_join2(h0, l0), v1 = rrc_join2(h1, l1);  // samples (2t, 2t+1) and (8+2t, 9+2t) of column 8j + g
      const uint64_t p0 = ptx::f2_pack(v0.x, v0.y), p1 = ptx::f2_pack(v1.x, v1.y);
      uint32_t bh0[D], bh1[D], bl0[D], bl1[D];
#pragma unroll
      for (int s = 0; s < D; ++s) {
        rrc_split2p(ptx::f2_mul(p0, wc[s][0]), bh0[s], bl0[s]);
        rrc_split2p(ptx::f2_mul(p1, wc[s][1]), bh1[s], bl1[s]);
      }
      // term-major over (copy, m-tile): two HMMAs into one accumulator are 2*D instructions apart
#pragma unroll
      for (int s = 0; s < D; ++s)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) rrc_mma(acc[mt][s * 4 + j], ahi[mt], bh0[s], bh1[s]);
#pragma unroll
      for (int s = 0; s < D; ++s)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) rrc_mma(acc[mt][s * 4 + j], alo[mt], bh0[s], bh1[s]);
#pragma unroll
      for (int s = 0; s < D; ++s)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) rrc_mma(acc[mt][s * 4 + j], ahi[mt], bl0[s], bl1[s]);
    }
    // the slot is free: refill it with this warp's chunk i + kRdStages
    __syncwarp();
    if (i + kRdStages < mine && lane == 0) {
      ptx::fence_proxy_async();
      issue(i + kRdStages);
    }
  }

  // ---- the CTA's warps, summed in a fixed order, into the packed gradient [a][r][sigma] ----
  __syncthreads();  // every warp is done with its ring: the memory is reused as red[warp][a][r][sigma]
  float *red = reinterpret_cast<float *>(smem);
  constexpr int kTile = 32 * 32 * D;
  {
    float *mineb = red + (size_t)warp * kTile;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int a = 16 * mt + g + 8 * (i >> 1), r = 8 * (j & 3) + 2 * t + (i & 1), s = j >> 2;
          mineb[(a * 32 + r) * D + s] = acc[mt][j][i];
        }
  }
  __syncthreads();
  const float sc = rrc_pow2(qmax - 2 * kTcKA);
  const int chi_w = M.chi;  // bond of the caller's gradient slabs (< 32 when the path runs zero-padded)
  const size_t site_off = (size_t)site * chi_w * chi_w * D + (site > c ? (size_t)chi_w * chi_w * D * (M.C - 1) : 0);
  float *gout = A.grads + site_off;
  for (int e = tid; e < kTile; e += kRdWarps * 32) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kRdWarps; ++w) v += red[(size_t)w * kTile + e];
    const int a = e / (32 * D), r = (e / D) & 31, s_ = e % D;
    if (a < chi_w && r < chi_w) atomicAdd(&gout[((size_t)a * chi_w + r) * D + s_], v * sc);
  }
}

}  // namespace tnb
