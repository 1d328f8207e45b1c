// K1 (small-bond regime of the float32 tensor family): the MPS chain at bond 32, register resident on `mma.sync`.
//
// Why not tcgen05 here: at chi = 32 a chain step is six tcgen05.mma of 128 x 64 x 16, and one MMA costs ~120 cycles to
// issue however small it is; the step is bound by instruction issue and by the TMEM -> registers -> shared-memory round
// trip between two dependent steps, not by flops or bytes.  The legacy warp-level MMA has no such
// cost, and at this bond the whole recurrence fits a warp's registers:
//
//   one WARP owns 16 samples (the M of `mma.sync.m16n8k16`); the environment row block v[16, 32] lives in the warp's
//   registers as fp16 hi/lo A fragments; a step is  acc[16, D*32] = v x W_site  (48 HMMAs for D = 2: hi*hi + lo*hi + hi*lo
//   per (k-step, n-tile), fp32 accumulate), the B fragments coming from the site's weight image in shared memory by
//   `ldmatrix` -- the image is the tcgen05 family's K-major stage image, whose 8 x 16-byte core matrices are exactly
//   ldmatrix tiles -- and the epilogue  v'[b, n] = 2^-e_b sum_s phi_s[b] acc[b, s*32 + n]  happens in registers: the
//   accumulator fragments of two adjacent n-tiles ARE the A fragment of one k-step, so the chain never leaves the
//   register file between sites.  The scale is the exact row maximum (two quad shuffles), rows peak in [2^13, 2^14).
//
//   The stash is the tcgen05 family's bond image (mps_tc.cuh): a thread's A-fragment register is four bytes of one
//   16-byte unit and the 32 lanes of a warp cover one 128-byte line, so an environment leaves the SM as 16 coalesced
//   32-bit stores per step and warp, and the weight-gradient GEMM (tc_da_kernel) reads the images unchanged.
//
// CTA: 8 warps (one 128-sample tile), two CTAs per SM, 128 registers per thread.  The site images (D * 4 KB per step)
// come through a ring of bulk async copies issued by the compute warps in turn (warp s % 8 issues the copy of step s + 4
// at the top of its step s; registers are allocated in units of four warps, so a 9th producer warp capped the kernel at
// 96 registers and spilled the ring cursor into the step loop: ncu long-scoreboard stalls on the barrier address).
// One-tile CTAs (rather than one 16-warp CTA per SM) because all CTAs take the same time: in the last, partly filled
// round an SM holds one CTA instead of two and that CTA runs 1.5 x faster (cfg-5 batch: 1024 tiles on 148 SMs, forward
// kernel 1.19 -> 1.09 ms).  The step's feature
// maps, weight scale and forward exponents are fetched ONE STEP AHEAD and the maps are evaluated underneath the step's
// HMMAs: with four warps per scheduler nothing else hides a global-load latency in front of every dependent step.
// Programs, bookkeeping (exps / esum / y / dy / wq / qmax / wqc / qmaxc) and results are those of tc_chain_kernel.
#pragma once
#include "mps_tc.cuh"

namespace tnb {

#ifndef TNB_RRC_WARPS
#define TNB_RRC_WARPS 8
#endif
constexpr int kRrcWarps = TNB_RRC_WARPS;             // warps of 16 samples each: 16 (two sample tiles of 128 per CTA) or 8 (one)
constexpr int kRrcTiles = kRrcWarps / 8;
constexpr int kRrcThreads = kRrcWarps * 32;
constexpr int kRrcStages = 6;                        // ring slots
constexpr int kRrcAhead = 4;                         // the copy of step s + kRrcAhead is issued at the top of step s

__device__ __forceinline__ void rrc_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// the same with a zero accumulator input (the first product term of a step: no accumulator clearing)
__device__ __forceinline__ void rrc_mma0(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};\n"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void rrc_ldsm4(uint32_t saddr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(saddr));
}
// mbarrier wait / arrive on a 32-bit shared-memory address (the step loop keeps the ring's barrier addresses as integers)
__device__ __forceinline__ void rrc_mbar_wait(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "RRC_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra RRC_WAIT_DONE;\n\t"
      "bra RRC_WAIT_LOOP;\n\t"
      "RRC_WAIT_DONE:\n\t"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void rrc_mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(addr) : "memory");
}
// two scaled floats -> fp16x2 hi and lo words
__device__ __forceinline__ void rrc_split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}
// a scaled fp32 pair -> fp16x2 hi and lo words (packed subtraction)
__device__ __forceinline__ void rrc_split2p(uint64_t w, uint32_t &hi, uint32_t &lo) {
  float a, b;
  ptx::f2_unpack(w, a, b);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  float la, lb;
  ptx::f2_unpack(ptx::f2_sub(w, ptx::f2_pack(hf.x, hf.y)), la, lb);
  const __half2 l = __floats2half2_rn(la, lb);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}
// Feature map of one input.  The headline map of the small-bond regime (trigonometric, k = 1, d = 2) takes one sincospif:
// (cos, sin)(pi x / 2) has unit norm to rounding, so embed()'s normaliser (embeddings.py:1374-1387) is the identity in float32.
template <int D>
__device__ __forceinline__ void rrc_embed(const EmbParams<float> &e, const float *xp, float (&ph)[D]) {
  if (D == 2 && e.kind == TNB_EMB_TRIG) {
    float sn, cs;
    sincospif(0.5f * (xp ? xp[0] : 0.5f), &sn, &cs);
    ph[0] = cs;
    ph[D - 1] = sn;
  } else {
    embed_site_d<D>(e, xp, ph);
  }
}
// 2^q as a float (q clamped to the normal range; the slow ldexpf only outside it)
__device__ __forceinline__ float rrc_pow2(int q) {
  return (q >= -126 && q <= 127) ? __int_as_float((q + 127) << 23) : ldexpf(1.f, q);
}
__device__ __forceinline__ float2 rrc_join2(uint32_t hi, uint32_t lo) {
  const float2 h = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
  const float2 l = __half22float2(*reinterpret_cast<const __half2 *>(&lo));
  return make_float2(h.x + l.x, h.y + l.y);
}

template <int D>
__global__ void __launch_bounds__(kRrcThreads, kRrcWarps == 16 ? 1 : 2) rr_chain_kernel(const __grid_constant__ TcArgs A, int adjoint) {
  constexpr int NT = 4 * D;                      // n-tiles of 8 accumulator columns: tile s*4 + j holds n = 8j .. 8j+7 of slice s
  constexpr uint32_t kStage = 2048u * D;         // one K-slab (16 rows) of the site image: hi plane, lo plane
  constexpr uint32_t kLoPlane = 1024u * D;
  constexpr uint32_t kSite = 2 * kStage;         // chi = 32: two K-slabs
  extern __shared__ __align__(1024) uint8_t smem[];
  const MpsArgs<float> &M = A.M;
  const int L = M.L, c = M.c, C = M.C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.y;
  const int nR = L - 1 - c, nL = c;

  // ---- the step program (as in tc_chain_kernel) ----
  int nseed, nchain, total;
  if (!adjoint) {
    nseed = 0; nchain = nR + nL; total = nchain + C;
  } else {
    if (half == 0 && c == 0) return;
    if (half == 1 && c == L - 1) return;
    nseed = C;
    nchain = half == 0 ? c - 1 : L - 2 - c;
    total = nseed + nchain;
  }
  auto step_image = [&](int step) -> const uint8_t * {
    int slab;
    bool tr;
    if (!adjoint) {
      if (step < nR) { slab = mps_slab(M, L - 1 - step); tr = true; }
      else if (step < nR + nL) { slab = mps_slab(M, step - nR); tr = false; }
      else { slab = c + (step - nR - nL); tr = false; }
    } else {
      if (step < nseed) { slab = c + step; tr = half == 0; }
      else if (half == 0) { slab = mps_slab(M, c - 1 - (step - nseed)); tr = true; }
      else { slab = mps_slab(M, c + 1 + (step - nseed)); tr = false; }
    }
    return (tr ? A.img_t : A.img_n) + (size_t)slab * A.site_img_bytes;
  };

  // step table in shared memory: site whose feature map / forward exponents step s needs (s == total: the adjoint's
  // boundary site) and the weight scale of its slab
  auto site_at = [&](int s) -> int {
    if (!adjoint) return s < nR ? L - 1 - s : (s < nR + nL ? s - nR : c);
    if (s < nseed) return c;
    if (s >= total) return half == 0 ? 0 : L - 1;
    return half == 0 ? c - 1 - (s - nseed) : c + 1 + (s - nseed);
  };
  auto slab_at = [&](int s) -> int {
    if (!adjoint) return s < nR + nL ? mps_slab(M, site_at(s)) : c;
    return (s < nseed || s >= total) ? c : mps_slab(M, site_at(s));
  };

  // the CTA's second sample tile may lie beyond the batch: its warps leave, the ring counts the others
  const long long ntiles = (M.B + kTcRows - 1) / kTcRows;
  const long long tile_first = A.tile0 / kTcRows + (long long)blockIdx.x * kRrcTiles;
  const int nact = tile_first + kRrcTiles <= ntiles ? kRrcWarps : 8;

  uint8_t *ring = smem;  // [kRrcStages][kSite]
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)kRrcStages * kSite), *empty = full + kRrcStages;
  int2 *tbl = reinterpret_cast<int2 *>(empty + kRrcStages);  // [total + 2]: (site, kW)
  if (tid == 0) {
    for (int i = 0; i < kRrcStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], (uint32_t)nact);
    }
    ptx::fence_mbar_init();
  }
  for (int s0 = tid; s0 <= total + 1; s0 += kRrcThreads) {
    const int sc = s0 <= total ? s0 : total;
    tbl[s0] = make_int2(site_at(sc), A.kW[slab_at(sc)]);
  }
  __syncthreads();

  // the first kRrcAhead site images (fresh ring slots: nothing to wait for)
  if (tid == 0) {
    for (int s0 = 0; s0 < kRrcAhead && s0 < total; ++s0) {
      ptx::mbar_expect_tx(&full[s0], kSite);
      ptx::bulk_g2s(ring + (size_t)s0 * kSite, step_image(s0), kSite, &full[s0]);
    }
  }
  if (warp >= nact) return;

  // ===== compute warps =====
  const int g = lane >> 2, t = lane & 3;
  const long long row0 = tile_first * kTcRows + warp * 16;  // a multiple of 8
  const long long bA = row0 + g, bB = bA + 8;               // this thread's two sample rows
  const bool vA = bA < M.B, vB = bB < M.B;
  const bool headlane = t < 2;                     // lane t == 0 keeps the books of row A, t == 1 of row B
  const long long bH = t == 0 ? bA : bB;
  const bool vH = headlane && (t == 0 ? vA : vB);
  const long long bE = (t & 1) ? bB : bA;          // the row whose feature maps this lane evaluates (shared by shuffle)
  const bool vE = (t & 1) ? vB : vA;
  const int qbase = lane & ~3;
  const bool scalar_in = M.emb.n_in == 1;
  const float *xrow = M.x + (size_t)(vE ? bE : 0) * L * M.emb.n_in;
  // byte offset of this thread's word in the image of a bond: unit (row, octet o, plane p) at (row/8)*1024 + p*512 + o*128 + (row%8)*16
  const size_t img_off = (size_t)(row0 >> 3) * 1024 + (size_t)lane * 4;
  // ldmatrix.x4 of one (k-step, n-tile): matrices {hi k-octet 0, hi k-octet 1, lo k-octet 0, lo k-octet 1}
  const uint32_t ld_base = ptx::smem_u32(ring) + (uint32_t)(lane & 7) * 16 + (uint32_t)((lane >> 3) & 1) * 128 +
                           (uint32_t)(lane >> 4) * kLoPlane;

  uint32_t ahi[2][4], alo[2][4];  // operand row block v[16, 32] as A fragments of the two k-steps
  float acc[NT][4];
  int stage = 0, gs = 0;  // ring slot and index of the next step
  uint32_t phase = 0;
  const uint32_t full_s = ptx::smem_u32(full), empty_s = ptx::smem_u32(empty);
  int turn = warp;        // steps until this warp's next turn to fetch a site image (every nact steps)

  auto set_e0 = [&](bool gateA, bool gateB) {  // boundary vector e0 * 2^14 (fp16 0x7400 in element 0)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int i = 0; i < 4; ++i) ahi[kk][i] = alo[kk][i] = 0u;
    if (t == 0) {
      ahi[0][0] = gateA ? 0x00007400u : 0u;
      ahi[0][1] = gateB ? 0x00007400u : 0u;
    }
  };
  auto img_slot = [&](uint8_t *base, int bond) -> uint8_t * {
    const int slot = (adjoint || A.img_stash) ? bond : (bond == c + 1 ? 0 : -1);
    if (slot < 0 || base == nullptr) return nullptr;
    return base + (size_t)slot * A.img_bond_bytes + img_off;
  };
  auto store_image = [&](uint8_t *dst) {  // the operand registers -> the image rows of the warp's 16 samples
    if (!dst) return;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // register i of k-step kk: row block (i & 1), octet 2*kk + (i >> 1)
        uint8_t *p = dst + (size_t)(i & 1) * 1024 + (size_t)(2 * kk + (i >> 1)) * 128;
        *reinterpret_cast<uint32_t *>(p) = ahi[kk][i];
        *reinterpret_cast<uint32_t *>(p + 512) = alo[kk][i];
      }
  };
  auto load_image = [&](const uint8_t *src, float2 (&v)[2][4]) {  // this thread's elements of an image, hi + lo
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint8_t *p = src + (size_t)(i & 1) * 1024 + (size_t)(2 * kk + (i >> 1)) * 128;
        v[kk][i] = rrc_join2(*reinterpret_cast<const uint32_t *>(p), *reinterpret_cast<const uint32_t *>(p + 512));
      }
  };
  // acc (+)= operand x site image of the ring's current stage
  auto mma_step = [&](bool accumulate) {
    // this warp's turn to fetch: the image of step gs + kRrcAhead, into the slot every warp has left at step gs - 2
    if (turn == 0) {
      turn = nact;
      if (gs + kRrcAhead < total && lane == 0) {
        const int sf = gs + kRrcAhead, slot = sf % kRrcStages, use = sf / kRrcStages;
        if (use > 0) rrc_mbar_wait(empty_s + (uint32_t)slot * 8, (uint32_t)(use - 1) & 1u);
        ptx::mbar_expect_tx(&full[slot], kSite);
        ptx::bulk_g2s(ring + (size_t)slot * kSite, step_image(sf), kSite, &full[slot]);
      }
      __syncwarp();
    }
    --turn;
    ++gs;
    rrc_mbar_wait(full_s + (uint32_t)stage * 8, phase);
    const uint32_t st = ld_base + (uint32_t)stage * kSite;
    // Four n-tiles at a time, one product term after the other: two HMMAs into the same accumulator are four
    // instructions apart (a dependent HMMA waits ~21 cycles; back to back the three terms of a tile were a serial chain).
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int j0 = 0; j0 < NT; j0 += 4) {
        uint32_t bh0[4], bh1[4], bl0[4], bl1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rrc_ldsm4(st + (uint32_t)kk * kStage + (uint32_t)(j0 + j) * 256, bh0[j], bh1[j], bl0[j], bl1[j]);
        if (kk == 0 && !accumulate) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rrc_mma0(acc[j0 + j], ahi[kk], bh0[j], bh1[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) rrc_mma(acc[j0 + j], ahi[kk], bh0[j], bh1[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) rrc_mma(acc[j0 + j], alo[kk], bh0[j], bh1[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) rrc_mma(acc[j0 + j], ahi[kk], bl0[j], bl1[j]);
      }
    __syncwarp();
    if (lane == 0) rrc_mbar_arrive(empty_s + (uint32_t)stage * 8);
    if (++stage == kRrcStages) { stage = 0; phase ^= 1; }
  };

  // ---- feature maps: a lane evaluates ONE map per two steps -- row A or B (t & 1) at step 2b + (t >> 1) of block b --
  // from an input fetched two steps earlier; the quad shares the values by shuffle.
  float phq[D];      // this lane's map of the current block
  float xq = 0.5f;   // this lane's input of the NEXT block (one-scalar maps)
  auto load_xq = [&](int blk) {
    const int sb = min(2 * blk + (t >> 1), total);
    if (scalar_in) xq = vE ? xrow[tbl[sb].x] : 0.5f;
  };
  auto eval_block = [&](int blk, float (&ph)[D]) {
    if (scalar_in) {
      const float xv = xq;
      rrc_embed<D>(M.emb, &xv, ph);
    } else {
      const int sb = min(2 * blk + (t >> 1), total);
      rrc_embed<D>(M.emb, vE ? xrow + (size_t)tbl[sb].x * M.emb.n_in : nullptr, ph);
    }
  };
  float pa[D], pb[D];  // maps of rows A / B at the current step
  auto take_phis = [&](int s) {  // (pa, pb) <- the maps of step s out of the quad's block values
    const int src = qbase + 2 * (s & 1);
#pragma unroll
    for (int q = 0; q < D; ++q) {
      pa[q] = __shfl_sync(0xffffffffu, phq[q], src);
      pb[q] = __shfl_sync(0xffffffffu, phq[q], src + 1);
    }
  };
  // called once per step s after its MMAs are issued and BEFORE (pa, pb) of step s are last used: evaluates the next block at
  // odd steps (underneath the HMMAs); finish_step then moves on to the maps of step s + 1.
  float phn[D];
  auto eval_ahead = [&](int s) {
    if (s & 1) {
      eval_block((s + 1) >> 1, phn);
      load_xq(((s + 1) >> 1) + 1);
    }
  };
  auto finish_step = [&](int s) {
    if (s & 1) {
#pragma unroll
      for (int q = 0; q < D; ++q) phq[q] = phn[q];
    }
    take_phis(s + 1);
  };
  load_xq(0);
  eval_block(0, phq);
  load_xq(1);
  take_phis(0);

  // o[j] (packed pairs; A rows: oA, B rows: oB) = sum_s phi_s acc[s*4 + j]
  auto contract = [&](uint64_t (&oA)[4], uint64_t (&oB)[4]) {
    uint64_t PA[D], PB[D];
#pragma unroll
    for (int q = 0; q < D; ++q) {
      PA[q] = ptx::f2_pack(pa[q], pa[q]);
      PB[q] = ptx::f2_pack(pb[q], pb[q]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      oA[j] = ptx::f2_mul(PA[0], ptx::f2_pack(acc[j][0], acc[j][1]));
      oB[j] = ptx::f2_mul(PB[0], ptx::f2_pack(acc[j][2], acc[j][3]));
#pragma unroll
      for (int q = 1; q < D; ++q) {
        oA[j] = ptx::f2_fma(PA[q], ptx::f2_pack(acc[q * 4 + j][0], acc[q * 4 + j][1]), oA[j]);
        oB[j] = ptx::f2_fma(PB[q], ptx::f2_pack(acc[q * 4 + j][2], acc[q * 4 + j][3]), oB[j]);
      }
    }
  };
  // Scale of a row of max-abs `m`: 2^(14 - e), m = f * 2^e, f in [0.5, 1)  ->  the scaled row peaks in [2^13, 2^14).
  auto row_scale = [&](float m, bool valid, int *e_out) -> float {
    const int ef = (int)((__float_as_uint(m) >> 23) & 0xffu);
    const bool ok = valid && ef >= 16 && ef < 255;  // (a zero row, or one below 2^-110: flushed)
    *e_out = ok ? ef - 126 : 0;
    return ok ? __uint_as_float((uint32_t)(127 + 126 + kTcKA - ef) << 23) : 0.f;
  };
  // contracted rows -> normalised fp16 hi/lo operand; dA / dB: the rows' exponents in units of the true recurrence
  // (e - 14 - kw; 0 and a zero row if flushed)
  auto renorm = [&](const uint64_t (&oA)[4], const uint64_t (&oB)[4], int kw, int &dA, int &dB) {
    float mA = 0.f, mB = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a0, a1, b0, b1;
      ptx::f2_unpack(oA[j], a0, a1);
      ptx::f2_unpack(oB[j], b0, b1);
      mA = fmaxf(mA, fmaxf(fabsf(a0), fabsf(a1)));
      mB = fmaxf(mB, fmaxf(fabsf(b0), fabsf(b1)));
    }
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
    int eA, eB;
    const float sA = row_scale(mA, vA, &eA), sB = row_scale(mB, vB, &eB);
    dA = sA != 0.f ? eA - kTcKA - kw : 0;
    dB = sB != 0.f ? eB - kTcKA - kw : 0;
    const uint64_t SA = ptx::f2_pack(sA, sA), SB = ptx::f2_pack(sB, sB);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      rrc_split2p(ptx::f2_mul(oA[2 * kk], SA), ahi[kk][0], alo[kk][0]);
      rrc_split2p(ptx::f2_mul(oB[2 * kk], SB), ahi[kk][1], alo[kk][1]);
      rrc_split2p(ptx::f2_mul(oA[2 * kk + 1], SA), ahi[kk][2], alo[kk][2]);
      rrc_split2p(ptx::f2_mul(oB[2 * kk + 1], SB), ahi[kk][3], alo[kk][3]);
    }
  };
  auto sel = [&](int a, int b) -> int { return t == 0 ? a : b; };  // the head lane's row

  // One chain step s: acc = operand x slab, next operand = normalised phi-contraction; moves (pa, pb) on to step s + 1.
  auto chain_step = [&](int s, int &dA, int &dB) {
    mma_step(false);
    eval_ahead(s);
    uint64_t oA[4], oB[4];
    contract(oA, oB);
    renorm(oA, oB, tbl[s].y, dA, dB);
    finish_step(s);
  };

  if (!adjoint) {
    // ---------------- forward program ----------------
    int eaA = 0, eaB = 0, ebA = 0, ebB = 0;  // exponent sums of the left / right chain
    set_e0(vA, vB);
    store_image(img_slot(A.fimg, L));
    store_image(img_slot(A.fimg, 0));
    set_e0(true, true);
    int s = 0;
    for (int i = 0; i < nR; ++i, ++s) {
      const int site = L - 1 - i;
      int dA, dB;
      chain_step(s, dA, dB);
      ebA += dA;
      ebB += dB;
      if (vH) M.exps[(size_t)site * M.B + bH] = sel(dA, dB);
      store_image(img_slot(A.fimg, site));
      if (i == nR - 1) set_e0(true, true);  // F[c+1] is in its image; the left chain starts from the boundary vector
    }
    for (int i = 0; i < nL; ++i, ++s) {
      int dA, dB;
      chain_step(s, dA, dB);
      eaA += dA;
      eaB += dB;
      if (vH) M.exps[(size_t)i * M.B + bH] = sel(dA, dB);
      store_image(img_slot(A.fimg, i + 1));
    }
    // ---------------- class site + head ----------------
    {
      const float ysc = ldexpf(1.f, -(2 * kTcKA + A.kW[c]));
      // F[c+1] of this thread's elements, from its image (written by this thread: at the start if c + 1 == L)
      float2 fr[2][4];
      load_image(img_slot(A.fimg, c + 1), fr);
      for (int k = 0; k < C; ++k, ++s) {
        mma_step(false);
        eval_ahead(s);
        uint64_t oA[4], oB[4];
        contract(oA, oB);
        float yA = 0.f, yB = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          float a0, a1, b0, b1;
          ptx::f2_unpack(oA[2 * kk], a0, a1);
          ptx::f2_unpack(oB[2 * kk], b0, b1);
          yA = fmaf(a0, fr[kk][0].x, fmaf(a1, fr[kk][0].y, yA));
          yB = fmaf(b0, fr[kk][1].x, fmaf(b1, fr[kk][1].y, yB));
          ptx::f2_unpack(oA[2 * kk + 1], a0, a1);
          ptx::f2_unpack(oB[2 * kk + 1], b0, b1);
          yA = fmaf(a0, fr[kk][2].x, fmaf(a1, fr[kk][2].y, yA));
          yB = fmaf(b0, fr[kk][3].x, fmaf(b1, fr[kk][3].y, yB));
        }
        yA += __shfl_xor_sync(0xffffffffu, yA, 1);
        yB += __shfl_xor_sync(0xffffffffu, yB, 1);
        yA += __shfl_xor_sync(0xffffffffu, yA, 2);
        yB += __shfl_xor_sync(0xffffffffu, yB, 2);
        if (vH) M.y[(size_t)bH * C + k] = (t == 0 ? yA : yB) * ysc;
        finish_step(s);
      }
      float lval = 0.f;
      if (vH) {
        const int ea = sel(eaA, eaB), eb = sel(ebA, ebB);
        M.esum[bH] = ea;
        M.esum[M.B + bH] = eb;
        float yl[kMaxOut], tl[kMaxOut];
        for (int k = 0; k < C; ++k) {
          yl[k] = M.y[(size_t)bH * C + k];
          tl[k] = M.targets ? M.targets[(size_t)bH * C + k] : 0.f;
          if (M.out) M.out[(size_t)bH * C + k] = yl[k];
        }
        if (M.out_exp) M.out_exp[bH] = ea + eb;
        lval = mps_head<float>(M.head, C, yl, ea + eb, tl, M.cw, (float *)nullptr);
        if (M.loss) M.loss[bH] = lval;
      }
      if (M.loss_sum) {  // one atomic per warp
        double ls = (double)lval;
        for (int off = 16; off > 0; off >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, off);
        if (lane == 0) atomicAdd(M.loss_sum, ls);
      }
    }
  } else {
    // ---------------- adjoint program ----------------
    // head derivative dLoss/dy' (times loss_scale): the head lanes, through M.dy (both halves write identical values)
    int edH = 0;
    if (headlane) {
      float yl[kMaxOut], tl[kMaxOut], dyl[kMaxOut];
      for (int k = 0; k < C; ++k) {
        yl[k] = vH ? M.y[(size_t)bH * C + k] : 1.f;
        tl[k] = (vH && M.targets) ? M.targets[(size_t)bH * C + k] : (k == 0 ? 1.f : 0.f);
      }
      const int E = vH ? M.esum[bH] + M.esum[M.B + bH] : 0;
      mps_head<float>(M.head, C, yl, E, tl, M.cw, dyl);
      float dmax = 0.f;
      for (int k = 0; k < C; ++k) {
        dyl[k] = vH ? dyl[k] * M.loss_scale : 0.f;
        dmax = fmaxf(dmax, fabsf(dyl[k]));
      }
      if (dmax > 0.f) frexpf(dmax, &edH);
      if (vH) {
        // class-site dA weights: wqc[(s,k)][b] = phi_s * dy'_k (this lane's own row: t == 0 -> A, 1 -> B)
        int qm = INT_MIN;
#pragma unroll
        for (int q = 0; q < D; ++q)
          for (int k = 0; k < C; ++k) {
            const float w = (t == 0 ? pa[q] : pb[q]) * dyl[k];
            A.wqc[((size_t)q * C + k) * M.B + bH] = w;
            if (w != 0.f) { int ew; frexpf(w, &ew); qm = max(qm, ew); }
          }
        for (int k = 0; k < C; ++k) M.dy[(size_t)bH * C + k] = dyl[k];
        if (qm != INT_MIN) atomicMax(A.qmaxc, qm);
      }
    }
    __syncwarp();
    const int edA = __shfl_sync(0xffffffffu, edH, qbase), edB = __shfl_sync(0xffffffffu, edH, qbase + 1);
    const float dscA = vA ? rrc_pow2(-edA) : 0.f, dscB = vB ? rrc_pow2(-edB) : 0.f;  // |dy_k * dsc| < 1
    // forward environment rows (operand units) of this thread's elements
    float2 vr[2][4];
    load_image(img_slot(A.fimg, half == 0 ? c + 1 : c), vr);
    int s = 0;
    for (int k = 0; k < C; ++k, ++s) {
      // seed operand of slab c + k: F row * dy_k * 2^-ed (rows normalised to 2^14, |factor| < 1)
      const float fA = vA ? M.dy[(size_t)bA * C + k] * dscA : 0.f, fB = vB ? M.dy[(size_t)bB * C + k] * dscB : 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float f = (i & 1) ? fB : fA;
          rrc_split2(vr[kk][i].x * f, vr[kk][i].y * f, ahi[kk][i], alo[kk][i]);
        }
      mma_step(k > 0);  // accumulates over k
      eval_ahead(s);
      if (k < C - 1) finish_step(s);  // (the class site's maps again)
    }
    // forward exponents of the two rows at the step's site, fetched one step ahead
    auto load_ds = [&](int sn, int &a, int &b) {
      const int site = tbl[sn].x;
      a = vA ? M.exps[(size_t)site * M.B + bA] : 0;
      b = vB ? M.exps[(size_t)site * M.B + bB] : 0;
    };
    int dsA, dsB;
    load_ds(nseed, dsA, dsB);
    int hA, hB;  // running exponents of the adjoint rows
    {
      uint64_t oA[4], oB[4];
      contract(oA, oB);
      int dA, dB;
      renorm(oA, oB, A.kW[c], dA, dB);
      hA = dA + edA;  // G = row * 2^h (the operand carried dy * 2^-ed)
      hB = dB + edB;
      store_image(img_slot(A.gimg, half == 0 ? c : c + 1));
      finish_step(nseed - 1);
    }
    for (int i = 0; i < nchain; ++i, ++s) {
      const int site = half == 0 ? c - 1 - i : c + 1 + i;
      // dA_site = sum_b 2^q (F x phi x G), q = h_in - delta_site, h_in = exponent of the input adjoint
      const int qA = hA - dsA, qB = hB - dsB;
      if (vH) {
        const float sc2 = rrc_pow2(sel(qA, qB));
#pragma unroll
        for (int q = 0; q < D; ++q) A.wq[((size_t)site * D + q) * M.B + bH] = (t == 0 ? pa[q] : pb[q]) * sc2;
      }
      // max_b q: one RED per warp and step on ONE address would serialise in L2 (16 k warps x ~100 steps); the value
      // settles after the first few warps, so the current maximum is read ahead of the step and most REDs are skipped
      // (no warp reduction either: every head lane compares its own row)
      const int qseen = vH ? *reinterpret_cast<const volatile int *>(&A.qmax[site]) : INT_MAX;
      const int qm = sel(qA, qB);
      int nA, nB;
      load_ds(s + 1, nA, nB);  // (s + 1 == total: the boundary site)
      int dA, dB;
      chain_step(s, dA, dB);
      if (qm > qseen) atomicMax(&A.qmax[site], qm);  // (qmax starts far below any exponent; invalid rows: qseen = INT_MAX)
      store_image(img_slot(A.gimg, half == 0 ? c - 1 - i : c + 2 + i));
      hA += dA - dsA;
      hB += dB - dsB;
      dsA = nA;
      dsB = nB;
    }
    // boundary site: no chain step, but it has a gradient ((pa, pb) and ds hold its inputs)
    {
      const int jb = half == 0 ? 0 : L - 1;
      const int qA = vA ? hA - dsA : INT_MIN, qB = vB ? hB - dsB : INT_MIN;
      if (vH) {
        const float sc2 = rrc_pow2(sel(qA, qB));
#pragma unroll
        for (int q = 0; q < D; ++q) A.wq[((size_t)jb * D + q) * M.B + bH] = (t == 0 ? pa[q] : pb[q]) * sc2;
      }
      int qm = max(qA, qB);
      for (int off = 16; off > 0; off >>= 1) qm = max(qm, __shfl_xor_sync(0xffffffffu, qm, off));
      if (lane == 0 && qm != INT_MIN) atomicMax(&A.qmax[jb], qm);
    }
  }
}

}  // namespace tnb
