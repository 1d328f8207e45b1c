// K2 + K3 (tensor-core family): the whole per-sample chain on tcgen05 / TMEM.
//
// One CTA owns 128 samples (sample row == TMEM lane) and runs a "program" of MMA steps; every step
// is one GEMM  acc[128, D*chi] = operand[128, chi] x W[chi, D*chi]  whose B operand is a pre-split
// weight image streamed by 1-D bulk copies, followed by an epilogue that turns the accumulator into
// the next A operand without leaving the SM.
//
//   forward (grid.y == 1):   right chain  (sites L-1 .. c+1, transposed images)
//                            left chain   (sites 0 .. c-1)
//                            class steps  (slabs c+k, k < C):  y'_k = <F[c] M_c^k , F[c+1]>
//                            head         (loss, out)                         -- K3
//   adjoint (grid.y == 2):   seed steps   (k < C, accumulating in TMEM):
//                                G[c]   = sum_k dy'_k M_c^k F[c+1]   (half 0)
//                                G[c+1] = sum_k dy'_k F[c] M_c^k     (half 1)
//                            adjoint chain towards the boundary, recording the per-sample weights
//                            wq / qmax of the dA GEMM (mps_tc.cuh, tc_da_kernel).
//
// The N = D*chi accumulator columns are issued as NH column chunks of NMMA = D*chi/NH columns;
// chunk h holds the outputs n in [h*chi/NH, (h+1)*chi/NH) for every physical index s
// (TMEM column h*NMMA + s*chi/NH + n - h*chi/NH) and has its own commit barrier, so the epilogue
// of chunk h runs underneath the MMAs of chunk h+1.
//
// A-priori scaling.  The next operand must be a row of bounded magnitude (fp16 range) times an exact power of two.
// Instead of the row maximum of the outputs (a reduction on the critical path) the scale comes from a bound known
// before the step runs: |o[n]| <= m * colsum, m = max-abs of the operand row (found while the NEXT step's MMAs run),
// colsum = max_n sum_{k,s} |W[k,s,n]| (tc_stats_kernel; |phi_s| <= 1).  The row then peaks a few bits below 2^14
// (the looseness of the bound, never accumulating because m is exact), which costs nothing in fp16 hi/lo.
//
// Warp roles (608 threads):
//   warps 0-15  epilogue.  A TMEM lane quarter is only visible to warps with warp % 4 == quarter, so the four
//               warps {q, q+4, q+8, q+12} share the 32 sample rows of quarter q and split a row's 16-column
//               groups between them (group j belongs to warp group j % 4).  ONE pass over TMEM: a group is
//               phi-contracted, scaled and split into fp16 hi/lo as soon as its column chunk is complete (under
//               the MMAs of the following chunks) and parked in the TMEM columns it came from (tcgen05.st; only
//               its own thread ever touches those columns), so no thread carries more than one group in registers.
//               When the last chunk is in, its groups are packed and everything is written to shared memory as the
//               next operand.  Row maximum (4 KB shared-memory exchange + a 128-thread named barrier per lane
//               quarter) and per-sample bookkeeping follow after the next step's MMAs have been released.
//   warp 16     TMEM owner + MMA issuer.
//   warp 17     weight producer: one bulk copy global -> shared per pipeline stage.
//   warp 18     image writer: the A operand IS the bond image of these 128 samples (same bytes); it goes to the HBM
//               stash as paced 4 KB bulk copies shared -> global while the step's MMAs run.  (Per-thread stores from
//               the epilogue put 128 KB per step on the critical path, 4.5 k cycles; one unpaced 128 KB bulk copy
//               starved the weight loads for as long.)
//
// Environments are carried as (mantissa row, integer exponent): every epilogue renormalises the
// row to max-abs in [0.5, 1) by an exact power of two, which is also what keeps the fp16 hi/lo
// operand split inside the fp16 range (see mps_tc.cuh for the fp16x3 arithmetic).
#pragma once
#include "mps_tc.cuh"

namespace tnb {

// Optional in-kernel timeline (built only with -DTNB_TC_TRACE, tools/trace_chain.py): CTA (0,0) records clock64() at
// fixed points of a few steady-state steps; read back with tnb_debug_trace_read.
#ifdef TNB_TC_TRACE
#ifndef TNB_TRACE_STEP0
#define TNB_TRACE_STEP0 20  // first traced step (short chains: build with -DTNB_TRACE_STEP0=2)
#endif
__device__ long long g_tc_trace[64 * 32];
#define TNB_TRACE(who, step_, ev)                                                                              \
  do {                                                                                                         \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (step_) >= TNB_TRACE_STEP0 && (step_) < TNB_TRACE_STEP0 + 32 && (threadIdx.x & 31) == 0 && (who) < 4) \
      g_tc_trace[((step_)-TNB_TRACE_STEP0) * 64 + (who)*16 + (ev)] = clock64();                                              \
  } while (0)
#define TNB_TRACE_SET(who, step_, ev, val)                                                                      \
  do {                                                                                                         \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (step_) >= TNB_TRACE_STEP0 && (step_) < TNB_TRACE_STEP0 + 32 && (threadIdx.x & 31) == 0 && (who) < 4) \
      g_tc_trace[((step_)-TNB_TRACE_STEP0) * 64 + (who)*16 + (ev)] = (val);                                                  \
  } while (0)
#define TNB_CLOCK() clock64()
#define TNB_DBG(bit) (A.dbg & (bit))   // timing experiments of the trace build (TNB_TC_DBG); some give wrong results
#else
#define TNB_TRACE(who, step_, ev) do { } while (0)
#define TNB_TRACE_SET(who, step_, ev, val) do { } while (0)
#define TNB_CLOCK() 0ll
#define TNB_DBG(bit) 0                 // release builds carry none of the experiment switches
#endif

// acc columns -> o[j] = sum_s ph[s] * acc[col + s*cstride + j], j < 16   (one 16-column group).
// The s = 0 slice is loaded straight into o (registers are the scarce resource in the chain kernel).
template <int D>
__device__ __forceinline__ void tc_combine16(uint32_t col, uint32_t cstride, const float (&ph)[D], float (&o)[16]) {
  ptx::tmem_ld16(col, o);
  if constexpr (D == 1) {
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] *= ph[0];
  } else {
    float t[16];
    ptx::tmem_ld16(col + cstride, t);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = fmaf(ph[1], t[j], ph[0] * o[j]);
#pragma unroll
    for (int s = 2; s < D; ++s) {
      ptx::tmem_ld16(col + (uint32_t)s * cstride, t);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = fmaf(ph[s], t[j], o[j]);
    }
  }
}

// Same contraction on packed fp32 pairs (phs = phi already multiplied by the step's power-of-two scale), followed by
// the fp16 hi/lo split.  pk[2*o + plane]: octet o (values 8o..8o+7), plane 0 = hi, 1 = lo.  Returns the max-abs of
// the 16 scaled values.
template <int D>
__device__ __forceinline__ float tc_combine_pack16(uint32_t col, uint32_t cstride, const float (&phs)[D], uint4 (&pk)[4]) {
  float t0[16];
  uint64_t w[8];
  ptx::tmem_ld16(col, t0);
  if constexpr (D == 1) {
    ptx::tmem_ld_wait();
    const uint64_t p0 = ptx::f2_pack(phs[0], phs[0]);
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = ptx::f2_mul(p0, ptx::f2_pack(t0[2 * j], t0[2 * j + 1]));
  } else {
    float t1[16];
    ptx::tmem_ld16(col + cstride, t1);
    ptx::tmem_ld_wait();
    const uint64_t p0 = ptx::f2_pack(phs[0], phs[0]), p1 = ptx::f2_pack(phs[1], phs[1]);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      w[j] = ptx::f2_fma(p1, ptx::f2_pack(t1[2 * j], t1[2 * j + 1]), ptx::f2_mul(p0, ptx::f2_pack(t0[2 * j], t0[2 * j + 1])));
#pragma unroll
    for (int s = 2; s < D; ++s) {
      ptx::tmem_ld16(col + (uint32_t)s * cstride, t1);
      ptx::tmem_ld_wait();
      const uint64_t ps = ptx::f2_pack(phs[s], phs[s]);
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = ptx::f2_fma(ps, ptx::f2_pack(t1[2 * j], t1[2 * j + 1]), w[j]);
    }
  }
  float m = 0.f;
  __half2 hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a, b;
    ptx::f2_unpack(w[j], a, b);
    m = fmaxf(m, fmaxf(fabsf(a), fabsf(b)));
    hi[j] = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(hi[j]);
    float la, lb;
    ptx::f2_unpack(ptx::f2_sub(w[j], ptx::f2_pack(hf.x, hf.y)), la, lb);
    lo[j] = __floats2half2_rn(la, lb);
  }
  auto u = [](__half2 h) { return *reinterpret_cast<uint32_t *>(&h); };
  pk[0] = make_uint4(u(hi[0]), u(hi[1]), u(hi[2]), u(hi[3]));
  pk[1] = make_uint4(u(lo[0]), u(lo[1]), u(lo[2]), u(lo[3]));
  pk[2] = make_uint4(u(hi[4]), u(hi[5]), u(hi[6]), u(hi[7]));
  pk[3] = make_uint4(u(lo[4]), u(lo[5]), u(lo[6]), u(lo[7]));
  return m;
}

// SLOTS: the largest number of 16-column groups an epilogue thread owns (chi / 64).
// TILES: sample tiles of 128 rows per CTA.  With one column chunk (chi <= 128) nothing hides a step's epilogue, so the
// CTA carries two independent tiles (own operand, own TMEM columns, own barriers): the MMAs of one tile run under the
// epilogue of the other.  At chi = 256 one tile fills TMEM and shared memory, and the column chunks do the hiding.
// CG: 1 = every CTA multiplies its own tile; 2 = CTA pair (cluster of two CTAs on the two SMs of a TPC): the leader issues
// `tcgen05.mma.cta_group::2` for both tiles (M = 256), each CTA holds -- and loads -- only its half of the B columns of a
// weight stage, which takes the operand fetch of an MMA from 96 to 64 B/clk per SM (shared-memory bandwidth is what caps
// the MMA rate of the one-CTA kernel).  Epilogue, operand, TMEM accumulator and stash writer stay per CTA.
// NGRP: epilogue warps per lane quarter (4 lane quarters x NGRP column-group owners).  4 (608 threads, one CTA per SM) in
// general; 2 (352 threads, TWO CTAs per SM) for steps of at most 256 accumulator columns and bonds up to 64, where half of the
// 16 epilogue warps would own no column group and a step is a handful of small MMAs: the second CTA's chain runs in the
// latencies of the first (issue, commit, TMEM read, publish), which nothing inside one CTA could hide there.
template <int D, int SLOTS, int TILES, int CG, int NGRP = 4>
__global__ void __launch_bounds__((4 * NGRP + 3) * 32, NGRP == 4 ? 1 : 2) tc_chain_kernel(const __grid_constant__ TcArgs A, int adjoint) {
  static_assert(CG == 1 || TILES == 1, "the CTA pair carries one tile per CTA");
  static_assert(NGRP == 4 || NGRP == 2, "4 or 2 column-group owners per lane quarter");
  constexpr int kEpiWarps = 4 * NGRP, kMmaWarp = kEpiWarps;  // + 1: weight producer warp, + 2: image writer warp
  extern __shared__ __align__(1024) uint8_t smem[];
  const MpsArgs<float> &M = A.M;
  const int chi = M.chip, L = M.L, c = M.c, C = M.C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.y;
  const long long b0 = A.tile0 + (long long)blockIdx.x * (kTcRows * TILES);
  const int nR = L - 1 - c, nL = c;
  const int NH = A.NH, chi_h = chi / NH;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;  // 0: leader of the pair (issues the MMAs)

  // ---- the step program (identical in every warp role) ----
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
  // step -> weight image
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

  // step -> bond whose image is the operand this step's epilogue leaves in shared memory (-1: none / not stashed).
  // (The last right-chain step of the forward program replaces the operand by e0: its image is stored by the
  // epilogue threads themselves.)
  auto step_out_bond = [&](int step) -> int {
    if (!adjoint) {
      if (!A.img_stash) return -1;
      if (step < nR) return step == nR - 1 ? -1 : L - 1 - step;
      if (step < nR + nL) return step - nR + 1;
      return -1;
    }
    if (step < nseed - 1) return -1;
    if (step == nseed - 1) return half == 0 ? c : c + 1;
    const int i = step - nseed;
    return half == 0 ? c - 1 - i : c + 2 + i;
  };

  // A operand of tile t, in the layout of a bond image: unit (row, octet o, plane p) at
  // (row/8)*chi*32 + p*chi*16 + o*128 + (row%8)*16
  const uint32_t a_bytes = (uint32_t)kTcRows * chi * 4;
  uint8_t *a_op = smem;  // [TILES][a_bytes]
  uint8_t *stages = smem + (size_t)TILES * a_bytes;
  float *xch = reinterpret_cast<float *>(stages + (size_t)A.nstage * A.stage_bytes);  // [TILES][2][4][128]
  uint32_t *gcol = reinterpret_cast<uint32_t *>(xch + TILES * 2 * 4 * kTcRows);       // [16] group -> TMEM column
  uint32_t *gchunk = gcol + 16;                                                       // [16] group -> column chunk
  uint64_t *bars = reinterpret_cast<uint64_t *>(gchunk + 16);
  uint64_t *full = bars, *empty = bars + A.nstage;
  uint64_t *acc_full = bars + 2 * A.nstage;        // [TILES][kTcMaxNH]
  uint64_t *a_ready = acc_full + TILES * kTcMaxNH;  // [TILES]
  uint64_t *img_done = a_ready + TILES, *img_half = img_done + TILES, *a_rel = img_half + TILES;
  uint64_t *a_mma = a_rel + TILES;  // CTA pair: both CTAs' operands are published (waited on by the leader's MMA warp)
  uint64_t *pfull = a_mma + 1;      // CTA pair: [nstage] the follower's half of a weight stage has landed (relayed)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pfull + A.nstage);

  if (tid == 0) {
    for (int i = 0; i < A.nstage; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int t = 0; t < TILES; ++t) {
      for (int h = 0; h < NH; ++h) ptx::mbar_init(&acc_full[t * kTcMaxNH + h], 1);
      ptx::mbar_init(&a_ready[t], kEpiWarps);
      ptx::mbar_init(&img_done[t], 1);
      ptx::mbar_init(&img_half[t], 1);
      ptx::mbar_init(&a_rel[t], 1);
    }
    ptx::mbar_init(a_mma, kEpiWarps * CG);
    for (int i = 0; i < A.nstage; ++i) ptx::mbar_init(&pfull[i], 1);
    ptx::fence_mbar_init();
  }
  if (tid < 16) {
    const int n0 = 16 * tid;
    const int h = n0 < chi ? n0 / chi_h : 0;
    gchunk[tid] = (uint32_t)h;
    gcol[tid] = (uint32_t)(h * A.NMMA + (n0 - h * chi_h));
  }
  if (warp == kMmaWarp) {
    if (CG == 2) {
      ptx::tmem_alloc2(tmem_slot, (uint32_t)A.tmem_cols);
      ptx::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(tmem_slot, (uint32_t)A.tmem_cols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();  // the peer's barriers are initialised before anything arrives on them remotely
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tile_cols = (uint32_t)A.tmem_cols / TILES;  // TMEM columns of one tile's accumulator
  const int KS = chi / 16;
  // Early release: once the last column chunk's MMAs have consumed the K-slabs [0, krel), the operand columns
  // n < 16*krel (= the outputs of all earlier chunks) may be overwritten with the next operand.
  const int krel = (NH - 1) * (KS / NH);
  // K-slab order of a (step, chunk).  tcgen05 accumulates into TMEM with truncation (round toward zero): every MMA that adds
  // to an accumulator already at its final magnitude shrinks it by about half an ulp.  With a fixed slab order and a
  // near-identity site tensor (the reference's initialisation, mps.py:348-355) the SAME columns take the most truncations at
  // every site, and the bias grows linearly along the chain (tools/parity_report.py: 1.1e-3 at L = 196, chi = 256).
  // Rotating the order from step to step spreads the truncations evenly over the columns: a uniform relative shrink, which
  // the per-sample normalisation cancels.  The last chunk keeps its low slabs [0, krel) first (early operand release).
  // (one modulo per (step, chunk), outside the per-stage loops: the issuer warp is the critical path)
  auto slab_rot = [&](int step, int h, int &r_lo, int &r_hi) {
    const bool split = h == NH - 1 && krel > 0;
    r_lo = (5 * step + 3 * h) % (split ? krel : KS);
    r_hi = split ? (5 * step + 3 * h) % (KS - krel) : 0;
  };
  auto slab_of = [&](int h, int q, int r_lo, int r_hi) -> int {
    if (h == NH - 1 && krel > 0) {
      if (q < krel) { const int k = q + r_lo; return k >= krel ? k - krel : k; }
      const int k = q + r_hi;
      return k >= KS ? k - (KS - krel) : k;
    }
    const int k = q + r_lo;
    return k >= KS ? k - KS : k;
  };

  if (warp == kMmaWarp + 2) {
    // ===== image writer: operand generation g (published as phase g of a_ready) -> HBM stash, in paced pieces =====
    // Columns n < 16*krel first (they are overwritten early, see krel), then the rest.
    if (lane == 0) {
      uint8_t *base = adjoint ? A.gimg : A.fimg;
      const uint32_t lo_bytes = (uint32_t)krel * 256, hi_bytes = (uint32_t)(KS - krel) * 256;  // per (row block, plane)
      const unsigned pace = (unsigned)(A.pace_ns > 0 ? (krel > 0 ? A.pace_ns / 2 : A.pace_ns) : 0);
      for (int g = 0; g <= total; ++g) {
        const int bond = g > 0 ? step_out_bond(g - 1) : -1;
        const bool on = bond >= 0 && base && !TNB_DBG(2);
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          ptx::mbar_wait(&a_ready[t], g & 1);
          const uint8_t *src = a_op + (size_t)t * a_bytes;
          uint8_t *dst = on ? base + (size_t)bond * A.img_bond_bytes + (size_t)((b0 + t * kTcRows) >> 3) * chi * 32 : nullptr;
          if (on && krel > 0) {
            for (uint32_t blk = 0; blk < 2 * kTcRows / 8; ++blk) {
              ptx::bulk_s2g(dst + (size_t)blk * chi * 16, src + (size_t)blk * chi * 16, lo_bytes);
              ptx::bulk_commit();
              if (pace) __nanosleep(pace);
            }
            ptx::bulk_wait_read0();
          }
          ptx::mbar_arrive(&img_half[t]);  // columns n < 16*krel may be overwritten
          if (on && krel > 0) {
            for (uint32_t blk = 0; blk < 2 * kTcRows / 8; ++blk) {
              ptx::bulk_s2g(dst + (size_t)blk * chi * 16 + lo_bytes, src + (size_t)blk * chi * 16 + lo_bytes, hi_bytes);
              ptx::bulk_commit();
              if (pace) __nanosleep(pace);
            }
            ptx::bulk_wait_read0();
          } else if (on) {  // one column chunk: the tile is one contiguous run
            for (uint32_t off = 0; off < a_bytes; off += 4096) {
              ptx::bulk_s2g(dst + off, src + off, 4096);
              ptx::bulk_commit();
              if (pace) __nanosleep(pace);
            }
            ptx::bulk_wait_read0();
          }
          ptx::mbar_arrive(&img_done[t]);  // the whole operand may be overwritten
        }
      }
      ptx::bulk_wait0();
    }
  } else if (warp == kMmaWarp + 1) {
    // ===== weight producer: one bulk async copy per pipeline stage (= one K-slab of one column chunk); the step's
    //       image is streamed once per tile =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int step = 0; step < total; ++step) {
        const uint8_t *img = step_image(step);
        for (int t = 0; t < TILES; ++t) {
          for (int h = 0, hk = 0; h < NH; ++h) {
           int r_lo, r_hi;
           slab_rot(step, h, r_lo, r_hi);
           for (int q = 0; q < KS; ++q, ++hk) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            const uint32_t nbytes = TNB_DBG(16) ? 1024u : A.stage_bytes;  // (trace-build experiment: short loads, wrong results)
            ptx::mbar_expect_tx(&full[stage], nbytes);
            const int src = h * KS + slab_of(h, q, r_lo, r_hi);  // the q-th stage of chunk h holds K-slab slab_of(h, q)
            // (CTA pair: the stage image is [rank][plane][columns / 2]; each CTA fetches its own half)
            ptx::bulk_g2s(stages + (size_t)stage * A.stage_bytes, img + ((size_t)src * CG + rank) * A.stage_bytes,
                          nbytes, &full[stage]);
            if (++stage == A.nstage) { stage = 0; phase ^= 1; }
           }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer (CTA pair: the leader issues for both CTAs; the follower relays its `full` barriers) =====
    if (lane == 0 && rank != 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int step = 0; step < total; ++step)
        for (int hk = 0; hk < NH * KS; ++hk) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::mbar_arrive_cluster(&pfull[stage], 0);
          if (++stage == A.nstage) { stage = 0; phase ^= 1; }
        }
    } else if (CG == 1 || rank == 0) {
      // The whole warp runs the loop and waits; the instructions are issued under `elect.sync` (with `lane == 0` the
      // compiler wraps every tcgen05 instruction in an ELECT / BRA.U.ANY loop, which costs the issuer more than the
      // waits do).
      const uint32_t idesc = ptx::idesc_f16_f32(kTcRows * CG, A.NMMA);
      const uint32_t sbo_a = (uint32_t)chi * 32;
      const uint32_t lo_plane = (uint32_t)(A.NMMA / CG) * 32;  // B columns held by this CTA: NMMA / CG
      int stage = 0;
      uint32_t phase = 0;
#ifdef TNB_TC_TRACE
      long long tr_full = 0, tr_pfull = 0, tr_issue = 0;
#endif
      for (int step = 0; step < total; ++step) {
        const bool accum_step = adjoint && step > 0 && step < nseed;  // seed steps accumulate over k
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          const uint32_t a_hi_s = ptx::smem_u32(a_op + (size_t)t * a_bytes), a_lo_s = a_hi_s + (uint32_t)chi * 16;
          ptx::mbar_wait(CG == 2 ? a_mma : &a_ready[t], step & 1);
          ptx::tc_fence_after();
          if (t == 0) TNB_TRACE(3, step, 0);
          for (int h = 0; h < NH; ++h) {
            const uint32_t d = tmem_base + (uint32_t)t * tile_cols + (uint32_t)h * A.NMMA;
            int r_lo, r_hi;
            slab_rot(step, h, r_lo, r_hi);
            for (int q = 0; q < KS; ++q) {
              const int ks = slab_of(h, q, r_lo, r_hi);  // operand K-slab of the q-th stage of this chunk
              const long long c0 = TNB_CLOCK();
              ptx::mbar_wait(&full[stage], phase);
              const long long c1 = TNB_CLOCK();
              if (CG == 2 && !TNB_DBG(4)) ptx::mbar_wait(&pfull[stage], phase);
              const long long c2 = TNB_CLOCK();
              (void)c0; (void)c1; (void)c2;
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                const uint64_t da_hi = ptx::smem_desc_kmajor(a_hi_s + ks * 256, 128, sbo_a);
                const uint64_t da_lo = ptx::smem_desc_kmajor(a_lo_s + ks * 256, 128, sbo_a);
                const uint32_t st = ptx::smem_u32(stages + (size_t)stage * A.stage_bytes);
                const uint64_t db_hi = ptx::smem_desc_kmajor(st, 128, 256);
                const uint64_t db_lo = ptx::smem_desc_kmajor(st + lo_plane, 128, 256);
                if (CG == 2) {
                  ptx::mma_f16_pair(d, da_hi, db_hi, idesc, (q > 0 || accum_step) ? 1u : 0u);
                  ptx::mma_f16_pair(d, da_lo, db_hi, idesc, 1u);
                  ptx::mma_f16_pair(d, da_hi, db_lo, idesc, 1u);
                  ptx::mma_commit_pair(&empty[stage]);
                  if (h == NH - 1 && q == krel - 1) ptx::mma_commit_pair(&a_rel[t]);
                  if (q == KS - 1) ptx::mma_commit_pair(&acc_full[t * kTcMaxNH + h]);
                } else {
                  ptx::mma_f16(d, da_hi, db_hi, idesc, (q > 0 || accum_step) ? 1u : 0u);
                  ptx::mma_f16(d, da_lo, db_hi, idesc, 1u);
                  ptx::mma_f16(d, da_hi, db_lo, idesc, 1u);
                  ptx::mma_commit(&empty[stage]);
                  if (h == NH - 1 && q == krel - 1) ptx::mma_commit(&a_rel[t]);
                  if (q == KS - 1) ptx::mma_commit(&acc_full[t * kTcMaxNH + h]);
                }
              }
              __syncwarp();
              if (++stage == A.nstage) { stage = 0; phase ^= 1; }
#ifdef TNB_TC_TRACE
              if (q == 0) tr_full = tr_pfull = tr_issue = 0;
              tr_full += c1 - c0; tr_pfull += c2 - c1; tr_issue += TNB_CLOCK() - c2;
#endif
            }
#ifdef TNB_TC_TRACE
            if (t == 0 && h < 2) {  // per column chunk: cycles waiting for the own stage / the peer's half / issuing
              TNB_TRACE_SET(3, step, 4 + 3 * h, tr_full);
              TNB_TRACE_SET(3, step, 5 + 3 * h, tr_pfull);
              TNB_TRACE_SET(3, step, 6 + 3 * h, tr_issue);
            }
#endif
            if (t == 0) TNB_TRACE(3, step, 1 + h);
          }
        }
      }
    }
  } else {
    // ===== epilogue: NGRP threads per sample row (one per warp group), each owning every NGRP-th 16-column group =====
    const int lq = warp & 3, grp = warp >> 2;
    const int row = lq * 32 + lane;
    const bool lead = grp == 0;  // the thread of the row that does the per-sample bookkeeping
    const uint32_t lo_off = (uint32_t)chi * 16;  // lo plane of a unit, in the operand and in the bond images
    const int nmine = (chi / 16 - grp + NGRP - 1) / NGRP;  // my 16-column groups: j = NGRP*i + grp, i < nmine (<= SLOTS)
    const int nslab = L - 1 + C;
    int step = 0;
#ifdef TNB_TC_TRACE
    const int tw = warp == 0 ? 0 : (warp == 15 ? 1 : (warp == 6 ? 2 : 63));  // traced epilogue warps (63: not recorded)
#endif
    // per-tile state of this thread (its row in each tile)
    struct Tile {
      long long b;
      size_t bs;
      bool valid;
      uint32_t trow;
      uint8_t *a_row;
      float m_op;    // max-abs of the current operand row (operand units)
      int ea, eb;    // forward: exponent sums of the left / right chain; adjoint: running exponent h (ea)
    } T[TILES];
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
      T[t].b = b0 + t * kTcRows + row;
      T[t].valid = T[t].b < M.B;
      T[t].bs = (size_t)(T[t].valid ? T[t].b : 0);
      T[t].trow = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)t * tile_cols;
      T[t].a_row = a_op + (size_t)t * a_bytes + (uint32_t)(row >> 3) * (uint32_t)chi * 32 + (uint32_t)(row & 7) * 16;
      T[t].m_op = 16384.f;  // the boundary vector is e0 * 2^14
      T[t].ea = T[t].eb = 0;
    }

    // a sample's 16-byte units inside the image of `bond` (see TcArgs)
    auto img_row = [&](uint8_t *base, int bond, long long b) -> uint8_t * {
      const int slot = A.img_stash ? bond : (bond == c + 1 ? 0 : -1);
      if (slot < 0 || base == nullptr) return nullptr;
      return base + (size_t)slot * A.img_bond_bytes + tc_img_row(b, chi);
    };
    // packed group j -> the unit row `dst` (operand row in smem or image row in HBM)
    auto store_pk = [&](int j, const uint4(&pk)[4], uint8_t *dst) {
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        *reinterpret_cast<uint4 *>(dst + (2 * j + o) * 128) = pk[2 * o];
        *reinterpret_cast<uint4 *>(dst + lo_off + (2 * j + o) * 128) = pk[2 * o + 1];
      }
    };
    // 16 values of an image row: hi + lo (exact in fp32), still carrying the 2^14 image scale
    auto load_img16 = [&](const uint8_t *irow, int j, float(&v)[16]) {
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const uint4 hi = *reinterpret_cast<const uint4 *>(irow + (2 * j + o) * 128);
        const uint4 lo = *reinterpret_cast<const uint4 *>(irow + lo_off + (2 * j + o) * 128);
        const __half2 *h = reinterpret_cast<const __half2 *>(&hi), *l = reinterpret_cast<const __half2 *>(&lo);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fh = __half22float2(h[q]), fl = __half22float2(l[q]);
          v[8 * o + 2 * q] = fh.x + fl.x;
          v[8 * o + 2 * q + 1] = fh.y + fl.y;
        }
      }
    };
    // boundary vector e0 * 2^14 into a unit row (my column groups only); `one` = false writes zeros
    auto write_e0 = [&](uint8_t *dst, bool one) {
      if (!dst) return;
#pragma unroll
      for (int i = 0; i < SLOTS; ++i) {
        if (i >= nmine) continue;
        const int j = NGRP * i + grp;
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          uint4 z = make_uint4(0, 0, 0, 0), zh = z;
          if (2 * j + o == 0 && one) zh.x = 0x00007400u;  // fp16(2^14) in element 0
          *reinterpret_cast<uint4 *>(dst + (2 * j + o) * 128) = zh;
          *reinterpret_cast<uint4 *>(dst + lo_off + (2 * j + o) * 128) = z;
        }
      }
    };
    auto write_operand_from_image = [&](const uint8_t *irow, float scale, uint8_t *a_row) {
#pragma unroll
      for (int i = 0; i < SLOTS; ++i) {
        if (i >= nmine) continue;
        const int j = NGRP * i + grp;
        float w[16];
        if (irow) load_img16(irow, j, w);
        else {
#pragma unroll
          for (int q = 0; q < 16; ++q) w[q] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q] *= scale;
        uint4 pk[4];
        split8(w, pk[0], pk[1]);
        split8(w + 8, pk[2], pk[3]);
        store_pk(j, pk, a_row);
      }
    };
    // the image writer has finished reading tile t's operand of this step (it may be overwritten)
    auto operand_free = [&](int t) { ptx::mbar_wait(&img_done[t], step & 1); };
    // operand (and TMEM reads) done -> the MMA warp may start the tile's next step.  Every warp first makes sure the
    // image writer has passed this step, so that no barrier phase can run more than one step ahead of a waiter.
    auto publish = [&](int t, bool first = false) {
      if (!first) operand_free(t);
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&a_ready[t]);                       // this CTA's image writer (and, one CTA, its MMA warp)
        if (CG == 2) ptx::mbar_arrive_cluster(a_mma, 0);     // the pair's MMA issuer
      }
    };
    // combine one float per thread across the 4 threads of a row (through smem, fixed order)
    auto row_exchange = [&](int t, float v) -> const float * {
      float *x = xch + (t * 2 + (step & 1)) * (4 * kTcRows);
      x[grp * kTcRows + row] = v;
      ptx::named_bar_sync(1 + lq, NGRP * 32);
      return x + row;
    };
    auto row_max = [&](int t, float v) -> float {
      const float *x = row_exchange(t, v);
      float m = fmaxf(x[0], x[kTcRows]);
      if (NGRP == 4) m = fmaxf(m, fmaxf(x[2 * kTcRows], x[3 * kTcRows]));
      return m;
    };
    // Scale of a step whose outputs are bounded by `bnd`: 2^(14 - e), bnd = f * 2^e, f in [0.5, 1).  *e_out = e.
    auto bound_scale = [&](float bnd, bool valid, int *e_out) -> float {
      const int ef = (int)((__float_as_uint(bnd) >> 23) & 0xffu);
      const bool ok = valid && ef >= 16 && ef < 255;  // (a zero row, or one below 2^-110: flushed)
      *e_out = ok ? ef - 126 : 0;
      return ok ? __uint_as_float((uint32_t)(127 + 126 + kTcKA - ef) << 23) : 0.f;
    };
    // Wait for the column chunks holding my groups of tile t, contract them with phi * sc, split to fp16 hi/lo and write
    // the row to `dst`: an image row in HBM (every group at once), or the next operand row in smem.  There a group of
    // the last chunk is stored at once (all MMAs of the step are done), an earlier group is parked in the TMEM columns
    // it came from and moved to smem as soon as the last chunk's MMAs have released those operand columns (a_rel) --
    // long before the last chunk itself is complete.  Returns the max-abs of my part of the scaled row.
    auto gather_pack = [&](int t, const float(&ph)[D], float sc, uint8_t *dst) -> float {
      float phs[D];
#pragma unroll
      for (int s = 0; s < D; ++s) phs[s] = ph[s] * sc;
      uint8_t *a_row = T[t].a_row;
      const uint32_t trow = T[t].trow;
      const bool to_op = dst == a_row;
      uint64_t *accf = acc_full + t * kTcMaxNH;
      auto move_parked = [&]() {
        ptx::tmem_st_wait();
        ptx::mbar_wait(&a_rel[t], step & 1);
        ptx::mbar_wait(&img_half[t], step & 1);
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
          if (i >= nmine) continue;
          const int j = NGRP * i + grp;
          if ((int)gchunk[j] == NH - 1) continue;
          uint4 pk[4];
          ptx::tmem_ld16u(trow + gcol[j], pk);
          ptx::tmem_ld_wait();
          store_pk(j, pk, a_row);
        }
        if (t == 0) TNB_TRACE(tw, step, 7);
      };
      float m = 0.f;
      int hprev = -1;
      bool moved = !to_op || NH == 1;
#pragma unroll
      for (int i = 0; i < SLOTS; ++i) {
        if (i >= nmine) continue;
        const int j = NGRP * i + grp;
        const int h = (int)gchunk[j];
        if (h != hprev) {
          if (h == NH - 1 && !moved) {
            move_parked();
            moved = true;
          }
          ptx::mbar_wait(&accf[h], step & 1);
          ptx::tc_fence_after();
          if (h == NH - 1 && to_op) operand_free(t);
          hprev = h;
          if (t == 0) TNB_TRACE(tw, step, h);
        }
        uint4 pk[4];
        m = fmaxf(m, tc_combine_pack16<D>(trow + gcol[j], (uint32_t)chi_h, phs, pk));
        if (!to_op) {
          if (dst) store_pk(j, pk, dst);
        } else if (h == NH - 1) {
          store_pk(j, pk, a_row);
        } else {
          ptx::tmem_st16(trow + gcol[j], pk);
        }
      }
      if (!moved) move_parked();
      if (hprev != NH - 1) {  // every thread has seen the whole step complete before it publishes
        ptx::mbar_wait(&accf[NH - 1], step & 1);
        ptx::tc_fence_after();
      }
      if (t == 0) TNB_TRACE(tw, step, 8);
      return m;
    };
    // One chain step of tile t.  mode 0: forward (records delta)   mode 1: adjoint (records wq/qmax).
    // `hexp` is the forward exponent sum (mode 0) or the running exponent h of the adjoint (mode 1); lead thread only.
    // next_op: 1 = the scaled row becomes the next operand (the image writer stashes it), 2 = the boundary vector e0
    // does and the row only goes to its image row `irow`, written here.
    auto chain_step = [&](int t, int site, int slab, int dir, int mode, int &hexp, int next_op, uint8_t *irow) {
      const bool valid = T[t].valid;
      const long long b = T[t].b;
      float ph[D];
      embed_site_d<D>(M.emb, valid ? M.x + (T[t].bs * L + site) * M.emb.n_in : nullptr, ph);
      const int dsite = (mode == 1 && valid && lead) ? M.exps[(size_t)site * M.B + b] : 0;
      const int kw = A.kW[slab];
      int e;
      const float sc = bound_scale(T[t].m_op * A.colsum[dir * nslab + slab], valid, &e);
      const bool nz = sc != 0.f;
      const float m = gather_pack(t, ph, sc, next_op == 2 ? irow : T[t].a_row);
      if (next_op == 2) {
        operand_free(t);
        write_e0(T[t].a_row, true);
      }
      if (t == 0) TNB_TRACE(tw, step, 11);
      publish(t);
      if (t == 0) TNB_TRACE(tw, step, 12);
      // off the critical path (the tile's next MMAs are running): row maximum of the new operand, bookkeeping
      T[t].m_op = next_op == 2 ? 16384.f : row_max(t, m);
      if (lead) {
        const int dtrue = nz ? e - kTcKA - kw : 0;  // exponent in units of the true recurrence
        if (mode == 0) {
          hexp += dtrue;
          if (valid && M.exps) M.exps[(size_t)site * M.B + b] = dtrue;
        } else {
          // dA_site = sum_b 2^q (F x phi x G), q = h_in - delta_site, h_in = exponent of the input adjoint
          const int q = hexp - dsite;
          if (valid) {
#pragma unroll
            for (int s = 0; s < D; ++s) A.wq[((size_t)site * D + s) * M.B + b] = ph[s] * ldexpf(1.f, q);
          }
          int qm = valid ? q : INT_MIN;
          for (int off = 16; off > 0; off >>= 1) qm = max(qm, __shfl_xor_sync(0xffffffffu, qm, off));
          if (lane == 0 && qm != INT_MIN) atomicMax(&A.qmax[site], qm);
          hexp += dtrue - dsite;
        }
      }
    };

    if (!adjoint) {
      // ---------------- forward program ----------------
#pragma unroll
      for (int t = 0; t < TILES; ++t) {
        write_e0(T[t].a_row, true);
        write_e0(img_row(A.fimg, L, T[t].b), T[t].valid);
        write_e0(img_row(A.fimg, 0, T[t].b), T[t].valid);
        publish(t, true);
      }
      for (int i = 0; i < nR; ++i) {
        const int site = L - 1 - i;
#pragma unroll
        for (int t = 0; t < TILES; ++t)
          chain_step(t, site, mps_slab(M, site), 1, 0, T[t].eb, i == nR - 1 ? 2 : 1, img_row(A.fimg, site, T[t].b));
        ++step;
      }
      for (int i = 0; i < nL; ++i) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) chain_step(t, i, mps_slab(M, i), 0, 0, T[t].ea, 1, nullptr);
        ++step;
      }
      // ---------------- class site + head ----------------
      const float ysc = ldexpf(1.f, -(2 * kTcKA + A.kW[c]));
      float y[TILES][kMaxOut];
      float phc[TILES][D];
#pragma unroll
      for (int t = 0; t < TILES; ++t) embed_site_d<D>(M.emb, T[t].valid ? M.x + (T[t].bs * L + c) * M.emb.n_in : nullptr, phc[t]);
      for (int k = 0; k < C; ++k) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          const uint8_t *fr = img_row(A.fimg, c + 1, T[t].b);  // each thread wrote its own column groups of that row
          uint64_t *accf = acc_full + t * kTcMaxNH;
          float acc = 0.f;
          int hprev = -1;
#pragma unroll
          for (int i = 0; i < SLOTS; ++i) {
            if (i >= nmine) continue;
            const int j = NGRP * i + grp;
            float f[16];
            if (fr) load_img16(fr, j, f);
            else {
#pragma unroll
              for (int q = 0; q < 16; ++q) f[q] = 0.f;
            }
            const int h = (int)gchunk[j];
            if (h != hprev) {
              ptx::mbar_wait(&accf[h], step & 1);
              ptx::tc_fence_after();
              hprev = h;
            }
            float o[16];
            tc_combine16<D>(T[t].trow + gcol[j], (uint32_t)chi_h, phc[t], o);
#pragma unroll
            for (int q = 0; q < 16; ++q) acc = fmaf(o[q], f[q], acc);
          }
          if (hprev != NH - 1) {
            ptx::mbar_wait(&accf[NH - 1], step & 1);
            ptx::tc_fence_after();
          }
          const float *x = row_exchange(t, acc);
          y[t][k] = (NGRP == 4 ? ((x[0] + x[kTcRows]) + (x[2 * kTcRows] + x[3 * kTcRows])) : (x[0] + x[kTcRows])) * ysc;
          publish(t);
        }
        ++step;
      }
      if (lead) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          const long long b = T[t].b;
          float lval = 0.f;
          if (T[t].valid) {
            M.esum[b] = T[t].ea;
            M.esum[M.B + b] = T[t].eb;
            float tl[kMaxOut];
            for (int k = 0; k < C; ++k) {
              tl[k] = M.targets ? M.targets[(size_t)b * C + k] : 0.f;
              M.y[(size_t)b * C + k] = y[t][k];
              if (M.out) M.out[(size_t)b * C + k] = y[t][k];
            }
            const int E = T[t].ea + T[t].eb;
            if (M.out_exp) M.out_exp[b] = E;
            lval = mps_head<float>(M.head, C, y[t], E, tl, M.cw, (float *)nullptr);
            if (M.loss) M.loss[b] = lval;
          }
          if (M.loss_sum) {  // one atomic per warp
            double ls = (double)lval;
            for (int off = 16; off > 0; off >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, off);
            if (lane == 0) atomicAdd(M.loss_sum, ls);
          }
        }
      }
    } else {
      // ---------------- adjoint program ----------------
      float dy[TILES][kMaxOut];
      float phc[TILES][D];
      float bnd[TILES], dsc[TILES];
      int ed[TILES];
      const uint8_t *vrow[TILES];
#pragma unroll
      for (int t = 0; t < TILES; ++t) {
        const bool valid = T[t].valid;
        const long long b = T[t].b;
        // head derivative dLoss/dy' (times loss_scale) for this sample
        {
          float yl[kMaxOut], tl[kMaxOut];
          for (int k = 0; k < C; ++k) {
            yl[k] = valid ? M.y[(size_t)b * C + k] : 1.f;
            tl[k] = (valid && M.targets) ? M.targets[(size_t)b * C + k] : (k == 0 ? 1.f : 0.f);
          }
          const int E = valid ? M.esum[b] + M.esum[M.B + b] : 0;
          mps_head<float>(M.head, C, yl, E, tl, M.cw, dy[t]);
          for (int k = 0; k < C; ++k) dy[t][k] = valid ? dy[t][k] * M.loss_scale : 0.f;
        }
        embed_site_d<D>(M.emb, valid ? M.x + (T[t].bs * L + c) * M.emb.n_in : nullptr, phc[t]);
        float dmax = 0.f;
        for (int k = 0; k < C; ++k) dmax = fmaxf(dmax, fabsf(dy[t][k]));
        ed[t] = 0;
        if (dmax > 0.f) frexpf(dmax, &ed[t]);
        if (valid && lead) {
          // class-site dA weights: wqc[(s,k)][b] = phi_s * dy'_k ; both halves write identical values
          int qm = INT_MIN;
#pragma unroll
          for (int s = 0; s < D; ++s)
            for (int k = 0; k < C; ++k) {
              const float w = phc[t][s] * dy[t][k];
              A.wqc[((size_t)s * C + k) * M.B + b] = w;
              if (w != 0.f) { int ew; frexpf(w, &ew); qm = max(qm, ew); }
            }
          for (int k = 0; k < C; ++k) M.dy[(size_t)b * C + k] = dy[t][k];
          if (qm != INT_MIN) atomicMax(A.qmaxc, qm);
        }
        vrow[t] = img_row(A.fimg, half == 0 ? c + 1 : c, b);  // forward environment row (operand units)
        dsc[t] = valid ? ldexpf(1.f, -ed[t]) : 0.f;           // |dy_k * dsc| < 1
        // bound of the accumulated seed: sum_k |dy_k dsc| * max|row| * colsum(slab c+k)
        float mrow = 0.f;
        if (vrow[t]) {
#pragma unroll
          for (int i = 0; i < SLOTS; ++i) {
            if (i >= nmine) continue;
            float w[16];
            load_img16(vrow[t], NGRP * i + grp, w);
#pragma unroll
            for (int q = 0; q < 16; ++q) mrow = fmaxf(mrow, fabsf(w[q]));
          }
        }
        mrow = row_max(t, mrow);
        float bs_ = 0.f;
        for (int k = 0; k < C; ++k) bs_ += fabsf(dy[t][k] * dsc[t]) * A.colsum[(half == 0 ? nslab : 0) + c + k];
        bnd[t] = bs_ * mrow * 1.001f;
        write_operand_from_image(vrow[t], dy[t][0] * dsc[t], T[t].a_row);
        publish(t, true);
      }
      const int jb = half == 0 ? 0 : L - 1;  // the boundary site: no chain step, but it has a gradient
      for (int k = 0; k < C; ++k) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          if (k < C - 1) {
            ptx::mbar_wait(&acc_full[t * kTcMaxNH + NH - 1], step & 1);  // MMAs of step k have finished reading the operand
            ptx::tc_fence_after();
            operand_free(t);
            write_operand_from_image(vrow[t], dy[t][k + 1] * dsc[t], T[t].a_row);
            publish(t);
          } else {
            // all C slabs accumulated: scale -> seed G[c] / G[c+1] (operand of the first chain step; the writer stashes it)
            int e;
            const float sc = bound_scale(bnd[t], T[t].valid, &e);
            const float m = gather_pack(t, phc[t], sc, T[t].a_row);
            T[t].ea = (sc != 0.f ? e - kTcKA - A.kW[c] : 0) + ed[t];  // G = row * 2^h (the operand carried dy * 2^-ed)
            publish(t);
            T[t].m_op = row_max(t, m);
          }
        }
        ++step;
      }
      for (int i = 0; i < nchain; ++i) {
        const int site = half == 0 ? c - 1 - i : c + 1 + i;
#pragma unroll
        for (int t = 0; t < TILES; ++t) chain_step(t, site, mps_slab(M, site), half == 0 ? 1 : 0, 1, T[t].ea, 1, nullptr);
        ++step;
      }
      // boundary site weights
      if (lead) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          int qm = INT_MIN;
          if (T[t].valid) {
            const long long b = T[t].b;
            const int q = T[t].ea - M.exps[(size_t)jb * M.B + b];
            float ph[D];
            embed_site_d<D>(M.emb, M.x + ((size_t)b * L + jb) * M.emb.n_in, ph);
#pragma unroll
            for (int s = 0; s < D; ++s) A.wq[((size_t)jb * D + s) * M.B + b] = ph[s] * ldexpf(1.f, q);
            qm = q;
          }
          for (int off = 16; off > 0; off >>= 1) qm = max(qm, __shfl_xor_sync(0xffffffffu, qm, off));
          if (lane == 0 && qm != INT_MIN) atomicMax(&A.qmax[jb], qm);
        }
      }
    }
  }

  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();  // both CTAs are done with each other's barriers and TMEM
  else __syncthreads();
  ptx::tc_fence_after();
  if (warp == kMmaWarp) {
    if (CG == 2) ptx::tmem_dealloc2(tmem_base, (uint32_t)A.tmem_cols);
    else ptx::tmem_dealloc(tmem_base, (uint32_t)A.tmem_cols);
  }
}

}  // namespace tnb
