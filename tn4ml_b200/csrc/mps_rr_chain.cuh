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
// CTA: 8 compute warps (one 128-sample tile, like the tcgen05 kernel) + 1 producer warp that streams one site image
// (D * 4 KB) per step through a ring of bulk async copies; two CTAs per SM for D = 2.
// Programs, bookkeeping (exps / esum / y / dy / wq / qmax / wqc / qmaxc) and results are those of tc_chain_kernel.
#pragma once
#include "mps_tc.cuh"

namespace tnb {

constexpr int kRrcWarps = 8;                         // compute warps: 16 samples each
constexpr int kRrcThreads = (kRrcWarps + 1) * 32;    // + weight producer warp
constexpr int kRrcStages = 4;                        // site images in flight

__device__ __forceinline__ void rrc_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void rrc_ldsm4(uint32_t saddr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(saddr));
}
// two scaled floats -> fp16x2 hi and lo words
__device__ __forceinline__ void rrc_split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ float2 rrc_join2(uint32_t hi, uint32_t lo) {
  const float2 h = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
  const float2 l = __half22float2(*reinterpret_cast<const __half2 *>(&lo));
  return make_float2(h.x + l.x, h.y + l.y);
}

template <int D>
__global__ void __launch_bounds__(kRrcThreads, D == 2 ? 2 : 1) rr_chain_kernel(TcArgs A, int adjoint) {
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

  uint8_t *ring = smem;  // [kRrcStages][kSite]
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)kRrcStages * kSite), *empty = full + kRrcStages;
  if (tid == 0) {
    for (int i = 0; i < kRrcStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], kRrcWarps);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == kRrcWarps) {
    // ===== weight producer: one bulk async copy (the whole site image) per step =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int step = 0; step < total; ++step) {
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full[stage], kSite);
        ptx::bulk_g2s(ring + (size_t)stage * kSite, step_image(step), kSite, &full[stage]);
        if (++stage == kRrcStages) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }

  // ===== compute warps =====
  const int g = lane >> 2, t = lane & 3;
  const long long row0 = A.tile0 + (long long)blockIdx.x * kTcRows + warp * 16;  // a multiple of 8
  const long long bA = row0 + g, bB = bA + 8;                                    // this thread's two sample rows
  const bool vA = bA < M.B, vB = bB < M.B;
  const bool headlane = t < 2;                     // lane t == 0 keeps the books of row A, t == 1 of row B
  const long long bH = t == 0 ? bA : bB;
  const bool vH = headlane && (t == 0 ? vA : vB);
  const long long bE = (t & 1) ? bB : bA;          // the row whose feature map this lane evaluates (shared by shuffle)
  const bool vE = (t & 1) ? vB : vA;
  const int srcA = lane & ~3, srcB = srcA + 1;
  // byte offset of this thread's word in the image of a bond: unit (row, octet o, plane p) at (row/8)*1024 + p*512 + o*128 + (row%8)*16
  const size_t img_off = (size_t)(row0 >> 3) * 1024 + (size_t)lane * 4;
  const uint32_t ring_s = ptx::smem_u32(ring);
  // ldmatrix.x4 of one (k-step, n-tile): matrices {hi k-octet 0, hi k-octet 1, lo k-octet 0, lo k-octet 1}
  const uint32_t ld_off = (uint32_t)(lane & 7) * 16 + (uint32_t)((lane >> 3) & 1) * 128 + (uint32_t)(lane >> 4) * kLoPlane;

  uint32_t ahi[2][4], alo[2][4];  // operand row block v[16, 32] as A fragments of the two k-steps
  float acc[NT][4];
  int stage = 0;
  uint32_t phase = 0;

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
  auto zero_acc = [&]() {
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  };
  // acc += operand x site image of the ring's current stage
  auto mma_step = [&]() {
    ptx::mbar_wait(&full[stage], phase);
    const uint32_t st = ring_s + (uint32_t)stage * kSite + ld_off;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t bh0, bh1, bl0, bl1;
        rrc_ldsm4(st + (uint32_t)kk * kStage + (uint32_t)j * 256, bh0, bh1, bl0, bl1);
        rrc_mma(acc[j], ahi[kk], bh0, bh1);
        rrc_mma(acc[j], alo[kk], bh0, bh1);
        rrc_mma(acc[j], ahi[kk], bl0, bl1);
      }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[stage]);
    if (++stage == kRrcStages) { stage = 0; phase ^= 1; }
  };
  // feature maps of the warp's rows at `site`: every lane evaluates one row, the quad shares them
  auto phis = [&](int site, float (&pa)[D], float (&pb)[D]) {
    float ph[D];
    embed_site_d<D>(M.emb, vE ? M.x + ((size_t)bE * L + site) * M.emb.n_in : nullptr, ph);
#pragma unroll
    for (int s = 0; s < D; ++s) {
      pa[s] = __shfl_sync(0xffffffffu, ph[s], srcA);
      pb[s] = __shfl_sync(0xffffffffu, ph[s], srcB);
    }
  };
  // o[j][i] = sum_s phi_s acc[s*4 + j][i]   (i = 0, 1: row A; 2, 3: row B)
  auto contract = [&](const float (&pa)[D], const float (&pb)[D], float (&o)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j][0] = pa[0] * acc[j][0];
      o[j][1] = pa[0] * acc[j][1];
      o[j][2] = pb[0] * acc[j][2];
      o[j][3] = pb[0] * acc[j][3];
#pragma unroll
      for (int s = 1; s < D; ++s) {
        o[j][0] = fmaf(pa[s], acc[s * 4 + j][0], o[j][0]);
        o[j][1] = fmaf(pa[s], acc[s * 4 + j][1], o[j][1]);
        o[j][2] = fmaf(pb[s], acc[s * 4 + j][2], o[j][2]);
        o[j][3] = fmaf(pb[s], acc[s * 4 + j][3], o[j][3]);
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
  // contracted rows -> normalised fp16 hi/lo operand; eA / eB: the rows' exponents (0 and a zero row if flushed)
  auto renorm = [&](const float (&o)[4][4], int &eA, int &eB, bool &nzA, bool &nzB) {
    float mA = 0.f, mB = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mA = fmaxf(mA, fmaxf(fabsf(o[j][0]), fabsf(o[j][1])));
      mB = fmaxf(mB, fmaxf(fabsf(o[j][2]), fabsf(o[j][3])));
    }
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
    const float sA = row_scale(mA, vA, &eA), sB = row_scale(mB, vB, &eB);
    nzA = sA != 0.f;
    nzB = sB != 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      rrc_split2(o[2 * kk][0] * sA, o[2 * kk][1] * sA, ahi[kk][0], alo[kk][0]);
      rrc_split2(o[2 * kk][2] * sB, o[2 * kk][3] * sB, ahi[kk][1], alo[kk][1]);
      rrc_split2(o[2 * kk + 1][0] * sA, o[2 * kk + 1][1] * sA, ahi[kk][2], alo[kk][2]);
      rrc_split2(o[2 * kk + 1][2] * sB, o[2 * kk + 1][3] * sB, ahi[kk][3], alo[kk][3]);
    }
  };
  auto sel = [&](int a, int b) -> int { return t == 0 ? a : b; };  // the head lane's row

  if (!adjoint) {
    // ---------------- forward program ----------------
    int eaA = 0, eaB = 0, ebA = 0, ebB = 0;  // exponent sums of the left / right chain
    set_e0(vA, vB);
    store_image(img_slot(A.fimg, L));
    store_image(img_slot(A.fimg, 0));
    set_e0(true, true);
    for (int i = 0; i < nR; ++i) {
      const int site = L - 1 - i, kw = A.kW[mps_slab(M, site)];
      float pa[D], pb[D], o[4][4];
      phis(site, pa, pb);
      zero_acc();
      mma_step();
      contract(pa, pb, o);
      int eA, eB;
      bool nzA, nzB;
      renorm(o, eA, eB, nzA, nzB);
      const int dA = nzA ? eA - kTcKA - kw : 0, dB = nzB ? eB - kTcKA - kw : 0;
      ebA += dA;
      ebB += dB;
      if (vH) M.exps[(size_t)site * M.B + bH] = sel(dA, dB);
      store_image(img_slot(A.fimg, site));
      if (i == nR - 1) set_e0(true, true);  // F[c+1] is in its image; the left chain starts from the boundary vector
    }
    for (int i = 0; i < nL; ++i) {
      const int kw = A.kW[i];
      float pa[D], pb[D], o[4][4];
      phis(i, pa, pb);
      zero_acc();
      mma_step();
      contract(pa, pb, o);
      int eA, eB;
      bool nzA, nzB;
      renorm(o, eA, eB, nzA, nzB);
      const int dA = nzA ? eA - kTcKA - kw : 0, dB = nzB ? eB - kTcKA - kw : 0;
      eaA += dA;
      eaB += dB;
      if (vH) M.exps[(size_t)i * M.B + bH] = sel(dA, dB);
      store_image(img_slot(A.fimg, i + 1));
    }
    // ---------------- class site + head ----------------
    {
      const float ysc = ldexpf(1.f, -(2 * kTcKA + A.kW[c]));
      float pa[D], pb[D];
      phis(c, pa, pb);
      // F[c+1] of this thread's elements, from its image (written by this thread: at the start if c + 1 == L)
      float2 fr[2][4];
      {
        const uint8_t *src = img_slot(A.fimg, c + 1);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint8_t *p = src + (size_t)(i & 1) * 1024 + (size_t)(2 * kk + (i >> 1)) * 128;
            fr[kk][i] = rrc_join2(*reinterpret_cast<const uint32_t *>(p), *reinterpret_cast<const uint32_t *>(p + 512));
          }
      }
      for (int k = 0; k < C; ++k) {
        float o[4][4];
        zero_acc();
        mma_step();
        contract(pa, pb, o);
        float yA = 0.f, yB = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          yA = fmaf(o[2 * kk][0], fr[kk][0].x, fmaf(o[2 * kk][1], fr[kk][0].y, yA));
          yB = fmaf(o[2 * kk][2], fr[kk][1].x, fmaf(o[2 * kk][3], fr[kk][1].y, yB));
          yA = fmaf(o[2 * kk + 1][0], fr[kk][2].x, fmaf(o[2 * kk + 1][1], fr[kk][2].y, yA));
          yB = fmaf(o[2 * kk + 1][2], fr[kk][3].x, fmaf(o[2 * kk + 1][3], fr[kk][3].y, yB));
        }
        yA += __shfl_xor_sync(0xffffffffu, yA, 1);
        yB += __shfl_xor_sync(0xffffffffu, yB, 1);
        yA += __shfl_xor_sync(0xffffffffu, yA, 2);
        yB += __shfl_xor_sync(0xffffffffu, yB, 2);
        if (vH) M.y[(size_t)bH * C + k] = (t == 0 ? yA : yB) * ysc;
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
        float ph[D];
        embed_site_d<D>(M.emb, M.x + ((size_t)bH * L + c) * M.emb.n_in, ph);
        int qm = INT_MIN;
#pragma unroll
        for (int s = 0; s < D; ++s)
          for (int k = 0; k < C; ++k) {
            const float w = ph[s] * dyl[k];
            A.wqc[((size_t)s * C + k) * M.B + bH] = w;
            if (w != 0.f) { int ew; frexpf(w, &ew); qm = max(qm, ew); }
          }
        for (int k = 0; k < C; ++k) M.dy[(size_t)bH * C + k] = dyl[k];
        if (qm != INT_MIN) atomicMax(A.qmaxc, qm);
      }
    }
    __syncwarp();
    const int edA = __shfl_sync(0xffffffffu, edH, srcA), edB = __shfl_sync(0xffffffffu, edH, srcB);
    const float dscA = vA ? ldexpf(1.f, -edA) : 0.f, dscB = vB ? ldexpf(1.f, -edB) : 0.f;  // |dy_k * dsc| < 1
    float pa[D], pb[D];
    phis(c, pa, pb);
    // forward environment rows (operand units) of this thread's elements
    float2 vr[2][4];
    {
      const uint8_t *src = img_slot(A.fimg, half == 0 ? c + 1 : c);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint8_t *p = src + (size_t)(i & 1) * 1024 + (size_t)(2 * kk + (i >> 1)) * 128;
          vr[kk][i] = rrc_join2(*reinterpret_cast<const uint32_t *>(p), *reinterpret_cast<const uint32_t *>(p + 512));
        }
    }
    zero_acc();
    for (int k = 0; k < C; ++k) {
      // seed operand of slab c + k: F row * dy_k * 2^-ed (rows normalised to 2^14, |factor| < 1)
      const float fA = vA ? M.dy[(size_t)bA * C + k] * dscA : 0.f, fB = vB ? M.dy[(size_t)bB * C + k] * dscB : 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float f = (i & 1) ? fB : fA;
          rrc_split2(vr[kk][i].x * f, vr[kk][i].y * f, ahi[kk][i], alo[kk][i]);
        }
      mma_step();  // accumulates over k
    }
    int hA, hB;  // running exponents of the adjoint rows
    {
      float o[4][4];
      contract(pa, pb, o);
      int eA, eB;
      bool nzA, nzB;
      renorm(o, eA, eB, nzA, nzB);
      const int kw = A.kW[c];
      hA = (nzA ? eA - kTcKA - kw : 0) + edA;  // G = row * 2^h (the operand carried dy * 2^-ed)
      hB = (nzB ? eB - kTcKA - kw : 0) + edB;
      store_image(img_slot(A.gimg, half == 0 ? c : c + 1));
    }
    for (int i = 0; i < nchain; ++i) {
      const int site = half == 0 ? c - 1 - i : c + 1 + i;
      const int kw = A.kW[mps_slab(M, site)];
      float qa[D], qb[D], o[4][4];
      phis(site, qa, qb);
      const int dsA = vA ? M.exps[(size_t)site * M.B + bA] : 0, dsB = vB ? M.exps[(size_t)site * M.B + bB] : 0;
      zero_acc();
      mma_step();
      contract(qa, qb, o);
      int eA, eB;
      bool nzA, nzB;
      renorm(o, eA, eB, nzA, nzB);
      store_image(img_slot(A.gimg, half == 0 ? c - 1 - i : c + 2 + i));
      // dA_site = sum_b 2^q (F x phi x G), q = h_in - delta_site, h_in = exponent of the input adjoint
      const int qA = hA - dsA, qB = hB - dsB;
      if (vH) {
        const float sc2 = ldexpf(1.f, sel(qA, qB));
#pragma unroll
        for (int s = 0; s < D; ++s) A.wq[((size_t)site * D + s) * M.B + bH] = (t == 0 ? qa[s] : qb[s]) * sc2;
      }
      int qm = max(vA ? qA : INT_MIN, vB ? qB : INT_MIN);
      for (int off = 16; off > 0; off >>= 1) qm = max(qm, __shfl_xor_sync(0xffffffffu, qm, off));
      if (lane == 0 && qm != INT_MIN) atomicMax(&A.qmax[site], qm);
      hA += (nzA ? eA - kTcKA - kw : 0) - dsA;
      hB += (nzB ? eB - kTcKA - kw : 0) - dsB;
    }
    // boundary site: no chain step, but it has a gradient
    {
      const int jb = half == 0 ? 0 : L - 1;
      float qa[D], qb[D];
      phis(jb, qa, qb);
      const int qA = vA ? hA - M.exps[(size_t)jb * M.B + bA] : INT_MIN, qB = vB ? hB - M.exps[(size_t)jb * M.B + bB] : INT_MIN;
      if (vH) {
        const float sc2 = ldexpf(1.f, sel(qA, qB));
#pragma unroll
        for (int s = 0; s < D; ++s) A.wq[((size_t)jb * D + s) * M.B + bH] = (t == 0 ? qa[s] : qb[s]) * sc2;
      }
      int qm = max(qA, qB);
      for (int off = 16; off > 0; off >>= 1) qm = max(qm, __shfl_xor_sync(0xffffffffu, qm, off));
      if (lane == 0 && qm != INT_MIN) atomicMax(&A.qmax[jb], qm);
    }
  }
}

}  // namespace tnb
