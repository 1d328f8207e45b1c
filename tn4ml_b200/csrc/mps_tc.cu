// tcgen05 family of the MPS path (float32, TNB_PREC_TENSOR): host-side launch code.
// Kernels: mps_tc.cuh (weight images, dA GEMM) and mps_tc_chain.cuh (chain kernel).  Entry points in host.h.
#include "host.h"
#include "mps_args.h"
#include "mps_tc.cuh"

namespace tnb {

int mps_rr_launch_sweep(const TcArgs &T, int d, long long B, int adjoint, cudaStream_t st);  // mps_rr.cu
int mps_rr_launch_da(const TcDaArgs &Da, int d, long long B, int nsites, cudaStream_t st);    // mps_rr.cu

static void fill_tc_args(const MpsPlan &P, void *ws, const MpsArgs<float> &M, TcArgs &T) {
  char *w = (char *)ws;
  T.M = M;
  T.img_n = (const uint8_t *)(w + P.off_img_n);
  T.img_t = (const uint8_t *)(w + P.off_img_t);
  T.kW = (const int *)(w + P.off_kW);
  T.tile0 = 0;
  T.cg = P.tc_cg;
#ifdef TNB_TC_TRACE
  T.dbg = knobs().tc_dbg;
#endif
  T.colsum = (const float *)(w + P.off_cs);
  // stash writer: pause between pieces.  A __nanosleep costs >= ~250 cycles whatever its argument, and the writer copies an operand
  // in 64 pieces: at chi = 256 (128 KB per operand, 16 k cycles per step) the pauses are hidden and keep the bulk copies from
  // crowding the weight loads; for every smaller operand they ARE the step (in-kernel timeline, cfg 4: operand published 7.4 k
  // cycles after the gather finished, waiting for the writer: chain kernels 0.65 / 0.74 -> 0.39 / 0.52 ms without pauses;
  // chi = 128: 6.75 / 7.10 -> 6.2 / 6.7 ms).
  T.pace_ns = knobs().tc_pace_ns >= 0 ? knobs().tc_pace_ns : (M.chip >= 256 ? P.tc_N * 2 / 5 : 0);
  T.wq = (float *)(w + P.off_wq);
  T.qmax = (int *)(w + P.off_qmax);
  T.wqc = (float *)(w + P.off_wqc);
  T.fimg = (uint8_t *)(w + P.off_fimg);
  T.gimg = M.stash ? (uint8_t *)(w + P.off_gimg) : nullptr;
  T.img_bond_bytes = P.img_bond_bytes;
  T.img_stash = M.stash;
  T.qmaxc = T.qmax + M.L;
  T.N = P.tc_N; T.NH = P.tc_NH; T.NMMA = P.tc_NMMA; T.nstage = P.tc_nstage; T.tmem_cols = P.tc_cols;
  T.stage_bytes = P.tc_stage_bytes; T.site_img_bytes = P.tc_site_img;
}

template <int D, int SLOTS, int TILES, int CG, int NGRP = 4> static cudaError_t tc_set_attr() {
  return cudaFuncSetAttribute(tc_chain_kernel<D, SLOTS, TILES, CG, NGRP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax);
}

// every instantiation tc_launch_sweep can reach + the dA GEMM, for the current device
cudaError_t mps_tc_set_attrs() {
  cudaError_t e = tc_set_attr<2, 4, 1, 2>();
  if (e == cudaSuccess) e = tc_set_attr<2, 4, 1, 1>();
  if (e == cudaSuccess) e = tc_set_attr<2, 2, 1, 1>();
  if (e == cudaSuccess) e = tc_set_attr<2, 2, 2, 1>();
  if (e == cudaSuccess) e = tc_set_attr<3, 4, 1, 1>();
  if (e == cudaSuccess) e = tc_set_attr<3, 2, 1, 1>();
  if (e == cudaSuccess) e = tc_set_attr<4, 2, 1, 1>();
  if (e == cudaSuccess) e = tc_set_attr<4, 2, 2, 1>();
  if (e == cudaSuccess) e = tc_set_attr<6, 2, 1, 1>();
  if (e == cudaSuccess) e = tc_set_attr<8, 2, 1, 1>();
  if (e == cudaSuccess) e = tc_set_attr<8, 2, 2, 1>();
  if (e == cudaSuccess) e = tc_set_attr<2, 2, 1, 1, 2>();
  if (e == cudaSuccess) e = tc_set_attr<3, 2, 1, 1, 2>();
  if (e == cudaSuccess) e = tc_set_attr<4, 2, 1, 1, 2>();
  if (e == cudaSuccess) e = tc_set_attr<6, 2, 1, 1, 2>();
  if (e == cudaSuccess) e = tc_set_attr<8, 2, 1, 1, 2>();
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_da_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax - 2048);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_da_kernel<kDaMaxCpg>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax - 2048);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_da_kernel<kDaMaxCpg, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax - 2048);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_da_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax - 2048);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_da_kernel<kDaMaxCpg, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax - 2048);
  return e;
}

template <int D, int SLOTS, int TILES, int CG = 1, int NGRP = 4>
static cudaError_t tc_launch_one(long long ntiles, long long tile0, size_t smem, TcArgs T, int adjoint, cudaStream_t st) {
  if (ntiles <= 0) return cudaSuccess;
  T.tile0 = tile0 * kTcRows;
  if (CG == 1) {
    dim3 grid((unsigned)(ntiles / TILES), adjoint ? 2 : 1);
    tc_chain_kernel<D, SLOTS, TILES, CG, NGRP><<<grid, (4 * NGRP + 3) * 32, smem, st>>>(T, adjoint);
    return cudaSuccess;
  }
  // CTA pair: clusters of two CTAs along x (consecutive sample tiles); an odd tile count gets an all-padding tile
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((ntiles + 1) / 2 * 2), adjoint ? 2 : 1);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, tc_chain_kernel<D, SLOTS, TILES, CG>, T, adjoint);
}

// launch the tcgen05 sweep (forward environments or adjoint chains)
static int tc_launch_sweep(const tnb_problem *p, const MpsPlan &P, const TcArgs &T, long long B, int adjoint,
                           cudaStream_t st) {
  if (P.tc_rr) return mps_rr_launch_sweep(T, p->d, B, adjoint, st);  // bond 32: register-resident chains on mma.sync
  cudaError_t e = cudaSuccess;
  const bool wide = P.tc_chi > 128;   // 16-column groups per epilogue thread: 4 (chi <= 256) or 2 (chi <= 128)
  const long long ntiles = (B + kTcRows - 1) / kTcRows;
  // Two-tile CTAs take whole rounds of 148 SMs (per half of the adjoint grid); the remaining tiles run as one-tile
  // CTAs behind them, so that the last round is not half empty with double-length CTAs.
  long long npair = 0;
  if (P.tc_tiles == 2) {
    const long long per_round = 2LL * 148 / (adjoint ? 2 : 1);
    npair = ntiles / per_round * per_round;
    if (ntiles - npair > per_round / 2) npair = ntiles / 2 * 2;  // the rest would need two more rounds anyway
  }
#define TNB_TC_GO(DD, SS)                                                                                     \
  do {                                                                                                        \
    if (npair > 0) e = tc_launch_one<DD, SS, 2>(npair, 0, P.tc_smem, T, adjoint, st);                         \
    if (e == cudaSuccess && ntiles > npair) e = tc_launch_one<DD, SS, 1>(ntiles - npair, npair, P.tc_smem, T, adjoint, st); \
  } while (0)
#define TNB_TC_GO1(DD, SS) e = tc_launch_one<DD, SS, 1>(ntiles, 0, P.tc_smem, T, adjoint, st)
#define TNB_TC_PAIR(DD, SS) e = tc_launch_one<DD, SS, 1, 2>(ntiles, 0, P.tc_smem, T, adjoint, st)
#define TNB_TC_SMALL(DD) e = tc_launch_one<DD, 2, 1, 1, 2>(ntiles, 0, P.tc_smem, T, adjoint, st)
  prof_begin(st);
  if (P.tc_ngrp == 2) {
    switch (p->d) {
      case 2: TNB_TC_SMALL(2); break;
      case 3: TNB_TC_SMALL(3); break;
      case 4: TNB_TC_SMALL(4); break;
      case 6: TNB_TC_SMALL(6); break;
      case 8: TNB_TC_SMALL(8); break;
      default: return fail("tensor path: unsupported physical dimension %d", p->d);
    }
  } else
  switch (p->d) {
    case 2:
      if (wide && P.tc_cg == 2) TNB_TC_PAIR(2, 4);
      else if (wide) TNB_TC_GO1(2, 4);
      else TNB_TC_GO(2, 2);
      break;
    case 3: if (wide) TNB_TC_GO1(3, 4); else TNB_TC_GO1(3, 2); break;
    case 4: TNB_TC_GO(4, 2); break;
    case 6: TNB_TC_GO1(6, 2); break;
    case 8: TNB_TC_GO(8, 2); break;
    default: return fail("tensor path: unsupported physical dimension %d", p->d);
  }
#undef TNB_TC_SMALL
#undef TNB_TC_GO
#undef TNB_TC_GO1
#undef TNB_TC_PAIR
  if (e != cudaSuccess) return fail("tc_chain launch: %s", cudaGetErrorString(e));
  TNB_LAUNCH_CHECK(adjoint ? "tc_chain_kernel(adj)" : "tc_chain_kernel(fwd)");
  return 0;
}

int mps_tc_forward(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                   const void *params, void *loss, double *loss_sum, void *out, int32_t *out_exp, int keep_stash,
                   void *ws, cudaStream_t st) {
  MpsArgs<float> A;
  fill_mps_args<float>(p, P, B, x, targets, ws, A);
  A.chip = P.tc_chi;  // the tensor path's bond (32 for a zero-padded smaller one); A.chi stays the caller's
  A.stash = keep_stash ? 1 : 0;
  A.loss = (float *)loss; A.loss_sum = loss_sum; A.out = (float *)out; A.out_exp = out_exp;
  // weight images, then ONE kernel for both chains, the class site and the head
  TcArgs Tc;
  fill_tc_args(P, ws, A, Tc);
  TNB_CUDA(cudaMemsetAsync((void *)Tc.kW, 0, (size_t)P.nslab * sizeof(int), st));
  TNB_CUDA(cudaMemsetAsync((void *)Tc.colsum, 0, (size_t)2 * P.nslab * sizeof(float), st));
  prof_begin(st);
  tc_stats_kernel<<<dim3(P.nslab, kTcStatGroups), 256, 0, st>>>((const float *)params, (int *)Tc.kW, (int *)Tc.colsum, p->L,
                                                                 p->chi, p->d, p->n_out, p->out_site);
  TNB_LAUNCH_CHECK("tc_stats_kernel");
  prof_begin(st);
  tc_finalize_kernel<<<1, 256, 0, st>>>((int *)Tc.kW, (float *)Tc.colsum, p->L, p->n_out, p->out_site);
  TNB_LAUNCH_CHECK("tc_finalize_kernel");
  size_t items = (size_t)P.tc_N * (P.tc_chi / 8) * P.nslab;
  int ib = (int)((items + 255) / 256);
  if (ib > 148 * 32) ib = 148 * 32;
  prof_begin(st);
  tc_image_kernel<<<ib, 256, 0, st>>>((const float *)params, Tc.kW, (uint8_t *)Tc.img_n, (uint8_t *)Tc.img_t, p->L,
                                      p->chi, P.tc_chi, p->d, p->n_out, p->out_site, P.tc_NH, P.tc_cg);
  TNB_LAUNCH_CHECK("tc_image_kernel");
  return tc_launch_sweep(p, P, Tc, B, 0, st);
}

int mps_tc_backward(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                    double loss_scale, void *grads, int accumulate, void *ws, cudaStream_t st, int part, int nparts) {
  MpsArgs<float> A;
  fill_mps_args<float>(p, P, B, x, targets, ws, A);
  A.chip = P.tc_chi;
  A.stash = 1;
  A.loss_scale = (float)loss_scale;
  TcArgs Tc;
  fill_tc_args(P, ws, A, Tc);
  if (part == 0) {
    if (!accumulate) TNB_CUDA(cudaMemsetAsync(grads, 0, (size_t)tnb_param_count(p) * sizeof(float), st));
    TNB_CUDA(cudaMemsetAsync(Tc.qmax, 0x80, ((size_t)p->L + 1) * sizeof(int), st));
    if (tc_launch_sweep(p, P, Tc, B, 1, st)) return 1;  // seeds through the class site + adjoint chains
  }
  int s0, s1;
  part_sites(p->L, part, nparts, &s0, &s1);
  if (s1 <= s0) return 0;
  TcDaArgs Da;
  Da.M = A;
  Da.grads = (float *)grads;
  Da.fimg = Tc.fimg;
  Da.gimg = Tc.gimg;
  Da.img_bond_bytes = Tc.img_bond_bytes;
  // stages sized by the bond: the U region is fixed (a virtual image of 256 rows), the V chunk is chi * 4 * kDaKC bytes.
  // Bonds up to 64: two CTAs of 8 producer warps per SM (see tc_da_kernel), each with half of the shared memory.
  const int tchi = P.tc_chi;
  Da.stage_bytes = (uint32_t)align_up((size_t)kDaU + (size_t)tchi * 4 * kDaKC, 1024);
  // da_small: largest bond that takes the several-CTAs-per-SM variants (knob TNB_DA_SMALL; 0 = never)
  const bool small = tchi <= knobs().da_small;
  const int ncta = !small ? 1 : ((tchi <= 32 && knobs().da_pw4) ? 3 : 2);   // CTAs per SM the stages are sized for
  const size_t smem_avail = (size_t)(kTcSmemMax - 2048) / ncta - 1024 - (ncta > 1 ? 1024 : 0);
  Da.nstage = (int)(smem_avail / Da.stage_bytes);
  if (Da.nstage > 6) Da.nstage = 6;
  if (knobs().da_stages >= 2 && knobs().da_stages < Da.nstage) Da.nstage = knobs().da_stages;
  Da.tmem_cols = 32;
  while (Da.tmem_cols < 2 * tchi) Da.tmem_cols <<= 1;
  const long long maxsp = (B + 2047) / 2048;
  Da.ra = tchi < 128 ? tchi : 128;
  Da.cpg = 256 / Da.ra;
  Da.site0 = s0;
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: every chain site of the range (copies sigma = s);  pass 1: the class site (copies sigma = (s, k))
    if (pass == 1 && (p->out_site < s0 || p->out_site >= s1)) break;
    if (pass == 0 && s1 - s0 == 1 && s0 == p->out_site) continue;
    Da.Dv = pass == 0 ? p->d : p->d * p->n_out;
    Da.only_site = pass == 0 ? -1 : p->out_site;
    Da.wq = pass == 0 ? Tc.wq : Tc.wqc;
    Da.qmax = pass == 0 ? Tc.qmax : Tc.qmaxc;
    if (pass == 0 && P.tc_rr && knobs().rr_da) {  // bond 32: the chain sites' gradients register resident on mma.sync
      if (mps_rr_launch_da(Da, p->d, B, s1 - s0, st)) return 1;
      continue;
    }
    Da.mgroups = ((tchi + Da.ra - 1) / Da.ra) * ((Da.Dv + Da.cpg - 1) / Da.cpg);
    // split K (the batch) so that the grid fills the 148 SMs about four times over with little wave quantisation
    const long long ctas = (long long)(pass == 0 ? s1 - s0 : 1) * Da.mgroups;
    // ... and so that no TMEM accumulator takes more than ~3 k MMAs: tcgen05 accumulates with truncation, and a coherent
    // sum of n terms loses about n / 4 ulp (1.2e-4 of the largest gradient entries at 43 k samples per split, measured
    // against the SIMT family); 16 k samples per split keep that below 5e-5.
    const long long nsp_min = (B + 16383) / 16384;
    long long best = nsp_min;
    double best_eff = 0.0;
    for (long long nsp = nsp_min; nsp <= maxsp && nsp <= 64; ++nsp) {
      const long long n = ctas * nsp;
      const double eff = (double)n / (double)((n + 147) / 148 * 148);
      const double score = eff - (n < 148 * 3 ? 1.0 : 0.0) - 0.002 * (double)nsp;  // prefer >= 3 waves, fewer splits
      if (nsp == nsp_min || score > best_eff) { best = nsp; best_eff = score; }
    }
    if (deterministic()) best = 1;  // one CTA per output tile: every gradient entry has a single writer, in a fixed order
    Da.nsplit = (int)best;
    Da.bchunk = ((B + Da.nsplit - 1) / Da.nsplit + kDaKC - 1) / kDaKC * kDaKC;
    dim3 gda((unsigned)(Da.mgroups * Da.nsplit), (unsigned)(pass == 0 ? s1 - s0 : 1));
    prof_begin(st);
    const size_t da_smem = (size_t)Da.nstage * Da.stage_bytes + 1024;
    if (Da.cpg == 2 && small) tc_da_kernel<2, 8><<<gda, 8 * 32 + 64, da_smem, st>>>(Da);
    else if (Da.cpg == 2) tc_da_kernel<2><<<gda, kDaThreads, da_smem, st>>>(Da);
    else if (ncta == 3) tc_da_kernel<kDaMaxCpg, 4><<<gda, 4 * 32 + 64, da_smem, st>>>(Da);
    else if (small) tc_da_kernel<kDaMaxCpg, 8><<<gda, 8 * 32 + 64, da_smem, st>>>(Da);
    else tc_da_kernel<kDaMaxCpg><<<gda, kDaThreads, da_smem, st>>>(Da);
    TNB_LAUNCH_CHECK(pass == 0 ? "tc_da_kernel" : "tc_da_kernel(class)");
  }
  return 0;
}

#ifdef TNB_TC_TRACE
/* measurement build only: copy the in-kernel timeline of the chain kernel (mps_tc_chain.cuh) to the host */
extern "C" int tnb_debug_trace_read(long long *host, int32_t n) {
  if (n > 64 * 32) n = 64 * 32;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host, tnb::g_tc_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? 0 : 1;
}
#endif

}  // namespace tnb
