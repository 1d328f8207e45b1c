// Register-resident MPS kernels at bond 32 (float32 tensor family, `mma.sync` fp16x3; smaller bonds run zero-padded): host-side
// launch code.  Kernels: mps_rr_chain.cuh (chains, class site, head, adjoint seeds), mps_rr_da.cuh (weight gradients of the chain
// sites).  Called from the tcgen05 family's launchers (mps_tc.cu) when the plan selects them.
#include "host.h"
#include "mps_rr_chain.cuh"
#include "mps_rr_da.cuh"

namespace tnb {

// ring of site images + barriers + the step table (site, kW) of at most L + n_out + 1 steps
#ifndef TNB_RRC_PAD_SMEM
#define TNB_RRC_PAD_SMEM 0  // (occupancy experiments: extra bytes that keep a second CTA off the SM)
#endif
static size_t rrc_smem(int d, int steps) { return (size_t)kRrcStages * 4096 * d + 128 + (size_t)(steps + 2) * 8 + TNB_RRC_PAD_SMEM; }

static constexpr size_t kRdSmem = (size_t)kRdWarps * kRdStages * kRdStage + (size_t)kRdWarps * kRdStages * 8;

cudaError_t mps_rr_set_attrs() {
  cudaError_t e = cudaFuncSetAttribute(rr_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rr_chain_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rr_da_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRdSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rr_da_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRdSmem);
  return e;
}

bool mps_rr_supported(const tnb_problem *p) {
  return knobs().rr_chain && p->kind == TNB_KIND_MPS && p->dtype == TNB_F32 && p->chi <= 32 && (p->d == 2 || p->d == 3) &&
         rrc_smem(p->d, p->L + p->n_out) <= (size_t)kTcSmemMax;
}

// forward (adjoint = 0) or adjoint (1) sweep over B samples; T as filled by the tcgen05 family (same images, same outputs)
int mps_rr_launch_sweep(const TcArgs &T, int d, long long B, int adjoint, cudaStream_t st) {
  const size_t smem = rrc_smem(d, T.M.L + T.M.C);
  const long long ntiles = (B + kTcRows - 1) / kTcRows;
  dim3 grid((unsigned)((ntiles + kRrcTiles - 1) / kRrcTiles), adjoint ? 2 : 1);  // kRrcTiles sample tiles per CTA
  prof_begin(st);
  if (d == 2) rr_chain_kernel<2><<<grid, kRrcThreads, smem, st>>>(T, adjoint);
  else rr_chain_kernel<3><<<grid, kRrcThreads, smem, st>>>(T, adjoint);
  TNB_LAUNCH_CHECK(adjoint ? "rr_chain_kernel(adj)" : "rr_chain_kernel(fwd)");
  return 0;
}

// weight gradients of the chain sites [Da.site0, Da.site0 + nsites) at bond 32 (the class site is skipped: tc_da_kernel's pass 1);
// Da as filled by mps_tc_backward (operands, weights, output)
int mps_rr_launch_da(const TcDaArgs &Da_in, int d, long long B, int nsites, cudaStream_t st) {
  TcDaArgs Da = Da_in;
  // split the batch so that the grid is a whole number of rounds of 2 CTAs per SM (all CTAs take the same time: a partly
  // filled last round is lost bandwidth), at least two rounds, every warp with at least 8 chunks of 16 samples; one split
  // (a single writer per entry) when deterministic
  const long long maxsp = (B + 1023) / 1024;
  const int nwork = nsites - ((Da.M.c >= Da.site0 && Da.M.c < Da.site0 + nsites) ? 1 : 0);  // (the class site's CTAs leave at once)
  const long long slots = d == 2 ? 296 : 148;  // CTAs resident on the machine (two per SM for d = 2)
  long long nsplit = 1;
  double best = -1.0;
  for (long long sp = 1; sp <= maxsp && sp <= 16; ++sp) {
    const long long n = (long long)(nwork > 0 ? nwork : 1) * sp;
    const double eff = (double)n / (double)((n + slots - 1) / slots * slots);
    const double score = eff - (n < 2 * slots ? 0.5 : 0.0) - 0.004 * (double)sp;
    if (score > best) { best = score; nsplit = sp; }
  }
  if (deterministic()) nsplit = 1;
  Da.nsplit = (int)nsplit;
  Da.bchunk = ((B + nsplit - 1) / nsplit + 127) / 128 * 128;
  dim3 grid((unsigned)nsplit, (unsigned)nsites);
  prof_begin(st);
  if (d == 2) rr_da_kernel<2><<<grid, kRdWarps * 32, kRdSmem, st>>>(Da);
  else rr_da_kernel<3><<<grid, kRdWarps * 32, kRdSmem, st>>>(Da);
  TNB_LAUNCH_CHECK("rr_da_kernel");
  return 0;
}

}  // namespace tnb
