// MpsArgs<T> from a problem + plan + caller buffers (shared by the SIMT and the tcgen05 translation units).
#pragma once
#include "host.h"
#include "mps_simt.cuh"

namespace tnb {

template <typename T>
static void fill_mps_args(const tnb_problem *p, const MpsPlan &P, long long B, const void *x, const void *targets,
                          void *ws, MpsArgs<T> &A) {
  char *w = (char *)ws;
  A.L = p->L; A.chi = p->chi; A.chip = P.chip; A.D = p->d; A.C = p->n_out; A.c = p->out_site; A.head = p->head;
  A.B = B;
  fill_emb<T>(p->emb, A.emb, p->d);
  for (int i = 0; i < kMaxOut; ++i) A.cw[i] = (T)p->class_weights[i];
  A.x = (const T *)x; A.targets = (const T *)targets;
  A.Wn = (const T *)(w + P.off_Wn); A.Wt = (const T *)(w + P.off_Wt);
  A.F = (T *)(w + P.off_F); A.G = (T *)(w + P.off_G);
  A.exps = (int *)(w + P.off_exps); A.esum = (int *)(w + P.off_esum);
  A.y = (T *)(w + P.off_y); A.dy = (T *)(w + P.off_dy);
  A.wq = (T *)(w + P.off_wsimt);
  A.loss = nullptr; A.loss_sum = nullptr; A.out = nullptr; A.out_exp = nullptr;
  A.stash = 0; A.KT = P.KT; A.TN = P.TN; A.WS = P.WS; A.dmma_sweep = P.dmma_sweep; A.dmma_class = P.dmma_class; A.vst = P.vst; A.wsz = P.wsz; A.loss_scale = T(1); A.seed_only = 0;
}

}  // namespace tnb
