#!/usr/bin/env python
"""Measurement aid: a small-bond float32 MPS classifier on its own kernels (SIMT family) against the same model zero-padded to
bond 32 (register-resident mma.sync family).   python tools/pad_experiment.py <chi> <batch> [L]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tn4ml_b200 import synthetic as O  # noqa: E402  (input generators)
import tn4ml_b200 as T  # noqa: E402
from tn4ml_b200 import metrics as M  # noqa: E402

chi, B = int(sys.argv[1]), int(sys.argv[2])
L = int(sys.argv[3]) if len(sys.argv) > 3 else 196
C = 10
arrays = O.make_mps_arrays(L, chi, 2, class_site=L // 2 - 1, C=C, std=1e-2, seed=42, dtype=torch.float32)
x = O.make_inputs(B, L, seed=1).float().cuda()
t = O.make_labels(B, C).float().cuda()
loss_fn = lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()  # noqa: E731


def pad(a, i):
    sh = list(a.shape)
    tgt = list(sh)
    if i > 0:
        tgt[0] = 32
    if i < L - 1:
        tgt[1] = 32
    out = torch.zeros(tgt, dtype=a.dtype)
    out[tuple(slice(0, s) for s in sh)] = a
    return out


def run(arrs, label):
    model = T.TensorNetwork(arrs, dtype=torch.float32, device="cuda:0")
    eng, _ = model._engine(T.TrigonometricEmbedding(), loss_fn, True, torch.float32)
    for _ in range(3):
        ls, g = eng.value_and_grad(model._packed, x, t)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ls, g = eng.value_and_grad(model._packed, x, t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"chi={chi} B={B} L={L} {label}: {ms:.3f} ms/step, {B / ms * 1e3:.3e} samples/s, loss {float(ls) / B:.6f}", flush=True)
    return g, model


g0, m0 = run(arrays, "own bond (SIMT family)")
g1, m1 = run([pad(a, i) for i, a in enumerate(arrays)], "padded to 32 (mma.sync family)")
v0, v1 = m0._views(g0), m1._views(g1)
err = max(float((b[tuple(slice(0, s) for s in a.shape)] - a).abs().max()) for a, b in zip(v0, v1)) / float(g0.abs().max())
print(f"max |g_padded - g| / max|g| = {err:.2e}")
