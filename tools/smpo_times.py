#!/usr/bin/env python
"""cfg 3 (SMPO L=784, S=8, chi=32, d=d_o=2, B=4096): forward-only and fwd+bwd times, float32 and float64 (CUDA events).
TNB_RR=0 in the environment selects the shared-memory family for comparison."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tn4ml_b200 as T  # noqa: E402
from tn4ml_b200 import metrics as M, synthetic as S  # noqa: E402


def run(dtype, B=4096, L=784, Sp=8, chi=32, steps=5):
    arrays = S.make_smpo_arrays(L, chi, 2, 2, Sp, std=1e-2, seed=42, dtype=dtype)
    model = T.SpacedMatrixProductOperator(arrays, spacing=Sp, dtype=dtype, device="cuda:0")
    eng, _ = model._engine(T.TrigonometricEmbedding(), M.LogQuadNorm, False, dtype)
    rng = np.random.default_rng(1)
    x = torch.from_numpy(1.0 - rng.random((B, L))).to("cuda:0", dtype)
    g = torch.zeros(eng.n_params, dtype=dtype, device="cuda:0")
    out = {}
    for name, fn in (("fwd", lambda: eng.forward(model._packed, x)), ("fwd+bwd", lambda: eng.value_and_grad(model._packed, x, grads=g))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / steps
    w = (Sp - 2) * 2 * chi**3 + 2 * chi**3 * 2 + 4 * chi**3 * 2 + Sp * 2 * chi**2 * 2 + 2 * chi**2 * 2
    fl = (L // Sp) * w * B
    print(f"{dtype} RR={os.environ.get('TNB_RR', '1')}: fwd {out['fwd']:.2f} ms ({fl / out['fwd'] / 1e9:.1f} TFLOP/s), "
          f"fwd+bwd {out['fwd+bwd']:.2f} ms ({3 * fl / out['fwd+bwd'] / 1e9:.1f} TFLOP/s, {B / out['fwd+bwd'] * 1e3:.0f} samples/s)", flush=True)


if __name__ == "__main__":
    for dt in (torch.float32, torch.float64):
        run(dt)
