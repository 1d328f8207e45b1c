#!/usr/bin/env python
"""One fwd+bwd of cfg 3 at a reduced batch (for ncu captures of the SMPO kernels): smpo_one.py [B] [f32|f64]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tn4ml_b200 as T  # noqa: E402
from tn4ml_b200 import metrics as M, synthetic as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
dtype = torch.float64 if len(sys.argv) > 2 and sys.argv[2] == "f64" else torch.float32
L, Sp, chi = 784, 8, 32
arrays = S.make_smpo_arrays(L, chi, 2, 2, Sp, std=1e-2, seed=42, dtype=dtype)
model = T.SpacedMatrixProductOperator(arrays, spacing=Sp, dtype=dtype, device="cuda:0")
eng, _ = model._engine(T.TrigonometricEmbedding(), M.LogQuadNorm, False, dtype)
x = torch.from_numpy(1.0 - np.random.default_rng(1).random((B, L))).to("cuda:0", dtype)
for _ in range(2):
    ls, g = eng.value_and_grad(model._packed, x)
torch.cuda.synchronize()
print("loss", float(ls) / B)
