#!/usr/bin/env python
"""Forward-only (no environment stash) timing of the cfg-5 chain at a given chi: separates the stash's HBM traffic from
the rest of the chain kernel.  python tools/fwd_only.py <chi> [batch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tn4ml_b200 as T  # noqa: E402
from oracle import tn_oracle as O  # noqa: E402
from tn4ml_b200 import metrics as M  # noqa: E402

chi = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
L, C = 196, 10
dev = torch.device("cuda", 0)
arrays = O.make_mps_arrays(L, chi, 2, class_site=L // 2 - 1, C=C, std=1e-2, seed=42, dtype=torch.float32)
model = T.TensorNetwork(arrays, dtype=torch.float32, device=dev)
loss = lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()  # noqa: E731
eng, _ = model._engine(T.TrigonometricEmbedding(), loss, True, torch.float32)
rng = np.random.default_rng(1)
x = torch.from_numpy((1.0 - rng.random((B, L))).astype(np.float32)).to(dev)
t = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, C, size=B)), C).to(dev, torch.float32)
for _ in range(2):
    eng.forward(model._packed, x, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.forward(model._packed, x, t)
e1.record()
torch.cuda.synchronize()
print(f"chi={chi} B={B} forward-only {e0.elapsed_time(e1) / 3:.3f} ms  (tiles env {os.environ.get('TNB_TC_TILES', 'default')})")
