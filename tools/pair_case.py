#!/usr/bin/env python
"""Small chi = 256 forward+backward against the oracle (the CTA-pair chain kernel's first gate).  Prints progress so a
hang can be located.  python tools/pair_case.py [B]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tn_oracle as O  # noqa: E402
import tn4ml_b200 as T  # noqa: E402
from tn4ml_b200 import _lib, metrics as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
L, chi, C_ = 6, 256, 3
arrays = O.make_mps_arrays(L, chi, 2, class_site=2, C=C_, std=0.05, seed=3)
x, t = O.make_inputs(B, L, seed=5), O.make_labels(B, C_)
arr32 = [a.float() for a in arrays]
prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
model = T.TensorNetwork(arr32, dtype=torch.float32, device="cuda:0")
loss_fn = lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()  # noqa: E731
eng, _ = model._engine(T.TrigonometricEmbedding(), loss_fn, True, torch.float32)
lib = _lib.lib()
ws = eng._workspace(B, True)
ls = torch.zeros(1, dtype=torch.float64, device="cuda")
g = torch.zeros(eng.n_params, device="cuda")
st = torch.cuda.current_stream().cuda_stream
xd, td = x.float().cuda(), t.float().cuda()
print("fwd rc", lib.tnb_forward(C.byref(eng.prob), B, xd.data_ptr(), td.data_ptr(), model._packed.data_ptr(), None,
                                ls.data_ptr(), None, None, 1, ws.data_ptr(), ws.numel(), st), flush=True)
torch.cuda.synchronize()
print("fwd done", float(ls) / B, flush=True)
print("bwd rc", lib.tnb_backward(C.byref(eng.prob), B, xd.data_ptr(), td.data_ptr(), model._packed.data_ptr(), 1.0 / B,
                                 g.data_ptr(), 0, ws.data_ptr(), ws.numel(), st), flush=True)
torch.cuda.synchronize()
print("bwd done", flush=True)
lref, gref = O.value_and_grad(prob, x.float().double(), [a.double() for a in arr32], t)
gmax = max(float(q.abs().max()) for q in gref)
err = max(float((a.double().cpu() - b).abs().max()) for a, b in zip(model._views(g), gref)) / gmax
print(f"loss {float(ls) / B:.8f} oracle {float(lref):.8f}  max grad err / max|g| {err:.2e}", flush=True)
