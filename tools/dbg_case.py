import sys, torch, ctypes as C
sys.path.insert(0, "/root/repo")
from oracle import tn_oracle as O
import tn4ml_b200 as T
from tn4ml_b200 import _lib, metrics as M
case = dict(L=8, chi=int(sys.argv[1]), d=2, cls=3, C=3, B=70)
arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], class_site=case["cls"], C=case["C"], std=0.05, seed=3)
x = O.make_inputs(case["B"], case["L"], seed=5).float().cuda()
t = O.make_labels(case["B"], case["C"]).float().cuda()
model = T.TensorNetwork(arrays, dtype=torch.float32, device="cuda:0")
model.precision = _lib.PREC_TENSOR
loss_fn = lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()
eng, _ = model._engine(T.TrigonometricEmbedding(), loss_fn, True, torch.float32)
lib = _lib.lib()
B = case["B"]
ws = eng._workspace(B, True)
ls = torch.zeros(1, dtype=torch.float64, device="cuda")
g = torch.zeros(eng.n_params, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("fwd", lib.tnb_forward(C.byref(eng.prob), B, x.data_ptr(), t.data_ptr(), model._packed.data_ptr(), None, ls.data_ptr(), None, None, 1, ws.data_ptr(), ws.numel(), st), flush=True)
torch.cuda.synchronize(); print("fwd done", float(ls)/B, flush=True)
print("bwd", lib.tnb_backward(C.byref(eng.prob), B, x.data_ptr(), t.data_ptr(), model._packed.data_ptr(), 1.0/B, g.data_ptr(), 0, ws.data_ptr(), ws.numel(), st), flush=True)
torch.cuda.synchronize(); print("bwd done", float(g.abs().max()), flush=True)
