import time, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
import tn4ml_b200 as T
from tn4ml_b200 import synthetic as S, metrics as M
L, chi, C, B = 196, 10, 10, 64
arrays = S.make_mps_arrays(L, chi, 2, class_site=97, C=C, std=1e-2, seed=42)
model = T.TensorNetwork(arrays, dtype=torch.float64, device="cuda:0")
loss = lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()
model.configure(strategy="sweeps", optimizer=T.optim.adam, loss=loss, train_type=T.TrainingType.SUPERVISED, learning_rate=1e-3, device=("gpu", 0))
x = S.make_inputs(B, L, seed=1).numpy(); t = S.make_labels(B, C).numpy()
t0 = time.time(); h = model.train(x, targets=t, batch_size=B, epochs=1, dtype=np.float64); t1 = time.time()
h = model.train(x, targets=t, batch_size=B, epochs=2, dtype=np.float64); t2 = time.time()
print("first epoch (incl. init loop)", round(t1 - t0, 2), "s; next two epochs", round(t2 - t1, 2), "s; losses", [round(v, 5) for v in h["loss"]], "pairs per sweep", 2 * (L - 1))
