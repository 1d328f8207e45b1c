#!/usr/bin/env python
"""Time one fwd+bwd step of every BASELINE.json config (device-resident inputs, CUDA events).
Not the driver's bench contract (that is bench.py); this fills the per-config table in DESIGN.md."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tn4ml_b200 as T  # noqa: E402
from tn4ml_b200 import synthetic as O  # noqa: E402  (synthetic-input generators)
from tn4ml_b200 import _lib, metrics as M  # noqa: E402


def mps_flops(L, chi, d, C, c):
    f = 0
    for i in range(L):
        cl, cr = (1 if i == 0 else chi), (1 if i == L - 1 else chi)
        f += 2 * cl * cr * d * (C if i == c else 1)
    return 3 * f


def smpo_flops(L, S, chi, d, do):
    w = (S - 2) * 2 * chi**3 + 2 * chi**3 * do + 4 * chi**3 * do + S * 2 * chi**2 * d + 2 * chi**2 * d * (do - 1)
    return 3 * (L // S) * w


CONFIGS = {
    "cfg1": dict(kind="cls", L=196, chi=10, d=2, C=10, B=64, emb=T.TrigonometricEmbedding(), head="xent"),
    "cfg2": dict(kind="cls", L=30, chi=32, d=3, C=2, B=65536, emb=T.PolynomialEmbedding(degree=2, n=1, include_bias=True),
                 head="wxent"),
    "cfg3": dict(kind="smpo", L=784, S=8, chi=32, d=2, do=2, B=4096, emb=T.TrigonometricEmbedding(), head="logquad"),
    "cfg4": dict(kind="nll", L=16, chi=64, d=8, B=65536, emb=T.FourierEmbedding(p=8), head="nll"),
    "cfg5_chi32": dict(kind="cls", L=196, chi=32, d=2, C=10, B=131072, emb=T.TrigonometricEmbedding(), head="xent"),
    "cfg5_chi128": dict(kind="cls", L=196, chi=128, d=2, C=10, B=131072, emb=T.TrigonometricEmbedding(), head="xent"),
    "cfg5": dict(kind="cls", L=196, chi=256, d=2, C=10, B=131072, emb=T.TrigonometricEmbedding(), head="xent"),
}


def run(name, cfg, dtype, steps=3, warmup=2):
    dev = torch.device("cuda", 0)
    L, chi, d, B = cfg["L"], cfg["chi"], cfg["d"], cfg["B"]
    rng = np.random.default_rng(1)
    x = torch.from_numpy(1.0 - rng.random((B, L))).to(dev, dtype)
    t = None
    if cfg["kind"] == "cls":
        c = L // 2 - 1
        arrays = O.make_mps_arrays(L, chi, d, class_site=c, C=cfg["C"], std=1e-2, seed=42, dtype=dtype)
        model = T.TensorNetwork(arrays, dtype=dtype, device=dev)
        t = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, cfg["C"], size=B)), cfg["C"]).to(dev, dtype)
        loss = (M.CrossEntropyWeighted([0.7, 1.6]) if cfg["head"] == "wxent" else
                (lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()))
        flops = mps_flops(L, chi, d, cfg["C"], c)
    elif cfg["kind"] == "nll":
        arrays = O.make_mps_arrays(L, chi, d, std=1e-2, seed=42, dtype=dtype, squeeze_ends=True, identity_all=True)
        model = T.MatrixProductState(arrays, dtype=dtype, device=dev)
        loss = M.NegLogLikelihood
        flops = mps_flops(L, chi, d, 1, L - 1)
    else:
        arrays = O.make_smpo_arrays(L, chi, d, cfg["do"], cfg["S"], std=1e-2, seed=42, dtype=dtype)
        model = T.SpacedMatrixProductOperator(arrays, spacing=cfg["S"], dtype=dtype, device=dev)
        loss = M.LogQuadNorm
        flops = smpo_flops(L, cfg["S"], chi, d, cfg["do"])
    eng, _ = model._engine(cfg["emb"], loss, t is not None, dtype)
    grads = torch.zeros(eng.n_params, dtype=dtype, device=dev)
    for _ in range(warmup):
        ls, _ = eng.value_and_grad(model._packed, x, t, grads=grads)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ls, _ = eng.value_and_grad(model._packed, x, t, grads=grads)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib = _lib.lib()
    out = {"config": name, "dtype": str(dtype).split(".")[-1], "batch": B, "ms_per_step": ms,
           "samples_per_s": B / ms * 1e3, "tflops": flops * B / ms / 1e9, "loss": float(ls.item()) / B,
           "microbatch": eng.microbatch(B, True), "finite_grads": bool(torch.isfinite(grads).all())}
    print(json.dumps(out), flush=True)
    del eng, model, grads, x
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    res = []
    for n in names:
        for dt in (torch.float32, torch.float64):
            if n in ("cfg5", "cfg5_chi128") and dt == torch.float64 and os.environ.get("SKIP_BIG_F64"):
                continue
            try:
                res.append(run(n, CONFIGS[n], dt))
            except Exception as exc:  # noqa: BLE001
                print(json.dumps({"config": n, "dtype": str(dt), "error": str(exc)[:300]}), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/configs.json", "w"), indent=1)
