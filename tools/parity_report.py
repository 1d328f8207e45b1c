#!/usr/bin/env python
"""Error-vs-(L, chi) table of the float32 kernel families against the float64 oracle (run on the GPU box).

For every shape: loss relative error, worst per-site relative Frobenius error of the gradient (sites below 1e-3 of the
largest gradient norm are measured against that floor, as in tests/test_gpu_parity.py), worst max-abs error over max|g|,
and the error profile along the chain (class site, quarter points, boundary).  Writes gpurun_out/<tag>_parity_table.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tn_oracle as O  # noqa: E402  (the checker)
import tn4ml_b200 as T  # noqa: E402
from tn4ml_b200 import metrics as M  # noqa: E402


def errors(L, chi, B, precision, dtype=torch.float32, std=1e-2, C=10, seed=42):
    arrays = O.make_mps_arrays(L, chi, 2, class_site=L // 2 - 1, C=C, std=std, seed=seed)
    x = O.make_inputs(B, L, seed=1239).to(dtype)
    t = O.make_labels(B, C).to(dtype)
    arr64 = [a.to(dtype).double() for a in arrays]
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    loss_ref, grads_ref = O.value_and_grad(prob, x.double(), arr64, t.double())
    model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
    model.precision = precision
    loss_fn = lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()  # noqa: E731
    eng, _ = model._engine(T.TrigonometricEmbedding(), loss_fn, True, dtype)
    ls, g = eng.value_and_grad(model._packed, x, t)
    loss = float(ls.item()) / B
    gmax = max(float(q.abs().max()) for q in grads_ref)
    gfro = max(float(q.norm()) for q in grads_ref)
    prof, wf, wm, scale = [], 0.0, 0.0, []
    for a, b in zip(model._views(g), grads_ref):
        d = a.double().cpu() - b
        ef = float(d.norm()) / max(float(b.norm()), 1e-3 * gfro)
        prof.append(ef)
        wf, wm = max(wf, ef), max(wm, float(d.abs().max()) / gmax)
        # least-squares scale factor of the computed gradient against the oracle's: 1 + a uniform shrink / growth
        scale.append(float((a.double().cpu() * b).sum() / (b * b).sum()) - 1.0)
    c = L // 2 - 1
    pick = sorted({0, c // 2, c, c + 1, (c + L) // 2, L - 1})
    return {"L": L, "chi": chi, "B": B, "family": "tcgen05 fp16x3" if precision == 1 else "SIMT fp32 FMA",
            "loss_rel_err": abs(loss - float(loss_ref)) / abs(float(loss_ref)), "site_fro_max": wf, "max_abs_over_max": wm,
            "profile": {str(i): prof[i] for i in pick}, "scale_minus_1": {str(i): scale[i] for i in pick}}


def big_batch(B, chi=256, L=196, C=10):
    """The accumulation over the BATCH (weight-gradient GEMM) cannot be checked by the oracle at full size; the float64 SIMT
    family on the same float32-rounded inputs can: error of both float32 families against it at batch B."""
    arrays = O.make_mps_arrays(L, chi, 2, class_site=L // 2 - 1, C=C, std=1e-2, seed=42)
    x = O.make_inputs(B, L, seed=1239).float()
    t = O.make_labels(B, C).float()
    loss_fn = lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()  # noqa: E731
    out = {}
    ref = None
    for name, dtype, precision in (("f64 SIMT (reference)", torch.float64, 0), ("tcgen05 fp16x3", torch.float32, 1),
                                   ("SIMT fp32 FMA", torch.float32, 0)):
        model = T.TensorNetwork([a.float() for a in arrays], dtype=dtype, device="cuda:0")
        model.precision = precision
        eng, _ = model._engine(T.TrigonometricEmbedding(), loss_fn, True, dtype)
        ls, g = eng.value_and_grad(model._packed, x.to(dtype), t.to(dtype))
        g = g.double().clone()
        if ref is None:
            ref = g
            continue
        views_r, views_g = model._views(ref), model._views(g)
        gmax = float(ref.abs().max())
        gfro = max(float(v.norm()) for v in views_r)
        wf = max(float((a - b).norm()) / max(float(b.norm()), 1e-3 * gfro) for a, b in zip(views_g, views_r))
        out[name] = {"site_fro_max": wf, "max_abs_over_max": float((g - ref).abs().max()) / gmax}
        del eng, model
        torch.cuda.empty_cache()
    return {"B": B, "chi": chi, "L": L, **out}


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
    rows = []
    for L, chi, B in [(196, 64, 256), (196, 128, 256), (196, 256, 256), (48, 256, 256), (96, 256, 256)]:
        for precision in (1, 0):
            if precision == 0 and chi == 256 and L == 196:
                B_ = 64
            else:
                B_ = B
            try:
                r = errors(L, chi, B_, precision)
            except Exception as exc:  # pragma: no cover
                r = {"L": L, "chi": chi, "error": str(exc)[:200]}
            rows.append(r)
            print(json.dumps(r), flush=True)
    for B in (256, 32768):
        try:
            r = big_batch(B)
        except Exception as exc:  # pragma: no cover
            r = {"B": B, "error": str(exc)[:300]}
        rows.append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"{tag}_parity_table.json"), "w"), indent=1)
