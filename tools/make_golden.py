#!/usr/bin/env python
"""Generate tests/golden/*.npz: small fixed problems with their losses and per-site gradients.

The reference (jax + quimb) cannot be imported or installed in this environment, so these vectors are
NOT outputs of the reference.  They come from the oracle's brute-force route: the full d^L amplitude tensor (MPS) /
the explicit dense vector over all output legs (SMPO) is built and contracted with the product state, and gradients
are taken by autograd through that dense expression -- no chain recurrence, no environment reuse.  They pin
(a) the chain-based oracle and (b) the CUDA kernels against an independent float64 evaluation of the same
mathematical object the reference defines (tn4ml/metrics.py:17-53, 56-81, 171-206, 423-487; smpo.py:269-404).

    python tools/make_golden.py        # rewrites tests/golden/*.npz (deterministic)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tn_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (kind, kwargs)
    "mps_classifier_trig": dict(kind="mps", L=7, chi=5, d=2, cls=3, C=3, B=12, emb=("trig", dict(k=1)), head="softmax_xent"),
    "mps_classifier_poly_weighted": dict(kind="mps", L=6, chi=4, d=3, cls=2, C=2, B=10,
                                         emb=("poly", dict(degree=2, include_bias=True)), head="softmax_xent_weighted",
                                         cw=(0.7, 1.6)),
    "mps_nll_fourier": dict(kind="mps", L=6, chi=4, d=4, B=9, emb=("fourier", dict(p=4)), head="nll"),
    "mps_mse_rbf": dict(kind="mps", L=6, chi=4, d=3, cls=5, C=4, B=8,
                        emb=("rbf", dict(gamma=0.8, centers=(0.0, 0.5, 1.0))), head="mse"),
    "smpo_logquad_trig": dict(kind="smpo", L=8, chi=3, d=2, d_o=2, S=3, B=6, emb=("trig", dict(k=1)), head="logquad"),
    "mpo_tsn_trig": dict(kind="smpo", L=5, chi=3, d=2, d_o=2, S=1, B=5, emb=("trig", dict(k=1)), head="tsn"),
}


def build(case):
    emb = O.EmbSpec(case["emb"][0], **case["emb"][1])
    x = O.make_inputs(case["B"], case["L"], seed=11)
    t = None
    if case["kind"] == "mps":
        if "cls" in case:
            arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], class_site=case["cls"], C=case["C"], std=0.3, seed=5)
            t = O.make_labels(case["B"], case["C"], seed=3)
        else:
            arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], std=0.3, seed=5, squeeze_ends=True,
                                       identity_all=True)
        prob = O.Problem("mps", emb, case["head"], class_weights=case.get("cw"))
    else:
        arrays = O.make_smpo_arrays(case["L"], case["chi"], case["d"], case["d_o"], case["S"], std=0.3, seed=5)
        prob = O.Problem("smpo", emb, case["head"], spacing=case["S"])
    leaves = [a.detach().clone().requires_grad_(True) for a in arrays]
    phi = O.embed_batch(x, emb)
    out = O.mps_dense_outputs(phi, leaves) if case["kind"] == "mps" else O.smpo_dense_sqnorm(phi, leaves, case["S"])
    cw = None if case.get("cw") is None else torch.tensor(case["cw"], dtype=x.dtype)
    per = O.head_loss(out, case["head"], t, cw)
    loss = per.mean()
    grads = torch.autograd.grad(loss, leaves)
    return prob, x, t, arrays, per.detach(), loss.detach(), [g.detach() for g in grads]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, case in CASES.items():
        prob, x, t, arrays, per, loss, grads = build(case)
        blob = {"x": x.numpy(), "per_sample_loss": per.numpy(), "loss": np.array(float(loss))}
        if t is not None:
            blob["targets"] = t.numpy()
        for i, (a, g) in enumerate(zip(arrays, grads)):
            blob[f"array_{i}"] = a.numpy()
            blob[f"grad_{i}"] = g.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
        print(name, "loss", float(loss), "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


if __name__ == "__main__":
    main()
