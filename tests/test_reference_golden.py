"""The oracle against golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE (tools/make_reference_golden.py:
/root/reference/tn4ml imported unmodified over stand-ins for its absent third-party stack, tools/ref_shim/README.md).

This is what pins the oracle: feature maps, embed()'s normaliser on both branches, the axis / index conventions of
plain and classifier MPS, every MPS loss head, and -- through central finite differences of the reference-executed mean loss --
individual entries of value_and_grad.  CPU only; the GPU parity tests then compare the kernels with the oracle.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tn_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SPECS = {
    "trig_k1": O.EmbSpec("trig", k=1), "trig_k2": O.EmbSpec("trig", k=2),
    "fourier_p4": O.EmbSpec("fourier", p=4), "fourier_p8": O.EmbSpec("fourier", p=8),
    "poly_d2_bias": O.EmbSpec("poly", degree=2, include_bias=True), "poly_d3": O.EmbSpec("poly", degree=3),
    "rbf": O.EmbSpec("rbf", gamma=0.8, centers=(0.1, 0.5, 0.9)),
}


@pytest.mark.parametrize("name", sorted(SPECS))
def test_feature_maps_match_the_reference_source(name):
    z = np.load(os.path.join(G, "reference_embeddings.npz"))
    phi = O.feature_map(torch.tensor(z["x"]), SPECS[name]).numpy()
    assert phi.shape[1] == int(z[f"dim_{name}"]) == SPECS[name].dim
    assert np.allclose(phi, z[f"phi_{name}"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("branch", ["short", "long"])
@pytest.mark.parametrize("name", ["poly_d2_bias", "trig_k1", "rbf"])
def test_embed_normaliser_matches_the_reference_source(branch, name):
    """embed() (embeddings.py:1374-1387): L <= 200 divides by norm**(1/L); L > 200 normalises site by site between QRs, which
    may move a sign from one site to the next -- the product state (what every contraction sees) is the same."""
    z = np.load(os.path.join(G, "reference_embeddings.npz"))
    x = torch.tensor(z[f"embed_{branch}_{name}_x"])
    ref = z[f"embed_{branch}_{name}"]
    assert abs(float(z[f"embed_{branch}_{name}_norm"]) - 1.0) < 1e-12
    mine = O.embed_batch(x[None, :], SPECS[name])[0].numpy()
    if branch == "short":
        assert np.allclose(mine, ref, rtol=1e-12, atol=1e-14)
        return
    k = np.abs(ref).argmax(axis=1)
    signs = np.sign(ref[np.arange(len(ref)), k]) * np.sign(mine[np.arange(len(ref)), k])
    assert np.allclose(mine * signs[:, None], ref, rtol=1e-10, atol=1e-13)
    assert np.prod(signs) == 1.0  # the signs the QRs move around cancel in the product state


MORE = {
    "lincomp_p2": O.EmbSpec("lincomp", p=2), "lincomp_p3": O.EmbSpec("lincomp", p=3),
    "legendre_d3": O.EmbSpec("legendre", degree=3), "legendre_d0": O.EmbSpec("legendre", degree=0),
    "laguerre_d3": O.EmbSpec("laguerre", degree=3), "hermite_d4": O.EmbSpec("hermite", degree=4),
    "hermite_d1": O.EmbSpec("hermite", degree=1),
    "poly_n3_d2": O.EmbSpec("poly", degree=2, n=3), "poly_n3_d2_bias": O.EmbSpec("poly", degree=2, n=3, include_bias=True),
    "poly_n3_d2_cross": O.EmbSpec("poly", degree=2, n=3, include_bias=True, cross=True),
    "poly_n2_d3_cross": O.EmbSpec("poly", degree=3, n=2, cross=True),
    "arrays_n3": O.EmbSpec("arrays", n=3), "arrays_n3_bias": O.EmbSpec("arrays", n=3, include_bias=True),
}


@pytest.mark.parametrize("name", sorted(MORE))
def test_remaining_feature_maps_match_the_reference_source(name):
    """LinearComplement / Legendre / Laguerre / Hermite (over the reference's tn4ml/scipy/special.py recurrences) / Polynomial
    with n > 1 and cross terms / JaxArrays, executed from the reference's files (tools/make_reference_golden.py)."""
    z = np.load(os.path.join(G, "reference_embeddings_more.npz"))
    spec = MORE[name]
    x = torch.tensor(z["xv"][:, :spec.n]) if spec.n > 1 else torch.tensor(z[f"x_{name}"])
    phi = O.feature_map(x, spec).numpy()
    assert phi.shape[1] == int(z[f"dim_{name}"]) == spec.dim
    assert np.allclose(phi, z[f"phi_{name}"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name,spec", [("hermite_d2", O.EmbSpec("hermite", degree=2)), ("arrays_n3", O.EmbSpec("arrays", n=3))])
def test_plain_mps_nll_over_remaining_maps_matches_the_reference_source(name, spec):
    z = np.load(os.path.join(G, "reference_embeddings_more.npz"))
    arrays = [torch.tensor(z[f"nll_{name}_site{i}"]) for i in range(6)]
    per = O.per_sample_loss(O.Problem("mps", spec, "nll"), torch.tensor(z[f"nll_{name}_x"]), arrays).numpy()
    assert np.allclose(per, z[f"nll_{name}"], rtol=1e-12)


def _sites(z):
    n = len([k for k in z.files if k.startswith("site")])
    return [torch.tensor(z[f"site{i}"]) for i in range(n)]


def _check_fd(prob, x, arrays, t, idx, fd):
    _, grads = O.value_and_grad(prob, x, arrays, t)
    for (s, k), g in zip(idx, fd):
        mine = float(grads[int(s)].reshape(-1)[int(k)])
        assert abs(mine - g) <= 1e-7 * max(1.0, abs(g)), (int(s), int(k), mine, g)


def test_plain_mps_nll_matches_the_reference_source():
    z = np.load(os.path.join(G, "reference_mps_nll.npz"))
    arrays, x = _sites(z), torch.tensor(z["x"])
    prob = O.Problem("mps", SPECS["poly_d2_bias"], "nll")
    per = O.per_sample_loss(prob, x, arrays).numpy()
    assert np.allclose(per, z["per_sample_loss"], rtol=1e-12)
    _check_fd(prob, x, arrays, None, z["fd_index"], z["fd_grad"])
    # len(model) < len(data): the reference sums the uncontracted data sites over their physical index (metrics.py:36-43)
    per_long = O.per_sample_loss(prob, torch.tensor(z["x_long"]), arrays).numpy()
    assert np.allclose(per_long, z["per_sample_loss_long"], rtol=1e-12)


@pytest.mark.parametrize("head", ["softmax_xent", "mse", "softmax_xent_weighted"])
def test_classifier_heads_match_the_reference_source(head):
    z = np.load(os.path.join(G, "reference_classifier.npz"))
    arrays, x, y = _sites(z), torch.tensor(z["x"]), torch.tensor(z["y"])
    cw = tuple(z["class_weights"]) if head == "softmax_xent_weighted" else None
    prob = O.Problem("mps", SPECS["trig_k1"], head, class_weights=cw)
    per = O.per_sample_loss(prob, x, arrays, y).numpy()
    assert np.allclose(per, z[f"loss_{head}"], rtol=1e-12)
    _check_fd(prob, x, arrays, y, z[f"fd_index_{head}"], z[f"fd_grad_{head}"])


def test_classifier_outputs_match_the_reference_source():
    """(model & data) ^ all with the class tensor built by the reference's MPS_initialize(arrays=..., class_index=...)."""
    z = np.load(os.path.join(G, "reference_classifier.npz"))
    arrays, x = _sites(z), torch.tensor(z["x"])
    out = O.mps_outputs(O.embed_batch(x, SPECS["trig_k1"]), arrays).numpy()
    # MPS_initialize normalises the model it builds (mps.py:470-490), so the raw outputs differ from those of the given arrays
    # by ONE positive factor (the heads above are scale free): same direction, same signs, constant ratio
    ratio = out / z["outputs"]
    assert ratio.min() > 0 and np.allclose(ratio, ratio.mean(), rtol=1e-10)
    assert int(z["class_site"]) == O.canon_mps_sites(arrays)[1]


@pytest.mark.parametrize("name", ["s2", "s4_lastout"])
def test_smpo_double_layer_matches_the_reference_source(name):
    """SpacedMatrixProductOperator built by the reference's constructor: <v|v> of the network TransformedSquaredNorm is defined
    on, and TransformedSquaredNorm / LogQuadNorm / QuadNorm executed from metrics.py through its len(model) < len(data) branch (the
    extra feature's site is summed over its physical index: a known factor on <v|v>); gradients by finite differences."""
    z = np.load(os.path.join(G, "reference_smpo.npz"))
    L, Sp = int(z[f"{name}_L"]), int(z[f"{name}_spacing"])
    arrays = [torch.tensor(z[f"{name}_site{i}"]) for i in range(L)]
    x = torch.tensor(z[f"{name}_x"])
    emb = O.EmbSpec("trig")
    n = O.smpo_sqnorm(O.embed_batch(x[:, :L], emb), arrays, Sp)
    assert np.allclose(n.numpy(), z[f"{name}_n_dense"], rtol=1e-12)
    assert np.allclose(O.per_sample_loss(O.Problem("smpo", emb, "tsn", spacing=Sp), x[:, :L], arrays).numpy(),
                       z[f"{name}_n_dense"] ** 2, rtol=1e-12)
    extra = O.feature_map(x[:, L], emb).sum(-1) ** 2          # the site beyond the model, summed over its physical index

    def heads(arrs):
        nn = O.smpo_sqnorm(O.embed_batch(x[:, :L], emb), arrs, Sp) * extra
        return {"tsn": O.head_loss(nn, "tsn"), "lqn": O.head_loss(nn, "logquad"), "qn": O.head_loss(nn, "quad")}

    h = heads(arrays)
    for k in ("tsn", "lqn", "qn"):
        assert np.allclose(h[k].numpy(), z[f"{name}_{k}_long"], rtol=1e-11), k
    leaves = [a.clone().requires_grad_(True) for a in arrays]
    grads = torch.autograd.grad(heads(leaves)["lqn"].mean(), leaves)
    for (s, k), g in zip(z[f"{name}_fd_index"], z[f"{name}_fd_grad"]):
        mine = float(grads[int(s)].reshape(-1)[int(k)])
        assert abs(mine - g) <= 1e-6 * max(1.0, abs(g)), (int(s), int(k), mine, g)
