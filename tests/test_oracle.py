"""Pins for the CPU oracle (the reference holds no golden vectors for this path, so the
oracle is validated by independent routes -- SURVEY.md 8c i-v -- and by the reference's
own property tests re-expressed against it)."""
import math

import numpy as np
import pytest
import torch

from oracle import tn_oracle as O

torch.manual_seed(0)
F64 = torch.float64

EMBS = [
    O.EmbSpec("trig", k=1),
    O.EmbSpec("trig", k=3),
    O.EmbSpec("fourier", p=8),
    O.EmbSpec("fourier", p=3),
    O.EmbSpec("rbf", gamma=0.7, centers=(0.0, 0.25, 0.5, 1.0)),
]


@pytest.mark.parametrize("spec", EMBS, ids=lambda s: f"{s.kind}{s.dim}")
def test_feature_maps_unit_norm(spec):
    # reference: tests/test_embeddings.py:10-129 (||phi(x)|| == approx(1))
    x = torch.linspace(0, 1, 101, dtype=F64)
    phi = O.feature_map(x, spec)
    assert phi.shape == (101, spec.dim)
    assert torch.allclose(phi.norm(dim=-1), torch.ones(101, dtype=F64), atol=1e-12)


def test_fourier_dirichlet_identity():
    # closed form of embeddings.py:407-423: |sin(pi y)/(p sin(pi y/p))|, y=(p-1)x-j
    p = 8
    x = torch.tensor(np.random.default_rng(0).random(1000), dtype=F64)
    phi = O.feature_map(x, O.EmbSpec("fourier", p=p))
    j = torch.arange(p, dtype=F64)
    y = (p - 1) * x[:, None] - j
    closed = torch.abs(torch.sin(math.pi * y) / (p * torch.sin(math.pi * y / p)))
    assert torch.allclose(phi, closed, atol=1e-12)


def test_trig_order_all_cos_then_all_sin():
    # embeddings.py:340-351: itertools.product([cos, sin], range(1,k+1))
    phi = O.feature_map(torch.tensor([0.3], dtype=F64), O.EmbSpec("trig", k=2))[0]
    exp = np.array([math.cos(math.pi * 0.3 / 2), math.cos(math.pi * 0.3 / 4),
                    math.sin(math.pi * 0.3 / 2), math.sin(math.pi * 0.3 / 4)]) / math.sqrt(2)
    assert np.allclose(phi.numpy(), exp, atol=1e-15)


def test_poly_and_rbf_values():
    phi = O.feature_map(torch.tensor([0.5], dtype=F64), O.EmbSpec("poly", degree=2, include_bias=True))[0]
    assert np.allclose(phi.numpy(), [1.0, 0.5, 0.25])
    phi = O.feature_map(torch.tensor([0.5], dtype=F64), O.EmbSpec("poly", degree=3))[0]
    assert np.allclose(phi.numpy(), [0.5, 0.25, 0.125])
    v = np.exp(-2.0 * (0.5 - np.array([0.0, 1.0])))  # no square: embeddings.py:601
    phi = O.feature_map(torch.tensor([0.5], dtype=F64), O.EmbSpec("rbf", gamma=2.0, centers=(0.0, 1.0)))[0]
    assert np.allclose(phi.numpy(), v / np.linalg.norm(v))


@pytest.mark.parametrize("L", [5, 30, 250])
@pytest.mark.parametrize("spec", EMBS + [O.EmbSpec("poly", degree=2, include_bias=True)],
                         ids=lambda s: f"{s.kind}{s.dim}")
def test_embed_norm_is_one(spec, L):
    # reference: tests/test_embeddings.py:132-173 (embed(x).norm() == approx(1)); both branches
    x = O.make_inputs(4, L, seed=1)
    phi = O.embed_batch(x, spec)
    norm = phi.norm(dim=-1).log().sum(dim=1).exp()
    assert torch.allclose(norm, torch.ones(4, dtype=F64), atol=1e-10)


def _mps_case(L=6, chi=3, d=2, cls=None, C=None, squeeze=False, shape_method="even", seed=3):
    return O.make_mps_arrays(L, chi, d, class_site=cls, C=C, std=0.3, seed=seed,
                             squeeze_ends=squeeze, shape_method=shape_method)


@pytest.mark.parametrize("order", ["gemm", "ref"])
def test_mps_chain_equals_dense_classifier(order):
    arrays = _mps_case(L=7, chi=4, d=3, cls=3, C=5)
    x = O.make_inputs(5, 7, seed=11)
    phi = O.embed_batch(x, O.EmbSpec("poly", degree=2, include_bias=True))
    y = O.mps_outputs(phi, arrays, order=order)
    yd = O.mps_dense_outputs(phi, arrays)
    assert y.shape == (5, 5)
    assert torch.allclose(y, yd, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("cls", [0, 6])
def test_mps_class_site_at_boundary(cls):
    arrays = _mps_case(L=7, chi=3, d=2, cls=cls, C=4)
    phi = O.embed_batch(O.make_inputs(3, 7, seed=5), O.EmbSpec("trig"))
    assert torch.allclose(O.mps_outputs(phi, arrays), O.mps_dense_outputs(phi, arrays), rtol=1e-12, atol=1e-14)


def test_mps_chain_equals_dense_plain_squeezed_and_noteven():
    for sm in ("even", "noteven"):
        arrays = _mps_case(L=8, chi=4, d=2, squeeze=True, shape_method=sm)
        assert arrays[0].ndim == 2 and arrays[-1].ndim == 2
        phi = O.embed_batch(O.make_inputs(6, 8, seed=2), O.EmbSpec("fourier", p=2))
        y = O.mps_outputs(phi, arrays)
        assert y.shape == (6, 1)
        assert torch.allclose(y, O.mps_dense_outputs(phi, arrays), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("L,S,do", [(8, 4, 2), (9, 4, 3), (6, 1, 2), (7, 2, 2), (8, 8, 2)])
def test_smpo_window_recursion_equals_dense(L, S, do):
    # double-layer recursion == explicit v.v  (SURVEY 8c v); (L-1)%S==0 puts an output leg on the last site
    arrays = O.make_smpo_arrays(L, 3, 2, do, S, std=0.3, seed=9)
    phi = O.embed_batch(O.make_inputs(4, L, seed=8), O.EmbSpec("trig"))
    n = O.smpo_sqnorm(phi, arrays, S)
    nd = O.smpo_dense_sqnorm(phi, arrays, S)
    assert torch.allclose(n, nd, rtol=1e-12)
    assert (n >= 0).all()  # tests/test_metrics.py:117-172 (TransformedSquaredNorm >= 0)


HEAD_CASES = [("softmax_xent", None), ("softmax_xent_weighted", (0.3, 1.7, 1.0, 2.0)),
              ("xent_softmax_argmax", None), ("mse", None)]


@pytest.mark.parametrize("head,cw", HEAD_CASES, ids=[h for h, _ in HEAD_CASES])
def test_mps_grad_autograd_vs_env_vs_fd(head, cw):
    L, C = 6, 4
    arrays = _mps_case(L=L, chi=3, d=2, cls=2, C=C)
    x = O.make_inputs(5, L, seed=4)
    t = O.make_labels(5, C)
    prob = O.Problem("mps", O.EmbSpec("trig"), head, class_weights=cw)
    loss, grads = O.value_and_grad(prob, x, arrays, t)
    # route iii: explicit environments, seeded with autograd through the head only
    phi = O.embed_batch(x, prob.emb)
    y = O.mps_outputs(phi, arrays).detach().requires_grad_(True)
    cwt = None if cw is None else torch.tensor(cw, dtype=F64)
    O.head_loss(y, head, t, cwt).mean().backward()
    g_env = O.mps_env_vjp(phi, arrays, y.grad)
    for g, ge in zip(grads, g_env):
        assert torch.allclose(g, ge, rtol=1e-10, atol=1e-14)
    # route ii: central finite differences on a few entries of every site
    rng = np.random.default_rng(0)
    for i, a in enumerate(arrays):
        for _ in range(3):
            idx = tuple(int(rng.integers(0, s)) for s in a.shape)
            h = 1e-6
            ap = [q.clone() for q in arrays]
            am = [q.clone() for q in arrays]
            ap[i][idx] += h
            am[i][idx] -= h
            fd = (O.loss_fn(prob, x, ap, t) - O.loss_fn(prob, x, am, t)) / (2 * h)
            assert abs(fd - grads[i][idx]) <= 1e-6 * max(1.0, abs(fd)), (i, idx, fd, grads[i][idx])


def test_nll_grad_and_nonneg():
    L = 6
    arrays = O.make_mps_arrays(L, 4, 8, std=0.2, seed=1, squeeze_ends=True)
    x = O.make_inputs(7, L, seed=6)
    prob = O.Problem("mps", O.EmbSpec("fourier", p=8), "nll")
    per = O.per_sample_loss(prob, x, arrays)
    # normalised model and data => |<A|phi>| <= 1 => NLL >= 0  (tests/test_metrics.py:40-113)
    assert (per >= 0).all()
    loss, grads = O.value_and_grad(prob, x, arrays)
    phi = O.embed_batch(x, prob.emb)
    y = O.mps_outputs(phi, arrays)
    dy = (-2.0 / y) / x.shape[0]
    for g, ge in zip(grads, O.mps_env_vjp(phi, arrays, dy)):
        assert torch.allclose(g, ge, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("head", ["tsn", "logquad", "quad"])
@pytest.mark.parametrize("L,S", [(8, 4), (9, 4), (6, 1)])
def test_smpo_grad_autograd_vs_env(head, L, S):
    arrays = O.make_smpo_arrays(L, 3, 2, 2, S, std=0.3, seed=5)
    x = O.make_inputs(4, L, seed=12)
    prob = O.Problem("smpo", O.EmbSpec("trig"), head, spacing=S)
    loss, grads = O.value_and_grad(prob, x, arrays)
    phi = O.embed_batch(x, prob.emb)
    n = O.smpo_sqnorm(phi, arrays, S).detach().requires_grad_(True)
    O.head_loss(n, head).mean().backward()
    for g, ge in zip(grads, O.smpo_env_vjp(phi, arrays, S, n.grad)):
        assert g.shape == ge.shape
        assert torch.allclose(g, ge, rtol=1e-9, atol=1e-13)


def test_ref_order_matches_gemm_order():
    arrays = _mps_case(L=12, chi=5, d=3, cls=5, C=3)
    x = O.make_inputs(9, 12, seed=3)
    t = O.make_labels(9, 3)
    prob = O.Problem("mps", O.EmbSpec("poly", degree=3), "softmax_xent")
    l1, g1 = O.value_and_grad(prob, x, arrays, t, order="gemm")
    l2, g2 = O.value_and_grad(prob, x, arrays, t, order="ref")
    assert abs(l1 - l2) < 1e-13
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-10, atol=1e-15)


def test_model_init_norm_is_one():
    # tests/test_mps.py:14-44 (norm() == approx(1) after init)
    arrays = O.make_mps_arrays(20, 6, 2, class_site=9, C=3)
    assert abs(float(O.mps_model_sqnorm(arrays)) - 1.0) < 1e-10


def test_sweeps_restatement_preserves_the_model_then_descends():
    """oracle.sweeps_train (strategy.py:51-258, model.py:753-773, 829-866): contracting and splitting a pair with max_bond = the
    bond's size is exact, so the first pair step sees the loss of the given model; later steps descend; the two-site gradient
    agrees with central finite differences of the loss of the merged tensor."""
    arrays = O.make_mps_arrays(6, 4, 2, class_site=2, C=3, std=0.2, seed=3)
    x, t = O.make_inputs(15, 6, seed=5), O.make_labels(15, 3)
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    l0 = float(O.loss_fn(prob, x, arrays, t))
    arr, hist, steps = O.sweeps_train(prob, x, arrays, t, epochs=2, lr=0.05)
    assert abs(steps[0] - l0) < 1e-12 and hist[1] < hist[0] and len(steps) == 2 * 10
    assert [tuple(a.shape) for a in arr] == [tuple(a.shape) for a in arrays]
    sites, cls = O.canon_mps_sites(arrays)
    for i in (1, 2, 4):  # class site on the right of the pair, on the left, away from it
        loss, T, g = O.pair_loss_and_grad(prob, x, sites, cls, i, t)
        assert abs(float(loss) - l0) < 1e-12
        rng = np.random.default_rng(i)
        for _ in range(4):
            idx = tuple(int(rng.integers(n)) for n in T.shape)
            h = 1e-6

            def f(delta):
                Tn = T.clone()
                Tn[idx] += delta
                A, B = sites[i], sites[i + 1]
                cl, cr, d, _, nk = Tn.shape
                left = Tn.permute(0, 1, 3, 4, 2).reshape(cl, cr * d * nk, d, 1)
                sel = torch.eye(cr * d * nk, dtype=Tn.dtype).reshape(cr * d * nk, cr, d, nk)
                new = list(sites)
                new[i], new[i + 1] = left, sel
                ncls = i + 1 if nk > 1 else cls
                return float(O.loss_fn(prob, x, [a if j == ncls else a[..., 0] for j, a in enumerate(new)], t))

            fd = (f(h) - f(-h)) / (2 * h)
            assert abs(fd - float(g[idx])) < 1e-7 * max(1.0, abs(fd))
