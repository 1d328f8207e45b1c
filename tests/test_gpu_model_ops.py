"""Model-only operations of the training loop on the device: norm(), normalize() on both branches,
canonicalize(center) and the Frobenius regularisers with their gradients (tnb_norm), against the oracle."""
import math

import numpy as np
import pytest
import torch

import tn4ml_b200 as T
from oracle import tn_oracle as O
from tn4ml_b200 import metrics as M

pytestmark = pytest.mark.gpu


def _models(dtype):
    out = []
    a = [0.9 * t for t in O.make_mps_arrays(12, 6, 3, class_site=5, C=4, std=0.3, seed=1, shape_method="noteven")]
    out.append(("classifier_noteven", T.TensorNetwork(a, dtype=dtype, device="cuda:0"), O.canon_mps_sites(a)[0], False))
    a = [1.2 * t for t in O.make_mps_arrays(9, 8, 2, std=0.3, seed=2, squeeze_ends=True)]  # (log norm away from the ReLU kink)
    out.append(("plain_mps", T.MatrixProductState(a, dtype=dtype, device="cuda:0"), O.canon_mps_sites(a)[0], False))
    a = O.make_smpo_arrays(17, 6, 2, 3, 4, std=0.3, seed=3)
    out.append(("smpo", T.SpacedMatrixProductOperator(a, spacing=4, dtype=dtype, device="cuda:0"), O.canon_smpo_sites(a, 4), True))
    a = O.make_smpo_arrays(6, 5, 2, 2, 1, std=0.3, seed=4)
    out.append(("mpo", T.MatrixProductOperator(a, dtype=dtype, device="cuda:0"), O.canon_smpo_sites(a, 1), False))
    a = [1.1 * t for t in O.make_mps_arrays(7, 130, 2, class_site=3, C=3, std=0.05, seed=5)]  # three tiles per edge, ragged last tile
    out.append(("chi130", T.TensorNetwork(a, dtype=dtype, device="cuda:0"), O.canon_mps_sites(a)[0], False))
    a = [0.3 * t for t in O.make_mps_arrays(784, 4, 2, std=0.2, seed=6, squeeze_ends=True)]  # <A|A> ~ 1e-800: log-space only
    out.append(("L784", T.MatrixProductState(a, dtype=dtype, device="cuda:0"), O.canon_mps_sites(a)[0], False))
    return out


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_norm_and_regularisers_match_the_oracle(dtype):
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    for name, model, sites4, trace in _models(dtype):
        sites4 = [s.to(dtype).double() for s in sites4]
        ref = float(O.model_log_sqnorm(sites4))
        l2, _ = model._log2_norm2()
        got = float(l2) * math.log(2.0)
        assert abs(got - ref) <= tol * max(1.0, abs(ref)), (name, got, ref)
        assert model._reg_norm_is_trace == trace
        for kind in ("log_frob", "log_pow_frob", "log_relu_frob", "quad_frob"):
            leaves = [s.clone().requires_grad_(True) for s in sites4]
            val = 0.7 * O.regulariser(leaves, kind, trace)
            grads_ref = torch.autograd.grad(val, leaves)
            expr = M.LossExpr(0, None, [M.RegExpr(kind, 0.7)])
            g = torch.zeros_like(model._packed)
            v = model._reg_value_and_grad(expr, g)
            assert abs(float(v) - float(val.detach())) <= tol * max(1.0, abs(float(val.detach()))), (name, kind)
            gmax = max(float(q.abs().max()) for q in grads_ref) or 1.0
            for i, q in enumerate(grads_ref):
                cl, cr, d, n = q.shape
                mine = model._canon4_of(g, i).double().cpu()
                assert float((mine - q).abs().max()) <= tol * 10 * gmax, (name, kind, i)
            # nothing leaks into the zero padding of the packed layout
            inside = sum(float(model._canon4_of(g, i).double().abs().sum()) for i in range(model.L))
            assert abs(float(g.double().abs().sum()) - inside) <= 1e-6 * float(g.double().abs().sum() + 1e-30)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_normalize_short_chain_spreads_the_norm_over_the_sites(dtype):
    """tn.py:103-107: every tensor divided by norm^(1/L); insert=i puts the whole factor on one site."""
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    a = [2.0 * t for t in O.make_mps_arrays(10, 6, 2, class_site=4, C=3, std=0.3, seed=7)]
    model = T.TensorNetwork(a, dtype=dtype, device="cuda:0")
    n0 = float(model.norm())
    before = [v.clone() for v in model.arrays]
    model.normalize()
    assert abs(float(model.norm()) - 1.0) <= tol
    f = n0 ** (-1.0 / model.L)
    for b, v in zip(before, model.arrays):
        assert torch.allclose(v, b * f, rtol=tol * 10, atol=0)
    model2 = T.TensorNetwork(a, dtype=dtype, device="cuda:0")
    model2.normalize(insert=3)
    assert abs(float(model2.norm()) - 1.0) <= tol
    assert torch.allclose(model2.arrays[3], before[3] / n0, rtol=tol * 10, atol=0)
    assert torch.equal(model2.arrays[2], before[2])
    with pytest.raises(IndexError):
        model2.normalize(insert=10)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_normalize_long_chain_is_the_reference_left_canonical_sweep(dtype):
    """L > 200 (tn.py:93-101): site i / ||site i||, then left_canonize_site(i).  Afterwards <A|A> = 1, every site but the last
    is a left isometry, and the state is the old one up to a positive factor (checked through the contraction itself)."""
    tol = 1e-10 if dtype == torch.float64 else 2e-4
    L, chi = 210, 6
    a = O.make_mps_arrays(L, chi, 2, std=0.2, seed=8, squeeze_ends=True, identity_all=True)
    model = T.MatrixProductState(a, dtype=dtype, device="cuda:0")
    x = O.make_inputs(5, L, seed=9)
    eng, _ = model._engine(T.TrigonometricEmbedding(), M.NegLogLikelihood, False, dtype)
    per0, _, _ = eng.forward(model._packed, x)
    model.normalize()
    assert abs(float(model.norm()) - 1.0) <= tol
    for i in range(L - 1):
        A = model._canon4(i).double()
        cl, cr, d, n = A.shape
        G = torch.einsum("lrsk,lmsk->rm", A, A)
        k = min(cl * d * n, cr)
        assert float((G[:k, :k] - torch.eye(k, device=G.device, dtype=G.dtype)).abs().max()) <= tol * 10, i
    per1, _, _ = eng.forward(model._packed, x)
    shift = (per1 - per0).double()  # -log s^2 moves by the same constant for every sample
    assert float((shift - shift.mean()).abs().max()) <= tol * 100 * max(1.0, float(per0.abs().max()))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_canonicalize_keeps_the_state_and_orthogonalises_around_the_center(dtype):
    tol = 1e-10 if dtype == torch.float64 else 2e-4
    L, c = 11, 4
    a = O.make_mps_arrays(L, 7, 3, class_site=6, C=3, std=0.3, seed=10)
    model = T.TensorNetwork(a, dtype=dtype, device="cuda:0")
    x = O.make_inputs(6, L, seed=11)
    emb = T.PolynomialEmbedding(degree=2, n=1, include_bias=True)
    out0 = model.forward(x.numpy(), embedding=emb, batch_size=6)
    n0 = float(model.norm())
    model.canonicalize(c)
    out1 = model.forward(x.numpy(), embedding=emb, batch_size=6)
    assert torch.allclose(out0, out1, rtol=tol * 100, atol=tol * float(out0.abs().max()))
    assert abs(float(model.norm()) - n0) <= tol * 10 * n0
    for i in range(L):
        A = model._canon4(i).double()
        if i < c:
            G = torch.einsum("lrsk,lmsk->rm", A, A)
            k = min(A.shape[0] * A.shape[2] * A.shape[3], A.shape[1])
        elif i > c:
            G = torch.einsum("lrsk,mrsk->lm", A, A)
            k = min(A.shape[1] * A.shape[2] * A.shape[3], A.shape[0])
        else:
            continue
        assert float((G[:k, :k] - torch.eye(k, device=G.device, dtype=G.dtype)).abs().max()) <= tol * 10, i


def test_train_with_regulariser_normalize_and_canonize_runs_and_descends():
    """CombinedLoss with alpha * LogReLUFrobNorm (metrics.py:535-592, the cfg-3 example's loss) + normalize + canonize."""
    from tn4ml_b200.initializers import randn
    model = T.SMPO_initialize(L=16, initializer=randn(1e-1), key=torch.Generator().manual_seed(3), shape_method="even", spacing=4,
                              bond_dim=6, phys_dim=(2, 2), add_identity=True, dtype=torch.float64)
    rng = np.random.default_rng(0)
    x = 1.0 - rng.random((64, 16))

    def loss(model_, data):
        return M.CombinedLoss(model_, data, error=M.LogQuadNorm, reg=lambda m: 0.4 * M.LogReLUFrobNorm(m))  # mnist_ad.py:202-207

    model.configure(optimizer=T.optim.adam, learning_rate=5e-3, loss=loss, train_type=T.TrainingType.UNSUPERVISED, device=("gpu", 0))
    h = model.train(x, batch_size=32, epochs=6, embedding=T.TrigonometricEmbedding(), normalize=True, canonize=(True, 3))
    assert len(h["loss"]) == 6 and all(np.isfinite(h["loss"]))
    assert h["loss"][-1] < h["loss"][0]
