"""The drop-in surface on a GPU: configure -> train -> evaluate / forward / accuracy / predict,
for the model kinds of the BASELINE configs (reference: tests/test_model.py:468-613 checks that
train() returns a finite history of the right length; here values are also checked against the oracle)."""
import numpy as np
import pytest
import torch

import tn4ml_b200 as T
from oracle import tn_oracle as O
from tn4ml_b200 import metrics as M

pytestmark = pytest.mark.gpu


def test_train_unsupervised_nll_history_and_descent():
    import tn4ml_b200 as T
    from tn4ml_b200.initializers import randn
    model = T.MPS_initialize(8, initializer=randn(0.3), key=7, bond_dim=6, phys_dim=2, device="cuda:0")
    model.configure(optimizer=T.optim.adam, strategy="global", loss=T.metrics.NegLogLikelihood,
                    train_type=T.TrainingType.UNSUPERVISED, learning_rate=5e-2, device=("gpu", 0))
    x = np.random.default_rng(0).random((40, 8))
    hist = model.train(x, batch_size=16, epochs=6, embedding=T.TrigonometricEmbedding(), normalize=True,
                       dtype=np.float64)
    assert len(hist["loss"]) == 6 and len(hist["epoch_time"]) == 6 and np.all(np.isfinite(hist["loss"]))
    assert hist["loss"][-1] < hist["loss"][0]
    assert abs(float(model.norm()) - 1.0) < 1e-9              # normalize=True keeps <A|A> = 1
    scores = model.evaluate(x, batch_size=16, embedding=T.TrigonometricEmbedding(), return_list=True,
                            metric=T.metrics.NegLogLikelihood, dtype=np.float64)
    assert scores.shape == (40,) and np.all(scores >= 0)       # tests/test_metrics.py:40-113
    ref = O.per_sample_loss(O.Problem("mps", O.EmbSpec("trig"), "nll"), torch.tensor(x),
                            [a.detach().cpu().double() for a in model.arrays])
    assert np.allclose(scores, ref.numpy(), rtol=1e-10)


def test_train_supervised_classifier_first_step_matches_oracle_adam():
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(10, 5, 3, class_site=4, C=3, std=0.2, seed=5)
    x, t = O.make_inputs(24, 10, seed=1), O.make_labels(24, 3)
    model = T.TensorNetwork(arrays, device="cuda:0")

    def loss(*a, **k):
        return T.metrics.OptaxWrapper(T.metrics.softmax_cross_entropy)(*a, **k).mean()

    model.configure(optimizer=T.optim.adam, loss=loss, train_type=T.TrainingType.SUPERVISED, learning_rate=1e-2,
                    device=("gpu", 0))
    emb = T.PolynomialEmbedding(degree=2, n=1, include_bias=True)
    hist = model.train(x.numpy(), targets=t.numpy(), batch_size=24, epochs=1, embedding=emb, dtype=np.float64)
    prob = O.Problem("mps", O.EmbSpec("poly", degree=2, include_bias=True), "softmax_xent")
    lref, gref = O.value_and_grad(prob, x, arrays, t)
    assert abs(hist["loss"][0] - float(lref)) < 1e-10
    for a_new, a_old, g in zip(model.arrays, arrays, gref):    # first Adam step: -lr * g / (|g| + eps)
        exp = a_old - 1e-2 * g / (g.abs() + 1e-8)
        mask = g.abs() > 1e-6                                   # sign(g) is ill-defined where g ~ 0
        assert torch.allclose(a_new.cpu()[mask], exp[mask], atol=1e-7)
    acc = model.accuracy(x.numpy(), t.numpy(), embedding=emb, batch_size=7, dtype=np.float64)
    assert 0.0 <= acc <= 1.0
    y = model.forward(x.numpy(), embedding=emb, batch_size=9, dtype=np.float64)
    yref = O.mps_outputs(O.embed_batch(x, prob.emb), [a.detach().cpu() for a in model.arrays])
    assert torch.allclose(y.cpu(), yref, rtol=1e-9, atol=1e-300)
    yn = model.predict(x[0].numpy(), embedding=emb, normalize=True)
    assert torch.allclose(yn.cpu(), yref[0] / yref[0].norm(), rtol=1e-9)
    with pytest.raises(ValueError):
        model.predict(x[0, :5].numpy(), embedding=emb)         # model.py:306-307


def test_train_smpo_combined_loss_with_regulariser():
    import tn4ml_b200 as T
    from tn4ml_b200.initializers import randn
    model = T.SMPO_initialize(16, randn(0.5), 3, spacing=4, bond_dim=4, phys_dim=(2, 2), device="cuda:0")

    def loss_combined(m, data, y_true=None, embedding=None):   # examples/unsupervised/mnist_ad.py:47-57
        return T.metrics.CombinedLoss(m, data, y_true, error=T.metrics.LogQuadNorm,
                                      reg=lambda P: 0.4 * T.metrics.LogReLUFrobNorm(P), embedding=embedding)

    model.configure(optimizer=T.optim.adam, loss=loss_combined, train_type=T.TrainingType.UNSUPERVISED,
                    learning_rate=1e-2, device=("gpu", 0))
    x = np.random.default_rng(1).random((12, 16))
    arrays0 = [a.detach().cpu().clone() for a in model.arrays]
    hist = model.train(x, batch_size=12, epochs=3, embedding=T.TrigonometricEmbedding(), dtype=np.float64)
    assert len(hist["loss"]) == 3 and np.all(np.isfinite(hist["loss"]))
    prob = O.Problem("smpo", O.EmbSpec("trig"), "logquad", spacing=4)
    l0 = float(O.loss_fn(prob, torch.tensor(x), arrays0))      # ||P|| = 1 after init, so the regulariser is 0
    assert abs(hist["loss"][0] - l0) < 1e-9 * max(1.0, abs(l0))


def test_float32_model_uses_tensor_path_by_default_and_trains():
    import tn4ml_b200 as T
    from tn4ml_b200 import _lib
    arrays = O.make_mps_arrays(12, 64, 2, class_site=5, C=4, std=0.05, seed=2, dtype=torch.float32)
    x, t = O.make_inputs(256, 12, seed=3), O.make_labels(256, 4)
    model = T.TensorNetwork(arrays, dtype=torch.float32, device="cuda:0")
    assert model.precision == _lib.PREC_TENSOR
    model.configure(optimizer=T.optim.adam, loss=T.metrics.CrossEntropySoftmax, train_type=T.TrainingType.SUPERVISED,
                    learning_rate=1e-3, device=("gpu", 0))
    hist = model.train(x.numpy(), targets=t.numpy(), batch_size=128, epochs=4, embedding=T.TrigonometricEmbedding(),
                       dtype=np.float32)
    assert np.all(np.isfinite(hist["loss"])) and hist["loss"][-1] < hist["loss"][0]


def test_cfg1_bond_float32_trains_on_the_register_resident_kernels():
    """BASELINE configs[0]'s shape family (chi = 10, 10 classes, batch 64) in float32 through Model.train(): the step runs
    zero-padded on the mma.sync kernels, the model keeps its own [10][10][2] site tensors, the loss goes down, and training in
    float32 follows the same run in float64 (SIMT family) to float32 accuracy."""
    import tn4ml_b200 as T
    from tests.test_gpu_tensor_path import _launched_kernels
    arrays = O.make_mps_arrays(24, 10, 2, class_site=11, C=10, std=0.05, seed=2)
    x, t = O.make_inputs(128, 24, seed=3), O.make_labels(128, 10)
    losses = {}
    for dt, npdt in ((torch.float32, np.float32), (torch.float64, np.float64)):
        model = T.TensorNetwork([a.to(dt) for a in arrays], dtype=dt, device="cuda:0")
        model.configure(optimizer=T.optim.adam, loss=T.metrics.CrossEntropySoftmax, train_type=T.TrainingType.SUPERVISED,
                        learning_rate=1e-3, device=("gpu", 0))
        run = lambda: losses.__setitem__(dt, model.train(x.numpy(), targets=t.numpy(), batch_size=64, epochs=5,  # noqa: E731
                                                         embedding=T.TrigonometricEmbedding(), dtype=npdt)["loss"])
        seen = _launched_kernels(run)
        if dt == torch.float32:
            assert "rr_chain_kernel(fwd)" in seen and "rr_da_kernel" in seen, seen
            assert tuple(model.arrays[3].shape) == (10, 10, 2)
    l32, l64 = np.asarray(losses[torch.float32]), np.asarray(losses[torch.float64])
    assert np.all(np.isfinite(l32)) and l32[-1] < l32[0]
    assert np.allclose(l32, l64, rtol=2e-4)


def test_missing_library_fails_loudly(monkeypatch):
    from tn4ml_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtn4ml_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_fused_adam_matches_the_optax_rule(dtype):
    """tnb_adam_update (one kernel on the packed buffer) against the same rule written with tensor ops."""
    import tn4ml_b200 as T
    g = torch.Generator(device="cpu").manual_seed(0)
    p0 = torch.randn(10007, generator=g, dtype=dtype).cuda()
    opt = T.optim.adam(T.optim.exponential_decay(3e-3, 2, 0.9), b1=0.8, b2=0.95, eps=1e-7, eps_root=1e-12)
    pa, pb = p0.clone(), p0.clone()
    sa, sb = opt.init(pa), opt.init(pb)
    for step in range(5):
        grad = torch.randn(10007, generator=g, dtype=dtype).cuda()
        scale = 0.5 if step == 3 else 1.0
        sa = opt.fused_step(pa, grad, sa, grad_scale=scale)
        upd, sb = opt.update(grad * scale, sb, pb)
        T.optim.apply_updates(pb, upd)
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    assert torch.allclose(pa, pb, rtol=tol, atol=tol)
    assert torch.allclose(sa["mu"], sb["mu"], rtol=tol, atol=tol) and torch.allclose(sa["nu"], sb["nu"], rtol=tol, atol=tol)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_gradient_clip_is_per_site_like_the_reference(dtype):
    """util.gradient_clip (tn4ml/util.py:131-155): every site tensor's gradient scaled by min(1, thr / (||g_site|| + 1e-6))
    -- NOT one global-norm clip."""
    arrays = O.make_mps_arrays(9, 12, 3, class_site=4, C=3, std=0.1, seed=1, shape_method="noteven")
    model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
    eng, _ = model._engine(T.PolynomialEmbedding(degree=2, n=1, include_bias=True),
                           lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean(), True, dtype)
    g = torch.zeros(eng.n_params, dtype=dtype, device="cuda:0")
    gen = torch.Generator().manual_seed(5)
    scales = [0.01, 5.0, 0.3, 40.0, 1.0, 0.0, 2.0, 0.5, 7.0]  # some sites below, some above the threshold; one all-zero
    for v, sc in zip(model._views(g), scales):
        v.copy_((torch.randn(v.shape, generator=gen, dtype=torch.float64) * sc / max(v.numel(), 1) ** 0.5).to(dtype))
    ref = []
    for v in model._views(g):
        n = float(torch.linalg.norm(v.double()))
        ref.append(v.double().cpu() * min(1.0, 1.5 / (n + 1e-6)))
    eng.gradient_clip(g, 1.5)
    tol = 1e-12 if dtype == torch.float64 else 1e-6
    for v, r in zip(model._views(g), ref):
        assert float((v.double().cpu() - r).abs().max()) <= tol * max(1.0, float(r.abs().max()))
    with pytest.raises(AssertionError):
        eng.gradient_clip(g, 0.0)


def test_engine_cache_is_keyed_on_the_embedding_contents():
    """two temporaries with different parameters must not share an engine."""
    arrays = O.make_mps_arrays(6, 4, 2, std=0.1, seed=1, squeeze_ends=True, identity_all=True)
    model = T.MatrixProductState(arrays, dtype=torch.float64, device="cuda:0")
    x = O.make_inputs(5, 6, seed=2)
    import numpy as np
    a = model.evaluate(x.numpy(), embedding=T.GaussianRBFEmbedding(centers=np.array([0.1, 0.9]), gamma=0.5), metric=M.NegLogLikelihood,
                       return_list=True)
    b = model.evaluate(x.numpy(), embedding=T.GaussianRBFEmbedding(centers=np.array([0.3, 0.6]), gamma=2.0), metric=M.NegLogLikelihood,
                       return_list=True)
    prob = O.Problem("mps", O.EmbSpec("rbf", gamma=2.0, centers=(0.3, 0.6)), "nll")
    ref = O.per_sample_loss(prob, x, arrays)
    assert not np.allclose(a, b)
    assert np.allclose(b, ref.numpy(), rtol=1e-10)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("head", ["nll", "softmax_xent"])
def test_model_shorter_than_the_data(head, dtype):
    """metrics.py:36-43 / 448-455: len(model) < len(data) -- the data sites beyond the model are summed over their physical
    index (a per-sample factor; embed() normalises over ALL data sites)."""
    from tests.test_gpu_parity import _heads
    Lm, Ld, B = 7, 10, 9
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    if head == "nll":
        arrays = O.make_mps_arrays(Lm, 5, 3, std=0.2, seed=4, squeeze_ends=True, identity_all=True)
        model, t = T.MatrixProductState(arrays, dtype=dtype, device="cuda:0"), None
    else:
        arrays = O.make_mps_arrays(Lm, 5, 3, class_site=3, C=3, std=0.2, seed=4)
        model, t = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0"), O.make_labels(B, 3)
    x = O.make_inputs(B, Ld, seed=6)
    spec = O.EmbSpec("poly", degree=2, include_bias=True)
    prob = O.Problem("mps", spec, head)
    arr = [a.to(dtype).double() for a in arrays]
    xr = x.to(dtype).double()
    loss_ref, grads_ref = O.value_and_grad(prob, xr, arr, t)
    per_ref = O.per_sample_loss(prob, xr, arr, t)
    eng, _ = model._engine(T.PolynomialEmbedding(degree=2, n=1, include_bias=True), _heads()[head], t is not None, dtype)
    ls, g = eng.value_and_grad(model._packed, x.to(dtype), None if t is None else t.to(dtype))
    assert abs(float(ls) / B - float(loss_ref)) <= tol * max(1.0, abs(float(loss_ref)))
    gmax = max(float(q.abs().max()) for q in grads_ref)
    for a, b in zip(model._views(g), grads_ref):
        assert float((a.double().cpu() - b).abs().max()) <= tol * gmax
    per, _, _ = eng.forward(model._packed, x.to(dtype), None if t is None else t.to(dtype))
    assert torch.allclose(per.double().cpu(), per_ref, rtol=tol * 10, atol=tol * 10)
    with pytest.raises(ValueError):  # metrics.py:49-51: the model must not be longer than the data
        eng.forward(model._packed, x[:, :5].to(dtype), None if t is None else t.to(dtype))


@pytest.mark.parametrize("dtype,chi,S", [(torch.float64, 8, 4), (torch.float32, 32, 8), (torch.float32, 12, 1)],
                         ids=["f64_shared_memory_family", "f32_register_resident_family", "f32_mpo"])
def test_semi_supervised_loss_matches_oracle(dtype, chi, S):
    """SemiSupervisedLoss (metrics.py:209-233): per-sample values through evaluate(), and loss + gradient through one plain
    gradient-descent step of train() (params_new = params - lr * grad), against the oracle's autograd."""
    L, B, lr = 16, 12, 1e-3
    arrays = [a.to(dtype) for a in O.make_smpo_arrays(L, chi, 2, 2, S, std=0.2 if chi <= 12 else 0.05, seed=8)]
    x = O.make_inputs(B, L, seed=9).to(dtype)
    y = torch.tensor(np.random.default_rng(3).integers(0, 2, size=B), dtype=dtype)
    cls = T.MatrixProductOperator if S == 1 else T.SpacedMatrixProductOperator
    model = cls(arrays, dtype=dtype, device="cuda:0", **({} if S == 1 else {"spacing": S}))
    model.configure(optimizer=T.optim.sgd, loss=M.SemiSupervisedLoss, train_type=T.TrainingType.SUPERVISED,
                    learning_rate=lr, device=("gpu", 0))
    emb = O.EmbSpec("trig")
    arr64 = [a.double() for a in arrays]
    trace = model._reg_norm_is_trace
    per_ref = O.semisup_per_sample(emb, x.double(), arr64, S, y.double(), norm_is_trace=trace)
    per = model.evaluate(x.cpu().numpy(), y.cpu().numpy(), batch_size=5, evaluate_type=T.TrainingType.SUPERVISED,
                         return_list=True, metric=M.SemiSupervisedLoss, dtype=dtype)
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    assert np.allclose(per, per_ref.numpy(), rtol=tol * 10, atol=tol)
    loss_ref, grads_ref = O.semisup_value_and_grad(emb, x.double(), arr64, S, y.double(), norm_is_trace=trace)
    hist = model.train(x.cpu().numpy(), targets=y.cpu().numpy(), batch_size=B, epochs=1, dtype=dtype)
    assert abs(hist["loss"][0] - float(loss_ref)) <= tol * max(1.0, abs(float(loss_ref)))
    gmax = max(float(g.abs().max()) for g in grads_ref)
    for a_new, a_old, g in zip(model.arrays, arrays, grads_ref):
        got = (a_old.double() - a_new.double().cpu()) / lr          # the gradient the step applied
        assert float((got - g).abs().max()) <= (1e-7 if dtype == torch.float64 else 2e-3) * gmax
