"""The Sweeps / DMRG-like strategy on the GPU (reference: tn4ml/strategy.py:51-258, tn4ml/models/model.py:753-773, 829-866)
against its CPU restatement in the oracle (oracle/tn_oracle.py: pair_loss_and_grad, sweeps_train).

  * the two-site gradient dLoss/dT of every pair, from the environment kernels + tnb_pair_grad, against autograd;
  * whole training runs: the loss of every epoch and the trained model's outputs.  QR / SVD fix their factors up to signs, so
    site tensors are compared through gauge-invariant quantities (losses, outputs); plain gradient descent and Adam are both
    covariant under those sign changes.
"""
import numpy as np
import pytest
import torch

import tn4ml_b200 as T
from oracle import tn_oracle as O

pytestmark = pytest.mark.gpu

CASES = {
    "plain_nll": dict(L=6, chi=4, d=2, cls=None, C=1, B=20, head="nll", emb=O.EmbSpec("trig")),
    "classifier": dict(L=7, chi=5, d=2, cls=3, C=3, B=23, head="softmax_xent", emb=O.EmbSpec("trig")),
    "classifier_poly_edge": dict(L=5, chi=6, d=3, cls=0, C=2, B=17, head="softmax_xent",
                                 emb=O.EmbSpec("poly", degree=2, include_bias=True)),
}


def _case(name, dtype=torch.float64):
    c = CASES[name]
    if c["cls"] is None:
        arrays = O.make_mps_arrays(c["L"], c["chi"], c["d"], std=0.2, seed=4, squeeze_ends=True, identity_all=True)
        t = None
        model = T.MatrixProductState(arrays, dtype=dtype, device="cuda:0")
    else:
        arrays = O.make_mps_arrays(c["L"], c["chi"], c["d"], class_site=c["cls"], C=c["C"], std=0.2, seed=3)
        t = O.make_labels(c["B"], c["C"])
        model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
    x = O.make_inputs(c["B"], c["L"], seed=6)
    return c, arrays, x, t, model, O.Problem("mps", c["emb"], c["head"])


def _loss_and_emb(c):
    from tests.test_gpu_parity import _emb, _heads
    return _heads()[c["head"]], _emb(c["emb"])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_pair_gradient_matches_oracle(name, dtype):
    """dLoss/dT of every neighbouring pair (class site on the left, on the right, away from the pair) at 1e-10 / 1e-4."""
    c, arrays, x, t, model, prob = _case(name, dtype)
    loss_fn, emb = _loss_and_emb(c)
    model.precision = 0
    eng, _ = model._engine(emb, loss_fn, t is not None, dtype)
    arr = [a.to(dtype).double() for a in arrays]
    sites, cls = O.canon_mps_sites(arr)
    xin = x.to(dtype)
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    worst = 0.0
    for i in range(c["L"] - 1):
        lref, _, gref = O.pair_loss_and_grad(prob, xin.double(), sites, cls, i, t)
        loss_sum, dT = eng.pair_value_and_grad(model._packed, xin, None if t is None else t.to(dtype), i)
        g = dT[:gref.shape[0], :gref.shape[1]].double().cpu()
        assert g.shape == gref.shape
        assert abs(float(loss_sum.item()) / c["B"] - float(lref)) <= tol * max(1.0, abs(float(lref)))
        err = float((g - gref).abs().max()) / float(gref.abs().max())
        worst = max(worst, err)
        assert err <= tol, f"pair ({i}, {i + 1}): {err:.2e}"
        # the padding of the packed layout carries no gradient
        assert float(dT.abs().sum()) == pytest.approx(float(dT[:gref.shape[0], :gref.shape[1]].abs().sum()), rel=1e-12)
    print(f"[sweeps] {name} {dtype}: worst pair-gradient error / max|g| {worst:.2e}")


@pytest.mark.parametrize("name,opt,two_way", [("plain_nll", "sgd", True), ("classifier", "sgd", True), ("classifier", "adam", True),
                                              ("classifier_poly_edge", "sgd", False)])
def test_sweeps_training_matches_oracle(name, opt, two_way):
    """configure(strategy=Sweeps(...)) -> train(): the epoch losses and the trained model against the oracle's restatement."""
    c, arrays, x, t, model, prob = _case(name)
    loss_fn, emb = _loss_and_emb(c)
    lr, epochs = 0.05, 2
    model.configure(strategy=T.Sweeps(two_way=two_way), optimizer=T.optim.sgd if opt == "sgd" else T.optim.adam, loss=loss_fn,
                    train_type=T.TrainingType.SUPERVISED if t is not None else T.TrainingType.UNSUPERVISED,
                    learning_rate=lr, device=("gpu", 0))
    hist = model.train(x.numpy(), targets=None if t is None else t.numpy(), batch_size=c["B"], epochs=epochs, embedding=emb,
                       dtype=np.float64)
    arr_ref, hist_ref, _ = O.sweeps_train(prob, x, arrays, t, epochs=epochs, lr=lr, two_way=two_way, optimizer=opt)
    assert len(hist["loss"]) == epochs
    assert np.allclose(hist["loss"], hist_ref, rtol=1e-8, atol=1e-10), (hist["loss"], hist_ref)
    assert hist["loss"][-1] < hist["loss"][0]
    mine = O.per_sample_loss(prob, x, [a.detach().cpu().double() for a in model.arrays], t)
    ref = O.per_sample_loss(prob, x, arr_ref, t)
    assert torch.allclose(mine, ref, rtol=1e-7, atol=1e-9)
    # shapes of the reference-shaped views are unchanged by the splits (max_bond = the bond's size)
    assert [tuple(a.shape) for a in model.arrays] == [tuple(a.shape) for a in arrays]


def test_sweeps_strategy_strings_and_errors():
    """configure(strategy=...) strings of the reference (model.py:186-214) and the grouping errors (strategy.py:100-103)."""
    c, arrays, x, t, model, prob = _case("plain_nll")
    model.configure(strategy="sweeps", loss=T.metrics.NegLogLikelihood, device=("gpu", 0))
    assert isinstance(model.strategy, T.Sweeps) and model.strategy.two_way and model.strategy.grouping == 2
    model.configure(strategy="sweeps-one-way", loss=T.metrics.NegLogLikelihood, device=("gpu", 0))
    assert not model.strategy.two_way
    assert list(model.strategy.iterate_sites(model)) == [(i, i + 1) for i in range(5)]
    assert list(T.Sweeps().iterate_sites(model)) == [(i, i + 1) for i in range(5)] + [(i, i - 1) for i in range(5, 0, -1)]
    with pytest.raises(ValueError):
        model.configure(strategy="sweeps-per-site")
    with pytest.raises(ValueError):
        T.Sweeps(grouping=3)
    # gradient clipping and renormalisation run
    model.configure(strategy=T.Sweeps(renormalize=True), optimizer=T.optim.adam, loss=T.metrics.NegLogLikelihood,
                    learning_rate=1e-2, device=("gpu", 0))
    hist = model.train(x.numpy(), batch_size=10, epochs=1, gradient_clip_threshold=0.5, dtype=np.float64)
    assert np.isfinite(hist["loss"]).all()
