"""Parity of the CUDA path (through the C-ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): relative error <= 1e-10 in float64, <= 1e-4 in float32,
on the loss and on every gradient tensor (max-abs error scaled by max-abs of the oracle gradient).
"""
import numpy as np
import pytest
import torch

from oracle import tn_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-10, torch.float32: 1e-4}
LAST = {}  # worst errors of the most recent _check (read by tools/parity_report.py)


def _heads():
    from tn4ml_b200 import metrics as M
    return {
        "nll": M.NegLogLikelihood,
        "softmax_xent": lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean(),
        "xent_softmax_argmax": M.CrossEntropySoftmax,
        "mse": M.MeanSquaredError,
        "tsn": M.TransformedSquaredNorm,
        "logquad": M.LogQuadNorm,
        "quad": M.QuadNorm,
    }


def _emb(spec):
    import tn4ml_b200 as T
    if spec.kind == "trig":
        return T.TrigonometricEmbedding(k=spec.k)
    if spec.kind == "fourier":
        return T.FourierEmbedding(p=spec.p)
    if spec.kind == "poly":
        return T.PolynomialEmbedding(degree=spec.degree, n=spec.n, include_bias=spec.include_bias,
                                     include_cross_terms=spec.cross)
    if spec.kind == "lincomp":
        return T.LinearComplementEmbedding(p=spec.p)
    if spec.kind in ("legendre", "laguerre", "hermite"):
        return {"legendre": T.LegendreEmbedding, "laguerre": T.LaguerreEmbedding,
                "hermite": T.HermiteEmbedding}[spec.kind](degree=spec.degree)
    if spec.kind == "arrays":
        return T.JaxArraysEmbedding(dim=spec.dim, add_bias=spec.include_bias, input_dim=spec.n)
    return T.GaussianRBFEmbedding(centers=np.array(spec.centers), gamma=spec.gamma)


def _errors(model, prob, loss_fn, x, arrays, t, dtype, precision=0):
    """Loss and gradient errors of the CUDA path (through the C-ABI) against the float64 oracle on identical inputs."""
    model.precision = precision  # 0 = SIMT kernels, 1 = tcgen05 kernels where the shape allows
    model._engine_cache = {}
    # identical inputs on both sides: round to the kernel dtype first, then run the oracle in float64
    arr64 = [a.to(dtype).double() for a in arrays]
    x = x.to(dtype).double()
    loss_ref, grads_ref = O.value_and_grad(prob, x, arr64, None if t is None else t.double())
    per_ref = O.per_sample_loss(prob, x, arr64, None if t is None else t.double())
    eng, _ = model._engine(_emb(prob.emb), loss_fn, t is not None, dtype)
    B = x.shape[0]
    loss_sum, g = eng.value_and_grad(model._packed, x.to(dtype), None if t is None else t.to(dtype))
    loss = float(loss_sum.item()) / B
    gmax = max(float(q.abs().max()) for q in grads_ref)
    gfro = max(float(q.norm()) for q in grads_ref)
    worst_fro, worst_max, worst_site = 0.0, 0.0, -1
    for i, (a, b) in enumerate(zip(model._views(g), grads_ref)):
        assert a.shape == b.shape
        diff = a.double().cpu() - b
        # ||g - g_ref|| / ||g_ref|| per site, and max-elementwise scaled by ||g_ref||_inf
        # (sites whose gradient is < 1e-3 of the largest are measured against that floor)
        efro = float(diff.norm()) / max(float(b.norm()), 1e-3 * gfro)
        emax = float(diff.abs().max()) / gmax
        if efro > worst_fro:
            worst_fro, worst_site = efro, i
        worst_max = max(worst_max, emax)
    per, out, oexp = eng.forward(model._packed, x.to(dtype), None if t is None else t.to(dtype))
    e = {"loss_rel": abs(loss - float(loss_ref)) / max(1.0, abs(float(loss_ref))), "site_fro": worst_fro,
         "site_max": worst_max, "site": worst_site, "loss": loss, "B": B, "L": len(arrays),
         "per": per.double().cpu(), "per_ref": per_ref}
    # the margin is part of the record: pytest -s / -rP shows it
    print(f"[parity] {dtype} precision={precision} B={B} L={len(arrays)}: loss rel err {e['loss_rel']:.2e}, worst per-site rel "
          f"Frobenius err {worst_fro:.2e} (site {worst_site}), worst max-abs err / max|g| {worst_max:.2e}")
    LAST.clear()
    LAST.update({k: v for k, v in e.items() if k not in ("per", "per_ref")})
    return e


def _check(model, prob, loss_fn, x, arrays, t, dtype, scale_tol=1.0, precision=0):
    tol = TOL[dtype] * scale_tol
    e = _errors(model, prob, loss_fn, x, arrays, t, dtype, precision)
    assert e["loss_rel"] <= tol, f"loss: {e['loss_rel']:.3e} > {tol:.0e}"
    assert e["site_fro"] <= tol, f"site {e['site']} (fro): {e['site_fro']:.3e} > {tol:.0e}"
    assert e["site_max"] <= tol, f"(max): {e['site_max']:.3e} > {tol:.0e}"
    assert torch.allclose(e["per"], e["per_ref"], rtol=tol * 10, atol=tol * 10)
    return e["loss"]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", [
    dict(L=12, chi=6, d=2, cls=5, C=3, B=40, emb=O.EmbSpec("trig"), head="softmax_xent"),
    dict(L=7, chi=4, d=3, cls=0, C=4, B=9, emb=O.EmbSpec("poly", degree=2, include_bias=True), head="mse"),
    dict(L=7, chi=5, d=4, cls=6, C=2, B=33, emb=O.EmbSpec("trig", k=2), head="xent_softmax_argmax"),
    dict(L=30, chi=32, d=3, cls=14, C=2, B=32, emb=O.EmbSpec("poly", degree=2, include_bias=True),
         head="softmax_xent_weighted", cw=(0.7, 1.6)),                      # cfg 2
    dict(L=20, chi=12, d=3, cls=9, C=5, B=130, emb=O.EmbSpec("rbf", gamma=0.8, centers=(0.0, 0.5, 1.0)),
         head="softmax_xent", shape_method="noteven"),
    dict(L=10, chi=40, d=2, cls=4, C=3, B=70, emb=O.EmbSpec("trig"), head="softmax_xent"),
], ids=["small", "cls_first", "cls_last", "cfg2", "noteven_rbf", "chi40"])
def test_mps_classifier_parity(case, dtype):
    import tn4ml_b200 as T
    from tn4ml_b200 import metrics as M
    arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], class_site=case["cls"], C=case["C"], std=0.2,
                               seed=3, shape_method=case.get("shape_method", "even"))
    x = O.make_inputs(case["B"], case["L"], seed=5)
    t = O.make_labels(case["B"], case["C"])
    cw = case.get("cw")
    prob = O.Problem("mps", case["emb"], case["head"], class_weights=cw)
    loss_fn = M.CrossEntropyWeighted(cw) if cw else _heads()[case["head"]]
    model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
    _check(model, prob, loss_fn, x, arrays, t, dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_cfg1_mnist_shaped_classifier(dtype):
    """BASELINE config 1: L=196, trig d=2, chi=10, 10 classes at site 97, batch 64."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(196, 10, 2, class_site=97, C=10, std=1e-2, seed=42)
    x = O.make_inputs(64, 196, seed=1235)
    t = O.make_labels(64, 10)
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
    _check(model, prob, _heads()["softmax_xent"], x, arrays, t, dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", [
    dict(L=16, chi=16, d=8, B=100, emb=O.EmbSpec("fourier", p=8)),         # cfg 4 shape, reduced chi
    dict(L=16, chi=64, d=8, B=257, emb=O.EmbSpec("fourier", p=8)),         # cfg 4 at full chi
    dict(L=9, chi=7, d=2, B=5, emb=O.EmbSpec("trig"), shape_method="noteven"),
], ids=["fourier16", "cfg4", "noteven"])
def test_mps_nll_parity(case, dtype):
    import tn4ml_b200 as T
    # identity on every physical component keeps <A|phi> away from 0 (where -log s^2 is singular
    # and its gradient arbitrarily ill-conditioned)
    arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], std=0.05, seed=4, squeeze_ends=True,
                               shape_method=case.get("shape_method", "even"), identity_all=True)
    x = O.make_inputs(case["B"], case["L"], seed=6)
    prob = O.Problem("mps", case["emb"], "nll")
    model = T.MatrixProductState(arrays, dtype=dtype, device="cuda:0")
    _check(model, prob, _heads()["nll"], x, arrays, None, dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", [
    dict(L=16, S=4, chi=8, d=2, do=2, B=10, head="logquad"),
    dict(L=17, S=4, chi=6, d=2, do=3, B=7, head="tsn"),        # output leg on the last site
    dict(L=12, S=1, chi=5, d=2, do=2, B=9, head="quad"),       # MPO
    dict(L=64, S=8, chi=32, d=2, do=2, B=11, head="logquad"),  # cfg 3 shape, shorter chain
    dict(L=10, S=5, chi=20, d=3, do=2, B=6, head="logquad"),
], ids=["s4", "s4_lastout", "mpo", "cfg3_short", "chi20"])
def test_smpo_parity(case, dtype):
    import tn4ml_b200 as T
    arrays = O.make_smpo_arrays(case["L"], case["chi"], case["d"], case["do"], case["S"], std=0.05, seed=8)
    x = O.make_inputs(case["B"], case["L"], seed=9)
    prob = O.Problem("smpo", O.EmbSpec("trig") if case["d"] == 2 else O.EmbSpec("poly", degree=3), case["head"],
                     spacing=case["S"])
    cls = T.MatrixProductOperator if case["S"] == 1 else T.SpacedMatrixProductOperator
    kw = {} if case["S"] == 1 else {"spacing": case["S"]}
    model = cls(arrays, dtype=dtype, device="cuda:0", **kw)
    _check(model, prob, _heads()[case["head"]], x, arrays, None, dtype)


def test_long_chain_does_not_leave_fp32_range():
    """H2: a 784-site chain whose raw product (~0.05^784) underflows float32 AND float64 is
    carried by the per-sample exponent; checked against the oracle evaluated in log space."""
    import tn4ml_b200 as T
    L = 784
    base = O.make_mps_arrays(L, 8, 2, std=0.05, seed=2, squeeze_ends=True, identity_all=True)
    arrays = [(0.05 * a).float() for a in base]
    x = O.make_inputs(16, L, seed=3).float()
    per_ref = O.mps_nll_logspace(O.embed_batch(x.double(), O.EmbSpec("trig")), [a.double() for a in arrays])
    assert not torch.isfinite(O.per_sample_loss(O.Problem("mps", O.EmbSpec("trig"), "nll"), x.double(),
                                                [a.double() for a in arrays])).any()
    model = T.MatrixProductState(arrays, dtype=torch.float32, device="cuda:0")
    eng, _ = model._engine(T.TrigonometricEmbedding(), _heads()["nll"], False, torch.float32)
    per, out, oexp = eng.forward(model._packed, x)
    assert torch.isfinite(per).all()
    assert torch.allclose(per.double().cpu(), per_ref, rtol=1e-5)


def test_embedding_call_matches_oracle():
    import tn4ml_b200 as T
    x = torch.linspace(0.01, 1, 57, dtype=torch.float64)
    for spec in [O.EmbSpec("trig", k=2), O.EmbSpec("fourier", p=8), O.EmbSpec("poly", degree=3),
                 O.EmbSpec("rbf", gamma=0.5, centers=(0.1, 0.9))]:
        phi = _emb(spec)(x.numpy())
        assert torch.allclose(phi.cpu(), O.feature_map(x, spec), atol=1e-12)


MORE_MAPS = [O.EmbSpec("lincomp", p=2), O.EmbSpec("lincomp", p=3), O.EmbSpec("legendre", degree=3), O.EmbSpec("laguerre", degree=2),
             O.EmbSpec("hermite", degree=4), O.EmbSpec("poly", degree=2, n=3, include_bias=True),
             O.EmbSpec("poly", degree=2, n=3, include_bias=True, cross=True), O.EmbSpec("poly", degree=3, n=2, cross=True),
             O.EmbSpec("arrays", n=3), O.EmbSpec("arrays", n=3, include_bias=True)]
MORE_IDS = ["lincomp2", "lincomp3", "legendre3", "laguerre2", "hermite4", "poly_n3", "poly_n3_cross", "poly_n2_d3_cross", "arrays3",
            "arrays3_bias"]


def _inputs_for(spec, B, L, seed):
    """x [B, L] (scalar maps) or [B, L, n] (maps with input_dim n > 1), in U(0, 1]."""
    rng = np.random.default_rng(seed)
    shape = (B, L) if spec.n == 1 else (B, L, spec.n)
    return torch.tensor(1.0 - rng.random(shape))


@pytest.mark.parametrize("spec", MORE_MAPS, ids=MORE_IDS)
def test_more_embeddings_call_matches_oracle(spec):
    """Embedding.__call__ of the remaining feature maps (embeddings.py:426-484, 605-717 with n > 1, 720-869, 872-928)."""
    x = _inputs_for(spec, 57, 1, 3)[:, 0]
    if spec.kind in ("legendre", "hermite"):
        x = 2.0 * x - 1.0
    phi = _emb(spec)(x.numpy())
    ref = O.feature_map(x, spec)
    assert phi.shape == ref.shape and phi.shape[-1] == spec.dim == _emb(spec).dim
    assert torch.allclose(phi.cpu(), ref, atol=1e-12)
    phi32 = _emb(spec)(x.float())
    assert torch.allclose(phi32.double().cpu(), ref, atol=2e-5, rtol=2e-5)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("spec", MORE_MAPS, ids=MORE_IDS)
def test_more_embeddings_mps_parity(spec, dtype):
    """value_and_grad of a plain MPS (NegLogLikelihood) and of a classifier over every remaining feature map, fused in-kernel
    like the four headline maps (the normaliser of embed() matters for all but lincomp)."""
    import tn4ml_b200 as T
    L, chi, B, C = 9, 12, 37, 3
    d = spec.dim
    x = _inputs_for(spec, B, L, 21)
    arrays = O.make_mps_arrays(L, chi, d, std=0.05, seed=4, squeeze_ends=True, identity_all=True)
    _check(T.MatrixProductState(arrays, dtype=dtype, device="cuda:0"), O.Problem("mps", spec, "nll"), _heads()["nll"], x, arrays,
           None, dtype)
    arrays = O.make_mps_arrays(L, chi, d, class_site=4, C=C, std=0.2, seed=3)
    t = O.make_labels(B, C)
    _check(T.TensorNetwork(arrays, dtype=dtype, device="cuda:0"), O.Problem("mps", spec, "softmax_xent"),
           _heads()["softmax_xent"], x, arrays, t, dtype)


@pytest.mark.parametrize("spec", [O.EmbSpec("lincomp", p=2), O.EmbSpec("legendre", degree=1), O.EmbSpec("arrays", n=2)],
                         ids=["lincomp2", "legendre1", "arrays2"])
def test_more_embeddings_tensor_and_smpo_families(spec):
    """The same maps through the tcgen05 chain kernels (float32, chi = 64) and the SMPO kernels (register-resident family at
    chi = 32, d = 2)."""
    import tn4ml_b200 as T
    L, chi, B, C = 10, 64, 200, 4
    x = _inputs_for(spec, B, L, 5)
    arrays = [a.float() for a in O.make_mps_arrays(L, chi, 2, class_site=4, C=C, std=0.05, seed=3)]
    t = O.make_labels(B, C)
    model = T.TensorNetwork(arrays, dtype=torch.float32, device="cuda:0")
    _check(model, O.Problem("mps", spec, "softmax_xent"), _heads()["softmax_xent"], x, arrays, t, torch.float32, precision=1)
    arrays = O.make_smpo_arrays(24, 32, 2, 2, 8, std=0.05, seed=8)
    xs = _inputs_for(spec, 11, 24, 9)
    for dtype in (torch.float64, torch.float32):
        model = T.SpacedMatrixProductOperator(arrays, spacing=8, dtype=dtype, device="cuda:0")
        _check(model, O.Problem("smpo", spec, "logquad", spacing=8), _heads()["logquad"], xs, arrays, None, dtype)


@pytest.mark.parametrize("case", [
    dict(L=5, chi=32, cls=2, C=3, B=19000),    # TN = 8: four row groups of warps
    dict(L=5, chi=40, cls=1, C=2, B=9600),     # padded positions (WS = 64 > chi)
    dict(L=4, chi=64, cls=0, C=3, B=9500),     # class site on the boundary
    dict(L=4, chi=128, cls=2, C=2, B=4800),    # TN = 32, one column group
    dict(L=4, chi=136, cls=1, C=2, B=4800),    # two column groups (chi > 128)
], ids=["chi32", "chi40", "chi64", "chi128", "chi136"])
def test_mps_classifier_parity_f64_tensor_pipe(case):
    """float64, d = 2, batches large enough for full 256-thread CTAs: the step GEMM of the sweep / class kernels and the
    128 x 64 dA tile run as DMMA (mma.sync.m8n8k4.f64) warp tiles -- same 1e-10 bar as the FMA loops."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(case["L"], case["chi"], 2, class_site=case["cls"], C=case["C"], std=0.1, seed=11)
    x = O.make_inputs(case["B"], case["L"], seed=12)
    t = O.make_labels(case["B"], case["C"])
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    model = T.TensorNetwork(arrays, dtype=torch.float64, device="cuda:0")
    _check(model, prob, _heads()["softmax_xent"], x, arrays, t, torch.float64)


@pytest.mark.parametrize("case", [
    dict(L=6, chi=32, d=3, cls=2, C=2, B=9700, emb=O.EmbSpec("poly", degree=2, include_bias=True)),   # cfg 2's map: odd d
    dict(L=5, chi=64, d=8, cls=None, C=1, B=9600, emb=O.EmbSpec("fourier", p=8)),                     # cfg 4's map
    dict(L=5, chi=136, d=4, cls=1, C=3, B=4900, emb=O.EmbSpec("trig", k=2)),                          # two column groups
], ids=["d3_chi32", "d8_chi64", "d4_chi136"])
def test_f64_tensor_pipe_general_physical_dimension(case):
    """float64 step GEMM on DMMA for d != 2: one pass per pair of physical indices (an odd d ends with a single one), full
    256-thread CTAs -- same 1e-10 bar."""
    import tn4ml_b200 as T
    if case["cls"] is None:
        arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], std=0.05, seed=4, squeeze_ends=True, identity_all=True)
        t, head = None, "nll"
        model = T.MatrixProductState(arrays, dtype=torch.float64, device="cuda:0")
    else:
        arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], class_site=case["cls"], C=case["C"], std=0.1, seed=11)
        t, head = O.make_labels(case["B"], case["C"]), "softmax_xent"
        model = T.TensorNetwork(arrays, dtype=torch.float64, device="cuda:0")
    x = O.make_inputs(case["B"], case["L"], seed=12)
    _check(model, O.Problem("mps", case["emb"], head), _heads()[head], x, arrays, t, torch.float64)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_edge_shapes(dtype):
    """Smallest and ragged shapes: one sample, two sites, bond 1, a batch one short / one past a tile, SMPO with one window."""
    import tn4ml_b200 as T
    tol_scale = 1.0
    # one sample, classifier on both kernel families' smallest legal bonds
    for chi, B, L, cls in ((3, 1, 5, 2), (32, 1, 4, 1), (32, 129, 3, 0), (32, 127, 3, 2), (1, 4, 6, 3)):
        if chi == 1 and dtype == torch.float32:
            continue  # (a product-state model: some site gradients cancel to 1e-3 of the largest; float32 keeps 2e-6 of max|g| there)
        arrays = O.make_mps_arrays(L, chi, 2, class_site=cls, C=3, std=0.2, seed=3)
        x, t = O.make_inputs(B, L, seed=5), O.make_labels(B, 3)
        model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
        _check(model, O.Problem("mps", O.EmbSpec("trig"), "softmax_xent"), _heads()["softmax_xent"], x, arrays, t, dtype, tol_scale,
               precision=1 if (dtype == torch.float32 and chi == 32) else 0)
    # two sites: plain MPS with squeezed boundaries
    arrays = O.make_mps_arrays(2, 4, 3, std=0.3, seed=4, squeeze_ends=True, identity_all=True)
    x = O.make_inputs(7, 2, seed=6)
    _check(T.MatrixProductState(arrays, dtype=dtype, device="cuda:0"), O.Problem("mps", O.EmbSpec("poly", degree=2, include_bias=True), "nll"),
           _heads()["nll"], x, arrays, None, dtype)
    # SMPO: a single window (spacing == L), one sample; and the register-resident family with one sample
    for L, S, chi, B in ((4, 4, 5, 1), (8, 8, 32, 1), (9, 8, 32, 3)):
        arrays = O.make_smpo_arrays(L, chi, 2, 2, S, std=0.2 if chi < 32 else 0.05, seed=8)
        x = O.make_inputs(B, L, seed=9)
        model = T.SpacedMatrixProductOperator(arrays, spacing=S, dtype=dtype, device="cuda:0")
        _check(model, O.Problem("smpo", O.EmbSpec("trig"), "logquad", spacing=S), _heads()["logquad"], x, arrays, None, dtype)
