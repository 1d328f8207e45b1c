#!/usr/bin/env python
"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE (run in this container; /root/reference is not on the
GPU box, the committed .npz files are what travels).

/root/reference/tn4ml is imported unmodified; its absent third-party stack (jax, optax, quimb, autoray, funcy) is replaced by
the stand-ins in tools/ref_shim (numpy semantics; a generic named-index tensor network -- see tools/ref_shim/README.md).
For every case the script stores the inputs and what the reference code returned:

  reference_embeddings.npz   Embedding.__call__ of the four feature maps on the hot path (embeddings.py:327-351, 394-423,
                             588-602, 680-717) and embed()'s normalised product state on both branches (:1374-1387)
  reference_mps_nll.npz      tn4ml.metrics.NegLogLikelihood(model, embed(x)) per sample, model = tn4ml.models.mps.
                             MatrixProductState(arrays) (plain MPS, squeezed boundaries) -- metrics.py:17-53
  reference_classifier.npz   OptaxWrapper(optax.softmax_cross_entropy), CrossEntropySoftmax-style heads, MeanSquaredError,
                             CrossEntropyWeighted on a classifier built by the reference's MPS_initialize(arrays=...,
                             class_index=...) (mps.py:296-330 index naming) -- metrics.py:291-333, 336-396, 423-532
  reference_smpo.npz         tn4ml.models.smpo.SpacedMatrixProductOperator(arrays) (constructor conventions, smpo.py:61-206) with
                             <v|v> by generic contraction and TransformedSquaredNorm / LogQuadNorm / QuadNorm executed through the
                             reference's len(model) < len(data) branch (metrics.py:56-81, 171-206)
  + central finite differences of the reference-executed mean loss with respect to a few site-tensor entries (the shim has no
    autodiff), which pin value_and_grad (model.py:554-566) entry by entry.

tests/test_reference_golden.py checks the oracle (CPU) against these files; the GPU parity tests check the kernels against the
oracle.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/tn4ml"
sys.path.insert(0, os.path.join(ROOT, "tools", "ref_shim"))
sys.path.insert(0, ROOT)
pkg = types.ModuleType("tn4ml")  # the package WITHOUT its __init__ (which imports plotting / sklearn helpers)
pkg.__path__ = [REF]
sys.modules["tn4ml"] = pkg

import optax  # noqa: E402  (stand-in)
import tn4ml.embeddings as RE  # noqa: E402  (reference source)
import tn4ml.metrics as RM  # noqa: E402
import tn4ml.models.mps as RMPS  # noqa: E402

from tn4ml_b200 import synthetic as S  # noqa: E402  (seeded data generators)

OUT = os.path.join(ROOT, "tests", "golden")


def embeddings_file():
    xs = np.linspace(0.013, 1.0, 23)
    out = {"x": xs}
    maps = {
        "trig_k1": RE.TrigonometricEmbedding(k=1), "trig_k2": RE.TrigonometricEmbedding(k=2),
        "fourier_p4": RE.FourierEmbedding(p=4), "fourier_p8": RE.FourierEmbedding(p=8),
        "poly_d2_bias": RE.PolynomialEmbedding(degree=2, n=1, include_bias=True),
        "poly_d3": RE.PolynomialEmbedding(degree=3, n=1, include_bias=False),
        "rbf": RE.GaussianRBFEmbedding(centers=np.array([0.1, 0.5, 0.9]), gamma=0.8),
    }
    for name, phi in maps.items():
        out[f"phi_{name}"] = np.stack([np.asarray(phi(np.asarray(x)), dtype=np.float64).reshape(-1) for x in xs])
        out[f"dim_{name}"] = np.array(phi.dim)
    # embed(): the normalised product state, L <= 200 branch and L > 200 branch
    rng = np.random.default_rng(5)
    for name, L in (("short", 7), ("long", 203)):
        x = 1.0 - rng.random(L)
        for mname in ("poly_d2_bias", "trig_k1", "rbf"):
            mps = RE.embed(x, maps[mname])
            arr = np.stack([np.asarray(t.data, dtype=np.float64).reshape(-1) for t in mps.tensors])
            out[f"embed_{name}_{mname}_x"] = x
            out[f"embed_{name}_{mname}"] = arr
            out[f"embed_{name}_{mname}_norm"] = np.array(float(mps.norm()))
    np.savez_compressed(os.path.join(OUT, "reference_embeddings.npz"), **out)
    return out


def embeddings_more_file():
    """The remaining scalar and vector-input feature maps: embeddings.py:426-484, 605-717 with n > 1,
    720-869 (over tn4ml/scipy/special.py), 872-928; and NegLogLikelihood of a plain MPS over two of them."""
    xs = np.linspace(0.013, 1.0, 23)
    out = {"x": xs}
    maps = {
        "lincomp_p2": RE.LinearComplementEmbedding(p=2), "lincomp_p3": RE.LinearComplementEmbedding(p=3),
        "legendre_d3": RE.LegendreEmbedding(degree=3), "laguerre_d3": RE.LaguerreEmbedding(degree=3),
        "hermite_d4": RE.HermiteEmbedding(degree=4), "legendre_d0": RE.LegendreEmbedding(degree=0),
        "hermite_d1": RE.HermiteEmbedding(degree=1),
    }
    for name, phi in maps.items():
        xin = 2.0 * xs - 1.0 if name.startswith("legendre") else (3.0 * xs if name.startswith("laguerre") else 2.5 * xs - 1.0
                                                                   if name.startswith("hermite") else xs)
        out[f"x_{name}"] = xin
        out[f"phi_{name}"] = np.stack([np.asarray(phi(np.asarray(x)), dtype=np.float64).reshape(-1) for x in xin])
        out[f"dim_{name}"] = np.array(phi.dim)
    rng = np.random.default_rng(9)
    xv = 1.0 - rng.random((11, 3))
    vmaps = {
        "poly_n3_d2": RE.PolynomialEmbedding(degree=2, n=3), "poly_n3_d2_bias": RE.PolynomialEmbedding(degree=2, n=3, include_bias=True),
        "poly_n3_d2_cross": RE.PolynomialEmbedding(degree=2, n=3, include_bias=True, include_cross_terms=True),
        "poly_n2_d3_cross": RE.PolynomialEmbedding(degree=3, n=2, include_cross_terms=True),
        "arrays_n3": RE.JaxArraysEmbedding(dim=3, input_dim=3), "arrays_n3_bias": RE.JaxArraysEmbedding(dim=4, add_bias=True, input_dim=3),
    }
    out["xv"] = xv
    for name, phi in vmaps.items():
        n = phi.input_dim
        out[f"phi_{name}"] = np.stack([np.asarray(phi(np.asarray(x[:n])), dtype=np.float64).reshape(-1) for x in xv])
        out[f"dim_{name}"] = np.array(phi.dim)
    # NegLogLikelihood of a plain MPS over embed(x, phi) for a scalar map and a vector-input map
    L, chi, B = 6, 4, 5
    for name, phi, xin in (("hermite_d2", RE.HermiteEmbedding(degree=2), S.make_inputs(B, L, seed=31).numpy()),
                           ("arrays_n3", vmaps["arrays_n3"], 1.0 - np.random.default_rng(32).random((B, L, 3)))):
        arrays = [a.numpy() for a in S.make_mps_arrays(L, chi, 3, std=0.2, seed=33, squeeze_ends=True, identity_all=True)]
        model = RMPS.MatrixProductState([np.array(a) for a in arrays])
        out[f"nll_{name}_x"] = xin
        out[f"nll_{name}"] = np.array([float(RM.NegLogLikelihood(model, RE.embed(xb, phi))) for xb in xin])
        for i, a in enumerate(arrays):
            out[f"nll_{name}_site{i}"] = a
    np.savez_compressed(os.path.join(OUT, "reference_embeddings_more.npz"), **out)
    return out


def mps_nll_file():
    L, chi, d, B = 8, 5, 3, 6
    arrays = [a.numpy() for a in S.make_mps_arrays(L, chi, d, std=0.2, seed=11, squeeze_ends=True, identity_all=True)]
    x = S.make_inputs(B, L, seed=12).numpy()
    phi = RE.PolynomialEmbedding(degree=2, n=1, include_bias=True)

    def losses(arrs):
        model = RMPS.MatrixProductState([np.array(a) for a in arrs])
        return np.array([float(RM.NegLogLikelihood(model, RE.embed(xb, phi))) for xb in x])

    per = losses(arrays)
    fd = _finite_differences(lambda arrs: losses(arrs).mean(), arrays, seed=3)
    # len(model) < len(data) (metrics.py:36-43): the same model on data with three more features
    x_long = S.make_inputs(B, L + 3, seed=13).numpy()
    model = RMPS.MatrixProductState([np.array(a) for a in arrays])
    per_long = np.array([float(RM.NegLogLikelihood(model, RE.embed(xb, phi))) for xb in x_long])
    np.savez_compressed(os.path.join(OUT, "reference_mps_nll.npz"), x=x, per_sample_loss=per, L=L, chi=chi, d=d,
                        fd_index=fd[0], fd_grad=fd[1], x_long=x_long, per_sample_loss_long=per_long,
                        **{f"site{i}": a for i, a in enumerate(arrays)})
    return per


def _finite_differences(f, arrays, seed, n=10, h=1e-5):
    """Central differences of f with respect to n random entries (site, flat index) of the site tensors."""
    rng = np.random.default_rng(seed)
    idx, grads = [], []
    for _ in range(n):
        s = int(rng.integers(len(arrays)))
        k = int(rng.integers(arrays[s].size))
        base = [a.copy() for a in arrays]
        e = np.zeros(arrays[s].size)
        e[k] = h
        base[s] = arrays[s] + e.reshape(arrays[s].shape)
        fp = f(base)
        base[s] = arrays[s] - e.reshape(arrays[s].shape)
        fm = f(base)
        idx.append((s, k))
        grads.append((fp - fm) / (2 * h))
    return np.array(idx), np.array(grads)


def smpo_file():
    """SMPO double layer (BASELINE configs[2]'s model): the reference's SpacedMatrixProductOperator constructor (axis orders
    rud / lru / lrud / lu(d), spacing detection, index names -- smpo.py:61-206) and its metrics executed from source.

      n_dense       <v|v> with v = (model & embed(x)) ^ all: the network TransformedSquaredNorm is defined on, contracted by
                    the stand-in's generic contraction (no hot-path knowledge)
      tsn / lqn / qn_long   TransformedSquaredNorm / LogQuadNorm / QuadNorm of the reference (metrics.py:56-81, 171-206) on data
                    with ONE more feature than the model has sites: that takes the reference's `len(model) < len(data)` branch,
                    which contracts the same network through `model.H & data` + `contract_ind` and leaves the extra site summed
                    over its physical index (a factor (sum_s phi_s(x_L))^2 on <v|v>).  The equal-length branch goes through
                    apply_mps, whose array bookkeeping depends on quimb's internal index ordering and is NOT reproduced by the
                    stand-in (tools/ref_shim/README.md).
      fd_*          central finite differences of the reference-executed mean LogQuadNorm."""
    import tn4ml.models.smpo as RSMPO
    phi = RE.TrigonometricEmbedding(k=1)
    out = {}
    for name, (L, Sp, chi, do) in {"s2": (6, 2, 3, 2), "s4_lastout": (9, 4, 4, 3)}.items():
        arrays = [a.numpy() for a in S.make_smpo_arrays(L, chi, 2, do, Sp, std=0.3, seed=8)]
        x = 1.0 - np.random.default_rng(17).random((4, L + 1))

        def build(arrs):
            return RSMPO.SpacedMatrixProductOperator([np.array(a) for a in arrs])

        model = build(arrays)
        assert model.spacing == Sp
        n_dense = []
        for xb in x:
            v = (model & RE.embed(xb[:L], phi)) ^ all
            n_dense.append(float((np.asarray(v.data) ** 2).sum()))

        def metric(fn, arrs):
            m = build(arrs)
            return np.array([float(fn(m, RE.embed(xb, phi))) for xb in x])

        out[f"{name}_L"], out[f"{name}_spacing"], out[f"{name}_chi"], out[f"{name}_do"] = L, Sp, chi, do
        out[f"{name}_x"] = x
        out[f"{name}_n_dense"] = np.array(n_dense)
        out[f"{name}_tsn_long"] = metric(RM.TransformedSquaredNorm, arrays)
        out[f"{name}_lqn_long"] = metric(RM.LogQuadNorm, arrays)
        out[f"{name}_qn_long"] = metric(RM.QuadNorm, arrays)
        fd = _finite_differences(lambda arrs: metric(RM.LogQuadNorm, arrs).mean(), arrays, seed=9, n=8)
        out[f"{name}_fd_index"], out[f"{name}_fd_grad"] = fd
        out.update({f"{name}_site{i}": a for i, a in enumerate(arrays)})
    np.savez_compressed(os.path.join(OUT, "reference_smpo.npz"), **out)
    return out


def classifier_file():
    L, chi, d, C, B, cls = 7, 4, 2, 3, 5, 3  # class tensor at 0-based site 3 == class_index 4 (1-based, mps.py:183-213)
    arrays = [a.numpy() for a in S.make_mps_arrays(L, chi, d, class_site=cls, C=C, std=0.2, seed=21)]
    x = S.make_inputs(B, L, seed=22).numpy()
    y = S.make_labels(B, C, seed=23).numpy()
    phi = RE.TrigonometricEmbedding(k=1)
    cw = np.array([0.7, 1.6, 1.1])

    def build(arrs):
        return RMPS.MPS_initialize(L, arrays=[np.array(a) for a in arrs], class_index=cls + 1)

    def run(loss, arrs, with_targets=True):
        model = build(arrs)
        out = []
        for xb, yb in zip(x, y):
            v = loss(model, RE.embed(xb, phi), yb) if with_targets else loss(model, RE.embed(xb, phi))
            out.append(float(np.asarray(v).reshape(-1)[0]))
        return np.array(out)

    heads = {
        "softmax_xent": RM.OptaxWrapper(optax.softmax_cross_entropy),
        "mse": RM.MeanSquaredError,
        "softmax_xent_weighted": RM.CrossEntropyWeighted(cw),
    }
    out = {"x": x, "y": y, "class_weights": cw, "L": L, "chi": chi, "d": d, "C": C, "class_site": cls}
    for name, loss in heads.items():
        out[f"loss_{name}"] = run(loss, arrays)
        fd = _finite_differences(lambda arrs, loss=loss: run(loss, arrs).mean(), arrays, seed=5)
        out[f"fd_index_{name}"], out[f"fd_grad_{name}"] = fd
    # raw outputs y[c] of the contraction (what Model.forward returns, model.py:312-316)
    model = build(arrays)
    ys = []
    for xb in x:
        t = (model & RE.embed(xb, phi)) ^ all
        ys.append(np.asarray(t.data).reshape(-1))
    out["outputs"] = np.stack(ys)
    out.update({f"site{i}": a for i, a in enumerate(arrays)})
    np.savez_compressed(os.path.join(OUT, "reference_classifier.npz"), **out)
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    e = embeddings_file()
    print("embeddings:", sorted(k for k in e if k.startswith("phi_")))
    m = embeddings_more_file()
    print("more embeddings:", sorted(k for k in m if k.startswith("phi_")), m["nll_hermite_d2"], m["nll_arrays_n3"])
    print("nll per sample:", mps_nll_file())
    sm = smpo_file()
    print("smpo:", {k: v for k, v in sm.items() if k.endswith("_n_dense") or k.endswith("_lqn_long")})
    c = classifier_file()
    print("classifier:", {k: v for k, v in c.items() if k.startswith("loss_")})
