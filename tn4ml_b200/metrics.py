"""Loss heads, mirroring ``tn4ml.metrics`` (same names and call signatures).

In the reference these functions are traced by JAX; here they are resolved *symbolically*:
``Model.train`` calls the user's loss once with tracer arguments and receives a
:class:`LossExpr` naming the fused CUDA head (``csrc/common.cuh``) plus any model-only
regulariser.  User closures written for the reference keep working, e.g.::

    def loss(*a, **kw): return OptaxWrapper(softmax_cross_entropy)(*a, **kw).mean()
    def loss(model, data, y=None): return CombinedLoss(model, data, y, error=LogQuadNorm,
                                                       reg=lambda P: 0.4 * LogReLUFrobNorm(P))
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import numpy as np

from . import _lib


class ModelTracer:
    """Stands for the model argument while a loss is being resolved."""

    def __init__(self, model):
        self.model = model
        self.tensors = model.tensors

    def __len__(self):
        return self.model.L


class DataTracer:
    """Stands for the embedded sample (``embed(x, phi)``)."""

    def __init__(self, n_sites, dim):
        self.n_sites, self.dim = n_sites, dim


class TargetTracer:
    """Stands for one target vector."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


@dataclass
class RegExpr:
    """coef * regulariser(model)  -- batch independent (metrics.py:89-168)."""

    kind: str
    coef: float = 1.0

    def __mul__(self, s):
        return RegExpr(self.kind, self.coef * float(s))

    __rmul__ = __mul__


@dataclass
class LossExpr:
    head: int
    class_weights: tuple | None = None
    regs: list = field(default_factory=list)
    semisup: float | None = None  # SemiSupervisedLoss: coefficient of LogReLUFrobNorm inside the per-sample value

    def mean(self, *a, **k):  # the per-sample loss is a scalar: .mean() is the identity
        return self

    def __add__(self, other):
        if isinstance(other, RegExpr):
            return LossExpr(self.head, self.class_weights, self.regs + [other], self.semisup)
        if isinstance(other, (int, float)) and other == 0:
            return self
        return NotImplemented

    __radd__ = __add__


def _check_lengths(model, data):
    if isinstance(model, ModelTracer) and isinstance(data, DataTracer):
        if len(model) > data.n_sites:
            raise ValueError(  # metrics.py:49-51
                "Number of tensors for input data MPS needs to be higher or equal number of tensors in model.")
        # len(model) < len(data): the sites beyond the model are summed over their physical index (metrics.py:36-43);
        # handled by Engine._split_extra


def NegLogLikelihood(model, data):  # noqa: N802
    """-log(<A|phi>^2)  (metrics.py:17-53)."""
    if isinstance(model, ModelTracer) and isinstance(data, DataTracer):
        assert model.model.phys_dim == data.dim  # metrics.py:34
    _check_lengths(model, data)
    return LossExpr(_lib.HEAD_NLL)


def TransformedSquaredNorm(model, data):  # noqa: N802
    """(<v|v>)^2 with v = model.apply(data)  (metrics.py:56-81)."""
    _check_lengths(model, data)
    return LossExpr(_lib.HEAD_TSN)


def LogQuadNorm(model, data):  # noqa: N802
    """(log TSN - 1)^2  (metrics.py:171-187)."""
    _check_lengths(model, data)
    return LossExpr(_lib.HEAD_LOGQUAD)


def QuadNorm(model, data):  # noqa: N802
    """(TSN - 1)^2  (metrics.py:190-206)."""
    _check_lengths(model, data)
    return LossExpr(_lib.HEAD_QUAD)


def SemiSupervisedLoss(model, data, y_true, **_kwargs):  # noqa: N802
    """(y / m + (1 - y) m)^2 with m = LogQuadNorm(model, data) + 0.3 LogReLUFrobNorm(model), y the target class fraction
    (metrics.py:209-233).  The per-sample LogQuadNorm and its gradient come from the SMPO kernels (the VJP takes the
    per-sample factor dLoss/dm as a weight); the regulariser from the transfer-matrix kernels."""
    _check_lengths(model, data)
    return LossExpr(_lib.HEAD_LOGQUAD, semisup=0.3)


def CrossEntropySoftmax(model, data, targets):  # noqa: N802
    """-log softmax(y/||y||)[argmax t]  (metrics.py:291-333)."""
    _check_lengths(model, data)
    return LossExpr(_lib.HEAD_XENT_SOFTMAX_ARGMAX)


def MeanSquaredError(model, data, targets):  # noqa: N802
    """mean((y/||y|| - t)^2)  (metrics.py:336-396)."""
    _check_lengths(model, data)
    return LossExpr(_lib.HEAD_MSE)


def softmax_cross_entropy(logits=None, labels=None):
    """Stand-in for ``optax.softmax_cross_entropy`` (optax is not a dependency here).
    Only its identity matters: ``OptaxWrapper`` maps it to the fused head."""
    raise RuntimeError("softmax_cross_entropy is a symbolic marker; pass it to OptaxWrapper")


def OptaxWrapper(optax_loss=None) -> Callable:  # noqa: N802
    """metrics.py:399-487.  Recognises ``softmax_cross_entropy`` (ours or optax's, by name)."""
    assert optax_loss is not None
    name = getattr(optax_loss, "__name__", "")
    if name != "softmax_cross_entropy":
        raise NotImplementedError(f"no fused head for optax loss '{name}' (softmax_cross_entropy is supported)")

    def loss_optax(model, data, y_true=None, **kwargs):
        _check_lengths(model, data)
        return LossExpr(_lib.HEAD_SOFTMAX_XENT)

    return loss_optax


def CrossEntropyWeighted(class_weights=None) -> Callable:  # noqa: N802
    """softmax-xent(y/||y||, t) * sum_k t_k w_k  (metrics.py:490-532)."""
    cw = tuple(float(v) for v in np.asarray(class_weights, dtype=np.float64).reshape(-1))
    if len(cw) > _lib.TNB_MAX_OUT:
        raise ValueError(f"at most {_lib.TNB_MAX_OUT} classes")

    def cross_entropy(model, data, y_true=None, **_kwargs):
        _check_lengths(model, data)
        return LossExpr(_lib.HEAD_SOFTMAX_XENT_WEIGHTED, class_weights=cw)

    return cross_entropy


def NoReg(_x):  # noqa: N802
    return 0


def LogFrobNorm(model):  # noqa: N802
    """log ||model||  (metrics.py:89-106)."""
    return RegExpr("log_frob")


def LogPowFrobNorm(model):  # noqa: N802
    """log ||model||^2  (metrics.py:109-126)."""
    return RegExpr("log_pow_frob")


def LogReLUFrobNorm(model):  # noqa: N802
    """max(0, log ||model||)  (metrics.py:129-147)."""
    return RegExpr("log_relu_frob")


def QuadFrobNorm(model):  # noqa: N802
    """(log ||model|| - 1)^2  (metrics.py:150-168)."""
    return RegExpr("quad_frob")


def CombinedLoss(model, data, y_true=None, error: Callable = LogQuadNorm, reg: Callable = NoReg,  # noqa: N802
                 embedding=None):
    """mean(error) + reg(model)  (metrics.py:535-592)."""
    if data is None:
        raise ValueError("Provide input data!")
    expr = error(model, data, y_true) if y_true is not None else error(model, data)
    return expr.mean() + reg(model)


def resolve(loss_fn: Callable, model, n_sites: int, dim: int, supervised: bool) -> LossExpr:
    """Call the user's loss once with tracers and return the LossExpr it builds."""
    if isinstance(loss_fn, LossExpr):
        return loss_fn
    m, d = ModelTracer(model), DataTracer(n_sites, dim)
    expr = loss_fn(m, d, TargetTracer(model.n_out)) if supervised else loss_fn(m, d)
    if not isinstance(expr, LossExpr):
        raise TypeError("the loss must be composed of tn4ml_b200.metrics functions (got %r)" % type(expr))
    return expr
