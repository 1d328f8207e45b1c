import numpy as _np
from . import initializers  # noqa: F401


def softmax(x, axis=-1):
    e = _np.exp(x - _np.max(x, axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def log_softmax(x, axis=-1):
    m = x - _np.max(x, axis=axis, keepdims=True)
    return m - _np.log(_np.exp(m).sum(axis=axis, keepdims=True))


def relu(x):
    return _np.maximum(x, 0)
