import numpy as _np


def do(fn, *args, like=None, **kwargs):
    if fn == "linalg.norm":
        return _np.linalg.norm(*args, **kwargs)
    if fn == "linalg.qr":
        return _np.linalg.qr(*args, **kwargs)
    return getattr(_np, fn)(*args, **kwargs)


def transpose(x, axes=None):
    return _np.transpose(x, axes)


def reshape(x, shape):
    return _np.reshape(x, shape)
