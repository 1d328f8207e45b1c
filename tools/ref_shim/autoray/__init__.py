import numpy as _np


def do(fn, *args, like=None, **kwargs):
    if fn == "linalg.norm":
        return _np.linalg.norm(*args, **kwargs)
    if fn == "linalg.qr":
        return _np.linalg.qr(*args, **kwargs)
    return getattr(_np, fn)(*args, **kwargs)
