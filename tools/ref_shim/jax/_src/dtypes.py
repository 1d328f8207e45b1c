import numpy as _np


def canonicalize_dtype(d):
    return _np.dtype(d)


float_ = _np.float64
