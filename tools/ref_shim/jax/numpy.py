"""jax.numpy -> numpy (float64)."""
from numpy import *  # noqa: F401,F403
import numpy as _np
from numpy import linalg, fft  # noqa: F401

float_ = _np.float64
ndarray = _np.ndarray


def array(x, dtype=None, **_kw):
    return _np.asarray(x, dtype=dtype)


asarray = array
