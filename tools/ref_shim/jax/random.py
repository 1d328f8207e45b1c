import numpy as _np


def PRNGKey(seed):
    return _np.random.default_rng(seed)


key = PRNGKey


def split(k, n=2):
    return [_np.random.default_rng(k.integers(1 << 31)) for _ in range(n)]


def normal(k, shape=(), dtype=_np.float64):
    return k.standard_normal(shape).astype(dtype)


def uniform(k, shape=(), dtype=_np.float64, minval=0.0, maxval=1.0):
    return k.uniform(minval, maxval, shape).astype(dtype)


def permutation(k, x):
    return k.permutation(x)
