"""Site-tensor initialisers (reference: tn4ml/initializers.py).  Initialisation runs once and
is outside the hot path; these follow the reference's closures ``init(key, shape, dtype)`` with a
``torch.Generator`` (or an int seed) standing in for the JAX PRNG key."""
from __future__ import annotations

import torch


def _gen(key):
    if isinstance(key, torch.Generator):
        return key
    g = torch.Generator()
    g.manual_seed(int(key) if key is not None else 0)
    return g


def randn(std: float = 1.0, mean: float = 0.0):
    def init(key, shape, dtype=torch.float64):
        return mean + std * torch.randn(tuple(shape), generator=_gen(key), dtype=torch.float64).to(dtype)
    init.__qualname__ = "randn.<locals>.init"
    return init


def zeros():
    def init(key, shape, dtype=torch.float64):
        return torch.zeros(tuple(shape), dtype=dtype)
    return init


def ones():
    def init(key, shape, dtype=torch.float64):
        return torch.ones(tuple(shape), dtype=dtype)
    return init
