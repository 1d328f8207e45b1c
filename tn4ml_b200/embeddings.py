"""Feature maps on the hot path, mirroring ``tn4ml.embeddings`` (names, constructor
arguments, ``dim`` / ``input_dim`` and error behaviour).

On this path an embedding is a *descriptor*: the feature map is evaluated inside the
first contraction stage of the CUDA kernels (``csrc/common.cuh``), so the embedded
``[B, L, d]`` tensor never reaches HBM.  ``__call__`` is kept for API fidelity and runs
the same device function through ``tnb_embed``.
"""
from __future__ import annotations

import abc
import ctypes as C
from typing import Any

import numpy as np
import torch

from . import _lib


class Embedding(abc.ABC):
    """Base class (reference: tn4ml/embeddings.py:16-73)."""

    def __init__(self, dtype: Any = np.float32):
        self.dtype = dtype

    @property
    @abc.abstractmethod
    def dim(self) -> int: ...

    @property
    def input_dim(self) -> int:
        return 1

    @abc.abstractmethod
    def _spec(self) -> _lib.TnbEmbedding: ...

    def __call__(self, x: Any) -> torch.Tensor:
        """phi(x) for a scalar or an array of scalars -> ``[..., dim]`` (CUDA tensor)."""
        if not torch.cuda.is_available():
            raise RuntimeError("tn4ml_b200 embeddings are evaluated on the GPU; no CUDA device is available")
        xt = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
        if xt.dtype not in (torch.float32, torch.float64):
            xt = xt.to(torch.float64 if np.dtype(self.dtype) == np.float64 else torch.float32)
        xt = xt.cuda().contiguous()
        n_in = self.input_dim or 1
        if n_in > 1:  # one input is a vector of input_dim values (the last axis)
            if xt.ndim < 1 or xt.shape[-1] != n_in:
                raise ValueError(f"expected inputs with a last axis of {n_in} values, got shape {tuple(xt.shape)}")
            lead = xt.shape[:-1]
        else:
            lead = xt.shape
        out = torch.empty(tuple(lead) + (self.dim,), dtype=xt.dtype, device=xt.device)
        spec = self._spec()
        _lib.check(_lib.lib().tnb_embed(C.byref(spec), _lib.F64 if xt.dtype == torch.float64 else _lib.F32,
                                        xt.numel() // n_in, xt.data_ptr(), out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
        return out


class TrigonometricEmbedding(Embedding):
    r"""phi(x) = k^{-1/2} [cos(pi x/2^i), i=1..k ; sin(pi x/2^i), i=1..k]  (embeddings.py:287-351)."""

    def __init__(self, k: int = 1, **kwargs):
        assert k >= 1, "k must be at least 1"
        self.k = k
        super().__init__(**kwargs)

    @property
    def dim(self) -> int:
        return self.k * 2

    def _spec(self):
        if self.dim > _lib.TNB_MAX_PHYS:
            raise ValueError(f"embedding dimension {self.dim} exceeds {_lib.TNB_MAX_PHYS}")
        s = _lib.TnbEmbedding()
        s.kind, s.k = _lib.EMB_TRIG, self.k
        return s


class FourierEmbedding(Embedding):
    r"""phi_j(x) = (1/p)|sum_k exp(2 pi i k((p-1)x - j)/p)|, dim == p  (embeddings.py:354-423)."""

    def __init__(self, p: int = 2, **kwargs):
        assert p >= 1, "Number of frequency components must be at least 1"
        self.p = p
        super().__init__(**kwargs)

    @property
    def dim(self) -> int:
        return self.p

    def _spec(self):
        if self.dim > _lib.TNB_MAX_PHYS:
            raise ValueError(f"embedding dimension {self.dim} exceeds {_lib.TNB_MAX_PHYS}")
        s = _lib.TnbEmbedding()
        s.kind, s.p = _lib.EMB_FOURIER, self.p
        return s


class GaussianRBFEmbedding(Embedding):
    r"""v_j = exp(-gamma (x - c_j)) (no square, as the reference), v/||v||  (embeddings.py:544-602)."""

    def __init__(self, centers: np.ndarray | None = None, gamma: float | None = None, **kwargs):
        self.centers = centers
        self.gamma = gamma
        super().__init__(**kwargs)

    @property
    def dim(self) -> int:
        return int(np.prod(np.array(np.asarray(self.centers).shape)))

    def _spec(self):
        c = np.asarray(self.centers, dtype=np.float64).reshape(-1)
        if c.size > _lib.TNB_MAX_PHYS:
            raise ValueError(f"embedding dimension {c.size} exceeds {_lib.TNB_MAX_PHYS}")
        s = _lib.TnbEmbedding()
        s.kind, s.n_centers, s.gamma = _lib.EMB_RBF, int(c.size), float(self.gamma)
        for i, v in enumerate(c):
            s.centers[i] = float(v)
        return s


class PolynomialEmbedding(Embedding):
    """[1?] + [x, x^2, ..., x^degree] for n = 1  (embeddings.py:605-717).

    Not unit norm: ``embed``'s normaliser (embeddings.py:1374-1387) is applied in-kernel.
    """

    def __init__(self, degree: int, n: int, include_bias: bool = False, include_cross_terms: bool = False,
                 **kwargs):
        if degree < 1:
            raise ValueError("Degree of PolynomialEmbedding must be at least 1.")
        self.degree = degree
        self.n = n
        self.include_bias = include_bias
        self.include_cross_terms = include_cross_terms
        super().__init__(**kwargs)

    @property
    def dim(self) -> int:
        import math
        total = 1 if self.include_bias else 0
        for d in range(1, self.degree + 1):
            total += math.comb(self.n + d - 1, d) if self.include_cross_terms else self.n
        return total

    @property
    def input_dim(self) -> int:
        return self.n

    def _spec(self):
        if self.dim > _lib.TNB_MAX_PHYS or self.n > _lib.TNB_MAX_PHYS:
            raise ValueError(f"embedding dimension {self.dim} exceeds {_lib.TNB_MAX_PHYS}")
        if self.include_cross_terms and self.n > 1 and self.degree > 8:
            raise ValueError("cross terms are built for degree <= 8")
        s = _lib.TnbEmbedding()
        s.kind, s.degree, s.include_bias = _lib.EMB_POLY, self.degree, int(bool(self.include_bias))
        s.n_in, s.cross = int(self.n), int(bool(self.include_cross_terms))
        return s


class LinearComplementEmbedding(Embedding):
    """[x, 1 - x] (p = 2) or [1, x, 1 - x] (p = 3), divided by its 2-norm  (embeddings.py:426-484)."""

    def __init__(self, p: int = 2, **kwargs):
        if p not in [2, 3]:
            raise ValueError("p must be 2 or 3")
        self.p = p
        super().__init__(**kwargs)

    @property
    def dim(self) -> int:
        return self.p

    def _spec(self):
        s = _lib.TnbEmbedding()
        s.kind, s.p = _lib.EMB_LINCOMP, self.p
        return s


class _OrthogonalPolynomialEmbedding(Embedding):
    """degree + 1 features from a three-term recurrence (tn4ml/scipy/special.py), evaluated in-kernel."""
    _kind = -1

    def __init__(self, degree: int = 2, **kwargs):
        self.degree = degree
        super().__init__(**kwargs)

    @property
    def dim(self) -> int:
        return self.degree + 1

    def _spec(self):
        if self.degree < 0 or self.dim > _lib.TNB_MAX_PHYS:
            raise ValueError(f"embedding dimension {self.dim} must be in [1, {_lib.TNB_MAX_PHYS}]")
        s = _lib.TnbEmbedding()
        s.kind, s.degree = self._kind, self.degree
        return s


class LegendreEmbedding(_OrthogonalPolynomialEmbedding):
    """[P_0(x) .. P_degree(x)], x in [-1, 1]  (embeddings.py:720-768)."""
    _kind = _lib.EMB_LEGENDRE


class LaguerreEmbedding(_OrthogonalPolynomialEmbedding):
    """exp(-x/2) [L_0(x) .. L_degree(x)], x >= 0  (embeddings.py:771-820)."""
    _kind = _lib.EMB_LAGUERRE


class HermiteEmbedding(_OrthogonalPolynomialEmbedding):
    """exp(-x^2/2) [H_0(x) .. H_degree(x)] (physicists' H)  (embeddings.py:823-869)."""
    _kind = _lib.EMB_HERMITE


class JaxArraysEmbedding(Embedding):
    """Pre-embedded inputs: site i of a sample is the vector x[i, :input_dim] itself, optionally behind a constant 1
    (embeddings.py:872-928).  The kernels read ``x [B, L, input_dim]``; named after the reference class it mirrors (the arrays
    here are torch / numpy)."""

    def __init__(self, dim: int | None = None, add_bias: bool = False, input_dim: int | None = None, **kwargs):
        super().__init__(**kwargs)
        self._dim = dim
        self.add_bias = add_bias
        self._input_dim = input_dim

    @property
    def dim(self) -> int:
        return self._dim

    @property
    def input_dim(self) -> int:
        return self._input_dim

    def _spec(self):
        n_in = self._input_dim if self._input_dim is not None else (self._dim - int(bool(self.add_bias)) if self._dim else None)
        if n_in is None or n_in < 1:
            raise ValueError("JaxArraysEmbedding needs dim or input_dim")
        d = n_in + int(bool(self.add_bias))
        if self._dim is not None and self._dim != d:
            raise ValueError(f"dim ({self._dim}) must equal input_dim ({n_in}) + add_bias")
        if d > _lib.TNB_MAX_PHYS:
            raise ValueError(f"embedding dimension {d} exceeds {_lib.TNB_MAX_PHYS}")
        s = _lib.TnbEmbedding()
        s.kind, s.n_in, s.include_bias = _lib.EMB_ARRAYS, int(n_in), int(bool(self.add_bias))
        return s


def spec_of(phi: Embedding) -> _lib.TnbEmbedding:
    if not isinstance(phi, Embedding):
        raise TypeError("Invalid embedding type")  # embed(): embeddings.py:1346-1349
    return phi._spec()
