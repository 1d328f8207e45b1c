"""MPS models (reference: tn4ml/models/mps.py, tn4ml/models/tn.py)."""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .. import _lib
from ..initializers import randn
from .model import Model, _torch_dtype


class MatrixProductState(Model):
    """Plain (regression / unsupervised) MPS: site 0 (chi, d) [r,p], bulk (chi_l, chi_r, d),
    last (chi, d) [l,p]  (mps.py:15-51, boundaries squeezed at mps.py:451)."""

    kind = _lib.KIND_MPS

    def __init__(self, arrays, **kwargs):
        super().__init__(arrays, n_out=1, out_site=len(arrays) - 1, **kwargs)

    def _canon_index(self, i, shape):
        L = self.L
        if len(shape) == 2:
            if i == 0:
                return (1, shape[0], shape[1], 1, (0, slice(0, shape[0]), slice(None), 0))
            if i == L - 1:
                return (shape[0], 1, shape[1], 1, (slice(0, shape[0]), 0, slice(None), 0))
            raise ValueError("only boundary sites may be 2-D")
        if len(shape) == 3:
            return (shape[0], shape[1], shape[2], 1, (slice(0, shape[0]), slice(0, shape[1]), slice(None), 0))
        raise ValueError(f"site {i}: unsupported shape {shape}")


class TensorNetwork(Model):
    """Classifier MPS ("ParametrizedTensorNetwork", tn.py:15-113): every site (chi_l, chi_r, d),
    one class site (chi_l, chi_r, d, C)  (mps.py:308-390)."""

    kind = _lib.KIND_MPS

    def __init__(self, arrays, **kwargs):
        cls = [i for i, a in enumerate(arrays) if len(a.shape) == 4]
        if len(cls) != 1:
            raise ValueError("a classifier network needs exactly one site with a class index")
        super().__init__(arrays, n_out=int(arrays[cls[0]].shape[3]), out_site=cls[0], **kwargs)

    def _canon_index(self, i, shape):
        if len(shape) == 3:
            return (shape[0], shape[1], shape[2], 1, (slice(0, shape[0]), slice(0, shape[1]), slice(None), 0))
        if len(shape) == 4:
            return (shape[0], shape[1], shape[2], shape[3],
                    (slice(0, shape[0]), slice(0, shape[1]), slice(None), slice(None)))
        raise ValueError("Tensors need to always be 3D or 4D in MPS for classification.")


def generate_shape(method: str, L: int, bond_dim: int = 2, phys_dim: int = 2, cyclic: bool = False,
                   position: int | None = None, class_index: int | None = None, class_dim: int | None = None):
    """Site shape for 1-based ``position`` (mps.py:70-153)."""
    has_cls = class_index is not None and position == class_index
    tail = (phys_dim, class_dim) if has_cls else (phys_dim,)
    if method == "even":
        cl = 1 if position == 1 else bond_dim
        cr = 1 if position == L else bond_dim
        if position == 1 and position == L:
            cl, cr = 1, 1
        return (cl, cr) + tail
    assert not cyclic
    j = (L + 1 - abs(2 * position - L - 1)) // 2 if position > L // 2 else position
    chir = min(bond_dim, phys_dim**j)
    chil = min(bond_dim, phys_dim ** (j - 1))
    if position > L // 2:
        chil, chir = chir, chil
    if position == 1:
        return (1, chir) + tail
    if position == L:
        return (chil, 1) + tail
    return (chil, chir) + tail


def MPS_initialize(L: int, arrays: list | None = None, initializer=None, key: Any = None, dtype: Any = np.float64,  # noqa: N802
                   shape_method: str = "even", bond_dim: int = 4, phys_dim: int = 2, cyclic: bool = False,
                   add_identity: bool = False, add_to_output: bool = False, boundary: str = "obc",
                   class_index: int | None = None, class_dim: int | None = None, compress: bool = False,
                   insert: int | None = None, canonical_center: int | None = None, device=None, **kwargs):
    """mps.py:218-493.  ``class_index`` is 1-based (position of the class site), as in the reference."""
    if cyclic:
        raise NotImplementedError("cyclic (pbc) chains need a matrix boundary environment (SURVEY 8f)")
    if compress or canonical_center is not None:
        raise NotImplementedError("compress / canonical_center are model-only QR/SVD sweeps outside the hot path")
    tdt = _torch_dtype(dtype)
    if initializer is None:
        initializer = randn()
    if arrays is not None:
        assert class_index is not None
    if class_index is not None:
        if class_index > L:
            raise ValueError("class_index should be less than L.")
        if arrays is None:
            arrays = []
            for i in range(1, L + 1):
                shape = generate_shape(shape_method, L, bond_dim, phys_dim, cyclic, i, class_index, class_dim)
                a = initializer(key, shape, tdt)
                if add_identity:
                    eye = torch.eye(shape[0], shape[1], dtype=tdt)
                    if a.ndim == 3:
                        a[:, :, 0] += eye
                    elif add_to_output:
                        a[:, :, 0, :] += eye.unsqueeze(-1)
                if boundary == "obc":
                    aux = torch.zeros_like(a)
                    if i == 1:
                        aux[:, 0] = a[:, 0]
                        a = aux
                    elif i == L:
                        aux[0] = a[0]
                        a = aux
                arrays.append(a)
        model = TensorNetwork(arrays, dtype=tdt, device=device)
        model.normalize()
        return model
    tensors = []
    for i in range(1, L + 1):
        shape = generate_shape(shape_method, L, bond_dim, phys_dim, cyclic, i)
        t = initializer(key, shape, tdt)
        if add_identity and t.ndim == 3:
            # (the reference's regression branch calls .at[].add() without keeping the result,
            #  mps.py:431-436, so add_identity is a no-op there; mirrored)
            pass
        if boundary == "obc":
            aux = torch.zeros_like(t)
            if i == 1:
                aux[:, 0, :] = t[:, 0, :]
                t = aux
            elif i == L:
                aux[0, :, :] = t[0, :, :]
                t = aux
        tensors.append(t.squeeze(0) if i == 1 else (t.squeeze(1) if i == L else t))
    if insert and insert < L and shape_method == "even":
        tensors[insert] = tensors[insert] / np.sqrt(phys_dim)
    model = MatrixProductState(tensors, dtype=tdt, device=device)
    model.normalize()
    return model
