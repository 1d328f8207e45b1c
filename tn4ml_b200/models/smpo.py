"""Spaced matrix product operator (reference: tn4ml/models/smpo.py)."""
from __future__ import annotations

import itertools
from typing import Any

import numpy as np
import torch

from .. import _lib
from .model import Model, _torch_dtype


class SpacedMatrixProductOperator(Model):
    """MPO with an output leg on sites ``i % spacing == 0`` (smpo.py:39-206), default 'lrud' order:
    site 0 (chi, d, d_o) [r,u,d]; output sites (chi_l, chi_r, d, d_o); others (chi_l, chi_r, d);
    last (chi, d) [l,u] or (chi, d, d_o)."""

    kind = _lib.KIND_SMPO
    _reg_norm_is_trace = True  # metrics.py:99-103: type(model) in [SpacedMatrixProductOperator]

    def __init__(self, arrays, output_inds=None, shape="lrud", spacing: int | None = None, **kwargs):
        if output_inds:
            raise NotImplementedError("uneven output_inds are not on the fused path (even spacing only)")
        if shape != "lrud":
            raise NotImplementedError("only the default 'lrud' axis order is supported")
        arrays = list(arrays)
        dims = [len(a.shape) for a in arrays]
        L = len(arrays)
        if spacing is None:  # smpo.py:128-142
            if dims.count(4) == 0:
                spacing = L if dims[-1] != 3 else L - 1
            elif dims.count(4) == 1:
                spacing = L if dims.index(4) == 0 else dims.index(4)
            else:
                spacing = dims.index(4, dims.index(4) + 1) - dims.index(4)
        d_o = int(arrays[0].shape[-1])
        self.spacing = spacing
        super().__init__(arrays, n_out=d_o, out_site=0, spacing=spacing, **kwargs)

    def _canon_index(self, i, shape):
        L, S = self.L, self.spacing
        has_out = i % S == 0
        full = slice(None)
        if i == 0:
            if len(shape) != 3:
                raise ValueError("site 0 must be (chi, d, d_o)")
            return (1, shape[0], shape[1], shape[2], (0, slice(0, shape[0]), full, full))
        if i == L - 1:
            if has_out:
                return (shape[0], 1, shape[1], shape[2], (slice(0, shape[0]), 0, full, full))
            return (shape[0], 1, shape[1], 1, (slice(0, shape[0]), 0, full, 0))
        if has_out:
            return (shape[0], shape[1], shape[2], shape[3], (slice(0, shape[0]), slice(0, shape[1]), full, full))
        return (shape[0], shape[1], shape[2], 1, (slice(0, shape[0]), slice(0, shape[1]), full, 0))


def generate_shape(method: str, L: int, has_out: bool = False, bond_dim: int = 2, phys_dim=(2, 2),
                   cyclic: bool = False, position: int | None = None, spacing: int | None = None):
    """smpo.py:496-573 (non-cyclic)."""
    if method == "even":
        cl = 1 if position == 1 else bond_dim
        cr = 1 if position == L else bond_dim
        return (cl, cr, *phys_dim) if has_out else (cl, cr, phys_dim[0])
    j = (L + 1 - abs(2 * position - L - 1)) // 2 if position > L // 2 else position
    chir = min(bond_dim, phys_dim[0] ** j * phys_dim[1] ** (j // spacing))
    chil = min(bond_dim, phys_dim[0] ** (j - 1) * phys_dim[1] ** ((j - 1) // spacing))
    if position > L // 2:
        chil, chir = chir, chil
    cl = 1 if position == 1 else chil
    cr = 1 if position == L else chir
    return (cl, cr, *phys_dim) if has_out else (cl, cr, phys_dim[0])


def SMPO_initialize(L: int, initializer, key: Any, dtype: Any = np.float64, shape_method: str = "even",  # noqa: N802
                    spacing: int = 2, bond_dim: int = 4, phys_dim=(2, 2), output_inds: list | None = None,
                    add_identity: bool = False, add_to_output: bool = False, boundary: str = "obc",
                    cyclic: bool = False, compress: bool = False, insert: int | None = None,
                    canonical_center: int | None = None, device=None, **kwargs):
    """smpo.py:576-764."""
    if output_inds:
        raise NotImplementedError("uneven output_inds are not on the fused path")
    if cyclic or compress or canonical_center is not None:
        raise NotImplementedError("cyclic / compress / canonical_center are outside the hot path")
    if spacing == 1:
        raise ValueError("Spacing must be > 1, otherwise is Matrix Product Operator.")
    tdt = _torch_dtype(dtype)
    tensors = []
    hasoutput = itertools.cycle([True, *[False] * (spacing - 1)])
    for i, has_out in zip(range(1, L + 1), hasoutput):
        shape = generate_shape(shape_method, L, has_out, bond_dim, phys_dim, cyclic, i, spacing)
        t = initializer(key, shape, tdt)
        # (add_identity uses .at[].set() without keeping the result in the reference,
        #  smpo.py:690-712: a no-op there; mirrored)
        if boundary == "obc":
            aux = torch.zeros_like(t)
            if i == 1:
                aux[:, 0] = t[:, 0]
                t = aux
            elif i == L:
                aux[0] = t[0]
                t = aux
        t = t / torch.linalg.norm(t)
        tensors.append(t.squeeze(0) if i == 1 else (t.squeeze(1) if i == L else t))
    if insert and insert < L and shape_method == "even":
        tensors[insert] = tensors[insert] / np.sqrt(min(bond_dim, phys_dim[0]))
    model = SpacedMatrixProductOperator(tensors, spacing=spacing, dtype=tdt, device=device)
    model.normalize()
    return model
