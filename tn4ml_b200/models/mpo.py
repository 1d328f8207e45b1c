"""Matrix product operator = SMPO with an output leg on every site (reference: tn4ml/models/mpo.py)."""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .model import _torch_dtype
from .smpo import SpacedMatrixProductOperator


class MatrixProductOperator(SpacedMatrixProductOperator):
    """Every site (chi_l, chi_r, d, d_o); boundaries squeezed (mpo.py:14-47, :99-104)."""

    _reg_norm_is_trace = False  # the reference's MPO is not a SpacedMatrixProductOperator: regularisers use model.norm()

    def __init__(self, arrays, **kwargs):
        super().__init__(arrays, spacing=1, **kwargs)


def MPO_initialize(L: int, initializer, key: Any, dtype: Any = np.float64, bond_dim: int = 4, phys_dim=(2, 2),  # noqa: N802
                   boundary: str = "obc", device=None, **kwargs):
    """mpo.py:126-254 (even shapes, open boundaries)."""
    tdt = _torch_dtype(dtype)
    tensors = []
    for i in range(1, L + 1):
        cl = 1 if i == 1 else bond_dim
        cr = 1 if i == L else bond_dim
        t = initializer(key, (cl, cr, *phys_dim), tdt)
        t = t / torch.linalg.norm(t)
        tensors.append(t.squeeze(0) if i == 1 else (t.squeeze(1) if i == L else t))
    model = MatrixProductOperator(tensors, dtype=tdt, device=device)
    model.normalize()
    return model
