"""Synthetic inputs of the five BASELINE.json configurations: seeded, framework-neutral DATA generators
(numpy / torch on the CPU).  Nothing of the contraction path lives here -- only the site-tensor initialisation the
reference's examples use (identity + noise, tn4ml/models/mps.py:348-355, then Model.normalize, tn4ml/models/tn.py:103-107),
uniform inputs and one-hot labels (tn4ml/util.py:175-199).  Used by bench.py (both arms), the tests and the oracle.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _canon4(arrays):
    """site tensors as (chi_l, chi_r, d, C or 1), boundaries un-squeezed"""
    L = len(arrays)
    out = []
    for i, a in enumerate(arrays):
        if a.ndim == 2:
            a = a.unsqueeze(0) if i == 0 else a.unsqueeze(1)
        if a.ndim == 3:
            a = a.unsqueeze(-1)
        out.append(a)
    return out


def make_mps_arrays(L, chi, d, class_site=None, C=None, std=1e-2, seed=42, dtype=torch.float64,
                    squeeze_ends=False, shape_method="even", identity_all=False):
    """A_i = I delta_{s,0} + std N(0,1) (mps.py:348-355 add_identity), then a global
    division by norm^(1/L) as Model.normalize does (tn.py:103-107)."""
    g = torch.Generator().manual_seed(seed)
    arrays = []
    for i in range(L):
        if shape_method == "even":
            cl = 1 if i == 0 else chi
            cr = 1 if i == L - 1 else chi
        else:  # mps.py:124-151 ("noteven": bonds truncated to min(chi, d^j))
            pos = i + 1
            j = (L + 1 - abs(2 * pos - L - 1)) // 2 if pos > L // 2 else pos
            chir, chil = min(chi, d**j), min(chi, d ** (j - 1))
            if pos > L // 2:
                chil, chir = chir, chil
            cl = 1 if pos == 1 else chil
            cr = 1 if pos == L else chir
        shape = (cl, cr, d) + ((C,) if class_site == i else ())
        a = std * torch.randn(shape, generator=g, dtype=torch.float64)
        eye = torch.eye(cl, cr, dtype=torch.float64)
        if a.ndim == 3:
            if identity_all:
                a += eye.unsqueeze(-1)
            else:
                a[:, :, 0] += eye
        else:
            a[:, :, 0, :] += eye.unsqueeze(-1)
        arrays.append(a)
    # global normalisation <A|A> = 1, spread over sites
    nrm2 = mps_model_sqnorm(arrays)
    scale = torch.exp(-0.5 * torch.log(nrm2) / L)
    arrays = [(a * scale).to(dtype) for a in arrays]
    if squeeze_ends and class_site is None:
        arrays[0] = arrays[0].squeeze(0)
        arrays[-1] = arrays[-1].squeeze(1)
    return arrays


def mps_model_sqnorm(arrays) -> torch.Tensor:
    """<A|A> of the model alone (tn.py:66-80): transfer-matrix sweep."""
    sites = _canon4(arrays)
    E = torch.ones(1, 1, dtype=sites[0].dtype)
    for A in sites:
        E = torch.einsum("lm,lrsk,mnsk->rn", E, A, A)
    return E[0, 0]


def make_smpo_arrays(L, chi, d, d_o, spacing, std=1e-2, seed=42, dtype=torch.float64):
    """SMPO site tensors in the reference's squeezed shapes, identity-plus-noise."""
    g = torch.Generator().manual_seed(seed)
    arrays = []
    for i in range(L):
        cl = 1 if i == 0 else chi
        cr = 1 if i == L - 1 else chi
        has_out = i % spacing == 0
        shape = (cl, cr, d) + ((d_o,) if has_out else ())
        a = std * torch.randn(shape, generator=g, dtype=torch.float64)
        eye = torch.eye(cl, cr, dtype=torch.float64)
        if has_out:
            a[:, :, 0, :] += eye.unsqueeze(-1) / math.sqrt(d_o)
        else:
            a[:, :, 0] += eye
        arrays.append(a)
    arrays[0] = arrays[0].squeeze(0)
    arrays[-1] = arrays[-1].squeeze(1)
    # crude normalisation so that n = O(1): divide every site by its Frobenius norm^(1/..)
    # (SMPO_initialize divides each tensor by its norm, smpo.py:728); here we keep the
    # spectral scale near 1 so the chain neither explodes nor vanishes.
    return [a.to(dtype) for a in arrays]


def make_inputs(B, L, seed, dtype=torch.float64):
    """x ~ U(0,1]  (avoids the phi=0 NaN of bias-free polynomials at x=0)."""
    rng = np.random.default_rng(seed)
    x = 1.0 - rng.random((B, L))
    return torch.tensor(x, dtype=dtype)


def make_labels(B, C, seed=7, dtype=torch.float64):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, C, size=B)
    return torch.nn.functional.one_hot(torch.tensor(idx), C).to(dtype)
