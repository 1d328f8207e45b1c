"""Training strategies, mirroring ``tn4ml.strategy`` (tn4ml/strategy.py:6-258).

``Sweeps`` is the DMRG-like strategy: for every neighbouring pair of sites -- left to right, then (two-way) right to left -- the
model is brought to canonical form around the pair, the two site tensors are contracted into one, the loss is differentiated
with respect to that merged tensor, the optimizer updates it, and a truncated SVD splits it back into two sites.

How the work is divided here:
  * loss and the two-site gradient: the environment kernels of the hot path (``tnb_forward`` with the stash kept, the adjoint
    sweep of ``tnb_backward_part``) and one batched outer-product kernel for the merged tensor (``tnb_pair_grad``) --
    :meth:`tn4ml_b200.engine.Engine.pair_value_and_grad`;
  * canonical form, the contraction of the pair, the optimizer update and the SVD split: small dense algebra on the device
    (torch / cuSOLVER), per pair -- plumbing around the kernels, like the reference's quimb calls around its jitted loss.
Like the reference (model.py:829-866), every pair step re-evaluates the chain for the whole batch.
"""
from __future__ import annotations

import abc

import torch


class Strategy:
    """Decides which tensors the gradients are taken of (strategy.py:6-48)."""

    def __init__(self, renormalize: bool = False):
        self.renormalize = renormalize

    def prehook(self, model, sites):
        """Modify `model` before computing gradient(s)."""

    def posthook(self, model, sites):
        """Modify `model` after optimizing tensors."""

    @abc.abstractmethod
    def iterate_sites(self, sites):
        """Iterate over selected tensors."""


def _check_model(model):
    if getattr(model, "kind", None) != 0 or not hasattr(model, "_canon4"):
        raise TypeError("Sweeps needs a matrix-product-state model (MatrixProductState / MPS classifier)")


class Sweeps(Strategy):
    """Sweeping DMRG-like strategy (strategy.py:51-258)."""

    def __init__(self, grouping: int = 2, two_way: bool = True, split_opts=None, **kwargs):
        if split_opts is None:
            split_opts = {"cutoff": 0.0}
        if grouping > 2:
            raise ValueError(f"grouping - {grouping=} > 2")
        if grouping == 1:
            raise ValueError("grouping == 1")
        self.grouping = grouping
        self.two_way = two_way
        self.split_opts = split_opts
        self.bond_dim_split = None
        self._merged = None  # (pair, merged tensor T[a, r, s, s', k]) between prehook and posthook
        super().__init__(**kwargs)

    def iterate_sites(self, model):
        """Pairs in sweep order: (0,1), (1,2), ..., (L-2,L-1), then (L-1,L-2), ..., (1,0)  (strategy.py:117-131)."""
        _check_model(model)
        L = model.L
        sites = [tuple(i + j for j in range(self.grouping)) for i in range(L - self.grouping + 1)]
        if self.two_way:
            sites.extend(tuple(i - j for j in range(self.grouping)) for i in reversed(range(self.grouping - 1, L)))
        yield from sites

    # ---- contract the pair (strategy.py:133-176) ------------------------------------------------
    def prehook(self, model, sites):
        _check_model(model)
        model.canonicalize(set(sites), inplace=True)
        i = min(sites)
        A, B = model._canon4(i), model._canon4(i + 1)       # (chi_l, chi_m, d, n_i)  [a, m, s, k] / [m, r, t, q]
        self.bond_dim_split = A.shape[1]                    # model.bond_size(sites[0], sites[1])
        T = torch.einsum("amsk,mrtq->arstkq", A, B)
        T = T.reshape(T.shape[0], T.shape[1], T.shape[2], T.shape[3], -1)  # one of k, q has extent 1
        self._merged = (i, T.contiguous())
        return self._merged[1]

    # ---- split it back (strategy.py:178-258; quimb tensor_split: SVD, absorb="both", max_bond, cutoff) -----------
    def posthook(self, model, sites):
        _check_model(model)
        i, T = self._merged
        assert i == min(sites)
        if self.renormalize:
            T = T / torch.linalg.norm(T)
        A, B = model._canon4(i), model._canon4(i + 1)
        cl, cm, d, na = A.shape
        cr, nb = B.shape[1], B.shape[3]
        if na > 1:    # class leg on the left site: rows (a, s, k), columns (r, s')
            M = T.permute(0, 2, 4, 1, 3).reshape(cl * d * na, cr * d)
        else:         # rows (a, s), columns (r, s', k)
            M = T.permute(0, 2, 1, 3, 4).reshape(cl * d, cr * d * nb)
        U, S, Vh = torch.linalg.svd(M, full_matrices=False)
        keep = min(self.bond_dim_split, S.numel())
        cutoff = float(self.split_opts.get("cutoff", 0.0))
        if cutoff > 0.0:  # quimb's default cutoff_mode="rel": drop singular values below cutoff * s_max
            keep = min(keep, max(1, int((S > cutoff * S[0]).sum())))
        rs = torch.sqrt(S[:keep])
        Un = U[:, :keep] * rs                      # (rows, keep)
        Vn = rs[:, None] * Vh[:keep, :]            # (keep, cols)
        A.zero_()                                  # (a bond the split leaves below its size stays zero padded)
        B.zero_()
        A[:, :keep] = Un.reshape(cl, d, na, keep).permute(0, 3, 1, 2)
        B[:keep] = Vn.reshape(keep, cr, d, nb)
        model._mark_modified(i, i + 1)
        self._merged = None
