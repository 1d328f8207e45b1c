"""CPU restatement of tn4ml's batched product-state contraction (fwd + bwd).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PINNED by golden vectors produced by executing the reference's own
source (tools/make_reference_golden.py over the stand-ins in tools/ref_shim; tests/test_reference_golden.py): feature maps,
embed()'s normaliser, MPS / classifier index conventions, every MPS loss head and finite differences of the executed loss.
The SMPO double layer is pinned through the reference's SpacedMatrixProductOperator constructor and its metrics' len(model) <
len(data) branch (same network, contracted by index names); apply_mps's own array bookkeeping is not executed (it depends on
quimb's internal index order).

Every function cites the reference file:line (relative to /root/reference) it
restates.  Arithmetic is torch on the CPU (float64 unless a dtype is passed);
gradients come from torch autograd *and*, independently, from the explicit
environment-based VJP in :func:`mps_env_vjp` / :func:`smpo_env_vjp`.

Conventions
-----------
``phi``      [B, L, d]   embedded, normalised product state (one row per sample)
MPS sites    list of L tensors, reference shapes:
             classifier (tn4ml/models/mps.py:308-390)  (chi_l, chi_r, d) with the
             class site (chi_l, chi_r, d, C); boundaries carry bond 1.
             plain MPS   (tn4ml/models/mps.py:412-464)  site 0 (chi, d) [r,p],
             bulk (chi_l, chi_r, d), last (chi, d) [l,p]   (boundaries squeezed).
SMPO sites   (tn4ml/models/smpo.py:161-205, 496-573, squeezed at :728)
             site 0 (chi, d, d_o) [r,u,d]; i % S == 0 -> (chi_l, chi_r, d, d_o);
             otherwise (chi_l, chi_r, d); last (chi, d) [l,u] or (chi, d, d_o).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np
import torch

# ----------------------------------------------------------------------------
# Embeddings  (tn4ml/embeddings.py)
# ----------------------------------------------------------------------------


@dataclass(frozen=True)
class EmbSpec:
    """Static description of one feature map (the four on the hot path)."""

    kind: str  # "trig" | "fourier" | "poly" | "rbf" | "lincomp" | "legendre" | "laguerre" | "hermite" | "arrays"
    k: int = 1  # trig: number of frequencies (dim = 2k)
    p: int = 2  # fourier: dim = p; lincomp: 2 or 3
    degree: int = 1  # poly, legendre, laguerre, hermite
    include_bias: bool = False  # poly; arrays (add_bias)
    gamma: float = 1.0  # rbf
    centers: tuple = field(default_factory=tuple)  # rbf
    n: int = 1  # poly / arrays: input values per site (Embedding.input_dim)
    cross: bool = False  # poly: include_cross_terms

    @property
    def dim(self) -> int:
        if self.kind == "trig":
            return 2 * self.k  # embeddings.py:317-320
        if self.kind == "fourier":
            return self.p  # embeddings.py:385-387 (returns p, not 2p)
        if self.kind == "poly":  # embeddings.py:655-673
            total = 1 if self.include_bias else 0
            for dg in range(1, self.degree + 1):
                total += math.comb(self.n + dg - 1, dg) if self.cross else self.n
            return total
        if self.kind == "rbf":
            return len(self.centers)  # embeddings.py:581-583
        if self.kind == "lincomp":
            return self.p  # embeddings.py:458-461
        if self.kind in ("legendre", "laguerre", "hermite"):
            return self.degree + 1  # embeddings.py:743-746, 794-797, 846-849
        if self.kind == "arrays":
            return self.n + (1 if self.include_bias else 0)  # embeddings.py:913-928
        raise ValueError(self.kind)


def feature_map(x: torch.Tensor, spec: EmbSpec) -> torch.Tensor:
    """phi(x) for every scalar in ``x`` (every vector x[..., :n] for the maps with input_dim n > 1) -> [..., d], *before*
    embed()'s normaliser."""
    dt = x.dtype
    if spec.kind == "arrays":
        # embeddings.py:913-928 : the input vector itself, behind a constant 1 if add_bias
        return torch.cat([torch.ones_like(x[..., :1]), x], dim=-1) if spec.include_bias else x.clone()
    if spec.kind == "poly" and spec.n > 1:
        # embeddings.py:693-717 : degree by degree, itertools.combinations_with_replacement over the inputs; cross terms
        # (more than one distinct input) are skipped unless include_cross_terms
        import itertools
        feats = [torch.ones_like(x[..., 0])] if spec.include_bias else []
        for dg in range(1, spec.degree + 1):
            for comb in itertools.combinations_with_replacement(range(spec.n), dg):
                if len(set(comb)) > 1 and not spec.cross:
                    continue
                feats.append(torch.prod(x[..., list(comb)], dim=-1))
        return torch.stack(feats, dim=-1)
    if spec.n > 1 and spec.kind == "poly":
        raise ValueError("unreachable")
    if spec.kind in ("legendre", "laguerre", "hermite"):
        # tn4ml/scipy/special.py:6-28, 57-79, 108-130 (three-term recurrences), embeddings.py:754-768, 805-820, 857-869
        vals = [torch.ones_like(x)]
        if spec.kind == "legendre":
            if spec.degree >= 1:
                vals.append(x.clone())
            for i in range(1, spec.degree):
                vals.append(((2 * i + 1) * x * vals[i] - i * vals[i - 1]) / (i + 1))
            w = torch.ones_like(x)
        elif spec.kind == "laguerre":
            if spec.degree >= 1:
                vals.append(1.0 - x)
            for i in range(2, spec.degree + 1):
                vals.append(((2 * i - 1 - x) * vals[i - 1] - (i - 1) * vals[i - 2]) / i)
            w = torch.exp(-x / 2)
        else:
            if spec.degree >= 1:
                vals.append(2.0 * x)
            for i in range(2, spec.degree + 1):
                vals.append(2 * x * vals[i - 1] - 2 * (i - 1) * vals[i - 2])
            w = torch.exp(-0.5 * x**2)
        return torch.stack([w * v for v in vals], dim=-1)
    if spec.kind == "lincomp":
        # embeddings.py:478-484 : [x, 1 - x] or [1, x, 1 - x], / ||.||
        comps = [x, 1.0 - x] if spec.p == 2 else [torch.ones_like(x), x, 1.0 - x]
        v = torch.stack(comps, dim=-1)
        return v / torch.linalg.norm(v, dim=-1, keepdim=True)
    x = x.unsqueeze(-1)
    if spec.kind == "trig":
        # embeddings.py:340-351 : 1/sqrt(k) * [cos(pi x/2^i) i=1..k, sin(pi x/2^i) i=1..k]
        i = torch.arange(1, spec.k + 1, dtype=dt)
        ang = math.pi * x / (2.0**i)
        return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1) / math.sqrt(spec.k)
    if spec.kind == "fourier":
        # embeddings.py:407-423 : (1/p) |sum_k exp(2 pi i k ((p-1)x - j)/p)|, j=0..p-1
        p = spec.p
        j = torch.arange(p, dtype=dt)
        kk = torch.arange(p, dtype=dt).reshape(p, 1)
        ang = 2.0 * math.pi * kk * (((p - 1) * x - j) / p).unsqueeze(-2)  # [..., k, j]
        re = torch.cos(ang).sum(dim=-2)
        im = torch.sin(ang).sum(dim=-2)
        return torch.sqrt(re * re + im * im) / p
    if spec.kind == "poly":
        # embeddings.py:693-717 with n=1, no cross terms: [1?] + [x, x^2, ..., x^degree]
        feats = []
        if spec.include_bias:
            feats.append(torch.ones_like(x))
        for dgr in range(1, spec.degree + 1):
            feats.append(x**dgr)
        return torch.cat(feats, dim=-1)
    if spec.kind == "rbf":
        # embeddings.py:601-602 : exp(-gamma (x - c_j)) (no square), then / ||.||
        c = torch.tensor(spec.centers, dtype=dt)
        v = torch.exp(-spec.gamma * (x - c))
        return v / torch.linalg.norm(v, dim=-1, keepdim=True)
    raise ValueError(spec.kind)


def embed_batch(x: torch.Tensor, spec: EmbSpec) -> torch.Tensor:
    """embed() for a batch: x [B, L] -> normalised product state phi [B, L, d].

    embeddings.py:1354-1358 builds one (1,1,d) tensor per feature; :1374-1387
    normalises: L <= 200 divides every site by norm**(1/L), norm = prod_i ||phi_i||
    (the MPS norm of a product state); L > 200 normalises site by site interleaved
    with QRs whose signs cancel pairwise.  Both give the unit-norm product state.
    """
    phi = feature_map(x, spec)  # [B, L, d]
    L = phi.shape[1]
    site_norm = torch.linalg.norm(phi, dim=-1)  # [B, L]
    if L > 200:
        return phi / site_norm.unsqueeze(-1)
    log_norm = torch.log(site_norm).sum(dim=1, keepdim=True)  # log prod ||phi_i||
    scale = torch.exp(log_norm / L)  # norm ** (1/L)
    return phi / scale.unsqueeze(-1)


# ----------------------------------------------------------------------------
# Site-tensor canonicalisation (layouts only)
# ----------------------------------------------------------------------------


def canon_mps_sites(arrays: Sequence[torch.Tensor]) -> tuple[list[torch.Tensor], int]:
    """Bring MPS site tensors to 4-D (chi_l, chi_r, d, C); return (sites, class_site).

    class_site = -1 when no site carries a class leg (plain MPS, mps.py:412-464).
    """
    L = len(arrays)
    out, cls = [], -1
    for i, a in enumerate(arrays):
        if a.ndim == 2:  # squeezed boundary of a plain MPS (mps.py:451)
            a = a.unsqueeze(0) if i == 0 else a.unsqueeze(1)  # (1,r,p) / (l,1,p)
            if L == 1:
                raise ValueError("single-site MPS not supported")
        if a.ndim == 3:
            a = a.unsqueeze(-1)
        elif a.ndim == 4:
            if cls >= 0:
                raise ValueError("more than one class site")
            cls = i
        else:
            raise ValueError(f"site {i}: ndim {a.ndim}")
        out.append(a)
    return out, cls


def canon_smpo_sites(arrays: Sequence[torch.Tensor], spacing: int) -> list[torch.Tensor]:
    """Bring SMPO site tensors to 4-D (chi_l, chi_r, d, d_o or 1)."""
    L = len(arrays)
    out = []
    for i, a in enumerate(arrays):
        has_out = i % spacing == 0
        if i == 0:
            a = a.unsqueeze(0)  # (chi, d, d_o) -> (1, chi, d, d_o)   smpo.py:186
        elif i == L - 1:
            a = a.unsqueeze(1)  # (chi, d[, d_o]) -> (chi, 1, d[, d_o])  smpo.py:199
        if not has_out:
            a = a.unsqueeze(-1)
        if a.ndim != 4:
            raise ValueError(f"SMPO site {i}: bad shape {tuple(a.shape)}")
        out.append(a)
    return out


# ----------------------------------------------------------------------------
# MPS contraction  (metrics.py:47, :473-476 ; model.py:312-316)
# ----------------------------------------------------------------------------


def mps_outputs(phi: torch.Tensor, arrays: Sequence[torch.Tensor], order: str = "gemm") -> torch.Tensor:
    """(model & data) ^ all for every sample -> y [B, C] (C = 1 without a class site).

    order="ref"  : what XLA receives from quimb -- materialise the per-sample site
                   matrices M_i[b] = sum_s A_i[:,:,s] phi_i[b,s], then a batched
                   vector-matrix chain.
    order="gemm" : (L_i[b,:] @ A_i) then the phi-weighted sum over s.
    """
    sites, cls = canon_mps_sites(arrays)
    B = phi.shape[0]
    L = len(sites)
    c = cls if cls >= 0 else L - 1

    def step_left(v, A, ph):  # v [B,chi_l], A (chi_l,chi_r,d,1) -> [B,chi_r]
        A3 = A[..., 0]
        if order == "ref":
            M = torch.einsum("lrs,bs->blr", A3, ph)
            return torch.bmm(v.unsqueeze(1), M).squeeze(1)
        T = (v @ A3.reshape(A3.shape[0], -1)).reshape(B, A3.shape[1], A3.shape[2])
        return (T * ph.unsqueeze(1)).sum(-1)

    def step_right(v, A, ph):  # v [B,chi_r] -> [B,chi_l]
        A3 = A[..., 0]
        if order == "ref":
            M = torch.einsum("lrs,bs->blr", A3, ph)
            return torch.bmm(M, v.unsqueeze(2)).squeeze(2)
        At = A3.permute(1, 0, 2).reshape(A3.shape[1], -1)
        T = (v @ At).reshape(B, A3.shape[0], A3.shape[2])
        return (T * ph.unsqueeze(1)).sum(-1)

    lv = torch.ones(B, 1, dtype=phi.dtype)
    for i in range(c):
        lv = step_left(lv, sites[i], phi[:, i])
    rv = torch.ones(B, 1, dtype=phi.dtype)
    for i in range(L - 1, c, -1):
        rv = step_right(rv, sites[i], phi[:, i])
    Ac = sites[c]  # (chi_l, chi_r, d, C)
    Mc = torch.einsum("lrsk,bs->blrk", Ac, phi[:, c])
    return torch.einsum("bl,blrk,br->bk", lv, Mc, rv)


def mps_nll_logspace(phi: torch.Tensor, arrays: Sequence[torch.Tensor]) -> torch.Tensor:
    """-log(<A|phi>^2) with the chain renormalised at every site (plain MPS only).  Same
    quantity as head_loss(mps_outputs(...), "nll") wherever that does not under/overflow."""
    sites, cls = canon_mps_sites(arrays)
    assert cls < 0
    B = phi.shape[0]
    v = torch.ones(B, 1, dtype=phi.dtype)
    logs = torch.zeros(B, dtype=phi.dtype)
    for i, A in enumerate(sites):
        M = torch.einsum("lrs,bs->blr", A[..., 0], phi[:, i])
        v = torch.bmm(v.unsqueeze(1), M).squeeze(1)
        n = v.abs().amax(dim=1)
        v = v / n[:, None]
        logs = logs + torch.log(n)
    return -2.0 * (torch.log(v[:, 0].abs()) + logs)


# ----------------------------------------------------------------------------
# SMPO / MPO contraction  (smpo.py:269-404 ; metrics.py:56-81)
# ----------------------------------------------------------------------------


def smpo_sqnorm(phi: torch.Tensor, arrays: Sequence[torch.Tensor], spacing: int) -> torch.Tensor:
    """n[b] = <v|v>, v = model.apply(data)  (the MPS over the output legs).

    apply_mps contracts every physical leg (smpo.py:302-303), merges the leg-less
    matrices of each spacing window into its output tensor (:309-331) and returns an
    MPS with one (chi, chi, d_o) tensor per window; <v|v> is its double-layer
    contraction (metrics.py:81 then squares it).
    """
    sites = canon_smpo_sites(arrays, spacing)
    B = phi.shape[0]
    E = torch.ones(B, 1, 1, dtype=phi.dtype)
    L = len(sites)
    i = 0
    while i < L:
        A = sites[i]
        T = torch.einsum("lrso,bs->bolr", A, phi[:, i])  # [B, d_o, chi_l, chi_r]
        j = i + 1
        while j < L and j % spacing != 0:
            M = torch.einsum("lrs,bs->blr", sites[j][..., 0], phi[:, j])
            T = torch.einsum("bolm,bmr->bolr", T, M)
            j += 1
        E = torch.einsum("bolr,blm,bomn->brn", T, E, T)
        i = j
    return E[:, 0, 0]


def smpo_log_sqnorm(phi: torch.Tensor, arrays: Sequence[torch.Tensor], spacing: int) -> torch.Tensor:
    """log n[b] of :func:`smpo_sqnorm` with the window environment renormalised after every window (the same quantity
    wherever the unscaled recursion stays inside the floating-point range).  Used to balance long synthetic chains
    (L = 784) so that the unscaled oracle -- which is what the reference computes -- neither under- nor overflows."""
    sites = canon_smpo_sites(arrays, spacing)
    B, L = phi.shape[0], len(sites)
    E = torch.ones(B, 1, 1, dtype=phi.dtype)
    logs = torch.zeros(B, dtype=phi.dtype)
    i = 0
    while i < L:
        T = torch.einsum("lrso,bs->bolr", sites[i], phi[:, i])
        j = i + 1
        while j < L and j % spacing != 0:
            T = torch.einsum("bolm,bmr->bolr", T, torch.einsum("lrs,bs->blr", sites[j][..., 0], phi[:, j]))
            j += 1
        E = torch.einsum("bolr,blm,bomn->brn", T, E, T)
        m = E.abs().amax(dim=(1, 2))
        E = E / m[:, None, None]
        logs = logs + torch.log(m)
        i = j
    return logs + torch.log(E[:, 0, 0])


# ----------------------------------------------------------------------------
# Loss heads  (tn4ml/metrics.py)
# ----------------------------------------------------------------------------

HEADS = (
    "nll",  # NegLogLikelihood            metrics.py:17-53
    "softmax_xent",  # OptaxWrapper(optax.softmax_cross_entropy)  :423-487
    "softmax_xent_weighted",  # CrossEntropyWeighted       :490-532
    "xent_softmax_argmax",  # CrossEntropySoftmax          :291-333
    "mse",  # MeanSquaredError                             :336-396
    "tsn",  # TransformedSquaredNorm                       :56-81
    "logquad",  # LogQuadNorm                              :171-187
    "quad",  # QuadNorm                                    :190-206
)


def head_loss(out: torch.Tensor, head: str, targets: torch.Tensor | None = None,
              class_weights: torch.Tensor | None = None) -> torch.Tensor:
    """Per-sample loss from the contraction result (``y`` [B,C] or ``n`` [B])."""
    if head == "nll":
        s = out.reshape(out.shape[0])
        return -torch.log(s**2)  # metrics.py:53
    if head == "tsn":
        return out**2  # metrics.py:81
    if head == "logquad":
        return (torch.log(out**2) - 1.0) ** 2  # metrics.py:187
    if head == "quad":
        return (out**2 - 1.0) ** 2  # metrics.py:206
    yhat = out / torch.linalg.norm(out, dim=-1, keepdim=True)  # metrics.py:479 / :331 / :394
    if head in ("softmax_xent", "softmax_xent_weighted"):
        # optax.softmax_cross_entropy(logits, labels) = -sum(labels * log_softmax(logits))
        l = -(targets * torch.log_softmax(yhat, dim=-1)).sum(-1)
        if head == "softmax_xent_weighted":
            l = l * (targets * class_weights).sum(-1)  # metrics.py:520-527
        return l
    if head == "xent_softmax_argmax":
        idx = targets.argmax(dim=-1)  # metrics.py:333
        return -torch.log_softmax(yhat, dim=-1).gather(1, idx[:, None]).squeeze(1)
    if head == "mse":
        return ((yhat - targets) ** 2).mean(-1)  # metrics.py:396
    raise ValueError(head)


# ----------------------------------------------------------------------------
# loss_fn / value_and_grad  (models/model.py:723-751, 554-566)
# ----------------------------------------------------------------------------


@dataclass
class Problem:
    kind: str  # "mps" | "smpo"
    emb: EmbSpec
    head: str
    spacing: int = 0  # smpo only (1 == MPO)
    class_weights: tuple | None = None


def per_sample_loss(prob: Problem, x: torch.Tensor, arrays: Sequence[torch.Tensor],
                    targets: torch.Tensor | None = None, order: str = "gemm") -> torch.Tensor:
    """evaluate()'s closure (model.py:1048-1076): embed, contract, head -> [B]."""
    phi = embed_batch(x, prob.emb)
    if prob.kind == "mps" and phi.shape[1] > len(arrays):
        # len(model) < len(data) (metrics.py:36-43, 312-319, 448-455): every k_i of the data is contracted; the sites the model
        # does not reach carry no partner and are summed over their physical index
        Lm = len(arrays)
        extra = phi[:, Lm:].sum(-1).prod(dim=1)  # [B]
        out = mps_outputs(phi[:, :Lm], arrays, order=order) * extra[:, None]
    elif prob.kind == "mps":
        out = mps_outputs(phi, arrays, order=order)
    elif prob.kind == "smpo":
        out = smpo_sqnorm(phi, arrays, prob.spacing)
    else:
        raise ValueError(prob.kind)
    cw = None if prob.class_weights is None else torch.tensor(prob.class_weights, dtype=x.dtype)
    return head_loss(out, prob.head, targets, cw)


def loss_fn(prob: Problem, x, arrays, targets=None, order: str = "gemm") -> torch.Tensor:
    """train()'s closure (model.py:723-751): mean over the batch of the per-sample loss."""
    return per_sample_loss(prob, x, arrays, targets, order).mean()


def value_and_grad(prob: Problem, x, arrays, targets=None, order: str = "gemm"):
    """create_train_step.value_and_grad (model.py:554-566): loss and one grad per site."""
    leaves = [a.detach().clone().requires_grad_(True) for a in arrays]
    loss = loss_fn(prob, x, leaves, targets, order)
    grads = torch.autograd.grad(loss, leaves)
    return loss.detach(), [g.detach() for g in grads]


# ----------------------------------------------------------------------------
# Independent gradient route: explicit left/right environments (no autograd
# through the chain).  Used to validate autograd (SURVEY 8c route iii) and as the
# blueprint of the CUDA VJP kernels.
# ----------------------------------------------------------------------------


def mps_env_vjp(phi: torch.Tensor, arrays: Sequence[torch.Tensor], dy: torch.Tensor):
    """Given dLoss/dy [B,C], return [dLoss/dA_i] by environments.

    F[j] = env on bond j (left env for j <= c, right env for j >= c+1);
    G[j] = its adjoint.  dA_j = sum_b F[j] (x) phi_j (x) G[j+1]   (j < c)
           dA_j = sum_b G[j] (x) phi_j (x) F[j+1]   (j > c)
           dA_c = sum_b F[c] (x) phi_c (x) F[c+1] (x) dy.
    """
    sites, cls = canon_mps_sites(arrays)
    B, L = phi.shape[0], len(sites)
    c = cls if cls >= 0 else L - 1
    one = torch.ones(B, 1, dtype=phi.dtype)
    M = [torch.einsum("lrsk,bs->blrk", sites[i], phi[:, i]) for i in range(L)]
    F = [None] * (L + 1)
    F[0], F[L] = one, one
    for i in range(c):
        F[i + 1] = torch.einsum("bl,blr->br", F[i], M[i][..., 0])
    for i in range(L - 1, c, -1):
        F[i] = torch.einsum("blr,br->bl", M[i][..., 0], F[i + 1])
    G = [None] * (L + 1)
    G[c] = torch.einsum("bk,blrk,br->bl", dy, M[c], F[c + 1])
    G[c + 1] = torch.einsum("bk,blrk,bl->br", dy, M[c], F[c])
    for i in range(c - 1, 0, -1):
        G[i] = torch.einsum("blr,br->bl", M[i][..., 0], G[i + 1])
    for i in range(c + 1, L - 1):
        G[i + 1] = torch.einsum("bl,blr->br", G[i], M[i][..., 0])
    grads = []
    for i in range(L):
        if i < c:
            g = torch.einsum("bl,bs,br->lrs", F[i], phi[:, i], G[i + 1]).unsqueeze(-1)
        elif i > c:
            g = torch.einsum("bl,bs,br->lrs", G[i], phi[:, i], F[i + 1]).unsqueeze(-1)
        else:
            g = torch.einsum("bl,bs,br,bk->lrsk", F[c], phi[:, c], F[c + 1], dy)
        grads.append(g.reshape(arrays[i].shape))
    return grads


def smpo_env_vjp(phi: torch.Tensor, arrays: Sequence[torch.Tensor], spacing: int, dn: torch.Tensor):
    """Given dLoss/dn [B], return [dLoss/dA_i] by window environments (no autograd)."""
    sites = canon_smpo_sites(arrays, spacing)
    B, L = phi.shape[0], len(sites)
    starts = list(range(0, L, spacing))
    # forward, keeping per-window intermediates
    E = [torch.ones(B, 1, 1, dtype=phi.dtype)]
    win = []
    for w0 in starts:
        w1 = min(w0 + spacing, L)
        Mo = torch.einsum("lrso,bs->bolr", sites[w0], phi[:, w0])
        Ms = [torch.einsum("lrs,bs->blr", sites[j][..., 0], phi[:, j]) for j in range(w0 + 1, w1)]
        Q = [None]  # prefix products Q_k = M_1 ... M_k ; Q_0 = identity (None)
        for Mj in Ms:
            Q.append(Mj if Q[-1] is None else torch.bmm(Q[-1], Mj))
        P = Q[-1]
        T = Mo if P is None else torch.einsum("bolm,bmr->bolr", Mo, P)
        U = torch.einsum("blm,bomr->bolr", E[-1], T)
        E.append(torch.einsum("bolr,boln->brn", T, U))
        win.append((w0, w1, Mo, Ms, Q, T, U))
    grads = [None] * L
    dE = dn.reshape(B, 1, 1).clone()
    for (w0, w1, Mo, Ms, Q, T, U), Ew in zip(reversed(win), reversed(E[:-1])):
        dEs = 0.5 * (dE + dE.transpose(1, 2))
        dT = 2.0 * torch.einsum("bolr,brn->boln", U, dEs)
        dE = torch.einsum("bolr,brn,bomn->blm", T, dEs, T)
        P = Q[-1]
        if P is None:
            dMo = dT
        else:
            dMo = torch.einsum("boln,bmn->bolm", dT, P)
            Gm = torch.einsum("bolm,boln->bmn", Mo, dT)  # dP
            for k in range(len(Ms), 0, -1):
                dMk = Gm if Q[k - 1] is None else torch.bmm(Q[k - 1].transpose(1, 2), Gm)
                j = w0 + k
                grads[j] = torch.einsum("blr,bs->lrs", dMk, phi[:, j]).reshape(arrays[j].shape)
                Gm = torch.bmm(Gm, Ms[k - 1].transpose(1, 2))
        grads[w0] = torch.einsum("bolr,bs->lrso", dMo, phi[:, w0]).reshape(arrays[w0].shape)
    return grads


# ----------------------------------------------------------------------------
# Brute-force dense contraction for tiny L (validation route i).
# util.from_mps_to_dense (tn4ml/util.py:336-365) shows the amplitude convention.
# ----------------------------------------------------------------------------


def mps_dense_outputs(phi: torch.Tensor, arrays: Sequence[torch.Tensor]) -> torch.Tensor:
    """Build the full d^L x C amplitude tensor of the MPS and contract it with the
    product state explicitly (exponential; L <= 10)."""
    sites, cls = canon_mps_sites(arrays)
    L = len(sites)
    if L > 10:
        raise ValueError("dense route is exponential")
    # amplitude tensor psi[s_0..s_{L-1}, k]
    cur = torch.ones(1, 1, 1, dtype=phi.dtype)  # [config, bond, k]
    for A in sites:
        # cur [n, l, k] x A (l, r, s, k') -> [n, s, r, k*k'] ; only one site has k' > 1
        nxt = torch.einsum("nlk,lrsq->nsrkq", cur, A)
        n, s, r, k, q = nxt.shape
        cur = nxt.reshape(n * s, r, k * q)
    psi = cur[:, 0, :]  # [d^L, C]
    B = phi.shape[0]
    amp = torch.ones(B, 1, dtype=phi.dtype)
    for i in range(L):
        amp = (amp.unsqueeze(-1) * phi[:, i].unsqueeze(1)).reshape(B, -1)
    return amp @ psi


def smpo_dense_sqnorm(phi: torch.Tensor, arrays: Sequence[torch.Tensor], spacing: int) -> torch.Tensor:
    """v = P|phi> built as an explicit dense vector over all output legs, then <v|v>."""
    sites = canon_smpo_sites(arrays, spacing)
    L = len(sites)
    if L > 12:
        raise ValueError("dense route is exponential")
    B = phi.shape[0]
    cur = torch.ones(B, 1, 1, dtype=phi.dtype)  # [b, outcfg, bond]
    for i, A in enumerate(sites):
        M = torch.einsum("lrso,bs->bolr", A, phi[:, i])
        nxt = torch.einsum("bnl,bolr->bnor", cur, M)
        cur = nxt.reshape(B, -1, nxt.shape[-1])
    v = cur[:, :, 0]
    return (v * v).sum(-1)


# ----------------------------------------------------------------------------
# Model-only quantities: <A|A>, the Frobenius regularisers, normalize
# ----------------------------------------------------------------------------


def model_log_sqnorm(sites4: Sequence[torch.Tensor]) -> torch.Tensor:
    """log <A|A> of a chain of 4-D sites (chi_l, chi_r, d, n) by the transfer-matrix sweep (tn.py:66-80: norm = sqrt(conj & self
    contracted)), renormalised per site so that long chains stay in range."""
    E = torch.ones(1, 1, dtype=sites4[0].dtype)
    logs = torch.zeros((), dtype=sites4[0].dtype)
    for A in sites4:
        E = torch.einsum("lm,lrsk,mnsk->rn", E, A, A)
        m = E.abs().max()
        E = E / m
        logs = logs + torch.log(m)
    return logs + torch.log(E[0, 0])


def regulariser(sites4: Sequence[torch.Tensor], kind: str, norm_is_trace: bool = False) -> torch.Tensor:
    """metrics.py:89-168.  l = log(norm); norm = sqrt(<A|A>) (model.norm()), or <A|A> itself when the model is a
    SpacedMatrixProductOperator (metrics.py:99-103: model.H.apply(model) fully contracted = Tr(P^dag P))."""
    l = model_log_sqnorm(sites4) * (1.0 if norm_is_trace else 0.5)
    if kind == "log_frob":
        return l  # metrics.py:105
    if kind == "log_pow_frob":
        return 2.0 * l  # log(norm^2)  metrics.py:124
    if kind == "log_relu_frob":
        return torch.clamp(l, min=0.0)  # metrics.py:146
    if kind == "quad_frob":
        return (l - 1.0) ** 2  # metrics.py:168
    raise ValueError(kind)


def semisup_per_sample(emb: EmbSpec, x, arrays, spacing: int, y, coef: float = 0.3, norm_is_trace: bool = True):
    """SemiSupervisedLoss (metrics.py:209-233): m = LogQuadNorm(model, data) + 0.3 LogReLUFrobNorm(model);
    loss = (y / m + (1 - y) m)^2 per sample, y the target class fraction."""
    phi = embed_batch(x, emb)
    lq = head_loss(smpo_sqnorm(phi, arrays, spacing), "logquad")
    m = lq + coef * regulariser(canon_smpo_sites(arrays, spacing), "log_relu_frob", norm_is_trace)
    y = y.reshape(-1)
    return (y / m + (1.0 - y) * m) ** 2


def semisup_value_and_grad(emb: EmbSpec, x, arrays, spacing: int, y, coef: float = 0.3, norm_is_trace: bool = True):
    leaves = [a.detach().clone().requires_grad_(True) for a in arrays]
    loss = semisup_per_sample(emb, x, leaves, spacing, y, coef, norm_is_trace).mean()
    grads = torch.autograd.grad(loss, leaves)
    return loss.detach(), [g.detach() for g in grads]


# ----------------------------------------------------------------------------
# Synthetic problems of the five BASELINE.json configs: framework-neutral DATA generators, shared with
# bench.py's product arm, live in tn4ml_b200/synthetic.py (no algorithm of the path in there); re-exported here so that
# the tests keep one namespace.
# ----------------------------------------------------------------------------
from tn4ml_b200.synthetic import (  # noqa: E402,F401
    make_inputs, make_labels, make_mps_arrays, make_smpo_arrays, mps_model_sqnorm)


# ----------------------------------------------------------------------------
# Sweeps / DMRG-like strategy  (tn4ml/strategy.py:51-258; training loop tn4ml/models/model.py:753-773, 829-866)
# ----------------------------------------------------------------------------


def _qr_left(sites, i):
    """left_canonize_site: A_i = Q R over rows (l, s, k); R into site i+1."""
    A = sites[i]
    cl, cr, d, n = A.shape
    Q, R = torch.linalg.qr(A.permute(0, 2, 3, 1).reshape(cl * d * n, cr))
    k = Q.shape[1]
    if k < cr:
        Q = torch.cat([Q, Q.new_zeros(Q.shape[0], cr - k)], dim=1)
        R = torch.cat([R, R.new_zeros(cr - k, cr)], dim=0)
    sites[i] = Q.reshape(cl, d, n, cr).permute(0, 3, 1, 2)
    sites[i + 1] = torch.einsum("ab,brsk->arsk", R, sites[i + 1])


def _qr_right(sites, i):
    A = sites[i]
    cl, cr, d, n = A.shape
    Q, R = torch.linalg.qr(A.permute(1, 2, 3, 0).reshape(cr * d * n, cl))
    k = Q.shape[1]
    if k < cl:
        Q = torch.cat([Q, Q.new_zeros(Q.shape[0], cl - k)], dim=1)
        R = torch.cat([R, R.new_zeros(cl - k, cl)], dim=0)
    sites[i] = Q.reshape(cr, d, n, cl).permute(3, 0, 1, 2)
    sites[i - 1] = torch.einsum("lask,ba->lbsk", sites[i - 1], R)


def _as_arrays(sites, cls):
    return [a if i == cls else a[..., 0] for i, a in enumerate(sites)]


def pair_loss_and_grad(prob: Problem, x, sites, cls, i, targets=None):
    """Mean loss and its gradient with respect to the merged tensor T[a, r, s, s', k] of sites (i, i+1), by autograd: the pair
    is replaced by T itself on site i (its right bond enlarged to (r, s', k)) and a constant selection tensor on site i+1."""
    A, B = sites[i], sites[i + 1]
    T = torch.einsum("amsk,mrtq->arstkq", A, B)
    cl, cr, d = T.shape[0], T.shape[1], T.shape[2]
    T = T.reshape(cl, cr, d, d, -1).detach().clone().requires_grad_(True)
    nk = T.shape[4]
    left = T.permute(0, 1, 3, 4, 2).reshape(cl, cr * d * nk, d, 1)           # [a, (r, s', k), s]
    sel = torch.eye(cr * d * nk, dtype=T.dtype).reshape(cr * d * nk, cr, d, nk)  # [(r, s', k), r, s', k]
    new = list(sites)
    new[i], new[i + 1] = left, sel
    new_cls = i + 1 if nk > 1 else cls
    arrays = [a if j == new_cls else a[..., 0] for j, a in enumerate(new)]
    loss = loss_fn(prob, x, arrays, targets)
    (g,) = torch.autograd.grad(loss, T)
    return loss.detach(), T.detach(), g


def sweeps_train(prob: Problem, x, arrays, targets=None, *, epochs=1, lr=1e-2, two_way=True, optimizer="sgd", clip=None,
                 b1=0.9, b2=0.999, eps=1e-8):
    """One full-batch training run with the Sweeps strategy.  Returns (site arrays, [loss per epoch], [loss per pair step])."""
    sites, cls = canon_mps_sites([a.clone() for a in arrays])
    L = len(sites)
    pairs = [(i, i + 1) for i in range(L - 1)]
    if two_way:
        pairs += [(i, i - 1) for i in reversed(range(1, L))]

    def canonicalize(lo, hi):
        for j in range(lo):
            _qr_left(sites, j)
        for j in range(L - 1, hi, -1):
            _qr_right(sites, j)

    def split(i, T, keep):
        cl, cr, d, _, nk = T.shape
        na = sites[i].shape[3]
        if na > 1:
            M = T.permute(0, 2, 4, 1, 3).reshape(cl * d * na, cr * d)
        else:
            M = T.permute(0, 2, 1, 3, 4).reshape(cl * d, cr * d * nk)
        U, S, Vh = torch.linalg.svd(M, full_matrices=False)
        k = min(keep, S.numel())
        rs = torch.sqrt(S[:k])
        A = torch.zeros(cl, keep, d, na, dtype=T.dtype)
        Bn = torch.zeros(keep, cr, d, sites[i + 1].shape[3], dtype=T.dtype)
        A[:, :k] = (U[:, :k] * rs).reshape(cl, d, na, k).permute(0, 3, 1, 2)
        Bn[:k] = (rs[:, None] * Vh[:k]).reshape(k, cr, d, sites[i + 1].shape[3])
        sites[i], sites[i + 1] = A, Bn

    state = {}
    for pr in pairs:  # model.py:753-773: the initialisation loop contracts and splits every pair once
        i = min(pr)
        canonicalize(i, i + 1)
        keep = sites[i].shape[1]
        T = torch.einsum("amsk,mrtq->arstkq", sites[i], sites[i + 1])
        T = T.reshape(T.shape[0], T.shape[1], T.shape[2], T.shape[3], -1)
        state[pr] = {"count": 0, "mu": torch.zeros_like(T), "nu": torch.zeros_like(T)}
        split(i, T, keep)
    history, steps = [], []
    for _ in range(epochs):
        acc = 0.0
        for pr in pairs:
            i = min(pr)
            canonicalize(i, i + 1)
            keep = sites[i].shape[1]
            loss, T, g = pair_loss_and_grad(prob, x, sites, cls, i, targets)
            if clip:
                g = g * min(1.0, clip / (float(g.norm()) + 1e-6))
            if optimizer == "sgd":
                T = T - lr * g
            else:  # optax.adam
                st = state[pr]
                st["count"] += 1
                t = st["count"]
                st["mu"] = b1 * st["mu"] + (1 - b1) * g
                st["nu"] = b2 * st["nu"] + (1 - b2) * g * g
                T = T - lr * (st["mu"] / (1 - b1**t)) / (torch.sqrt(st["nu"] / (1 - b2**t)) + eps)
            split(i, T, keep)
            steps.append(float(loss))
            acc += float(loss)
        history.append(acc / len(pairs))
    return _as_arrays(sites, cls), history, steps
