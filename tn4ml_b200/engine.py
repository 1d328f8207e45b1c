"""Host side of the drop-in seam: ``loss_fn`` / ``value_and_grad`` on the C-ABI.

Replaces, for one (model, embedding, loss) triple,
    loss_fn(data, targets, *params)            tn4ml/models/model.py:723-751
    value_and_grad(params, data, targets)      tn4ml/models/model.py:554-566
    evaluate()'s per-sample closure            tn4ml/models/model.py:1048-1076
    predict()/forward()                        tn4ml/models/model.py:278-368
PyTorch is used for device memory and streams only; every number is produced by the
kernels in ``csrc/`` through ``tnb_forward`` / ``tnb_backward``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .embeddings import Embedding, spec_of
from .metrics import LossExpr


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float64:
        return _lib.F64
    if dt == torch.float32:
        return _lib.F32
    raise TypeError(f"dtype {dt} is not supported (float32 / float64)")


class Engine:
    """One contraction problem bound to a CUDA device."""

    def __init__(self, *, kind: int, L: int, chi: int, d: int, n_out: int, out_site: int, spacing: int,
                 embedding: Embedding, loss: LossExpr | None, dtype: torch.dtype, device: torch.device,
                 precision: int = _lib.PREC_SIMT, max_microbatch: int | None = None):
        if device.type != "cuda":
            raise RuntimeError("tn4ml_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.dtype, self.device = dtype, device
        _lib.check(self.lib.tnb_init(device.index if device.index is not None else torch.cuda.current_device()))
        p = _lib.TnbProblem()
        p.kind, p.dtype, p.precision = kind, _dtype_code(dtype), precision
        p.L, p.chi, p.d, p.n_out, p.out_site, p.spacing = L, chi, d, n_out, out_site, max(spacing, 1)
        head = loss.head if loss is not None else (_lib.HEAD_TSN if kind == _lib.KIND_SMPO else
                                                   (_lib.HEAD_NLL if n_out == 1 else _lib.HEAD_SOFTMAX_XENT))
        p.head = head
        p.emb = spec_of(embedding)
        if loss is not None and loss.class_weights is not None:
            for i, v in enumerate(loss.class_weights):
                p.class_weights[i] = v
        self.prob = p
        self.n_params = int(self.lib.tnb_param_count(C.byref(p)))
        if self.lib.tnb_workspace_bytes(C.byref(p), 1, 0) < 0:
            msg = self.lib.tnb_last_error().decode()
            if "physical dimension" in msg:
                raise AssertionError(msg)  # metrics.py:34
            raise ValueError("tn4ml_b200: " + msg)
        self.max_microbatch = max_microbatch
        self._ws = {}
        self._mb = {}
        self._loss_sum = torch.zeros(1, dtype=torch.float64, device=device)

    # -- workspace ----------------------------------------------------------------
    def _workspace(self, batch: int, need_grad: bool) -> torch.Tensor:
        key = (batch, need_grad)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = self.lib.tnb_workspace_bytes(C.byref(self.prob), batch, int(need_grad))
            if nbytes < 0:
                raise RuntimeError("tn4ml_b200: " + self.lib.tnb_last_error().decode())
            self._ws.clear()  # keep one workspace alive at a time
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def microbatch(self, batch: int, need_grad: bool) -> int:
        """Largest micro-batch whose workspace stays under ~60% of free HBM."""
        if self.max_microbatch:
            return min(batch, self.max_microbatch)
        key = (batch, need_grad)
        if key in self._mb:
            return self._mb[key]
        self._ws.clear()
        free, _ = torch.cuda.mem_get_info(self.device)
        mb = batch
        while mb > 1024:
            nbytes = self.lib.tnb_workspace_bytes(C.byref(self.prob), mb, int(need_grad))
            if nbytes <= 0.6 * free:
                break
            mb = (mb + 1) // 2
        self._mb[key] = mb
        return mb

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _prep(self, x, targets):
        x = torch.as_tensor(x)
        if x.device != self.device or x.dtype != self.dtype:
            x = x.to(self.device, self.dtype, non_blocking=True)
        x = x.contiguous()
        n_in = max(1, int(self.prob.emb.n_in))  # Embedding.input_dim: values per site (pre-embedded inputs, polynomial n > 1)
        if n_in > 1 and (x.ndim != 3 or x.shape[2] != n_in):
            raise ValueError(f"Input data must have shape [batch, sites, {n_in}] for this embedding, got {tuple(x.shape)}")
        if x.ndim != (3 if n_in > 1 else 2) or x.shape[1] < self.prob.L:
            raise ValueError(f"Input data must have at least {self.prob.L} elements!")  # model.py:306-307
        if targets is not None:
            targets = torch.as_tensor(targets)
            if targets.device != self.device or targets.dtype != self.dtype:
                targets = targets.to(self.device, self.dtype, non_blocking=True)
            targets = targets.contiguous().reshape(x.shape[0], -1)
            if targets.shape[1] != self.prob.n_out:
                raise ValueError(f"targets must have {self.prob.n_out} columns, got {targets.shape[1]}")
        return x, targets

    # -- len(model) < len(data) ------------------------------------------------------
    def _split_extra(self, x):
        """The reference's branch for a model shorter than the data (metrics.py:36-43, 312-319, 357-364, 448-455): every
        physical index of the data is contracted, and the sites the model does not reach are summed over their physical
        index.  That multiplies the contraction by T[b] = prod_{i >= L} sum_s phi_hat_i[b, s], a per-sample constant that
        does not depend on the model: for NegLogLikelihood it adds -2 ln|T[b]| to the loss and leaves the gradient alone;
        the classifier heads act on y / ||y|| and are unchanged for T > 0.  Returns (x[:, :L], -2 ln|T| or None)."""
        L = self.prob.L
        if x.shape[1] == L:
            return x, None
        if self.prob.kind == _lib.KIND_SMPO:
            raise NotImplementedError("len(model) < len(data) for SMPO/MPO heads (the factor enters the head non-linearly) "
                                      "is not built")
        xe = x[:, L:].contiguous()
        phi = torch.empty(xe.shape[:2] + (self.prob.d,), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tnb_embed(C.byref(self.prob.emb), _dtype_code(self.dtype), xe.shape[0] * xe.shape[1],
                                          xe.data_ptr(), phi.data_ptr(), self._stream()))
        # embed() normalises the product state over ALL sites of the data: the same as unit-norm sites (embeddings.py:1374-1387)
        t = phi.sum(-1) / torch.linalg.norm(phi, dim=-1)
        if self.prob.head != _lib.HEAD_NLL and bool((t.prod(dim=1) <= 0).any()):
            raise NotImplementedError("negative feature sums on the sites beyond the model flip the sign of the logits")
        return x[:, :L].contiguous(), -2.0 * torch.log(t.abs()).sum(dim=1)

    # -- forward only -------------------------------------------------------------
    def forward(self, params: torch.Tensor, x, targets=None, want_loss=True):
        """Per-sample losses [B], output mantissas [B, n_out] (or [B]) and exponents [B]."""
        x, targets = self._prep(x, targets)
        x, extra = self._split_extra(x)
        if extra is not None:
            loss, out, oexp = self.forward(params, x, targets, want_loss)
            if self.prob.head == _lib.HEAD_NLL:
                loss = loss + extra.to(loss.dtype)
                # raw output s = mantissa * 2^exp picks up T = exp(-extra / 2)
                out = out * torch.exp(-0.5 * extra).to(out.dtype)[:, None]
            return loss, out, oexp
        B = x.shape[0]
        smpo = self.prob.kind == _lib.KIND_SMPO
        loss = torch.empty(B, dtype=self.dtype, device=self.device)
        out = torch.empty((B,) if smpo else (B, self.prob.n_out), dtype=self.dtype, device=self.device)
        oexp = torch.empty(B, dtype=torch.int32, device=self.device)
        if want_loss and not smpo and self.prob.head != _lib.HEAD_NLL and targets is None:
            raise ValueError("supervised heads need targets")
        if not want_loss and targets is None and not smpo and self.prob.head != _lib.HEAD_NLL:
            targets = torch.zeros(B, self.prob.n_out, dtype=self.dtype, device=self.device)
            targets[:, 0] = 1
        mb = self.microbatch(B, False)
        with torch.cuda.device(self.device):
            for s in range(0, B, mb):
                e = min(B, s + mb)
                ws = self._workspace(e - s, False)
                t_ptr = targets[s:e].data_ptr() if targets is not None else None
                _lib.check(self.lib.tnb_forward(
                    C.byref(self.prob), e - s, x[s:e].data_ptr(), t_ptr, params.data_ptr(), loss[s:e].data_ptr(),
                    None, out[s:e].data_ptr(), oexp[s:e].data_ptr(), 0, ws.data_ptr(), ws.numel(), self._stream()))
        return loss, out, oexp

    # -- forward + backward -------------------------------------------------------
    def value_and_grad(self, params: torch.Tensor, x, targets=None, *, denom: int | None = None,
                       grads: torch.Tensor | None = None, accumulate: bool = False, reducer=None):
        """(sum_b loss_b as a device float64 scalar, packed gradients of sum_b loss_b / denom).

        ``reducer`` (:class:`tn4ml_b200.distributed.OverlappedAllReduce`): data-parallel step -- the last micro-batch's
        backward pass runs in site ranges and each range's slice of ``grads`` is all-reduced as soon as it is final, under
        the GEMMs of the next range; the loss sum is all-reduced too.  On return both are the sums over all ranks
        (in stream order)."""
        x, targets = self._prep(x, targets)
        x, extra = self._split_extra(x)
        B = x.shape[0]
        denom = B if denom is None else denom
        if grads is None:
            grads = torch.empty(self.n_params, dtype=self.dtype, device=self.device)
            accumulate = False
        self._loss_sum.zero_()
        if extra is not None and self.prob.head == _lib.HEAD_NLL:
            self._loss_sum.add_(extra.double().sum())  # (a model-independent constant: no gradient)
        mb = self.microbatch(B, True)
        with torch.cuda.device(self.device):
            for s in range(0, B, mb):
                e = min(B, s + mb)
                ws = self._workspace(e - s, True)
                t_ptr = targets[s:e].data_ptr() if targets is not None else None
                _lib.check(self.lib.tnb_forward(
                    C.byref(self.prob), e - s, x[s:e].data_ptr(), t_ptr, params.data_ptr(), None,
                    self._loss_sum.data_ptr(), None, None, 1, ws.data_ptr(), ws.numel(), self._stream()))
                acc = int(accumulate or s > 0)
                if reducer is None or e < B:
                    _lib.check(self.lib.tnb_backward(
                        C.byref(self.prob), e - s, x[s:e].data_ptr(), t_ptr, params.data_ptr(), 1.0 / denom,
                        grads.data_ptr(), acc, ws.data_ptr(), ws.numel(), self._stream()))
                    continue
                reducer.reduce(self._loss_sum)  # (every micro-batch's forward has been issued)
                nparts = reducer.nparts
                for part in range(nparts):
                    _lib.check(self.lib.tnb_backward_part(
                        C.byref(self.prob), e - s, x[s:e].data_ptr(), t_ptr, params.data_ptr(), 1.0 / denom,
                        grads.data_ptr(), acc, ws.data_ptr(), ws.numel(), self._stream(), part, nparts))
                    lo, hi = self.part_range(part, nparts)
                    if hi > lo:
                        reducer.reduce(grads[lo:hi])
                reducer.finish()
        return self._loss_sum, grads

    # -- SemiSupervisedLoss (metrics.py:209-233) on the SMPO kernels ---------------------------------------
    @staticmethod
    def semisup_value(lq: torch.Tensor, y: torch.Tensor, reg: torch.Tensor):
        """f = (y / m + (1 - y) m)^2 and df/dm for m = LogQuadNorm_b + reg (all device tensors; y in [0, 1])."""
        m = lq.double() + reg.double()
        u = y.double() / m + (1.0 - y.double()) * m
        return u * u, 2.0 * u * (-y.double() / (m * m) + (1.0 - y.double()))

    def semisup_value_and_grad(self, params: torch.Tensor, x, y, reg: torch.Tensor, *, denom: int, grads: torch.Tensor):
        """(sum_b f_b, grads += d/dparams of sum_b f_b / denom THROUGH LogQuadNorm_b, sum_b df_b/dm_b): the forward pass returns
        the per-sample LogQuadNorm, the VJP takes df_b/dm_b as per-sample weights (tnb_backward, TNB_KIND_SMPO).  The part of
        the gradient that goes through the model-only regulariser is added by the caller (tnb_norm)."""
        if self.prob.kind != _lib.KIND_SMPO or self.prob.head != _lib.HEAD_LOGQUAD:
            raise ValueError("SemiSupervisedLoss is defined for SMPO / MPO models (LogQuadNorm inside)")
        x, _ = self._prep(x, None)
        B = x.shape[0]
        y = torch.as_tensor(y).to(self.device, self.dtype).reshape(-1)
        if y.numel() != B:
            raise ValueError(f"SemiSupervisedLoss takes one target value per sample, got {tuple(y.shape)} for {B} samples")
        lq = torch.empty(B, dtype=self.dtype, device=self.device)
        loss_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        wsum = torch.zeros((), dtype=torch.float64, device=self.device)
        mb = self.microbatch(B, True)
        with torch.cuda.device(self.device):
            for s in range(0, B, mb):
                e = min(B, s + mb)
                ws = self._workspace(e - s, True)
                _lib.check(self.lib.tnb_forward(
                    C.byref(self.prob), e - s, x[s:e].data_ptr(), None, params.data_ptr(), lq[s:e].data_ptr(), None, None,
                    None, 1, ws.data_ptr(), ws.numel(), self._stream()))
                f, dfdm = self.semisup_value(lq[s:e], y[s:e], reg)
                w = dfdm.to(self.dtype).contiguous()
                _lib.check(self.lib.tnb_backward(
                    C.byref(self.prob), e - s, x[s:e].data_ptr(), w.data_ptr(), params.data_ptr(), 1.0 / denom,
                    grads.data_ptr(), 1, ws.data_ptr(), ws.numel(), self._stream()))
                loss_sum += f.sum()
                wsum += dfdm.sum()
        return loss_sum, grads, wsum

    # -- Sweeps strategy: loss and the gradient with respect to a merged pair of sites --------------------
    def pair_value_and_grad(self, params: torch.Tensor, x, targets, site: int, *, denom: int | None = None):
        """(sum_b loss_b as a device float64 scalar, dT [chi, chi, d, d, n_k]): the gradient of sum_b loss_b / denom with
        respect to T = A_site . A_site+1 (strategy.py:135-176, model.py:829-866), from the environments of one forward pass
        and one adjoint sweep (tnb_forward with the stash kept, tnb_backward_part part 0 of L, tnb_pair_grad)."""
        if self.prob.kind != _lib.KIND_MPS:
            raise NotImplementedError("Sweeps is built for matrix product states (the reference's strategy.py)")
        x, targets = self._prep(x, targets)
        x, extra = self._split_extra(x)
        B = x.shape[0]
        denom = B if denom is None else denom
        L = self.prob.L
        cls = self.prob.n_out > 1 and site in (self.prob.out_site, self.prob.out_site - 1)
        nk = self.prob.n_out if cls else 1
        chi, d = self.prob.chi, self.prob.d
        dT = torch.zeros(chi, chi, d, d, nk, dtype=self.dtype, device=self.device)
        if getattr(self, "_pair_scratch", None) is None:
            self._pair_scratch = (torch.empty(self.n_params, dtype=self.dtype, device=self.device), torch.empty_like(dT))
        scratch_g, part = self._pair_scratch
        if part.shape != dT.shape:
            part = torch.empty_like(dT)
            self._pair_scratch = (scratch_g, part)
        self._loss_sum.zero_()
        if extra is not None and self.prob.head == _lib.HEAD_NLL:
            self._loss_sum.add_(extra.double().sum())
        mb = self.microbatch(B, True)
        with torch.cuda.device(self.device):
            for s in range(0, B, mb):
                e = min(B, s + mb)
                ws = self._workspace(e - s, True)
                t_ptr = targets[s:e].data_ptr() if targets is not None else None
                _lib.check(self.lib.tnb_forward(
                    C.byref(self.prob), e - s, x[s:e].data_ptr(), t_ptr, params.data_ptr(), None,
                    self._loss_sum.data_ptr(), None, None, 1, ws.data_ptr(), ws.numel(), self._stream()))
                # the adjoint sweep (part 0 of L also forms the gradient of site 0 alone, which is discarded)
                _lib.check(self.lib.tnb_backward_part(
                    C.byref(self.prob), e - s, x[s:e].data_ptr(), t_ptr, params.data_ptr(), 1.0 / denom,
                    scratch_g.data_ptr(), 0, ws.data_ptr(), ws.numel(), self._stream(), 0, L))
                _lib.check(self.lib.tnb_pair_grad(
                    C.byref(self.prob), e - s, x[s:e].data_ptr(), t_ptr, site, part.data_ptr(), ws.data_ptr(), ws.numel(),
                    self._stream()))
                dT.add_(part)
        return self._loss_sum, dT

    def part_range(self, part: int, nparts: int) -> tuple[int, int]:
        """Element range of the packed gradient that is final after part `part` of `nparts` (tnb_part_sites)."""
        s0, s1 = C.c_int32(), C.c_int32()
        _lib.check(self.lib.tnb_part_sites(C.byref(self.prob), part, nparts, C.byref(s0), C.byref(s1)))
        return (int(self.lib.tnb_site_offset(C.byref(self.prob), s0.value)),
                int(self.lib.tnb_site_offset(C.byref(self.prob), s1.value)))

    def gradient_clip(self, grads: torch.Tensor, threshold: float, pre_scale: float = 1.0) -> torch.Tensor:
        """util.gradient_clip (tn4ml/util.py:131-155) in place on the packed gradient: one factor per site tensor."""
        if not threshold > 0:
            raise AssertionError("Threshold must be positive.")
        if getattr(self, "_clip_scratch", None) is None:
            n = int(self.lib.tnb_gradient_clip_scratch_bytes(C.byref(self.prob)))
            self._clip_scratch = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tnb_gradient_clip(C.byref(self.prob), grads.data_ptr(), float(threshold), float(pre_scale),
                                                  self._clip_scratch.data_ptr(), self._stream()))
        return grads
