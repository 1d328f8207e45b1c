"""``Model``: the training seam of tn4ml (tn4ml/models/model.py) on the B200-native kernels.

Same public surface as the reference for the data-parallel path: ``configure``, ``train``,
``evaluate``, ``predict``, ``forward``, ``accuracy``, ``update_tensors``, ``arrays``,
``tensors``, ``norm``, ``normalize``, ``nparams``, ``save`` / ``load_model``.  The site
tensors are views of one packed device buffer (layout in ``include/tn4ml_b200.h``); the
``loss_fn -> value_and_grad`` step runs in ``csrc/`` through :class:`tn4ml_b200.engine.Engine`.
"""
from __future__ import annotations

import copy as _copy
import logging
import pickle
from time import time
from typing import Any, Callable, Sequence

import numpy as np
import torch

from .. import _lib, optim
from ..embeddings import Embedding, TrigonometricEmbedding, spec_of
from ..engine import Engine
from ..metrics import LossExpr, resolve
from ..strategy import Strategy, Sweeps
from ..util import EarlyStopping, TrainingType

logger = logging.getLogger(__name__)


def _torch_dtype(dtype) -> torch.dtype:
    if isinstance(dtype, torch.dtype):
        return dtype
    return torch.float64 if np.dtype(dtype) == np.float64 else torch.float32


class SiteTensor:
    """Minimal stand-in for ``quimb.tensor.Tensor``: ``.data``, ``.shape``, ``.modify(data=)``."""

    def __init__(self, model: "Model", i: int):
        self._m, self._i = model, i

    @property
    def data(self) -> torch.Tensor:
        return self._m.arrays[self._i]

    @property
    def shape(self):
        return tuple(self._m._shapes[self._i])

    def modify(self, data=None):
        if data is not None:
            self._m.arrays[self._i].copy_(torch.as_tensor(data).to(self._m._packed))


class Model:
    """Trainable tensor network whose site tensors live in one packed device buffer."""

    kind = _lib.KIND_MPS

    def __init__(self, arrays: Sequence[Any], *, n_out: int = 1, out_site: int | None = None, spacing: int = 0,
                 dtype=None, device: Any = None):
        arrays = [torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a) for a in arrays]
        if dtype is None:
            dtype = arrays[0].dtype if arrays[0].dtype in (torch.float32, torch.float64) else torch.float64
        self.dtype = _torch_dtype(dtype)
        self.L = len(arrays)
        self.n_out, self.spacing = n_out, spacing
        self.out_site = out_site if out_site is not None else self.L - 1
        self._shapes = [tuple(a.shape) for a in arrays]
        self._canon = [self._canon_index(i, s) for i, s in enumerate(self._shapes)]
        self.chi = max(max(c[0], c[1]) for c in self._canon)
        self.phys_dim = self._canon[0][2]
        self.loss: Callable | None = None
        self.strategy: Any = "global"
        self.optimizer: Any = optim.adam
        self.learning_rate = 1e-3
        self.train_type = TrainingType.UNSUPERVISED
        self.gradient_transforms = None
        self.device = ("gpu", 0)
        self._engine_cache: dict = {}
        # float32 chains run on the tcgen05 kernels wherever the shape allows (the library falls back to
        # the SIMT family otherwise); float64 always runs on the SIMT/FP64 family
        self.precision = _lib.PREC_TENSOR
        self.process_group = None  # torch.distributed group for data-parallel training
        if device is None:  # CPU tensors are storage only (layout tests); every compute call needs CUDA
            device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        self._build_packed(arrays, torch.device(device))

    # ---- layout -----------------------------------------------------------------
    def _site_nout(self, i: int) -> int:
        if self.kind == _lib.KIND_SMPO:
            return self.n_out if i % self.spacing == 0 else 1
        return self.n_out if (i == self.out_site and self.n_out > 1) else 1

    def _canon_index(self, i: int, shape):
        """(chi_l, chi_r, d, n_i, index) : how the reference-shaped site maps into its slab."""
        raise NotImplementedError

    def _site_offsets(self):
        per = self.chi * self.chi * self.phys_dim
        offs, o = [], 0
        for i in range(self.L):
            offs.append(o)
            o += per * self._site_nout(i)
        return offs, o

    def _views(self, packed: torch.Tensor):
        offs, _ = self._site_offsets()
        out = []
        for i in range(self.L):
            cl, cr, d, nj, idx = self._canon[i]
            slab = packed[offs[i]: offs[i] + self.chi * self.chi * d * nj].view(self.chi, self.chi, d, nj)
            out.append(slab[idx])
        return tuple(out)

    def _build_packed(self, arrays, dev: torch.device):
        _, total = self._site_offsets()
        self._packed = torch.zeros(total, dtype=self.dtype, device=dev)
        self._arrays = self._views(self._packed)
        for v, a in zip(self._arrays, arrays):
            if tuple(v.shape) != tuple(a.shape):
                raise ValueError(f"site shape {tuple(a.shape)} does not fit the layout {tuple(v.shape)}")
            v.copy_(a.to(dev, self.dtype))

    # ---- reference surface --------------------------------------------------------
    @property
    def arrays(self):
        return self._arrays

    @property
    def tensors(self):
        return [SiteTensor(self, i) for i in range(self.L)]

    def nparams(self) -> int:
        return int(sum(int(np.prod(s)) for s in self._shapes))  # model.py:146-153

    def update_tensors(self, params):  # model.py:467-490
        for v, a in zip(self._arrays, params):
            if v.data_ptr() != torch.as_tensor(a).data_ptr():
                v.copy_(torch.as_tensor(a).to(v))

    def copy(self, virtual: bool = False, deep: bool = False):
        new = _copy.copy(self)
        new._engine_cache = {}
        if not virtual:
            new._packed = self._packed.clone()
            new._arrays = new._views(new._packed)
        if deep and hasattr(self, "history"):
            new.history = _copy.deepcopy(self.history)
        return new

    def to(self, device):
        dev = torch.device(device)
        self._packed = self._packed.to(dev)
        self._arrays = self._views(self._packed)
        self._engine_cache = {}
        return self

    def save(self, model_name, dir_name="~", tn=False):  # model.py:105-144 (pickle of the model)
        import os
        path = os.path.join(os.path.expanduser(dir_name), f"{model_name}.pkl")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        state = self.__dict__.copy()
        state["_packed"] = self._packed.cpu()
        for k in ("_arrays", "_engine_cache", "step", "opt_state", "loss_func", "process_group"):
            state.pop(k, None)
        with open(path, "wb") as f:
            pickle.dump((self.__class__, state), f)

    # ---- model-only operations (plain device tensor algebra) -----
    def _canon4(self, i: int) -> torch.Tensor:
        return self._canon4_of(self._packed, i)

    def _canon4_of(self, packed: torch.Tensor, i: int) -> torch.Tensor:
        """Site i of a packed buffer (parameters or gradients) as its (chi_l, chi_r, d, n_i) block."""
        offs, _ = self._site_offsets()
        cl, cr, d, nj, _ = self._canon[i]
        return packed[offs[i]: offs[i] + self.chi * self.chi * d * nj].view(self.chi, self.chi, d, nj)[:cl, :cr]

    # the reference's regularisers take "norm" = Tr(P^dag P) for a SpacedMatrixProductOperator (exact type) and
    # sqrt(<A|A>) for everything else (metrics.py:99-103)
    _reg_norm_is_trace = False

    def _norm_problem(self):
        p = _lib.TnbProblem()
        p.kind, p.dtype = self.kind, (_lib.F64 if self.dtype == torch.float64 else _lib.F32)
        p.L, p.chi, p.d, p.n_out, p.out_site, p.spacing = self.L, self.chi, self.phys_dim, self.n_out, self.out_site, max(self.spacing, 1)
        return p

    def _log2_norm2(self, regs=None, grads: torch.Tensor | None = None):
        """log2 <A|A> (device float64 [1]) by the transfer-matrix kernels (tnb_norm); with `regs`, also the value of
        sum coef * reg(model) (device float64 [1]) and its gradient ADDED to `grads`."""
        import ctypes as C
        if not self._packed.is_cuda:
            # CPU tensors are STORAGE ONLY (layout tests, initialisers run before the model is moved to its GPU): the
            # initialiser's one-off normalisation is plain torch; anything on the training path needs the device
            if regs or grads is not None:
                raise RuntimeError("tn4ml_b200 runs on CUDA devices only (no CPU fallback)")
            E = torch.ones(1, 1, dtype=torch.float64)
            acc = torch.zeros((), dtype=torch.float64)
            for i in range(self.L):
                A = self._canon4(i).double()
                E = torch.einsum("lm,lrsk,mnsk->rn", E, A, A)
                m = E.abs().max()
                E, acc = E / m, acc + torch.log2(m)
            return acc + torch.log2(E[0, 0]), None
        lib = _lib.lib()
        prob = self._norm_problem()
        dev = self._packed.device
        ws = getattr(self, "_norm_ws", None)
        need = int(lib.tnb_norm_workspace_bytes(C.byref(prob)))
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = self._norm_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = torch.zeros(2, dtype=torch.float64, device=dev)
        kinds = coefs = None
        n = 0
        if regs:
            names = {"log_frob": _lib.REG_LOG_FROB, "log_pow_frob": _lib.REG_LOG_POW_FROB,
                     "log_relu_frob": _lib.REG_LOG_RELU_FROB, "quad_frob": _lib.REG_QUAD_FROB}
            n = len(regs)
            kinds = (C.c_int32 * n)(*[names[r.kind] + (4 if self._reg_norm_is_trace else 0) for r in regs])
            coefs = (C.c_double * n)(*[float(r.coef) for r in regs])
        with torch.cuda.device(dev):
            _lib.check(lib.tnb_norm(C.byref(prob), self._packed.data_ptr(), out.data_ptr(), n, kinds, coefs,
                                    None if grads is None else grads.data_ptr(), out[1:].data_ptr() if n else None,
                                    ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return out[0], out[1]

    def norm(self, **_opts) -> torch.Tensor:
        """sqrt(<A|A>) by a device transfer-matrix sweep (tn.py:66-80, smpo.py:234-248)."""
        l2, _ = self._log2_norm2()
        return torch.exp2(0.5 * l2).to(self.dtype)

    def normalize(self, insert=None):
        """tn.py:82-113 / mps.py:34-51 / smpo.py:208-232.  L <= 200: every site divided by norm^(1/L) (or one site by norm);
        L > 200: the reference's site-by-site scheme -- site i divided by its Frobenius norm, then left-canonised (QR, R into
        site i+1) -- which leaves a left-canonical chain with <A|A> = 1."""
        import ctypes as C
        if self.L > 200:
            self._qr_sweep(range(self.L - 1), left=True, normalize_sites=True)
            return
        l2, _ = self._log2_norm2()
        if not self._packed.is_cuda:  # storage-only model (see _log2_norm2)
            if insert is None:
                self._packed.mul_(float(torch.exp2(-0.5 * l2 / self.L)))
            else:
                self._arrays[insert].mul_(float(torch.exp2(-0.5 * l2)))
            return
        lib = _lib.lib()
        code = _lib.F64 if self.dtype == torch.float64 else _lib.F32
        dev = self._packed.device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            if insert is None:
                _lib.check(lib.tnb_scale(code, self._packed.numel(), self._packed.data_ptr(), l2.data_ptr(), -0.5 / self.L, st))
            else:
                if not (0 <= insert < self.L):
                    raise IndexError(f"Insert index {insert} is out of bounds for the tensor list.")  # tn.py:109-112
                offs, total = self._site_offsets()
                end = offs[insert + 1] if insert + 1 < self.L else total
                seg = self._packed[offs[insert]:end]
                _lib.check(lib.tnb_scale(code, seg.numel(), seg.data_ptr(), l2.data_ptr(), -0.5, st))

    # ---- QR sweeps: canonicalize(center) and the L > 200 normalize (model.py:879-883, tn.py:93-101) -----------------
    def _qr_sweep(self, sites, left: bool, normalize_sites: bool = False):
        """left=True: for i in sites: (optionally A_i /= ||A_i||), A_i = Q R over (chi_l * d * n_i, chi_r), A_i <- Q,
        A_{i+1} <- R A_{i+1}.  left=False: the mirror image (rows = chi_l, R into site i-1).  Householder QR (LAPACK geqrf
        convention, the one jnp.linalg.qr uses) through torch.linalg.qr on the device: plumbing, not the hot path."""
        for i in sites:
            A = self._canon4(i)                      # (cl, cr, d, n) view of the packed slab
            if normalize_sites and i > 0:
                A.div_(torch.linalg.norm(A))
            cl, cr, d, n = A.shape
            nb = i + 1 if left else i - 1
            B = self._canon4(nb)
            if left:
                M = A.permute(0, 2, 3, 1).reshape(cl * d * n, cr)
                Q, R = torch.linalg.qr(M)            # Q (rows, k), R (k, cr), k = min(rows, cr)
                k = Q.shape[1]
                if k < cr:                            # fewer rows than columns: pad back to the slab's bond
                    Q = torch.cat([Q, Q.new_zeros(Q.shape[0], cr - k)], dim=1)
                    R = torch.cat([R, R.new_zeros(cr - k, cr)], dim=0)
                A.copy_(Q.reshape(cl, d, n, cr).permute(0, 3, 1, 2))
                B.copy_(torch.einsum("ab,brsk->arsk", R, B))
            else:
                M = A.permute(1, 2, 3, 0).reshape(cr * d * n, cl)
                Q, R = torch.linalg.qr(M)
                k = Q.shape[1]
                if k < cl:
                    Q = torch.cat([Q, Q.new_zeros(Q.shape[0], cl - k)], dim=1)
                    R = torch.cat([R, R.new_zeros(cl - k, cl)], dim=0)
                A.copy_(Q.reshape(cr, d, n, cl).permute(3, 0, 1, 2))
                B.copy_(torch.einsum("lask,ba->lbsk", B, R))
        if normalize_sites:
            last = self._canon4(self.L - 1)
            last.div_(torch.linalg.norm(last))

    def canonicalize(self, where, inplace: bool = True, **_opts):
        """Mixed-canonical form around site `where` -- or around the section min(where) .. max(where) of a set of sites, as
        the Sweeps strategy asks for (strategy.py:150-154) -- (quimb's canonicalize, used by train(canonize=(True, c)),
        model.py:882-883): sites left of it left-orthogonal, sites right of it right-orthogonal; the state is unchanged.

        Inside a sweep the sites known to be canonical already (`_canon_known`, kept by the strategy's hooks) are skipped:
        a QR of an isometry returns it up to signs, so this only fixes a sign gauge the reference leaves to LAPACK."""
        lo, hi = (where, where) if isinstance(where, (int, np.integer)) else (min(where), max(where))
        if not (0 <= lo <= hi < self.L):
            raise ValueError(f"canonical center {where} out of range")
        m = self if inplace else self.copy()
        known = getattr(m, "_canon_known", None)
        first, last = (0, m.L - 1) if known is None else (min(known[0], lo), max(known[1], hi))
        m._qr_sweep(range(first, lo), left=True)
        m._qr_sweep(range(last, hi, -1), left=False)
        if known is not None:
            m._canon_known = (lo, hi)
        return m

    def _mark_modified(self, lo: int, hi: int):
        """Sites lo .. hi were rewritten: they are no longer known to be canonical."""
        known = getattr(self, "_canon_known", None)
        if known is not None:
            self._canon_known = (min(known[0], lo), max(known[1], hi))

    canonize = canonicalize

    # ---- configure (model.py:155-276) ---------------------------------------------
    def configure(self, **kwargs):
        for key, value in kwargs.items():
            if key == "strategy":  # model.py:186-214
                if isinstance(value, Strategy):
                    self.strategy = value
                elif value in ["sweeps", "local", "dmrg", "dmrg-like", "sweeps-one-way",
                               "sweeps-one-way-per-site", "sweeps-per-site"]:
                    if value == "sweeps-one-way-per-site":
                        self.strategy = Sweeps(two_way=False, grouping=1)  # (raises like the reference: grouping == 1)
                    elif value == "sweeps-per-site":
                        self.strategy = Sweeps(two_way=True, grouping=1)
                    elif value == "sweeps-one-way":
                        self.strategy = Sweeps(two_way=False, grouping=2)
                    else:
                        self.strategy = Sweeps()
                elif value in ["global", "sgd", "gd", "gradient_descent"]:
                    self.strategy = "global"
                else:
                    raise ValueError(f'Strategy "{value}" not found')
            elif key in ["optimizer", "loss", "train_type", "learning_rate", "gradient_transforms", "device",
                         "precision", "process_group"]:
                setattr(self, key, value)
            else:
                raise AttributeError(f"Attribute {key} not found")
        if self.train_type not in [TrainingType.UNSUPERVISED, TrainingType.SUPERVISED, TrainingType.TARGET_TN]:
            raise AttributeError("Specify type of training: UNSUPERVISED, SUPERVISED, or TARGET_TN!")
        if self.train_type == TrainingType.TARGET_TN:
            raise NotImplementedError("TARGET_TN has no batched contraction; it stays on the reference path")
        if self.gradient_transforms:
            self._opt = optim.chain(*self.gradient_transforms)
        elif callable(self.optimizer) and not isinstance(self.optimizer, optim.GradientTransformation):
            self._opt = self.optimizer(learning_rate=self.learning_rate)
        elif isinstance(self.optimizer, optim.GradientTransformation):
            self._opt = self.optimizer
        else:
            self._opt = optim.adam(learning_rate=self.learning_rate)
        if isinstance(self.device, str):  # README form: device='gpu'  
            self.device = (self.device, 0)
        if len(self.device) != 2 or not isinstance(self.device, tuple):
            raise AttributeError("Device must be a tuple of (str, int)!")
        if self.device[0] not in ["cpu", "gpu"]:
            raise AttributeError("Device must be 'cpu' or 'gpu'!")
        if self.device[0] == "cpu":
            raise RuntimeError("Device 'cpu' was requested but tn4ml_b200 has no CPU path; use ('gpu', i).")
        if not torch.cuda.is_available() or self.device[1] >= torch.cuda.device_count():
            raise RuntimeError(f"Device '{self.device[0]}' was requested but no such device is available.")
        self.to(torch.device("cuda", self.device[1]))
        logger.info("backend=tn4ml_b200 | requested=%s:%s", self.device[0], self.device[1])

    # ---- engine -------------------------------------------------------------------
    def _engine(self, embedding: Embedding, loss_fn, supervised: bool, dtype) -> tuple[Engine, LossExpr | None]:
        dt = _torch_dtype(dtype) if dtype is not None else self.dtype
        if dt != self.dtype:
            self._packed = self._packed.to(dt)
            self.dtype = dt
            self._arrays = self._views(self._packed)
            self._engine_cache = {}
        expr = None
        if loss_fn is not None:
            expr = resolve(loss_fn, self, self.L, embedding.dim, supervised)
        # keyed on the CONTENTS of the feature-map spec (id() of a temporary embedding is reused by CPython)
        key = (bytes(spec_of(embedding)), None if expr is None else (expr.head, expr.class_weights), dt, self.precision)
        if expr is not None and expr.semisup is not None and self.kind != _lib.KIND_SMPO:
            raise ValueError("SemiSupervisedLoss is defined for SpacedMatrixProductOperator / MatrixProductOperator models")
        eng = self._engine_cache.get(key)
        if eng is None:
            eng = Engine(kind=self.kind, L=self.L, chi=self.chi, d=self.phys_dim, n_out=self.n_out,
                         out_site=self.out_site, spacing=self.spacing, embedding=embedding, loss=expr, dtype=dt,
                         device=self._packed.device, precision=self.precision)
            self._engine_cache[key] = eng
        return eng, expr

    def _reg_value_and_grad(self, expr: LossExpr, grads: torch.Tensor | None = None):
        """Model-only regularisers (metrics.py:89-168): value (device scalar) from the transfer-matrix kernels; the gradient is
        ADDED to `grads` in place by the same call (tnb_norm).  Batch independent."""
        if not expr or not expr.regs:
            return None
        _, val = self._log2_norm2(expr.regs, grads)
        return val

    def create_train_step(self, params=None, loss_func=None, *, embedding=None, dtype=None):
        """model.py:529-618: returns (train_step, opt_state).  ``train_step(params, opt_state, data,
        grad_clip_threshold)`` -> (params, opt_state, loss) with the fused value-and-grad inside."""
        supervised = self.train_type == TrainingType.SUPERVISED
        eng, expr = self._engine(embedding or TrigonometricEmbedding(), self.loss, supervised, dtype)
        opt_state = self._opt.init(self._packed)
        grads = torch.zeros_like(self._packed)
        group = self.process_group

        reducer = None
        if group is not None:
            from ..distributed import OverlappedAllReduce
            reducer = OverlappedAllReduce(group, self._packed.device,
                                          grad_bytes=self._packed.numel() * self._packed.element_size())
        n_cnt = torch.zeros(1, dtype=torch.float64, device=self._packed.device)

        semisup = expr.semisup if expr is not None else None
        if semisup is not None and group is not None:
            raise NotImplementedError("SemiSupervisedLoss is single-device here")

        def train_step(params, opt_state, data=None, grad_clip_threshold=None):
            if isinstance(data, tuple) and len(data) == 2:
                x, t = data
            else:
                x, t = data, None
            n_local = x.shape[0]
            if semisup is not None:
                # metrics.py:209-233: the per-sample value depends on the model-only term too, so the regulariser's gradient
                # enters with the batch sum of df/dm as its coefficient (one host read of that sum per step)
                from ..metrics import RegExpr
                _, reg = self._log2_norm2([RegExpr("log_relu_frob", semisup)])
                grads.zero_()
                loss_sum, g, wsum = eng.semisup_value_and_grad(self._packed, x, t, reg, denom=n_local, grads=grads)
                self._log2_norm2([RegExpr("log_relu_frob", semisup * float(wsum.item()) / n_local)], grads=g)
                loss = loss_sum / n_local
                gscale = 1.0
            elif group is None:
                loss_sum, g = eng.value_and_grad(self._packed, x, t, denom=n_local, grads=grads)
                loss = loss_sum / n_local
                gscale = 1.0
            else:
                # K6: shards may differ in size (distributed.shard_bounds), so the gradient of the loss SUM is reduced and
                # the division by the true global sample count happens on the device, after the all-reduce of the counts
                import torch.distributed as dist
                n_cnt.fill_(float(n_local))
                dist.all_reduce(n_cnt, op=dist.ReduceOp.SUM, group=group)
                loss_sum, g = eng.value_and_grad(self._packed, x, t, denom=1, grads=grads, reducer=reducer)
                loss = loss_sum / n_cnt
                gscale = (1.0 / n_cnt).to(self._packed.dtype)
            fused = (isinstance(self._opt, optim._Adam) and g.is_cuda and g.dtype == self._packed.dtype
                     and g.is_contiguous())
            has_reg = bool(expr and expr.regs)
            if has_reg or grad_clip_threshold or not fused:
                if torch.is_tensor(gscale):
                    g.mul_(gscale)
                    gscale = 1.0
            if has_reg:
                loss = loss + self._reg_value_and_grad(expr, g)  # value; gradient added to g by the kernel
            if grad_clip_threshold:  # util.gradient_clip (util.py:131-155): one factor per SITE tensor, eps 1e-6
                eng.gradient_clip(g, grad_clip_threshold)
            if fused:  # optimizer + update_tensors of train_step (model.py:591-614) as one kernel on the packed buffer
                opt_state = self._opt.fused_step(self._packed, g, opt_state, grad_scale=gscale)
                return self.arrays, opt_state, loss
            updates, opt_state = self._opt.update(g, opt_state, self._packed)
            optim.apply_updates(self._packed, updates.to(self._packed.dtype))
            return self.arrays, opt_state, loss

        return train_step, opt_state

    # ---- Sweeps strategy (strategy.py:51-258; model.py:753-773, 829-866) -------------------------------------------
    def create_pair_train_step(self, *, embedding=None, dtype=None):
        """train_step(T, opt_state, data, sites, grad_clip_threshold) -> (T, opt_state, loss) for the merged tensor T of one
        pair of sites: loss and dLoss/dT from the environment kernels (Engine.pair_value_and_grad), util.gradient_clip on the
        one tensor, the optimizer's update applied to T in place.  What create_train_step(params=params_i, ...) returns in the
        reference when the strategy is Sweeps."""
        supervised = self.train_type == TrainingType.SUPERVISED
        if self.process_group is not None:
            raise NotImplementedError("Sweeps is single-device (as in the reference)")
        prec = self.precision
        self.precision = _lib.PREC_SIMT  # tnb_pair_grad reads the SIMT family's environment stash
        try:
            eng, expr = self._engine(embedding or TrigonometricEmbedding(), self.loss, supervised, dtype)
        finally:
            self.precision = prec
        if expr is not None and expr.regs:
            raise NotImplementedError("model-only regularisers inside a Sweeps step are not built")

        def train_step(T, opt_state, data=None, sites=None, grad_clip_threshold=None):
            if isinstance(data, tuple) and len(data) == 2:
                x, t = data
            else:
                x, t = data, None
            i = min(sites)
            loss_sum, dT = eng.pair_value_and_grad(self._packed, x, t, i, denom=x.shape[0])
            g = dT[:T.shape[0], :T.shape[1]]
            if grad_clip_threshold:  # util.gradient_clip on the single tensor (util.py:131-155)
                g = g * torch.clamp(grad_clip_threshold / (torch.linalg.norm(g) + 1e-6), max=1.0)
            if opt_state is None or opt_state_shape(opt_state) != tuple(T.shape):
                opt_state = self._opt.init(T)  # (model.py:596-604: the moments are re-initialised when the shape changed)
            updates, opt_state = self._opt.update(g.contiguous(), opt_state, T)
            T.add_(updates.to(T.dtype))
            return T, opt_state, loss_sum / x.shape[0]

        def opt_state_shape(state):
            for v in (state.values() if isinstance(state, dict) else ()):
                if torch.is_tensor(v):
                    return tuple(v.shape)
            return None

        return train_step

    # ---- train (model.py:620-976) ----------------------------------------------------
    def train(self, *args, **kwargs):
        """model.py:620-976 (same arguments as the reference's train)."""
        try:
            return self._train(*args, **kwargs)
        finally:
            self._canon_known = None  # (the canonical-section bookkeeping of a sweep is only valid inside it)

    def _train(self, inputs=None, val_inputs=None, targets=None, val_targets=None, tn_target=None, batch_size=None,
               epochs=1, embedding=None, normalize=False, canonize=(False, None), time_limit=None, earlystop=None,
               gradient_clip_threshold=None, val_batch_size=None, eval_metric=None, display_val_acc=False,
               accuracy_fn=None, dtype=None, shuffle=False, seed=42, alternate_flip=False):
        if embedding is None:
            embedding = TrigonometricEmbedding()
        sweeps = isinstance(self.strategy, Sweeps)
        if not sweeps and self.strategy != "global":
            raise ValueError("Only Global Gradient Descent and DMRG Sweeping strategy is supported for now!")
        if not hasattr(self, "_opt"):
            raise AttributeError("Provide 'optimizer' or sequence of 'gradient_transforms'! ")
        inputs = np.asarray(inputs)
        if batch_size is None:
            batch_size = len(inputs)
        num_batches = max(1, len(inputs) // batch_size)
        if targets is not None:
            targets = np.asarray(targets)
            if targets.ndim == 1:
                targets = np.expand_dims(targets, axis=-1)
        if val_inputs is not None and eval_metric is None:
            eval_metric = self.loss
        self.batch_size = batch_size
        n_batches = max(1, len(inputs) // self.batch_size)
        if not hasattr(self, "history"):
            self.history = {"loss": [], "epoch_time": [], "unfinished": False}
            if val_inputs is not None:
                if val_batch_size is None:
                    raise ValueError("Validation batch size must be provided!")
                self.history["val_loss"] = []
                if display_val_acc:
                    self.history["val_acc"] = []
        return_value = 0
        if earlystop:
            earlystop.on_begin_train(self.history, self)
        if sweeps:
            # model.py:753-773: one optimizer state per pair and direction; the reference's initialisation loop contracts and
            # splits every pair once (a change of gauge: the splits are exact)
            self.step = self.create_pair_train_step(embedding=embedding, dtype=dtype)
            self.opt_states = {}
            self._canon_known = None
            self.canonicalize(0)
            self._canon_known = (0, 0)
            for sites in self.strategy.iterate_sites(self):
                T = self.strategy.prehook(self, sites)
                self.opt_states[sites] = self._opt.init(T)
                self.strategy.posthook(self, sites)
        else:
            self.step, self.opt_state = self.create_train_step(embedding=embedding, dtype=dtype)
        start_train = time()
        for epoch in range(epochs):
            time_epoch = time()
            loss_batch = torch.zeros((), dtype=torch.float64, device=self._packed.device)
            for batch_data in _batch_iterator(inputs, targets, batch_size, dtype=self.dtype, shuffle=shuffle,
                                              seed=seed, alternate_flip=alternate_flip):
                dev = self._packed.device
                if isinstance(batch_data, tuple):
                    batch_data = tuple(_to_device(b, dev, self.dtype) for b in batch_data)
                else:
                    batch_data = _to_device(batch_data, dev, self.dtype)
                if sweeps:  # model.py:829-866
                    loss_curr = torch.zeros((), dtype=torch.float64, device=dev)
                    site_count = 0
                    for sites in self.strategy.iterate_sites(self):
                        site_count += 1
                        T = self.strategy.prehook(self, sites)
                        _, self.opt_states[sites], loss_group = self.step(
                            T, self.opt_states[sites], batch_data, sites=sites, grad_clip_threshold=gradient_clip_threshold)
                        self.strategy.posthook(self, sites)
                        loss_curr = loss_curr + loss_group.reshape(())
                    loss_curr = loss_curr / site_count
                else:
                    _, self.opt_state, loss_curr = self.step(self.arrays, self.opt_state, batch_data,
                                                             grad_clip_threshold=gradient_clip_threshold)
                loss_batch = loss_batch + loss_curr.reshape(())
                if normalize:
                    self.normalize()
                    self._canon_known = None if getattr(self, "_canon_known", None) is None else (0, self.L - 1)
                if canonize and canonize[0]:  # model.py:882-883
                    self.canonicalize(canonize[1], inplace=True)
            loss_epoch = float((loss_batch / n_batches).item())  # the one D2H sync per epoch (model.py:887)
            self.history["loss"].append(loss_epoch)
            self.history["epoch_time"].append(time() - time_epoch)
            if time_limit is not None and (time() - start_train + np.mean(self.history["epoch_time"]) >= time_limit):
                self.history["unfinished"] = True
                return self.history
            if val_inputs is not None:
                loss_val = self.evaluate(val_inputs, val_targets, batch_size=val_batch_size, embedding=embedding,
                                         evaluate_type=self.train_type, metric=eval_metric, dtype=dtype,
                                         shuffle=shuffle, seed=seed, alternate_flip=alternate_flip)
                self.history["val_loss"].append(loss_val)
                if display_val_acc:
                    self.history["val_acc"].append(self.accuracy(
                        val_inputs, val_targets, batch_size=val_batch_size, embedding=embedding, shuffle=shuffle,
                        accuracy_fn=accuracy_fn, dtype=dtype, seed=seed, alternate_flip=alternate_flip))
                if earlystop and earlystop.monitor == "val_loss":
                    return_value = earlystop.on_end_epoch(loss_val, epoch, self)
            elif earlystop:
                current = loss_epoch if earlystop.monitor == "loss" else \
                    sum(self.history[earlystop.monitor][-num_batches:]) / num_batches
                return_value = earlystop.on_end_epoch(current, epoch, self)
            if earlystop and return_value == 1:
                return earlystop.memory["best_model"].history
        return self.history

    # ---- evaluate (model.py:978-1151) --------------------------------------------------
    def evaluate(self, inputs=None, targets=None, tn_target=None, batch_size=None, embedding=None,
                 evaluate_type: int = TrainingType.UNSUPERVISED, return_list=False, metric=None, dtype=None,
                 shuffle=False, seed=42, alternate_flip=False):
        if embedding is None:
            embedding = TrigonometricEmbedding()
        if evaluate_type not in [TrainingType.UNSUPERVISED, TrainingType.SUPERVISED, TrainingType.TARGET_TN]:
            raise ValueError("Specify type of evaluation: UNSUPERVISED, SUPERVISED, or TARGET_TN!")
        if evaluate_type == TrainingType.TARGET_TN:
            raise NotImplementedError("TARGET_TN has no batched contraction; it stays on the reference path")
        if hasattr(self, "batch_size") and batch_size is None:
            batch_size = self.batch_size
        if not hasattr(self, "batch_size"):
            self.batch_size = batch_size
        supervised = evaluate_type == TrainingType.SUPERVISED
        eng, expr = self._engine(embedding, metric or self.loss, supervised, dtype)
        inputs = np.asarray(inputs)
        if batch_size is None:
            batch_size = len(inputs)
        if targets is not None:
            targets = np.asarray(targets)
            if targets.ndim == 1:
                targets = np.expand_dims(targets, axis=-1)
        losses, loss_value, nb = [], 0.0, 0
        for batch_data in _batch_iterator(inputs, targets if supervised else None, batch_size, dtype=self.dtype,
                                          shuffle=shuffle, seed=seed, alternate_flip=alternate_flip):
            x, y = batch_data if isinstance(batch_data, tuple) else (batch_data, None)
            per, _, _ = eng.forward(self._packed, x, None if (expr is not None and expr.semisup is not None) else y)
            if expr is not None and expr.semisup is not None:  # SemiSupervisedLoss: the per-sample LogQuadNorm is post-processed
                from ..metrics import RegExpr
                _, reg = self._log2_norm2([RegExpr("log_relu_frob", expr.semisup)])
                yv = torch.as_tensor(y).to(per.device, per.dtype).reshape(-1)
                per = eng.semisup_value(per, yv, reg)[0].to(per.dtype)
            losses.append(per)
            nb += 1
        if nb == 0:
            raise ValueError("No evaluation batches were produced.")
        rv = self._reg_value_and_grad(expr)
        if return_list:
            out = torch.cat(losses)
            if rv is not None:
                out = out + rv.to(out.dtype)
            return out.cpu().numpy()
        loss_value = sum(float(l.mean().item()) for l in losses) / nb  # mean of batch means (model.py:1128-1141)
        if rv is not None:
            loss_value += float(rv)
        return float(loss_value)

    # ---- predict / forward / accuracy (model.py:278-465) ------------------------------
    def forward(self, data, embedding=None, batch_size=64, normalize=False, dtype=None, seed=42,
                alternate_flip=False):
        """Raw model outputs for a batch -> CUDA tensor [N, n_out] (or [N])."""
        if embedding is None:
            embedding = TrigonometricEmbedding()
        eng, _ = self._engine(embedding, None, False, dtype)
        data = np.asarray(data) if not torch.is_tensor(data) else data
        if data.ndim == 1:
            data = data.reshape(1, -1)
        outs = []
        for x in _batch_iterator(data, None, batch_size, dtype=self.dtype, shuffle=False, seed=seed,
                                 alternate_flip=alternate_flip):
            _, out, oexp = eng.forward(self._packed, x, None, want_loss=False)
            if normalize and out.ndim == 2:
                out = out / torch.linalg.norm(out, dim=-1, keepdim=True)  # scale-free: exponent cancels
            elif out.ndim == 2:
                out = torch.ldexp(out, oexp[:, None])
            else:
                out = torch.ldexp(out, oexp)
            outs.append(out.squeeze(-1) if out.ndim == 2 and out.shape[1] == 1 else out)
        return torch.cat(outs, dim=0)

    def predict(self, sample, embedding=None, return_tn=False, normalize=False):
        if return_tn:
            raise NotImplementedError("return_tn needs a quimb tensor network; not part of the fused path")
        sample = np.asarray(sample)
        if len(sample.flatten()) < self.L:
            raise ValueError(f"Input data must have at least {self.L} elements!")
        return self.forward(sample.reshape(1, -1), embedding=embedding, batch_size=1, normalize=normalize)[0]

    def accuracy(self, data, y_true=None, embedding=None, batch_size=64, shuffle=False, normalize=False,
                 accuracy_fn=None, dtype=None, seed=42, alternate_flip=False) -> float:
        if y_true is None:
            raise ValueError("For unsupervised learning you must provide target data!")
        y_pred = self.forward(data, embedding=embedding, batch_size=batch_size, normalize=normalize, dtype=dtype,
                              seed=seed, alternate_flip=alternate_flip)
        if accuracy_fn is not None:
            y_pred = accuracy_fn(y_pred)
        y = torch.as_tensor(np.asarray(y_true)).to(y_pred.device)
        return float((_as_class_labels(y_pred) == _as_class_labels(y)).sum().item()) / len(y)


def _as_class_labels(a: torch.Tensor) -> torch.Tensor:
    if a.ndim >= 2 and a.shape[-1] > 1:
        return a.argmax(dim=-1)
    return a.reshape(-1).round().long() if a.is_floating_point() else a.reshape(-1)


def _to_device(a, dev, dtype):
    t = torch.as_tensor(a)
    if t.device.type == "cpu":
        t = t.to(dtype).pin_memory()
    return t.to(dev, dtype, non_blocking=True)


def load_model(model_name, dir_name=None):
    """model.py:1167-1183."""
    import os
    path = f"{model_name}.pkl" if dir_name is None else os.path.join(os.path.expanduser(dir_name), f"{model_name}.pkl")
    with open(path, "rb") as f:
        cls, state = pickle.load(f)
    obj = cls.__new__(cls)
    obj.__dict__.update(state)
    obj._engine_cache = {}
    obj.process_group = None
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    obj._packed = obj._packed.to(dev)
    obj._arrays = obj._views(obj._packed)
    return obj


def _batch_iterator(x, y=None, batch_size=2, dtype=torch.float64, shuffle=True, seed=0, alternate_flip=False):
    """model.py:1201-1277: optional shuffle, fixed-size chunks (last one may be smaller),
    optional flip of every odd batch along the feature axis."""
    xt = torch.as_tensor(x)
    yt = None if y is None else torch.as_tensor(y)
    n = len(xt)
    if shuffle:
        g = torch.Generator().manual_seed(int(seed))
        perm = torch.randperm(n, generator=g)
        xt = xt[perm.to(xt.device)]
        if yt is not None:
            yt = yt[perm.to(yt.device)]
    for bi, s in enumerate(range(0, n, batch_size)):
        xc = xt[s: s + batch_size]
        if alternate_flip and bi % 2 == 1:
            xc = torch.flip(xc, dims=[1])
        if yt is not None:
            yield xc, yt[s: s + batch_size]
        else:
            yield xc
