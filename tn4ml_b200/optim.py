"""Optimisers with optax's update rules, operating on the packed parameter buffer.

The reference runs eager optax over an L-leaf pytree (tn4ml/models/model.py:591-614); the
site tensors here are views of one packed device buffer, so one fused update covers all
sites.  (optax itself is not installable in this environment; semantics follow
``optax.adam`` / ``optax.sgd`` / ``optax.clip_by_global_norm`` / ``optax.chain``.)
"""
from __future__ import annotations

import torch


class GradientTransformation:
    def init(self, params: torch.Tensor):
        return None

    def update(self, grads: torch.Tensor, state, params=None):
        raise NotImplementedError


class _Adam(GradientTransformation):
    def __init__(self, learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
        self.lr, self.b1, self.b2, self.eps, self.eps_root = learning_rate, b1, b2, eps, eps_root

    def init(self, params):
        return {"count": 0, "mu": torch.zeros_like(params), "nu": torch.zeros_like(params)}

    def update(self, grads, state, params=None):
        state["count"] += 1
        t = state["count"]
        state["mu"].mul_(self.b1).add_(grads, alpha=1 - self.b1)
        state["nu"].mul_(self.b2).addcmul_(grads, grads, value=1 - self.b2)
        mu_hat = state["mu"] / (1 - self.b1**t)
        nu_hat = state["nu"] / (1 - self.b2**t)
        lr = self.lr(t - 1) if callable(self.lr) else self.lr
        return -lr * mu_hat / (torch.sqrt(nu_hat + self.eps_root) + self.eps), state

    def fused_step(self, params, grads, state, grad_scale=1.0):
        """update() + apply_updates() as ONE kernel over the packed CUDA buffers (tnb_adam_update).  ``grad_scale`` may be
        a Python float or a one-element device tensor of the parameters' dtype (read by the kernel: no host round trip)."""
        from . import _lib
        state["count"] += 1
        t = state["count"]
        lr = self.lr(t - 1) if callable(self.lr) else self.lr
        code = _lib.F64 if params.dtype == torch.float64 else _lib.F32
        gs_dev = None
        if torch.is_tensor(grad_scale):
            gs_dev = grad_scale.reshape(1).to(params.dtype)
            grad_scale = 1.0
        with torch.cuda.device(params.device):
            _lib.check(_lib.lib().tnb_adam_update_dev(
                code, params.numel(), params.data_ptr(), grads.data_ptr(), state["mu"].data_ptr(), state["nu"].data_ptr(),
                float(lr), self.b1, self.b2, self.eps, self.eps_root, t, float(grad_scale),
                None if gs_dev is None else gs_dev.data_ptr(), torch.cuda.current_stream(params.device).cuda_stream))
        return state


class _Sgd(GradientTransformation):
    def __init__(self, learning_rate, momentum=None):
        self.lr, self.momentum = learning_rate, momentum

    def init(self, params):
        return {"count": 0, "trace": torch.zeros_like(params) if self.momentum else None}

    def update(self, grads, state, params=None):
        state["count"] += 1
        lr = self.lr(state["count"] - 1) if callable(self.lr) else self.lr
        if self.momentum:
            state["trace"].mul_(self.momentum).add_(grads)
            grads = state["trace"]
        return -lr * grads, state


class _ClipByGlobalNorm(GradientTransformation):
    def __init__(self, max_norm):
        self.max_norm = max_norm

    def update(self, grads, state, params=None):
        g = torch.linalg.norm(grads)
        return grads * torch.clamp(self.max_norm / (g + 1e-16), max=1.0), state


class _Chain(GradientTransformation):
    def __init__(self, *ts):
        self.ts = ts

    def init(self, params):
        return [t.init(params) for t in self.ts]

    def update(self, grads, state, params=None):
        new = []
        for t, s in zip(self.ts, state):
            grads, s = t.update(grads, s, params)
            new.append(s)
        return grads, new


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    return _Adam(learning_rate, b1, b2, eps, eps_root)


def sgd(learning_rate, momentum=None):
    return _Sgd(learning_rate, momentum)


def clip_by_global_norm(max_norm):
    return _ClipByGlobalNorm(max_norm)


def chain(*transforms):
    return _Chain(*transforms)


def exponential_decay(init_value, transition_steps, decay_rate, transition_begin=0, staircase=False, end_value=None):
    """optax.exponential_decay schedule."""
    def schedule(count):
        c = max(count - transition_begin, 0)
        p = c / transition_steps
        if staircase:
            p = float(int(p))
        v = init_value * (decay_rate**p) if count > transition_begin else init_value
        if end_value is not None:
            v = max(v, end_value) if decay_rate < 1 else min(v, end_value)
        return v
    return schedule


def apply_updates(params: torch.Tensor, updates: torch.Tensor) -> torch.Tensor:
    return params.add_(updates)
