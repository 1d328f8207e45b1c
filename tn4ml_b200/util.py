"""Host utilities the training seam depends on (reference: tn4ml/util.py)."""
from __future__ import annotations

import copy
from enum import IntEnum

import numpy as np


class TrainingType(IntEnum):
    """tn4ml/util.py:368-373."""

    UNSUPERVISED = 0
    SUPERVISED = 1
    TARGET_TN = 2


def integer_to_one_hot(labels, num_classes=None):
    """tn4ml/util.py:175-199."""
    labels = np.asarray(labels).astype(np.int64).reshape(-1)
    if num_classes is None:
        num_classes = int(labels.max()) + 1
    out = np.zeros((labels.size, num_classes), dtype=np.float64)
    out[np.arange(labels.size), labels] = 1.0
    return out


class EarlyStopping:
    """Stop training when the monitored metric stops improving (tn4ml/util.py:376-469)."""

    def __init__(self, monitor="loss", min_delta=0.0, patience=0, mode="min"):
        if mode not in ("min", "max"):
            raise ValueError("mode must be 'min' or 'max'")
        self.monitor, self.min_delta, self.patience, self.mode = monitor, abs(min_delta), patience, mode
        self.memory = {}

    def on_begin_train(self, history, model):
        if self.monitor not in history and self.monitor != "loss":
            raise ValueError(f"metric '{self.monitor}' is not recorded in history")
        self.memory = {"best": np.inf if self.mode == "min" else -np.inf, "wait": 0, "best_epoch": 0,
                       "best_model": None}

    def on_end_epoch(self, current, epoch, model):
        better = (current < self.memory["best"] - self.min_delta) if self.mode == "min" else \
                 (current > self.memory["best"] + self.min_delta)
        if better or self.memory["best_model"] is None:
            self.memory.update(best=current, wait=0, best_epoch=epoch, best_model=model.copy(deep=True))
            return 0
        self.memory["wait"] += 1
        return 1 if self.memory["wait"] >= self.patience else 0
