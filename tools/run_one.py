#!/usr/bin/env python
"""Run a couple of fwd+bwd steps of one BASELINE config (for ncu captures): run_one.py <cfg> <f32|f64> [batch]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_configs import CONFIGS, run  # noqa: E402

name, dt = sys.argv[1], sys.argv[2]
cfg = dict(CONFIGS[name])
if len(sys.argv) > 3:
    cfg["B"] = int(sys.argv[3])
run(name, cfg, torch.float32 if dt == "f32" else torch.float64, steps=1, warmup=1)
