#!/usr/bin/env python
"""Per-kernel device times of one fwd+bwd step of a BASELINE config (tnb_profile): cfg_kernel_times.py <cfg> <f32|f64> [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tn4ml_b200 import _lib  # noqa: E402
from tools.bench_configs import CONFIGS, run  # noqa: E402

name, dt = sys.argv[1], sys.argv[2]
cfg = dict(CONFIGS[name])
if len(sys.argv) > 3:
    cfg["B"] = int(sys.argv[3])
dtype = torch.float32 if dt == "f32" else torch.float64
lib = _lib.lib()
run(name, cfg, dtype, steps=2, warmup=2)
lib.tnb_profile(1)
run(name, cfg, dtype, steps=3, warmup=0)
torch.cuda.synchronize()
per = bench.read_profile(lib)
lib.tnb_profile(0)
print(name, dt, "  ".join(f"{k.replace('_kernel', '')}={sum(v) / 3:.3f}" for k, v in per.items()), "(ms per step)")
