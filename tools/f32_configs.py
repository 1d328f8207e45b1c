import sys, torch
sys.path.insert(0, ".")
from tools.bench_configs import CONFIGS, run
for n in ("cfg2", "cfg4", "cfg5_chi32", "cfg5_chi128"):
    run(n, CONFIGS[n], torch.float32, steps=5, warmup=3)
