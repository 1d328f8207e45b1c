#!/usr/bin/env python
"""Per-kernel times of the float64 (SIMT) family at chi = 32 / 64 / 128, B = 16384 (measurement aid)."""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for chi in (sys.argv[1:] or ["32", "64", "128"]):
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--dtype", "f64", "--chi", chi, "--batch", "16384",
                          "--steps", "3", "--warmup", "2", "--no-cpu-baseline", "--no-chi-sweep"], capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print("FAILED", out.stderr[-500:])
        continue
    d = json.loads(line[-1])
    k = d["roofline"].get("kernel_ms_per_step") or d["roofline"]["kernel_ms"]
    print(f"f64 chi={chi} step={d['ms_per_step']:.2f} ms {d['value']:.0f} samples/s {d['roofline'].get('step_tflops', 0):.2f} TF/s  "
          + "  ".join(f"{n.replace('mps_', '').replace('_kernel', '')}={v:.2f}" for n, v in k.items() if v > 0.05)
          + f"  loss={d['loss']:.9f}")
