#!/usr/bin/env python
"""One line per run: step time and per-kernel times of bench.py at a given chi (measurement aid).
   python tools/kernel_times.py <chi> [label]"""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--chi", sys.argv[1], "--steps", "3", "--warmup", "3",
                      "--no-cpu-baseline", "--no-chi-sweep"], capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print("FAILED", out.stderr[-500:])
    sys.exit(1)
d = json.loads(line[-1])
k = d["roofline"].get("kernel_ms_per_step") or d["roofline"]["kernel_ms"]
print(f"chi={sys.argv[1]} {sys.argv[2] if len(sys.argv) > 2 else ''} step={d['ms_per_step']:.3f} ms  "
      + "  ".join(f"{n.replace('tc_', '').replace('_kernel', '')}={v:.3f}" for n, v in k.items() if v > 0.05)
      + f"  loss={d['loss']:.6f}")
