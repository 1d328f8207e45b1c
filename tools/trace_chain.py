#!/usr/bin/env python
"""In-kernel timeline of the tcgen05 chain kernel (measurement aid).

Builds a second copy of the library with -DTNB_TC_TRACE (tools/libtn4ml_b200_trace.so), runs one forward+backward of
cfg 5 at a small batch and prints, for CTA (0,0) of the LAST chain launch (the adjoint), clock64() deltas per step:
    mma: a_ready seen -> chunk commits issued;  epilogue warps 0 / 15 / 6: chunk h seen, parked groups moved to the
    operand, gather done, published.
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, "tools", "libtn4ml_b200_trace.so")


def build():
    import __graft_entry__ as G
    step0 = os.environ.get("TRACE_STEP0")  # first traced step (default 20; short chains such as cfg 4: TRACE_STEP0=2)
    flags = ("-DTNB_TC_TRACE",) + ((f"-DTNB_TRACE_STEP0={int(step0)}",) if step0 else ())
    G.build(extra_flags=flags, target=SO)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
        sys.exit(0)
    import torch
    from tn4ml_b200 import _lib
    _lib.LIB_PATH = SO
    from tools.bench_configs import CONFIGS, run
    name = os.environ.get("TRACE_CFG", "cfg5")
    cfg = dict(CONFIGS[name])
    cfg["B"] = int(os.environ.get("TRACE_B", "18944"))
    fwd_only = "--fwd" in sys.argv
    run(name, cfg, torch.float32, steps=1, warmup=1)
    lib = _lib.lib()
    lib.tnb_debug_trace_read.restype = C.c_int
    n = 64 * 32
    buf = (C.c_longlong * n)()
    assert lib.tnb_debug_trace_read(buf, n) == 0
    if "--da" in sys.argv:
        # weight-gradient GEMM, CTA (0, site 1), chunks 200..231: mma = ready seen / MMAs issued; copy = stage free;
        # prod = loaded seen / transformed (warp 0).  Cycles relative to the previous chunk's "ready seen".
        prev = None
        for st in range(32):
            e = buf[st * 64 + 32:st * 64 + 40]
            if e[0] == 0:
                continue
            if prev:
                print(f"chunk {200 + st}: period={e[0] - prev} issue={e[1] - e[0]} stage_free={e[2] - e[0]} "
                      f"loaded={e[3] - e[0]} transformed={e[4] - e[0]}")
            prev = e[0]
        sys.exit(0)
    prev_pub = None
    for st in range(32):
        row = buf[st * 64:(st + 1) * 64]
        mma = row[48:64]
        t0 = mma[0]
        if t0 == 0:
            continue
        def rel(v):
            return "-" if v == 0 else str(v - t0)
        line = f"step {20 + st:3d} mma: aready=0 commits=" + ",".join(rel(v) for v in mma[1:4])
        if "--stages" in sys.argv:  # per column chunk: cycles the issuer waited for its stage / the peer's half / spent issuing
            line += " chunk[wait_full,wait_pfull,issue]=" + " ".join(
                "(" + ",".join(str(mma[4 + 3 * k + j]) for j in range(3)) + ")" for k in range(2))
        for w, name in ((0, "w0"), (1, "w15"), (2, "w6")):
            e = row[w * 16:(w + 1) * 16]
            line += f" | {name}: seen=" + ",".join(rel(v) for v in e[0:4]) + \
                    f" moved={rel(e[7])} gath={rel(e[8])} pub={rel(e[12])}"
        if prev_pub:
            line += f" | step_len={t0 - prev_pub}"
        prev_pub = t0
        print(line)
