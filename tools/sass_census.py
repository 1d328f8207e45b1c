#!/usr/bin/env python
"""SASS census of the shipped library: per kernel, how many tensor-core / TMEM / bulk-copy / local-memory instructions the
binary holds (cuobjdump -sass), next to registers and stack from cuobjdump -res-usage.  Run here (no GPU needed):
    python tools/sass_census.py > profiles/<tag>_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tn4ml_b200", "libtn4ml_b200.so")
PATTERNS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "HMMA", "DMMA", "DFMA", "FFMA", "STL", "LDL", "RED", "ATOM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = (re.search(r"REG:(\d+)", line).group(1), re.search(r"STACK:(\d+)", line).group(1),
                          re.search(r"SHARED:(\d+)", line).group(1))
            cur = None
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for p in PATTERNS:
            if op.startswith(p):
                counts[cur][p] += 1
                break
        counts[cur]["total"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass / -res-usage of tn4ml_b200/libtn4ml_b200.so (sm_100a); counts are static instructions in the binary")
    print("kernel,regs,stack,static_smem,instructions," + ",".join(PATTERNS))
    for (name, c), dm in zip(counts.items(), demangle):
        r = usage.get(name, ("?", "?", "?"))
        short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("tnb::", "")
        print(f"\"{short}\",{r[0]},{r[1]},{r[2]},{c['total']}," + ",".join(str(c[p]) for p in PATTERNS))


if __name__ == "__main__":
    sys.exit(main())
