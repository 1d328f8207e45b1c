#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (read on the CPU box).

  python tools/ncu_summary.py launches gpurun_out/launches_<tag>.csv "<command>"   > profiles/<tag>_launches_summary.csv
  python tools/ncu_summary.py full gpurun_out/prof_<tag>.ncu-rep "<command>"       > profiles/<tag>_full_capture.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__shared_mem_per_block_dynamic",
]


def launches(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {cmd}")
    print("# (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)")
    print("kernel,launches,total_ns,share")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"\"{k}\",{a[0]},{a[1]:.0f},{a[1] / tot:.4f}")


def full(path, cmd):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {}
    for i, h in enumerate(hdr):
        for m in METRICS:
            if h.endswith(m) and m not in idx:
                idx[m] = i
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full --clock-control none --import-source on: {cmd}")
    print("# units: " + "; ".join(f"{m} [{units[idx[m]]}]" for m in METRICS if m in idx))
    print("kernel," + ",".join(m for m in METRICS if m in idx))
    for r in rows[2:]:
        print(f"\"{r[ki]}\"," + ",".join(r[idx[m]] for m in METRICS if m in idx))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
