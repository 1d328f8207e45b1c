#!/usr/bin/env python
"""bench.py -- train samples/s (fwd+bwd) of the MPS classifier hot path on B200.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): MPS classifier,
L=196 (14x14 MNIST-shaped), trigonometric embedding d=2, bond chi (default 256), 10 classes at
site 97, normalised softmax cross-entropy, batch 131072 per GPU, float32 -- synthetic inputs
x ~ U(0,1], identity-plus-noise site tensors, fixed seeds.

One "step" = loss + all L site gradients for the whole batch (micro-batched inside the step),
plus the NCCL all-reduce of the packed gradient when N > 1.  Prints ONE JSON line.

  python bench.py [--gpus N --steps K --warmup W] [--chi 256] [--batch 131072] [--dtype f32|f64]
  python bench.py --impl reference ...     # CPU restatement of the reference path on the host cores
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--L", type=int, default=196)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--batch", type=int, default=131072, help="samples per GPU per step (weak scaling)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--precision", default="auto", choices=["auto", "simt", "tensor"])
    ap.add_argument("--microbatch", type=int, default=0)
    ap.add_argument("--cpu-batch", type=int, default=0, help="samples per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chi-sweep", action="store_true", help="skip the chi=32/128 side measurements")
    return ap.parse_args()


def flops_per_sample(L, chi, d, C, c):
    """Algorithmic work of one fwd+bwd step per sample: 3 * sum_i 2 chi_l chi_r d (x C at the class site)."""
    f = 0
    for i in range(L):
        cl = 1 if i == 0 else chi
        cr = 1 if i == L - 1 else chi
        f += 2 * cl * cr * d * (C if i == c else 1)
    return 3 * f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_problem(args, dtype):
    from oracle import tn_oracle as O  # synthetic-input generators only (framework-neutral data)
    c = args.L // 2 - 1
    arrays = O.make_mps_arrays(args.L, args.chi, 2, class_site=c, C=args.classes, std=1e-2, seed=42, dtype=dtype)
    return arrays, c


def cpu_step_fn(args, arrays, B_cpu, dtype):
    """The reference path restated on the CPU (oracle, gemm_order; the faster of the two orders)."""
    from oracle import tn_oracle as O
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    x = O.make_inputs(B_cpu, args.L, seed=1239, dtype=dtype)
    t = O.make_labels(B_cpu, args.classes, dtype=dtype)

    def step():
        loss, grads = O.value_and_grad(prob, x, arrays, t, order="gemm")
        return float(loss)
    return step


def time_cpu(args, arrays, dtype, steps, warmup):
    torch.set_num_threads(os.cpu_count())
    B_cpu = args.cpu_batch or max(64, min(4096, int(2.0e11 / flops_per_sample(args.L, args.chi, 2, args.classes, args.L // 2 - 1))))
    step = cpu_step_fn(args, arrays, B_cpu, dtype)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    return B_cpu / dt, dt, B_cpu


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    arrays, c = make_problem(args, dtype)
    sps, dt, B_cpu = time_cpu(args, arrays, dtype, max(1, args.steps), max(1, min(args.warmup, 2)))
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd)", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"MPS classifier L={args.L} chi={args.chi} d=2 C={args.classes} softmax-xent "
                               f"(BASELINE configs[4]); CPU sample of {B_cpu} samples/step",
                   "batch_per_step": B_cpu},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{B_cpu} samples/step, torch-CPU {args.dtype}, gemm_order, {torch.get_num_threads()} threads"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def quick_point(args, chi, dtype, dev, steps=2):
    """Device-resident fwd+bwd throughput at another bond dimension (the metric's chi = 32 / 128 points)."""
    import copy
    import tn4ml_b200 as T
    from tn4ml_b200 import metrics
    a2 = copy.copy(args)
    a2.chi = chi
    arrays, c = make_problem(a2, dtype)
    model = T.TensorNetwork(arrays, dtype=dtype, device=dev)
    loss_fn = lambda *a, **k: metrics.OptaxWrapper(metrics.softmax_cross_entropy)(*a, **k).mean()  # noqa: E731
    eng, _ = model._engine(T.TrigonometricEmbedding(), loss_fn, True, dtype)
    rng = np.random.default_rng(7)
    B = args.batch
    x = torch.from_numpy((1.0 - rng.random((B, args.L))).astype(np.float32 if dtype == torch.float32 else np.float64)).to(dev)
    t = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, args.classes, size=B)), args.classes).to(dev, dtype)
    g = torch.zeros(eng.n_params, dtype=dtype, device=dev)
    for _ in range(3):
        eng.value_and_grad(model._packed, x, t, grads=g)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.value_and_grad(model._packed, x, t, grads=g)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    fl = flops_per_sample(args.L, chi, 2, args.classes, c) * B
    out = {"samples_per_s": B / ms * 1e3, "ms_per_step": ms, "tflops": fl / ms / 1e9}
    del eng, model, g, x, t
    torch.cuda.empty_cache()
    return out


def run_native(args):
    import torch.distributed as dist
    import tn4ml_b200 as T
    from tn4ml_b200 import _lib, metrics

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    lib = _lib.lib()

    arrays, c = make_problem(args, dtype)
    model = T.TensorNetwork(arrays, dtype=dtype, device=dev)
    precision = {"auto": None, "simt": _lib.PREC_SIMT, "tensor": _lib.PREC_TENSOR}[args.precision]
    if precision is None:
        precision = _lib.PREC_TENSOR if dtype == torch.float32 else _lib.PREC_SIMT
    loss_fn = lambda *a, **k: metrics.OptaxWrapper(metrics.softmax_cross_entropy)(*a, **k).mean()  # noqa: E731
    model.configure(optimizer=T.optim.adam, learning_rate=1e-4, loss=loss_fn, train_type=T.TrainingType.SUPERVISED,
                    device=("gpu", local), precision=precision,
                    process_group=dist.group.WORLD if world > 1 else None)
    emb = T.TrigonometricEmbedding()
    eng, _ = model._engine(emb, loss_fn, True, dtype)
    if args.microbatch:
        eng.max_microbatch = args.microbatch
    B = args.batch
    rng = np.random.default_rng(1239 + rank)
    x_host = torch.from_numpy((1.0 - rng.random((B, args.L))).astype(np.float32 if dtype == torch.float32 else np.float64)).pin_memory()
    lab = rng.integers(0, args.classes, size=B)
    t_host = torch.nn.functional.one_hot(torch.from_numpy(lab), args.classes).to(dtype).pin_memory()
    x_dev, t_dev = x_host.to(dev), t_host.to(dev)
    grads = torch.zeros(eng.n_params, dtype=dtype, device=dev)
    denom = B * world

    def step_resident():
        loss_sum, g = eng.value_and_grad(model._packed, x_dev, t_dev, denom=denom, grads=grads)
        if world > 1:
            dist.all_reduce(g)
            dist.all_reduce(loss_sum)
        return loss_sum

    train_step, opt_state = model.create_train_step(embedding=emb, dtype=dtype)

    def step_e2e(state):
        xb = x_host.to(dev, non_blocking=True)
        tb = t_host.to(dev, non_blocking=True)
        _, state, loss = train_step(model.arrays, state, (xb, tb))
        return state, float(loss.item())

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.tnb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    sync_all()
    launches = int(lib.tnb_launch_count() - l0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # per-kernel timing for the roofline of the dominant kernel (separate pass, events on the launch stream)
    lib.tnb_profile(1)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    step_resident()
    pe1.record()
    torch.cuda.synchronize(dev)
    profile_step_ms = pe0.elapsed_time(pe1)
    nmax = 8192
    names = C.create_string_buffer(48 * nmax)
    msbuf = (C.c_float * nmax)()
    n = lib.tnb_profile_read(names, msbuf, nmax)
    lib.tnb_profile(0)
    per = {}
    for i in range(n):
        nm = names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
        per.setdefault(nm, []).append(msbuf[i])
    tot = sum(sum(v) for v in per.values()) or 1.0
    top = max(per, key=lambda k: sum(per[k])) if per else "none"
    fps = flops_per_sample(args.L, args.chi, 2, args.classes, c)
    # share of the step's algorithmic flops carried by each kernel family (fwd sweep 1/3, adjoint 1/3, dA 1/3)
    share = {"mps_sweep_kernel(fwd)": 1 / 3, "mps_sweep_kernel(adj)": 1 / 3, "mps_da_kernel": 1 / 3,
             "tc_chain_kernel(fwd)": 1 / 3, "tc_chain_kernel(adj)": 1 / 3, "tc_da_kernel": 1 / 3}
    mb_launches = len(per.get(top, [1]))
    top_ms = float(np.mean(per[top])) if per else float("nan")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if dtype == torch.float32:
        peak = 0.5 * peaks.get("bf16_tflops_sustained", 1400.0)
        peak_note = ("0.5 x measured sustained bf16 (TF32:BF16 nominal ratio), of measured" if peaks else
                     "0.5 x 1.4 PFLOP/s, of fallback")
    else:
        peak, peak_note = 40.0, "FP64 nominal 40 TFLOP/s (no measured figure on file), of nominal"
    k_flops = fps * B * share.get(top, 1 / 3) / max(mb_launches, 1)
    achieved = k_flops / (top_ms * 1e-3) / 1e12 if top_ms == top_ms else None
    # DRAM traffic of the dominant kernel per launch: from the ncu --set full capture of the same problem at 296 sample
    # tiles (profiles/r01d_full_capture.csv: dram read + write of tc_chain_kernel), scaled to this launch's tile count.
    # Only for the profiled configuration; null otherwise.
    traffic, traffic_note = None, None
    ncu_gb_per_296_tiles = {"tc_chain_kernel(fwd)": 0.378655 + 7.633898, "tc_chain_kernel(adj)": 0.445621 + 7.622838}
    if (dtype == torch.float32 and args.chi == 256 and args.L == 196 and args.classes == 10 and top in ncu_gb_per_296_tiles):
        tiles = (B // max(mb_launches, 1) + 127) // 128
        traffic = ncu_gb_per_296_tiles[top] * 1e9 * tiles / 296.0
        traffic_note = "bytes per launch; ncu dram__bytes_read+write.sum at 296 tiles (profiles/r01d_full_capture.csv) x tiles/296"
    roofline = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_note,
                "peak_source": peak_note,
                "kernel_ms_avg": top_ms, "kernel_share_of_step": sum(per.get(top, [0])) / tot,
                "step_tflops": fps * B / (ms_per_step * 1e-3) / 1e12,
                "kernel_ms": {k: float(np.sum(v)) for k, v in per.items()},
                "profiled_step_ms": profile_step_ms, "profiled_kernel_sum_ms": float(tot)}

    # end to end through the public train step: pinned host inputs -> H2D -> step -> loss D2H
    state = opt_state
    for _ in range(2):
        state, _ = step_e2e(state)
    sync_all()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        state, lval = step_e2e(state)
    e1.record()
    sync_all()
    ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = B * world / (float(ems.item()) / args.steps * 1e-3)
    h2d = x_host.numel() * x_host.element_size() + t_host.numel() * t_host.element_size()

    sweep = None
    if rank == 0 and world == 1 and not args.no_chi_sweep:
        del x_dev, t_dev, grads
        eng._ws.clear()
        torch.cuda.empty_cache()
        sweep = {str(args.chi): {"samples_per_s": value, "ms_per_step": ms_per_step,
                                 "tflops": fps * B / (ms_per_step * 1e-3) / 1e12}}
        for chi2 in (128, 32):
            if chi2 != args.chi:
                try:
                    sweep[str(chi2)] = quick_point(args, chi2, dtype, dev)
                except Exception as exc:  # pragma: no cover
                    sweep[str(chi2)] = {"error": str(exc)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sps, dt, B_cpu = time_cpu(args, [a.cpu() for a in arrays], dtype, 3, 1)
            cpu = {"value": sps, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{B_cpu} samples/step x 3 steps, torch-CPU {args.dtype}, gemm_order, "
                             f"{torch.get_num_threads()} threads"}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": "train samples/sec (fwd+bwd)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"MPS classifier L={args.L} chi={args.chi} d=2 C={args.classes} at site {c}, "
                                   f"trig embedding, softmax-xent (BASELINE configs[4])",
                       "batch_per_gpu": B, "global_batch": B * world, "microbatch": eng.microbatch(B, True),
                       "precision": "tensor" if precision == _lib.PREC_TENSOR else "simt",
                       "parallelism": f"dp{world}", "l2": "working set (env stash) >> L2"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "includes": "H2D of x and targets from pinned memory, value_and_grad, all-reduce, Adam update, loss D2H"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "chi_sweep": sweep,
            "loss": lval,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
