#!/usr/bin/env python
"""bench.py -- train samples/s (fwd+bwd) of the MPS classifier hot path on B200.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): MPS classifier,
L=196 (14x14 MNIST-shaped), trigonometric embedding d=2, bond chi (default 256), 10 classes at
site 97, normalised softmax cross-entropy, batch 131072 per GPU, float32 -- synthetic inputs
x ~ U(0,1], identity-plus-noise site tensors, fixed seeds.

One "step" = loss + all L site gradients for the whole batch (micro-batched inside the step),
plus the NCCL all-reduce of the packed gradient when N > 1.  Prints ONE JSON line.

  python bench.py [--gpus N --steps K --warmup W] [--chi 256] [--batch 131072] [--dtype f32|f64]
                  [--scaling weak|strong]
  python bench.py --impl reference ...     # CPU restatement of the reference path on the host cores

The product arm never imports oracle/: the synthetic-input generators live in tn4ml_b200/synthetic.py; the oracle is
executed only by the `cpu_baseline` leg, the `parity` checker (after the timed regions) and `--impl reference`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import json
import os
import re
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--L", type=int, default=196)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--batch", type=int, default=131072, help="samples per GPU per step (weak) / per step in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--precision", default="auto", choices=["auto", "simt", "tensor"])
    ap.add_argument("--microbatch", type=int, default=0)
    ap.add_argument("--cpu-batch", type=int, default=0, help="samples per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chi-sweep", action="store_true", help="skip the chi=32/128 and float64 side measurements")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="one trailing all-reduce instead of the overlapped one")
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
        sm, mx, pw, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_median": float(np.median(pw)) if pw else None}


def make_problem(args, dtype, chi=None):
    from tn4ml_b200 import synthetic as S  # data generators only (identity + noise site tensors)
    c = args.L // 2 - 1
    arrays = S.make_mps_arrays(args.L, chi or args.chi, 2, class_site=c, C=args.classes, std=1e-2, seed=42, dtype=dtype)
    return arrays, c


def make_batch(args, B, dtype, seed):
    rng = np.random.default_rng(seed)
    x = torch.from_numpy((1.0 - rng.random((B, args.L))).astype(np.float32 if dtype == torch.float32 else np.float64))
    lab = rng.integers(0, args.classes, size=B)
    t = torch.nn.functional.one_hot(torch.from_numpy(lab), args.classes).to(dtype)
    return x, t


# --------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# --------------------------------------------------------------------------------------------------
def time_cpu(args, arrays, dtype, steps, warmup, order):
    """The reference path restated on the CPU (oracle): `order` = "gemm" (L_i @ A_i, then the phi-weighted sum) or
    "ref" (what XLA receives from quimb: per-sample site matrices, then a batched vector-matrix chain)."""
    from oracle import tn_oracle as O
    torch.set_num_threads(os.cpu_count())
    fps = flops_per_sample(args.L, args.chi, 2, args.classes, args.L // 2 - 1)
    budget = 2.0e11 if order == "gemm" else 0.6e11  # ~ seconds of work per step on a few dozen cores
    B_cpu = args.cpu_batch or max(32, min(4096, int(budget / fps)))
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    x, t = make_batch(args, B_cpu, dtype, 1239)

    def step():
        loss, grads = O.value_and_grad(prob, x, arrays, t, order=order)
        return float(loss)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    return B_cpu / dt, dt, B_cpu


def cpu_both_orders(args, arrays, dtype, steps, warmup):
    """time both contraction orders, report the faster (labelled) with the other beside it."""
    res = {}
    for order in ("gemm", "ref"):
        try:
            sps, dt, B_cpu = time_cpu(args, arrays, dtype, steps, warmup, order)
            res[order] = {"samples_per_s": sps, "s_per_step": dt, "batch": B_cpu}
        except Exception as exc:  # pragma: no cover
            res[order] = {"error": str(exc)[:200]}
    ok = {k: v for k, v in res.items() if "samples_per_s" in v}
    best = max(ok, key=lambda k: ok[k]["samples_per_s"]) if ok else None
    return best, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    arrays, c = make_problem(args, dtype)
    best, res = cpu_both_orders(args, arrays, dtype, max(1, args.steps), max(1, min(args.warmup, 2)))
    cores = os.cpu_count()
    r = res[best]
    sps, dt, B_cpu = r["samples_per_s"], r["s_per_step"], r["batch"]
    line = {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd)", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"MPS classifier L={args.L} chi={args.chi} d=2 C={args.classes} at site {c}, trig embedding, "
                               f"softmax-xent (BASELINE configs[4]); CPU sample of {B_cpu} samples/step",
                   "batch_per_step": B_cpu},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{B_cpu} samples/step, torch-CPU {args.dtype}, {best}_order (the faster of the two "
                                   f"orders), {torch.get_num_threads()} threads",
                         "orders": res},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU legs
# --------------------------------------------------------------------------------------------------
def read_profile(lib):
    nmax = 16384
    names = C.create_string_buffer(48 * nmax)
    msbuf = (C.c_float * nmax)()
    n = lib.tnb_profile_read(names, msbuf, nmax)
    per = {}
    for i in range(n):
        nm = names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
        per.setdefault(nm, []).append(msbuf[i])
    return per


def ncu_traffic(kernel: str, tiles: int):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full summary (profiles/*_full_capture.csv),
    scaled from the capture's sample-tile count to this launch's.  None if no capture lists the kernel."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_full_capture.csv")))
    want = kernel.split("(")[0]
    adj = "(adj)" in kernel
    for path in reversed(files):
        try:
            lines = open(path).read().splitlines()
        except Exception:
            continue
        m = re.search(r"(\d+) sample tiles", lines[0]) if lines else None
        if not m:
            continue
        cap_tiles = int(m.group(1))
        hdr = next((ln for ln in lines if ln.startswith("kernel,")), None)
        if not hdr:
            continue
        cols = hdr.split(",")
        try:
            ir, iw = cols.index("dram__bytes_read.sum"), cols.index("dram__bytes_write.sum")
        except ValueError:
            continue
        rows = [ln for ln in lines if ln.startswith('"') and re.search(r'\b' + re.escape(want) + r'\b', ln.split('"')[1])]
        if want == "tc_chain_kernel" and len(rows) >= 2:
            rows = [rows[1 if adj else 0]]  # first launch of the step = forward sweep, second = adjoint
        if not rows:
            continue
        f = rows[0].rsplit('"', 1)[1].split(",")  # fields after the quoted kernel signature
        gb = float(f[ir]) + float(f[iw])
        return gb * 1e9 * tiles / cap_tiles, f"{os.path.basename(path)}: dram read+write at {cap_tiles} tiles x {tiles}/{cap_tiles}"
    return None, None


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written) or the profiling recipe's fallback figures."""
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pk = {}
    return {"hbm_gbs": float(pk.get("hbm_gbs", 6454.0)), "bf16_tflops_sustained": float(pk.get("bf16_tflops_sustained", 1400.0))}


def quick_point(args, chi, dtype, dev, steps=10, batch=None, microbatch=0):
    """Device-resident fwd+bwd throughput at another bond dimension / dtype (the metric's chi = 32 / 128 points and the
    float64 variant), >= 10 timed steps."""
    import tn4ml_b200 as T
    from tn4ml_b200 import metrics
    arrays, c = make_problem(args, dtype, chi)
    model = T.TensorNetwork(arrays, dtype=dtype, device=dev)
    loss_fn = lambda *a, **k: metrics.OptaxWrapper(metrics.softmax_cross_entropy)(*a, **k).mean()  # noqa: E731
    eng, _ = model._engine(T.TrigonometricEmbedding(), loss_fn, True, dtype)
    if microbatch:
        eng.max_microbatch = microbatch
    B = batch or args.batch
    xh, th = make_batch(args, B, dtype, 7)
    x, t = xh.to(dev), th.to(dev)
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
    out = {"samples_per_s": B / ms * 1e3, "ms_per_step": ms, "tflops": fl / ms / 1e9, "steps": steps, "batch": B}
    # fractions of both rooflines: tensor pipe (TF32 denominator, float32 only) and HBM -- algorithmic bytes = x +
    # params in + grads out + loss + ONE environment set written and read once; the small-bond point is bound by the second
    w = 4 if dtype == torch.float32 else 8
    nbytes = B * args.L * w + 2 * eng.n_params * w + B * w + 2 * B * args.L * chi * w
    peaks = measured_peaks()
    out["hbm_frac"] = nbytes / (ms * 1e-3) / (peaks["hbm_gbs"] * 1e9)
    if dtype == torch.float32:
        out["tensor_frac"] = fl / ms / 1e9 / (0.5 * peaks["bf16_tflops_sustained"])
    del eng, model, g, x, t
    torch.cuda.empty_cache()
    return out


def other_configs(dev):
    """Device-resident fwd+bwd throughput of the other BASELINE.json configs (the parity-test cases of the contract; quoted here
    so that the driver's record holds them too): cfg 1 (MPS classifier, chi = 10, batch 64), cfg 2 (MPS classifier, polynomial d = 3, chi = 32), cfg 3 (SMPO L = 784,
    spacing 8, chi = 32, batch 4096; float32 and the reference's float64), cfg 4 (MPS NLL, Fourier d = 8, chi = 64)."""
    import tn4ml_b200 as T
    from tn4ml_b200 import metrics as M, synthetic as S
    res = {}

    def timed(eng, model, x, t, B, flops, steps):
        g = torch.zeros(eng.n_params, dtype=model.dtype, device=dev)
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
        return {"samples_per_s": B / ms * 1e3, "ms_per_step": ms, "tflops": flops * B / ms / 1e9, "steps": steps, "batch": B}

    rng = np.random.default_rng(1)
    for name, dtype, steps in (("cfg3_smpo_f32", torch.float32, 10), ("cfg3_smpo_f64", torch.float64, 3)):
        try:
            L, Sp, chi, B = 784, 8, 32, 4096
            model = T.SpacedMatrixProductOperator(S.make_smpo_arrays(L, chi, 2, 2, Sp, std=1e-2, seed=42, dtype=dtype), spacing=Sp,
                                                  dtype=dtype, device=dev)
            eng, _ = model._engine(T.TrigonometricEmbedding(), M.LogQuadNorm, False, dtype)
            x = torch.from_numpy(1.0 - rng.random((B, L))).to(dev, dtype)
            w = (Sp - 2) * 2 * chi**3 + 2 * chi**3 * 2 + 4 * chi**3 * 2 + Sp * 2 * chi**2 * 2 + 2 * chi**2 * 2
            res[name] = timed(eng, model, x, None, B, 3 * (L // Sp) * w, steps)
            del eng, model, x
        except Exception as exc:  # pragma: no cover
            res[name] = {"error": str(exc)[:200]}
        torch.cuda.empty_cache()
    try:
        L, chi, d, B = 16, 64, 8, 65536
        model = T.MatrixProductState(S.make_mps_arrays(L, chi, d, std=1e-2, seed=42, dtype=torch.float32, squeeze_ends=True,
                                                       identity_all=True), dtype=torch.float32, device=dev)
        eng, _ = model._engine(T.FourierEmbedding(p=8), M.NegLogLikelihood, False, torch.float32)
        x = torch.from_numpy(1.0 - rng.random((B, L))).to(dev, torch.float32)
        res["cfg4_nll_fourier_f32"] = timed(eng, model, x, None, B, flops_per_sample(L, chi, d, 1, L - 1), 20)
        del eng, model, x
    except Exception as exc:  # pragma: no cover
        res["cfg4_nll_fourier_f32"] = {"error": str(exc)[:200]}
    torch.cuda.empty_cache()
    try:
        L, chi, d, C, B = 30, 32, 3, 2, 65536
        model = T.TensorNetwork(S.make_mps_arrays(L, chi, d, class_site=14, C=C, std=1e-2, seed=42, dtype=torch.float32),
                                dtype=torch.float32, device=dev)
        eng, _ = model._engine(T.PolynomialEmbedding(degree=2, n=1, include_bias=True), M.CrossEntropyWeighted([0.7, 1.6]), True,
                               torch.float32)
        x = torch.from_numpy(1.0 - rng.random((B, L))).to(dev, torch.float32)
        t = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, C, size=B)), C).to(dev, torch.float32)
        res["cfg2_poly_classifier_f32"] = timed(eng, model, x, t, B, flops_per_sample(L, chi, d, C, 14), 20)
        del eng, model, x, t
    except Exception as exc:  # pragma: no cover
        res["cfg2_poly_classifier_f32"] = {"error": str(exc)[:200]}
    torch.cuda.empty_cache()
    for name, dtype in (("cfg1_classifier_f32", torch.float32), ("cfg1_classifier_f64", torch.float64)):
        try:  # BASELINE configs[0]: L = 196, chi = 10, 10 classes, batch 64 (the reference's own CPU-runnable case; latency bound)
            L, chi, C, B = 196, 10, 10, 64
            model = T.TensorNetwork(S.make_mps_arrays(L, chi, 2, class_site=97, C=C, std=1e-2, seed=42, dtype=dtype), dtype=dtype, device=dev)
            eng, _ = model._engine(T.TrigonometricEmbedding(), lambda *a, **k: M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean(),
                                   True, dtype)
            x = torch.from_numpy(1.0 - rng.random((B, L))).to(dev, dtype)
            t = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, C, size=B)), C).to(dev, dtype)
            res[name] = timed(eng, model, x, t, B, flops_per_sample(L, chi, 2, C, 97), 50)
            del eng, model, x, t
        except Exception as exc:  # pragma: no cover
            res[name] = {"error": str(exc)[:200]}
    torch.cuda.empty_cache()
    return res


def parity_check(args, eng, model, emb, loss_fn, x_dev, t_dev, dtype, n=256):
    """Ties the perf line to correctness: the oracle (float64, CPU) on the first `n` samples of the timed batch against the
    same kernels on the same samples -- per-sample losses and every site gradient.  Runs after the timed regions."""
    from oracle import tn_oracle as O
    n = min(n, x_dev.shape[0])
    xs, ts = x_dev[:n].contiguous(), t_dev[:n].contiguous()
    arrays = [a.detach().double().cpu() for a in model.arrays]
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    t0 = time.perf_counter()
    loss_ref, grads_ref = O.value_and_grad(prob, xs.double().cpu(), arrays, ts.double().cpu())
    per_ref = O.per_sample_loss(prob, xs.double().cpu(), arrays, ts.double().cpu())
    t_oracle = time.perf_counter() - t0
    ls, g = eng.value_and_grad(model._packed, xs, ts)
    per, _, _ = eng.forward(model._packed, xs, ts)
    loss = float(ls.item()) / n
    gmax = max(float(q.abs().max()) for q in grads_ref)
    gfro = max(float(q.norm()) for q in grads_ref)
    worst_fro, worst_max = 0.0, 0.0
    for a, b in zip(model._views(g), grads_ref):
        diff = a.double().cpu() - b
        worst_fro = max(worst_fro, float(diff.norm()) / max(float(b.norm()), 1e-3 * gfro))
        worst_max = max(worst_max, float(diff.abs().max()) / gmax)
    per_err = float(((per.double().cpu() - per_ref).abs() / per_ref.abs().clamp_min(1e-30)).max())
    tol = 1e-4 if dtype == torch.float32 else 1e-10
    err = max(abs(loss - float(loss_ref)) / abs(float(loss_ref)), worst_fro, worst_max)
    out = {"samples": n, "oracle": "oracle/tn_oracle.py float64 on the first samples of the timed batch",
           "loss_rel_err": abs(loss - float(loss_ref)) / abs(float(loss_ref)), "per_sample_loss_max_rel_err": per_err,
           "grad_site_fro_rel_err_max": worst_fro, "grad_max_abs_err_over_max": worst_max, "max_rel_err": err,
           "tolerance": tol, "oracle_seconds": t_oracle}
    bar = tol
    if dtype == torch.float32 and eng.prob.precision == 1:
        # the yardstick of what float32 arithmetic itself reaches on these samples: the fp32-FMA (SIMT) family, same inputs
        import copy
        m2 = copy.copy(model)
        m2._engine_cache = {}
        m2.precision = 0
        eng2, _ = m2._engine(emb, loss_fn, True, dtype)
        _, g2 = eng2.value_and_grad(model._packed, xs, ts)
        f2 = m2_max = 0.0
        for a, b in zip(model._views(g2), grads_ref):
            diff = a.double().cpu() - b
            f2 = max(f2, float(diff.norm()) / max(float(b.norm()), 1e-3 * gfro))
            m2_max = max(m2_max, float(diff.abs().max()) / gmax)
        out["fp32_fma_family_max_rel_err"] = max(f2, m2_max)
        bar = max(tol, 4.0 * max(f2, m2_max))
        out["bar"] = "max(1e-4, 4 x the fp32-FMA family's error on the same samples): 196 sites of chi = 256 are at the edge of 1e-4 for float32 itself, and an fp16 hi/lo pair carries 22-23 significand bits against float32's 24 (measured ratios 0.7-3.3, profiles/r02_parity_table.md)"
    out["meets_1e-4"] = bool(err <= tol)
    out["ok"] = bool(err <= bar)
    return out


def run_native(args):
    import torch.distributed as dist
    import tn4ml_b200 as T
    from tn4ml_b200 import _lib, metrics
    from tn4ml_b200.distributed import OverlappedAllReduce

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
    # weak scaling: --batch samples per GPU; strong scaling: --batch samples per step in total, split over the ranks
    B = args.batch if args.scaling == "weak" else args.batch // world
    x_host, t_host = make_batch(args, B, dtype, 1239 + rank)
    x_host, t_host = x_host.pin_memory(), t_host.pin_memory()
    x_dev, t_dev = x_host.to(dev), t_host.to(dev)
    grads = torch.zeros(eng.n_params, dtype=dtype, device=dev)
    denom = B * world
    reducer = (OverlappedAllReduce(dist.group.WORLD, dev, grad_bytes=eng.n_params * grads.element_size())
               if (world > 1 and not args.no_overlap) else None)

    def step_resident():
        if reducer is not None:
            return eng.value_and_grad(model._packed, x_dev, t_dev, denom=denom, grads=grads, reducer=reducer)[0]
        loss_sum, g = eng.value_and_grad(model._packed, x_dev, t_dev, denom=denom, grads=grads)
        if world > 1:
            dist.all_reduce(g)
            dist.all_reduce(loss_sum)
        return loss_sum

    train_step, opt_state = model.create_train_step(embedding=emb, dtype=dtype)
    copy_stream = torch.cuda.Stream(dev)
    # double-buffered device inputs: the H2D of step i+1 runs on a copy stream under the kernels of step i
    bufs = [(torch.empty_like(x_dev), torch.empty_like(t_dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i % 2])
            bufs[i % 2][0].copy_(x_host, non_blocking=True)
            bufs[i % 2][1].copy_(t_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    warm = max(args.warmup, 3)
    for _ in range(warm):
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
    launches = int(lib.tnb_launch_count() - l0)
    # The same K steps again, back to back with the timed ones (no host sync in between: the GPU never idles, so these run
    # at the same sustained clocks), with CUDA events around every launch: the dominant kernel's STEADY-STATE duration.
    lib.tnb_profile(1)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(args.steps):
        step_resident()
    pe1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    per = read_profile(lib)
    lib.tnb_profile(0)
    profiled_ms_per_step = pe0.elapsed_time(pe1) / args.steps
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = B * world / (ms_per_step * 1e-3)

    tot = sum(sum(v) for v in per.values()) or 1.0
    top = max(per, key=lambda k: sum(per[k])) if per else "none"
    fps = flops_per_sample(args.L, args.chi, 2, args.classes, c)
    # share of the step's algorithmic flops carried by each kernel family (fwd sweep 1/3, adjoint 1/3, dA 1/3)
    share = {"mps_sweep_kernel(fwd)": 1 / 3, "mps_sweep_kernel(adj)": 1 / 3, "mps_da_kernel": 1 / 3,
             "tc_chain_kernel(fwd)": 1 / 3, "tc_chain_kernel(adj)": 1 / 3, "tc_da_kernel": 1 / 3}
    launches_per_step = max(len(per.get(top, [1])) // max(args.steps, 1), 1)
    top_ms = float(np.mean(per[top])) if per else float("nan")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if dtype == torch.float32:
        # timed inside a long back-to-back loop under the power cap -> the SUSTAINED measured figure
        peak = 0.5 * peaks.get("bf16_tflops_sustained", 1400.0)
        peak_note = ("0.5 x measured sustained bf16 (nominal TF32:BF16 ratio), kernel timed in the steady-state loop; of measured"
                     if peaks else "0.5 x 1.4 PFLOP/s sustained, of fallback")
    else:
        peak, peak_note = 40.0, "FP64 nominal 40 TFLOP/s, of nominal"
        try:
            dm = json.load(open(os.path.join(ROOT, "profiles", "r02_fp64_dmma_peak.json")))
            peak = float(dm["fp64_dmma_tflops_sustained"])
            peak_note = "measured sustained mma.sync.m8n8k4.f64 rate (tools/ubench/dmma_peak.cu, profiles/r02_fp64_dmma_peak.json); of measured"
        except Exception:
            pass
    k_flops = fps * B * share.get(top, 1 / 3) / launches_per_step
    achieved = k_flops / (top_ms * 1e-3) / 1e12 if top_ms == top_ms else None
    tiles = (B // launches_per_step + 127) // 128
    traffic, traffic_note = ncu_traffic(top, tiles) if dtype == torch.float32 and args.chi == 256 else (None, None)
    roofline = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_note,
                "peak_source": peak_note,
                "kernel_ms_avg": top_ms, "kernel_launches_timed": len(per.get(top, [])),
                "kernel_share_of_step": sum(per.get(top, [0])) / tot,
                "step_tflops": fps * B / (ms_per_step * 1e-3) / 1e12,
                "step_frac": fps * B / (ms_per_step * 1e-3) / 1e12 / peak,
                "kernel_ms_per_step": {k: float(np.sum(v)) / args.steps for k, v in per.items()},
                "profiled_ms_per_step": profiled_ms_per_step}

    # end to end through the public train step: pinned host inputs -> H2D (prefetched on a copy stream) -> step -> loss D2H
    state = opt_state
    for ev in freed:
        ev.record()
    prefetch(0)
    lval = float("nan")

    def step_e2e(i, state):
        cur = i % 2
        torch.cuda.current_stream(dev).wait_event(ready[cur])
        prefetch(i + 1)
        _, state, loss = train_step(model.arrays, state, bufs[cur])
        freed[cur].record()
        return state, float(loss.item())

    k = 0
    for _ in range(2):
        state, _ = step_e2e(k, state)
        k += 1
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        state, lval = step_e2e(k, state)
        k += 1
    e1.record()
    sync_all()
    ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = B * world / (float(ems.item()) / args.steps * 1e-3)
    h2d = x_host.numel() * x_host.element_size() + t_host.numel() * t_host.element_size()

    parity = None
    if rank == 0 and not args.no_parity:
        try:
            # the e2e loop moved the parameters (Adam): the check uses the model as it is now, on the timed batch's samples
            parity = parity_check(args, eng, model, emb, loss_fn, x_dev, t_dev, dtype)
        except Exception as exc:  # pragma: no cover
            parity = {"error": str(exc)[:300]}

    sweep = None
    if rank == 0 and world == 1 and not args.no_chi_sweep:
        del x_dev, t_dev, grads, bufs
        eng._ws.clear()
        torch.cuda.empty_cache()
        sweep = {str(args.chi): {"samples_per_s": value, "ms_per_step": ms_per_step,
                                 "tflops": fps * B / (ms_per_step * 1e-3) / 1e12, "steps": args.steps, "batch": B}}
        for chi2 in (128, 64, 32):  # (the metric's chi = 32 / 128 / 256, and the smallest bond of the north star's chi >= 64 target)
            if chi2 != args.chi:
                try:
                    sweep[str(chi2)] = quick_point(args, chi2, dtype, dev)
                except Exception as exc:  # pragma: no cover
                    sweep[str(chi2)] = {"error": str(exc)[:200]}
        if dtype == torch.float32:
            # the float64 variant configs[4] names: chi = 128 at the full batch (micro-batched) and chi = 256 at B = 16384
            for chi2, b2, st in ((128, args.batch, 10), (256, 16384, 10)):
                try:
                    sweep[f"f64_chi{chi2}"] = quick_point(args, chi2, torch.float64, dev, steps=st, batch=b2)
                except Exception as exc:  # pragma: no cover
                    sweep[f"f64_chi{chi2}"] = {"error": str(exc)[:200]}

    others = None
    if rank == 0 and world == 1 and not args.no_chi_sweep:
        try:
            others = other_configs(dev)
        except Exception as exc:  # pragma: no cover
            others = {"error": str(exc)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            best, res = cpu_both_orders(args, [a.cpu() for a in arrays], dtype, 3, 1)
            r = res[best]
            cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{r['batch']} samples/step x 3 steps, torch-CPU {args.dtype}, {best}_order (the faster of the "
                             f"two orders), {torch.get_num_threads()} threads",
                   "orders": res}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": "train samples/sec (fwd+bwd)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"MPS classifier L={args.L} chi={args.chi} d=2 C={args.classes} at site {c}, "
                                   f"trig embedding, softmax-xent (BASELINE configs[4])",
                       "batch_per_gpu": B, "global_batch": B * world, "microbatch": eng.microbatch(B, True),
                       "precision": "tensor" if precision == _lib.PREC_TENSOR else "simt",
                       "parallelism": f"dp{world}", "l2": "working set (env stash) >> L2",
                       "allreduce": ("overlapped per site range" if reducer is not None else "one trailing ncclAllReduce") if world > 1 else None},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "includes": "H2D of x and targets from pinned memory (double-buffered on a copy stream), value_and_grad, "
                                "all-reduce, Adam update, loss D2H"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity, "chi_sweep": sweep, "other_configs": others,
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
