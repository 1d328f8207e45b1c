#!/usr/bin/env python
"""BASELINE configs[3] as the north star states it: MPS anomaly detection, L = 16, Fourier d = 8, chi = 64, NegLogLikelihood,
a GLOBAL batch of 65536 sharded over the ranks (strong scaling), gradient all-reduce overlapped per site range.

    python tools/scale_cfg4.py                                             (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/scale_cfg4.py

One JSON line from rank 0: samples/s of the whole job (device time, max over ranks), fwd+bwd incl. the all-reduce.
Measurement aid for profiles/ (the driver's bench contract is bench.py)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tn4ml_b200 as T  # noqa: E402
from tn4ml_b200 import metrics as M, synthetic as S  # noqa: E402
from tn4ml_b200.distributed import OverlappedAllReduce, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-overlap", action="store_true")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    L, chi, d = 16, 64, 8
    arrays = S.make_mps_arrays(L, chi, d, std=1e-2, seed=42, dtype=dtype, squeeze_ends=True, identity_all=True)
    model = T.MatrixProductState(arrays, dtype=dtype, device=dev)
    eng, _ = model._engine(T.FourierEmbedding(p=8), M.NegLogLikelihood, False, dtype)
    lo, hi = shard_bounds(args.batch, rank, world)
    rng = np.random.default_rng(1238)
    x = torch.from_numpy(1.0 - rng.random((args.batch, L)))[lo:hi].to(dev, dtype)
    grads = torch.zeros(eng.n_params, dtype=dtype, device=dev)
    reducer = (OverlappedAllReduce(dist.group.WORLD, dev, grad_bytes=eng.n_params * grads.element_size())
               if (world > 1 and not args.no_overlap) else None)

    def step():
        if reducer is not None:
            return eng.value_and_grad(model._packed, x, None, denom=args.batch, grads=grads, reducer=reducer)[0]
        loss_sum, g = eng.value_and_grad(model._packed, x, None, denom=args.batch, grads=grads)
        if world > 1:
            dist.all_reduce(g)
            dist.all_reduce(loss_sum)
        return loss_sum

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms.item())
        flops = 3 * sum(2 * (1 if i == 0 else chi) * (1 if i == L - 1 else chi) * d for i in range(L)) * args.batch
        print(json.dumps({"workload": "cfg 4: MPS NLL L=16 Fourier d=8 chi=64", "global_batch": args.batch, "n_gpus": world,
                          "scaling": "strong", "dtype": args.dtype, "ms_per_step": ms, "samples_per_s": args.batch / ms * 1e3,
                          "tflops": flops / ms / 1e9, "grad_bytes": int(eng.n_params * grads.element_size()),
                          "allreduce": "none" if world == 1 else ("trailing" if args.no_overlap else "overlapped per site range"),
                          "loss": float(loss.item()) / args.batch}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
