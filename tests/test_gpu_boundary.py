"""The C-ABI boundary's own promises (include/tn4ml_b200.h), checked on the device:

  * tnb_forward + tnb_backward (+ tnb_adam_update) are CUDA-graph capturable: a captured step, replayed, reproduces the eager
    result (SURVEY 8b "Threading": no host sync, no lazy attribute setup, no environment reads on the call path);
  * tnb_pack_params / tnb_unpack_grads agree with the Python layout (ragged "noteven" bonds, squeezed boundaries);
  * a second device in the same process works (the opt-in shared-memory attributes are per device);
  * 2 GPUs: the packed gradient after the NCCL all-reduce of two batch shards equals the 1-GPU gradient of the whole batch.
"""
import ctypes as C
import os
import socket

import pytest
import torch

from oracle import tn_oracle as O
from tests.test_gpu_parity import _heads

pytestmark = pytest.mark.gpu


def _mps_case(dtype, chi, L=10, C_=4, B=300, seed=3):
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(L, chi, 2, class_site=L // 2, C=C_, std=0.05, seed=seed)
    x = O.make_inputs(B, L, seed=5).to(dtype)
    t = O.make_labels(B, C_).to(dtype)
    model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
    eng, _ = model._engine(T.TrigonometricEmbedding(), _heads()["softmax_xent"], True, dtype)
    return model, eng, x.cuda(), t.cuda()


@pytest.mark.parametrize("dtype,chi", [(torch.float32, 64), (torch.float64, 24), (torch.float32, 256), (torch.float32, 32)],
                         ids=["f32_tensor_chi64", "f64_simt", "f32_tensor_chi256", "f32_bond32_rr"])
def test_step_is_cuda_graph_capturable(dtype, chi):
    """Capture forward(keep_stash) + backward + the fused Adam update into one CUDA graph, replay it twice, and compare with
    the same two steps run eagerly (bitwise for the loss sum of step 1; gradients to accumulation-order noise)."""
    from tn4ml_b200 import _lib
    model, eng, x, t = _mps_case(dtype, chi)
    lib = _lib.lib()
    B, n = x.shape[0], eng.n_params
    p0 = model._packed.clone()

    def alloc():
        return dict(p=p0.clone(), g=torch.zeros(n, dtype=dtype, device="cuda"), mu=torch.zeros(n, dtype=dtype, device="cuda"),
                    nu=torch.zeros(n, dtype=dtype, device="cuda"), ls=torch.zeros(1, dtype=torch.float64, device="cuda"))

    ws = eng._workspace(B, True)
    code = _lib.F64 if dtype == torch.float64 else _lib.F32

    def step(s, count):
        st = torch.cuda.current_stream().cuda_stream
        s["ls"].zero_()
        _lib.check(lib.tnb_forward(C.byref(eng.prob), B, x.data_ptr(), t.data_ptr(), s["p"].data_ptr(), None,
                                   s["ls"].data_ptr(), None, None, 1, ws.data_ptr(), ws.numel(), st))
        _lib.check(lib.tnb_backward(C.byref(eng.prob), B, x.data_ptr(), t.data_ptr(), s["p"].data_ptr(), 1.0 / B,
                                    s["g"].data_ptr(), 0, ws.data_ptr(), ws.numel(), st))
        _lib.check(lib.tnb_adam_update(code, n, s["p"].data_ptr(), s["g"].data_ptr(), s["mu"].data_ptr(), s["nu"].data_ptr(),
                                       1e-3, 0.9, 0.999, 1e-8, 0.0, count, 1.0, st))

    eager = alloc()
    step(eager, 1)
    ls1 = eager["ls"].clone()
    g1 = eager["g"].clone()
    torch.cuda.synchronize()

    cap = alloc()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(cap, 1)  # warm-up on the capture stream (not part of the graph); state reset below
    torch.cuda.current_stream().wait_stream(side)
    for k, v in alloc().items():
        cap[k].copy_(v)
    graph = torch.cuda.CUDAGraph()
    l0 = lib.tnb_launch_count()
    with torch.cuda.graph(graph):
        step(cap, 1)
    assert lib.tnb_launch_count() > l0  # the library's launches went into the graph
    cap["ls"].zero_()
    graph.replay()
    torch.cuda.synchronize()
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    assert abs(float(cap["ls"]) - float(ls1)) <= tol * abs(float(ls1))
    gmax = float(g1.abs().max())
    assert float((cap["g"] - g1).abs().max()) <= tol * gmax
    assert float((cap["p"] - eager["p"]).abs().max()) <= 1e-6 * float(p0.abs().max()) + tol
    # a second replay runs the step again from the updated parameters: it must move them further
    p_after1 = cap["p"].clone()
    graph.replay()
    torch.cuda.synchronize()
    assert float((cap["p"] - p_after1).abs().max()) > 0


def test_smpo_step_is_cuda_graph_capturable():
    import tn4ml_b200 as T
    from tn4ml_b200 import _lib
    dtype = torch.float32
    arrays = O.make_smpo_arrays(32, 32, 2, 2, 8, std=0.05, seed=8)
    model = T.SpacedMatrixProductOperator(arrays, spacing=8, dtype=dtype, device="cuda:0")
    eng, _ = model._engine(T.TrigonometricEmbedding(), _heads()["logquad"], False, dtype)
    x = O.make_inputs(200, 32, seed=9).to(dtype).cuda()
    ls_e, g_e = eng.value_and_grad(model._packed, x)
    ls_e, g_e = ls_e.clone(), g_e.clone()
    torch.cuda.synchronize()
    g = torch.zeros_like(g_e)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ls_c, _ = eng.value_and_grad(model._packed, x, grads=g)
    graph.replay()
    torch.cuda.synchronize()
    assert abs(float(ls_c) - float(ls_e)) <= 1e-5 * abs(float(ls_e))
    assert float((g - g_e).abs().max()) <= 1e-5 * float(g_e.abs().max())


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_pack_params_and_unpack_grads_match_the_python_layout(dtype):
    import tn4ml_b200 as T
    from tn4ml_b200 import _lib
    L, chi, C_ = 9, 12, 3
    arrays = O.make_mps_arrays(L, chi, 3, class_site=4, C=C_, std=0.1, seed=1, shape_method="noteven")
    model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
    eng, _ = model._engine(T.PolynomialEmbedding(degree=2, n=1, include_bias=True), _heads()["softmax_xent"], True, dtype)
    lib = _lib.lib()
    sites = [a.to("cuda", dtype).contiguous() for a in arrays]
    ptrs = (C.c_void_p * L)(*[s.data_ptr() for s in sites])
    cl = (C.c_int32 * L)(*[a.shape[0] for a in arrays])
    cr = (C.c_int32 * L)(*[a.shape[1] for a in arrays])
    packed = torch.full((eng.n_params,), 7.0, dtype=dtype, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.tnb_pack_params(C.byref(eng.prob), ptrs, cl, cr, packed.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(packed, model._packed)  # same bytes as the Python-side layout, padding zero-filled
    outs = [torch.empty_like(s) for s in sites]
    optrs = (C.c_void_p * L)(*[o.data_ptr() for o in outs])
    g = torch.randn(eng.n_params, dtype=dtype, device="cuda")
    _lib.check(lib.tnb_unpack_grads(C.byref(eng.prob), g.data_ptr(), optrs, cl, cr, st))
    torch.cuda.synchronize()
    for o, v in zip(outs, model._views(g)):
        assert torch.equal(o, v)
    bad = (C.c_int32 * L)(*[chi + 1] * L)
    assert lib.tnb_pack_params(C.byref(eng.prob), ptrs, bad, cr, packed.data_ptr(), st) != 0
    assert b"do not fit" in lib.tnb_last_error()


def test_second_device_in_the_same_process():
    """the opt-in shared-memory attributes are per device; the chi=256 tensor kernels need > 48 KB."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(8, 256, 2, class_site=4, C=3, std=0.05, seed=3)
    x = O.make_inputs(130, 8, seed=5).float()
    t = O.make_labels(130, 3).float()
    res = []
    for dev in ("cuda:0", "cuda:1"):
        model = T.TensorNetwork(arrays, dtype=torch.float32, device=dev)
        eng, _ = model._engine(T.TrigonometricEmbedding(), _heads()["softmax_xent"], True, torch.float32)
        ls, g = eng.value_and_grad(model._packed, x, t)
        res.append((float(ls), g.cpu()))
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[0][0])
    assert float((res[0][1] - res[1][1]).abs().max()) <= 1e-5 * float(res[0][1].abs().max())


def _dp_worker(rank, world, port, dtype_name, chi, out):
    import torch.distributed as dist
    import tn4ml_b200 as T
    from tn4ml_b200.distributed import allreduce_step, shard_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dtype = getattr(torch, dtype_name)
    L, C_, B = 12, 4, 1000 + 37  # an odd batch: the shards differ in size
    arrays = O.make_mps_arrays(L, chi, 2, class_site=5, C=C_, std=0.05, seed=3)
    x = O.make_inputs(B, L, seed=5).to(dtype)
    t = O.make_labels(B, C_).to(dtype)
    model = T.TensorNetwork(arrays, dtype=dtype, device=f"cuda:{rank}")
    eng, _ = model._engine(T.TrigonometricEmbedding(), _heads()["softmax_xent"], True, dtype)
    lo, hi = shard_bounds(B, rank, world)
    ls, g = eng.value_and_grad(model._packed, x[lo:hi], t[lo:hi], denom=B)
    ls, g = allreduce_step(ls.clone(), g.clone())
    if rank == 0:
        ls1, g1 = eng.value_and_grad(model._packed, x, t, denom=B)  # the whole batch on one GPU
        gmax = float(g1.abs().max())
        out.put((abs(float(ls) - float(ls1)) / abs(float(ls1)), float((g - g1).abs().max()) / gmax))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,chi", [("float64", 24), ("float32", 64)], ids=["f64_simt", "f32_tensor"])
def test_two_gpu_allreduced_gradient_equals_one_gpu_gradient(dtype_name, chi):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, dtype_name, chi, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    loss_err, grad_err = q.get(timeout=10)
    print(f"[2-GPU] {dtype_name}: loss rel err {loss_err:.2e}, max grad err / max|g| {grad_err:.2e}")
    tol = 1e-12 if dtype_name == "float64" else 1e-5
    assert loss_err <= tol and grad_err <= tol


@pytest.mark.parametrize("dtype,chi", [(torch.float32, 64), (torch.float32, 32), (torch.float64, 24)],
                         ids=["f32_tensor", "f32_bond32_rr", "f64_simt"])
def test_deterministic_mode_gives_bitwise_identical_gradients(dtype, chi):
    """SURVEY 8e "Determinism": with tnb_set_deterministic(1) every gradient entry has one writer in a fixed order."""
    from tn4ml_b200 import _lib
    model, eng, x, t = _mps_case(dtype, chi, L=8, B=9000)
    lib = _lib.lib()
    _, g0 = eng.value_and_grad(model._packed, x, t)
    g0 = g0.clone()
    try:
        lib.tnb_set_deterministic(1)
        _, g1 = eng.value_and_grad(model._packed, x, t)
        g1 = g1.clone()
        _, g2 = eng.value_and_grad(model._packed, x, t)
        assert torch.equal(g1, g2)
    finally:
        lib.tnb_set_deterministic(0)
    tol = 1e-4 if dtype == torch.float32 else 1e-12  # (another accumulation order: float32 rounding, not bitwise)
    assert float((g1 - g0).abs().max()) <= tol * float(g0.abs().max())


@pytest.mark.parametrize("chi", [32, 64], ids=["bond32_rr", "bond64_tcgen05"])
def test_backward_in_site_ranges_equals_one_piece(chi):
    """tnb_backward_part (the data-parallel step's site ranges) on one GPU with a reducer that does nothing: the packed gradient
    equals the one-piece backward pass up to the order of the float32 atomics."""
    model, eng, x, t = _mps_case(torch.float32, chi, L=11, B=700)

    class _NoReduce:
        nparts = 4
        calls = 0

        def reduce(self, tensor):
            self.calls += 1

        def finish(self):
            pass

    ls0, g0 = eng.value_and_grad(model._packed, x, t)
    ls0, g0 = float(ls0.item()), g0.clone()
    red = _NoReduce()
    ls1, g1 = eng.value_and_grad(model._packed, x, t, reducer=red)
    assert red.calls >= 2
    assert abs(float(ls1.item()) - ls0) <= 1e-6 * abs(ls0)
    assert float((g1 - g0).abs().max()) <= 1e-5 * float(g0.abs().max())


@pytest.mark.parametrize("dtype,chi,d,B", [
    (torch.float32, 32, 2, 333),    # register-resident mma.sync kernels, ragged last tile
    (torch.float32, 10, 2, 77),     # ... zero-padded from bond 10
    (torch.float32, 24, 3, 130),    # ... d = 3
    (torch.float32, 50, 2, 130),    # tcgen05 kernels, zero-padded to 64
    (torch.float32, 128, 2, 200),   # two-tile CTAs
    (torch.float32, 256, 2, 129),   # CTA pair
    (torch.float32, 256, 2, 64),    # CTA pair with one sample tile: the second CTA of the cluster is all padding
    (torch.float64, 24, 2, 100),    # SIMT / DMMA family
], ids=["rr32", "rr_pad10", "rr_pad24_d3", "tc_pad50", "tc128", "tc256", "tc256_one_tile", "f64"])
def test_no_kernel_writes_outside_its_buffers(dtype, chi, d, B):
    """Guard bands (a byte pattern) around the workspace and the gradient buffer survive a forward + backward step and a
    forward-only call: no kernel of any family stores outside what tnb_workspace_bytes / tnb_param_count promised."""
    import tn4ml_b200 as T
    L, C_ = 9, 3
    arrays = O.make_mps_arrays(L, chi, d, class_site=4, C=C_, std=0.05, seed=3)
    x = O.make_inputs(B, L, seed=5).to(dtype).cuda()
    t = O.make_labels(B, C_).to(dtype).cuda()
    model = T.TensorNetwork(arrays, dtype=dtype, device="cuda:0")
    emb = T.TrigonometricEmbedding() if d == 2 else T.PolynomialEmbedding(degree=2, n=1, include_bias=True)
    eng, _ = model._engine(emb, _heads()["softmax_xent"], True, dtype)
    G = 1 << 20
    held = {}

    def guarded(batch, need_grad):
        key = (batch, need_grad)
        if key not in held:
            n = int(eng.lib.tnb_workspace_bytes(C.byref(eng.prob), batch, int(need_grad)))
            held[key] = torch.full((n + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda:0")
        buf = held[key]
        return buf[G: buf.numel() - G]

    eng._workspace = guarded
    esz = 4 if dtype == torch.float32 else 8
    gbuf = torch.full((eng.n_params * esz + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda:0")
    grads = gbuf[G: G + eng.n_params * esz].view(dtype)
    ls, g = eng.value_and_grad(model._packed, x, t, grads=grads)
    per, out, _ = eng.forward(model._packed, x, t)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(g).all()) and bool(torch.isfinite(per).all())
    for buf in list(held.values()) + [gbuf]:
        assert bool((buf[:G] == 0xA5).all()) and bool((buf[-G:] == 0xA5).all())
