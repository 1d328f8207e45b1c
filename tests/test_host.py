"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol the header
declares, packed layouts, symbolic loss resolution, error behaviour mirrored from the reference,
and the data-parallel shard/all-reduce logic on a 2-rank gloo group."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch

import tn4ml_b200 as T
from oracle import tn_oracle as O
from tn4ml_b200 import _lib, metrics as M
from tn4ml_b200.initializers import randn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _built()
    header = open(os.path.join(ROOT, "include", "tn4ml_b200.h")).read()
    declared = set(re.findall(r"\b(tnb_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/tn4ml_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    assert b"tn4ml_b200" in _lib.lib().tnb_version()


def test_param_count_and_offsets_match_python_layout():
    _built()
    lib = _lib.lib()
    m = T.MPS_initialize(9, initializer=randn(0.1), key=1, bond_dim=6, phys_dim=3, class_index=4, class_dim=5,
                         device="cpu")
    p = _lib.TnbProblem()
    p.kind, p.dtype, p.L, p.chi, p.d, p.n_out, p.out_site, p.spacing = 0, 1, m.L, m.chi, m.phys_dim, m.n_out, m.out_site, 1
    offs, total = m._site_offsets()
    assert lib.tnb_param_count(ctypes.byref(p)) == total == m._packed.numel()
    for i, o in enumerate(offs):
        assert lib.tnb_site_offset(ctypes.byref(p), i) == o
        assert lib.tnb_site_nout(ctypes.byref(p), i) == m._site_nout(i)
    s = T.SMPO_initialize(10, randn(0.2), 2, spacing=4, bond_dim=3, phys_dim=(2, 3), device="cpu")
    p.kind, p.L, p.chi, p.d, p.n_out, p.spacing = 1, s.L, s.chi, s.phys_dim, s.n_out, s.spacing
    offs, total = s._site_offsets()
    assert lib.tnb_param_count(ctypes.byref(p)) == total
    for i, o in enumerate(offs):
        assert lib.tnb_site_offset(ctypes.byref(p), i) == o


def test_validation_errors_come_back_as_status_and_message():
    _built()
    lib = _lib.lib()
    p = _lib.TnbProblem()
    p.kind, p.dtype, p.L, p.chi, p.d, p.n_out, p.out_site, p.spacing, p.head = 0, 1, 8, 4, 2, 1, 7, 1, 0
    p.emb.kind, p.emb.k = _lib.EMB_TRIG, 2          # embedding dim 4 != model phys dim 2  (metrics.py:34)
    assert lib.tnb_workspace_bytes(ctypes.byref(p), 16, 1) < 0
    assert b"physical dimension" in lib.tnb_last_error()
    p.emb.k = 1
    assert lib.tnb_workspace_bytes(ctypes.byref(p), 16, 1) > 0
    p.L = 1
    assert lib.tnb_workspace_bytes(ctypes.byref(p), 16, 1) < 0


def test_site_views_share_the_packed_buffer_and_keep_reference_shapes():
    m = T.MPS_initialize(7, initializer=randn(0.3), key=3, bond_dim=4, phys_dim=2, shape_method="noteven", device="cpu")
    assert [tuple(a.shape) for a in m.arrays] == [(2, 2), (2, 4, 2), (4, 4, 2), (4, 4, 2), (4, 4, 2), (4, 2, 2), (2, 2)]
    assert abs(float(m.norm()) - 1.0) < 1e-12                      # tests/test_mps.py:14-44
    before = m._packed.clone()
    m.arrays[2].mul_(2.0)
    assert not torch.equal(before, m._packed)
    m.update_tensors([a.clone() * 0 + 1 for a in m.arrays])
    assert float(m.arrays[0].sum()) == 4.0
    assert m.nparams() == sum(int(np.prod(a.shape)) for a in m.arrays)
    # zero padding outside the true bond dimensions stays zero
    assert float(m._packed.sum()) == float(sum(a.sum() for a in m.arrays))


def test_classifier_and_smpo_shapes_follow_the_reference():
    m = T.MPS_initialize(6, initializer=randn(0.3), key=1, bond_dim=5, phys_dim=3, class_index=3, class_dim=4,
                         add_identity=True, device="cpu")
    assert tuple(m.arrays[2].shape) == (5, 5, 3, 4) and m.out_site == 2 and m.n_out == 4   # test_mps.py:57-92
    assert tuple(m.arrays[0].shape) == (1, 5, 3) and tuple(m.arrays[-1].shape) == (5, 1, 3)
    s = T.SMPO_initialize(9, randn(0.5), 2, spacing=4, bond_dim=3, phys_dim=(2, 2), device="cpu")
    assert [a.ndim for a in s.arrays] == [3, 3, 3, 3, 4, 3, 3, 3, 3]
    assert tuple(s.arrays[0].shape) == (3, 2, 2) and tuple(s.arrays[-1].shape) == (3, 2, 2)  # (L-1) % S == 0
    assert abs(float(s.norm()) - 1.0) < 1e-12                      # tests/test_smpo.py:13-51
    with pytest.raises(ValueError):
        T.SMPO_initialize(8, randn(0.5), 2, spacing=1, device="cpu")                        # smpo.py:660-661
    with pytest.raises(ValueError):
        T.MPS_initialize(4, class_index=9, class_dim=2, device="cpu")                       # mps.py:311-312


def test_loss_resolution_mirrors_reference_closures():
    m = T.MPS_initialize(6, initializer=randn(0.3), key=1, bond_dim=4, phys_dim=2, class_index=3, class_dim=3,
                         device="cpu")

    def crossentropy_loss(*a, **k):  # docs/source/examples/supervised/breast_class.py:28-29
        return M.OptaxWrapper(M.softmax_cross_entropy)(*a, **k).mean()

    e = M.resolve(crossentropy_loss, m, 6, 2, True)
    assert e.head == _lib.HEAD_SOFTMAX_XENT and not e.regs

    def weighted(*a, **k):           # breast_class.py:101-102
        return M.CrossEntropyWeighted(class_weights=[0.5, 2.0, 1.0])(*a, **k).mean()

    e = M.resolve(weighted, m, 6, 2, True)
    assert e.head == _lib.HEAD_SOFTMAX_XENT_WEIGHTED and e.class_weights == (0.5, 2.0, 1.0)

    s = T.SMPO_initialize(8, randn(0.5), 2, spacing=4, bond_dim=3, phys_dim=(2, 2), device="cpu")

    def loss_combined(model, data, y_true=None, embedding=None):  # examples/unsupervised/mnist_ad.py:47-57
        return M.CombinedLoss(model, data, y_true, error=M.LogQuadNorm, reg=lambda P: 0.4 * M.LogReLUFrobNorm(P),
                              embedding=embedding)

    e = M.resolve(loss_combined, s, 8, 2, False)
    assert e.head == _lib.HEAD_LOGQUAD and e.regs[0].kind == "log_relu_frob" and e.regs[0].coef == 0.4
    with pytest.raises(ValueError):    # metrics.py:49-51: more model sites than data sites
        M.resolve(M.NegLogLikelihood, T.MPS_initialize(6, device="cpu"), 4, 2, False)
    with pytest.raises(AssertionError):  # metrics.py:34: physical dimensions must match
        M.resolve(M.NegLogLikelihood, T.MPS_initialize(6, phys_dim=3, device="cpu"), 6, 2, False)
    with pytest.raises(TypeError):
        M.resolve(lambda model, data: 1.0, m, 6, 2, False)


def test_configure_error_behaviour_matches_reference():
    m = T.MPS_initialize(6, device="cpu")
    with pytest.raises(AttributeError):
        m.configure(foo=1)                                  # model.py:216
    with pytest.raises(ValueError):
        m.configure(strategy="nonsense")                    # model.py:214
    with pytest.raises(AttributeError):
        m.configure(train_type=7)                           # model.py:227
    with pytest.raises(AttributeError):
        m.configure(train_type=0, device=("tpu", 0))        # model.py:255
    with pytest.raises(RuntimeError):
        m.configure(train_type=0, device=("cpu", 0))        # no CPU path in this build
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m.configure(train_type=0, device=("gpu", 0))    # model.py:262-266


def test_semi_supervised_loss_resolves_to_the_logquad_head_with_a_model_term():
    """metrics.SemiSupervisedLoss (metrics.py:209-233): LogQuadNorm head + 0.3 LogReLUFrobNorm inside the per-sample value."""
    m = T.SMPO_initialize(8, randn(0.1), 1, spacing=4, device="cpu")
    e = M.resolve(M.SemiSupervisedLoss, m, 8, 2, True)
    assert e.head == _lib.HEAD_LOGQUAD and e.semisup == 0.3 and not e.regs
    from tn4ml_b200.engine import Engine
    f, dfdm = Engine.semisup_value(torch.tensor([0.5, 2.0]), torch.tensor([1.0, 0.0]), torch.tensor(0.25))
    assert torch.allclose(f, torch.tensor([(1 / 0.75) ** 2, 2.25 ** 2], dtype=torch.float64))
    assert torch.allclose(dfdm, torch.tensor([2 * (1 / 0.75) * (-1 / 0.75**2), 2 * 2.25], dtype=torch.float64))


def test_sweeps_strategy_host_logic():
    """Sweeps (strategy.py:51-131): pair order of a two-way / one-way sweep, constructor errors, the split of a contracted pair
    (storage-only model on the CPU: prehook / posthook are plain tensor algebra, exact for max_bond = the bond's size)."""
    m = T.MPS_initialize(5, bond_dim=4, phys_dim=2, device="cpu")
    assert list(T.Sweeps().iterate_sites(m)) == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 3), (3, 2), (2, 1), (1, 0)]
    assert list(T.Sweeps(two_way=False).iterate_sites(m)) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    with pytest.raises(ValueError):
        T.Sweeps(grouping=1)
    with pytest.raises(ValueError):
        T.Sweeps(grouping=3)
    with pytest.raises(TypeError):
        list(T.Sweeps().iterate_sites(T.SMPO_initialize(6, randn(0.1), 1, spacing=2, device="cpu")))
    x = torch.rand(3, 5, dtype=m.dtype)
    prob = O.Problem("mps", O.EmbSpec("trig"), "nll")
    before = O.per_sample_loss(prob, x.double(), [a.double() for a in m.arrays])
    s = T.Sweeps()
    for sites in s.iterate_sites(m):
        Tm = s.prehook(m, sites)
        assert Tm.shape[2:] == (2, 2, 1)
        s.posthook(m, sites)
    after = O.per_sample_loss(prob, x.double(), [a.double() for a in m.arrays])
    assert torch.allclose(before, after, rtol=1e-9)


def test_embedding_descriptors():
    assert T.TrigonometricEmbedding(k=3).dim == 6 and T.FourierEmbedding(p=8).dim == 8   # embeddings.py:385-387
    assert T.PolynomialEmbedding(degree=2, n=1, include_bias=True).dim == 3
    assert T.GaussianRBFEmbedding(centers=np.array([0.0, 0.5, 1.0]), gamma=1.0).dim == 3
    with pytest.raises(AssertionError):
        T.TrigonometricEmbedding(k=0)
    with pytest.raises(ValueError):
        T.PolynomialEmbedding(degree=0, n=1)
    with pytest.raises(TypeError):
        from tn4ml_b200.embeddings import spec_of
        spec_of(object())                                    # embeddings.py:1346-1349


def test_batch_iterator_matches_reference_chunking():
    from tn4ml_b200.models.model import _batch_iterator
    x = np.arange(22, dtype=np.float64).reshape(11, 2)
    chunks = list(_batch_iterator(x, None, 4, shuffle=False))
    assert [len(c) for c in chunks] == [4, 4, 3]              # last partial chunk kept (model.py:1186-1198)
    flipped = list(_batch_iterator(x, None, 4, shuffle=False, alternate_flip=True))
    assert torch.equal(flipped[1], torch.flip(chunks[1], dims=[1])) and torch.equal(flipped[0], chunks[0])
    xs, ys = zip(*_batch_iterator(x, np.arange(11), 5, shuffle=True, seed=3))
    assert sorted(torch.cat(ys).tolist()) == list(range(11))
    assert all(torch.equal(a[:, 0] / 2, b.double()) for a, b in zip(xs, ys))


def test_optax_style_adam_matches_closed_form():
    opt = T.optim.adam(learning_rate=0.1)
    p = torch.tensor([1.0, -2.0], dtype=torch.float64)
    g = torch.tensor([0.5, 0.25], dtype=torch.float64)
    st = opt.init(p)
    upd, st = opt.update(g, st, p)
    assert torch.allclose(upd, -0.1 * g / (g.abs() + 1e-8))  # first Adam step: -lr * g / (|g| + eps)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from tn4ml_b200.distributed import allreduce_step, shard_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arrays = O.make_mps_arrays(8, 4, 2, class_site=3, C=3, std=0.2, seed=1)
    x, t = O.make_inputs(11, 8, seed=2), O.make_labels(11, 3)      # 11 rows: uneven shards
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    xs, ts = shard_batch(x, t)
    # local compute stands in for the CUDA engine: sum of per-sample losses and grads of sum / N_global
    leaves = [a.clone().requires_grad_(True) for a in arrays]
    per = O.per_sample_loss(prob, xs, leaves, ts)
    gl = torch.autograd.grad(per.sum() / len(x), leaves)
    packed = torch.cat([g.reshape(-1) for g in gl])
    loss_sum, packed = allreduce_step(per.sum().detach().reshape(1), packed)
    if rank == 0:
        out.put((float(loss_sum) / len(x), packed.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_shards_and_allreduce_equal_full_batch_gloo():
    import torch.multiprocessing as mp
    from tn4ml_b200.distributed import shard_bounds
    assert [shard_bounds(11, r, 2) for r in range(2)] == [(0, 6), (6, 11)]
    assert [shard_bounds(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    loss, packed = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    arrays = O.make_mps_arrays(8, 4, 2, class_site=3, C=3, std=0.2, seed=1)
    x, t = O.make_inputs(11, 8, seed=2), O.make_labels(11, 3)
    lref, gref = O.value_and_grad(O.Problem("mps", O.EmbSpec("trig"), "softmax_xent"), x, arrays, t)
    assert abs(loss - float(lref)) < 1e-12
    assert torch.allclose(packed, torch.cat([g.reshape(-1) for g in gref]), rtol=1e-10, atol=1e-14)
