"""Parity of the tcgen05 (fp16x3) sweep against the CPU oracle: float32 tolerance 1e-4."""
import pytest
import torch

from oracle import tn_oracle as O
from tests.test_gpu_parity import _check, _heads

pytestmark = pytest.mark.gpu


def _tensor_model(cls, arrays):
    import tn4ml_b200 as T
    from tn4ml_b200 import _lib
    model = cls(arrays, dtype=torch.float32, device="cuda:0")
    model.precision = _lib.PREC_TENSOR
    return model


@pytest.mark.parametrize("case", [
    dict(L=6, chi=64, d=2, cls=3, C=3, B=128),
    dict(L=12, chi=64, d=2, cls=5, C=4, B=200),       # ragged last tile
    dict(L=9, chi=128, d=2, cls=0, C=3, B=130),       # class site on the left boundary
    dict(L=9, chi=32, d=4, cls=8, C=2, B=64),         # class site on the right boundary, d=4
    dict(L=8, chi=256, d=2, cls=4, C=10, B=256),      # cfg 5 bond dimension
    dict(L=10, chi=128, d=3, cls=4, C=2, B=90),       # N = 384 -> two MMA halves of 192
    dict(L=14, chi=32, d=2, cls=6, C=10, B=300),      # chi = 32 (the metric's small-bond point): dA tile half empty
    dict(L=30, chi=32, d=3, cls=14, C=2, B=160),      # cfg 2 shape: N = 96, partial dA tile
    dict(L=8, chi=48, d=2, cls=3, C=3, B=70),         # chi = 48: N = 96
    dict(L=6, chi=256, d=2, cls=2, C=3, B=64),        # one sample tile on the CTA-pair kernel (its second CTA is all padding)
    dict(L=6, chi=192, d=2, cls=3, C=2, B=130),       # a wide bond that stays on the one-CTA kernel by default
], ids=["chi64", "chi64_ragged", "chi128_cls0", "chi32_d4_clsL", "chi256", "chi128_d3", "chi32", "cfg2_shape", "chi48",
        "chi256_one_tile", "chi192"])
def test_tensor_path_classifier(case):
    import tn4ml_b200 as T
    from tn4ml_b200 import _lib
    arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], class_site=case["cls"], C=case["C"], std=0.05, seed=3)
    x = O.make_inputs(case["B"], case["L"], seed=5)
    t = O.make_labels(case["B"], case["C"])
    emb = {2: O.EmbSpec("trig"), 3: O.EmbSpec("poly", degree=2, include_bias=True), 4: O.EmbSpec("trig", k=2)}[case["d"]]
    prob = O.Problem("mps", emb, "softmax_xent")
    model = _tensor_model(T.TensorNetwork, arrays)
    l0 = _lib.lib().tnb_launch_count()
    _check(model, prob, _heads()["softmax_xent"], x, arrays, t, torch.float32, precision=1)
    assert _lib.lib().tnb_launch_count() > l0


def test_tensor_path_cfg4_nll():
    """BASELINE config 4 shape: L=16, Fourier d=8, chi=64, NLL."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(16, 64, 8, std=0.05, seed=4, squeeze_ends=True, identity_all=True)
    x = O.make_inputs(300, 16, seed=6)
    prob = O.Problem("mps", O.EmbSpec("fourier", p=8), "nll")
    model = _tensor_model(T.MatrixProductState, arrays)
    _check(model, prob, _heads()["nll"], x, arrays, None, torch.float32, precision=1)


def test_tensor_path_long_chain():
    """L=196 at chi=64: the accuracy budget of the fp16x3 split over the full cfg-5 chain length."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(196, 64, 2, class_site=97, C=10, std=1e-2, seed=42)
    x = O.make_inputs(256, 196, seed=1)
    t = O.make_labels(256, 10)
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    model = _tensor_model(T.TensorNetwork, arrays)
    _check(model, prob, _heads()["softmax_xent"], x, arrays, t, torch.float32, precision=1)


def test_tensor_path_two_tile_rounds_and_remainder():
    """chi = 128 (N = 256, one column chunk): a sweep over 297 sample tiles runs as two-tile CTAs for the whole rounds of
    148 SMs and one-tile CTAs for the remainder (forward: 296 + 1 tiles; adjoint: 148-tile rounds per half)."""
    import tn4ml_b200 as T
    B = 297 * 128 - 5
    arrays = O.make_mps_arrays(5, 128, 2, class_site=2, C=3, std=0.05, seed=3)
    x = O.make_inputs(B, 5, seed=5)
    t = O.make_labels(B, 3)
    prob = O.Problem("mps", O.EmbSpec("trig"), "softmax_xent")
    model = _tensor_model(T.TensorNetwork, arrays)
    _check(model, prob, _heads()["softmax_xent"], x, arrays, t, torch.float32, precision=1)


@pytest.mark.parametrize("precision", [0, 1], ids=["simt", "tensor"])
def test_micro_batched_step_equals_one_piece(precision):
    """The host splits a step into micro-batches when the environment stash would not fit (cfg 5 at full size); the
    gradient accumulates across tnb_backward calls (accumulate=1) and the loss sum across tnb_forward calls."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(9, 64, 2, class_site=4, C=3, std=0.05, seed=3)
    x = O.make_inputs(300, 9, seed=5)
    t = O.make_labels(300, 3)
    model = _tensor_model(T.TensorNetwork, arrays)
    model.precision = precision
    model._engine_cache = {}
    eng, _ = model._engine(T.TrigonometricEmbedding(), _heads()["softmax_xent"], True, torch.float32)
    ls1, g1 = eng.value_and_grad(model._packed, x.float(), t.float())
    ls1, g1 = float(ls1.item()), g1.clone()
    eng.max_microbatch = 77                      # 300 = 77 + 77 + 77 + 69: ragged tiles in every piece
    eng._ws.clear()
    ls2, g2 = eng.value_and_grad(model._packed, x.float(), t.float())
    assert abs(float(ls2.item()) - ls1) <= 1e-5 * abs(ls1)
    assert float((g2 - g1).abs().max()) <= 2e-5 * float(g1.abs().max())


@pytest.mark.parametrize("cg", ["1", "2"], ids=["one_cta", "cta_pair"])
def test_chain_kernel_variants_at_chi_256(cg):
    """Both chain kernels of the wide-bond case against the oracle at chi = 256 with an odd number of sample tiles: the CTA pair
    (TNB_TC_CG=2: cta_group::2 MMAs, each CTA of a cluster of two holding half of the weight columns, the pair's second tile all
    padding here; the default at chi = 256) and the one-CTA kernel (TNB_TC_CG=1).  The knob is read once per process, hence the
    subprocess (tools/pair_case.py prints the errors)."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TNB_TC_CG=cg)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "pair_case.py"), "300"], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"loss ([0-9.eE+-]+) oracle ([0-9.eE+-]+)\s+max grad err / max\|g\| ([0-9.eE+-]+)", out.stdout)
    assert m, out.stdout[-2000:]
    loss, ref, err = (float(v) for v in m.groups())
    assert abs(loss - ref) <= 1e-4 * abs(ref)
    assert err <= 1e-4


# ---- bond 32: the register-resident chain kernel (mps_rr_chain.cuh, `mma.sync` fp16x3) ------------------------------------
@pytest.mark.parametrize("case", [
    dict(L=10, d=2, cls=0, C=3, B=130, head="softmax_xent"),        # class site on the left boundary: no left chain
    dict(L=10, d=2, cls=9, C=4, B=17, head="softmax_xent"),         # ... on the right boundary; one partial warp
    dict(L=7, d=3, cls=3, C=2, B=129, head="softmax_xent"),         # d = 3, one sample in the second tile
    dict(L=12, d=2, cls=5, C=10, B=1, head="mse"),                  # a single sample
    dict(L=12, d=2, cls=6, C=5, B=257, head="xent_softmax_argmax"),
], ids=["cls0", "clsL", "d3", "one_sample_mse", "argmax"])
def test_rr_chain_classifier(case):
    import tn4ml_b200 as T
    from tn4ml_b200 import _lib
    arrays = O.make_mps_arrays(case["L"], 32, case["d"], class_site=case["cls"], C=case["C"], std=0.05, seed=11)
    x = O.make_inputs(case["B"], case["L"], seed=12)
    t = O.make_labels(case["B"], case["C"])
    emb = {2: O.EmbSpec("trig"), 3: O.EmbSpec("poly", degree=2, include_bias=True)}[case["d"]]
    model = _tensor_model(T.TensorNetwork, arrays)
    _check(model, O.Problem("mps", emb, case["head"]), _heads()[case["head"]], x, arrays, t, torch.float32, precision=1)


def test_rr_chain_nll_and_long_chain():
    """Unsupervised NLL at bond 32 (d = 2 trig), and the cfg-5 chain length at the metric's small-bond point (L = 196, 10 classes)."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(20, 32, 2, std=0.05, seed=4, squeeze_ends=True, identity_all=True)
    x = O.make_inputs(200, 20, seed=6)
    model = _tensor_model(T.MatrixProductState, arrays)
    _check(model, O.Problem("mps", O.EmbSpec("trig"), "nll"), _heads()["nll"], x, arrays, None, torch.float32, precision=1)
    arrays = O.make_mps_arrays(196, 32, 2, class_site=97, C=10, std=1e-2, seed=42)
    x = O.make_inputs(300, 196, seed=1)
    t = O.make_labels(300, 10)
    model = _tensor_model(T.TensorNetwork, arrays)
    _check(model, O.Problem("mps", O.EmbSpec("trig"), "softmax_xent"), _heads()["softmax_xent"], x, arrays, t, torch.float32,
           precision=1)


def _launched_kernels(fn):
    """Names of the library's kernels launched by fn() (tnb_profile: every launch bracketed with CUDA events)."""
    import ctypes as C
    from tn4ml_b200 import _lib
    lib = _lib.lib()
    lib.tnb_profile(1)
    try:
        fn()
        torch.cuda.synchronize()
        nmax = 4096
        names = C.create_string_buffer(48 * nmax)
        ms = (C.c_float * nmax)()
        n = lib.tnb_profile_read(names, ms, nmax)
        return {names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode() for i in range(n)}
    finally:
        lib.tnb_profile(0)


@pytest.mark.parametrize("case", [
    dict(L=20, chi=10, d=2, cls=9, C=10, B=64, kw={}),                          # BASELINE cfg 1's bond and batch
    dict(L=9, chi=24, d=3, cls=4, C=2, B=130, kw={}),                           # d = 3 (252-register dA variant)
    dict(L=8, chi=16, d=2, cls=3, C=3, B=77, kw={}),
    dict(L=10, chi=31, d=2, cls=0, C=4, B=200, kw={}),                          # odd bond, class site on the boundary
], ids=["chi10_cfg1", "chi24_d3", "chi16", "chi31_cls0"])
def test_small_bonds_run_zero_padded_on_the_register_resident_family(case):
    """float32 bonds of 8 .. 31 (d = 2, 3): the library builds bond-32 weight images from the caller's [chi][chi][d] slabs and
    runs the mma.sync kernels; gradients come back in the caller's layout.  Oracle parity at 1e-4 and the kernels that ran."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], class_site=case["cls"], C=case["C"], std=0.05, seed=21)
    x = O.make_inputs(case["B"], case["L"], seed=22)
    t = O.make_labels(case["B"], case["C"])
    emb = {2: O.EmbSpec("trig"), 3: O.EmbSpec("poly", degree=2, include_bias=True)}[case["d"]]
    model = _tensor_model(T.TensorNetwork, arrays)
    seen = _launched_kernels(lambda: _check(model, O.Problem("mps", emb, "softmax_xent"), _heads()["softmax_xent"], x, arrays, t,
                                            torch.float32, precision=1))
    assert {"rr_chain_kernel(fwd)", "rr_chain_kernel(adj)", "rr_da_kernel", "tc_da_kernel(class)"} <= seen, seen
    assert not any(k.startswith("mps_sweep") or k.startswith("tc_chain") for k in seen), seen


def test_small_bond_ragged_shapes_and_nll_zero_padded():
    """Ragged ("noteven") bonds and squeezed boundaries below 32, NLL head: the padded path against the oracle."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(14, 12, 2, std=0.05, seed=4, squeeze_ends=True, shape_method="noteven", identity_all=True)
    x = O.make_inputs(150, 14, seed=6)
    model = _tensor_model(T.MatrixProductState, arrays)
    seen = _launched_kernels(lambda: _check(model, O.Problem("mps", O.EmbSpec("trig"), "nll"), _heads()["nll"], x, arrays, None,
                                            torch.float32, precision=1))
    assert "rr_da_kernel" in seen, seen


@pytest.mark.parametrize("case", [
    dict(L=8, chi=50, d=2, cls=3, C=3, B=140),      # -> 64: the two-CTA small-step kernel, general dA epilogue
    dict(L=7, chi=20, d=4, cls=2, C=2, B=70),       # -> 32 with d = 4: tcgen05 kernels (the mma.sync family has d = 2, 3)
    dict(L=7, chi=100, d=2, cls=3, C=4, B=130),     # -> 112: N = 224, one column chunk
    dict(L=6, chi=120, d=2, cls=2, C=3, B=129),     # -> 128: two-tile CTAs, the transposing dA epilogue with both guards
    dict(L=6, chi=200, d=2, cls=3, C=2, B=64),      # -> 208: N = 416, two column chunks
], ids=["chi50", "chi20_d4", "chi100", "chi120", "chi200"])
def test_bonds_between_the_tensor_kernels_sizes_run_zero_padded(case):
    """float32 bonds that are not multiples of 16 run zero-padded to the next one on the tcgen05 kernels (weight images padded,
    gradients in the caller's layout): oracle parity at 1e-4, and no SIMT sweep in the launch list."""
    import tn4ml_b200 as T
    arrays = O.make_mps_arrays(case["L"], case["chi"], case["d"], class_site=case["cls"], C=case["C"], std=0.05, seed=31)
    x = O.make_inputs(case["B"], case["L"], seed=32)
    t = O.make_labels(case["B"], case["C"])
    emb = {2: O.EmbSpec("trig"), 4: O.EmbSpec("trig", k=2)}[case["d"]]
    model = _tensor_model(T.TensorNetwork, arrays)
    seen = _launched_kernels(lambda: _check(model, O.Problem("mps", emb, "softmax_xent"), _heads()["softmax_xent"], x, arrays, t,
                                            torch.float32, precision=1))
    assert "tc_chain_kernel(fwd)" in seen and "tc_da_kernel" in seen, seen
    assert not any(k.startswith("mps_sweep") for k in seen), seen
