"""GPU parity of the Self-ONN variants: the streaming ops behind the operational layers and whole models per
layer, against the oracle.

NOT YET RUN ON HARDWARE: this file was written after the round's GPU budget was spent.  The planner side is pinned by the float64
CPU emulator tests (tests/test_plan_families_cpu.py::test_*_self_onn_family); the three kernels touched (eltwise ops 4 / 5, tanh in
the activation helpers, b2seg_outact_fwd / bwd) are element-wise.  The file sorts last so that a surprise here cannot mask the
results of the suites above it under `pytest -x`.

Inputs are scaled to [0, 0.5): the Self-ONN encoder is linear and un-normalised (unet_variants.py:782-786), so with inputs in [0, 1) at
random init its cubes of cubes reach 1e17-1e38 (oracle, float64) — outside any useful bf16 range and not a property of the kernels.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from b2seg import _lib as L  # noqa: E402
from b2seg import lowering as lw  # noqa: E402
from b2seg.models1d import UNet  # noqa: E402
from b2seg.models2d import unet_model_builder  # noqa: E402
from oracle.ref_models import Ref1D, Ref2D  # noqa: E402
from test_gpu_kernels import bf, rel_l2, stream, tv  # noqa: E402
from test_gpu_model import check_per_layer  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _dev():
    L.check(L.load().b2seg_device_check(0), "device_check")
    torch.manual_seed(0)
    yield
    torch.cuda.synchronize()


def test_self_onn_pow_tanh_outact():
    """the streaming ops behind the operational layers (onn_layers.py:6-48): x^p and its backward (eltwise ops 4 / 5), tanh in
    the activation kernels (forward and backward from the pre-activation), and an Activation used as a model output"""
    dev = "cuda"
    N, H, W, Cc = 2, 16, 24, 64
    nv = lw.NULL_VIEW.to_c()
    x = bf(torch.randn(N, H, W, Cc, device=dev))
    g = bf(torch.randn(N, H, W, Cc, device=dev))
    for p in (2, 3, 4):
        y, dx = torch.zeros_like(x), torch.zeros_like(x)
        L.call("b2seg_eltwise", L.EltwiseDesc(4, tv(x).to_c(), nv, nv, tv(y).to_c(), p), stream())
        L.call("b2seg_eltwise", L.EltwiseDesc(5, tv(x).to_c(), tv(g).to_c(), nv, tv(dx).to_c(), p), stream())
        torch.cuda.synchronize()
        assert torch.equal(y, (x.float() ** p).to(torch.bfloat16)) or rel_l2(y.float(), x.float() ** p) < 4e-3
        assert rel_l2(dx.float(), p * x.float() ** (p - 1) * g.float()) < 4e-3
    # tanh through the activation kernel (no BatchNorm) and its backward
    o = torch.zeros_like(x)
    d = L.BnActDesc()
    d.x, d.act, d.n_out = tv(x).to_c(), L.ACT_TANH, 1
    d.out[0] = tv(o).to_c()
    L.call("b2seg_bn_act", d, stream())
    dz = torch.zeros_like(x)
    bd = L.BnBwdDesc()
    bd.x, bd.act, bd.n_src = tv(x).to_c(), L.ACT_TANH, 1
    bd.src[0] = L.GradSrc(tv(g).to_c(), 0, 1, 1)
    bd.count, bd.dx = 1.0, tv(dz).to_c()
    L.call("b2seg_bn_bwd", bd, stream())
    torch.cuda.synchronize()
    t = torch.tanh(x.float())
    assert rel_l2(o.float(), t) < 4e-3
    assert rel_l2(dz.float(), g.float() * (1 - t * t)) < 5e-3
    # Activation as a model output: fp32 [pixel][cout] from an 8-channel bf16 block, and the bf16 gradient back
    for cout, act, cp in ((1, L.ACT_SIGMOID, 8), (4, L.ACT_SOFTMAX, 8), (3, L.ACT_NONE, 8), (11, L.ACT_SOFTMAX, 16), (9, L.ACT_SIGMOID, 16)):
        z = bf(torch.randn(N, H, W, cp, device=dev) * 3)
        yv = torch.full((N, H, W, cout), -7.0, device=dev)
        dl = torch.randn(N, H, W, cout, device=dev)
        dzo = torch.full_like(z, 5.0)
        L.call("b2seg_outact_fwd", L.OutActDesc(tv(z).to_c(), cout, act, yv.data_ptr(), 0, nv), stream())
        L.call("b2seg_outact_bwd", L.OutActDesc(nv, cout, act, 0, dl.data_ptr(), tv(dzo).to_c()), stream())
        torch.cuda.synchronize()
        zf = z.float()[..., :cout]
        want = torch.sigmoid(zf) if act == L.ACT_SIGMOID else (torch.softmax(zf, -1) if act == L.ACT_SOFTMAX else zf)
        assert rel_l2(yv, want) < 1e-4      # __expf
        assert torch.equal(dzo[..., :cout], dl.to(torch.bfloat16)) and float(dzo.float()[..., cout:].abs().max()) == 0


SELF_CASES = [
    ("SelfUNet", dict(ds=1), 64, 16, 3),                              # :644: operational transposed up-sampling (tanh), BN + tanh nodes
    ("SelfUNetPP", dict(ds=1, is_transconv=False, q=2), 64, 16, 2),   # :667
    ("SelfUNet3P", dict(ds=1, output_nums=3, final_activation="softmax"), 64, 16, 2),   # :713: stride-2 operational level heads
]


@pytest.mark.parametrize("dec,kw,size,width,depth", SELF_CASES, ids=[c[0] + "-" + "-".join(f"{k}{v}" for k, v in c[1].items()) for c in SELF_CASES])
def test_2d_self_onn_per_layer(dec, kw, size, width, depth):
    kw = dict(num_channels=3, **kw)
    m = unet_model_builder(dec, size, size, width, depth, train_mode="from_scratch", **kw).ResNet50()
    rng = np.random.default_rng(21)
    x = 0.5 * rng.random((4, size, size, 3), dtype=np.float32)
    targets, losses = [], []
    for i, n in enumerate(m.graph.outputs):
        H, W, C = n.shape
        if i == 0 and kw.get("final_activation") == "softmax":
            targets.append(np.eye(C, dtype=np.float32)[rng.integers(0, C, (4, H, W))]); losses.append("cce")
        elif i == 0:
            targets.append((rng.random((4, H, W, C)) > 0.6).astype(np.float32)); losses.append("bce")
        else:
            targets.append(rng.standard_normal((4, H, W, C)).astype(np.float32)); losses.append("mse")
    # (the outputs carry the nested models' auto-names, oper2d_k, not 'out' / 'level k': unet_variants.py:1107-1108, :653)
    check_per_layer(m, Ref2D(dec, size, size, width, depth, **kw), 2, x, targets, losses, e2e_bound=1.0, mask_outputs=[m.output_names[0]])


@pytest.mark.parametrize("var,kw", [("SelfUNetPP", dict(ds=1)), ("SelfR2UNetPP", dict(ds=1, t=2, q=2)), ("SelfUNet3P", dict(ds=1, q=2))],
                         ids=["SelfUNetPP", "SelfR2UNetPP-q2", "SelfUNet3P-q2"])
def test_1d_self_onn_per_layer(var, kw):
    """1DCNN/Models/unet_variants.py:1312-1583 (Oper1D pairs, Oper1DTranspose with kernel 4, Self_Recurrent_Conv_Block)"""
    m = getattr(UNet(256, 2, 2, 16, 3, problem_type="Regression", output_nums=1, **kw), var)()
    rng = np.random.default_rng(22)
    x = (0.5 * (rng.random((4, 256, 2)) - 0.5)).astype(np.float32)
    targets = [rng.standard_normal((4,) + tuple(n.shape[1:])).astype(np.float32) for n in m.graph.outputs]
    check_per_layer(m, Ref1D(var, 256, 2, 2, 16, 3, problem_type="Regression", output_nums=1, **kw), 1, x, targets, ["mse"] * len(targets),
                    e2e_bound=1.0)


def test_ds_target_pyramid_on_device():
    """compile(ds_targets=...) (input pipeline): b2seg_target_pool against torch, and a training step fed the bare
    mask against the same step fed the reference's host-built dictionary (helper_functions.py:359-380)"""
    import torch.nn.functional as F
    from b2seg.helpers import prepareTrainDict
    from b2seg.model import Adam
    dev = "cuda"
    src = torch.randn(3, 32, 48, 2, device=dev)
    for (ph, pw, mode) in ((2, 2, 0), (8, 8, 0), (1, 4, 1)):
        dst = torch.zeros(3, 32 // ph, 48 // pw, 2, device=dev)
        L.call("b2seg_target_pool", L.TPoolDesc(src.data_ptr(), dst.data_ptr(), 3, 32, 48, 2, ph, pw, mode), stream())
        torch.cuda.synchronize()
        t = src.permute(0, 3, 1, 2)
        want = (F.max_pool2d(t, (ph, pw)) if mode == 0 else F.avg_pool2d(t, (ph, pw))).permute(0, 2, 3, 1)
        assert torch.allclose(dst, want, atol=1e-6)
    rng = np.random.default_rng(23)
    x = rng.random((4, 64, 64, 3), dtype=np.float32)
    mask = (rng.random((4, 64, 64, 1)) > 0.6).astype(np.float32)
    losses = {}
    for how in ("host", "device"):
        m = unet_model_builder("UNet", 64, 64, 16, 3, ds=1, train_mode="from_scratch").ResNet50()
        m.compile(loss={"out": "binary_crossentropy", "level1": "mse", "level2": "mse", "level3": "mse"}, optimizer=Adam(1e-3),
                  ds_targets="UNet" if how == "device" else None)
        losses[how] = m.train_on_batch(x, mask if how == "device" else prepareTrainDict(mask, 3, "UNet"))
    # same weights, same targets: equal up to the summation order of the loss kernel's atomics
    assert abs(losses["host"] - losses["device"]) < 1e-5 * max(1.0, abs(losses["host"])), losses


def test_wide_softmax_head_per_layer():
    """an 11-class softmax head: above the 8 classes of the pointwise-head kernels the 1x1 convolution runs on the tensor-core
    kernels and b2seg_outact applies the softmax (unet_variants.py:1106 with output_nums = 11)"""
    kw = dict(num_channels=3, ds=1, output_nums=11, final_activation="softmax")
    m = unet_model_builder("UNet", 64, 64, 16, 3, train_mode="from_scratch", **kw).ResNet50()
    rng = np.random.default_rng(24)
    x = rng.random((4, 64, 64, 3), dtype=np.float32)
    targets, losses = [], []
    for n in m.graph.outputs:
        H, W, C = n.shape
        if n.name == "out":
            targets.append(np.eye(C, dtype=np.float32)[rng.integers(0, C, (4, H, W))]); losses.append("cce")
        else:
            targets.append(rng.standard_normal((4, H, W, C)).astype(np.float32)); losses.append("mse")
    check_per_layer(m, Ref2D("UNet", 64, 64, 16, 3, **kw), 2, x, targets, losses, e2e_bound=1.0)
