"""Buffer reuse by liveness (Planner._assign_memory) on the float64 emulator: the arena is filled with NaN before every step, so an op
that reads bytes no op of the same step has written poisons the result.  Reuse and no-reuse engines must agree EXACTLY (same
arithmetic, same order), over several steps, for every graph family of the BASELINE configs and the odd ones (strided dgrad of
the attention gates, gapped MultiRes layouts, ConvLSTM frames, Dense bottleneck, inference plans)."""
import numpy as np
import pytest
import torch

import b2seg.engine
from b2seg.model import Adam
from b2seg.models1d import BCDUNet, UNet
from b2seg.models2d import unet_model_builder
from cpu_engine import CpuEngine


@pytest.fixture()
def cpu_engine(monkeypatch):
    monkeypatch.setattr(b2seg.engine, "Engine", CpuEngine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)


CASES = {
    "unet": lambda: unet_model_builder("UNet", 32, 32, 8, 3, train_mode="from_scratch", num_channels=3).ResNet50(),
    "unetpp-ds-ag-softmax": lambda: unet_model_builder("UNetPP", 32, 32, 8, 2, train_mode="from_scratch", num_channels=3, ds=1, ag=1, output_nums=3,
                                                       final_activation="softmax").ResNet50(),
    "multires": lambda: unet_model_builder("MultiResUNet", 32, 32, 16, 2, train_mode="from_scratch", num_channels=1).ResNet50(),
    "bcdunet2d": lambda: unet_model_builder("UNet", 32, 32, 16, 2, train_mode="from_scratch", num_channels=3, lstm=1, dense_loop=2).ResNet50(),
    "unet3p-ds": lambda: unet_model_builder("UNet3P", 32, 32, 8, 2, train_mode="from_scratch", num_channels=3, ds=1).ResNet50(),
    "ae-bilinear": lambda: unet_model_builder("UNet", 32, 32, 8, 2, train_mode="from_scratch", num_channels=3, ae=1, feature_number=16, is_transconv=False).ResNet50(),
    "unet1d-ds-ag": lambda: UNet(64, 2, 2, 8, 3, ds=1, ag=1).UNet(),
    "bcdunet1d": lambda: BCDUNet(64, 2, 2, 16, 3, ds=1, ag=1, lstm=1, dense_loop=2).BCDUNet(),
    "r2unet1d": lambda: UNet(64, 2, 2, 8, 3, ds=1, t=2).R2UNet(),
}


@pytest.mark.parametrize("case", list(CASES))
def test_reuse_equals_no_reuse_with_poisoned_arena(cpu_engine, case):
    a, b = CASES[case](), CASES[case]()
    a.keep_activations, b.keep_activations = True, False
    rng = np.random.default_rng(3)
    H, W, C = a.graph.inputs[0].shape
    x = rng.random((2, H, W, C), dtype=np.float32) if a.graph.ndim == 2 else rng.standard_normal((2, W, C)).astype(np.float32)
    targets, losses = [], []
    for n in a.graph.outputs:
        shp = (2,) + (tuple(n.shape) if a.graph.ndim == 2 else tuple(n.shape[1:]))
        fn = n.attrs.get("activation") if n.op == "conv" else n.attrs.get("fn")
        if fn == "softmax":
            targets.append(np.eye(shp[-1], dtype=np.float32)[rng.integers(0, shp[-1], shp[:-1])]); losses.append("cce")
        elif fn == "sigmoid":
            targets.append((rng.random(shp) > 0.5).astype(np.float32)); losses.append("bce")
        else:
            targets.append(rng.standard_normal(shp).astype(np.float32)); losses.append("mse")
    for m in (a, b):
        m.compile(loss=losses if len(losses) > 1 else losses[0], optimizer=Adam(2e-3))
    b.set_weight_dict(a.get_weight_dict())
    tg = targets if len(targets) > 1 else targets[0]
    for step in range(3):
        la, lb = a.train_on_batch(x, tg), b.train_on_batch(x, tg)
        assert np.isfinite(lb) and la == lb, (case, step, la, lb)
    ea, eb = a._engine(2, True), b._engine(2, True)
    assert eb.reuse and not ea.reuse
    for oa, ob in zip(ea.outputs, eb.outputs):
        assert torch.equal(oa["y"], ob["y"])
    assert torch.equal(ea.g, eb.g) and torch.equal(ea.w, eb.w) and torch.equal(ea.moving, eb.moving)
    st = eb.planner.reuse_stats
    assert st["arena_bytes"] < st["tensor_bytes"]
    # inference plans keep almost nothing
    pa, pb = a.predict(x, batch_size=2), b.predict(x, batch_size=2)
    for qa, qb in zip(pa if isinstance(pa, list) else [pa], pb if isinstance(pb, list) else [pb]):
        assert np.array_equal(qa, qb)
    si = b._engine(2, False).planner.reuse_stats
    assert si["arena_bytes"] <= 0.7 * si["tensor_bytes"], si


def test_arena_size_of_the_baseline_configs():
    """the packing itself at BASELINE shapes (planning only, no memory touched): config 3 at batch 32 and config 4 at batch 8 must fit
    comfortably in one B200 (18.4 GB at batch 8 / 43.6 GB at batch 8 without reuse)"""
    from b2seg.planner import Planner

    def plan(model, batch, losses):
        top = [1 << 40]

        def alloc(nbytes, tag="act"):
            top[0] += (nbytes + 1023) // 1024 * 1024
            return top[0] - (nbytes + 1023) // 1024 * 1024
        return Planner(model.graph, batch, alloc, training=True, losses=losses, reuse=True).build()
    m3 = unet_model_builder("UNetPP", 256, 256, 64, 5, train_mode="from_scratch", num_channels=3, output_nums=4, ds=1, ag=1, final_activation="softmax").ResNet50()
    p3 = plan(m3, 32, ["cce"] + ["mse"] * 5)
    m4 = unet_model_builder("MultiResUNet", 512, 512, 64, 5, train_mode="from_scratch", num_channels=1, alpha=1.0).ResNet50()
    p4 = plan(m4, 8, ["bce"])
    gb3, gb4 = p3.reuse_stats["arena_bytes"] / 2 ** 30, p4.reuse_stats["arena_bytes"] / 2 ** 30
    print(f"cfg3 batch 32: arena {gb3:.1f} GB of {p3.reuse_stats['tensor_bytes'] / 2 ** 30:.1f} GB; cfg4 batch 8: arena {gb4:.1f} GB of "
          f"{p4.reuse_stats['tensor_bytes'] / 2 ** 30:.1f} GB")
    assert gb3 < 60 and gb4 <= 25
