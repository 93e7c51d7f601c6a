"""One-process training with Adam cut into buckets that run beside backward (Model._step_overlapped_adam, the default on a GPU whenever
the parameter arena is larger than one bucket).  On the emulator engine the side stream is program order: bucket i's Adam runs right
after the backward op the schedule names, BEFORE the rest of backward — so if any later backward op still read a weight of that bucket
(a dgrad after its layer's wgrad, the attention gate's second pass, a ConvLSTM's recurrent kernel) or still added to one of its gradients,
the step would differ from the plain forward / backward / Adam order.  Three steps, every weight and moment, several families."""
import numpy as np
import pytest
import torch

import b2seg.engine
from b2seg.model import Adam
from b2seg.models1d import UNet
from b2seg.models2d import unet_model_builder
from cpu_engine import CpuEngine


@pytest.fixture()
def cpu_engine(monkeypatch):
    monkeypatch.setattr(b2seg.engine, "Engine", CpuEngine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)


def _user_shape(m, shape):
    """graph shapes are (H, W, C); a 1D model's are (1, L, C) and its user-facing arrays (N, L, C)"""
    shape = tuple(shape)
    return shape[1:] if m.graph.ndim == 1 else shape


def _targets(m, rng, n):
    ts, ls = [], []
    for o in m.graph.outputs:
        shp = (n,) + _user_shape(m, o.shape)
        if o.attrs.get("activation") == "softmax":
            ts.append(np.eye(shp[-1], dtype=np.float32)[rng.integers(0, shp[-1], shp[:-1])]); ls.append("categorical_crossentropy")
        elif o.attrs.get("activation") == "sigmoid":
            ts.append((rng.random(shp) > 0.5).astype(np.float32)); ls.append("binary_crossentropy")
        else:
            ts.append(rng.standard_normal(shp).astype(np.float32)); ls.append("mse")
    return ts, ls


CASES = [
    ("unet2d", lambda: unet_model_builder("UNet", 16, 16, 8, 2, num_channels=2, train_mode="from_scratch").ResNet50()),
    ("unetpp2d-ds-ag-softmax", lambda: unet_model_builder("UNetPP", 16, 16, 8, 2, num_channels=2, output_nums=3, ds=1, ag=1,
                                                           final_activation="softmax", train_mode="from_scratch").ResNet50()),
    ("multires2d", lambda: unet_model_builder("MultiResUNet", 16, 16, 16, 2, num_channels=2, train_mode="from_scratch").ResNet50()),
    ("bcdunet2d-lstm-dense2", lambda: unet_model_builder("UNet", 16, 16, 16, 2, num_channels=2, lstm=1, dense_loop=2,
                                                         train_mode="from_scratch").ResNet50()),
    ("unet3p2d-ds", lambda: unet_model_builder("UNet3P", 16, 16, 8, 2, num_channels=2, ds=1, train_mode="from_scratch").ResNet50()),
    ("unet1d-ds-ag", lambda: UNet(64, 2, 2, 8, 3, ds=1, ag=1).UNet()),
    ("r2unet1d", lambda: UNet(64, 2, 2, 8, 3, t=2).R2UNet()),
]


@pytest.mark.parametrize("name,build", CASES, ids=[c[0] for c in CASES])
def test_bucketed_adam_beside_backward_equals_the_plain_order(name, build, cpu_engine):
    rng = np.random.default_rng(5)
    a, b = build(), build()
    a.adam_overlap_bytes, b.adam_overlap_bytes = 0, 16384          # b: 4096-element buckets
    n = 3
    x = rng.random((n,) + _user_shape(a, a.graph.inputs[0].shape), dtype=np.float32)
    ts, ls = _targets(a, rng, n)
    for m in (a, b):
        m.compile(loss=ls if len(ls) > 1 else ls[0], optimizer=Adam(5e-3))
    b.set_weight_dict(a.get_weight_dict())
    tg = ts if len(ts) > 1 else ts[0]
    for step in range(3):
        la, lb = a.train_on_batch(x, tg), b.train_on_batch(x, tg)
        assert abs(la - lb) <= 1e-9 * max(1.0, abs(la)), (step, la, lb)
    ea, eb = a._engine(n, True), b._engine(n, True)
    assert len(ea.planner.ops[2]) == 1 and len(eb.planner.ops[2]) >= 3, (len(ea.planner.ops[2]), len(eb.planner.ops[2]))
    wa, wb = a.get_weight_dict(), b.get_weight_dict()
    for k in wa:
        assert np.allclose(wa[k], wb[k], rtol=1e-9, atol=1e-12), k
    for arena in ("m", "v"):
        assert torch.allclose(getattr(ea, arena), getattr(eb, arena), rtol=1e-9, atol=1e-14), arena
