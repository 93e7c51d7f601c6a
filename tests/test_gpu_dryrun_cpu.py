"""Dry run (CPU) of GPU tests that could not be run on hardware yet: the same test functions, with b2seg.engine.Engine replaced by
the float64 emulator engine (tests/cpu_engine.py).  Catches mistakes in the tests' own Python — names, shapes, targets, oracle
plumbing — before they cost a GPU run; says nothing about the kernels (in float64 the tolerances are met by ten orders of magnitude)."""
import pytest
import torch

import b2seg.engine
from cpu_engine import CpuEngine


@pytest.fixture()
def cpu_engine(monkeypatch):
    monkeypatch.setattr(b2seg.engine, "Engine", CpuEngine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)


def test_dryrun_2d_self_onn_per_layer(cpu_engine):
    from test_gpu_zz_self_onn import SELF_CASES, test_2d_self_onn_per_layer
    for dec, kw, _size, _width, _depth in SELF_CASES:
        test_2d_self_onn_per_layer(dec, kw, 32, 8, 2)          # reduced sizes: the emulator is a Python interpreter of descriptors


def test_dryrun_1d_self_onn_per_layer(cpu_engine):
    from test_gpu_zz_self_onn import test_1d_self_onn_per_layer
    for var, kw in (("SelfUNetPP", dict(ds=1)), ("SelfR2UNetPP", dict(ds=1, t=2, q=2)), ("SelfUNet3P", dict(ds=1, q=2))):
        test_1d_self_onn_per_layer(var, kw)


def test_dryrun_wide_head_and_ds_targets(cpu_engine, monkeypatch):
    import test_gpu_zz_self_onn as z
    z.test_wide_softmax_head_per_layer()
    # the model half of test_ds_target_pyramid_on_device (its first half calls the CUDA kernel directly)
    import numpy as np
    from b2seg.helpers import prepareTrainDict
    from b2seg.model import Adam
    from b2seg.models2d import unet_model_builder
    rng = np.random.default_rng(23)
    x = rng.random((2, 32, 32, 3), dtype=np.float32)
    mask = (rng.random((2, 32, 32, 1)) > 0.6).astype(np.float32)
    losses = {}
    for how in ("host", "device"):
        m = unet_model_builder("UNet", 32, 32, 8, 2, ds=1, train_mode="from_scratch").ResNet50()
        m.compile(loss={"out": "binary_crossentropy", "level1": "mse", "level2": "mse"}, optimizer=Adam(1e-3),
                  ds_targets="UNet" if how == "device" else None)
        losses[how] = m.train_on_batch(x, mask if how == "device" else prepareTrainDict(mask, 2, "UNet"))
    assert abs(losses["host"] - losses["device"]) < 1e-12


def test_dryrun_existing_end_to_end_tests(cpu_engine):
    """two tests that HAVE passed on the B200, replayed here as a check of the dry-run harness itself (and of later planner changes)"""
    import test_gpu_model as g
    g.test_unet1d_shallow_end_to_end()
    g.test_unet2d_multiclass_mse_end_to_end()


@pytest.mark.parametrize("case", ["self2d", "wide", "self1d"])
def test_inference_plans_match_the_oracle(cpu_engine, case):
    """predict() (moving statistics, no backward) of the families added last — Self-ONN heads are Activations, wide heads run on
    the convolution kernels — against the oracle in inference mode, through the facade and the emulator engine; weights perturbed so
    that BatchNorm's moving statistics and every bias matter"""
    import numpy as np
    from b2seg.models1d import UNet
    from b2seg.models2d import unet_model_builder
    from oracle.keras_ref import KerasRef
    from oracle.ref_models import Ref1D, Ref2D
    rng = np.random.default_rng(31)
    if case == "self2d":
        kw = dict(num_channels=2, ds=1, q=3)
        m = unet_model_builder("SelfUNet", 16, 16, 8, 2, train_mode="from_scratch", **kw).ResNet50()
        ref, ndim, x = Ref2D("SelfUNet", 16, 16, 8, 2, **kw), 2, 0.5 * rng.random((3, 16, 16, 2), dtype=np.float32)
    elif case == "wide":
        kw = dict(num_channels=2, ds=1, output_nums=10, final_activation="softmax")
        m = unet_model_builder("UNet", 16, 16, 8, 2, train_mode="from_scratch", **kw).ResNet50()
        ref, ndim, x = Ref2D("UNet", 16, 16, 8, 2, **kw), 2, rng.random((3, 16, 16, 2), dtype=np.float32)
    else:
        kw = dict(ds=1, t=2, q=2)
        m = UNet(32, 2, 2, 8, 3, **kw).SelfR2UNetPP()
        ref, ndim, x = Ref1D("SelfR2UNetPP", 32, 2, 2, 8, 3, **kw), 1, (0.5 * rng.standard_normal((3, 32, 2))).astype(np.float32)
    w = m.get_weight_dict()
    for k in w:
        if k.endswith(("/gamma", "/moving_variance")):
            w[k] = (w[k] * (1 + 0.3 * rng.random(w[k].shape))).astype(np.float32)
        elif k.endswith(("/beta", "/bias", "/moving_mean")):
            w[k] = (w[k] + 0.1 * rng.standard_normal(w[k].shape)).astype(np.float32)
    m.set_weight_dict(w)
    got = m.predict(x, batch_size=2)                    # 3 samples in batches of 2: the ragged tail is zero-padded and cut
    got = got if isinstance(got, list) else [got]
    k = KerasRef(ndim, params={kk: torch.from_numpy(v).double() for kk, v in w.items()}, dtype=torch.float64, training=False, strict=True)
    want = ref(k, torch.from_numpy(x).double())
    assert len(got) == len(want)
    for g_, w_ in zip(got, want):
        assert g_.shape == tuple(w_.shape) and np.allclose(g_, w_.detach().numpy(), atol=2e-6, rtol=1e-6), float(np.abs(g_ - w_.detach().numpy()).max())


@pytest.mark.parametrize("case", ["unetpp_ds_ag", "self_unet", "bcdunet1d"])
def test_three_adam_steps_track_the_oracle(cpu_engine, case):
    """train_on_batch x 3 through the facade on the emulator engine against the oracle's trajectory (Keras-2 Adam rule with its step
    counter, BatchNorm moving statistics carried from step to step, several weighted outputs): losses and final weights in float64"""
    import numpy as np
    from b2seg.model import Adam
    from b2seg.models1d import BCDUNet
    from b2seg.models2d import unet_model_builder
    from oracle.keras_ref import KerasRef, keras_adam_step, keras_loss
    from oracle.ref_models import Ref1D, Ref2D
    rng = np.random.default_rng(41)
    if case == "unetpp_ds_ag":
        kw = dict(num_channels=2, ds=1, ag=1, output_nums=3, final_activation="softmax")
        m, ref, ndim = unet_model_builder("UNetPP", 16, 16, 8, 2, train_mode="from_scratch", **kw).ResNet50(), Ref2D("UNetPP", 16, 16, 8, 2, **kw), 2
        x = rng.random((2, 16, 16, 2), dtype=np.float32)
    elif case == "self_unet":
        kw = dict(num_channels=2, ds=1, q=3)
        m, ref, ndim = unet_model_builder("SelfUNet", 16, 16, 8, 2, train_mode="from_scratch", **kw).ResNet50(), Ref2D("SelfUNet", 16, 16, 8, 2, **kw), 2
        x = 0.5 * rng.random((2, 16, 16, 2), dtype=np.float32)
    else:
        kw = dict(ds=1, lstm=1, ag=1, dense_loop=2)
        m, ref, ndim = BCDUNet(32, 2, 2, 16, 3, **kw).BCDUNet(), Ref1D("BCDUNet", 32, 2, 2, 16, 3, **kw), 1
        x = rng.standard_normal((2, 32, 2)).astype(np.float32)
    targets, losses = [], []
    for i, n in enumerate(m.graph.outputs):
        shp = (2,) + (tuple(n.shape) if ndim == 2 else tuple(n.shape[1:]))
        fn = n.attrs.get("activation") if n.op == "conv" else n.attrs.get("fn")
        if fn == "softmax":
            targets.append(np.eye(shp[-1], dtype=np.float32)[rng.integers(0, shp[-1], shp[:-1])]); losses.append("cce")
        elif fn == "sigmoid":
            targets.append((rng.random(shp) > 0.5).astype(np.float32)); losses.append("bce")
        else:
            targets.append(rng.standard_normal(shp).astype(np.float32)); losses.append("mse")
    lw = [1.0 - 0.15 * i for i in range(len(targets))]
    m.compile(loss=losses, optimizer=Adam(2e-3), loss_weights=lw)
    tp = {k: torch.from_numpy(v.copy()).double() for k, v in m.get_weight_dict().items()}
    st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in tp.items()}
    got, want = [], []
    for t in range(1, 4):
        got.append(m.train_on_batch(x, targets if len(targets) > 1 else targets[0]))
        k = KerasRef(ndim, params=tp, dtype=torch.float64, training=True, strict=True)
        outs = ref(k, torch.from_numpy(x).double())
        total = sum(w * keras_loss(kind, o, torch.from_numpy(tg).double(), logits=k.logits.get(name))
                    for w, kind, o, tg, name in zip(lw, losses, outs, targets, m.output_names))
        total.backward()
        want.append(float(total))
        with torch.no_grad():
            for key in k.trainable:
                if tp[key].grad is not None:
                    keras_adam_step(tp[key], tp[key].grad, st[key][0], st[key][1], t, lr=2e-3)
                    tp[key].grad = None
            for key, v in k.new_moving.items():
                tp[key] = v
    assert np.allclose(got, want, rtol=2e-5, atol=1e-7), (got, want)
    final = m.get_weight_dict()
    worst = max(float(np.abs(final[key] - tp[key].detach().numpy()).max()) for key in tp)
    # the emulator keeps float64 but the facade moves weights through float32 (as the device does): 1e-5 absolute
    assert worst < 2e-5, worst


def test_dryrun_fit_history_and_callbacks(cpu_engine, tmp_path, monkeypatch):
    """the fit()-level GPU tests (history keys incl. training and validation metrics, callbacks, validation_split, checkpoint files)"""
    import test_gpu_model as g
    g.test_fit_and_history_api()
    g.test_fit_with_reference_callbacks_and_validation_split(tmp_path)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    g.test_fit_pipelined_input_matches_train_on_batch()


def test_dryrun_every_model_test_of_the_gpu_suite(cpu_engine):
    """All train_on_batch / predict based tests of tests/test_gpu_model.py at their real sizes, on the emulator engine: after a planner
    or facade change, the Python half of the GPU suite is known to work before a GPU is spent on it (about 40 s)."""
    import test_gpu_model as g
    g.test_unet2d_shallow_end_to_end()
    g.test_unet2d_autoencoder_bottleneck_end_to_end()
    g.test_unet1d_autoencoder_bottleneck_end_to_end()
    g.test_training_reduces_loss_and_matches_oracle_trajectory()
    g.test_predict_uses_moving_statistics()
    g.test_1d_bcdunet_lstm_ag_ds_per_layer()
    for case in g.FAMILY_CASES:
        g.test_2d_families_per_layer(*case)
    for kw in (dict(), dict(ds=1, ag=1)):
        g.test_fpn_per_layer(kw)
    for var, kw in [("RUNet", dict(ds=1, t=2)), ("R2UNet", dict(ds=1, ag=1, t=2)), ("R2UNetPP", dict(ds=1, t=1)), ("R2UNet3P", dict(ds=1, t=1)),
                    ("UNet4P", dict(ds=1, ag=1)), ("MultiResUNet3P", dict(ds=1))]:
        g.test_1d_recurrent_unets_per_layer(var, kw)


def test_dryrun_round2_parity_tests(cpu_engine, monkeypatch):
    """the GPU tests added in round 2 (1D nested variants; BASELINE configs 3 / 4 / 5 — here at reduced size, the emulator being a
    Python interpreter of descriptors; float32 oracle path included)"""
    import test_gpu_model as g
    for var, kw in [("UNetE", dict(ds=1)), ("UNetP", dict(ds=1)), ("UNetPP", dict(ds=1)), ("UNetPP", dict(ds=1, ag=1, is_transconv=False)),
                    ("UNet3P", dict(ds=1)), ("MultiResUNet", dict(ds=1)), ("MultiResUNet", dict(ds=0, ag=1, alpha=1.5))]:
        g.test_1d_nested_unets_per_layer(var, kw)
    for (cid, dec, kw, _size, _width, _depth, batch, dtype) in g.BASELINE_SHAPE_CASES:
        g.test_baseline_configs_3_4_5_at_their_shapes_per_layer((cid, dec, kw, 32, 16, 2, 2, dtype), monkeypatch)


def test_dryrun_golden_fixture_replay(cpu_engine):
    """tests/test_gpu_golden.py on the emulator engine: the committed fixtures still load into the product and reproduce"""
    import test_gpu_golden as t
    for spec in t.CASES:
        t.test_training_step_matches_golden(spec)


def test_dryrun_smoke_entry(cpu_engine, monkeypatch, capsys):
    """__graft_entry__.smoke() (the driver's first GPU step of every round) on the emulator engine"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as entry
    monkeypatch.setattr(CpuEngine, "launches", [0, 0, 0], raising=False)
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
