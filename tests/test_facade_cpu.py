"""The Keras-Model facade driven the way the reference's scripts drive it (2DCNN/Train.py:216, 281-300, 320-325, 354-415), on the float64
emulator engine (tests/cpu_engine.py): data sources, Adam bookkeeping across engines and recompiles, partial weight updates, logs."""
import numpy as np
import pytest
import torch

import b2seg.engine
from b2seg.model import Adam
from b2seg.models2d import unet_model_builder
from cpu_engine import CpuEngine
from oracle.keras_ref import KerasRef, keras_adam_step, keras_loss
from oracle.ref_models import Ref2D


@pytest.fixture()
def cpu_engine(monkeypatch):
    monkeypatch.setattr(b2seg.engine, "Engine", CpuEngine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)


KW = dict(num_channels=2, output_nums=1, ds=1)


def _model(lr=2e-3, **compile_kw):
    m = unet_model_builder("UNet", 16, 16, 8, 2, train_mode="from_scratch", **KW).ResNet50()
    m.compile(loss={"out": "binary_crossentropy", "level1": "mse", "level2": "mse"}, optimizer=Adam(lr), **compile_kw)
    return m


class Gen:
    """what 2DCNN/utils/DataGenerator.py:CustomDataGenerator is to Keras: __len__, __getitem__ -> (x, dict of targets), on_epoch_end"""

    def __init__(self, x, ys, bs):
        self.x, self.ys, self.bs, self.epochs_seen = x, ys, bs, 0

    def __len__(self):
        return -(-self.x.shape[0] // self.bs)

    def __getitem__(self, i):
        s = slice(i * self.bs, (i + 1) * self.bs)
        return self.x[s], {k: v[s] for k, v in self.ys.items()}

    def on_epoch_end(self):
        self.epochs_seen += 1


def _data(n, rng):
    x = rng.random((n, 16, 16, 2), dtype=np.float32)
    ys = {"out": (x[..., :1] > 0.5).astype(np.float32), "level1": rng.standard_normal((n, 8, 8, 1)).astype(np.float32),
          "level2": rng.standard_normal((n, 4, 4, 1)).astype(np.float32)}
    return x, ys


def test_fit_sequence_with_sequence_validation_and_dict_targets(cpu_engine):
    """Train.py:281-300: model.fit(train_ds, validation_data=val_ds, ...) with Sequence objects whose items carry dict targets; the
    ragged last batch of the training Sequence runs through its own engine; same result as the array form"""
    rng = np.random.default_rng(0)
    x, ys = _data(10, rng)
    vx, vys = _data(6, rng)
    a, b = _model(metrics=["accuracy"]), _model(metrics=["accuracy"])
    b.set_weight_dict(a.get_weight_dict())
    tr, va = Gen(x, ys, 4), Gen(vx, vys, 4)
    ha = a.fit(tr, validation_data=va, epochs=2, verbose=0)
    hb = b.fit(x, ys, batch_size=4, validation_data=(vx, vys), epochs=2, shuffle=False, verbose=0)
    assert tr.epochs_seen == 2 and va.epochs_seen == 2
    keys = {"loss", "out_loss", "level1_loss", "level2_loss", "out_accuracy", "level1_accuracy", "level2_accuracy"}
    assert set(ha.history) == keys | {"val_" + k for k in keys}
    for k in ha.history:
        assert np.allclose(ha.history[k], hb.history[k], rtol=1e-9, atol=1e-12), k
    wa, wb = a.get_weight_dict(), b.get_weight_dict()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)


def test_fit_generators_and_zip(cpu_engine):
    """Train.py:216: zip(image_generator, mask_generator) with steps_per_epoch, validation from another generator with validation_steps;
    validation_split with a generator raises like Keras"""
    rng = np.random.default_rng(1)
    x, ys = _data(12, rng)

    def forever(a, bs):
        while True:
            for s in range(0, a.shape[0], bs):
                yield a[s:s + bs]
    m = unet_model_builder("UNet", 16, 16, 8, 2, train_mode="from_scratch", num_channels=2, output_nums=1).ResNet50()
    m.compile(loss="binary_crossentropy", optimizer=Adam(2e-3), metrics=["accuracy"])
    ref = unet_model_builder("UNet", 16, 16, 8, 2, train_mode="from_scratch", num_channels=2, output_nums=1).ResNet50()
    ref.compile(loss="binary_crossentropy", optimizer=Adam(2e-3), metrics=["accuracy"])
    ref.set_weight_dict(m.get_weight_dict())
    h = m.fit(zip(forever(x, 4), forever(ys["out"], 4)), steps_per_epoch=3, epochs=2, verbose=0,
              validation_data=zip(forever(x, 4), forever(ys["out"], 4)), validation_steps=2)
    assert set(h.history) == {"loss", "accuracy", "val_loss", "val_accuracy"} and len(h.history["loss"]) == 2
    # the generator keeps running across epochs: 6 steps over batches 0,1,2,0,1,2
    losses = [ref.train_on_batch(x[s:s + 4], ys["out"][s:s + 4]) for s in (0, 4, 8, 0, 4, 8)]
    assert np.allclose(h.history["loss"], [np.mean(losses[:3]), np.mean(losses[3:])], rtol=1e-9)
    with pytest.raises(ValueError, match="validation_split"):
        m.fit(zip(forever(x, 4), forever(ys["out"], 4)), steps_per_epoch=1, validation_split=0.2, verbose=0)
    with pytest.raises(ValueError, match="steps_per_epoch"):
        m.fit(zip(forever(x, 4), forever(ys["out"], 4)), epochs=2, verbose=0)


def test_adam_counter_is_per_model_and_compile_resets_the_optimizer(cpu_engine):
    """ragged batches run on a second engine that shares weights and Adam moments with the first: one step counter for both.
    Recompiling gives a fresh optimizer (zero moments, t = 0) but keeps the weights, like Keras (Train.py:320-325 recompiles the
    same model for every fold).  Trajectory against the oracle's Keras-2 Adam."""
    rng = np.random.default_rng(2)
    x = rng.random((7, 16, 16, 2), dtype=np.float32)
    y = (x[..., :1] > 0.5).astype(np.float32)
    kw = dict(num_channels=2, output_nums=1)
    m = unet_model_builder("UNet", 16, 16, 8, 2, train_mode="from_scratch", **kw).ResNet50()
    m.compile(loss="binary_crossentropy", optimizer=Adam(1e-2))
    ref = Ref2D("UNet", 16, 16, 8, 2, **kw)
    tp = {k: torch.from_numpy(v.copy()).double() for k, v in m.get_weight_dict().items()}
    state = {"st": {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in tp.items()}, "t": 0}

    def oracle_step(bx, by):
        state["t"] += 1
        k = KerasRef(2, params=tp, dtype=torch.float64, training=True, strict=True)
        out = ref(k, torch.from_numpy(bx).double())[0]
        loss = keras_loss("bce", out, torch.from_numpy(by).double(), logits=k.logits["out"])
        loss.backward()
        with torch.no_grad():
            for key in k.trainable:
                if tp[key].grad is not None:
                    keras_adam_step(tp[key], tp[key].grad, state["st"][key][0], state["st"][key][1], state["t"], lr=1e-2)
                    tp[key].grad = None
            for key, v in k.new_moving.items():
                tp[key] = v
        return float(loss)

    got, want = [], []
    for _ in range(2):                                   # two epochs of batches 4 + 3 (ragged: second engine)
        for s in (0, 4):
            got.append(m.train_on_batch(x[s:s + 4], y[s:s + 4]))
            want.append(oracle_step(x[s:s + 4], y[s:s + 4]))
    assert m._adam_step == 4 and len(m._engines) == 2
    m.compile(loss="binary_crossentropy", optimizer=Adam(1e-2))     # fresh optimizer
    state["st"] = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in tp.items()}
    state["t"] = 0
    for s in (0, 4):
        got.append(m.train_on_batch(x[s:s + 4], y[s:s + 4]))
        want.append(oracle_step(x[s:s + 4], y[s:s + 4]))
    assert np.allclose(got, want, rtol=2e-5, atol=1e-7), (got, want)
    final = m.get_weight_dict()
    assert max(float(np.abs(final[key] - tp[key].detach().numpy()).max()) for key in tp) < 2e-5


def test_partial_weight_update_keeps_the_trained_weights(cpu_engine):
    rng = np.random.default_rng(3)
    x, ys = _data(4, rng)
    m = _model()
    m.train_on_batch(x, ys)
    trained = m.get_weight_dict()
    m._weights = {k: np.zeros_like(v) for k, v in m._weights.items()}      # a stale host copy must not come back
    new_bias = np.full_like(trained["out/bias"], 0.25)
    m.set_weight_dict({"out/bias": new_bias})
    after = m.get_weight_dict()
    assert np.array_equal(after["out/bias"], new_bias)
    assert all(np.array_equal(after[k], trained[k]) for k in trained if k != "out/bias")
