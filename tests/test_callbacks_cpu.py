"""Keras-2 semantics of the three callbacks the reference's Train.py builds (TensorFlow/2DCNN/Train.py:372-391), driven the way
b2seg.model.Model.fit drives them; a scripted stand-in for the model supplies the metric sequence."""
import numpy as np
import pytest

from b2seg.callbacks import EarlyStopping, ModelCheckpoint, ReduceLROnPlateau


class _Opt:
    learning_rate = 1e-3


class _FakeModel:
    def __init__(self):
        self.stop_training = False
        self.optimizer = _Opt()
        self.w = [np.zeros(2)]
        self.saved = []

    def get_weights(self):
        return [w.copy() for w in self.w]

    def set_weights(self, ws):
        self.w = [w.copy() for w in ws]

    def save_weights(self, path):
        self.saved.append(path)


def _drive(cbs, values, key="val_loss"):
    """the loop of Model.fit: one on_epoch_end per epoch until a callback stops training"""
    m = _FakeModel()
    for cb in cbs:
        cb.set_model(m)
        cb.on_train_begin({})
    epochs = 0
    for ep, v in enumerate(values):
        m.w = [np.full(2, float(ep))]            # "weights after epoch ep"
        logs = {key: v}
        for cb in cbs:
            cb.on_epoch_end(ep, logs)
        epochs += 1
        if m.stop_training:
            break
    for cb in cbs:
        cb.on_train_end({})
    return m, epochs


def test_early_stopping_patience_and_restore():
    es = EarlyStopping(monitor="val_loss", patience=2, restore_best_weights=True)
    m, n = _drive([es], [1.0, 0.8, 0.9, 0.85, 0.7, 0.6])
    assert n == 4 and es.stopped_epoch == 3 and es.best == 0.8 and es.best_epoch == 1    # two epochs without beating 0.8
    assert float(m.w[0][0]) == 1.0                                                       # weights of epoch 1 restored
    # max mode is inferred from an accuracy-like monitor; min_delta must be beaten
    es = EarlyStopping(monitor="val_accuracy", patience=1, min_delta=0.05)
    assert es.mode == "max"
    _, n = _drive([es], [0.5, 0.54, 0.56], key="val_accuracy")
    assert n == 2                                                                        # +0.04 is not an improvement
    # patience 0 never stops on epoch 0 (Keras: `epoch > 0`)
    _, n = _drive([EarlyStopping(patience=0)], [1.0, 2.0, 3.0])
    assert n == 2
    with pytest.warns(UserWarning):
        _drive([EarlyStopping(monitor="missing")], [1.0])


def test_model_checkpoint_best_only_and_formatting():
    ck = ModelCheckpoint("run/w_{epoch:02d}_{val_loss:.2f}.keras", monitor="val_loss", save_best_only=True, mode="min")
    m, _ = _drive([ck], [1.0, 1.2, 0.9, 0.95])
    assert m.saved == ["run/w_01_1.00.keras", "run/w_03_0.90.keras"] and ck.best == 0.9
    m, _ = _drive([ModelCheckpoint("w.keras")], [1.0, 1.2])
    assert m.saved == ["w.keras", "w.keras"]                                             # every epoch without save_best_only
    assert ModelCheckpoint("x", monitor="val_acc").mode == "max"


def test_reduce_lr_on_plateau_cooldown_and_floor():
    rl = ReduceLROnPlateau(monitor="val_loss", factor=0.5, patience=2, min_delta=1e-4, cooldown=1, min_lr=2e-4)
    m = _FakeModel()
    rl.set_model(m)
    rl.on_train_begin({})
    lrs = []
    for ep, v in enumerate([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]):
        logs = {"val_loss": v}
        rl.on_epoch_end(ep, logs)
        assert "lr" in logs
        lrs.append(m.optimizer.learning_rate)
    # epoch 0 sets the best; epochs 1-2 wait -> halve at epoch 2; the cooldown epoch 3 ends the cooldown and (Keras' order of the
    # two tests) already counts as a waiting epoch, so the next halving comes at epoch 4, then epoch 6 hits the floor 2e-4
    assert np.allclose(lrs, [1e-3, 1e-3, 5e-4, 5e-4, 2.5e-4, 2.5e-4, 2e-4, 2e-4, 2e-4])
    with pytest.raises(ValueError):
        ReduceLROnPlateau(factor=1.0)
    assert ReduceLROnPlateau(monitor="val_acc").mode == "max"


# ---- deep-supervision target pyramids (b2seg.helpers; the device version is checked in tests/test_gpu_zz_self_onn.py) ----------------
def test_prepare_train_dict_matches_reference_semantics():
    """2D: MaxPooling2D(2^i) of the mask for 'UNet', the mask for 'UNetPP' (helper_functions.py:359-380);
    1D: window mean, restated as the notebook's explicit loops (1D_Segmentation.ipynb cell 31)"""
    import torch
    import torch.nn.functional as F
    from b2seg.helpers import derive_targets_host, prepareTrainDict, prepareTrainDict1D
    rng = np.random.default_rng(0)
    m = (rng.random((3, 32, 48)) > 0.7).astype(np.float32)          # rank 3: a channel axis is appended (:367-368)
    d = prepareTrainDict(m, 3, "UNet")
    assert list(d) == ["out", "level1", "level2", "level3"] and d["out"].shape == (3, 32, 48, 1)
    for i in (1, 2, 3):
        want = F.max_pool2d(torch.from_numpy(m)[:, None], 2 ** i)[:, 0, :, :, None].numpy()
        assert np.array_equal(d[f"level{i}"], want)
    dpp = prepareTrainDict(m, 2, "UNetPP")
    assert all(np.array_equal(dpp[k], dpp["out"]) for k in ("level1", "level2"))
    with pytest.raises(KeyError):
        prepareTrainDict(m, 2, "FPN")
    y = rng.standard_normal((4, 64, 2))
    d1 = prepareTrainDict1D(y, 3, 64, "UNet", num_channel=2)
    for i in (1, 2, 3):
        w = 2 ** i
        want = np.zeros((4, 64 // w, 2))
        for j in range(2):
            for s in range(0, 64, w):
                want[:, s // w, j] = np.mean(y[:, s:s + w, j], axis=1)
        assert np.allclose(d1[f"level{i}"], want, atol=1e-12)
    # what Model.evaluate / the device path derive from a bare mask: same arrays, picked by output shape
    mask = d["out"]
    got = derive_targets_host(mask, [(3, 32, 48, 1), (3, 16, 24, 1), (3, 32, 48, 1), (3, 4, 6, 1)], 2)
    assert np.array_equal(got[1], d["level1"]) and np.array_equal(got[3], d["level3"]) and got[0] is mask and got[2] is mask
    with pytest.raises(ValueError):
        derive_targets_host(mask, [(3, 5, 48, 1)], 2)


def test_compile_ds_targets_target_lists():
    """compile(ds_targets=...): a bare mask becomes [mask, None, ...] for the device path and a full host pyramid for evaluate()"""
    from b2seg.models2d import unet_model_builder
    m = unet_model_builder("UNet", 32, 32, 8, 2, ds=1, train_mode="from_scratch").ResNet50()
    m.compile(loss="mse", optimizer="adam", ds_targets="UNet")
    mask = (np.random.default_rng(1).random((2, 32, 32, 1)) > 0.5).astype(np.float32)
    dev = m._targets(mask)
    assert dev[0].shape == (2, 32, 32, 1) and dev[1:] == [None, None]
    host = m._targets(mask, host=True)
    assert [t.shape for t in host] == [(2, 32, 32, 1), (2, 16, 16, 1), (2, 8, 8, 1)]     # outputs: out, level1 (/2), level2 (/4)
    from b2seg.helpers import prepareTrainDict
    d = prepareTrainDict(mask, 2, "UNet")
    assert all(np.array_equal(h, d[n]) for h, n in zip(host, m.output_names))
    # explicit dicts / lists keep working unchanged
    assert all(t is not None for t in m._targets(d))
    with pytest.raises(ValueError):
        m.compile(loss="mse", ds_targets="pyramid")


def test_fit_host_batching_with_ds_targets(monkeypatch):
    """fit()'s host-side slicing with compile(ds_targets=...): only the mask travels (the device derives the other targets) through
    the full batches, the ragged last batch and the validation split; evaluate() rebuilds the pyramid on the host.  The device
    side is replaced by recorders here (the real path is exercised by tests/test_gpu_zz_self_onn.py and tests/test_facade_cpu.py)."""
    from b2seg.models2d import unet_model_builder
    from b2seg.planner import LOSS_BUF_FLOATS
    m = unet_model_builder("UNet", 32, 32, 8, 2, ds=1, train_mode="from_scratch").ResNet50()
    m.compile(loss={"out": "bce", "level1": "mse", "level2": "mse"}, optimizer="adam", ds_targets="UNet")
    seen = {"stream": [], "eval": []}

    def fake_stream(batches):
        recs = []
        for bx, by in batches:
            seen["stream"].append((bx, by, m._targets(by)))
            rec = np.zeros(LOSS_BUF_FLOATS)
            rec[0] = 0.5 if bx.shape[0] == 8 else 0.25
            recs.append((bx.shape[0], rec))
        return recs

    def fake_evaluate(x, y=None, batch_size=32, **kw):
        vx, vy = x
        seen["eval"].append((vx, m._targets(vy, host=True)))
        return {"loss": 0.125} if kw.get("return_dict") else 0.125
    monkeypatch.setattr(m, "_train_stream", fake_stream)
    monkeypatch.setattr(m, "evaluate", fake_evaluate)
    rng = np.random.default_rng(3)
    x = rng.random((25, 32, 32, 3), dtype=np.float32)
    mask = (rng.random((25, 32, 32, 1)) > 0.5).astype(np.float32)
    h = m.fit(x, mask, batch_size=8, epochs=1, shuffle=False, verbose=0, validation_split=0.2)
    # 25 samples, 20 % held out -> 20 for training: two full batches of 8 and a ragged one of 4
    assert [bx.shape[0] for bx, _, _ in seen["stream"]] == [8, 8, 4]
    for bi, (bx, by, bys) in enumerate(seen["stream"]):
        assert np.array_equal(bx, x[:20][8 * bi:8 * bi + 8]) and np.array_equal(by, mask[:20][8 * bi:8 * bi + 8])
        assert np.array_equal(bys[0], mask[:20][8 * bi:8 * bi + 8]) and bys[1:] == [None, None]
    (vx, vt), = seen["eval"]
    assert vx.shape == (5, 32, 32, 3) and [t.shape for t in vt] == [(5, 32, 32, 1), (5, 16, 16, 1), (5, 8, 8, 1)]
    # Keras weights the batch losses by their sample counts
    assert np.array_equal(vt[0], mask[20:]) and h.history["loss"] == [pytest.approx((8 * 0.5 + 8 * 0.5 + 4 * 0.25) / 20)] and h.history["val_loss"] == [0.125]
    # without ds_targets the same call is refused: three outputs need three target arrays
    m.compile(loss={"out": "bce", "level1": "mse", "level2": "mse"}, optimizer="adam")
    monkeypatch.setattr(m, "_train_stream", lambda batches: [(bx.shape[0], m._targets(by)) for bx, by in batches])
    with pytest.raises(ValueError, match="expected 3 target arrays"):
        m.fit(x, mask, batch_size=8, epochs=1, verbose=0)


# ---- Keras weight-file layouts (b2seg.keras_io) on an h5py-shaped fake: the build image has no HDF5 library ---------------------------
class _FakeGroup(dict):
    """quacks like an h5py group: mapping with .attrs, nested paths through '/'"""
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.attrs = {}

    def __getitem__(self, key):
        node = self
        for part in str(key).split("/"):
            node = dict.__getitem__(node, part)
        return node

    def put(self, path, value):
        node = self
        parts = path.split("/")
        for part in parts[:-1]:
            if part not in node.keys():
                dict.__setitem__(node, part, _FakeGroup())
            node = dict.__getitem__(node, part)
        dict.__setitem__(node, parts[-1], value)


def test_keras_weight_file_layouts():
    from b2seg.keras_io import select_for_model, weight_order, weights_from_tree
    from b2seg.models2d import unet_model_builder
    m = unet_model_builder("UNet", 16, 16, 8, 1, train_mode="from_scratch").ResNet50()
    specs = m.graph.param_specs()
    w = m.get_weight_dict()
    order = weight_order(specs)
    assert order["conv2d"] == ["kernel", "bias"] and order["batch_normalization"] == ["gamma", "beta", "moving_mean", "moving_variance"]
    # Keras-2 save_weights layout, plain and nested under model_weights (model.save)
    legacy = _FakeGroup()
    legacy.attrs["layer_names"] = [l.encode() for l in order] + [b"max_pooling2d"]
    for layer, names in order.items():
        legacy.put(layer, _FakeGroup())
        legacy[layer].attrs["weight_names"] = [f"{layer}/{n}:0".encode() for n in names]
        for n in names:
            legacy[layer].put(f"{layer}/{n}:0", w[f"{layer}/{n}"])
    legacy.put("max_pooling2d", _FakeGroup())                      # parameter-free layers have empty groups
    legacy["max_pooling2d"].attrs["weight_names"] = []
    for root in (legacy, _FakeGroup(model_weights=legacy)):
        got, extra = select_for_model(weights_from_tree(root, order), specs)
        assert extra == [] and set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
    # .keras (v3) layout: unnamed variables in creation order
    v3 = _FakeGroup()
    for layer, names in order.items():
        for i, n in enumerate(names):
            v3.put(f"layers/{layer}/vars/{i}", w[f"{layer}/{n}"])
    v3.put("layers/max_pooling2d/vars", _FakeGroup())
    got, _ = select_for_model(weights_from_tree(v3, order), specs)
    assert all(np.array_equal(got[k], w[k]) for k in w)
    # errors: a missing weight, a wrong shape, an unknown layout, a variable-count mismatch
    broken = dict(got)
    del broken["out/bias"]
    with pytest.raises(ValueError, match="lacks 1 of"):
        select_for_model(broken, specs)
    with pytest.raises(ValueError, match="shape"):
        select_for_model(dict(got, **{"out/bias": np.zeros(3, np.float32)}), specs)
    with pytest.raises(ValueError, match="not a Keras weight file"):
        weights_from_tree(_FakeGroup(), order)
    v3.put("layers/conv2d/vars/2", np.zeros(1))
    with pytest.raises(ValueError, match="holds 3 variables"):
        weights_from_tree(v3, order)
    # without h5py the facade says what to do instead of failing obscurely
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(NotImplementedError, match="keras_weights_to_npz"):
            m.load_weights("/nonexistent/best.h5")


def test_validation_metrics_of_the_reference_configuration(monkeypatch):
    """The shipped Train_Configs.ini compiles with metrics=[MeanSquaredError] and lets its callbacks monitor val_mean_squared_error
    (Train_Configs.ini:36,44; Train.py:324,372-391): fit() must log that key.  Metrics are computed on the host from predict()."""
    from b2seg.models2d import unet_model_builder

    class MeanSquaredError:            # stands in for tf.keras.metrics.MeanSquaredError(name='mean_squared_error')
        name = "mean_squared_error"
    rng = np.random.default_rng(5)
    x = rng.random((6, 16, 16, 3), dtype=np.float32)
    y = (rng.random((6, 16, 16, 1)) > 0.5).astype(np.float32)
    pred = rng.random((6, 16, 16, 1)).astype(np.float32)
    m = unet_model_builder("UNet", 16, 16, 8, 1, train_mode="from_scratch").ResNet50()
    m.compile(loss="binary_crossentropy", optimizer="adam", metrics=[MeanSquaredError(), "BinaryAccuracy", "acc", "AUC"])
    monkeypatch.setattr(m, "predict", lambda x_, batch_size=None, **kw: pred[:len(x_)])
    logs = m.evaluate(x, y, return_dict=True)
    assert set(logs) == {"loss", "mean_squared_error", "binary_accuracy", "accuracy"}          # AUC is not a host-side metric: left out
    assert logs["mean_squared_error"] == pytest.approx(float(((pred.astype(np.float64) - y) ** 2).mean()))
    assert logs["binary_accuracy"] == logs["accuracy"] == pytest.approx(float(((pred > 0.5) == (y > 0.5)).mean()))
    assert isinstance(m.evaluate(x, y), float) and m.evaluate(x, y) == pytest.approx(logs["loss"])
    from b2seg.planner import LOSS_BUF_FLOATS
    monkeypatch.setattr(m, "_train_stream", lambda batches: [(bx.shape[0], np.full(LOSS_BUF_FLOATS, 0.7)) for bx, _ in batches])
    from b2seg.callbacks import EarlyStopping
    es = EarlyStopping(monitor="val_mean_squared_error", patience=0)
    h = m.fit(x[:4], y[:4], batch_size=2, epochs=3, verbose=0, validation_data=(x[4:], y[4:]), callbacks=[es], shuffle=False)
    assert "val_mean_squared_error" in h.history and "val_loss" in h.history
    assert len(h.history["loss"]) == 2                      # constant predictions: no improvement in epoch 2 -> stopped by the monitor
    # several outputs: Keras prefixes the output name
    m2 = unet_model_builder("UNet", 16, 16, 8, 1, ds=1, train_mode="from_scratch").ResNet50()
    m2.compile(loss={"out": "bce", "level1": "mse"}, optimizer="adam", metrics=["mse"], ds_targets="UNet")
    monkeypatch.setattr(m2, "predict", lambda x_, batch_size=None, **kw: [pred[:len(x_)], pred[:len(x_), ::2, ::2]])
    logs2 = m2.evaluate(x, y, return_dict=True)
    assert set(logs2) == {"loss", "out_loss", "level1_loss", "out_mean_squared_error", "level1_mean_squared_error"}
