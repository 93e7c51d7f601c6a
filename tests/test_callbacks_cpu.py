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
