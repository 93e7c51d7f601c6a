"""Keras-2 callback semantics for `b2seg.model.Model.fit` — the three callbacks the reference's training script builds
(`from keras.callbacks import EarlyStopping, ModelCheckpoint, ReduceLROnPlateau`, TensorFlow/2DCNN/Train.py:8, :372-391):
same constructor arguments, same monitor / mode / patience / min_delta / cooldown rules as tf.keras 2.15, driven by the hooks
`Model.fit` calls (`set_model`, `on_train_begin`, `on_epoch_end(epoch, logs)`, `on_train_end`).

They only need the Keras `Model` protocol: `model.stop_training`, `model.optimizer.learning_rate`, `model.get_weights()` /
`model.set_weights()`, `model.save_weights(path)` (weights are written as `.npz` keyed by Keras layer names: reading / writing
`.keras` / `.h5` containers needs h5py, which this image does not have — SURVEY 8(f) rank 1).
"""
from __future__ import annotations

import warnings

import numpy as np


class Callback:
    def __init__(self):
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass


def _monitor_mode(monitor: str, mode: str, acc_like) -> str:
    if mode not in ("auto", "min", "max"):
        warnings.warn(f"mode {mode} is unknown, fallback to auto mode.")
        mode = "auto"
    if mode == "auto":
        mode = "max" if acc_like(monitor) else "min"
    return mode


class EarlyStopping(Callback):
    """tf.keras.callbacks.EarlyStopping (Train.py:372-374)."""

    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", baseline=None,
                 restore_best_weights=False, start_from_epoch=0):
        super().__init__()
        self.monitor, self.patience, self.verbose, self.baseline = monitor, patience, verbose, baseline
        self.min_delta = abs(min_delta)
        self.restore_best_weights, self.start_from_epoch = restore_best_weights, start_from_epoch
        self.mode = _monitor_mode(monitor, mode, lambda m: m.endswith("acc") or m.endswith("accuracy") or m.endswith("auc"))
        self.monitor_op = np.greater if self.mode == "max" else np.less
        self.min_delta *= 1 if self.mode == "max" else -1
        self.wait = self.stopped_epoch = self.best_epoch = 0
        self.best, self.best_weights = None, None

    def on_train_begin(self, logs=None):
        self.wait = self.stopped_epoch = self.best_epoch = 0
        self.best = np.inf if self.mode == "min" else -np.inf
        self.best_weights = None

    def _is_improvement(self, value, reference):
        return bool(self.monitor_op(value - self.min_delta, reference))

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            warnings.warn(f"Early stopping conditioned on metric `{self.monitor}` which is not available. "
                          f"Available metrics are: {','.join(list((logs or {}).keys()))}")
            return
        if epoch < self.start_from_epoch:
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = self.model.get_weights()
        self.wait += 1
        if self._is_improvement(current, self.best):
            self.best, self.best_epoch = current, epoch
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
            # only restart the wait if we beat both the baseline and our previous best
            if self.baseline is None or self._is_improvement(current, self.baseline):
                self.wait = 0
            return
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True
            if self.restore_best_weights and self.best_weights is not None:
                if self.verbose > 0:
                    print(f"Restoring model weights from the end of the best epoch: {self.best_epoch + 1}.")
                self.model.set_weights(self.best_weights)

    def on_train_end(self, logs=None):
        if self.stopped_epoch > 0 and self.verbose > 0:
            print(f"Epoch {self.stopped_epoch + 1}: early stopping")


class ModelCheckpoint(Callback):
    """tf.keras.callbacks.ModelCheckpoint with save_freq='epoch' (Train.py:375-379: best-only checkpoint of the fold)."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False, mode="auto",
                 save_freq="epoch", initial_value_threshold=None):
        super().__init__()
        if save_freq != "epoch":
            raise ValueError("only save_freq='epoch' is supported")
        self.filepath, self.monitor, self.verbose = str(filepath), monitor, verbose
        self.save_best_only, self.save_weights_only = save_best_only, save_weights_only
        self.mode = _monitor_mode(monitor, mode, lambda m: "acc" in m or m.startswith("fmeasure"))
        self.monitor_op = np.greater if self.mode == "max" else np.less
        self.best = initial_value_threshold
        if self.best is None:
            self.best = -np.inf if self.mode == "max" else np.inf
        self.last_saved = None

    def _save(self, epoch, logs):
        path = self.filepath.format(epoch=epoch + 1, **(logs or {}))
        self.model.save_weights(path)
        self.last_saved = path

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        if not self.save_best_only:
            if self.verbose > 0:
                print(f"\nEpoch {epoch + 1}: saving model to {self.filepath}")
            self._save(epoch, logs)
            return
        current = logs.get(self.monitor)
        if current is None:
            warnings.warn(f"Can save best model only with {self.monitor} available, skipping.")
            return
        if self.monitor_op(current, self.best):
            if self.verbose > 0:
                print(f"\nEpoch {epoch + 1}: {self.monitor} improved from {self.best:.5f} to {current:.5f}, saving model to {self.filepath}")
            self.best = current
            self._save(epoch, logs)
        elif self.verbose > 0:
            print(f"\nEpoch {epoch + 1}: {self.monitor} did not improve from {self.best:.5f}")


class ReduceLROnPlateau(Callback):
    """tf.keras.callbacks.ReduceLROnPlateau (Train.py:380-387)."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", min_delta=1e-4, cooldown=0, min_lr=0, **kwargs):
        super().__init__()
        if factor >= 1.0:
            raise ValueError(f"ReduceLROnPlateau does not support a factor >= 1.0. Got {factor}")
        if "epsilon" in kwargs:
            min_delta = kwargs.pop("epsilon")
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.min_delta, self.cooldown, self.min_lr = min_delta, cooldown, min_lr
        self.mode = _monitor_mode(monitor, mode, lambda m: "acc" in m)
        self._reset()

    def _reset(self):
        if self.mode == "min":
            self.monitor_op = lambda a, b: np.less(a, b - self.min_delta)
            self.best = np.inf
        else:
            self.monitor_op = lambda a, b: np.greater(a, b + self.min_delta)
            self.best = -np.inf
        self.cooldown_counter = 0
        self.wait = 0

    def in_cooldown(self):
        return self.cooldown_counter > 0

    def on_train_begin(self, logs=None):
        self._reset()

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        logs["lr"] = float(self.model.optimizer.learning_rate)
        current = logs.get(self.monitor)
        if current is None:
            warnings.warn(f"Learning rate reduction is conditioned on metric `{self.monitor}` which is not available. "
                          f"Available metrics are: {','.join(list(logs.keys()))}")
            return
        if self.in_cooldown():
            self.cooldown_counter -= 1
            self.wait = 0
        if self.monitor_op(current, self.best):
            self.best = current
            self.wait = 0
        elif not self.in_cooldown():
            self.wait += 1
            if self.wait >= self.patience:
                old_lr = float(self.model.optimizer.learning_rate)
                if old_lr > np.float32(self.min_lr):
                    new_lr = max(old_lr * self.factor, self.min_lr)
                    self.model.optimizer.learning_rate = new_lr
                    if self.verbose > 0:
                        print(f"\nEpoch {epoch + 1}: ReduceLROnPlateau reducing learning rate to {new_lr}.")
                    self.cooldown_counter = self.cooldown
                    self.wait = 0
