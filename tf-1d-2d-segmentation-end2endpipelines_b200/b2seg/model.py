"""Keras-`Model`-shaped facade over the B200 engine: the object the builder classes return.

Mirrors the subset of the tf.keras.Model protocol the reference's callers use:
compile / fit / predict / train_on_batch / evaluate / load_weights / save_weights / get_weights / set_weights /
summary / count_params / trainable_weights / non_trainable_weights  (2DCNN/Train.py:322-415, Test.py:114-164,
1DCNN/1D_Segmentation.ipynb cells 35-41).  Inputs/outputs are NumPy float32 channels-last arrays; with deep
supervision `predict` returns the list [out, level1, ..., level_d] like Keras.
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional

import numpy as np

from .graph import Graph, init_params

_LOSS_ALIASES = {
    "binary_crossentropy": "bce", "bce": "bce", "binarycrossentropy": "bce",
    "categorical_crossentropy": "cce", "cce": "cce", "categoricalcrossentropy": "cce",
    "mean_squared_error": "mse", "mse": "mse", "meansquarederror": "mse",
    "mean_absolute_error": "mae", "mae": "mae", "meanabsoluteerror": "mae",
}


class Adam:
    """tf.keras.optimizers.Adam(learning_rate, beta_1, beta_2, epsilon) (utils/tf_optimizers.py:11)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, **kw):
        if amsgrad:
            raise NotImplementedError("amsgrad=True is outside the hot-path scope")
        self.learning_rate = float(kw.get("lr", learning_rate))
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)


class _WeightRef:
    def __init__(self, name, shape):
        self.name, self.shape = name, tuple(shape)

    def get_shape(self):
        return self.shape


class History:
    def __init__(self):
        self.history: Dict[str, List[float]] = {}
        self.epoch: List[int] = []


# Keras metric identifiers the reference passes to compile(metrics=...) (2DCNN/utils/tf_metrics.py; the shipped INI uses MeanSquaredError
# and monitors val_mean_squared_error, Train_Configs.ini:36,44).  Computed on the host from predict() for the VALIDATION logs of fit()
# (what the reference's callbacks monitor); per-batch training metrics would need every output read back and are not produced.
def _m_bin_acc(t, p):
    return float(((p > 0.5) == (t > 0.5)).mean())


def _m_cat_acc(t, p):
    return float((p.argmax(-1) == t.argmax(-1)).mean())


_METRICS = {
    "mean_squared_error": lambda t, p: float(((p - t) ** 2).mean()), "mean_absolute_error": lambda t, p: float(np.abs(p - t).mean()),
    "binary_accuracy": _m_bin_acc, "categorical_accuracy": _m_cat_acc,
    "accuracy": lambda t, p: _m_cat_acc(t, p) if p.shape[-1] > 1 else _m_bin_acc(t, p),    # Keras picks by the output shape
    "binary_crossentropy": lambda t, p: float(-(t * np.log(np.clip(p, 1e-7, 1 - 1e-7)) + (1 - t) * np.log(1 - np.clip(p, 1e-7, 1 - 1e-7))).mean()),
}
_METRIC_ALIASES = {"mse": "mean_squared_error", "mae": "mean_absolute_error", "acc": "accuracy", "meansquarederror": "mean_squared_error",
                   "meanabsoluteerror": "mean_absolute_error", "binaryaccuracy": "binary_accuracy", "categoricalaccuracy": "categorical_accuracy",
                   "binarycrossentropy": "binary_crossentropy"}


def _metric_name(obj):
    """canonical name of a metric given as a string or as an object with .name (tf.keras.metrics.MeanSquaredError(name=...)); None if
    it is not one of the host-side metrics above"""
    key = obj if isinstance(obj, str) else (getattr(obj, "name", None) or type(obj).__name__)
    key = str(key)
    low = key.lower().replace(" ", "")
    name = key if key in _METRICS else _METRIC_ALIASES.get(low, low if low in _METRICS else None)
    return name


def _loss_name(obj) -> str:
    if isinstance(obj, str):
        key = obj.lower().replace(" ", "")
    else:
        key = getattr(obj, "name", None) or type(obj).__name__
        key = key.lower()
    if key not in _LOSS_ALIASES:
        raise NotImplementedError(f"loss '{obj}' is outside the hot-path scope (supported: bce, cce, mse, mae)")
    return _LOSS_ALIASES[key]


class Model:
    def __init__(self, graph: Graph):
        self.graph = graph
        self.name = graph.name
        self.output_names = [n.attrs.get("output_name", n.name) for n in graph.outputs]   # (a head with a non-fused activation: graph.conv)
        self._weights: Dict[str, np.ndarray] = init_params(graph)
        self._engines: Dict[tuple, object] = {}
        self._primary = None
        self._losses: Optional[List[str]] = None
        self._loss_weights: Optional[List[float]] = None
        self.optimizer: Optional[Adam] = None
        self.stop_training = False
        self.world_size = 1
        self._dist = None
        self.exchange_bucket_bytes = 64 << 20   # gradient-exchange bucket (data parallel): overlap granularity vs launch count
        self._ds_targets = None

    # ---- introspection -----------------------------------------------------------------------------------
    @property
    def input_shape(self):
        H, W, C = self.graph.inputs[0].shape
        return (None, H, W, C) if self.graph.ndim == 2 else (None, W, C)

    @property
    def trainable_weights(self):
        return [_WeightRef(f"{l}/{w}", s) for (l, w, s, _, t) in self.graph.param_specs() if t]

    @property
    def non_trainable_weights(self):
        return [_WeightRef(f"{l}/{w}", s) for (l, w, s, _, t) in self.graph.param_specs() if not t]

    def count_params(self):
        return sum(self.graph.count_params())

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"')
        print_fn(f"{'Layer (type)':<40}{'Output Shape':<28}{'Param #':>12}")
        per_layer: Dict[str, int] = {}
        for (l, _, s, _, _) in self.graph.param_specs():
            per_layer[l] = per_layer.get(l, 0) + int(np.prod(s))
        for n in self.graph.nodes:
            shp = (None,) + (n.shape if self.graph.ndim == 2 else n.shape[1:])
            print_fn(f"{(n.name + ' (' + n.op + ')'):<40}{str(shp):<28}{per_layer.get(n.name, 0):>12}")
        tr, nt = self.graph.count_params()
        print_fn(f"Total params: {tr + nt}\nTrainable params: {tr}\nNon-trainable params: {nt}")

    # ---- weights -----------------------------------------------------------------------------------------
    def _sync_from_device(self):
        if self._primary is not None:
            self._weights = self._primary.get_weights()

    def get_weight_dict(self) -> Dict[str, np.ndarray]:
        self._sync_from_device()
        return {k: v.copy() for k, v in self._weights.items()}

    def set_weight_dict(self, params: Dict[str, np.ndarray]):
        for k, v in params.items():
            if k not in self._weights:
                raise KeyError(f"unknown weight {k}")
            if tuple(np.shape(v)) != self._weights[k].shape:
                raise ValueError(f"weight {k}: shape {np.shape(v)} != {self._weights[k].shape}")
            self._weights[k] = np.asarray(v, np.float32).copy()
        if self._primary is not None:
            self._primary.set_weights(self._weights)

    def get_weights(self):
        self._sync_from_device()
        return [self._weights[f"{l}/{w}"].copy() for (l, w, _, _, _) in self.graph.param_specs()]

    def set_weights(self, weights):
        specs = self.graph.param_specs()
        if len(weights) != len(specs):
            raise ValueError(f"You called `set_weights(weights)` with a weight list of length {len(weights)}, but the model was expecting {len(specs)} weights.")
        self.set_weight_dict({f"{l}/{w}": a for (l, w, _, _, _), a in zip(specs, weights)})

    def save_weights(self, path):
        self._sync_from_device()
        np.savez(path if str(path).endswith(".npz") else str(path) + ".npz", **{k.replace("/", "::"): v for k, v in self._weights.items()})

    def load_weights(self, path):
        p = str(path)
        if not os.path.exists(p) and os.path.exists(p + ".npz"):
            p = p + ".npz"
        if p.endswith((".h5", ".hdf5", ".keras")):
            # Keras weight files (Train.py:363,375): needs h5py, absent from the build image — see b2seg/keras_io.py
            from .keras_io import read_keras_weights, select_for_model
            weights, _extra = select_for_model(read_keras_weights(p, self.graph.param_specs()), self.graph.param_specs())
            self.set_weight_dict(weights)
            return
        with np.load(p) as z:
            self.set_weight_dict({k.replace("::", "/"): z[k] for k in z.files})

    # ---- compile -----------------------------------------------------------------------------------------
    def compile(self, loss=None, optimizer=None, metrics=None, loss_weights=None, ds_targets=None, **kw):
        """ds_targets (extension; None = Keras behaviour): 'UNet' or 'UNetPP' — fit / train_on_batch / evaluate then take the mask
        alone and the deep-supervision targets of the other outputs are derived from it ON THE DEVICE (the reference's
        prepareTrainDict, helper_functions.py:359-380 / notebook cell 31, run per batch on the host): an output of the mask's
        shape gets the mask, a smaller one its window max (2D) / window mean (1D)."""
        if ds_targets not in (None, "UNet", "UNetPP"):
            raise ValueError("ds_targets must be None, 'UNet' or 'UNetPP'")
        self._ds_targets = ds_targets
        names = self.output_names
        if isinstance(loss, dict):
            self._losses = [_loss_name(loss[n]) for n in names]
        elif isinstance(loss, (list, tuple)):
            self._losses = [_loss_name(l) for l in loss]
        else:
            self._losses = [_loss_name(loss)] * len(names)
        if loss_weights is None:
            self._loss_weights = [1.0] * len(names)
        elif isinstance(loss_weights, dict):
            self._loss_weights = [float(loss_weights.get(n, 1.0)) for n in names]
        else:
            self._loss_weights = [float(w) for w in loss_weights]
        if optimizer is None or (isinstance(optimizer, str) and optimizer.lower() == "adam"):
            optimizer = Adam()
        if not isinstance(optimizer, Adam):
            if hasattr(optimizer, "learning_rate") and type(optimizer).__name__ == "Adam":
                optimizer = Adam(float(optimizer.learning_rate), getattr(optimizer, "beta_1", 0.9), getattr(optimizer, "beta_2", 0.999),
                                 getattr(optimizer, "epsilon", 1e-7))
            else:
                raise NotImplementedError("only the Adam optimizer is inside the hot-path scope (utils/tf_optimizers.py:11)")
        self.optimizer = optimizer
        self.metrics = metrics or []
        self._engines = {k: e for k, e in self._engines.items() if not k[1]}

    # ---- engines -----------------------------------------------------------------------------------------
    def _engine(self, batch: int, training: bool):
        from .engine import Engine
        key = (batch, training)
        if key not in self._engines:
            adam = None
            if training:
                if self._losses is None:
                    raise RuntimeError("You must compile your model before training/testing. Use `model.compile(optimizer, loss)`.")
                o = self.optimizer
                adam = dict(lr=o.learning_rate, beta1=o.beta_1, beta2=o.beta_2, eps=o.epsilon)
            bucket = 0
            if training:
                # data parallel (a process group exists when the engine is built): one Adam op per exchange bucket
                import torch.distributed as dist
                if self.world_size > 1 or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
                    bucket = self.exchange_bucket_bytes
            eng = Engine(self.graph, batch, training=training, losses=self._losses, loss_weights=self._loss_weights, adam=adam,
                         share_params_from=self._primary, adam_bucket_bytes=bucket)
            if self._primary is None:
                self._primary = eng
                eng.set_weights(self._weights)
            if bucket:
                import torch.distributed as dist
                if dist.get_backend(getattr(self, "_pg", None)) == "nccl":
                    eng.reserve_sms_for_exchange(eng.planner.exchange_schedule(self.exchange_bucket_bytes), group=getattr(self, "_pg", None))
            self._engines[key] = eng
        return self._engines[key]

    def _to_nhwc(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        if self.graph.ndim == 1:
            if a.ndim != 3:
                raise ValueError(f"expected input of rank 3 (N, length, channels), got shape {a.shape}")
            return a[:, None, :, :]
        if a.ndim != 4:
            raise ValueError(f"expected input of rank 4 (N, H, W, channels), got shape {a.shape}")
        return a

    def _targets(self, y, host=False) -> List[np.ndarray]:
        """target arrays in output order; with compile(ds_targets=...) and a bare mask: [mask, None, ...] (None = derived from the
        mask on the device by Engine.derive_targets), or all of them derived on the host when `host` is set"""
        if self._ds_targets is not None and not isinstance(y, (dict, list, tuple)) and len(self.output_names) > 1:
            mask = self._to_nhwc(y)
            if not host:
                return [mask] + [None] * (len(self.output_names) - 1)
            from .helpers import derive_targets_host
            shapes = [(mask.shape[0],) + tuple(n.shape) for n in self.graph.outputs]
            return derive_targets_host(mask, shapes, self.graph.ndim)
        if isinstance(y, dict):
            ys = [y[n] for n in self.output_names]
        elif isinstance(y, (list, tuple)):
            ys = list(y)
        else:
            ys = [y]
        if len(ys) != len(self.output_names):
            raise ValueError(f"expected {len(self.output_names)} target arrays, got {len(ys)}")
        return [self._to_nhwc(t) for t in ys]

    # ---- data parallel (one process per GPU; torch.distributed/NCCL only moves the flat gradient arena) -----
    def distribute(self, process_group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dist = dist
        self._pg = process_group
        self.world_size = dist.get_world_size(process_group)
        return self

    def broadcast_weights(self, src=0):
        eng = self._primary
        if eng is None:
            raise RuntimeError("build an engine first (train_on_batch / predict)")
        for t in (eng.w, eng.moving, eng.m, eng.v):
            self._dist.broadcast(t, src, group=self._pg)
        eng.wb.copy_(eng.w.to(eng.wb.dtype))

    # ---- train / predict ---------------------------------------------------------------------------------
    def train_on_batch(self, x, y, return_loss=True):
        import torch
        xs = self._to_nhwc(x)
        ys = self._targets(y)
        eng = self._engine(xs.shape[0], True)
        eng.x_dev.copy_(torch.from_numpy(xs), non_blocking=True)
        for o, t in zip(eng.outputs, ys):
            if t is None:
                continue
            if tuple(t.shape) != tuple(o["shape"]):
                raise ValueError(f"target for '{o['name']}' has shape {t.shape}, expected {o['shape']}")
            o["target"].copy_(torch.from_numpy(t), non_blocking=True)
        if any(t is None for t in ys):
            eng.derive_targets()
        return self._step(eng, return_loss)

    def _exchange_schedule(self, eng):
        sched = getattr(eng, "_exchange", None)
        if sched is None:
            sched = eng._exchange = eng.planner.exchange_schedule(self.exchange_bucket_bytes)
        return sched

    def _step(self, eng, return_loss=True):
        eng.forward()
        scale = 1.0
        if self.world_size > 1:
            # backward with the gradient exchange overlapped: as soon as the ops that finish a slice of the gradient
            # arena are enqueued, its all-reduce starts on the collective's stream while the remaining backward ops run
            import torch.distributed as dist
            from .dist import wait_all
            works, done = [], 0
            for (n_ops, lo, hi) in self._exchange_schedule(eng):
                if n_ops > done:
                    eng.run_range(1, done, n_ops - done)
                    done = n_ops
                works.append(dist.all_reduce(eng.g[lo:hi], group=self._pg, async_op=True))
            n_total = eng.planner.num_launch_ops(1)
            if n_total > done:
                eng.run_range(1, done, n_total - done)
            scale = 1.0 / self.world_size
            if eng.adam_bucket_bytes == self.exchange_bucket_bytes and len(eng.planner.ops[2]) == len(works):
                # bucket i's Adam runs as soon as ITS all-reduce has landed; the later buckets are still on the wire
                eng.optimizer_begin(self.optimizer.learning_rate, scale)
                for i, w in enumerate(works):
                    w.wait()
                    eng.run_range(2, i, 1)
                if return_loss:
                    return float(eng.loss_buf.item())
                return None
            wait_all(works)
        else:
            eng.backward()
        eng.optimizer_step(self.optimizer.learning_rate, scale)
        if return_loss:
            return float(eng.loss_buf.item())
        return None

    def predict(self, x, batch_size=None, verbose=0, **kw):
        import torch
        xs = self._to_nhwc(x)
        n = xs.shape[0]
        bs = int(batch_size) if batch_size else min(n, 32)
        bs = max(1, min(bs, n))
        eng = self._engine(bs, False)
        outs = [np.empty((n,) + tuple(o["shape"][1:]), np.float32) for o in eng.outputs]
        for s in range(0, n, bs):
            chunk = xs[s:s + bs]
            m = chunk.shape[0]
            if m < bs:
                chunk = np.concatenate([chunk, np.zeros((bs - m,) + chunk.shape[1:], np.float32)], 0)
            eng.x_dev.copy_(torch.from_numpy(chunk), non_blocking=True)
            eng.forward()
            for i, o in enumerate(eng.outputs):
                outs[i][s:s + m] = o["y"][:m].cpu().numpy()
        if self.graph.ndim == 1:
            outs = [o[:, 0] for o in outs]
        return outs if len(outs) > 1 else outs[0]

    def evaluate(self, x, y, batch_size=32, verbose=0, **kw):
        """mean of the compiled (weighted) loss over batches, computed on the host from predict()"""
        ys = self._targets(y, host=True)
        pred = self.predict(x, batch_size=batch_size)
        pred = pred if isinstance(pred, list) else [pred]
        total = 0.0
        for p, t, kind, w in zip(pred, ys, self._losses, self._loss_weights):
            t = t[:, 0] if self.graph.ndim == 1 else t
            p = p.astype(np.float64)
            if kind == "bce":
                pc = np.clip(p, 1e-7, 1 - 1e-7)
                l = -(t * np.log(pc) + (1 - t) * np.log(1 - pc)).mean()
            elif kind == "cce":
                l = -(t * np.log(np.clip(p, 1e-7, 1))).sum(-1).mean()
            elif kind == "mse":
                l = ((p - t) ** 2).mean()
            else:
                l = np.abs(p - t).mean()
            total += w * float(l)
        if kw.get("return_dict"):
            logs = {"loss": total}
            names = [n for n in (_metric_name(mt) for mt in (self.metrics or [])) if n is not None]
            for out_name, p, t in zip(self.output_names, pred, ys):
                t = t[:, 0] if self.graph.ndim == 1 else t
                for n in names:        # Keras prefixes the output's name when the model has several outputs
                    logs[n if len(pred) == 1 else f"{out_name}_{n}"] = _METRICS[n](np.asarray(t, np.float64), p.astype(np.float64))
            return logs
        return total

    def _train_batches_pipelined(self, batches):
        """Train on a list of equally sized (x, [targets]) host batches with the input pipeline overlapped: batch i+1 is copied
        host -> device on a copy stream (into one of two staging slots) while step i computes, and every step's loss is read back
        asynchronously into pinned memory and collected at the end — what Keras' fit() does with its prefetching data adapter
        (2DCNN/Train.py:394-415).  Pinned source arrays make the copies truly asynchronous; pageable ones still work."""
        import torch
        B = batches[0][0].shape[0]
        eng = self._engine(B, True)
        cur = torch.cuda.current_stream(eng.dev)
        if not hasattr(eng, "_stage"):
            eng._copy_stream = torch.cuda.Stream(device=eng.dev)
            eng._stage = [dict(x=torch.empty_like(eng.x_dev), t=[torch.empty_like(o["target"]) for o in eng.outputs],
                               ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
        derive = any(t is None for t in batches[0][1])    # compile(ds_targets=...): only the mask is staged
        loss_host = torch.empty(len(batches), dtype=torch.float32).pin_memory()
        for i, (bx, bys) in enumerate(batches):
            st = eng._stage[i % 2]
            with torch.cuda.stream(eng._copy_stream):
                eng._copy_stream.wait_event(st["free"])          # the step that used this slot has copied it out
                st["x"].copy_(torch.from_numpy(bx), non_blocking=True)
                for dst, t in zip(st["t"], bys):
                    if t is None:
                        continue
                    if tuple(t.shape) != tuple(dst.shape):
                        raise ValueError(f"target shape {t.shape} != {tuple(dst.shape)}")
                    dst.copy_(torch.from_numpy(np.ascontiguousarray(t, np.float32)), non_blocking=True)
                st["ready"].record(eng._copy_stream)
            cur.wait_event(st["ready"])
            eng.x_dev.copy_(st["x"], non_blocking=True)
            for o, t, src in zip(eng.outputs, st["t"], bys):
                if src is not None:
                    o["target"].copy_(t, non_blocking=True)
            if derive:
                eng.derive_targets()
            st["free"].record(cur)
            self._step(eng, return_loss=False)
            loss_host[i].copy_(eng.loss_buf[0], non_blocking=True)
        cur.synchronize()
        return [float(v) for v in loss_host.tolist()]

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None, shuffle=True,
            initial_epoch=0, steps_per_epoch=None, **kw):
        hist = History()
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin({})
        sequence = x if (y is None and hasattr(x, "__getitem__") and hasattr(x, "__len__") and not isinstance(x, np.ndarray)) else None
        if sequence is None:
            xs = np.asarray(x, np.float32)
            ys = self._targets(y)
            bs = int(batch_size or 32)
            split = float(kw.get("validation_split") or 0.0)
            if split and validation_data is None:
                # Keras holds out the LAST fraction of the samples, before any shuffling (Train.py:411-415)
                if not 0.0 < split < 1.0:
                    raise ValueError(f"`validation_split` must be between 0 and 1, received: {split}")
                cut = int(xs.shape[0] * (1.0 - split))
                held = [t for t in ys if t is not None]          # (ds_targets: the mask alone; evaluate() derives the rest)
                validation_data = (xs[cut:], [t[cut:][:, 0] if self.graph.ndim == 1 else t[cut:] for t in held])
                if len(held) == 1:
                    validation_data = (validation_data[0], validation_data[1][0])
                xs, ys = xs[:cut], [None if t is None else t[:cut] for t in ys]
            n = xs.shape[0]
        rng = np.random.default_rng(0)
        self.stop_training = False
        for ep in range(initial_epoch, epochs):
            t0 = time.time()
            losses = []
            if sequence is not None:
                for bi in range(len(sequence)):
                    bx, by = sequence[bi][:2]
                    losses.append(self.train_on_batch(bx, by))
                if hasattr(sequence, "on_epoch_end"):
                    sequence.on_epoch_end()
            else:
                order = rng.permutation(n) if shuffle else None
                starts = list(range(0, n, bs))
                if steps_per_epoch:
                    starts = starts[:int(steps_per_epoch)]
                full = [s for s in starts if s + bs <= n]
                xn = self._to_nhwc(xs)

                def take(a, s):   # contiguous slice (keeps pinned memory pinned) unless shuffled
                    if a is None:
                        return None
                    return a[s:s + bs] if order is None else a[order[s:s + bs]]
                if full:
                    losses += self._train_batches_pipelined([(take(xn, s), [take(t, s) for t in ys]) for s in full])
                for s in starts[len(full):]:    # ragged last batch: its own engine
                    by = [take(t, s)[:, 0] if self.graph.ndim == 1 else take(t, s) for t in ys if t is not None]
                    losses.append(self.train_on_batch(take(xs, s), by if len(by) > 1 else by[0]))
            logs = {"loss": float(np.mean(losses))}
            if validation_data is not None:
                vx, vy = validation_data[:2]
                for k_, v_ in self.evaluate(vx, vy, batch_size=batch_size or 32, return_dict=True).items():
                    logs[f"val_{k_}"] = v_
            if verbose:
                print(f"Epoch {ep + 1}/{epochs} - {time.time() - t0:.1f}s - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(ep, logs)
            # Keras' History callback runs after the user's callbacks: what they add to `logs` (ReduceLROnPlateau's `lr`) is recorded
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            hist.epoch.append(ep)
            if self.stop_training:
                break
        for cb in callbacks:
            if hasattr(cb, "on_train_end"):
                cb.on_train_end({})
        self.history = hist
        return hist

    # ---- parity taps --------------------------------------------------------------------------------------
    def layer_output(self, name, batch, training=True, grad=False):
        return self._engine(batch, training).tap(name, grad=grad)
