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

# Keras identifiers (class names of 2DCNN/utils/tf_losses.py:8-46, their `name=` strings and the usual short forms) -> canonical
# loss names of the planner (b2seg.planner.LOSS_KINDS)
_LOSS_ALIASES = {}
for _canon, _names in {
        "bce": ("binary_crossentropy", "bce", "binarycrossentropy"),
        "cce": ("categorical_crossentropy", "cce", "categoricalcrossentropy"),
        "mse": ("mean_squared_error", "mse", "meansquarederror"),
        "mae": ("mean_absolute_error", "mae", "meanabsoluteerror"),
        "msle": ("mean_squared_logarithmic_error", "msle", "meansquaredlogarithmicerror"),
        "huber": ("huber_loss", "huber"),
        "logcosh": ("log_cosh", "logcosh"),
        "focal": ("binary_focal_crossentropy", "binaryfocalcrossentropy"),
        "poisson": ("poisson",),
        "kld": ("kl_divergence", "kld", "kldivergence", "kullback_leibler_divergence"),
        "hinge": ("hinge",),
        "squared_hinge": ("squared_hinge", "squaredhinge"),
        "mape": ("mean_absolute_percentage_error", "mape", "meanabsolutepercentageerror"),
        "categorical_hinge": ("categorical_hinge", "categoricalhinge"),
        "cosine": ("cosine_similarity", "cosinesimilarity")}.items():
    for _n in _names:
        _LOSS_ALIASES[_n] = _canon


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
    "binary_crossentropy": lambda t, p: float(-(t * np.log(np.clip(p, 1e-7, 1 - 1e-7)) + (1 - t) * np.log(1 - np.clip(p, 1e-7, 1 - 1e-7))).mean()),
}
_METRIC_ALIASES = {"mse": "mean_squared_error", "mae": "mean_absolute_error", "acc": "accuracy", "meansquarederror": "mean_squared_error",
                   "meanabsoluteerror": "mean_absolute_error", "binaryaccuracy": "binary_accuracy", "categoricalaccuracy": "categorical_accuracy",
                   "binarycrossentropy": "binary_crossentropy"}


def _accuracy_kind(loss_kind, cout):
    """what the string 'accuracy' means for one output (keras/engine/compile_utils.py get_metric_function): binary accuracy for a
    one-channel output or a binary cross-entropy loss, else categorical accuracy"""
    return "binary_accuracy" if (cout == 1 or loss_kind in ("bce", "focal")) else "categorical_accuracy"


def _metric_value(name, loss_kind, t, p):
    if name == "accuracy":
        name = _accuracy_kind(loss_kind, p.shape[-1])
    return _METRICS[name](t, p)


def _host_loss(kind, t, p):
    """float64 NumPy value of a compiled loss on activated outputs (evaluate / validation logs; the training path uses b2seg_loss)"""
    eps = 1e-7
    if kind in ("bce", "focal"):
        pc = np.clip(p, eps, 1 - eps)
        bce = -(t * np.log(pc) + (1 - t) * np.log(1 - pc))
        if kind == "focal":
            bce = (1 - (t * p + (1 - t) * (1 - p))) ** 2 * bce
        return float(bce.mean())
    if kind == "cce":
        return float(-(t * np.log(np.clip(p / p.sum(-1, keepdims=True), eps, 1))).sum(-1).mean())
    if kind == "mse":
        return float(((p - t) ** 2).mean())
    if kind == "mae":
        return float(np.abs(p - t).mean())
    if kind == "msle":
        return float(((np.log(np.maximum(p, eps) + 1) - np.log(np.maximum(t, eps) + 1)) ** 2).mean())
    if kind == "huber":
        d = np.abs(p - t)
        return float(np.where(d <= 1, 0.5 * d * d, d - 0.5).mean())
    if kind == "logcosh":
        d = p - t
        return float((d + np.logaddexp(0.0, -2 * d) - np.log(2.0)).mean())
    if kind == "poisson":
        return float((p - t * np.log(p + eps)).mean())
    if kind == "kld":
        tc, pk = np.clip(t, eps, 1), np.clip(p, eps, 1)
        return float((tc * np.log(tc / pk)).sum(-1).mean())
    if kind in ("hinge", "squared_hinge"):
        y = 2 * t - 1 if np.isin(t, (0.0, 1.0)).all() else t
        m = np.maximum(1 - y * p, 0)
        return float((m if kind == "hinge" else m * m).mean())
    if kind == "mape":
        return float((100 * np.abs((t - p) / np.maximum(np.abs(t), eps))).mean())
    if kind == "categorical_hinge":
        return float(np.maximum(((1 - t) * p).max(-1) - (t * p).sum(-1) + 1, 0).mean())
    if kind == "cosine":
        def l2n(v):
            return v / np.sqrt(np.maximum((v * v).sum(-1, keepdims=True), 1e-12))
        return float(-(l2n(t) * l2n(p)).sum(-1).mean())
    raise ValueError(kind)


def _metric_name(obj):
    """canonical name of a metric given as a string or as an object with .name (tf.keras.metrics.MeanSquaredError(name=...)); None if
    it is not one of the host-side metrics above"""
    key = obj if isinstance(obj, str) else (getattr(obj, "name", None) or type(obj).__name__)
    key = str(key)
    low = key.lower().replace(" ", "")
    if low == "accuracy":
        return "accuracy"
    name = key if key in _METRICS else _METRIC_ALIASES.get(low, low if low in _METRICS else None)
    return name


def _loss_name(obj) -> str:
    if isinstance(obj, str):
        key = obj.lower().replace(" ", "")
    else:
        key = getattr(obj, "name", None) or type(obj).__name__
        key = key.lower()
    if key not in _LOSS_ALIASES:
        raise NotImplementedError(f"loss '{obj}' is not lowered (supported: every loss of utils/tf_losses.py except SparseCategoricalCrossentropy: "
                                  f"{sorted(set(_LOSS_ALIASES.values()))})")
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
        self.exchange_bucket_bytes = int(os.environ.get("B2SEG_BUCKET_MB", "64")) << 20   # gradient-exchange bucket (data parallel): overlap granularity vs launch count
        # one process, one GPU: Adam is cut into buckets of this size and each bucket runs on a side stream as soon as backward has
        # finished its gradients, beside the remaining (tensor-core bound) backward kernels instead of after them.  0 = one Adam
        # launch after backward.
        self.adam_overlap_bytes = int(os.environ.get("B2SEG_ADAM_OVERLAP_MB", "8")) << 20
        self._ds_targets = None
        # False (default): activation / gradient buffers share one arena by liveness (Planner._assign_memory).  True: every layer's
        # tensors stay readable after a step (layer_output / the per-layer parity tests); B2SEG_KEEP_ACTIVATIONS=1 forces it.
        self.keep_activations = bool(os.environ.get("B2SEG_KEEP_ACTIVATIONS"))
        # data parallel: "sharded" = per bucket reduce-scatter of the fp32 gradients -> Adam on this rank's 1/world slice -> all-gather
        # of the bf16 weights the kernels read (25 % fewer bytes on the wire than an all-reduce, Adam 1/world of the work);
        # "allreduce" = all-reduce + replicated Adam (round 1)
        self.dp_mode = os.environ.get("B2SEG_DP_MODE", "sharded")
        self._wversion = 0              # bumped whenever the weights change: inference engines refold BatchNorm into their kernels lazily
        self._master_version = 0        # _wversion at which the fp32 masters were last complete on this rank (sharded data parallel)
        self._adam_step = 0             # Adam's t: one counter per model (the moments are shared by the engines of every batch size)

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
            self.gather_master_weights()
            self._weights = self._primary.get_weights()

    def gather_master_weights(self):
        """Sharded data-parallel training keeps each fp32 master weight up to date only on the rank that owns its slice (the kernels
        read the all-gathered bf16 copy).  This all-gathers the fp32 slices so that every rank holds the full-precision weights —
        a COLLECTIVE: get_weights / save_weights / a checkpoint callback must then run on every rank, as under MirroredStrategy."""
        eng = next((e for (b, tr), e in self._engines.items() if tr and e.planner.shard[1] > 1), None)
        if eng is None or self._master_version == self._wversion:
            return
        from .dist import all_gather_bucket_
        rank, world = eng.planner.shard
        works = [all_gather_bucket_(eng.w, lo, hi, rank, world, self._pg) for (_n, lo, hi) in self._exchange_schedule(eng)]
        for w_ in works:
            w_.wait()
        self._master_version = self._wversion

    def get_weight_dict(self) -> Dict[str, np.ndarray]:
        self._sync_from_device()
        return {k: v.copy() for k, v in self._weights.items()}

    def set_weight_dict(self, params: Dict[str, np.ndarray]):
        self._sync_from_device()       # a partial update must not push stale host copies of the OTHER weights back to the device
        for k, v in params.items():
            if k not in self._weights:
                raise KeyError(f"unknown weight {k}")
            if tuple(np.shape(v)) != self._weights[k].shape:
                raise ValueError(f"weight {k}: shape {np.shape(v)} != {self._weights[k].shape}")
            self._weights[k] = np.asarray(v, np.float32).copy()
        self._wversion += 1
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
        # Keras builds a fresh optimizer on every compile (the reference recompiles the same model for every fold, Train.py:320-325):
        # zero moments, t = 0.  The weights stay.
        self._adam_step = 0
        if self._primary is not None:
            self._primary.reset_optimizer()

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
            shard = (0, 1)
            if bucket and self.dp_mode == "sharded":
                import torch.distributed as dist
                shard = (dist.get_rank(getattr(self, "_pg", None)), dist.get_world_size(getattr(self, "_pg", None)))
            exchange = bucket > 0
            if training and not bucket:
                bucket = self.adam_overlap_bytes
            eng = Engine(self.graph, batch, training=training, losses=self._losses, loss_weights=self._loss_weights, adam=adam,
                         share_params_from=self._primary, adam_bucket_bytes=bucket, reuse=not self.keep_activations, shard=shard)
            if self._primary is None:
                self._primary = eng
                eng.set_weights(self._weights)
            if exchange:
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
        self._wversion += 1

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
            sched = eng._exchange = eng.planner.exchange_schedule(eng.adam_bucket_bytes or self.exchange_bucket_bytes)
        return sched

    def _step(self, eng, return_loss=True):
        self._wversion += 1
        eng.forward()
        scale = 1.0
        if self.world_size > 1 and eng.planner.shard[1] > 1:
            return self._step_sharded(eng, return_loss)
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
                self._adam_step += 1
                eng.optimizer_begin(self.optimizer.learning_rate, scale, step=self._adam_step)
                for i, w in enumerate(works):
                    w.wait()
                    eng.run_range(2, i, 1)
                if return_loss:
                    return float(eng.loss_buf.item())
                return None
            wait_all(works)
        elif eng.adam_bucket_bytes and len(eng.planner.ops[2]) > 1:
            return self._step_overlapped_adam(eng, return_loss)
        else:
            eng.backward()
        self._adam_step += 1
        eng.optimizer_step(self.optimizer.learning_rate, scale, step=self._adam_step)
        if return_loss:
            return float(eng.loss_buf.item())
        return None

    def _step_overlapped_adam(self, eng, return_loss):
        """One process: backward is replayed in the ranges of the bucket schedule (Planner.exchange_schedule — bucket i of the flat
        gradient arena is final, and its weights are no longer read, once the first n_ops backward ops have run); bucket i's Adam is
        enqueued on a side stream behind exactly those ops and runs beside the rest of backward (HBM-bound Adam under the
        tensor-core-bound convolutions).  The step's stream rejoins the side stream before anything else runs."""
        import torch
        sched = self._exchange_schedule(eng)
        cuda = eng.dev.type == "cuda"
        self._adam_step += 1
        eng.optimizer_begin(self.optimizer.learning_rate, 1.0, step=self._adam_step)
        if cuda:
            if not hasattr(eng, "_side_stream"):
                eng._side_stream = torch.cuda.Stream(device=eng.dev)
            main, side = torch.cuda.current_stream(eng.dev), eng._side_stream
        done = 0
        for i, (n_ops, _lo, _hi) in enumerate(sched):
            if n_ops > done:
                eng.run_range(1, done, n_ops - done)
                done = n_ops
            if cuda:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    eng.run_range(2, i, 1)
            else:
                eng.run_range(2, i, 1)
        n_total = eng.planner.num_launch_ops(1)
        if n_total > done:
            eng.run_range(1, done, n_total - done)
        if cuda:
            main.wait_stream(side)
        if return_loss:
            return float(eng.loss_buf.item())
        return None

    def _step_sharded(self, eng, return_loss):
        """Data-parallel step with a sharded optimizer.  Backward is replayed in the exchange schedule's ranges; as soon as the ops
        that finish a bucket of the gradient arena are enqueued its reduce-scatter starts on the collective's stream.  On a side
        stream, bucket by bucket: wait for the reduce-scatter, Adam on this rank's slice (fp32 master, moments, bf16 copy), all-gather
        of the bucket's bf16 weights — all of it beside the remaining backward kernels.  The step ends when the last all-gather has
        landed; only the last bucket's reduce-scatter -> Adam -> all-gather chain is exposed."""
        import torch
        from .dist import all_gather_bucket_, reduce_scatter_bucket_
        rank, world = eng.planner.shard
        sched = self._exchange_schedule(eng)
        cuda = eng.dev.type == "cuda"
        if cuda and not hasattr(eng, "_side_stream"):
            eng._side_stream = torch.cuda.Stream(device=eng.dev)
        self._adam_step += 1
        eng.optimizer_begin(self.optimizer.learning_rate, 1.0 / world, step=self._adam_step)
        done, gathers = 0, []
        main = torch.cuda.current_stream(eng.dev) if cuda else None
        for i, (n_ops, lo, hi) in enumerate(sched):
            if n_ops > done:
                eng.run_range(1, done, n_ops - done)
                done = n_ops
            rs = reduce_scatter_bucket_(eng.g, lo, hi, rank, world, self._pg)
            if cuda:
                with torch.cuda.stream(eng._side_stream):
                    rs.wait()
                    eng.run_range(2, i, 1)
                    gathers.append(all_gather_bucket_(eng.wb, lo, hi, rank, world, self._pg))
            else:
                rs.wait()
                eng.run_range(2, i, 1)
                gathers.append(all_gather_bucket_(eng.wb, lo, hi, rank, world, self._pg))
        n_total = eng.planner.num_launch_ops(1)
        if n_total > done:
            eng.run_range(1, done, n_total - done)
        for ag in gathers:
            ag.wait()
        if return_loss:
            return float(eng.loss_buf.item())
        return None

    def predict(self, x, batch_size=None, verbose=0, **kw):
        """tf.keras.Model.predict (2DCNN/Test.py:149-164): moving-statistics BatchNorm folded into the convolution kernels (refolded
        lazily when the weights have changed), batches pipelined: while batch i computes, batch i+1 is staged host -> device on one
        copy stream and the outputs of batch i-1 travel device -> host on another (pinned staging on both sides)."""
        import torch
        xs = self._to_nhwc(x)
        n = xs.shape[0]
        bs = int(batch_size) if batch_size else min(n, 32)
        bs = max(1, min(bs, n))
        eng = self._engine(bs, False)
        if getattr(eng, "_fold_version", None) != self._wversion:
            eng.run(2)                                   # phase 2 of an inference plan: b2seg_fold_bn of every Conv + BatchNorm pair
            eng._fold_version = self._wversion
        outs = [np.empty((n,) + tuple(o["shape"][1:]), np.float32) for o in eng.outputs]
        starts = list(range(0, n, bs))
        if eng.dev.type != "cuda":                       # (emulator engine of the CPU test-suite)
            for s in starts:
                chunk = xs[s:s + bs]
                m = chunk.shape[0]
                if m < bs:
                    chunk = np.concatenate([chunk, np.zeros((bs - m,) + chunk.shape[1:], np.float32)], 0)
                eng.x_dev.copy_(torch.from_numpy(chunk))
                eng.forward()
                for i, o in enumerate(eng.outputs):
                    outs[i][s:s + m] = o["y"][:m].cpu().numpy()
        else:
            st = getattr(eng, "_predict_stage", None)
            if st is None:
                st = eng._predict_stage = dict(
                    h2d=torch.cuda.Stream(device=eng.dev), d2h=torch.cuda.Stream(device=eng.dev),
                    slots=[dict(hin=torch.empty(tuple(eng.x_dev.shape), dtype=torch.float32).pin_memory(), xin=torch.empty_like(eng.x_dev),
                                y=[torch.empty_like(o["y"]) for o in eng.outputs],
                                hout=[torch.empty(tuple(o["y"].shape), dtype=torch.float32).pin_memory() for o in eng.outputs],
                                ready=torch.cuda.Event(), done=torch.cuda.Event(), landed=torch.cuda.Event(), consumed=torch.cuda.Event())
                           for _ in range(2)])
            cur = torch.cuda.current_stream(eng.dev)

            def collect(j):
                sl = st["slots"][j % 2]
                sl["landed"].synchronize()
                s_ = starts[j]
                m_ = min(bs, n - s_)
                for i in range(len(outs)):
                    outs[i][s_:s_ + m_] = sl["hout"][i][:m_].numpy()
            for j, s in enumerate(starts):
                sl = st["slots"][j % 2]
                m = min(bs, n - s)
                if j >= 2:
                    collect(j - 2)                        # frees this slot's host buffers (and overlaps with batch j-1 on the device)
                sl["hin"][:m].copy_(torch.from_numpy(xs[s:s + m]))
                if m < bs:
                    sl["hin"][m:].zero_()
                with torch.cuda.stream(st["h2d"]):
                    st["h2d"].wait_event(sl["consumed"])  # the forward that read this slot's device buffer has copied it out
                    sl["xin"].copy_(sl["hin"], non_blocking=True)
                    sl["ready"].record(st["h2d"])
                cur.wait_event(sl["ready"])
                eng.x_dev.copy_(sl["xin"], non_blocking=True)
                sl["consumed"].record(cur)
                eng.forward()
                for i, o in enumerate(eng.outputs):
                    sl["y"][i].copy_(o["y"], non_blocking=True)
                sl["done"].record(cur)
                with torch.cuda.stream(st["d2h"]):
                    st["d2h"].wait_event(sl["done"])
                    for i in range(len(outs)):
                        sl["hout"][i].copy_(sl["y"][i], non_blocking=True)
                    sl["landed"].record(st["d2h"])
            for j in range(max(0, len(starts) - 2), len(starts)):
                collect(j)
        if self.graph.ndim == 1:
            outs = [o[:, 0] for o in outs]
        return outs if len(outs) > 1 else outs[0]

    def evaluate(self, x, y=None, batch_size=32, verbose=0, steps=None, **kw):
        """the compiled (weighted) loss and metrics over x, y — arrays, a Sequence or an iterator of (x, y) batches (2DCNN/Train.py:216,
        281-300 passes a CustomDataGenerator or zip(generators) as validation_data).  Every batch runs predict() on the device; the
        loss / metric values are formed on the host and averaged with the batch sizes as weights, like Keras."""
        tot_w, acc = 0, {}
        for bx, by in self._host_batches(x, y, batch_size or 32, steps):
            ys = self._targets(by, host=True)
            pred = self.predict(bx, batch_size=batch_size or 32)
            pred = pred if isinstance(pred, list) else [pred]
            n = pred[0].shape[0]
            logs = {"loss": 0.0}
            names = [nm for nm in (_metric_name(mt) for mt in (self.metrics or [])) if nm is not None]
            for out_name, p, t, kind, w in zip(self.output_names, pred, ys, self._losses, self._loss_weights):
                t = np.asarray(t[:, 0] if self.graph.ndim == 1 else t, np.float64)
                p = p.astype(np.float64)
                l = _host_loss(kind, t, p)
                logs["loss"] += w * l
                if len(pred) > 1:
                    logs[f"{out_name}_loss"] = l
                for nm in names:        # Keras prefixes the output's name when the model has several outputs
                    logs[nm if len(pred) == 1 else f"{out_name}_{nm}"] = _metric_value(nm, kind, t, p)
            for k_, v_ in logs.items():
                acc[k_] = acc.get(k_, 0.0) + n * v_
            tot_w += n
        if not tot_w:
            raise ValueError("evaluate: no data")
        logs = {k_: v_ / tot_w for k_, v_ in acc.items()}
        return logs if kw.get("return_dict") else logs["loss"]

    # ---- input handling: arrays, keras.utils.Sequence-like objects, iterators / generators / zip(generators) ------------------
    @staticmethod
    def _is_sequence(x):
        return hasattr(x, "__getitem__") and hasattr(x, "__len__") and not isinstance(x, (np.ndarray, list, tuple, dict))

    @staticmethod
    def _is_iterator(x):
        return hasattr(x, "__next__") or (hasattr(x, "__iter__") and not isinstance(x, (np.ndarray, list, tuple, dict)) and not hasattr(x, "__getitem__"))

    @staticmethod
    def _split_item(item):
        """(x, y) of one generator / Sequence item; sample weights (a third element) are not supported"""
        if not isinstance(item, (tuple, list)) or len(item) < 2:
            raise ValueError("a data generator must yield (inputs, targets) tuples")
        if len(item) > 2 and item[2] is not None:
            raise NotImplementedError("per-sample weights from a generator are outside the hot-path scope")
        return item[0], item[1]

    def _host_batches(self, x, y, batch_size, steps=None, order=None):
        """(x batch, y batch) pairs of one pass over the data, whatever its form"""
        if isinstance(x, (tuple, list)) and y is None and len(x) in (2, 3) and not self._is_sequence(x):
            x, y = self._split_item(tuple(x))          # validation_data=(x, y[, sample_weight])
        if y is None and self._is_sequence(x):
            n = len(x) if steps is None else min(len(x), int(steps))
            for i in range(n):
                yield self._split_item(x[i])
            return
        if y is None and self._is_iterator(x):
            it = iter(x)
            i = 0
            while steps is None or i < int(steps):
                try:
                    item = next(it)
                except StopIteration:
                    return
                yield self._split_item(item)
                i += 1
            return
        xs = np.asarray(x)
        n = xs.shape[0]
        starts = list(range(0, n, batch_size))
        if steps is not None:
            starts = starts[:int(steps)]

        def take(a, s_):
            if isinstance(a, dict):
                return {k_: take(v_, s_) for k_, v_ in a.items()}
            if isinstance(a, (list, tuple)):
                return [take(v_, s_) for v_ in a]
            return a[s_:s_ + batch_size] if order is None else a[order[s_:s_ + batch_size]]      # contiguous slices keep pinned memory pinned
        for s_ in starts:
            yield take(xs, s_), take(y, s_)

    def _train_stream(self, batches):
        """Train on a stream of (x, y) host batches with the input pipeline overlapped, the way Keras' fit() drives its data adapter
        (2DCNN/Train.py:281,394-415): a producer thread pulls the next batch from the caller's arrays / Sequence / generator, stages it
        host -> device on a copy stream (two slots per batch shape; pinned sources copy asynchronously, pageable ones block only
        the producer) while the main thread keeps the device busy with the previous step.  No host synchronisation per batch: every
        step's 1 KB log record (loss, per-output losses, metric sums) is copied back asynchronously and read once at the end.
        Returns [(n_samples, log record)]."""
        import queue
        import threading
        import torch
        from .planner import LOSS_BUF_FLOATS
        state = {"slots": {}, "err": None}
        ready_q: "queue.Queue" = queue.Queue(maxsize=2)
        out_shapes = [tuple(n.shape) for n in self.graph.outputs]
        dev = None
        # the first batch decides the device path: engines are built lazily per batch size on the main thread
        it = iter(batches)

        def normalise(bx, by):
            xs = self._to_nhwc(bx)
            ys = self._targets(by)
            ys = [None if t is None else np.ascontiguousarray(t, np.float32) for t in ys]
            for t, shp in zip(ys, out_shapes):
                if t is not None and tuple(t.shape[1:]) != shp:
                    raise ValueError(f"target shape {t.shape} does not match the output shape (None,) + {shp}")
            return xs, ys

        try:
            bx, by = next(it)
        except StopIteration:
            return []
        xs0, ys0 = normalise(bx, by)
        eng0 = self._engine(xs0.shape[0], True)
        on_gpu = eng0.dev.type == "cuda"
        records = []
        if not on_gpu:
            # (emulator engine of the CPU test-suite: same order of operations, synchronous copies)
            def run_sync(xs, ys):
                eng = self._engine(xs.shape[0], True)
                eng.x_dev.copy_(torch.from_numpy(xs))
                for o, t in zip(eng.outputs, ys):
                    if t is not None:
                        o["target"].copy_(torch.from_numpy(t))
                if any(t is None for t in ys):
                    eng.derive_targets()
                self._step(eng, return_loss=False)
                records.append((xs.shape[0], eng.logs_buf.clone().double().numpy()))
            run_sync(xs0, ys0)
            for bx, by in it:
                run_sync(*normalise(bx, by))
            return records

        dev = eng0.dev
        cur = torch.cuda.current_stream(dev)
        copy_stream = torch.cuda.Stream(device=dev)

        def slots_for(B):
            if B not in state["slots"]:
                H, W, Cin = self.graph.inputs[0].shape
                sl = [dict(x=torch.empty((B, H, W, Cin), dtype=torch.float32, device=dev),
                           t=[torch.empty((B,) + shp, dtype=torch.float32, device=dev) for shp in out_shapes],
                           ready=torch.cuda.Event(), free=torch.cuda.Event(), B=B, idx=i) for i in range(2)]
                fq: "queue.Queue" = queue.Queue()
                for sl_ in sl:
                    fq.put(sl_)
                state["slots"][B] = fq
            return state["slots"][B]

        def stage(xs, ys):
            st = slots_for(xs.shape[0]).get()             # blocks until the main thread has released a slot of this shape
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(st["free"])
                st["x"].copy_(torch.from_numpy(xs), non_blocking=True)
                for dst, t in zip(st["t"], ys):
                    if t is not None:
                        dst.copy_(torch.from_numpy(t), non_blocking=True)
                st["ready"].record(copy_stream)
            st["src"] = (xs, ys)                            # keep the host arrays alive until the copy has been consumed
            return st

        def producer():
            try:
                torch.cuda.set_device(dev)
                ready_q.put(stage(xs0, ys0))
                for bx_, by_ in it:
                    ready_q.put(stage(*normalise(bx_, by_)))
            except BaseException as e:      # noqa: BLE001 - re-raised on the main thread
                state["err"] = e
            ready_q.put(None)

        th = threading.Thread(target=producer, daemon=True)
        th.start()
        ring = 256
        log_host = torch.empty((ring, LOSS_BUF_FLOATS), dtype=torch.float32).pin_memory()
        pending = []
        i = 0
        while True:
            st = ready_q.get()
            if st is None:
                break
            eng = self._engine(st["B"], True)
            cur.wait_event(st["ready"])
            eng.x_dev.copy_(st["x"], non_blocking=True)
            derive = False
            for o, t, src in zip(eng.outputs, st["t"], st["src"][1]):
                if src is not None:
                    o["target"].copy_(t, non_blocking=True)
                else:
                    derive = True
            if derive:
                eng.derive_targets()
            st["free"].record(cur)
            state["slots"][st["B"]].put(st)
            self._step(eng, return_loss=False)
            if i and i % ring == 0:         # the ring of pinned log records is about to wrap: collect what is in flight
                cur.synchronize()
                records += [(n_, log_host[j % ring].double().numpy().copy()) for (j, n_) in pending]
                pending = []
            log_host[i % ring].copy_(eng.logs_buf, non_blocking=True)
            pending.append((i, st["B"]))
            i += 1
        th.join()
        cur.synchronize()
        if state["err"] is not None:
            raise state["err"]
        records += [(n_, log_host[j % ring].double().numpy().copy()) for (j, n_) in pending]
        return records

    def _train_logs(self, records):
        """Keras' epoch logs from the per-step device records: sample-weighted means of the total loss, the per-output losses (models
        with several outputs) and the compiled metrics, under Keras' key names"""
        tot = float(sum(n for n, _ in records))
        logs = {"loss": sum(n * float(r[0]) for n, r in records) / tot}
        names = [nm for nm in (_metric_name(mt) for mt in (self.metrics or [])) if nm is not None]
        multi = len(self.output_names) > 1
        for oi, (out_name, node, kind) in enumerate(zip(self.output_names, self.graph.outputs, self._losses)):
            H, W, C = node.shape
            base = 8 + 8 * oi
            if multi:
                logs[f"{out_name}_loss"] = sum(n * float(r[base]) for n, r in records) / tot
            for nm in names:
                which = _accuracy_kind(kind, C) if nm == "accuracy" else nm
                col, per = {"mean_squared_error": (1, H * W * C), "mean_absolute_error": (2, H * W * C), "binary_accuracy": (3, H * W * C),
                            "categorical_accuracy": (4, H * W)}.get(which, (None, None))
                if col is None:
                    continue        # (a metric the training pass does not accumulate: reported for validation only)
                logs[nm if not multi else f"{out_name}_{nm}"] = sum(float(r[base + col]) for n, r in records) / (tot * per)
        return logs

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None, shuffle=True,
            initial_epoch=0, steps_per_epoch=None, validation_steps=None, validation_batch_size=None, **kw):
        """tf.keras.Model.fit for the call forms of the reference (2DCNN/Train.py:281-300, 394-415; 1D notebook cells 36-40): NumPy
        arrays with dict / list targets, a keras.utils.Sequence (CustomDataGenerator), or a generator / zip(generators) with
        steps_per_epoch; validation_data as a tuple, a Sequence or a generator (+ validation_steps); validation_split on arrays."""
        hist = History()
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin({})
        arrays = not (y is None and (self._is_sequence(x) or self._is_iterator(x)))
        split = float(kw.get("validation_split") or 0.0)
        if split and not arrays:
            raise ValueError("`validation_split` is only supported for Tensors or NumPy arrays, found following types in the input: "
                             f"{type(x)}")
        if not arrays and self._is_iterator(x) and not self._is_sequence(x) and steps_per_epoch is None and epochs - initial_epoch > 1:
            raise ValueError("When passing a generator that is consumed once, specify `steps_per_epoch` (the same generator is iterated "
                             "across epochs, as Keras does)")
        bs = int(batch_size or 32)
        if arrays:
            xs = np.asarray(x) if not isinstance(x, np.ndarray) else x
            ys = y
            if split and validation_data is None:
                # Keras holds out the LAST fraction of the samples, before any shuffling (Train.py:411-415)
                if not 0.0 < split < 1.0:
                    raise ValueError(f"`validation_split` must be between 0 and 1, received: {split}")
                cut = int(xs.shape[0] * (1.0 - split))

                def part(a, lo, hi):
                    if isinstance(a, dict):
                        return {k_: v_[lo:hi] for k_, v_ in a.items()}
                    if isinstance(a, (list, tuple)):
                        return [v_[lo:hi] for v_ in a]
                    return a[lo:hi]
                validation_data = (xs[cut:], part(ys, cut, None))
                xs, ys = xs[:cut], part(ys, 0, cut)
            n = xs.shape[0]
        rng = np.random.default_rng(0)
        self.stop_training = False
        gen_iter = iter(x) if (not arrays and not self._is_sequence(x)) else None
        for ep in range(initial_epoch, epochs):
            t0 = time.time()
            if arrays:
                order = rng.permutation(n) if shuffle else None
                stream = self._host_batches(xs, ys, bs, steps_per_epoch, order)
            elif gen_iter is not None:
                stream = self._host_batches(gen_iter, None, bs, steps_per_epoch)
            else:
                stream = self._host_batches(x, None, bs, steps_per_epoch)
            records = self._train_stream(stream)
            if not records:
                raise ValueError("fit: the data source produced no batches")
            if not arrays and hasattr(x, "on_epoch_end"):
                x.on_epoch_end()
            logs = self._train_logs(records)
            if validation_data is not None:
                for k_, v_ in self.evaluate(validation_data, None, batch_size=validation_batch_size or batch_size or 32,
                                            steps=validation_steps, return_dict=True).items():
                    logs[f"val_{k_}"] = v_
                if hasattr(validation_data, "on_epoch_end"):
                    validation_data.on_epoch_end()
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
