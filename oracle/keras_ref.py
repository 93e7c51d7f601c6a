"""ORACLE — test infrastructure, NOT product code.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.

CPU restatement (PyTorch eager, float32 or float64) of the tf.keras (Keras-2 API, TF 2.13-2.15) layer semantics
the reference relies on.  The reference (Sakib1263/TF-1D-2D-Segmentation-End2EndPipelines) contains no arithmetic
of its own: every FLOP is a tf.keras.layers call.  TensorFlow is not installable in the build container and the
reference ships no tests, golden vectors or fixtures, so **parity is unpinned by the reference**: the semantics
below are a reading of Keras-2 (items marked † there), cross-checked against independent
NumPy restatements in tests/test_oracle_cpu.py.

Each layer function cites the reference call site it stands in for.  Weights use Keras layouts and Keras-2
auto-names generated in call order (fresh counters per model = after clear_session()).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


class _TeacherForce(torch.autograd.Function):
    """forward: the other implementation's tensor, bit for bit; backward: the gradient goes to the oracle's own tensor.  Lets
    the oracle differentiate ITS layers at the forward values another implementation produced (max-pool routing, ReLU masks and
    every Jacobian are then evaluated on identical numbers)."""

    @staticmethod
    def forward(ctx, y, dev):
        return dev.clone()

    @staticmethod
    def backward(ctx, g):
        return g, None


class KerasRef:
    """Eager mini functional API.  ndim: 1 or 2.  Tensors are channels-last like Keras: (N, L, C) / (N, H, W, C)."""

    def __init__(self, ndim: int, params: Optional[Dict[str, torch.Tensor]] = None, dtype=torch.float32, training=True, seed=1234,
                 strict=False):
        self.ndim = ndim
        self.dtype = dtype
        self.training = training
        self.params: Dict[str, torch.Tensor] = params if params is not None else {}
        self.strict = strict          # strict: every weight must already exist in `params` with the right shape
        self.seed = seed
        self.counters: Dict[str, int] = {}
        self.acts: Dict[str, torch.Tensor] = {}      # layer name -> output tensor (graph retained for .grad)
        self.used: List[str] = []                    # parameter keys in creation order
        self.new_moving: Dict[str, torch.Tensor] = {}
        self.trainable: List[str] = []
        self.logits: Dict[str, torch.Tensor] = {}
        self.override: Optional[Dict[str, torch.Tensor]] = None   # teacher forcing, see _rec
        self.local_err: Dict[str, float] = {}
        self.local_out: Dict[str, torch.Tensor] = {}
        # teacher forcing with gradient flow: tensors named in `cut` are replaced by leaves (the graph is cut there), every other
        # overridden tensor takes the supplied VALUE but stays connected to the oracle graph (straight-through), see _rec
        self.cut: Optional[set] = None
        self._own: Dict[int, tuple] = {}          # id(teacher-forced tensor) -> (tensor, the oracle's own value of it)
        # max-pool layers whose arg-max the other implementation takes on the UNROUNDED activations it recomputes in its backward
        # pass (see MaxPooling); every other pooling layer routes to the first maximum of the stored values
        self.pool_argmax_unrounded: set = set()

    # ---- naming: keras.backend.unique_object_name semantics -----------------------------------------------
    def _name(self, base: str, name: Optional[str]) -> str:
        if name is not None:
            return name
        k = self.counters.get(base, 0)
        self.counters[base] = k + 1
        return base if k == 0 else f"{base}_{k}"

    def _sfx(self):
        return "2d" if self.ndim == 2 else "1d"

    # ---- weights -------------------------------------------------------------------------------------------
    def _weight(self, layer, wname, shape, init, trainable=True):
        key = f"{layer}/{wname}"
        if key in self.params:
            w = self.params[key]
            if tuple(w.shape) != tuple(shape):
                raise ValueError(f"oracle: weight {key} has shape {tuple(w.shape)}, the reference graph needs {tuple(shape)}")
        else:
            if self.strict:
                raise KeyError(f"oracle: weight {key} {tuple(shape)} missing from supplied params")
            w = self._init(shape, init, len(self.used))
            self.params[key] = w
        if w.dtype != self.dtype:
            w = w.to(self.dtype)
            self.params[key] = w
        if trainable and not w.requires_grad and self.training:
            w.requires_grad_(True)
        self.used.append(key)
        if trainable:
            self.trainable.append(key)
        return w

    def _init(self, shape, init, pos):
        g = torch.Generator().manual_seed(self.seed * 7919 + pos)
        if init == "zeros":
            return torch.zeros(shape, dtype=self.dtype)
        if init == "ones":
            return torch.ones(shape, dtype=self.dtype)
        if len(shape) == 2:
            fan_in, fan_out = shape
        else:
            rf = int(np.prod(shape[:-2]))
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        if init == "he_uniform":      # keras VarianceScaling(2, fan_in, uniform): limit sqrt(6/fan_in)
            lim = math.sqrt(6.0 / fan_in)
            return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(self.dtype)
        if init == "glorot_uniform":  # limit sqrt(6/(fan_in+fan_out))
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(self.dtype)
        if init in ("he_normal", "orthogonal"):  # scale only matters for conditioning of the tests
            return (torch.randn(shape, generator=g, dtype=torch.float64) * math.sqrt(2.0 / fan_in)).clamp(-2, 2).to(self.dtype)
        raise ValueError(init)

    # ---- helpers --------------------------------------------------------------------------------------------
    def _cf(self, x):  # channels-last -> torch NCHW (1D: N,C,1,L)
        # .contiguous(): torch 2.11's multi-threaded CPU conv backward crashes on permuted (N,C,1,L) views
        return (x.permute(0, 3, 1, 2) if self.ndim == 2 else x.permute(0, 2, 1).unsqueeze(2)).contiguous()

    def _cl(self, x):
        return x.permute(0, 2, 3, 1) if self.ndim == 2 else x.squeeze(2).permute(0, 2, 1)

    @staticmethod
    def _pair(v, ndim):
        if isinstance(v, (tuple, list)):
            return (int(v[0]), int(v[1])) if len(v) == 2 else (1, int(v[0]))
        return (int(v), int(v)) if ndim == 2 else (1, int(v))

    def _rec(self, name, y):
        """Record a layer output.  Teacher forcing (`override`): when the caller supplies the tensor another implementation
        produced for this layer, the layer-local error is recorded (this layer evaluated by the oracle on the *other
        implementation's* inputs vs what that implementation stored) and the supplied tensor replaces the oracle's for all
        consumers — so rounding noise does not compound through a deep random-init network and each layer is judged on
        its own arithmetic.  `local_out[name]` keeps the graph-connected oracle output for layer-local backward checks."""
        if self.override is not None and name in self.override:
            dev = self.override[name].to(self.dtype)
            if tuple(dev.shape) != tuple(y.shape) and dev.numel() == y.numel() and y.dim() == 2:
                dev = dev.reshape(y.shape)       # a Dense output: (N, units) here, (N, 1, 1, units) in a channels-last implementation
            if tuple(dev.shape) != tuple(y.shape):
                raise ValueError(f"override for {name}: shape {tuple(dev.shape)} != {tuple(y.shape)}")
            self.local_err[name] = float((dev - y.detach()).norm() / (y.detach().norm() + 1e-30))
            self.local_out[name] = y
            if self.cut is not None and name not in self.cut and y.requires_grad:
                # gradient teacher forcing (tests/test_gpu_model.py check_per_layer): value of the other implementation, Jacobian
                # of the oracle — only the tensors in `cut` start a new graph, where the caller injects the other
                # implementation's gradient
                own = y.detach()
                y = _TeacherForce.apply(y, dev)
                self._own[id(y)] = (y, own)
            else:
                y = dev.clone().requires_grad_(self.training)
        if self.training and y.requires_grad:
            y.retain_grad()
        self.acts[name] = y
        return y

    @staticmethod
    def activation_fn(fn, x):
        """tf.keras.layers.Activation(fn): 'LeakyReLU' resolves to the LeakyReLU layer, alpha 0.3 (Keras 2) †."""
        if fn in (None, "linear"):
            return x
        if fn in ("relu", "ReLU"):
            return F.relu(x)
        if fn == "LeakyReLU":
            return F.leaky_relu(x, 0.3)
        if fn == "sigmoid":
            return torch.sigmoid(x)
        if fn in ("softmax", "Softmax"):          # 'Softmax' resolves to the Softmax layer (axis -1), like 'ReLU' / 'LeakyReLU'
            return torch.softmax(x, dim=-1)
        if fn == "tanh":
            return torch.tanh(x)
        raise ValueError(f"Unknown activation function: {fn}")

    # ---- layers ---------------------------------------------------------------------------------------------
    def Input(self, x):
        return x.to(self.dtype)

    def Conv(self, x, filters, kernel, strides=1, padding="valid", activation=None, kernel_initializer="glorot_uniform", name=None):
        """tf.keras.layers.Conv2D / Conv1D (2DCNN/models/unet_variants.py:9,69,1106; 1DCNN/Models/unet_variants.py:55,156,308).
        Cross-correlation; kernel (kh,kw,Cin,Cout); SAME pads total k-1 with the smaller half first †."""
        name = self._name("conv" + self._sfx(), name)
        kh, kw = self._pair(kernel, self.ndim)
        sh, sw = self._pair(strides, self.ndim)
        cin = x.shape[-1]
        kshape = (kh, kw, cin, filters) if self.ndim == 2 else (kw, cin, filters)
        w = self._weight(name, "kernel", kshape, kernel_initializer)
        b = self._weight(name, "bias", (filters,), "zeros")
        w4 = w if self.ndim == 2 else w.unsqueeze(0)
        xt = self._cf(x)
        if padding == "same":
            H, W = xt.shape[2], xt.shape[3]
            th = max((-(-H // sh) - 1) * sh + kh - H, 0)
            tw = max((-(-W // sw) - 1) * sw + kw - W, 0)
            xt = F.pad(xt, (tw // 2, tw - tw // 2, th // 2, th - th // 2))
        y = F.conv2d(xt, w4.permute(3, 2, 0, 1), b, stride=(sh, sw))
        y = self._cl(y)
        if activation not in (None, "linear"):
            self.logits[name] = y  # Keras 2 keeps the logits of sigmoid/softmax heads for BCE/CCE
        y = self.activation_fn(activation, y)
        return self._rec(name, y)

    def ConvTranspose(self, x, filters, kernel, strides, padding="same", name=None):
        """Conv2DTranspose(4x4, s2, 'same') (unet_variants.py:19) == torch ConvTranspose2d(k4,s2,p1);
        Conv1DTranspose(2, s2, 'same') (1DCNN :104) == ConvTranspose1d(k2,s2,p0).  Keras kernel (kh,kw,Cout,Cin), no flip †."""
        name = self._name("conv" + self._sfx() + "_transpose", name)
        kh, kw = self._pair(kernel, self.ndim)
        sh, sw = self._pair(strides, self.ndim)
        cin = x.shape[-1]
        kshape = (kh, kw, filters, cin) if self.ndim == 2 else (kw, filters, cin)
        w = self._weight(name, "kernel", kshape, "glorot_uniform")
        b = self._weight(name, "bias", (filters,), "zeros")
        w4 = w if self.ndim == 2 else w.unsqueeze(0)
        assert padding == "same"
        # 'same': output = input*stride; total crop k - s split evenly (k=4,s=2 -> 1/1; k=2,s=2 -> 0)
        ph, pw = (kh - sh) // 2, (kw - sw) // 2
        assert (kh - sh) % 2 == 0 and (kw - sw) % 2 == 0
        y = F.conv_transpose2d(self._cf(x), w4.permute(3, 2, 0, 1), b, stride=(sh, sw), padding=(ph, pw))
        return self._rec(name, self._cl(y))

    def BatchNormalization(self, x, name=None):
        """tf.keras.layers.BatchNormalization() (unet_variants.py:11): axis -1, eps 1e-3, momentum 0.99; training uses the
        batch mean and biased variance; moving variance takes the Bessel-corrected batch variance on 4-D inputs
        (fused op) and the biased one on 3-D inputs †."""
        name = self._name("batch_normalization", name)
        c = x.shape[-1]
        gamma = self._weight(name, "gamma", (c,), "ones")
        beta = self._weight(name, "beta", (c,), "zeros")
        mm = self._weight(name, "moving_mean", (c,), "zeros", trainable=False)
        mv = self._weight(name, "moving_variance", (c,), "ones", trainable=False)
        red = tuple(range(x.dim() - 1))
        if self.training:
            mean = x.mean(red)
            var = x.var(red, unbiased=False)
            n = x.numel() // c
            uv = var * n / max(n - 1, 1) if self.ndim == 2 else var
            self.new_moving[f"{name}/moving_mean"] = (mm * 0.99 + mean.detach() * 0.01).detach()
            self.new_moving[f"{name}/moving_variance"] = (mv * 0.99 + uv.detach() * 0.01).detach()
        else:
            mean, var = mm, mv
        y = (x - mean) * torch.rsqrt(var + 1e-3) * gamma + beta
        return self._rec(name, y)

    def Activation(self, x, fn, name=None):
        name = self._name("activation", name)
        return self._rec(name, self.activation_fn(fn, x))

    def Oper(self, x, filters, kernel, q=1, strides=1, padding="same", activation=None, transpose=False):
        """Self-ONN operational layer: Oper2D / Oper2DTranspose (2DCNN/models/onn_layers.py:6-25, 29-48) and their 1D twins
        (1DCNN/Models/ONN_layers.py).  A nested tf.keras.Model `oper2d[_k]` with q explicitly named convolutions:
            y = ONN_Conv_1(x) + sum_{i=1}^{q-1} ONN_Conv_{i+1}(tf.math.pow(x, i+1));  y = Activation(activation)(y) if given.
        Recorded tensors: every convolution (`<model>/ONN_Conv_i`), every power (`<model>/tf_math_pow{i}`), the running sum after
        terms 3, 5, ... (`<model>/add[_k]`), and the model output under the model's own name.  Weight names
        `<model>/ONN_Conv_i/{kernel,bias}` †(the nesting of variable names under the sub-model is unverified)."""
        base = self._name("oper" + self._sfx() + ("_transpose" if transpose else ""), None)
        stem = "ONN_TransConv" if transpose else "ONN_Conv"
        one = (lambda t, nm: self.ConvTranspose(t, filters, kernel, strides, padding=padding, name=nm)) if transpose else \
              (lambda t, nm: self.Conv(t, filters, kernel, strides=strides, padding=padding, name=nm))
        y = one(x, f"{base}/{stem}_1")
        k = 0
        for i in range(1, int(q)):
            xp = self._rec(f"{base}/tf_math_pow{i}", torch.pow(x, i + 1))
            y = y + one(xp, f"{base}/{stem}_{i + 1}")
            if (i + 1) % 2 == 1 or i + 1 == q:        # the product sums three terms per pass; same check points
                last = (i + 1 == q) and activation is None
                y = self._rec(base if last else f"{base}/add" + (f"_{k}" if k else ""), y)
                k += 1
        if activation is not None:
            if activation in ("sigmoid", "softmax", "Softmax"):
                self.logits[base] = y
            y = self._rec(base, self.activation_fn(activation, y))
        return y

    def MaxPooling(self, x, size):
        """MaxPooling2D((p,p)) / MaxPooling1D(p): stride = pool, 'valid' (unet_variants.py:357,790)."""
        name = self._name("max_pooling" + self._sfx(), None)
        ph, pw = self._pair(size, self.ndim)
        if name in self.pool_argmax_unrounded and id(x) in self._own and self._own[id(x)][0] is x:
            # Teacher-forced gradient check of an implementation that stores activations in bf16 but routes the pooling gradient
            # to the arg-max of the activations it RECOMPUTES in fp32 (b2seg's fused BatchNorm/pool backward): bf16 rounding makes
            # ~1 % of the windows tie, and the implementation breaks those ties by the unrounded values — which is also what an
            # fp32 reference does.  Same rule here: maximum of the stored values, ties broken by the oracle's own (unrounded)
            # values of the same tensor, then first in scan order.
            xt, own = self._cf(x), self._cf(self._own[id(x)][1])
            N, C, H, W = xt.shape

            def win(t):
                return t[:, :, :H // ph * ph, :W // pw * pw].reshape(N, C, H // ph, ph, W // pw, pw).permute(0, 1, 2, 4, 3, 5).reshape(N, C, H // ph, W // pw, ph * pw)
            xw, ow = win(xt), win(own)
            top = xw.detach().max(-1, keepdim=True).values
            idx = torch.where(xw.detach() == top, ow, torch.full_like(ow, -float("inf"))).argmax(-1, keepdim=True)
            return self._rec(name, self._cl(xw.gather(-1, idx).squeeze(-1)))
        return self._rec(name, self._cl(F.max_pool2d(self._cf(x), (ph, pw))))

    def UpSampling(self, x, size, interpolation="nearest"):
        """UpSampling2D(size, 'bilinear') = tf.image.resize half-pixel bilinear (align_corners=False) †;
        UpSampling1D(size) = repeat (unet_variants.py:37; 1DCNN :122)."""
        name = self._name("up_sampling" + self._sfx(), None)
        fh, fw = self._pair(size, self.ndim)
        xt = self._cf(x)
        if interpolation == "bilinear":
            y = F.interpolate(xt, scale_factor=(fh, fw), mode="bilinear", align_corners=False)
        else:
            y = xt.repeat_interleave(fh, 2).repeat_interleave(fw, 3)
        return self._rec(name, self._cl(y))

    def concatenate(self, xs):
        name = self._name("concatenate", None)
        return self._rec(name, torch.cat(list(xs), dim=-1))

    def add(self, xs):
        name = self._name("add", None)
        y = xs[0]
        for t in xs[1:]:
            y = y + t
        return self._rec(name, y)

    def multiply(self, a, b):
        name = self._name("tf.math.multiply", None)
        return self._rec(name, a * b)

    def ConvLSTM(self, xs, filters, kernel, name=None):
        """ConvLSTM2D/1D(filters, 3, 'same', return_sequences=False, go_backwards=True, kernel_initializer='he_normal') on a
        length-1 sequence made by Reshape(1,...) + concatenate(axis=-1) (unet_variants.py:145-149; 1DCNN :296-299).
        Gate order i,f,c,o; recurrent_activation hard_sigmoid = clip(0.2x+0.5,0,1) (Keras 2) †; h0 = c0 = 0."""
        name = self._name("conv_lstm" + self._sfx(), name)
        x = torch.cat(list(xs), dim=-1)
        kh, kw = self._pair(kernel, self.ndim)
        cin = x.shape[-1]
        k1 = (kh, kw, cin, 4 * filters) if self.ndim == 2 else (kw, cin, 4 * filters)
        k2 = (kh, kw, filters, 4 * filters) if self.ndim == 2 else (kw, filters, 4 * filters)
        w = self._weight(name, "kernel", k1, "he_normal")
        u = self._weight(name, "recurrent_kernel", k2, "orthogonal")
        if f"{name}/bias" not in self.params and not self.strict:
            b0 = torch.zeros(4 * filters, dtype=self.dtype)
            b0[filters:2 * filters] = 1.0  # unit_forget_bias
            self.params[f"{name}/bias"] = b0
        b = self._weight(name, "bias", (4 * filters,), "zeros")
        w4 = w if self.ndim == 2 else w.unsqueeze(0)
        u4 = u if self.ndim == 2 else u.unsqueeze(0)
        xt = self._cf(x)
        z = F.conv2d(F.pad(xt, ((kw - 1) // 2, kw - 1 - (kw - 1) // 2, (kh - 1) // 2, kh - 1 - (kh - 1) // 2)), w4.permute(3, 2, 0, 1), b)
        h0 = torch.zeros(xt.shape[0], filters, xt.shape[2], xt.shape[3], dtype=self.dtype)
        z = z + F.conv2d(F.pad(h0, ((kw - 1) // 2, kw - 1 - (kw - 1) // 2, (kh - 1) // 2, kh - 1 - (kh - 1) // 2)), u4.permute(3, 2, 0, 1))
        zi, zf, zc, zo = torch.split(z, filters, dim=1)
        # the three live gate pre-activations [i | c | o] as one recorded tensor (teacher forcing / gradient cut point: an
        # implementation that stores them in bf16 evaluates the gate non-linearities on the stored values)
        live = self._cf(self._rec(f"{name}/gates", self._cl(torch.cat([zi, zc, zo], dim=1))))
        zi, zc, zo = torch.split(live, filters, dim=1)
        hs = hard_sigmoid
        c0 = torch.zeros_like(zi)
        c1 = hs(zf) * c0 + hs(zi) * torch.tanh(zc)
        h1 = hs(zo) * torch.tanh(c1)
        return self._rec(name, self._cl(h1))

    def Flatten(self, x):
        self._name("flatten", None)
        return x.reshape(x.shape[0], -1)

    def Dense(self, x, units, name=None):
        name = self._name("dense", name)
        w = self._weight(name, "kernel", (x.shape[-1], units), "glorot_uniform")
        b = self._weight(name, "bias", (units,), "zeros")
        return self._rec(name, x @ w + b)

    def Reshape(self, x, shape):
        self._name("reshape", None)
        return x.reshape((x.shape[0],) + tuple(shape))


def hard_sigmoid(x):
    """tf.keras.activations.hard_sigmoid, Keras 2: 0 if x < -2.5, 1 if x > 2.5, else 0.2 * x + 0.5 (the recurrent activation of
    ConvLSTM2D/1D as the reference builds it, unet_variants.py:145-149).  Pinned by the docstring example in tests/golden/keras_doc_kats.json."""
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


# ---- losses (SUM_OVER_BATCH_SIZE reduction = mean over all elements / pixels) and Keras Adam ----------------
def _maybe_convert_labels(y_true):
    """keras.losses._maybe_convert_labels: labels that are all 0 / 1 become -1 / 1 (hinge family)"""
    if bool(((y_true == 0) | (y_true == 1)).all()):
        return 2.0 * y_true - 1.0
    return y_true


def keras_loss(kind: str, y_pred, y_true, logits=None):
    """tf.keras.losses.<X>()(y_true, y_pred) with the default SUM_OVER_BATCH_SIZE reduction, for the classes 2DCNN/utils/tf_losses.py:8-46
    offers (Keras 2.13-2.15 `keras/losses.py` + `keras/backend.py`, restated †).  With a sigmoid/softmax head Keras 2 evaluates the
    cross-entropies from the cached logits † (SURVEY hazard 19); pass them via `logits`."""
    eps = 1e-7
    if kind == "bce":
        if logits is not None:
            return F.binary_cross_entropy_with_logits(logits, y_true)
        p = y_pred.clamp(eps, 1 - eps)                     # backend.binary_crossentropy: clip, then log(p + eps)
        return -(y_true * (p + eps).log() + (1 - y_true) * (1 - p + eps).log()).mean()
    if kind == "cce":
        if logits is not None:
            return -(y_true * torch.log_softmax(logits, -1)).sum(-1).mean()
        p = y_pred / y_pred.sum(-1, keepdim=True)
        return -(y_true * p.clamp(eps, 1 - eps).log()).sum(-1).mean()
    if kind == "mse":
        return ((y_pred - y_true) ** 2).mean()
    if kind == "mae":
        return (y_pred - y_true).abs().mean()
    if kind == "msle":                                     # log(max(., eps) + 1)
        return ((torch.log(y_pred.clamp_min(eps) + 1.0) - torch.log(y_true.clamp_min(eps) + 1.0)) ** 2).mean()
    if kind == "huber":                                    # delta = 1 (tf_losses.py:23)
        d = y_pred - y_true
        return torch.where(d.abs() <= 1.0, 0.5 * d * d, d.abs() - 0.5).mean()
    if kind == "logcosh":
        d = y_pred - y_true
        return (d + F.softplus(-2.0 * d) - math.log(2.0)).mean()
    if kind == "focal":                                    # BinaryFocalCrossentropy(gamma=2, apply_class_balancing=False)
        p_t = y_true * y_pred + (1 - y_true) * (1 - y_pred)
        if logits is not None:
            bce = F.binary_cross_entropy_with_logits(logits, y_true, reduction="none")
        else:
            p = y_pred.clamp(eps, 1 - eps)
            bce = -(y_true * (p + eps).log() + (1 - y_true) * (1 - p + eps).log())
        return ((1.0 - p_t) ** 2 * bce).mean()
    if kind == "poisson":
        return (y_pred - y_true * torch.log(y_pred + eps)).mean()
    if kind == "kld":
        t, p = y_true.clamp(eps, 1.0), y_pred.clamp(eps, 1.0)
        return (t * torch.log(t / p)).sum(-1).mean()
    if kind == "hinge":
        return (1.0 - _maybe_convert_labels(y_true) * y_pred).clamp_min(0).mean()
    if kind == "squared_hinge":
        return ((1.0 - _maybe_convert_labels(y_true) * y_pred).clamp_min(0) ** 2).mean()
    if kind == "mape":
        return (100.0 * ((y_true - y_pred) / y_true.abs().clamp_min(eps)).abs()).mean()
    if kind == "categorical_hinge":
        pos = (y_true * y_pred).sum(-1)
        neg = ((1.0 - y_true) * y_pred).max(-1).values
        return (neg - pos + 1.0).clamp_min(0).mean()
    if kind == "cosine":
        def l2n(v):
            return v * torch.rsqrt((v * v).sum(-1, keepdim=True).clamp_min(1e-12))
        return -(l2n(y_true) * l2n(y_pred)).sum(-1).mean()
    raise ValueError(kind)


def keras_adam_step(w, g, m, v, t, lr=2e-4, b1=0.9, b2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam (utils/tf_optimizers.py:11), Keras-2 update: epsilon is added to sqrt(v), un-corrected †."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    alpha = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    w.sub_(alpha * m / (v.sqrt() + eps))
