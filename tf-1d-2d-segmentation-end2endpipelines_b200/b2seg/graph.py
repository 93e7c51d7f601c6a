"""Static layer graph emitted by the builder classes (the role tf.keras' functional API plays in the reference).

Nodes are Keras-level layers with Keras-2 auto-names generated in *call order* (one counter per layer class,
fresh per model build — what tf.keras.backend.clear_session() followed by a builder call produces), so weights can be
exchanged with the reference by layer name.  Tensors are channels-last; 1D tensors are (L, C) and are stored
as (H=1, W=L, C).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass(eq=False)
class Node:
    op: str                      # input conv tconv bn act pool up concat add mul pow convlstm flatten dense reshape
    name: str
    inputs: List["Node"]
    shape: Tuple[int, int, int]  # (H, W, C) of the output, batch excluded
    attrs: dict = field(default_factory=dict)
    idx: int = -1                # creation index

    def __repr__(self):
        return f"<{self.op} {self.name} {self.shape}>"

    @property
    def C(self):
        return self.shape[2]


_BASE_NAMES = {
    ("conv", 2): "conv2d", ("conv", 1): "conv1d", ("tconv", 2): "conv2d_transpose", ("tconv", 1): "conv1d_transpose",
    ("bn", 0): "batch_normalization", ("act", 0): "activation", ("pool", 2): "max_pooling2d", ("pool", 1): "max_pooling1d",
    ("up", 2): "up_sampling2d", ("up", 1): "up_sampling1d", ("concat", 0): "concatenate", ("add", 0): "add",
    ("mul", 0): "tf.math.multiply", ("convlstm", 2): "conv_lstm2d", ("convlstm", 1): "conv_lstm1d", ("flatten", 0): "flatten",
    ("dense", 0): "dense", ("reshape", 0): "reshape", ("input", 0): "input",
    ("oper", 2): "oper2d", ("oper", 1): "oper1d", ("opert", 2): "oper2d_transpose", ("opert", 1): "oper1d_transpose",
}


# tf.keras.activations.get(name) in Keras 2: the functions of keras.activations by their (lower-case) names, plus the advanced
# activation LAYER classes by their class names — which is why the reference's Activation('ReLU') / Activation('LeakyReLU')
# (2DCNN unet_variants.py:7,17) work, why final_activation="Softmax" (Train.py:54 suggests that spelling) works, and why
# "Sigmoid" or "Linear" raise.
_KERAS_ACTIVATION_FUNCTIONS = {"linear", "relu", "sigmoid", "softmax", "tanh", "elu", "selu", "gelu", "swish", "softplus", "softsign",
                               "exponential", "hard_sigmoid", "leaky_relu", "relu6", "silu", "mish", "log_softmax"}
_KERAS_ACTIVATION_LAYERS = {"ReLU": "ReLU", "LeakyReLU": "LeakyReLU", "Softmax": "softmax", "PReLU": "PReLU", "ELU": "elu",
                            "ThresholdedReLU": "ThresholdedReLU"}


def keras_activation_name(name):
    """canonical name of a Keras-2 activation identifier; ValueError for what tf.keras.activations.get would not resolve"""
    if name is None or callable(name):
        return name
    if name in _KERAS_ACTIVATION_FUNCTIONS:
        return name
    if name in _KERAS_ACTIVATION_LAYERS:
        return _KERAS_ACTIVATION_LAYERS[name]
    raise ValueError(f"Unknown activation function: '{name}'. Please ensure you are using a `keras.utils.custom_object_scope` "
                     f"and that this object is included in the scope.")


class Graph:
    """Builder-side graph: layer factory methods mirror the tf.keras.layers calls the reference makes."""

    def __init__(self, ndim: int):
        assert ndim in (1, 2)
        self.ndim = ndim
        self.nodes: List[Node] = []
        self._counters: Dict[str, int] = {}
        self.inputs: List[Node] = []
        self.outputs: List[Node] = []
        self.name = "model"

    # -- naming ---------------------------------------------------------------------------------------------
    def _auto_name(self, op: str) -> str:
        base = _BASE_NAMES.get((op, self.ndim)) or _BASE_NAMES[(op, 0)]
        if base == "input":
            k = self._counters.get(base, 0) + 1
            self._counters[base] = k
            return f"input_{k}"
        k = self._counters.get(base, 0)
        self._counters[base] = k + 1
        return base if k == 0 else f"{base}_{k}"

    def _add(self, op, inputs, shape, name=None, **attrs) -> Node:
        # Keras consumes an auto-name counter only for unnamed layers
        n = Node(op, name if name is not None else self._auto_name(op), list(inputs), tuple(int(s) for s in shape), attrs, len(self.nodes))
        self.nodes.append(n)
        return n

    # -- layers ---------------------------------------------------------------------------------------------
    def input(self, H, W, C) -> Node:
        n = self._add("input", [], (H, W, C))
        self.inputs.append(n)
        return n

    @staticmethod
    def _pair(v, ndim):
        if isinstance(v, (tuple, list)):
            return (int(v[0]), int(v[1])) if len(v) == 2 else (1, int(v[0]))
        return (int(v), int(v)) if ndim == 2 else (1, int(v))

    @staticmethod
    def _positive_filters(filters):
        # Keras' Conv.__init__: MultiResBlock asks for int(alpha * W * 0.167) filters (unet_variants.py:87-89), which is 0 below W = 6
        if int(filters) <= 0:
            raise ValueError(f"Invalid value for argument `filters`. Expected a strictly positive value. Received filters={filters}.")

    def conv(self, x: Node, filters, kernel, strides=1, padding="valid", activation=None, kernel_initializer="glorot_uniform", name=None) -> Node:
        self._positive_filters(filters)
        kh, kw = self._pair(kernel, self.ndim)
        sh, sw = self._pair(strides, self.ndim)
        H, W, _ = x.shape
        if padding == "same":
            Ho, Wo = -(-H // sh), -(-W // sw)
        else:
            Ho, Wo = (H - kh) // sh + 1, (W - kw) // sw + 1
        act = keras_activation_name(activation)
        fused = act if act in (None, "linear", "sigmoid", "softmax") else None
        c = self._add("conv", [x], (Ho, Wo, filters), name, filters=int(filters), kernel=(kh, kw), strides=(sh, sw), padding=padding,
                      activation=fused, init=kernel_initializer)
        if fused is act:
            return c
        # Conv(..., activation='relu' | 'tanh' | ...) — only ever a head here (final_activation, 2DCNN unet_variants.py:1106): the layer is
        # lowered as the convolution (which keeps the layer's name, hence its weight keys) followed by an Activation that carries
        # the layer's name as a model output; the tensor called `name` in Keras is the activated one, so the raw one is not tapped
        c.attrs["no_tap"] = True
        return self._add("act", [c], c.shape, f"{c.name}/activation", fn=act, output_name=c.name)

    def tconv(self, x: Node, filters, kernel, strides, padding="same", name=None) -> Node:
        self._positive_filters(filters)
        kh, kw = self._pair(kernel, self.ndim)
        sh, sw = self._pair(strides, self.ndim)
        assert padding == "same"
        H, W, _ = x.shape
        return self._add("tconv", [x], (H * sh, W * sw, filters), name, filters=int(filters), kernel=(kh, kw), strides=(sh, sw),
                         padding=padding, init="glorot_uniform")

    def bn(self, x: Node, name=None) -> Node:
        return self._add("bn", [x], x.shape, name, eps=1e-3, momentum=0.99)

    def act(self, x: Node, fn: str, name=None) -> Node:
        return self._add("act", [x], x.shape, name, fn=keras_activation_name(fn))

    def pool(self, x: Node, size) -> Node:
        ph, pw = self._pair(size, self.ndim)
        H, W, C = x.shape
        return self._add("pool", [x], (H // ph, W // pw, C), size=(ph, pw))

    def up(self, x: Node, size, interpolation="nearest") -> Node:
        fh, fw = self._pair(size, self.ndim)
        H, W, C = x.shape
        return self._add("up", [x], (H * fh, W * fw, C), size=(fh, fw), interpolation=interpolation)

    def concat(self, xs: List[Node]) -> Node:
        H, W, _ = xs[0].shape
        for t in xs:
            if t.shape[:2] != (H, W):
                raise ValueError(f"A `Concatenate` layer requires inputs with matching shapes except for the concatenation axis. Received: {[t.shape for t in xs]}")
        return self._add("concat", xs, (H, W, sum(t.C for t in xs)))

    def add(self, xs: List[Node]) -> Node:
        for t in xs:
            if t.shape != xs[0].shape:
                raise ValueError(f"Inputs have incompatible shapes. Received shapes {[t.shape for t in xs]}")
        return self._add("add", xs, xs[0].shape)

    def mul(self, a: Node, b: Node) -> Node:
        """a * b with b broadcast over channels when b has one channel (tf operator overload in Attention_Block)."""
        if a.shape[:2] != b.shape[:2] or b.C not in (1, a.C):
            raise ValueError(f"Incompatible shapes for multiply: {a.shape} vs {b.shape}")
        return self._add("mul", [a, b], a.shape)

    def pow(self, x: Node, p: int, name=None) -> Node:
        """tf.math.pow(x, p), integer p >= 2 (operational layers, 2DCNN/models/onn_layers.py:19,41)"""
        if not 2 <= int(p) <= 8:
            raise ValueError(f"power {p} outside 2..8")
        return self._add("pow", [x], x.shape, name, p=int(p))

    def oper(self, x: Node, filters, kernel, q=1, strides=1, padding="same", activation=None, transpose=False) -> Node:
        """Self-ONN operational layer Oper2D / Oper1D / Oper2DTranspose / Oper1DTranspose (2DCNN/models/onn_layers.py:6-48,
        1DCNN/Models/ONN_layers.py): a nested tf.keras.Model named oper2d[_k] holding q convolutions ONN_Conv_1..q
        (ONN_TransConv_1..q); call() = conv_1(x) + sum_{p=2..q} conv_p(x ** p), then an optional Activation.
        Lowered as what it is: q convolutions, q-1 element-wise powers and the running sum.  Only the nested model consumes an
        auto-name counter (its convolutions and its Activation are explicitly named)."""
        base = self._auto_name("opert" if transpose else "oper")
        stem = "ONN_TransConv" if transpose else "ONN_Conv"
        terms = []
        for i in range(1, int(q) + 1):
            xi = x if i == 1 else self.pow(x, i, name=f"{base}/tf_math_pow{i - 1}")
            if transpose:
                terms.append(self.tconv(xi, filters, kernel, strides, padding=padding, name=f"{base}/{stem}_{i}"))
            else:
                terms.append(self.conv(xi, filters, kernel, strides=strides, padding=padding, name=f"{base}/{stem}_{i}"))
        acc = terms[0]
        rest = terms[1:]
        k = 0
        while rest:                      # `x += ...` (onn_layers.py:19): a left fold; three-way sums keep it to one pass for q = 3
            take, rest = rest[:2], rest[2:]
            last = not rest and activation is None
            acc = self._add("add", [acc] + take, acc.shape, base if last else f"{base}/add" + (f"_{k}" if k else ""))
            k += 1
        if activation is not None:
            acc = self._add("act", [acc], acc.shape, base, fn=keras_activation_name(activation))
        # the node called `base` is the nested model's output (what Keras reports as the output of layer oper2d[_k]);
        # with q = 1 and no activation that is the convolution itself, which keeps its ONN_Conv_1 name
        return acc

    def convlstm(self, xs: List[Node], filters, kernel, name=None) -> Node:
        """ConvLSTM over a length-1 sequence whose single frame is the channel-concat of xs
        (Reshape(1,...) + concatenate(axis=-1) + ConvLSTM(return_sequences=False, go_backwards=True))."""
        kh, kw = self._pair(kernel, self.ndim)
        H, W, _ = xs[0].shape
        for t in xs:
            if t.shape[:2] != (H, W):
                raise ValueError("ConvLSTM inputs must share spatial shape")
        return self._add("convlstm", xs, (H, W, filters), name, filters=int(filters), kernel=(kh, kw), cin=sum(t.C for t in xs), init="he_normal")

    def flatten(self, x: Node) -> Node:
        H, W, C = x.shape
        return self._add("flatten", [x], (1, 1, H * W * C))

    def dense(self, x: Node, units, name=None) -> Node:
        return self._add("dense", [x], (1, 1, units), name, units=int(units), init="glorot_uniform")

    def reshape(self, x: Node, H, W, C) -> Node:
        assert H * W * C == x.shape[0] * x.shape[1] * x.shape[2]
        return self._add("reshape", [x], (H, W, C))

    # -- finalisation ---------------------------------------------------------------------------------------
    def finalize(self, outputs: List[Node], name="model") -> "Graph":
        """tf.keras.Model(inputs, outputs): keep only ancestors of the outputs (dangling layers are dropped)."""
        self.outputs = list(outputs)
        self.name = name
        keep = set()
        stack = list(outputs)
        while stack:
            n = stack.pop()
            if id(n) in keep:
                continue
            keep.add(id(n))
            stack.extend(n.inputs)
        self.nodes = [n for n in self.nodes if id(n) in keep]
        return self

    def consumers(self) -> Dict[int, List[Node]]:
        cons: Dict[int, List[Node]] = {id(n): [] for n in self.nodes}
        for n in self.nodes:
            for i in n.inputs:
                cons[id(i)].append(n)
        return cons

    # -- parameters (Keras layouts) -------------------------------------------------------------------------
    def param_specs(self) -> List[Tuple[str, str, Tuple[int, ...], str, bool]]:
        """[(layer_name, weight_name, keras_shape, initializer, trainable)] in layer creation order."""
        out = []
        for n in self.nodes:
            a = n.attrs
            if n.op == "conv":
                kh, kw = a["kernel"]
                cin = n.inputs[0].C
                ks = (kh, kw, cin, a["filters"]) if self.ndim == 2 else (kw, cin, a["filters"])
                out += [(n.name, "kernel", ks, a["init"], True), (n.name, "bias", (a["filters"],), "zeros", True)]
            elif n.op == "tconv":
                kh, kw = a["kernel"]
                cin = n.inputs[0].C
                ks = (kh, kw, a["filters"], cin) if self.ndim == 2 else (kw, a["filters"], cin)
                out += [(n.name, "kernel", ks, a["init"], True), (n.name, "bias", (a["filters"],), "zeros", True)]
            elif n.op == "bn":
                c = n.C
                out += [(n.name, "gamma", (c,), "ones", True), (n.name, "beta", (c,), "zeros", True),
                        (n.name, "moving_mean", (c,), "zeros", False), (n.name, "moving_variance", (c,), "ones", False)]
            elif n.op == "convlstm":
                kh, kw = a["kernel"]
                F, cin = a["filters"], a["cin"]
                k1 = (kh, kw, cin, 4 * F) if self.ndim == 2 else (kw, cin, 4 * F)
                k2 = (kh, kw, F, 4 * F) if self.ndim == 2 else (kw, F, 4 * F)
                out += [(n.name, "kernel", k1, a["init"], True), (n.name, "recurrent_kernel", k2, "orthogonal", True),
                        (n.name, "bias", (4 * F,), "lstm_bias", True)]
            elif n.op == "dense":
                fin = n.inputs[0].shape[2]
                out += [(n.name, "kernel", (fin, a["units"]), a["init"], True), (n.name, "bias", (a["units"],), "zeros", True)]
        return out

    def count_params(self):
        tr = sum(int(np.prod(s)) for (_, _, s, _, t) in self.param_specs() if t)
        nt = sum(int(np.prod(s)) for (_, _, s, _, t) in self.param_specs() if not t)
        return tr, nt


def init_params(graph: Graph, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Keras default initialisers: he_uniform / glorot_uniform / he_normal(truncated) / zeros / ones.
    Deterministic per (seed, position of the layer among parameterised layers)."""
    params: Dict[str, np.ndarray] = {}
    layer_pos: Dict[str, int] = {}
    for (layer, wname, shape, init, _tr) in graph.param_specs():
        pos = layer_pos.setdefault(layer, len(layer_pos))
        rng = np.random.default_rng([seed, pos, sum(ord(c) for c in wname)])
        key = f"{layer}/{wname}"
        if init == "zeros":
            params[key] = np.zeros(shape, np.float32)
        elif init == "ones":
            params[key] = np.ones(shape, np.float32)
        elif init == "lstm_bias":  # unit_forget_bias: [0_F, 1_F, 0_2F]
            F = shape[0] // 4
            b = np.zeros(shape, np.float32)
            b[F:2 * F] = 1.0
            params[key] = b
        elif init == "orthogonal":
            rows = int(np.prod(shape[:-1]))
            a = rng.standard_normal((max(rows, shape[-1]), min(rows, shape[-1])))
            q, r = np.linalg.qr(a)
            q = q * np.sign(np.diag(r))
            if rows < shape[-1]:
                q = q.T
            params[key] = q[:rows, :shape[-1]].reshape(shape).astype(np.float32)
        else:
            # fan computation as in keras.initializers: receptive field x channels
            if len(shape) == 2:
                fan_in, fan_out = shape
            else:
                rf = int(np.prod(shape[:-2]))
                fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
            if init == "he_uniform":
                lim = np.sqrt(6.0 / fan_in)
                params[key] = rng.uniform(-lim, lim, shape).astype(np.float32)
            elif init == "glorot_uniform":
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                params[key] = rng.uniform(-lim, lim, shape).astype(np.float32)
            elif init == "he_normal":
                std = np.sqrt(2.0 / fan_in) / 0.87962566103423978
                v = rng.standard_normal(shape)
                bad = np.abs(v) > 2
                while bad.any():
                    v[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(v) > 2
                params[key] = (v * std).astype(np.float32)
            else:
                raise ValueError(f"unknown initializer {init}")
    return params
