"""Planner: lowers a layer graph (b2seg.graph.Graph) to the flat list of C-ABI descriptors a b2seg plan replays.

It plays the role TensorFlow's graph executor + autodiff play for the reference (Model.fit / Model.predict on the
graphs of 2DCNN/models/unet_variants.py and 1DCNN/Models/unet_variants.py):

* fuses Conv -> BatchNormalization -> Activation (-> MaxPooling) chains into
  [tcgen05 conv + statistics] -> [finalize] -> [apply + act (+ pool)],
* realises `concatenate` by handing producers channel windows of one NHWC buffer (no copy),
* emits the backward pass (BN/activation backward with gradient-source summation and max-pool routing, dgrad,
  wgrad, bias gradients, head + loss seed) in reverse order, and the fused Adam update,
* lays all trainable weights out in one flat fp32 arena (+ bf16 shadow) in kernel-friendly layouts and records how
  each maps back to the Keras layout for weight exchange by layer name.

The planner only does address arithmetic through an `alloc(nbytes) -> int` callback, so the same plan can be bound to
device memory (b2seg.engine) or to the CPU descriptor emulator used by the CPU tests.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from . import _lib as L
from . import lowering as lw
from .graph import Graph, Node
from .lowering import TView


def ceil8(c):
    return (c + 7) // 8 * 8


def pick_box(N, H, W, pixels):
    """Mirror of b2::pick_box (csrc/conv_gemm.cu): (bw, bh, bn) with bw*bh*bn == pixels."""
    def best(extent, cap):
        d = 1
        while d * 2 <= cap and extent % (d * 2) == 0:
            d <<= 1
        if d >= 8 or d == extent:
            return d
        p = 1
        while p < extent and p < cap:
            p <<= 1
        return p
    w = best(W, pixels)
    h = best(H, pixels // w)
    return w, h, pixels // (w * h)


def conv_num_mtiles(N, H, W):
    bw, bh, bn = pick_box(N, H, W, 128)
    return -(-W // bw) * -(-H // bh) * -(-N // bn)


def halo_geometry(desc):
    """Mirror of b2::halo_geometry (csrc/conv_halo.cu): (bw, bh, bn) when the halo-tile kernel applies, else None."""
    import os
    if os.environ.get("B2SEG_NO_HALO"):
        return None
    wins = {}
    for g in range(desc.n_groups):
        for t in range(desc.taps_per_group):
            tp = desc.taps[g * desc.taps_per_group + t]
            w = wins.setdefault((g, tp.src), [tp.dh, tp.dh, tp.dw, tp.dw])
            w[0], w[1], w[2], w[3] = min(w[0], tp.dh), max(w[1], tp.dh), min(w[2], tp.dw), max(w[3], tp.dw)
    if len(wins) > 16 or len(wins) % desc.n_groups:
        return None
    KH = max(w[1] - w[0] + 1 for w in wins.values())
    KW = max(w[3] - w[2] + 1 for w in wins.values())
    if KH * KW < 2:
        return None
    per_group = len(wins) // desc.n_groups
    for g in range(desc.n_groups):
        if sum(1 for (gg, _) in wins if gg == g) != per_group:
            return None
    o = desc.out[0]
    if o.H == 1:
        if KH != 1 or o.W % 128:
            return None
        bw, bh = 128, 1
    else:
        if o.W % 8 or o.H % 16:
            return None
        bw, bh = 8, 16
    if (bh + KH - 1) * (bw + KW - 1) * 128 > 24576 or bw + KW - 1 > 256 or bh + KH - 1 > 256:
        return None
    for i in range(desc.n_src):
        if desc.src[i].N != o.N:
            return None
    return bw, bh, 1


def conv_stat_rows(desc, num_sms=148):
    """Mirror of b2::conv_num_stat_rows (csrc/conv_gemm.cu) for a given SM count."""
    o = desc.out[0]
    geo = halo_geometry(desc)
    if geo is not None:
        bw, bh, bn = geo
        m_tiles = -(-o.W // bw) * -(-o.H // bh) * -(-o.N // bn)
    else:
        m_tiles = conv_num_mtiles(o.N, o.H, o.W)
    bn = desc.block_n or (64 if o.C <= 64 else (128 if o.C <= 128 else 256))
    n_tiles = -(-o.C // bn)
    total = desc.n_groups * m_tiles * n_tiles
    return min(total, num_sms) if n_tiles == 1 else desc.n_groups * m_tiles


# ----------------------------------------------------------------------------------------------------------
@dataclass
class ParamEntry:
    key: str                 # "layer/weight" (Keras names)
    keras_shape: Tuple[int, ...]
    kind: str                # conv | tconv | vec | head | dense
    offset: int              # element offset into the flat arenas (trainable) or the moving-stat arena
    size: int                # padded element count
    trainable: bool
    meta: dict = field(default_factory=dict)


@dataclass
class Phys:
    """Physical realisation of a logical tensor: a bf16 view + map from logical channels to physical channels."""
    view: TView
    C: int
    segs: List[Tuple[int, int]]  # [(physical offset, count)] in logical channel order

    @property
    def Cp(self):
        return self.view.C


@dataclass
class GSrc:
    view: TView
    kind: int = 0      # 0 direct, 1 pooled
    pool: Tuple[int, int] = (1, 1)
    premasked: bool = False
    colsums: Optional[Tuple[int, int, int]] = None   # (fp32 rows ptr, n_rows, pitch): per-CTA column sums written by the producer
    head: Optional[dict] = None   # kind 2: a pointwise head whose backward is folded into the producer's BN backward (b2seg_gradsrc)
    pool_name: Optional[str] = None   # kind 1: the MaxPooling layer the gradient is routed through (bookkeeping for the parity tests)


class PlanError(NotImplementedError):
    pass


# canonical loss names (b2seg.model._LOSS_ALIASES maps the Keras identifiers onto them) -> b2seg_loss_desc.kind
LOSS_KINDS = {"bce": 0, "binary_crossentropy": 0, "cce": 1, "categorical_crossentropy": 1, "mse": 2, "mean_squared_error": 2,
              "mae": 3, "mean_absolute_error": 3, "msle": 4, "huber": 5, "logcosh": 6, "focal": 7, "poisson": 8, "kld": 9, "hinge": 10,
              "squared_hinge": 11, "mape": 12, "categorical_hinge": 13, "cosine": 14}
LOSS_BUF_FLOATS = 256          # [0] total weighted loss; [8 + 8 i ..] per-output {loss, sum sq err, sum abs err, binary hits, arg-max hits}
MAX_OUTPUTS = (LOSS_BUF_FLOATS - 8) // 8


VBASE = 1 << 60          # virtual addresses handed out while planning with reuse=True (no real pointer or size lives up there)
ARENA_ALIGN = 1024


class Planner:
    def __init__(self, graph: Graph, batch: int, alloc: Callable[[int], int], training: bool = True,
                 losses: Optional[List[str]] = None, loss_weights: Optional[List[float]] = None, adam=None,
                 stat_rows_fn: Optional[Callable] = None, adam_bucket_bytes: int = 0, reuse: bool = False, shard: Tuple[int, int] = (0, 1)):
        self.shard = shard                  # (rank, world) of a sharded-optimizer data-parallel plan; (0, 1) = every rank updates everything
        self.stat_rows_fn = stat_rows_fn or conv_stat_rows
        # reuse: activation / gradient buffers share one arena by liveness (see _assign_memory); off = one allocation per tensor,
        # every tensor stays readable after the step (what the per-layer parity tests tap)
        self.reuse = reuse
        self._vbufs: List[list] = []        # [virtual address, nbytes, tag]
        self._vnext = VBASE
        # > 0 (data parallel): the optimizer phase is one Adam op per gradient-exchange bucket, in exchange order, so the
        # update of a bucket can run as soon as ITS all-reduce has landed while later buckets are still on the wire
        self.adam_bucket_bytes = adam_bucket_bytes
        self.fuse_heads = os.environ.get("B2SEG_NO_HEAD_FUSION") is None
        self.g = graph
        self.N = batch
        self.alloc_fn = alloc
        self.training = training
        self.ndim = graph.ndim
        self.ops: Dict[int, List[Tuple[int, object, str]]] = {0: [], 1: [], 2: []}
        self.cons = graph.consumers()
        self.phys: Dict[int, Phys] = {}
        self.gsrc: Dict[int, List[GSrc]] = {}
        self.params: List[ParamEntry] = []
        self.pindex: Dict[str, ParamEntry] = {}
        self.n_train = 0
        self.n_moving = 0
        self.buffers: List[Tuple[str, int, int]] = []  # (tag, ptr, nbytes)
        self.taps: Dict[str, Tuple[TView, int, str]] = {}   # layer name -> (view, logical C, 'act'|'raw')
        self.grad_taps: Dict[str, Tuple[TView, int]] = {}
        self.outputs: List[dict] = []
        self.losses = losses
        self.loss_weights = loss_weights
        self.adam = adam or dict(lr=2e-4, beta1=0.9, beta2=0.999, eps=1e-7)
        self.input_ptr = 0
        self.loss_ptr = 0
        self.units: List[dict] = []
        self.catbuf: Dict[int, Phys] = {}
        self.act_bytes = 0
        self.op_info: Dict[Tuple[int, int], dict] = {}
        self._grad_touched = set()
        self.head_prologue: Dict[int, Tuple[int, int, int]] = {}   # tensor id -> (scale ptr, shift ptr, activation) applied by its head
        self.pool_routing: Dict[str, str] = {}   # max-pool layer -> "recomputed" (see _bwd_conv); absent = routed on the stored tensor
        self.grad_ready: Dict[str, int] = {}   # param key -> number of backward ops after which its gradient is final
        self._analyse()
        self._layout_params()

    # ---------------------------------------------------------------------------------------- memory helpers
    def alloc(self, nbytes, tag="act"):
        nbytes = (int(nbytes) + 255) // 256 * 256
        if tag in ("act", "grad"):
            self.act_bytes += nbytes
            if self.reuse:
                nbytes = (nbytes + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN
                ptr = self._vnext
                self._vbufs.append([ptr, nbytes, tag])
                self._vnext += nbytes
                return ptr
        ptr = self.alloc_fn(nbytes, tag)
        self.buffers.append((tag, ptr, nbytes))
        return ptr

    # ---------------------------------------------------------------------------------------- liveness / buffer reuse
    @staticmethod
    def _walk_ptrs(obj, fn):
        """apply fn to every c_uint64 field (device addresses; sizes and strides are signed) of a ctypes structure, recursively"""
        import ctypes as C
        if isinstance(obj, C.Structure):
            for name, typ in obj._fields_:
                if typ is C.c_uint64:
                    new = fn(getattr(obj, name))
                    if new is not None:
                        setattr(obj, name, new)
                elif isinstance(typ, type) and issubclass(typ, (C.Structure, C.Array)):
                    Planner._walk_ptrs(getattr(obj, name), fn)
        elif isinstance(obj, C.Array):
            if obj._type_ is C.c_uint64:
                for i in range(len(obj)):
                    new = fn(obj[i])
                    if new is not None:
                        obj[i] = new
            elif isinstance(obj._type_, type) and issubclass(obj._type_, (C.Structure, C.Array)):
                for i in range(len(obj)):
                    Planner._walk_ptrs(obj[i], fn)

    def _assign_memory(self):
        """Buffer reuse by liveness.  Every op runs on one stream in program order (forward, backward, optimizer), so a tensor is live
        from the first op that touches it to the last; the forward activations the backward pass re-reads stay live across the two
        phases, gradients and streaming-only intermediates die within a few ops, and an inference plan keeps almost nothing.
        Planning hands out virtual addresses; here the program is scanned for the first / last use of every virtual buffer, the
        buffers are packed into one arena (time-ordered best fit; a buffer that dies at op t cannot share memory with one born at
        op t: no op ever reads and writes aliased tensors) and every address in every descriptor is relocated."""
        import bisect
        bufs = self._vbufs
        starts = [b[0] for b in bufs]

        def find(addr):
            i = bisect.bisect_right(starts, addr) - 1
            assert i >= 0 and addr < bufs[i][0] + bufs[i][1], hex(addr)
            return i
        first, last = {}, {}
        t = 0
        for phase in (0, 1, 2):
            for (_op, desc, _note) in self.ops[phase]:
                def see(v, t=t):
                    if v >= VBASE:
                        i = find(v)
                        first.setdefault(i, t)
                        last[i] = t
                    return None
                self._walk_ptrs(desc, see)
                t += 1
        # pack: events in time order, allocations of time t before the releases of time t
        order = sorted(first, key=lambda i: (first[i], -bufs[i][1]))
        by_end: Dict[int, List[int]] = {}
        for i in order:
            by_end.setdefault(last[i], []).append(i)
        free: List[List[int]] = []          # [offset, size], sorted by offset
        top = 0
        offset: Dict[int, int] = {}
        released_upto = -1
        for i in order:
            # release everything that died strictly before this buffer is born
            while released_upto < first[i] - 1:
                released_upto += 1
                for j in by_end.get(released_upto, []):
                    o, n = offset[j], bufs[j][1]
                    k = bisect.bisect_left([f[0] for f in free], o)
                    free.insert(k, [o, n])
                    if k + 1 < len(free) and free[k][0] + free[k][1] == free[k + 1][0]:
                        free[k][1] += free[k + 1][1]
                        del free[k + 1]
                    if k > 0 and free[k - 1][0] + free[k - 1][1] == free[k][0]:
                        free[k - 1][1] += free[k][1]
                        del free[k]
            need = bufs[i][1]
            best = None
            for k, (o, n) in enumerate(free):
                if n >= need and (best is None or n < free[best][1]):
                    best = k
            if best is not None:
                o, n = free[best]
                offset[i] = o
                if n == need:
                    del free[best]
                else:
                    free[best] = [o + need, n - need]
            elif free and free[-1][0] + free[-1][1] == top:       # grow the arena from its last free block
                o, n = free.pop()
                offset[i] = o
                top = o + need
            else:
                offset[i] = top
                top += need
        self.arena_bytes = max(top, ARENA_ALIGN)
        base = self.alloc_fn(self.arena_bytes, "arena")
        self.buffers.append(("arena", base, self.arena_bytes))
        assert base % 256 == 0

        def reloc(v):
            if v < VBASE:
                return None
            i = find(v)
            return base + offset.get(i, 0) + (v - bufs[i][0])
        for phase in (0, 1, 2):
            for (_op, desc, _note) in self.ops[phase]:
                self._walk_ptrs(desc, reloc)
        from dataclasses import replace as _replace

        def rv(view):
            return _replace(view, ptr=reloc(view.ptr)) if view.ptr >= VBASE else view
        self.taps = {k: (rv(v[0]),) + tuple(v[1:]) for k, v in self.taps.items()}
        self.grad_taps = {k: (rv(v[0]),) + tuple(v[1:]) for k, v in self.grad_taps.items()}
        self.reuse_stats = dict(tensors=len(bufs), tensor_bytes=sum(b[1] for b in bufs), arena_bytes=self.arena_bytes)

    def new_act(self, H, W, Cp, tag="act") -> TView:
        return TView.dense(self.alloc(self.N * H * W * Cp * 2, tag), self.N, H, W, Cp)

    def emit(self, phase, op, desc, note="", flops=0.0, bytes_=0.0):
        """flops: algorithmic FLOPs (2*MAC on logical channels) of a tensor-core op; bytes_: algorithmic HBM bytes of a streaming op"""
        self.op_info[(phase, len(self.ops[phase]))] = dict(flops=float(flops), bytes=float(bytes_), note=note, op=op)
        self.ops[phase].append((op, desc, note))

    def _conv_flops(self, n: Node) -> float:
        kh, kw = n.attrs["kernel"]
        cin, co = n.inputs[0].C, n.attrs["filters"]
        if n.op == "tconv":
            H, W, _ = n.inputs[0].shape
        else:
            H, W, _ = n.shape
        return 2.0 * self.N * H * W * cin * co * kh * kw

    # ---------------------------------------------------------------------------------------- graph analysis
    def _sole(self, n: Node, op: str) -> Optional[Node]:
        c = self.cons[id(n)]
        if len(c) == 1 and c[0].op == op and n not in self.g.outputs:
            return c[0]
        return None

    def _analyse(self):
        g = self.g
        absorbed = set()
        self.flat_concat: Dict[int, List[Node]] = {}
        self.absorbed_concat = set()
        # flatten nested concatenations (Concat_Block is a left fold of concatenate layers)
        # a ConvLSTM's frame is the channel-concat of its inputs: it owns an input concat buffer like a concatenate layer
        for n in g.nodes:
            if n.op in ("concat", "convlstm"):
                parts = []
                for i in n.inputs:
                    if i.op == "concat" and len(self.cons[id(i)]) == 1 and i not in g.outputs:
                        parts += self.flat_concat[id(i)]
                        absorbed.add(id(i))
                        self.absorbed_concat.add(id(i))
                    else:
                        parts.append(i)
                self.flat_concat[id(n)] = parts
        # ---- channel-layout classes.  Tensors related by an element-wise layer (BN, activation, pool, up-sampling, add, the
        # attention multiply) must share one physical channel layout.  A class containing a concatenate output inherits that
        # concat's (possibly gapped: odd channel counts are padded per slot) layout, and the convolutions feeding the class
        # are made to PRODUCE it by scattering their weight rows — MultiResBlock's `add([shortcut, BN(concat(...))])`.
        parent = {id(n): id(n) for n in g.nodes}

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        def union(a, b):
            ra, rb = find(id(a)), find(id(b))
            if ra != rb:
                parent[ra] = rb

        for n in g.nodes:
            if n.op in ("bn", "act", "pool", "up", "pow"):
                union(n, n.inputs[0])
            elif n.op == "add":
                for i in n.inputs:
                    union(n, i)
            elif n.op == "mul":
                union(n, n.inputs[0])
        self._find = find
        self._cls_concat: Dict[int, List[Node]] = {}
        for n in g.nodes:
            if n.op == "concat" and id(n) not in self.absorbed_concat:
                self._cls_concat.setdefault(find(id(n)), []).append(n)
        self._segs_memo: Dict[int, List[Tuple[int, int]]] = {}
        self._cphys_memo: Dict[int, int] = {}
        self._match_gates(absorbed)
        for n in g.nodes:
            if id(n) in absorbed:
                continue
            if n.op == "input":
                self.units.append(dict(kind="input", node=n, out=n))
            elif n.op in ("conv", "tconv") and id(n) in self.gate_proj:
                # a 1x1 projection of a fused attention gate: convolution + atomic column sums only (b2seg_gate_fwd does the rest)
                self.units.append(dict(kind="conv", node=n, bn=None, act=None, pool=None, out=n, gate=self.gate_proj[id(n)]))
            elif n.op in ("conv", "tconv"):
                a = n.attrs
                is_head = (n.op == "conv" and a["kernel"] == (1, 1) and a["filters"] <= 8 and n in g.outputs)
                if is_head:
                    self.units.append(dict(kind="head", node=n, out=n))
                    continue
                wide_head = None
                if a.get("activation") not in (None, "linear"):
                    if not (n in g.outputs and a["activation"] in ("sigmoid", "softmax")):
                        raise PlanError(f"conv {n.name}: fused activation only supported on output heads")
                    # an output head the pointwise-head kernels do not take (more than 8 classes, or a k > 1 kernel): the
                    # convolution runs on the tensor-core kernels, its activation in the output op
                    wide_head = a["activation"]
                u = dict(kind="conv", node=n, bn=None, act=None, pool=None)
                if a.get("no_tap"):
                    u["no_tap"] = True
                if wide_head is not None:
                    u["out"] = n
                    u["no_tap"] = True        # the tensor called `n.name` is the activated output, not these logits
                    self.units.append(u)
                    self.units.append(dict(kind="outact", node=n, out=None, src=n, fn=wide_head))
                    continue
                last = n
                b = self._sole(last, "bn")
                if b is not None:
                    u["bn"] = b
                    absorbed.add(id(b))
                    last = b
                c = self._sole(last, "act")
                # tanh (Self-ONN decoders) lives in the streaming kernels only: fused when a BatchNorm sits in between
                if c is not None and c not in g.outputs and (c.attrs["fn"] in ("relu", "ReLU", "LeakyReLU", "sigmoid")
                                                             or (c.attrs["fn"] == "tanh" and u["bn"] is not None)):
                    u["act"] = c
                    absorbed.add(id(c))
                    last = c
                u["out"] = last
                if u["bn"] is not None and last not in g.outputs:
                    want = (2, 2) if self.ndim == 2 else (1, 2)
                    for p in self.cons[id(last)]:
                        if p.op == "pool" and p.attrs["size"] == want and last.shape[0] % want[0] == 0 and last.shape[1] % want[1] == 0:
                            u["pool"] = p
                            absorbed.add(id(p))
                            break
                self.units.append(u)
            elif n.op == "concat":
                self.units.append(dict(kind="concat", node=n, out=n, parts=self.flat_concat[id(n)]))
            elif n.op == "add":
                u = dict(kind="add", node=n, out=n, act=None)
                c = self._sole(n, "act")
                if c is not None and c not in g.outputs and c.attrs["fn"] in ("relu", "ReLU", "LeakyReLU"):
                    u["act"], u["out"] = c, c
                    absorbed.add(id(c))
                self.units.append(u)
            elif n.op == "up":
                u = dict(kind="up", node=n, out=n, act=None)
                c = self._sole(n, "act")
                if c is not None and c not in g.outputs and c.attrs["fn"] in ("relu", "ReLU", "LeakyReLU", "sigmoid", "tanh"):
                    u["act"], u["out"] = c, c
                    absorbed.add(id(c))
                self.units.append(u)
            elif n.op == "pool":
                self.units.append(dict(kind="pool", node=n, out=n))
            elif n.op == "bn":
                u = dict(kind="bn", node=n, out=n, act=None)
                c = self._sole(n, "act")
                if c is not None and c not in g.outputs and c.attrs["fn"] in ("relu", "ReLU", "LeakyReLU", "sigmoid", "tanh"):
                    u["act"], u["out"] = c, c
                    absorbed.add(id(c))
                self.units.append(u)
            elif n.op == "act":
                if n in g.outputs and n.attrs["fn"] in ("sigmoid", "softmax", "linear"):
                    # a model output that is an Activation over a tensor (the Self-ONN head, unet_variants.py:1107-1108):
                    # fp32 probabilities for the loss straight from the bf16 logits
                    self.units.append(dict(kind="outact", node=n, out=n, src=n.inputs[0], fn=n.attrs["fn"]))
                    continue
                if n.attrs["fn"] not in ("relu", "ReLU", "LeakyReLU", "sigmoid", "tanh"):
                    raise PlanError(f"standalone activation '{n.attrs['fn']}' ({n.name}) is not lowered")
                self.units.append(dict(kind="act", node=n, out=n))
            elif n.op == "pow":
                self.units.append(dict(kind="pow", node=n, out=n))
            elif n.op == "mul" and id(n) in self.gates:
                self.units.append(dict(kind="gate", node=n, out=n, gate=self.gates[id(n)]))
            elif n.op == "mul":
                self.units.append(dict(kind="mul", node=n, out=n))
            elif n.op == "convlstm":
                self.units.append(dict(kind="convlstm", node=n, out=n, parts=self.flat_concat[id(n)]))
            elif n.op in ("flatten", "dense", "reshape"):
                # Feature_Extraction_Block (ae=1; 2DCNN unet_variants.py:41-48, 1DCNN :127-135): Flatten and Reshape are views of
                # the channels-last buffer, Dense is a 1x1 convolution over a (N, 1, 1, features) tensor on the same kernels
                # (a Reshape output may feed a concatenation — the recurrent blocks re-concatenate their input, which with ae=1 is
                # the Feature_Extraction_Block's Reshape, 1DCNN unet_variants.py:63-72,1008 — it is then copied into the slot)
                if n.op != "reshape" and any(c.op in ("concat", "convlstm") for c in self.cons[id(n)]):
                    raise PlanError(f"{n.name}: a {n.op} output feeding a concatenation is not lowered")
                self.units.append(dict(kind=n.op, node=n, out=n))
            else:
                raise PlanError(f"layer type '{n.op}' ({n.name}) is not lowered yet")
            if n in g.outputs and self.units[-1]["kind"] not in ("head", "outact"):
                # any other tensor used as a model output (a Self-ONN deep-supervision level is the bare sum of q pointwise
                # convolutions, :653): produced by its own unit, then exported as a linear output
                last = self.units[-1]
                if last["out"] is not n:
                    raise PlanError(f"output {n.name} ({n.op}) is not lowered")
                self.units.append(dict(kind="outact", node=n, out=None, src=n, fn="linear"))
        self.unit_of_out = {id(u["out"]): u for u in self.units if u["out"] is not None}
        for u in self.units:
            if u.get("pool") is not None:
                self.unit_of_out[id(u["pool"])] = u
        self._plan_bn_glue()
    # ---------------------------------------------------------------------------------------- MultiResBlock / ResPath glue
    def _plan_bn_glue(self):
        """Per-level fusion of the MultiResBlock / ResPath chains (2DCNN/models/unet_variants.py:85-122; 1DCNN :173-219).  Every
        BatchNormalization is a grid-wide reduction, i.e. a kernel boundary in training; what goes is every pass that only moves data:
          * a BatchNormalization whose input is a concat of Conv_Block outputs, or the output of Add + ReLU, gets its batch
            statistics from the kernels that WRITE that input (bn_act.out_stats) instead of a statistics pass of its own;
          * BatchNormalization(concat) -> Add([shortcut, .]) -> ReLU is one pass (bn_act.add): the normalised tensor and the sum are
            never written;
          * the backward of that ReLU is folded into the backward of the BatchNormalization that reads it (bn_bwd.x_relu_mask).
        B2SEG_NO_BN_GLUE=1 keeps the op-by-op lowering."""
        if os.environ.get("B2SEG_NO_BN_GLUE"):
            return
        relu = ("relu", "ReLU")

        def add_relu_unit(t):
            """the Add + ReLU unit producing tensor t, if t has no other reader"""
            ua = self.unit_of_out.get(id(t))
            if ua is None or ua["kind"] != "add" or ua["act"] is None or ua["act"].attrs["fn"] not in relu or ua["out"] is not t:
                return None
            return ua if (len(self.cons[id(t)]) == 1 and t not in self.g.outputs and len(ua["node"].inputs) == 2) else None

        for ub in self.units:
            if ub["kind"] != "bn":
                continue
            t = ub["node"].inputs[0]
            dense = list(self._segs(t)) == [(0, t.C)] or True      # (gapped layouts are fine: statistics and masks are per physical lane)
            # ---- deferred apply: BN (no activation) whose only reader is Add + ReLU
            if ub["act"] is None and ub["out"] not in self.g.outputs and len(self.cons[id(ub["out"])]) == 1:
                ua = self.unit_of_out.get(id(self.cons[id(ub["out"])][0]))
                if ua is None:       # the consumer is the add node itself (its unit's out may be the absorbed activation)
                    ua = next((u_ for u_ in self.units if u_["kind"] == "add" and u_["node"] is self.cons[id(ub["out"])][0]), None)
                if (ua is not None and ua["kind"] == "add" and ua["act"] is not None and ua["act"].attrs["fn"] in relu and len(ua["node"].inputs) == 2
                        and "fused_bn" not in ua and ua["node"].inputs[0] is not ua["node"].inputs[1]):
                    ub["defer_to"], ua["fused_bn"] = ua, ub
            if not self.training:
                continue
            # ---- statistics from the producers
            ua = add_relu_unit(t)
            if ua is not None and "stats_for" not in ua:
                ua["stats_for"], ub["stats_from"] = ub, "producer"
                ub["x_relu_mask"], ua["bwd_premasked"] = True, True
            elif t.op == "concat" and id(t) not in self.absorbed_concat and len(self.cons[id(t)]) >= 1:
                parts = self.flat_concat[id(t)]
                pus = [self.unit_of_out.get(id(p)) for p in parts]
                ok = all(pu is not None and pu["kind"] == "conv" and pu.get("gate") is None and pu["bn"] is not None and pu["pool"] is None
                         and pu["out"] is p and "stats_for" not in pu and (pu["act"] is None or pu["act"].attrs["fn"] in relu + ("LeakyReLU",))
                         and not self._head_prologue_candidate(pu) for pu, p in zip(pus, parts)) and len(set(map(id, parts))) == len(parts)
                if ok:
                    off = 0
                    for pu, p in zip(pus, parts):
                        pu["stats_for"], pu["stats_off"] = ub, off
                        off += self._cphys(p)
                    ub["stats_from"] = "parts"
            if "stats_from" in ub:
                self._acc_floats += 2 * self._cphys(t)

    def _head_prologue_candidate(self, u) -> bool:
        t = u["out"]
        c = self.cons[id(t)]
        return len(c) == 1 and c[0].op == "conv" and c[0].attrs["kernel"] == (1, 1) and c[0].attrs["filters"] <= 8 and c[0] in self.g.outputs

    def _acc_take(self, n_floats: int) -> int:
        """n_floats of the per-step zeroed accumulator region (BatchNorm sums added into by the kernels that write the tensor)"""
        ptr = self._acc_ptr + 4 * self._acc_used
        self._acc_used += n_floats
        assert self._acc_used <= self._acc_floats
        return ptr

    def _stats_acc(self, ub) -> int:
        if "acc" not in ub:
            ub["acc"] = self._acc_take(2 * self._cphys(ub["node"].inputs[0]))
        return ub["acc"]

    # ---------------------------------------------------------------------------------------- fused attention gates
    def _match_gates(self, absorbed):
        """Recognise Attention_Block (2DCNN/models/unet_variants.py:67-82) by structure:
              mul(skip, add(up2x_bilinear(m), LeakyReLU(tconv4x4s2(m) -> 1)))  with
              m = sigmoid(bn(conv1x1 -> 1 (relu(add(bn(conv1x1 s2 (skip)), bn(conv1x1(gate)))))))
        and lower it onto b2seg_gate_fwd / b2seg_gate_bwd (csrc/gate.cu).  Every interior tensor must have the gate as its only
        consumer; channel counts must be 8 * 2^k with dense layouts.  Anything else (the 1D gate, whose transposed-conv branch has a
        fourth BatchNorm; odd MultiRes widths) keeps the op-by-op lowering.  B2SEG_NO_GATE_FUSION=1 disables the fusion."""
        self.gates: Dict[int, dict] = {}
        self.gate_proj: Dict[int, dict] = {}
        self._acc_floats = 0
        if self.ndim != 2 or os.environ.get("B2SEG_NO_GATE_FUSION"):
            return
        g, cons = self.g, self.cons

        def sole(t, op=None):
            c = cons[id(t)]
            return len(c) == 1 and t not in g.outputs and (op is None or c[0].op == op)

        def p2(c):
            return c % 8 == 0 and (c // 8) & (c // 8 - 1) == 0

        for n in g.nodes:
            if n.op != "mul" or n.inputs[1].C != 1:
                continue
            skip, r = n.inputs
            if r.op != "add" or len(r.inputs) != 2 or not sole(r):
                continue
            r1, r2 = r.inputs
            if not (r1.op == "up" and r1.attrs["size"] == (2, 2) and r1.attrs["interpolation"] == "bilinear" and sole(r1)):
                continue
            if not (r2.op == "act" and r2.attrs["fn"] == "LeakyReLU" and sole(r2)):
                continue
            tc = r2.inputs[0]
            if not (tc.op == "tconv" and tc.attrs["filters"] == 1 and tc.attrs["kernel"] == (4, 4) and tc.attrs["strides"] == (2, 2) and sole(tc)):
                continue
            m = r1.inputs[0]
            if tc.inputs[0] is not m or not (m.op == "act" and m.attrs["fn"] == "sigmoid") or m in g.outputs or len(cons[id(m)]) != 2:
                continue
            bn3 = m.inputs[0]
            if not (bn3.op == "bn" and sole(bn3)):
                continue
            c3 = bn3.inputs[0]
            if not (c3.op == "conv" and c3.attrs["filters"] == 1 and c3.attrs["kernel"] == (1, 1) and c3.attrs["strides"] == (1, 1)
                    and c3.attrs.get("activation") in (None, "linear") and sole(c3)):
                continue
            relu = c3.inputs[0]
            if not (relu.op == "act" and relu.attrs["fn"] in ("relu", "ReLU") and sole(relu)):
                continue
            ab = relu.inputs[0]
            if not (ab.op == "add" and len(ab.inputs) == 2 and sole(ab)):
                continue
            bna, bnb = ab.inputs
            if not (bna.op == "bn" and bnb.op == "bn" and sole(bna) and sole(bnb)):
                continue
            ca, cb = bna.inputs[0], bnb.inputs[0]
            ok = all(c.op == "conv" and c.attrs["kernel"] == (1, 1) and c.attrs.get("activation") in (None, "linear") and sole(c) for c in (ca, cb))
            if not ok or ca.attrs["strides"] != (2, 2) or cb.attrs["strides"] != (1, 1) or ca.inputs[0] is not skip:
                continue
            C, Cs = ca.attrs["filters"], skip.C
            if cb.attrs["filters"] != C or not p2(C) or not p2(Cs) or C > 4096 or skip.shape[0] % 2 or skip.shape[1] % 2:
                continue
            if self._cphys(skip) != Cs or list(self._segs(skip)) != [(0, Cs)] or self._cphys(n) != Cs:
                continue
            gate = dict(mul=n, skip=skip, gating=cb.inputs[0], conv_a=ca, conv_b=cb, bn_a=bna, bn_b=bnb, conv3=c3, bn3=bn3, tconv=tc, C=C, Cs=Cs,
                        interior=[bna, bnb, ab, relu, c3, bn3, m, r1, tc, r2, r])
            self.gates[id(n)] = gate
            self.gate_proj[id(ca)] = gate
            self.gate_proj[id(cb)] = gate
            for t in gate["interior"]:
                absorbed.add(id(t))
        self._acc_floats += sum(4 * gt["C"] + 8 for gt in self.gates.values())   # per gate: sums_a [2][C], sums_b [2][C], sums3 [2] (+ pad)

    def _fwd_gate_proj(self, u):
        """one of the two 1x1 projections of a fused gate: raw output + column sums added into the gate's accumulators"""
        n, gt = u["node"], u["gate"]
        a = n.attrs
        x = self.phys[id(n.inputs[0])]
        pe = self.pindex[f"{n.name}/kernel"]
        pe.meta["segs"] = list(x.segs)
        assert pe.meta["cin_p"] == x.Cp, (n.name, pe.meta["cin_p"], x.Cp)
        C = gt["C"]
        H, W, _ = n.shape
        z = self.new_act(H, W, C)
        self.phys[id(n)] = Phys(z, C, [(0, C)])
        strided = a["strides"] != (1, 1)
        xin = x.view.parity(0, 0, a["strides"][0], a["strides"][1]) if strided else x.view
        cd = lw.conv_fprop(xin, self.pwb(pe.key), C, 1, 1, x.Cp, z, bias=self.pw(f"{n.name}/bias"), act=L.ACT_NONE, stats=0)
        if self.training:
            which = "sums_a" if n is gt["conv_a"] else "sums_b"
            cd.stats, cd.stats_atomic = self._gate_acc(gt)[which], 1
        self.emit(0, L.OP_CONV, cd, n.name, flops=self._conv_flops(n))
        u["y"] = z
        self.taps[n.name] = (z, C, "raw")

    def _gate_acc(self, gt):
        """addresses of the gate's statistics accumulators inside the per-step zeroed region"""
        if "acc" not in gt:
            C = gt["C"]
            base = self._acc_take(4 * C + 8)
            gt["acc"] = dict(sums_a=base, sums_b=base + 8 * C, sums3=base + 16 * C)
        return gt["acc"]

    def _gate_desc(self, gt) -> "L.GateDesc":
        if "desc" in gt:
            return gt["desc"]
        C, Cs = gt["C"], gt["Cs"]
        h, w, _ = gt["conv_a"].shape
        d = L.GateDesc()
        d.za, d.zb = self.phys[id(gt["conv_a"])].view.to_c(), self.phys[id(gt["conv_b"])].view.to_c()
        if self.training:
            acc = self._gate_acc(gt)
            d.sums_a, d.sums_b, d.sums3 = acc["sums_a"], acc["sums_b"], acc["sums3"]
        for tag, bn in (("a", gt["bn_a"]), ("b", gt["bn_b"]), ("3", gt["bn3"])):
            setattr(d, "gamma" + ("_" + tag if tag != "3" else "3"), self.pw(f"{bn.name}/gamma"))
            setattr(d, "beta" + ("_" + tag if tag != "3" else "3"), self.pw(f"{bn.name}/beta"))
            setattr(d, "mm" + ("_" + tag if tag != "3" else "3"), self.pmov(f"{bn.name}/moving_mean"))
            setattr(d, "mv" + ("_" + tag if tag != "3" else "3"), self.pmov(f"{bn.name}/moving_variance"))
        vec = self.alloc(8 * C * 4, "scratch")
        d.vec_a, d.vec_b = vec, vec + 4 * C * 4
        c3, tc = gt["conv3"], gt["tconv"]
        self.pindex[f"{c3.name}/kernel"].meta["segs"] = [(0, C)]
        self.pindex[f"{tc.name}/kernel"].meta["segs"] = [(0, 1)]
        assert self.pindex[f"{c3.name}/kernel"].meta["cin_p"] == C and self.pindex[f"{tc.name}/kernel"].meta["taps"] == 16
        d.w3, d.b3 = self.pw(f"{c3.name}/kernel"), self.pw(f"{c3.name}/bias")
        d.wt, d.bt, d.wt_stride = self.pw(f"{tc.name}/kernel"), self.pw(f"{tc.name}/bias"), self.pindex[f"{tc.name}/kernel"].meta["cin_p"]
        d.z = self.alloc(2 * self.N * h * w * 4, "scratch")
        d.m = d.z + self.N * h * w * 4
        d.training, d.bessel = (1 if self.training else 0), 1
        d.eps, d.momentum, d.count = gt["bn3"].attrs["eps"], gt["bn3"].attrs["momentum"], float(self.N * h * w)
        d.skip = self.phys[id(gt["skip"])].view.to_c()
        gt["desc"] = d
        return d

    def _fwd_gate(self, u):
        n, gt = u["node"], u["gate"]
        skip = self.phys[id(gt["skip"])]
        dests = self._dests(n)
        self.phys[id(n)] = Phys(dests[0], n.C, list(skip.segs))
        d = L.GateDesc.from_buffer_copy(self._gate_desc(gt))
        d.out = dests[0].to_c()
        flops = 0.0
        self.emit(0, L.OP_GATE_FWD, d, n.name, flops=flops)
        self._copy_extra(dests[0], dests[1:])
        self.taps[n.name] = (dests[0], n.C, "act")

    def _bwd_gate(self, u):
        n, gt = u["node"], u["gate"]
        dout = self._single_grad(n)
        if dout is None:
            return
        C = gt["C"]
        h, w, _ = gt["conv_a"].shape
        d = L.GateDesc.from_buffer_copy(self._gate_desc(gt))
        dza, dzb = self.new_act(h, w, C, "grad"), self.new_act(h, w, C, "grad")
        # two-pass form: dskip is written by a second op once the stride-2 projection's dgrad exists (emitted by _bwd_conv of conv_a)
        d.dout, d.dza, d.dzb = dout.to_c(), dza.to_c(), dzb.to_c()
        gt["bwd_desc"] = d
        scr = self.alloc((self.N * 4 * h * w + self.N * h * w + 8 + 3 * C) * 4, "scratch")
        d.dr, d.g3 = scr, scr + self.N * 4 * h * w * 4
        d.bsums3 = d.g3 + self.N * h * w * 4
        d.bsums_ab = d.bsums3 + 32
        for tag, bn in (("_a", gt["bn_a"]), ("_b", gt["bn_b"]), ("3", gt["bn3"])):
            setattr(d, "dgamma" + tag, self.pg(f"{bn.name}/gamma"))
            setattr(d, "dbeta" + tag, self.pg(f"{bn.name}/beta"))
        d.dw3, d.db3 = self.pg(f"{gt['conv3'].name}/kernel"), 0     # (conv3's bias feeds a BatchNorm: exactly zero gradient, not accumulated)
        d.dwt, d.dbt = self.pg(f"{gt['tconv'].name}/kernel"), self.pg(f"{gt['tconv'].name}/bias")
        self.emit(1, L.OP_GATE_BWD, d, f"gate bwd {n.name}")
        self._add_gsrc(gt["conv_a"], GSrc(dza))
        self._add_gsrc(gt["conv_b"], GSrc(dzb))

    # ---------------------------------------------------------------------------------------- parameters
    def _add_param(self, key, keras_shape, kind, size, trainable, **meta):
        size_p = (size + 63) // 64 * 64
        if trainable:
            off = self.n_train
            self.n_train += size_p
        else:
            off = self.n_moving
            self.n_moving += size_p
        e = ParamEntry(key, tuple(keras_shape), kind, off, size_p, trainable, meta)
        self.params.append(e)
        self.pindex[key] = e
        return e

    def _im2col_of(self, n: Node):
        """(kh, kw, Cin) when convolution n reads the network input through a K-packed im2col copy (a 1x1 convolution over kh*kw*Cin
        channels: ONE tensor-core tap instead of kh*kw taps of 8 mostly-padding channels), else None"""
        if os.environ.get("B2SEG_NO_IM2COL") or n.op != "conv" or n.inputs[0].op != "input" or id(n) in self.gate_proj:
            return None
        a = n.attrs
        kh, kw = a["kernel"]
        cin = n.inputs[0].C
        if kh * kw < 2 or a["strides"] != (1, 1) or a["padding"] != "same" or kh * kw * cin > 64:
            return None
        u = self.unit_of_out.get(id(n))
        if u is not None and u["kind"] == "head":
            return None
        return kh, kw, cin

    def _layout_params(self):
        """Assign arena offsets in layer creation order.  Input-channel maps are resolved at emission time."""
        specs = {}
        for (layer, wname, shape, init, tr) in self.g.param_specs():
            specs[(layer, wname)] = (shape, tr)
        for n in self.g.nodes:
            if n.op in ("conv", "tconv"):
                kh, kw = n.attrs["kernel"]
                cin_p = self._cin_phys(n.inputs[0])
                co = n.attrs["filters"]
                u = self.unit_of_out.get(id(n))
                if u is not None and u["kind"] == "head":
                    self._add_param(f"{n.name}/kernel", specs[(n.name, "kernel")][0], "head", cin_p * co, True, cin_p=cin_p, cout=co)
                    self._add_param(f"{n.name}/bias", (co,), "vec", co, True, C=co)
                elif self._im2col_of(n) is not None:
                    ikh, ikw, icin = self._im2col_of(n)
                    cop, osegs = self._cphys(n), self._segs(n)
                    kp = ceil8(ikh * ikw * icin)
                    self._add_param(f"{n.name}/kernel", specs[(n.name, "kernel")][0], "conv", cop * kp, True, cout=co, cout_p=cop, taps=1, cin_p=kp,
                                    kh=1, kw=1, out_segs=osegs, segs=[(0, ikh * ikw * icin)], im2col=(ikh, ikw, icin))
                    self._add_param(f"{n.name}/bias", (co,), "vec", cop, True, C=co, vsegs=osegs, Cp=cop)
                else:
                    cop, osegs = self._cphys(n), self._segs(n)   # the conv produces the layout of its element-wise class
                    self._add_param(f"{n.name}/kernel", specs[(n.name, "kernel")][0], n.op, cop * kh * kw * cin_p, True,
                                    cout=co, cout_p=cop, taps=kh * kw, cin_p=cin_p, kh=kh, kw=kw, out_segs=osegs)
                    self._add_param(f"{n.name}/bias", (co,), "vec", cop, True, C=co, vsegs=osegs, Cp=cop)
            elif n.op == "bn":
                cp, vs = self._cphys(n), self._segs(n)
                self._add_param(f"{n.name}/gamma", (n.C,), "vec", cp, True, C=n.C, fill=1.0, vsegs=vs, Cp=cp)
                self._add_param(f"{n.name}/beta", (n.C,), "vec", cp, True, C=n.C, vsegs=vs, Cp=cp)
                self._add_param(f"{n.name}/moving_mean", (n.C,), "vec", cp, False, C=n.C, vsegs=vs, Cp=cp)
                self._add_param(f"{n.name}/moving_variance", (n.C,), "vec", cp, False, C=n.C, fill=1.0, vsegs=vs, Cp=cp)
            elif n.op == "dense":
                fin, units = n.inputs[0].shape[2], n.attrs["units"]
                fin_p, isegs = fin, [(0, fin)]
                if n.inputs[0].op == "flatten":      # Flatten of a padded / gapped tensor: the kernel rows follow its physical layout
                    t = n.inputs[0].inputs[0]
                    fin_p = t.shape[0] * t.shape[1] * self._cphys(t)
                    isegs = self._flat_segs(t.shape[0], t.shape[1], self._cphys(t), self._segs(t))
                if fin_p % 8 or units % 8:
                    raise PlanError(f"{n.name}: Dense {fin} -> {units}: feature counts must be multiples of 8")
                self._add_param(f"{n.name}/kernel", (fin, units), "conv", units * fin_p, True, cout=units, cout_p=units, taps=1, cin_p=fin_p,
                                kh=1, kw=1, out_segs=[(0, units)], segs=isegs)
                self._add_param(f"{n.name}/bias", (units,), "vec", units, True, C=units, vsegs=[(0, units)], Cp=units)
            elif n.op == "convlstm":
                kh, kw = n.attrs["kernel"]
                F = n.attrs["filters"]
                if F % 8:
                    raise PlanError(f"{n.name}: ConvLSTM filters must be a multiple of 8 (got {F})")
                cin_p = sum(self._cphys(p) for p in self.flat_concat[id(n)])
                # rows ordered [i | g | o | f]: the three live gates form a contiguous [3F][taps][cin] GEMM operand
                self._add_param(f"{n.name}/kernel", specs[(n.name, "kernel")][0], "lstm", 4 * F * kh * kw * cin_p, True,
                                F=F, taps=kh * kw, cin_p=cin_p, kh=kh, kw=kw)
                self._add_param(f"{n.name}/recurrent_kernel", specs[(n.name, "recurrent_kernel")][0], "blob",
                                int(np.prod(specs[(n.name, "recurrent_kernel")][0])), True)
                self._add_param(f"{n.name}/bias", (4 * F,), "lstm_bias", 4 * F, True, F=F)

    def _concat_layout(self, c: Node) -> Tuple[List[Tuple[int, int]], int]:
        segs, off = [], 0
        for p in self.flat_concat[id(c)]:
            for (o, cc) in self._segs(p):
                segs.append((off + o, cc))
            off += self._cphys(p)
        return segs, off

    def _class_layout(self, t: Node) -> Tuple[List[Tuple[int, int]], int]:
        """(segments, physical channel count) shared by every tensor of t's element-wise class (pure function of the graph)"""
        key = self._find(id(t))
        if key not in self._segs_memo:
            fixed = self._cls_concat.get(key, [])
            if t.op == "concat" and id(t) in self.absorbed_concat:
                fixed = []
            if fixed:
                segs, cp = self._concat_layout(fixed[0])
                for other in fixed[1:]:
                    if self._concat_layout(other) != (segs, cp):
                        raise PlanError(f"{fixed[0].name} and {other.name} meet in one element-wise class with different channel layouts")
            else:
                segs, cp = [(0, t.C)], ceil8(t.C)
            self._segs_memo[key], self._cphys_memo[key] = segs, cp
        return self._segs_memo[key], self._cphys_memo[key]

    def _segs(self, t: Node) -> List[Tuple[int, int]]:
        """logical->physical channel segments a tensor will have once materialised"""
        if t.op == "concat" and id(t) in self.absorbed_concat:
            return self._concat_layout(t)[0]
        return self._class_layout(t)[0]

    def _cphys(self, t: Node) -> int:
        if t.op == "concat" and id(t) in self.absorbed_concat:
            return self._concat_layout(t)[1]
        return self._class_layout(t)[1]

    def _cin_phys(self, t: Node) -> int:
        return self._cphys(t)

    def logical_channels(self, name: str) -> List[int]:
        """physical channel index of every logical channel of a tapped layer (identity unless the layout is gapped)"""
        if not hasattr(self, "_by_name"):
            self._by_name = {n.name: n for n in self.g.nodes}
        if name.endswith("/gates"):      # ConvLSTM gate pre-activations: a dense 3F-channel buffer of its own
            return list(range(3 * self._by_name[name[:-6]].attrs["filters"]))
        return [po + i for (po, c) in self._segs(self._by_name[name]) for i in range(c)]

    # ---------------------------------------------------------------------------------------- arenas
    @property
    def arena_elems(self) -> int:
        """length of the flat parameter / gradient / moment arenas: the trainable elements rounded up to 4096, so that any gradient
        exchange bucket cut at a multiple of 4096 splits evenly over 1, 2, 4 or 8 ranks (64-element aligned shards)"""
        return (max(self.n_train, 64) + 4095) // 4096 * 4096

    def bind_arenas(self):
        n = self.arena_elems
        self.w_ptr = self.alloc(n * 4, "param_w")
        self.g_ptr = self.alloc(n * 4, "param_g")
        self.m_ptr = self.alloc(n * 4, "param_m")
        self.v_ptr = self.alloc(n * 4, "param_v")
        self.wb_ptr = self.alloc(n * 2, "param_wb")
        self.mov_ptr = self.alloc(max(self.n_moving, 64) * 4, "moving")

    def pw(self, key):   # fp32 master address
        return self.w_ptr + 4 * self.pindex[key].offset

    def pg(self, key):
        self._grad_touched.add(key)   # the backward emitter of the current unit writes this gradient
        return self.g_ptr + 4 * self.pindex[key].offset

    def pwb(self, key):  # bf16 shadow address
        return self.wb_ptr + 2 * self.pindex[key].offset

    def pmov(self, key):
        return self.mov_ptr + 4 * self.pindex[key].offset

    # ---------------------------------------------------------------------------------------- build
    def build(self):
        self.bind_arenas()
        self._acc_used = 0
        self._acc_ptr = self.alloc(self._acc_floats * 4, "scratch") if (self._acc_floats and self.training) else 0
        if self._acc_ptr:
            # BatchNorm sum accumulators of the fused gates: zeroed once per step, then added into by the projection kernels
            self.emit(0, L.OP_MEMSET, L.MemsetDesc(self._acc_ptr, self._acc_floats * 4), "zero gate statistics")
        self.loss_ptr = self.alloc(LOSS_BUF_FLOATS * 4, "loss")
        if len(self.g.outputs) > MAX_OUTPUTS:
            raise PlanError(f"{len(self.g.outputs)} model outputs (max {MAX_OUTPUTS})")
        for u in self.units:
            getattr(self, "_fwd_" + u["kind"])(u)
        if self.training:
            self.emit(1, L.OP_MEMSET, L.MemsetDesc(self.g_ptr, self.arena_elems * 4), "zero grads")
            self.emit(1, L.OP_MEMSET, L.MemsetDesc(self.loss_ptr, LOSS_BUF_FLOATS * 4), "zero loss")
            for u in reversed(self.units):
                self._grad_touched = set()
                getattr(self, "_bwd_" + u["kind"])(u)
                for key in self._grad_touched:
                    self.grad_ready[key] = len(self.ops[1])
            a = self.adam
            ranges = [(0, self.arena_elems)]
            if self.adam_bucket_bytes > 0:
                ranges = [(lo, hi) for (_n, lo, hi) in self.exchange_schedule(self.adam_bucket_bytes)]
                rank, world = self.shard
                if world > 1:
                    # sharded optimizer: after the bucket's reduce-scatter this rank owns (and updates) one 1/world slice of it;
                    # the bf16 shadow of the whole bucket comes back through an all-gather (Model._step)
                    ranges = [(lo + rank * ((hi - lo) // world), lo + (rank + 1) * ((hi - lo) // world)) for (lo, hi) in ranges]
            for (lo, hi) in ranges:
                self.emit(2, L.OP_ADAM, L.AdamDesc(self.w_ptr + 4 * lo, self.g_ptr + 4 * lo, self.m_ptr + 4 * lo, self.v_ptr + 4 * lo,
                                                   self.wb_ptr + 2 * lo, hi - lo, a["lr"], a["beta1"], a["beta2"], a["eps"], 1.0, 1),
                          "adam" if len(ranges) == 1 else f"adam [{lo}, {hi})")
        if self.reuse:
            self._assign_memory()
        return self

    # -- destinations: where a produced tensor must live ------------------------------------------------------
    def _concat_root(self, c: Node) -> Node:
        """outermost concat a (possibly nested, flattened) concat node belongs to"""
        while id(c) in self.absorbed_concat:
            c = self.cons[id(c)][0]
        return c

    def _concat_buffer(self, c: Node) -> TView:
        """the NHWC buffer holding the channel-concatenation consumed by a concatenate layer (= its output) or a ConvLSTM
        (= its input frame); producers write channel windows of it"""
        if id(c) not in self.catbuf:
            parts = self.flat_concat[id(c)]
            H, W, _ = parts[0].shape
            segs, off = [], 0
            for p in parts:
                for (o, cc) in self._segs(p):
                    segs.append((off + o, cc))
                off += self._cphys(p)
            view = self.new_act(H, W, off)
            self.catbuf[id(c)] = Phys(view, sum(p.C for p in parts), segs)
            if c.op == "concat":
                self.phys[id(c)] = self.catbuf[id(c)]
        return self.catbuf[id(c)].view

    def _dests(self, t: Node) -> List[TView]:
        """views the producer of t should write: one slot per consuming concat / ConvLSTM frame, else one dense buffer"""
        dests = []
        for c in self.cons[id(t)]:
            if c.op in ("concat", "convlstm"):
                root = self._concat_root(c)
                parts = self.flat_concat[id(root)]
                buf = self._concat_buffer(root)
                off = 0
                for p in parts:
                    if p is t:
                        dests.append(buf.chan(off, self._cphys(t)))
                    off += self._cphys(p)
        H, W, _ = t.shape
        if not dests:
            dests.append(self.new_act(H, W, self._cphys(t)))
        self.phys[id(t)] = Phys(dests[0], t.C, list(self._segs(t)))
        return dests

    def _copy_extra(self, src: TView, extra: List[TView]):
        for d in extra:
            self.emit(0, L.OP_ELTWISE, L.EltwiseDesc(1, src.to_c(), lw.NULL_VIEW.to_c(), lw.NULL_VIEW.to_c(), d.to_c()), "concat copy")

    # -- forward emitters ------------------------------------------------------------------------------------
    def _fwd_input(self, u):
        n = u["node"]
        H, W, C = n.shape
        self.input_ptr = self.alloc(self.N * H * W * C * 4, "input")
        self._im2col_views: Dict[Tuple[int, int], TView] = {}
        windows = [self._im2col_of(c) for c in self.cons[id(n)]]
        for win in windows:
            if win is not None and win[:2] not in self._im2col_views:
                kh, kw, _ = win
                col = self.new_act(H, W, ceil8(kh * kw * C))
                self.emit(0, L.OP_CAST, L.CastDesc(self.input_ptr, self.N, H, W, C, col.to_c(), kh, kw), f"{n.name} im2col {kh}x{kw}")
                self._im2col_views[(kh, kw)] = col
        if any(win is None for win in windows) or not windows:
            dests = self._dests(n)      # consumers that read the input as it is
            self.emit(0, L.OP_CAST, L.CastDesc(self.input_ptr, self.N, H, W, C, dests[0].to_c(), 0, 0), n.name)
            self._copy_extra(dests[0], dests[1:])

    def _act_code(self, node: Optional[Node]):
        return L.ACT_NONE if node is None else L.ACT_CODES[node.attrs["fn"]]

    def _fwd_conv(self, u):
        if u.get("gate") is not None:
            return self._fwd_gate_proj(u)
        n: Node = u["node"]
        a = n.attrs
        kh, kw = a["kernel"]
        pe = self.pindex[f"{n.name}/kernel"]
        ic = self._im2col_of(n)
        if ic is not None:      # a 1x1 convolution over the K-packed im2col copy of the network input
            x = Phys(self._im2col_views[ic[:2]], ic[0] * ic[1] * ic[2], [(0, ic[0] * ic[1] * ic[2])])
            kh, kw = 1, 1
        else:
            x = self.phys[id(n.inputs[0])]
            pe.meta["segs"] = list(x.segs)
        assert pe.meta["cin_p"] == x.Cp, (n.name, pe.meta["cin_p"], x.Cp)
        co, cop, cin_p = a["filters"], pe.meta["cout_p"], x.Cp
        H, W, _ = n.shape
        out_node = u["out"]
        act = self._act_code(u["act"])
        bias = self.pw(f"{n.name}/bias")
        strided = n.op == "conv" and a["strides"] != (1, 1)
        if strided and not (a["kernel"] == (1, 1) and (a["padding"] == "valid" or (n.inputs[0].shape[0] % a["strides"][0] == 0
                                                                                  and n.inputs[0].shape[1] % a["strides"][1] == 0))):
            # ('same' pads nothing for a 1x1 kernel when the stride divides the map: Oper2D(1, (1,1), strides=(2,2)), :745)
            raise PlanError(f"{n.name}: only 1x1 strided convolutions without padding are lowered")

        def conv_desc(out_view, act_code, stats_ptr, weights=None, bias_ptr=None):
            wts = self.pwb(pe.key) if weights is None else weights
            bs = bias if bias_ptr is None else bias_ptr
            if n.op == "tconv":
                return lw.tconv_fprop(x.view, wts, cop, kh, kw, cin_p, out_view, bias=bs, act=act_code, stats=stats_ptr)
            xin = x.view.parity(0, 0, a["strides"][0], a["strides"][1]) if strided else x.view
            if a["padding"] == "same" or (kh, kw) == (1, 1):
                return lw.conv_fprop(xin, wts, cop, kh, kw, cin_p, out_view, bias=bs, act=act_code, stats=stats_ptr)
            raise PlanError(f"{n.name}: padding '{a['padding']}' with kernel {a['kernel']} is not lowered")

        if u["bn"] is None:
            dests = self._dests(out_node)
            self.emit(0, L.OP_CONV, conv_desc(dests[0], act, 0), n.name, flops=self._conv_flops(n))
            self._copy_extra(dests[0], dests[1:])
            u["y"] = dests[0]
            if not u.get("no_tap"):
                self.taps[out_node.name] = (dests[0], co, "act")
            if u["act"] is not None:
                self.taps[n.name] = (dests[0], co, "post")  # pre-activation is not materialised
            return
        bn = u["bn"]
        if (not self.training and not os.environ.get("B2SEG_NO_BN_FOLD") and act in (L.ACT_NONE, L.ACT_RELU, L.ACT_LEAKY, L.ACT_SIGMOID)
                and self._cvalid(co, self._segs(n), act) == 0):
            # Inference: BatchNorm (moving statistics) folded into the kernel and bias (b2seg_fold_bn, phase 2, replayed when the weights
            # change); the convolution epilogue writes act(BN(conv)) straight into its destination — no raw tensor, no BatchNorm pass
            row = pe.meta["taps"] * pe.meta["cin_p"]
            wf, bfold = self.alloc(cop * row * 2, "fold_w"), self.alloc(cop * 4, "scratch")
            self.emit(2, L.OP_FOLD_BN, L.FoldDesc(self.pw(pe.key), bias, self.pw(f"{bn.name}/gamma"), self.pw(f"{bn.name}/beta"),
                                                 self.pmov(f"{bn.name}/moving_mean"), self.pmov(f"{bn.name}/moving_variance"), bn.attrs["eps"],
                                                 cop, row, wf, bfold), f"fold {bn.name} into {n.name}")
            dests = self._dests(out_node)
            self.emit(0, L.OP_CONV, conv_desc(dests[0], act, 0, wf, bfold), n.name, flops=self._conv_flops(n))
            self._copy_extra(dests[0], dests[1:])
            if u["pool"] is not None:
                pn = u["pool"]
                pd = self._dests(pn)
                d = L.BnActDesc()
                d.x, d.act, d.n_out = dests[0].to_c(), L.ACT_NONE, 0
                d.pool_h, d.pool_w = pn.attrs["size"]
                d.pooled = pd[0].to_c()
                self.emit(0, L.OP_BN_ACT, d, pn.name)
                self._copy_extra(pd[0], pd[1:])
                self.taps[pn.name] = (pd[0], pn.C, "act")
            u["y"] = dests[0]
            self.taps[out_node.name] = (dests[0], co, "act")
            return
        z = self.new_act(H, W, cop)
        u["z"] = z
        cd = conv_desc(z, L.ACT_NONE, 0)
        n_part = int(self.stat_rows_fn(cd))
        stats = self.alloc(n_part * 2 * cop * 4, "scratch") if self.training else 0
        cd.stats = stats
        self.emit(0, L.OP_CONV, cd, n.name, flops=self._conv_flops(n))
        vec = self.alloc(4 * cop * 4, "scratch")
        u["scale"], u["shift"], u["mean"], u["rstd"] = vec, vec + cop * 4, vec + 2 * cop * 4, vec + 3 * cop * 4
        count = float(self.N * H * W)
        self.emit(0, L.OP_BN_FINALIZE, L.BnFinalizeDesc(
            stats, n_part, cop, count, self.pw(f"{bn.name}/gamma"), self.pw(f"{bn.name}/beta"),
            self.pmov(f"{bn.name}/moving_mean"), self.pmov(f"{bn.name}/moving_variance"),
            1 if self.training else 0, 1 if self.ndim == 2 else 0, bn.attrs["eps"], bn.attrs["momentum"],
            u["scale"], u["shift"], u["mean"], u["rstd"], 0 if self.training else 1), bn.name)
        if self._head_prologue_ok(u):
            # the head applies BatchNorm + activation itself and its backward is folded into this layer's: no activated tensor
            self.phys[id(out_node)] = Phys(z, co, list(self._segs(n)))
            self.head_prologue[id(out_node)] = (u["scale"], u["shift"], act)
            u["y"] = z
            self.taps[n.name] = (z, co, "raw")
            return
        dests = self._dests(out_node)
        d = L.BnActDesc()
        d.x, d.scale, d.shift, d.act = z.to_c(), u["scale"], u["shift"], act
        d.n_out = min(len(dests), 2)
        for i in range(d.n_out):
            d.out[i] = dests[i].to_c()
        d.c_valid = self._cvalid(co, self._segs(n), act)
        if self.training and u.get("stats_for") is not None and d.c_valid == 0:
            # this tensor is one channel window of a concat a BatchNormalization reads: its sums are that layer's batch statistics
            d.out_stats = self._stats_acc(u["stats_for"]) + 4 * u["stats_off"]
            d.out_stats_pitch = self._cphys(u["stats_for"]["node"].inputs[0])
        if u["pool"] is not None:
            pn = u["pool"]
            pd = self._dests(pn)
            d.pool_h, d.pool_w = pn.attrs["size"]
            d.pooled = pd[0].to_c()
            self.taps[pn.name] = (pd[0], pn.C, "act")
        self.emit(0, L.OP_BN_ACT, d, out_node.name)
        self._copy_extra(dests[0], dests[2:])
        if u["pool"] is not None:
            self._copy_extra(pd[0], pd[1:])
        u["y"] = dests[0]
        self.taps[n.name] = (z, co, "raw")
        self.taps[out_node.name] = (dests[0], co, "act")

    def _fwd_concat(self, u):
        n = u["node"]
        self._concat_buffer(n)  # producers already wrote their slots
        self.taps[n.name] = (self.phys[id(n)].view, n.C, "concat")
        # A concatenation with several consumers is not folded into an outer concatenation that also consumes it (the recurrent
        # blocks of RUNet / R2UNet concatenate the block input — itself a concat in the decoder — again and again,
        # 1DCNN/Models/unet_variants.py:63-72): its buffer is complete here, copy it into the outer slots.
        slots = []
        for c in self.cons[id(n)]:
            if c.op in ("concat", "convlstm"):
                root = self._concat_root(c)
                buf = self._concat_buffer(root)
                off = 0
                for p in self.flat_concat[id(root)]:
                    if p is n:
                        slots.append(buf.chan(off, self._cphys(n)))
                    off += self._cphys(p)
        self._copy_extra(self.phys[id(n)].view, slots)

    def _cvalid(self, C, segs=None, act=0):
        """channel count to pass as c_valid so padding lanes are forced to zero (0 = nothing to mask).  In a gapped layout
        (odd-channel concat slots) padding lanes are interleaved; they stay zero by construction for every activation
        with f(0) = 0, so only sigmoid needs the mask and is refused there."""
        if segs is not None and list(segs) != [(0, C)]:
            if act == L.ACT_SIGMOID:
                raise PlanError("sigmoid over a gapped (odd-channel concat) channel layout is not lowered")
            return 0
        # Padding lanes need no mask for an activation with f(0) = 0: their pre-activation is exactly 0 (zero weight rows, zero bias;
        # BatchNorm of an all-zero channel has mean 0 and beta 0), and an unmasked descriptor takes the row-walking fast kernels
        # (MultiResUNet's 10 / 21-channel branches used to fall back to the generic ones).  Only sigmoid(0) = 0.5 must be forced to 0.
        return C if (C % 8 and act == L.ACT_SIGMOID) else 0

    def _fwd_add(self, u):
        n = u["node"]
        if len(n.inputs) not in (2, 3):
            raise PlanError("add with != 2 or 3 inputs")
        if u.get("fused_bn") is not None:
            return self._fwd_add_fused(u)
        ins = [self.phys[id(i)] for i in n.inputs]
        for p in ins[1:]:
            if p.segs != ins[0].segs:
                raise PlanError(f"{n.name}: operands of add have different channel layouts (odd-channel concat; not lowered yet)")
        out_node = u["out"]
        dests = self._dests(out_node)
        self.phys[id(out_node)] = Phys(dests[0], out_node.C, list(ins[0].segs))
        c = ins[2].view.to_c() if len(ins) == 3 else lw.NULL_VIEW.to_c()
        if self.training and u.get("stats_for") is not None and len(ins) == 2:
            # Add + ReLU whose result a BatchNormalization reads (ResPath): the apply kernel with identity scale adds the operands and
            # accumulates that layer's batch statistics on the way
            d = L.BnActDesc()
            d.x, d.add, d.act = ins[0].view.to_c(), ins[1].view.to_c(), self._act_code(u["act"])
            d.n_out = min(len(dests), 2)
            for i in range(d.n_out):
                d.out[i] = dests[i].to_c()
            d.out_stats = self._stats_acc(u["stats_for"])
            self.emit(0, L.OP_BN_ACT, d, out_node.name)
            self._copy_extra(dests[0], dests[2:])
        else:
            self.emit(0, L.OP_ELTWISE, L.EltwiseDesc(0 if len(ins) == 2 else 3, ins[0].view.to_c(), ins[1].view.to_c(), c, dests[0].to_c(),
                                                     self._act_code(u["act"])), out_node.name)
            self._copy_extra(dests[0], dests[1:])
        u["y"] = dests[0]
        self.taps[out_node.name] = (dests[0], n.C, "act")

    def _fwd_add_fused(self, u):
        """ReLU(shortcut + BatchNormalization(x)) in one pass over x and the shortcut (+ the column sums of the result when a
        BatchNormalization reads it): MultiResBlock unet_variants.py:96-98, ResPath :110-111"""
        n, ub = u["node"], u["fused_bn"]
        other = n.inputs[0] if n.inputs[1] is ub["out"] else n.inputs[1]
        xb, sc = ub["xphys"], self.phys[id(other)]
        if list(xb.segs) != list(sc.segs) or xb.Cp != sc.Cp:
            raise PlanError(f"{n.name}: operands of add have different channel layouts")
        out_node = u["out"]
        dests = self._dests(out_node)
        self.phys[id(out_node)] = Phys(dests[0], out_node.C, list(sc.segs))
        d = L.BnActDesc()
        d.x, d.scale, d.shift, d.act = xb.view.to_c(), ub["scale"], ub["shift"], self._act_code(u["act"])
        d.add = sc.view.to_c()
        d.n_out = min(len(dests), 2)
        for i in range(d.n_out):
            d.out[i] = dests[i].to_c()
        if self.training and u.get("stats_for") is not None:
            d.out_stats = self._stats_acc(u["stats_for"])
        self.emit(0, L.OP_BN_ACT, d, out_node.name)
        self._copy_extra(dests[0], dests[2:])
        u["y"] = dests[0]
        self.taps[out_node.name] = (dests[0], n.C, "act")

    def _fwd_up(self, u):
        n = u["node"]
        x = self.phys[id(n.inputs[0])]
        out_node = u["out"]
        dests = self._dests(out_node)
        fh, fw = n.attrs["size"]
        mode = 1 if n.attrs["interpolation"] == "bilinear" else 0
        act = self._act_code(u["act"])
        d = L.ResizeDesc(x.view.to_c(), dests[0].to_c(), lw.NULL_VIEW.to_c(), fh, fw, mode, act, 0)
        if act == L.ACT_SIGMOID:
            if list(x.segs) != [(0, n.C)]:
                # gapped layout (MultiResBlock outputs in MultiResUNet3P / KSSNet): sigmoid(0) = 0.5 must not leak into the padding lanes
                if len(x.segs) > 8:
                    raise PlanError(f"{out_node.name}: sigmoid over a channel layout with {len(x.segs)} segments (max 8)")
                d.n_vseg = len(x.segs)
                for i, (po, c) in enumerate(x.segs):
                    d.vseg_off[i], d.vseg_cnt[i] = po, c
            else:
                d.c_valid = self._cvalid(n.C, x.segs, act)
        self.emit(0, L.OP_RESIZE_FWD, d, out_node.name)
        self._copy_extra(dests[0], dests[1:])
        u["y"] = dests[0]
        self.taps[out_node.name] = (dests[0], n.C, "act")

    def _fwd_pool(self, u):
        n = u["node"]
        x = self.phys[id(n.inputs[0])]
        dests = self._dests(n)
        d = L.BnActDesc()
        d.x, d.act, d.n_out = x.view.to_c(), L.ACT_NONE, 0
        d.pool_h, d.pool_w = n.attrs["size"]
        d.pooled = dests[0].to_c()
        self.emit(0, L.OP_BN_ACT, d, n.name)
        self._copy_extra(dests[0], dests[1:])
        self.taps[n.name] = (dests[0], n.C, "act")

    def _fwd_act(self, u):
        n = u["node"]
        x = self.phys[id(n.inputs[0])]
        dests = self._dests(n)
        self.phys[id(n)] = Phys(dests[0], n.C, list(x.segs))
        d = L.BnActDesc()
        d.x, d.act, d.n_out = x.view.to_c(), self._act_code(n), min(len(dests), 2)
        for i in range(d.n_out):
            d.out[i] = dests[i].to_c()
        d.c_valid = self._cvalid(n.C, x.segs, d.act) if d.act == L.ACT_SIGMOID else 0
        self.emit(0, L.OP_BN_ACT, d, n.name)
        self._copy_extra(dests[0], dests[2:])
        u["x"] = x.view
        self.taps[n.name] = (dests[0], n.C, "act")

    def _fwd_bn(self, u):
        """BatchNormalization whose input is not a convolution output (MultiResBlock / ResPath): statistics kernel first"""
        n = u["node"]
        x = self.phys[id(n.inputs[0])]
        H, W, _ = n.shape
        cp = x.Cp
        out_node = u["out"]
        act = self._act_code(u["act"])
        nb = max(1, min(592, (self.N * H * W) // 64))
        if self.training and u.get("stats_from"):
            stats, nb = self._stats_acc(u), 1          # the kernels that wrote x added its column sums into this accumulator
        else:
            stats = self.alloc(nb * 2 * cp * 4, "scratch") if self.training else 0
            if self.training:
                self.emit(0, L.OP_COLSTATS, L.ColstatsDesc(x.view.to_c(), stats, nb), f"stats {n.name}")
        vec = self.alloc(4 * cp * 4, "scratch")
        u["scale"], u["shift"], u["mean"], u["rstd"] = vec, vec + cp * 4, vec + 2 * cp * 4, vec + 3 * cp * 4
        self.emit(0, L.OP_BN_FINALIZE, L.BnFinalizeDesc(
            stats, nb, cp, float(self.N * H * W), self.pw(f"{n.name}/gamma"), self.pw(f"{n.name}/beta"),
            self.pmov(f"{n.name}/moving_mean"), self.pmov(f"{n.name}/moving_variance"),
            1 if self.training else 0, 1 if self.ndim == 2 else 0, n.attrs["eps"], n.attrs["momentum"],
            u["scale"], u["shift"], u["mean"], u["rstd"], 0 if self.training else 1), n.name)
        u["x"] = x.view
        if u.get("defer_to") is not None:
            # applied by the Add + ReLU that reads it (bn_act.add): the normalised tensor is not materialised
            u["xphys"] = x
            self.phys[id(out_node)] = Phys(x.view, n.C, list(x.segs))    # (geometry only: nothing reads the normalised tensor as data)
            return
        dests = self._dests(out_node)
        d = L.BnActDesc()
        d.x, d.scale, d.shift, d.act = x.view.to_c(), u["scale"], u["shift"], act
        d.n_out = min(len(dests), 2)
        for i in range(d.n_out):
            d.out[i] = dests[i].to_c()
        d.c_valid = self._cvalid(n.C, x.segs, act)
        self.emit(0, L.OP_BN_ACT, d, out_node.name)
        self._copy_extra(dests[0], dests[2:])
        u["y"] = dests[0]
        self.taps[out_node.name] = (dests[0], n.C, "act")

    def _fwd_mul(self, u):
        n = u["node"]
        a, b = self.phys[id(n.inputs[0])], self.phys[id(n.inputs[1])]
        if n.inputs[1].C != 1:
            raise PlanError(f"{n.name}: only (N,H,W,C) * (N,H,W,1) broadcast multiplies are lowered")
        dests = self._dests(n)
        self.phys[id(n)] = Phys(dests[0], n.C, list(a.segs))
        nv = lw.NULL_VIEW.to_c()
        self.emit(0, L.OP_MULBC_FWD, L.MulbcDesc(a.view.to_c(), b.view.to_c(), dests[0].to_c(), nv, nv, nv), n.name)
        self._copy_extra(dests[0], dests[1:])
        self.taps[n.name] = (dests[0], n.C, "act")

    def _fwd_convlstm(self, u):
        n = u["node"]
        a = n.attrs
        kh, kw = a["kernel"]
        F = a["filters"]
        self._concat_buffer(n)
        x = self.catbuf[id(n)]
        pe = self.pindex[f"{n.name}/kernel"]
        pe.meta["segs"] = list(x.segs)
        assert pe.meta["cin_p"] == x.Cp
        H, W, _ = n.shape
        z = self.new_act(H, W, 3 * F)
        u["z"] = z
        flops = 2.0 * self.N * H * W * x.C * 3 * F * kh * kw   # live gates only (3F, no recurrent conv)
        self.emit(0, L.OP_CONV, lw.conv_fprop(x.view, self.pwb(pe.key), 3 * F, kh, kw, x.Cp, z, bias=self.pw(f"{n.name}/bias")),
                  n.name, flops=flops)
        dests = self._dests(n)
        nv = lw.NULL_VIEW.to_c()
        self.emit(0, L.OP_LSTM_FWD, L.LstmDesc(z.to_c(), dests[0].to_c(), nv, nv, F), f"gates {n.name}")
        self._copy_extra(dests[0], dests[1:])
        self.taps[n.name] = (dests[0], F, "act")
        self.taps[f"{n.name}/gates"] = (z, 3 * F, "raw")      # live gate pre-activations [i | g | o]

    # -- Feature_Extraction_Block: Flatten -> Dense -> Dense -> Reshape -------------------------------------------------------
    @staticmethod
    def _flat_segs(H, W, Cp, segs):
        """logical -> physical segments of Flatten over an (H, W, Cp) buffer whose pixels carry `segs`: Keras' Flatten order is
        (h, w, logical channel), the buffer's is (h, w, physical channel) — identical when the tensor is unpadded"""
        if list(segs) == [(0, Cp)]:
            return [(0, H * W * Cp)]
        return [(pix * Cp + o, c) for pix in range(H * W) for (o, c) in segs]

    def _flat_view(self, t: Node, what: str) -> TView:
        """the channels-last buffer of t seen as (N, 1, 1, H*W*Cp).  Padding lanes and gaps (odd MultiRes channel counts) stay in
        the flattened vector: they hold zeros and the Dense layer that reads it keeps zero weights there (segments in its
        parameter layout), exactly like a convolution reading a gapped concat buffer."""
        p = self.phys[id(t)]
        H, W, C = t.shape
        v = p.view
        if v.C != p.Cp or v.sw != p.Cp or (H > 1 and v.sh != W * p.Cp) or v.sn != H * W * p.Cp:
            raise PlanError(f"{what}: needs a contiguous channels-last tensor (got {t.name}: C={C}, physical {p.Cp}, strides {v.sn},{v.sh},{v.sw})")
        return TView.dense(v.ptr, self.N, 1, 1, H * W * p.Cp)

    def _fwd_flatten(self, u):
        n = u["node"]
        t = n.inputs[0]
        H, W, _ = t.shape
        p = self.phys[id(t)]
        self.phys[id(n)] = Phys(self._flat_view(t, n.name), n.C, self._flat_segs(H, W, p.Cp, p.segs))

    def _fwd_reshape(self, u):
        n = u["node"]
        H, W, C = n.shape
        src = self.phys[id(n.inputs[0])]
        if C % 8 or src.Cp != H * W * C:
            raise PlanError(f"{n.name}: Reshape to {n.shape} needs C % 8 == 0 and an unpadded source")
        view = TView.dense(src.view.ptr, self.N, H, W, C)
        if any(c.op in ("concat", "convlstm") for c in self.cons[id(n)]):
            # a view cannot live inside a concat buffer: copy it into every slot; direct consumers keep reading the view
            for d in self._dests(n):
                self.emit(0, L.OP_ELTWISE, L.EltwiseDesc(1, view.to_c(), lw.NULL_VIEW.to_c(), lw.NULL_VIEW.to_c(), d.to_c()), f"concat copy {n.name}")
        self.phys[id(n)] = Phys(view, C, [(0, C)])
        self.taps[n.name] = (view, C, "act")

    def _fwd_dense(self, u):
        n = u["node"]
        x = self.phys[id(n.inputs[0])]
        fin, units = n.inputs[0].shape[2], n.attrs["units"]
        pe = self.pindex[f"{n.name}/kernel"]
        if (pe.meta["cin_p"], list(pe.meta["segs"])) != (x.Cp, list(x.segs)):
            raise PlanError(f"{n.name}: input layout {x.Cp} / {len(x.segs)} segments differs from the parameter layout {pe.meta['cin_p']}")
        out = self.new_act(1, 1, units)
        self.emit(0, L.OP_CONV, lw.conv_fprop(x.view, self.pwb(f"{n.name}/kernel"), units, 1, 1, x.Cp, out, bias=self.pw(f"{n.name}/bias")),
                  n.name, flops=2.0 * self.N * fin * units)
        self.phys[id(n)] = Phys(out, units, [(0, units)])
        u["x"] = x.view
        if n.name not in self.taps:
            self.taps[n.name] = (out, units, "act")

    def _bwd_flatten(self, u):
        n = u["node"]
        g = self._single_grad(n)
        if g is not None:
            H, W, _ = n.inputs[0].shape
            self._add_gsrc(n.inputs[0], GSrc(TView.dense(g.ptr, self.N, H, W, self.phys[id(n.inputs[0])].Cp)))

    def _bwd_reshape(self, u):
        n = u["node"]
        g = self._single_grad(n)
        if g is not None:
            H, W, C = n.shape
            if (g.C, g.sw, g.sh, g.sn) != (C, C, W * C, H * W * C):
                # the only gradient is a channel window of a concat gradient buffer: make it dense before re-viewing it as (N,1,1,F)
                dense = self._grad_like(n)
                self.emit(1, L.OP_ELTWISE, L.EltwiseDesc(1, g.to_c(), lw.NULL_VIEW.to_c(), lw.NULL_VIEW.to_c(), dense.to_c()), f"dense grad {n.name}")
                g = dense
            self._add_gsrc(n.inputs[0], GSrc(TView.dense(g.ptr, self.N, 1, 1, H * W * C)))

    def _bwd_dense(self, u):
        n = u["node"]
        dz = self._single_grad(n)
        if dz is None:
            return
        fin, units = n.inputs[0].shape[2], n.attrs["units"]
        fin_p = u["x"].C                      # physical width of the input (a flattened padded tensor keeps its zero lanes)
        self.grad_taps[n.name] = (dz, units)
        flops = 2.0 * self.N * fin * units
        self.emit(1, L.OP_WGRAD, lw.conv_wgrad(dz, u["x"], self.pg(f"{n.name}/kernel"), units, 1, 1, fin_p), f"wgrad {n.name}", flops=flops)
        self.emit(1, L.OP_COLSUM, L.ColsumDesc(dz.to_c(), self.pg(f"{n.name}/bias"), 0, 0), f"bias grad {n.name}")
        dx = self.new_act(1, 1, fin_p, "grad")
        self.emit(1, L.OP_CONV, lw.conv_dgrad(dz, self.pwb(f"{n.name}/kernel"), units, 1, 1, fin_p, dx), f"dgrad {n.name}", flops=flops)
        self._add_gsrc(n.inputs[0], GSrc(dx))

    def _fwd_pow(self, u):
        """tf.math.pow(x, p) of an operational layer (onn_layers.py:19): one element-wise pass (padding lanes stay 0)"""
        n = u["node"]
        x = self.phys[id(n.inputs[0])]
        dests = self._dests(n)
        self.phys[id(n)] = Phys(dests[0], n.C, list(x.segs))
        nv = lw.NULL_VIEW.to_c()
        self.emit(0, L.OP_ELTWISE, L.EltwiseDesc(4, x.view.to_c(), nv, nv, dests[0].to_c(), n.attrs["p"]), n.name)
        self._copy_extra(dests[0], dests[1:])
        self.taps[n.name] = (dests[0], n.C, "act")

    def _bwd_pow(self, u):
        n = u["node"]
        if n.inputs[0].op == "input":
            return
        dy = self._single_grad(n)
        if dy is None:
            return
        x = self.phys[id(n.inputs[0])]
        dx = self._grad_like(n.inputs[0])
        nv = lw.NULL_VIEW.to_c()
        self.emit(1, L.OP_ELTWISE, L.EltwiseDesc(5, x.view.to_c(), dy.to_c(), nv, dx.to_c(), n.attrs["p"]), f"pow bwd {n.name}")
        self._add_gsrc(n.inputs[0], GSrc(dx))

    def _fwd_outact(self, u):
        n, src = u["node"], u["src"]
        x = self.phys[id(src)]
        if x.Cp != ceil8(src.C) or list(x.segs) != [(0, src.C)]:
            raise PlanError(f"output {n.name}: expected one dense channel block, got {x.Cp} channels / {x.segs}")
        H, W, co = n.shape
        npix = self.N * H * W
        y = self.alloc(npix * co * 4, "output")
        dl = self.alloc(npix * co * 4, "grad") if self.training else 0
        tgt = self.alloc(npix * co * 4, "target") if self.training else 0
        act = L.ACT_CODES[u["fn"]]
        u["desc"] = L.OutActDesc(x.view.to_c(), co, act, y, dl, lw.NULL_VIEW.to_c())
        self.emit(0, L.OP_OUTACT_FWD, u["desc"], f"output {n.name}")
        self.outputs.append(dict(index=self.g.outputs.index(n), name=n.name, ptr=y, shape=(self.N, H, W, co), target_ptr=tgt, act=act,
                                 dlogits=dl, npix=npix, cout=co))

    def _loss_kind(self, o) -> int:
        """loss code of output o (b2seg_loss_desc.kind).  Every loss of 2DCNN/utils/tf_losses.py except the sparse one is lowered, on
        linear / sigmoid / softmax heads alike (cross-entropies on their own activation from the cached logits like Keras 2, else
        from clipped probabilities).  Refused: a cross-entropy on the OTHER activation — binary (focal) cross-entropy on a softmax
        head, categorical on a sigmoid head: Keras 2 finds the cached `_keras_logits` of whichever activation ran and feeds them to
        the wrong *_cross_entropy_with_logits, a quirk nobody means — and SparseCategoricalCrossentropy (integer targets)."""
        name = self.losses[o["index"]] if self.losses else "bce"
        kind, act = LOSS_KINDS[name], o["act"]
        if act not in (L.ACT_NONE, L.ACT_SIGMOID, L.ACT_SOFTMAX) or (kind in (0, 7) and act == L.ACT_SOFTMAX) or (kind == 1 and act == L.ACT_SIGMOID):
            raise PlanError(f"output {o['name']}: loss '{name}' on a head with activation code {act} is not lowered (a cross-entropy needs "
                            f"its own activation or a linear head)")
        return kind

    def _emit_loss(self, o, note):
        idx = o["index"]
        kind = self._loss_kind(o)
        wgt = float(self.loss_weights[idx]) if self.loss_weights else 1.0
        self.emit(1, L.OP_LOSS, L.LossDesc(o["ptr"], o["target_ptr"], o["npix"], o["cout"], kind, o["act"], wgt, o["dlogits"], self.loss_ptr,
                                           self.loss_ptr + 4 * (8 + 8 * idx)), note)

    def _bwd_outact(self, u):
        n, src = u["node"], u["src"]
        o = next(o for o in self.outputs if o["name"] == n.name)
        self._emit_loss(o, f"loss {n.name}")
        dx = self._grad_like(src)
        d = L.OutActDesc.from_buffer_copy(u["desc"])
        d.dx = dx.to_c()
        self.emit(1, L.OP_OUTACT_BWD, d, f"output bwd {n.name}")
        self._add_gsrc(src, GSrc(dx))

    def _fwd_head(self, u):
        n = u["node"]
        a = n.attrs
        x = self.phys[id(n.inputs[0])]
        pe = self.pindex[f"{n.name}/kernel"]
        pe.meta["segs"] = list(x.segs)
        co = a["filters"]
        st = a["strides"][1]
        if a["strides"][0] not in (1, st):
            raise PlanError("anisotropic head stride")
        Ho, Wo, _ = n.shape
        npix = self.N * Ho * Wo
        y = self.alloc(npix * co * 4, "output")
        dl = self.alloc(npix * co * 4, "grad") if self.training else 0
        tgt = self.alloc(npix * co * 4, "target") if self.training else 0
        act = L.ACT_CODES[a.get("activation")]
        d = L.HeadDesc()
        d.x, d.w, d.b, d.cout, d.act, d.stride = x.view.to_c(), self.pw(pe.key), self.pw(f"{n.name}/bias"), co, act, st
        d.y, d.dlogits = y, dl
        d.dw, d.db = self.pg(pe.key), self.pg(f"{n.name}/bias")
        if id(n.inputs[0]) in self.head_prologue:
            d.bn_scale, d.bn_shift, d.bn_act = self.head_prologue[id(n.inputs[0])]
        u["desc"] = d
        self.emit(0, L.OP_HEAD_FWD, d, n.name)
        idx = self.g.outputs.index(n)
        self.outputs.append(dict(index=idx, name=n.name, ptr=y, shape=(self.N, Ho, Wo, co), target_ptr=tgt, act=act, dlogits=dl, npix=npix, cout=co))

    # -- backward emitters -----------------------------------------------------------------------------------
    def _add_gsrc(self, t: Node, s: GSrc):
        self.gsrc.setdefault(id(t), []).append(s)

    def _bwd_head(self, u):
        n = u["node"]
        o = next(o for o in self.outputs if o["name"] == n.name)
        self._emit_loss(o, f"loss {n.name}")
        t = n.inputs[0]
        x = self.phys[id(t)]
        if self._head_foldable(n):
            # the head reads act(BN(conv)): its backward (input gradient, dW, db) is folded into that layer's BN backward,
            # so the 2 x H x W x C gradient tensor is never written or read (b2seg_gradsrc kind 2)
            self._add_gsrc(t, GSrc(x.view, 2, (1, 1), head=dict(unit=u, name=n.name, dlogits=o["dlogits"], cout=o["cout"])))
            return
        self._emit_head_bwd(u)

    def _head_foldable(self, n: Node) -> bool:
        """head n reads act(BN(conv)) of a Conv_Block whose backward can take the head's backward in (b2seg_gradsrc kind 2)"""
        t = n.inputs[0]
        pu = self.unit_of_out.get(id(t))
        return bool(self.fuse_heads and pu is not None and pu["kind"] == "conv" and pu.get("gate") is None and pu["bn"] is not None and pu["out"] is t
                    and pu["pool"] is None and pu["act"] is not None and self._act_code(pu["act"]) in (L.ACT_RELU, L.ACT_LEAKY)
                    and n.attrs["filters"] <= 2 and n.attrs["strides"] == (1, 1) and self.N * t.shape[0] * t.shape[1] * self._cphys(t) < 2 ** 31)

    def _head_prologue_ok(self, u) -> bool:
        """the LAST Conv_Block: its activated tensor has one reader, a foldable head, whose forward can apply BatchNorm + activation
        itself (b2seg_head_desc.bn_scale) — then neither direction of the head needs the tensor and it is not materialised"""
        t = u["out"]
        c = self.cons[id(t)]
        if os.environ.get("B2SEG_NO_HEAD_PROLOGUE") or t in self.g.outputs or len(c) != 1 or self.unit_of_out.get(id(c[0]), {}).get("kind") != "head":
            return False
        cv = self._cphys(t) // 8
        dense = list(self._segs(t)) == [(0, t.C)] and self._cphys(t) == t.C
        return self.training and self._head_foldable(c[0]) and dense and cv <= 32 and cv & (cv - 1) == 0 and self.losses is not None

    def _emit_head_bwd(self, u) -> Optional[GSrc]:
        """stand-alone head backward: dW, db and (unless the head reads the network input) dx as a dense gradient tensor"""
        n = u["node"]
        d: L.HeadDesc = u["desc"]
        x = self.phys[id(n.inputs[0])]
        H, W, _ = n.inputs[0].shape
        d2 = L.HeadDesc.from_buffer_copy(d)
        d2.dw, d2.db = self.pg(f"{n.name}/kernel"), self.pg(f"{n.name}/bias")   # (also marks them written by this unit)
        src = None
        if n.inputs[0].op != "input":
            dx = self.new_act(H, W, x.Cp, "grad")
            d2.dx = dx.to_c()
            src = GSrc(dx)
            self._add_gsrc(n.inputs[0], src)
        self.emit(1, L.OP_HEAD_BWD, d2, f"head bwd {n.name}")
        return src

    def _bwd_input(self, u):
        pass

    def _grad_like(self, t: Node) -> TView:
        """a fresh gradient buffer with the physical layout of tensor t"""
        p = self.phys[id(t)]
        H, W, _ = t.shape
        return self.new_act(H, W, p.Cp, "grad")

    def _direct_sources(self, t: Node, srcs: List[GSrc]) -> List[GSrc]:
        """turn max-pool-routed sources into dense ones (generic pool backward against the stored forward tensor)"""
        out = []
        for s in srcs:
            if s.kind == 0:
                out.append(s)
            elif s.kind == 2:      # a head that could not be folded into a BN backward after all: run its own backward now
                out.append(self._unfuse_head(t, s))
            else:
                dx = self._grad_like(t)
                self.emit(1, L.OP_POOL_BWD, L.PoolBwdDesc(self.phys[id(t)].view.to_c(), s.view.to_c(), dx.to_c(), s.pool[0], s.pool[1]),
                          f"pool bwd -> {t.name}")
                out.append(GSrc(dx))
        return out

    def _unfuse_head(self, t: Node, s: GSrc) -> GSrc:
        lst = self.gsrc.get(id(t), [])
        n_before = len(lst)
        src = self._emit_head_bwd(s.head["unit"])
        del lst[n_before:]          # _emit_head_bwd registered the dense gradient on t; the caller consumes it directly
        return src

    def _kernel_sources(self, t: Node, srcs: List[GSrc], allow_head: bool = False) -> List[GSrc]:
        """sources in the form bn_bwd accepts: pooled sources only if they share one window of 2 or 4 elements; a folded head
        (kind 2) only next to direct sources of a BN + ReLU/LeakyReLU layer, and only one of them"""
        heads = [s for s in srcs if s.kind == 2]
        if heads and (not allow_head or len(heads) > 1 or any(s.kind == 1 for s in srcs)):
            srcs = [self._unfuse_head(t, s) if s.kind == 2 else s for s in srcs]
        wins = {s.pool for s in srcs if s.kind == 1}
        if len(wins) > 1 or any(w[0] * w[1] not in (2, 4) for w in wins):
            srcs = self._direct_sources(t, srcs)
        for s in srcs:
            if s.kind == 1 and s.pool_name:
                # the BatchNorm / activation / pooling backward kernel recomputes the activations in fp32 and routes the pooled
                # gradient to THEIR first maximum (not to the first maximum of the bf16-rounded stored tensor, which is what
                # b2seg_pool_bwd does); recorded for the parity tests
                self.pool_routing[s.pool_name] = "recomputed"
        return self._reduce_sources(srcs, self.phys[id(t)].view)

    def _single_grad(self, t: Node) -> Optional[TView]:
        """one dense tensor holding the total gradient w.r.t. t (None if no gradient reaches t)"""
        srcs = self.gsrc.get(id(t), [])
        if not srcs:
            return None
        srcs = self._direct_sources(t, srcs)
        if len(srcs) == 1:
            return srcs[0].view
        nv = lw.NULL_VIEW.to_c()
        acc = srcs[0].view
        i = 1
        while i < len(srcs):
            tmp = self._grad_like(t)
            if i + 1 < len(srcs):
                self.emit(1, L.OP_ELTWISE, L.EltwiseDesc(3, acc.to_c(), srcs[i].view.to_c(), srcs[i + 1].view.to_c(), tmp.to_c(), 0), "grad sum")
                i += 2
            else:
                self.emit(1, L.OP_ELTWISE, L.EltwiseDesc(0, acc.to_c(), srcs[i].view.to_c(), nv, tmp.to_c(), 0), "grad sum")
                i += 1
            acc = tmp
        return acc

    def _act_bwd_noBN(self, t_out: Node, x_view: TView, act: int, note: str) -> Optional[TView]:
        """dz = (sum of gradient sources of t_out) * act'(.) through the bn_bwd kernel without BatchNorm"""
        srcs = self.gsrc.get(id(t_out), [])
        if not srcs:
            return None
        srcs = self._kernel_sources(t_out, srcs)
        dz = self._grad_like(t_out)
        d = L.BnBwdDesc()
        d.x, d.act, d.n_src = x_view.to_c(), act, len(srcs)
        for i, s in enumerate(srcs):
            d.src[i] = L.GradSrc(s.view.to_c(), s.kind, s.pool[0], s.pool[1])
        d.count = 1.0
        d.dx = dz.to_c()
        self.emit(1, L.OP_BN_BWD, d, note)
        return dz

    def _bwd_add(self, u):
        n, out_node = u["node"], u["out"]
        if u["act"] is None:
            # a max-pool-routed gradient must be routed against the SUM (the tensor that was pooled), not against each addend:
            # make it dense here (R2UNet pools `Add([shortcut, recurrent pair])`, 1DCNN/Models/unet_variants.py:1060-1064)
            for s in self._direct_sources(n, self.gsrc.get(id(n), [])):
                for i in n.inputs:
                    self._add_gsrc(i, s)
            return
        act = self._act_code(u["act"])
        if u.get("bwd_premasked"):
            # the BatchNormalization that reads this ReLU applied its mask already (bn_bwd.x_relu_mask): its dx IS the gradient of the sum
            for s in self._direct_sources(out_node, self.gsrc.get(id(out_node), [])):
                for i in n.inputs:
                    self._add_gsrc(i, s)
            return
        dz = self._act_bwd_noBN(out_node, u["y"], act, f"act bwd {out_node.name}")  # relu / leaky: sign(y) == sign(pre-activation)
        if dz is not None:
            for i in n.inputs:
                self._add_gsrc(i, GSrc(dz))

    def _bwd_up(self, u):
        n, out_node = u["node"], u["out"]
        dy = self._single_grad(out_node)
        if dy is None or n.inputs[0].op == "input":
            return
        dx = self._grad_like(n.inputs[0])
        fh, fw = n.attrs["size"]
        mode = 1 if n.attrs["interpolation"] == "bilinear" else 0
        act = self._act_code(u["act"])
        yf = u["y"].to_c() if act != L.ACT_NONE else lw.NULL_VIEW.to_c()
        self.emit(1, L.OP_RESIZE_BWD, L.ResizeDesc(dx.to_c(), dy.to_c(), yf, fh, fw, mode, act, 0), f"up bwd {n.name}")
        self._add_gsrc(n.inputs[0], GSrc(dx))

    def _pool_sources(self, pool_node: Node) -> List[GSrc]:
        """dense gradient sources of a max-pool output; many of them (an operational layer consumes its input q times, a
        recurrent block t + 1 times) are summed first so that the pooled tensor's backward sees one routed source, not a dozen"""
        srcs = self._direct_sources(pool_node, self.gsrc.get(id(pool_node), []))
        if len(srcs) > L.MAX_GRADSRC - 2:
            self.gsrc[id(pool_node)] = srcs
            srcs = [GSrc(self._single_grad(pool_node))]
        return srcs

    def _bwd_pool(self, u):
        n = u["node"]
        srcs = self._pool_sources(n)
        for s in srcs:
            self._add_gsrc(n.inputs[0], GSrc(s.view, 1, tuple(n.attrs["size"]), pool_name=n.name))

    def _bwd_act(self, u):
        n = u["node"]
        dz = self._act_bwd_noBN(n, u["x"], self._act_code(n), f"act bwd {n.name}")  # x = pre-activation: exact for every activation
        if dz is not None and n.inputs[0].op != "input":
            self._add_gsrc(n.inputs[0], GSrc(dz))

    def _bwd_bn(self, u):
        n, out_node = u["node"], u["out"]
        srcs = self.gsrc.get(id(out_node), [])
        if not srcs:
            return
        srcs = self._kernel_sources(out_node, srcs)
        H, W, _ = n.shape
        cp = u["x"].C
        dx = self._grad_like(n.inputs[0])
        d = L.BnBwdDesc()
        d.x, d.scale, d.shift, d.mean, d.rstd = u["x"].to_c(), u["scale"], u["shift"], u["mean"], u["rstd"]
        d.act, d.n_src = self._act_code(u["act"]), len(srcs)
        for i, s in enumerate(srcs):
            d.src[i] = L.GradSrc(s.view.to_c(), s.kind, s.pool[0], s.pool[1])
        d.count = float(self.N * H * W)
        nb = max(1, min(1184, (self.N * H * W) // 64))
        d.partials, d.n_blocks = self.alloc(nb * 2 * cp * 4, "scratch"), nb
        d.dgamma, d.dbeta = self.pg(f"{n.name}/gamma"), self.pg(f"{n.name}/beta")
        d.dx = dx.to_c()
        d.accumulate = 1   # "zero grads" opens the backward phase
        d.x_relu_mask = 1 if u.get("x_relu_mask") else 0
        self.emit(1, L.OP_BN_BWD, d, f"bn bwd {n.name}")
        self._add_gsrc(n.inputs[0], GSrc(dx))

    def _bwd_mul(self, u):
        n = u["node"]
        dout = self._single_grad(n)
        if dout is None:
            return
        a, b = self.phys[id(n.inputs[0])], self.phys[id(n.inputs[1])]
        da, db = self._grad_like(n.inputs[0]), self._grad_like(n.inputs[1])
        self.emit(1, L.OP_MULBC_BWD, L.MulbcDesc(a.view.to_c(), b.view.to_c(), lw.NULL_VIEW.to_c(), dout.to_c(), da.to_c(), db.to_c()),
                  f"mul bwd {n.name}")
        self._add_gsrc(n.inputs[0], GSrc(da))
        self._add_gsrc(n.inputs[1], GSrc(db))

    def _bwd_convlstm(self, u):
        n = u["node"]
        a = n.attrs
        kh, kw = a["kernel"]
        F = a["filters"]
        dh = self._single_grad(n)
        if dh is None:
            return
        x = self.catbuf[id(n)]
        pe = self.pindex[f"{n.name}/kernel"]
        H, W, _ = n.shape
        dz = self.new_act(H, W, 3 * F, "grad")
        nv = lw.NULL_VIEW.to_c()
        self.emit(1, L.OP_LSTM_BWD, L.LstmDesc(u["z"].to_c(), nv, dh.to_c(), dz.to_c(), F), f"gates bwd {n.name}")
        self.grad_taps[f"{n.name}/gates"] = (dz, 3 * F)
        if (dh.C, dh.sw) == (F, F):
            self.grad_taps[n.name] = (dh, F)          # gradient w.r.t. the ConvLSTM output h (when it is one dense tensor)
        flops = 2.0 * self.N * H * W * x.C * 3 * F * kh * kw
        self.emit(1, L.OP_WGRAD, lw.conv_wgrad(dz, x.view, self.pg(pe.key), 3 * F, kh, kw, x.Cp), f"wgrad {n.name}", flops=flops)
        self.emit(1, L.OP_COLSUM, L.ColsumDesc(dz.to_c(), self.pg(f"{n.name}/bias"), 0, 0), f"bias grad {n.name}")
        dx = self.new_act(H, W, x.Cp, "grad")
        self.emit(1, L.OP_CONV, lw.conv_dgrad(dz, self.pwb(pe.key), 3 * F, kh, kw, x.Cp, dx), f"dgrad {n.name}", flops=flops)
        off = 0
        for p in u["parts"]:
            cp = self._cphys(p)
            self._add_gsrc(p, GSrc(dx.chan(off, cp)))
            off += cp

    def _bwd_concat(self, u):
        n = u["node"]
        srcs = self.gsrc.get(id(n), [])
        off = 0
        for p in u["parts"]:
            cp = self._cphys(p)
            for s in srcs:
                if s.kind != 0:
                    raise PlanError("pooled gradient of a concat")
                self._add_gsrc(p, GSrc(s.view.chan(off, cp), 0, (1, 1), s.premasked and off == 0, s.colsums if off == 0 else None))
            off += cp

    def _reduce_sources(self, srcs: List[GSrc], shape_like: TView) -> List[GSrc]:
        """bn_bwd takes at most MAX_GRADSRC sources: pre-sum surplus direct sources."""
        while len(srcs) > L.MAX_GRADSRC:
            direct = [s for s in srcs if s.kind == 0]
            if len(direct) < 2:
                raise PlanError("too many pooled gradient sources")
            a, b = direct[0], direct[1]
            tmp = TView.dense(self.alloc(shape_like.N * shape_like.H * shape_like.W * shape_like.C * 2, "grad"),
                              shape_like.N, shape_like.H, shape_like.W, shape_like.C)
            self.emit(1, L.OP_ELTWISE, L.EltwiseDesc(0, a.view.to_c(), b.view.to_c(), lw.NULL_VIEW.to_c(), tmp.to_c()), "grad pre-sum")
            srcs = [s for s in srcs if s is not a and s is not b] + [GSrc(tmp)]
        return srcs

    def _bwd_conv(self, u):
        n: Node = u["node"]
        a = n.attrs
        kh, kw = a["kernel"]
        out_node = u["out"]
        srcs = list(self.gsrc.get(id(out_node), []))
        if u["pool"] is not None:
            ph, pw_ = u["pool"].attrs["size"]
            for s in self._pool_sources(u["pool"]):
                srcs.append(GSrc(s.view, 1, (ph, pw_), pool_name=u["pool"].name))
        if not srcs:
            return  # dead branch (no gradient reaches it)
        ic = self._im2col_of(n)
        if ic is not None:
            x = Phys(self._im2col_views[ic[:2]], ic[0] * ic[1] * ic[2], [(0, ic[0] * ic[1] * ic[2])])
            kh, kw = 1, 1
        else:
            x = self.phys[id(n.inputs[0])]
        pe = self.pindex[f"{n.name}/kernel"]
        co, cop, cin_p = a["filters"], pe.meta["cout_p"], x.Cp
        H, W, _ = n.shape
        act = self._act_code(u["act"])
        y: TView = u["y"]
        srcs = self._kernel_sources(out_node, srcs, allow_head=u["bn"] is not None and act in (L.ACT_RELU, L.ACT_LEAKY))
        bias_rows = None
        # ---- dZ: gradient w.r.t. the raw convolution output
        if u["bn"] is not None:
            bn = u["bn"]
            dz = self.new_act(H, W, cop, "grad")
            d = L.BnBwdDesc()
            d.x, d.scale, d.shift, d.mean, d.rstd = u["z"].to_c(), u["scale"], u["shift"], u["mean"], u["rstd"]
            d.act, d.n_src = act, len(srcs)
            for i, s in enumerate(srcs):
                d.src[i] = L.GradSrc(s.view.to_c(), s.kind, s.pool[0], s.pool[1])
                if s.kind == 2:
                    hk, hb = f"{s.head['name']}/kernel", f"{s.head['name']}/bias"
                    d.src[i].dlogits, d.src[i].cout = s.head["dlogits"], s.head["cout"]
                    d.src[i].head_w, d.src[i].head_dw, d.src[i].head_db = self.pw(hk), self.pg(hk), self.pg(hb)   # pg(): written by THIS unit
            d.count = float(self.N * H * W)
            nb = max(1, min(1184, (self.N * H * W) // 64))
            d.partials, d.n_blocks = self.alloc(nb * 2 * cop * 4, "scratch"), nb
            d.dgamma, d.dbeta = self.pg(f"{bn.name}/gamma"), self.pg(f"{bn.name}/beta")
            d.dx = dz.to_c()
            d.accumulate = 1   # "zero grads" opens the backward phase
            self.emit(1, L.OP_BN_BWD, d, f"bn bwd {bn.name}")
            has_bias_grad = False
        else:
            if len(srcs) == 1 and srcs[0].kind == 0 and (act == L.ACT_NONE or srcs[0].premasked):
                dz = srcs[0].view
                bias_rows = srcs[0].colsums if srcs[0].premasked else None
            else:
                dz = self.new_act(H, W, cop, "grad")
                d = L.BnBwdDesc()
                d.x, d.act, d.n_src = y.to_c(), act, len(srcs)
                if act == L.ACT_SIGMOID:
                    raise PlanError("sigmoid epilogue backward")
                for i, s in enumerate(srcs):
                    d.src[i] = L.GradSrc(s.view.to_c(), s.kind, s.pool[0], s.pool[1])
                d.count = 1.0
                d.dx = dz.to_c()
                self.emit(1, L.OP_BN_BWD, d, f"act bwd {out_node.name}")
            has_bias_grad = True
        if not u.get("no_tap"):
            self.grad_taps[n.name] = (dz, co)
        # ---- weight / bias gradients
        strided = n.op == "conv" and a["strides"] != (1, 1)
        xin = x.view.parity(0, 0, a["strides"][0], a["strides"][1]) if strided else x.view
        if n.op == "tconv":
            self.emit(1, L.OP_WGRAD, lw.tconv_wgrad(dz, x.view, self.pg(pe.key), cop, kh, kw, cin_p), f"wgrad {n.name}", flops=self._conv_flops(n))
        else:
            self.emit(1, L.OP_WGRAD, lw.conv_wgrad(dz, xin, self.pg(pe.key), cop, kh, kw, cin_p), f"wgrad {n.name}", flops=self._conv_flops(n))
        if u.get("gate") is not None:
            has_bias_grad = False       # the projection feeds a BatchNormalization: its bias gradient is analytically zero (exact 0 emitted)
        if has_bias_grad and bias_rows is not None:
            self.emit(1, L.OP_ROWSUM, L.RowsumDesc(bias_rows[0], bias_rows[1], bias_rows[2], cop, self.pg(f"{n.name}/bias"), 0), f"bias grad {n.name}")
        elif has_bias_grad:
            self.emit(1, L.OP_COLSUM, L.ColsumDesc(dz.to_c(), self.pg(f"{n.name}/bias"), 0, 0), f"bias grad {n.name}")
        # ---- input gradient
        src_node = n.inputs[0]
        if src_node.op == "input":
            return
        Hi, Wi, _ = src_node.shape
        dx = self.new_act(Hi, Wi, cin_p, "grad")
        if strided and u.get("gate") is not None and n is u["gate"]["conv_a"] and "bwd_desc" in u["gate"]:
            # the stride-2 projection of a fused gate: dgrad into a DENSE low-resolution tensor; the gate's second backward pass adds
            # it at the even pixels while it writes dskip = dout * r (no skip-sized memset, strided write or three-tensor sum)
            gt = u["gate"]
            h_, w_, _ = n.shape
            da_low = self.new_act(h_, w_, cin_p, "grad")
            self.emit(1, L.OP_CONV, lw.conv_dgrad(dz, self.pwb(pe.key), cop, kh, kw, cin_p, da_low), f"dgrad {n.name}", flops=self._conv_flops(n))
            d2 = L.GateDesc.from_buffer_copy(gt["bwd_desc"])
            dskip = self._grad_like(gt["skip"])
            d2.dskip, d2.da_low = dskip.to_c(), da_low.to_c()
            # this pass recomputes the resampler from the gate's fp32 parameters: under a sharded / per-bucket optimizer they must not
            # be updated before it has run, so their gradients count as "written" by this unit too (exchange_schedule readiness)
            for key in (f"{gt['tconv'].name}/kernel", f"{gt['tconv'].name}/bias", f"{gt['bn3'].name}/gamma", f"{gt['bn3'].name}/beta"):
                self.pg(key)
            self.emit(1, L.OP_GATE_BWD, d2, f"gate bwd dskip {gt['mul'].name}")
            self._add_gsrc(src_node, GSrc(dskip))
            return
        if strided:
            # 1x1 'valid' stride-s conv reads pixels (s*i, s*j): its input gradient lives on that sub-grid; the other
            # pixels of dx are never written and stay at the zeros the buffer was allocated with (with buffer reuse the
            # memory has had other tenants: zero it first)
            if self.reuse:
                self.emit(1, L.OP_MEMSET, L.MemsetDesc(dx.ptr, self.N * Hi * Wi * cin_p * 2), f"zero dx {n.name}")
            self.emit(1, L.OP_CONV, lw.conv_dgrad(dz, self.pwb(pe.key), cop, kh, kw, cin_p, dx.parity(0, 0, a["strides"][0], a["strides"][1])),
                      f"dgrad {n.name}", flops=self._conv_flops(n))
            self._add_gsrc(src_node, GSrc(dx))
            return
        mul_view, mul_mode, premasked = None, 0, False
        if src_node.op == "concat" and n.op == "conv":
            first = u_first = self.flat_concat[id(src_node)][0]
            pu = self.unit_of_out.get(id(first))
            if (pu is not None and pu["kind"] == "conv" and pu["bn"] is None and pu["act"] is not None and pu["out"] is first
                    and len(self.cons[id(first)]) == 1 and len(self.cons[id(src_node)]) == 1
                    and self._act_code(pu["act"]) in (L.ACT_RELU, L.ACT_LEAKY)):
                mul_view, mul_mode, premasked = x.view.chan(0, self._cphys(first)), self._act_code(pu["act"]), True
        colsums = None
        if n.op == "tconv":
            self.emit(1, L.OP_CONV, lw.tconv_dgrad(dz, self.pwb(pe.key), cop, kh, kw, cin_p, dx), f"dgrad {n.name}", flops=self._conv_flops(n))
        else:
            dd = lw.conv_dgrad(dz, self.pwb(pe.key), cop, kh, kw, cin_p, dx, mul_view, mul_mode)
            if premasked:
                # the masked channels of dx are dL/d(pre-activation) of the transposed conv: its bias gradient is their
                # column sum, which the epilogue's statistics path produces for free (consumed by OP_ROWSUM)
                rows = int(self.stat_rows_fn(dd))
                dd.stats = self.alloc(rows * 2 * cin_p * 4, "scratch")
                colsums = (dd.stats, rows, 2 * cin_p)
            self.emit(1, L.OP_CONV, dd, f"dgrad {n.name}", flops=self._conv_flops(n))
        self._add_gsrc(src_node, GSrc(dx, 0, (1, 1), premasked, colsums))

    # ---------------------------------------------------------------------------------------- weights <-> Keras
    def to_internal(self, key: str, arr: np.ndarray) -> np.ndarray:
        """Keras-layout array -> flat internal layout (padded, kernel-friendly)."""
        e = self.pindex[key]
        m = e.meta
        out = np.zeros(e.size, np.float32)
        arr = np.asarray(arr, np.float32)
        if e.kind == "vec":
            if m.get("vsegs"):
                if m.get("fill") is not None:
                    out[:m["Cp"]] = m["fill"]
                src = 0
                for (po, c) in m["vsegs"]:
                    out[po:po + c] = arr[src:src + c]
                    src += c
            else:
                out[:m["C"]] = arr
                if m.get("fill") is not None:
                    out[m["C"]:ceil8(m["C"])] = m["fill"]
        elif e.kind in ("conv", "tconv"):
            k = arr if arr.ndim == 4 else (arr[None, None] if arr.ndim == 2 else arr[None])   # (kh,kw,Cin,Cout) | tconv (kh,kw,Cout,Cin) | Dense (in,out)
            if m.get("im2col"):
                k = k.reshape(1, 1, -1, k.shape[-1])        # a 1x1 kernel over the K-packed window: index (i*kw + j)*Cin + c
            if e.kind == "conv":
                k = np.transpose(k, (3, 0, 1, 2))               # -> (Cout,kh,kw,Cin)
            else:
                k = np.transpose(k, (2, 0, 1, 3))               # -> (Cout,kh,kw,Cin)
            wl = np.zeros((m["cout"], m["taps"], m["cin_p"]), np.float32)   # logical rows, physical columns
            src = 0
            for (po, c) in m["segs"]:
                wl[:, :, po:po + c] = k[:, :, :, src:src + c].reshape(m["cout"], m["taps"], c)
                src += c
            w = np.zeros((m["cout_p"], m["taps"], m["cin_p"]), np.float32)
            src = 0
            for (po, c) in m.get("out_segs") or [(0, m["cout"])]:
                w[po:po + c] = wl[src:src + c]
                src += c
            out[:w.size] = w.reshape(-1)
        elif e.kind == "head":
            k = arr.reshape(-1, m["cout"])                      # (1,1,Cin,Cout) -> (Cin,Cout)
            w = np.zeros((m["cin_p"], m["cout"]), np.float32)
            src = 0
            for (po, c) in m["segs"]:
                w[po:po + c] = k[src:src + c]
                src += c
            out[:w.size] = w.reshape(-1)
        elif e.kind == "lstm":
            # Keras (kh,kw,Cin,4F), gate order i,f,c,o -> rows [i | c(g) | o | f] x taps x cin_p
            F = m["F"]
            k = arr if arr.ndim == 4 else arr[None]
            k = np.transpose(k, (3, 0, 1, 2))                   # (4F,kh,kw,Cin)
            k = np.concatenate([k[0:F], k[2 * F:3 * F], k[3 * F:4 * F], k[F:2 * F]], 0)
            w = np.zeros((4 * F, m["taps"], m["cin_p"]), np.float32)
            src = 0
            for (po, c) in m["segs"]:
                w[:, :, po:po + c] = k[:, :, :, src:src + c].reshape(4 * F, m["taps"], c)
                src += c
            out[:w.size] = w.reshape(-1)
        elif e.kind == "lstm_bias":
            F = m["F"]
            out[:4 * F] = np.concatenate([arr[0:F], arr[2 * F:3 * F], arr[3 * F:4 * F], arr[F:2 * F]])
        elif e.kind == "blob":
            out[:arr.size] = arr.reshape(-1)
        else:
            raise PlanError(e.kind)
        return out

    def from_internal(self, key: str, flat: np.ndarray) -> np.ndarray:
        e = self.pindex[key]
        m = e.meta
        flat = np.asarray(flat, np.float32)
        if e.kind == "vec":
            if m.get("vsegs"):
                return np.concatenate([flat[po:po + c] for (po, c) in m["vsegs"]]).reshape(e.keras_shape)
            return flat[:m["C"]].copy().reshape(e.keras_shape)
        if e.kind in ("conv", "tconv"):
            w = flat[:m["cout_p"] * m["taps"] * m["cin_p"]].reshape(m["cout_p"], m["taps"], m["cin_p"])
            w = np.concatenate([w[po:po + c] for (po, c) in (m.get("out_segs") or [(0, m["cout"])])], 0)   # logical rows
            cin = sum(c for _, c in m["segs"])
            k = np.zeros((m["cout"], m["taps"], cin), np.float32)
            src = 0
            for (po, c) in m["segs"]:
                k[:, :, src:src + c] = w[:m["cout"], :, po:po + c]
                src += c
            k = k.reshape(m["cout"], m["kh"], m["kw"], cin)
            k = np.transpose(k, (1, 2, 3, 0)) if e.kind == "conv" else np.transpose(k, (1, 2, 0, 3))
            return k.reshape(e.keras_shape)
        if e.kind == "head":
            w = flat[:m["cin_p"] * m["cout"]].reshape(m["cin_p"], m["cout"])
            cin = sum(c for _, c in m["segs"])
            k = np.zeros((cin, m["cout"]), np.float32)
            src = 0
            for (po, c) in m["segs"]:
                k[src:src + c] = w[po:po + c]
                src += c
            return k.reshape(e.keras_shape)
        if e.kind == "lstm":
            F = m["F"]
            w = flat[:4 * F * m["taps"] * m["cin_p"]].reshape(4 * F, m["taps"], m["cin_p"])
            cin = sum(c for _, c in m["segs"])
            k = np.zeros((4 * F, m["taps"], cin), np.float32)
            src = 0
            for (po, c) in m["segs"]:
                k[:, :, src:src + c] = w[:, :, po:po + c]
                src += c
            k = np.concatenate([k[0:F], k[3 * F:4 * F], k[F:2 * F], k[2 * F:3 * F]], 0)   # back to i,f,c,o
            k = k.reshape(4 * F, m["kh"], m["kw"], cin)
            return np.transpose(k, (1, 2, 3, 0)).reshape(e.keras_shape)
        if e.kind == "lstm_bias":
            F = m["F"]
            v = flat[:4 * F]
            return np.concatenate([v[0:F], v[3 * F:4 * F], v[F:2 * F], v[2 * F:3 * F]]).reshape(e.keras_shape)
        if e.kind == "blob":
            return flat[:int(np.prod(e.keras_shape))].copy().reshape(e.keras_shape)
        raise PlanError(e.kind)

    def exchange_schedule(self, bucket_bytes: int = 64 << 20) -> List[Tuple[int, int, int]]:
        """Data-parallel gradient exchange overlapped with backward: [(n_ops, lo, hi)] in execution
        order — once the first n_ops backward ops are enqueued, the gradient arena elements [lo, hi) are final and
        their exchange may start.  Backward walks the layers in reverse, and the arena is laid out in layer order,
        so finished gradients form a suffix of the arena that grows towards offset 0.  Buckets are cut top-down at
        multiples of 4096 elements (Adam is element-wise, so a bucket may split a tensor; 4096 makes every bucket
        divisible into 64-aligned shards for 2, 4 or 8 ranks).  A bucket is ready when every tensor it touches is;
        gradients no backward op writes (conv biases feeding a BatchNormalization: analytically zero) are final once
        the two memsets that open the backward phase have run."""
        n_total = len(self.ops[1])
        untouched = min(2, n_total)
        step = max(4096, bucket_bytes // 4 // 4096 * 4096)
        entries = sorted((e for e in self.params if e.trainable), key=lambda e: e.offset)
        out, hi = [], self.arena_elems
        while hi > 0:
            lo = max(0, hi - step)
            if lo < step // 2:           # do not leave a sliver for the last bucket
                lo = 0
            ready = untouched
            for e in entries:
                if e.offset < hi and e.offset + e.size > lo:
                    ready = max(ready, self.grad_ready.get(e.key, untouched))
            out.append((ready, lo, hi))
            hi = lo
        # a later bucket can never start before an earlier one (single stream of backward ops)
        fixed, floor = [], 0
        for (r, lo, h) in out:
            floor = max(floor, r)
            fixed.append((floor, lo, h))
        return fixed

    def num_launch_ops(self, phase):
        return len(self.ops[phase])
