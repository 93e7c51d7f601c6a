"""Device engine: binds a Planner program to B200 memory and replays it through the C ABI (libb2seg.so).

PyTorch is used only for plumbing — device allocation, streams, pinned staging buffers and (multi-GPU)
torch.distributed/NCCL; every kernel that touches activations, gradients or weights is ours.
There is no CPU fallback: constructing an Engine without a CUDA sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from .graph import Graph
from .planner import Planner


class Engine:
    def __init__(self, graph: Graph, batch: int, training: bool = True, losses: Optional[List[str]] = None,
                 loss_weights: Optional[List[float]] = None, adam: Optional[dict] = None, device: Optional[int] = None,
                 share_params_from: Optional["Engine"] = None, adam_bucket_bytes: int = 0, reuse: bool = False, shard=(0, 1)):
        if not torch.cuda.is_available():
            raise L.B2SegError("b2seg needs a CUDA sm_100 (B200) device: the hot path has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else device
        self.lib = L.load()
        L.check(self.lib.b2seg_device_check(self.device), "device_check")
        self.graph, self.batch, self.training = graph, batch, training
        self._bufs: Dict[int, torch.Tensor] = {}
        self._tagged: Dict[str, torch.Tensor] = {}
        self._shared = share_params_from
        self.dev = torch.device("cuda", self.device)
        self.planner = Planner(graph, batch, self._alloc, training=training, losses=losses, loss_weights=loss_weights, adam=adam,
                               stat_rows_fn=lambda d: self.lib.b2seg_conv_num_stat_rows(C.byref(d)),
                               adam_bucket_bytes=adam_bucket_bytes, reuse=reuse, shard=shard).build()
        self.reuse = reuse
        self.adam_bucket_bytes = adam_bucket_bytes
        p = self.planner
        n = p.arena_elems
        self.w = self._typed("param_w", torch.float32, n)
        self.g = self._typed("param_g", torch.float32, n)
        self.m = self._typed("param_m", torch.float32, n)
        self.v = self._typed("param_v", torch.float32, n)
        self.wb = self._typed("param_wb", torch.bfloat16, n)
        self.moving = self._typed("moving", torch.float32, max(p.n_moving, 64))
        self.loss_buf = self._typed("loss", torch.float32, 1)
        from .planner import LOSS_BUF_FLOATS
        self.logs_buf = self._typed("loss", torch.float32, LOSS_BUF_FLOATS)   # [0] total loss, [8 + 8 i ..] per-output loss + metric sums
        H, W, Cin = graph.inputs[0].shape
        self.input_shape = (batch, H, W, Cin)
        self.x_dev = self._typed("input", torch.float32, batch * H * W * Cin).view(batch, H, W, Cin)
        self.outputs = []
        for o in sorted(p.outputs, key=lambda o: o["index"]):
            numel = int(np.prod(o["shape"]))
            y = self._bufs[o["ptr"]].view(torch.float32)[:numel].view(o["shape"])
            t = self._bufs[o["target_ptr"]].view(torch.float32)[:numel].view(o["shape"]) if training else None
            self.outputs.append(dict(name=o["name"], y=y, target=t, shape=o["shape"]))
        self.plan = C.c_void_p()
        self.reserved_ops = frozenset()
        self._build_plan()
        self.step = 0

    def derive_targets(self):
        """compile(ds_targets=...): fill the targets of outputs 1.. from the mask in output 0's target buffer, on the device —
        a copy where the shapes agree (UNet++-style full-resolution levels), else b2seg_target_pool: window max in 2D
        (helper_functions.py:359-380), window mean in 1D (1D_Segmentation.ipynb cell 31)"""
        src = self.outputs[0]
        N, H, W, Cm = src["shape"]
        for o in self.outputs[1:]:
            n, h, w, c = o["shape"]
            if (n, h, w, c) == (N, H, W, Cm):
                o["target"].copy_(src["target"], non_blocking=True)
                continue
            if c != Cm or h < 1 or w < 1 or H % h or W % w:
                raise ValueError(f"cannot derive the target of '{o['name']}' {o['shape']} from a mask of shape {src['shape']}")
            d = L.TPoolDesc(src["target"].data_ptr(), o["target"].data_ptr(), N, H, W, Cm, H // h, W // w, 0 if self.graph.ndim == 2 else 1)
            L.call("b2seg_target_pool", d, self._stream())

    def _build_plan(self, reserved_ops=frozenset(), reserve_sms=0):
        """(re)create the C plan; backward ops whose index is in reserved_ops size their grids for (SMs - reserve_sms)"""
        if self.plan.value:
            self.lib.b2seg_plan_destroy(self.plan)
            self.plan = C.c_void_p()
        L.check(self.lib.b2seg_plan_create(C.byref(self.plan)), "plan_create")
        try:
            for phase in (0, 1, 2):
                for i, (op, desc, note) in enumerate(self.planner.ops[phase]):
                    r = reserve_sms if (phase == 1 and i in reserved_ops) else 0
                    L.check(self.lib.b2seg_set_backward_sm_reserve(r), "set_backward_sm_reserve")
                    rc = self.lib.b2seg_plan_add(self.plan, phase, op, C.byref(desc), C.sizeof(desc))
                    L.check(rc, f"plan_add[{note}]")
        finally:
            self.lib.b2seg_set_backward_sm_reserve(0)
        self.reserved_ops, self.reserve_sms = frozenset(reserved_ops), reserve_sms
        self.launches = [self.lib.b2seg_plan_num_launches(self.plan, ph) for ph in (0, 1, 2)]

    def timed_phase(self, phase, reps=2):
        """device time of every op of a phase in ms (CUDA events after each op; last of `reps` replays)"""
        n_ops = self.lib.b2seg_plan_num_ops(self.plan, phase)
        buf = (C.c_float * n_ops)()
        for _ in range(reps):
            L.check(self.lib.b2seg_plan_run_timed(self.plan, phase, C.c_void_p(self._stream()), buf, n_ops), "run_timed")
        return [float(v) for v in buf]

    def reserve_sms_for_exchange(self, schedule, group=None, reserve_sms=None):
        """Data parallel.  Every tensor-core kernel here is persistent with one CTA per SM and a static tile split, so the SMs
        the all-reduce's CTAs hold while it runs turn each concurrent launch into two waves (measured at 8 GPUs with NVLS'
        24 CTAs: +1.5 ms per step).  Decide WHICH backward ops overlap the gradient exchange and give only those a grid of
        (SMs - reserve): time the backward ops and one all-reduce once, replay the exchange schedule on those numbers
        (bucket j starts when its last producer op has finished and the wire is free), and rebuild the plan.  The timing
        replays run on whatever the buffers hold; kernel times do not depend on the data."""
        import torch.distributed as dist
        if reserve_sms is None:
            reserve_sms = int(os.environ.get("B2SEG_BWD_SM_RESERVE", "16"))
        if reserve_sms <= 0 or not schedule:
            return
        ms = self.timed_phase(1)
        # wire speed: the largest bucket, twice, second one timed
        (_n, lo, hi) = max(schedule, key=lambda b: b[2] - b[1])
        probe = torch.zeros(hi - lo, dtype=torch.float32, device=self.dev)
        dist.all_reduce(probe, group=group)
        torch.cuda.synchronize(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(probe, group=group)
        e1.record()
        torch.cuda.synchronize(self.dev)
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)    # every rank must take the same decision
        ms_per_byte = float(t.item()) / ((hi - lo) * 4)
        tms = torch.tensor(ms, device=self.dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX, group=group)
        ms = [float(v) for v in tms.tolist()]
        self.exchange_calibration = {"allreduce_GBps": 1e-6 / ms_per_byte, "backward_ms": sum(ms)}
        from .dist import ops_overlapping_exchange
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        reserved, windows = ops_overlapping_exchange(ms, schedule, ms_per_byte, sms / float(sms - reserve_sms))
        self.exchange_calibration["reserved_ops"] = len(reserved)
        self.exchange_calibration["exchange_windows_ms"] = [(round(a, 3), round(b, 3)) for (a, b) in windows]
        self._build_plan(reserved, reserve_sms)

    # ---- memory ------------------------------------------------------------------------------------------
    def _alloc(self, nbytes, tag):
        if self._shared is not None and tag in ("param_w", "param_g", "param_m", "param_v", "param_wb", "moving"):
            t = self._shared._tagged[tag]
            assert t.numel() >= nbytes
        else:
            t = torch.zeros(nbytes, dtype=torch.uint8, device=self.dev)
        self._bufs[t.data_ptr()] = t
        if tag in ("param_w", "param_g", "param_m", "param_v", "param_wb", "moving", "loss", "input"):
            self._tagged[tag] = t
        return t.data_ptr()

    def _typed(self, tag, dtype, numel):
        return self._tagged[tag].view(dtype)[:numel]

    def memory_bytes(self):
        return sum(t.numel() for t in self._bufs.values())

    # ---- weights -----------------------------------------------------------------------------------------
    def set_weights(self, params: Dict[str, np.ndarray], strict=True):
        p = self.planner
        w_host = self.w.cpu().numpy()
        mov_host = self.moving.cpu().numpy()
        for e in p.params:
            if e.key not in params:
                if strict:
                    raise KeyError(f"missing weight {e.key}")
                continue
            arr = np.asarray(params[e.key], np.float32)
            if tuple(arr.shape) != tuple(e.keras_shape):
                raise ValueError(f"weight {e.key}: shape {arr.shape} != {e.keras_shape}")
            flat = p.to_internal(e.key, arr)
            (w_host if e.trainable else mov_host)[e.offset:e.offset + e.size] = flat
        self.w.copy_(torch.from_numpy(w_host))
        self.moving.copy_(torch.from_numpy(mov_host))
        self.wb.copy_(self.w.to(torch.bfloat16))

    def get_weights(self) -> Dict[str, np.ndarray]:
        p = self.planner
        w_host, mov_host = self.w.cpu().numpy(), self.moving.cpu().numpy()
        return {e.key: p.from_internal(e.key, (w_host if e.trainable else mov_host)[e.offset:e.offset + e.size]) for e in p.params}

    def get_grads(self) -> Dict[str, np.ndarray]:
        p = self.planner
        g_host = self.g.cpu().numpy()
        return {e.key: p.from_internal(e.key, g_host[e.offset:e.offset + e.size]) for e in p.params if e.trainable}

    def reset_optimizer(self):
        self.m.zero_()
        self.v.zero_()
        self.step = 0

    # ---- execution ---------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def run(self, phase):
        L.check(self.lib.b2seg_plan_run(self.plan, phase, C.c_void_p(self._stream())), f"plan_run[{phase}]")

    def run_range(self, phase, first_op, n_ops):
        L.check(self.lib.b2seg_plan_run_range(self.plan, phase, first_op, n_ops, C.c_void_p(self._stream())), f"plan_run_range[{phase}]")

    def forward(self):
        self.run(0)

    def backward(self):
        self.run(1)

    def optimizer_begin(self, lr: float, grad_scale: float = 1.0, step: Optional[int] = None):
        """step: Adam's t for this update.  The moments live in arenas that several engines of one Model share (one engine per
        batch size), so the Model owns the counter and passes it in; a stand-alone engine counts its own steps."""
        self.step = self.step + 1 if step is None else int(step)
        L.check(self.lib.b2seg_plan_set_adam(self.plan, lr, self.step, grad_scale), "set_adam")

    def optimizer_step(self, lr: float, grad_scale: float = 1.0, step: Optional[int] = None):
        self.optimizer_begin(lr, grad_scale, step)
        self.run(2)

    def tap(self, name: str, grad=False) -> torch.Tensor:
        """activation (or raw-conv-output gradient) of a layer as an fp32 NHWC tensor with logical channels"""
        if self.reuse:
            raise L.B2SegError("this engine reuses activation memory (a tensor's bytes are overwritten once nothing in the step needs them): "
                               "build the model with keep_activations=True (or B2SEG_KEEP_ACTIVATIONS=1) to tap layers")
        p = self.planner
        view, Cn = (p.grad_taps[name] if grad else p.taps[name][:2])
        base = None
        for ptr, t in self._bufs.items():
            if ptr <= view.ptr < ptr + t.numel():
                base = t
                break
        off = (view.ptr - base.data_ptr()) // 2
        flat = base.view(torch.bfloat16)
        out = torch.as_strided(flat, (view.N, view.H, view.W, view.C), (view.sn, view.sh, view.sw, 1), off)
        idx = p.logical_channels(name)
        if idx == list(range(Cn)):
            return out[..., :Cn].float()
        return out[..., torch.tensor(idx, device=out.device)].float()

    def __del__(self):
        try:
            if getattr(self, "plan", None) is not None and self.plan.value:
                self.lib.b2seg_plan_destroy(self.plan)
                self.plan = C.c_void_p()
        except Exception:
            pass
