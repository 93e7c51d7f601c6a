"""Test infrastructure: a stand-in for b2seg.engine.Engine that replays the planner's program on the float64 CPU descriptor emulator
(tests/desc_emulator.py) instead of the GPU.  It exists so that the PYTHON side of the GPU test-suite — model facade, target
handling, tap names, oracle teacher forcing, tolerances plumbing — can be dry-run where no GPU exists (tests/test_gpu_dryrun_cpu.py).
It proves nothing about the CUDA kernels and is never importable from the product: the product's Engine raises without a B200."""
from typing import Dict

import numpy as np
import torch

from b2seg import _lib as L
from b2seg.helpers import derive_targets_host
from b2seg.planner import Planner
from desc_emulator import PlanMem, run_phase


_PARAM_TAGS = ("param_w", "param_g", "param_m", "param_v", "param_wb", "moving")


class CpuEngine:
    def __init__(self, graph, batch, training=True, losses=None, loss_weights=None, adam=None, device=None, share_params_from=None,
                 adam_bucket_bytes=0, reuse=False, shard=(0, 1)):
        self.graph, self.batch, self.training = graph, batch, training
        self.mem = PlanMem()
        self._tag_ptr = {}
        share = share_params_from
        if share is not None:
            # like the real Engine, alias the primary's parameter / gradient / Adam-moment / moving-statistics arenas: same addresses,
            # same storage; this engine's own buffers live above the primary's address range
            self.mem.next = share.mem.next + (1 << 32)
            for tag in _PARAM_TAGS:
                ptr = share._tag_ptr[tag]
                self.mem.bufs.append(next(b for b in share.mem.bufs if b[0] == ptr))

        def alloc(nbytes, tag="act"):
            if share is not None and tag in _PARAM_TAGS:
                return share._tag_ptr[tag]
            ptr = self.mem.alloc_bytes(nbytes, tag)
            if tag in _PARAM_TAGS:
                self._tag_ptr[tag] = ptr
            return ptr
        self.planner = p = Planner(graph, batch, alloc, training=training, losses=losses, loss_weights=loss_weights,
                                   adam=adam, adam_bucket_bytes=adam_bucket_bytes, reuse=reuse, shard=shard).build()
        self.reuse = reuse
        self.adam_bucket_bytes = adam_bucket_bytes
        H, W, Cin = graph.inputs[0].shape
        self.x_dev = torch.zeros(batch, H, W, Cin, dtype=torch.float32)
        self.outputs = []
        for o in sorted(p.outputs, key=lambda o: o["index"]):
            numel = int(np.prod(o["shape"]))
            y = self.mem.f32(o["ptr"], numel).view(o["shape"])
            t = self.mem.f32(o["target_ptr"], numel).view(o["shape"]) if training else None
            self.outputs.append(dict(name=o["name"], y=y, target=t, shape=o["shape"]))
        self.loss_buf = self.mem.f32(p.loss_ptr, 1)
        from b2seg.planner import LOSS_BUF_FLOATS
        self.logs_buf = self.mem.f32(p.loss_ptr, LOSS_BUF_FLOATS)
        n = p.arena_elems
        self.w, self.g = self.mem.f32(p.w_ptr, n), self.mem.f32(p.g_ptr, n)          # flat arenas (views), as Engine exposes them
        self.m, self.v = self.mem.f32(p.m_ptr, n), self.mem.f32(p.v_ptr, n)
        self.moving = self.mem.f32(p.mov_ptr, max(p.n_moving, 64))
        self.wb = self.mem.f32(p.wb_ptr, n)                                           # the copy the emulated kernels read (bf16 on the device)
        self.step = 0
        self.dev = torch.device("cpu")
        self._shared = share_params_from

    # ---- weights
    def _arena(self, ptr, e):
        return self.mem.f32(ptr + 4 * e.offset, e.size)

    def set_weights(self, params: Dict[str, np.ndarray], strict=True):
        p = self.planner
        for e in p.params:
            if e.key not in params:
                if strict:
                    raise KeyError(f"missing weight {e.key}")
                continue
            flat = torch.from_numpy(p.to_internal(e.key, np.asarray(params[e.key], np.float32))).double()
            if e.trainable:
                self._arena(p.w_ptr, e)[:] = flat
                wb, off = self.mem.resolve(p.wb_ptr + 2 * e.offset)
                wb[off:off + e.size] = flat
            else:
                self._arena(p.mov_ptr, e)[:] = flat

    def get_weights(self):
        p = self.planner
        return {e.key: p.from_internal(e.key, self._arena(p.w_ptr if e.trainable else p.mov_ptr, e).numpy().astype(np.float32)) for e in p.params}

    def get_grads(self):
        p = self.planner
        return {e.key: p.from_internal(e.key, self._arena(p.g_ptr, e).numpy().astype(np.float32)) for e in p.params if e.trainable}

    # ---- execution
    def _poison(self):
        """reuse mode: the arena holds NaN when a step starts, so a read of bytes no op of THIS step has written shows up in the result"""
        if self.reuse:
            for tag, ptr, nbytes in self.planner.buffers:
                if tag == "arena":
                    self.mem.resolve(ptr)[0].fill_(float("nan"))

    def forward(self):
        self._poison()
        self.mem.f32(self.planner.input_ptr, self.x_dev.numel())[:] = self.x_dev.reshape(-1).double()
        run_phase(self.mem, self.planner, 0)

    def backward(self):
        run_phase(self.mem, self.planner, 1)

    def run(self, phase):
        if phase == 0:
            return self.forward()
        run_phase(self.mem, self.planner, phase)

    def run_range(self, phase, first_op, n_ops):
        if phase == 0 and first_op == 0:
            self._poison()
            self.mem.f32(self.planner.input_ptr, self.x_dev.numel())[:] = self.x_dev.reshape(-1).double()
        run_phase(self.mem, self.planner, phase, first_op, n_ops)

    def reset_optimizer(self):
        self.m.zero_()
        self.v.zero_()
        self.step = 0

    def optimizer_begin(self, lr, grad_scale=1.0, step=None):
        self.step = self.step + 1 if step is None else int(step)
        for (op, d, _note) in self.planner.ops[2]:
            if op == L.OP_ADAM:
                d.lr, d.step, d.grad_scale = lr, self.step, grad_scale

    def optimizer_step(self, lr, grad_scale=1.0, step=None):
        self.optimizer_begin(lr, grad_scale, step)
        run_phase(self.mem, self.planner, 2)

    def derive_targets(self):
        mask = self.outputs[0]["target"].numpy()
        for o, t in zip(self.outputs[1:], derive_targets_host(mask, [o["shape"] for o in self.outputs[1:]], self.graph.ndim)):
            o["target"].copy_(torch.from_numpy(np.ascontiguousarray(t)))

    def tap(self, name, grad=False):
        assert not self.reuse, "taps need keep_activations=True"
        p = self.planner
        view = (p.grad_taps[name] if grad else p.taps[name])[0]
        return self.mem.gather_view(view.to_c())[..., p.logical_channels(name)].float()
