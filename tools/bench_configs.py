"""Train-step throughput of the five BASELINE.json configs at their full sizes (not the bench contract — bench.py measures
config 2 — but the evidence that every graph family runs at scale through the same kernels).
usage (GPU box): python tools/bench_configs.py [cfg numbers ...]   e.g.  python tools/bench_configs.py 1 3 4 5"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200"))

from b2seg.model import Adam  # noqa: E402
from b2seg.models1d import UNet  # noqa: E402
from b2seg.models2d import unet_model_builder  # noqa: E402


def cfg(n):
    """(name, model, x shape, targets, loss, loss_weights, fwd GFLOP / sample from SURVEY 8(d))"""
    rng = np.random.default_rng(n)
    if n == 1:
        m = UNet(1024, 5, 1, 64, 3, problem_type="Classification", output_nums=2, ds=0, ae=0, ag=0, lstm=0, is_transconv=True).UNet()
        x = rng.standard_normal((32, 1024, 1)).astype(np.float32)
        y = np.eye(2, dtype=np.float32)[(x[..., 0] > 0).astype(np.int64)]
        return "cfg1 1D UNet d5 w64 L1024", m, x, y, "categorical_crossentropy", None, 5.23
    kw = dict(train_mode="from_scratch", is_transconv=True)
    if n in (2, 6):   # 6 = config 2 with the auto-encoder bottleneck (ae=1: two Dense layers of 134 M parameters each)
        m = unet_model_builder("UNet", 256, 256, 64, 5, num_channels=3, output_nums=1, dense_loop=1, ae=1 if n == 6 else 0, **kw).ResNet50()
        B, S, c, loss, gf = 32, 256, 3, "binary_crossentropy", 91.77 + (0.537 if n == 6 else 0.0)
    elif n == 3:
        m = unet_model_builder("UNetPP", 256, 256, 64, 5, num_channels=3, output_nums=4, ds=1, ag=1, final_activation="softmax", **kw).ResNet50()
        B, S, c, loss, gf = 8, 256, 3, None, 342.21
    elif n == 4:
        m = unet_model_builder("MultiResUNet", 512, 512, 64, 5, num_channels=1, output_nums=1, alpha=1.0, **kw).ResNet50()
        B, S, c, loss, gf = 8, 512, 1, "binary_crossentropy", 529.76
    else:
        m = unet_model_builder("UNet", 256, 256, 64, 5, num_channels=3, output_nums=1, lstm=1, dense_loop=3, **kw).ResNet50()
        B, S, c, loss, gf = 32, 256, 3, "binary_crossentropy", 186.0
    x = rng.random((B, S, S, c), dtype=np.float32)
    if n == 3:
        lab = rng.integers(0, 4, (B, S, S))
        y = {"out": np.eye(4, dtype=np.float32)[lab]}
        for name in m.output_names[1:]:
            y[name] = (lab > 0).astype(np.float32)[..., None]
        loss = {"out": "categorical_crossentropy", **{name: "mse" for name in m.output_names[1:]}}
        return "cfg3 2D UNet++ DS+AG 4 classes", m, x, y, loss, None, gf
    y = (rng.random((B, S, S, 1)) > 0.7).astype(np.float32)
    return {2: "cfg2 2D UNet d5 w64 256^2", 6: "cfg2 + ae=1 (Dense 131072 -> 1024 -> 131072)", 4: "cfg4 2D MultiResUNet 512^2x1", 5: "cfg5 2D UNet lstm=1 dense_loop=3 (BCDUNet)"}[n], m, x, y, loss, None, gf


def main():
    which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5]
    for n in which:
        name, m, x, y, loss, lw, gf = cfg(n)
        t0 = time.time()
        m.compile(loss=loss, optimizer=Adam(2e-4), loss_weights=lw)
        l0 = m.train_on_batch(x, y)
        eng = m._engine(x.shape[0], True)
        build_s = time.time() - t0
        for _ in range(2):
            m._step(eng, return_loss=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        e0.record()
        for _ in range(steps):
            m._step(eng, return_loss=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        l1 = m.train_on_batch(x, y)
        B = x.shape[0]
        print(json.dumps({"config": name, "batch": B, "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1),
                          "train_tflops": round(3 * gf * B / ms, 1), "loss_first": round(float(l0), 5), "loss_after_9_steps": round(float(l1), 5),
                          "params": int(m.count_params()), "device_memory_gb": round(eng.memory_bytes() / 2 ** 30, 2),
                          "launches_per_step": int(sum(eng.launches)), "build_s": round(build_s, 1)}), flush=True)
        if os.environ.get("BENCH_CONFIGS_OPS"):
            # where the step goes: device time per op kind (CUDA events after every op of one replay)
            import collections
            from b2seg import _lib as L
            names = {v: k for k, v in vars(L).items() if k.startswith("OP_") and isinstance(v, int)}
            agg = collections.defaultdict(lambda: [0, 0.0])
            for phase in (0, 1, 2):
                for i, t in enumerate(eng.timed_phase(phase)):
                    info = eng.planner.op_info[(phase, i)]
                    key = ("fwd", "bwd", "opt")[phase] + " " + names.get(info["op"], str(info["op"]))
                    agg[key][0] += 1
                    agg[key][1] += t
            tot = sum(v[1] for v in agg.values())
            for k2, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
                print(f"    {k2:<22} x{v[0]:<4d} {v[1]:8.3f} ms  {100 * v[1] / tot:5.1f}%", flush=True)
        del m, eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
