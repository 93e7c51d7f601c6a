"""GPU diagnostic: layer by layer, the device against the CPU numerics model (float64 descriptor emulator with bf16 storage) on
a golden fixture — where do the two part?  usage: python tools/diag_numerics_model.py [fixture name]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tf-1d-2d-segmentation-end2endpipelines_b200", "tests", "tests/golden", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import b2seg.engine  # noqa: E402
import desc_emulator  # noqa: E402
from b2seg.model import Adam  # noqa: E402
from cpu_engine import CpuEngine  # noqa: E402
from make_golden import CASES  # noqa: E402
from test_gpu_golden import _model  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "multires2d"
    spec = [c for c in CASES if c["name"] == name][0]
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    losses = spec["losses"]
    params = {k[len("param/"):]: gold[k] for k in gold.files if k.startswith("param/")}
    targets = [gold[f"target{i}"] for i in range(len(losses))]

    def run():
        m = _model(spec)
        m.keep_activations = True
        m.compile(loss=losses if len(losses) > 1 else losses[0], optimizer=Adam(1e-3))
        m.set_weight_dict({k: params[k] for k in m.get_weight_dict()})
        loss = m.train_on_batch(gold["x"], targets if len(targets) > 1 else targets[0])
        eng = m._engine(gold["x"].shape[0], True)
        taps = {n: np.array(eng.tap(n).cpu()) for n in eng.planner.taps}
        gtaps = {n: np.array(eng.tap(n, grad=True).cpu()) for n in eng.planner.grad_taps}
        return loss, taps, gtaps, list(eng.planner.taps), list(eng.planner.grad_taps)
    loss_d, taps_d, g_d, order, gorder = run()
    torch.cuda.synchronize()
    b2seg.engine.Engine = CpuEngine
    desc_emulator.ROUND_BF16 = "1"
    loss_m, taps_m, g_m, _, _ = run()
    print("loss", loss_d, loss_m)
    for n in order:
        a, b = taps_d[n], taps_m[n]
        nd = int((a != b).sum())
        print(f"fwd {n:36s} rel {rel(a, b):.2e}  differing elements {nd}/{a.size}  max|d| {float(np.abs(a - b).max()):.3e}  max|v| {float(np.abs(b).max()):.3e}")
    for n in gorder:
        a, b = g_d[n], g_m[n]
        print(f"bwd {n:36s} rel {rel(a, b):.2e}")


if __name__ == "__main__":
    main()
