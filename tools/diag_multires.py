"""GPU diagnostic: is a MultiResUNet step deterministic, and where do keep_activations=True / False engines part?
Prints per-layer activation differences between two identical keep engines (run-to-run noise) and per-gradient differences
between keep/keep and keep/reuse."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200"))
from b2seg.model import Adam  # noqa: E402
from b2seg.models2d import unet_model_builder  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def main():
    size, width, depth = 64, 32, 3
    rng = np.random.default_rng(19)
    x = rng.random((4, size, size, 3), dtype=np.float32)
    y = (rng.random((4, size, size, 1)) > 0.6).astype(np.float32)
    ms = []
    for keep in (True, True, False, False):
        m = unet_model_builder("MultiResUNet", size, size, width, depth, train_mode="from_scratch", num_channels=3).ResNet50()
        m.keep_activations = keep
        m.compile(loss="bce", optimizer=Adam(1e-3))
        if ms:
            m.set_weight_dict(ms[0].get_weight_dict())
        ms.append(m)
    losses = [m.train_on_batch(x, y) for m in ms]
    torch.cuda.synchronize()
    print("losses", losses)
    engs = [m._engine(4, True) for m in ms]
    outs = [np.array(e.outputs[0]["y"].float().cpu()) for e in engs]
    for i in range(1, 4):
        print(f"out[{i}] vs out[0]: {rel(outs[i], outs[0]):.2e}")
    grads = [e.get_grads() for e in engs]
    print("gradients: keep/keep   keep/reuse   reuse/reuse")
    for k in grads[0]:
        if k.endswith("/kernel"):
            print(f"  {k:40s} {rel(grads[1][k], grads[0][k]):.2e}  {rel(grads[2][k], grads[0][k]):.2e}  {rel(grads[3][k], grads[2][k]):.2e}")
    print("activations keep/keep (first 60 layers in plan order, forward taps then gradient taps)")
    p = engs[0].planner
    n = 0
    for name in p.taps:
        a, b = engs[0].tap(name).cpu().numpy(), engs[1].tap(name).cpu().numpy()
        r = rel(b, a)
        if r > 1e-3 or n < 8:
            print(f"  fwd {name:40s} {r:.2e}")
        n += 1
    for name in p.grad_taps:
        a, b = engs[0].tap(name, grad=True).cpu().numpy(), engs[1].tap(name, grad=True).cpu().numpy()
        r = rel(b, a)
        if r > 1e-2:
            print(f"  bwd {name:40s} {r:.2e}")


if __name__ == "__main__":
    main()
