"""Debug/report tool (not a test): per-layer rel-L2 of the B200 path against the float64 oracle, in layer order.
Usage (GPU box):  python tests/tools_layer_errors.py [1d|2d] [batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200"))

from b2seg.model import Adam  # noqa: E402
from b2seg.models1d import UNet  # noqa: E402
from b2seg.models2d import unet_model_builder  # noqa: E402
from oracle.keras_ref import KerasRef, keras_loss  # noqa: E402
from oracle.ref_models import Ref1D, Ref2D  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "1d"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rng = np.random.default_rng(2)
    if which == "1d":
        m = UNet(1024, 5, 1, 64, 3, problem_type="Classification", output_nums=2, ds=0, is_transconv=True).UNet()
        ref = Ref1D("UNet", 1024, 5, 1, 64, 3, problem_type="Classification", output_nums=2, ds=0)
        x = rng.standard_normal((N, 1024, 1)).astype(np.float32)
        y = np.eye(2, dtype=np.float32)[(x[..., 0] > 0).astype(np.int64)]
        loss, ndim = "cce", 1
    else:
        kw = dict(num_channels=3, output_nums=1, dense_loop=1, is_transconv=True)
        m = unet_model_builder("UNet", 64, 64, 64, 5, train_mode="from_scratch", **kw).ResNet50()
        ref = Ref2D("UNet", 64, 64, 64, 5, **kw)
        x = rng.random((N, 64, 64, 3), dtype=np.float32)
        y = (rng.random((N, 64, 64, 1)) > 0.7).astype(np.float32)
        loss, ndim = "bce", 2
    m.compile(loss=loss, optimizer=Adam(1e-3))
    params = m.get_weight_dict()
    params = {k: torch.from_numpy(v).to(torch.bfloat16).float().numpy() if k.endswith("/kernel") else v for k, v in params.items()}
    m.set_weight_dict(params)
    m.train_on_batch(x, y)
    eng = m._engine(N, True)
    torch.cuda.synchronize()
    tp = {k: torch.from_numpy(v.copy()).double() for k, v in params.items()}
    k = KerasRef(ndim, params=tp, dtype=torch.float64, training=True, strict=True)
    out = ref(k, torch.from_numpy(x).double())[0]
    keras_loss(loss, out, torch.from_numpy(y).double(), logits=k.logits["out"]).backward()
    print(f"{'layer':<28}{'kind':<8}{'act rel-L2':>12}{'grad rel-L2':>12}")
    for name, (view, C, kind) in eng.planner.taps.items():
        if name not in k.acts or kind in ("post", "concat"):
            continue
        got = eng.tap(name).cpu()
        got = got[:, 0] if ndim == 1 else got
        e = rel(got, k.acts[name].detach())
        ge = ""
        if kind == "raw" and name in eng.planner.grad_taps and k.acts[name].grad is not None:
            gg = eng.tap(name, grad=True).cpu()
            gg = gg[:, 0] if ndim == 1 else gg
            ge = f"{rel(gg, k.acts[name].grad):.2e}"
        print(f"{name:<28}{kind:<8}{e:>12.2e}{ge:>12}")
    got = eng.outputs[0]["y"].cpu()
    got = got[:, 0] if ndim == 1 else got
    print("out", rel(got, out.detach()))
    grads = eng.get_grads()
    for key, g in grads.items():
        if tp[key].grad is not None and key.endswith("kernel"):
            print(f"  dW {key:<36}{rel(torch.from_numpy(g), tp[key].grad):.2e}")


if __name__ == "__main__":
    main()
