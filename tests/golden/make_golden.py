"""Writes the golden fixtures of tests/golden/*.npz from the float64 oracle (oracle/keras_ref.py + oracle/ref_models.py).

The reference is TensorFlow/Keras code; TensorFlow is not installable in the build container, so these vectors come from the
oracle, not from a run of the reference (parity unpinned).  They serve two purposes: tests/test_oracle_cpu.py
re-runs the oracle against them (a change of the oracle's semantics cannot slip in unnoticed), and tests/test_gpu_golden.py
loads the stored weights into the product, runs one training step on the B200 and compares outputs / loss / gradients with the
stored values — on the GPU box nothing is recomputed on the CPU for these cases.

usage (repo root):  python tests/golden/make_golden.py        # rewrites every fixture
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.keras_ref import KerasRef, keras_loss  # noqa: E402
from oracle.ref_models import Ref1D, Ref2D  # noqa: E402

# name, dims, builder arguments (the reference constructor signatures), batch, losses per output
CASES = [
    dict(name="unet2d_d2_w8", ndim=2, variant="UNet", args=(16, 16, 8, 2), kw=dict(num_channels=3), batch=2, losses=["bce"]),
    dict(name="unet1d_d3_w8_k3", ndim=1, variant="UNet", args=(64, 3, 1, 8, 3), kw=dict(problem_type="Classification", output_nums=2, ds=0),
         batch=3, losses=["cce"]),
    dict(name="unetpp2d_ds_ag", ndim=2, variant="UNetPP", args=(16, 16, 8, 2), kw=dict(num_channels=2, ds=1, ag=1, output_nums=4, final_activation="softmax"),
         batch=2, losses=["cce", "mse", "mse"]),
    # (width 32: branches of 5 / 10 / 16 channels.  At width 8 the first kernel has 9 elements and its free-running gradient error is a coin toss.)
    dict(name="multires2d", ndim=2, variant="MultiResUNet", args=(16, 16, 32, 2), kw=dict(num_channels=1), batch=2, losses=["bce"]),
    dict(name="bcdunet1d_lstm_ds", ndim=1, variant="BCDUNet", args=(32, 2, 2, 16, 3), kw=dict(ds=1, lstm=1), batch=2, losses=["mse", "mse", "mse"]),
]


def _ref(spec):
    return (Ref2D if spec["ndim"] == 2 else Ref1D)(spec["variant"], *spec["args"], **spec["kw"])


def run_case(spec):
    """-> {array name: float32/float64 ndarray}; deterministic"""
    rng = np.random.default_rng(abs(hash(spec["name"])) % 1000 if False else sum(map(ord, spec["name"])))
    ndim = spec["ndim"]
    if ndim == 2:
        H, W = spec["args"][0], spec["args"][1]
        x = rng.random((spec["batch"], H, W, spec["kw"].get("num_channels", 3)), dtype=np.float32)
    else:
        x = rng.standard_normal((spec["batch"], spec["args"][0], spec["args"][2])).astype(np.float32)
    torch.manual_seed(0)
    # weights: the oracle's own Keras-style initialisers, stored as float32 (what a Keras weight file holds), made
    # bf16-representable for conv kernels so the B200 path (bf16 tensor-core operands) starts from identical values
    k0 = KerasRef(ndim, params={}, dtype=torch.float64, training=True, seed=11, strict=False)
    _ref(spec)(k0, torch.from_numpy(x).double())
    params = {}
    for key, v in k0.params.items():
        a = v.detach().float()
        if key.endswith("/kernel"):
            a = a.to(torch.bfloat16).float()
        elif key.endswith("/gamma"):
            a = a + 0.2 * torch.from_numpy(rng.standard_normal(tuple(a.shape)).astype(np.float32))
        elif key.endswith("/beta") or key.endswith("/bias"):
            a = a + 0.1 * torch.from_numpy(rng.standard_normal(tuple(a.shape)).astype(np.float32))
        params[key] = a.numpy()
    tp = {kk: torch.from_numpy(v.copy()).double() for kk, v in params.items()}
    k = KerasRef(ndim, params=tp, dtype=torch.float64, training=True, strict=True)
    outs = _ref(spec)(k, torch.from_numpy(x).double())
    assert len(outs) == len(spec["losses"]), (spec["name"], len(outs))
    res = {"x": x}
    total = 0
    for i, (o, kind) in enumerate(zip(outs, spec["losses"])):
        name = next(n for n, t in k.acts.items() if t is o)
        if kind == "bce":
            t = (rng.random(tuple(o.shape)) > 0.6).astype(np.float32)
        elif kind == "cce":
            t = np.eye(o.shape[-1], dtype=np.float32)[rng.integers(0, o.shape[-1], tuple(o.shape[:-1]))]
        else:
            t = rng.standard_normal(tuple(o.shape)).astype(np.float32)
        res[f"target{i}"] = t
        res[f"out{i}"] = o.detach().numpy()
        total = total + keras_loss(kind, o, torch.from_numpy(t).double(), logits=k.logits.get(name))
    total.backward()
    res["loss"] = np.array(float(total.detach()))
    for key, v in params.items():
        res["param/" + key] = v
    kernels = [kk for kk in tp if kk.endswith("/kernel") and tp[kk].grad is not None]
    gammas = [kk for kk in tp if kk.endswith("/gamma") and tp[kk].grad is not None]
    for key in [kernels[0], kernels[len(kernels) // 2], kernels[-1], gammas[0], gammas[-1]]:
        res["grad/" + key] = tp[key].grad.numpy()
    mk = sorted(k.new_moving)[0]
    res["moving/" + mk] = k.new_moving[mk].numpy()
    return res


def main():
    for spec in CASES:
        res = run_case(spec)
        path = os.path.join(HERE, spec["name"] + ".npz")
        np.savez_compressed(path, **res)
        print(f"{path}: {len(res)} arrays, {os.path.getsize(path) / 1024:.0f} KiB, loss {float(res['loss']):.6f}")


if __name__ == "__main__":
    main()
