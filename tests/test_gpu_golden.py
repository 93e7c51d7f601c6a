"""B200 path vs the committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the float64 oracle):
the stored Keras-layout weights are loaded by layer name, one training step runs through the builder API -> planner -> C ABI ->
kernels, and outputs, loss, gradients and BN moving statistics are compared with the stored values.  Nothing is recomputed on
the CPU here and /root/reference is not needed.

Tolerances: these are free-running (not teacher-forced) comparisons of small models in bf16, see tests/test_gpu_model.py for why
outputs get 3e-2 and gradients a loose bound; the per-layer 1e-2 criterion is asserted there."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
from make_golden import CASES  # noqa: E402  (specs only; run_case is not called on the GPU box)

from b2seg.model import Adam  # noqa: E402
from b2seg.models1d import BCDUNet, UNet  # noqa: E402
from b2seg.models2d import unet_model_builder  # noqa: E402


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def _model(spec):
    if spec["ndim"] == 2:
        H, W, width, depth = spec["args"]
        return unet_model_builder(spec["variant"], H, W, width, depth, train_mode="from_scratch", **spec["kw"]).ResNet50()
    if spec["variant"] == "BCDUNet":
        return BCDUNet(*spec["args"], **spec["kw"]).BCDUNet()
    return getattr(UNet(*spec["args"], **spec["kw"]), spec["variant"])()


@pytest.mark.parametrize("spec", CASES, ids=[c["name"] for c in CASES])
def test_training_step_matches_golden(spec):
    gold = np.load(os.path.join(GOLDEN, spec["name"] + ".npz"))
    m = _model(spec)
    losses = spec["losses"]
    m.compile(loss=losses if len(losses) > 1 else losses[0], optimizer=Adam(1e-3))
    params = {k[len("param/"):]: gold[k] for k in gold.files if k.startswith("param/")}
    mine = m.get_weight_dict()
    # the oracle also initialises layers of dangling branches that Keras prunes (MultiResUNet's last ResPath): ignore those
    assert set(mine) <= set(params), sorted(set(mine) - set(params))[:5]
    m.set_weight_dict({k: params[k] for k in mine})
    targets = [gold[f"target{i}"] for i in range(len(losses))]
    loss = m.train_on_batch(gold["x"], targets if len(targets) > 1 else targets[0])
    eng = m._engine(gold["x"].shape[0], True)
    torch.cuda.synchronize()
    assert abs(loss - float(gold["loss"])) < 2e-2 * max(1.0, abs(float(gold["loss"]))), (loss, float(gold["loss"]))
    for i, o in enumerate(eng.outputs):
        got = o["y"].cpu().numpy()
        if spec["ndim"] == 1:
            got = got[:, 0]
        assert rel_l2(got, gold[f"out{i}"]) < 5e-2, (o["name"], rel_l2(got, gold[f"out{i}"]))
    grads = eng.get_grads()
    gkeys = [k for k in gold.files if k.startswith("grad/")]   # stored order: first / middle / last kernel, first / last gamma
    for pos, k in enumerate(gkeys):
        e = rel_l2(grads[k[len("grad/"):]], gold[k])
        # gradients next to the loss see almost no accumulated noise; the first layers sit behind every ReLU mask of the model,
        # where bf16 storage noise flips a fraction of the masks (tests/tools_bf16_noise_sim.py): loose bound only
        print(f"[golden {spec['name']}] {k}: rel-L2 {e:.3f}")
        assert e < (0.05 if pos in (2, 4) else 0.6), (k, e)
    after = m.get_weight_dict()
    for k in gold.files:
        if k.startswith("moving/"):
            assert np.allclose(after[k[len("moving/"):]], gold[k], rtol=2e-2, atol=2e-3), k
