"""Every loss of 2DCNN/utils/tf_losses.py the planner lowers (b2seg_loss kinds 0..14), on sigmoid, softmax and linear heads: the
explicit derivative formulas of the kernel (mirrored in float64 by tests/desc_emulator.py:emu_loss) against autograd through the
oracle's restatement of the Keras definitions (oracle/keras_ref.py:keras_loss) — loss value, every gradient, one Adam step."""
import numpy as np
import pytest
import torch

from b2seg.models1d import UNet
from b2seg.models2d import unet_model_builder
from oracle.ref_models import Ref1D, Ref2D
from test_plan_cpu import _run

LOSSES = ["bce", "cce", "mse", "mae", "msle", "huber", "logcosh", "focal", "poisson", "kld", "hinge", "squared_hinge", "mape",
          "categorical_hinge", "cosine"]


def _target(kind, shape, rng):
    C = shape[-1]
    if kind in ("cce", "categorical_hinge", "kld") and C > 1:
        return torch.from_numpy(np.eye(C, dtype=np.float32)[rng.integers(0, C, shape[:-1])])
    if kind in ("bce", "focal", "hinge", "squared_hinge"):
        return torch.from_numpy((rng.random(shape) > 0.6).astype(np.float32))
    if kind in ("msle", "poisson", "kld"):
        return torch.from_numpy(rng.random(shape).astype(np.float32) * 2.0)
    if kind == "mape":
        return torch.from_numpy((0.5 + rng.random(shape)).astype(np.float32))
    return torch.from_numpy(rng.standard_normal(shape).astype(np.float32))


@pytest.mark.parametrize("kind", LOSSES)
@pytest.mark.parametrize("head", ["sigmoid", "softmax", "linear"])
def test_loss_on_head_matches_oracle(kind, head):
    rng = np.random.default_rng(sum(map(ord, kind + head)))
    if (kind in ("bce", "focal") and head == "softmax") or (kind == "cce" and head == "sigmoid"):
        from b2seg.planner import PlanError, Planner
        g = unet_model_builder("UNet", 16, 16, 8, 2, train_mode="from_scratch", num_channels=2, output_nums=2, final_activation=head).build_graph()
        with pytest.raises(PlanError, match="cross-entropy needs its own activation"):   # (Keras 2 would feed the other activation's logits)
            Planner(g, 2, lambda n, tag="act": 1 << 20, training=True, losses=[kind]).build()
        return
    if kind == "poisson" and head == "linear":
        pytest.skip("log of negative predictions: NaN in Keras as well")
    if head == "linear":
        g = UNet(32, 2, 2, 8, 3, problem_type="Regression", output_nums=3, ds=0).UNet().graph
        ref, ndim = Ref1D("UNet", 32, 2, 2, 8, 3, problem_type="Regression", output_nums=3, ds=0), 1
        x = torch.from_numpy(rng.standard_normal((2, 32, 2)).astype(np.float32))
        shape = (2, 32, 3)
    else:
        kw = dict(num_channels=2, output_nums=3 if head == "softmax" else 2, final_activation=head)
        g = unet_model_builder("UNet", 16, 16, 8, 2, train_mode="from_scratch", **kw).build_graph()
        ref, ndim = Ref2D("UNet", 16, 16, 8, 2, **kw), 2
        x = torch.from_numpy(rng.random((2, 16, 16, 2), dtype=np.float32))
        shape = (2, 16, 16, kw["output_nums"])
    y = _target(kind, shape, rng)
    # (a linear head produces values outside (0, 1) / negative ones: the clipped-probability branches and max(., eps) are exercised)
    _run(g, ref, x, [y], [kind], ndim, grad_rtol=2e-6, loss_rtol=2e-7)


def test_msle_is_what_the_shipped_config_compiles():
    """Train_Configs.ini:42 `loss_function = MeanSquaredLogarithmicError`: the class name, the Keras `name=` and the short form all
    resolve; an unknown loss names the supported ones"""
    from b2seg.model import _loss_name

    class MeanSquaredLogarithmicError:
        name = "mean_squared_logarithmic_error"
    assert _loss_name("MeanSquaredLogarithmicError") == _loss_name(MeanSquaredLogarithmicError()) == _loss_name("msle") == "msle"
    assert _loss_name("Huber") == "huber" and _loss_name("BinaryFocalCrossentropy") == "focal" and _loss_name("KLDivergence") == "kld"
    with pytest.raises(NotImplementedError, match="SparseCategoricalCrossentropy"):
        _loss_name("SparseCategoricalCrossentropy")
