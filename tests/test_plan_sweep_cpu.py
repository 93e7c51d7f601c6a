"""Coverage sweep (CPU): every builder method x {ds, ag, lstm, ae, is_transconv} combination is pushed through the planner at a small
size.  A combination must either lower completely, or fail with the error the REFERENCE itself raises for it (SURVEY §9 item 17:
`lstm=1` needs `is_transconv=True` because of the Reshape before the ConvLSTM; the MultiResUNet / KSSNet / 1D MultiResUNet3P `lstm=1`
branches use undefined names) — never with a PlanError, i.e. there is no valid reference configuration the product refuses.
The arithmetic of the families is checked elsewhere (tests/test_plan_families_cpu.py); this only proves coverage."""
import itertools

import pytest

from b2seg.models1d import BCDUNet, UNet
from b2seg.models2d import IN_SCOPE_DECODERS, fpn_model_builder, unet_model_builder
from b2seg.planner import Planner
from desc_emulator import PlanMem

ADAM = dict(lr=2e-4, beta1=0.9, beta2=0.999, eps=1e-7)
FLAGS = list(itertools.product((0, 1), (0, 1), (0, 1), (0, 1), (True, False)))   # ds, ag, lstm, ae, is_transconv
VARIANTS_1D = ["UNet", "UNetE", "UNetP", "UNetPP", "UNet3P", "UNet4P", "MultiResUNet", "MultiResUNet3P", "RUNet", "R2UNet", "R2UNetPP",
               "R2UNet3P", "SelfUNetPP", "SelfR2UNetPP", "SelfUNet3P", "BCDUNet"]


def _plan(graph):
    mem = PlanMem()
    Planner(graph, 2, mem.alloc_bytes, training=True, losses=["mse"] * len(graph.outputs), adam=ADAM).build()


def _reference_error(exc, lstm, tc):
    """the failures the reference has for these flags"""
    if isinstance(exc, NameError):
        return lstm == 1                                   # 2D MultiResUNet :477 / KSSNet, 1D MultiResUNet3P: undefined names in the lstm branch
    if isinstance(exc, ValueError):
        msg = str(exc)
        # Reshape((1, H, W, C)) of an up-sampled (not transposed-conv) tensor with twice the channels: Keras' Reshape error;
        # FPN's add-merge of an up-sampled tensor with twice the channels: Keras' Add error
        return lstm == 1 and not tc and ("total size of new array must be unchanged" in msg) or \
            (not tc and "Inputs have incompatible shapes" in msg)
    return False


@pytest.mark.parametrize("dec", list(IN_SCOPE_DECODERS) + ["FPN"])
def test_every_2d_flag_combination_lowers_or_fails_like_the_reference(dec):
    lowered = 0
    for ds, ag, lstm, ae, tc in FLAGS:
        kw = dict(ds=ds, ag=ag, lstm=lstm, ae=ae, feature_number=16, is_transconv=tc, train_mode="from_scratch")
        try:
            b = fpn_model_builder("FPN", 32, 32, 16, 3, **kw) if dec == "FPN" else unet_model_builder(dec, 32, 32, 16, 3, **kw)
            _plan(b.build_graph())
            lowered += 1
        except Exception as e:  # noqa: BLE001
            assert _reference_error(e, lstm, tc), (dec, kw, type(e).__name__, str(e)[:200])
    assert lowered >= 16, (dec, lowered)


@pytest.mark.parametrize("var", VARIANTS_1D)
def test_every_1d_flag_combination_lowers_or_fails_like_the_reference(var):
    lowered = 0
    for ds, ag, lstm, ae, tc in FLAGS:
        kw = dict(ds=ds, ag=ag, lstm=lstm, ae=ae, feature_number=16, is_transconv=tc)
        try:
            g = BCDUNet(64, 3, 2, 16, 3, dense_loop=2, **kw).BCDUNet().graph if var == "BCDUNet" else getattr(UNet(64, 3, 2, 16, 3, **kw), var)().graph
            _plan(g)
            lowered += 1
        except Exception as e:  # noqa: BLE001
            assert _reference_error(e, lstm, tc), (var, kw, type(e).__name__, str(e)[:200])
    assert lowered >= 16, (var, lowered)
