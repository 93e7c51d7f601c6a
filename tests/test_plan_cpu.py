"""CPU check of the planner: run the emitted descriptor program on the float64 descriptor emulator and compare
outputs, per-layer activations, gradients, Adam-updated weights and BN moving statistics with the oracle
(oracle/ref_models.py, float64) on identical weights and inputs.  Tolerance 1e-8: both sides are float64, so any
mismatch is a planner bug (fusion, concat slots, gradient routing, weight layout), not rounding."""
import numpy as np
import pytest
import torch

from b2seg.graph import init_params
from b2seg.models1d import BCDUNet, UNet
from b2seg.models2d import unet_model_builder
from b2seg.planner import Planner
from desc_emulator import PlanMem, run_phase
from oracle.keras_ref import KerasRef, keras_adam_step, keras_loss
from oracle.ref_models import Ref1D, Ref2D


def _run(graph, ref, x, targets, losses, ndim, lr=1e-2, loss_weights=None, strict=True, act_atol=1e-8, loss_rtol=1e-7, adam_atol=2e-6, act_rtol=0.0, grad_rtol=1e-6):
    N = x.shape[0]
    params = init_params(graph, seed=7)
    # make BN affine and biases non-trivial so their handling is actually tested
    rng = np.random.default_rng(0)
    for k, v in params.items():
        if k.endswith("/gamma"):
            params[k] = (1 + 0.2 * rng.standard_normal(v.shape)).astype(np.float32)
        elif k.endswith("/beta") or k.endswith("/bias"):
            params[k] = (0.1 * rng.standard_normal(v.shape)).astype(np.float32)
    mem = PlanMem()
    pl = Planner(graph, N, mem.alloc_bytes, training=True, losses=losses, loss_weights=loss_weights,
                 adam=dict(lr=lr, beta1=0.9, beta2=0.999, eps=1e-7)).build()
    for e in pl.params:
        flat = torch.from_numpy(pl.to_internal(e.key, params[e.key])).double()
        if e.trainable:
            mem.f32(pl.w_ptr + 4 * e.offset, e.size)[:] = flat
            wb, off = mem.resolve(pl.wb_ptr + 2 * e.offset)
            wb[off:off + e.size] = flat
        else:
            mem.f32(pl.mov_ptr + 4 * e.offset, e.size)[:] = flat
    xin = x if ndim == 2 else x.unsqueeze(1)
    mem.f32(pl.input_ptr, xin.numel())[:] = xin.reshape(-1).double()
    for o in pl.outputs:
        mem.f32(o["target_ptr"], targets[o["index"]].numel())[:] = targets[o["index"]].reshape(-1).double()
    run_phase(mem, pl, 0)
    run_phase(mem, pl, 1)

    # ---- oracle
    tp = {k: torch.from_numpy(v.copy()).double() for k, v in params.items()}
    k = KerasRef(ndim, params=tp, dtype=torch.float64, training=True, strict=strict)
    outs = ref(k, x.double())
    total = 0
    for i, (o, t) in enumerate(zip(outs, targets)):
        name = graph.outputs[i].name
        li = keras_loss(losses[i], o, t.double(), logits=k.logits.get(name))
        total = total + (loss_weights[i] if loss_weights else 1.0) * li
    total.backward()

    # outputs
    for o in pl.outputs:
        got = mem.f32(o["ptr"], int(np.prod(o["shape"]))).view(o["shape"])
        want = outs[o["index"]].detach()
        if ndim == 1:
            got = got.squeeze(1)
        assert torch.allclose(got, want, atol=1e-8 + act_rtol * float(want.abs().max())), (o["name"], float((got - want).abs().max()))
    assert abs(float(mem.f32(pl.loss_ptr, 1)) - float(total.detach())) < loss_rtol * max(1.0, abs(float(total.detach())))
    # per-layer activations
    checked = 0
    for name, (view, C, kind) in pl.taps.items():
        if name not in k.acts or kind == "post":
            continue
        if kind == "concat":
            continue
        got = mem.gather_view(view.to_c())[..., pl.logical_channels(name)]
        want = k.acts[name].detach()
        if ndim == 1:
            got = got.squeeze(1)
        got = got.reshape(want.shape)        # (a Dense output is (N, 1, 1, units) in the plan, (N, units) in the oracle)
        # act_rtol is relative to the tensor's largest value (sums of large cancelling terms in the un-normalised Self-ONN graphs)
        assert torch.allclose(got, want, atol=act_atol + act_rtol * float(want.abs().max())), (name, float((got - want).abs().max()))
        checked += 1
    assert checked > 3
    # activation gradients (w.r.t. raw conv outputs)
    for name, (view, C) in pl.grad_taps.items():
        if name in k.acts and k.acts[name].grad is not None and pl.taps.get(name, (0, 0, ""))[2] == "raw":
            got = mem.gather_view(view.to_c())[..., pl.logical_channels(name)]
            if ndim == 1:
                got = got.squeeze(1)
            got = got.reshape(k.acts[name].grad.shape)
            # descriptors carry eps / loss weights as float32 (relative 5e-8): tolerance relative to the gradient's scale
            gtol = 1e-9 + 5e-7 * float(k.acts[name].grad.abs().max())
            assert torch.allclose(got, k.acts[name].grad, atol=gtol), ("grad", name, float((got - k.acts[name].grad).abs().max()))
    # parameter gradients
    bn_convs = {u["node"].name for u in pl.units if u["kind"] == "conv" and u["bn"] is not None}
    gmax = max([float(v.grad.abs().max()) for v in tp.values() if v.grad is not None] + [1.0])   # "analytically zero" is relative to this
    for e in pl.params:
        if not e.trainable:
            continue
        got = torch.from_numpy(pl.from_internal(e.key, mem.f32(pl.g_ptr + 4 * e.offset, e.size).numpy().astype(np.float32))).double()
        got64 = mem.f32(pl.g_ptr + 4 * e.offset, e.size)
        want = tp[e.key].grad
        if want is None:
            assert float(got64.abs().max()) == 0, e.key
            continue
        gi = torch.from_numpy(pl.to_internal(e.key, want.numpy().astype(np.float32))).double()  # layout round trip of the oracle grad
        if e.key.endswith("/bias") and e.key.rsplit("/", 1)[0] in [u["node"].name for u in pl.units if u["kind"] == "conv" and u["bn"] is not None]:
            # conv bias followed by BN: gradient is analytically zero; the product emits exact zeros
            assert float(got64.abs().max()) == 0 and float(want.abs().max()) < 1e-9 * gmax, e.key
            continue
        scale = float(want.abs().max()) + 1e-12
        assert float((got - want).abs().max()) < grad_rtol * scale + 1e-9, (e.key, float((got - want).abs().max()), scale)
        assert got.shape == want.shape and gi.numel() == e.size
    # Adam + moving statistics
    run_phase(mem, pl, 2)
    for e in pl.params:
        if e.trainable:
            w = tp[e.key].detach().clone()
            g = tp[e.key].grad if tp[e.key].grad is not None else torch.zeros_like(w)
            if e.key.endswith("/bias") and e.key.rsplit("/", 1)[0] in bn_convs and float(g.abs().max()) < 1e-9 * gmax:
                g = torch.zeros_like(w)
            keras_adam_step(w, g, torch.zeros_like(w), torch.zeros_like(w), 1, lr=lr)
            got = torch.from_numpy(pl.from_internal(e.key, mem.f32(pl.w_ptr + 4 * e.offset, e.size).numpy().astype(np.float32))).double()
            assert torch.allclose(got, w, atol=adam_atol), (e.key, float((got - w).abs().max()))
        else:
            got = torch.from_numpy(pl.from_internal(e.key, mem.f32(pl.mov_ptr + 4 * e.offset, e.size).numpy().astype(np.float32))).double()
            assert torch.allclose(got, k.new_moving[e.key], atol=1e-6), e.key
    return pl


def test_plan_unet2d_matches_oracle():
    torch.manual_seed(0)
    kw = dict(num_channels=3, output_nums=1, dense_loop=1, is_transconv=True)
    g = unet_model_builder("UNet", 16, 16, 8, 2, train_mode="from_scratch", **kw).build_graph()
    x = torch.rand(2, 16, 16, 3)
    y = (torch.rand(2, 16, 16, 1) > 0.6).float()
    pl = _run(g, Ref2D("UNet", 16, 16, 8, 2, **kw), x, [y], ["bce"], 2)
    kinds = [op for (op, _, _) in pl.ops[0]]
    assert kinds.count(11) == 1  # one input cast; concat realised without copy ops
    assert not any(note == "concat copy" for (_, _, note) in pl.ops[0])
    # the `out` head's backward (input gradient, dW, db) is folded into batch_normalization_4's backward: no OP_HEAD_BWD (= 8) left
    assert 8 not in [op for (op, _, _) in pl.ops[1]]
    # transposed-conv bias gradients come from the consumer dgrad's epilogue statistics (OP_ROWSUM = 22), not from a pass over dY
    assert [note for (op, _, note) in pl.ops[1] if op == 22] == ["bias grad conv2d_transpose_1", "bias grad conv2d_transpose"]


def test_plan_unet2d_depth3_dense2_mse():
    torch.manual_seed(1)
    kw = dict(num_channels=1, output_nums=2, dense_loop=2, is_transconv=True, final_activation="linear")
    g = unet_model_builder("UNet", 16, 24, 8, 3, train_mode="from_scratch", **kw).build_graph()
    x = torch.rand(2, 16, 24, 1)
    y = torch.randn(2, 16, 24, 2)
    _run(g, Ref2D("UNet", 16, 24, 8, 3, **kw), x, [y], ["mse"], 2)


def test_plan_unet1d_matches_oracle():
    torch.manual_seed(2)
    m = UNet(64, 2, 1, 8, 3, problem_type="Classification", output_nums=2, ds=0, is_transconv=True).UNet()
    x = torch.randn(3, 64, 1)
    y = torch.nn.functional.one_hot(torch.randint(0, 2, (3, 64)), 2).float()
    _run(m.graph, Ref1D("UNet", 64, 2, 1, 8, 3, problem_type="Classification", output_nums=2, ds=0), x, [y], ["cce"], 1)


def test_plan_unet1d_regression_k5():
    torch.manual_seed(3)
    m = UNet(32, 2, 2, 8, 5, problem_type="Regression", output_nums=1, ds=0, is_transconv=True).UNet()
    x = torch.randn(2, 32, 2)
    y = torch.randn(2, 32, 1)
    _run(m.graph, Ref1D("UNet", 32, 2, 2, 8, 5, problem_type="Regression", output_nums=1, ds=0), x, [y], ["mae"], 1)


def test_unsupported_loss_head_pairs_are_refused():
    """b2seg_loss seeds the backward pass with dL/dlogits; pairs it cannot form that for must fail at plan time, not train wrongly"""
    from b2seg.planner import PlanError
    g = unet_model_builder("UNet", 16, 16, 8, 2, ds=1, train_mode="from_scratch").build_graph()       # out: sigmoid, levels: linear
    with pytest.raises(PlanError, match="is not lowered"):      # categorical cross-entropy on the sigmoid head
        Planner(g, 2, PlanMem().alloc_bytes, training=True, losses=["cce", "mse", "mse"], adam=dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7)).build()
    # cross-entropies on the LINEAR deep-supervision heads are what Keras accepts (clipped probabilities), and so does the planner
    for losses in (["bce", "bce", "mse"], ["mse", "cce", "mse"], ["mse", "mae", "mse"], ["msle", "huber", "logcosh"]):
        Planner(g, 2, PlanMem().alloc_bytes, training=True, losses=losses, adam=dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7)).build()
    # a tanh head is a convolution + an Activation output, i.e. a "linear" output for the loss (the Activation has its own backward):
    # tanh caches no logits in Keras, so even a cross-entropy is the clipped-probability form
    g = unet_model_builder("UNet", 16, 16, 8, 2, final_activation="tanh", train_mode="from_scratch").build_graph()
    for loss in ("mse", "bce"):
        Planner(g, 2, PlanMem().alloc_bytes, training=True, losses=[loss], adam=dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7)).build()


def test_activation_identifiers_resolve_like_keras2():
    """Activation('ReLU') / ('LeakyReLU') / final_activation='Softmax' name advanced-activation LAYERS; 'Sigmoid' and 'Linear' name nothing"""
    g = unet_model_builder("UNet", 16, 16, 8, 2, output_nums=3, final_activation="Softmax", train_mode="from_scratch").build_graph()
    assert g.outputs[0].attrs["activation"] == "softmax"
    for bad in ("Sigmoid", "Linear", "softMax"):
        with pytest.raises(ValueError, match="Unknown activation function"):
            unet_model_builder("UNet", 16, 16, 8, 2, final_activation=bad, train_mode="from_scratch").build_graph()
    k = KerasRef(2, dtype=torch.float64)
    x = torch.randn(2, 3, 3, 4, dtype=torch.float64)
    assert torch.equal(k.activation_fn("Softmax", x), torch.softmax(x, -1))
    with pytest.raises(ValueError, match="Unknown activation function"):
        k.activation_fn("Sigmoid", x)


def test_zero_filter_convolutions_fail_like_keras():
    """MultiResBlock's int(alpha * W * 0.167) is 0 for narrow models: Keras refuses the layer, and so does the builder"""
    with pytest.raises(ValueError, match="strictly positive"):
        unet_model_builder("MultiResUNet", 16, 16, 4, 2, train_mode="from_scratch").build_graph()
    with pytest.raises(ValueError, match="strictly positive"):
        UNet(32, 2, 1, 4, 3).MultiResUNet()
