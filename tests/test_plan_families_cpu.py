"""Planner vs oracle (float64, CPU emulator) for the decoder families and flags beyond plain UNet:
up-sampling decoders, attention gates, deep supervision, nested decoders (UNetE / UNet+ / UNet++), UNet3+, ConvLSTM skip
fusion (the 2D 'BCDUNet' = lstm=1), and their 1D counterparts incl. BCDUNet."""
import numpy as np
import pytest
import torch

from b2seg.models1d import BCDUNet, UNet
from b2seg.models2d import fpn_model_builder, model_selector, unet_model_builder
from oracle.ref_models import Ref1D, Ref2D, RefFPN
from test_plan_cpu import _run


def _targets(graph, N, rng, ndim):
    ts, losses = [], []
    for n in graph.outputs:
        H, W, C = n.shape
        shape = (N, H, W, C) if ndim == 2 else (N, W, C)
        final = n.attrs.get("activation") if n.name == "out" else (n.attrs.get("fn") if n.op == "act" else None)   # Self-ONN head: an Activation
        if final == "sigmoid":
            ts.append(torch.from_numpy((rng.random(shape) > 0.6).astype(np.float32))); losses.append("bce")
        elif final == "softmax":
            lab = rng.integers(0, C, shape[:-1])
            ts.append(torch.from_numpy(np.eye(C, dtype=np.float32)[lab])); losses.append("cce")
        else:
            ts.append(torch.from_numpy(rng.standard_normal(shape).astype(np.float32))); losses.append("mse")
    return ts, losses


CASES_2D = [
    ("UNet", dict(is_transconv=False)),
    ("UNet", dict(ds=1)),
    ("UNet", dict(ag=1)),
    ("UNet", dict(ag=1, ds=1, is_transconv=False, output_nums=3, final_activation="softmax")),
    ("UNet", dict(lstm=1, dense_loop=2)),
    ("UNetE", dict(ds=1)),
    ("UNetP", dict(ag=1)),
    ("UNetPP", dict(ds=1, ag=1, output_nums=4, final_activation="softmax")),
    ("UNetPP", dict(lstm=1)),
    ("UNet3P", dict(ds=1)),
    ("MultiResUNet", dict()),                      # odd channel counts (w/6, w/3, w/2): gapped concat layouts
    ("MultiResUNet", dict(ds=1, is_transconv=False)),
    ("UNet", dict(ae=1, feature_number=24)),       # Feature_Extraction_Block: Flatten -> Dense('features') -> Dense -> Reshape
    ("UNetPP", dict(ae=1, ds=1, feature_number=16)),
    ("UNet", dict(ds=1, output_nums=11, final_activation="softmax")),   # more than 8 classes: tensor-core 1x1 convolution + output activation op
    ("UNetPP", dict(output_nums=9, final_activation="sigmoid")),
    ("UNet", dict(ds=1, final_activation="tanh")),             # regression heads: convolution + named Activation output
    ("UNet3P", dict(output_nums=2, final_activation="relu")),
    ("MultiResUNet", dict(final_activation="LeakyReLU")),
    ("MultiResUNet", dict(ae=1, feature_number=16)),   # Flatten of an odd-channel (gapped, padded) tensor: the Dense kernel rows follow its physical layout
]


@pytest.mark.parametrize("dec,kw", CASES_2D, ids=[f"{d}-{'-'.join(f'{k}{v}' for k, v in kw.items())}" for d, kw in CASES_2D])
def test_2d_family(dec, kw):
    torch.manual_seed(0)
    rng = np.random.default_rng(1)
    depth = 2
    W = 16 if kw.get("lstm") else 8
    kw = dict(num_channels=2, **kw)
    g = unet_model_builder(dec, 16, 16, W, depth, train_mode="from_scratch", **kw).build_graph()
    x = torch.from_numpy(rng.random((2, 16, 16, 2), dtype=np.float32))
    ts, losses = _targets(g, 2, rng, 2)
    lw = [1.0 - 0.1 * i for i in range(len(ts))]
    # MultiResUNet builds a ResPath on the deepest encoder level that no decoder level consumes: Keras prunes it, the
    # eager oracle still evaluates it with weights of its own
    _run(g, Ref2D(dec, 16, 16, W, depth, **kw), x, ts, losses, 2, loss_weights=lw, strict=dec != "MultiResUNet")


CASES_2D_DEEP = [
    ("UNet4P", dict()),                             # UNet++ grid + dense sigmoid-pooled encoder links + anti-diagonal up-links (:379, :758-781)
    ("UNet4P", dict(ds=1, ag=1)),
    ("UNet4PV2", dict(ds=1)),                       # UNet3+ decoder on the dense-link encoder
    ("AHNet", dict(ds=1)),                          # UNet4P with ResPaths on every link (:523)
    ("MultiResUNet3P", dict(ds=1)),                 # :490
    ("KSSNet", dict(ds=1, ag=1)),                   # :603
    ("KSSNet", dict(is_transconv=False)),
]


@pytest.mark.parametrize("dec,kw", CASES_2D_DEEP, ids=[f"{d}-{'-'.join(f'{k}{v}' for k, v in kw.items())}" for d, kw in CASES_2D_DEEP])
def test_2d_family_depth3(dec, kw):
    """the remaining decoders of unet_variants.py at depth 3, where their extra links first appear"""
    rng = np.random.default_rng(4)
    kw = dict(num_channels=2, **kw)
    g = unet_model_builder(dec, 32, 32, 8, 3, train_mode="from_scratch", **kw).build_graph()
    x = torch.from_numpy(rng.random((2, 32, 32, 2), dtype=np.float32))
    ts, losses = _targets(g, 2, rng, 2)
    strict = dec not in ("MultiResUNet3P", "KSSNet", "AHNet")   # dangling ResPath branches that Keras prunes (see test_2d_family)
    # descriptors carry eps / momentum as float32 (relative 6e-8): through these 60-90 layer graphs a pre-activation within 1e-7 of
    # zero can land on the other side of the ReLU, hence 2e-7 instead of 1e-8 on the activations
    _run(g, Ref2D(dec, 32, 32, 8, 3, **kw), x, ts, losses, 2, loss_weights=[1.0 - 0.1 * i for i in range(len(ts))], strict=strict, act_atol=2e-7)


CASES_SELF = [
    ("SelfUNet", dict()),                                              # :644 operational layers, transposed operational up-sampling
    ("SelfUNet", dict(ds=1, is_transconv=False, q=4)),                 # two sum passes: (c1 + c2 + c3), (+ c4)
    ("SelfUNet", dict(ds=1, output_nums=3, final_activation="softmax", dense_loop=2)),
    ("SelfUNetPP", dict(ds=1)),                                        # :667
    ("SelfUNetPP", dict(is_transconv=False, q=2)),
    ("SelfUNet3P", dict(ds=1)),                                        # :713 (stride-2 operational deep-supervision heads)
    ("SelfUNet", dict(ae=1, feature_number=16, q=1)),                  # q = 1: an operational layer is a plain convolution
]


@pytest.mark.parametrize("dec,kw", CASES_SELF, ids=[f"{d}-{'-'.join(f'{k}{v}' for k, v in kw.items())}" for d, kw in CASES_SELF])
def test_2d_self_onn_family(dec, kw):
    """Self-ONN decoders (unet_variants.py:59-64, 644-747, 782-786, 1107-1108; onn_layers.py:6-48)"""
    rng = np.random.default_rng(5)
    depth = 3 if dec == "SelfUNetPP" else 2     # (deeper un-normalised cubic encoders make the Adam check ill-conditioned)
    S = 8 * 2 ** depth // 2
    kw = dict(num_channels=2, **kw)
    g = unet_model_builder(dec, S, S, 8, depth, train_mode="from_scratch", **kw).build_graph()
    # the Self-ONN encoder is linear and un-normalised (:782-786): powers of powers overflow quickly, keep the input small
    x = torch.from_numpy(0.5 * (rng.random((2, S, S, 2), dtype=np.float32) - 0.3))
    ts, losses = _targets(g, 2, rng, 2)
    # SelfUNet3P feeds un-normalised cubic features to the sigmoid head (:741, :1108): at random init the logits saturate, where the
    # REPORTED loss scalar (probabilities clipped to [1e-7, 1 - 1e-7] by b2seg_loss) departs from Keras' from-logits value; the
    # gradient seed (p - y) / n and every tensor below are still held to the tight tolerances
    _run(g, Ref2D(dec, S, S, 8, depth, **kw), x, ts, losses, 2, loss_weights=[1.0 - 0.1 * i for i in range(len(ts))], act_atol=2e-7,
         loss_rtol=1e-2 if dec == "SelfUNet3P" else 1e-7)


CASES_FPN = [dict(), dict(ds=1), dict(ag=1, ds=1, output_nums=3, final_activation="softmax"), dict(lstm=1), dict(ae=1, feature_number=16)]


@pytest.mark.parametrize("kw", CASES_FPN, ids=["-".join(f"{k}{v}" for k, v in kw.items()) or "plain" for kw in CASES_FPN])
def test_fpn_family(kw):
    """FPN genre (fpn_variants.py:132-169): add-merge decoder + multi-scale bilinear concat head, through the same kernels"""
    rng = np.random.default_rng(3)
    depth = 3
    W = 16 if kw.get("lstm") else 8
    kw = dict(num_channels=2, **kw)
    g = fpn_model_builder("FPN", 32, 32, W, depth, train_mode="from_scratch", **kw).build_graph()
    x = torch.from_numpy(rng.random((2, 32, 32, 2), dtype=np.float32))
    ts, losses = _targets(g, 2, rng, 2)
    _run(g, RefFPN("FPN", 32, 32, W, depth, **kw), x, ts, losses, 2, loss_weights=[1.0 - 0.1 * i for i in range(len(ts))])


def test_fpn_builder_api():
    with pytest.raises(ValueError):
        fpn_model_builder("FPN", 32, 32, 8, 2, train_mode="nope")
    with pytest.raises(ValueError):   # Keras Add() of a bilinearly up-sampled tensor and a skip with half its channels
        fpn_model_builder("FPN", 32, 32, 8, 2, is_transconv=False, train_mode="from_scratch").build_graph()
    g = fpn_model_builder("FPN", 64, 64, 16, 3, num_channels=3, ds=1, train_mode="from_scratch").build_graph("VGG16")
    assert g.name == "VGG16_FPN" and [n.name for n in g.outputs] == ["out", "level1", "level2", "level3"]
    assert g.outputs[0].inputs[0].C == 16 * (4 + 2 + 1)          # the head reads the concat of all decoder levels
    m = model_selector("FPN", "resnet50", "FPN", 64, 64, 16, 3, train_mode="from_scratch").segmentation_model()   # model_selector.py:717
    assert m.name == "ResNet50_FPN" and m.output_names == ["out"]
    assert model_selector("fpn", "chexnet", "FPN", 64, 64, 16, 3, train_mode="from_scratch").segmentation_model().name == "DenseNet121(CheXNet)_FPN"


CASES_1D = [
    ("UNet", dict(ds=1, ag=1, is_transconv=False)),
    ("UNet", dict(ds=0, lstm=1)),
    ("UNetPP", dict(ds=1, ag=1)),
    ("UNet3P", dict(ds=1)),
    ("BCDUNet", dict(ds=1, lstm=1, dense_loop=2)),
    ("BCDUNet", dict(ds=0, lstm=0, ag=1)),
    ("MultiResUNet", dict(ds=0)),
    ("MultiResUNet", dict(ds=1, ag=1)),
    ("UNet", dict(ae=1, ds=0, feature_number=24)),
    ("BCDUNet", dict(ae=1, ds=1, lstm=1, feature_number=16)),
    ("RUNet", dict(ds=1, t=2)),                    # Recurrent_Conv_Block (uv.py:63-72): t rounds of conv + concat with the block input
    ("R2UNet", dict(ds=0, ag=1, t=1)),             # + 1x1 Conv_Block shortcut and Add around each pair
    ("R2UNet", dict(ds=1, lstm=1, t=2)),
    ("R2UNetPP", dict(ds=1, ag=1, t=1)),           # UNet++ grid of shortcut + ONE recurrent block nodes (:1119)
    ("R2UNet3P", dict(ds=1, t=1)),                 # :1226; the m-loop drops its first recurrent block (names still consumed)
    ("MultiResUNet3P", dict(ds=1, ag=1)),          # :899, with its overwritten dense links and its dead bottleneck block
    ("MultiResUNet3P", dict(ds=0, is_transconv=False)),
    ("R2UNet", dict(ds=1, ae=1, feature_number=16, t=2)),   # the bottleneck's recurrent blocks concatenate the Feature_Extraction_Block's Reshape output
    ("RUNet", dict(ds=0, ae=1, feature_number=16, t=1)),
    ("UNet", dict(ds=1, problem_type="Classification", output_nums=12)),   # 12-class softmax head
    ("MultiResUNet", dict(ds=1, ae=1, feature_number=16)),                 # Flatten of a pooled 60-channel tensor stored in 64 lanes
    ("BCDUNet", dict(ae=1, ds=0, lstm=1, dense_loop=2, feature_number=16)),   # Flatten of a concatenation (concat-dense block, BCDUNet.py:70-76)
]


@pytest.mark.parametrize("var,kw", CASES_1D, ids=[f"{d}-{'-'.join(f'{k}{v}' for k, v in kw.items())}" for d, kw in CASES_1D])
def test_1d_family(var, kw):
    rng = np.random.default_rng(2)
    L_, depth, ch, ks = 32, 2, 2, 3
    W = 16 if kw.get("lstm") else 8
    if var == "BCDUNet":
        g = BCDUNet(L_, depth, ch, W, ks, **kw).BCDUNet().graph
    else:
        g = getattr(UNet(L_, depth, ch, W, ks, **kw), var)().graph
    kw = {k_: v for k_, v in kw.items() if k_ != "t" or var in ("RUNet", "R2UNet", "R2UNetPP", "R2UNet3P")}
    x = torch.from_numpy(rng.standard_normal((2, L_, ch)).astype(np.float32))
    ts, losses = _targets(g, 2, rng, 1)
    # with lstm=0 the 1D BCDUNet drops its skip connections, so an attention gate built on them is a dangling branch that
    # Keras prunes: the oracle (eager) still evaluates it with weights of its own
    _run(g, Ref1D(var, L_, depth, ch, W, ks, **kw), x, ts, losses, 1, strict=not ((var == "BCDUNet" and not kw.get("lstm")) or var in ("MultiResUNet", "R2UNet3P", "MultiResUNet3P")))


CASES_1D_SELF = [
    ("SelfUNetPP", dict(ds=1)),                        # uv.py:1412: Oper1D pairs + Oper1DTranspose(kernel 4, tanh)
    ("SelfUNetPP", dict(ds=0, ag=1, is_transconv=False, q=2)),
    ("SelfUNetPP", dict(ds=1, lstm=1, q=2, ae=1, feature_number=16)),
    ("SelfR2UNetPP", dict(ds=1, t=2)),                 # :1312: Self_Recurrent_Conv_Block encoder (bottom level with q=1), single Oper1D nodes
    ("SelfR2UNetPP", dict(ds=0, t=1, ag=1)),
    ("SelfR2UNetPP", dict(ds=1, t=2, ae=1, feature_number=16)),
    ("SelfUNet3P", dict(ds=1)),                        # :1515
]


@pytest.mark.parametrize("var,kw", CASES_1D_SELF, ids=[f"{d}-{'-'.join(f'{k}{v}' for k, v in kw.items())}" for d, kw in CASES_1D_SELF])
def test_1d_self_onn_family(var, kw):
    """1D Self-ONN variants (1DCNN/Models/unet_variants.py:75-84, 1312-1583; ONN_layers.py:7-52)"""
    rng = np.random.default_rng(6)
    L_, depth, ch, ks = 32, 2, 2, 3
    W = 16 if kw.get("lstm") else 8
    g = getattr(UNet(L_, depth, ch, W, ks, **kw), var)().graph
    # two un-normalised operational layers per level: powers of powers (3^12 at depth 2) overflow for |x| > 1 — keep the signal small
    x = torch.from_numpy((0.3 * (rng.random((2, L_, ch)) - 0.5)).astype(np.float32))
    ts, losses = _targets(g, 2, rng, 1)
    # Adam's first step is lr * g / (|g| + eps): an element 1e-6 below its tensor's largest gradient (these cubic networks spread
    # gradients over six decades) turns the float32 descriptor rounding the gradient check allows into a visible step difference
    # ... and activations reach 1e4 (cubes of cubes), so the float32-descriptor noise (relative 5e-8, amplified 3x per cube) is
    # judged relative to the value
    _run(g, Ref1D(var, L_, depth, ch, W, ks, **kw), x, ts, losses, 1, act_atol=2e-7, act_rtol=2e-6, grad_rtol=1e-5, adam_atol=1e-4, loss_rtol=1e-6)


@pytest.mark.parametrize("kw", [dict(ds=1), dict(ds=1, ag=1, is_transconv=False)], ids=["ds1", "ds1-ag1-upsample"])
def test_1d_unet4p_depth3(kw):
    """1D UNet4+ (uv.py:717-834) at depth 3, where its dense encoder links (levels 1 .. i-1) and anti-diagonal up-links first appear"""
    rng = np.random.default_rng(5)
    g = UNet(64, 3, 2, 8, 3, **kw).UNet4P().graph
    x = torch.from_numpy(rng.standard_normal((2, 64, 2)).astype(np.float32))
    ts, losses = _targets(g, 2, rng, 1)
    _run(g, Ref1D("UNet4P", 64, 3, 2, 8, 3, **kw), x, ts, losses, 1, act_atol=2e-7)


EXCHANGE_CASES = [
    ("UNetPP", dict(ds=1, ag=1, output_nums=4, final_activation="softmax")),   # BASELINE config 3 family
    ("MultiResUNet", dict()),                                                  # config 4 family
    ("UNet", dict(lstm=1, dense_loop=2)),                                      # config 5 family
    ("UNet", dict(ae=1, feature_number=16, ds=1)),
    ("UNet3P", dict(ds=1)),
    ("SelfUNet", dict(ds=1)),
    ("UNet", dict(output_nums=10, final_activation="softmax")),
    ("KSSNet", dict(ag=1)),
]


@pytest.mark.parametrize("dec,kw", EXCHANGE_CASES, ids=[f"{d}-{'-'.join(f'{k}{v}' for k, v in kw.items())}" for d, kw in EXCHANGE_CASES])
def test_exchange_schedule_slices_are_final_when_released(dec, kw):
    """Data parallel: Model._step all-reduces slice [lo, hi) of the gradient arena right after backward op n_ops.  For
    every family: the slices tile the arena, and each one already holds its FINAL value at that point (no later op adds to it) —
    checked by replaying the backward phase in the schedule's ranges on the emulator and comparing with the finished gradient."""
    from b2seg.graph import init_params
    from b2seg.planner import Planner
    from desc_emulator import PlanMem, run_phase
    rng = np.random.default_rng(8)
    W = 16 if kw.get("lstm") else 8
    kw = dict(num_channels=2, **kw)
    g = unet_model_builder(dec, 16, 16, W, 2, train_mode="from_scratch", **kw).build_graph()
    ts, losses = _targets(g, 2, rng, 2)
    mem = PlanMem()
    pl = Planner(g, 2, mem.alloc_bytes, training=True, losses=losses, adam=dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7),
                 adam_bucket_bytes=16384).build()
    params = init_params(g, seed=5)
    for e in pl.params:
        flat = torch.from_numpy(pl.to_internal(e.key, params[e.key])).double()
        if e.trainable:
            mem.f32(pl.w_ptr + 4 * e.offset, e.size)[:] = flat
            wb, off = mem.resolve(pl.wb_ptr + 2 * e.offset)
            wb[off:off + e.size] = flat
        else:
            mem.f32(pl.mov_ptr + 4 * e.offset, e.size)[:] = flat
    x = 0.5 * rng.random((2, 16, 16, 2))
    mem.f32(pl.input_ptr, x.size)[:] = torch.from_numpy(x.reshape(-1))
    for o in pl.outputs:
        mem.f32(o["target_ptr"], ts[o["index"]].numel())[:] = ts[o["index"]].reshape(-1).double()
    run_phase(mem, pl, 0)
    n = pl.arena_elems
    grads = mem.f32(pl.g_ptr, n)
    sched = pl.exchange_schedule(bucket_bytes=16384)
    covered = sorted((lo, hi) for (_, lo, hi) in sched)
    assert covered[0][0] == 0 and covered[-1][1] == n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    assert len(sched) >= 3 and sched[0][0] < len(pl.ops[1])
    assert len(pl.ops[2]) == len(sched)                       # one Adam op per bucket, over the same slices
    for (op, d, _note), (_n_ops, lo, hi) in zip(pl.ops[2], sched):
        assert (d.n, d.g) == (hi - lo, pl.g_ptr + 4 * lo)
    done, released = 0, []
    for (n_ops, lo, hi) in sched:
        run_phase(mem, pl, 1, done, n_ops - done)
        done = n_ops
        released.append((lo, hi, grads[lo:hi].clone()))
    run_phase(mem, pl, 1, done, len(pl.ops[1]) - done)
    assert float(grads.abs().max()) > 0
    for lo, hi, snap in released:
        assert torch.equal(snap, grads[lo:hi]), (dec, lo, hi, float((snap - grads[lo:hi]).abs().max()))
