"""Pins the oracle (oracle/keras_ref.py, PyTorch eager) against INDEPENDENT restatements of the Keras-2 layer arithmetic written as
plain NumPy loops straight from the Keras / TensorFlow documentation formulas — no torch.nn.functional on the checking side.
The reference ships no tests or golden vectors and TensorFlow cannot run here (parity unpinned by the reference), so
this is the strongest pin available: two independent implementations of the same reading of Keras-2 must agree to float64
round-off, and the committed fixtures under tests/golden/ (written by tests/golden/make_golden.py from the float64 oracle) must
keep reproducing, so a silent change of the oracle's semantics cannot go unnoticed."""
import math
import os

import numpy as np
import pytest
import torch

from oracle.keras_ref import KerasRef, keras_adam_step, keras_loss

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
rng = np.random.default_rng(42)


def _conv_same_np(x, w, b, sh=1, sw=1):
    """tf.nn.conv2d 'SAME': out = ceil(in / s); total pad = max((out-1)*s + k - in, 0), the smaller half first (TF docs)."""
    N, H, W, Ci = x.shape
    kh, kw, _, Co = w.shape
    Ho, Wo = -(-H // sh), -(-W // sw)
    th, tw = max((Ho - 1) * sh + kh - H, 0), max((Wo - 1) * sw + kw - W, 0)
    pt, pl = th // 2, tw // 2
    y = np.zeros((N, Ho, Wo, Co))
    for n in range(N):
        for i in range(Ho):
            for j in range(Wo):
                for a in range(kh):
                    for c in range(kw):
                        hi, wi = i * sh + a - pt, j * sw + c - pl
                        if 0 <= hi < H and 0 <= wi < W:
                            y[n, i, j] += x[n, hi, wi] @ w[a, c]
    return y + b


@pytest.mark.parametrize("kh,kw,sh,sw", [(3, 3, 1, 1), (2, 2, 1, 1), (4, 4, 1, 1), (1, 5, 1, 1), (1, 1, 2, 2), (3, 3, 2, 2)])
def test_conv_same_padding_and_strides(kh, kw, sh, sw):
    x = rng.standard_normal((2, 6, 7, 3))
    w = rng.standard_normal((kh, kw, 3, 4))
    b = rng.standard_normal(4)
    k = KerasRef(2, params={"conv2d/kernel": torch.from_numpy(w), "conv2d/bias": torch.from_numpy(b)}, dtype=torch.float64, strict=True)
    y = k.Conv(torch.from_numpy(x), 4, (kh, kw), strides=(sh, sw), padding="same")
    assert np.allclose(y.detach().numpy(), _conv_same_np(x, w, b, sh, sw), atol=1e-12)


def test_conv1d_same_even_kernel():
    x = rng.standard_normal((2, 9, 2))
    w = rng.standard_normal((4, 2, 3))
    k = KerasRef(1, params={"conv1d/kernel": torch.from_numpy(w), "conv1d/bias": torch.zeros(3, dtype=torch.float64)}, dtype=torch.float64, strict=True)
    y = k.Conv(torch.from_numpy(x), 3, 4, padding="same")
    want = _conv_same_np(x[:, None], w[None], np.zeros(3))[:, 0]
    assert np.allclose(y.detach().numpy(), want, atol=1e-12)


def _tconv_np(x, w, b, s):
    """Conv2DTranspose 'same', stride s, kernel (kh,kw,Cout,Cin): scatter form.  Full output size (in-1)*s + k, of which Keras keeps
    in*s rows starting at (k - s) // 2 (deconv_output_length + the gradient-of-conv definition)."""
    N, H, W, Ci = x.shape
    kh, kw, Co, _ = w.shape
    full = np.zeros((N, (H - 1) * s + kh, (W - 1) * s + kw, Co))
    for n in range(N):
        for i in range(H):
            for j in range(W):
                for a in range(kh):
                    for c in range(kw):
                        full[n, i * s + a, j * s + c] += w[a, c] @ x[n, i, j]
    ch, cw = (kh - s) // 2, (kw - s) // 2
    return full[:, ch:ch + H * s, cw:cw + W * s] + b


def test_conv2d_transpose_4x4_s2():
    x = rng.standard_normal((2, 3, 4, 3))
    w = rng.standard_normal((4, 4, 5, 3))
    b = rng.standard_normal(5)
    k = KerasRef(2, params={"conv2d_transpose/kernel": torch.from_numpy(w), "conv2d_transpose/bias": torch.from_numpy(b)}, dtype=torch.float64, strict=True)
    y = k.ConvTranspose(torch.from_numpy(x), 5, (4, 4), (2, 2))
    assert y.shape == (2, 6, 8, 5) and np.allclose(y.detach().numpy(), _tconv_np(x, w, b, 2), atol=1e-12)


def test_conv1d_transpose_k2_s2():
    x = rng.standard_normal((2, 5, 3))
    w = rng.standard_normal((2, 4, 3))
    k = KerasRef(1, params={"conv1d_transpose/kernel": torch.from_numpy(w), "conv1d_transpose/bias": torch.zeros(4, dtype=torch.float64)}, dtype=torch.float64, strict=True)
    y = k.ConvTranspose(torch.from_numpy(x), 4, 2, 2)
    want = np.zeros((2, 10, 4))
    for n in range(2):
        for i in range(5):
            for a in range(2):
                want[n, 2 * i + a] += w[a] @ x[n, i]
    assert np.allclose(y.detach().numpy(), want, atol=1e-12)


def test_operational_layers_against_numpy_loops():
    """Oper2D / Oper2DTranspose / Oper1DTranspose(kernel 4) (onn_layers.py:6-48, ONN_layers.py:31-52): y = sum_p conv_p(x ** p), then the
    optional activation — restated with the NumPy loop convolutions above, one power at a time"""
    q = 3
    x = 0.7 * rng.standard_normal((2, 5, 6, 3))
    ws = [rng.standard_normal((3, 3, 3, 4)) for _ in range(q)]
    bs = [rng.standard_normal(4) for _ in range(q)]
    params = {}
    for i in range(q):
        params[f"oper2d/ONN_Conv_{i + 1}/kernel"], params[f"oper2d/ONN_Conv_{i + 1}/bias"] = torch.from_numpy(ws[i]), torch.from_numpy(bs[i])
    k = KerasRef(2, params=params, dtype=torch.float64, strict=True)
    y = k.Oper(torch.from_numpy(x), 4, (3, 3), q=q, activation="tanh")
    want = np.tanh(sum(_conv_same_np(x ** (i + 1), ws[i], bs[i]) for i in range(q)))
    assert np.allclose(y.detach().numpy(), want, atol=1e-12)
    assert set(k.acts) >= {"oper2d/ONN_Conv_1", "oper2d/ONN_Conv_3", "oper2d/tf_math_pow1", "oper2d/tf_math_pow2", "oper2d/add", "oper2d"}
    # second instance: the nested model takes the next auto-name, its sub-layers keep their explicit names
    wt = [rng.standard_normal((4, 4, 2, 4)) for _ in range(2)]
    params2 = {f"oper2d_transpose/ONN_TransConv_{i + 1}/kernel": torch.from_numpy(wt[i]) for i in range(2)}
    params2.update({f"oper2d_transpose/ONN_TransConv_{i + 1}/bias": torch.zeros(2, dtype=torch.float64) for i in range(2)})
    k2 = KerasRef(2, params=params2, dtype=torch.float64, strict=True)
    x2 = torch.from_numpy(want)
    y2 = k2.Oper(x2, 2, (4, 4), q=2, strides=(2, 2), transpose=True)
    want2 = _tconv_np(want, wt[0], 0.0, 2) + _tconv_np(want ** 2, wt[1], 0.0, 2)
    assert y2.shape == (2, 10, 12, 2) and np.allclose(y2.detach().numpy(), want2, atol=1e-12)
    # 1D, transposed, kernel 4, stride 2, 'same': output position 2i - 1 + a
    x1 = rng.standard_normal((2, 5, 3))
    w1 = rng.standard_normal((4, 2, 3))
    k3 = KerasRef(1, params={"oper1d_transpose/ONN_TransConv_1/kernel": torch.from_numpy(w1),
                             "oper1d_transpose/ONN_TransConv_1/bias": torch.zeros(2, dtype=torch.float64)}, dtype=torch.float64, strict=True)
    y3 = k3.Oper(torch.from_numpy(x1), 2, 4, q=1, strides=2, transpose=True)
    want3 = np.zeros((2, 10, 2))
    for n in range(2):
        for i in range(5):
            for a in range(4):
                o = 2 * i - 1 + a
                if 0 <= o < 10:
                    want3[n, o] += w1[a] @ x1[n, i]
    assert np.allclose(y3.detach().numpy(), want3, atol=1e-12)


@pytest.mark.parametrize("ndim", [1, 2])
def test_batchnorm_training_and_moving_statistics(ndim):
    shape = (3, 4, 5, 6) if ndim == 2 else (3, 7, 6)
    x = rng.standard_normal(shape) * 2 + 1
    gamma, beta = rng.standard_normal(6), rng.standard_normal(6)
    mm, mv = rng.standard_normal(6), rng.random(6) + 0.5
    p = {"batch_normalization/gamma": gamma, "batch_normalization/beta": beta, "batch_normalization/moving_mean": mm, "batch_normalization/moving_variance": mv}
    k = KerasRef(ndim, params={a: torch.from_numpy(v) for a, v in p.items()}, dtype=torch.float64, strict=True, training=True)
    y = k.BatchNormalization(torch.from_numpy(x))
    flat = x.reshape(-1, 6)
    n = flat.shape[0]
    mean = flat.sum(0) / n
    var = ((flat - mean) ** 2).sum(0) / n
    want = (x - mean) / np.sqrt(var + 1e-3) * gamma + beta
    assert np.allclose(y.detach().numpy(), want, atol=1e-12)
    uv = var * n / (n - 1) if ndim == 2 else var     # fused (4-D) BN feeds the Bessel-corrected variance to the moving average
    assert np.allclose(k.new_moving["batch_normalization/moving_mean"].detach().numpy(), 0.99 * mm + 0.01 * mean, atol=1e-12)
    assert np.allclose(k.new_moving["batch_normalization/moving_variance"].detach().numpy(), 0.99 * mv + 0.01 * uv, atol=1e-12)
    k2 = KerasRef(ndim, params={a: torch.from_numpy(v) for a, v in p.items()}, dtype=torch.float64, strict=True, training=False)
    y2 = k2.BatchNormalization(torch.from_numpy(x))
    assert np.allclose(y2.detach().numpy(), (x - mm) / np.sqrt(mv + 1e-3) * gamma + beta, atol=1e-12)


def test_activations_pool_upsampling():
    x = rng.standard_normal((2, 4, 6, 3))
    t = torch.from_numpy(x)
    assert np.allclose(KerasRef.activation_fn("LeakyReLU", t).detach().numpy(), np.where(x > 0, x, 0.3 * x))      # Keras-2 default alpha 0.3
    assert np.allclose(KerasRef.activation_fn("relu", t).detach().numpy(), np.maximum(x, 0))
    assert np.allclose(KerasRef.activation_fn("sigmoid", t).detach().numpy(), 1 / (1 + np.exp(-x)))
    e = np.exp(x - x.max(-1, keepdims=True))
    assert np.allclose(KerasRef.activation_fn("softmax", t).detach().numpy(), e / e.sum(-1, keepdims=True))
    k = KerasRef(2, dtype=torch.float64)
    pooled = k.MaxPooling(t, (2, 2)).detach().numpy()
    want = x.reshape(2, 2, 2, 3, 2, 3).max(axis=(2, 4))
    assert np.allclose(pooled, want)
    up = k.UpSampling(t, (2, 2)).detach().numpy()
    assert np.allclose(up, x.repeat(2, 1).repeat(2, 2))
    # bilinear, half-pixel centres (tf.image.resize, align_corners=False): src = (dst + 0.5) / 2 - 0.5, edge-clamped
    bi = k.UpSampling(t, (2, 2), "bilinear").detach().numpy()
    H, W = 4, 6
    want = np.zeros((2, 8, 12, 3))
    for i in range(8):
        si = min(max((i + 0.5) / 2 - 0.5, 0), H - 1)
        i0 = int(math.floor(si)); i1 = min(i0 + 1, H - 1); fi = si - i0
        for j in range(12):
            sj = min(max((j + 0.5) / 2 - 0.5, 0), W - 1)
            j0 = int(math.floor(sj)); j1 = min(j0 + 1, W - 1); fj = sj - j0
            want[:, i, j] = (1 - fi) * ((1 - fj) * x[:, i0, j0] + fj * x[:, i0, j1]) + fi * ((1 - fj) * x[:, i1, j0] + fj * x[:, i1, j1])
    assert np.allclose(bi, want, atol=1e-12)


def test_convlstm_single_step_gates():
    """ConvLSTM2D on a length-1 sequence with h0 = c0 = 0: h = hs(o) * tanh(hs(i) * tanh(g)), hs = clip(0.2 z + 0.5, 0, 1)."""
    F_ = 2
    x = rng.standard_normal((1, 4, 4, 3))
    w = rng.standard_normal((3, 3, 3, 4 * F_))
    u = rng.standard_normal((3, 3, F_, 4 * F_))
    b = rng.standard_normal(4 * F_)
    p = {"conv_lstm2d/kernel": w, "conv_lstm2d/recurrent_kernel": u, "conv_lstm2d/bias": b}
    k = KerasRef(2, params={a: torch.from_numpy(v) for a, v in p.items()}, dtype=torch.float64, strict=True)
    h = k.ConvLSTM([torch.from_numpy(x)], F_, (3, 3)).detach().numpy()
    z = _conv_same_np(x, w, b)
    hs = lambda t: np.clip(0.2 * t + 0.5, 0, 1)
    zi, zc, zo = z[..., :F_], z[..., 2 * F_:3 * F_], z[..., 3 * F_:]
    assert np.allclose(h, hs(zo) * np.tanh(hs(zi) * np.tanh(zc)), atol=1e-12)


def test_losses_and_adam():
    z = rng.standard_normal((2, 5, 5, 1)) * 3
    y = (rng.random((2, 5, 5, 1)) > 0.5).astype(np.float64)
    p = 1 / (1 + np.exp(-z))
    bce = np.mean(np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z))))          # tf sigmoid_cross_entropy_with_logits
    assert abs(float(keras_loss("bce", torch.from_numpy(p), torch.from_numpy(y), logits=torch.from_numpy(z))) - bce) < 1e-12
    zc = rng.standard_normal((2, 5, 4))
    yc = np.eye(4)[rng.integers(0, 4, (2, 5))]
    lse = np.log(np.exp(zc - zc.max(-1, keepdims=True)).sum(-1, keepdims=True)) + zc.max(-1, keepdims=True)
    cce = np.mean(-(yc * (zc - lse)).sum(-1))
    sm = np.exp(zc - lse)
    assert abs(float(keras_loss("cce", torch.from_numpy(sm), torch.from_numpy(yc), logits=torch.from_numpy(zc))) - cce) < 1e-12
    a, t = rng.standard_normal((3, 4)), rng.standard_normal((3, 4))
    assert abs(float(keras_loss("mse", torch.from_numpy(a), torch.from_numpy(t))) - np.mean((a - t) ** 2)) < 1e-12
    assert abs(float(keras_loss("mae", torch.from_numpy(a), torch.from_numpy(t))) - np.mean(np.abs(a - t))) < 1e-12
    # Keras-2 Adam: m, v updates; alpha_t = lr sqrt(1-b2^t)/(1-b1^t); w -= alpha_t m / (sqrt(v) + eps)
    w0, g = rng.standard_normal(7), rng.standard_normal(7)
    m0, v0 = rng.standard_normal(7) * 0.1, rng.random(7) * 0.01
    w, m, v = (torch.from_numpy(q.copy()) for q in (w0, m0, v0))
    keras_adam_step(w, torch.from_numpy(g), m, v, 3, lr=1e-3)
    m1 = 0.9 * m0 + 0.1 * g
    v1 = 0.999 * v0 + 0.001 * g * g
    al = 1e-3 * math.sqrt(1 - 0.999 ** 3) / (1 - 0.9 ** 3)
    assert np.allclose(w.detach().numpy(), w0 - al * m1 / (np.sqrt(v1) + 1e-7), atol=1e-15) and np.allclose(m.detach().numpy(), m1) and np.allclose(v.detach().numpy(), v1)


# ---- committed fixtures ------------------------------------------------------------------------------------------------------
def _golden_cases():
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import CASES
    return CASES


@pytest.mark.parametrize("case", [c["name"] for c in _golden_cases()])
def test_oracle_reproduces_golden_fixture(case):
    from make_golden import CASES, run_case
    spec = next(c for c in CASES if c["name"] == case)
    want = np.load(os.path.join(GOLDEN, case + ".npz"))
    got = run_case(spec)
    assert set(got) == set(want.files)
    for key in want.files:
        assert np.allclose(got[key], want[key], rtol=1e-6, atol=1e-7), (case, key)


# ---- known answers published by the real dependency ------------------------------------------------------------------------------
def _kats():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keras_doc_kats.json")) as f:
        return json.load(f)


def test_oracle_reproduces_the_keras_documentation_examples():
    """tests/golden/keras_doc_kats.json: docstring examples of tf.keras (the numbers TensorFlow itself printed).  The oracle's loss
    definitions, hard_sigmoid, LeakyReLU slope, Adam rule, nearest up-sampling, max-pooling and BatchNorm formula must reproduce
    every one of them to the digits quoted."""
    from oracle.keras_ref import KerasRef, keras_adam_step, keras_loss, hard_sigmoid
    k = _kats()
    for case in k["losses"]:
        yt, yp = torch.tensor(case["y_true"], dtype=torch.float64), torch.tensor(case["y_pred"], dtype=torch.float64)
        if case.get("logits"):
            got = float(keras_loss(case["kind"], torch.sigmoid(yp), yt, logits=yp))
        else:
            got = float(keras_loss(case["kind"], yp, yt))
        assert abs(got - case["value"]) < 0.6 * 10 ** (-case["digits"]), (case["source"], got, case["value"])
    hs = k["hard_sigmoid"]
    assert torch.allclose(hard_sigmoid(torch.tensor(hs["x"], dtype=torch.float64)), torch.tensor(hs["y"], dtype=torch.float64), atol=1e-12)
    lr_ = k["leaky_relu"]
    assert torch.allclose(KerasRef.activation_fn("LeakyReLU", torch.tensor(lr_["x"], dtype=torch.float64)), torch.tensor(lr_["y"], dtype=torch.float64), atol=1e-12)
    ad = k["adam"]
    w = torch.tensor([ad["w0"]], dtype=torch.float64)
    keras_adam_step(w, w.clone(), torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64), 1, lr=ad["lr"])   # d/dw (w^2 / 2) = w
    assert abs(float(w) - ad["w1"]) < 0.6 * 10 ** (-ad["digits"])
    u1 = k["upsampling1d"]
    r1 = KerasRef(1, dtype=torch.float64, training=False)
    x1 = torch.arange(12, dtype=torch.float64).reshape(u1["x_shape"])
    assert torch.equal(r1.UpSampling(x1, u1["size"]), torch.tensor(u1["y"], dtype=torch.float64))
    u2 = k["upsampling2d"]
    r2 = KerasRef(2, dtype=torch.float64, training=False)
    x2 = torch.arange(12, dtype=torch.float64).reshape(u2["x_shape"])
    assert torch.equal(r2.UpSampling(x2, tuple(u2["size"])), torch.tensor(u2["y"], dtype=torch.float64))
    mp = k["maxpool2d"]
    xm = torch.tensor(mp["x"], dtype=torch.float64)[None, :, :, None]
    assert torch.equal(r2.MaxPooling(xm, mp["pool"])[0, :, :, 0], torch.tensor(mp["y"], dtype=torch.float64))
    bn = k["batchnorm"]
    r3 = KerasRef(1, dtype=torch.float64, training=True)
    yb = r3.BatchNormalization(torch.tensor(bn["x"], dtype=torch.float64)[:, None, :])          # (N, L=1, C=1)
    assert torch.allclose(yb[:, 0, :], torch.tensor(bn["y"], dtype=torch.float64), atol=1e-12)
    assert abs(float(r3.new_moving["batch_normalization/moving_mean"]) - bn["moving_mean_after"]) < 1e-12


def test_loss_kernel_semantics_reproduce_the_keras_documentation_examples():
    """the same published loss values through the float64 mirror of the CUDA loss kernel (tests/desc_emulator.py:emu_loss == csrc loss_kernel
    formulas, which the GPU test compares with the device)"""
    import types
    from b2seg import _lib as L
    from b2seg.planner import LOSS_KINDS
    from desc_emulator import PlanMem, emu_loss
    for case in _kats()["losses"]:
        yt, yp = torch.tensor(case["y_true"], dtype=torch.float64), torch.tensor(case["y_pred"], dtype=torch.float64)
        act = L.ACT_NONE
        if case.get("logits"):
            yp, act = torch.sigmoid(yp), L.ACT_SIGMOID      # the kernel sees probabilities + the head's activation (Keras' cached-logits path)
        mem = PlanMem()
        n = yp.numel()
        p_ptr, t_ptr, l_ptr = mem.alloc_bytes(4 * n, "output"), mem.alloc_bytes(4 * n, "target"), mem.alloc_bytes(64, "loss")
        mem.f32(p_ptr, n)[:] = yp.reshape(-1)
        mem.f32(t_ptr, n)[:] = yt.reshape(-1)
        d = types.SimpleNamespace(y_pred=p_ptr, y_true=t_ptr, n_pix=yp.shape[0], cout=yp.shape[1], kind=LOSS_KINDS[case["kind"]], act=act, weight=1.0,
                                  dlogits=0, loss=l_ptr, metrics=0)
        emu_loss(mem, d)
        got = float(mem.f32(l_ptr, 1))
        assert abs(got - case["value"]) < 0.6 * 10 ** (-case["digits"]), (case["source"], got, case["value"])
