"""GPU parity of each C-ABI kernel (include/b2seg.h) against a plain PyTorch fp32 evaluation of the same op.

Tolerances: inputs are bf16-rounded before both paths, accumulation is fp32 on both, outputs are rounded to
bf16 by the kernel => rel-L2 <= 4e-3 (one bf16 rounding, 2^-9 relative, plus accumulation-order noise);
fp32 outputs (wgrad, statistics, Adam) <= 1e-4.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from b2seg import _lib as L  # noqa: E402
from b2seg import lowering as lw  # noqa: E402


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def tv(t, c_off=0, Cn=None):
    N, H, W, Ct = t.shape
    esz = t.element_size()
    return lw.TView(t.data_ptr() + c_off * esz, N, H, W, Ct - c_off if Cn is None else Cn, t.stride(0), t.stride(1), t.stride(2), esz)


def stream():
    return torch.cuda.current_stream().cuda_stream


def bf(x):
    return x.to(torch.bfloat16)


@pytest.fixture(scope="module", autouse=True)
def _dev():
    L.check(L.load().b2seg_device_check(0), "device_check")
    torch.manual_seed(0)
    yield
    torch.cuda.synchronize()


CONV_CASES = [
    # N, H, W, Cin, Cout, kh, kw, act, stats
    (2, 16, 16, 64, 64, 3, 3, L.ACT_NONE, False),
    (2, 16, 16, 64, 128, 3, 3, L.ACT_RELU, True),
    (4, 8, 8, 128, 256, 3, 3, L.ACT_NONE, True),
    (2, 8, 8, 320, 512, 3, 3, L.ACT_LEAKY, False),
    (1, 32, 32, 8, 64, 3, 3, L.ACT_NONE, True),      # Cin = 3 padded to 8 (K tail zero-filled by TMA)
    (3, 12, 20, 24, 40, 3, 3, L.ACT_NONE, True),     # ragged: tiles overhang the image, odd channel counts
    (2, 1, 256, 64, 64, 1, 3, L.ACT_NONE, True),     # 1D, kernel 3
    (5, 1, 32, 64, 128, 1, 5, L.ACT_NONE, False),    # 1D, kernel 5, short rows span several samples
    (2, 16, 16, 64, 64, 1, 1, L.ACT_SIGMOID, False),
    (8, 4, 4, 256, 256, 3, 3, L.ACT_NONE, True),
    # halo-tile kernel (H % 16 == 0, W % 8 == 0 / 1D W % 128 == 0); the big ones keep the weights resident in smem
    (16, 64, 64, 64, 64, 3, 3, L.ACT_RELU, True),
    (16, 64, 64, 128, 64, 3, 3, L.ACT_NONE, True),
    (16, 64, 64, 8, 64, 3, 3, L.ACT_NONE, True),
    (4, 32, 32, 256, 256, 3, 3, L.ACT_NONE, True),
    (2, 32, 64, 128, 320, 3, 3, L.ACT_LEAKY, False),
    (4, 1, 1024, 64, 64, 1, 3, L.ACT_NONE, True),
    (40, 1, 1024, 64, 128, 1, 5, L.ACT_RELU, True),
]


@pytest.mark.parametrize("N,H,W,Cin,Cout,kh,kw,act,stats", CONV_CASES)
def test_conv_fprop(N, H, W, Cin, Cout, kh, kw, act, stats):
    dev = "cuda"
    x = bf(torch.randn(N, H, W, Cin, device=dev))
    w = bf(torch.randn(Cout, kh * kw, Cin, device=dev) * (1.0 / (kh * kw * Cin) ** 0.5))
    bias = torch.randn(Cout, device=dev)
    out = torch.zeros(N, H, W, Cout, device=dev, dtype=torch.bfloat16)
    d = lw.conv_fprop(tv(x), w.data_ptr(), Cout, kh, kw, Cin, tv(out), bias=bias.data_ptr(), act=act)
    mt = L.load().b2seg_conv_num_stat_rows(C.byref(d))
    from b2seg.planner import conv_stat_rows
    assert mt == conv_stat_rows(d, torch.cuda.get_device_properties(0).multi_processor_count)
    st = torch.zeros(mt, 2, Cout, device=dev) if stats else None
    if stats:
        d.stats = st.data_ptr()
    L.call("b2seg_conv", d, stream())
    torch.cuda.synchronize()
    wt = w.float().view(Cout, kh, kw, Cin).permute(0, 3, 1, 2)
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xp = F.pad(x.float().permute(0, 3, 1, 2), (pw, kw - 1 - pw, ph, kh - 1 - ph))
    ref = F.conv2d(xp, wt, bias)
    if act == L.ACT_RELU:
        ref = F.relu(ref)
    elif act == L.ACT_LEAKY:
        ref = F.leaky_relu(ref, 0.3)
    elif act == L.ACT_SIGMOID:
        ref = torch.sigmoid(ref)
    ref = ref.permute(0, 2, 3, 1)
    assert rel_l2(out.float(), ref) < 4e-3
    if stats:
        s = st.sum(0)
        o = out.float().view(-1, Cout)
        assert torch.allclose(s[0], o.sum(0), rtol=1e-3, atol=1e-2)
        assert torch.allclose(s[1], (o * o).sum(0), rtol=1e-3, atol=1e-2)


def test_conv_concat_slot_and_offsets():
    """input read from, and output written to, channel windows of wider buffers (concat without copies)"""
    dev = "cuda"
    N, H, W = 2, 16, 16
    xbuf = bf(torch.randn(N, H, W, 192, device=dev))
    obuf = torch.full((N, H, W, 256), 7.0, device=dev, dtype=torch.bfloat16)
    w = bf(torch.randn(64, 9, 128, device=dev) * 0.03)
    d = lw.conv_fprop(tv(xbuf, 64, 128), w.data_ptr(), 64, 3, 3, 128, tv(obuf, 128, 64))
    L.call("b2seg_conv", d, stream())
    torch.cuda.synchronize()
    ref = F.conv2d(xbuf[..., 64:192].float().permute(0, 3, 1, 2), w.float().view(64, 3, 3, 128).permute(0, 3, 1, 2), padding=1)
    assert rel_l2(obuf[..., 128:192].float(), ref.permute(0, 2, 3, 1)) < 4e-3
    assert float((obuf[..., :128].float() - 7).abs().max()) == 0 and float((obuf[..., 192:].float() - 7).abs().max()) == 0


DGRAD_CASES = [(16, 64, 64, 128, 64, 3, 3), (4, 32, 32, 256, 256, 3, 3), (40, 1, 1024, 64, 128, 1, 3), (2, 16, 16, 64, 64, 3, 3), (2, 8, 8, 256, 128, 3, 3), (4, 8, 8, 128, 320, 3, 3), (3, 12, 20, 24, 40, 3, 3),
               (2, 1, 128, 64, 128, 1, 3), (2, 16, 16, 64, 64, 1, 1)]


@pytest.mark.parametrize("N,H,W,Cin,Cout,kh,kw", DGRAD_CASES)
def test_conv_dgrad(N, H, W, Cin, Cout, kh, kw):
    dev = "cuda"
    dy = bf(torch.randn(N, H, W, Cout, device=dev))
    w = bf(torch.randn(Cout, kh * kw, Cin, device=dev) * 0.05)
    dx = torch.zeros(N, H, W, Cin, device=dev, dtype=torch.bfloat16)
    L.call("b2seg_conv", lw.conv_dgrad(tv(dy), w.data_ptr(), Cout, kh, kw, Cin, tv(dx)), stream())
    torch.cuda.synchronize()
    x = torch.zeros(N, Cin, H, W, device=dev, requires_grad=True)
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    y = F.conv2d(F.pad(x, (pw, kw - 1 - pw, ph, kh - 1 - ph)), w.float().view(Cout, kh, kw, Cin).permute(0, 3, 1, 2))
    y.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_l2(dx.float(), x.grad.permute(0, 2, 3, 1)) < 4e-3


def test_conv_dgrad_mask_fusion():
    dev = "cuda"
    N, H, W, Cin, Cout = 2, 16, 16, 128, 64
    dy = bf(torch.randn(N, H, W, Cout, device=dev))
    w = bf(torch.randn(Cout, 9, Cin, device=dev) * 0.05)
    yfwd = bf(torch.randn(N, H, W, Cin, device=dev))
    dx = torch.zeros(N, H, W, Cin, device=dev, dtype=torch.bfloat16)
    d = lw.conv_dgrad(tv(dy), w.data_ptr(), Cout, 3, 3, Cin, tv(dx), mul_view=tv(yfwd, 0, 64), mul_mode=L.ACT_LEAKY)
    L.call("b2seg_conv", d, stream())
    torch.cuda.synchronize()
    x = torch.zeros(N, Cin, H, W, device=dev, requires_grad=True)
    F.conv2d(x, w.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2), padding=1).backward(dy.float().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1).clone()
    ref[..., :64] *= torch.where(yfwd[..., :64].float() > 0, 1.0, 0.3)
    assert rel_l2(dx.float(), ref) < 6e-3
    # statistics after the mask + row sum = bias gradient of the layer whose derivative was fused (two image sizes: the
    # 16x16 map takes the per-tile-box kernel, 32x64 the halo kernel)
    for (N2, H2, W2) in ((2, 16, 16), (3, 32, 64)):
        dy = bf(torch.randn(N2, H2, W2, Cout, device=dev))
        yfwd = bf(torch.randn(N2, H2, W2, Cin, device=dev))
        dx = torch.zeros(N2, H2, W2, Cin, device=dev, dtype=torch.bfloat16)
        d = lw.conv_dgrad(tv(dy), w.data_ptr(), Cout, 3, 3, Cin, tv(dx), mul_view=tv(yfwd, 0, 64), mul_mode=L.ACT_LEAKY)
        rows = L.load().b2seg_conv_num_stat_rows(C.byref(d))
        stats = torch.zeros(rows, 2, Cin, device=dev)
        d.stats = stats.data_ptr()
        L.call("b2seg_conv", d, stream())
        db = torch.full((64,), 7.0, device=dev)
        L.call("b2seg_rowsum", L.RowsumDesc(stats.data_ptr(), rows, 2 * Cin, 64, db.data_ptr(), 0), stream())
        db2 = torch.ones(64, device=dev)
        L.call("b2seg_rowsum", L.RowsumDesc(stats.data_ptr(), rows, 2 * Cin, 64, db2.data_ptr(), 1), stream())
        torch.cuda.synchronize()
        want = dx[..., :64].float().sum((0, 1, 2))
        assert torch.allclose(db, want, rtol=1e-4, atol=1e-3), float((db - want).abs().max())
        assert torch.allclose(db2, want + 1.0, rtol=1e-4, atol=1e-3)


WGRAD_CASES = [(2, 16, 16, 64, 64, 3, 3, 0), (2, 16, 16, 64, 64, 3, 3, 1), (4, 8, 8, 128, 256, 3, 3, 0), (2, 8, 8, 320, 136, 3, 3, 0),
               (3, 12, 20, 24, 40, 3, 3, 0), (2, 1, 256, 64, 64, 1, 3, 0), (1, 32, 32, 8, 64, 3, 3, 0), (2, 16, 16, 64, 64, 1, 1, 0),
               # fused-tap halo kernel shapes: Cin 128 (3 taps x N=128), 256 / 512 (tap pairs x N=256), 192, several Cin tiles,
               # small and ragged images (H < 16, W not a multiple of 8), 1D kernels 3 / 5 / 7, forced split-K
               (4, 32, 32, 128, 128, 3, 3, 0), (2, 32, 32, 256, 128, 3, 3, 0), (2, 16, 16, 512, 256, 3, 3, 0), (2, 16, 24, 192, 64, 3, 3, 0),
               (8, 4, 4, 256, 256, 3, 3, 0), (3, 6, 12, 128, 64, 3, 3, 0), (2, 2, 4, 64, 128, 3, 3, 0), (4, 64, 64, 64, 128, 3, 3, 0),
               (2, 1, 512, 128, 128, 1, 3, 0), (2, 1, 256, 256, 64, 1, 5, 0), (3, 1, 96, 64, 64, 1, 7, 0), (2, 1, 64, 16, 32, 1, 5, 0),
               (4, 32, 32, 128, 128, 3, 3, 3), (2, 32, 32, 320, 128, 3, 3, 2)]


@pytest.mark.parametrize("N,H,W,Cin,Cout,kh,kw,ksplit", WGRAD_CASES)
def test_conv_wgrad(N, H, W, Cin, Cout, kh, kw, ksplit):
    dev = "cuda"
    dy = bf(torch.randn(N, H, W, Cout, device=dev))
    xin = bf(torch.randn(N, H, W, Cin, device=dev))
    dw = torch.zeros(Cout, kh * kw, Cin, device=dev)
    L.call("b2seg_wgrad", lw.conv_wgrad(tv(dy), tv(xin), dw.data_ptr(), Cout, kh, kw, Cin, ksplit=ksplit), stream())
    torch.cuda.synchronize()
    w = torch.zeros(Cout, Cin, kh, kw, device=dev, requires_grad=True)
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    y = F.conv2d(F.pad(xin.float().permute(0, 3, 1, 2), (pw, kw - 1 - pw, ph, kh - 1 - ph)), w)
    y.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_l2(dw.view(Cout, kh, kw, Cin), w.grad.permute(0, 2, 3, 1)) < 1e-4


TCONV_CASES = [(2, 16, 16, 128, 64, 4, 4), (8, 32, 32, 64, 64, 4, 4), (2, 8, 8, 128, 64, 4, 4), (2, 4, 4, 256, 128, 4, 4), (2, 1, 64, 128, 64, 1, 2), (1, 6, 10, 24, 16, 4, 4)]


@pytest.mark.parametrize("N,H,W,Cin,Cout,kh,kw", TCONV_CASES)
def test_tconv_all(N, H, W, Cin, Cout, kh, kw):
    dev = "cuda"
    sh = 2 if kh > 1 else 1
    pad = (1 if kh == 4 else 0, 1 if kw == 4 else 0)
    x = bf(torch.randn(N, H, W, Cin, device=dev))
    w = bf(torch.randn(Cout, kh * kw, Cin, device=dev) * 0.05)  # internal layout
    bias = torch.randn(Cout, device=dev)
    out = torch.zeros(N, H * sh, W * 2, Cout, device=dev, dtype=torch.bfloat16)
    L.call("b2seg_conv", lw.tconv_fprop(tv(x), w.data_ptr(), Cout, kh, kw, Cin, tv(out), bias=bias.data_ptr(), act=L.ACT_LEAKY), stream())
    torch.cuda.synchronize()
    wt = w.float().view(Cout, kh, kw, Cin).permute(3, 0, 1, 2).contiguous().requires_grad_(True)  # torch (Cin,Cout,kh,kw)
    xt = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    pre = F.conv_transpose2d(xt, wt, bias, stride=(sh, 2), padding=pad)
    ref = F.leaky_relu(pre, 0.3)
    assert rel_l2(out.float(), ref.permute(0, 2, 3, 1)) < 4e-3
    dy = bf(torch.randn(N, H * sh, W * 2, Cout, device=dev))
    pre.backward(dy.float().permute(0, 3, 1, 2))
    dx = torch.zeros(N, H, W, Cin, device=dev, dtype=torch.bfloat16)
    L.call("b2seg_conv", lw.tconv_dgrad(tv(dy), w.data_ptr(), Cout, kh, kw, Cin, tv(dx)), stream())
    dw = torch.zeros(Cout, kh * kw, Cin, device=dev)
    L.call("b2seg_wgrad", lw.tconv_wgrad(tv(dy), tv(x), dw.data_ptr(), Cout, kh, kw, Cin), stream())
    torch.cuda.synchronize()
    assert rel_l2(dx.float(), xt.grad.permute(0, 2, 3, 1)) < 4e-3
    assert rel_l2(dw.view(Cout, kh, kw, Cin), wt.grad.permute(1, 2, 3, 0)) < 1e-4


def test_bn_finalize_act_pool_and_backward():
    dev = "cuda"
    N, H, W, Cc = 4, 16, 16, 64
    z = bf(torch.randn(N, H, W, Cc, device=dev) * 2 + 0.5)
    gamma = torch.rand(Cc, device=dev) + 0.5
    beta = torch.randn(Cc, device=dev) * 0.1
    zf = z.float().view(-1, Cc)
    part = torch.stack([zf.sum(0), (zf * zf).sum(0)]).view(1, 2, Cc).contiguous()
    mm, mv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    scale, shift, mean, rstd = (torch.zeros(Cc, device=dev) for _ in range(4))
    fd = L.BnFinalizeDesc(part.data_ptr(), 1, Cc, float(N * H * W), gamma.data_ptr(), beta.data_ptr(), mm.data_ptr(), mv.data_ptr(),
                          1, 1, 1e-3, 0.99, scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 0)
    L.call("b2seg_bn_finalize", fd, stream())
    a = torch.zeros_like(z)
    cat = torch.zeros(N, H, W, 2 * Cc, device=dev, dtype=torch.bfloat16)
    pooled = torch.zeros(N, H // 2, W // 2, Cc, device=dev, dtype=torch.bfloat16)
    ad = L.BnActDesc()
    ad.x, ad.scale, ad.shift, ad.act, ad.n_out = tv(z).to_c(), scale.data_ptr(), shift.data_ptr(), L.ACT_RELU, 2
    ad.out[0], ad.out[1] = tv(a).to_c(), tv(cat, Cc, Cc).to_c()
    ad.pool_h, ad.pool_w, ad.pooled = 2, 2, tv(pooled).to_c()
    L.call("b2seg_bn_act", ad, stream())
    torch.cuda.synchronize()
    zt = z.float().requires_grad_(True)
    mu, var = zt.view(-1, Cc).mean(0), zt.view(-1, Cc).var(0, unbiased=False)
    y = F.relu((zt - mu) / torch.sqrt(var + 1e-3) * gamma + beta)
    assert torch.allclose(mean, mu.detach(), atol=1e-4) and torch.allclose(mm, 0.01 * mu.detach(), atol=1e-5)
    cnt = N * H * W
    assert torch.allclose(mv, 0.99 + 0.01 * var.detach() * cnt / (cnt - 1), atol=1e-4)
    assert rel_l2(a.float(), y.detach()) < 4e-3
    assert torch.equal(a, cat[..., Cc:])
    yp = F.max_pool2d(y.permute(0, 3, 1, 2), 2)
    assert rel_l2(pooled.float(), yp.detach().permute(0, 2, 3, 1)) < 4e-3
    # backward with a direct and a pool-routed gradient source
    g1 = bf(torch.randn(N, H, W, Cc, device=dev))
    g2 = bf(torch.randn(N, H // 2, W // 2, Cc, device=dev))
    (y * g1.float()).sum().backward(retain_graph=True)
    (yp * g2.float().permute(0, 3, 1, 2)).sum().backward()
    nb = 32
    partials = torch.zeros(nb, 2, Cc, device=dev)
    dgamma, dbeta = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    dz = torch.zeros_like(z)
    bd = L.BnBwdDesc()
    bd.x, bd.scale, bd.shift, bd.mean, bd.rstd = tv(z).to_c(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    bd.act, bd.n_src = L.ACT_RELU, 2
    bd.src[0] = L.GradSrc(tv(g1).to_c(), 0, 1, 1)
    bd.src[1] = L.GradSrc(tv(g2).to_c(), 1, 2, 2)
    bd.count, bd.partials, bd.n_blocks = float(cnt), partials.data_ptr(), nb
    bd.dgamma, bd.dbeta, bd.dx = dgamma.data_ptr(), dbeta.data_ptr(), tv(dz).to_c()
    L.call("b2seg_bn_bwd", bd, stream())
    torch.cuda.synchronize()
    assert rel_l2(dz.float(), zt.grad) < 8e-3
    # accumulate = 1 (plan mode): the caller zeroes dgamma / dbeta, the op adds into them; same results
    dgamma_a, dbeta_a, dz_a = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev), torch.zeros_like(z)
    bd.dgamma, bd.dbeta, bd.dx, bd.accumulate = dgamma_a.data_ptr(), dbeta_a.data_ptr(), tv(dz_a).to_c(), 1
    L.call("b2seg_bn_bwd", bd, stream())
    torch.cuda.synchronize()
    assert rel_l2(dgamma_a, dgamma) < 1e-5 and rel_l2(dbeta_a, dbeta) < 1e-5 and rel_l2(dz_a.float(), dz.float()) < 1e-3
    # dgamma/dbeta against autograd through gamma/beta
    g_ = gamma.clone().requires_grad_(True)
    b_ = beta.clone().requires_grad_(True)
    zz = z.float()
    y2 = F.relu((zz - mu.detach()) / torch.sqrt(var.detach() + 1e-3) * g_ + b_)
    ((y2 * g1.float()).sum() + (F.max_pool2d(y2.permute(0, 3, 1, 2), 2) * g2.float().permute(0, 3, 1, 2)).sum()).backward()
    assert rel_l2(dgamma, g_.grad) < 2e-3 and rel_l2(dbeta, b_.grad) < 2e-3


@pytest.mark.parametrize("cout,act,extra", [(1, L.ACT_RELU, False), (2, L.ACT_LEAKY, True), (1, L.ACT_RELU, True)])
def test_bn_bwd_with_folded_head(cout, act, extra):
    """b2seg_gradsrc kind 2: the pointwise head's input gradient is formed on the fly and its dW / db fall out of pass 0"""
    dev = "cuda"
    N, H, W, Cc = 3, 20, 24, 64
    z = bf(torch.randn(N, H, W, Cc, device=dev) * 2 + 0.3)
    gamma, beta = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
    hw = torch.randn(Cc, cout, device=dev) * 0.2
    dl = torch.randn(N, H, W, cout, device=dev) * 0.1
    g_extra = bf(torch.randn(N, H, W, Cc, device=dev) * 0.05)
    zt = z.float().requires_grad_(True)
    gt, bt, hwt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True), hw.clone().requires_grad_(True)
    mu, var = zt.view(-1, Cc).mean(0), zt.view(-1, Cc).var(0, unbiased=False)
    pre = (zt - mu) / torch.sqrt(var + 1e-3) * gt + bt
    a = F.relu(pre) if act == L.ACT_RELU else F.leaky_relu(pre, 0.3)
    hb = torch.zeros(cout, device=dev, requires_grad=True)
    loss = ((a @ hwt + hb) * dl).sum() + ((a * g_extra.float()).sum() if extra else 0.0)
    loss.backward()
    rstd = 1.0 / torch.sqrt(var.detach() + 1e-3)
    scale = (gamma * rstd).contiguous()
    shift = (beta - mu.detach() * gamma * rstd).contiguous()
    mean = mu.detach().contiguous()
    dgamma, dbeta, dw, db = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev), torch.zeros(Cc, cout, device=dev), torch.zeros(cout, device=dev)
    dz = torch.zeros_like(z)
    bd = L.BnBwdDesc()
    bd.x, bd.scale, bd.shift, bd.mean, bd.rstd = tv(z).to_c(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    bd.act, bd.n_src = act, 2 if extra else 1
    bd.src[0] = L.GradSrc(tv(z).to_c(), 2, 1, 1, dl.data_ptr(), hw.data_ptr(), dw.data_ptr(), db.data_ptr(), cout)
    if extra:
        bd.src[1] = L.GradSrc(tv(g_extra).to_c(), 0, 1, 1)
    bd.count, bd.partials, bd.n_blocks = float(N * H * W), 0, 0
    bd.dgamma, bd.dbeta, bd.dx, bd.accumulate = dgamma.data_ptr(), dbeta.data_ptr(), tv(dz).to_c(), 1
    L.call("b2seg_bn_bwd", bd, stream())
    torch.cuda.synchronize()
    assert rel_l2(dz.float(), zt.grad) < 8e-3
    assert rel_l2(dgamma, gt.grad) < 2e-3 and rel_l2(dbeta, bt.grad) < 2e-3
    assert rel_l2(dw, hwt.grad) < 2e-3 and rel_l2(db, hb.grad) < 1e-4


def test_adam_keras_rule():
    dev = "cuda"
    n = 4096 + 8
    w = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev) * 1e-3
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    wb = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    w0 = w.clone()
    lr, b1, b2, eps = 2e-4, 0.9, 0.999, 1e-7
    mr, vr, wr = torch.zeros_like(w), torch.zeros_like(w), w0.clone().double()
    for t in (1, 2, 3):
        L.call("b2seg_adam", L.AdamDesc(w.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), wb.data_ptr(), n, lr, b1, b2, eps, 1.0, t), stream())
        mr = b1 * mr + (1 - b1) * g
        vr = b2 * vr + (1 - b2) * g * g
        alpha = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
        wr = wr - alpha * mr.double() / (vr.double().sqrt() + eps)
    torch.cuda.synchronize()
    assert torch.allclose(w.double(), wr, atol=1e-6)
    assert torch.equal(wb, w.to(torch.bfloat16))


@pytest.mark.parametrize("cout,act,kind,Cin", [(1, L.ACT_SIGMOID, 0, 64), (4, L.ACT_SOFTMAX, 1, 64), (1, L.ACT_NONE, 2, 128), (2, L.ACT_NONE, 3, 64),
                                                (8, L.ACT_SOFTMAX, 1, 64), (3, L.ACT_NONE, 2, 256), (5, L.ACT_NONE, 2, 320), (1, L.ACT_SIGMOID, 0, 8)])
def test_head_and_loss(cout, act, kind, Cin):
    dev = "cuda"
    N, H, W = 2, 16, 16
    x = bf(torch.randn(N, H, W, Cin, device=dev))
    w = torch.randn(Cin, cout, device=dev) * 0.1
    b = torch.randn(cout, device=dev) * 0.1
    y = torch.zeros(N, H, W, cout, device=dev)
    if kind == 1:
        tgt = F.one_hot(torch.randint(0, cout, (N, H, W), device=dev), cout).float()
    elif kind == 0:
        tgt = (torch.rand(N, H, W, cout, device=dev) > 0.7).float()
    else:
        tgt = torch.randn(N, H, W, cout, device=dev)
    dl = torch.zeros_like(y)
    loss = torch.zeros(1, device=dev)
    dx = torch.zeros_like(x)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    hd = L.HeadDesc()
    hd.x, hd.w, hd.b, hd.cout, hd.act, hd.stride = tv(x).to_c(), w.data_ptr(), b.data_ptr(), cout, act, 1
    hd.y, hd.dlogits, hd.dx, hd.dw, hd.db = y.data_ptr(), dl.data_ptr(), tv(dx).to_c(), dw.data_ptr(), db.data_ptr()
    L.call("b2seg_head_fwd", hd, stream())
    L.call("b2seg_loss", L.LossDesc(y.data_ptr(), tgt.data_ptr(), N * H * W, cout, kind, act, 1.0, dl.data_ptr(), loss.data_ptr()), stream())
    L.call("b2seg_head_bwd", hd, stream())
    torch.cuda.synchronize()
    xt = x.float().requires_grad_(True)
    wt, bt = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    z = xt @ wt + bt
    if kind == 0:
        ref_y, ref_loss = torch.sigmoid(z), F.binary_cross_entropy_with_logits(z, tgt)
    elif kind == 1:
        ref_y, ref_loss = torch.softmax(z, -1), -(tgt * torch.log_softmax(z, -1)).sum(-1).mean()
    elif kind == 2:
        ref_y, ref_loss = z, ((z - tgt) ** 2).mean()
    else:
        ref_y, ref_loss = z, (z - tgt).abs().mean()
    ref_loss.backward()
    assert torch.allclose(y, ref_y.detach(), atol=2e-5, rtol=1e-4)
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * max(1.0, abs(float(ref_loss)))
    assert rel_l2(dx.float(), xt.grad) < 4e-3
    assert rel_l2(dw, wt.grad) < 1e-4 and rel_l2(db, bt.grad) < 1e-4


@pytest.mark.parametrize("cout", [1, 2, 6])
def test_head_on_channel_window_ragged(cout):
    """pixel-contiguous fast path on a channel window of a wider buffer (pitch > C) with a pixel count that is no multiple
    of the thread tiling; outputs outside the window must stay untouched"""
    dev = "cuda"
    N, H, W, Cin = 3, 5, 7, 64
    xb = bf(torch.randn(N, H, W, 3 * Cin, device=dev))
    dxb = torch.full((N, H, W, 2 * Cin), 9.0, device=dev, dtype=torch.bfloat16)
    w = torch.randn(Cin, cout, device=dev) * 0.1
    b = torch.randn(cout, device=dev) * 0.1
    y = torch.zeros(N, H, W, cout, device=dev)
    dl = torch.randn(N, H, W, cout, device=dev)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    hd = L.HeadDesc()
    hd.x, hd.w, hd.b, hd.cout, hd.act, hd.stride = tv(xb, Cin, Cin).to_c(), w.data_ptr(), b.data_ptr(), cout, L.ACT_NONE, 1
    hd.y, hd.dlogits, hd.dx, hd.dw, hd.db = y.data_ptr(), dl.data_ptr(), tv(dxb, Cin, Cin).to_c(), dw.data_ptr(), db.data_ptr()
    L.call("b2seg_head_fwd", hd, stream())
    L.call("b2seg_head_bwd", hd, stream())
    torch.cuda.synchronize()
    x = xb[..., Cin:2 * Cin].float()
    assert torch.allclose(y, x @ w + b, atol=2e-5, rtol=1e-4)
    assert rel_l2(dxb[..., Cin:].float(), dl @ w.t()) < 4e-3
    assert torch.all(dxb[..., :Cin] == 9.0)
    assert rel_l2(dw, torch.einsum("nhwc,nhwo->co", x, dl)) < 1e-4 and rel_l2(db, dl.sum((0, 1, 2))) < 1e-4


def test_cast_eltwise_colsum():
    dev = "cuda"
    N, H, W = 2, 8, 8
    src = torch.rand(N, H, W, 3, device=dev)
    out = torch.full((N, H, W, 8), 5.0, device=dev, dtype=torch.bfloat16)
    L.call("b2seg_cast_input", L.CastDesc(src.data_ptr(), N, H, W, 3, tv(out).to_c()), stream())
    a, b = bf(torch.randn(N, H, W, 64, device=dev)), bf(torch.randn(N, H, W, 64, device=dev))
    o = torch.zeros_like(a)
    L.call("b2seg_eltwise", L.EltwiseDesc(0, tv(a).to_c(), tv(b).to_c(), lw.NULL_VIEW.to_c(), tv(o).to_c()), stream())
    cs = torch.zeros(64, device=dev)
    L.call("b2seg_colsum", L.ColsumDesc(tv(a).to_c(), cs.data_ptr(), 0, 0), stream())
    torch.cuda.synchronize()
    assert torch.equal(out[..., :3], src.to(torch.bfloat16)) and float(out[..., 3:].float().abs().max()) == 0
    assert torch.equal(o, (a.float() + b.float()).to(torch.bfloat16))
    assert torch.allclose(cs, a.float().view(-1, 64).sum(0), atol=1e-3)


# ------------------------------------------------------------------------------------------------------------------
# kernels of the decoder variants beyond plain UNet
@pytest.mark.parametrize("N,H,W,Cc,fh,fw,mode,act", [(2, 8, 8, 64, 2, 2, 1, L.ACT_NONE), (2, 4, 6, 16, 4, 4, 1, L.ACT_SIGMOID),
                                                     (1, 2, 2, 8, 16, 16, 1, L.ACT_SIGMOID), (3, 1, 32, 64, 1, 2, 0, L.ACT_NONE),
                                                     (2, 1, 16, 8, 1, 4, 0, L.ACT_SIGMOID), (2, 16, 16, 8, 2, 2, 1, L.ACT_NONE)])
def test_resize_fwd_bwd(N, H, W, Cc, fh, fw, mode, act):
    dev = "cuda"
    x = bf(torch.randn(N, H, W, Cc, device=dev))
    y = torch.zeros(N, H * fh, W * fw, Cc, device=dev, dtype=torch.bfloat16)
    L.call("b2seg_resize_fwd", L.ResizeDesc(tv(x).to_c(), tv(y).to_c(), lw.NULL_VIEW.to_c(), fh, fw, mode, act, 0), stream())
    torch.cuda.synchronize()
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    if mode == 1:
        ref = F.interpolate(xt, scale_factor=(fh, fw), mode="bilinear", align_corners=False)
    else:
        ref = xt.repeat_interleave(fh, 2).repeat_interleave(fw, 3)
    if act == L.ACT_SIGMOID:
        ref = torch.sigmoid(ref)
    assert rel_l2(y.float(), ref.permute(0, 2, 3, 1)) < 4e-3
    dy = bf(torch.randn_like(y.float()))
    # backward uses the STORED (bf16) forward output for the activation derivative
    yt = y.float().permute(0, 3, 1, 2)
    g = dy.float().permute(0, 3, 1, 2) * (yt * (1 - yt) if act == L.ACT_SIGMOID else 1.0)
    lin = F.interpolate(xt, scale_factor=(fh, fw), mode="bilinear", align_corners=False) if mode == 1 else xt.repeat_interleave(fh, 2).repeat_interleave(fw, 3)
    lin.backward(g)
    dx = torch.zeros_like(x)
    L.call("b2seg_resize_bwd", L.ResizeDesc(tv(dx).to_c(), tv(dy).to_c(), tv(y).to_c(), fh, fw, mode, act, 0), stream())
    torch.cuda.synchronize()
    assert rel_l2(dx.float(), xt.grad.permute(0, 2, 3, 1)) < 5e-3


def test_mulbc_colstats_lstm_poolbwd_eltwise_act():
    dev = "cuda"
    N, H, W, Cc = 2, 16, 16, 64
    a = bf(torch.randn(N, H, W, Cc, device=dev))
    b = bf(torch.rand(N, H, W, 8, device=dev))
    out = torch.zeros_like(a)
    nv = lw.NULL_VIEW.to_c()
    L.call("b2seg_mulbc_fwd", L.MulbcDesc(tv(a).to_c(), tv(b).to_c(), tv(out).to_c(), nv, nv, nv), stream())
    g = bf(torch.randn_like(a.float()))
    da, db = torch.zeros_like(a), torch.full_like(b, 3.0)
    L.call("b2seg_mulbc_bwd", L.MulbcDesc(tv(a).to_c(), tv(b).to_c(), nv, tv(g).to_c(), tv(da).to_c(), tv(db).to_c()), stream())
    torch.cuda.synchronize()
    assert rel_l2(out.float(), a.float() * b.float()[..., :1]) < 4e-3
    assert rel_l2(da.float(), g.float() * b.float()[..., :1]) < 4e-3
    assert rel_l2(db.float()[..., 0], (g.float() * a.float()).sum(-1)) < 4e-3 and float(db.float()[..., 1:].abs().max()) == 0
    # column statistics
    nb = 16
    part = torch.zeros(nb, 2, Cc, device=dev)
    L.call("b2seg_colstats", L.ColstatsDesc(tv(a).to_c(), part.data_ptr(), nb), stream())
    torch.cuda.synchronize()
    af = a.float().view(-1, Cc)
    assert torch.allclose(part[:, 0].sum(0), af.sum(0), atol=1e-2, rtol=1e-4) and torch.allclose(part[:, 1].sum(0), (af * af).sum(0), atol=1e-2, rtol=1e-4)
    # ConvLSTM gates
    Fg = 32
    z = bf(torch.randn(N, H, W, 3 * Fg, device=dev) * 2)
    # hard_sigmoid has a kink at z = +-2.5, which bf16 represents exactly: keep samples off it (fused vs two-step
    # rounding of 0.2 z + 0.5 decides the side there, and a flipped derivative is a full-size error)
    z = torch.where((z.float().abs() - 2.5).abs() < 0.02, bf(z.float() * 0.9), z).contiguous()
    h = torch.zeros(N, H, W, Fg, device=dev, dtype=torch.bfloat16)
    L.call("b2seg_lstm_fwd", L.LstmDesc(tv(z).to_c(), tv(h).to_c(), nv, nv, Fg), stream())
    zt = z.float().requires_grad_(True)
    hs = lambda t: torch.clamp(0.2 * t + 0.5, 0, 1)
    href = hs(zt[..., 2 * Fg:]) * torch.tanh(hs(zt[..., :Fg]) * torch.tanh(zt[..., Fg:2 * Fg]))
    dh = bf(torch.randn_like(h.float()))
    href.backward(dh.float())
    dz = torch.zeros_like(z)
    L.call("b2seg_lstm_bwd", L.LstmDesc(tv(z).to_c(), nv, tv(dh).to_c(), tv(dz).to_c(), Fg), stream())
    torch.cuda.synchronize()
    assert rel_l2(h.float(), href.detach()) < 4e-3 and rel_l2(dz.float(), zt.grad) < 5e-3
    # max-pool backward, 4x4 window
    y = bf(torch.randn(N, H, W, Cc, device=dev))
    dp = bf(torch.randn(N, H // 4, W // 4, Cc, device=dev))
    dx = torch.zeros_like(y)
    L.call("b2seg_pool_bwd", L.PoolBwdDesc(tv(y).to_c(), tv(dp).to_c(), tv(dx).to_c(), 4, 4), stream())
    torch.cuda.synchronize()
    yt = y.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.max_pool2d(yt, 4).backward(dp.float().permute(0, 3, 1, 2))
    assert rel_l2(dx.float(), yt.grad.permute(0, 2, 3, 1)) < 1e-6
    # add + relu
    o = torch.zeros_like(a)
    L.call("b2seg_eltwise", L.EltwiseDesc(0, tv(a).to_c(), tv(y).to_c(), nv, tv(o).to_c(), L.ACT_RELU), stream())
    torch.cuda.synchronize()
    assert torch.equal(o, torch.relu(a.float() + y.float()).to(torch.bfloat16))


def test_strided_1x1_dgrad_and_bn_act_cvalid():
    dev = "cuda"
    N, H, W, Cin, Cout = 2, 16, 16, 64, 64
    dy = bf(torch.randn(N, H // 2, W // 2, Cout, device=dev))
    w = bf(torch.randn(Cout, 1, Cin, device=dev) * 0.1)
    dx = torch.zeros(N, H, W, Cin, device=dev, dtype=torch.bfloat16)
    L.call("b2seg_conv", lw.conv_dgrad(tv(dy), w.data_ptr(), Cout, 1, 1, Cin, tv(dx).parity(0, 0, 2, 2)), stream())
    torch.cuda.synchronize()
    x = torch.zeros(N, Cin, H, W, device=dev, requires_grad=True)
    F.conv2d(x, w.float().view(Cout, 1, 1, Cin).permute(0, 3, 1, 2), stride=2).backward(dy.float().permute(0, 3, 1, 2))
    assert rel_l2(dx.float(), x.grad.permute(0, 2, 3, 1)) < 4e-3
    z = bf(torch.randn(N, H, W, 8, device=dev))
    o = torch.zeros_like(z)
    d = L.BnActDesc()
    d.x, d.act, d.n_out, d.c_valid = tv(z).to_c(), L.ACT_SIGMOID, 1, 1
    d.out[0] = tv(o).to_c()
    L.call("b2seg_bn_act", d, stream())
    torch.cuda.synchronize()
    assert rel_l2(o.float()[..., 0], torch.sigmoid(z.float()[..., 0])) < 4e-3 and float(o.float()[..., 1:].abs().max()) == 0


# ------------------------------------------------------------------------------------------------------------------
# every loss of utils/tf_losses.py on every head activation: value, seed dL/dlogits and the metric sums vs autograd through the oracle
_LOSS_NAMES = ["bce", "cce", "mse", "mae", "msle", "huber", "logcosh", "focal", "poisson", "kld", "hinge", "squared_hinge", "mape",
               "categorical_hinge", "cosine"]


@pytest.mark.parametrize("kind", range(15), ids=_LOSS_NAMES)
@pytest.mark.parametrize("act", [L.ACT_NONE, L.ACT_SIGMOID, L.ACT_SOFTMAX], ids=["linear", "sigmoid", "softmax"])
def test_loss_kinds_and_metrics(kind, act):
    from oracle.keras_ref import keras_loss
    name = _LOSS_NAMES[kind]
    if (kind in (0, 7) and act == L.ACT_SOFTMAX) or (kind == 1 and act == L.ACT_SIGMOID):
        buf = torch.zeros(64, device="cuda")          # (valid memory: a library that failed to refuse must not fault)
        with pytest.raises(L.B2SegError):
            L.call("b2seg_loss", L.LossDesc(buf.data_ptr(), buf.data_ptr(), 4, 2, kind, act, 1.0, 0, 0, 0), stream())
        return
    if kind == 8 and act == L.ACT_NONE:
        pytest.skip("Poisson of negative predictions is NaN in Keras too")
    dev, npix, co = "cuda", 3 * 37 * 5, 4
    g = torch.Generator(device="cpu").manual_seed(100 + 3 * kind + act)
    # float32 reference: the clipped-probability branches differ visibly between float32 and float64 (1 - (1 - 1e-7) is 1.19e-7 in
    # float32), and TensorFlow computes them in float32 like the kernel
    z = (torch.randn(npix, co, generator=g) * 1.5).requires_grad_(True)
    if name in ("cce", "categorical_hinge", "kld"):
        t = F.one_hot(torch.randint(0, co, (npix,), generator=g), co).float()
    elif name in ("bce", "focal", "hinge", "squared_hinge"):
        t = (torch.rand(npix, co, generator=g) > 0.6).float()
    elif name in ("msle", "poisson", "mape"):
        t = torch.rand(npix, co, generator=g) * 2 + 0.25
    else:
        t = torch.randn(npix, co, generator=g)
    p = z if act == L.ACT_NONE else (torch.sigmoid(z) if act == L.ACT_SIGMOID else torch.softmax(z, -1))
    own = (kind in (0, 7) and act == L.ACT_SIGMOID) or (kind == 1 and act == L.ACT_SOFTMAX)
    want = keras_loss(name, p, t, logits=z if own else None)
    (dz,) = torch.autograd.grad(want, z)
    y, tt = p.detach().float().to(dev).contiguous(), t.float().to(dev).contiguous()
    dl, loss, met = torch.zeros_like(y), torch.zeros(1, device=dev), torch.zeros(8, device=dev)
    L.call("b2seg_loss", L.LossDesc(y.data_ptr(), tt.data_ptr(), npix, co, kind, act, 0.7, dl.data_ptr(), loss.data_ptr(), met.data_ptr()), stream())
    torch.cuda.synchronize()
    assert abs(float(loss) - 0.7 * float(want)) < 1e-4 * max(1.0, abs(float(want))), (float(loss), 0.7 * float(want))
    assert abs(float(met[0]) - float(want)) < 1e-4 * max(1.0, abs(float(want)))
    assert rel_l2(dl.cpu().double(), 0.7 * dz.double()) < 1e-3, rel_l2(dl.cpu().double(), 0.7 * dz.double())
    pe, te = p.detach().double(), t.double()
    assert abs(float(met[1]) - float(((pe - te) ** 2).sum())) < 1e-4 * float(((pe - te) ** 2).sum()) + 1e-3
    assert abs(float(met[2]) - float((pe - te).abs().sum())) < 1e-4 * float((pe - te).abs().sum()) + 1e-3
    assert abs(float(met[3]) - float(((y.cpu() > 0.5) == (te > 0.5)).sum())) <= 2      # (a prediction within float32 rounding of 0.5)
    assert abs(float(met[4]) - float((y.cpu().argmax(-1) == te.argmax(-1)).sum())) <= 2


# ------------------------------------------------------------------------------------------------------------------
# fused attention gate (csrc/gate.cu) against the float64 restatement the CPU emulator uses (tests/desc_emulator.py:_gate_forward)
@pytest.mark.parametrize("N,h,w,C,Cs,training", [(2, 8, 8, 64, 64, 1), (3, 5, 7, 16, 8, 1), (2, 16, 12, 128, 256, 1), (1, 4, 4, 1024, 512, 1),
                                                  (2, 33, 9, 8, 32, 1), (2, 8, 8, 64, 64, 0)])
def test_fused_attention_gate(N, h, w, C, Cs, training):
    import types
    from desc_emulator import _gate_forward
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(7 * C + h)
    rnd = lambda *s: torch.randn(*s, generator=g)
    za, zb = bf(rnd(N, h, w, C).to(dev)), bf((rnd(N, h, w, C) * 0.7 + 0.2).to(dev))
    skipb = bf(rnd(N, 2 * h, 2 * w, 3 * Cs).to(dev))             # the skip and the output are channel windows of wider buffers
    outb = torch.full((N, 2 * h, 2 * w, 2 * Cs), 7.0, device=dev, dtype=torch.bfloat16)
    prm = dict(gamma_a=1 + 0.2 * rnd(C), beta_a=0.1 * rnd(C), mm_a=0.1 * rnd(C), mv_a=1 + 0.2 * torch.rand(C, generator=g),
               gamma_b=1 + 0.2 * rnd(C), beta_b=0.1 * rnd(C), mm_b=0.1 * rnd(C), mv_b=1 + 0.2 * torch.rand(C, generator=g),
               w3=rnd(C) / C ** 0.5, b3=0.1 * rnd(1), gamma3=1 + 0.2 * rnd(1), beta3=0.1 * rnd(1), mm3=0.1 * rnd(1),
               mv3=1 + 0.2 * torch.rand(1, generator=g), wt=0.5 * rnd(16), bt=0.1 * rnd(1))
    stride = 8
    P = {k: v.to(dev).float().contiguous() for k, v in prm.items()}
    wt_dev = torch.zeros(16 * stride, device=dev)
    wt_dev[::stride] = P["wt"]
    moving0 = {k: P[k].clone() for k in ("mm_a", "mv_a", "mm_b", "mv_b", "mm3", "mv3")}
    sums_a = torch.cat([za.float().reshape(-1, C).sum(0), (za.float() ** 2).reshape(-1, C).sum(0)]).contiguous()
    sums_b = torch.cat([zb.float().reshape(-1, C).sum(0), (zb.float() ** 2).reshape(-1, C).sum(0)]).contiguous()
    sums3 = torch.zeros(8, device=dev)
    vec = torch.zeros(8 * C, device=dev)
    z = torch.zeros(N * h * w, device=dev)
    mmap = torch.zeros(N * h * w, device=dev)
    d = L.GateDesc()
    d.za, d.zb = tv(za).to_c(), tv(zb).to_c()
    d.sums_a, d.sums_b, d.sums3 = sums_a.data_ptr(), sums_b.data_ptr(), sums3.data_ptr()
    for k_ in ("gamma_a", "beta_a", "mm_a", "mv_a", "gamma_b", "beta_b", "mm_b", "mv_b", "w3", "b3", "gamma3", "beta3", "mm3", "mv3", "bt"):
        setattr(d, k_, P[k_].data_ptr())
    d.wt, d.wt_stride = wt_dev.data_ptr(), stride
    d.vec_a, d.vec_b, d.z, d.m = vec.data_ptr(), vec.data_ptr() + 16 * C, z.data_ptr(), mmap.data_ptr()
    d.training, d.bessel, d.eps, d.momentum, d.count = training, 1, 1e-3, 0.99, float(N * h * w)
    d.skip, d.out = tv(skipb, Cs, Cs).to_c(), tv(outb, Cs, Cs).to_c()
    L.call("b2seg_gate_fwd", d, stream())
    torch.cuda.synchronize()
    # ---- reference
    cfg = types.SimpleNamespace(training=training, eps=1e-3)
    R = {k: v.double().clone().requires_grad_(not k.startswith("m")) for k, v in prm.items()}
    za64 = za.cpu().double().requires_grad_(True)
    zb64 = zb.cpu().double().requires_grad_(True)
    sk64 = skipb[..., Cs:2 * Cs].cpu().double().requires_grad_(True)
    out64, z64, st = _gate_forward(None, cfg, za64, zb64, sk64, R)
    assert rel_l2(z.cpu().double(), z64.detach().reshape(-1)) < 2e-5
    assert rel_l2(outb[..., Cs:].float().cpu().double(), out64.detach()) < 3e-3
    assert torch.all(outb[..., :Cs] == 7.0)
    if training:
        n = N * h * w
        for tag, key in (("_a", "a"), ("_b", "b"), ("3", "c")):
            mean, var = st[key]
            assert torch.allclose(P["mm" + tag].cpu().double(), moving0["mm" + tag].cpu().double() * 0.99 + mean.detach() * 0.01, atol=1e-5)
            assert torch.allclose(P["mv" + tag].cpu().double(), moving0["mv" + tag].cpu().double() * 0.99 + var.detach() * n / (n - 1) * 0.01, atol=1e-5)
        assert abs(float(sums3[0]) - float(z64.sum())) < 1e-3 * (1 + abs(float(z64.sum()))) and abs(float(sums3[1]) - float((z64 ** 2).sum())) < 1e-3 * float((z64 ** 2).sum())
    else:
        assert all(torch.equal(P[k_], moving0[k_]) for k_ in moving0)
        return
    # ---- backward
    doutb = bf(rnd(N, 2 * h, 2 * w, 2 * Cs).to(dev))
    dskipb = torch.full((N, 2 * h, 2 * w, Cs), 3.0, device=dev, dtype=torch.bfloat16)
    dza, dzb = torch.zeros_like(za), torch.zeros_like(zb)
    G = {k_: torch.zeros_like(P[k_]) for k_ in ("gamma_a", "beta_a", "gamma_b", "beta_b", "gamma3", "beta3", "w3", "b3", "bt")}
    dwt = torch.zeros(16 * stride, device=dev)
    scr = torch.zeros(N * 4 * h * w + N * h * w + 8 + 3 * C, device=dev)
    d.dout, d.dskip, d.dza, d.dzb = tv(doutb, 0, Cs).to_c(), tv(dskipb).to_c(), tv(dza).to_c(), tv(dzb).to_c()
    d.dr, d.g3 = scr.data_ptr(), scr.data_ptr() + 4 * N * 4 * h * w
    d.bsums3 = d.g3 + 4 * N * h * w
    d.bsums_ab = d.bsums3 + 32
    for k_ in ("gamma_a", "beta_a", "gamma_b", "beta_b", "gamma3", "beta3"):
        setattr(d, "d" + k_, G[k_].data_ptr())
    d.dw3, d.db3, d.dwt, d.dbt = G["w3"].data_ptr(), G["b3"].data_ptr(), dwt.data_ptr(), G["bt"].data_ptr()
    for rep in range(2):              # the scratch sums are zeroed by the op itself; dw3 / db3 / dwt / dbt accumulate
        L.call("b2seg_gate_bwd", d, stream())
    torch.cuda.synchronize()
    out64.backward(doutb[..., :Cs].cpu().double())
    assert rel_l2(dskipb.float().cpu().double(), sk64.grad) < 4e-3
    assert rel_l2(dza.float().cpu().double(), za64.grad) < 6e-3, rel_l2(dza.float().cpu().double(), za64.grad)
    assert rel_l2(dzb.float().cpu().double(), zb64.grad) < 6e-3, rel_l2(dzb.float().cpu().double(), zb64.grad)
    for k_ in ("gamma_a", "beta_a", "gamma_b", "beta_b", "gamma3", "beta3"):
        assert rel_l2(G[k_].cpu().double(), R[k_].grad) < 2e-3, (k_, rel_l2(G[k_].cpu().double(), R[k_].grad))
    for k_ in ("w3", "bt"):
        assert rel_l2(G[k_].cpu().double(), 2 * R[k_].grad) < 2e-3, (k_, rel_l2(G[k_].cpu().double(), 2 * R[k_].grad))
    # b3 sits in front of a BatchNorm: analytically zero (the sum of a BatchNorm-backward output); the kernel's value is its rounding noise
    assert abs(float(G["b3"])) < 1e-3 * float(G["w3"].abs().max()) * C ** 0.5 + 1e-4
    assert rel_l2(dwt[::stride].cpu().double(), 2 * R["wt"].grad) < 2e-3 and float(dwt.view(16, stride)[:, 1:].abs().max()) == 0


@pytest.mark.parametrize("N,H,W,Cc,Ctot,off,use_add", [(2, 16, 16, 72, 72, 0, True), (3, 7, 9, 16, 40, 8, False), (2, 32, 8, 64, 64, 0, True), (1, 5, 5, 24, 72, 48, False)])
def test_bn_apply_with_fused_add_and_output_statistics(N, H, W, Cc, Ctot, off, use_add):
    """MultiResBlock / ResPath glue: y = ReLU(x * scale + shift + add) with the column sums of y (as stored) added into an accumulator
    row pair of a WIDER tensor (pitch = Ctot, this tensor is the channel window [off, off + Cc)); then the BatchNorm backward of a
    layer whose input is that ReLU folds the ReLU mask in (x_relu_mask)"""
    dev = "cuda"
    x = bf(torch.randn(N, H, W, Cc, device=dev))
    add = bf(torch.randn(N, H, W, Cc, device=dev)) if use_add else None
    scale, shift = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.2
    wide = torch.full((N, H, W, Ctot), 5.0, device=dev, dtype=torch.bfloat16)
    acc = torch.zeros(2 * Ctot, device=dev)
    d = L.BnActDesc()
    d.x, d.scale, d.shift, d.act, d.n_out = tv(x).to_c(), scale.data_ptr(), shift.data_ptr(), L.ACT_RELU, 1
    d.out[0] = tv(wide, off, Cc).to_c()
    if use_add:
        d.add = tv(add).to_c()
    d.out_stats, d.out_stats_pitch = acc.data_ptr() + 4 * off, Ctot
    for _ in range(2):                      # accumulates: two launches = twice the sums
        L.call("b2seg_bn_act", d, stream())
    torch.cuda.synchronize()
    y = F.relu(x.float() * scale + shift + (add.float() if use_add else 0.0))
    got = wide[..., off:off + Cc]
    assert rel_l2(got.float(), y) < 4e-3
    if off:
        assert torch.all(wide[..., :off] == 5.0)
    gs = got.float().reshape(-1, Cc)
    assert torch.allclose(acc[off:off + Cc], 2 * gs.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(acc[Ctot + off:Ctot + off + Cc], 2 * (gs * gs).sum(0), rtol=1e-4, atol=1e-2)
    assert float(acc[:off].abs().max() if off else 0.0) == 0.0
    # ---- BatchNorm backward of the layer that reads got (= ReLU output): dx masked by got > 0
    s = got.contiguous()
    sf = s.float().reshape(-1, Cc)
    gamma, beta = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
    mean, var = sf.mean(0), sf.var(0, unbiased=False)
    rstd = torch.rsqrt(var + 1e-3)
    sc2, sh2 = (gamma * rstd).contiguous(), (beta - mean * gamma * rstd).contiguous()
    g = bf(torch.randn(N, H, W, Cc, device=dev))
    dgamma, dbeta, dx = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev), torch.zeros_like(s)
    part = torch.zeros(32 * 2 * Cc, device=dev)
    bd = L.BnBwdDesc()
    bd.x, bd.scale, bd.shift, bd.mean, bd.rstd = tv(s).to_c(), sc2.data_ptr(), sh2.data_ptr(), mean.contiguous().data_ptr(), rstd.contiguous().data_ptr()
    bd.act, bd.n_src = L.ACT_NONE, 1
    bd.src[0] = L.GradSrc(tv(g).to_c(), 0, 1, 1)
    bd.count, bd.partials, bd.n_blocks = float(N * H * W), part.data_ptr(), 32
    bd.dgamma, bd.dbeta, bd.dx, bd.x_relu_mask = dgamma.data_ptr(), dbeta.data_ptr(), tv(dx).to_c(), 1
    L.call("b2seg_bn_bwd", bd, stream())
    torch.cuda.synchronize()
    pre = (x.float() * scale + shift + (add.float() if use_add else 0.0)).requires_grad_(True)   # gradient in front of the ReLU
    st = F.relu(pre)
    # the forward values are the stored ones: straight-through so that the statistics match what the kernel saw
    st = st + (s.float() - st).detach()
    mu_, var_ = st.reshape(-1, Cc).mean(0), st.reshape(-1, Cc).var(0, unbiased=False)
    ((st - mu_) * torch.rsqrt(var_ + 1e-3) * gamma + beta).mul(g.float()).sum().backward()
    assert rel_l2(dx.float(), pre.grad) < 8e-3, rel_l2(dx.float(), pre.grad)


def test_fused_attention_gate_second_backward_pass():
    """two-pass form of b2seg_gate_bwd: pass 1 without dskip, pass 2 writes dskip = dout * r + the projection's input gradient at the
    even pixels; must equal the one-pass dskip plus the scattered tensor"""
    import types
    from desc_emulator import _gate_forward
    dev = "cuda"
    N, h, w, C, Cs = 2, 12, 10, 32, 64
    g = torch.Generator(device="cpu").manual_seed(5)
    rnd = lambda *s: torch.randn(*s, generator=g)
    za, zb = bf(rnd(N, h, w, C).to(dev)), bf(rnd(N, h, w, C).to(dev))
    skip = bf(rnd(N, 2 * h, 2 * w, Cs).to(dev))
    out = torch.zeros_like(skip)
    prm = dict(gamma_a=1 + 0.2 * rnd(C), beta_a=0.1 * rnd(C), mm_a=torch.zeros(C), mv_a=torch.ones(C), gamma_b=1 + 0.2 * rnd(C), beta_b=0.1 * rnd(C),
               mm_b=torch.zeros(C), mv_b=torch.ones(C), w3=rnd(C) / C ** 0.5, b3=0.1 * rnd(1), gamma3=1 + 0.2 * rnd(1), beta3=0.1 * rnd(1),
               mm3=torch.zeros(1), mv3=torch.ones(1), wt=0.5 * rnd(16), bt=0.1 * rnd(1))
    P = {k: v.to(dev).float().contiguous() for k, v in prm.items()}
    sums_a = torch.cat([za.float().reshape(-1, C).sum(0), (za.float() ** 2).reshape(-1, C).sum(0)]).contiguous()
    sums_b = torch.cat([zb.float().reshape(-1, C).sum(0), (zb.float() ** 2).reshape(-1, C).sum(0)]).contiguous()
    sums3, vec, z, mmap = torch.zeros(8, device=dev), torch.zeros(8 * C, device=dev), torch.zeros(N * h * w, device=dev), torch.zeros(N * h * w, device=dev)
    d = L.GateDesc()
    d.za, d.zb, d.sums_a, d.sums_b, d.sums3 = tv(za).to_c(), tv(zb).to_c(), sums_a.data_ptr(), sums_b.data_ptr(), sums3.data_ptr()
    for k_ in ("gamma_a", "beta_a", "mm_a", "mv_a", "gamma_b", "beta_b", "mm_b", "mv_b", "w3", "b3", "gamma3", "beta3", "mm3", "mv3", "bt", "wt"):
        setattr(d, k_, P[k_].data_ptr())
    d.wt_stride, d.vec_a, d.vec_b, d.z, d.m = 1, vec.data_ptr(), vec.data_ptr() + 16 * C, z.data_ptr(), mmap.data_ptr()
    d.training, d.bessel, d.eps, d.momentum, d.count = 1, 1, 1e-3, 0.99, float(N * h * w)
    d.skip, d.out = tv(skip).to_c(), tv(out).to_c()
    L.call("b2seg_gate_fwd", d, stream())
    dout, da_low = bf(rnd(N, 2 * h, 2 * w, Cs).to(dev)), bf(rnd(N, h, w, Cs).to(dev))
    dskip1, dskip2, dza, dzb = torch.zeros_like(skip), torch.zeros_like(skip), torch.zeros_like(za), torch.zeros_like(zb)
    G = torch.zeros(6 * C + 64, device=dev)
    scr = torch.zeros(N * 4 * h * w + N * h * w + 8 + 3 * C, device=dev)
    d.dout, d.dza, d.dzb = tv(dout).to_c(), tv(dza).to_c(), tv(dzb).to_c()
    d.dr, d.g3 = scr.data_ptr(), scr.data_ptr() + 4 * N * 4 * h * w
    d.bsums3 = d.g3 + 4 * N * h * w
    d.bsums_ab = d.bsums3 + 32
    base = G.data_ptr()
    d.dgamma_a, d.dbeta_a, d.dgamma_b, d.dbeta_b, d.dw3 = base, base + 4 * C, base + 8 * C, base + 12 * C, base + 16 * C
    d.dgamma3, d.dbeta3, d.dwt, d.dbt = base + 20 * C, base + 20 * C + 4, base + 20 * C + 64, base + 20 * C + 160
    one = L.GateDesc.from_buffer_copy(d)
    one.dskip = tv(dskip1).to_c()
    L.call("b2seg_gate_bwd", one, stream())                    # one-pass form
    dza1 = dza.clone()
    L.call("b2seg_gate_bwd", d, stream())                      # pass 1 of the two-pass form: no dskip
    two = L.GateDesc.from_buffer_copy(d)
    two.dskip, two.da_low = tv(dskip2).to_c(), tv(da_low).to_c()
    L.call("b2seg_gate_bwd", two, stream())
    torch.cuda.synchronize()
    assert torch.equal(dza, dza1)
    want = dskip1.float()
    want[:, ::2, ::2] += da_low.float()
    assert rel_l2(dskip2.float(), want) < 6e-3, rel_l2(dskip2.float(), want)
    assert float(dskip2.float().abs().max()) > 0
