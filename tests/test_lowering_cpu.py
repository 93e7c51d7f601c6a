"""CPU checks of the descriptor builders: emulate the C-ABI conv/wgrad semantics and compare with torch."""
import pytest
import torch
import torch.nn.functional as F

from b2seg import lowering as lw
from desc_emulator import FakeMem, run_conv, run_wgrad


def _mk(mem, N, H, W, C, data=None):
    base, t = mem.alloc(N * H * W * C)
    v = lw.TView.dense(base, N, H, W, C)
    if data is not None:
        mem.write_view(v.to_c(), data)
    return v


def _w_internal(mem, w_oihw_like):  # [cout][taps][cin]
    base, t = mem.alloc(w_oihw_like.numel())
    t[:] = w_oihw_like.reshape(-1).double()
    return base


@pytest.mark.parametrize("kh,kw,H,W", [(3, 3, 6, 8), (1, 1, 4, 4), (1, 3, 1, 16), (1, 4, 1, 16), (1, 5, 1, 8)])
def test_conv_fprop_dgrad_wgrad_taps(kh, kw, H, W):
    torch.manual_seed(0)
    N, Cin, Cout = 2, 8, 16
    x = torch.randn(N, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, kh, kw, dtype=torch.float64, requires_grad=True)
    # TF SAME: left pad (k-1)//2, right pad k-1-(k-1)//2
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xp = F.pad(x, (pw, kw - 1 - pw, ph, kh - 1 - ph))
    y = F.conv2d(xp, w)
    gy = torch.randn_like(y)
    y.backward(gy)

    mem = FakeMem()
    xv = _mk(mem, N, H, W, Cin, x.detach().permute(0, 2, 3, 1))
    wp = _w_internal(mem, w.detach().permute(0, 2, 3, 1))
    yv = _mk(mem, N, H, W, Cout)
    run_conv(mem, lw.conv_fprop(xv, wp, Cout, kh, kw, Cin, yv))
    got = mem.gather_view(yv.to_c()).permute(0, 3, 1, 2)
    assert torch.allclose(got, y.detach(), atol=1e-9)

    gyv = _mk(mem, N, H, W, Cout, gy.permute(0, 2, 3, 1))
    dxv = _mk(mem, N, H, W, Cin)
    run_conv(mem, lw.conv_dgrad(gyv, wp, Cout, kh, kw, Cin, dxv))
    assert torch.allclose(mem.gather_view(dxv.to_c()).permute(0, 3, 1, 2), x.grad, atol=1e-9)

    dwp, dwt = mem.alloc(w.numel(), esize=4)
    run_wgrad(mem, lw.conv_wgrad(gyv, xv, dwp, Cout, kh, kw, Cin))
    assert torch.allclose(dwt.view(Cout, kh, kw, Cin).permute(0, 3, 1, 2), w.grad, atol=1e-9)


@pytest.mark.parametrize("kh,kw,H,W,pad", [(4, 4, 4, 6, 1), (1, 2, 1, 8, 0)])
def test_tconv_taps(kh, kw, H, W, pad):
    torch.manual_seed(1)
    N, Cin, Cout = 2, 8, 8
    sh = 2 if kh > 1 else 1
    x = torch.randn(N, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cin, Cout, kh, kw, dtype=torch.float64, requires_grad=True)  # torch ConvTranspose layout
    y = F.conv_transpose2d(x, w, stride=(sh, 2), padding=(pad if kh > 1 else 0, pad))
    assert y.shape[2:] == (H * sh, W * 2)
    gy = torch.randn_like(y)
    y.backward(gy)

    mem = FakeMem()
    xv = _mk(mem, N, H, W, Cin, x.detach().permute(0, 2, 3, 1))
    wp = _w_internal(mem, w.detach().permute(1, 2, 3, 0))  # [cout][kh*kw][cin]
    yv = _mk(mem, N, H * sh, W * 2, Cout)
    run_conv(mem, lw.tconv_fprop(xv, wp, Cout, kh, kw, Cin, yv))
    assert torch.allclose(mem.gather_view(yv.to_c()).permute(0, 3, 1, 2), y.detach(), atol=1e-9)

    gyv = _mk(mem, N, H * sh, W * 2, Cout, gy.permute(0, 2, 3, 1))
    dxv = _mk(mem, N, H, W, Cin)
    run_conv(mem, lw.tconv_dgrad(gyv, wp, Cout, kh, kw, Cin, dxv))
    assert torch.allclose(mem.gather_view(dxv.to_c()).permute(0, 3, 1, 2), x.grad, atol=1e-9)

    dwp, dwt = mem.alloc(w.numel(), esize=4)
    run_wgrad(mem, lw.tconv_wgrad(gyv, xv, dwp, Cout, kh, kw, Cin))
    assert torch.allclose(dwt.view(Cout, kh, kw, Cin).permute(3, 0, 1, 2), w.grad, atol=1e-9)


def test_strided_1x1_valid():
    torch.manual_seed(2)
    N, Cin, Cout, H, W = 1, 8, 8, 6, 6
    x = torch.randn(N, Cin, H, W, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 1, 1, dtype=torch.float64)
    y = F.conv2d(x, w, stride=2)
    mem = FakeMem()
    xv = _mk(mem, N, H, W, Cin, x.permute(0, 2, 3, 1))
    wp = _w_internal(mem, w.permute(0, 2, 3, 1))
    yv = _mk(mem, N, H // 2, W // 2, Cout)
    run_conv(mem, lw.conv_s2_fprop(xv, wp, Cout, Cin, yv))
    assert torch.allclose(mem.gather_view(yv.to_c()).permute(0, 3, 1, 2), y, atol=1e-9)
