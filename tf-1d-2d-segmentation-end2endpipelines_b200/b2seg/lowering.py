"""Descriptor builders: Keras layer semantics -> b2seg C-ABI descriptors (include/b2seg.h).

Everything here is address arithmetic on NHWC views; no tensor library is involved.  The tap tables encode the
reference's layer semantics:

* Conv2D / Conv1D, padding 'same' (reference 2DCNN/models/unet_variants.py:9, 1DCNN/Models/unet_variants.py:55):
  tap (i, j) reads the input at offset (i - (kh-1)//2, j - (kw-1)//2) — TensorFlow's SAME split puts the smaller
  half of the padding first.
* Conv2DTranspose(4x4, strides 2, 'same') (unet_variants.py:19): output o = 2*i - 1 + k, i.e. four output-parity
  classes, each a 2x2-tap convolution of the input.  Conv1DTranspose(2, strides 2) (1DCNN :104): o = 2*i + k.
* Their gradients w.r.t. input (TF Conv2DBackpropInput / strided Conv2D) and filter (Conv2DBackpropFilter).
"""
from __future__ import annotations

from dataclasses import dataclass, replace

from . import _lib as L


@dataclass(frozen=True)
class TView:
    """A strided NHWC window (element strides, channel stride 1) into a bf16 buffer."""
    ptr: int
    N: int
    H: int
    W: int
    C: int
    sn: int
    sh: int
    sw: int
    esize: int = 2

    @staticmethod
    def dense(ptr, N, H, W, C, esize=2):
        return TView(ptr, N, H, W, C, H * W * C, W * C, C, esize)

    def chan(self, c_off, C):
        assert 0 <= c_off and c_off + C <= self.C, (c_off, C, self.C)
        return replace(self, ptr=self.ptr + c_off * self.esize, C=C)

    def parity(self, a, b, fh=2, fw=2):
        """Sub-grid of pixels (fh*i + a, fw*j + b)."""
        return replace(self, ptr=self.ptr + (a * self.sh + b * self.sw) * self.esize,
                       H=(self.H - a + fh - 1) // fh, W=(self.W - b + fw - 1) // fw, sh=self.sh * fh, sw=self.sw * fw)

    def to_c(self):
        return L.View(self.ptr, self.N, self.H, self.W, self.C, self.sn, self.sh, self.sw)


NULL_VIEW = TView(0, 0, 0, 0, 0, 0, 0, 0)


def _set_taps(desc, taps):
    for i, (src, dh, dw, widx) in enumerate(taps):
        desc.taps[i] = L.Tap(src, dh, dw, widx)


def conv_fprop(x: TView, w_ptr, cout, kh, kw, cin, out: TView, bias=0, act=L.ACT_NONE, stats=0, block_n=0):
    """Conv, stride 1, SAME.  Weights bf16 [cout][kh*kw][cin]; out.C == cout (padded)."""
    d = L.ConvDesc()
    d.n_src = 1
    d.src[0] = x.to_c()
    d.weights, d.w_cout, d.w_taps, d.w_cin = w_ptr, cout, kh * kw, cin
    d.b_mn_major = 0
    d.n_groups, d.taps_per_group = 1, kh * kw
    _set_taps(d, [(0, i - (kh - 1) // 2, j - (kw - 1) // 2, i * kw + j) for i in range(kh) for j in range(kw)])
    d.out[0] = out.to_c()
    d.bias, d.act, d.stats, d.block_n = bias, act, stats, block_n
    return d


def conv_dgrad(dy: TView, w_ptr, cout, kh, kw, cin, dx: TView, mul_view: TView | None = None, mul_mode=0, block_n=0):
    """dx = Conv2DBackpropInput(dy): dx[p] = sum_t dy[p - d_t] . W_t^T; optional dx *= act'(mul_view)."""
    d = L.ConvDesc()
    d.n_src = 1
    d.src[0] = dy.to_c()
    d.weights, d.w_cout, d.w_taps, d.w_cin = w_ptr, cout, kh * kw, cin
    d.b_mn_major = 1
    d.n_groups, d.taps_per_group = 1, kh * kw
    _set_taps(d, [(0, -(i - (kh - 1) // 2), -(j - (kw - 1) // 2), i * kw + j) for i in range(kh) for j in range(kw)])
    d.out[0] = dx.to_c()
    d.block_n = block_n
    if mul_view is not None and mul_mode:
        d.mul_view, d.mul_mode = mul_view.to_c(), mul_mode
    return d


def conv_wgrad(dy: TView, x: TView, dw_ptr, cout, kh, kw, cin, ksplit=0):
    d = L.WgradDesc()
    d.n_pair = 1
    d.dy[0], d.x[0] = dy.to_c(), x.to_c()
    d.gN, d.gH, d.gW = dy.N, dy.H, dy.W
    d.n_taps = kh * kw
    for i in range(kh):
        for j in range(kw):
            d.taps[i * kw + j] = L.WgradTap(0, 0, 0, i - (kh - 1) // 2, j - (kw - 1) // 2, i * kw + j)
    d.dw, d.w_cout, d.w_taps, d.w_cin = dw_ptr, cout, kh * kw, cin
    d.ksplit = ksplit
    return d


def conv_s2_fprop(x: TView, w_ptr, cout, cin, out: TView, bias=0, act=L.ACT_NONE, stats=0):
    """Conv 1x1, strides 2, padding 'valid' (attention gate, unet_variants.py:69): reads pixels (2i, 2j)."""
    return conv_fprop(x.parity(0, 0, 2 if x.H > 1 else 1, 2), w_ptr, cout, 1, 1, cin, out, bias, act, stats)


# ---- transposed convolution --------------------------------------------------------------------------------
def _tconv_axis(k, s):
    """For one axis: list over output parity a of [(kernel index, input offset)].
    k=4,s=2,'same': o = 2i - 1 + kk ; k=2,s=2: o = 2i + kk ; k=1 (degenerate H axis of 1D): o = i."""
    if k == 1:
        return [[(0, 0)]]
    if k == 4 and s == 2:
        return [[(1, 0), (3, -1)], [(0, 1), (2, 0)]]
    if k == 2 and s == 2:
        return [[(0, 0)], [(1, 0)]]
    raise ValueError(f"unsupported transposed conv kernel {k} stride {s}")


def tconv_fprop(x: TView, w_ptr, cout, kh, kw, cin, out: TView, bias=0, act=L.ACT_NONE, stats=0, block_n=0):
    """Conv2DTranspose(kh x kw, strides 2 (1 on a k=1 axis), 'same').  out is the full (upsampled) view."""
    ah, aw = _tconv_axis(kh, 2), _tconv_axis(kw, 2)
    fh, fw = len(ah), len(aw)
    d = L.ConvDesc()
    d.n_src = 1
    d.src[0] = x.to_c()
    d.weights, d.w_cout, d.w_taps, d.w_cin = w_ptr, cout, kh * kw, cin
    d.b_mn_major = 0
    d.n_groups = fh * fw
    taps = []
    for a in range(fh):
        for b in range(fw):
            g = a * fw + b
            grp = [(0, dh, dw, ki * kw + kj) for (ki, dh) in ah[a] for (kj, dw) in aw[b]]
            taps.append(grp)
            ov = out.parity(a, b, fh, fw)
            assert (ov.N, ov.H, ov.W) == (x.N, x.H, x.W), (ov, x)
            d.out[g] = ov.to_c()
    d.taps_per_group = len(taps[0])
    _set_taps(d, [t for grp in taps for t in grp])
    d.bias, d.act, d.stats, d.block_n = bias, act, stats, block_n
    return d


def tconv_dgrad(dy: TView, w_ptr, cout, kh, kw, cin, dx: TView, block_n=0):
    """dx[i] = sum_k dy[2i - 1 + k] . W_k^T  (a stride-2 convolution of dy), via the parity sub-grids of dy."""
    ah, aw = _tconv_axis(kh, 2), _tconv_axis(kw, 2)
    fh, fw = len(ah), len(aw)
    d = L.ConvDesc()
    d.n_src = fh * fw
    taps = []
    for a in range(fh):
        for b in range(fw):
            d.src[a * fw + b] = dy.parity(a, b, fh, fw).to_c()
            # forward: out(parity a) at m reads x[m + dh]  =>  x at i gets dy(parity a) at i - dh
            taps += [(a * fw + b, -dh, -dw, ki * kw + kj) for (ki, dh) in ah[a] for (kj, dw) in aw[b]]
    d.weights, d.w_cout, d.w_taps, d.w_cin = w_ptr, cout, kh * kw, cin
    d.b_mn_major = 1
    d.n_groups, d.taps_per_group = 1, len(taps)
    _set_taps(d, taps)
    d.out[0] = dx.to_c()
    d.block_n = block_n
    return d


def tconv_wgrad(dy: TView, x: TView, dw_ptr, cout, kh, kw, cin, ksplit=0):
    ah, aw = _tconv_axis(kh, 2), _tconv_axis(kw, 2)
    fh, fw = len(ah), len(aw)
    d = L.WgradDesc()
    d.n_pair = fh * fw
    t = 0
    for a in range(fh):
        for b in range(fw):
            p = a * fw + b
            d.dy[p], d.x[p] = dy.parity(a, b, fh, fw).to_c(), x.to_c()
            for (ki, dh) in ah[a]:
                for (kj, dw) in aw[b]:
                    # dW_k = sum_m dy_par[m] * x[m + dh]
                    d.taps[t] = L.WgradTap(p, 0, 0, dh, dw, ki * kw + kj)
                    t += 1
    d.gN, d.gH, d.gW = x.N, x.H, x.W
    d.n_taps = t
    d.dw, d.w_cout, d.w_taps, d.w_cin = dw_ptr, cout, kh * kw, cin
    d.ksplit = ksplit
    return d
