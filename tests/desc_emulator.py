"""CPU emulation of the b2seg conv / wgrad descriptor semantics (include/b2seg.h) on torch tensors.

Test infrastructure only: it lets the tap tables produced by b2seg.lowering be checked against
torch.nn.functional on the CPU, without a GPU.  Views are resolved against a dict {base_ptr: flat fp32 tensor}
that fakes device memory (element size 2 like bf16 addresses).
"""
import torch


import os

# B2SEG_EMU_BF16=1: every tensor written into a 2-byte (bf16 on the device) buffer is rounded to bf16 — a CPU model of the device's
# storage rounding, used to calibrate the tolerances of the GPU tests where no GPU exists
ROUND_BF16 = os.environ.get("B2SEG_EMU_BF16", "")     # "1": activations and gradients; "act" / "grad": only those buffers (diagnosis)


class FakeMem:
    def __init__(self):
        self.bufs = []  # (base, nbytes, tensor)
        self.next = 1 << 20

    def alloc(self, numel, esize=2):
        base = self.next
        t = torch.zeros(numel, dtype=torch.float64)
        self.bufs.append((base, numel * esize, t, esize))
        self.next += ((numel * esize + 1023) // 1024 + 1) * 1024
        return base, t

    def tag_of(self, ptr):
        for base, nbytes, t, esize in self.bufs:
            if base <= ptr < base + nbytes:
                return getattr(self, "tags", {}).get(base, "act")
        raise KeyError(ptr)

    def esize_of(self, ptr):
        for base, nbytes, t, esize in self.bufs:
            if base <= ptr < base + nbytes:
                return esize
        raise KeyError(ptr)

    def resolve(self, ptr):
        for base, nbytes, t, esize in self.bufs:
            if base <= ptr < base + nbytes:
                return t, (ptr - base) // esize
        raise KeyError(ptr)

    def read_view(self, v, n, h, w):
        """channel vector at pixel (n,h,w) or zeros if outside"""
        if not (0 <= n < v.N and 0 <= h < v.H and 0 <= w < v.W):
            return torch.zeros(v.C, dtype=torch.float64)
        t, off = self.resolve(v.ptr)
        o = off + n * v.sn + h * v.sh + w * v.sw
        return t[o:o + v.C]

    def gather_view(self, v):
        t, off = self.resolve(v.ptr)
        idx = (off + torch.arange(v.N).view(-1, 1, 1, 1) * v.sn + torch.arange(v.H).view(1, -1, 1, 1) * v.sh
               + torch.arange(v.W).view(1, 1, -1, 1) * v.sw + torch.arange(v.C).view(1, 1, 1, -1))
        return t[idx]

    def shifted(self, v, dh, dw, N, H, W):
        """tensor [N,H,W,C] of v sampled at (h+dh, w+dw) with zero fill"""
        full = self.gather_view(v)
        out = torch.zeros(N, H, W, v.C, dtype=torch.float64)
        for h in range(H):
            hs = h + dh
            if not 0 <= hs < v.H:
                continue
            w_lo, w_hi = max(0, -dw), min(W, v.W - dw)
            if w_hi > w_lo:
                out[:min(N, v.N), h, w_lo:w_hi] = full[:N, hs, w_lo + dw:w_hi + dw]
        return out

    def write_view(self, v, data):
        t, off = self.resolve(v.ptr)
        if ROUND_BF16 and self.esize_of(v.ptr) == 2 and (ROUND_BF16 == "1" or ROUND_BF16 == self.tag_of(v.ptr)):
            data = data.to(torch.float32).to(torch.bfloat16)    # numerics model of the device: stored activations / gradients are bf16
        idx = (off + torch.arange(v.N).view(-1, 1, 1, 1) * v.sn + torch.arange(v.H).view(1, -1, 1, 1) * v.sh
               + torch.arange(v.W).view(1, 1, -1, 1) * v.sw + torch.arange(v.C).view(1, 1, 1, -1))
        t[idx] = data.to(torch.float64)


def run_conv(mem: FakeMem, d):
    wt, woff = mem.resolve(d.weights)
    Wm = wt[woff:woff + d.w_cout * d.w_taps * d.w_cin].view(d.w_cout, d.w_taps, d.w_cin)
    for g in range(d.n_groups):
        o = d.out[g]
        acc = torch.zeros(o.N, o.H, o.W, o.C, dtype=torch.float64)
        for t in range(d.taps_per_group):
            tap = d.taps[g * d.taps_per_group + t]
            xs = mem.shifted(d.src[tap.src], tap.dh, tap.dw, o.N, o.H, o.W)
            if d.b_mn_major == 0:
                acc += xs[..., :d.w_cin] @ Wm[:o.C, tap.widx, :].T
            else:
                acc += xs[..., :d.w_cout] @ Wm[:, tap.widx, :o.C]
        mem.write_view(o, acc)


def run_wgrad(mem: FakeMem, d):
    wt, woff = mem.resolve(d.dw)
    dW = wt[woff:woff + d.w_cout * d.w_taps * d.w_cin].view(d.w_cout, d.w_taps, d.w_cin)
    for t in range(d.n_taps):
        tap = d.taps[t]
        dy = mem.shifted(d.dy[tap.pair], tap.dyh, tap.dyw, d.gN, d.gH, d.gW)
        x = mem.shifted(d.x[tap.pair], tap.dh, tap.dw, d.gN, d.gH, d.gW)
        dW[:, tap.widx, :] += torch.einsum("nhwo,nhwi->oi", dy, x)


# ============================================================================================================
# Whole-plan emulation: every op of include/b2seg.h in float64 on FakeMem.  Used by the CPU tests to check the
# planner (fusion, concat slots, gradient routing, weight layouts) against the oracle without a GPU.
import math

import numpy as np

from b2seg import _lib as L

_ESIZE = {"act": 2, "grad": 2, "param_wb": 2, "arena": 2, "fold_w": 2}


class PlanMem(FakeMem):
    def alloc_bytes(self, nbytes, tag="act"):
        es = _ESIZE.get(tag, 4)
        base, _ = self.alloc(nbytes // es, es)
        if not hasattr(self, "tags"):
            self.tags = {}
        self.tags[base] = tag
        return base

    def f32(self, ptr, n):
        t, off = self.resolve(ptr)
        return t[off:off + n]


def _act(x, code):
    if code == L.ACT_RELU:
        return torch.relu(x)
    if code == L.ACT_LEAKY:
        return torch.where(x > 0, x, 0.3 * x)
    if code == L.ACT_SIGMOID:
        return torch.sigmoid(x)
    if code == L.ACT_TANH:
        return torch.tanh(x)
    return x


def _dact_from_y(y, code):
    if code == L.ACT_RELU:
        return (y > 0).double()
    if code == L.ACT_LEAKY:
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.3))
    if code == L.ACT_SIGMOID:
        return y * (1 - y)
    if code == L.ACT_TANH:
        return 1 - y * y
    return torch.ones_like(y)


def emu_conv(mem, d):
    wt, woff = mem.resolve(d.weights)
    Wm = wt[woff:woff + d.w_cout * d.w_taps * d.w_cin].view(d.w_cout, d.w_taps, d.w_cin)
    tot_s, tot_q = None, None
    for g in range(d.n_groups):
        o = d.out[g]
        acc = torch.zeros(o.N, o.H, o.W, o.C, dtype=torch.float64)
        for t in range(d.taps_per_group):
            tap = d.taps[g * d.taps_per_group + t]
            xs = mem.shifted(d.src[tap.src], tap.dh, tap.dw, o.N, o.H, o.W)
            if d.b_mn_major == 0:
                acc += xs[..., :d.w_cin] @ Wm[:o.C, tap.widx, :xs.shape[-1]].T
            else:
                acc += xs[..., :d.w_cout] @ Wm[:xs.shape[-1], tap.widx, :o.C]
        if d.bias:
            acc = acc + mem.f32(d.bias, o.C)
        acc = _act(acc, d.act)
        if d.mul_mode:
            mv = d.mul_view
            y = mem.gather_view(mv)
            acc[..., :mv.C] = acc[..., :mv.C] * _dact_from_y(y, d.mul_mode)
        mem.write_view(o, acc)
        if d.stats:
            if ROUND_BF16 and mem.esize_of(o.ptr) == 2:
                acc = mem.gather_view(o)        # like the kernel: statistics of exactly what was stored
            s, q = acc.reshape(-1, o.C).sum(0), (acc.reshape(-1, o.C) ** 2).sum(0)
            tot_s = s if tot_s is None else tot_s + s
            tot_q = q if tot_q is None else tot_q + q
    if d.stats:
        C = d.out[0].C
        st = mem.f32(d.stats, 2 * C)   # emulator writes totals into partial 0; the rest stay zero
        if d.stats_atomic:             # one caller-zeroed row every CTA adds into
            st[:C] += tot_s
            st[C:] += tot_q
        else:
            st[:C], st[C:] = tot_s, tot_q


def emu_bn_finalize(mem, d):
    C = d.C
    mm, mv = mem.f32(d.moving_mean, C), mem.f32(d.moving_var, C)
    gamma, beta = mem.f32(d.gamma, C), mem.f32(d.beta, C)
    if d.inference:
        mean, var = mm.clone(), mv.clone()
    else:
        part = mem.f32(d.partials, d.n_partials * 2 * C).view(d.n_partials, 2, C)
        s, q = part[:, 0].sum(0), part[:, 1].sum(0)
        mean = s / d.count
        var = (q / d.count - mean * mean).clamp_min(0)
        if d.update_moving:
            uv = var * d.count / (d.count - 1) if (d.bessel and d.count > 1) else var
            mm[:] = mm * d.momentum + mean * (1 - d.momentum)
            mv[:] = mv * d.momentum + uv * (1 - d.momentum)
    rstd = 1.0 / torch.sqrt(var + d.eps)
    mem.f32(d.scale, C)[:] = gamma * rstd
    mem.f32(d.shift, C)[:] = beta - mean * gamma * rstd
    if d.mean:
        mem.f32(d.mean, C)[:] = mean
    if d.rstd:
        mem.f32(d.rstd, C)[:] = rstd


def emu_bn_act(mem, d):
    x = mem.gather_view(d.x)
    C = d.x.C
    sc = mem.f32(d.scale, C) if d.scale else torch.ones(C, dtype=torch.float64)
    sf = mem.f32(d.shift, C) if d.shift else torch.zeros(C, dtype=torch.float64)
    pre = x * sc + sf
    if d.add.ptr:
        pre = pre + mem.gather_view(d.add)
    y = _act(pre, d.act)
    if d.c_valid:
        y[..., d.c_valid:] = 0
    for i in range(d.n_out):
        mem.write_view(d.out[i], y)
    if d.out_stats:          # sums of the values as stored (bf16 on the device), added into caller-zeroed accumulators
        ys = mem.gather_view(d.out[0]).reshape(-1, C)
        pitch = d.out_stats_pitch or C
        st = mem.f32(d.out_stats, pitch + C)
        st[:C] += ys.sum(0)
        st[pitch:pitch + C] += (ys * ys).sum(0)
    ph, pw = max(d.pool_h, 1), max(d.pool_w, 1)
    if ph > 1 or pw > 1:
        N, H, W, _ = y.shape
        p = y.view(N, H // ph, ph, W // pw, pw, C).amax(dim=(2, 4))
        mem.write_view(d.pooled, p)


def emu_bn_bwd(mem, d):
    x = mem.gather_view(d.x)
    N, H, W, C = x.shape
    has_bn = bool(d.scale)
    sc = mem.f32(d.scale, C) if has_bn else torch.ones(C, dtype=torch.float64)
    sf = mem.f32(d.shift, C) if d.shift else torch.zeros(C, dtype=torch.float64)
    mu = mem.f32(d.mean, C) if d.mean else torch.zeros(C, dtype=torch.float64)
    rs = mem.f32(d.rstd, C) if d.rstd else torch.ones(C, dtype=torch.float64)
    y = _act(x * sc + sf, d.act)
    g = torch.zeros_like(x)
    for i in range(d.n_src):
        s = d.src[i]
        if s.kind == 2:   # folded pointwise head: g = dlogits . W^T; the head's dW, db accumulate here (caller-zeroed slots)
            dl = mem.f32(s.dlogits, N * H * W * s.cout).view(N, H, W, s.cout)
            hw = mem.f32(s.head_w, C * s.cout).view(C, s.cout)
            g += dl @ hw.T
            dw = mem.f32(s.head_dw, C * s.cout).view(C, s.cout)
            dw += torch.einsum("nhwc,nhwo->co", y, dl)
            db = mem.f32(s.head_db, s.cout)
            db += dl.reshape(-1, s.cout).sum(0)
            continue
        gv = mem.gather_view(s.g)
        if s.kind == 0:
            g += gv
        else:
            ph, pw = s.pool_h, s.pool_w
            yw = y.view(N, H // ph, ph, W // pw, pw, C).permute(0, 1, 3, 5, 2, 4).reshape(N, H // ph, W // pw, C, ph * pw)
            first = yw.argmax(dim=-1)  # torch returns the first maximal index
            onehot = torch.nn.functional.one_hot(first, ph * pw).double() * gv.unsqueeze(-1)
            g += onehot.view(N, H // ph, W // pw, C, ph, pw).permute(0, 1, 4, 2, 5, 3).reshape(N, H, W, C)
    g = g * _dact_from_y(y, d.act)
    if has_bn:
        xh = (x - mu) * rs
        dbeta = g.reshape(-1, C).sum(0)
        dgamma = (g * xh).reshape(-1, C).sum(0)
        mem.f32(d.dbeta, C)[:] = dbeta
        mem.f32(d.dgamma, C)[:] = dgamma
        dx = sc * (g - dbeta / d.count - xh * dgamma / d.count)
    else:
        dx = g
    if d.x_relu_mask:
        dx = dx * (x > 0)
    mem.write_view(d.dx, dx)


def emu_wgrad(mem, d):
    run_wgrad(mem, d)


def emu_adam(mem, d):
    n = d.n
    w, g, m, v = mem.f32(d.w, n), mem.f32(d.g, n), mem.f32(d.m, n), mem.f32(d.v, n)
    gg = g * d.grad_scale
    m[:] = d.beta1 * m + (1 - d.beta1) * gg
    v[:] = d.beta2 * v + (1 - d.beta2) * gg * gg
    alpha = d.lr * math.sqrt(1 - d.beta2 ** d.step) / (1 - d.beta1 ** d.step)
    w[:] = w - alpha * m / (v.sqrt() + d.eps)
    wb, off = mem.resolve(d.w_bf16)
    wb[off:off + n] = w


def _head_geom(d):
    st = max(d.stride, 1)
    return st, (d.x.H + st - 1) // st, (d.x.W + st - 1) // st


def emu_head_fwd(mem, d):
    st, Ho, Wo = _head_geom(d)
    x = mem.gather_view(d.x)[:, ::st, ::st]
    if d.bn_scale:          # BatchNorm + activation of the last Conv_Block applied by the head itself
        x = _act(x * mem.f32(d.bn_scale, d.x.C) + mem.f32(d.bn_shift, d.x.C), d.bn_act)
    w = mem.f32(d.w, d.x.C * d.cout).view(d.x.C, d.cout)
    z = x @ w + mem.f32(d.b, d.cout)
    n = z.numel()
    if d.logits:
        mem.f32(d.logits, n)[:] = z.reshape(-1)
    y = torch.softmax(z, -1) if d.act == L.ACT_SOFTMAX else _act(z, d.act)
    mem.f32(d.y, n)[:] = y.reshape(-1)


def emu_head_bwd(mem, d):
    st, Ho, Wo = _head_geom(d)
    xfull = mem.gather_view(d.x)
    x = xfull[:, ::st, ::st]
    w = mem.f32(d.w, d.x.C * d.cout).view(d.x.C, d.cout)
    dl = mem.f32(d.dlogits, x.shape[0] * Ho * Wo * d.cout).view(x.shape[0], Ho, Wo, d.cout)
    if d.dx.ptr:
        dx = torch.zeros_like(xfull)
        dx[:, ::st, ::st] = dl @ w.T
        mem.write_view(d.dx, dx)
    mem.f32(d.dw, d.x.C * d.cout)[:] += torch.einsum("nhwc,nhwo->co", x, dl).reshape(-1)
    mem.f32(d.db, d.cout)[:] += dl.reshape(-1, d.cout).sum(0)


def emu_outact_fwd(mem, d):
    assert d.x.C % 8 == 0 and 1 <= d.cout <= d.x.C
    z = mem.gather_view(d.x)[..., :d.cout]
    y = torch.softmax(z, -1) if d.act == L.ACT_SOFTMAX else _act(z, d.act)
    mem.f32(d.y, y.numel())[:] = y.reshape(-1)


def emu_outact_bwd(mem, d):
    assert d.dx.C % 8 == 0 and 1 <= d.cout <= d.dx.C
    n = d.dx.N * d.dx.H * d.dx.W
    dx = torch.zeros(d.dx.N, d.dx.H, d.dx.W, d.dx.C, dtype=torch.float64)
    dx[..., :d.cout] = mem.f32(d.dlogits, n * d.cout).view(d.dx.N, d.dx.H, d.dx.W, d.cout)
    mem.write_view(d.dx, dx)


def emu_loss(mem, d):
    """float64 mirror of loss_kernel (csrc/stream_kernels.cu): explicit derivative formulas, no autograd"""
    n = d.n_pix * d.cout
    p, t = mem.f32(d.y_pred, n).view(d.n_pix, d.cout), mem.f32(d.y_true, n).view(d.n_pix, d.cout)
    eps, kind, act = 1e-7, d.kind, d.act
    own = (kind in (0, 7) and act == L.ACT_SIGMOID) or (kind == 1 and act == L.ACT_SOFTMAX)
    chan_sum = kind in (1, 9, 13, 14)
    scale = 1.0 / d.n_pix if chan_sum else 1.0 / n
    inside = (p > eps) & (p < 1 - eps)
    pc = p.clamp(eps, 1 - eps)
    if kind == 0:
        if own:
            l, g = -(t * pc.log() + (1 - t) * (1 - pc).log()), torch.zeros_like(p)
        else:
            l = -(t * (pc + eps).log() + (1 - t) * (1 - pc + eps).log())
            g = torch.where(inside, -(t / (pc + eps) - (1 - t) / (1 - pc + eps)), torch.zeros_like(p))
    elif kind == 1:
        if own:
            l, g = -t * p.clamp_min(eps).log(), torch.zeros_like(p)
        else:
            S = p.sum(-1, keepdim=True)
            q = p / S
            qin = (q > eps) & (q < 1 - eps)
            l = -t * q.clamp(eps, 1 - eps).log()
            g = -torch.where(qin, t / p, torch.zeros_like(p)) + (t * qin).sum(-1, keepdim=True) / S
    elif kind == 2:
        l, g = (p - t) ** 2, 2 * (p - t)
    elif kind == 3:
        l, g = (p - t).abs(), torch.sign(p - t)
    elif kind == 4:
        a, b = (p.clamp_min(eps) + 1).log(), (t.clamp_min(eps) + 1).log()
        l, g = (a - b) ** 2, torch.where(p > eps, 2 * (a - b) / (p + 1), torch.zeros_like(p))
    elif kind == 5:
        dd = p - t
        l = torch.where(dd.abs() <= 1, 0.5 * dd * dd, dd.abs() - 0.5)
        g = torch.where(dd.abs() <= 1, dd, torch.sign(dd))
    elif kind == 6:
        dd = p - t
        l, g = dd + torch.nn.functional.softplus(-2 * dd) - math.log(2.0), torch.tanh(dd)
    elif kind == 7:
        om = 1 - (t * p + (1 - t) * (1 - p))
        if own:
            bce = -(t * pc.log() + (1 - t) * (1 - pc).log())
            l = om * om * bce
            g = -2 * om * (2 * t - 1) * bce + om * om * (p - t) / (p * (1 - p)).clamp_min(1e-30)
        else:
            bce = -(t * (pc + eps).log() + (1 - t) * (1 - pc + eps).log())
            dbce = torch.where(inside, -(t / (pc + eps) - (1 - t) / (1 - pc + eps)), torch.zeros_like(p))
            l, g = om * om * bce, -2 * om * (2 * t - 1) * bce + om * om * dbce
    elif kind == 8:
        l, g = p - t * (p + eps).log(), 1 - t / (p + eps)
    elif kind == 9:
        tc, pk = t.clamp(eps, 1.0), p.clamp(eps, 1.0)
        l, g = tc * (tc / pk).log(), torch.where((p > eps) & (p < 1), -tc / pk, torch.zeros_like(p))
    elif kind in (10, 11):
        y = torch.where((t == 0) | (t == 1), 2 * t - 1, t)
        m = (1 - y * p).clamp_min(0)
        l, g = (m, torch.where(m > 0, -y, torch.zeros_like(p))) if kind == 10 else (m * m, -2 * m * y)
    elif kind == 12:
        den = t.abs().clamp_min(eps)
        dd = (t - p) / den
        l, g = 100 * dd.abs(), -100 * torch.sign(dd) / den
    elif kind == 13:
        pos, (neg, j) = (t * p).sum(-1), ((1 - t) * p).max(-1)
        margin = (neg - pos + 1).unsqueeze(-1)
        l = torch.zeros_like(p)
        l[:, 0] = margin[:, 0].clamp_min(0)
        hot = torch.nn.functional.one_hot(j, d.cout).double()
        g = torch.where(margin > 0, hot * (1 - t) - t, torch.zeros_like(p))
    else:
        a = torch.rsqrt((t * t).sum(-1, keepdim=True).clamp_min(1e-12))
        b = torch.rsqrt((p * p).sum(-1, keepdim=True).clamp_min(1e-12))
        c = (t * p).sum(-1, keepdim=True)
        l = torch.zeros_like(p)
        l[:, 0] = (-c * a * b)[:, 0]
        g = -a * (t * b - torch.where(b > 9.9e5, torch.zeros_like(c), c * p * b ** 3))
    if kind in (0, 1) and own:
        dz = p - t
    elif act == L.ACT_SIGMOID:
        dz = g * p * (1 - p)
    elif act == L.ACT_SOFTMAX:
        dz = p * (g - (g * p).sum(-1, keepdim=True))
    else:
        dz = g
    loss = l.sum() * scale
    if d.dlogits:
        mem.f32(d.dlogits, n)[:] = (d.weight * dz * scale).reshape(-1)
    if d.loss:
        mem.f32(d.loss, 1)[:] += d.weight * loss
    if d.metrics:
        mt = mem.f32(d.metrics, 5)
        e = p - t
        mt[0] += loss
        mt[1] += (e * e).sum()
        mt[2] += e.abs().sum()
        mt[3] += ((p > 0.5) == (t > 0.5)).double().sum()
        mt[4] += (p.argmax(-1) == t.argmax(-1)).double().sum()


def emu_eltwise(mem, d):
    a = mem.gather_view(d.a)
    if d.op == 0:
        o = _act(a + mem.gather_view(d.b), d.act)
    elif d.op == 3:
        o = _act(a + mem.gather_view(d.b) + mem.gather_view(d.c), d.act)
    elif d.op == 2:
        bb = mem.gather_view(d.b)
        o = a * torch.where(bb > 0, torch.ones_like(bb), torch.full_like(bb, 0.3))
    elif d.op == 4:            # a^p, p = d.act
        o = a ** int(d.act)
    elif d.op == 5:            # p * a^(p-1) * b
        o = int(d.act) * a ** (int(d.act) - 1) * mem.gather_view(d.b)
    else:
        assert d.op == 1, d.op
        o = a
    mem.write_view(d.out, o)


def emu_cast(mem, d):
    src = mem.f32(d.src, d.N * d.H * d.W * d.C).view(d.N, d.H, d.W, d.C)
    o = torch.zeros(d.N, d.H, d.W, d.out.C, dtype=torch.float64)
    if d.kh * d.kw > 1:          # K-packed im2col of the input
        ph, pw = (d.kh - 1) // 2, (d.kw - 1) // 2
        pad = torch.nn.functional.pad(src, (0, 0, pw, d.kw - 1 - pw, ph, d.kh - 1 - ph))
        for i in range(d.kh):
            for j in range(d.kw):
                t = i * d.kw + j
                o[..., t * d.C:(t + 1) * d.C] = pad[:, i:i + d.H, j:j + d.W, :]
    else:
        o[..., :d.C] = src
    mem.write_view(d.out, o)


def emu_colsum(mem, d):
    g = mem.gather_view(d.g)
    mem.f32(d.out, d.g.C)[:] += g.reshape(-1, d.g.C).sum(0)


def emu_memset(mem, d):
    t, off = mem.resolve(d.ptr)
    es = next(es for (b, nb, tt, es) in mem.bufs if tt is t)
    t[off:off + d.bytes // es] = 0


def _resize_matrix(in_size, f, mode):
    """[out, in] interpolation matrix of one axis (nearest repeat, or half-pixel bilinear with edge clamp)"""
    out = in_size * f
    M = torch.zeros(out, in_size, dtype=torch.float64)
    for o in range(out):
        if mode == 0:
            M[o, o // f] = 1.0
        else:
            src = (o + 0.5) / f - 0.5
            fl = math.floor(src)
            lam = src - fl
            i0, i1 = min(max(fl, 0), in_size - 1), min(max(fl + 1, 0), in_size - 1)
            M[o, i0] += 1.0 - lam
            M[o, i1] += lam
    return M


def emu_resize_fwd(mem, d):
    x = mem.gather_view(d.x)
    Mh, Mw = _resize_matrix(d.x.H, d.fh, d.mode), _resize_matrix(d.x.W, d.fw, d.mode)
    y = torch.einsum("ah,nhwc->nawc", Mh, x)
    y = torch.einsum("bw,nawc->nabc", Mw, y)
    y = _act(y, d.act)
    if d.n_vseg:
        real = torch.zeros(y.shape[-1], dtype=torch.bool)
        for i in range(d.n_vseg):
            real[d.vseg_off[i]:d.vseg_off[i] + d.vseg_cnt[i]] = True
        y[..., ~real] = 0
    elif d.c_valid:
        y[..., d.c_valid:] = 0
    mem.write_view(d.y, y)


def emu_resize_bwd(mem, d):
    g = mem.gather_view(d.y)
    if d.act != L.ACT_NONE:
        g = g * _dact_from_y(mem.gather_view(d.yfwd), d.act)
    Mh, Mw = _resize_matrix(d.x.H, d.fh, d.mode), _resize_matrix(d.x.W, d.fw, d.mode)
    dx = torch.einsum("ah,nabc->nhbc", Mh, g)
    dx = torch.einsum("bw,nhbc->nhwc", Mw, dx)
    mem.write_view(d.x, dx)


def emu_mulbc_fwd(mem, d):
    mem.write_view(d.out, mem.gather_view(d.a) * mem.gather_view(d.b)[..., :1])


def emu_mulbc_bwd(mem, d):
    a, b, g = mem.gather_view(d.a), mem.gather_view(d.b), mem.gather_view(d.dout)
    mem.write_view(d.da, g * b[..., :1])
    db = torch.zeros(d.db.N, d.db.H, d.db.W, d.db.C, dtype=torch.float64)
    db[..., 0] = (g * a).sum(-1)
    mem.write_view(d.db, db)


def emu_colstats(mem, d):
    x = mem.gather_view(d.x).reshape(-1, d.x.C)
    part = mem.f32(d.partials, d.n_blocks * 2 * d.x.C).view(d.n_blocks, 2, d.x.C)
    part.zero_()
    part[0, 0], part[0, 1] = x.sum(0), (x * x).sum(0)


def emu_rowsum(mem, d):
    part = mem.f32(d.partials, d.n_rows * d.pitch).view(d.n_rows, d.pitch)
    out = mem.f32(d.out, d.C)
    tot = part[:, :d.C].sum(0)
    out[:] = out + tot if d.accumulate else tot


def _hs(t):
    return torch.clamp(0.2 * t + 0.5, 0.0, 1.0)


def _hs_grad(t):
    u = 0.2 * t + 0.5
    return torch.where((u >= 0) & (u <= 1), torch.full_like(u, 0.2), torch.zeros_like(u))


def emu_lstm_fwd(mem, d):
    z = mem.gather_view(d.z)
    F = d.F
    zi, zg, zo = z[..., :F], z[..., F:2 * F], z[..., 2 * F:3 * F]
    mem.write_view(d.h, _hs(zo) * torch.tanh(_hs(zi) * torch.tanh(zg)))


def emu_lstm_bwd(mem, d):
    z = mem.gather_view(d.z)
    F = d.F
    zi, zg, zo = z[..., :F], z[..., F:2 * F], z[..., 2 * F:3 * F]
    dh = mem.gather_view(d.dh)
    gi, gg, go = _hs(zi), torch.tanh(zg), _hs(zo)
    c = gi * gg
    tc = torch.tanh(c)
    dc = dh * go * (1 - tc * tc)
    dz = torch.zeros_like(z)
    dz[..., :F] = dc * gg * _hs_grad(zi)
    dz[..., F:2 * F] = dc * gi * (1 - gg * gg)
    dz[..., 2 * F:3 * F] = dh * tc * _hs_grad(zo)
    mem.write_view(d.dz, dz)


def emu_pool_bwd(mem, d):
    y = mem.gather_view(d.y)
    N, H, W, C = y.shape
    ph, pw = d.ph, d.pw
    yw = y.view(N, H // ph, ph, W // pw, pw, C).permute(0, 1, 3, 5, 2, 4).reshape(N, H // ph, W // pw, C, ph * pw)
    first = yw.argmax(dim=-1)
    g = mem.gather_view(d.dp)
    onehot = torch.nn.functional.one_hot(first, ph * pw).double() * g.unsqueeze(-1)
    dx = onehot.view(N, H // ph, W // pw, C, ph, pw).permute(0, 1, 4, 2, 5, 3).reshape(N, H, W, C)
    mem.write_view(d.dx, dx)


def _gate_forward(mem, d, za, zb, skip, prm):
    """float64 restatement of csrc/gate.cu's forward on torch tensors (channels-last); prm: dict of parameter tensors.  Returns
    (out, z, batch statistics)"""
    import torch.nn.functional as F
    N, h, w, C = za.shape

    def bn(x, gamma, beta, mm, mv):
        if d.training:
            mean = x.reshape(-1, x.shape[-1]).mean(0)
            var = ((x.reshape(-1, x.shape[-1]) - mean) ** 2).mean(0)
        else:
            mean, var = mm, mv
        return (x - mean) * torch.rsqrt(var + d.eps) * gamma + beta, mean, var
    a, mean_a, var_a = bn(za, prm["gamma_a"], prm["beta_a"], prm["mm_a"], prm["mv_a"])
    b, mean_b, var_b = bn(zb, prm["gamma_b"], prm["beta_b"], prm["mm_b"], prm["mv_b"])
    z = (torch.relu(a + b) * prm["w3"]).sum(-1, keepdim=True) + prm["b3"]
    zn, mean3, var3 = bn(z, prm["gamma3"], prm["beta3"], prm["mm3"], prm["mv3"])
    m = torch.sigmoid(zn).permute(0, 3, 1, 2)                                     # N,1,h,w
    r1 = F.interpolate(m, scale_factor=2, mode="bilinear", align_corners=False)
    pre = F.conv_transpose2d(m, prm["wt"].view(1, 1, 4, 4), prm["bt"], stride=2, padding=1)
    r = (r1 + torch.where(pre > 0, pre, 0.3 * pre)).permute(0, 2, 3, 1)
    return skip * r, z, dict(a=(mean_a, var_a), b=(mean_b, var_b), c=(mean3, var3))


def _gate_params(mem, d, C, grad=False):
    prm = dict(gamma_a=mem.f32(d.gamma_a, C), beta_a=mem.f32(d.beta_a, C), mm_a=mem.f32(d.mm_a, C), mv_a=mem.f32(d.mv_a, C),
               gamma_b=mem.f32(d.gamma_b, C), beta_b=mem.f32(d.beta_b, C), mm_b=mem.f32(d.mm_b, C), mv_b=mem.f32(d.mv_b, C),
               w3=mem.f32(d.w3, C), b3=mem.f32(d.b3, 1), gamma3=mem.f32(d.gamma3, 1), beta3=mem.f32(d.beta3, 1), mm3=mem.f32(d.mm3, 1),
               mv3=mem.f32(d.mv3, 1), wt=mem.f32(d.wt, 16 * d.wt_stride)[::d.wt_stride], bt=mem.f32(d.bt, 1))
    if grad:
        prm = {k: v.clone().requires_grad_(not k.startswith("m")) for k, v in prm.items()}
    return prm


def emu_gate_fwd(mem, d):
    za, zb, skip = mem.gather_view(d.za), mem.gather_view(d.zb), mem.gather_view(d.skip)
    N, h, w, C = za.shape
    prm = _gate_params(mem, d, C)
    out, z, st = _gate_forward(mem, d, za, zb, skip, prm)
    mem.f32(d.z, N * h * w)[:] = z.reshape(-1)
    mem.write_view(d.out, out)
    if d.training:
        # the kernel derives the branch statistics from the accumulated column sums: they must agree with the direct ones
        for tag, x in (("a", za), ("b", zb)):
            sums = mem.f32(getattr(d, "sums_" + tag), 2 * C)
            flat = x.reshape(-1, C)
            assert torch.allclose(sums[:C], flat.sum(0), rtol=1e-9, atol=1e-9) and torch.allclose(sums[C:], (flat * flat).sum(0), rtol=1e-9, atol=1e-9), \
                "gate: projection statistics accumulators do not hold the column sums"
        s3 = mem.f32(d.sums3, 2)
        s3[0] += z.sum()
        s3[1] += (z * z).sum()
        n = d.count
        for tag, key in (("a", "a"), ("b", "b"), ("3", "c")):
            mean, var = st[key]
            uv = var * n / (n - 1) if (d.bessel and n > 1) else var
            mm, mv = prm["mm_" + tag if tag != "3" else "mm3"], prm["mv_" + tag if tag != "3" else "mv3"]
            mm[:] = mm * d.momentum + mean * (1 - d.momentum)
            mv[:] = mv * d.momentum + uv * (1 - d.momentum)
    for tag in ("a", "b"):
        mean, var = st[tag] if d.training else (prm["mm_" + tag], prm["mv_" + tag])
        rstd = torch.rsqrt(var + d.eps)
        vec = mem.f32(getattr(d, "vec_" + tag), 4 * C)
        vec[:C], vec[C:2 * C] = prm["gamma_" + tag] * rstd, prm["beta_" + tag] - mean * prm["gamma_" + tag] * rstd
        vec[2 * C:3 * C], vec[3 * C:] = mean, rstd


def emu_gate_bwd(mem, d):
    if d.da_low.ptr:          # second pass: dskip = dout * r + the stride-2 projection's input gradient at the even pixels
        za, zb, skip = mem.gather_view(d.za), mem.gather_view(d.zb), mem.gather_view(d.skip)
        out, _z, _st = _gate_forward(mem, d, za, zb, torch.ones_like(skip), _gate_params(mem, d, za.shape[-1]))     # = r broadcast over channels
        dskip = mem.gather_view(d.dout) * out
        dskip[:, ::2, ::2] += mem.gather_view(d.da_low)
        mem.write_view(d.dskip, dskip)
        return
    za = mem.gather_view(d.za).clone().requires_grad_(True)
    zb = mem.gather_view(d.zb).clone().requires_grad_(True)
    skip = mem.gather_view(d.skip).clone().requires_grad_(True)      # (a leaf here: gradient through the multiply only)
    N, h, w, C = za.shape
    prm = _gate_params(mem, d, C, grad=True)
    out, _z, _st = _gate_forward(mem, d, za, zb, skip, prm)
    out.backward(mem.gather_view(d.dout))
    if d.dskip.ptr:
        mem.write_view(d.dskip, skip.grad)
    mem.write_view(d.dza, za.grad)
    mem.write_view(d.dzb, zb.grad)
    for name in ("gamma_a", "beta_a", "gamma_b", "beta_b"):
        mem.f32(getattr(d, "d" + name), C)[:] = prm[name].grad
    mem.f32(d.dgamma3, 1)[:] = prm["gamma3"].grad
    mem.f32(d.dbeta3, 1)[:] = prm["beta3"].grad
    mem.f32(d.dw3, C)[:] += prm["w3"].grad
    if d.db3:
        mem.f32(d.db3, 1)[:] += prm["b3"].grad
    mem.f32(d.dwt, 16 * d.wt_stride)[::d.wt_stride] += prm["wt"].grad
    mem.f32(d.dbt, 1)[:] += prm["bt"].grad


def emu_fold_bn(mem, d):
    w = mem.f32(d.w, d.cout_p * d.row).view(d.cout_p, d.row)
    s_ = mem.f32(d.gamma, d.cout_p) * torch.rsqrt(mem.f32(d.moving_var, d.cout_p) + d.eps)
    wf, off = mem.resolve(d.w_folded)
    wf[off:off + d.cout_p * d.row] = (w * s_[:, None]).reshape(-1)
    b = mem.f32(d.bias, d.cout_p) if d.bias else torch.zeros(d.cout_p, dtype=torch.float64)
    mem.f32(d.bias_folded, d.cout_p)[:] = b * s_ + mem.f32(d.beta, d.cout_p) - mem.f32(d.moving_mean, d.cout_p) * s_


EMU = {L.OP_FOLD_BN: emu_fold_bn, L.OP_GATE_FWD: emu_gate_fwd, L.OP_GATE_BWD: emu_gate_bwd, L.OP_CONV: emu_conv, L.OP_WGRAD: emu_wgrad, L.OP_BN_FINALIZE: emu_bn_finalize, L.OP_BN_ACT: emu_bn_act,
       L.OP_BN_BWD: emu_bn_bwd, L.OP_ADAM: emu_adam, L.OP_HEAD_FWD: emu_head_fwd, L.OP_HEAD_BWD: emu_head_bwd,
       L.OP_LOSS: emu_loss, L.OP_ELTWISE: emu_eltwise, L.OP_CAST: emu_cast, L.OP_COLSUM: emu_colsum,
       L.OP_MEMSET: emu_memset, L.OP_RESIZE_FWD: emu_resize_fwd, L.OP_RESIZE_BWD: emu_resize_bwd,
       L.OP_MULBC_FWD: emu_mulbc_fwd, L.OP_MULBC_BWD: emu_mulbc_bwd, L.OP_COLSTATS: emu_colstats,
       L.OP_LSTM_FWD: emu_lstm_fwd, L.OP_LSTM_BWD: emu_lstm_bwd, L.OP_POOL_BWD: emu_pool_bwd,
       L.OP_ROWSUM: emu_rowsum, L.OP_OUTACT_FWD: emu_outact_fwd, L.OP_OUTACT_BWD: emu_outact_bwd}


def run_phase(mem, planner, phase, first_op=0, n_ops=None):
    """replay ops [first_op, first_op + n_ops) of a phase (default: the whole phase), like b2seg_plan_run_range"""
    ops = planner.ops[phase]
    end = len(ops) if n_ops is None else first_op + n_ops
    for (op, desc, _note) in ops[first_op:end]:
        EMU[op](mem, desc)
