"""CPU emulation of the b2seg conv / wgrad descriptor semantics (include/b2seg.h) on torch tensors.

Test infrastructure only: it lets the tap tables produced by b2seg.lowering be checked against
torch.nn.functional on the CPU, without a GPU.  Views are resolved against a dict {base_ptr: flat fp32 tensor}
that fakes device memory (element size 2 like bf16 addresses).
"""
import torch


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
