"""Micro-benchmark tool (not a test): times the HBM-streaming ops of a cfg2 training step through the C ABI with CUDA
events and prints achieved GB/s on the algorithmic bytes (elements read + written x element size).
Usage (GPU box): python tests/tools_stream_bench.py [filter] [reps]
  ncu --set full --clock-control none --import-source on -k regex:bn_ -c 6 -o gpurun_out/prof python tests/tools_stream_bench.py bnbwd_pool 1"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200"))

from b2seg import _lib as L  # noqa: E402
from b2seg import lowering as lw  # noqa: E402


def tv(t, c_off=0, Cn=None):
    N, H, W, Ct = t.shape
    esz = t.element_size()
    return lw.TView(t.data_ptr() + c_off * esz, N, H, W, Ct - c_off if Cn is None else Cn, t.stride(0), t.stride(1), t.stride(2), esz)


def bfr(*shape):
    return torch.randn(*shape, device="cuda").to(torch.bfloat16)


def bn_vectors(Cc):
    return [torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1, torch.randn(Cc, device="cuda") * 0.1, torch.rand(Cc, device="cuda") + 0.5]


def case_bnact(N, H, W, Cc, pool, n_out=1):
    z = bfr(N, H, W, Cc)
    sc, sf, _, _ = bn_vectors(Cc)
    cat = torch.empty(N, H, W, 2 * Cc, device="cuda", dtype=torch.bfloat16)
    a = torch.empty(N, H, W, Cc, device="cuda", dtype=torch.bfloat16)
    d = L.BnActDesc()
    d.x, d.scale, d.shift, d.act, d.n_out = tv(z).to_c(), sc.data_ptr(), sf.data_ptr(), L.ACT_RELU, n_out
    d.out[0] = tv(cat, Cc, Cc).to_c()
    if n_out > 1:
        d.out[1] = tv(a).to_c()
    nbytes = z.numel() * 2 * (1 + n_out)
    keep = [z, sc, sf, cat, a]
    if pool:
        p = torch.empty(N, H // 2, W // 2, Cc, device="cuda", dtype=torch.bfloat16)
        d.pool_h, d.pool_w, d.pooled = 2, 2, tv(p).to_c()
        nbytes += p.numel() * 2
        keep.append(p)
    return L.OP_BN_ACT, d, nbytes, keep


def case_bnbwd(N, H, W, Cc, pooled_src, direct_in_concat=True):
    z = bfr(N, H, W, Cc)
    sc, sf, mu, rs = bn_vectors(Cc)
    gcat = bfr(N, H, W, 2 * Cc) if direct_in_concat else bfr(N, H, W, Cc)
    dz = torch.empty_like(z)
    nb = max(1, min(1184, (N * H * W) // 64))
    partials = torch.zeros(nb, 2, Cc, device="cuda")
    dgamma, dbeta = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    d = L.BnBwdDesc()
    d.x, d.scale, d.shift, d.mean, d.rstd = tv(z).to_c(), sc.data_ptr(), sf.data_ptr(), mu.data_ptr(), rs.data_ptr()
    d.act = L.ACT_RELU
    d.src[0] = L.GradSrc((tv(gcat, Cc, Cc) if direct_in_concat else tv(gcat)).to_c(), 0, 1, 1)
    d.n_src = 1
    keep = [z, sc, sf, mu, rs, gcat, dz, partials, dgamma, dbeta]
    reads = 2.0
    if pooled_src:
        g2 = bfr(N, H // 2, W // 2, Cc)
        d.src[1] = L.GradSrc(tv(g2).to_c(), 1, 2, 2)
        d.n_src = 2
        keep.append(g2)
        reads += 0.25
    d.count, d.partials, d.n_blocks = float(N * H * W), partials.data_ptr(), nb
    d.dgamma, d.dbeta, d.dx = dgamma.data_ptr(), dbeta.data_ptr(), tv(dz).to_c()
    nbytes = z.numel() * 2 * (2 * reads + 1)   # pass 0 reads, pass 1 reads + one write
    return L.OP_BN_BWD, d, nbytes, keep


def case_colsum(N, H, W, Cc):
    g = bfr(N, H, W, 2 * Cc)
    out = torch.zeros(Cc, device="cuda")
    nb = 592
    scratch = torch.zeros(nb, Cc, device="cuda")
    d = L.ColsumDesc(tv(g, 0, Cc).to_c(), out.data_ptr(), scratch.data_ptr(), nb)
    return L.OP_COLSUM, d, g.numel(), [g, out, scratch]   # reads Cc of 2*Cc channels


def case_head(N, H, W, Cc, bwd):
    x = bfr(N, H, W, Cc)
    w = torch.randn(Cc, 1, device="cuda") * 0.1
    b = torch.zeros(1, device="cuda")
    y = torch.empty(N, H, W, 1, device="cuda")
    dl = torch.randn(N, H, W, 1, device="cuda")
    dx = torch.empty_like(x)
    dw, db = torch.zeros(Cc, 1, device="cuda"), torch.zeros(1, device="cuda")
    d = L.HeadDesc()
    d.x, d.w, d.b, d.cout, d.act, d.stride = tv(x).to_c(), w.data_ptr(), b.data_ptr(), 1, L.ACT_SIGMOID, 1
    d.y, d.dlogits, d.dx, d.dw, d.db = y.data_ptr(), dl.data_ptr(), tv(dx).to_c(), dw.data_ptr(), db.data_ptr()
    nbytes = x.numel() * 2 * (2 if bwd else 1) + y.numel() * 4
    return (L.OP_HEAD_BWD if bwd else L.OP_HEAD_FWD), d, nbytes, [x, w, b, y, dl, dx, dw, db]


def case_adam(n):
    w, g, m, v = (torch.randn(n, device="cuda") * 0.01 for _ in range(4))
    v = v.abs()
    wb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    d = L.AdamDesc(w.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), wb.data_ptr(), n, 2e-4, 0.9, 0.999, 1e-7, 1.0, 1)
    return L.OP_ADAM, d, n * 30, [w, g, m, v, wb]


CASES = [
    ("bnact_256_64", lambda: case_bnact(32, 256, 256, 64, False)),
    ("bnact_pool_256_64", lambda: case_bnact(32, 256, 256, 64, True)),
    ("bnact_pool_128_128", lambda: case_bnact(32, 128, 128, 128, True)),
    ("bnact_pool_32_512", lambda: case_bnact(32, 32, 32, 512, True)),
    ("bnbwd_256_64", lambda: case_bnbwd(32, 256, 256, 64, False, False)),
    ("bnbwd_128_128", lambda: case_bnbwd(32, 128, 128, 128, False, False)),
    ("bnbwd_pool_256_64", lambda: case_bnbwd(32, 256, 256, 64, True)),
    ("bnbwd_pool_128_128", lambda: case_bnbwd(32, 128, 128, 128, True)),
    ("bnbwd_pool_64_256", lambda: case_bnbwd(32, 64, 64, 256, True)),
    ("bnbwd_pool_32_512", lambda: case_bnbwd(32, 32, 32, 512, True)),
    ("colsum_256_64", lambda: case_colsum(32, 256, 256, 64)),
    ("colsum_128_128", lambda: case_colsum(32, 128, 128, 128)),
    ("head_fwd_256_64", lambda: case_head(32, 256, 256, 64, False)),
    ("head_bwd_256_64", lambda: case_head(32, 256, 256, 64, True)),
    ("adam_170M", lambda: case_adam(170_529_856)),
]


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    lib = L.load()
    L.check(lib.b2seg_device_check(0), "device")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    for name, make in CASES:
        if flt and not name.startswith(flt):
            continue
        op, d, nbytes, keep = make()
        plan = C.c_void_p()
        L.check(lib.b2seg_plan_create(C.byref(plan)), "plan_create")
        L.check(lib.b2seg_plan_add(plan, 0, op, C.byref(d), C.sizeof(d)), "plan_add")
        L.check(lib.b2seg_plan_run(plan, 0, C.c_void_p(st)), "plan_run")
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()   # evict the working set from L2 between repetitions
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.b2seg_plan_run(plan, 0, C.c_void_p(st))
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        lib.b2seg_plan_destroy(plan)
        ms = tot / reps
        print(f"{name:<22}{ms:8.3f} ms  {nbytes / 1e6:9.1f} MB  {nbytes / ms / 1e6:8.0f} GB/s", flush=True)
        del keep


if __name__ == "__main__":
    main()
