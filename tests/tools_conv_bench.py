"""Micro-benchmark tool (not a test): times single conv / dgrad / wgrad launches of the cfg2 layer shapes through the
C ABI with CUDA events.  Usage (GPU box): python tests/tools_conv_bench.py [filter] ; profile one shape with
  ncu --set full --clock-control none --import-source on -k regex:conv_ -c 2 -o gpurun_out/prof python tests/tools_conv_bench.py conv2d_12"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200"))

from b2seg import _lib as L  # noqa: E402
from b2seg import lowering as lw  # noqa: E402

# name, kind, N, H, W, Cin, Cout  (cfg2 at batch 32)
SHAPES = [
    ("conv2d", "fprop", 32, 256, 256, 8, 64),
    ("probe_64_64", "fprop", 32, 256, 256, 64, 64),
    ("probe_16_64", "fprop", 32, 256, 256, 16, 64),
    ("conv2d_1", "fprop", 32, 128, 128, 64, 128),
    ("conv2d_2", "fprop", 32, 64, 64, 128, 256),
    ("conv2d_3", "fprop", 32, 32, 32, 256, 512),
    ("conv2d_6", "fprop", 32, 8, 8, 1024, 2048),
    ("conv2d_10", "fprop", 32, 64, 64, 512, 256),
    ("conv2d_11", "fprop", 32, 128, 128, 256, 128),
    ("conv2d_12", "fprop", 32, 256, 256, 128, 64),
    ("dgrad_conv2d_12", "dgrad", 32, 256, 256, 128, 64),
    ("dgrad_conv2d_11", "dgrad", 32, 128, 128, 256, 128),
    ("tconv_4", "tconv", 32, 128, 128, 128, 64),
    ("wgrad_conv2d_12", "wgrad", 32, 256, 256, 128, 64),
    ("wgrad_conv2d_11", "wgrad", 32, 128, 128, 256, 128),
    ("wgrad_conv2d_10", "wgrad", 32, 64, 64, 512, 256),
    ("wgrad_conv2d_9", "wgrad", 32, 32, 32, 1024, 512),
    ("wgrad_conv2d_8", "wgrad", 32, 16, 16, 2048, 1024),
    ("wgrad_conv2d_6", "wgrad", 32, 8, 8, 2048, 2048),
    ("wgrad_conv2d_4", "wgrad", 32, 16, 16, 512, 1024),
    ("wgrad_conv2d_3", "wgrad", 32, 32, 32, 256, 512),
    ("wgrad_conv2d_2", "wgrad", 32, 64, 64, 128, 256),
    ("wgrad_conv2d_1", "wgrad", 32, 128, 128, 64, 128),
    ("wgrad_conv2d", "wgrad", 32, 256, 256, 8, 64),
    ("wgrad_tconv_4", "wgrad_t", 32, 128, 128, 128, 64),
    ("wgrad_tconv_2", "wgrad_t", 32, 32, 32, 512, 256),
    ("wgrad_tconv", "wgrad_t", 32, 8, 8, 2048, 1024),
]


def tv(t, c_off=0, Cn=None):
    N, H, W, Ct = t.shape
    return lw.TView(t.data_ptr() + c_off * 2, N, H, W, Ct - c_off if Cn is None else Cn, t.stride(0), t.stride(1), t.stride(2))


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    L.check(L.load().b2seg_device_check(0), "device")
    dev = "cuda"
    st = torch.cuda.current_stream().cuda_stream
    for (name, kind, N, H, W, Cin, Cout) in SHAPES:
        if flt and flt != name and not (flt.endswith("*") and name.startswith(flt[:-1])):
            continue
        x = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
        taps = 16 if kind == "tconv" else 9
        w = (torch.randn(Cout, taps, Cin, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.zeros(Cout, device=dev)
        if kind == "fprop":
            out = torch.empty(N, H, W, Cout, device=dev, dtype=torch.bfloat16)
            d = lw.conv_fprop(tv(x), w.data_ptr(), Cout, 3, 3, Cin, tv(out), bias=bias.data_ptr())
            rows = L.load().b2seg_conv_num_stat_rows(__import__("ctypes").byref(d))
            stats = torch.zeros(rows, 2, Cout, device=dev)
            if not os.environ.get("CONVBENCH_NO_STATS"):
                d.stats = stats.data_ptr()
            fn, flops = "b2seg_conv", 2.0 * N * H * W * Cin * Cout * 9
        elif kind == "dgrad":
            dy = torch.randn(N, H, W, Cout, device=dev).to(torch.bfloat16)
            dx = torch.empty(N, H, W, Cin, device=dev, dtype=torch.bfloat16)
            d = lw.conv_dgrad(tv(dy), w.data_ptr(), Cout, 3, 3, Cin, tv(dx))
            fn, flops = "b2seg_conv", 2.0 * N * H * W * Cin * Cout * 9
        elif kind == "tconv":
            out = torch.empty(N, 2 * H, 2 * W, Cout, device=dev, dtype=torch.bfloat16)
            d = lw.tconv_fprop(tv(x), w.data_ptr(), Cout, 4, 4, Cin, tv(out), bias=bias.data_ptr(), act=L.ACT_LEAKY)
            fn, flops = "b2seg_conv", 2.0 * N * H * W * Cin * Cout * 16
        elif kind == "wgrad_t":
            dy = torch.randn(N, 2 * H, 2 * W, Cout, device=dev).to(torch.bfloat16)
            dw = torch.zeros(Cout, 16, Cin, device=dev)
            d = lw.tconv_wgrad(tv(dy), tv(x), dw.data_ptr(), Cout, 4, 4, Cin)
            fn, flops = "b2seg_wgrad", 2.0 * N * H * W * Cin * Cout * 16
        else:
            dy = torch.randn(N, H, W, Cout, device=dev).to(torch.bfloat16)
            dw = torch.zeros(Cout, 9, Cin, device=dev)
            d = lw.conv_wgrad(tv(dy), tv(x), dw.data_ptr(), Cout, 3, 3, Cin)
            fn, flops = "b2seg_wgrad", 2.0 * N * H * W * Cin * Cout * 9
        # prepare once (tensor maps, tables), replay: the way a training plan runs it
        import ctypes as C
        lib = L.load()
        plan = C.c_void_p()
        L.check(lib.b2seg_plan_create(C.byref(plan)), "plan_create")
        L.check(lib.b2seg_plan_add(plan, 0, L.OP_CONV if fn == "b2seg_conv" else L.OP_WGRAD, C.byref(d), C.sizeof(d)), "plan_add")
        L.check(lib.b2seg_plan_run(plan, 0, C.c_void_p(st)), "plan_run")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lib.b2seg_plan_run(plan, 0, C.c_void_p(st))
        e1.record()
        torch.cuda.synchronize()
        lib.b2seg_plan_destroy(plan)
        ms = e0.elapsed_time(e1) / reps
        print(f"{name:<18}{kind:<7} N{N} {H}x{W} {Cin}->{Cout}  {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)
        if os.environ.get("B2SEG_TRACE") and fn == "b2seg_conv":
            # per-tile clock64() stamps of CTA 0 (include/b2seg.h: b2seg_debug_read_trace), relative to the first stamp
            buf = (C.c_uint64 * (48 * 8))()
            lib.b2seg_debug_read_trace.argtypes = [C.POINTER(C.c_uint64), C.c_int]
            if lib.b2seg_debug_read_trace(buf, 48 * 8) > 0:
                t = [[int(buf[i * 8 + j]) for j in range(8)] for i in range(48)]
                base = min(v for row in t for v in row[:7] if v)
                print("  tile | mma: acc free, A landed, issued | producer TMA | epi: acc full, tmem read, done   (clocks since start)")
                for i in range(8, 24):
                    r = [v - base if v else -1 for v in t[i][:7]]
                    print(f"  {i:4d} | {r[0]:8d} {r[1]:8d} {r[2]:8d} | {r[3]:8d} | {r[4]:8d} {r[5]:8d} {r[6]:8d}")


if __name__ == "__main__":
    main()
