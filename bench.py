#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: 2D UNet (depth 5, width 64, 256x256x3, batch 32/GPU, bf16) training
images/s on N B200s, with the tensor-core roofline of the dominant kernel and the CPU oracle timed beside it.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference        # the reference's CPU path (oracle port: TensorFlow is not installable here)

One "step" = forward + backward + Adam over one batch of 32 synthetic images per GPU.
`value`   : K steps with the batch already resident in HBM (CUDA events, barrier + synchronize on both sides, max over ranks).
`e2e`     : the same K steps through Model.fit over K batches in pinned HOST memory (H2D of each step's x and y, D2H of each step's loss;
            fit overlaps the copy of batch i+1 with the compute of step i).
`roofline`: algorithmic FLOPs of the implicit-GEMM conv kernel launches / their CUDA-event durations vs MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

TRAIN_GFLOP_PER_IMAGE = 275.088  # BASELINE.md §2, config 2 (fwd + dgrad + wgrad, first-layer dgrad omitted)
CONV_FAM = "conv_halo_kernel+conv_gemm_kernel"      # fprop / dgrad / transposed-conv launches (b2seg_conv)
WGRAD_FAM = "wgrad_halo_kernel+wgrad_kernel"       # weight-gradient launches (b2seg_wgrad)
WORKLOAD = "2D UNet depth5 width64 256x256x3 from_scratch dense_loop=1 transconv, BCE + Adam(2e-4), batch 32/GPU"


def synth_batch(batch, size, seed):
    """x ~ U[0,1); y = blobs (box-blurred uniform noise > threshold, ~30 % foreground)."""
    rng = np.random.default_rng(seed)
    x = rng.random((batch, size, size, 3), dtype=np.float32)
    u = rng.random((batch, size, size), dtype=np.float32)
    k = 9
    c = np.cumsum(np.cumsum(np.pad(u, ((0, 0), (k, k), (k, k)), mode="wrap"), 1), 2)
    blur = (c[:, k:, k:] - c[:, :-k, k:] - c[:, k:, :-k] + c[:, :-k, :-k])[:, :size, :size] / (k * k)
    y = (blur > np.quantile(blur, 0.7)).astype(np.float32)[..., None]
    return x, y


class ClockSampler:
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU oracle leg
def oracle_train_step_time(batch, size, steps, warmup, depth=5, width=64):
    """the reference's CPU path for this graph, restated in PyTorch fp32 (oracle/): forward + backward + Keras Adam"""
    import torch
    from oracle.keras_ref import KerasRef, keras_adam_step, keras_loss
    from oracle.ref_models import Ref2D
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = Ref2D("UNet", size, size, width, depth, num_channels=3, output_nums=1, dense_loop=1, is_transconv=True)
    x, y = synth_batch(batch, size, 2)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    params, state = {}, {}
    times = []
    for t in range(1, warmup + steps + 1):
        t0 = time.perf_counter()
        k = KerasRef(2, params=params, dtype=torch.float32, training=True)
        out = ref(k, xt)[0]
        loss = keras_loss("bce", out, yt, logits=k.logits["out"])
        loss.backward()
        with torch.no_grad():
            for key in k.trainable:
                w = params[key]
                if w.grad is None:
                    continue
                if key not in state:
                    state[key] = (torch.zeros_like(w), torch.zeros_like(w))
                keras_adam_step(w, w.grad, state[key][0], state[key][1], t, lr=2e-4)
                w.grad = None
            for key, v in k.new_moving.items():
                params[key] = v
        if t > warmup:
            times.append(time.perf_counter() - t0)
    return float(np.sum(times)), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2
    total, cores = oracle_train_step_time(batch, args.size, args.steps, args.warmup)
    ips = batch * args.steps / total
    line = {"impl": "reference", "metric": "2D UNet 256^2 train images/s", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference arm = oracle port of the reference's TF/Keras CPU path (TensorFlow is not installable in this image)"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} train steps of batch {batch} at {args.size}x{args.size} (same graph, fp32, PyTorch CPU)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2seg", choices=["b2seg", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ops-json", default="", help="dump per-op device times (profiling aid)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2seg" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from b2seg import _lib as L
    from b2seg.model import Adam
    from b2seg.models2d import unet_model_builder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # the backward kernels that overlap the gradient all-reduce leave B2SEG_BWD_SM_RESERVE SMs to it (default 16; measured
        # at 8 GPUs: no reserve 12.82, 24 -> 12.34, 16 -> 12.24 ms/step); keep NCCL inside that budget (the variable must be
        # set before the communicator exists)
        if int(os.environ.get("B2SEG_BWD_SM_RESERVE", "16")) > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("B2SEG_BWD_SM_RESERVE", "16"))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, S = args.batch, args.size
    model = unet_model_builder("UNet", S, S, 64, 5, num_channels=3, output_nums=1, ds=0, ae=0, ag=0, lstm=0, dense_loop=1,
                               is_transconv=True, final_activation="sigmoid", train_mode="from_scratch").ResNet50()
    model.compile(loss="binary_crossentropy", optimizer=Adam(2e-4))
    x, y = synth_batch(B, S, 2 + rank)
    xp, yp = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    eng = model._engine(B, True)
    if world > 1:
        model.distribute()
        model.broadcast_weights(0)
    eng.x_dev.copy_(xp)
    eng.outputs[0]["target"].copy_(yp)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- device-resident arm (value)
    for _ in range(args.warmup):
        model._step(eng, return_loss=False)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(lambda: model._step(eng, return_loss=False), args.steps)
    clocks = sampler.stop()
    value = world * B * args.steps / (ms / 1e3)

    # ---- end-to-end arm: the public API the reference's Train.py drives (Model.fit over host arrays, 2DCNN/Train.py:394-415).
    # Every step copies ITS batch from pinned host memory (a different 33.5 MB slice per step) and reads its loss back;
    # fit() overlaps the copy of batch i+1 with the compute of step i (two staging slots + a copy stream).
    ns = args.steps
    xe = torch.from_numpy(np.concatenate([x] * ns, 0)).pin_memory()
    ye = torch.from_numpy(np.concatenate([y] * ns, 0)).pin_memory()
    xh, yh = xe.numpy(), ye.numpy()
    model.fit(xh[:2 * B], yh[:2 * B], batch_size=B, epochs=1, shuffle=False, verbose=0)   # warm-up (staging buffers, streams)
    last = {}

    def e2e_run():
        h = model.fit(xh, yh, batch_size=B, epochs=1, shuffle=False, verbose=0)
        last["loss"] = h.history["loss"][-1]

    ms_e2e = timed(e2e_run, 1)
    e2e_value = world * B * ns / (ms_e2e / 1e3)

    # ---- per-op device times (one extra, untimed-for-throughput replay with an event after every op)
    roof = None
    if rank == 0:
        peak_tf, peak_gbs, peak_src = measured_peaks()
        agg = {}
        op_rows = []
        for phase in (0, 1, 2):
            n_ops = eng.lib.b2seg_plan_num_ops(eng.plan, phase)
            buf = (C.c_float * n_ops)()
            reps = 3
            acc = np.zeros(n_ops)
            for _ in range(reps):
                L.check(eng.lib.b2seg_plan_run_timed(eng.plan, phase, C.c_void_p(eng._stream()), buf, n_ops), "run_timed")
                acc += np.array(list(buf))
            acc /= reps
            for i in range(n_ops):
                info = eng.planner.op_info[(phase, i)]
                fam = {L.OP_CONV: CONV_FAM, L.OP_WGRAD: WGRAD_FAM}.get(info["op"], "streaming")
                op_rows.append({"phase": phase, "i": i, "op": info["op"], "note": info["note"], "ms": float(acc[i]),
                                "tflops": info["flops"] / (acc[i] / 1e3) / 1e12 if info["flops"] and acc[i] > 0 else None})
                a = agg.setdefault(fam, dict(ms=0.0, flops=0.0, launches=0))
                a["ms"] += float(acc[i]); a["flops"] += info["flops"]; a["launches"] += 1
        if args.ops_json:
            os.makedirs(os.path.dirname(os.path.abspath(args.ops_json)), exist_ok=True)
            with open(args.ops_json, "w") as f:
                json.dump(op_rows, f, indent=0)
        step_ms_ops = sum(a["ms"] for a in agg.values())
        dom = max((CONV_FAM, WGRAD_FAM), key=lambda k: agg.get(k, {"ms": 0})["ms"])
        a = agg[dom]
        achieved = a["flops"] / (a["ms"] / 1e3) / 1e12
        # DRAM bytes per launch of the dominant family from the committed ncu capture of this same command (tools/launch_summary.py)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_launches_final.json")
        if os.path.exists(tpath) and S == 256 and B == 32:
            with open(tpath) as f:
                tj = json.load(f)
            key = "conv" if dom == CONV_FAM else "wgrad"
            traffic = tj[key]["dram_bytes_per_launch"]
            traffic_src = f"profiles/r1_launches_final.json: ncu dram__bytes_read.sum + dram__bytes_write.sum over the {tj[key]['launches']} {key} launches of one step / launches"
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launches_per_step": a["launches"], "kernel_ms_per_step": a["ms"],
                "kernel_share_of_step": a["ms"] / step_ms_ops,
                "families": {k: {"ms_per_step": v["ms"], "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["flops"] else None,
                                 "launches": v["launches"]} for k, v in agg.items()},
                "whole_step_tflops": value / world * TRAIN_GFLOP_PER_IMAGE / 1e3 * (S * S) / (256 * 256),
                "whole_step_frac_of_peak": value / world * TRAIN_GFLOP_PER_IMAGE / 1e3 * (S * S) / (256 * 256) / peak_tf}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        total, cores = oracle_train_step_time(4, S, 4, 1)
        cpu = {"value": 4 * 4 / total, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"4 train steps of batch 4 at {S}x{S} after 1 warm-up (oracle: PyTorch-CPU fp32 restatement of the reference graph)"}

    if rank == 0:
        launches = sum(eng.launches)
        line = {"metric": "2D UNet 256^2 train images/s", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD if S == 256 and B == 32 else f"{WORKLOAD} [overridden: size {S}, batch {B}]",
                           "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2_policy": "per-step working set (activations+weights+Adam state, >5 GB) exceeds the 126 MB L2; no flush needed"},
                "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(x.nbytes + y.nbytes), "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps, "last_loss": last.get("loss"),
                        "api": "Model.fit(x, y, batch_size, epochs=1, shuffle=False) over steps x batch samples in pinned host memory"},
                "gpu_launches": launches * args.steps, "launches_per_step": launches,
                "device_memory_gb": eng.memory_bytes() / 2 ** 30}
        if getattr(eng, "exchange_calibration", None):
            line["exchange"] = eng.exchange_calibration
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
