"""world_size-2 gloo test of the data-parallel path on the CPU: each rank runs forward + backward of the SAME
planner program on its own batch shard through the float64 descriptor emulator; the gradient exchange is interleaved with
the backward ops exactly as Model._step does it on the GPU (Planner.exchange_schedule: a slice of the flat gradient arena
is sum-all-reduced as soon as the ops that finish it have run, while later ops are still to come) and Adam applies
grad_scale = 1/world.  Both ranks must end with identical weights,
equal to a single-process run that averages the two per-replica gradients by hand."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(batch, adam_bucket_bytes=0):
    from b2seg.graph import init_params
    from b2seg.models2d import unet_model_builder
    from b2seg.planner import Planner
    from desc_emulator import PlanMem
    g = unet_model_builder("UNet", 16, 16, 8, 2, num_channels=1, train_mode="from_scratch").build_graph()
    mem = PlanMem()
    pl = Planner(g, batch, mem.alloc_bytes, training=True, losses=["bce"], adam=dict(lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-7),
                 adam_bucket_bytes=adam_bucket_bytes).build()
    params = init_params(g, seed=3)
    for e in pl.params:
        flat = torch.from_numpy(pl.to_internal(e.key, params[e.key])).double()
        if e.trainable:
            mem.f32(pl.w_ptr + 4 * e.offset, e.size)[:] = flat
            wb, off = mem.resolve(pl.wb_ptr + 2 * e.offset)
            wb[off:off + e.size] = flat
        else:
            mem.f32(pl.mov_ptr + 4 * e.offset, e.size)[:] = flat
    return g, mem, pl


def _fwd_bwd(mem, pl, x, y):
    from desc_emulator import run_phase
    mem.f32(pl.input_ptr, x.numel())[:] = x.reshape(-1).double()
    mem.f32(pl.outputs[0]["target_ptr"], y.numel())[:] = y.reshape(-1).double()
    run_phase(mem, pl, 0)
    run_phase(mem, pl, 1)
    return mem.f32(pl.g_ptr, pl.arena_elems)


def _data():
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.random((4, 16, 16, 1), dtype=np.float32))
    y = (x > 0.5).float()
    return x, y


def _worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tf-1d-2d-segmentation-end2endpipelines_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b2seg.dist import allreduce_flat_, shard_range, wait_all
    from desc_emulator import run_phase
    x, y = _data()
    b, e = shard_range(x.shape[0], rank, world)
    g, mem, pl = _build(e - b, adam_bucket_bytes=4096)   # data-parallel plan: one Adam op per exchange bucket
    xs, ys = x[b:e], y[b:e]
    mem.f32(pl.input_ptr, xs.numel())[:] = xs.reshape(-1).double()
    mem.f32(pl.outputs[0]["target_ptr"], ys.numel())[:] = ys.reshape(-1).double()
    run_phase(mem, pl, 0)
    grads = mem.f32(pl.g_ptr, pl.arena_elems)
    sched = pl.exchange_schedule(bucket_bytes=4096)
    assert len(sched) >= 3 and sched[0][0] < len(pl.ops[1])          # the exchange really starts before backward ends
    covered = sorted((lo, hi) for (_, lo, hi) in sched)
    assert covered[0][0] == 0 and covered[-1][1] == pl.arena_elems and all(a[1] == b_[0] for a, b_ in zip(covered, covered[1:]))
    done = 0
    works = []
    for (n_ops, lo, hi) in sched:
        run_phase(mem, pl, 1, done, n_ops - done)
        done = n_ops
        works.append(allreduce_flat_(grads[lo:hi]))
    run_phase(mem, pl, 1, done, len(pl.ops[1]) - done)
    # optimizer phase of a data-parallel plan: bucket i's Adam as soon as bucket i's all-reduce has landed (Model._step)
    assert len(pl.ops[2]) == len(sched)
    for i, ((op, desc, _n), (_ops, lo, hi)) in enumerate(zip(pl.ops[2], sched)):
        assert desc.n == hi - lo and desc.g == pl.g_ptr + 4 * lo
        desc.grad_scale = 1.0 / world
        wait_all(works[i])
        run_phase(mem, pl, 2, i, 1)
    torch.save(mem.f32(pl.w_ptr, pl.arena_elems).clone(), os.path.join(out, f"w{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_data_parallel_step(tmp_path):
    from b2seg.dist import shard_range, shard_sizes
    assert shard_sizes(5, 2) == [3, 2] and shard_range(5, 1, 2) == (3, 5)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w0, w1 = torch.load(tmp_path / "w0.pt"), torch.load(tmp_path / "w1.pt")
    assert torch.equal(w0, w1)
    # single-process reference: average the per-replica gradients by hand
    from desc_emulator import run_phase
    x, y = _data()
    gs = []
    for r in range(2):
        g, mem, pl = _build(2)
        gs.append(_fwd_bwd(mem, pl, x[2 * r:2 * r + 2], y[2 * r:2 * r + 2]).clone())
    g, mem, pl = _build(2)
    _fwd_bwd(mem, pl, x[:2], y[:2])
    mem.f32(pl.g_ptr, pl.arena_elems)[:] = (gs[0] + gs[1]) / 2
    run_phase(mem, pl, 2)
    ref = mem.f32(pl.w_ptr, pl.arena_elems)
    assert torch.allclose(w0, ref, atol=1e-12)
    assert float((w0 - ref).abs().max()) < 1e-12


def test_ops_overlapping_exchange_replay():
    """host-side replay that decides which backward ops leave SMs to the all-reduce (Engine.reserve_sms_for_exchange)"""
    from b2seg.dist import ops_overlapping_exchange
    op_ms = [1.0] * 10                                    # ten 1 ms ops
    # one 1 MB bucket ready after op 3 has finished (t = 4 ms) at 1 ms per MB -> on the wire during [4, 5.03]: ops 4 and 5, and op 3
    # whose end touches the window within the 0.05 ms margin
    res, win = ops_overlapping_exchange(op_ms, [(4, 0, 262144)], 1.0 / (1 << 20), 1.0)
    assert abs(win[0][0] - 4.0) < 1e-9 and abs(win[0][1] - 5.03) < 1e-6
    assert res == [3, 4, 5]
    # a second bucket ready at the same time queues behind the first; the slowdown stretches the reserved ops (op 3 now ends at 4.25,
    # which is when the wire starts) and the fixed point is stable
    res2, win2 = ops_overlapping_exchange(op_ms, [(4, 0, 262144), (4, 262144, 524288)], 1.0 / (1 << 20), 1.25)
    assert abs(win2[0][0] - 4.25) < 1e-9 and win2[1][0] == win2[0][1] and res2 == [3, 4, 5]
    # nothing to exchange -> nothing reserved; a bucket that is only ready after the last op reserves nothing but the tail
    assert ops_overlapping_exchange(op_ms, [], 1e-6, 1.2)[0] == []
    assert ops_overlapping_exchange(op_ms, [(10, 0, 1024)], 1e-9, 1.2)[0] == [9]


# ---- the same through the facade: Model.distribute() + Model._step's data-parallel branch (the host code that runs on 8 GPUs) ----------
def _facade_model():
    from b2seg.model import Adam
    from b2seg.models2d import unet_model_builder
    m = unet_model_builder("UNetPP", 16, 16, 8, 2, num_channels=1, ds=1, ag=1, train_mode="from_scratch").ResNet50()
    m.compile(loss={"out": "bce", "level1": "mse", "level2": "mse"}, optimizer=Adam(1e-2), ds_targets="UNetPP")
    m.exchange_bucket_bytes = 4096
    return m


def _facade_worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tf-1d-2d-segmentation-end2endpipelines_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import b2seg.engine
    from cpu_engine import CpuEngine
    b2seg.engine.Engine = CpuEngine
    from b2seg.dist import shard_range
    x, y = _data()
    b, e = shard_range(x.shape[0], rank, world)
    m = _facade_model().distribute()
    eng = m._engine(e - b, True)
    assert len(eng.planner.ops[2]) == len(eng.planner.exchange_schedule(4096)) >= 3      # a data-parallel plan: one Adam op per bucket
    if rank == 1:                                # rank 1 starts from different weights: broadcast_weights must repair that
        eng.w.mul_(1.5)
    m.broadcast_weights(0)
    loss = m.train_on_batch(x[b:e].numpy(), y[b:e].numpy())
    assert eng.planner.shard == (rank, world)    # sharded optimizer: each rank updated its slice of every bucket, the kernels' copy was all-gathered
    wb = eng.wb.clone()
    own = torch.cat([eng.w[lo + rank * ((hi - lo) // world):lo + (rank + 1) * ((hi - lo) // world)] for (_n, lo, hi) in m._exchange_schedule(eng)])
    m.gather_master_weights()                    # collective: the fp32 masters of the other ranks' slices
    assert torch.equal(wb, eng.w)                # (the emulator's "bf16" copy is float64: identical to the gathered masters)
    loss2 = m.train_on_batch(x[b:e].numpy(), y[b:e].numpy())    # a second step starts from the all-gathered weights
    m.gather_master_weights()
    torch.save(dict(w=eng.w.clone(), loss=loss, loss2=loss2, own=own), os.path.join(out, f"facade{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_step_through_the_facade(tmp_path, monkeypatch):
    """Model.distribute / broadcast_weights / _step with world_size 2 over gloo, each rank on the emulator engine: identical weights on
    both ranks, equal to one process that averages the two replicas' gradients by hand and applies one Adam step"""
    port = _free_port()
    mp.spawn(_facade_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "facade0.pt"), torch.load(tmp_path / "facade1.pt")
    assert torch.equal(r0["w"], r1["w"])
    import b2seg.engine
    from cpu_engine import CpuEngine
    monkeypatch.setattr(b2seg.engine, "Engine", CpuEngine)
    x, y = _data()
    grads = []
    for (b, e) in ((0, 2), (2, 4)):
        m = _facade_model()
        eng = m._engine(e - b, True)
        eng.x_dev.copy_(x[b:e])
        eng.outputs[0]["target"].copy_(y[b:e])
        eng.derive_targets()
        eng.forward()
        eng.backward()
        grads.append(eng.g.clone())
    eng.g[:] = (grads[0] + grads[1]) / 2
    eng.optimizer_step(1e-2, 1.0)
    # second step of the reference: both replicas from the updated weights
    w1 = eng.w.clone()
    grads = []
    for (b, e) in ((0, 2), (2, 4)):
        eng.x_dev.copy_(x[b:e])
        eng.outputs[0]["target"].copy_(y[b:e])
        eng.derive_targets()
        eng.w.copy_(w1)
        eng.wb.copy_(w1)
        eng.forward()
        eng.backward()
        grads.append(eng.g.clone())
    eng.w.copy_(w1)
    eng.g[:] = (grads[0] + grads[1]) / 2
    eng.optimizer_step(1e-2, 1.0)
    assert torch.allclose(r0["w"], eng.w, atol=1e-12, rtol=0), float((r0["w"] - eng.w).abs().max())
    assert r0["loss2"] < r0["loss"] or r1["loss2"] < r1["loss"]
