"""world_size-2 gloo tests (CPU) of the data-parallel host logic: ray sharding covers the frame exactly,
and summing equal-shard local-mean gradients then scaling by 1/world equals the global-batch gradient
of the reference loss (train.py:460) -- checked with the oracle as the per-rank "kernel"."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_shard_range_partitions_exactly():
    par = importlib.import_module("sfm-nerf_b200.parallel")
    for n in (0, 1, 7, 640000, 160000, 10001):
        for world in (1, 2, 3, 4, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert par.shard_range(640000, 3, 8) == (240000, 320000)          # C4: 80,000 rays per GPU
    with pytest.raises(ValueError):
        par.shard_range(10, 2, 2)
    seen = {par.batch_counter(it, r, 4) for it in range(10) for r in range(4)}
    assert len(seen) == 40


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    par = importlib.import_module("sfm-nerf_b200.parallel")
    from oracle import nerf_oracle as orc
    sd = orc.init_state_dict(0)
    gen = torch.Generator().manual_seed(9)
    cam = orc.make_camera_info(32, 32)
    pose = orc.look_at_pose(10.0, -25.0)
    px = torch.randint(0, 32, (8, 2), generator=gen).numpy()
    rays = orc.rays_from_indices(np.zeros((1, 32, 32, 3)), pose[None], cam, np.zeros(8, int), px[:, 0], px[:, 1])
    o, d = torch.tensor(rays[:, 0:3]), torch.tensor(rays[:, 3:6])
    target = torch.rand(8, 3, generator=gen)
    b, e = par.shard_range(8, rank, world)
    _, g_local = orc.loss_and_grads(sd, o[b:e], d[b:e], target[b:e], 16)
    flat = torch.cat([g.reshape(-1) for g in g_local.values()])
    par.allreduce_sum_(flat)
    flat *= par.mean_scale(world)
    # the exchange object on a CPU group: no symmetric memory -> the NCCL/gloo allreduce mode, double-buffered by step parity
    ex = par.GradExchange(flat.numel(), "cpu", world)
    assert ex.mode == "nccl" and ex.handle is None
    assert ex.local(0).data_ptr() != ex.local(1).data_ptr() and ex.local(0).data_ptr() == ex.local(2).data_ptr()
    assert ex.local(1).numel() == flat.numel() and (ex.local(1).data_ptr() - ex.local(0).data_ptr()) % 256 == 0
    ex.barrier()                                     # no-op in this mode
    # frame assembly
    frame = par.gather_frame(torch.arange(b, e, dtype=torch.float32).unsqueeze(1), 8)
    if rank == 0:
        _, g_full = orc.loss_and_grads(sd, o, d, target, 16)
        full = torch.cat([g.reshape(-1) for g in g_full.values()])
        out.put((float((flat - full).abs().max()), float(full.abs().max()), frame.squeeze(1).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_global_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale, frame = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err <= 1e-6 * max(scale, 1.0) + 1e-9, (err, scale)
    assert frame == [float(i) for i in range(8)]
