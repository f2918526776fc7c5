#!/usr/bin/env python
"""bench.py -- headline benchmark: NeRF train-step (and render) rays/s at 64 samples/ray.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): lego-shaped synthetic scene, 400x400 views, train step of
4096 rays x 64 stratified samples (jittered), live PE model (661,636 params), Adam.  A "step" is
generateBatch + render + MSE + backward + Adam + weight repack.  `value` = rays/s of the whole job
with every input resident in HBM (rays are drawn on the GPU); `e2e` = the same step fed from a pinned
HOST ray batch (H2D copy inside the timed region) with the loss read back to the host every step,
i.e. the reference's train.py:447-468 loop body.  N > 1: data-parallel, 4096 rays per rank (weak
scaling); the gradient exchange is fused into the Adam kernel (NVLS multimem.ld_reduce / peer loads over symmetric memory,
NCCL allreduce as fallback).  Timing: CUDA events on the launch stream,
barrier + synchronize on both sides, max over ranks; the headline is the MEDIAN of WINDOWS windows of
`--steps` steps each (a single 20-step window is 26 ms: one hiccup of one rank would decide it), every
window's per-rank times are in the line.  The per-step working set (activation stash +
gradient stash, ~2.8 GB) is far larger than the 126 MB L2, so no explicit L2 flush is needed.

Both arms start a step from a generateBatch() result in HOST memory ([n_rays, 9] fp32): the GPU `e2e` arm uploads
it inside the timed region, the CPU arm (`--impl reference`, `cpu_baseline`) turns it into a tensor inside the timed
region; producing the batch (the reference's per-ray Python loop, train.py:104-127) is outside both timed regions
and reported separately (`generate_batch_ms`).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, N_SAMPLE, IMG = 4096, 64, 400
FLOP_FWD = 2 * 658944            # per sample, unpadded, biases/activations excluded
FLOP_DGRAD = 2 * 623232
FLOP_WGRAD = 2 * 658944
BYTES_COMPOSITE_FWD = 20         # sigma 4 + rgb 12 + z 4 per sample (jittered)
BYTES_COMPOSITE_BWD = 36
# algorithmic HBM bytes per sample of the training kernels (DESIGN.md "Data layout"): bf16 stash blocks
# [PE 128 | DE 64 | h0..h8, feature 10 x 512 | r1 256] + 11 ReLU bit masks x 32 B; dY stash 10 x 512 + 256 + 16
BYTES_FWD_TRAIN = 128 + 64 + 10 * 512 + 256 + 11 * 32 + 16 + 4
BYTES_DGRAD = 11 * 32 + 16 + 16 + 10 * 512 + 256 + 16
# wgrad, MINIMAL count (every operand block once): X = PE 128 + DE 64 + h0..h8 9 x 512 + feature 512 + r1 256 = 5568;
# dY = 10 x 512 + 256 + 16 = 5392.  What the kernels actually move on top of this is reported as traffic / algorithmic.
BYTES_WGRAD = (128 + 64 + 9 * 512 + 512 + 256) + (10 * 512 + 256 + 16)
WINDOWS = 5


def workload_config(n_rays, world):
    return {"workload": f"lego-shaped synthetic scene {IMG}x{IMG} (20 views), train step {n_rays} rays x {N_SAMPLE} samples per GPU, "
                        "live PE model 661,636 params, Adam",
            "samples_per_ray": N_SAMPLE, "parallelism": f"dp{world}" if world > 1 else "single",
            "l2": "working set (stash+dstash ~2.8 GB/step) >> 126 MB L2; no explicit flush"}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rays", type=int, default=N_RAYS)
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"tflops": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "tflops_burst": float(d["bf16_tflops"]),
                "hbm": float(d["hbm_gbs"]), "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port of train.py:447-465) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_steps(n_rays, steps, warmup):
    """Times `steps` train steps of the reference's loop body (train.py:447-465: tensor from the generateBatch() array, render,
    nn.MSELoss, backward, Adam step, StepLR step, loss.item()) on the host cores.  kind "reference": the UNMODIFIED reference
    from oracle/_ref (its own generateBatch / render / NeRFmodel + torch.optim.Adam); kind "port": the oracle restatement, only
    when oracle/_ref was not vendored.  Returns (step seconds, generateBatch seconds, threads, kind)."""
    import argparse as _ap
    import numpy as np
    import torch
    from oracle import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(0)
    times, gen_times = [], []
    if ref_loader.available():
        import random
        ref = ref_loader.import_reference()
        from oracle import nerf_oracle as orc
        cam = orc.make_camera_info(IMG, IMG)
        poses = orc.sphere_poses(8, seed=0)
        images = rng.random((8, IMG, IMG, 3))          # colours only feed the loss; content is irrelevant to timing
        args = _ap.Namespace(n_sample=N_SAMPLE, perturb=True, chunk_size=1024 * 32, n_rays_batch=n_rays, near=2.0, far=6.0)
        torch.manual_seed(0)
        random.seed(0)
        model = ref.NeRFmodel()                                                     # CPU: render follows its inputs' device
        optimizer = torch.optim.Adam(model.parameters(), lr=5e-4)                  # train.py:401
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=50000, gamma=0.1)
        for it in range(warmup + steps):
            tg = time.perf_counter()
            batch = ref.generateBatch(images, poses, cam, args)                    # train.py:447 (outside the timed region, both arms)
            tg = time.perf_counter() - tg
            t0 = time.perf_counter()
            rays_batch = torch.tensor(batch)                                        # train.py:448
            optimizer.zero_grad()
            rendered = ref.render(model, rays_batch[:, :3], rays_batch[:, 3:6], args, near=args.near, far=args.far)
            loss = torch.nn.MSELoss()(rendered, rays_batch[:, 6:9])
            loss.backward()
            optimizer.step()
            scheduler.step()
            _ = loss.item()                                                          # train.py:468
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt); gen_times.append(tg)
        return times, gen_times, torch.get_num_threads(), "reference"
    from oracle import nerf_oracle as orc
    sd = orc.init_state_dict(0)
    opt = orc.Adam(sd)
    cam = orc.make_camera_info(IMG, IMG)
    poses = orc.sphere_poses(8, seed=0)
    images = rng.random((8, 16, 16, 3))
    gen = torch.Generator().manual_seed(0)
    for it in range(warmup + steps):
        k = rng.integers(0, 8, n_rays); x = rng.integers(0, IMG, n_rays); y = rng.integers(0, IMG, n_rays)
        tg = time.perf_counter()
        batch = np.empty((n_rays, 9), np.float32)
        for i in range(n_rays):
            o, d = orc.pixel_to_ray(cam, poses[k[i]], (int(x[i]), int(y[i])))
            batch[i, 0:3], batch[i, 3:6], batch[i, 6:9] = o, d, images[k[i]][y[i] % 16, x[i] % 16]
        tg = time.perf_counter() - tg
        t0 = time.perf_counter()
        o, d, tgt = (torch.tensor(batch[:, j:j + 3]) for j in (0, 3, 6))
        t_rand = torch.rand(n_rays, N_SAMPLE, generator=gen)
        loss, grads = orc.loss_and_grads(sd, o, d, tgt, N_SAMPLE, t_rand=t_rand)
        sd = opt.step(sd, grads)
        _ = loss.item()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt); gen_times.append(tg)
    return times, gen_times, torch.get_num_threads(), "port"


def cpu_sample_text(kind, n, n_rays, threads):
    what = ("unmodified reference from oracle/_ref: generateBatch array -> ref.render + nn.MSELoss + backward + torch.optim.Adam + StepLR"
            if kind == "reference" else "oracle port of train.py:447-465")
    return (f"{n} train steps of {n_rays} rays x {N_SAMPLE} samples ({what}, torch CPU fp32, {threads} threads; "
            "generateBatch's per-ray Python loop is outside the timed region of both arms and reported as generate_batch_ms)")


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    # calibrate the bounded sample: keep the whole run within ~2.5 minutes
    cal, _, threads, kind = cpu_train_steps(256, 1, 1)
    per_ray = cal[0] / 256
    budget = 150.0 / max(1, args.steps + args.warmup)
    n_rays = int(min(args.rays, max(256, budget / per_ray)))
    n_rays -= n_rays % 64
    times, gen_times, threads, kind = cpu_train_steps(n_rays, args.steps, args.warmup)
    total = sum(times)
    value = n_rays * len(times) / total
    line = {
        "metric": "train_step_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": workload_config(args.rays, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": kind,
                         "sample": cpu_sample_text(kind, len(times), n_rays, threads),
                         "generate_batch_ms": 1e3 * sum(gen_times) / len(gen_times)},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def exchange_check(s, E, dist, dev, images, poses, cam, rank, world):
    """Correctness of the data-parallel gradient exchange, run before anything is timed (world > 1): from identical initial
    weights, 3 steps with the fused exchange (NVLS multimem.ld_reduce / peer loads inside the Adam kernel, CUDA-graph replay)
    and 3 steps with NCCL's allreduce -- replicas must stay bit-identical across ranks in both, the two must agree to within
    the fp32 summation order (3 Adam steps of lr 5e-4), and the graph capture must not have fallen back.  Raises on violation."""
    import torch
    out = {}
    flats = {}
    for mode in ("fused", "nccl"):
        old = os.environ.get("NERF_GRAD_EXCHANGE")
        if mode == "nccl":
            os.environ["NERF_GRAD_EXCHANGE"] = "nccl"
        torch.manual_seed(0)
        m = s.NeRFmodel().to(dev)
        tr = E.Trainer(m, images, poses, cam, 1024, N_SAMPLE, perturb=False, seed=3, rank=rank, world=world)
        if mode == "nccl":
            if old is None:
                os.environ.pop("NERF_GRAD_EXCHANGE", None)
            else:
                os.environ["NERF_GRAD_EXCHANGE"] = old
        for _ in range(3):
            tr.step()
        torch.cuda.synchronize()
        flat = tr.st.flat.detach().clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        identical = all(torch.equal(gathered[0], g) for g in gathered)
        flats[mode] = flat
        out[mode] = {"exchange_mode": tr.exchange.mode, "replicas_bit_identical": identical, "graphs": len(tr._graphs),
                     "graph_error": tr.graph_error, "loss": float(tr.loss.item())}
        tr.close()
        del tr, m
    diff = float((flats["fused"] - flats["nccl"]).abs().max())
    moved = float((flats["fused"] - torch.cat([p.detach().flatten() for p in _fresh_params(s, dev)])).abs().max())
    out["max_abs_weight_diff_fused_vs_nccl_after_3_steps"] = diff
    out["max_abs_weight_change_after_3_steps"] = moved
    ok = (out["fused"]["replicas_bit_identical"] and out["nccl"]["replicas_bit_identical"] and diff < 5e-4 and moved > 5e-4
          and out["fused"]["graph_error"] == "" and (out["fused"]["exchange_mode"] == "nccl" or out["fused"]["graphs"] > 0))
    out["ok"] = bool(ok)
    return out


def _fresh_params(s, dev):
    import torch
    torch.manual_seed(0)
    return s.NeRFmodel().to(dev).ordered_parameters()


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import sfmnerf_b200 as s
    E = s.engine

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    s._lib.check(s._lib.load().nerf_device_check(), "nerf_device_check")
    n_rays = args.rays
    K = args.steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def per_rank(values):
        """[len(values)] per rank -> [world][len(values)] on every rank."""
        t = torch.tensor(values, dtype=torch.float64, device=dev)
        if world == 1:
            return [t.tolist()]
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [p.tolist() for p in parts]

    def timed_windows(*step_fns):
        """WINDOWS windows of K steps per step function, each bracketed by barrier + synchronize; per-window time = max over
        ranks.  Several step functions are timed in ALTERNATING windows (a, b, a, b, ...), so the slow climb of the step time on
        a power-capped board (5-8 % over the first hundred steps) enters their medians alike."""
        mine = [[] for _ in step_fns]
        for w in range(WINDOWS):
            for f, step_fn in enumerate(step_fns):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(K):
                    step_fn(w * K + i)
                e1.record()
                barrier()
                mine[f].append(e0.elapsed_time(e1))
        out = []
        for f in range(len(step_fns)):
            table = per_rank(mine[f])                                    # [world][WINDOWS]
            win = [max(table[r][w] for r in range(world)) for w in range(WINDOWS)]
            med = sorted(win)[WINDOWS // 2]
            rank_ms = [sorted(table[r])[WINDOWS // 2] / K for r in range(world)]          # per-rank median step time
            out.append((med, {"windows": WINDOWS, "window_ms_per_step": [round(x / K, 5) for x in win],
                              "per_rank_ms_per_step": {"min": round(min(rank_ms), 5), "median": round(sorted(rank_ms)[world // 2], 5),
                                                       "max": round(max(rank_ms), 5), "slowest_rank": int(np.argmax(rank_ms))}}))
        return out[0] if len(out) == 1 else out

    # ---- synthetic scene (data generation, untimed) ---------------------------------------
    cam = s.synthetic.make_camera_info(IMG, IMG)
    poses_np = s.synthetic.sphere_poses(20, seed=0)
    images = s.synthetic.render_lego_images(poses_np, cam, n_sample=128, device=dev).contiguous()
    poses = E.poses_to_device(poses_np, dev)

    # ---- data parallel: the gradient exchange is checked before it is timed ----------------
    exch_check = None
    if world > 1:
        exch_check = exchange_check(s, E, dist, dev, images, poses, cam, rank, world)
        if not exch_check["ok"]:
            if rank == 0:
                print(json.dumps({"error": "gradient exchange check failed", "grad_exchange": {"check": exch_check}}), flush=True)
            dist.destroy_process_group()
            sys.exit(3)

    torch.manual_seed(0)
    model = s.NeRFmodel().to(dev)
    trainer = E.Trainer(model, images, poses, cam, n_rays, N_SAMPLE, perturb=True, seed=0, rank=rank, world=world)

    # ---- device-resident train steps ------------------------------------------------------
    # Trainer.step() = ONE nerf_train_step call, replayed as a CUDA graph from the second step on (world > 1: one graph per
    # gradient-buffer parity with the cross-rank barrier captured inside).
    warm = max(args.warmup, 3)
    for _ in range(warm):
        trainer.step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- end to end through the public API with host buffers: prepared here, timed in windows that alternate with the device-resident ones
    host_batches = []
    rng = np.random.default_rng(1234 + rank)
    for _ in range(4):
        k = torch.from_numpy(rng.integers(0, len(poses_np), n_rays).astype(np.int32)).to(dev)
        x = torch.from_numpy(rng.integers(0, IMG, n_rays).astype(np.int32)).to(dev)
        y = torch.from_numpy(rng.integers(0, IMG, n_rays).astype(np.int32)).to(dev)
        o, d, tgt = E.raygen_batch(poses, cam, k, x, y, images)          # a generateBatch() result ...
        host_batches.append(torch.cat([o, d, tgt], 1).cpu().pin_memory())  # ... living in pinned host memory
    loss_box = [0.0]

    def host_step(i):
        # pinned host batch n_rays x 9 fp32 (train.py:448) -> H2D inside step(); the loss comes back as a Python float every
        # step (train.py:468).  Forward graph, async D2H of the loss, backward graph; the host waits for the loss only.
        loss_box[0] = trainer.step(rays=host_batches[i % 4], loss_to_host=True)

    for i in range(3):
        host_step(i)
    for _ in range(3):
        trainer.step()
    (ms_total, win_info), (e2e_ms, e2e_info) = timed_windows(lambda i: trainer.step(), host_step)
    loss_host = loss_box[0]
    launches = trainer.launches_per_step * K * WINDOWS            # this repo's kernels only (the gradient memset is the driver's)
    graph_replay = len(trainer._graphs) > 0
    graph_error = trainer.graph_error
    if graph_error or not graph_replay:                           # the headline path is the graph-replayed step: never time a silent fallback
        if rank == 0:
            print(json.dumps({"error": "CUDA-graph capture of nerf_train_step failed; the step ran as direct launches", "graph_error": graph_error}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        sys.exit(4)
    # per-kernel durations: the same K steps again with the same kernels launched one by one, each bracketed by CUDA events on
    # the launch stream (events cannot be read from inside a graph replay); 3 untimed steps first (the first launches of the
    # per-kernel sequence pay one-off costs: lazy module load of the un-graphed barrier kernel, allocator growth)
    for _ in range(3):
        trainer.step(timers={})
    barrier()
    timers = {}
    ev0i, ev1i = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0i.record()
    for _ in range(K):
        trainer.step(timers=timers)
    ev1i.record()
    barrier()
    ms_instrumented = max_over_ranks(ev0i.elapsed_time(ev1i))
    grad_exchange = "none" if trainer.exchange is None else trainer.exchange.mode
    grad_exchange_note = "" if trainer.exchange is None else trainer.exchange.why
    final_loss = float(trainer.loss.item())
    barrier_wait = None
    if "adam_reduce" in timers:
        mine = sum(a.elapsed_time(b) for a, b in timers["adam_reduce"]) / len(timers["adam_reduce"])
        tab = [r[0] for r in per_rank([mine])]
        barrier_wait = {"adam_reduce_ms_per_rank": [round(x, 4) for x in tab],
                        "note": "duration of the reduce-Adam launch per rank (per-kernel pass): the cross-rank barrier is fused into it, so "
                                "a rank that arrives early waits here for the slowest one; the minimum over ranks is the cost of "
                                "barrier + NVLink reduction + update alone"}

    clocks = sampler.result()

    # ---- render throughput (inference, no stash) for the same metric family -----------------
    args_r = argparse.Namespace(n_sample=N_SAMPLE, perturb=True, chunk_size=32768, n_rays_batch=IMG * IMG, near=2.0, far=6.0)
    r_begin, r_end = s.parallel.shard_range(IMG * IMG, rank, world)
    per_rank_rays = r_end - r_begin

    def time_render(mdl, cam_, n_samp, begin, count, frames):
        a_ = argparse.Namespace(n_sample=n_samp, perturb=True, chunk_size=32768, n_rays_batch=count, near=2.0, far=6.0)
        with torch.no_grad():
            for _ in range(2):
                o_, d_ = E.raygen_image(poses, 0, cam_, begin, count)
                s.render(mdl, o_, d_, a_)
            barrier()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            for f in range(frames):
                o_, d_ = E.raygen_image(poses, f % len(poses_np), cam_, begin, count)
                s.render(mdl, o_, d_, a_)
            eb.record()
            barrier()
        return max_over_ranks(ea.elapsed_time(eb)) / frames

    n_frames = 10
    render_ms = time_render(model, cam, N_SAMPLE, r_begin, per_rank_rays, n_frames) * n_frames
    # BASELINE config 4: full-resolution 800x800 frame x 192 samples, rays sharded contiguously over the ranks
    cam4 = s.synthetic.make_camera_info(800, 800)
    b4, e4 = s.parallel.shard_range(800 * 800, rank, world)
    c4_ms = time_render(model, cam4, 192, b4, e4 - b4, 4)
    # the fused MLP inference kernel alone on this rank's slice of the frame (tensor-pipe fraction)
    with torch.no_grad():
        st = model.kernel_state()
        o, d = E.raygen_image(poses, 0, cam, r_begin, per_rank_rays)
        zr = E.sample_z(per_rank_rays, N_SAMPLE, 2.0, 6.0, dev, perturb=True, seed=5)
        for _ in range(2):
            E.mlp_forward_rays(st, o, d, zr, None)
        ev6, ev7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev6.record()
        for _ in range(5):
            E.mlp_forward_rays(st, o, d, zr, None)
        ev7.record()
        torch.cuda.synchronize()
    mlp_infer_ms = ev6.elapsed_time(ev7) / 5
    # K5 alone (HBM-bound; 20 B/sample forward, 36 B/sample backward -- SURVEY 8d) on this rank's share of 640,000 rays x 64
    # samples: large enough that a launch (0.1-0.3 ms) is not hidden behind the host's per-call overhead
    n_comp = (e4 - b4)
    with torch.no_grad():
        sr = torch.rand(n_comp, N_SAMPLE, 4, device=dev)
        sr[..., 0] *= 4.0
        zc = E.sample_z(n_comp, N_SAMPLE, 2.0, 6.0, dev, perturb=True, seed=9)
        g_rgb = torch.rand(n_comp, 3, device=dev)
        comp_ms = {}
        for name, fn in (("forward", lambda: E.composite_forward(sr, zc, extras=True)), ("backward", lambda: E.composite_backward(sr, zc, g_rgb))):
            for _ in range(3):
                fn()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            for _ in range(20):
                fn()
            eb.record()
            torch.cuda.synchronize()
            comp_ms[name] = ea.elapsed_time(eb) / 20
        del sr, g_rgb, zc
    render_rays_s = IMG * IMG * n_frames / render_ms * 1e3
    trainer.close()

    # ---- BASELINE config 3: the no-positional-encoding variant, render + train, 1 GPU (rank 0's numbers at N > 1) ----------
    # (a) the live class with pos_encoding=False (NeRFModel.py:113-114, 172-174; 627,844 params): train step + render;
    # (b) a checkpoint in the layout of Phase2/checkpoints_no_pos_encoding imported through legacy.py: render.
    torch.manual_seed(0)
    model_c3 = s.NeRFmodel(pos_encoding=False).to(dev)
    tr3 = E.Trainer(model_c3, images, poses, cam, n_rays, N_SAMPLE, perturb=True, seed=0, rank=0, world=1)
    for _ in range(3):
        tr3.step()
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(K):
        tr3.step()
    eb.record()
    torch.cuda.synchronize()
    c3_train_ms = ea.elapsed_time(eb) / K
    c3_loss = float(tr3.loss.item())
    tr3.close()
    c3_render_ms = time_render(model_c3, cam, N_SAMPLE, r_begin, per_rank_rays, 5)
    gen3 = torch.Generator().manual_seed(3)
    legacy_sd = {}
    for name, (o_f, i_f) in (("fc1.0", (256, 3)), *[(f"block1.{k}", (256, 256)) for k in (0, 2, 4, 6, 8, 10)], ("block1.12", (257, 256)),
                             ("density_fc", (1, 256)), ("rgb_fc.0", (128, 259)), ("rgb_fc.2", (3, 128))):
        bound = 1.0 / np.sqrt(i_f)                                # nn.Linear's default init range
        legacy_sd[name + ".weight"] = (torch.rand(o_f, i_f, generator=gen3) * 2 - 1) * bound
        legacy_sd[name + ".bias"] = (torch.rand(o_f, generator=gen3) * 2 - 1) * bound
    model_legacy = s.load_legacy_checkpoint(legacy_sd, dev)
    c3_legacy_ms = time_render(model_legacy, cam, N_SAMPLE, r_begin, per_rank_rays, 5)
    del tr3, model_c3, model_legacy

    # BASELINE config 1 (the reference's CPU-runnable case): test_image of one 100x100 view x 64 deterministic samples through the
    # public API -- host image in, prediction + MSE (a host float) out; rank 0 only
    c1 = None
    if rank == 0:
        cam1 = s.synthetic.make_camera_info(100, 100)
        img1 = np.random.default_rng(7).random((100, 100, 3)).astype(np.float32)
        args_c1 = argparse.Namespace(n_sample=N_SAMPLE, perturb=False, chunk_size=32768, n_rays_batch=4096, near=2.0, far=6.0)
        torch.manual_seed(0)
        model_c1 = s.NeRFmodel().to(dev)              # random-init weights, as the config says (the trained model above renders black)
        for _ in range(2):
            pred1, mse1 = s.test_image(model_c1, img1, poses_np[0], cam1, args_c1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            pred1, mse1 = s.test_image(model_c1, img1, poses_np[0], cam1, args_c1)
        torch.cuda.synchronize()
        c1 = {"gpu_ms": 1e3 * (time.perf_counter() - t0) / 5, "pred": pred1.cpu(), "mse": mse1, "img": img1, "cam": cam1,
              "sd": {k: v.detach().cpu().clone() for k, v in model_c1.state_dict().items()}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel averages and roofline -------------------------------------------------
    peaks = measured_peaks()
    rows = n_rays * N_SAMPLE
    work = {"mlp_fwd": rows * FLOP_FWD, "mlp_dgrad": rows * FLOP_DGRAD, "mlp_wgrad": rows * FLOP_WGRAD}
    traffic_alg = {"mlp_fwd": rows * BYTES_FWD_TRAIN, "mlp_dgrad": rows * BYTES_DGRAD, "mlp_wgrad": rows * BYTES_WGRAD,
                   "composite_train": rows * BYTES_COMPOSITE_BWD}     # forward + MSE + backward in one launch: 20 B read + 16 B written
    kern = {}
    for name, pairs in timers.items():
        ms = sum(a.elapsed_time(b) for a, b in pairs) / len(pairs)
        kern[name] = {"ms": round(ms, 4)}
        if name in work:
            kern[name]["tflops"] = round(work[name] / ms / 1e9, 1)
            kern[name]["tensor_frac"] = round(work[name] / ms / 1e9 / peaks["tflops"], 4)
        if name in traffic_alg:
            kern[name]["gbs"] = round(traffic_alg[name] / ms / 1e6, 1)
            kern[name]["hbm_frac"] = round(traffic_alg[name] / ms / 1e6 / peaks["hbm"], 4)
        if name == "adam_reduce":      # every rank pulls the (world - 1) remote copies of the 2.6 MB gradient through NVLink while it updates
            n_par = trainer.st.n_params
            kern[name]["nvlink_bytes_in_per_rank"] = 4 * n_par * (world - 1)
            kern[name]["nvlink_gbs_in_per_rank"] = round(4 * n_par * (world - 1) / ms / 1e6, 1)
    dom = max(work, key=lambda k: kern[k]["ms"])
    measured = {}
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        measured = json.load(open(tpath))
    k = kern[dom]
    if k["hbm_frac"] >= k["tensor_frac"]:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": k["gbs"], "peak": peaks["hbm"], "unit": "GB/s", "frac": k["hbm_frac"],
                    "bytes_per_sample": traffic_alg[dom] // rows, "algorithmic_bytes": traffic_alg[dom]}
    else:
        roofline = {"bound": "tensor", "kernel": dom, "achieved": k["tflops"], "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": k["tensor_frac"], "flop_per_sample": work[dom] // rows}
    traffic = measured.get(dom)
    roofline.update({"traffic": traffic, "traffic_over_algorithmic": (round(traffic / traffic_alg[dom], 4) if traffic else None),
                     "peak_source": peaks["src"], "tensor_frac": k["tensor_frac"], "hbm_frac": k["hbm_frac"],
                     "step_tflops": round(rows * (FLOP_FWD + FLOP_DGRAD + FLOP_WGRAD) * K / ms_total / 1e9, 1),
                     "step_tensor_frac": round(rows * (FLOP_FWD + FLOP_DGRAD + FLOP_WGRAD) * K / ms_total / 1e9 / peaks["tflops"], 4)})

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        times, gen_times, threads, kind = cpu_train_steps(n_rays, 3, 1)
        v = n_rays * len(times) / sum(times)
        cpu_baseline = {"value": v, "unit": "rays/s", "cores": threads, "kind": kind, "sample": cpu_sample_text(kind, len(times), n_rays, threads),
                        "generate_batch_ms": 1e3 * sum(gen_times) / len(gen_times)}

    render_c1 = None
    if c1 is not None:
        render_c1 = {"workload": "test_image: one 100x100 view x 64 deterministic samples, random-init weights, wall clock incl. host<->device copies",
                     "gpu_ms": c1["gpu_ms"], "gpu_rays_per_s": 1e4 / c1["gpu_ms"] * 1e3}
        if world == 1 and not args.no_cpu_baseline:
            import torch as _t
            from oracle import ref_loader
            _t.set_num_threads(os.cpu_count() or 1)
            args_c1 = argparse.Namespace(n_sample=N_SAMPLE, perturb=False, chunk_size=32768, n_rays_batch=4096, near=2.0, far=6.0)
            if ref_loader.available():
                ref = ref_loader.import_reference()
                mref = ref.NeRFmodel(); mref.load_state_dict(c1["sd"])
                t0 = time.perf_counter()
                ref_pred, ref_mse = ref.test_image(mref, c1["img"], poses_np[0], c1["cam"], args_c1)    # the reference's own function
                cpu_s = time.perf_counter() - t0
                ckind = "reference"
            else:
                from oracle import nerf_oracle as orc
                t0 = time.perf_counter()
                ref_pred, ref_mse = orc.test_image(c1["sd"], c1["img"], poses_np[0], c1["cam"], N_SAMPLE, 4096)
                cpu_s = time.perf_counter() - t0
                ckind = "port"
            render_c1.update({"cpu_s": cpu_s, "cpu_rays_per_s": 1e4 / cpu_s, "cpu_cores": _t.get_num_threads(), "cpu_kind": ckind,
                              "max_abs_rgb_err_vs_cpu": float((c1["pred"] - ref_pred.detach().float()).abs().max()),
                              "mse_gpu": c1["mse"], "mse_cpu": ref_mse})
    value = n_rays * world * K / ms_total * 1e3
    line = {
        "metric": "train_step_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K,
        "warmup": warm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(n_rays, world),
        "timing": dict(win_info, what=f"median of {WINDOWS} windows of {K} steps, each window barrier + synchronize on both sides, "
                                      "CUDA events, max over ranks; the windows alternate with the e2e windows (device, e2e, device, ...)"),
        "samples_per_s": value * N_SAMPLE,
        "render": {"rays_per_s": render_rays_s, "samples_per_s": render_rays_s * N_SAMPLE,
                   "workload": f"{IMG}x{IMG} frame x {N_SAMPLE} samples, rays sharded over {world} GPU(s), ray-gen + sampling + MLP + compositing",
                   "mlp_kernel": {"ms": round(mlp_infer_ms, 4), "rows": per_rank_rays * N_SAMPLE,
                                  "tflops": round(per_rank_rays * N_SAMPLE * FLOP_FWD / mlp_infer_ms / 1e9, 1),
                                  "tensor_frac": round(per_rank_rays * N_SAMPLE * FLOP_FWD / mlp_infer_ms / 1e9 / measured_peaks()["tflops"], 4),
                                  "samples_per_s": per_rank_rays * N_SAMPLE / mlp_infer_ms * 1e3}},
        "compositing": {"workload": f"{n_comp} rays x {N_SAMPLE} samples per GPU, kernels alone, synthetic sigma / rgb",
                        "forward": {"ms": round(comp_ms["forward"], 4), "gbs": round(n_comp * N_SAMPLE * BYTES_COMPOSITE_FWD / comp_ms["forward"] / 1e6, 1),
                                    "hbm_frac": round(n_comp * N_SAMPLE * BYTES_COMPOSITE_FWD / comp_ms["forward"] / 1e6 / measured_peaks()["hbm"], 4)},
                        "backward": {"ms": round(comp_ms["backward"], 4), "gbs": round(n_comp * N_SAMPLE * BYTES_COMPOSITE_BWD / comp_ms["backward"] / 1e6, 1),
                                     "hbm_frac": round(n_comp * N_SAMPLE * BYTES_COMPOSITE_BWD / comp_ms["backward"] / 1e6 / measured_peaks()["hbm"], 4)}},
        "render_c1": render_c1,
        "train_c3": {"workload": f"BASELINE config 3, live NeRFmodel(pos_encoding=False) 627,844 params: train step {n_rays} rays x {N_SAMPLE} samples, 1 GPU, "
                                 "device-resident batches, CUDA-graph replay", "ms_per_step": c3_train_ms, "rays_per_s": n_rays / c3_train_ms * 1e3,
                     "loss": c3_loss},
        "render_c3": {"workload": f"BASELINE config 3: {IMG}x{IMG} frame x {N_SAMPLE} samples (this rank's shard), ray-gen + sampling + MLP + compositing",
                      "live_nope": {"ms_per_frame": c3_render_ms, "rays_per_s": per_rank_rays / c3_render_ms * 1e3},
                      "legacy_nope_checkpoint_layout": {"ms_per_frame": c3_legacy_ms, "rays_per_s": per_rank_rays / c3_legacy_ms * 1e3,
                                                        "note": "state_dict in the layout of Phase2/checkpoints_no_pos_encoding (random values), "
                                                                "imported with legacy.load_legacy_checkpoint (ReLU density head, layout 3)"}},
        "render_c4": {"workload": f"800x800 frame x 192 samples, rays sharded over {world} GPU(s)", "ms_per_frame": c4_ms,
                      "rays_per_s": 800 * 800 / c4_ms * 1e3, "samples_per_s": 800 * 800 * 192 / c4_ms * 1e3},
        "e2e": {"value": n_rays * world * K / e2e_ms * 1e3, "unit": "rays/s", "h2d_bytes_per_step": n_rays * 9 * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / K, "timing": e2e_info},
        "gpu_launches": launches,
        "step_launch": {"mode": ("cuda graph replay of nerf_train_step (7 kernels + 1 memset per step" + (", cross-rank barrier fused into the reduce-Adam kernel)" if world > 1 else ")")) if graph_replay else
                                "per-kernel sequence from the host",
                        "graph_error": graph_error, "ms_per_step_instrumented": ms_instrumented / K,
                        "note": "`kernels` / `roofline` durations: the same steps launched kernel by kernel with CUDA events, timed right after the headline region"},
        "grad_exchange": {"mode": grad_exchange, "note": grad_exchange_note, "check": exch_check, "barrier_wait": barrier_wait,
                          "what": "nvls / p2p: gradient sum over ranks read through NVLink inside the Adam kernel, which also holds the cross-rank barrier (symmetric memory); nccl: allreduce"},
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kern,
        "cpu_baseline": cpu_baseline,
        "final_loss": final_loss, "loss_e2e": loss_host,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
