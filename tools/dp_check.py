"""Data-parallel gradient exchange check (run under torchrun, one rank per GPU):
for every exchange mode (nccl allreduce / p2p peer loads / nvls multimem.ld_reduce, the latter two fused into the Adam
kernel) train a few steps from the same start and verify that (a) the replicated weights stay bit-identical across
ranks, (b) the fused modes agree with the NCCL path, and report the step time of each mode.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import sfmnerf_b200 as s
E = s.engine

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cam = s.synthetic.make_camera_info(64, 64)
poses_np = s.synthetic.sphere_poses(6, seed=0)
images = s.synthetic.render_lego_images(poses_np, cam, n_sample=64, device=dev, field=s.synthetic.dense_field).contiguous()
poses = E.poses_to_device(poses_np, dev)
res, ref = {}, None
for mode in ("nccl", "nccl_again", "p2p", "nvls"):
    os.environ["NERF_GRAD_EXCHANGE"] = mode.split("_")[0]
    torch.manual_seed(0)
    m = s.NeRFmodel().to(dev)
    try:
        tr = E.Trainer(m, images, poses, cam, 1024, 64, perturb=False, seed=3, rank=rank, world=world)
    except Exception as e:
        res[mode] = {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
        continue
    for _ in range(3):
        tr.step()
    torch.cuda.synchronize(); dist.barrier()
    flat3 = tr.st.flat.detach().clone()                        # after 3 steps: before chaos amplifies the fp32 summation order
    for _ in range(2):
        tr.step()
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(20):
        tr.step()
    ev1.record(); torch.cuda.synchronize()
    flat = tr.st.flat.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(gathered[0], g) for g in gathered)
    if mode == "nccl":
        ref = flat3
    diff = float((flat3 - ref).abs().max()) if ref is not None else None
    res[mode] = {"exchange_mode": tr.exchange.mode, "ms_per_step": ev0.elapsed_time(ev1) / 20, "replicas_bit_identical": identical,
                 "max_abs_weight_diff_vs_nccl_after_3_steps": diff, "loss": float(tr.loss.item()),
                 "cuda_graphs": len(tr._graphs), "graph_error": tr.graph_error}
    tr.close()
    assert identical, f"{mode}: replicas diverged"
    if ref is not None and mode != "nccl":
        assert diff < 5e-4, (mode, diff)          # 3 Adam steps of lr 5e-4: only the fp32 summation order differs (cf. nccl_again)
if rank == 0:
    print(json.dumps(res, indent=1))
dist.destroy_process_group()
