"""Quick per-kernel timing probe on one GPU (CUDA events, warm, inputs > L2 for the MLP kernels)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sfmnerf_b200 as s
E = s.engine

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[len(ts) // 2]

def main():
    n_rays, S = int(os.environ.get("RAYS", 4096)), 64
    torch.manual_seed(0)
    m = s.NeRFmodel().cuda()
    st = m.kernel_state()
    o = torch.zeros(n_rays, 3, device="cuda"); o[:, 2] = -4
    d = torch.nn.functional.normalize(torch.randn(n_rays, 3, device="cuda") * 0.2 + torch.tensor([0, 0, 1.0], device="cuda"), dim=-1)
    z = E.sample_z(n_rays, S, 2.0, 6.0, "cuda", perturb=True, seed=1)
    rows = n_rays * S
    stash, dstash = E.alloc_stash(st, rows), E.alloc_dstash(st, rows)
    grads = torch.zeros(st.n_params, device="cuda")
    sr = E.mlp_forward_rays(st, o, d, z, stash)
    g = torch.randn(n_rays, 3, device="cuda")
    dsr = E.composite_backward(sr, z, g)
    call, ptr, sp = s._lib.call, s._lib.ptr, s._lib.stream_ptr
    res = {}
    res["mlp_fwd_infer_ms"] = timeit(lambda: E.mlp_forward_rays(st, o, d, z, None))
    res["mlp_fwd_train_ms"] = timeit(lambda: E.mlp_forward_rays(st, o, d, z, stash))
    res["dgrad_ms"] = timeit(lambda: call("nerf_mlp_dgrad", ptr(st.packed), st.layout, ptr(stash), ptr(sr), ptr(dsr), rows, ptr(dstash), sp()))
    res["wgrad_ms"] = timeit(lambda: call("nerf_mlp_wgrad", st.layout, ptr(stash), ptr(dstash), rows, ptr(grads), 0, sp()))
    res["composite_fwd_ms"] = timeit(lambda: E.composite_forward(sr, z))
    res["composite_bwd_ms"] = timeit(lambda: E.composite_backward(sr, z, g))
    res["pack_ms"] = timeit(lambda: st.pack())
    fl = 1317888.0
    res["fwd_infer_TFLOPs"] = rows * fl / res["mlp_fwd_infer_ms"] / 1e9
    res["fwd_train_TFLOPs"] = rows * fl / res["mlp_fwd_train_ms"] / 1e9
    res["dgrad_TFLOPs"] = rows * 2 * 622592.0 / res["dgrad_ms"] / 1e9
    res["wgrad_TFLOPs"] = rows * 2 * 658944.0 / res["wgrad_ms"] / 1e9
    res["infer_samples_per_s"] = rows / res["mlp_fwd_infer_ms"] * 1e3
    cam = s.synthetic.make_camera_info(400, 400)
    poses = E.poses_to_device(s.synthetic.sphere_poses(8), "cuda")
    images = torch.rand(8, 400, 400, 3, device="cuda")
    tr = E.Trainer(m, images, poses, cam, n_rays, S, perturb=True)
    res["train_step_ms"] = timeit(lambda: tr.step(), n=20)
    res["train_rays_per_s"] = n_rays / res["train_step_ms"] * 1e3
    res["rays"] = n_rays
    print(json.dumps(res, indent=1))

main()
