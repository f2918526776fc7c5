"""Host-side glue between torch tensors (device memory, streams, autograd, torch.distributed) and the
C ABI of libnerf_b200.so.  No arithmetic of the hot path happens here: every function below
validates shapes, allocates outputs and enqueues kernels on the current CUDA stream.
"""
from __future__ import annotations

import torch

from . import _lib, parallel
from ._lib import call, ptr, stream_ptr


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and t.device.type != "cuda":
            raise _lib.NerfB200Error("libnerf_b200 needs CUDA tensors (no CPU fallback); got a tensor on %s" % t.device)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


# ------------------------------------------------------------------------------------------
# Model state: flat fp32 parameters (state_dict order) + packed bf16 UMMA weight images
# ------------------------------------------------------------------------------------------
class ModelState:
    def __init__(self, layout, flat):
        lib = _lib.load()
        self.layout = int(layout)
        self.flat = flat
        self.device = flat.device
        self.n_params = int(lib.nerf_param_count(self.layout))
        assert flat.numel() == self.n_params
        self.offsets = [int(lib.nerf_param_offset(self.layout, i)) for i in range(26)]
        self.packed = torch.empty(int(lib.nerf_mlp_packed_bytes(self.layout)), dtype=torch.uint8, device=flat.device)
        self.frozen = False          # fast paths set this while they manage repacking themselves

    @classmethod
    def adopt(cls, layout, params):
        """Move the 26 parameter tensors into one flat buffer and re-point them at views of it."""
        lib = _lib.load()
        n = int(lib.nerf_param_count(layout))
        if n < 0:
            _lib.check(n, "nerf_param_count")
        dev = params[0].device
        _need_cuda(params[0])
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        offs = [int(lib.nerf_param_offset(layout, i)) for i in range(26)]
        if sum(p.numel() for p in params) != n:
            raise _lib.NerfB200Error("parameter shapes do not match layout %d" % layout)
        with torch.no_grad():
            for p, off in zip(params, offs):
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.detach().to(torch.float32))
                p.data = view
        return cls(layout, flat)

    def owns(self, params):
        base = self.flat.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in zip(params, self.offsets))

    def pack(self):
        call("nerf_mlp_pack_weights", ptr(self.flat), self.layout, ptr(self.packed), stream_ptr())

    def refresh_packed(self, params=None):
        if not self.frozen:
            self.pack()

    def split(self, flat_like):
        """Views of a flat vector shaped like the 26 parameters."""
        out = []
        for i, off in enumerate(self.offsets):
            nxt = self.offsets[i + 1] if i + 1 < 26 else self.n_params
            out.append(flat_like[off:nxt])
        return out


# ------------------------------------------------------------------------------------------
# K1-K3, K5 wrappers
# ------------------------------------------------------------------------------------------
def encode(x, L):
    """NeRFModel.py:142-156 via nerf_encode."""
    _need_cuda(x)
    lead = x.shape[:-1]
    xf = _f32c(x.reshape(-1, 3))
    out = torch.empty((xf.shape[0], 3 * (2 * L + 1)), dtype=torch.float32, device=x.device)
    call("nerf_encode", ptr(xf), xf.shape[0], int(L), ptr(out), stream_ptr())
    return out.reshape(*lead, 3 * (2 * L + 1))


def _intrinsics(camera_info):
    K = camera_info["camera_matrix"]
    return float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])


def poses_to_device(poses, device):
    """[n,4,4] (or [4,4]) camera-to-world matrices -> fp64 device tensor (the dtype K1 computes in)."""
    t = torch.as_tensor(poses, dtype=torch.float64)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    return t.contiguous().to(device)


def raygen_image(poses_dev, pose_index, camera_info, ray_begin=0, ray_count=None):
    """All (or a contiguous slice of) the rays of one image, ray id = y*W + x (train.py:735-742)."""
    _need_cuda(poses_dev)
    W, H = int(camera_info["width"]), int(camera_info["height"])
    if ray_count is None:
        ray_count = W * H - ray_begin
    fx, fy, cx, cy = _intrinsics(camera_info)
    o = torch.empty((ray_count, 3), dtype=torch.float32, device=poses_dev.device)
    d = torch.empty_like(o)
    call("nerf_raygen_image", ptr(poses_dev), int(pose_index), fx, fy, cx, cy, W, H, int(ray_begin), int(ray_count),
         ptr(o), ptr(d), stream_ptr())
    return o, d


def raygen_batch(poses_dev, camera_info, img_idx, px, py, images_dev=None):
    """Rays (and target colours) for pre-drawn (image, x, y) triples (train.py:119-127)."""
    _need_cuda(poses_dev, img_idx)
    W, H = int(camera_info["width"]), int(camera_info["height"])
    fx, fy, cx, cy = _intrinsics(camera_info)
    n = img_idx.numel()
    o = torch.empty((n, 3), dtype=torch.float32, device=poses_dev.device)
    d = torch.empty_like(o)
    tgt = torch.empty_like(o) if images_dev is not None else None
    call("nerf_raygen_batch", ptr(poses_dev), int(poses_dev.shape[0]), fx, fy, cx, cy, W, H, ptr(img_idx), ptr(px), ptr(py),
         ptr(images_dev), n, ptr(o), ptr(d), ptr(tgt), stream_ptr())
    return o, d, tgt


def draw_pixels(seed, offset, n_images, W, H, n_rays, device):
    idx = torch.empty((3, n_rays), dtype=torch.int32, device=device)
    call("nerf_draw_pixels", int(seed), int(offset), int(n_images), int(W), int(H), int(n_rays),
         ptr(idx[0]), ptr(idx[1]), ptr(idx[2]), stream_ptr())
    return idx[0], idx[1], idx[2]


def sample_z(n_rays, n_sample, near, far, device, perturb=False, t_rand=None, seed=0, offset=0):
    """train.py:205-217."""
    z = torch.empty((n_rays, n_sample), dtype=torch.float32, device=device)
    if t_rand is not None:
        t_rand = _f32c(t_rand)
        _need_cuda(t_rand)
    call("nerf_sample_z", int(n_rays), int(n_sample), float(near), float(far), 1 if perturb else 0, ptr(t_rand),
         int(seed), int(offset), ptr(z), stream_ptr())
    return z


def composite_forward(sigma_rgb, z, extras=True):
    n_rays, S = z.shape
    rgb = torch.empty((n_rays, 3), dtype=torch.float32, device=z.device)
    depth = torch.empty((n_rays,), dtype=torch.float32, device=z.device) if extras else None
    acc = torch.empty((n_rays,), dtype=torch.float32, device=z.device) if extras else None
    call("nerf_composite_forward", ptr(sigma_rgb), ptr(z), n_rays, S, ptr(rgb), ptr(depth), ptr(acc), stream_ptr())
    return rgb, depth, acc


def composite_backward(sigma_rgb, z, g_rgb):
    n_rays, S = z.shape
    d = torch.empty((n_rays, S, 4), dtype=torch.float32, device=z.device)
    call("nerf_composite_backward", ptr(sigma_rgb), ptr(z), ptr(_f32c(g_rgb)), n_rays, S, ptr(d), stream_ptr())
    return d


# ------------------------------------------------------------------------------------------
# K4 wrappers
# ------------------------------------------------------------------------------------------
def alloc_stash(st, n_rows):
    lib = _lib.load()
    return torch.empty(int(lib.nerf_mlp_stash_bytes(st.layout, int(n_rows))), dtype=torch.uint8, device=st.device)


def alloc_dstash(st, n_rows):
    lib = _lib.load()
    return torch.empty(int(lib.nerf_mlp_dstash_bytes(st.layout, int(n_rows))), dtype=torch.uint8, device=st.device)


def mlp_forward(st, pos, dirs, stash=None):
    n = pos.shape[0]
    out = torch.empty((n, 4), dtype=torch.float32, device=pos.device)
    call("nerf_mlp_forward", ptr(st.packed), st.layout, ptr(pos), ptr(dirs), n, ptr(out), ptr(stash), stream_ptr())
    return out


def mlp_forward_rays(st, rays_o, rays_d, z, stash=None):
    n_rays, S = z.shape
    out = torch.empty((n_rays, S, 4), dtype=torch.float32, device=z.device)
    call("nerf_mlp_forward_rays", ptr(st.packed), st.layout, ptr(rays_o), ptr(rays_d), ptr(z), n_rays, S, ptr(out),
         ptr(stash), stream_ptr())
    return out


def mlp_render_rays(st, rays_o, rays_d, z, extras=True):
    """Fused render of 64-sample rays: the MLP and the volume compositing (NeRFModel.py:158-202 + train.py:248-269) in ONE kernel;
    bit-identical to ``composite_forward(mlp_forward_rays(...), z)``."""
    n_rays, S = z.shape
    rgb = torch.empty((n_rays, 3), dtype=torch.float32, device=z.device)
    depth = torch.empty((n_rays,), dtype=torch.float32, device=z.device) if extras else None
    acc = torch.empty((n_rays,), dtype=torch.float32, device=z.device) if extras else None
    call("nerf_mlp_render_rays", ptr(st.packed), st.layout, ptr(rays_o), ptr(rays_d), ptr(z), n_rays, S, ptr(rgb), ptr(depth), ptr(acc),
         stream_ptr())
    return rgb, depth, acc


def mlp_backward(st, stash, sigma_rgb, d_sigma_rgb, grads, dstash=None, accumulate=False):
    n_rows = sigma_rgb.numel() // 4
    if dstash is None:
        dstash = alloc_dstash(st, n_rows)
    call("nerf_mlp_backward", ptr(st.packed), st.layout, ptr(stash), ptr(sigma_rgb), ptr(d_sigma_rgb), n_rows, ptr(dstash),
         ptr(grads), 1 if accumulate else 0, stream_ptr())
    return grads


def wants_grad(params):
    """True when a backward pass can follow: grad mode is on and some parameter requires grad.  (Inside
    autograd.Function.forward grad mode is always off and ctx.needs_input_grad ignores torch.no_grad(), so the callers decide:
    without this the no_grad() render path would allocate and write the training stash.)"""
    return torch.is_grad_enabled() and any(p.requires_grad for p in params)


def mlp_inference(st, pos, dirs):
    """NeRFmodel.forward without autograd: the inference kernel (no stash)."""
    _need_cuda(pos, dirs)
    return mlp_forward(st, pos, dirs, None)


def render_inference(st, rays_o, rays_d, z):
    """render without autograd: inference kernel + compositing -> (rgb, depth, acc)."""
    _need_cuda(rays_o, rays_d, z)
    return composite_forward(mlp_forward_rays(st, rays_o, rays_d, z, None), z, extras=True)


def render_image(st, poses_dev, pose_index, camera_info, n_sample, near=2.0, far=6.0, ray_begin=0, ray_count=None, perturb=False,
                 seed=0, offset=0, workspace=None):
    """A frame (or the slice [ray_begin, ray_begin + ray_count) of its ray ids) by one nerf_render_image call:
    ray generation, depths, inference MLP, compositing -> (rgb [n,3], depth [n], acc [n])."""
    _need_cuda(poses_dev)
    W, H = int(camera_info["width"]), int(camera_info["height"])
    if ray_count is None:
        ray_count = W * H - ray_begin
    fx, fy, cx, cy = _intrinsics(camera_info)
    dev = poses_dev.device
    need = int(_lib.load().nerf_render_workspace_bytes(int(ray_count), int(n_sample)))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    rgb = torch.empty((ray_count, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((ray_count,), dtype=torch.float32, device=dev)
    acc = torch.empty((ray_count,), dtype=torch.float32, device=dev)
    call("nerf_render_image", ptr(st.packed), st.layout, ptr(poses_dev), int(pose_index), fx, fy, cx, cy, W, H, int(ray_begin),
         int(ray_count), int(n_sample), float(near), float(far), 1 if perturb else 0, int(seed), int(offset), ptr(rgb), ptr(depth),
         ptr(acc), ptr(workspace), stream_ptr())
    return rgb, depth, acc


class MLPFunction(torch.autograd.Function):
    """NeRFmodel.forward: (pos [n,3], dir [n,3]) -> sigma_rgb [n,4]; gradients reach the parameters."""

    @staticmethod
    def forward(ctx, st, pos, dirs, *params):
        _need_cuda(pos, dirs)
        need = any(ctx.needs_input_grad[3:])
        stash = alloc_stash(st, pos.shape[0]) if need else None
        out = mlp_forward(st, pos, dirs, stash)
        ctx.st, ctx.stash = st, stash
        ctx.shapes = [p.shape for p in params]
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        st = ctx.st
        grads = torch.empty(st.n_params, dtype=torch.float32, device=out.device)
        mlp_backward(st, ctx.stash, out, _f32c(g), grads)
        ctx.stash = None
        return (None, None, None) + tuple(v.view(s) for v, s in zip(st.split(grads), ctx.shapes))


class RenderFunction(torch.autograd.Function):
    """render (train.py:187-271) for given rays and depths: -> (rgb [N,3], depth [N], acc [N])."""

    @staticmethod
    def forward(ctx, st, rays_o, rays_d, z, *params):
        _need_cuda(rays_o, rays_d, z)
        need = any(ctx.needs_input_grad[4:])
        n_rays, S = z.shape
        stash = alloc_stash(st, n_rays * S) if need else None
        sigma_rgb = mlp_forward_rays(st, rays_o, rays_d, z, stash)
        rgb, depth, acc = composite_forward(sigma_rgb, z, extras=True)
        ctx.st, ctx.stash = st, stash
        ctx.shapes = [p.shape for p in params]
        if need:
            ctx.save_for_backward(sigma_rgb, z)
        ctx.mark_non_differentiable(depth, acc)
        return rgb, depth, acc

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_acc):
        sigma_rgb, z = ctx.saved_tensors
        st = ctx.st
        d = composite_backward(sigma_rgb, z, g_rgb)
        grads = torch.empty(st.n_params, dtype=torch.float32, device=z.device)
        mlp_backward(st, ctx.stash, sigma_rgb, d, grads)
        ctx.stash = None
        return (None, None, None, None) + tuple(v.view(s) for v, s in zip(st.split(grads), ctx.shapes))


class _timed:
    """Brackets a launch with CUDA events on the current stream when a timers dict is given."""

    def __init__(self, timers, name):
        self.timers, self.name = timers, name

    def __enter__(self):
        if self.timers is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.timers is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.timers.setdefault(self.name, []).append((self.a, b))
        return False


# ------------------------------------------------------------------------------------------
# Fused train step (train.py:447-465) over the flat buffers: one call = one optimisation step.
# ------------------------------------------------------------------------------------------
class Trainer:
    """generateBatch -> render -> MSE -> backward -> [allreduce] -> Adam -> repack, all on device.

    ``images_dev`` [n,H,W,3] fp32 and ``poses_dev`` [n,4,4] fp64 stay resident in HBM; rays are drawn
    on the GPU (counter-based RNG: (seed, step, rank) fully determine the batch), so a step moves no
    bytes across PCIe unless the caller asks for the loss.  With ``world > 1`` each rank renders
    ``n_rays`` of the global batch; the gradient sum over ranks is read through NVLink inside the Adam
    kernel (parallel.GradExchange; NCCL allreduce as fallback) and scaled by 1/world -- weights stay
    bit-identical on every rank.  ``step()`` is one ``nerf_train_step`` call, replayed as a CUDA graph
    (``graph=False`` or ``NERF_TRAIN_GRAPH=0``: plain calls); ``step(timers=...)`` launches kernel by kernel.
    """

    def __init__(self, model, images_dev, poses_dev, camera_info, n_rays, n_sample, near=2.0, far=6.0, lr=5e-4,
                 betas=(0.9, 0.999), eps=1e-8, lr_decay_steps=50000, lr_decay=0.1, perturb=True, seed=0, rank=0, world=1,
                 process_group=None, graph=True):
        self.model = model
        self.st = model.kernel_state()
        self.st.frozen = True                      # this class repacks right after each Adam step
        self.images, self.poses, self.cam = images_dev, poses_dev, camera_info
        self.n_rays, self.n_sample, self.near, self.far = int(n_rays), int(n_sample), float(near), float(far)
        self.lr0, self.betas, self.eps = float(lr), betas, float(eps)
        self.lr_decay_steps, self.lr_decay = int(lr_decay_steps), float(lr_decay)
        self.perturb, self.seed, self.rank, self.world, self.pg = bool(perturb), int(seed), int(rank), int(world), process_group
        dev = self.st.device
        n = self.st.n_params
        # world > 1: symmetric (peer-mapped) gradient buffers; the exchange happens inside the Adam kernel (parallel.GradExchange)
        self.exchange = parallel.GradExchange(n, dev, self.world, process_group) if self.world > 1 else None
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev) if self.exchange is None else self.exchange.local(0)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        rows = self.n_rays * self.n_sample
        self.stash = alloc_stash(self.st, rows)
        self.dstash = alloc_dstash(self.st, rows)
        self.comp_ws, self.comp_ws_rays = None, 0     # workspace of nerf_composite_train, sized for the largest batch seen
        self.iteration = 0
        self.launches_per_step = 0
        # whole-step entry point (nerf_train_step): fixed per-step buffers + the device-side iteration counter, so that the
        # step has no step-dependent launch argument and replays as a CUDA graph (world > 1: one graph per step parity, the
        # symmetric-memory barrier captured with it; the NCCL-allreduce fallback keeps the per-kernel sequence below)
        import os
        self.use_graph = bool(graph) and os.environ.get("NERF_TRAIN_GRAPH", "1") != "0"
        self.step_state = torch.zeros(1, dtype=torch.int64, device=dev)
        self.exchange_epoch = torch.zeros(1, dtype=torch.int64, device=dev)    # exchanges done so far (data parallel; only ever grows)
        self._dev_iteration = 0
        self._fixed = None                          # (odt, z, sigma_rgb, dsr, batch9) once the first fast step allocates them
        self._cfg = {}                              # (from_batch9, step parity) -> NerfTrainStep
        self._graphs, self._graph_tried = {}, set() # (from_batch9, step parity, segment) -> torch.cuda.CUDAGraph
        self.graph_error = ""
        self._loss_host, self._loss_event = None, None
        self._b9_buf, self._copy_stream, self._copy_event = 0, None, None    # double-buffered batch upload of the whole-step path
        self._b9_read = [None, None]                # per input buffer: event behind the launches that last read it

    def lr(self):
        return self.lr0 * self.lr_decay ** (self.iteration // self.lr_decay_steps)

    def _check_batch(self, n_rays):
        """The activation / gradient stashes were sized for ``self.n_rays`` rays in __init__: a larger batch would run the MLP
        kernels past their end."""
        if n_rays > self.n_rays:
            raise _lib.NerfB200Error(f"batch of {n_rays} rays exceeds the {self.n_rays} rays this Trainer was built for")
        if n_rays <= 0:
            raise _lib.NerfB200Error("empty ray batch")

    def close(self):
        self.st.frozen = False

    def _train_step_cfg(self, from_batch9, parity, buf=0):
        import ctypes
        st, dev, n, S = self.st, self.st.device, self.n_rays, self.n_sample
        if self._fixed is None:
            ws = torch.zeros(int(_lib.load().nerf_composite_train_workspace_bytes(n)) // 4 + 1, dtype=torch.float32, device=dev)
            self._fixed = (torch.empty((3, n, 3), dtype=torch.float32, device=dev), torch.empty((n, S), dtype=torch.float32, device=dev),
                           torch.empty((n, S, 4), dtype=torch.float32, device=dev), torch.empty((n, S, 4), dtype=torch.float32, device=dev),
                           torch.empty((2, n, 9), dtype=torch.float32, device=dev), ws)
        odt, z, sigma_rgb, dsr, batch9, ws = self._fixed
        c = _lib.NerfTrainStep()
        c.struct_bytes = ctypes.sizeof(_lib.NerfTrainStep)
        c.layout, c.n_sample, c.perturb, c.world, c.rank = st.layout, S, 1 if self.perturb else 0, self.world, self.rank
        c.n_rays, c.near, c.far, c.seed, c.z_seed = n, self.near, self.far, self.seed, self.seed ^ 0x5bd1e995
        c.lr0, c.lr_decay, c.lr_decay_steps = self.lr0, self.lr_decay, self.lr_decay_steps
        c.beta1, c.beta2, c.eps, c.grad_scale = self.betas[0], self.betas[1], self.eps, parallel.mean_scale(self.world)
        if from_batch9:
            c.batch9 = batch9[buf].data_ptr()              # two input buffers: the upload of a step overlaps the previous step's backward
        else:
            c.fx, c.fy, c.cx, c.cy = _intrinsics(self.cam)
            c.n_images, c.width, c.height = int(self.images.shape[0]), int(self.cam["width"]), int(self.cam["height"])
            c.poses, c.images = self.poses.data_ptr(), self.images.data_ptr()
        c.rays_o, c.rays_d, c.target, c.z = odt[0].data_ptr(), odt[1].data_ptr(), odt[2].data_ptr(), z.data_ptr()
        c.sigma_rgb, c.d_sigma_rgb, c.loss, c.composite_ws = sigma_rgb.data_ptr(), dsr.data_ptr(), self.loss.data_ptr(), ws.data_ptr()
        c.stash, c.dstash, c.packed = self.stash.data_ptr(), self.dstash.data_ptr(), st.packed.data_ptr()
        c.params, c.exp_avg, c.exp_avg_sq = st.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        c.step_state = self.step_state.data_ptr()
        if self.exchange is None:
            c.grads = self.grads.data_ptr()
        else:
            # data parallel: wgrad writes this step's symmetric gradient buffer (double-buffered by step parity); the Adam
            # kernel reads the sum over ranks through the peers' mappings / the multicast address
            c.grads = self.exchange.local(parity).data_ptr()
            peers, mc = self.exchange.peer_pointers(parity)
            flags = self.exchange.flag_pointers()
            for r in range(self.world):
                c.peer_grads[r] = peers[r]
                c.peer_flags[r] = flags[r]
            c.multicast_grads = mc
            c.exchange_epoch = self.exchange_epoch.data_ptr()          # the cross-rank barrier runs inside the reduce-Adam kernel
        return c

    def _run_segment(self, name, fn):
        """Run ``fn`` (a fixed sequence of nerf_train_step phases [+ the cross-rank barrier]) directly the first time, then
        capture the same sequence into a CUDA graph and replay it from then on."""
        g = self._graphs.get(name)
        if g is not None:
            g.replay()
            return
        fn()
        if self.use_graph and name not in self._graph_tried:
            self._graph_tried.add(name)
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):                            # capture only records; nothing executes
                    fn()
                self._graphs[name] = g
            except Exception as e:                                   # noqa: BLE001 -- keep training with direct calls
                self.use_graph, self.graph_error = False, f"{type(e).__name__}: {e}"
                torch.cuda.synchronize()

    def _mark_b9_read(self, buf):
        if self._b9_read[buf] is None:
            self._b9_read[buf] = torch.cuda.Event()
        self._b9_read[buf].record()

    def _step_whole(self, batch9, loss_to_host=False):
        """The whole step through nerf_train_step (one C call per phase group, CUDA-graph replays after the first step).
        ``loss_to_host``: the forward phase, an async D2H copy of the loss, then the rest; the host waits for that copy only
        (the backward phases are still running when it returns), so a loop that reads every step's loss (train.py:468)
        keeps the GPU busy.  world > 1: forward + gradients, the stream-ordered cross-rank barrier, then the update phase
        whose Adam kernel reads the gradient sum over ranks through NVLink."""
        import ctypes
        key, ex = batch9 is not None, self.exchange
        parity = self.iteration & 1 if ex is not None else 0
        buf = 0
        if key:
            self._b9_buf ^= 1
            buf = self._b9_buf
        if (key, parity, buf) not in self._cfg:
            self._cfg[(key, parity, buf)] = self._train_step_cfg(key, parity, buf)
        cfg = ctypes.byref(self._cfg[(key, parity, buf)])
        if self._dev_iteration != self.iteration:            # the per-kernel path ran in between, or a checkpoint was resumed
            self.step_state.fill_(self.iteration)
            self._dev_iteration = self.iteration
        if key:
            # H2D (pinned host batch) or D2D, straight into the step's input buffer -- on a side stream, into the buffer the PREVIOUS
            # step did not use, so the copy runs while that step's backward kernels are still busy; the forward waits for it.
            # (The buffer's last reader, the batch kernel two steps back, finished before the host got that step's loss.)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.st.device)
                self._copy_event = torch.cuda.Event()
            main = torch.cuda.current_stream()
            if batch9.device.type == "cuda":
                self._copy_stream.wait_stream(main)            # a device-resident batch may still be being produced on this stream
            if self._b9_read[buf] is not None:
                self._copy_stream.wait_event(self._b9_read[buf])   # the batch kernel that last read this buffer has finished
            with torch.cuda.stream(self._copy_stream):
                self._fixed[4][buf].copy_(batch9, non_blocking=True)
                self._copy_event.record()
            main.wait_event(self._copy_event)

        def phases(mask):
            call("nerf_train_step", cfg, mask, stream_ptr())

        def rest(first):                                      # everything after (or including) the forward phase
            phases(first | 6)                                 # data parallel: the cross-rank barrier is inside the reduce-Adam kernel

        if loss_to_host:
            if self._loss_host is None:
                self._loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
                self._loss_event = torch.cuda.Event()
            self._run_segment((key, parity, buf, "forward"), lambda: phases(1))
            if key:
                self._mark_b9_read(buf)
            self._loss_host.copy_(self.loss, non_blocking=True)
            self._loss_event.record()
            self._run_segment((key, parity, buf, "backward"), lambda: rest(0))
        else:
            self._run_segment((key, parity, buf, "step"), lambda: rest(1))
            if key:
                self._mark_b9_read(buf)
        self.iteration += 1
        self._dev_iteration += 1
        self.launches_per_step = 7                  # batch, mlp fwd, composite_train, dgrad, wgrad (pairs), adam, pack (+1 memset)
        if loss_to_host:
            self._loss_event.synchronize()
            return float(self._loss_host[0])
        return self.loss

    def step(self, rays=None, timers=None, loss_to_host=False):
        """One optimisation step.  Returns the loss as a device tensor, or with ``loss_to_host`` as a Python float (read back
        as soon as the forward phase has finished, see _step_whole).  ``rays`` = optional batch: a generateBatch() array [N,9] already on the device (e.g. uploaded
        from the host like train.py:448) or (rays_o, rays_d, target) device tensors; default draws the batch on the GPU.
        ``timers`` = optional dict: CUDA-event pairs per kernel are appended to timers[name]."""
        st, S = self.st, self.n_sample
        if (self.exchange is None or self.exchange.mode != "nccl") and timers is None:
            if rays is None and self.images is not None:
                return self._step_whole(None, loss_to_host)
            if torch.is_tensor(rays) and rays.dim() == 2 and tuple(rays.shape) == (self.n_rays, 9) and rays.dtype == torch.float32:
                return self._step_whole(rays, loss_to_host)
        if torch.is_tensor(rays) and rays.device.type != "cuda":
            rays = rays.to(st.device, non_blocking=True)
        W, H = int(self.cam["width"]), int(self.cam["height"])
        dev = st.device
        launches = 0
        ctr = parallel.batch_counter(self.iteration, self.rank, self.world)
        z_seed = self.seed ^ 0x5bd1e995
        if rays is None or (torch.is_tensor(rays) and rays.dim() == 2 and rays.shape[1] == 9):
            # one launch: draw + ray generation + target gather (or the split of a generateBatch() [N,9] array) + depths
            n_rays = self.n_rays if rays is None else int(rays.shape[0])
            self._check_batch(n_rays)
            odt = torch.empty((3, n_rays, 3), dtype=torch.float32, device=dev)
            o, d, tgt = odt[0], odt[1], odt[2]
            z = torch.empty((n_rays, S), dtype=torch.float32, device=dev)
            z_off = ctr * ((n_rays * S + 3) // 4)
            if rays is None:
                fx, fy, cx, cy = _intrinsics(self.cam)
                call("nerf_train_batch", ptr(self.poses), int(self.poses.shape[0]), fx, fy, cx, cy, W, H, ptr(self.images), self.seed,
                     ctr * self.n_rays, n_rays, ptr(o), ptr(d), ptr(tgt), S, self.near, self.far, 1 if self.perturb else 0, z_seed, z_off,
                     ptr(z), stream_ptr())
            else:
                _need_cuda(rays)
                call("nerf_unpack_batch", ptr(_f32c(rays)), n_rays, ptr(o), ptr(d), ptr(tgt), S, self.near, self.far,
                     1 if self.perturb else 0, z_seed, z_off, ptr(z), stream_ptr())
            launches += 1
        else:
            o, d, tgt = (_f32c(t) for t in rays)
            _need_cuda(o, d, tgt)
            n_rays = o.shape[0]
            if d.shape != o.shape or tgt.shape != o.shape or o.dim() != 2 or o.shape[1] != 3:
                raise _lib.NerfB200Error("rays must be three [n,3] tensors (origins, directions, target colours)")
            self._check_batch(n_rays)
            z = sample_z(n_rays, S, self.near, self.far, dev, perturb=self.perturb, seed=z_seed, offset=ctr * ((n_rays * S + 3) // 4))
            launches += 1
        with _timed(timers, "mlp_fwd"):
            sigma_rgb = mlp_forward_rays(st, o, d, z, self.stash)
        # compositing forward + MSE + compositing backward in one launch; local mean over this rank's rays (the sum over
        # ranks is scaled by 1/world inside Adam)
        if n_rays > self.comp_ws_rays:
            self.comp_ws = torch.zeros(int(_lib.load().nerf_composite_train_workspace_bytes(n_rays)) // 4 + 1, dtype=torch.float32, device=dev)
            self.comp_ws_rays = n_rays
        with _timed(timers, "composite_train"):
            dsr = torch.empty_like(sigma_rgb)
            call("nerf_composite_train", ptr(sigma_rgb), ptr(z), ptr(tgt), n_rays, S, 1.0, ptr(self.loss), ptr(dsr),
                 ptr(self.comp_ws), stream_ptr())
        rows = n_rays * S
        with _timed(timers, "mlp_dgrad"):
            call("nerf_mlp_dgrad", ptr(st.packed), st.layout, ptr(self.stash), ptr(sigma_rgb), ptr(dsr), rows,
                 ptr(self.dstash), stream_ptr())
        ex = self.exchange
        if ex is not None:
            self.grads = ex.local(self.iteration)
        with _timed(timers, "mlp_wgrad"):
            call("nerf_mlp_wgrad", st.layout, ptr(self.stash), ptr(self.dstash), rows, ptr(self.grads), 0, stream_ptr())
        launches += 4                               # mlp fwd, composite_train, dgrad, wgrad (pairs) (+1 memset)
        lr = self.lr()
        if ex is not None and ex.mode != "nccl":
            # the gradient exchange is fused into the optimiser: barrier, then Adam reads the sum over ranks through NVLink
            peers, mc = ex.peer_pointers(self.iteration)
            self.iteration += 1
            # Adam reading the gradient sum over ranks through NVLink (multimem / peer loads), the cross-rank barrier fused in: the
            # time of this launch = wait for the slowest rank + the reduction + the update
            with _timed(timers, "adam_reduce"):
                call("nerf_adam_step_reduce", ptr(st.flat), peers, self.world, mc, ptr(self.exp_avg), ptr(self.exp_avg_sq), st.n_params,
                     lr, self.betas[0], self.betas[1], self.eps, self.iteration, parallel.mean_scale(self.world), ex.flag_pointers(),
                     self.rank, ptr(self.exchange_epoch), stream_ptr())
            self.exchange_epoch += 1
        else:
            if self.world > 1:
                with _timed(timers, "allreduce"):
                    parallel.allreduce_sum_(self.grads, self.pg)
            self.iteration += 1
            with _timed(timers, "adam"):
                call("nerf_adam_step", ptr(st.flat), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), st.n_params,
                     lr, self.betas[0], self.betas[1], self.eps, self.iteration, parallel.mean_scale(self.world), stream_ptr())
        with _timed(timers, "pack"):
            st.pack()
        launches += 2                               # adam, pack
        self.launches_per_step = launches
        return float(self.loss.item()) if loss_to_host else self.loss
