"""Multi-GPU host logic (one process per GPU, torch.distributed): how rays shard and how gradients
combine.  Rays are independent through K1-K5, so
  * inference: rank r renders the contiguous ray-id slice ``shard_range(H*W, r, world)`` of the frame
    into its own slice of the output -- no collective on the data path;
  * training: every rank draws ``n_rays`` rays of the global batch (distinct counter ranges of the
    device RNG), back-propagates its LOCAL mean-MSE, and the flat fp32 gradient (661,636 floats,
    2.6 MB) is sum-allreduced over NCCL/NVLink; Adam applies it scaled by 1/world, which equals the
    gradient of the global-batch mean (train.py:460) because the shards are equal-sized.
Pure host functions: usable (and tested) with the gloo backend on CPU.

On GPUs the exchange is fused into the optimiser kernel (``GradExchange``): wgrad writes the local gradient into a
symmetric (peer-mapped) buffer, one stream-ordered cross-rank barrier follows, and ``nerf_adam_step_reduce`` reads the
gradient as the sum over ranks while applying Adam -- through the NVLink multicast address (NVLS, ``multimem.ld_reduce``)
or by loading every peer's buffer in rank order.  NCCL's allreduce remains the fallback where symmetric memory is not
available (and the only mode on CPU / gloo).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [begin, end) slice of ``n_items`` ray ids owned by ``rank``; sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def batch_counter(iteration: int, rank: int, world: int) -> int:
    """Index of the (iteration, rank) ray batch in the global counter-based RNG stream: batches of
    different ranks / iterations never overlap."""
    return int(iteration) * int(world) + int(rank)


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of the flat gradient vector (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def mean_scale(world: int) -> float:
    """Factor applied to the summed local-mean gradients (inside the fused Adam kernel)."""
    return 1.0 / float(world)


def gather_frame(local_rgb: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """Assemble the full frame from the per-rank slices (outside any timed region)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rgb
    world = dist.get_world_size(group)
    sizes = [e - b for b, e in (shard_range(n_items, r, world) for r in range(world))]
    pad = torch.zeros((max(sizes),) + tuple(local_rgb.shape[1:]), dtype=local_rgb.dtype, device=local_rgb.device)
    pad[: local_rgb.shape[0]] = local_rgb
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


class GradExchange:
    """Double-buffered symmetric gradient buffers of one data-parallel group.

    ``mode``: ``"nvls"`` (multicast address available), ``"p2p"`` (peer pointers only) or ``"nccl"`` (symmetric memory
    unavailable: plain local buffers + ``allreduce_sum_``).  ``NERF_GRAD_EXCHANGE=nccl|p2p|nvls`` forces a mode.
    Buffer ``step % 2`` is used in step ``step``: a rank can be at most one step ahead of another (each step has one
    barrier), so the buffer a slow rank is still reading is never the one a fast rank is already overwriting.
    """

    def __init__(self, n_params: int, device, world: int, group=None):
        import os
        self.n, self.world, self.group, self.device = int(n_params), int(world), group, torch.device(device)
        self.n_pad = (self.n + 63) // 64 * 64                                  # 256-byte aligned second buffer
        self.n_flags = 64                                                      # barrier words behind the two gradient buffers (8 used)
        self.mode, self.why, self.handle = "nccl", "", None
        want = os.environ.get("NERF_GRAD_EXCHANGE", "")
        if self.world > 1 and self.device.type == "cuda" and want != "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                buf = symm_mem.empty(2 * self.n_pad + self.n_flags, dtype=torch.float32, device=self.device)
                buf.zero_()
                self.handle = symm_mem.rendezvous(buf, grp)
                self.buf = buf
                self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
                mc = int(self.handle.multicast_ptr or 0)                      # 0: no NVLS multicast mapping for this buffer
                self.mc_ptr = mc
                self.mode = "nvls" if (mc and want != "p2p") else "p2p"
                if want == "nvls" and not mc:
                    raise RuntimeError("NERF_GRAD_EXCHANGE=nvls but the symmetric buffer has no multicast address")
            except Exception as e:                                             # noqa: BLE001 -- any failure means: use NCCL
                if want in ("p2p", "nvls"):
                    raise
                self.mode, self.why, self.handle = "nccl", f"{type(e).__name__}: {e}", None
        # The mode must be the SAME on every rank (a rank alone in dist.all_reduce while the others wait in the symmetric-memory
        # barrier would hang, or read gradients that were never exchanged): agree on the weakest mode any rank reached.
        if self.world > 1 and self.device.type == "cuda" and want != "nccl" and dist.is_available() and dist.is_initialized():
            code = {"nccl": 0, "p2p": 1, "nvls": 2}[self.mode]
            t = torch.tensor([code], dtype=torch.int32, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            agreed = ("nccl", "p2p", "nvls")[int(t.item())]
            if agreed != self.mode:
                self.why = (self.why + "; " if self.why else "") + f"another rank only reached '{agreed}' (this rank: '{self.mode}')"
                self.mode = agreed
                if agreed == "nccl":
                    self.handle = None
        if self.handle is None:
            self.buf = torch.zeros(2 * self.n_pad + self.n_flags, dtype=torch.float32, device=self.device)

    def local(self, step: int) -> torch.Tensor:
        """This rank's gradient buffer of ``step`` (wgrad writes it)."""
        o = (step & 1) * self.n_pad
        return self.buf[o:o + self.n]

    def barrier(self):
        """Stream-ordered cross-rank barrier (torch's symmetric-memory barrier kernel).  The train step does not use it any more:
        the barrier is fused into the reduce-Adam kernel through ``flag_pointers`` (one launch less, no 16 us barrier kernel)."""
        if self.handle is not None:
            self.handle.barrier(channel=0)

    def flag_pointers(self):
        """ctypes array of every rank's barrier-word array (8 unsigned words behind its gradient buffers), in rank order."""
        import ctypes as C
        off = 2 * self.n_pad * 4
        return (C.c_void_p * self.world)(*[p + off for p in self.peer_ptrs])

    def peer_pointers(self, step: int):
        """(ctypes array of the ranks' buffer addresses in rank order, multicast address or None) for ``step``."""
        import ctypes as C
        off = (step & 1) * self.n_pad * 4
        arr = (C.c_void_p * self.world)(*[p + off for p in self.peer_ptrs])
        mc = C.c_void_p(self.mc_ptr + off) if self.mode == "nvls" else None
        return arr, mc
