"""Data-parallel plumbing for the Hang2020 hot path: one process per GPU, one gradient
exchange per step.

The reference never runs multi-GPU itself; with ``gpus > 1`` Lightning would wrap the module
in DDP (/root/reference/train.py:89-98): full parameter replica per rank, per-rank BatchNorm
statistics (no SyncBN), gradients averaged over ranks.  SURVEY.md 8(e).  Here the backward
already leaves every float gradient in ONE flat buffer (see Hang2020._FusedNetFunction), so the
exchange is a single all-reduce of 2.9 MB over NCCL (NVLink 5 / NVSwitch) plus alpha's 8 bytes,
coalesced into the same NCCL group launch.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """(rank, world_size, local_rank); initialises torch.distributed from the torchrun env
    (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT) when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total: int, rank: int, world: int):
    """Contiguous, even split of ``total`` crops over ``world`` ranks (the first ``total % world``
    ranks take one extra crop)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GradSync:
    """Averages the gradients of ``module`` over the process group after ``backward()``.

    Fast path: the fused backward left all float32 gradients as views of one flat buffer ->
    one all-reduce on it (alpha's float64 gradient rides in the same coalesced launch).
    Generic path (any module, e.g. the CPU oracle in the gloo tests, or gradients that were
    accumulated/replaced): flatten per dtype, all-reduce, scatter back.
    """

    def __init__(self, module: torch.nn.Module, group=None):
        self.module, self.group = module, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.last_path = None

    def _flat_views_intact(self, spec) -> bool:
        flat = getattr(spec, "flat_grad", None) if spec is not None else None
        if flat is None:
            return False
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        for p in self.module.parameters():
            g = p.grad
            if g is None or g.dtype != torch.float32:
                continue
            if not (lo <= g.data_ptr() < hi):
                return False
        return True

    def sync(self):
        if self.world == 1:
            self.last_path = "single"
            return
        spec = self.module.fused_spec() if hasattr(self.module, "fused_spec") else None
        if self._flat_views_intact(spec):
            flat, galpha = spec.flat_grad, spec.alpha_grad
            tensors = [flat] + ([galpha] if galpha is not None else [])
            if flat.is_cuda:
                # NCCL averages in the collective itself; both tensors go out in one group launch
                with dist._coalescing_manager(group=self.group, device=flat.device, async_ops=False):
                    for t in tensors:
                        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
            else:
                for t in tensors:
                    dist.all_reduce(t, group=self.group)
                    t.div_(self.world)
            self.last_path = "flat"
            return
        by_dtype = {}
        for p in self.module.parameters():
            if p.grad is not None:
                by_dtype.setdefault(p.grad.dtype, []).append(p.grad)
        for grads in by_dtype.values():
            flat = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(flat, group=self.group)
            flat.div_(self.world)
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(f)
        self.last_path = "generic"


def broadcast_buffers(module: torch.nn.Module, src: int = 0, group=None):
    """DDP's default ``broadcast_buffers=True``: rank ``src``'s BatchNorm running statistics
    replace every rank's before a forward."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for b in module.buffers():
        dist.broadcast(b, src=src, group=group)
