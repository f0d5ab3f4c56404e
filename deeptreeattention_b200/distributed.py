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
    one all-reduce on it (plus 8 bytes for alpha's float64 gradient).
    Generic path (any module, e.g. the CPU oracle in the gloo tests, or gradients that were
    accumulated/replaced): flatten per dtype, all-reduce, scatter back.
    """

    def __init__(self, module: torch.nn.Module, group=None, peer: bool = True):
        self.module, self.group = module, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.last_path = None
        self.peer = None
        self.peer_error = None
        if peer and self.world > 1 and hasattr(module, "fused_spec") and next(module.parameters()).is_cuda:
            try:
                self.peer = _PeerBuffers(module, group)
            except Exception as e:      # no symmetric memory on this system / torch build: NCCL path
                self.peer_error = f"{type(e).__name__}: {e}"

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

    def _refresh_alpha(self, averaged):
        """alpha's float64 gradient is reduced in the buffer the backward wrote it to; autograd's AccumulateGrad keeps its OWN
        copy in ``p.grad`` whenever that buffer is referenced elsewhere (it always is: the spec / the symmetric buffer), so the
        mean has to be copied over it."""
        for p in self.module.parameters():
            if p.dtype == torch.float64 and p.grad is not None and p.grad.data_ptr() != averaged.data_ptr():
                p.grad.copy_(averaged)

    def sync(self):
        if self.world == 1:
            self.last_path = "single"
            return
        spec = self.module.fused_spec() if hasattr(self.module, "fused_spec") else None
        if self.peer is not None and spec is not None and spec.flat_grad is self.peer.flat and self._flat_views_intact(spec):
            if self.peer.in_backward and self.peer.exchanged():
                # dta_backward already exchanged (two launches overlapped with conv1's weight gradient, dta_set_grad_exchange)
                self.last_path = "peer-multimem-in-backward"
                return
            # ONE kernel over NVLink peer memory: sum over ranks (in the switch when NVLS is available), mean, in place
            self.peer.allreduce()
            self._refresh_alpha(self.peer.alpha)
            self.last_path = "peer-multimem" if self.peer.multicast_ptr else "peer-p2p"
            return
        if self._flat_views_intact(spec):
            flat, galpha = spec.flat_grad, spec.alpha_grad
            tensors = [flat] + ([galpha] if galpha is not None else [])
            if flat.is_cuda:
                # NCCL averages in the collective itself (alpha's float64 gradient cannot be coalesced with float32)
                for t in tensors:
                    dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
            else:
                for t in tensors:
                    dist.all_reduce(t, group=self.group)
                    t.div_(self.world)
            if galpha is not None:
                self._refresh_alpha(galpha)
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


class _PeerBuffers:
    """Symmetric-memory gradient buffers of one module + the fused all-reduce call (dta_grad_allreduce)."""

    def __init__(self, module: torch.nn.Module, group=None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _capi
        params = list(module.parameters())
        dev = params[0].device
        nfloat = sum(p.numel() for p in params if p.dtype == torch.float32)
        ndouble = sum(p.numel() for p in params if p.dtype == torch.float64)
        if any(p.dtype not in (torch.float32, torch.float64) for p in params) or ndouble > 1:
            raise ValueError("peer gradient buffers support float32 parameters plus the float64 alpha")
        self.n4 = (nfloat + 3) // 4
        self.nd = ndouble
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        lib = _capi.lib()
        buf_bytes, flags_off, scratch_bytes = C.c_size_t(), C.c_size_t(), C.c_size_t()
        if lib.dta_grad_allreduce_sizes(self.n4, self.nd, self.world, C.byref(buf_bytes), C.byref(flags_off), C.byref(scratch_bytes)) != 0:
            raise ValueError(f"world size {self.world} not supported by the peer all-reduce")
        with torch.cuda.device(dev):
            self.buf = symm_mem.empty(buf_bytes.value, dtype=torch.uint8, device=dev)
            self.buf.zero_()
            self.handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            self.scratch = torch.empty(scratch_bytes.value, dtype=torch.uint8, device=dev)
            self.sync_words = torch.zeros(8, dtype=torch.int32, device=dev)
        self.peer_ptrs = (C.c_void_p * 16)(*[int(p) for p in self.handle.buffer_ptrs] + [None] * (16 - self.world))
        try:
            self.multicast_ptr = int(self.handle.multicast_ptr) if self.handle.has_multicast_support(dev.type, dev.index) else 0
        except Exception:
            self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.flat = self.buf[:nfloat * 4].view(torch.float32)
        self.alpha = self.buf[self.n4 * 16:self.n4 * 16 + 8].view(torch.float64).reshape(()) if ndouble else torch.zeros((), dtype=torch.float64, device=dev)
        self.device, self._capi, self._C = dev, _capi, C
        module.__dict__["_grad_buffers"] = (self.flat, self.alpha)
        # Optional (DTA_EXCHANGE_IN_BACKWARD=1): the backward pass exchanges its gradients itself in two launches, everything but
        # conv1's weights under conv1's weight-gradient kernel.  Measured SLOWER than the single exchange kernel after the
        # backward pass (N = 2: 1.057 vs 1.033 ms per step, N = 8: 1.011 vs 0.972 ms; profiles/r06y_*): the spinning exchange CTAs
        # take issue slots from the tensor kernel they hide under.  Off by default.
        self.in_backward = False
        if self.multicast_ptr and os.environ.get("DTA_EXCHANGE_IN_BACKWARD", "0") == "1":
            handle = _capi.context(dev.index)
            rc = lib.dta_set_grad_exchange(handle, self.rank, self.world, C.byref(self.peer_ptrs), self.multicast_ptr, self.n4, self.nd,
                                           self.sync_words.data_ptr())
            _capi.check(handle, rc, "dta_set_grad_exchange")
            self.in_backward = True
        torch.cuda.synchronize(dev)
        dist.barrier(group)          # every rank's flag words are zero before anyone signals

    def exchanged(self) -> bool:
        return bool(self._capi.get_option(self.device.index, "exchanged"))

    def allreduce(self):
        dev = self.device
        handle = self._capi.context(dev.index)
        rc = self._capi.lib().dta_grad_allreduce(handle, self.rank, self.world, self._C.byref(self.peer_ptrs), self.multicast_ptr or None,
                                                 self.n4, self.nd, self.scratch.data_ptr(), self.sync_words.data_ptr(),
                                                 torch.cuda.current_stream(dev).cuda_stream)
        self._capi.check(handle, rc, "dta_grad_allreduce")
