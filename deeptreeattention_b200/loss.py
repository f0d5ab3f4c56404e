"""Fused weighted cross-entropy over classifier heads (the loss either side of the Hang2020 path).

``cross_entropy_heads([s1, ..., sk], y, weight)`` equals ``sum(F.cross_entropy(s, y, weight=weight) for s in heads)``
-- with one head it is exactly the reference's ``TreeModel.training_step`` loss
(/root/reference/src/main.py:78) -- but runs as two kernel launches that also leave the score
gradients behind, so ``backward()`` launches nothing for the loss.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _capi


class _CrossEntropyHeads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, weight, *heads):
        dev = heads[0].device
        B, classes = heads[0].shape
        n = len(heads)
        lib = _capi.lib()
        handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
        heads = [h.contiguous() for h in heads]
        need = C.c_size_t()
        if lib.dta_loss_workspace_bytes(B, n, C.byref(need)) != 0:
            raise ValueError(f"cross_entropy_heads: unsupported batch={B} / heads={n} (at most 7 heads)")
        with torch.cuda.device(dev):
            loss = torch.empty(n + 1, dtype=torch.float32, device=dev)
            flat = torch.empty((n, B, classes), dtype=torch.float32, device=dev)   # one buffer: backward scales it in one kernel
            grads = [flat[i] for i in range(n)]
            work = torch.empty(need.value, dtype=torch.uint8, device=dev)
            sp = (C.c_void_p * 8)(*[h.data_ptr() for h in heads] + [None] * (8 - n))
            gp = (C.c_void_p * 8)(*[g.data_ptr() for g in grads] + [None] * (8 - n))
            rc = lib.dta_cross_entropy_heads(handle, B, classes, n, C.byref(sp), y.data_ptr(),
                                             weight.data_ptr() if weight is not None else None, loss.data_ptr(), C.byref(gp),
                                             work.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _capi.check(handle, rc, "dta_cross_entropy_heads")
        ctx.flat = flat
        ctx.head_losses = loss[:n]
        return loss[n]

    @staticmethod
    def backward(ctx, gout):
        scaled = ctx.flat * gout
        return (None, None) + tuple(scaled[i] for i in range(scaled.shape[0]))


def cross_entropy_heads(heads: Sequence[torch.Tensor], y: torch.Tensor, weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum over ``heads`` of the class-weighted mean cross-entropy (``F.cross_entropy(s, y, weight=weight)``)."""
    heads = list(heads)
    if not heads:
        raise ValueError("cross_entropy_heads needs at least one head")
    if not heads[0].is_cuda:
        raise RuntimeError("deeptreeattention_b200 has no CPU path: scores must live on a CUDA (sm_100) device")
    if y.dtype != torch.int64 or y.dim() != 1 or y.shape[0] != heads[0].shape[0]:
        raise ValueError("labels must be int64 of shape (batch,)")
    if y.device != heads[0].device or (weight is not None and weight.device != heads[0].device):
        raise RuntimeError(f"labels / class weights must be on the scores' device ({heads[0].device})")
    for h in heads:
        if h.dtype != torch.float32 or h.shape != heads[0].shape or h.dim() != 2:
            raise ValueError("every head must be float32 of the same (batch, classes) shape")
    if weight is not None and (weight.dtype != torch.float32 or weight.numel() != heads[0].shape[1]):
        raise ValueError("weight must be float32 of shape (classes,)")
    return _CrossEntropyHeads.apply(y.contiguous(), weight.contiguous() if weight is not None else None, *heads)
