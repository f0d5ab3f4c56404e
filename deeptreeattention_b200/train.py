"""One training step of the regime the north star names -- forward, loss = sum over the heads of the class-weighted
cross-entropy, backward -- as ONE library call (``dta_train_step``, include/dta_b200.h).

``loss = fused_train_step(model, x, y, weight)`` leaves in ``p.grad`` exactly what

    out = model(x); loss = cross_entropy_heads(heads, y, weight); loss.backward()       # heads: every head of the network

leaves there after ``p.grad = None`` (bit-identical; tests/test_train_step.py) -- the caller of the reference's
``TreeModel.training_step`` (/root/reference/src/main.py:71-80) with the head losses summed -- but the loss kernel and the alpha
blend leave the critical path between the forward and the backward pass, and nothing goes through autograd.  Optimizers,
``distributed.GradSync`` and CUDA-graph capture work as with the autograd path: the gradients live in the same flat buffer.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi
from .Hang2020 import _FusedNet, _fill_tensors, _head_of_param, _stream_ptr


def fused_train_step(model: _FusedNet, x: torch.Tensor, y: torch.Tensor, weight: Optional[torch.Tensor] = None,
                     want_joint: bool = False) -> torch.Tensor:
    """Forward + summed weighted cross-entropy over every head + backward of a fused network (``Hang2020``,
    ``spectral_network``, ``spatial_network``, ``vanilla_CNN``).  Returns the loss (0-dim tensor, detached); sets ``p.grad`` of
    every parameter (``None`` for ``alpha``, which this loss does not reach, like autograd) and ``model.head_scores`` /
    ``model.head_losses``.  ``want_joint``: also compute ``Hang2020``'s blended scores (``model.joint_scores``).
    Gradients are OVERWRITTEN, not accumulated (the step is ``p.grad = None; ...; loss.backward()`` in one call): for
    gradient accumulation over micro-batches use the autograd path (``model(x)`` / ``cross_entropy_heads`` / ``backward()``)."""
    if not isinstance(model, _FusedNet):
        raise TypeError("fused_train_step needs one of the fused networks of deeptreeattention_b200.Hang2020")
    if not x.is_cuda:
        raise RuntimeError("deeptreeattention_b200 has no CPU path: move the model and the crops to a CUDA (sm_100) device")
    if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != model._bands or x.shape[2] != 11 or x.shape[3] != 11:
        raise ValueError(f"expected float32 crops of shape (B, {model._bands}, 11, 11), got {x.dtype} {tuple(x.shape)}")
    if y.dtype != torch.int64 or y.dim() != 1 or y.shape[0] != x.shape[0] or y.device != x.device:
        raise ValueError("labels must be int64 of shape (batch,) on the crops' device")
    x, y = x.contiguous(), y.contiguous()
    names, params, buffers, spec = model._call_cache()
    dev = x.device
    if params[0].device != dev:
        raise RuntimeError(f"parameter on {params[0].device} but crops on {dev}")
    if weight is not None:
        if weight.dtype != torch.float32 or weight.numel() != spec.classes or weight.device != dev:
            raise ValueError("weight must be float32 of shape (classes,) on the crops' device")
        weight = weight.contiguous()
    lib = _capi.lib()
    handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
    B, training = x.shape[0], model.training
    shape = _capi.Shape(spec.kind, B, spec.bands, spec.classes, int(training))
    sizes = _capi.query_sizes(spec.kind, B, spec.bands, spec.classes, training)
    ptr_of = {n: p.data_ptr() for n, p in zip(names, params)}
    ptr_of.update({n: b.data_ptr() for n, b in buffers.items()})
    table = spec.table_for(ptr_of)
    need = C.c_size_t()
    if lib.dta_loss_workspace_bytes(B, spec.n_heads, C.byref(need)) != 0:
        raise ValueError(f"fused_train_step: unsupported batch={B}")
    with torch.cuda.device(dev):
        scores = [torch.empty((B, spec.classes), dtype=torch.float32, device=dev) for _ in range(spec.n_heads)]
        dscores = torch.empty((spec.n_heads, B, spec.classes), dtype=torch.float32, device=dev)
        joint = torch.empty((B, spec.classes), dtype=torch.float32, device=dev) if (want_joint and spec.kind == _capi.NET_HANG2020) else None
        loss = torch.empty(spec.n_heads + 1, dtype=torch.float32, device=dev)
        saved = torch.empty(sizes.saved_bytes, dtype=torch.uint8, device=dev)
        work_f = torch.empty(max(sizes.workspace_fwd, 256), dtype=torch.uint8, device=dev)
        work_b = torch.empty(max(sizes.workspace_bwd, 256), dtype=torch.uint8, device=dev)
        work_l = torch.empty(max(need.value, 256), dtype=torch.uint8, device=dev)
        numels = [p.numel() if p.dtype == torch.float32 else 0 for p in params]
        bufs = model.__dict__.get("_grad_buffers")
        if bufs is not None and bufs[0].numel() == sum(numels) and bufs[0].device == dev:
            flat, galpha = bufs                       # distributed.GradSync's symmetric-memory buffers, reduced in place
        else:
            flat = torch.empty(sum(numels), dtype=torch.float32, device=dev)
            galpha = torch.empty((), dtype=torch.float64, device=dev)
        if _capi.POISON_GRADS:
            flat.fill_(float("nan"))
            galpha.fill_(float("nan"))
        grads, off = [], 0
        for p, n in zip(params, numels):
            if p.dtype == torch.float32:
                grads.append(flat[off:off + n].view(p.shape))
                off += n
            else:
                grads.append(galpha)
        gtable = _fill_tensors(spec.kind, {n: g.data_ptr() for n, g in zip(names, grads)})
        sp = (C.c_void_p * 6)(*[s.data_ptr() for s in scores] + [None] * (6 - spec.n_heads))
        dp = (C.c_void_p * 6)(*[dscores[i].data_ptr() for i in range(spec.n_heads)] + [None] * (6 - spec.n_heads))
        rc = lib.dta_train_step(handle, C.byref(shape), x.data_ptr(), C.byref(table), y.data_ptr(),
                                weight.data_ptr() if weight is not None else None, C.byref(sp),
                                joint.data_ptr() if joint is not None else None, loss.data_ptr(), C.byref(dp), C.byref(gtable),
                                saved.data_ptr(), work_f.data_ptr(), work_b.data_ptr(), work_l.data_ptr(), _stream_ptr(dev))
    _capi.check(handle, rc, "dta_train_step")
    spec.flat_grad, spec.alpha_grad = flat, None
    for name, p, g in zip(names, params, grads):
        if not p.requires_grad or name == "alpha":
            p.grad = None                             # the summed head losses do not reach alpha (autograd: grad None)
        else:
            p.grad = g
    model.head_scores = scores if spec.kind != _capi.NET_VANILLA else None
    model.head_losses = loss[:spec.n_heads]
    model.joint_scores = joint
    # saved / workspaces / dscores go back to the caching allocator here: the library joined its side streams into the caller's
    # stream before returning, so stream-ordered reuse is safe (same as the work buffers of the autograd path)
    return loss[spec.n_heads]


class GraphedFusedTrainStep:
    """``step(x, y) -> loss`` replaying ONE captured ``fused_train_step`` (+ ``after_backward()``, e.g. ``GradSync.sync``, and an
    optional capturable ``optimizer.step()``): the counterpart of ``graph.GraphedTrainStep`` for the one-call step."""

    def __init__(self, model, x_example: torch.Tensor, y_example: torch.Tensor, weight: Optional[torch.Tensor] = None,
                 after_backward=None, warmup: int = 3, optimizer=None):
        if not x_example.is_cuda:
            raise RuntimeError("GraphedFusedTrainStep needs CUDA tensors (no CPU path)")
        self.model = model
        self.x = torch.empty_like(x_example)
        self.y = torch.empty_like(y_example)
        self.x.copy_(x_example)
        self.y.copy_(y_example)

        def body():
            loss = fused_train_step(model, self.x, self.y, weight)
            if after_backward is not None:
                after_backward()
            if optimizer is not None:
                optimizer.step()
            return loss

        side = torch.cuda.Stream(self.x.device)
        side.wait_stream(torch.cuda.current_stream(self.x.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                body()
        torch.cuda.current_stream(self.x.device).wait_stream(side)
        torch.cuda.synchronize(self.x.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = body()

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if y is not None and y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
