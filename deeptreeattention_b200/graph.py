"""CUDA-graph capture of one training step (forward + loss + backward [+ gradient all-reduce]).

The hot path is ~60 kernel launches of a few tens of microseconds each on up to three streams (the library forks its
side streams with events, which a capture records as a DAG); replaying them as one CUDA graph
removes the host launch overhead and the gaps between kernels.  The library only *enqueues* work on the
current stream and takes every buffer from the caller (torch's caching allocator), so a step is capturable
as is: buffers allocated during capture live in the graph's private pool and keep their addresses, which
is also why ``p.grad`` tensors stay valid views of the flat gradient buffer across replays.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class GraphedTrainStep:
    """``step(x, y) -> loss`` replaying a captured forward+loss+backward of ``model``.

    ``loss_fn(model, out, y)`` must be built from CUDA ops only (e.g. ``loss.cross_entropy_heads``).
    ``after_backward()`` (optional, e.g. ``GradSync.sync``) is captured too.  Gradients are left in
    ``p.grad`` exactly as an eager ``loss.backward()`` would leave them after ``p.grad = None``.
    ``optimizer`` (optional): its ``step()`` is captured after the gradient exchange; it must keep its step counter and
    learning rate on the device (``optim.FusedAdam(..., capturable=True)``).
    """

    def __init__(self, model: torch.nn.Module, x_example: torch.Tensor, y_example: torch.Tensor,
                 loss_fn: Callable, after_backward: Optional[Callable[[], None]] = None, warmup: int = 3,
                 optimizer: Optional[torch.optim.Optimizer] = None):
        if not x_example.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (no CPU path)")
        self.model = model
        self.x = torch.empty_like(x_example)
        self.y = torch.empty_like(y_example)
        self.x.copy_(x_example)
        self.y.copy_(y_example)
        self.params = [p for p in model.parameters() if p.requires_grad]

        def body():
            for p in self.params:
                p.grad = None
            out = model(self.x)
            loss = loss_fn(model, out, self.y)
            loss.backward()
            if after_backward is not None:
                after_backward()
            if optimizer is not None:
                optimizer.step()
            return loss

        # gradients are produced on the capture / warm-up stream while the AccumulateGrad nodes were created on the
        # default stream: intended here (everything is re-recorded into the graph), so silence the advisory
        try:
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:      # older torch
            pass
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
        # BatchNorm buffers advanced during warm-up/capture like any other training step would have

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if y is not None and y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
