"""The callers of the year ensemble in the reference's ``src/models/multi_stage.py``: ``base_model`` (:17-33) and the
arithmetic of ``MultiStage.predict_step`` (:306-318) -- every level's model on the same per-year crops, softmax per level.

``predict_step`` here is that loop re-planned for the GPU: per year, the level networks that share the crop tensor are
evaluated two at a time by ``dta_forward_pair`` (the crops are read once per pair and block 1's convolution runs with both
networks' filters side by side), the zero-year flags are computed once on the device for all levels, and each level's
masked year mean is fused with its softmax (``dta_ensemble_mean``).  The Lightning plumbing around it (datasets, loaders,
logging, ``torchmetrics``) is out of scope."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch
from torch.nn import Module

from . import Hang2020, _capi
from .year import crops_nonzero, ensemble_mean, learned_ensemble, _check_years


class base_model(Module):
    """``base_model(years, classes, config)``: wraps a ``learned_ensemble`` as ``self.model`` (reference :17-33).  The reference
    also builds a ``torchmetrics.MetricCollection`` here (:23-28); metrics are logging plumbing and not part of this package."""

    def __init__(self, years, classes, config):
        super().__init__()
        self.model = learned_ensemble(classes=classes, years=years, config=config)
        self.metrics = None

    def forward(self, x):
        return self.model(x)


def _ensemble_of(m) -> learned_ensemble:
    e = m.model if isinstance(m, base_model) else m
    if not isinstance(e, learned_ensemble):
        raise TypeError("predict_step expects base_model / learned_ensemble level models")
    return e


def paired_last_heads(net_a, net_b, x: torch.Tensor):
    """Last-head scores of two ``spectral_network``s (eval mode) on the same crops through ONE ``dta_forward_pair`` call."""
    for n in (net_a, net_b):
        if not isinstance(n, Hang2020.spectral_network):
            raise TypeError("paired_last_heads pairs spectral_network modules (the year models of the reference)")
        if n.training:
            raise RuntimeError("paired_last_heads is inference only: call .eval() on the level models")
    if net_a._bands != net_b._bands or x.shape[1] != net_a._bands:
        raise ValueError("both networks and the crops must have the same number of bands")
    dev = x.device
    pa = next(net_a.parameters())
    if pa.device != dev or next(net_b.parameters()).device != dev:
        raise RuntimeError("networks and crops must live on the same CUDA device")
    x = x.contiguous()
    B = x.shape[0]
    ca, cb = net_a._classes, net_b._classes
    # the pair reuses the two-branch table of Hang2020: branch 0 <- net_a, branch 1 <- net_b
    ptr_of = {}
    for prefix, net in (("spectral_network.", net_a), ("spatial_network.", net_b)):
        for k, v in net.state_dict(keep_vars=True).items():
            ptr_of[prefix + k] = v.data_ptr()
    table = Hang2020._fill_tensors(_capi.NET_HANG2020, ptr_of)
    sizes = _capi.query_sizes(_capi.NET_SPECTRAL_PAIR, B, net_a._bands, max(ca, cb), False)
    shape = _capi.Shape(_capi.NET_SPECTRAL_PAIR, B, net_a._bands, ca, 0)
    handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
    with torch.cuda.device(dev):
        scores = [torch.empty((B, ca if i < 3 else cb), dtype=torch.float32, device=dev) for i in range(6)]
        saved = torch.empty(sizes.saved_bytes, dtype=torch.uint8, device=dev)
        work = torch.empty(max(sizes.workspace_fwd, 256), dtype=torch.uint8, device=dev)
        sp = (C.c_void_p * 6)(*[s.data_ptr() for s in scores])
        rc = _capi.lib().dta_forward_pair(handle, C.byref(shape), cb, x.data_ptr(), C.byref(table), C.byref(sp), saved.data_ptr(),
                                          work.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _capi.check(handle, rc, "dta_forward_pair")
    return scores[2], scores[5]


@torch.no_grad()
def predict_step(models: Sequence[Module], images: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """``[F.softmax(model(images), dim=1) for model in models]`` (reference ``MultiStage.predict_step``, :306-318) with every
    level model in eval mode: one pass over each year's crops per PAIR of levels, no host synchronisation."""
    ensembles = [_ensemble_of(m) for m in models]
    if not ensembles:
        return []
    _check_years(images)
    for e in ensembles:
        if e.training:
            raise RuntimeError("predict_step is inference only: call .eval() on the level models")
        if len(e.year_models) != len(images):
            raise ValueError("every level model needs one year network per crop tensor")
    flags = crops_nonzero(images)
    per_level: List[List[torch.Tensor]] = [[] for _ in ensembles]
    for y, x in enumerate(images):
        nets = [e.year_models[y] for e in ensembles]
        i = 0
        while i + 1 < len(nets):
            sa, sb = paired_last_heads(nets[i], nets[i + 1], x)
            per_level[i].append(sa)
            per_level[i + 1].append(sb)
            i += 2
        if i < len(nets):
            per_level[i].append(nets[i](x)[-1])
    return [ensemble_mean(scores, flags, softmax=True) for scores in per_level]
