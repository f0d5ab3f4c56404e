"""Drop-in for the reference's ``src/models/year.py``: one spectral network per year, all-zero years skipped,
mean of the last-head scores (reference :9-33).  Every year network runs through the CUDA library.

Neither mode leaves the device.  ``dta_crops_nonzero`` computes the per-year "not all zero" flags (the reference's
``if x.sum() == 0: continue``, a device->host sync per year, year.py:27).
Inference (``eval()`` under ``torch.no_grad()``: validation / predict): every year network runs and ``dta_ensemble_mean``
averages the flagged years (optionally fused with the softmax of ``MultiStage.predict_step``).
Training: every year network runs with its flag registered as the library's update gate (``dta_set_update_gate``: the
BatchNorm running statistics of a flagged-out year stay untouched, as if it had not run), and the flagged years' last-head
scores are averaged with the flags as weights, so a skipped year receives exactly zero gradient (the reference leaves its
``.grad`` at ``None``).  The whole ensemble is stream-ordered and CUDA-graph capturable; the price is the arithmetic of the
skipped years.  ``host_skip=True`` restores the reference's control flow (one host sync for all years, skipped years not
computed)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
from torch import nn
from torch.nn import Module

from . import Hang2020, _capi


def _handle(dev):
    return _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())


def _check_years(images: Sequence[torch.Tensor]):
    if len(images) == 0 or len(images) > 16:
        raise ValueError("expected between 1 and 16 per-year crop tensors")
    for x in images:
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError("deeptreeattention_b200 has no CPU path: every year's crops must live on a CUDA (sm_100) device")
        if x.dtype != torch.float32 or x.shape != images[0].shape or x.device != images[0].device:
            raise ValueError("every year's crops must be float32 tensors of one shape on one device")


def crops_nonzero(images: Sequence[torch.Tensor]) -> torch.Tensor:
    """float32 (years,) device tensor: 1 where ``images[y].sum() != 0`` (the reference's skip test, year.py:27), else 0."""
    _check_years(images)
    imgs = [x.contiguous() for x in images]
    dev = imgs[0].device
    handle = _handle(dev)
    with torch.cuda.device(dev):
        flags = torch.empty(len(imgs), dtype=torch.float32, device=dev)
        work = torch.empty(len(imgs) * 4096, dtype=torch.uint8, device=dev)
        ptrs = (C.c_void_p * 16)(*[x.data_ptr() for x in imgs] + [None] * (16 - len(imgs)))
        rc = _capi.lib().dta_crops_nonzero(handle, len(imgs), C.byref(ptrs), imgs[0].numel(), flags.data_ptr(), work.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream)
    _capi.check(handle, rc, "dta_crops_nonzero")
    return flags


def ensemble_mean(scores: Sequence[torch.Tensor], flags: Optional[torch.Tensor], softmax: bool = False) -> torch.Tensor:
    """Mean over the flagged years of (batch, classes) score tensors, optionally followed by ``F.softmax(dim=1)``
    (``torch.stack(year_scores, axis=1).mean(axis=1)``, year.py:33; softmax: multi_stage.py:302,314).  No autograd."""
    scores = [s.detach().contiguous() for s in scores]
    if not scores or len(scores) > 16:
        raise ValueError("expected between 1 and 16 score tensors")
    dev = scores[0].device
    if dev.type != "cuda":
        raise RuntimeError("deeptreeattention_b200 has no CPU path: scores must live on a CUDA (sm_100) device")
    B, K = scores[0].shape
    for s in scores:
        if s.dtype != torch.float32 or s.shape != (B, K) or s.device != dev:
            raise ValueError("every year's scores must be float32 (batch, classes) on one device")
    if flags is not None and (flags.dtype != torch.float32 or flags.numel() != len(scores) or flags.device != dev):
        raise ValueError("flags must be float32 of shape (years,) on the scores' device")
    handle = _handle(dev)
    with torch.cuda.device(dev):
        out = torch.empty((B, K), dtype=torch.float32, device=dev)
        ptrs = (C.c_void_p * 16)(*[s.data_ptr() for s in scores] + [None] * (16 - len(scores)))
        rc = _capi.lib().dta_ensemble_mean(handle, len(scores), C.byref(ptrs), flags.data_ptr() if flags is not None else None, B, K,
                                           int(softmax), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _capi.check(handle, rc, "dta_ensemble_mean")
    return out


class learned_ensemble(Module):
    """``learned_ensemble(years, classes, config)`` with ``config["bands"]`` and ``config["pretrain_state_dict"]``
    exactly as in the reference (year.py:10-22)."""

    def __init__(self, years, classes, config, host_skip: bool = False):
        super().__init__()
        self.year_models = nn.ModuleList()
        self.years = years
        self.host_skip = host_skip
        for _ in range(years):
            if config.get("pretrain_state_dict"):
                base_model = Hang2020.load_from_backbone(state_dict=config["pretrain_state_dict"], classes=classes, bands=config["bands"])
            else:
                base_model = Hang2020.spectral_network(bands=config["bands"], classes=classes)
            self.year_models.append(base_model)

    def forward(self, images):
        """``images``: one (B, bands, 11, 11) tensor per year.  A year whose tensor sums to zero is skipped (reference :27)."""
        _check_years(images)
        if not self.training and not torch.is_grad_enabled():
            # inference: flags, networks and the masked mean all stay on the device (no host sync)
            flags = crops_nonzero(images)
            year_scores = [self.year_models[index](x)[-1] for index, x in enumerate(images)]
            return ensemble_mean(year_scores, flags)
        if self.host_skip:
            # the reference's control flow: a skipped year does not run at all -- one host sync for all years
            sums = torch.stack([x.sum() for x in images]).tolist()
            year_scores: List[torch.Tensor] = []
            for index, x in enumerate(images):
                if sums[index] == 0:
                    continue
                year_scores.append(self.year_models[index](x)[-1])
            return torch.stack(year_scores, axis=1).mean(axis=1)
        # training / autograd on the device: flags gate the BatchNorm buffer updates and weight the mean
        flags = crops_nonzero(images)
        dev = images[0].device
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        year_scores = []
        try:
            for y, x in enumerate(images):
                _capi.set_update_gate(index, flags[y:y + 1].data_ptr() if self.training else None)
                year_scores.append(self.year_models[y](x)[-1])
        finally:
            _capi.set_update_gate(index, None)
        stacked = torch.stack(year_scores, dim=1)                        # (B, years, classes)
        weights = (flags / flags.sum()).view(1, -1, 1)                   # no active year: NaN rows (the reference raises)
        return (stacked * weights).sum(dim=1)
