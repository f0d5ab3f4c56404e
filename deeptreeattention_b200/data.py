"""Device-side crop preprocessing (the data format either side of the hot path).

``preprocess_crops(raw)`` is the reference's ``utils.preprocess_image`` (/root/reference/src/utils.py:36-57) for a
batch of raw int16 crown crops that are already 11 x 11 and channel-first: clip the first/last 10 bands (when there
are more than 3), cast to float32 and min-max scale every pixel's spectrum to [0, 1] -- bit-identical to the
sklearn arithmetic the reference uses, but on the GPU, so the crops can cross PCIe as int16.
"""
from __future__ import annotations

import torch

from . import _capi


def preprocess_crops(raw: torch.Tensor, clip: int | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """raw: int16 CUDA tensor (B, bands_in, 11, 11) -> float32 (B, bands_in - 2*clip, 11, 11) (written into ``out`` if given)."""
    if not raw.is_cuda:
        raise RuntimeError("deeptreeattention_b200 has no CPU path: move the raw crops to a CUDA (sm_100) device")
    if raw.dtype != torch.int16:
        raise TypeError(f"raw crops must be int16, got {raw.dtype}")
    if raw.dim() != 4 or raw.shape[2] != 11 or raw.shape[3] != 11:
        raise ValueError(f"expected raw crops of shape (B, bands, 11, 11), got {tuple(raw.shape)}")
    B, C = raw.shape[0], raw.shape[1]
    if clip is None:
        clip = 10 if C > 3 else 0            # src/utils.py:40-42
    if C - 2 * clip <= 0:
        raise ValueError(f"{C} bands cannot lose {clip} at each end")
    raw = raw.contiguous()
    dev = raw.device
    handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
    shape = (B, C - 2 * clip, 11, 11)
    if out is not None and (out.dtype != torch.float32 or tuple(out.shape) != shape or not out.is_contiguous() or out.device != dev):
        raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape}")
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=dev)
        rc = _capi.lib().dta_preprocess_crops(handle, raw.data_ptr(), B, C, clip, out.data_ptr(),
                                              torch.cuda.current_stream(dev).cuda_stream)
    _capi.check(handle, rc, "dta_preprocess_crops")
    return out
