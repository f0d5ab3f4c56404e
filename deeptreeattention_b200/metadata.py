"""Drop-in for the ``metadata`` / ``metadata_sensor_fusion`` modules of the reference's ``src/models/metadata.py``
(:9-44, BASELINE config 5): site-embedding MLP fused late with the Hang2020 joint scores.

Same constructor signatures, ``forward`` signatures and ``state_dict`` keys as the reference.  The torch layers are
parameter containers only: the sensor model is the CUDA-library ``Hang2020`` and everything around it -- embedding gather,
BatchNorm1d (batch statistics / running-stat update), dropout, both linear layers, the concatenation and both ReLUs, and
their backward -- is ``dta_metadata_forward`` / ``dta_metadata_backward`` (csrc/dta_metadata.cuh).  No CPU path.

Dropout: ``nn.Dropout(p=0.7)`` draws its mask from torch's generator; here the mask comes from a counter-based generator
inside the kernel, keyed by a seed drawn from torch's CPU generator per call (so ``torch.manual_seed`` makes runs
repeatable), or is supplied by the caller (``keep_mask=``, a (B, 16) bool tensor: what the parity tests do to share the
reference's mask).  Site indices outside ``[0, sites)`` are clamped on the device (``nn.Embedding`` raises)."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn
from torch.nn import Module

from . import _capi
from .Hang2020 import Hang2020, _stream_ptr

_META_KEYS = (("embedding", "embedding.weight"), ("bn_w", "batch_norm.weight"), ("bn_b", "batch_norm.bias"),
              ("bn_rm", "batch_norm.running_mean"), ("bn_rv", "batch_norm.running_var"),
              ("bn_nbt", "batch_norm.num_batches_tracked"), ("mlp_w", "mlp.weight"), ("mlp_b", "mlp.bias"))


def _table(meta_tensors, fc_w=None, fc_b=None) -> _capi.MetadataTensors:
    t = _capi.MetadataTensors()
    for field, _ in _META_KEYS:
        v = meta_tensors.get(field)
        setattr(t, field, v.data_ptr() if v is not None else None)
    t.fc_w = fc_w.data_ptr() if fc_w is not None else None
    t.fc_b = fc_b.data_ptr() if fc_b is not None else None
    return t


class _MetadataFunction(torch.autograd.Function):
    """One dta_metadata_forward / dta_metadata_backward pair.  Differentiable inputs: sensor scores (or None) and the six
    float parameters (embedding, BN weight/bias, mlp weight/bias, and fc weight/bias when fused)."""

    @staticmethod
    def forward(ctx, site, sensor, training, sites, buffers, keep_mask, seed, emb, bn_w, bn_b, mlp_w, mlp_b, fc_w, fc_b):
        dev = emb.device
        lib = _capi.lib()
        handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
        B, classes = site.shape[0], mlp_w.shape[0]
        saved_bytes, work_bytes = C.c_size_t(), C.c_size_t()
        if lib.dta_metadata_sizes(B, classes, C.byref(saved_bytes), C.byref(work_bytes)) != _capi.DTA_OK:
            raise ValueError(f"dta_metadata_sizes rejected batch={B}, classes={classes}")
        tensors = {"embedding": emb, "bn_w": bn_w, "bn_b": bn_b, "mlp_w": mlp_w, "mlp_b": mlp_b, **buffers}
        table = _table(tensors, fc_w, fc_b)
        with torch.cuda.device(dev):
            out = torch.empty((B, classes), dtype=torch.float32, device=dev)
            saved = torch.empty(saved_bytes.value, dtype=torch.uint8, device=dev)
            rc = lib.dta_metadata_forward(handle, B, sites, classes, int(training), site.data_ptr(),
                                          sensor.data_ptr() if sensor is not None else None, C.byref(table),
                                          keep_mask.data_ptr() if keep_mask is not None else None, seed, out.data_ptr(),
                                          saved.data_ptr(), _stream_ptr(dev))
        _capi.check(handle, rc, "dta_metadata_forward")
        ctx.training, ctx.sites, ctx.work_bytes, ctx.fused = training, sites, work_bytes.value, sensor is not None
        ctx.buffers = buffers
        ctx.save_for_backward(site, sensor, saved, out, emb, bn_w, bn_b, mlp_w, mlp_b, fc_w, fc_b)
        return out

    @staticmethod
    def backward(ctx, dout):
        site, sensor, saved, out, emb, bn_w, bn_b, mlp_w, mlp_b, fc_w, fc_b = ctx.saved_tensors
        dev = emb.device
        lib = _capi.lib()
        handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
        B, classes = site.shape[0], mlp_w.shape[0]
        dout = dout.contiguous().float()
        tensors = {"embedding": emb, "bn_w": bn_w, "bn_b": bn_b, "mlp_w": mlp_w, "mlp_b": mlp_b, **ctx.buffers}
        table = _table(tensors, fc_w, fc_b)
        with torch.cuda.device(dev):
            g = {k: torch.empty_like(v) for k, v in (("embedding", emb), ("bn_w", bn_w), ("bn_b", bn_b), ("mlp_w", mlp_w), ("mlp_b", mlp_b))}
            g_fc_w = torch.empty_like(fc_w) if ctx.fused else None
            g_fc_b = torch.empty_like(fc_b) if ctx.fused else None
            dsensor = torch.empty_like(sensor) if ctx.fused else None
            work = torch.empty(max(ctx.work_bytes, 256), dtype=torch.uint8, device=dev)
            gtable = _table(g, g_fc_w, g_fc_b)
            rc = lib.dta_metadata_backward(handle, B, ctx.sites, classes, int(ctx.training), site.data_ptr(),
                                           sensor.data_ptr() if ctx.fused else None, C.byref(table), saved.data_ptr(), out.data_ptr(),
                                           dout.data_ptr(), C.byref(gtable), dsensor.data_ptr() if ctx.fused else None,
                                           work.data_ptr(), _stream_ptr(dev))
        _capi.check(handle, rc, "dta_metadata_backward")
        return (None, dsensor, None, None, None, None, None,
                g["embedding"], g["bn_w"], g["bn_b"], g["mlp_w"], g["mlp_b"], g_fc_w, g_fc_b)


def _draw_seed() -> int:
    """64-bit key for the in-kernel dropout stream, from torch's CPU generator (repeatable under torch.manual_seed)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _check_inputs(site, emb):
    if not isinstance(site, torch.Tensor) or site.dim() != 1:
        raise ValueError("metadata must be a 1-D tensor of site indices")
    if not site.is_cuda or not emb.is_cuda:
        raise RuntimeError("deeptreeattention_b200 has no CPU path: move the model and the site indices to a CUDA (sm_100) device")
    if site.device != emb.device:
        raise RuntimeError(f"parameters on {emb.device} but site indices on {site.device}")
    return site.to(torch.int64).contiguous()


def _keep(keep_mask, B, dev):
    if keep_mask is None:
        return None
    if tuple(keep_mask.shape) != (B, 16):
        raise ValueError(f"keep_mask must have shape ({B}, 16)")
    return keep_mask.to(device=dev, dtype=torch.uint8).contiguous()


class metadata(Module):
    """Embedding(sites, 16) -> BatchNorm1d -> Dropout(0.7) -> Linear(16, classes) -> ReLU (reference :9-24)."""

    def __init__(self, sites, classes):
        super().__init__()
        self.embedding = nn.Embedding(sites, 16)
        self.batch_norm = nn.BatchNorm1d(16)
        self.mlp = nn.Linear(in_features=16, out_features=classes)
        self.dropout = nn.Dropout(p=0.7)

    def _bn_buffers(self):
        bn = self.batch_norm
        return {"bn_rm": bn.running_mean, "bn_rv": bn.running_var, "bn_nbt": bn.num_batches_tracked}

    def forward(self, x, keep_mask=None):
        site = _check_inputs(x, self.embedding.weight)
        return _MetadataFunction.apply(site, None, self.training, self.embedding.num_embeddings, self._bn_buffers(),
                                       _keep(keep_mask, site.shape[0], site.device), _draw_seed() if (self.training and keep_mask is None) else 0,
                                       self.embedding.weight, self.batch_norm.weight, self.batch_norm.bias, self.mlp.weight,
                                       self.mlp.bias, None, None)


class metadata_sensor_fusion(Module):
    """Joint fusion of the HSI sensor model and the site metadata (reference :26-44)."""

    def __init__(self, bands, sites, classes):
        super().__init__()
        self.metadata_model = metadata(sites, classes)
        self.sensor_model = Hang2020(bands, classes)
        self.fc1 = nn.Linear(in_features=classes * 2, out_features=classes)

    def forward(self, images, metadata, keep_mask=None):
        mm = self.metadata_model
        site = _check_inputs(metadata, mm.embedding.weight)
        sensor_softmax = self.sensor_model(images)          # the reference's name; these are the joint scores (:39)
        if site.shape[0] != sensor_softmax.shape[0]:
            raise ValueError("images and metadata must have the same batch size")
        return _MetadataFunction.apply(site, sensor_softmax, self.training, mm.embedding.num_embeddings, mm._bn_buffers(),
                                       _keep(keep_mask, site.shape[0], site.device), _draw_seed() if (self.training and keep_mask is None) else 0,
                                       mm.embedding.weight, mm.batch_norm.weight, mm.batch_norm.bias, mm.mlp.weight, mm.mlp.bias,
                                       self.fc1.weight, self.fc1.bias)
