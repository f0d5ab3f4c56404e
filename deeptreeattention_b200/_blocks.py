"""Autograd bindings of the stand-alone building blocks (include/dta_b200.h: dta_plane_mean, dta_conv_module_*,
dta_attention_*, dta_classifier_*).  The reference's tests call ``conv_module``, ``spatial_attention``,
``spectral_attention`` and ``Classifier`` on their own (/root/reference/tests/test_Hang2020.py:8-33); inside the
networks they run fused (``_FusedNetFunction``).  Arithmetic is CUDA only: CPU tensors raise."""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi


def _require_cuda(x: torch.Tensor, what: str, dims: int):
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what}: expected a tensor")
    if not x.is_cuda:
        raise RuntimeError("deeptreeattention_b200 has no CPU path: move the module and its input to a CUDA (sm_100) device")
    if x.dtype != torch.float32:
        raise TypeError(f"{what}: input must be float32, got {x.dtype}")
    if x.dim() != dims:
        raise ValueError(f"{what}: expected a {dims}-d tensor, got shape {tuple(x.shape)}")


def _handle(dev):
    return _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _dense(g):
    return g.contiguous().float() if g is not None else None


class _PlaneMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, Cc, H, W = x.shape
        dev = x.device
        handle = _handle(dev)
        with torch.cuda.device(dev):
            out = torch.empty((B, Cc, 1), dtype=torch.float32, device=dev)
            rc = _capi.lib().dta_plane_mean(handle, x.data_ptr(), B * Cc, H * W, out.data_ptr(), _stream(dev))
        _capi.check(handle, rc, "dta_plane_mean")
        ctx.shape = (B, Cc, H, W)
        return out

    @staticmethod
    def backward(ctx, gout):
        B, Cc, H, W = ctx.shape
        gout = _dense(gout)
        dev = gout.device
        handle = _handle(dev)
        with torch.cuda.device(dev):
            din = torch.empty((B, Cc, H, W), dtype=torch.float32, device=dev)
            rc = _capi.lib().dta_plane_mean_backward(handle, gout.data_ptr(), B * Cc, H * W, din.data_ptr(), _stream(dev))
        _capi.check(handle, rc, "dta_plane_mean_backward")
        return din


def plane_mean(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, "global_spectral_pool", 4)
    return _PlaneMean.apply(x.contiguous())


class _ConvModule(torch.autograd.Function):
    """conv_module.forward (reference Hang2020.py:24-31): conv 3x3 same + BatchNorm2d + ReLU (+ MaxPool2d)."""

    @staticmethod
    def forward(ctx, x, pool_h, pool_w, training, rm, rv, nbt, conv_w, conv_b, bn_w, bn_b):
        B, Cin, H, W = x.shape
        filters = conv_w.shape[0]
        dev = x.device
        handle = _handle(dev)
        plane = _capi.Plane(B, Cin, H, W)
        params = _capi.ConvBlock(conv_w.data_ptr(), conv_b.data_ptr(), bn_w.data_ptr(), bn_b.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                 _ptr(nbt))
        with torch.cuda.device(dev):
            z = torch.empty((B, filters, H, W), dtype=torch.float32, device=dev)
            stat = torch.empty(2 * filters, dtype=torch.float32, device=dev)
            out = torch.empty((B, filters, H // pool_h, W // pool_w), dtype=torch.float32, device=dev)
            rc = _capi.lib().dta_conv_module_forward(handle, C.byref(plane), filters, pool_h, pool_w, int(training), x.data_ptr(),
                                                     C.byref(params), z.data_ptr(), stat.data_ptr(), out.data_ptr(), _stream(dev))
        _capi.check(handle, rc, "dta_conv_module_forward")
        ctx.geom = (B, Cin, H, W, filters, pool_h, pool_w, bool(training))
        ctx.save_for_backward(x, z, stat, conv_w, bn_w, bn_b)
        return out

    @staticmethod
    def backward(ctx, gout):
        B, Cin, H, W, filters, pool_h, pool_w, training = ctx.geom
        x, z, stat, conv_w, bn_w, bn_b = ctx.saved_tensors
        gout = _dense(gout)
        dev = x.device
        handle = _handle(dev)
        lib = _capi.lib()
        plane = _capi.Plane(B, Cin, H, W)
        need = C.c_size_t()
        lib.dta_conv_module_workspace_bytes(C.byref(plane), filters, C.byref(need))
        with torch.cuda.device(dev):
            g_w = torch.empty_like(conv_w)
            g_cb = torch.empty(filters, dtype=torch.float32, device=dev)
            g_bw = torch.empty_like(bn_w)
            g_bb = torch.empty_like(bn_b)
            dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            work = torch.empty(max(need.value, 256), dtype=torch.uint8, device=dev)
            params = _capi.ConvBlock(conv_w.data_ptr(), None, bn_w.data_ptr(), bn_b.data_ptr(), None, None, None)
            grads = _capi.ConvBlock(g_w.data_ptr(), g_cb.data_ptr(), g_bw.data_ptr(), g_bb.data_ptr(), None, None, None)
            rc = lib.dta_conv_module_backward(handle, C.byref(plane), filters, pool_h, pool_w, int(training), x.data_ptr(), C.byref(params),
                                              z.data_ptr(), stat.data_ptr(), gout.data_ptr(), C.byref(grads), _ptr(dx), work.data_ptr(),
                                              _stream(dev))
        _capi.check(handle, rc, "dta_conv_module_backward")
        return dx, None, None, None, None, None, None, g_w, g_cb, g_bw, g_bb


def conv_module_forward(module, x: torch.Tensor, pool: bool) -> torch.Tensor:
    _require_cuda(x, "conv_module", 4)
    conv, bn = module.conv_layer, module.bn1
    if x.shape[1] != conv.in_channels:
        raise ValueError(f"conv_module: expected {conv.in_channels} input channels, got {x.shape[1]}")
    if conv.weight.device != x.device:
        raise RuntimeError(f"parameter on {conv.weight.device} but input on {x.device}")
    ph, pw = (1, 1)
    if pool:
        k = module.max_pool.kernel_size            # AttributeError without maxpool_kernel, like the reference
        ph, pw = (k, k) if isinstance(k, int) else tuple(k)
    return _ConvModule.apply(x.contiguous(), int(ph), int(pw), module.training, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                             conv.weight, conv.bias, bn.weight, bn.bias)


class _Attention(torch.autograd.Function):
    """spectral_attention.forward / spatial_attention.forward (reference Hang2020.py:146-168 / 103-124)."""

    @staticmethod
    def forward(ctx, x, kind, *params):
        B, Cc, H, W = x.shape
        dev = x.device
        handle = _handle(dev)
        lib = _capi.lib()
        plane = _capi.Plane(B, Cc, H, W)
        nfeat, nsaved, nwork = C.c_size_t(), C.c_size_t(), C.c_size_t()
        rc = lib.dta_attention_sizes(kind, C.byref(plane), C.byref(nfeat), C.byref(nsaved), C.byref(nwork))
        if rc != 0:
            raise ValueError("Unknown incoming kernel size {} for attention layers".format(Cc))
        table = _attention_table(kind, params)
        with torch.cuda.device(dev):
            out = torch.empty_like(x)
            feat = torch.empty((B, nfeat.value), dtype=torch.float32, device=dev)
            saved = torch.empty(B * nsaved.value, dtype=torch.float32, device=dev)
            rc = lib.dta_attention_forward(handle, kind, C.byref(plane), x.data_ptr(), C.byref(table), out.data_ptr(), feat.data_ptr(),
                                           saved.data_ptr(), _stream(dev))
        _capi.check(handle, rc, "dta_attention_forward")
        ctx.kind, ctx.geom, ctx.nwork = kind, (B, Cc, H, W), nwork.value
        ctx.save_for_backward(x, saved, *params)
        return out, feat

    @staticmethod
    def backward(ctx, gout, gfeat):
        x, saved, *params = ctx.saved_tensors
        B, Cc, H, W = ctx.geom
        gout, gfeat = _dense(gout), _dense(gfeat)
        dev = x.device
        handle = _handle(dev)
        plane = _capi.Plane(B, Cc, H, W)
        table = _attention_table(ctx.kind, params)
        with torch.cuda.device(dev):
            grads = [torch.empty_like(p) for p in params]
            gtable = _attention_table(ctx.kind, grads)
            dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            work = torch.empty(max(ctx.nwork, 256), dtype=torch.uint8, device=dev)
            rc = _capi.lib().dta_attention_backward(handle, ctx.kind, C.byref(plane), x.data_ptr(), C.byref(table), saved.data_ptr(),
                                                    _ptr(gout), _ptr(gfeat), _ptr(dx), C.byref(gtable), work.data_ptr(), _stream(dev))
        _capi.check(handle, rc, "dta_attention_backward")
        return (dx, None) + tuple(grads)


def _attention_table(kind, tensors):
    """dta_attention table from (w0, b0, w1, b1) [spectral] or (pool_w, pool_b, w0, b0, w1, b1) [spatial]."""
    p = [t.data_ptr() for t in tensors]
    if kind == _capi.ATTN_SPECTRAL:
        return _capi.Attention(None, None, p[0], p[1], p[2], p[3])
    return _capi.Attention(p[0], p[1], p[2], p[3], p[4], p[5])


def attention_forward(module, kind: int, x: torch.Tensor):
    _require_cuda(x, "attention", 4)
    c1, c2 = module.attention_conv1, module.attention_conv2
    if c1.weight.device != x.device:
        raise RuntimeError(f"parameter on {c1.weight.device} but input on {x.device}")
    if kind == _capi.ATTN_SPECTRAL:
        params = (c1.weight, c1.bias, c2.weight, c2.bias)
        filters = c1.in_channels
    else:
        params = (module.channel_pool.weight, module.channel_pool.bias, c1.weight, c1.bias, c2.weight, c2.bias)
        filters = module.channel_pool.in_channels
    if x.shape[1] != filters:
        raise ValueError(f"attention: expected {filters} channels, got {x.shape[1]}")
    return _Attention.apply(x.contiguous(), kind, *params)


class _Classifier(torch.autograd.Function):
    """Classifier.forward (reference Hang2020.py:63-66)."""

    @staticmethod
    def forward(ctx, feat, w, b):
        B, F = feat.shape
        K = w.shape[0]
        dev = feat.device
        handle = _handle(dev)
        with torch.cuda.device(dev):
            out = torch.empty((B, K), dtype=torch.float32, device=dev)
            rc = _capi.lib().dta_classifier_forward(handle, B, F, K, feat.data_ptr(), w.data_ptr(), _ptr(b), out.data_ptr(), _stream(dev))
        _capi.check(handle, rc, "dta_classifier_forward")
        ctx.has_bias = b is not None
        ctx.save_for_backward(feat, w)
        return out

    @staticmethod
    def backward(ctx, gout):
        feat, w = ctx.saved_tensors
        B, F = feat.shape
        K = w.shape[0]
        gout = _dense(gout)
        dev = feat.device
        handle = _handle(dev)
        with torch.cuda.device(dev):
            dfeat = torch.empty_like(feat) if ctx.needs_input_grad[0] else None
            dw = torch.empty_like(w)
            db = torch.empty(K, dtype=torch.float32, device=dev) if ctx.has_bias else None
            rc = _capi.lib().dta_classifier_backward(handle, B, F, K, feat.data_ptr(), w.data_ptr(), gout.data_ptr(), _ptr(dfeat),
                                                     dw.data_ptr(), _ptr(db), _stream(dev))
        _capi.check(handle, rc, "dta_classifier_backward")
        return dfeat, dw, db


def classifier_forward(module, feat: torch.Tensor) -> torch.Tensor:
    _require_cuda(feat, "Classifier", 2)
    fc = module.fc1
    if feat.shape[1] != fc.in_features:
        raise ValueError(f"Classifier: expected {fc.in_features} features, got {feat.shape[1]}")
    if fc.weight.device != feat.device:
        raise RuntimeError(f"parameter on {fc.weight.device} but input on {feat.device}")
    return _Classifier.apply(feat.contiguous(), fc.weight, fc.bias)
