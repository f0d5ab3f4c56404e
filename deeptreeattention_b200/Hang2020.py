"""Drop-in namespace for the reference's ``src/models/Hang2020.py`` backed by libdta_b200.so.

Same module-level names, constructor signatures, ``forward`` signatures, ``state_dict()``
keys/shapes/dtypes and train()/eval() semantics as the reference
(/root/reference/src/models/Hang2020.py:7-278), so an instance can be handed
to ``TreeModel(model=...)`` (src/main.py:33,50,77) unchanged.  The torch.nn layers below are
only *parameter containers* (they give the reference's default initialisation and key names);
the arithmetic of a network forward/backward is one call each into the CUDA library.

No CPU path exists here: a CPU tensor, a missing library or a non-sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch
from torch import nn
from torch.nn import Module

from . import _blocks, _capi

_KIND_PREFIXES = {
    _capi.NET_HANG2020: ("spectral_network.", "spatial_network."),
    _capi.NET_SPECTRAL: ("",),
    _capi.NET_SPATIAL: ("",),
    _capi.NET_VANILLA: ("",),
}
_CONV_FIELDS = {
    "conv_layer.weight": "conv_w", "conv_layer.bias": "conv_b", "bn1.weight": "bn_w", "bn1.bias": "bn_b",
    "bn1.running_mean": "bn_rm", "bn1.running_var": "bn_rv", "bn1.num_batches_tracked": "bn_nbt",
}
_ATTN_FIELDS = {
    "channel_pool.weight": "pool_w", "channel_pool.bias": "pool_b",
    "attention_conv1.weight": "w0", "attention_conv1.bias": "b0",
    "attention_conv2.weight": "w1", "attention_conv2.bias": "b1",
}


# Every (re)registration of a parameter, buffer or sub-module anywhere in the process bumps this counter (torch's global
# registration hooks fire from Module.__setattr__ / register_*): the fused networks cache their name / tensor lists and only
# trust them while the counter stands still.  Swapping a head (net.classifier3 = Classifier(...)), a sub-network or a single
# Parameter is therefore picked up like the reference nn.Module does, at the cost of one integer comparison per call.
_STRUCTURE_VERSION = [0]


def _bump_structure_version(*_args):
    _STRUCTURE_VERSION[0] += 1
    return None


nn.modules.module.register_module_parameter_registration_hook(_bump_structure_version)
nn.modules.module.register_module_buffer_registration_hook(_bump_structure_version)
nn.modules.module.register_module_module_registration_hook(_bump_structure_version)


def _fill_tensors(kind: int, ptr_of: Dict[str, int]) -> _capi.Tensors:
    """Build the C parameter (or gradient) table from {state_dict key: device pointer}."""
    t = _capi.Tensors()
    t.alpha = ptr_of.get("alpha", None)
    for g, prefix in enumerate(_KIND_PREFIXES[kind]):
        br = t.branch[g]
        for k in (1, 2, 3):
            for key, field in _CONV_FIELDS.items():
                setattr(br.conv[k - 1], field, ptr_of.get(f"{prefix}conv{k}.{key}", None))
            for key, field in _ATTN_FIELDS.items():
                setattr(br.attn[k - 1], field, ptr_of.get(f"{prefix}attention_{k}.{key}", None))
            br.fc_w[k - 1] = ptr_of.get(f"{prefix}classifier{k}.fc1.weight", None)
            br.fc_b[k - 1] = ptr_of.get(f"{prefix}classifier{k}.fc1.bias", None)
        if kind == _capi.NET_VANILLA:
            br.fc_w[2] = ptr_of.get("fc1.weight", None)
            br.fc_b[2] = ptr_of.get("fc1.bias", None)
    return t


def _head_of_param(kind: int, name: str):
    """Index of the head whose upstream gradient alone reaches this parameter (classifier
    weights), else None."""
    if kind == _capi.NET_VANILLA:
        return None
    for g, prefix in enumerate(_KIND_PREFIXES[kind]):
        for k in (1, 2, 3):
            if name.startswith(f"{prefix}classifier{k}."):
                return g * 3 + (k - 1)
    return None


class _NetSpec:
    """Static description handed to the autograd function."""

    def __init__(self, kind, bands, classes):
        self.kind, self.bands, self.classes = kind, bands, classes
        self.n_heads = {_capi.NET_HANG2020: 6, _capi.NET_SPECTRAL: 3, _capi.NET_SPATIAL: 3, _capi.NET_VANILLA: 1}[kind]
        # set by the last backward: the flat fp32 gradient buffer every float parameter's
        # .grad is a view of, and alpha's fp64 gradient (distributed.GradSync reduces these)
        self.flat_grad = None
        self.alpha_grad = None
        # optional persistent (flat float32, 0-dim float64) gradient buffers supplied by distributed.GradSync: symmetric
        # memory that the peer all-reduce kernel works on in place
        self.grad_buffers = None
        self.last_saved = None          # diagnostics only (_capi.KEEP_SAVED)
        self._table_key = None
        self._table = None

    def table_for(self, ptr_of: Dict[str, int]):
        """C parameter table for these device pointers (rebuilt only when a pointer moved)."""
        key = tuple(ptr_of.values())
        if key != self._table_key:
            self._table, self._table_key = _fill_tensors(self.kind, ptr_of), key
        return self._table


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _FusedNetFunction(torch.autograd.Function):
    """One dta_forward / dta_backward pair (include/dta_b200.h)."""

    @staticmethod
    def forward(ctx, x, spec: _NetSpec, training: bool, names: List[str], buffers: Dict[str, torch.Tensor], *params):
        dev = x.device
        lib = _capi.lib()
        handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
        B = x.shape[0]
        shape = _capi.Shape(spec.kind, B, spec.bands, spec.classes, int(training))
        sizes = _capi.query_sizes(spec.kind, B, spec.bands, spec.classes, training)
        ptr_of = {n: p.data_ptr() for n, p in zip(names, params)}
        ptr_of.update({n: b.data_ptr() for n, b in buffers.items()})
        table = spec.table_for(ptr_of)
        with torch.cuda.device(dev):
            scores = [torch.empty((B, spec.classes), dtype=torch.float32, device=dev) for _ in range(spec.n_heads)]
            joint = torch.empty((B, spec.classes), dtype=torch.float32, device=dev) if spec.kind == _capi.NET_HANG2020 else None
            saved = torch.empty(sizes.saved_bytes, dtype=torch.uint8, device=dev)
            work = torch.empty(max(sizes.workspace_fwd, 256), dtype=torch.uint8, device=dev)
            score_ptrs = (C.c_void_p * 6)(*[s.data_ptr() for s in scores] + [None] * (6 - spec.n_heads))
            rc = lib.dta_forward(handle, C.byref(shape), x.data_ptr(), C.byref(table), C.byref(score_ptrs),
                                 joint.data_ptr() if joint is not None else None, saved.data_ptr(), work.data_ptr(),
                                 _stream_ptr(dev))
        _capi.check(handle, rc, "dta_forward")
        spec.last_saved = saved if _capi.KEEP_SAVED else None
        ctx.spec, ctx.training, ctx.names, ctx.buffers = spec, training, names, buffers
        ctx.save_for_backward(x, saved, *params)
        ctx.set_materialize_grads(False)
        outs = tuple(scores) + ((joint,) if joint is not None else ())
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        spec = ctx.spec
        x, saved, *params = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("deeptreeattention_b200: gradient w.r.t. the crops is not built; "
                                      "the reference feeds requires_grad=False images (src/main.py:75-77)")
        dev = x.device
        lib = _capi.lib()
        handle = _capi.context(dev.index if dev.index is not None else torch.cuda.current_device())
        B = x.shape[0]
        shape = _capi.Shape(spec.kind, B, spec.bands, spec.classes, int(ctx.training))
        sizes = _capi.query_sizes(spec.kind, B, spec.bands, spec.classes, ctx.training)
        dheads = [g.contiguous().float() if g is not None else None for g in gouts[:spec.n_heads]]
        djoint = gouts[spec.n_heads] if spec.kind == _capi.NET_HANG2020 else None
        if djoint is not None:
            djoint = djoint.contiguous().float()
        ptr_of = {n: p.data_ptr() for n, p in zip(ctx.names, params)}
        ptr_of.update({n: b.data_ptr() for n, b in ctx.buffers.items()})
        table = spec.table_for(ptr_of)
        with torch.cuda.device(dev):
            # every float gradient lives in ONE flat buffer (a single all-reduce covers it)
            numels = [p.numel() if p.dtype == torch.float32 else 0 for p in params]
            bufs = spec.grad_buffers
            persistent = bufs is not None and bufs[0].numel() == sum(numels) and bufs[0].device == dev
            if persistent:
                # A parameter whose .grad is STILL a view of the persistent buffer (gradient accumulation over micro-batches,
                # zero_grad(set_to_none=False), two forwards feeding one backward) must not see that buffer wiped: this
                # backward then writes into a temporary and autograd's AccumulateGrad adds it onto the live views.
                lo, hi = bufs[0].data_ptr(), bufs[0].data_ptr() + bufs[0].numel() * 4
                persistent = not any(p.grad is not None and (lo <= p.grad.data_ptr() < hi or p.grad.data_ptr() == bufs[1].data_ptr())
                                     for p in params)
                accumulating = not persistent
            else:
                accumulating = False
            # dta_backward writes EVERY gradient of the table (kernels for the reached ones, its own zero-fill for dead Conv1d
            # taps and unreached heads, on its side stream), so the buffers need no zero-fill in front of it: two fill kernels
            # less on the step's critical path.  _capi.POISON_GRADS (tests) fills them with NaN to prove that coverage.
            if persistent:
                flat, galpha = bufs
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
            gtable = _fill_tensors(spec.kind, {n: g.data_ptr() for n, g in zip(ctx.names, grads)})
            work = torch.empty(max(sizes.workspace_bwd, 256), dtype=torch.uint8, device=dev)
            dptrs = (C.c_void_p * 6)(*[(d.data_ptr() if d is not None else None) for d in dheads] + [None] * (6 - spec.n_heads))
            rc = lib.dta_backward(handle, C.byref(shape), x.data_ptr(), C.byref(table), saved.data_ptr(), C.byref(dptrs),
                                  djoint.data_ptr() if djoint is not None else None, C.byref(gtable), None,
                                  work.data_ptr(), _stream_ptr(dev))
        _capi.check(handle, rc, "dta_backward")
        if not accumulating:     # (accumulating: the sums live on in the persistent buffer spec.flat_grad already names)
            spec.flat_grad, spec.alpha_grad = flat, (galpha if djoint is not None else None)
        out = []
        for name, g, need in zip(ctx.names, grads, ctx.needs_input_grad[5:]):
            if not need:
                out.append(None)
                continue
            h = _head_of_param(spec.kind, name)
            if h is not None:
                reached = dheads[h] is not None or (djoint is not None and h in (2, 5))
                out.append(g if reached else None)     # reference: grad None for unused heads
            elif name == "alpha":
                out.append(g if djoint is not None else None)
            else:
                out.append(g)
        return (None, None, None, None, None) + tuple(out)


class _FusedNet(Module):
    """Mixin: runs the whole network through the library."""
    _net_kind = None

    def _fused(self, x: torch.Tensor):
        if not isinstance(x, torch.Tensor):
            raise TypeError("expected a tensor of crops (B, bands, 11, 11)")
        if not x.is_cuda:
            raise RuntimeError("deeptreeattention_b200 has no CPU path: move the model and the crops to a CUDA (sm_100) device")
        if x.dtype != torch.float32:
            raise TypeError(f"crops must be float32, got {x.dtype}")
        if x.dim() != 4 or x.shape[1] != self._bands or x.shape[2] != 11 or x.shape[3] != 11:
            raise ValueError(f"expected crops of shape (B, {self._bands}, 11, 11), got {tuple(x.shape)}")
        x = x.contiguous()
        names, params, buffers, spec = self._call_cache()
        spec.grad_buffers = self.__dict__.get("_grad_buffers")
        if params[0].device != x.device:
            raise RuntimeError(f"parameter on {params[0].device} but crops on {x.device}")
        return _FusedNetFunction.apply(x, spec, self.training, names, buffers, *params)

    def _call_cache(self):
        """(names, parameters, buffers, spec) of this module as the library sees it, rebuilt when the module tree changed."""
        # the cached name/tensor lists are valid only while no parameter, buffer or sub-module has been (re)registered
        ident = _STRUCTURE_VERSION[0]
        cache = self.__dict__.get("_fused_cache")
        if cache is not None and cache[4] != ident:
            cache = None
        if cache is None:
            names, params, buffers = [], [], {}
            for k, v in self.state_dict(keep_vars=True).items():
                if isinstance(v, nn.Parameter):
                    names.append(k)
                    params.append(v)
                else:
                    buffers[k] = v
            # the class count is whatever the last head says now (a swapped head may have changed it)
            head = [v for k, v in zip(names, params) if k.endswith("fc1.weight")][-1]
            heads = [v.shape[0] for k, v in zip(names, params) if k.endswith("fc1.weight")]
            if any(h != head.shape[0] for h in heads):
                raise ValueError(f"every classifier head must have the same number of classes, got {heads}")
            self._classes = int(head.shape[0])
            cache = (names, params, buffers, _NetSpec(self._net_kind, self._bands, self._classes), ident)
            self.__dict__["_fused_cache"] = cache
        return cache[:4]

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_fused_cache", None)      # .cuda()/.to() replace buffers (and maybe parameters)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_fused_cache", None)
        return super().load_state_dict(*args, **kwargs)

    def __getstate__(self):
        """copy.deepcopy / pickle / torch.save(model): the call cache (ctypes pointer tables, flat gradient views) and the
        symmetric-memory gradient buffers belong to THIS object on THIS device and are rebuilt on the first call."""
        state = self.__dict__.copy()
        state.pop("_fused_cache", None)
        state.pop("_grad_buffers", None)
        state.pop("head_scores", None)     # last call's outputs (autograd graph attached): not part of the module's state
        for k in ("head_losses", "joint_scores"):
            state.pop(k, None)
        return state

    def fused_spec(self):
        """Spec of the last fused call (holds the flat gradient buffer after a backward)."""
        cache = self.__dict__.get("_fused_cache")
        return cache[3] if cache is not None else None


# ------------------------------------------------------------------- reference API surface
def global_spectral_pool(x):
    """Mean over H, W keeping a trailing singleton: (B, C, H, W) -> (B, C, 1) (reference Hang2020.py:7-12).
    Stand-alone kernel (``dta_plane_mean``); inside the networks the squeeze is fused into the attention kernels."""
    return _blocks.plane_mean(x)


class conv_module(Module):
    """Conv2d 3x3 'same' + BatchNorm2d + ReLU (+ MaxPool2d) block (reference :14-31)."""

    def __init__(self, in_channels, filters, maxpool_kernel=None):
        super().__init__()
        self.conv_layer = nn.Conv2d(in_channels, out_channels=filters, kernel_size=(3, 3), padding="same")
        self.bn1 = nn.BatchNorm2d(filters)
        self.maxpool_kernal = maxpool_kernel
        if maxpool_kernel:
            self.max_pool = nn.MaxPool2d(maxpool_kernel)

    def forward(self, x, pool=False):
        """Stand-alone block forward (``dta_conv_module_forward``; any channel count / plane size).  The network
        modules never call this: they run all their blocks fused."""
        return _blocks.conv_module_forward(self, x, pool)


class Classifier(Module):
    """Linear head (reference :55-66)."""

    def __init__(self, in_features, classes):
        super().__init__()
        self.fc1 = nn.Linear(in_features=in_features, out_features=classes)

    def forward(self, features):
        return _blocks.classifier_forward(self, features)


class spatial_attention(Module):
    """Pixel gate: 1x1 channel pool, two k x k stencils, class max-pool (reference :68-124)."""

    def __init__(self, filters):
        super().__init__()
        sizes = {32: (7, 4), 64: (5, 2), 128: (3, 1)}
        if filters not in sizes:
            raise ValueError("Unknown incoming kernel size {} for attention layers".format(filters))
        kernel_size, pool = sizes[filters]
        self.channel_pool = nn.Conv2d(in_channels=filters, out_channels=1, kernel_size=1)
        self.attention_conv1 = nn.Conv2d(1, 1, kernel_size=kernel_size, padding="same")
        self.attention_conv2 = nn.Conv2d(1, 1, kernel_size=kernel_size, padding="same")
        self.class_pool = nn.MaxPool2d((pool, pool))

    def forward(self, x):
        """Returns (gated feature map, flattened class-pooled features) like the reference (:103-124)."""
        return _blocks.attention_forward(self, _capi.ATTN_SPATIAL, x)


class spectral_attention(Module):
    """Channel gate: squeeze, two Conv1d on a length-1 sequence, sigmoid (reference :126-168)."""

    def __init__(self, filters):
        super().__init__()
        sizes = {32: 3, 64: 5, 128: 7}
        if filters not in sizes:
            raise ValueError("Unknown incoming kernel size {} for attention layers".format(filters))
        kernel_size = sizes[filters]
        self.attention_conv1 = nn.Conv1d(filters, filters, kernel_size=kernel_size, padding="same")
        self.attention_conv2 = nn.Conv1d(filters, filters, kernel_size=kernel_size, padding="same")

    def forward(self, x):
        """Returns (gated feature map, pooled features) like the reference (:146-168)."""
        return _blocks.attention_forward(self, _capi.ATTN_SPECTRAL, x)


class _Branch(_FusedNet):
    _attention = None
    _features = None

    def __init__(self, bands, classes):
        super().__init__()
        self._bands, self._classes = int(bands), int(classes)
        cin = bands
        for k, (c, f) in enumerate(zip((32, 64, 128), self._features), start=1):
            setattr(self, f"conv{k}", conv_module(in_channels=cin, filters=c, maxpool_kernel=(2, 2) if k > 1 else None))
            setattr(self, f"attention_{k}", self._attention(filters=c))
            setattr(self, f"classifier{k}", Classifier(classes=classes, in_features=f))
            cin = c

    def forward(self, x):
        """Returns [scores1, scores2, scores3] like the reference (:190-204 / :226-240)."""
        return list(self._fused(x))


class spatial_network(_Branch):
    """Reference :170-204."""
    _net_kind = _capi.NET_SPATIAL
    _attention = spatial_attention
    _features = (128, 256, 512)


class spectral_network(_Branch):
    """Reference :206-240."""
    _net_kind = _capi.NET_SPECTRAL
    _attention = spectral_attention
    _features = (32, 64, 128)


class vanilla_CNN(_FusedNet):
    """Baseline without attention (reference :33-53)."""
    _net_kind = _capi.NET_VANILLA

    def __init__(self, bands, classes):
        super().__init__()
        self._bands, self._classes = int(bands), int(classes)
        self.conv1 = conv_module(in_channels=bands, filters=32)
        self.conv2 = conv_module(in_channels=32, filters=64, maxpool_kernel=(2, 2))
        self.conv3 = conv_module(in_channels=64, filters=128, maxpool_kernel=(2, 2))
        self.fc1 = nn.Linear(in_features=512, out_features=classes)

    def forward(self, x):
        return self._fused(x)[0]


class Hang2020(_FusedNet):
    """Both branches on the same crops, alpha-blended last heads (reference :242-263)."""
    _net_kind = _capi.NET_HANG2020

    def __init__(self, bands, classes):
        super().__init__()
        self._bands, self._classes = int(bands), int(classes)
        self.spectral_network = spectral_network(bands, classes)
        self.spatial_network = spatial_network(bands, classes)
        self.alpha = nn.Parameter(torch.tensor(0.5, dtype=float), requires_grad=True)

    @property
    def weighted_average(self):
        """sigmoid(alpha), the blend weight of the spectral branch (reference :260 stores it in forward; nothing in the
        reference reads it back, so it is evaluated on access instead of costing a kernel launch every step)."""
        return torch.sigmoid(self.alpha.detach())

    def forward(self, x):
        outs = self._fused(x)
        self.head_scores = list(outs[:6])        # spectral 1-3, spatial 1-3 (extra over the reference)
        return outs[6]


def load_from_backbone(state_dict, classes, bands):
    """Fresh spectral_network with every non-classifier tensor taken from a saved
    spectral_network state_dict file (reference :266-278)."""
    saved = torch.load(state_dict, map_location="cpu")
    model = spectral_network(classes=classes, bands=bands)
    merged = model.state_dict()
    merged.update({k: v for k, v in saved.items() if "classifier" not in k})
    model.load_state_dict(merged)
    return model
