"""CPU oracle for the Hang2020 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product package
(``deeptreeattention_b200``) never does; it fails loudly when the CUDA library is
missing instead of falling back to anything here.

What it is: a functional restatement (plain functions over a ``{name: tensor}``
parameter table, stock ATen ops on the CPU) of the forward pass the reference builds
out of ``torch.nn`` layers in ``/root/reference/src/models/Hang2020.py``; the backward
pass is whatever autograd derives from it, exactly as in the reference.  Every
function cites the reference lines it follows.  Parameter names are the reference's
``state_dict`` keys, so a table can be loaded into the
reference module (done by ``tests/golden/make_golden.py`` in the build container)
and into the B200 module unchanged.

Parity status: **pinned by execution of the reference itself**.  The reference's own
tests hold no golden values (shape-only, ``tests/test_Hang2020.py:8-75``), so
``tests/golden/make_golden.py`` imported the reference module in the build container,
ran it on seeded inputs/parameters and committed its outputs and gradients under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against those
fixtures on every CPU run.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

NET_KINDS = ("hang2020", "spectral", "spatial", "vanilla")
FILTERS = (32, 64, 128)
BN_EPS = 1e-5          # nn.BatchNorm2d default, Hang2020.py:19
BN_MOMENTUM = 0.1      # nn.BatchNorm2d default, Hang2020.py:19
IMAGE_SIZE = 11        # config.yml:50


# --------------------------------------------------------------------------- params
def spectral_kernel_size(filters: int) -> int:
    """Conv1d kernel width per block width (Hang2020.py:136-144)."""
    try:
        return {32: 3, 64: 5, 128: 7}[filters]
    except KeyError:
        raise ValueError(f"Unknown incoming kernel size {filters} for attention layers")


def spatial_kernel_size(filters: int) -> int:
    """k x k attention stencil per block width (Hang2020.py:77-85)."""
    try:
        return {32: 7, 64: 5, 128: 3}[filters]
    except KeyError:
        raise ValueError(f"Unknown incoming kernel size {filters} for attention layers")


def spatial_class_pool(filters: int) -> Tuple[int, int]:
    """(pool window, head in_features) per block width (Hang2020.py:91-101)."""
    try:
        return {32: (4, 128), 64: (2, 256), 128: (1, 512)}[filters]
    except KeyError:
        raise ValueError("Unknown filter size for max pooling")


def param_shapes(kind: str, bands: int, classes: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, role) in ``state_dict()`` order.

    role in {"w", "b", "bn_w", "bn_b", "bn_rm", "bn_rv", "bn_nbt", "alpha"}; for "w"/"b"
    the initialiser needs fan_in, recovered from the weight shape.
    """
    def conv_block(prefix, cin, cout):
        return [
            (f"{prefix}.conv_layer.weight", (cout, cin, 3, 3), "w"),
            (f"{prefix}.conv_layer.bias", (cout,), "b"),
            (f"{prefix}.bn1.weight", (cout,), "bn_w"),
            (f"{prefix}.bn1.bias", (cout,), "bn_b"),
            (f"{prefix}.bn1.running_mean", (cout,), "bn_rm"),
            (f"{prefix}.bn1.running_var", (cout,), "bn_rv"),
            (f"{prefix}.bn1.num_batches_tracked", (), "bn_nbt"),
        ]

    def branch(prefix, attn):
        out = []
        cin = bands
        for k, c in enumerate(FILTERS, start=1):
            out += conv_block(f"{prefix}conv{k}", cin, c)
            if attn == "spectral":
                ks = spectral_kernel_size(c)
                out += [
                    (f"{prefix}attention_{k}.attention_conv1.weight", (c, c, ks), "w"),
                    (f"{prefix}attention_{k}.attention_conv1.bias", (c,), "b"),
                    (f"{prefix}attention_{k}.attention_conv2.weight", (c, c, ks), "w"),
                    (f"{prefix}attention_{k}.attention_conv2.bias", (c,), "b"),
                ]
                feat = c
            else:
                ks = spatial_kernel_size(c)
                out += [
                    (f"{prefix}attention_{k}.channel_pool.weight", (1, c, 1, 1), "w"),
                    (f"{prefix}attention_{k}.channel_pool.bias", (1,), "b"),
                    (f"{prefix}attention_{k}.attention_conv1.weight", (1, 1, ks, ks), "w"),
                    (f"{prefix}attention_{k}.attention_conv1.bias", (1,), "b"),
                    (f"{prefix}attention_{k}.attention_conv2.weight", (1, 1, ks, ks), "w"),
                    (f"{prefix}attention_{k}.attention_conv2.bias", (1,), "b"),
                ]
                feat = spatial_class_pool(c)[1]
            out += [
                (f"{prefix}classifier{k}.fc1.weight", (classes, feat), "w"),
                (f"{prefix}classifier{k}.fc1.bias", (classes,), "b"),
            ]
            cin = c
        return out

    if kind == "hang2020":
        return ([("alpha", (), "alpha")]
                + branch("spectral_network.", "spectral")
                + branch("spatial_network.", "spatial"))
    if kind == "spectral":
        return branch("", "spectral")
    if kind == "spatial":
        return branch("", "spatial")
    if kind == "vanilla":
        out = []
        cin = bands
        for k, c in enumerate(FILTERS, start=1):
            out += conv_block(f"conv{k}", cin, c)
            cin = c
        out += [("fc1.weight", (classes, 512), "w"), ("fc1.bias", (classes,), "b")]
        return out
    raise ValueError(f"unknown net kind {kind!r}")


def init_params(kind: str, bands: int, classes: int, seed: int,
                perturb_bn: bool = False, dtype=torch.float32) -> Params:
    """Deterministic, torch-RNG-independent parameter table.

    Same *distribution* as the reference's default init (PyTorch: kaiming_uniform(a=sqrt 5)
    == U(+-1/sqrt(fan_in)) for conv/linear weights and biases; BN gamma=1, beta=0,
    running_mean=0, running_var=1; alpha=0.5 float64, Hang2020.py:249) but drawn from
    numpy's PCG64 so that fixtures do not depend on the torch build.  ``perturb_bn``
    randomises the BN affine/running tensors so eval-mode tests are non-trivial.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    table: Params = {}
    fan_in = 1
    for name, shape, role in param_shapes(kind, bands, classes):
        if role == "w":
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            t = torch.from_numpy(rng.uniform(-bound, bound, size=shape)).to(dtype)
        elif role == "b":
            bound = 1.0 / math.sqrt(fan_in)
            t = torch.from_numpy(rng.uniform(-bound, bound, size=shape)).to(dtype)
        elif role == "bn_w":
            t = torch.ones(shape, dtype=dtype)
            if perturb_bn:
                t = torch.from_numpy(rng.uniform(0.5, 1.5, size=shape)).to(dtype)
        elif role in ("bn_b", "bn_rm"):
            t = torch.zeros(shape, dtype=dtype)
            if perturb_bn:
                t = torch.from_numpy(rng.uniform(-0.2, 0.2, size=shape)).to(dtype)
        elif role == "bn_rv":
            t = torch.ones(shape, dtype=dtype)
            if perturb_bn:
                t = torch.from_numpy(rng.uniform(0.5, 1.5, size=shape)).to(dtype)
        elif role == "bn_nbt":
            t = torch.zeros((), dtype=torch.int64)
        elif role == "alpha":
            t = torch.tensor(0.5, dtype=torch.float64)
        else:  # pragma: no cover
            raise AssertionError(role)
        table[name] = t
    return table


def is_buffer(name: str) -> bool:
    return name.endswith(("running_mean", "running_var", "num_batches_tracked"))


def trainable(params: Params) -> Params:
    return {k: v for k, v in params.items() if not is_buffer(k)}


def make_inputs(batch: int, bands: int, classes: int, seed: int, dist: str = "uniform"):
    """Synthetic crops: U[0,1) like the loader's per-pixel min-max scaling (src/utils.py:49)
    or randn like the reference tests (tests/test_Hang2020.py:10); labels uniform int64."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    if dist == "uniform":
        x = rng.random(size=(batch, bands, IMAGE_SIZE, IMAGE_SIZE), dtype=np.float32)
    elif dist == "normal":
        x = rng.standard_normal(size=(batch, bands, IMAGE_SIZE, IMAGE_SIZE), dtype=np.float32)
    else:
        raise ValueError(dist)
    y = rng.integers(0, classes, size=(batch,), dtype=np.int64)
    return torch.from_numpy(x), torch.from_numpy(y)


# --------------------------------------------------------------------------- forward
Z_RECORD: Optional[Dict[str, torch.Tensor]] = None   # tests: set to a dict to collect every block's convolution output
# tests / fixture screening: "split_bf16x3" evaluates every 3x3 convolution's VALUE in the arithmetic of the tensor-core kernels
# (operands split v = hi + lo into two round-to-nearest bf16, products hi*hi + hi*lo + lo*hi, csrc/dta_tc.cuh split2) while
# the derivative stays the exact one: shows which ReLU / max-pool decisions that arithmetic can move
CONV_ARITH: Optional[str] = None


def split_bf16(v: torch.Tensor):
    hi = v.float().bfloat16().float()
    lo = (v.float() - hi).bfloat16().float()
    return hi.to(v.dtype), lo.to(v.dtype)


def conv_block(p: Params, prefix: str, u: torch.Tensor, pool: bool, training: bool,
               z_values: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """conv3x3 'same' -> BatchNorm2d -> ReLU -> optional 2x2 floor max-pool.

    Follows conv_module.forward, Hang2020.py:24-31 (layers declared :18-22).  In
    training mode F.batch_norm updates running_mean/var in place (momentum 0.1,
    unbiased variance) and num_batches_tracked is bumped like nn.BatchNorm2d does.

    ``z_values`` (tests only): ``{prefix: tensor}`` of convolution outputs measured on another implementation.  The
    block then continues from THOSE values while gradients still flow through this oracle's own convolution
    (value substitution, derivative unchanged): every ReLU / max-pool decision downstream is taken on the same
    numbers as in the implementation under test, which is what makes a gradient comparison well defined for a
    piecewise-linear network (see tests/test_gpu_parity.py).  ``z_values["bn:" + prefix]``: likewise the values AFTER
    BatchNorm, i.e. the numbers whose sign the ReLU tests (an implementation that normalises as fmaf(z, scale, shift) in
    float32 can disagree with a float64 BatchNorm about the sign of an activation that is 1e-8 from zero).
    """
    z = F.conv2d(u, p[f"{prefix}.conv_layer.weight"], p[f"{prefix}.conv_layer.bias"], padding=1)
    if CONV_ARITH == "split_bf16x3":
        with torch.no_grad():
            uh, ul = split_bf16(u)
            wh, wl = split_bf16(p[f"{prefix}.conv_layer.weight"])
            z_em = (F.conv2d(uh, wh, p[f"{prefix}.conv_layer.bias"], padding=1) + F.conv2d(uh, wl, None, padding=1)
                    + F.conv2d(ul, wh, None, padding=1))
        z = z + (z_em - z).detach()
    if Z_RECORD is not None:
        Z_RECORD[prefix] = z.detach().clone()
    if z_values is not None and prefix in z_values:
        z = z + (z_values[prefix].to(z.dtype) - z).detach()
    a = F.batch_norm(z, p[f"{prefix}.bn1.running_mean"], p[f"{prefix}.bn1.running_var"],
                     p[f"{prefix}.bn1.weight"], p[f"{prefix}.bn1.bias"],
                     training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
    if training:
        p[f"{prefix}.bn1.num_batches_tracked"] += 1
    if z_values is not None and ("bn:" + prefix) in z_values:      # the other implementation's post-BatchNorm values (see above)
        a = a + (z_values["bn:" + prefix].to(a.dtype) - a).detach()
    r = F.relu(a)
    if pool:
        r = F.max_pool2d(r, (2, 2))
    return r


def spectral_gate(p: Params, prefix: str, r: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Channel attention.  spectral_attention.forward, Hang2020.py:149-168 with
    global_spectral_pool :7-12.  The Conv1d runs on a length-1 sequence so only its
    centre tap contributes; it is still evaluated as a Conv1d here so that the dead
    taps receive their (zero) gradient exactly as in the reference."""
    ks = p[f"{prefix}.attention_conv1.weight"].shape[-1]
    g = r.mean(dim=(2, 3)).unsqueeze(-1)
    h = F.relu(F.conv1d(g, p[f"{prefix}.attention_conv1.weight"],
                        p[f"{prefix}.attention_conv1.bias"], padding=ks // 2))
    s = torch.sigmoid(F.conv1d(h, p[f"{prefix}.attention_conv2.weight"],
                               p[f"{prefix}.attention_conv2.bias"], padding=ks // 2))
    out = r * s.unsqueeze(-1)
    feat = out.mean(dim=(2, 3))
    return out, feat


def spatial_gate(p: Params, prefix: str, r: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pixel attention.  spatial_attention.forward, Hang2020.py:105-124: 1x1 channel
    pool + ReLU, two k x k 'same' stencils (ReLU, sigmoid), gate, class max-pool
    (window 4/2/1, floor) and channel-major flatten."""
    c = r.shape[1]
    ks = p[f"{prefix}.attention_conv1.weight"].shape[-1]
    window, _ = spatial_class_pool(c)
    q = F.relu(F.conv2d(r, p[f"{prefix}.channel_pool.weight"], p[f"{prefix}.channel_pool.bias"]))
    t = F.relu(F.conv2d(q, p[f"{prefix}.attention_conv1.weight"],
                        p[f"{prefix}.attention_conv1.bias"], padding=ks // 2))
    s = torch.sigmoid(F.conv2d(t, p[f"{prefix}.attention_conv2.weight"],
                               p[f"{prefix}.attention_conv2.bias"], padding=ks // 2))
    out = r * s
    feat = torch.flatten(F.max_pool2d(out, (window, window)), start_dim=1)
    return out, feat


def branch_forward(p: Params, prefix: str, attn: str, x: torch.Tensor, training: bool, z_values=None) -> List[torch.Tensor]:
    """Three (conv block -> attention -> head) stages; returns the three head scores.
    spectral_network.forward Hang2020.py:226-240 / spatial_network.forward :190-204."""
    gate = spectral_gate if attn == "spectral" else spatial_gate
    scores = []
    u = x
    for k in (1, 2, 3):
        r = conv_block(p, f"{prefix}conv{k}", u, pool=(k > 1), training=training, z_values=z_values)
        u, feat = gate(p, f"{prefix}attention_{k}", r)
        scores.append(F.linear(feat, p[f"{prefix}classifier{k}.fc1.weight"],
                               p[f"{prefix}classifier{k}.fc1.bias"]))   # Classifier, :63-66
    return scores


def forward(kind: str, p: Params, x: torch.Tensor, training: bool, z_values=None):
    """Returns (result, heads): ``result`` is what the reference module returns
    (joint scores for hang2020 :251-263, list of 3 for the sub-networks, scores for
    vanilla_CNN :45-53); ``heads`` is the list of every head's scores in branch-major
    order (spectral 1-3 then spatial 1-3 for hang2020)."""
    if kind == "hang2020":
        spec = branch_forward(p, "spectral_network.", "spectral", x, training, z_values)
        spat = branch_forward(p, "spatial_network.", "spatial", x, training, z_values)
        w = torch.sigmoid(p["alpha"])                     # float64 0-dim, :259
        joint = spec[-1] * w + spat[-1] * (1 - w)         # float32 result, :260
        return joint, spec + spat
    if kind in ("spectral", "spatial"):
        heads = branch_forward(p, "", kind, x, training, z_values)
        return heads, heads
    if kind == "vanilla":
        u = x
        for k in (1, 2, 3):
            u = conv_block(p, f"conv{k}", u, pool=(k > 1), training=training, z_values=z_values)
        s = F.linear(torch.flatten(u, start_dim=1), p["fc1.weight"], p["fc1.bias"])
        return s, [s]
    raise ValueError(f"unknown net kind {kind!r}")


def weighted_ce(scores: torch.Tensor, y: torch.Tensor, weight: Optional[torch.Tensor] = None):
    """TreeModel.training_step loss, src/main.py:78 (class weights = ones on CPU, :69)."""
    return F.cross_entropy(scores, y, weight=weight)


def loss_regime(regime: str, result, heads, y, weight=None):
    """R1: the reference's CE(module output) (main.py:71-80; for the sub-networks the
    production caller takes the last head, src/models/year.py:29-30).
    R2: the north-star regime, sum of CE over every head."""
    if regime == "R1":
        out = result[-1] if isinstance(result, list) else result
        return weighted_ce(out, y, weight)
    if regime == "R2":
        return sum(weighted_ce(s, y, weight) for s in heads)
    raise ValueError(regime)


def step(kind: str, params: Params, x, y, regime: str = "R1", training: bool = True,
         weight=None, num_threads: Optional[int] = None, z_values=None):
    """One forward + loss + backward.  Returns (loss, result, heads, grads) with grads a
    ``{name: tensor-or-None}`` table over the trainable parameters."""
    if num_threads:
        torch.set_num_threads(num_threads)
    p = {k: (v.clone().requires_grad_(True) if not is_buffer(k) else v.clone())
         for k, v in params.items()}
    result, heads = forward(kind, p, x, training, z_values)
    loss = loss_regime(regime, result, heads, y, weight)
    names = [k for k in p if not is_buffer(k)]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    buffers = {k: v for k, v in p.items() if is_buffer(k)}
    return loss.detach(), result, heads, dict(zip(names, grads)), buffers


class OracleModule(torch.nn.Module):
    """nn.Module shell over the functional oracle (used as the CPU baseline arm in
    bench.py and as the comparison model in tests): parameters live in a ParameterDict-
    like table keyed by the reference's state_dict names."""

    def __init__(self, kind: str, bands: int, classes: int, seed: int = 0):
        super().__init__()
        self.kind = kind
        self._names = []
        for name, t in init_params(kind, bands, classes, seed).items():
            attr = name.replace(".", "__")
            self._names.append((name, attr))
            if is_buffer(name):
                self.register_buffer(attr, t)
            else:
                self.register_parameter(attr, torch.nn.Parameter(t))

    def table(self) -> Params:
        return {name: getattr(self, attr) for name, attr in self._names}

    def load_table(self, table: Params):
        with torch.no_grad():
            for name, attr in self._names:
                getattr(self, attr).copy_(table[name])

    def forward(self, x):
        result, self.heads = forward(self.kind, self.table(), x, self.training)
        return result
