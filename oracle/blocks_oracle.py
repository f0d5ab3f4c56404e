"""CPU oracle for the STAND-ALONE building blocks of the reference's src/models/Hang2020.py -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import anything under oracle/; the product path
(deeptreeattention_b200/) never does.  Functional restatements with stock ATen CPU ops (backward = autograd, as in the
reference) of

  global_spectral_pool   /root/reference/src/models/Hang2020.py:7-12
  conv_module.forward    :24-31  (any pooling kernel handed to the constructor, :20-22)
  Classifier.forward     :63-66
  spatial_attention      :103-124  (via hang2020_oracle.spatial_gate)
  spectral_attention     :146-168  (via hang2020_oracle.spectral_gate)

on the shapes the reference's own tests exercise (tests/test_Hang2020.py:8-33).  Parity is pinned by
tests/golden/blocks.npz, produced by tests/golden/make_block_golden.py from the reference modules themselves.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import hang2020_oracle as orc

Params = Dict[str, torch.Tensor]

# name, block, constructor args, input shape, training, pool flag
CASES = [
    dict(name="conv_369_32_train", block="conv_module", args=dict(in_channels=369, filters=32), shape=(4, 369, 11, 11), training=True, pool=False),
    dict(name="conv_32_64_pool_train", block="conv_module", args=dict(in_channels=32, filters=64, maxpool_kernel=(2, 2)), shape=(5, 32, 11, 11),
         training=True, pool=True),
    dict(name="conv_64_128_pool_eval", block="conv_module", args=dict(in_channels=64, filters=128, maxpool_kernel=(2, 2)), shape=(3, 64, 5, 5),
         training=False, pool=True),
    dict(name="conv_7_12_pool32_train", block="conv_module", args=dict(in_channels=7, filters=12, maxpool_kernel=(3, 2)), shape=(4, 7, 9, 6),
         training=True, pool=True),
    dict(name="spectral_32", block="spectral_attention", args=dict(filters=32), shape=(4, 32, 11, 11), training=True, pool=False),
    dict(name="spectral_64", block="spectral_attention", args=dict(filters=64), shape=(4, 64, 5, 5), training=True, pool=False),
    dict(name="spectral_128", block="spectral_attention", args=dict(filters=128), shape=(4, 128, 2, 2), training=True, pool=False),
    dict(name="spatial_32", block="spatial_attention", args=dict(filters=32), shape=(4, 32, 11, 11), training=True, pool=False),
    dict(name="spatial_64", block="spatial_attention", args=dict(filters=64), shape=(4, 64, 5, 5), training=True, pool=False),
    dict(name="spatial_128", block="spatial_attention", args=dict(filters=128), shape=(4, 128, 2, 2), training=True, pool=False),
    dict(name="classifier_128_10", block="Classifier", args=dict(in_features=128, classes=10), shape=(6, 128), training=True, pool=False),
    dict(name="pool_mean", block="global_spectral_pool", args=dict(), shape=(3, 5, 4, 7), training=True, pool=False),
]


def param_shapes(case) -> List[Tuple[str, Tuple[int, ...]]]:
    a, blk = case["args"], case["block"]
    if blk == "conv_module":
        c, cin = a["filters"], a["in_channels"]
        return [("conv_layer.weight", (c, cin, 3, 3)), ("conv_layer.bias", (c,)), ("bn1.weight", (c,)), ("bn1.bias", (c,)),
                ("bn1.running_mean", (c,)), ("bn1.running_var", (c,)), ("bn1.num_batches_tracked", ())]
    if blk == "spectral_attention":
        c, ks = a["filters"], orc.spectral_kernel_size(a["filters"])
        return [("attention_conv1.weight", (c, c, ks)), ("attention_conv1.bias", (c,)), ("attention_conv2.weight", (c, c, ks)),
                ("attention_conv2.bias", (c,))]
    if blk == "spatial_attention":
        c, ks = a["filters"], orc.spatial_kernel_size(a["filters"])
        return [("channel_pool.weight", (1, c, 1, 1)), ("channel_pool.bias", (1,)), ("attention_conv1.weight", (1, 1, ks, ks)),
                ("attention_conv1.bias", (1,)), ("attention_conv2.weight", (1, 1, ks, ks)), ("attention_conv2.bias", (1,))]
    if blk == "Classifier":
        return [("fc1.weight", (a["classes"], a["in_features"])), ("fc1.bias", (a["classes"],))]
    return []


def build_case(case, seed: int):
    """Seeded (numpy PCG64, independent of the torch RNG) parameters, input and upstream gradients of a case."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p: Params = {}
    for name, shape in param_shapes(case):
        if name.endswith("num_batches_tracked"):
            p[name] = torch.tensor(3, dtype=torch.int64)
        elif name.endswith("running_var"):
            p[name] = torch.from_numpy(rng.uniform(0.5, 1.5, size=shape).astype(np.float32))
        elif name.endswith("running_mean"):
            p[name] = torch.from_numpy(rng.normal(0.0, 0.2, size=shape).astype(np.float32))
        elif name == "bn1.weight":
            p[name] = torch.from_numpy(rng.uniform(0.5, 1.5, size=shape).astype(np.float32))
        elif name == "bn1.bias":
            p[name] = torch.from_numpy(rng.normal(0.0, 0.2, size=shape).astype(np.float32))
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
            bound = 1.0 / np.sqrt(max(fan_in, 1))
            p[name] = torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))
    x = torch.from_numpy(rng.normal(0.0, 1.0, size=case["shape"]).astype(np.float32))
    return p, x, rng


def upstream(rng, outs):
    return [torch.from_numpy(rng.normal(0.0, 1.0, size=tuple(o.shape)).astype(np.float32)) for o in outs]


def forward(case, p: Params, x: torch.Tensor):
    """Tuple of outputs of the block, restated with ATen CPU ops."""
    blk, a = case["block"], case["args"]
    if blk == "global_spectral_pool":
        return (torch.mean(x, dim=(2, 3)).unsqueeze(-1),)                                   # :9-10
    if blk == "conv_module":
        z = F.conv2d(x, p["conv_layer.weight"], p["conv_layer.bias"], padding=1)              # :25 (padding="same", k=3)
        y = F.batch_norm(z, p["bn1.running_mean"], p["bn1.running_var"], p["bn1.weight"], p["bn1.bias"],
                         training=case["training"], momentum=orc.BN_MOMENTUM, eps=orc.BN_EPS)  # :26
        if case["training"]:
            p["bn1.num_batches_tracked"] += 1
        y = F.relu(y)                                                                         # :27
        if case["pool"]:
            y = F.max_pool2d(y, a["maxpool_kernel"])                                          # :28-29
        return (y,)
    if blk == "Classifier":
        return (F.linear(x, p["fc1.weight"], p["fc1.bias"]),)                                 # :64
    table = {"m." + k: v for k, v in p.items()}
    if blk == "spectral_attention":
        return orc.spectral_gate(table, "m", x)
    if blk == "spatial_attention":
        return orc.spatial_gate(table, "m", x)
    raise ValueError(blk)


def step(case, seed: int):
    """(outputs, {param or 'x': gradient}, buffers after the call) for L = sum_k <out_k, G_k> with seeded G_k."""
    p, x, rng = build_case(case, seed)
    x.requires_grad_(True)
    leaves = {k: v.requires_grad_(True) for k, v in p.items() if v.is_floating_point() and "running" not in k}
    outs = forward(case, p, x)
    gs = upstream(rng, outs)
    loss = sum((o * g).sum() for o, g in zip(outs, gs))
    loss.backward()
    grads = {k: v.grad.detach() for k, v in leaves.items()}
    grads["x"] = x.grad.detach()
    bufs = {k: v.detach() for k, v in p.items() if "running" in k or k.endswith("num_batches_tracked")}
    return [o.detach() for o in outs], grads, bufs
