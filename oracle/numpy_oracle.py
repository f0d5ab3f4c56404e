"""Second, independent CPU oracle: the Hang2020 FORWARD pass in plain numpy float64 -- TEST INFRASTRUCTURE ONLY.

``hang2020_oracle`` restates the reference with the same ATen kernels the reference's torch.nn layers dispatch to; this
file shares nothing with it but the parameter table: explicit shifted-slice convolutions, explicit BatchNorm statistics,
explicit pooling loops.  It pins the SEMANTICS (padding, pooling floor, biased/unbiased variance, centre-tap Conv1d,
flatten order, alpha blend) independently of PyTorch; tests/test_oracle_golden.py checks it against the golden vectors the
reference produced.  Line references: /root/reference/src/models/Hang2020.py.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _f64(t) -> np.ndarray:
    return np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float64)


def conv_same(u: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """k x k 'same' cross-correlation with zero padding (:18, :87-88): sum over taps of shifted slices."""
    B, cin, H, W = u.shape
    k = w.shape[-1]
    r = k // 2
    up = np.zeros((B, cin, H + 2 * r, W + 2 * r))
    up[:, :, r:r + H, r:r + W] = u
    out = np.zeros((B, w.shape[0], H, W))
    for dy in range(k):
        for dx in range(k):
            out += np.einsum("bchw,oc->bohw", up[:, :, dy:dy + H, dx:dx + W], w[:, :, dy, dx])
    return out + b[None, :, None, None]


def batch_norm(z, gamma, beta, rm, rv, training: bool, buffers_out: Dict[str, np.ndarray], prefix: str):
    """nn.BatchNorm2d (:19, :26): batch statistics with biased variance in train(), running statistics in eval();
    the running update uses momentum 0.1 and the unbiased variance."""
    if training:
        mean = z.mean(axis=(0, 2, 3))
        var = z.var(axis=(0, 2, 3))
        n = z.shape[0] * z.shape[2] * z.shape[3]
        buffers_out[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
        buffers_out[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * n / max(n - 1, 1)
    else:
        mean, var = rm, rv
    zhat = (z - mean[None, :, None, None]) / np.sqrt(var[None, :, None, None] + BN_EPS)
    return zhat * gamma[None, :, None, None] + beta[None, :, None, None]


def max_pool(a: np.ndarray, k: int) -> np.ndarray:
    """nn.MaxPool2d((k, k)): stride = kernel, floor (:20-22, :91-101)."""
    B, C, H, W = a.shape
    Ho, Wo = H // k, W // k
    return a[:, :, :Ho * k, :Wo * k].reshape(B, C, Ho, k, Wo, k).max(axis=(3, 5))


def conv_block(p, prefix, u, pool, training, buffers_out):
    z = conv_same(u, _f64(p[f"{prefix}.conv_layer.weight"]), _f64(p[f"{prefix}.conv_layer.bias"]))     # :25
    a = batch_norm(z, _f64(p[f"{prefix}.bn1.weight"]), _f64(p[f"{prefix}.bn1.bias"]), _f64(p[f"{prefix}.bn1.running_mean"]),
                   _f64(p[f"{prefix}.bn1.running_var"]), training, buffers_out, f"{prefix}.bn1")       # :26
    r = np.maximum(a, 0.0)                                                                              # :27
    return max_pool(r, 2) if pool else r                                                                # :28-29


def sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def spectral_gate(p, prefix, r):
    """:146-168.  A Conv1d with 'same' padding over a length-1 sequence only ever sees its centre tap."""
    w1, w2 = _f64(p[f"{prefix}.attention_conv1.weight"]), _f64(p[f"{prefix}.attention_conv2.weight"])
    mid = w1.shape[-1] // 2
    g = r.mean(axis=(2, 3))                                                                             # :7-12
    h = np.maximum(g @ w1[:, :, mid].T + _f64(p[f"{prefix}.attention_conv1.bias"]), 0.0)
    s = sigmoid(h @ w2[:, :, mid].T + _f64(p[f"{prefix}.attention_conv2.bias"]))
    out = r * s[:, :, None, None]
    return out, out.mean(axis=(2, 3))


def spatial_gate(p, prefix, r):
    """:103-124; class pool 4 / 2 / 1 for 32 / 64 / 128 filters (:91-101), channel-major flatten."""
    c = r.shape[1]
    window = {32: 4, 64: 2, 128: 1}[c]
    q = np.maximum(conv_same(r, _f64(p[f"{prefix}.channel_pool.weight"]), _f64(p[f"{prefix}.channel_pool.bias"])), 0.0)
    t = np.maximum(conv_same(q, _f64(p[f"{prefix}.attention_conv1.weight"]), _f64(p[f"{prefix}.attention_conv1.bias"])), 0.0)
    s = sigmoid(conv_same(t, _f64(p[f"{prefix}.attention_conv2.weight"]), _f64(p[f"{prefix}.attention_conv2.bias"])))
    out = r * s
    return out, max_pool(out, window).reshape(r.shape[0], -1)


def branch(p, prefix, attn, x, training, buffers_out) -> List[np.ndarray]:
    gate = spectral_gate if attn == "spectral" else spatial_gate
    scores, u = [], x
    for k in (1, 2, 3):
        r = conv_block(p, f"{prefix}conv{k}", u, k > 1, training, buffers_out)
        u, feat = gate(p, f"{prefix}attention_{k}", r)
        scores.append(feat @ _f64(p[f"{prefix}classifier{k}.fc1.weight"]).T + _f64(p[f"{prefix}classifier{k}.fc1.bias"]))   # :64
    return scores


def forward(kind: str, p, x, training: bool):
    """(result, heads, BatchNorm buffers after the call) in float64; same conventions as hang2020_oracle.forward."""
    x = _f64(x)
    bufs: Dict[str, np.ndarray] = {}
    if kind == "hang2020":
        spec = branch(p, "spectral_network.", "spectral", x, training, bufs)
        spat = branch(p, "spatial_network.", "spatial", x, training, bufs)
        w = sigmoid(float(_f64(p["alpha"])))                                                            # :259
        return spec[-1] * w + spat[-1] * (1.0 - w), spec + spat, bufs                                   # :260
    if kind in ("spectral", "spatial"):
        heads = branch(p, "", kind, x, training, bufs)
        return heads[-1], heads, bufs
    if kind == "vanilla":
        u = x
        for k in (1, 2, 3):
            u = conv_block(p, f"conv{k}", u, k > 1, training, bufs)
        s = u.reshape(u.shape[0], -1) @ _f64(p["fc1.weight"]).T + _f64(p["fc1.bias"])                   # :50-51
        return s, [s], bufs
    raise ValueError(kind)
