"""CPU oracle for BASELINE config 5 (site metadata + late fusion) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional restatement of ``/root/reference/src/models/metadata.py`` over a ``{state_dict key: tensor}`` table with stock
ATen CPU ops (backward = autograd, as in the reference).  Only ``tests/`` and ``bench.py``'s CPU legs import it.

Parity status: pinned by execution of the reference itself -- ``tests/golden/make_metadata_golden.py`` imports the reference
``metadata`` / ``metadata_sensor_fusion`` modules (``src.main`` stubbed), runs them on seeded inputs and commits
outputs, losses, gradients and the dropout masks they drew under ``tests/golden/metadata_*.npz``;
``tests/test_metadata.py`` checks this file against those fixtures on every CPU run.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import hang2020_oracle as orc

EMBED_DIM = 16      # nn.Embedding(sites, 16), metadata.py:12
DROP_P = 0.7        # nn.Dropout(p=0.7), metadata.py:15


def init_meta_params(sites: int, classes: int, seed: int, fused: bool, perturb_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded table for the metadata branch (prefix ``metadata_model.`` and ``fc1.*`` when ``fused``), numpy PCG64 like
    ``hang2020_oracle.init_params``: N(0,1) embedding (nn.Embedding default), U(+-1/sqrt(fan_in)) linears."""
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    pre = "metadata_model." if fused else ""
    t = {}
    t[pre + "embedding.weight"] = torch.from_numpy(rng.standard_normal((sites, EMBED_DIM)).astype(np.float32))
    t[pre + "batch_norm.weight"] = torch.from_numpy(rng.uniform(0.5, 1.5, EMBED_DIM).astype(np.float32)) if perturb_bn else torch.ones(EMBED_DIM)
    t[pre + "batch_norm.bias"] = torch.from_numpy(rng.uniform(-0.2, 0.2, EMBED_DIM).astype(np.float32)) if perturb_bn else torch.zeros(EMBED_DIM)
    t[pre + "batch_norm.running_mean"] = torch.from_numpy(rng.uniform(-0.2, 0.2, EMBED_DIM).astype(np.float32)) if perturb_bn else torch.zeros(EMBED_DIM)
    t[pre + "batch_norm.running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, EMBED_DIM).astype(np.float32)) if perturb_bn else torch.ones(EMBED_DIM)
    t[pre + "batch_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    b = 1.0 / math.sqrt(EMBED_DIM)
    t[pre + "mlp.weight"] = torch.from_numpy(rng.uniform(-b, b, (classes, EMBED_DIM)).astype(np.float32))
    t[pre + "mlp.bias"] = torch.from_numpy(rng.uniform(-b, b, classes).astype(np.float32))
    if fused:
        b = 1.0 / math.sqrt(2 * classes)
        t["fc1.weight"] = torch.from_numpy(rng.uniform(-b, b, (classes, 2 * classes)).astype(np.float32))
        t["fc1.bias"] = torch.from_numpy(rng.uniform(-b, b, classes).astype(np.float32))
    return t


def init_fusion_params(bands: int, sites: int, classes: int, seed: int) -> Dict[str, torch.Tensor]:
    """Full ``metadata_sensor_fusion(bands, sites, classes).state_dict()``-shaped table."""
    t = init_meta_params(sites, classes, seed, fused=True)
    for k, v in orc.init_params("hang2020", bands, classes, seed, perturb_bn=True).items():
        t["sensor_model." + k] = v
    return t


def make_sites(batch: int, sites: int, seed: int) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64(seed + 15485863))
    return torch.from_numpy(rng.integers(0, sites, size=(batch,), dtype=np.int64))


def is_buffer(name: str) -> bool:
    return orc.is_buffer(name)


def metadata_forward(p, prefix: str, site: torch.Tensor, training: bool, keep_mask: Optional[torch.Tensor]):
    """metadata.forward, metadata.py:17-24.  ``keep_mask`` (B,16) bool: the dropout's Bernoulli(0.3) draw, shared with the
    implementation under test (train mode only; scale 1/(1-p) like F.dropout)."""
    x = F.embedding(site, p[prefix + "embedding.weight"])                                  # :18
    x = F.batch_norm(x, p[prefix + "batch_norm.running_mean"], p[prefix + "batch_norm.running_var"],
                     p[prefix + "batch_norm.weight"], p[prefix + "batch_norm.bias"], training=training,
                     momentum=orc.BN_MOMENTUM, eps=orc.BN_EPS)                             # :19
    if training:
        p[prefix + "batch_norm.num_batches_tracked"] += 1
        if keep_mask is None:
            raise ValueError("train-mode parity needs the dropout mask")
        x = x * keep_mask.to(x.dtype) / (1.0 - DROP_P)                                      # :20
    x = F.linear(x, p[prefix + "mlp.weight"], p[prefix + "mlp.bias"])                      # :21
    return F.relu(x)                                                                       # :22


def fusion_forward(p, images: torch.Tensor, site: torch.Tensor, training: bool, keep_mask=None, z_values=None):
    """metadata_sensor_fusion.forward, metadata.py:37-44."""
    meta = metadata_forward(p, "metadata_model.", site, training, keep_mask)               # :38
    sensor_table = {k[len("sensor_model."):]: v for k, v in p.items() if k.startswith("sensor_model.")}
    sensor, _ = orc.forward("hang2020", sensor_table, images, training, z_values)         # :39
    cat = torch.cat([meta, sensor.to(meta.dtype)], dim=1)                                  # :40
    return F.relu(F.linear(cat, p["fc1.weight"], p["fc1.bias"]))                           # :41-42


def fusion_step(table, images, site, y, training=True, keep_mask=None, z_values=None):
    """forward + CE + backward; returns (loss, out, grads, buffers)."""
    p = {k: (v.clone().requires_grad_(True) if not is_buffer(k) else v.clone()) for k, v in table.items()}
    out = fusion_forward(p, images, site, training, keep_mask, z_values)
    loss = F.cross_entropy(out, y)
    names = [k for k in p if not is_buffer(k)]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return loss.detach(), out.detach(), dict(zip(names, grads)), {k: v for k, v in p.items() if is_buffer(k)}


def metadata_step(table, site, y, training=True, keep_mask=None):
    p = {k: (v.clone().requires_grad_(True) if not is_buffer(k) else v.clone()) for k, v in table.items()}
    out = metadata_forward(p, "", site, training, keep_mask)
    loss = F.cross_entropy(out, y)
    names = [k for k in p if not is_buffer(k)]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return loss.detach(), out.detach(), dict(zip(names, grads)), {k: v for k, v in p.items() if is_buffer(k)}
