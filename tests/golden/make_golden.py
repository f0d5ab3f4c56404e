"""Generate golden vectors by executing the REFERENCE module (build container only).

Run here (where /root/reference exists):   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests only read the committed ``*.npz`` files.

For every case the reference ``src/models/Hang2020.py`` module is instantiated, loaded
with the seeded parameter table from ``oracle.hang2020_oracle.init_params`` (numpy
PCG64, independent of the torch RNG), run on ``make_inputs`` crops, and its outputs,
loss, gradients and BatchNorm buffers are stored.  Large gradient tensors are stored
as a fixed strided sample plus sum / abs-sum / l2 so the fixtures stay small.

Kink screening.  The network is piecewise smooth (ReLU, max-pool argmax): a pre-activation
that lies within fp32 rounding noise of a kink can fall on either side under a different
summation order, which changes individual gradient elements by O(1e-2) relative -- in the
reference itself too (CPU oneDNN vs cuDNN).  Element-wise gradient parity at 1e-3 is only
defined away from kinks, so every case's seed is advanced until the reference's gradients are
stable (5e-4 relative) under three random 1e-5 relative perturbations of crops and weights
(above the 2^-18 = 3.8e-6 operand rounding of the split-bf16 tensor-core convolutions, and far above
the noise of an fp32 evaluation with a different summation order).  The accepted seed
is recorded in cases.json (the same stability is required between the exact evaluation and one whose
convolution values are computed in the tensor-core kernels' split-bf16 arithmetic, oracle CONV_ARITH) and the measured per-tensor sensitivity is stored next to each
gradient (``grad/<name>/sens``) so tests can widen the tolerance by the reference's own
conditioning instead of guessing.

Live-gate screening.  A spatial attention block whose 1x1 channel pool is negative for every
pixel of every crop sits behind a dead ReLU (``Hang2020.py:108-109``): its ``channel_pool`` /
stencil gradients are identically zero and the block's backward is never exercised.  With the
default initialisation that happens for about half of all (seed, block) pairs, so seeds are also
advanced until every ``channel_pool.weight`` gradient of the case is non-zero.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import hang2020_oracle as orc  # noqa: E402

REF = "/root/reference/src/models/Hang2020.py"
SAMPLE = 2048

# name, kind, bands, classes, batch, dist, regime, training, perturb_bn, seed
CASES = [
    ("cfg1_vanilla_b3_c2_train", "vanilla", 3, 2, 4, "uniform", "R1", True, False, 1),
    ("cfg1_vanilla_b3_c2_eval", "vanilla", 3, 2, 4, "uniform", "R1", False, True, 2),
    ("hang_b3_c10_randn_train_R1", "hang2020", 3, 10, 6, "normal", "R1", True, False, 3),
    ("hang_b3_c10_randn_train_R2", "hang2020", 3, 10, 5, "normal", "R2", True, True, 4),
    ("hang_b369_c50_train_R1", "hang2020", 369, 50, 8, "uniform", "R1", True, False, 5),
    ("hang_b369_c50_train_R2", "hang2020", 369, 50, 8, "uniform", "R2", True, True, 6),
    ("hang_b369_c50_eval_R2", "hang2020", 369, 50, 8, "normal", "R2", False, True, 7),
    ("spectral_b369_c20_train_R2", "spectral", 369, 20, 8, "uniform", "R2", True, False, 8),
    ("spectral_b349_c10_eval_R1", "spectral", 349, 10, 5, "uniform", "R1", False, True, 9),
    ("spatial_b349_c10_train_R2", "spatial", 349, 10, 6, "normal", "R2", True, True, 10),
    ("vanilla_b369_c10_train", "vanilla", 369, 10, 7, "normal", "R1", True, True, 11),
]


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_Hang2020", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sample_index(n):
    if n <= SAMPLE:
        return np.arange(n)
    return np.unique(np.linspace(0, n - 1, SAMPLE).astype(np.int64))


def kink_stable(ref, case, seed, eps=1e-5, tol=5e-4, draws=3):
    """(stable, {name: sensitivity}): whether the reference's gradients move smoothly under
    tiny perturbations, and by how much each tensor moved."""
    name, kind, bands, classes, batch, dist, regime, training, perturb, _ = case
    table = orc.init_params(kind, bands, classes, seed, perturb_bn=perturb)
    x, y = orc.make_inputs(batch, bands, classes, seed, dist)
    base = orc.step(kind, table, x, y, regime=regime, training=training)[3]
    for k, g in base.items():      # live-gate screening: every spatial gate must pass gradient
        if k.endswith("channel_pool.weight") and g is not None and float(g.abs().max()) == 0.0:
            return False, {}
    gen = torch.Generator().manual_seed(seed)
    sens = {k: 0.0 for k, g in base.items() if g is not None}
    stable = True
    # the tensor-core kernels' own arithmetic (split-bf16 operands, three products) must not move a decision either
    orc.CONV_ARITH = "split_bf16x3"
    try:
        g_tc = orc.step(kind, table, x, y, regime=regime, training=training)[3]
    finally:
        orc.CONV_ARITH = None
    for k, g in base.items():
        if g is not None and float((g - g_tc[k]).abs().max()) > tol * float(g.abs().max()) + 2e-6:
            return False, {}
    for _ in range(draws):
        t2 = {k: (v * (1 + eps * torch.randn(v.shape, generator=gen, dtype=v.dtype))
                  if (v.is_floating_point() and not orc.is_buffer(k)) else v.clone()) for k, v in table.items()}
        x2 = x * (1 + eps * torch.randn(x.shape, generator=gen))
        g2 = orc.step(kind, t2, x2, y, regime=regime, training=training)[3]
        for k, g in base.items():
            if g is None:
                continue
            scale = float(g.abs().max())
            d = float((g - g2[k]).abs().max())
            sens[k] = max(sens[k], d)
            if d > tol * scale + 2e-6:
                stable = False
        if not stable:
            break
    return stable, sens


def run_case(ref, case):
    name, kind, bands, classes, batch, dist, regime, training, perturb, seed = case
    tries = 0
    while True:
        stable, sens = kink_stable(ref, case, seed)
        if stable:
            break
        seed += 1000
        tries += 1
        if tries > 400:
            raise RuntimeError(f"no kink-stable seed found for {name}")
    case = case[:-1] + (seed,)
    table = orc.init_params(kind, bands, classes, seed, perturb_bn=perturb)
    x, y = orc.make_inputs(batch, bands, classes, seed, dist)
    cls = {"hang2020": ref.Hang2020, "spectral": ref.spectral_network,
           "spatial": ref.spatial_network, "vanilla": ref.vanilla_CNN}[kind]
    m = cls(bands, classes)
    m.load_state_dict(table, strict=True)
    m.train(training)
    if kind == "hang2020":
        # every head, via the sub-networks on the reference model (SURVEY 0.3 regime R2)
        spec = m.spectral_network(x)
        spat = m.spatial_network(x)
        w = torch.sigmoid(m.alpha)
        heads = spec + spat
        result = spec[-1] * w + spat[-1] * (1 - w)
        if regime == "R1":
            # the reference regime must go through Hang2020.forward itself; redo on a
            # fresh copy so BN buffers are stepped exactly once
            m = cls(bands, classes)
            m.load_state_dict(table, strict=True)
            m.train(training)
            result = m(x)
            with torch.no_grad():
                m2 = cls(bands, classes)
                m2.load_state_dict(table, strict=True)
                m2.train(training)
                heads = m2.spectral_network(x) + m2.spatial_network(x)
    elif kind == "vanilla":
        result = m(x)
        heads = [result]
    else:
        heads = m(x)
        result = heads
    loss = orc.loss_regime(regime, result, heads, y)
    loss.backward()
    out = {"loss": loss.detach().numpy()}
    res = result[-1] if isinstance(result, list) else result
    out["result"] = res.detach().numpy()
    for i, h in enumerate(heads):
        out[f"head{i}"] = h.detach().numpy()
    sd = m.state_dict()
    for k, v in sd.items():
        if orc.is_buffer(k):
            out[f"buf/{k}"] = v.detach().numpy()
    for k, prm in m.named_parameters():
        if prm.grad is None:
            out[f"gradnone/{k}"] = np.array(1, dtype=np.int8)
            continue
        g = prm.grad.detach().numpy().reshape(-1)
        idx = sample_index(g.size)
        out[f"grad/{k}/sample"] = g[idx]
        g64 = g.astype(np.float64)
        out[f"grad/{k}/stats"] = np.array([g64.sum(), np.abs(g64).sum(), np.sqrt((g64 * g64).sum())])
        out[f"grad/{k}/sens"] = np.array(sens[k], dtype=np.float64)
    return name, out, case


def main():
    torch.set_num_threads(8)
    ref = load_reference()
    meta = []
    for case in CASES:
        name, out, case = run_case(ref, case)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        meta.append(case)
        print(name, "seed", case[-1], "loss", float(out["loss"]), "keys", len(out), flush=True)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        import json
        json.dump({"torch": torch.__version__, "reference": "weecology/DeepTreeAttention@cae13f1",
                   "fields": ["name", "kind", "bands", "classes", "batch", "dist", "regime",
                              "training", "perturb_bn", "seed"],
                   "cases": meta}, f, indent=1)


if __name__ == "__main__":
    main()
