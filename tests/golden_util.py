"""Shared helpers for reading the committed golden fixtures (tests/golden/*.npz)."""
import json
import os

import numpy as np
import torch

from oracle import hang2020_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = 2048


def cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        meta = json.load(f)
    return [dict(zip(meta["fields"], c)) for c in meta["cases"]]


def load(case):
    return dict(np.load(os.path.join(GOLDEN, case["name"] + ".npz")))


def build(case):
    table = orc.init_params(case["kind"], case["bands"], case["classes"], case["seed"],
                            perturb_bn=case["perturb_bn"])
    x, y = orc.make_inputs(case["batch"], case["bands"], case["classes"], case["seed"], case["dist"])
    return table, x, y


def sample_index(n):
    if n <= SAMPLE:
        return np.arange(n)
    return np.unique(np.linspace(0, n - 1, SAMPLE).astype(np.int64))


def check_grads(gold, grads, rtol=1e-3, atol=2e-6, where=""):
    """grads: {name: tensor or None}.  Weight grads: |d| <= atol + rtol*max|ref| per tensor
    (conv biases under train-mode BN have true gradient 0 and the reference itself holds
    ~1e-7 noise there, SURVEY Appendix C, so the absolute term matters)."""
    seen = 0
    for key in gold:
        if key.startswith("gradnone/"):
            name = key[len("gradnone/"):]
            g = grads.get(name)
            assert g is None or float(torch.as_tensor(g).abs().max()) == 0.0, f"{where}{name}: expected no grad"
        elif key.endswith("/sample"):
            name = key[len("grad/"):-len("/sample")]
            g = grads[name]
            assert g is not None, f"{where}{name}: missing grad"
            g = g.detach().cpu().numpy().reshape(-1)
            ref = gold[key]
            got = g[sample_index(g.size)]
            assert got.dtype == ref.dtype, f"{where}{name}: dtype {got.dtype} vs {ref.dtype}"
            scale = float(np.abs(ref).max())
            err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max())
            assert err <= atol + rtol * scale, f"{where}{name}: grad err {err:.3e} scale {scale:.3e}"
            stats = gold[f"grad/{name}/stats"]
            g64 = g.astype(np.float64)
            l2 = np.sqrt((g64 * g64).sum())
            assert abs(l2 - stats[2]) <= atol * np.sqrt(g.size) + rtol * stats[2], \
                f"{where}{name}: l2 {l2:.6e} vs {stats[2]:.6e}"
            seen += 1
    assert seen > 0
