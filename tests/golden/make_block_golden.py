"""Golden vectors of the STAND-ALONE blocks, produced by executing the REFERENCE modules (build container only).

Run here (where /root/reference exists):   python tests/golden/make_block_golden.py
Writes tests/golden/blocks.npz + blocks.json.  For every case of ``oracle.blocks_oracle.CASES`` the reference
module (/root/reference/src/models/Hang2020.py: conv_module, spectral_attention, spatial_attention, Classifier,
global_spectral_pool) is built, loaded with the seeded parameters of ``blocks_oracle.build_case``, run on the seeded
input in the case's mode, and back-propagated from L = sum_k <out_k, G_k> with seeded G_k.  Outputs, parameter and
input gradients and BatchNorm buffers are stored (large tensors as a strided sample + l2 norm).  Seeds are advanced until
the reference's gradients are stable under 3e-6 relative input noise (ReLU / max-pool kinks, see make_golden.py).
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import blocks_oracle as bo  # noqa: E402
import golden_util as gu  # noqa: E402

REF = "/root/reference/src/models/Hang2020.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_Hang2020", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference(ref, case, seed, noise=0.0):
    p, x, rng = bo.build_case(case, seed)
    if noise:
        gen = torch.Generator().manual_seed(seed)
        x = x * (1 + noise * torch.randn(x.shape, generator=gen))
    x.requires_grad_(True)
    if case["block"] == "global_spectral_pool":
        outs = (ref.global_spectral_pool(x),)
        m = None
    else:
        m = getattr(ref, case["block"])(**case["args"])
        m.load_state_dict(p, strict=True)
        m.train(case["training"])
        if case["block"] == "conv_module":
            outs = (m(x, pool=case["pool"]),)
        elif case["block"] == "Classifier":
            outs = (m(x),)
        else:
            outs = m(x)
    gs = bo.upstream(rng, outs)
    sum((o * g).sum() for o, g in zip(outs, gs)).backward()
    grads = {k: v.grad.detach() for k, v in m.named_parameters()} if m is not None else {}
    grads["x"] = x.grad.detach()
    bufs = {k: v.detach() for k, v in m.state_dict().items() if "running" in k or k.endswith("num_batches_tracked")} if m is not None else {}
    return [o.detach() for o in outs], grads, bufs


def main():
    torch.set_num_threads(8)
    ref = load_reference()
    out, meta = {}, {}
    for case in bo.CASES:
        seed = 100 + len(meta)
        for _ in range(200):
            base = run_reference(ref, case, seed)[1]
            top = max(float(g.abs().max()) for g in base.values())   # conv biases under batch statistics have true gradient 0
            stable = True
            for d in range(2):
                pert = run_reference(ref, case, seed, noise=3e-6 * (d + 1))[1]
                for k, g in base.items():
                    if float((g - pert[k]).abs().max()) > 5e-4 * float(g.abs().max()) + 5e-7 * top + 2e-6:
                        stable = False
            if stable:
                break
            seed += 1000
        else:
            raise RuntimeError("no kink-stable seed for " + case["name"])
        outs, grads, bufs = run_reference(ref, case, seed)
        # the oracle restatement must agree with the reference before the fixture is written
        o_outs, o_grads, o_bufs = bo.step(case, seed)
        for a, b in zip(outs, o_outs):
            assert float((a - b).abs().max()) <= 2e-6, case["name"]
        n = case["name"]
        for i, o in enumerate(outs):
            out[f"{n}/out{i}"] = o.numpy()
        for k, g in grads.items():
            flat = g.numpy().reshape(-1)
            out[f"{n}/grad/{k}/sample"] = flat[gu.sample_index(flat.size)]
            out[f"{n}/grad/{k}/l2"] = np.array(np.sqrt((flat.astype(np.float64) ** 2).sum()))
        for k, b in bufs.items():
            out[f"{n}/buf/{k}"] = b.numpy()
        meta[n] = seed
        print(n, "seed", seed, "outs", [tuple(o.shape) for o in outs], flush=True)
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), **out)
    with open(os.path.join(HERE, "blocks.json"), "w") as f:
        json.dump({"torch": torch.__version__, "reference": "weecology/DeepTreeAttention@cae13f1", "seeds": meta}, f, indent=1)


if __name__ == "__main__":
    main()
