"""Golden vectors for BASELINE config 5 from the REFERENCE's ``src/models/metadata.py`` (build container only).

    python tests/golden/make_metadata_golden.py

The reference module is loaded with ``src.main`` stubbed (oracle/ref_loader.py), given the seeded parameter
table of ``oracle.metadata_oracle`` and run on seeded crops / site ids / labels.  In train mode the Bernoulli(0.3) mask
that ``nn.Dropout`` drew is recovered with a forward hook (output != 0) and stored in the fixture, so that every other
implementation can be evaluated on the SAME mask.  Stored: output scores, CE loss, BatchNorm1d buffers, every gradient
of the metadata branch and the fusion layer (complete; they are small) and alpha's.  The sensor model's other gradients
are covered by the Hang2020 fixtures and, at this batch size, by the decision-matched oracle (tests/test_gpu_parity.py).

Kink screening: a ReLU input of the fusion layer within 3e-5 of zero could flip under 1e-5-accurate sensor scores (the site
MLP's inputs are exact: 1e-6 there); seeds are advanced until none is.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import hang2020_oracle as orc  # noqa: E402
from oracle import metadata_oracle as mo  # noqa: E402
from oracle import ref_loader  # noqa: E402

# name, fused, bands, sites, classes, batch, training, seed
CASES = [
    ("metadata_fusion_cfg5_b369_s32_c50_B512_train", True, 369, 32, 50, 512, True, 51),
    ("metadata_fusion_cfg5_b369_s32_c50_B512_eval", True, 369, 32, 50, 512, False, 52),
    ("metadata_fusion_b3_s1_c10_B20_train", True, 3, 1, 10, 20, True, 53),       # shapes of tests/test_metadata.py:17-23
    ("metadata_fusion_b30_s5_c10_B20_eval", True, 30, 5, 10, 20, False, 54),
    ("metadata_s1_c10_B20_train", False, 0, 1, 10, 20, True, 55),               # shapes of tests/test_metadata.py:11-15
    ("metadata_s32_c50_B512_train", False, 0, 32, 50, 512, True, 56),
    ("metadata_s32_c50_B64_eval", False, 0, 32, 50, 64, False, 57),
]


def run(ref, case):
    name, fused, bands, sites, classes, batch, training, seed = case
    for attempt in range(400):
        table = mo.init_fusion_params(bands, sites, classes, seed) if fused else mo.init_meta_params(sites, classes, seed, fused=False)
        site = mo.make_sites(batch, sites, seed)
        x, y = orc.make_inputs(batch, max(bands, 1), classes, seed)
        m = ref.metadata_sensor_fusion(bands, sites, classes) if fused else ref.metadata(sites, classes)
        m.load_state_dict(table, strict=True)
        m.train(training)
        drop = m.metadata_model.dropout if fused else m.dropout
        mlp = m.metadata_model.mlp if fused else m.mlp
        seen = {}
        h1 = drop.register_forward_hook(lambda mod, inp, out: seen.__setitem__("keep", (out != 0) | (inp[0] == 0)))
        h2 = mlp.register_forward_hook(lambda mod, inp, out: seen.__setitem__("pre_mlp", out.detach()))
        h3 = m.fc1.register_forward_hook(lambda mod, inp, out: seen.__setitem__("pre_fc", out.detach())) if fused else None
        torch.manual_seed(seed)
        out = m(x, site) if fused else m(site)
        h1.remove(); h2.remove()
        if h3:
            h3.remove()
        margin_mlp = float(seen["pre_mlp"].abs().min())
        margin = float(seen["pre_fc"].abs().min()) if fused else margin_mlp
        if margin_mlp > 1e-6 and (not fused or margin > 3e-5):
            break
        seed += 1000
    else:
        raise RuntimeError(f"no kink-free seed for {name}")
    loss = F.cross_entropy(out, y)
    loss.backward()
    rec = {"out": out.detach().numpy(), "loss": loss.detach().numpy(), "relu_margin": np.array(margin)}
    if training:
        rec["keep_mask"] = seen["keep"].numpy().astype(np.uint8)
    for k, v in m.state_dict().items():
        if orc.is_buffer(k) and not k.startswith("sensor_model."):
            rec[f"buf/{k}"] = v.detach().numpy()
    for k, p in m.named_parameters():
        if k.startswith("sensor_model.") and k != "sensor_model.alpha":
            continue
        rec[f"grad/{k}"] = p.grad.detach().numpy()
    return name, rec, case[:-1] + (seed,)


def main():
    torch.set_num_threads(8)
    ref = ref_loader.load("metadata")
    assert ref is not None, "needs /root/reference (or oracle/_ref)"
    meta = []
    for case in CASES:
        name, rec, case = run(ref, case)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        meta.append(case)
        print(name, "seed", case[-1], "loss", float(rec["loss"]), "margin", float(rec["relu_margin"]), flush=True)
    with open(os.path.join(HERE, "metadata_cases.json"), "w") as f:
        json.dump({"torch": torch.__version__, "reference": "weecology/DeepTreeAttention@cae13f1 src/models/metadata.py",
                   "fields": ["name", "fused", "bands", "sites", "classes", "batch", "training", "seed"], "cases": meta}, f, indent=1)


if __name__ == "__main__":
    main()
