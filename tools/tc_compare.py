"""Compare conv_impl 1 (tcgen05 split-bf16) against conv_impl 0 (fp32 SIMT) and the CPU oracle on one
seeded case; prints where the largest conv1 weight-gradient difference sits.
Usage (GPU box): python tools/tc_compare.py kind bands classes batch dist regime"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deeptreeattention_b200 import Hang2020 as H, _capi  # noqa: E402
from oracle import hang2020_oracle as orc  # noqa: E402

CLS = {"hang2020": H.Hang2020, "spectral": H.spectral_network, "spatial": H.spatial_network, "vanilla": H.vanilla_CNN}


def run(kind, bands, classes, table, x, y, regime, impl):
    _capi.set_option(0, "conv_impl", impl)
    m = CLS[kind](bands, classes)
    m.load_state_dict(table)
    m = m.cuda().train()
    out = m(x.cuda())
    heads = m.head_scores if kind == "hang2020" else ([out] if kind == "vanilla" else out)
    loss = orc.loss_regime(regime, out, heads, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    return [h.detach().cpu() for h in heads], {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}


def main():
    kind, bands, classes, batch, dist, regime = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
    seed = 1000 + batch
    table = orc.init_params(kind, bands, classes, seed, perturb_bn=True)
    x, y = orc.make_inputs(batch, bands, classes, seed, dist)
    _, _, rheads, rgrads, _ = orc.step(kind, table, x, y, regime=regime, training=True)
    h0, g0 = run(kind, bands, classes, table, x, y, regime, 0)
    h1, g1 = run(kind, bands, classes, table, x, y, regime, 1)
    for i, (a, b, r) in enumerate(zip(h0, h1, rheads)):
        print(f"head{i}: |tc-simt| {float((a - b).abs().max()):.3e}  |simt-oracle| {float((a - r.detach()).abs().max()):.3e}  |tc-oracle| {float((b - r.detach()).abs().max()):.3e}")
    for k in g0:
        if not ("conv" in k and "weight" in k and "attention" not in k):
            print(f"  {k}: scale {float(rgrads[k].abs().max()):.3e} |tc-simt| {float((g0[k] - g1[k]).abs().max()):.3e} |simt-oracle| {float((g0[k] - rgrads[k]).abs().max()):.3e}")
            continue
        d01 = (g0[k] - g1[k]).abs()
        print(f"{k}: scale {float(rgrads[k].abs().max()):.3e} |tc-simt| {float(d01.max()):.3e} |simt-oracle| {float((g0[k] - rgrads[k]).abs().max()):.3e} "
              f"|tc-oracle| {float((g1[k] - rgrads[k]).abs().max()):.3e}")
        if "conv1" in k and g0[k].dim() == 4:
            idx = torch.nonzero(d01 == d01.max())[0].tolist()
            print("   worst at (co, ci, ky, kx) =", idx, "tc", float(g1[k][tuple(idx)]), "simt", float(g0[k][tuple(idx)]), "oracle", float(rgrads[k][tuple(idx)]))
            per_tap = d01.amax(dim=(0, 1))
            print("   per-tap max diff:", [f"{v:.2e}" for v in per_tap.flatten().tolist()])
            per_ci = d01.amax(dim=(0, 2, 3))
            top = torch.topk(per_ci, 5)
            print("   worst ci:", top.indices.tolist(), [f"{v:.2e}" for v in top.values.tolist()])


if __name__ == "__main__":
    main()
