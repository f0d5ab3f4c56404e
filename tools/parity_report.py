"""Run every golden case (and a few oracle cases) through the CUDA path and print per-tensor
errors without stopping at the first mismatch.  Usage (GPU box): python tools/parity_report.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402
from oracle import hang2020_oracle as orc  # noqa: E402
from deeptreeattention_b200 import Hang2020 as H  # noqa: E402

CLS = {"hang2020": H.Hang2020, "spectral": H.spectral_network, "spatial": H.spatial_network, "vanilla": H.vanilla_CNN}


def run_cuda(kind, bands, classes, table, x, y, regime, training):
    m = CLS[kind](bands, classes)
    m.load_state_dict(table)
    m = m.cuda().train(training)
    xd, yd = x.cuda(), y.cuda()
    out = m(xd)
    if kind == "hang2020":
        heads, result = m.head_scores, out
    elif kind == "vanilla":
        heads, result = [out], out
    else:
        heads, result = out, out
    loss = orc.loss_regime(regime, result, heads, yd)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: (p.grad.detach().cpu() if p.grad is not None else None) for k, p in m.named_parameters()}
    bufs = {k: v.detach().cpu() for k, v in m.state_dict().items() if orc.is_buffer(k)}
    res = result[-1] if isinstance(result, list) else result
    return float(loss), res.detach().cpu(), [h.detach().cpu() for h in heads], grads, bufs


def report(name, loss, res, heads, grads, bufs, ref):
    rloss, rres, rheads, rgrads, rbufs = ref
    worst = 0.0
    print(f"== {name}: loss {loss:.6f} ref {rloss:.6f}")
    e = (res - rres).abs().max().item()
    flips = int((res.argmax(1) != rres.argmax(1)).sum())
    print(f"   result max|d| {e:.3e}  argmax flips {flips}/{res.shape[0]}")
    for i, (h, rh) in enumerate(zip(heads, rheads)):
        print(f"   head{i} max|d| {(h - rh).abs().max().item():.3e}")
    for k, rg in rgrads.items():
        g = grads.get(k)
        if rg is None:
            tag = "None" if g is None else f"max {g.abs().max().item():.2e} (ref None)"
            print(f"   grad {k}: {tag}")
            continue
        if g is None:
            print(f"   grad {k}: MISSING (ref max {rg.abs().max().item():.2e})")
            worst = max(worst, 1.0)
            continue
        err = (g.double() - rg.double()).abs().max().item()
        scale = rg.abs().max().item()
        rel = err / (scale + 1e-30)
        flag = "" if err <= 1e-3 * scale + 1e-5 else "   <<<<<<"
        worst = max(worst, rel if scale > 1e-5 else 0.0)
        print(f"   grad {k}: err {err:.3e} scale {scale:.3e} rel {rel:.2e}{flag}")
    for k, rb in rbufs.items():
        err = (bufs[k].double() - rb.double()).abs().max().item()
        flag = "" if err <= 1e-5 else "   <<<<<<"
        print(f"   buf {k}: err {err:.3e}{flag}")
    return worst


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    only = sys.argv[1:] or None
    for case in gu.cases():
        if only and case["name"] not in only:
            continue
        table, x, y = gu.build(case)
        ref_loss, ref_res, ref_heads, ref_grads, ref_bufs = orc.step(case["kind"], table, x, y, regime=case["regime"],
                                                                    training=case["training"])
        ref_res = ref_res[-1] if isinstance(ref_res, list) else ref_res
        try:
            out = run_cuda(case["kind"], case["bands"], case["classes"], table, x, y, case["regime"], case["training"])
        except Exception as exc:  # keep going: report every case
            print(f"== {case['name']}: EXCEPTION {type(exc).__name__}: {exc}")
            continue
        report(case["name"], *out, (float(ref_loss), ref_res.detach(), [h.detach() for h in ref_heads], ref_grads, ref_bufs))


if __name__ == "__main__":
    main()
