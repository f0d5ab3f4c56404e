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


def check_grads(gold, grads, rtol=1e-3, atol=2e-6, where="", sens_factor=4.0):
    """grads: {name: tensor or None}.  Weight grads: |d| <= atol + rtol*max|ref| + 4*sens per
    tensor.  (Conv biases under train-mode BN have true gradient 0 and the reference itself
    holds ~1e-7 noise there, SURVEY Appendix C, so the absolute term matters; ``sens`` is the
    reference's own measured movement under 3e-6 relative input/weight perturbations, see
    tests/golden/make_golden.py -- small for the screened, kink-stable cases.)"""
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
            sens = float(gold[f"grad/{name}/sens"]) if f"grad/{name}/sens" in gold else 0.0
            err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max())
            assert err <= atol + rtol * scale + sens_factor * sens, \
                f"{where}{name}: grad err {err:.3e} scale {scale:.3e} sens {sens:.3e}"
            stats = gold[f"grad/{name}/stats"]
            g64 = g.astype(np.float64)
            l2 = np.sqrt((g64 * g64).sum())
            assert abs(l2 - stats[2]) <= atol * np.sqrt(g.size) + rtol * stats[2] + sens_factor * sens * np.sqrt(g.size), \
                f"{where}{name}: l2 {l2:.6e} vs {stats[2]:.6e}"
            seen += 1
    assert seen > 0


# ----------------------------------------------------------------------------- kink-robust gradient metric
def rel_l2(a, b) -> float:
    """||a - b||_2 / ||b||_2 over a whole tensor, in float64."""
    a = torch.as_tensor(a).double().reshape(-1)
    b = torch.as_tensor(b).double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def grad_error(name, g, ref, all_refs) -> float:
    """Relative L2 error of one gradient tensor -- except for a ONE-element bias gradient, which is a batch sum with
    cancellation (|sum| can be 1e-4 of the sum of |terms|): its ABSOLUTE error is held to the scale of the sibling weight
    gradient, which sums the same terms."""
    ref = torch.as_tensor(ref)
    g = torch.as_tensor(g)
    sib = name[:-4] + "weight"
    if ref.numel() == 1 and name.endswith(".bias") and all_refs.get(sib) is not None:
        den = max(float(ref.double().abs().max()), float(torch.as_tensor(all_refs[sib]).double().norm()))
        return float((g.double().reshape(-1) - ref.double().reshape(-1)).abs().max() / max(den, 1e-300))
    return rel_l2(g, ref)


def to_fp64(table):
    return {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in table.items()}


def oracle_step_fp64(kind, table, x, y, regime, training):
    """The oracle evaluated in float64 ("truth" of SURVEY.md 8c: decides which side of a tolerance miss is wrong)."""
    return orc.step(kind, to_fp64(table), x.double(), y, regime=regime, training=training)


def l2_conditioning(kind, table, x, y, regime, training, base_grads64, eps=1e-5, draws=2, seed=0):
    """Per-tensor relative-L2 movement of the FLOAT64 oracle's gradients when crops and weights carry ``eps``-relative
    noise -- the size of the split-bf16 rounding of the tensor-core convolutions (and, at eps ~ 1e-6, of any fp32
    evaluation with another summation order).  The network is piecewise smooth: a ReLU / max-pool decision whose
    pre-activation lies inside the noise flips, and each flip switches one element of a conv-output gradient on or off.
    In L2 over a weight-gradient tensor n flips out of N active elements cost ~sqrt(n / N) relative, which is what this
    measures; float64 arithmetic keeps rounding of the oracle itself out of the number."""
    gen = torch.Generator().manual_seed(seed)
    t64, x64 = to_fp64(table), x.double()
    cond = {k: 0.0 for k, g in base_grads64.items() if g is not None}
    for _ in range(draws):
        t2 = {k: (v * (1 + eps * torch.randn(v.shape, generator=gen, dtype=v.dtype))
                  if (v.is_floating_point() and not orc.is_buffer(k)) else v.clone()) for k, v in t64.items()}
        x2 = x64 * (1 + eps * torch.randn(x64.shape, generator=gen, dtype=torch.float64))
        g2 = orc.step(kind, t2, x2, y, regime=regime, training=training)[3]
        for k in cond:
            cond[k] = max(cond[k], rel_l2(g2[k], base_grads64[k]))
    return cond


# ----------------------------------------------------------------------------- decision-matched oracle
_KIND_ID = {"hang2020": 0, "spectral": 1, "spatial": 2, "vanilla": 3}


def cuda_conv_outputs(model, kind, batch, bands, classes, training):
    """``{oracle block prefix: z, "bn:" + prefix: a}``: the three convolution outputs z (pre-BatchNorm, bias included) the CUDA
    forward just left in its ``saved`` buffer and the post-BatchNorm values a = fmaf(z, scale, shift) every kernel derives from
    them (scale / shift read from the same buffer; the fused multiply-add is reproduced exactly: float32 product and sum are
    exact in float64, then ONE rounding).  ``dta_saved_region``; needs ``_capi.KEEP_SAVED``.  Split per branch, on the CPU."""
    from deeptreeattention_b200 import _capi
    saved = model.fused_spec().last_saved
    assert saved is not None, "set _capi.KEEP_SAVED before the forward"
    out = {}
    nb = 2 if kind == "hang2020" else 1

    def region(blk, which):
        off, n = _capi.saved_region(_KIND_ID[kind], batch, bands, classes, training, blk, which)
        return saved[off:off + 4 * n].view(torch.float32).cpu()

    for blk, (c, s) in enumerate(((32, 11), (64, 11), (128, 5))):
        z = region(blk, 0).view(batch, nb * c, s, s)
        scale, shift = region(blk, 1).double().view(1, -1, 1, 1), region(blk, 2).double().view(1, -1, 1, 1)
        a = (z.double() * scale + shift).float()
        names = [f"spectral_network.conv{blk + 1}", f"spatial_network.conv{blk + 1}"] if kind == "hang2020" else [f"conv{blk + 1}"]
        for g, name in enumerate(names):
            out[name] = z[:, g * c:(g + 1) * c].contiguous()
            out["bn:" + name] = a[:, g * c:(g + 1) * c].contiguous()
    return out


def matched_oracle_step_fp64(kind, table, x, y, regime, training, z_values):
    """The float64 oracle continued from the convolution outputs (and their BatchNorm images) of the implementation under test (value substitution only;
    every derivative is the oracle's own).  A Hang2020 network is piecewise linear in its activations: its gradient is only
    comparable between two evaluations that sit on the same linear piece, i.e. take the same ReLU / max-pool decisions.
    Forward parity (scores against the UNMATCHED oracle) bounds how far the substituted values are from the oracle's own."""
    return orc.step(kind, to_fp64(table), x.double(), y, regime=regime, training=training,
                    z_values={k: v.double() for k, v in z_values.items()})
