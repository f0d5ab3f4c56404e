"""BASELINE config 5: ``metadata`` / ``metadata_sensor_fusion`` (reference src/models/metadata.py:9-44).

Fixtures (tests/golden/metadata_*.npz) were produced by the reference's own modules (tests/golden/make_metadata_golden.py),
at cfg 5's real shape (369 bands, 32 sites, 50 classes, 512 crops; train with the reference's dropout mask, and eval) and at
the shapes of the reference's tests/test_metadata.py.  CPU: the oracle restatement against them.  GPU: the CUDA path
(dta_metadata_forward / dta_metadata_backward + dta_forward / dta_backward through the drop-in modules) against them and
against the decision-matched float64 oracle for the sensor model's gradients."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_util as gu
from oracle import hang2020_oracle as orc
from oracle import metadata_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cases():
    with open(os.path.join(GOLDEN, "metadata_cases.json")) as f:
        meta = json.load(f)
    return [dict(zip(meta["fields"], c)) for c in meta["cases"]]


def build(case):
    if case["fused"]:
        table = mo.init_fusion_params(case["bands"], case["sites"], case["classes"], case["seed"])
    else:
        table = mo.init_meta_params(case["sites"], case["classes"], case["seed"], fused=False)
    site = mo.make_sites(case["batch"], case["sites"], case["seed"])
    x, y = orc.make_inputs(case["batch"], max(case["bands"], 1), case["classes"], case["seed"])
    gold = dict(np.load(os.path.join(GOLDEN, case["name"] + ".npz")))
    keep = torch.from_numpy(gold["keep_mask"]).bool() if "keep_mask" in gold else None
    return table, x, site, y, gold, keep


@pytest.mark.parametrize("case", cases(), ids=lambda c: c["name"])
def test_oracle_matches_reference_golden(case):
    torch.set_num_threads(8)
    table, x, site, y, gold, keep = build(case)
    if case["fused"]:
        loss, out, grads, bufs = mo.fusion_step(table, x, site, y, case["training"], keep)
    else:
        loss, out, grads, bufs = mo.metadata_step(table, site, y, case["training"], keep)
    np.testing.assert_allclose(out.numpy(), gold["out"], rtol=0, atol=2e-6)
    assert abs(float(loss) - float(gold["loss"])) < 2e-6
    seen = 0
    for key, ref in gold.items():
        if key.startswith("grad/"):
            g = grads[key[5:]].numpy()
            assert g.dtype == ref.dtype
            # absolute floor: with a single site BatchNorm1d sees identical rows, the true embedding / BN-weight gradient is 0
            # and the reference itself holds ~3e-6 of rounding noise there (istd = 1/sqrt(eps) = 316 amplifies it)
            np.testing.assert_allclose(g, ref, rtol=1e-4, atol=1e-5 + 1e-4 * np.abs(ref).max())
            seen += 1
        elif key.startswith("buf/"):
            np.testing.assert_allclose(bufs[key[4:]].numpy(), ref, rtol=1e-6, atol=1e-6)
    assert seen >= 5


def _cuda_module(case, table):
    from deeptreeattention_b200 import metadata as M
    m = M.metadata_sensor_fusion(case["bands"], case["sites"], case["classes"]) if case["fused"] else M.metadata(case["sites"], case["classes"])
    m.load_state_dict(table)
    return m.cuda().train(case["training"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", cases(), ids=lambda c: c["name"])
def test_cuda_matches_reference_golden(case):
    from deeptreeattention_b200 import _capi
    torch.set_num_threads(8)
    table, x, site, y, gold, keep = build(case)
    m = _cuda_module(case, table)
    big = case["fused"] and case["batch"] >= 256
    _capi.KEEP_SAVED = big
    try:
        if case["fused"]:
            out = m(x.cuda(), site.cuda(), keep_mask=keep)
            if big:
                torch.cuda.synchronize()
                zvals = gu.cuda_conv_outputs(m.sensor_model, "hang2020", case["batch"], case["bands"], case["classes"], case["training"])
                m.sensor_model.fused_spec().last_saved = None
        else:
            out = m(site.cuda(), keep_mask=keep)
    finally:
        _capi.KEEP_SAVED = False
    loss = F.cross_entropy(out, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    tol = 1e-3 if case["fused"] else 1e-5        # the fused output sits behind the sensor model's convolutions
    if case["sites"] == 1 and case["training"]:
        tol = max(tol, 1e-4)   # one site: BatchNorm1d sees identical rows, variance 0, invstd = 316 amplifies the rounding of (x - mean)
    np.testing.assert_allclose(out.detach().cpu().numpy(), gold["out"], rtol=0, atol=tol)
    assert abs(float(loss) - float(gold["loss"])) < tol
    sd = m.state_dict()
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}
    seen = 0
    for key, ref in gold.items():
        if key.startswith("buf/"):
            np.testing.assert_allclose(sd[key[4:]].cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
        elif key.startswith("grad/"):
            name = key[5:]
            g = grads[name].numpy()
            assert g.dtype == ref.dtype and g.shape == ref.shape, name
            scale = float(np.abs(ref).max())
            if scale < 1e-5:        # true gradient 0 (single site: identical rows under BatchNorm1d), reference holds rounding noise
                assert float(np.abs(g).max()) <= 1e-5, name
            else:
                assert gu.rel_l2(g, ref) <= 1e-3, f"{name}: rel-L2 {gu.rel_l2(g, ref):.3e}"
                assert float(np.abs(g - ref).max()) <= 1e-3 * scale + 1e-6, name
            seen += 1
    assert seen >= 5
    if big:
        # every gradient, the sensor model's included, against the float64 oracle continued from the CUDA convolution outputs
        z64 = {"sensor_model." + k: v for k, v in zvals.items()}
        t64 = gu.to_fp64(table)
        _, _, g64, _ = mo.fusion_step(t64, x.double(), site, y, case["training"], keep,
                                      z_values={k[len("sensor_model."):]: v.double() for k, v in z64.items()})
        for k, rg in g64.items():
            if rg is None:
                assert k not in grads or float(grads[k].abs().max()) == 0.0, k
                continue
            if case["training"] and k.endswith("conv_layer.bias"):
                assert float(grads[k].abs().max()) <= 1e-5, k
                continue
            if float(rg.abs().max()) == 0.0:
                assert float(grads[k].abs().max()) <= 1e-7, k
                continue
            assert gu.grad_error(k, grads[k], rg, g64) <= 1e-3, f"{k}: rel-L2 {gu.grad_error(k, grads[k], rg, g64):.3e}"


@pytest.mark.gpu
def test_in_kernel_dropout_stream():
    """Without a caller-supplied mask the kernel draws its own: ~30 % kept, survivors scaled by 1/0.3, repeatable under
    torch.manual_seed, different between calls, identity in eval mode."""
    from deeptreeattention_b200 import metadata as M
    torch.manual_seed(3)
    m = M.metadata(sites=8, classes=12).cuda().train()
    with torch.no_grad():
        m.mlp.weight.copy_(torch.eye(12, 16))          # mlp output i = relu(dropped feature i + bias)
        m.mlp.bias.fill_(10.0)
        m.batch_norm.bias.fill_(1.0)
    site = torch.randint(0, 8, (4096,)).cuda()
    torch.manual_seed(11)
    a = m(site)
    torch.manual_seed(11)
    b = m(site)
    c = m(site)
    assert torch.equal(a, b) and not torch.equal(a, c)
    kept = (a != 10.0).float().mean().item()           # a dropped feature leaves exactly the bias
    assert 0.28 < kept < 0.32, kept
    m.eval()
    e1, e2 = m(site), m(site)
    assert torch.equal(e1, e2)


@pytest.mark.gpu
def test_reference_test_shapes_and_errors():
    """tests/test_metadata.py:11-23 of the reference: (20,) site ids -> (20, 10) for both modules; CPU tensors raise."""
    from deeptreeattention_b200 import metadata as M
    m = M.metadata(sites=1, classes=10).cuda()
    assert tuple(m(torch.zeros(20, dtype=torch.int64).cuda()).shape) == (20, 10)
    f = M.metadata_sensor_fusion(bands=3, sites=1, classes=10).cuda()
    out = f(torch.randn(20, 3, 11, 11).cuda(), torch.zeros(20, dtype=torch.int64).cuda())
    assert tuple(out.shape) == (20, 10)
    with pytest.raises(RuntimeError):
        m(torch.zeros(20, dtype=torch.int64))
    with pytest.raises(ValueError):
        f(torch.randn(20, 3, 11, 11).cuda(), torch.zeros(19, dtype=torch.int64).cuda())


def test_state_dict_contract_matches_reference_shapes():
    """Keys / shapes / dtypes of metadata_sensor_fusion(bands, sites, classes).state_dict()."""
    from deeptreeattention_b200 import metadata as M
    m = M.metadata_sensor_fusion(bands=12, sites=7, classes=5)
    sd = m.state_dict()
    table = mo.init_fusion_params(12, 7, 5, 0)
    assert set(sd.keys()) == set(table.keys())
    for k, v in table.items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
    assert list(sd.keys())[:9] == ["metadata_model.embedding.weight", "metadata_model.batch_norm.weight", "metadata_model.batch_norm.bias",
                                   "metadata_model.batch_norm.running_mean", "metadata_model.batch_norm.running_var",
                                   "metadata_model.batch_norm.num_batches_tracked", "metadata_model.mlp.weight", "metadata_model.mlp.bias",
                                   "sensor_model.alpha"]
