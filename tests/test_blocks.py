"""Stand-alone building blocks (conv_module, spectral_attention, spatial_attention, Classifier, global_spectral_pool):
the oracle against the golden vectors produced by the reference modules (CPU), and the CUDA kernels behind the drop-in
modules against the same vectors and the reference's own shape tests (GPU, /root/reference/tests/test_Hang2020.py:8-33).
Tolerances: outputs 1e-4 abs (exact-fp32 kernels, different summation order only); gradients 1e-3 of the tensor's max
+ 1e-6 of the block's largest gradient (conv biases under batch statistics have true gradient 0) + 1e-5."""
import json
import os

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import blocks_oracle as bo

GOLD = dict(np.load(os.path.join(gu.GOLDEN, "blocks.npz")))
with open(os.path.join(gu.GOLDEN, "blocks.json")) as f:
    SEEDS = json.load(f)["seeds"]


def check_against_golden(name, outs, grads, bufs, out_tol):
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.detach().cpu().numpy(), GOLD[f"{name}/out{i}"], rtol=0, atol=out_tol)
    keys = [k[len(name) + 6:-7] for k in GOLD if k.startswith(name + "/grad/") and k.endswith("/sample")]
    top = max(float(np.abs(GOLD[f"{name}/grad/{k}/sample"]).max()) for k in keys)
    assert keys
    for k in keys:
        g = grads[k].detach().cpu().numpy().reshape(-1)
        ref = GOLD[f"{name}/grad/{k}/sample"]
        got = g[gu.sample_index(g.size)]
        assert got.dtype == ref.dtype, k
        tol = 1e-3 * float(np.abs(ref).max()) + 1e-6 * top + 1e-5
        err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max())
        assert err <= tol, f"{name} {k}: err {err:.3e} tol {tol:.3e}"
        l2 = float(np.sqrt((g.astype(np.float64) ** 2).sum()))
        assert abs(l2 - float(GOLD[f"{name}/grad/{k}/l2"])) <= tol * np.sqrt(g.size), f"{name} {k}: l2"
    for k, b in bufs.items():
        np.testing.assert_allclose(b.detach().cpu().numpy(), GOLD[f"{name}/buf/{k}"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", bo.CASES, ids=lambda c: c["name"])
def test_block_oracle_matches_reference_golden(case):
    outs, grads, bufs = bo.step(case, SEEDS[case["name"]])
    check_against_golden(case["name"], outs, grads, bufs, out_tol=2e-6)


def test_blocks_refuse_cpu_tensors():
    from deeptreeattention_b200 import Hang2020 as H
    with pytest.raises(RuntimeError):
        H.conv_module(in_channels=3, filters=4)(torch.randn(2, 3, 11, 11))
    with pytest.raises(RuntimeError):
        H.spectral_attention(filters=32)(torch.randn(2, 32, 11, 11))
    with pytest.raises(RuntimeError):
        H.spatial_attention(filters=64)(torch.randn(2, 64, 5, 5))
    with pytest.raises(RuntimeError):
        H.Classifier(in_features=8, classes=3)(torch.randn(2, 8))
    with pytest.raises(RuntimeError):
        H.global_spectral_pool(torch.ones(2, 3, 4, 4))


def run_cuda_block(case, seed):
    from deeptreeattention_b200 import Hang2020 as H
    p, x, rng = bo.build_case(case, seed)
    x = x.cuda().requires_grad_(True)
    if case["block"] == "global_spectral_pool":
        m = None
        outs = (H.global_spectral_pool(x),)
    else:
        m = getattr(H, case["block"])(**case["args"])
        m.load_state_dict(p, strict=True)
        m = m.cuda().train(case["training"])
        if case["block"] == "conv_module":
            outs = (m(x, pool=case["pool"]),)
        elif case["block"] == "Classifier":
            outs = (m(x),)
        else:
            outs = tuple(m(x))
    gs = [g.cuda() for g in bo.upstream(rng, [o.detach().cpu() for o in outs])]
    sum((o * g).sum() for o, g in zip(outs, gs)).backward()
    torch.cuda.synchronize()
    grads = {k: v.grad for k, v in m.named_parameters()} if m is not None else {}
    grads["x"] = x.grad
    bufs = {k: v for k, v in m.state_dict().items() if "running" in k or k.endswith("num_batches_tracked")} if m is not None else {}
    return outs, grads, bufs


@pytest.mark.gpu
@pytest.mark.parametrize("case", bo.CASES, ids=lambda c: c["name"])
def test_cuda_block_matches_reference_golden(case):
    outs, grads, bufs = run_cuda_block(case, SEEDS[case["name"]])
    check_against_golden(case["name"], outs, grads, bufs, out_tol=1e-4)


@pytest.mark.gpu
def test_cuda_spectral_dead_taps_exact_zero_and_partial_upstream():
    """Only one of the two outputs used downstream (the other gradient is None), dead Conv1d taps exactly 0."""
    from deeptreeattention_b200 import Hang2020 as H
    case = [c for c in bo.CASES if c["name"] == "spectral_64"][0]
    p, x, _ = bo.build_case(case, 5)
    m = H.spectral_attention(filters=64)
    m.load_state_dict(p)
    m = m.cuda()
    out, feat = m(x.cuda())
    feat.sum().backward()
    g = m.attention_conv1.weight.grad
    assert float(g[:, :, [0, 1, 3, 4]].abs().max()) == 0.0 and float(g[:, :, 2].abs().max()) > 0.0
    ro, rf = bo.forward(case, {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in p.items()}, x)
    np.testing.assert_allclose(feat.detach().cpu().numpy(), rf.detach().numpy(), rtol=0, atol=1e-5)


# ---- the reference's own tests for these modules, run against the drop-in namespace (tests/test_Hang2020.py:8-33) ----
@pytest.mark.gpu
def test_conv_module():
    from deeptreeattention_b200 import Hang2020
    m = Hang2020.conv_module(in_channels=369, filters=32).cuda()
    image = torch.randn(20, 369, 11, 11).cuda()
    output = m(image)
    assert output.shape == (20, 32, 11, 11)


@pytest.mark.gpu
def test_conv_module_maxpooling():
    from deeptreeattention_b200 import Hang2020
    m = Hang2020.conv_module(in_channels=32, filters=64, maxpool_kernel=(2, 2)).cuda()
    image = torch.randn(20, 32, 11, 11).cuda()
    output = m(image, pool=True)
    assert output.shape == (20, 64, 5, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("conv_dimension", [(20, 32, 11, 11), (20, 64, 5, 5), (20, 128, 2, 2)])
def test_spatial_attention(conv_dimension):
    from deeptreeattention_b200 import Hang2020
    m = Hang2020.spatial_attention(filters=conv_dimension[1]).cuda()
    image = torch.randn(conv_dimension).cuda()
    attention, scores = m(image)
    assert attention.shape == conv_dimension and scores.shape == (20, 4 * conv_dimension[1])


@pytest.mark.gpu
@pytest.mark.parametrize("conv_dimension", [(20, 32, 11, 11), (20, 64, 5, 5), (20, 128, 2, 2)])
def test_spectral_attention(conv_dimension):
    from deeptreeattention_b200 import Hang2020
    m = Hang2020.spectral_attention(filters=conv_dimension[1]).cuda()
    image = torch.randn(conv_dimension).cuda()
    attention, scores = m(image)
    assert attention.shape == conv_dimension and scores.shape == (20, conv_dimension[1])


@pytest.mark.gpu
def test_blocks_compose_like_the_fused_network():
    """vanilla_CNN rebuilt from its stand-alone blocks equals the fused vanilla_CNN (same parameters, eval mode)."""
    from deeptreeattention_b200 import Hang2020 as H
    from oracle import hang2020_oracle as orc
    table = orc.init_params("vanilla", 12, 5, 3, perturb_bn=True)
    x, _ = orc.make_inputs(6, 12, 5, 3)
    net = H.vanilla_CNN(12, 5)
    net.load_state_dict(table)
    net = net.cuda().eval()
    xd = x.cuda()
    fused = net(xd)
    h = net.conv1(xd)
    h = net.conv2(h, pool=True)
    h = net.conv3(h, pool=True)
    head = H.Classifier(in_features=512, classes=5).cuda()
    head.fc1 = net.fc1
    composed = head(torch.flatten(h, start_dim=1))
    np.testing.assert_allclose(composed.detach().cpu().numpy(), fused.detach().cpu().numpy(), rtol=0, atol=1e-4)


@pytest.mark.gpu
def test_block_argument_errors():
    from deeptreeattention_b200 import Hang2020 as H
    with pytest.raises(ValueError):
        H.conv_module(in_channels=8, filters=4).cuda()(torch.randn(2, 7, 11, 11).cuda())        # channel mismatch
    with pytest.raises(AttributeError):
        H.conv_module(in_channels=8, filters=4).cuda()(torch.randn(2, 8, 11, 11).cuda(), pool=True)   # no max_pool, like the reference
    with pytest.raises(ValueError):
        H.spatial_attention(filters=64).cuda()(torch.randn(2, 32, 5, 5).cuda())
    with pytest.raises(TypeError):
        H.spectral_attention(filters=32).cuda()(torch.randn(2, 32, 5, 5).cuda().double())
    with pytest.raises(ValueError):
        H.Classifier(in_features=16, classes=3).cuda()(torch.randn(2, 8).cuda())
    # a plane larger than the crops of the networks works too (generic kernels)
    m = H.spatial_attention(filters=32).cuda()
    out, feat = m(torch.randn(2, 32, 16, 13).cuda())
    assert out.shape == (2, 32, 16, 13) and feat.shape == (2, 32 * 4 * 3)
