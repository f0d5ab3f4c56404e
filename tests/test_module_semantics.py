"""nn.Module semantics the reference gets for free from stock torch layers and the fused CUDA modules have to honour
explicitly: gradient accumulation into persistent gradient buffers, swapping sub-modules after the first call, deepcopy /
pickle of a used module, optimizer checkpoints of the graph-captured Adam step."""
import copy
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hang2020_oracle as orc

pytestmark = pytest.mark.gpu


def _model(bands=20, classes=6, seed=3):
    from deeptreeattention_b200 import Hang2020 as H
    m = H.Hang2020(bands, classes)
    m.load_state_dict(orc.init_params("hang2020", bands, classes, seed, perturb_bn=True))
    return m.cuda().train()


def _persistent_buffers(m):
    """What distributed.GradSync(peer=True) installs: one flat float32 buffer + alpha's float64 scalar the backward writes
    its gradients into every step (symmetric memory there, plain device memory here)."""
    n = sum(p.numel() for p in m.parameters() if p.dtype == torch.float32)
    flat = torch.zeros(n, device="cuda")
    alpha = torch.zeros((), dtype=torch.float64, device="cuda")
    m.__dict__["_grad_buffers"] = (flat, alpha)
    return flat, alpha


def test_gradient_accumulation_with_persistent_buffers():
    """Two backward passes without clearing .grad must leave g1 + g2 (micro-batch accumulation, zero_grad(set_to_none=False)),
    also when the gradients live in the persistent exchange buffer; after p.grad = None the buffer is reused in place."""
    x1, y1 = orc.make_inputs(6, 20, 6, 1)
    x2, y2 = orc.make_inputs(6, 20, 6, 2)
    ref = _model()
    g = []
    for x, y in ((x1, y1), (x2, y2)):
        for p in ref.parameters():
            p.grad = None
        F.cross_entropy(ref(x.cuda()), y.cuda()).backward()
        g.append({k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None})
    m = _model()
    flat, _ = _persistent_buffers(m)
    F.cross_entropy(m(x1.cuda()), y1.cuda()).backward()
    w = m.spectral_network.conv1.conv_layer.weight
    assert flat.data_ptr() <= w.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4      # views of the persistent buffer
    # BatchNorm running statistics moved between the two reference passes too: replay them in the same order
    F.cross_entropy(m(x2.cuda()), y2.cuda()).backward()
    for k, p in m.named_parameters():
        if k in g[0]:
            want = g[0][k] + g[1][k]
            assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-7 + 1e-6 * float(want.abs().max())), k
    assert m.fused_spec().flat_grad is flat                # GradSync's fast path still names the buffer holding the sums
    for p in m.parameters():
        p.grad = None
    F.cross_entropy(m(x2.cuda()), y2.cuda()).backward()
    assert flat.data_ptr() <= w.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4


def test_swapped_submodules_are_picked_up():
    """Replacing a head, one tensor or a whole branch after the first forward: the next call uses (and trains) the new one."""
    from deeptreeattention_b200 import Hang2020 as H
    m = _model()
    x, y = orc.make_inputs(5, 20, 6, 4)
    xd, yd = x.cuda(), y.cuda()
    m(xd)
    # 1. swap the last spatial head
    new_head = H.Classifier(in_features=512, classes=6).cuda()
    old_w = m.spatial_network.classifier3.fc1.weight
    m.spatial_network.classifier3 = new_head
    F.cross_entropy(m(xd), yd).backward()
    assert new_head.fc1.weight.grad is not None and old_w.grad is None
    table = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    m.eval()
    ref_eval = orc.forward("hang2020", table, x, training=False)[0]
    with torch.no_grad():
        np.testing.assert_allclose(m(xd).cpu().numpy(), ref_eval.detach().numpy(), atol=1e-3, rtol=0)
    m.train()
    # 2. re-assign one parameter
    m.alpha = torch.nn.Parameter(torch.tensor(-0.3, dtype=torch.float64, device="cuda"))
    out = m(xd)
    F.cross_entropy(out, yd).backward()
    assert m.alpha.grad is not None
    w_blend = torch.sigmoid(torch.tensor(-0.3, dtype=torch.float64))
    heads = m.head_scores
    np.testing.assert_allclose(out.detach().cpu().numpy(),
                               (heads[2].detach().cpu().double() * w_blend + heads[5].detach().cpu().double() * (1 - w_blend)).float().numpy(),
                               atol=1e-5, rtol=0)
    # 3. a different class count in one head only is an error, in all heads it is the new class count
    m.spatial_network.classifier3 = H.Classifier(in_features=512, classes=9).cuda()
    with pytest.raises(ValueError):
        m(xd)


def test_deepcopy_and_pickle_after_use():
    m = _model()
    x, y = orc.make_inputs(4, 20, 6, 5)
    _persistent_buffers(m)
    F.cross_entropy(m(x.cuda()), y.cuda()).backward()
    c = copy.deepcopy(m)                     # EMA / SWA callbacks do this
    assert "_fused_cache" not in c.__dict__ and "_grad_buffers" not in c.__dict__
    m.eval(); c.eval()
    with torch.no_grad():
        assert torch.equal(m(x.cuda()), c(x.cuda()))
    buf = io.BytesIO()
    torch.save(m, buf)                       # whole-module checkpoint
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    with torch.no_grad():
        assert torch.equal(m(x.cuda()), r(x.cuda()))


def test_capturable_adam_checkpoint_carries_the_device_step():
    """FusedAdam(capturable=True): the step counter lives on the device and advances on CUDA-graph replays; state_dict() must
    save THAT count so a resumed run continues the bias correction."""
    from deeptreeattention_b200.graph import GraphedTrainStep
    from deeptreeattention_b200.loss import cross_entropy_heads
    from deeptreeattention_b200.optim import FusedAdam
    m = _model()
    x, y = orc.make_inputs(6, 20, 6, 6)
    opt = FusedAdam(m.parameters(), lr=1e-3, capturable=True)
    step = GraphedTrainStep(m, x.cuda(), y.cuda(), lambda mm, out, yy: cross_entropy_heads([out], yy), warmup=1, optimizer=opt)
    for _ in range(7):
        step()
    torch.cuda.synchronize()
    sd = opt.state_dict()
    steps = {int(s["step"]) for s in sd["state"].values() if "step" in s}
    assert steps == {1 + 7}, steps           # one warm-up + seven replays (the capture pass records, it does not execute)
    opt2 = FusedAdam(m.parameters(), lr=1e-3, capturable=True)
    opt2.load_state_dict(sd)
    assert int(opt2._flat[0]["step_dev"].item()) == 8
