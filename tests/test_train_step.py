"""dta_train_step / train.fused_train_step: forward + summed weighted cross-entropy over every head + backward in one library
call must leave exactly what the three-call sequence (model(x); loss.cross_entropy_heads; loss.backward()) leaves -- the caller
of TreeModel.training_step (/root/reference/src/main.py:71-80) with the head losses summed -- and match the oracle."""
import numpy as np
import pytest
import torch

from oracle import hang2020_oracle as orc

pytestmark = pytest.mark.gpu


def _modules():
    from deeptreeattention_b200 import Hang2020 as H
    return {"hang2020": H.Hang2020, "spectral": H.spectral_network, "spatial": H.spatial_network, "vanilla": H.vanilla_CNN}


def _fresh(kind, bands, classes, table, training):
    m = _modules()[kind](bands, classes)
    m.load_state_dict(table)
    return m.cuda().train(training)


def _heads(kind, m, out):
    return m.head_scores if kind == "hang2020" else ([out] if kind == "vanilla" else out)


@pytest.mark.parametrize("kind,bands,classes,batch,training,weighted", [
    ("hang2020", 369, 50, 40, True, True), ("hang2020", 30, 7, 6, True, False), ("hang2020", 30, 7, 6, False, True),
    ("spectral", 30, 7, 5, True, True), ("spatial", 12, 4, 9, True, False), ("vanilla", 12, 4, 9, True, True),
    ("hang2020", 20, 5, 1100, True, True)])
def test_fused_train_step_equals_three_call_sequence(kind, bands, classes, batch, training, weighted):
    from deeptreeattention_b200 import _capi
    from deeptreeattention_b200.loss import cross_entropy_heads
    from deeptreeattention_b200.train import fused_train_step
    table = orc.init_params(kind, bands, classes, 37, perturb_bn=True)
    x, y = orc.make_inputs(batch, bands, classes, 37)
    w = (torch.rand(classes, generator=torch.Generator().manual_seed(3)) + 0.5).cuda() if weighted else None
    y = y.clone()
    xd, yd = x.cuda(), y.cuda()

    ma = _fresh(kind, bands, classes, table, training)
    out = ma(xd)
    loss_a = cross_entropy_heads(_heads(kind, ma, out), yd, w)
    loss_a.backward()

    mb = _fresh(kind, bands, classes, table, training)
    _capi.POISON_GRADS = True          # the one-call step must write every gradient too
    try:
        loss_b = fused_train_step(mb, xd, yd, w, want_joint=True)
    finally:
        _capi.POISON_GRADS = False
    torch.cuda.synchronize()
    assert float(loss_b) == float(loss_a)
    for ha, hb in zip(_heads(kind, ma, out), mb.head_scores if kind != "vanilla" else [None]):
        if hb is not None:
            assert torch.equal(ha.detach(), hb)
    if kind == "hang2020":
        assert torch.equal(out.detach(), mb.joint_scores)
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert (pa.grad is None) == (pb.grad is None), k
        if pa.grad is not None:
            assert not torch.isnan(pb.grad).any(), k
            assert torch.equal(pa.grad, pb.grad), k
    assert not torch.isnan(mb.fused_spec().flat_grad).any()
    for (k, ba), (_, bb) in zip(ma.named_buffers(), mb.named_buffers()):
        assert torch.equal(ba, bb), k


def test_fused_train_step_matches_oracle():
    from deeptreeattention_b200.train import fused_train_step
    kind, bands, classes, batch = "hang2020", 40, 6, 8
    table = orc.init_params(kind, bands, classes, 41, perturb_bn=True)
    x, y = orc.make_inputs(batch, bands, classes, 41)
    loss_ref, _, heads_ref, grads_ref, _ = orc.step(kind, table, x, y, regime="R2", training=True)
    m = _fresh(kind, bands, classes, table, True)
    loss = fused_train_step(m, x.cuda(), y.cuda())
    assert abs(float(loss) - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    for h, hr in zip(m.head_scores, heads_ref):
        np.testing.assert_allclose(h.cpu().numpy(), hr.detach().numpy(), rtol=0, atol=1e-4)
    for k, p in m.named_parameters():
        g = grads_ref.get(k)
        if g is None or p.grad is None:
            assert k == "alpha" or (g is None) == (p.grad is None), k
            continue
        # (conv biases in front of a train-mode BatchNorm have an exactly zero gradient here, rounding noise in the oracle)
        num, den = (p.grad.cpu().double() - g.double()).norm().item(), g.double().norm().item()
        assert num <= 2e-3 * den + 2e-6, (k, num, den)


def test_graphed_fused_train_step_equals_eager():
    from deeptreeattention_b200.train import GraphedFusedTrainStep, fused_train_step
    table = orc.init_params("hang2020", 40, 6, 11, perturb_bn=True)
    x, y = orc.make_inputs(12, 40, 6, 11)
    x2, y2 = orc.make_inputs(12, 40, 6, 12)
    me = _fresh("hang2020", 40, 6, table, True)
    loss_e = fused_train_step(me, x2.cuda(), y2.cuda())
    mg = _fresh("hang2020", 40, 6, table, True)
    step = GraphedFusedTrainStep(mg, x.cuda(), y.cuda(), warmup=1)
    loss_g = step(x2.cuda(), y2.cuda())
    torch.cuda.synchronize()
    assert float(loss_g) == float(loss_e)
    for (k, pe), (_, pg) in zip(me.named_parameters(), mg.named_parameters()):
        assert (pe.grad is None) == (pg.grad is None), k
        if pe.grad is not None:
            assert torch.equal(pe.grad, pg.grad), k


def test_fused_train_step_rejects_cpu_and_bad_labels():
    from deeptreeattention_b200.train import fused_train_step
    m = _modules()["spectral"](12, 4).cuda()
    with pytest.raises(RuntimeError):
        fused_train_step(m, torch.zeros(2, 12, 11, 11), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(ValueError):
        fused_train_step(m, torch.zeros(2, 12, 11, 11, device="cuda"), torch.zeros(3, dtype=torch.int64, device="cuda"))
