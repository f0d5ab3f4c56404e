"""FusedAdam (dta_adam_step) against torch.optim.Adam, the optimizer the reference configures
(/root/reference/src/main.py:135-136): same parameters after several steps within float32 rounding."""
import pytest
import torch


def test_fused_adam_refuses_cpu_parameters():
    from deeptreeattention_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.randn(4))
    opt = FusedAdam([p], lr=1e-3)
    p.grad = torch.randn(4)
    with pytest.raises(RuntimeError):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)


def _params(seed, device):
    g = torch.Generator().manual_seed(seed)
    shapes = [(32, 369, 3, 3), (32,), (64, 32, 3, 3), (5000,), (1, 1, 7, 7), (50, 512), (3,)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(device)) for s in shapes]
    ps.append(torch.nn.Parameter(torch.tensor(0.5, dtype=torch.float64, device=device)))     # Hang2020.alpha
    return ps


@pytest.mark.gpu
@pytest.mark.parametrize("weight_decay,capturable", [(0.0, False), (0.01, False), (0.0, True)])
def test_fused_adam_matches_torch_adam(weight_decay, capturable):
    from deeptreeattention_b200.optim import FusedAdam
    ours, ref = _params(0, "cuda"), _params(0, "cuda")
    a = FusedAdam(ours, lr=1e-3, weight_decay=weight_decay, capturable=capturable)
    b = torch.optim.Adam(ref, lr=1e-3, weight_decay=weight_decay, foreach=False, fused=False)
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        for i, (p, q) in enumerate(zip(ours, ref)):
            if i == 3 and step % 2 == 1 and not capturable:      # a parameter that sometimes has no gradient (its step count lags)
                p.grad, q.grad = None, None
                continue
            grad = torch.randn(p.shape, generator=g, dtype=torch.float32).to(p.dtype).cuda() * (0.1 if i != 4 else 5.0)
            p.grad, q.grad = grad.clone(), grad.clone()
        if step == 3:
            for group in list(a.param_groups) + list(b.param_groups):
                group["lr"] = 5e-4                                   # what ReduceLROnPlateau does (main.py:138-147)
        a.step()
        b.step()
    torch.cuda.synchronize()
    for p, q in zip(ours, ref):
        assert p.dtype == q.dtype
        err = float((p.detach().double() - q.detach().double()).abs().max())
        assert err <= 2e-6 + 1e-6 * float(q.detach().abs().max()), err
    sa, sb = a.state_dict()["state"], b.state_dict()["state"]
    assert set(sa.keys()) == set(sb.keys())
    for k in sa:
        assert int(sa[k]["step"]) == int(sb[k]["step"]) or capturable
        assert torch.allclose(sa[k]["exp_avg"].float(), sb[k]["exp_avg"].float(), rtol=1e-5, atol=1e-7)
        assert torch.allclose(sa[k]["exp_avg_sq"].float(), sb[k]["exp_avg_sq"].float(), rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_fused_adam_trains_the_fused_model_and_reloads_state():
    from deeptreeattention_b200 import Hang2020 as H
    from deeptreeattention_b200.optim import FusedAdam
    from oracle import hang2020_oracle as orc
    table = orc.init_params("hang2020", 30, 6, 9)
    x, y = orc.make_inputs(16, 30, 6, 9)
    xd, yd = x.cuda(), y.cuda()
    m1, m2 = H.Hang2020(30, 6), H.Hang2020(30, 6)
    m1.load_state_dict(table), m2.load_state_dict(table)
    m1, m2 = m1.cuda().train(), m2.cuda().train()
    o1 = FusedAdam(m1.parameters(), lr=1e-3)
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-3, foreach=False, fused=False)
    losses = []
    for _ in range(3):
        for m, o in ((m1, o1), (m2, o2)):
            o.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(m(xd), yd)
            loss.backward()
        # both optimizers must see the SAME gradients: after the first step the two models differ by an ulp, which can flip
        # a ReLU and move single gradient elements by far more than the optimizers' own rounding (Adam then amplifies it)
        for p, q in zip(m1.parameters(), m2.parameters()):
            if q.grad is not None:
                p.grad.copy_(q.grad)
        o1.step()
        o2.step()
        losses.append(float(loss))
    for (k, p), q in zip(m1.named_parameters(), m2.parameters()):
        assert float((p.detach().double() - q.detach().double()).abs().max()) <= 1e-5 + 1e-5 * float(q.detach().abs().max()), k
    # checkpoint / resume through torch's own state_dict format
    o3 = FusedAdam(m1.parameters(), lr=1e-3)
    o3.load_state_dict(o1.state_dict())
    for p in m1.parameters():
        if p in o1.state:
            assert torch.equal(o3.state[p]["exp_avg"], o1.state[p]["exp_avg"])
            assert int(o3.state[p]["step"]) == int(o1.state[p]["step"])


@pytest.mark.gpu
def test_graph_captured_step_with_fused_adam_matches_eager_torch_adam():
    """forward + loss + backward + FusedAdam(capturable) replayed as ONE CUDA graph equals the eager loop with torch.optim.Adam."""
    from deeptreeattention_b200 import Hang2020 as H
    from deeptreeattention_b200.graph import GraphedTrainStep
    from deeptreeattention_b200.loss import cross_entropy_heads
    from deeptreeattention_b200.optim import FusedAdam
    from oracle import hang2020_oracle as orc
    table = orc.init_params("hang2020", 24, 5, 13)
    x, y = orc.make_inputs(10, 24, 5, 13)
    xd, yd = x.cuda(), y.cuda()
    mg, me = H.Hang2020(24, 5), H.Hang2020(24, 5)
    mg.load_state_dict(table), me.load_state_dict(table)
    mg, me = mg.cuda().train(), me.cuda().train()

    def loss_fn(m, out, yy):
        return cross_entropy_heads(m.head_scores + [out], yy)

    step = GraphedTrainStep(mg, xd, yd, loss_fn, warmup=1, optimizer=FusedAdam(mg.parameters(), lr=2e-3, capturable=True))
    for _ in range(2):
        loss_g = step(xd, yd)
    opt = torch.optim.Adam(me.parameters(), lr=2e-3, foreach=False, fused=False)
    for _ in range(3):                                   # 1 warm-up step + 2 replays
        opt.zero_grad(set_to_none=True)
        loss_e = loss_fn(me, me(xd), yd)
        loss_e.backward()
        opt.step()
    torch.cuda.synchronize()
    assert abs(float(loss_g) - float(loss_e)) < 1e-4
    for (k, p), q in zip(mg.named_parameters(), me.parameters()):
        assert float((p.detach().double() - q.detach().double()).abs().max()) <= 2e-5 + 2e-5 * float(q.detach().abs().max()), k
