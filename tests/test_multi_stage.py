"""Year ensemble and prediction fan-out (reference src/models/year.py:9-33, src/models/multi_stage.py:17-33,306-318)
against the oracle: every level model on the same per-year crops, zero years skipped, softmax per level."""
import numpy as np
import pytest
import torch

from oracle import hang2020_oracle as orc


def test_multi_stage_surface_and_cpu_refusal():
    from deeptreeattention_b200 import multi_stage as MS, year
    m = MS.base_model(years=2, classes=3, config={"bands": 5, "pretrain_state_dict": None})
    assert isinstance(m.model, year.learned_ensemble) and len(m.model.year_models) == 2
    assert [k for k in m.state_dict() if k.endswith("classifier3.fc1.bias")] == [
        "model.year_models.0.classifier3.fc1.bias", "model.year_models.1.classifier3.fc1.bias"]
    with pytest.raises(RuntimeError):
        MS.predict_step([m.eval()], [torch.randn(1, 5, 11, 11) for _ in range(2)])
    with pytest.raises(RuntimeError):
        year.crops_nonzero([torch.zeros(1, 5, 11, 11)])


def _levels(classes_per_level, years, bands):
    from deeptreeattention_b200 import multi_stage as MS
    torch.manual_seed(3)
    return [MS.base_model(years=years, classes=c, config={"bands": bands, "pretrain_state_dict": None}).cuda().eval()
            for c in classes_per_level]


def _oracle_level(model, images, keep):
    scores = []
    for y in keep:
        table = {k: v.detach().cpu() for k, v in model.model.year_models[y].state_dict().items()}
        scores.append(orc.forward("spectral", table, images[y], training=False)[0][-1])
    return torch.softmax(torch.stack(scores, 1).mean(1), dim=1)


@pytest.mark.gpu
@pytest.mark.parametrize("classes_per_level", [(2, 5, 7), (4, 4), (3,), (2, 3, 9, 5, 6)])
def test_predict_step_matches_oracle(classes_per_level):
    from deeptreeattention_b200 import multi_stage as MS
    bands, B, years = 45, 6, 3
    models = _levels(classes_per_level, years, bands)
    for m in models:                                   # running statistics away from their initial values
        for k, v in m.state_dict().items():
            if k.endswith("running_mean"):
                v.normal_(0, 0.1)
            elif k.endswith("running_var"):
                v.uniform_(0.5, 1.5)
    images = [orc.make_inputs(B, bands, 2, 70, "uniform")[0], torch.zeros(B, bands, 11, 11), orc.make_inputs(B, bands, 2, 71, "normal")[0]]
    y_hats = MS.predict_step(models, [x.cuda() for x in images])
    assert len(y_hats) == len(models)
    for m, c, y_hat in zip(models, classes_per_level, y_hats):
        assert y_hat.shape == (B, c)
        ref = _oracle_level(m, images, keep=(0, 2))
        np.testing.assert_allclose(y_hat.cpu().numpy(), ref.numpy(), rtol=0, atol=1e-4)
        assert np.array_equal(y_hat.cpu().numpy().argmax(1), ref.numpy().argmax(1))
        # the reference loop itself (model.forward + softmax) through the drop-in modules agrees too
        with torch.no_grad():
            loop = torch.softmax(m([x.cuda() for x in images]), dim=1)
        np.testing.assert_allclose(loop.cpu().numpy(), y_hat.cpu().numpy(), rtol=0, atol=2e-5)


@pytest.mark.gpu
def test_crops_nonzero_and_ensemble_mean():
    from deeptreeattention_b200 import year
    g = torch.Generator().manual_seed(0)
    crops = [torch.rand(5, 7, 11, 11, generator=g), torch.zeros(5, 7, 11, 11), torch.rand(5, 7, 11, 11, generator=g) * 1e-30,
             torch.zeros(5, 7, 11, 11)]
    crops[3][4, 6, 10, 10] = 1e-20                       # a single tiny element is not a zero year
    flags = year.crops_nonzero([c.cuda() for c in crops])
    assert flags.cpu().tolist() == [float(c.sum() != 0) for c in crops] == [1.0, 0.0, 1.0, 1.0]
    odd = [torch.rand(3, 5, 11, 11, generator=g)[:, :, :, :].contiguous()[1:], torch.zeros(2, 5, 11, 11)]   # element count not a multiple of 4 rows
    assert year.crops_nonzero([c.cuda() for c in odd]).cpu().tolist() == [1.0, 0.0]
    scores = [torch.randn(9, 37, generator=g) for _ in range(4)]
    out = year.ensemble_mean([s.cuda() for s in scores], flags, softmax=False).cpu()
    ref = torch.stack([scores[0], scores[2], scores[3]], 1).mean(1)
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-6, atol=1e-6)
    out = year.ensemble_mean([s.cuda() for s in scores], None, softmax=True).cpu()
    ref = torch.softmax(torch.stack(scores, 1).mean(1), dim=1)
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("host_skip", [False, True])
def test_training_ensemble_leaves_skipped_year_untouched(host_skip):
    """train(): the all-zero year (year.py:27) must not move its BatchNorm buffers, must not contribute to the mean and gets no
    gradient -- on the device (flag-gated buffer update, flag-weighted mean; no host sync) as with the reference's control
    flow (host_skip=True); both agree with each other and with the oracle's mean over the live years."""
    from deeptreeattention_b200 import year
    torch.manual_seed(1)
    m = year.learned_ensemble(years=3, classes=4, config={"bands": 20, "pretrain_state_dict": None}, host_skip=host_skip).cuda().train()
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    xs = [orc.make_inputs(5, 20, 4, 4)[0], torch.zeros(5, 20, 11, 11), orc.make_inputs(5, 20, 4, 5)[0]]
    y = orc.make_inputs(5, 20, 4, 5)[1]
    out = m([x.cuda() for x in xs])
    torch.nn.functional.cross_entropy(out, y.cuda()).backward()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    # buffers: years 0 and 2 stepped once, year 1 bit-identical to before
    for k, v in m.state_dict().items():
        if orc.is_buffer(k):
            if k.startswith("year_models.1."):
                assert torch.equal(v, before[k]), k
            elif k.endswith("num_batches_tracked"):
                assert int(v) == 1, k
    # the mean over the live years, against the oracle in train mode on the same parameters
    ref = []
    for i in (0, 2):
        table = {k[len(f"year_models.{i}."):]: v.cpu() for k, v in before.items() if k.startswith(f"year_models.{i}.")}
        ref.append(orc.forward("spectral", table, xs[i], training=True)[0][-1])
    np.testing.assert_allclose(out.detach().cpu().numpy(), torch.stack(ref, 1).mean(1).detach().numpy(), rtol=0, atol=1e-3)
    g1 = m.year_models[1].conv1.conv_layer.weight.grad
    assert g1 is None or float(g1.abs().max()) == 0.0           # reference: None (the year never ran)
    for i in (0, 2):
        g = m.year_models[i].conv1.conv_layer.weight.grad
        assert g is not None and float(g.abs().max()) > 0.0


@pytest.mark.gpu
def test_training_ensemble_device_skip_equals_host_skip():
    """Same seed, same crops: the device-gated path and the reference's host-side skip give the same scores and gradients."""
    from deeptreeattention_b200 import year
    xs = [orc.make_inputs(6, 12, 3, 8)[0].cuda(), torch.zeros(6, 12, 11, 11).cuda(), orc.make_inputs(6, 12, 3, 9)[0].cuda()]
    y = orc.make_inputs(6, 12, 3, 9)[1].cuda()
    runs = []
    for host_skip in (False, True):
        torch.manual_seed(5)
        m = year.learned_ensemble(years=3, classes=3, config={"bands": 12, "pretrain_state_dict": None}, host_skip=host_skip).cuda().train()
        out = m(xs)
        torch.nn.functional.cross_entropy(out, y).backward()
        runs.append((out.detach(), {k: p.grad for k, p in m.named_parameters()}, {k: v.clone() for k, v in m.state_dict().items()}))
    np.testing.assert_allclose(runs[0][0].cpu().numpy(), runs[1][0].cpu().numpy(), rtol=0, atol=1e-6)
    for k, g in runs[1][1].items():
        if g is None:
            assert runs[0][1][k] is None or float(runs[0][1][k].abs().max()) == 0.0, k
        else:
            scale = float(g.abs().max())
            assert float((runs[0][1][k] - g).abs().max()) <= 1e-5 * scale + 1e-8, k
    for k, v in runs[1][2].items():
        assert torch.allclose(runs[0][2][k].float(), v.float(), rtol=1e-6, atol=1e-7), k
