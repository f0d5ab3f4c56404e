"""CPU: the C-ABI library builds/loads and exports every symbol the header declares; host-side
logic of the drop-in modules (state_dict contract, error behaviour) -- no compute calls."""
import ctypes
import os
import re

import pytest
import torch

from oracle import hang2020_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dta_b200.h")).read()
    return sorted(set(re.findall(r"\b(dta_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from deeptreeattention_b200 import _capi
    path = _capi.build()
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert set(syms) == set(_capi.EXPORTS), (syms, _capi.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), s
    assert _capi.lib().dta_abi_version() == 1


def test_query_sizes_and_bad_shapes():
    from deeptreeattention_b200 import _capi
    s = _capi.query_sizes(_capi.NET_HANG2020, 1024, 369, 50, True)
    assert s.n_heads == 6 and s.saved_bytes > 1024 * 118 * 1024 and s.workspace_bwd > 0
    assert _capi.query_sizes(_capi.NET_VANILLA, 4, 3, 2, True).n_heads == 1
    with pytest.raises(ValueError):
        _capi.query_sizes(9, 4, 3, 2, True)
    with pytest.raises(ValueError):
        _capi.query_sizes(_capi.NET_HANG2020, 0, 3, 2, True)


def test_host_side_size_queries_of_blocks_and_pairs():
    """Pure host arithmetic of the newer entry points (no GPU needed): block workspaces, attention geometry, pair kinds."""
    import ctypes as C
    from deeptreeattention_b200 import _capi
    L = _capi.lib()
    plane = _capi.Plane(20, 369, 11, 11)
    need = C.c_size_t()
    assert L.dta_conv_module_workspace_bytes(C.byref(plane), 32, C.byref(need)) == 0
    assert need.value == (20 * 32 * 121 + 3 * 32) * 4
    assert L.dta_conv_module_workspace_bytes(C.byref(plane), 0, C.byref(need)) == -1
    feat, saved, work = C.c_size_t(), C.c_size_t(), C.c_size_t()
    for kind, filters, side, nfeat, nsaved in ((_capi.ATTN_SPECTRAL, 32, 11, 32, 96), (_capi.ATTN_SPECTRAL, 128, 2, 128, 384),
                                               (_capi.ATTN_SPATIAL, 32, 11, 128, 363), (_capi.ATTN_SPATIAL, 64, 5, 256, 75),
                                               (_capi.ATTN_SPATIAL, 128, 2, 512, 12)):
        pl = _capi.Plane(20, filters, side, side)
        assert L.dta_attention_sizes(kind, C.byref(pl), C.byref(feat), C.byref(saved), C.byref(work)) == 0
        assert (feat.value, saved.value) == (nfeat, nsaved) and work.value > 0
    bad = _capi.Plane(20, 48, 11, 11)                      # reference: ValueError for filters outside {32, 64, 128}
    assert L.dta_attention_sizes(_capi.ATTN_SPATIAL, C.byref(bad), C.byref(feat), C.byref(saved), C.byref(work)) == -2
    pair = _capi.query_sizes(_capi.NET_SPECTRAL_PAIR, 64, 369, 9, False)
    single = _capi.query_sizes(_capi.NET_SPECTRAL, 64, 369, 9, False)
    assert pair.n_heads == 6 and single.n_heads == 3 and pair.saved_bytes > single.saved_bytes
    hyper = _capi.AdamHyper(1e-3, 0.9, 0.999, 1e-8, 0.0, 1)
    assert L.dta_adam_step(None, 0, None, None, None, None, None, None, None, None, None, C.byref(hyper), None, None, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from deeptreeattention_b200 import _capi
    with pytest.raises(_capi.DtaError) as e:
        _capi.context(0)
    assert "no CPU fallback" in str(e.value) or "no CUDA device" in str(e.value)


@pytest.mark.parametrize("kind,cls", [("hang2020", "Hang2020"), ("spectral", "spectral_network"),
                                      ("spatial", "spatial_network"), ("vanilla", "vanilla_CNN")])
def test_state_dict_contract(kind, cls):
    from deeptreeattention_b200 import Hang2020 as H
    m = getattr(H, cls)(bands=369, classes=10)
    sd = m.state_dict()
    want = orc.param_shapes(kind, 369, 10)
    assert list(sd.keys()) == [n for n, _, _ in want]
    for n, shape, role in want:
        assert tuple(sd[n].shape) == tuple(shape), n
        if role == "alpha":
            assert sd[n].dtype == torch.float64
        elif role == "bn_nbt":
            assert sd[n].dtype == torch.int64
        else:
            assert sd[n].dtype == torch.float32
    m.load_state_dict(orc.init_params(kind, 369, 10, 0))


def test_reference_api_surface_and_errors(tmp_path):
    from deeptreeattention_b200 import Hang2020 as H
    for name in ("global_spectral_pool", "conv_module", "vanilla_CNN", "Classifier", "spatial_attention",
                 "spectral_attention", "spatial_network", "spectral_network", "Hang2020", "load_from_backbone"):
        assert hasattr(H, name), name
    with pytest.raises(ValueError):
        H.spectral_attention(filters=48)
    with pytest.raises(ValueError):
        H.spatial_attention(filters=48)
    with pytest.raises(RuntimeError):
        H.global_spectral_pool(torch.ones(2, 3, 4, 4))     # no CPU path for the helper either
    m = H.Hang2020(bands=3, classes=10)
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 3, 11, 11))          # no CPU path, must not silently fall back
    # load_from_backbone: 10 -> 20 classes keeps every non-classifier tensor (tests/test_Hang2020.py:66-75)
    path = str(tmp_path / "state_dict.pt")
    torch.save(m.spectral_network.state_dict(), path)
    m20 = H.load_from_backbone(state_dict=path, classes=20, bands=3)
    assert m20.classifier3.fc1.weight.shape == (20, 128)
    assert torch.equal(m20.conv1.conv_layer.weight, m.spectral_network.conv1.conv_layer.weight)
    assert torch.equal(m20.attention_2.attention_conv1.weight, m.spectral_network.attention_2.attention_conv1.weight)


def test_same_seed_same_init_as_torch_layers():
    """Parameter containers are real torch.nn layers created in the reference's order, so a
    seeded construction consumes the RNG identically (init under manual_seed(0))."""
    from deeptreeattention_b200 import Hang2020 as H
    torch.manual_seed(0)
    a = H.spectral_network(5, 3).state_dict()
    torch.manual_seed(0)
    b = H.spectral_network(5, 3).state_dict()
    for k in a:
        assert torch.equal(a[k], b[k])
    bound = 1.0 / (5 * 9) ** 0.5
    assert float(a["conv1.conv_layer.weight"].abs().max()) <= bound


def test_year_and_metadata_modules_mirror_the_reference_surface():
    """src/models/year.py:9-33 and src/models/metadata.py:9-44: constructors, parameter trees, CPU inputs refused."""
    from deeptreeattention_b200 import metadata as M, year
    e = year.learned_ensemble(years=3, classes=4, config={"bands": 5, "pretrain_state_dict": None})
    assert len(e.year_models) == 3 and e.years == 3
    assert sorted(k for k in e.state_dict() if k.startswith("year_models.0.classifier3")) == [
        "year_models.0.classifier3.fc1.bias", "year_models.0.classifier3.fc1.weight"]
    with pytest.raises(RuntimeError):
        e([torch.randn(1, 5, 11, 11) for _ in range(3)])
    f = M.metadata_sensor_fusion(bands=3, sites=2, classes=10)
    keys = list(f.state_dict().keys())
    assert keys[0] == "metadata_model.embedding.weight" and "sensor_model.alpha" in keys and keys[-1] == "fc1.bias"
    assert f.fc1.weight.shape == (10, 20)
    with pytest.raises(RuntimeError):                 # the site MLP runs in the CUDA library too: no CPU path
        M.metadata(sites=1, classes=10)(torch.zeros(20).int())
    with pytest.raises(RuntimeError):
        f(torch.randn(2, 3, 11, 11), torch.zeros(2).int())


def test_fused_train_step_host_side_errors():
    """train.fused_train_step checks its arguments on the host before any library call: no CPU path, fused networks only."""
    import torch
    from deeptreeattention_b200 import Hang2020 as H
    from deeptreeattention_b200.train import fused_train_step
    m = H.spectral_network(12, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fused_train_step(m, torch.zeros(2, 12, 11, 11), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(TypeError):
        fused_train_step(torch.nn.Linear(3, 3), torch.zeros(2, 12, 11, 11), torch.zeros(2, dtype=torch.int64))


def test_product_code_never_touches_the_oracle_or_a_cpu_fallback():
    """The oracle is test infrastructure: nothing under deeptreeattention_b200/ may import it, and no product module may
    import the reference tree either."""
    import ast
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deeptreeattention_b200")
    for name in sorted(os.listdir(root)):
        if not name.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(root, name)).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            for m in mods:
                assert not m.split(".")[0] in ("oracle", "src"), f"{name} imports {m}"
