"""Crop preprocessing (src/utils.py:36-57): the numpy restatement against the sklearn call the reference makes
(CPU), and the CUDA kernel against the oracle, bit for bit (GPU)."""
import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po


def raw_crops(batch, bands, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    raw = rng.integers(-50, 10000, size=(batch, bands, 11, 11)).astype(np.int16)   # NEON reflectance x 10000, some nodata < 0
    raw[0, :, 3, 4] = 1234                                                           # a constant spectrum -> all zeros
    return raw


@pytest.mark.parametrize("bands", [369, 30, 3])
def test_numpy_restatement_matches_sklearn_path(bands):
    raw = raw_crops(2, bands, 1)[0]
    ref = po.preprocess_image(raw)
    clip = 10 if bands > 3 else 0
    kept = raw[clip:bands - clip].astype(np.float32)
    mine = po.minmax_rows_float32(kept.reshape(kept.shape[0], -1).T).T.reshape(kept.shape)
    assert ref.dtype == np.float32 and ref.shape == (bands - 2 * clip, 11, 11)
    assert np.array_equal(ref, mine)
    assert ref.min() == 0.0 and abs(float(ref.max()) - 1.0) < 1e-6      # two float32 roundings: 1.0000001 does occur
    assert np.all(ref[:, 3, 4] == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("batch,bands", [(5, 369), (2, 30), (3, 3)])
def test_cuda_preprocess_is_bit_identical(batch, bands):
    from deeptreeattention_b200.data import preprocess_crops
    raw = raw_crops(batch, bands, 2)
    got = preprocess_crops(torch.from_numpy(raw).cuda()).cpu().numpy()
    ref = np.stack([po.preprocess_image(r) for r in raw])
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_preprocess_rejects_cpu_and_wrong_dtype():
    from deeptreeattention_b200.data import preprocess_crops
    with pytest.raises(RuntimeError):
        preprocess_crops(torch.zeros(1, 30, 11, 11, dtype=torch.int16))
