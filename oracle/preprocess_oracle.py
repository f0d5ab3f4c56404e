"""CPU oracle for the crop preprocessing -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as hang2020_oracle).

``preprocess_image`` restates /root/reference/src/utils.py:36-57 line for line (band clipping :39-42, float32 cast
:44, sklearn ``minmax_scale(data, axis=1)`` on the (pixels, bands) matrix :45-49, reshape :50) without the geo
imports of that file (rasterio, which is not installed here).  ``minmax_rows_float32`` restates the float32
arithmetic sklearn performs (MinMaxScaler.partial_fit / transform) in plain numpy; the test suite checks the two
against each other, which pins the restatement to the library call the reference makes."""
import warnings

import numpy as np


def preprocess_image(image: np.ndarray) -> np.ndarray:
    """image (C, H, W) any integer/float dtype -> float32 (C', H, W) scaled per pixel to [0, 1]."""
    from sklearn import preprocessing
    if image.shape[0] > 3:
        image = image[10:, :, :]
        image = image[:-10, :, :]
    img = np.asarray(image, dtype="float32")
    data = img.reshape(img.shape[0], np.prod(img.shape[1:])).T
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        data = preprocessing.minmax_scale(data, axis=1).T
    return data.reshape(img.shape)


def minmax_rows_float32(data: np.ndarray) -> np.ndarray:
    """data (pixels, bands) float32: out = x*scale + offset with scale = 1/range (range < 10 eps -> 1) and
    offset = 0 - min*scale, every operation rounded to float32 -- sklearn's MinMaxScaler on data.T."""
    data = np.asarray(data, dtype=np.float32)
    mn = data.min(axis=1, keepdims=True)
    mx = data.max(axis=1, keepdims=True)
    rng = (mx - mn).astype(np.float32)
    rng = np.where(rng < 10 * np.finfo(np.float32).eps, np.float32(1.0), rng).astype(np.float32)
    scale = (np.float32(1.0) / rng).astype(np.float32)
    offset = (np.float32(0.0) - (mn * scale).astype(np.float32)).astype(np.float32)
    return ((data * scale).astype(np.float32) + offset).astype(np.float32)
