"""deeptreeattention_b200 -- B200-native (sm_100a) Hang2020 hot path for DeepTreeAttention.

``Hang2020``     mirrors the reference's ``src.models.Hang2020`` namespace (networks fused, blocks stand-alone)
``year``         ``src.models.year.learned_ensemble`` (+ device-side zero-year flags / masked mean)
``multi_stage``  ``base_model`` and the arithmetic of ``MultiStage.predict_step`` (``src.models.multi_stage``)
``metadata``     ``src.models.metadata`` (site MLP + fusion around the CUDA Hang2020)
``loss``         fused weighted cross-entropy over the heads (``TreeModel.training_step``'s loss line)
``train``        forward + summed head losses + backward as ONE library call (``TreeModel.training_step`` + ``backward()``)
``optim``        ``FusedAdam``: the optimizer the reference configures, one launch per step
``data``         int16 crop preprocessing on the device (``utils.preprocess_image``)
``graph``        CUDA-graph capture of a whole training step
``distributed``  one process per GPU: crop sharding and the gradient exchange
``_capi``        ctypes binding of the C-ABI library ``libdta_b200.so`` (``include/dta_b200.h``)

Sub-modules other than ``Hang2020`` and ``_capi`` are imported on demand (``from deeptreeattention_b200 import optim``).
"""
from . import _capi  # noqa: F401
from . import Hang2020  # noqa: F401

__all__ = ["Hang2020", "_capi", "year", "multi_stage", "metadata", "loss", "train", "optim", "data", "graph", "distributed"]
__version__ = "0.2.0"
