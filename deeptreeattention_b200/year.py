"""Drop-in for the reference's ``src/models/year.py``: one spectral network per year, all-zero years skipped,
mean of the last-head scores (reference :9-33).  Every year network runs through the CUDA library."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import Module

from . import Hang2020


class learned_ensemble(Module):
    """``learned_ensemble(years, classes, config)`` with ``config["bands"]`` and ``config["pretrain_state_dict"]``
    exactly as in the reference (year.py:10-22)."""

    def __init__(self, years, classes, config):
        super().__init__()
        self.year_models = nn.ModuleList()
        self.years = years
        for _ in range(years):
            if config.get("pretrain_state_dict"):
                base_model = Hang2020.load_from_backbone(state_dict=config["pretrain_state_dict"], classes=classes, bands=config["bands"])
            else:
                base_model = Hang2020.spectral_network(bands=config["bands"], classes=classes)
            self.year_models.append(base_model)

    def forward(self, images):
        """``images``: one (B, bands, 11, 11) tensor per year.  A year whose tensor sums to zero is skipped
        (reference :27; the test is a device->host sync there too)."""
        # one host sync for all years instead of one per year
        sums = torch.stack([x.sum() for x in images]).tolist()
        year_scores = []
        for index, x in enumerate(images):
            if sums[index] == 0:
                continue
            year_scores.append(self.year_models[index](x)[-1])
        return torch.stack(year_scores, axis=1).mean(axis=1)
