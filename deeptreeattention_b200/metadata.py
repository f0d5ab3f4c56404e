"""Drop-in for the ``metadata`` / ``metadata_sensor_fusion`` modules of the reference's ``src/models/metadata.py``
(:9-44, BASELINE config 5): site-embedding MLP fused late with the Hang2020 joint scores.  The sensor model is the
CUDA-library Hang2020; the 16-wide site MLP and the 2C -> C fusion layer are plain torch layers (their cost is
noise next to the crops: SURVEY.md 8a row a12)."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import Module
from torch.nn import functional as F

from .Hang2020 import Hang2020


class metadata(Module):
    """Embedding(sites, 16) -> BatchNorm1d -> Dropout(0.7) -> Linear(16, classes) -> ReLU (reference :9-24)."""

    def __init__(self, sites, classes):
        super().__init__()
        self.embedding = nn.Embedding(sites, 16)
        self.batch_norm = nn.BatchNorm1d(16)
        self.mlp = nn.Linear(in_features=16, out_features=classes)
        self.dropout = nn.Dropout(p=0.7)

    def forward(self, x):
        x = self.embedding(x)
        x = self.batch_norm(x)
        x = self.dropout(x)
        x = self.mlp(x)
        return F.relu(x)


class metadata_sensor_fusion(Module):
    """Joint fusion of the HSI sensor model and the site metadata (reference :26-44)."""

    def __init__(self, bands, sites, classes):
        super().__init__()
        self.metadata_model = metadata(sites, classes)
        self.sensor_model = Hang2020(bands, classes)
        self.fc1 = nn.Linear(in_features=classes * 2, out_features=classes)

    def forward(self, images, metadata):
        metadata_softmax = self.metadata_model(metadata)
        sensor_softmax = self.sensor_model(images)
        concat_features = torch.cat([metadata_softmax, sensor_softmax], dim=1)
        return F.relu(self.fc1(concat_features))
