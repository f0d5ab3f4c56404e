"""Minimal stand-in for the reference's ``TreeModel`` LightningModule -- TEST INFRASTRUCTURE.

``/root/reference/src/main.py`` cannot be imported anywhere here (pytorch_lightning, torchmetrics, geopandas, rasterio,
deepforest, comet_ml are not installed).  This class restates, without Lightning, exactly the methods
through which ``TreeModel`` touches the model it is given (the drop-in boundary of SURVEY.md 8b):

  * ``__init__(model, classes, label_dict, loss_weight, config)``        main.py:33-69  (model stored as ``self.model``,
    ``loss_weight`` a float tensor on the model's device, ones when none is given)
  * ``training_step`` / ``validation_step``                              main.py:71-94  (``self.model.forward(images)`` on
    ``inputs["HSI"]``, ``F.cross_entropy(y_hat, y, weight=self.loss_weight)``)
  * ``configure_optimizers``                                             main.py:135-149 (Adam(lr) + ReduceLROnPlateau with
    the reference's arguments; ``verbose`` dropped: removed from current torch)
  * ``predict``                                                          main.py:152-163 (``self.model(images)``, result on
    the CPU)

plus ``fit_steps``, the dozen lines of ``Trainer.fit`` that matter for a parity run (zero_grad / backward / step and the
scheduler fed with the monitored validation loss).  Nothing here knows which implementation ``model`` is.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import optim


class TreeModel:
    def __init__(self, model, classes, label_dict, loss_weight=None, config=None):
        self.config = config if config is not None else {"lr": 1e-4, "top_k": 1}
        self.classes = classes
        self.label_to_index = label_dict
        self.index_to_label = {v: k for k, v in label_dict.items()}
        self.model = model
        device = next(model.parameters()).device
        self.device = device
        if loss_weight is not None:                                   # main.py:66-67
            self.loss_weight = torch.tensor(loss_weight, device=device, dtype=torch.float)
        else:                                                         # main.py:69
            self.loss_weight = torch.ones((classes), device=device)
        self.logged = {}

    def log(self, name, value, **kwargs):
        self.logged.setdefault(name, []).append(float(value))

    def training_step(self, batch, batch_idx):                        # main.py:71-80
        individual, inputs, y = batch
        images = inputs["HSI"]
        y_hat = self.model.forward(images)
        loss = F.cross_entropy(y_hat, y, weight=self.loss_weight)
        return loss

    def validation_step(self, batch, batch_idx):                      # main.py:82-94
        individual, inputs, y = batch
        images = inputs["HSI"]
        y_hat = self.model.forward(images)
        loss = F.cross_entropy(y_hat, y, weight=self.loss_weight)
        self.log("val_loss", loss, on_epoch=True)
        return loss

    def configure_optimizers(self):                                   # main.py:135-149
        optimizer = optim.Adam(self.model.parameters(), lr=self.config["lr"])
        scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.75, patience=8, threshold=0.0001,
                                                         threshold_mode='rel', cooldown=0, min_lr=0.0000001, eps=1e-08)
        return {'optimizer': optimizer, 'lr_scheduler': scheduler, "monitor": 'val_loss'}

    def predict(self, inputs):                                        # main.py:152-163
        images = inputs["HSI"]
        if "cuda" == self.device.type:
            images = images.cuda() if torch.is_tensor(images) else [x.cuda() for x in images]
            pred = self.model(images).cpu()
        else:
            pred = self.model(images)
        return pred

    # ---- what Trainer.fit does with the hooks above ----------------------------------------------------------------
    def fit_steps(self, train_batches, val_batch=None):
        """One optimizer step per training batch; after the last one, a validation pass feeds ReduceLROnPlateau.
        Returns the list of training losses."""
        opt = self.configure_optimizers()
        optimizer, scheduler = opt["optimizer"], opt["lr_scheduler"]
        losses = []
        self.model.train()
        for i, batch in enumerate(train_batches):
            optimizer.zero_grad()
            loss = self.training_step(batch, i)
            loss.backward()
            optimizer.step()
            losses.append(float(loss.detach()))
        if val_batch is not None:
            self.model.eval()
            with torch.no_grad():
                val = self.validation_step(val_batch, 0)
            scheduler.step(float(val))
            self.model.train()
        self.optimizer, self.scheduler = optimizer, scheduler
        return losses
