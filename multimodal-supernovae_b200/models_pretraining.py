"""Drop-in mirror of the masked-light-curve pretraining objective of the reference's src/models_pretraining.py
(`get_random_mask`, `get_continous_random_mask`, `MaskedLightCurveEncoder`): the encoder is the fused sequence encoder with
`agg="pretraining"` (per-token outputs), the removable `last_layer` a C-ABI GEMM, and the loss a masked mean-squared error kernel
(`mvn_masked_mse_fwd/bwd`) over the (B, T) grid -- `nn.MSELoss()(x[mask_pred], x_pred[mask_pred])` without the two gathers.

The mask generators stay on the host like the reference's (they draw from Python's `random` / `torch.randperm`); they consume the
random streams in the same order, so a seeded run selects the SAME positions as the reference
(tests/test_host.py::test_pretraining_masks_match_reference_golden)."""
from __future__ import annotations

import random
from typing import Any, Dict, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .models_multimodal import _Base
from .transformer_utils import TransformerWithTimeEmbeddings


def get_random_mask(padding_mask: Tensor, f_mask: float = 0.15) -> Tuple[Tensor, Tensor]:
    """reference: src/models_pretraining.py:17-56.  Per sample, int(n_obs * f_mask) valid positions chosen by one
    `torch.randperm(n_obs)` draw: `mask` = padding mask without them (encoder input), `mask_pred` = only them (loss)."""
    pm = padding_mask.bool()
    mask, mask_pred = pm.clone(), torch.zeros_like(pm)
    pm_cpu = pm.cpu()
    for i in range(pm.shape[0]):
        valid = torch.nonzero(pm_cpu[i]).flatten()
        n_pred = int(valid.numel() * f_mask)
        chosen = valid[torch.randperm(valid.numel())[:n_pred]].to(pm.device)
        mask[i, chosen] = False
        mask_pred[i, chosen] = True
    return mask, mask_pred


def get_continous_random_mask(padding_mask: Tensor, nbands: int, f_mask: float = 0.15) -> Tuple[Tensor, Tensor]:
    """reference: src/models_pretraining.py:59-103.  Per sample and band one contiguous run of int(n_obs * f_mask) slots starting at
    `random.randint(band_start, band_start + n_obs - n_to_mask)` (one draw per (sample, band), in that order): `mask_pred` keeps the
    run's VALID slots, `mask` drops the run from the padding mask."""
    pm = padding_mask.bool()
    B, N = pm.shape
    bandsize = N // nbands
    n_obs = pm[:, :bandsize * nbands].reshape(B, nbands, bandsize).sum(dim=2).cpu().tolist()      # one device read for the whole batch
    lower = torch.empty(B, nbands, dtype=torch.long)
    upper = torch.empty(B, nbands, dtype=torch.long)
    for i in range(B):
        for k in range(nbands):
            n_to_mask = int(n_obs[i][k] * f_mask)
            lo = random.randint(bandsize * k, bandsize * k + n_obs[i][k] - n_to_mask)
            lower[i, k], upper[i, k] = lo, lo + n_to_mask
    pos = torch.arange(N, device=pm.device)
    band = torch.clamp(pos // bandsize, max=nbands - 1)                                           # slots past nbands*bandsize: untouched by either mask
    in_band = pos < bandsize * nbands
    lo = lower.to(pm.device)[:, band]
    up = upper.to(pm.device)[:, band]
    run = (pos[None, :] >= lo) & (pos[None, :] < up) & in_band[None, :]
    mask = pm & ~run
    mask_pred = pm & (run | ~in_band[None, :])
    return mask, mask_pred


class MaskedLightCurveEncoder(_Base):
    """reference: src/models_pretraining.py:106-259.  Same constructor arguments and state_dict keys (`net.*`, `last_layer.*`)."""

    def __init__(self, f_mask: float = 0.2, nband: int = 1,
                 transformer_kwargs: Dict = {"n_out": 1, "emb": 128, "heads": 2, "depth": 4},
                 optimizer_kwargs: Dict = {}, lr_scheduler_kwargs: Dict = {}, lr: float = 1e-3) -> None:
        super().__init__()
        self.nband = nband
        self.optimizer_kwargs = optimizer_kwargs
        self.lr_scheduler_kwargs = lr_scheduler_kwargs
        self.lr = lr
        self.f_mask = f_mask
        self.net = TransformerWithTimeEmbeddings(nband=nband, agg="pretraining", **transformer_kwargs)
        self.last_layer = nn.Linear(transformer_kwargs["emb"], 1)

    def forward(self, x: Tensor, t: Tensor, mask: Tensor = None) -> Tensor:
        tokens = self.net(x[..., None], t, mask)                              # (B, T, emb), zero on padding
        B, T, E = tokens.shape
        y = ops.linear(tokens.reshape(B * T, E), self.last_layer.weight, self.last_layer.bias, 0)
        return y.reshape(B, T)

    def configure_optimizers(self) -> Dict[str, Any]:
        # the reference's torch.optim.RAdam (src/models_pretraining.py:168-181); the parameters are plain nn.Parameters here
        optimizer = torch.optim.RAdam(self.parameters(), lr=self.lr, **self.optimizer_kwargs)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, **self.lr_scheduler_kwargs)
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "monitor": "val_loss", "interval": "epoch", "frequency": 1}}

    def masked_forward(self, x: Tensor, t: Tensor, padding_mask: Tensor, f_mask: float = 0.15):
        """-> (x_pred (B, T), mask_pred (B, T)): the input has the selected run zeroed, the encoder still attends over the whole
        padding mask (src/models_pretraining.py:201-203)."""
        mask_in, mask_pred = get_continous_random_mask(padding_mask, self.nband, f_mask=f_mask)
        x_masked = torch.where(mask_in, x, torch.zeros_like(x))               # index permutation / select only: no arithmetic
        return self(x_masked, t, mask=padding_mask), mask_pred

    def masked_pred(self, x: Tensor, t: Tensor, padding_mask: Tensor, f_mask: float = 0.15) -> Tuple[Tensor, Tensor]:
        """reference: src/models_pretraining.py:183-204 -- (true, predicted) values at the selected positions."""
        x_pred, mask_pred = self.masked_forward(x, t, padding_mask, f_mask)
        return x[mask_pred], x_pred[mask_pred]

    def masked_loss(self, batch) -> Tensor:
        if len(batch) == 3:
            t, x, padding_mask = batch
        else:
            _, x, t, padding_mask, _spec, _freq, _maskspec, _redshift, _ = batch
        x_pred, mask_pred = self.masked_forward(x, t, padding_mask, f_mask=self.f_mask)
        loss, count = ops.MaskedMSEFn.apply(x_pred, x, mask_pred)
        return ops.dp_weighted_mean(loss, count)

    def training_step(self, batch, batch_idx: int) -> Tensor:
        loss = self.masked_loss(batch)
        self.log("train_loss", loss, on_epoch=True, on_step=False, prog_bar=True)
        return loss

    def validation_step(self, batch, batch_idx: int) -> Tensor:
        loss = self.masked_loss(batch)
        self.log("val_loss", loss, on_epoch=True, on_step=False, prog_bar=True)
        return loss
