"""Drop-in mirror of the retrieval metrics of the reference's src/utils.py (get_ROC_data :380-413, get_AUC :416-426), the
O(N^2) Python loop that LightCurveImageCLIP.on_validation_epoch_end runs over every pair of modalities
(src/models_multimodal.py:527-553).  On the device it is two kernels: the rank of every true partner in its source's
similarity ranking (mvn_retrieval_ranks, streamed similarity -- the N x N matrix is never stored) and the count of ranks
below each of the 100 integer cut-offs (mvn_retrieval_curve); the 100 counts come back to the host for the trapezoid."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import ops


def get_ROC_data(embs1: torch.Tensor, embs2: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    """-> (thresholds, fraction of sources whose true partner is within the top int(threshold * N))."""
    if embs1.shape != embs2.shape or embs1.dim() != 2:
        raise ValueError(f"get_ROC_data: embeddings must both be (N, D), got {tuple(embs1.shape)} and {tuple(embs2.shape)}")
    n = embs1.shape[0]
    thresholds = np.linspace(0, 1, 100)
    k = torch.tensor([int(t * n) for t in thresholds], dtype=torch.int32).to(embs1.device, non_blocking=True)
    ranks = ops.retrieval_ranks(embs1.float(), embs2.float())
    counts = ops.retrieval_curve(ranks, k)
    return thresholds, counts.cpu().numpy().astype(np.float64) / n


def get_AUC(embs1: torch.Tensor, embs2: torch.Tensor) -> float:
    thresholds, fraction_correct = get_ROC_data(embs1, embs2)
    return float(np.sum((fraction_correct[1:] + fraction_correct[:-1]) * np.diff(thresholds)) / 2.0)     # np.trapz
