"""Drop-in mirror of the reference's src/loss.py (softmax CLIP loss only: every reference driver hard-codes
loss="softmax"; the SigLIP variants are out of scope).

clip_loss / clip_loss_multimodal keep the reference signatures.  Forward and backward are streamed kernels
(mvn_clip_loss_fwd/_bwd): the N x N logits are never materialised.  When a data-parallel group is registered with
`maven_b200.ops.set_data_parallel_group`, embeddings are all-gathered so negatives span the global batch and the
returned loss is the global one (identical on every rank).
"""
from __future__ import annotations

import torch

from . import ops


def clip_loss(embs1, embs2, logit_scale=1.0, logit_bias=0.0, image_encoder=None, lightcurve_encoder=None, prec: int = 0):
    """reference: src/loss.py:14-38.  logit_scale is the LOG of the scale (exponentiated inside)."""
    dev = embs1.device
    if not torch.is_tensor(logit_scale):
        logit_scale = torch.tensor(float(logit_scale), device=dev)
    if not torch.is_tensor(logit_bias):
        logit_bias = torch.tensor(float(logit_bias), device=dev)
    return ops.ClipLossFn.apply(embs1, embs2, logit_scale, logit_bias, prec)


def clip_loss_multimodal(embeddings, logit_scales=1.0, logit_biases=0.0, prec: int = 0):
    """reference: src/loss.py:41-65.  Sum over unordered pairs (i<j) in list order."""
    n = len(embeddings)
    npair = n * (n - 1) // 2
    dev = embeddings[0].device
    if not torch.is_tensor(logit_scales):
        logit_scales = torch.tensor(float(logit_scales), device=dev)
    if not torch.is_tensor(logit_biases):
        logit_biases = torch.tensor(float(logit_biases), device=dev)
    if logit_scales.dim() == 0 and logit_biases.dim() == 0 and n >= 2:
        # one scale / bias for every pair (the model's case): all pairs in one op -> one all-gather of the embeddings, one of the LSEs
        return ops.ClipLossMultiFn.apply(logit_scales, logit_biases, prec, *embeddings)
    loss_total = 0
    count = 0
    for i in range(n - 1):
        for j in range(i + 1, n):
            ls = logit_scales if logit_scales.dim() == 0 else logit_scales[count]
            lb = logit_biases if logit_biases.dim() == 0 else logit_biases[count]
            loss_total = loss_total + clip_loss(embeddings[i], embeddings[j], ls, lb, prec=prec)
            count += 1
    assert count == npair
    return loss_total
