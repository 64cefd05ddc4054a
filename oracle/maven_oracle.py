"""CPU oracle for the CLIP contrastive training step of multimodal-supernovae.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may.  The product path (``maven_b200``) fails loudly when its CUDA library is missing and
never routes through here.

This is a functional (state_dict-driven) restatement of the reference's arithmetic, written
against plain ``torch`` CPU ops in whatever dtype the inputs carry (fp32 for parity with the
reference, fp64 for a tighter truth).  Every function cites the reference lines it restates;
paths are relative to the upstream repository root.

Pinning status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference
modules inside the build container, runs them on seeded inputs and commits inputs, weights and
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays those through this
file.  It also replays the reference's own shipped known-answer vectors (the ``lc-reg`` entries of
``evaluation_metrics/collect_regression_results.pkl`` against ``models/lc_reg/*`` checkpoints).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# A1  TimePositionalEncoding.forward                       src/transformer_utils.py:166-176
# --------------------------------------------------------------------------------------
def time_div_term(d_emb: int, norm: float) -> Tensor:
    """fp32 frequency vector exactly as the reference forms it (:168-170)."""
    return torch.exp(torch.arange(0, d_emb, 2).float() * (-math.log(norm) / d_emb))


def time_positional_encoding(t: Tensor, d_emb: int, norm: float) -> Tensor:
    div = time_div_term(d_emb, norm).to(t.dtype)
    arg = t.unsqueeze(2) * div[None, None, :]
    pe = torch.zeros(t.shape[0], t.shape[1], d_emb, dtype=t.dtype)
    pe[:, :, 0::2] = torch.sin(arg)
    pe[:, :, 1::2] = torch.cos(arg)
    return pe


# --------------------------------------------------------------------------------------
# A3  SelfAttention.forward                                 src/transformer_utils.py:36-89
# --------------------------------------------------------------------------------------
def self_attention(sd: SD, p: str, x: Tensor, mask: Optional[Tensor], heads: int) -> Tensor:
    b, t, e = x.shape
    h, s = heads, e // heads
    k = F.linear(x, sd[p + "tokeys.weight"])
    q = F.linear(x, sd[p + "toqueries.weight"])
    v = F.linear(x, sd[p + "tovalues.weight"])

    def fold(z):
        return z.view(b, t, h, s).transpose(1, 2).contiguous().view(b * h, t, s)

    k, q, v = fold(k), fold(q), fold(v)
    q = q / (e ** (1 / 4))          # both scaled by emb**0.25 -> scores / sqrt(emb)   (:63-64)
    k = k / (e ** (1 / 4))
    dot = torch.bmm(q, k.transpose(1, 2))
    if mask is not None:            # keys only, fill value -1e7 (not -inf)             (:71-77)
        m = mask.unsqueeze(1).unsqueeze(2).expand(b, h, 1, t).reshape(b * h, 1, t)
        dot = dot.masked_fill(~m, float("-1e7"))
    dot = F.softmax(dot, dim=2)
    out = torch.bmm(dot, v).view(b, h, t, s).transpose(1, 2).contiguous().view(b, t, s * h)
    return F.linear(out, sd[p + "unifyheads.weight"], sd[p + "unifyheads.bias"])


# --------------------------------------------------------------------------------------
# A4  TransformerBlock.forward (post-norm, ReLU FFN)      src/transformer_utils.py:109-116
# A5  Transformer.forward                                  src/transformer_utils.py:143-153
# Dropout (:112,:115,:147) is the identity unless the caller supplies the multiplicative keep
# factors `drop` (0 or 1/(1-p), shape (B,T,E)) for each site -- torch's RNG stream cannot be
# matched by a fused kernel, so dropout parity is defined on a GIVEN mask.
# --------------------------------------------------------------------------------------
def transformer_block(sd: SD, p: str, x: Tensor, mask: Optional[Tensor], heads: int,
                      drop1: Optional[Tensor] = None, drop2: Optional[Tensor] = None) -> Tensor:
    e = x.shape[-1]
    a = self_attention(sd, p + "attention.", x, mask, heads)
    x = F.layer_norm(a + x, (e,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    if drop1 is not None:
        x = x * drop1
    f = F.linear(x, sd[p + "ff.0.weight"], sd[p + "ff.0.bias"])
    f = F.linear(torch.relu(f), sd[p + "ff.2.weight"], sd[p + "ff.2.bias"])
    x = F.layer_norm(f + x, (e,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    return x if drop2 is None else x * drop2


def transformer(sd: SD, p: str, x: Tensor, mask: Optional[Tensor], heads: int, depth: int,
                drop_scales: Optional[Sequence[Tensor]] = None) -> Tensor:
    """drop_scales: 1 + 2*depth keep factors in site order [input, (norm1, norm2) per layer]."""
    if drop_scales is not None:
        x = x * drop_scales[0]
    for i in range(depth):
        d1 = None if drop_scales is None else drop_scales[1 + 2 * i]
        d2 = None if drop_scales is None else drop_scales[2 + 2 * i]
        x = transformer_block(sd, f"{p}tblocks.{i}.", x, mask, heads, d1, d2)
    return x


# --------------------------------------------------------------------------------------
# A2 + A6 + A7(first half)  TransformerWithTimeEmbeddings.forward
#                                                          src/transformer_utils.py:209-253
# --------------------------------------------------------------------------------------
def band_index(T: int, nband: int) -> Tensor:
    """int64 band id per position: first T//nband positions band 0, next band 1, ... (:220-227)."""
    return torch.arange(nband, dtype=torch.int64).repeat_interleave(T // nband)


def seq_embed(sd: SD, p: str, x: Tensor, t: Tensor, emb: int, nband: int, time_norm: float) -> Tensor:
    pe = time_positional_encoding(t, emb, time_norm)
    h = F.linear(x, sd[p + "embedding_mag.weight"], sd[p + "embedding_mag.bias"]) + pe
    if nband > 1:
        h = h + sd[p + "band_emb.weight"][band_index(x.shape[1], nband)].unsqueeze(0)
    return h


def seq_encoder(sd: SD, p: str, x: Tensor, t: Tensor, mask: Tensor, *, emb: int, heads: int,
                depth: int, nband: int = 1, agg: str = "mean", time_norm: float = 10000.0,
                drop_scales: Optional[Sequence[Tensor]] = None) -> Tensor:
    """x (B,T,1), t (B,T), mask (B,T) bool -> (B,n_out)   [or (B,T,emb) for agg='pretraining']."""
    h = seq_embed(sd, p, x, t, emb, nband, time_norm)
    h = transformer(sd, p + "transformer.", h, mask, heads, depth, drop_scales)
    h = h * mask[:, :, None]
    if agg == "mean":
        h = h.sum(dim=1) / mask.sum(dim=1)[:, None]
    elif agg == "max":
        h = h.max(dim=1)[0]
    elif agg == "attn":
        # nn.MultiheadAttention(emb, 2 heads, batch_first) with a learnable query and NO key
        # padding mask: zeroed padded rows take part as k=b_k, v=b_v (:241-247).
        B, T, E = h.shape
        nh, hd = 2, E // 2
        w, bb = sd[p + "agg_attn.in_proj_weight"], sd[p + "agg_attn.in_proj_bias"]
        q = F.linear(sd[p + "query"].view(1, 1, E).expand(B, 1, E), w[:E], bb[:E])
        k = F.linear(h, w[E:2 * E], bb[E:2 * E])
        v = F.linear(h, w[2 * E:], bb[2 * E:])
        q = q.view(B, 1, nh, hd).transpose(1, 2)
        k = k.view(B, T, nh, hd).transpose(1, 2)
        v = v.view(B, T, nh, hd).transpose(1, 2)
        a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, E)
        h = F.linear(o, sd[p + "agg_attn.out_proj.weight"], sd[p + "agg_attn.out_proj.bias"])
    elif agg == "pretraining":
        return h
    else:                           # reference falls through to projection on (B,T,E)
        pass
    return F.linear(h, sd[p + "projection.weight"], sd[p + "projection.bias"])


# --------------------------------------------------------------------------------------
# N4  masked-light-curve pretraining objective            src/models_pretraining.py:147-226
# net = TransformerWithTimeEmbeddings(agg="pretraining") -> last_layer Linear(emb, 1) -> squeeze; the loss is
# nn.MSELoss()(x[mask_pred], x_pred[mask_pred]) with the run selected by mask_in zeroed in the input and the encoder
# attending over the whole padding mask (:201-203).  The masks are inputs here (the reference draws them on the host).
# --------------------------------------------------------------------------------------
def masked_lc_pred(sd: SD, x: Tensor, t: Tensor, padding_mask: Tensor, mask_in: Tensor, *, emb: int, heads: int, depth: int,
                   nband: int = 1, time_norm: float = 10000.0) -> Tensor:
    xm = x.clone()
    xm[~mask_in] = 0
    h = seq_encoder(sd, "net.", xm[..., None], t, padding_mask, emb=emb, heads=heads, depth=depth, nband=nband,
                    agg="pretraining", time_norm=time_norm)
    return F.linear(h, sd["last_layer.weight"], sd["last_layer.bias"]).squeeze(2)


def masked_lc_loss(sd: SD, x: Tensor, t: Tensor, padding_mask: Tensor, mask_in: Tensor, mask_pred: Tensor, **kw) -> Tensor:
    x_pred = masked_lc_pred(sd, x, t, padding_mask, mask_in, **kw)
    return F.mse_loss(x_pred[mask_pred], x[mask_pred])


# --------------------------------------------------------------------------------------
# A8  ConvMixer.forward                                   src/models_multimodal.py:38-95
# BatchNorm in train mode uses biased batch variance to normalise and updates running stats
# with momentum 0.1 / unbiased variance.  `stats_out`, when given, receives the new buffers.
# --------------------------------------------------------------------------------------
def _bn(sd: SD, p: str, x: Tensor, training: bool, stats_out: Optional[SD]) -> Tensor:
    w, b = sd[p + "weight"], sd[p + "bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            n = x.numel() / x.shape[1]
            stats_out[p + "running_mean"] = 0.9 * sd[p + "running_mean"] + 0.1 * mean.detach()
            stats_out[p + "running_var"] = 0.9 * sd[p + "running_var"] + 0.1 * var.detach() * n / (n - 1)
            stats_out[p + "num_batches_tracked"] = sd[p + "num_batches_tracked"] + 1
    else:
        mean, var = sd[p + "running_mean"], sd[p + "running_var"]
    xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + 1e-5)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def convmixer(sd: SD, p: str, x: Tensor, *, depth: int, kernel_size: int, patch_size: int,
              training: bool = True, stats_out: Optional[SD] = None, drop_scales: Optional[dict] = None) -> Tensor:
    """drop_scales (optional): keep factors (0 or 1/(1-p)) of the nn.Dropout layers (src/models_multimodal.py:62-77,85-87),
    keyed by site: s in 1..2*depth follows BatchNorm s (NCHW feature-map shape), 1+2*depth follows the head's GELU
    ((B, 1024)).  Dropout parity is defined on a GIVEN mask, like the sequence encoder's."""
    ds = drop_scales or {}
    dim = sd[p + "net.0.weight"].shape[0]
    x = F.conv2d(x, sd[p + "net.0.weight"], None, stride=patch_size)
    x = _bn(sd, p + "net.2.", F.gelu(x), training, stats_out)
    for d in range(depth):
        q = f"{p}net.{3 + d}."
        y = F.conv2d(x, sd[q + "0.fn.0.weight"], sd[q + "0.fn.0.bias"], groups=dim, padding="same")
        y = _bn(sd, q + "0.fn.2.", F.gelu(y), training, stats_out)
        x = (y * ds[2 * d + 1] if 2 * d + 1 in ds else y) + x
        y = F.conv2d(x, sd[q + "1.weight"], sd[q + "1.bias"])
        x = _bn(sd, q + "3.", F.gelu(y), training, stats_out)
        if 2 * d + 2 in ds:
            x = x * ds[2 * d + 2]
    x = x.mean(dim=(2, 3))
    x = F.gelu(F.linear(x, sd[p + "projection.2.weight"], sd[p + "projection.2.bias"]))
    if 2 * depth + 1 in ds:
        x = x * ds[2 * depth + 1]
    return F.linear(x, sd[p + "projection.5.weight"], sd[p + "projection.5.bias"])


# --------------------------------------------------------------------------------------
# meta encoder (N4)                          src/models_multimodal.py:295-304, 834-857
# --------------------------------------------------------------------------------------
def meta_encoder(sd: SD, classification: Tensor, redshift: Tensor, input_dim: int, num_layers: int,
                 drop_scales: Optional[Sequence[Tensor]] = None) -> Tensor:
    """drop_scales (optional): one keep-factor tensor (B, hidden) per hidden layer (Linear -> ReLU -> Dropout, :842-853)."""
    x = torch.cat([sd["class_emb.weight"][classification.long()],
                   redshift.unsqueeze(1).repeat(1, input_dim // 2)], dim=-1)
    for i in range(num_layers):
        x = torch.relu(F.linear(x, sd[f"meta_encoder.layers.{3 * i}.weight"], sd[f"meta_encoder.layers.{3 * i}.bias"]))
        if drop_scales is not None:
            x = x * drop_scales[i]
    j = 3 * num_layers
    return F.linear(x, sd[f"meta_encoder.layers.{j}.weight"], sd[f"meta_encoder.layers.{j}.bias"])


# --------------------------------------------------------------------------------------
# A10 / A11  clip_loss, clip_loss_multimodal                       src/loss.py:14-65
# --------------------------------------------------------------------------------------
def clip_loss(embs1: Tensor, embs2: Tensor, logit_scale: Tensor, logit_bias: Tensor) -> Tensor:
    logits = (embs2 @ embs1.T) * logit_scale.exp() + logit_bias
    l_row = -torch.log_softmax(logits, dim=1).diag()
    l_col = -torch.log_softmax(logits, dim=0).diag()
    n = min(len(embs1), len(embs2))
    return (l_row.sum() / n + l_col.sum() / n) / 2


def clip_loss_multimodal(embeddings: Sequence[Tensor], logit_scales: Tensor, logit_biases: Tensor) -> Tensor:
    m = len(embeddings)
    npair = m * (m - 1) // 2
    if logit_scales.dim() == 0:
        logit_scales = logit_scales.repeat(npair)
    if logit_biases.dim() == 0:
        logit_biases = logit_biases.repeat(npair)
    total, c = 0, 0
    for i in range(m - 1):
        for j in range(i + 1, m):
            total = total + clip_loss(embeddings[i], embeddings[j], logit_scales[c], logit_biases[c])
            c += 1
    return total


def clip_loss_grads_closed_form(e1: Tensor, e2: Tensor, logit_scale: Tensor, logit_bias: Tensor):
    """Closed-form backward of A10: G=(P_row+P_col-2I)/(2N)."""
    n = e1.shape[0]
    s = logit_scale.exp()
    z = (e2 @ e1.T) * s + logit_bias
    g = (torch.softmax(z, 1) + torch.softmax(z, 0) - 2 * torch.eye(n, dtype=z.dtype)) / (2 * n)
    return s * g.T @ e2, s * g @ e1, (g * (z - logit_bias)).sum(), g.sum()


# --------------------------------------------------------------------------------------
# A12  LightCurveImageCLIP.forward / training_step   src/models_multimodal.py:203-366
# `cfg` keys: combinations (iterable), transformer_kwargs, transformer_spectral_kwargs,
# conv_kwargs, meta_kwargs, nband, regression, classification, n_classes.
# --------------------------------------------------------------------------------------
def _l2n(x: Tensor) -> Tensor:
    return x / x.norm(dim=-1, keepdim=True)      # no epsilon  (:279,286,293)


def _enc_kwargs(kw: dict, nband: int) -> dict:
    return dict(emb=kw["emb"], heads=kw["heads"], depth=kw["depth"], nband=nband,
                agg=kw.get("agg", "mean"), time_norm=kw.get("time_norm", 10000.0))


def model_forward(sd: SD, cfg: dict, batch: Sequence[Optional[Tensor]], *, training: bool = True,
                  stats_out: Optional[SD] = None):
    x_img, x_lc, t_lc, mask_lc, x_sp, t_sp, mask_sp, redshift, cls = batch
    comb = set(cfg["combinations"])
    head = cfg.get("regression", False) or cfg.get("classification", False)
    norm = (lambda z: z) if head else _l2n
    out: List[Tensor] = []
    if "host_galaxy" in comb:      # fixed order img, lc, sp, meta   (:260-273)
        ck = cfg["conv_kwargs"]
        z = convmixer(sd, "image_encoder.", x_img, depth=ck["depth"], kernel_size=ck["kernel_size"],
                      patch_size=ck["patch_size"], training=training, stats_out=stats_out)
        out.append(norm(F.linear(z, sd["image_projection.weight"], sd["image_projection.bias"])))
    if "lightcurve" in comb:
        z = seq_encoder(sd, "lightcurve_encoder.", x_lc[..., None], t_lc, mask_lc,
                        **_enc_kwargs(cfg["transformer_kwargs"], cfg.get("nband", 1)))
        out.append(norm(F.linear(z, sd["lightcurve_projection.weight"], sd["lightcurve_projection.bias"])))
    if "spectral" in comb:
        z = seq_encoder(sd, "spectral_encoder.", x_sp[..., None], t_sp, mask_sp,
                        **_enc_kwargs(cfg["transformer_spectral_kwargs"], 1))
        out.append(norm(F.linear(z, sd["spectral_projection.weight"], sd["spectral_projection.bias"])))
    if "meta" in comb:
        mk = cfg["meta_kwargs"]
        out.append(norm(meta_encoder(sd, cls, redshift, mk["input_dim"], mk["num_layers"])))
    if head:
        return F.linear(torch.cat(out, dim=-1), sd["linear.weight"], sd["linear.bias"])
    return out


def training_loss(sd: SD, cfg: dict, batch, *, stats_out: Optional[SD] = None) -> Tensor:
    """training_step's loss (:312-366), softmax CLIP / weighted CE / MSE."""
    x = model_forward(sd, cfg, batch, training=True, stats_out=stats_out)
    if cfg.get("regression", False):
        return F.mse_loss(x.squeeze(), batch[7])
    if cfg.get("classification", False):
        nc = cfg.get("n_classes", 5)
        w = {5: [0.3, 0.08, 1.0, 0.01, 0.2], 3: [0.33, 0.06, 1.0]}.get(nc, [1.0] * nc)
        return F.cross_entropy(x.squeeze(), batch[8].long(), weight=torch.tensor(w, dtype=x.dtype))
    return clip_loss_multimodal(x, sd["logit_scale"], sd["logit_bias"]).mean()


# --------------------------------------------------------------------------------------
# A13  torch.optim.RAdam (coupled L2 weight decay), restated from the published algorithm
# (Liu et al. 2020, as implemented by torch.optim.RAdam; the reference calls it at
# src/models_multimodal.py:306-310).  torch itself is the un-vendored third-party dependency
# (requirements.txt: unpinned `torch`; oracle build: torch 2.11.0); pinned in
# tests/test_oracle_golden.py against torch.optim.RAdam.
# --------------------------------------------------------------------------------------
def radam_scalars(step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999):
    """Host-side per-step scalars: (bias_correction1, rect*adaptive-scale or None)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    rho_inf = 2 / (1 - beta2) - 1
    rho_t = rho_inf - 2 * step * (beta2 ** step) / bc2
    rect = None
    if rho_t > 5.0:
        rect = math.sqrt((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t))
    return bc1, bc2, rect


def radam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
               weight_decay: float = 0.0, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """In-place single-tensor RAdam update; returns nothing."""
    if weight_decay != 0:
        g = g + weight_decay * p
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2, rect = radam_scalars(step, lr, beta1, beta2)
    mhat = m / bc1
    if rect is not None:
        adaptive = math.sqrt(bc2) / (v.sqrt() + eps)
        p.add_(mhat * lr * adaptive * rect, alpha=-1.0)
    else:
        p.add_(mhat * lr, alpha=-1.0)


# --------------------------------------------------------------------------------------
# N3  retrieval rank of the true pair (count of sims strictly above the diagonal), the core
# of get_ROC_data                                                  src/utils.py:380-426
# --------------------------------------------------------------------------------------
def retrieval_ranks(e1: Tensor, e2: Tensor) -> Tensor:
    """rank[j] = #{i : cos(e1_i, e2_j) > cos(e1_j, e2_j)} -- position of the true partner of
    source e2_j in the descending argsort over e1 (ties aside)."""
    sim = F.normalize(e1, dim=-1) @ F.normalize(e2, dim=-1).T
    return (sim > sim.diag()[None, :]).sum(dim=0)


def roc_data(e1: Tensor, e2: Tensor, n_thr: int = 100):
    """get_ROC_data (src/utils.py:380-413): thresholds = linspace(0,1,100); source j is "right" at threshold t when its true
    partner is among the first int(t*N) entries of its descending similarity ranking, i.e. rank_j < int(t*N).
    -> (thresholds float64 [n_thr], fraction_correct float64 [n_thr])."""
    import numpy as np
    N = e1.shape[0]
    thresholds = np.linspace(0, 1, n_thr)
    ranks = retrieval_ranks(e1, e2).numpy()
    k = np.array([int(t * N) for t in thresholds])
    return thresholds, (ranks[None, :] < k[:, None]).sum(axis=1) / N


def auc(e1: Tensor, e2: Tensor) -> float:
    """get_AUC (src/utils.py:416-426): trapezoid area under the curve above."""
    import numpy as np
    thr, frac = roc_data(e1, e2)
    return float(np.sum((frac[1:] + frac[:-1]) * np.diff(thr)) / 2.0)


# --------------------------------------------------------------------------------------
# N1  NoisyDataLoader.__iter__ on GIVEN random tensors             src/dataloader.py:88-287
# (the reference draws rand_like(images), randn_like(mag), randn_like(spec), randint(0,4) from
# torch's global generator; tests/golden/make_golden_noisy.py replays those draws against the
# unmodified class, so parity is defined on the recovered tensors)
# --------------------------------------------------------------------------------------
def noisy_images(imgs: Tensor, img_u: Tensor, rot_k: Tensor, max_noise_intensity: float) -> Tensor:
    noise_range = max_noise_intensity * torch.std(imgs)                       # :93
    noisy = imgs + (2 * img_u - 1) * noise_range                              # :96-98
    # RandomRotation([a, a]) with a = 90*k is exactly torch.rot90(img, k, (1, 2))  (:101-109; checked by make_golden_noisy.py)
    return torch.stack([torch.rot90(noisy[i], int(rot_k[i]), (1, 2)) for i in range(noisy.shape[0])])


def noisy_seq(x: Tensor, err: Tensor, noise: Tensor, noise_level: float) -> Tensor:
    return x + noise * err * noise_level                                      # :125, :136
