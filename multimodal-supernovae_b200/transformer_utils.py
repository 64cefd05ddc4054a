"""Drop-in mirror of the reference's src/transformer_utils.py for the CLIP training path.

Same class names, constructor arguments, forward signatures and state_dict keys; the arithmetic
runs in libmaven_sm100.so.  `TransformerWithTimeEmbeddings.forward` is one fused library call (pack ragged tokens,
embed, depth x block, pool, project); the finer-grained classes are built from the per-op kernels and compute every
(also padded) position exactly like the reference, because callers of those see the padded rows.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import SeqCfg

_AGG = {"mean": _lib.MVN_AGG_MEAN, "max": _lib.MVN_AGG_MAX, "pretraining": _lib.MVN_AGG_NONE, "attn": _lib.MVN_AGG_NONE}
_PREC = {"fp32": 0, "tf32": 1, "fused": 2}


def _prec_of(module) -> int:
    return _PREC[getattr(module, "precision", "fp32")]


def set_precision(module: nn.Module, precision: str) -> nn.Module:
    """precision in {'fp32','tf32','fused'}: arithmetic tier of every maven_b200 submodule (storage stays fp32).
    'fused' = the tf32 arithmetic with the fused block kernels of the whole-encoder call (per-layer intermediates stay on chip);
    modules built from the per-op kernels treat it like 'tf32'."""
    if precision not in _PREC:
        raise ValueError(f"precision must be one of {list(_PREC)}, got {precision!r}")
    for m in module.modules():
        m.precision = precision
    return module


class SelfAttention(nn.Module):
    """reference: src/transformer_utils.py:8-89."""

    def __init__(self, emb, heads=2):
        super().__init__()
        assert emb % heads == 0, f"Embedding dimension ({emb}) should be divisible by nr. of heads ({heads})"
        self.emb = emb
        self.heads = heads
        self.tokeys = nn.Linear(emb, emb, bias=False)
        self.toqueries = nn.Linear(emb, emb, bias=False)
        self.tovalues = nn.Linear(emb, emb, bias=False)
        self.unifyheads = nn.Linear(emb, emb)

    def heads_out(self, x, mask=None):
        """softmax(QK^T/sqrt(emb) with key padding mask) V, heads concatenated -- everything before unifyheads."""
        b, t, e = x.size()
        assert e == self.emb, f"Input embedding dim ({e}) should match layer embedding dim ({self.emb})"
        wqkv = torch.cat([self.toqueries.weight, self.tokeys.weight, self.tovalues.weight], dim=0)
        qkv = ops.linear(x, wqkv, None, _prec_of(self))
        return ops.AttentionFn.apply(qkv, mask, b, t, e, self.heads)

    def forward(self, x, mask=None):
        return ops.linear(self.heads_out(x, mask), self.unifyheads.weight, self.unifyheads.bias, _prec_of(self))


class TransformerBlock(nn.Module):
    """reference: src/transformer_utils.py:92-116 (post-norm, ReLU feed-forward)."""

    def __init__(self, emb, heads, ff_hidden_mult=6, dropout=0.0):
        super().__init__()
        self.attention = SelfAttention(emb, heads=heads)
        self.norm1 = nn.LayerNorm(emb)
        self.norm2 = nn.LayerNorm(emb)
        self.ff = nn.Sequential(nn.Linear(emb, ff_hidden_mult * emb), nn.ReLU(), nn.Linear(ff_hidden_mult * emb, emb))
        self.do = nn.Dropout(dropout)

    def forward(self, x, mask=None):
        p = float(self.do.p) if self.training else 0.0
        seed = ops.next_dropout_seed(self) if p > 0.0 else 0    # sites: 0 after norm1, 1 after norm2 (:112,115)
        prec = _prec_of(self)
        a = self.attention.heads_out(x, mask)
        u = self.attention.unifyheads
        x = ops.LinearResLNFn.apply(a, u.weight, u.bias, x, self.norm1.weight, self.norm1.bias, self.norm1.eps, prec)
        x = ops.dropout(x, p, seed, 0)
        x = ops.FFNResLNFn.apply(x, self.ff[0].weight, self.ff[0].bias, self.ff[2].weight, self.ff[2].bias,
                                 self.norm2.weight, self.norm2.bias, self.norm2.eps, prec)
        return ops.dropout(x, p, seed, 1)


class Transformer(nn.Module):
    """reference: src/transformer_utils.py:119-153."""

    def __init__(self, emb, heads, depth, ff_hidden_mult=4, dropout=0.0):
        super().__init__()
        self.tblocks = nn.ModuleList([TransformerBlock(emb=emb, heads=heads, ff_hidden_mult=ff_hidden_mult, dropout=dropout)
                                      for _ in range(depth)])
        self.do = nn.Dropout(dropout)
        self.emb, self.heads, self.depth, self.ff_hidden_mult = emb, heads, depth, ff_hidden_mult

    def forward(self, x, mask=None):
        p = float(self.do.p) if self.training else 0.0
        x = ops.dropout(x, p, ops.next_dropout_seed(self) if p > 0.0 else 0, 0)      # input dropout (:147)
        for tblock in self.tblocks:
            x = tblock(x, mask)
        return x


class TimePositionalEncoding(nn.Module):
    """reference: src/transformer_utils.py:156-176."""

    def __init__(self, d_emb, norm=10000.0):
        super().__init__()
        self.d_emb = d_emb
        self.norm = norm
        self._div = None

    def div_term(self, device) -> torch.Tensor:
        # the reference's own expression, evaluated once in fp32 on the host (parity trap B-1) then cached per device
        if self._div is None or self._div.device != device:
            self._div = torch.exp(torch.arange(0, self.d_emb, 2).float() * (-math.log(self.norm) / self.d_emb)).to(device)
        return self._div

    def forward(self, t):
        L = _lib.lib()
        t = ops._req(t, "t")
        B, T = t.shape
        E = self.d_emb
        dev = t.device
        cu = torch.empty(B + 1, dtype=torch.int32, device=dev)
        tok = torch.empty(B * T, dtype=torch.int32, device=dev)
        kv = torch.empty(B * T, dtype=torch.uint8, device=dev)
        _lib.check(L.mvn_pack_plan(None, B, T, 0, ops._p(cu), ops._p(tok), ops._p(kv), ops._stream()), "pack_plan")
        zero = torch.zeros(E, dtype=torch.float32, device=dev)
        out = torch.empty(B, T, E, dtype=torch.float32, device=dev)
        xz = torch.zeros_like(t)
        _lib.check(L.mvn_embed_fwd(ops._p(xz), ops._p(t), ops._p(cu), ops._p(tok), ops._p(self.div_term(dev)), ops._p(zero), ops._p(zero),
                                   None, B, T, E, 1, ops._p(out), ops._stream()), "embed_fwd")
        ops._count(5)
        return out


class TransformerWithTimeEmbeddings(nn.Module):
    """reference: src/transformer_utils.py:179-253.  forward(x (B,T,1), t (B,T), mask (B,T) bool) -> (B,n_out)."""

    def __init__(self, n_out, nband=1, agg="mean", time_norm=10000.0, **kwargs):
        super().__init__()
        self.agg = agg
        self.nband = nband
        self.n_out = n_out
        emb = kwargs["emb"]
        self.embedding_mag = nn.Linear(in_features=1, out_features=emb)
        self.embedding_t = TimePositionalEncoding(emb, time_norm)
        self.transformer = Transformer(**kwargs)
        if nband > 1:
            self.band_emb = nn.Embedding(nband, emb)
        self.projection = nn.Linear(emb, n_out)
        if self.agg == "attn":
            self.query = nn.Parameter(torch.rand(emb))
            self.agg_attn = nn.MultiheadAttention(embed_dim=emb, num_heads=2, dropout=0.0, batch_first=True)
        self._own_group: Optional[ops.FlatParams] = None

    # ---- flat parameter layout (must match mvn_seq_cfg's documented order) -----------------------------
    def core_params(self):
        """encoder parameters in the library's flat order, without the projection head."""
        ps = [self.embedding_mag.weight, self.embedding_mag.bias]
        if self.nband > 1:
            ps.append(self.band_emb.weight)
        for blk in self.transformer.tblocks:
            a = blk.attention
            ps += [a.toqueries.weight, a.tokeys.weight, a.tovalues.weight, a.unifyheads.weight, a.unifyheads.bias,
                   blk.norm1.weight, blk.norm1.bias, blk.ff[0].weight, blk.ff[0].bias, blk.ff[2].weight, blk.ff[2].bias,
                   blk.norm2.weight, blk.norm2.bias]
        return ps

    def head_params(self):
        return [self.projection.weight, self.projection.bias]

    def make_cfg(self, agg_code: int, enc_dim: int, normalize: bool) -> SeqCfg:
        tr = self.transformer
        p = float(tr.do.p) if self.training else 0.0
        seed = ops.next_dropout_seed(self) if p > 0.0 else 0
        return SeqCfg(B=0, T=0, E=tr.emb, H=tr.heads, depth=tr.depth, nband=self.nband, n_out=self.n_out, enc_dim=enc_dim,
                      agg=agg_code, normalize=1 if normalize else 0, prec=_prec_of(self), ff_mult=tr.ff_hidden_mult,
                      ln_eps=1e-5, dropout_p=p, seed=seed)

    def run_fused(self, x, t, mask, *, group: ops.FlatParams, pidx, extra_params, agg_code: int, enc_dim: int, normalize: bool,
                  gbuf=None, goff=0):
        """One library call for the whole encoder.  `group` holds the parameters (core [+head [+modality proj]]) as
        group.params[pidx[0]:pidx[1]] in flat order; `extra_params` are those parameter objects (for autograd)."""
        if x.dim() == 3:
            x = x[..., 0]
        if mask is None:
            mask = torch.ones(x.shape, dtype=torch.bool, device=x.device)
        flat = group.ensure()
        off = group.offsets[pidx[0]]
        count = group.offsets[pidx[1]] - off
        cfg = self.make_cfg(agg_code, enc_dim, normalize)
        call = ops.SeqCall(cfg, flat, off, count, group, pidx, self.embedding_t.div_term(x.device), gbuf, goff)
        return ops.SeqEncoderFn.apply(x, t, mask, call, *extra_params)

    def _standalone_group(self, with_head: bool):
        ps = self.core_params() + (self.head_params() if with_head else [])
        g = self._own_group
        if g is None or len(g.params) != len(ps) or any(a is not b for a, b in zip(g.params, ps)):
            g = ops.FlatParams(ps)
            self._own_group = g
        return g, ps

    def forward(self, x, t, mask=None):
        agg_code = _AGG.get(self.agg)
        if agg_code is None:
            raise ValueError(f"unknown agg {self.agg!r}")
        if self.agg in ("mean", "max"):
            g, ps = self._standalone_group(True)
            return self.run_fused(x, t, mask, group=g, pidx=(0, len(ps)), extra_params=ps, agg_code=agg_code, enc_dim=0, normalize=False)
        g, ps = self._standalone_group(False)
        tokens = self.run_fused(x, t, mask, group=g, pidx=(0, len(ps)), extra_params=ps, agg_code=_lib.MVN_AGG_NONE, enc_dim=0, normalize=False)
        if self.agg == "pretraining":
            return tokens
        pooled = self._attn_pool(tokens, mask)
        return ops.linear(pooled, self.projection.weight, self.projection.bias, 0)

    attn_pool_closed_form = True        # False: the per-op path (k|v projection of all B*T tokens), kept for shapes the kernel does not cover and for A/B

    def _attn_pool(self, tokens, mask=None):
        """agg='attn' (:241-247): nn.MultiheadAttention(emb, 2 heads) with a learnable query over the zero-padded
        tokens and NO key mask.  One query per sequence."""
        B, T, E = tokens.shape
        H = self.agg_attn.num_heads
        w, b = self.agg_attn.in_proj_weight, self.agg_attn.in_proj_bias
        if self.attn_pool_closed_form and mask is not None and ops.attn_pool_supported(T, E, H):
            return ops.AttnPoolFn.apply(tokens, mask, self.query, w, b, self.agg_attn.out_proj.weight, self.agg_attn.out_proj.bias, H)
        q = ops.linear(self.query.view(1, E), w[:E], b[:E], 0)                       # (1,E) same for every sequence
        kv = ops.linear(tokens, w[E:], b[E:], 0)                                     # (B,T,2E) = k|v
        o = ops.QueryPoolFn.apply(q, kv, B, T, E, H)                                 # (B,E)
        return ops.linear(o, self.agg_attn.out_proj.weight, self.agg_attn.out_proj.bias, 0)
