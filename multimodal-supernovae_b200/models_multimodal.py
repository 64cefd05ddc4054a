"""Drop-in mirror of the hot-path classes of the reference's src/models_multimodal.py:
Residual, ConvMixer, LightCurveImageCLIP (forward / *_embeddings_with_projection / training_step /
configure_optimizers) and MLP.  Same constructor arguments, forward signatures and state_dict keys, so the shipped
checkpoints load with load_state_dict.  Validation hooks, metric logging and checkpoint plumbing stay with the
reference's Python: they are not on the training-step path.
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from .loss import clip_loss_multimodal
from .transformer_utils import TransformerWithTimeEmbeddings, _prec_of

try:                                    # the reference subclasses pl.LightningModule; keep that when Lightning exists
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:                       # pragma: no cover - Lightning is not in this image
    class _Base(nn.Module):
        def log(self, *args, **kwargs):
            pass


class Residual(nn.Module):
    """reference: src/models_multimodal.py:24-35 (container only; ConvMixer.forward is one fused call)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x) + x


class ConvMixer(nn.Module):
    """reference: src/models_multimodal.py:38-95.  The nn.Sequential layout is kept for state_dict compatibility;
    forward runs the library's ConvMixer kernels (patch conv -> GELU -> BN, depthwise/pointwise mixer layers,
    avg-pool, MLP head)."""

    def __init__(self, dim, depth, channels=1, kernel_size=5, patch_size=8, n_out=128, dropout_prob=0.5):
        super().__init__()
        self.dim, self.depth, self.channels = dim, depth, channels
        self.kernel_size, self.patch_size, self.n_out, self.dropout_prob = kernel_size, patch_size, n_out, dropout_prob
        self.net = nn.Sequential(nn.Conv2d(channels, dim, kernel_size=patch_size, stride=patch_size, bias=False), nn.GELU(), nn.BatchNorm2d(dim))
        for _ in range(depth):
            self.net.append(nn.Sequential(
                Residual(nn.Sequential(nn.Conv2d(dim, dim, kernel_size, groups=dim, padding="same"), nn.GELU(), nn.BatchNorm2d(dim),
                                       nn.Dropout(dropout_prob))),
                nn.Conv2d(dim, dim, kernel_size=1), nn.GELU(), nn.BatchNorm2d(dim), nn.Dropout(dropout_prob)))
        self.projection = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(), nn.Linear(dim, 1024), nn.GELU(),
                                        nn.Dropout(dropout_prob), nn.Linear(1024, n_out))
        self._own_group: Optional[ops.FlatParams] = None

    def bn_layers(self) -> List[nn.BatchNorm2d]:
        out = [self.net[2]]
        for d in range(self.depth):
            out += [self.net[3 + d][0].fn[2], self.net[3 + d][3]]
        return out

    def running_flat(self, device) -> torch.Tensor:
        """BatchNorm running_mean/var of every BN as views of one [nbn, 2, dim] buffer (state_dict names unchanged)."""
        bns = self.bn_layers()
        flat = getattr(self, "_running", None)
        ok = flat is not None and flat.device == device
        if ok:
            base = flat.data_ptr()
            ok = all(bn.running_mean.data_ptr() == base + 4 * (2 * i) * self.dim and bn.running_var.data_ptr() == base + 4 * (2 * i + 1) * self.dim
                     for i, bn in enumerate(bns))
        if not ok:
            flat = torch.empty(len(bns), 2, self.dim, dtype=torch.float32, device=device)
            for i, bn in enumerate(bns):
                flat[i, 0].copy_(bn.running_mean); flat[i, 1].copy_(bn.running_var)
                bn._buffers["running_mean"] = flat[i, 0]
                bn._buffers["running_var"] = flat[i, 1]
            self._running = flat
        return flat

    def core_params(self):
        ps = [self.net[0].weight, self.net[2].weight, self.net[2].bias]
        for d in range(self.depth):
            blk = self.net[3 + d]
            ps += [blk[0].fn[0].weight, blk[0].fn[0].bias, blk[0].fn[2].weight, blk[0].fn[2].bias,
                   blk[1].weight, blk[1].bias, blk[3].weight, blk[3].bias]
        ps += [self.projection[2].weight, self.projection[2].bias, self.projection[5].weight, self.projection[5].bias]
        return ps

    def run_fused(self, x, *, group: ops.FlatParams, pidx, extra_params, enc_dim: int, normalize: bool, gbuf=None, goff=0):
        p_drop = float(self.dropout_prob) if self.training else 0.0
        seed = ops.next_dropout_seed(self) if p_drop > 0.0 else 0
        flat = group.ensure()
        off = group.offsets[pidx[0]]
        count = group.offsets[pidx[1]] - off
        call = ops.ConvCall(self, flat, off, count, group, pidx, enc_dim, normalize, _prec_of(self), gbuf, goff, p_drop, seed)
        return ops.ConvMixerFn.apply(x, call, *extra_params)

    def forward(self, x):
        ps = self.core_params()
        g = self._own_group
        if g is None or len(g.params) != len(ps) or any(a is not b for a, b in zip(g.params, ps)):
            g = ops.FlatParams(ps)
            self._own_group = g
        return self.run_fused(x, group=g, pidx=(0, len(ps)), extra_params=ps, enc_dim=0, normalize=False)


class MLP(nn.Module):
    """reference: src/models_multimodal.py:834-857 (meta encoder / fine-tune heads): Linear+ReLU stacks."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, dropout):
        super().__init__()
        self.input_dim, self.hidden_dim, self.output_dim, self.num_layers, self.dropout = input_dim, hidden_dim, output_dim, num_layers, dropout
        self.layers = nn.ModuleList()
        self.layers.append(nn.Linear(input_dim, hidden_dim)); self.layers.append(nn.ReLU()); self.layers.append(nn.Dropout(dropout))
        for _ in range(num_layers - 1):
            self.layers.append(nn.Linear(hidden_dim, hidden_dim)); self.layers.append(nn.ReLU()); self.layers.append(nn.Dropout(dropout))
        self.layers.append(nn.Linear(hidden_dim, output_dim))

    def forward(self, x):
        p = float(self.dropout) if self.training else 0.0
        seed = ops.next_dropout_seed(self) if p > 0.0 else 0
        lins = [m for m in self.layers if isinstance(m, nn.Linear)]
        for site, lin in enumerate(lins[:-1]):                  # Linear -> ReLU -> Dropout (site = hidden layer index)
            x = ops.dropout(ops.LinearReluFn.apply(x, lin.weight, lin.bias, 0), p, seed, site)
        return ops.linear(x, lins[-1].weight, lins[-1].bias, 0)


_SIDE_STREAMS: Dict[int, List["torch.cuda.Stream"]] = {}       # per device; kept off the module so models stay copyable


class LightCurveImageCLIP(_Base):
    """reference: src/models_multimodal.py:98-366."""

    def __init__(self, enc_dim: int = 128, logit_scale: float = 10.0, nband: int = 1,
                 transformer_kwargs: Dict[str, Any] = {"n_out": 128, "emb": 256, "heads": 2, "depth": 8, "time_norm": 10000.0},
                 transformer_spectral_kwargs: Dict[str, Any] = {"n_out": 128, "emb": 256, "heads": 2, "depth": 8, "time_norm": 10000.0},
                 conv_kwargs: Dict[str, Any] = {"dim": 32, "depth": 8, "channels": 3, "kernel_size": 5, "patch_size": 10, "n_out": 128},
                 meta_kwargs: Dict[str, Any] = {"input_dim": 128, "hidden_dim": 128, "num_layers": 2},
                 combinations: List[str] = ["host_galaxy", "spectral"], optimizer_kwargs: Dict = {}, lr: float = 1e-4,
                 loss: str = "sigmoid", regression: bool = False, classification: bool = False, n_classes: int = 5):
        super().__init__()
        self.lr = lr
        self.optimizer_kwargs = optimizer_kwargs
        self.enc_dim = enc_dim
        self.combinations = set(combinations)
        self.regression = regression
        self.classification = classification
        if self.classification:
            self.n_classes = n_classes
        self.logit_scale = nn.Parameter(torch.tensor(math.log(logit_scale)), requires_grad=True)
        self.logit_bias = nn.Parameter(torch.tensor(-10.0), requires_grad=True)
        if "lightcurve" in self.combinations:
            self.lightcurve_encoder = TransformerWithTimeEmbeddings(nband=nband, **transformer_kwargs)
            self.lightcurve_projection = nn.Linear(transformer_kwargs["n_out"], enc_dim)
        if "spectral" in self.combinations:
            self.spectral_encoder = TransformerWithTimeEmbeddings(nband=1, **transformer_spectral_kwargs)
            self.spectral_projection = nn.Linear(transformer_spectral_kwargs["n_out"], enc_dim)
        if "host_galaxy" in self.combinations:
            self.image_encoder = ConvMixer(**conv_kwargs)
            self.image_projection = nn.Linear(conv_kwargs["n_out"], enc_dim)
        if "meta" in self.combinations:
            self.len_meta_input = meta_kwargs["input_dim"]
            self.class_emb = nn.Embedding(n_classes, self.len_meta_input // 2)
            self.meta_encoder = MLP(output_dim=enc_dim, **meta_kwargs)
        self.loss = loss
        if loss != "softmax" and not (regression or classification):
            raise NotImplementedError(f"maven_b200 builds the softmax CLIP loss only (every reference driver sets loss='softmax'); got loss={loss!r}. "
                                      "Keep the reference module for the SigLIP variant.")
        self.linear_out = 1
        if self.classification:
            self.linear_out = self.n_classes
        if self.regression or self.classification:
            self.linear = nn.Linear(enc_dim * len(self.combinations), self.linear_out)
        self._group: Optional[ops.FlatParams] = None
        self._segments: Dict[str, tuple] = {}
        self._gbuf: Optional[torch.Tensor] = None
        self.y_pred, self.y_true = [], []
        self.track_predictions = False      # the reference's epoch-end metric hooks (out of scope) consume these lists
        # One CUDA stream per modality encoder (see _run_modalities); MVN_CONCURRENT=0 keeps everything on the caller's stream.
        self.concurrent_modalities = os.environ.get("MVN_CONCURRENT", "1") != "0"
        # data parallel: start each encoder's gradient all-reduce as soon as its backward has been enqueued (maven_b200.ops.finish_grad_reduce)
        self.overlap_grad_reduce = os.environ.get("MVN_OVERLAP_GRADS", "1") != "0"

    # ---- flat parameter group over the whole model ---------------------------------------------------------
    # The flat layout is cached between steps (walking ~130 parameters through nn.Module.parameters() costs ~0.7 ms of
    # host time per call, 3 calls per step).  Anything that can replace Parameter objects bumps `_layout_version`.
    def _apply(self, fn, *args, **kwargs):
        self._layout_version = getattr(self, "_layout_version", 0) + 1
        return super()._apply(fn, *args, **kwargs)

    def register_parameter(self, name, param):
        self.__dict__["_layout_version"] = self.__dict__.get("_layout_version", 0) + 1
        return super().register_parameter(name, param)

    def add_module(self, name, module):
        self.__dict__["_layout_version"] = self.__dict__.get("_layout_version", 0) + 1
        return super().add_module(name, module)

    def invalidate_flat_layout(self):
        """Call after replacing a Parameter object of a submodule by hand (module.weight = nn.Parameter(...))."""
        self._layout_version = getattr(self, "_layout_version", 0) + 1

    def flat_group(self) -> ops.FlatParams:
        """All parameters as views of one buffer: fused encoders first (each contiguous in library order), then the
        rest in named_parameters order.  Rebuilt if parameter objects were replaced."""
        ver = getattr(self, "_layout_version", 0)
        if self._group is not None and getattr(self, "_group_version", -1) == ver:
            return self._group
        ordered: List[nn.Parameter] = []
        seg: Dict[str, tuple] = {}

        def add(name, ps):
            seg[name] = (len(ordered), len(ordered) + len(ps))
            ordered.extend(ps)

        def enc_ps(enc, proj):
            if enc.agg in ("mean", "max"):
                return enc.core_params() + enc.head_params() + [proj.weight, proj.bias]
            return enc.core_params()
        if "lightcurve" in self.combinations:
            add("lightcurve", enc_ps(self.lightcurve_encoder, self.lightcurve_projection))
        if "spectral" in self.combinations:
            add("spectral", enc_ps(self.spectral_encoder, self.spectral_projection))
        if "host_galaxy" in self.combinations:
            add("host_galaxy", self.image_encoder.core_params() + [self.image_projection.weight, self.image_projection.bias])
        seen = {id(p) for p in ordered}
        rest = [p for p in self.parameters() if id(p) not in seen]
        add("rest", rest)
        g = self._group
        if g is None or len(g.params) != len(ordered) or any(a is not b for a, b in zip(g.params, ordered)):
            g = ops.FlatParams(ordered)
            self._group = g
        self._segments = seg
        self._group_version = ver
        return g

    def gather_grads(self) -> torch.Tensor:
        """One flat gradient buffer for the whole model: gradients written by the fused ops already live in it; the few
        produced by per-op functions (logit_scale, heads) are staged into their slots and `p.grad` is re-pointed at the
        slot.  Used by FusedRAdam and by the data-parallel all-reduce (one NCCL call over the returned tensor)."""
        g = self.flat_group()
        flat = g.ensure()
        gbuf = self._gbuf
        if gbuf is None or gbuf.numel() != g.total or gbuf.device != flat.device:
            gbuf = torch.zeros_like(flat)
            self._gbuf = gbuf
        gbase = gbuf.data_ptr()
        with torch.no_grad():
            for k, p in enumerate(g.params):
                gr = p.grad
                if gr is None:
                    continue
                o = g.offsets[k]
                if gr.data_ptr() != gbase + 4 * o:
                    v = gbuf[o:o + g.sizes[k]].view(p.shape)
                    v.copy_(gr)
                    p.grad = v
        return gbuf

    def reduce_gradients(self, group=None) -> torch.Tensor:
        """Data-parallel SUM all-reduce of the flat gradient buffer (call after backward, before the optimizer step): the encoder
        segments were launched during the backward (see ops.begin_grad_overlap), this waits for them and reduces the rest."""
        gbuf = self.gather_grads()
        ops.finish_grad_reduce(gbuf, group)
        return gbuf

    def _new_gbuf(self, g: ops.FlatParams, device):
        if torch.is_grad_enabled():
            # zeros, not empty: slots of parameters that receive no gradient this step (logit_scale of a classifier, a frozen
            # backbone) are part of what the data-parallel all-reduce and any norm / clipping over the flat buffer read
            self._gbuf = torch.zeros(g.total, dtype=torch.float32, device=device)
        else:
            self._gbuf = None
        ops.begin_grad_overlap(self._gbuf if (self.training and self.overlap_grad_reduce) else None)
        return self._gbuf

    def _seq_embed(self, which: str, x, t, mask, g, normalize: bool):
        enc = self.lightcurve_encoder if which == "lightcurve" else self.spectral_encoder
        proj = self.lightcurve_projection if which == "lightcurve" else self.spectral_projection
        i0, i1 = self._segments[which]
        if enc.agg in ("mean", "max"):
            return enc.run_fused(x, t, mask, group=g, pidx=(i0, i1), extra_params=g.params[i0:i1], agg_code=_lib.MVN_AGG_MEAN if enc.agg == "mean" else _lib.MVN_AGG_MAX,
                                 enc_dim=self.enc_dim, normalize=normalize, gbuf=self._gbuf, goff=g.offsets[i0])
        # agg in {"attn"}: fused token encoder, then the attention pooling + projections from per-op kernels
        tokens = enc.run_fused(x, t, mask, group=g, pidx=(i0, i1), extra_params=g.params[i0:i1], agg_code=_lib.MVN_AGG_NONE, enc_dim=0,
                               normalize=False, gbuf=self._gbuf, goff=g.offsets[i0])
        if enc.agg == "pretraining":
            z = tokens
        else:
            z = ops.linear(enc._attn_pool(tokens, mask), enc.projection.weight, enc.projection.bias, 0)
        z = ops.linear(z, proj.weight, proj.bias, 0)
        return ops.L2NormFn.apply(z) if normalize else z

    def _img_embed(self, x_img, g, normalize: bool):
        i0, i1 = self._segments["host_galaxy"]
        return self.image_encoder.run_fused(x_img, group=g, pidx=(i0, i1), extra_params=g.params[i0:i1], enc_dim=self.enc_dim,
                                            normalize=normalize, gbuf=self._gbuf, goff=g.offsets[i0])

    def _meta_embed(self, classification, redshift, normalize: bool):
        x_meta = ops.MetaInputFn.apply(self.class_emb.weight, classification, redshift)      # [class_emb[cls] | redshift repeated], one kernel
        z = self.meta_encoder(x_meta)
        return ops.L2NormFn.apply(z) if normalize else z

    def _run_modalities(self, jobs, dev):
        """The modality encoders are independent until the loss, and each is a chain of short persistent kernels whose
        launch / prologue / tail gaps leave the SMs idle.  Enqueue every encoder on its own stream so the block scheduler
        fills one chain's gaps with the other chain's CTAs; autograd replays each encoder's backward on the stream its
        forward ran on, so the backward chains interleave the same way.  Outputs are joined on the caller's stream."""
        concurrent = self.concurrent_modalities and len(jobs) > 1 and dev.type == "cuda"
        # programmatic dependent launches help one kernel chain and hurt two long concurrent ones (measured, csrc/api.cu): off when
        # both sequence encoders run side by side
        n_seq = ("lightcurve" in self.combinations) + ("spectral" in self.combinations)
        _lib.lib().mvn_set_pdl(0 if (concurrent and n_seq >= 2) else 1)
        if not concurrent:
            return [j() for j in jobs]
        main = torch.cuda.current_stream(dev)
        side = _SIDE_STREAMS.setdefault(dev.index if dev.index is not None else torch.cuda.current_device(), [])
        while len(side) < len(jobs) - 1:
            side.append(torch.cuda.Stream(device=dev))
        streams = [main] + side[:len(jobs) - 1]
        for s in streams[1:]:
            s.wait_stream(main)
        outs = []
        for j, s in zip(jobs, streams):
            with torch.cuda.stream(s):
                outs.append(j())
        for o, s in zip(outs[1:], streams[1:]):
            main.wait_stream(s)
            o.record_stream(main)
        return outs

    def forward(self, x_img, x_lc, t_lc, mask_lc, x_sp, t_sp, mask_sp, redshift=None, classification=None):
        head = self.regression or self.classification
        g = self.flat_group()
        dev = next(self.parameters()).device
        g.ensure()
        self._new_gbuf(g, dev)
        jobs = []                      # list order is fixed [host_galaxy, lightcurve, spectral, meta] (reference :260-273)
        if "host_galaxy" in self.combinations:
            jobs.append(lambda: self._img_embed(x_img, g, not head))
        if "lightcurve" in self.combinations:
            jobs.append(lambda: self._seq_embed("lightcurve", x_lc, t_lc, mask_lc, g, not head))
        if "spectral" in self.combinations:
            jobs.append(lambda: self._seq_embed("spectral", x_sp, t_sp, mask_sp, g, not head))
        if "meta" in self.combinations:
            jobs.append(lambda: self._meta_embed(classification, redshift, not head))
        x = self._run_modalities(jobs, dev)
        if head:
            return ops.linear(torch.cat(x, dim=-1) if len(x) > 1 else x[0], self.linear.weight, self.linear.bias, 0)
        return x

    def _single(self, which, *args):
        g = self.flat_group()
        g.ensure()
        self._new_gbuf(g, next(self.parameters()).device)
        if which == "host_galaxy":
            return self._img_embed(args[0], g, True)
        return self._seq_embed(which, args[0], args[1], args[2], g, True)

    def image_embeddings_with_projection(self, x_img):
        return self._single("host_galaxy", x_img)

    def lightcurve_embeddings_with_projection(self, x_lc, t_lc, mask_lc=None):
        return self._single("lightcurve", x_lc, t_lc, mask_lc)

    def spectral_embeddings_with_projection(self, x_lc, t_lc, mask_lc=None):
        return self._single("spectral", x_lc, t_lc, mask_lc)

    def meta_embeddings_with_projection(self, classification, redshift):
        return self._meta_embed(classification, redshift, True)

    def configure_optimizers(self):
        from .optim import FusedRAdam
        optimizer = FusedRAdam(self.parameters(), lr=self.lr, model=self, **self.optimizer_kwargs)
        return {"optimizer": optimizer}

    def training_loss(self, batch):
        """The loss of training_step (src/models_multimodal.py:312-361) without Lightning bookkeeping."""
        x_img, x_lc, t_lc, mask_lc, x_sp, t_sp, mask_sp, redshift, classification = batch
        x = self(x_img, x_lc, t_lc, mask_lc, x_sp, t_sp, mask_sp, redshift, classification)
        if self.regression:
            if self.track_predictions:
                self.y_pred.append(x.flatten()); self.y_true.append(redshift)
            return ops.dp_weighted_mean(ops.MSEFn.apply(x.squeeze(), redshift), float(redshift.numel()))
        if self.classification:
            if self.n_classes == 5:
                w = [0.3, 0.08, 1.0, 0.01, 0.2]
            elif self.n_classes == 3:
                w = [0.33, 0.06, 1.0]
            else:
                w = [1.0] * self.n_classes
            cw = getattr(self, "_class_w", None)
            if cw is None or cw.device != x.device:
                cw = torch.tensor(w, dtype=torch.float32, device=x.device)
                self._class_w = cw
            if self.track_predictions:
                self.y_pred.append(x); self.y_true.append(classification)
            loss, wsum = ops.WeightedCEFn.apply(x.squeeze(), classification, cw)
            return ops.dp_weighted_mean(loss, wsum)
        if self.loss == "softmax":
            return clip_loss_multimodal(x, self.logit_scale, self.logit_bias, prec=min(_prec_of(self), 1))   # 0-dim already: .mean() is the identity
        raise NotImplementedError("maven_b200 builds the softmax CLIP loss only (every reference driver sets loss='softmax')")

    def training_step(self, batch, batch_idx):
        loss = self.training_loss(batch)
        self.log("train_loss", loss, on_epoch=True, on_step=False, prog_bar=True, logger=True)
        return loss

    # ---- validation, CLIP branch (src/models_multimodal.py:418-553): collect the embeddings, log the retrieval AUC of every
    # modality pair at epoch end.  The reference's O(N^2) Python loop is two kernels here (maven_b200.utils.get_AUC).
    def on_validation_start(self) -> None:
        self.embs_list = [[] for _ in range(len(self.combinations))]
        if self.regression or self.classification:
            self.y_pred_val, self.y_true_val = [], []

    def validation_step(self, batch, batch_idx):
        x_img, x_lc, t_lc, mask_lc, x_sp, t_sp, mask_sp, redshift, classification = batch
        if self.regression or self.classification:
            track, self.track_predictions = self.track_predictions, False
            try:
                with torch.no_grad():
                    loss = self.training_loss(batch)
            finally:
                self.track_predictions = track
        else:
            with torch.no_grad():
                x = self(x_img, x_lc, t_lc, mask_lc, x_sp, t_sp, mask_sp, redshift, classification)
                for i in range(len(self.embs_list)):
                    self.embs_list[i].append(x[i])
                loss = clip_loss_multimodal(x, self.logit_scale, self.logit_bias, prec=min(_prec_of(self), 1))
        self.log("val_loss", loss, on_epoch=True, on_step=False, prog_bar=True, logger=True)
        return loss

    def on_validation_epoch_end(self) -> None:
        if self.regression or self.classification or not getattr(self, "embs_list", None):
            return
        from .utils import get_AUC
        embs = [torch.cat(e, dim=0) for e in self.embs_list]
        if len(self.combinations) == 2:
            self.log("AUC_val", get_AUC(embs[0], embs[1]), on_epoch=True, on_step=False, prog_bar=True, logger=True)
        else:
            count = 1
            for i in range(len(self.combinations) - 1):
                for j in range(i + 1, len(self.combinations)):
                    self.log(f"AUC_val{count}", get_AUC(embs[i], embs[j]), on_epoch=True, on_step=False, prog_bar=True, logger=True)
                    count += 1
        self.embs_list = None


class ClipMLP(_Base):
    """reference: src/models_multimodal.py:859-1060 -- a (pre-trained) CLIP backbone with an MLP head for redshift regression or
    classification (finetune_clip.py).  The backbone's embeddings come from the fused encoder calls, the head from the per-op
    kernels; same constructor arguments, state_dict keys (`clip_model.*`, `mlp_model.*`) and training_step as the reference."""

    def __init__(self, clip_model, mlp_kwargs, optimizer_kwargs, lr, combinations=["lightcurve"], regression=True,
                 classification=False, n_classes=5):
        super().__init__()
        enc_dim = 0
        if "lightcurve" in combinations:
            enc_dim += clip_model.lightcurve_projection.out_features
        if "spectral" in combinations:
            enc_dim += clip_model.spectral_projection.out_features
        mlp_kwargs["input_dim"] = enc_dim
        self.clip_model = clip_model
        self.mlp_model = MLP(**mlp_kwargs)
        self.optimizer_kwargs = optimizer_kwargs
        self.lr = lr
        self.combinations = combinations
        self.regression = regression
        self.classification = classification
        self.n_classes = n_classes
        self.y_pred, self.y_true = [], []
        self.track_predictions = False

    def forward(self, x_lc=None, t_lc=None, mask_lc=None, x_sp=None, t_sp=None, mask_sp=None):
        x = []
        if "lightcurve" in self.combinations:
            x.append(self.clip_model.lightcurve_embeddings_with_projection(x_lc, t_lc, mask_lc))
        if "spectral" in self.combinations:
            x.append(self.clip_model.spectral_embeddings_with_projection(x_sp, t_sp, mask_sp))
        x = torch.cat(x, dim=-1) if len(x) > 1 else x[0]
        return self.mlp_model(x)

    def configure_optimizers(self):
        # the reference's torch.optim.RAdam over backbone + head (src/models_multimodal.py:923-927); the parameters are views of
        # flat buffers, which torch's foreach implementation updates in place
        return {"optimizer": torch.optim.RAdam(self.parameters(), lr=self.lr, **self.optimizer_kwargs)}

    def training_loss(self, batch):
        _, x_lc, t_lc, mask_lc, x_sp, t_sp, mask_sp, redshift, classification = batch
        x = self(x_lc, t_lc, mask_lc, x_sp, t_sp, mask_sp)
        if self.regression:
            if self.track_predictions:
                self.y_pred.append(x.flatten()); self.y_true.append(redshift)
            return ops.dp_weighted_mean(ops.MSEFn.apply(x.squeeze(), redshift), float(redshift.numel()))
        if self.classification:
            w = {5: [0.3, 0.08, 1.0, 0.01, 0.2], 3: [0.33, 0.06, 1.0]}.get(self.n_classes, [1.0] * self.n_classes)
            cw = getattr(self, "_class_w", None)
            if cw is None or cw.device != x.device:
                cw = torch.tensor(w, dtype=torch.float32, device=x.device)
                self._class_w = cw
            if self.track_predictions:
                self.y_pred.append(x); self.y_true.append(classification)
            loss, wsum = ops.WeightedCEFn.apply(x.squeeze(), classification, cw)
            return ops.dp_weighted_mean(loss, wsum)
        raise ValueError("ClipMLP needs regression=True or classification=True")

    def training_step(self, batch, batch_idx):
        loss = self.training_loss(batch)
        self.log("train_loss", loss, on_epoch=True, on_step=False, prog_bar=True, logger=True)
        return loss
