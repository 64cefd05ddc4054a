"""torch.autograd.Function wrappers over the C ABI (include/maven_sm100.h).

PyTorch supplies device memory, the current stream and autograd bookkeeping; every arithmetic step is a kernel in
libmaven_sm100.so.  CPU tensors raise: there is no fallback path.

Parameter handling: a fused op consumes its parameters as ONE flat fp32 buffer (layout documented in the header).
`FlatParams` re-homes a module's nn.Parameters as views into such a buffer (state_dict names and shapes are
untouched); backward writes one flat gradient buffer and hands autograd per-parameter views of it, so
`p.grad` ends up as views of a single allocation that the fused optimizer and the gradient all-reduce can use.
"""
from __future__ import annotations

import ctypes
import functools
import math
import os
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import SeqCfg, check, lib

_LAUNCHES = 0          # kernels enqueued by this process through the library (bench.py reads it)


def launches() -> int:
    return _LAUNCHES


def _count(n: int):
    global _LAUNCHES
    _LAUNCHES += n


# NVTX ranges: MVN_NVTX=1 brackets every library call group (encoder / ConvMixer / loss forward and backward, optimizer,
# graph replay) with a named range so that a timeline (nsys, ncu --nvtx) reads in the reference's terms.  Off by default: no cost.
_NVTX = os.environ.get("MVN_NVTX", "0") == "1"


def nvtx_range(name: str):
    """Decorator: run the function inside the NVTX range `maven/<name>` when MVN_NVTX=1."""
    def deco(fn):
        if not _NVTX:
            return fn

        @functools.wraps(fn)
        def wrapped(*a, **k):
            torch.cuda.nvtx.range_push("maven/" + name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"maven_b200: `{name}` is on {t.device}; this path has CUDA kernels only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"maven_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _mask_u8(mask: Optional[torch.Tensor], name="mask") -> Optional[torch.Tensor]:
    if mask is None:
        return None
    if not mask.is_cuda:
        raise RuntimeError(f"maven_b200: `{name}` is on {mask.device}; CUDA only")
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    return (mask != 0).contiguous().view(torch.uint8)


# ======================================================================================================
# flat parameter groups
# ======================================================================================================
class FlatParams:
    """Keeps an ordered list of nn.Parameters as contiguous views of one fp32 CUDA buffer."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.total = self.offsets[-1]
        self.buffer: Optional[torch.Tensor] = None

    def _is_bound(self) -> bool:
        b = self.buffer
        if b is None or not self.params:
            return False
        base = b.data_ptr()
        first, last = self.params[0], self.params[-1]
        return (first.data_ptr() == base and last.data_ptr() == base + 4 * self.offsets[-2]
                and first.device == b.device)

    def ensure(self) -> torch.Tensor:
        """Returns the flat buffer, (re)building it if the parameters were moved or replaced."""
        if self._is_bound():
            base = self.buffer.data_ptr()
            ok = all(p.data_ptr() == base + 4 * o for p, o in zip(self.params, self.offsets))
            if ok:
                return self.buffer
        dev = self.params[0].device          # any device: this is storage plumbing; the kernels reject CPU tensors
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("maven_b200: parameters must be float32 (bf16/tf32 arithmetic is selected per op, storage stays fp32)")
        flat = torch.empty(self.total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o, n in zip(self.params, self.offsets, self.sizes):
                v = flat[o:o + n].view(p.shape)
                v.copy_(p.data)
                p.data = v
        self.buffer = flat
        return flat

    def grad_views(self, gflat: torch.Tensor, needs: Sequence[bool]) -> List[Optional[torch.Tensor]]:
        return [gflat[o:o + n].view(p.shape) if need else None
                for p, o, n, need in zip(self.params, self.offsets, self.sizes, needs)]


# ======================================================================================================
# whole sequence encoder  (A1-A7)
# ======================================================================================================
class SeqCall:
    """Bookkeeping for one SeqEncoderFn call (kept out of the autograd argument list)."""
    __slots__ = ("cfg", "flat", "off", "count", "group", "pidx", "div_term", "gbuf", "goff")

    def __init__(self, cfg, flat, off, count, group, pidx, div_term, gbuf=None, goff=0):
        self.cfg, self.flat, self.off, self.count, self.group, self.pidx = cfg, flat, off, count, group, pidx
        self.div_term, self.gbuf, self.goff = div_term, gbuf, goff


class SeqEncoderFn(torch.autograd.Function):
    """mvn_seq_encoder_fwd / _bwd.  Inputs: x (B,T) fp32, t (B,T) fp32, mask (B,T) bool, a SeqCall, then the
    parameters (only so that autograd routes their gradients; the kernels read the flat buffer)."""

    @staticmethod
    @nvtx_range("seq_encoder.forward")
    def forward(ctx, x, t, mask, call: SeqCall, *params):
        L = lib()
        x = _req(x, "x"); t = _req(t, "t")
        if x.dim() != 2 or x.shape != t.shape:
            raise ValueError(f"seq encoder: x and t must both be (B, T), got {tuple(x.shape)} and {tuple(t.shape)}")
        m8 = _mask_u8(mask)
        B, T = x.shape
        cfg = SeqCfg.from_buffer_copy(call.cfg)
        cfg.B, cfg.T = B, T
        need = L.mvn_seq_param_count(ctypes.byref(cfg))
        if need != call.count:
            raise RuntimeError(f"maven_b200: parameter layout mismatch (library expects {need} floats, module packed {call.count}): "
                               + lib().mvn_last_error().decode())
        ws_bytes = L.mvn_seq_workspace_bytes(ctypes.byref(cfg))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        if cfg.agg == _lib.MVN_AGG_NONE:
            out = torch.empty(B, T, cfg.E, dtype=torch.float32, device=x.device)
        else:
            out = torch.empty(B, cfg.enc_dim if cfg.enc_dim > 0 else cfg.n_out, dtype=torch.float32, device=x.device)
        pview = call.flat[call.off:call.off + call.count]
        check(L.mvn_seq_encoder_fwd(ctypes.byref(cfg), _p(pview), _p(call.div_term), _p(x), _p(t), _p(m8), _p(out), _p(ws), ws_bytes, _stream()),
              "seq_encoder_fwd")
        _count(4 + 1 + 5 * cfg.depth + 5)
        ctx.cfg, ctx.ws, ctx.x, ctx.pview, ctx.call = cfg, ws, x, pview, call
        return out

    @staticmethod
    @nvtx_range("seq_encoder.backward")
    def backward(ctx, dout):
        L = lib()
        dout = _req(dout, "grad_output")
        call: SeqCall = ctx.call
        if call.gbuf is not None:
            g = call.gbuf[call.goff:call.goff + call.count]
        else:
            g = torch.empty(call.count, dtype=torch.float32, device=dout.device)
        cfg = ctx.cfg
        check(L.mvn_seq_encoder_bwd(ctypes.byref(cfg), _p(ctx.pview), _p(ctx.x), _p(dout), _p(g), _p(ctx.ws), ctx.ws.numel(), _stream()),
              "seq_encoder_bwd")
        _count(6 + 12 * cfg.depth + 2)
        _segment_done(call.gbuf, call.goff, call.count)        # data parallel: this encoder's gradients start their all-reduce now
        needs = ctx.needs_input_grad[4:]
        grp: FlatParams = call.group
        i0, i1 = call.pidx
        base = grp.offsets[i0]
        grads = []
        for k, need in zip(range(i0, i1), needs):
            if need:
                o = grp.offsets[k] - base
                grads.append(g[o:o + grp.sizes[k]].view(grp.params[k].shape))
            else:
                grads.append(None)
        ctx.ws = None
        return (None, None, None, None) + tuple(grads)


# ======================================================================================================
# per-op functions (standalone SelfAttention / TransformerBlock / Transformer modules, heads)
# ======================================================================================================
class LinearFn(torch.autograd.Function):
    """Y = X W^T + b over the last dimension (nn.Linear)."""

    @staticmethod
    @nvtx_range("linear.forward")
    def forward(ctx, x, w, b, prec: int):
        L = lib()
        shp = x.shape
        x2 = _req(x, "x").reshape(-1, shp[-1])
        w = _req(w, "weight")
        M, K = x2.shape
        N = w.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        check(L.mvn_linear_fwd(_p(x2), _p(w), _p(b), _p(y), None, M, N, K, _lib.MVN_ACT_NONE, prec, _stream()), "linear_fwd")
        _count(1)
        ctx.save_for_backward(x2, w)
        ctx.prec, ctx.has_b, ctx.shp = prec, b is not None, shp
        return y.view(*shp[:-1], N)

    @staticmethod
    @nvtx_range("linear.backward")
    def backward(ctx, dy):
        L = lib()
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dy = _req(dy, "grad_output").reshape(M, N)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
            check(L.mvn_linear_bwd_input(_p(dy), _p(w), _p(dx), None, None, 0, None, M, N, K, ctx.prec, _stream()), "linear_bwd_input")
            _count(1)
            dx = dx.view(ctx.shp)
        if ctx.needs_input_grad[1] or (ctx.has_b and ctx.needs_input_grad[2]):
            wsb = L.mvn_linear_bwd_weight_workspace_bytes(M, N, K)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dy.device)
            dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
            db = torch.empty(N, dtype=torch.float32, device=dy.device) if ctx.has_b else None
            check(L.mvn_linear_bwd_weight(_p(dy), _p(x2), _p(dw), _p(db), None, M, N, K, 0, _p(ws), wsb, ctx.prec, _stream()), "linear_bwd_weight")
            _count(3)
        return dx, dw, db, None


def linear(x, w, b=None, prec: int = 0):
    return LinearFn.apply(x, w, b, prec)


class LinearReluFn(torch.autograd.Function):
    """relu(X W^T + b) (MLP heads): fused ReLU epilogue forward, mask + the two GEMMs backward."""

    @staticmethod
    @nvtx_range("linear_relu.forward")
    def forward(ctx, x, w, b, prec: int):
        L = lib()
        shp = x.shape
        x2 = _req(x, "x").reshape(-1, shp[-1])
        M, K = x2.shape
        N = w.shape[0]
        h = torch.empty(M, N, dtype=torch.float32, device=x.device)
        check(L.mvn_linear_fwd(_p(x2), _p(_req(w, "weight")), _p(b), _p(h), None, M, N, K, _lib.MVN_ACT_RELU, prec, _stream()), "linear_fwd")
        _count(1)
        ctx.save_for_backward(x2, w, h)
        ctx.prec, ctx.has_b, ctx.shp = prec, b is not None, shp
        return h.view(*shp[:-1], N)

    @staticmethod
    @nvtx_range("linear_relu.backward")
    def backward(ctx, dh):
        L = lib()
        x2, w, h = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dh = _req(dh, "grad_output").reshape(M, N)
        dpre = torch.empty_like(dh)
        check(L.mvn_relu_bwd(_p(dh), _p(h), M * N, _p(dpre), _stream()), "relu_bwd")
        dx = torch.empty(M, K, dtype=torch.float32, device=dh.device)
        check(L.mvn_linear_bwd_input(_p(dpre), _p(w), _p(dx), None, None, 0, None, M, N, K, ctx.prec, _stream()), "linear_bwd_input")
        wsb = L.mvn_linear_bwd_weight_workspace_bytes(M, N, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dh.device)
        dw = torch.empty(N, K, dtype=torch.float32, device=dh.device)
        db = torch.empty(N, dtype=torch.float32, device=dh.device) if ctx.has_b else None
        check(L.mvn_linear_bwd_weight(_p(dpre), _p(x2), _p(dw), _p(db), None, M, N, K, 0, _p(ws), wsb, ctx.prec, _stream()), "linear_bwd_weight")
        _count(5)
        return dx.view(ctx.shp), dw, db, None


def next_dropout_seed(module) -> int:
    """A fresh 64-bit dropout seed per training-mode call (the kernels regenerate their counter-based masks from it in
    the backward): torch's seed, a per-module call counter and the data-parallel rank (shards draw different masks), so
    runs are reproducible under torch.manual_seed; `module.dropout_seed` pins it."""
    fixed = getattr(module, "dropout_seed", None)
    if fixed is not None:
        return int(fixed) & 0xFFFFFFFFFFFFFFFF
    module._drop_calls = getattr(module, "_drop_calls", 0) + 1
    rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank()
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + module._drop_calls * 0xD1B54A32D192ED03 + (id(module) % 65521) * 0x2545F491
            + rank * 0x9FB21C651E98DF25) & 0xFFFFFFFFFFFFFFFF


_STEP_COUNTER = None       # keeps the registered device counter alive for as long as the library points at it


def set_step_counter(counter: Optional[torch.Tensor]):
    """Registers a 1-element int32 CUDA tensor (or None) as the library's device step counter (mvn_set_step_counter):
    every dropout site mixes its value into the mask hash, which is what varies the masks across CUDA-graph replays."""
    global _STEP_COUNTER
    if counter is not None and not (counter.is_cuda and counter.dtype == torch.int32 and counter.numel() == 1):
        raise TypeError("maven_b200: the step counter must be a 1-element int32 CUDA tensor")
    lib().mvn_set_step_counter(None if counter is None else ctypes.c_void_p(counter.data_ptr()))
    _STEP_COUNTER = counter


class DropoutFn(torch.autograd.Function):
    """nn.Dropout over the last dimension's rows with the library's counter-based mask (mvn_dropout_apply); the backward
    regenerates the mask from (seed, site)."""

    @staticmethod
    def forward(ctx, x, p: float, seed: int, site: int):
        L = lib()
        x2 = _req(x, "x").reshape(-1, x.shape[-1])
        y = torch.empty_like(x2)
        check(L.mvn_dropout_apply(_p(x2), _p(y), x2.shape[0], x2.shape[1], seed, site, p, _stream()), "dropout_apply")
        _count(1)
        ctx.meta = (p, seed, site)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        L = lib()
        p, seed, site = ctx.meta
        d2 = _req(dy, "grad_output").reshape(-1, dy.shape[-1])
        dx = torch.empty_like(d2)
        check(L.mvn_dropout_apply(_p(d2), _p(dx), d2.shape[0], d2.shape[1], seed, site, p, _stream()), "dropout_apply")
        _count(1)
        return dx.view(dy.shape), None, None, None


def dropout(x, p: float, seed: int, site: int = 0):
    return x if not (p > 0.0) else DropoutFn.apply(x, float(p), int(seed), int(site))


class L2NormFn(torch.autograd.Function):
    """x / ||x||_2 over the last dimension, no epsilon (src/models_multimodal.py:279,286,293)."""

    @staticmethod
    def forward(ctx, x):
        L = lib()
        x2 = _req(x, "x").reshape(-1, x.shape[-1])
        B, D = x2.shape
        y = torch.empty_like(x2)
        nrm = torch.empty(B, dtype=torch.float32, device=x.device)
        check(L.mvn_l2norm_fwd(_p(x2), _p(y), _p(nrm), B, D, _stream()), "l2norm_fwd")
        _count(1)
        ctx.save_for_backward(y, nrm)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        L = lib()
        y, nrm = ctx.saved_tensors
        B, D = y.shape
        dy2 = _req(dy, "grad_output").reshape(B, D)
        dx = torch.empty_like(y)
        check(L.mvn_l2norm_bwd(_p(dy2), _p(y), _p(nrm), _p(dx), B, D, _stream()), "l2norm_bwd")
        _count(1)
        return dx.view(dy.shape)


class LinearResLNFn(torch.autograd.Function):
    """Y = LayerNorm(X W^T + b + R) gamma + beta (one kernel); backward = LN-bwd, dX GEMM, dW GEMM."""

    @staticmethod
    @nvtx_range("linear_res_ln.forward")
    def forward(ctx, x, w, b, r, gamma, beta, eps: float, prec: int):
        L = lib()
        shp = r.shape
        x2 = _req(x, "x").reshape(-1, x.shape[-1])
        r2 = _req(r, "residual").reshape(-1, shp[-1])
        M, K = x2.shape
        N = w.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        xhat = torch.empty_like(y)
        rstd = torch.empty(M, dtype=torch.float32, device=x.device)
        check(L.mvn_linear_res_ln_fwd(_p(x2), _p(_req(w, "w")), _p(b), _p(r2), _p(gamma), _p(beta), _p(y), _p(xhat), _p(rstd), None, M, N, K,
                                      eps, prec, _stream()), "linear_res_ln_fwd")
        _count(1)
        ctx.save_for_backward(x2, w, xhat, rstd, gamma)
        ctx.prec, ctx.shp, ctx.xshape = prec, shp, x.shape
        return y.view(shp)

    @staticmethod
    @nvtx_range("linear_res_ln.backward")
    def backward(ctx, dy):
        L = lib()
        x2, w, xhat, rstd, gamma = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dy = _req(dy, "grad_output").reshape(M, N)
        dz = torch.empty_like(dy)
        dg = torch.empty(N, dtype=torch.float32, device=dy.device)
        dbt = torch.empty_like(dg)
        wsb = max(L.mvn_linear_bwd_weight_workspace_bytes(M, N, K), L.mvn_num_slabs() * 2 * N * 4)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dy.device)
        check(L.mvn_layernorm_bwd(_p(dy), _p(xhat), _p(rstd), _p(gamma), _p(dz), _p(dg), _p(dbt), None, M, N, _p(ws), wsb, _stream()), "layernorm_bwd")
        dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
        check(L.mvn_linear_bwd_input(_p(dz), _p(w), _p(dx), None, None, 0, None, M, N, K, ctx.prec, _stream()), "linear_bwd_input")
        dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
        db = torch.empty(N, dtype=torch.float32, device=dy.device)
        check(L.mvn_linear_bwd_weight(_p(dz), _p(x2), _p(dw), _p(db), None, M, N, K, 0, _p(ws), wsb, ctx.prec, _stream()), "linear_bwd_weight")
        _count(3 + 1 + 3)
        return dx.view(ctx.xshape), dw, db, dz.view(ctx.shp), dg, dbt, None, None


class FFNResLNFn(torch.autograd.Function):
    """Y = LayerNorm(relu(X W1^T + b1) W2^T + b2 + X) gamma + beta  -- the feed-forward half of a TransformerBlock."""

    @staticmethod
    @nvtx_range("ffn_res_ln.forward")
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, eps: float, prec: int):
        L = lib()
        shp = x.shape
        x2 = _req(x, "x").reshape(-1, shp[-1])
        M, E = x2.shape
        F = w1.shape[0]
        dev = x.device
        h = torch.empty(M, F, dtype=torch.float32, device=dev)
        check(L.mvn_linear_fwd(_p(x2), _p(_req(w1, "w1")), _p(b1), _p(h), None, M, F, E, _lib.MVN_ACT_RELU, prec, _stream()), "linear_fwd")
        y = torch.empty(M, E, dtype=torch.float32, device=dev)
        xhat = torch.empty_like(y)
        rstd = torch.empty(M, dtype=torch.float32, device=dev)
        check(L.mvn_linear_res_ln_fwd(_p(h), _p(_req(w2, "w2")), _p(b2), _p(x2), _p(gamma), _p(beta), _p(y), _p(xhat), _p(rstd), None, M, E, F,
                                      eps, prec, _stream()), "linear_res_ln_fwd")
        _count(2)
        ctx.save_for_backward(x2, w1, w2, h, xhat, rstd, gamma)
        ctx.prec, ctx.shp = prec, shp
        return y.view(shp)

    @staticmethod
    @nvtx_range("ffn_res_ln.backward")
    def backward(ctx, dy):
        L = lib()
        x2, w1, w2, h, xhat, rstd, gamma = ctx.saved_tensors
        M, E = x2.shape
        F = w1.shape[0]
        dev = dy.device
        dy = _req(dy, "grad_output").reshape(M, E)
        f32 = dict(dtype=torch.float32, device=dev)
        wsb = max(L.mvn_linear_bwd_weight_workspace_bytes(M, F, E), L.mvn_num_slabs() * 2 * E * 4)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        dz = torch.empty(M, E, **f32); dg = torch.empty(E, **f32); dbt = torch.empty(E, **f32)
        check(L.mvn_layernorm_bwd(_p(dy), _p(xhat), _p(rstd), _p(gamma), _p(dz), _p(dg), _p(dbt), None, M, E, _p(ws), wsb, _stream()), "layernorm_bwd")
        dw2 = torch.empty(E, F, **f32); db2 = torch.empty(E, **f32)
        check(L.mvn_linear_bwd_weight(_p(dz), _p(h), _p(dw2), _p(db2), None, M, E, F, 0, _p(ws), wsb, ctx.prec, _stream()), "linear_bwd_weight")
        dh = torch.empty(M, F, **f32)
        check(L.mvn_linear_bwd_input(_p(dz), _p(w2), _p(dh), None, _p(h), 1, None, M, E, F, ctx.prec, _stream()), "linear_bwd_input")
        dw1 = torch.empty(F, E, **f32); db1 = torch.empty(F, **f32)
        check(L.mvn_linear_bwd_weight(_p(dh), _p(x2), _p(dw1), _p(db1), None, M, F, E, 0, _p(ws), wsb, ctx.prec, _stream()), "linear_bwd_weight")
        dx = torch.empty(M, E, **f32)
        check(L.mvn_linear_bwd_input(_p(dh), _p(w1), _p(dx), _p(dz), None, 0, None, M, F, E, ctx.prec, _stream()), "linear_bwd_input")
        _count(3 + 3 + 1 + 3 + 1)
        return dx.view(ctx.shp), dw1, db1, dw2, db2, dg, dbt, None, None


class AttentionFn(torch.autograd.Function):
    """Packed padding-masked attention core on qkv (M,3E) with dense plan (all B*T rows, keys masked)."""

    @staticmethod
    @nvtx_range("attention.forward")
    def forward(ctx, qkv, mask, B: int, T: int, E: int, H: int):
        L = lib()
        qkv2 = _req(qkv, "qkv").reshape(B * T, 3 * E)
        m8 = _mask_u8(mask)
        dev = qkv.device
        cu = torch.empty(B + 1, dtype=torch.int32, device=dev)
        tok = torch.empty(B * T, dtype=torch.int32, device=dev)
        kv = torch.empty(B * T, dtype=torch.uint8, device=dev)
        check(L.mvn_pack_plan(_p(m8), B, T, 0, _p(cu), _p(tok), _p(kv), _stream()), "pack_plan")
        out = torch.empty(B * T, E, dtype=torch.float32, device=dev)
        lse = torch.empty(B * T, H, dtype=torch.float32, device=dev)
        scale = 1.0 / math.sqrt(E)
        check(L.mvn_attention_fwd(_p(qkv2), _p(cu), _p(kv), _p(out), _p(lse), B, E, H, scale, 0, _stream()), "attention_fwd")
        _count(5)
        ctx.save_for_backward(qkv2, cu, kv, out, lse)
        ctx.dims = (B, T, E, H, scale)
        return out.view(B, T, E)

    @staticmethod
    @nvtx_range("attention.backward")
    def backward(ctx, dout):
        L = lib()
        qkv2, cu, kv, out, lse = ctx.saved_tensors
        B, T, E, H, scale = ctx.dims
        dout = _req(dout, "grad_output").reshape(B * T, E)
        dqkv = torch.empty_like(qkv2)
        check(L.mvn_attention_bwd(_p(qkv2), _p(cu), _p(kv), _p(out), _p(lse), _p(dout), _p(dqkv), B, E, H, scale, 0, _stream()), "attention_bwd")
        _count(1)
        return dqkv.view(B, T, 3 * E), None, None, None, None, None


class ConvCall:
    __slots__ = ("module", "flat", "off", "count", "group", "pidx", "enc_dim", "normalize", "prec", "gbuf", "goff", "dropout_p", "seed")

    def __init__(self, module, flat, off, count, group, pidx, enc_dim, normalize, prec, gbuf=None, goff=0, dropout_p=0.0, seed=0):
        self.module, self.flat, self.off, self.count, self.group, self.pidx = module, flat, off, count, group, pidx
        self.enc_dim, self.normalize, self.prec, self.gbuf, self.goff = enc_dim, normalize, prec, gbuf, goff
        self.dropout_p, self.seed = dropout_p, seed


class ConvMixerFn(torch.autograd.Function):
    """mvn_convmixer_fwd_stage / _bwd_stage, one stage per BatchNorm.  Under a data-parallel group the per-channel
    batch sums are all-reduced between stages (SyncBN semantics: statistics of the GLOBAL batch)."""

    @staticmethod
    @nvtx_range("convmixer.forward")
    def forward(ctx, x, call: ConvCall, *params):
        from ._lib import ConvCfg
        L = lib()
        x = _req(x, "x_img")
        if x.dim() != 4:
            raise ValueError(f"ConvMixer expects (B, C, H, W), got {tuple(x.shape)}")
        m = call.module
        B, C, H, W = x.shape
        if C != m.channels:
            raise ValueError(f"ConvMixer built for {m.channels} channels, got {C}")
        grp = _DP_GROUP if m.training else None
        world = 1
        if grp is not None:
            import torch.distributed as dist
            world = dist.get_world_size(grp)
        P = (H // m.patch_size) * (W // m.patch_size)
        bn0 = m.net[2]
        cfg = ConvCfg(B=B, C=C, H=H, W=W, dim=m.dim, depth=m.depth, kernel_size=m.kernel_size, patch_size=m.patch_size, n_out=m.n_out,
                      enc_dim=call.enc_dim, hidden=m.projection[2].out_features, normalize=1 if call.normalize else 0,
                      training=1 if m.training else 0, prec=call.prec, bn_eps=bn0.eps, bn_momentum=bn0.momentum if bn0.momentum is not None else 0.1,
                      global_count=B * P * world, dropout_p=call.dropout_p if m.training else 0.0, seed=call.seed)
        need = L.mvn_conv_param_count(ctypes.byref(cfg))
        if need != call.count:
            raise RuntimeError(f"maven_b200: ConvMixer parameter layout mismatch (library {need}, module {call.count}): "
                               + L.mvn_last_error().decode())
        nbn = 1 + 2 * m.depth
        wsb = L.mvn_conv_workspace_bytes(ctypes.byref(cfg))
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        stats = torch.zeros(nbn, 2 * m.dim, dtype=torch.float64, device=x.device)
        running = m.running_flat(x.device)
        D = call.enc_dim if call.enc_dim > 0 else m.n_out
        out = torch.empty(B, D, dtype=torch.float32, device=x.device)
        pview = call.flat[call.off:call.off + call.count]
        for s in range(nbn + 1):
            check(L.mvn_convmixer_fwd_stage(ctypes.byref(cfg), s, _p(pview), _p(x), _p(running), _p(stats), _p(out), _p(ws), wsb, _stream()),
                  "convmixer_fwd_stage")
            if world > 1 and s < nbn:
                import torch.distributed as dist
                dist.all_reduce(stats[s], group=grp)
        _count(5 + 6 * m.depth * 2 + 8)
        if m.training:
            torch._foreach_add_([bn.num_batches_tracked for bn in m.bn_layers()], 1)
        ctx.cfg, ctx.ws, ctx.stats, ctx.x, ctx.pview, ctx.call, ctx.world, ctx.grp = cfg, ws, stats, x, pview, call, world, grp
        return out

    @staticmethod
    @nvtx_range("convmixer.backward")
    def backward(ctx, dout):
        L = lib()
        dout = _req(dout, "grad_output")
        call: ConvCall = ctx.call
        cfg = ctx.cfg
        nbn = 1 + 2 * cfg.depth
        if call.gbuf is not None:
            g = call.gbuf[call.goff:call.goff + call.count]
        else:
            g = torch.empty(call.count, dtype=torch.float32, device=dout.device)
        sb = torch.zeros(nbn, 2 * cfg.dim, dtype=torch.float64, device=dout.device)
        for s in range(nbn, -1, -1):
            check(L.mvn_convmixer_bwd_stage(ctypes.byref(cfg), s, _p(ctx.pview), _p(ctx.x), _p(ctx.stats), _p(sb), _p(dout), _p(g), _p(ctx.ws),
                                            ctx.ws.numel(), _stream()), "convmixer_bwd_stage")
            if ctx.world > 1 and s > 0:
                import torch.distributed as dist
                dist.all_reduce(sb[s - 1], group=ctx.grp)
        _count(10 + 8 * cfg.depth * 2)
        _segment_done(call.gbuf, call.goff, call.count)
        needs = ctx.needs_input_grad[2:]
        grp: FlatParams = call.group
        i0, i1 = call.pidx
        base = grp.offsets[i0]
        grads = []
        for k, need in zip(range(i0, i1), needs):
            if need:
                o = grp.offsets[k] - base
                grads.append(g[o:o + grp.sizes[k]].view(grp.params[k].shape))
            else:
                grads.append(None)
        ctx.ws = None
        return (None, None) + tuple(grads)


class QueryPoolFn(torch.autograd.Function):
    """agg='attn' pooling core (mvn_query_pool_fwd/_bwd): q (1,E) projected query, kv (B,T,2E) projected keys|values."""

    @staticmethod
    def forward(ctx, q, kv, B: int, T: int, E: int, H: int):
        L = lib()
        q2 = _req(q, "query").reshape(E)
        kv2 = _req(kv, "kv").reshape(B * T, 2 * E)
        out = torch.empty(B, E, dtype=torch.float32, device=kv.device)
        probs = torch.empty(B, H, T, dtype=torch.float32, device=kv.device)
        check(L.mvn_query_pool_fwd(_p(q2), _p(kv2), B, T, E, H, _p(out), _p(probs), _stream()), "query_pool_fwd")
        _count(1)
        ctx.save_for_backward(q2, kv2, probs)
        ctx.dims = (B, T, E, H, q.shape, kv.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = lib()
        q2, kv2, probs = ctx.saved_tensors
        B, T, E, H, qshape, kvshape = ctx.dims
        dout = _req(dout, "grad_output")
        dkv = torch.empty_like(kv2)
        dq = torch.empty(E, dtype=torch.float32, device=dout.device)
        ws = torch.empty(B * E * 4, dtype=torch.uint8, device=dout.device)
        check(L.mvn_query_pool_bwd(_p(q2), _p(kv2), _p(probs), _p(dout), B, T, E, H, _p(dkv), _p(dq), _p(ws), ws.numel(), _stream()), "query_pool_bwd")
        _count(2)
        return dq.view(qshape), dkv.view(kvshape), None, None, None, None


class AttnPoolFn(torch.autograd.Function):
    """agg='attn' pooling in closed form (mvn_attn_pool_fwd/_bwd): tokens (B,T,E) zero on padding, bool mask (B,T), the learnable
    query and nn.MultiheadAttention's in_proj / out_proj parameters -> pooled (B,E).  One kernel forward; the k|v projection of the
    B*T tokens of the per-op path is never formed."""

    @staticmethod
    @nvtx_range("attn_pool.forward")
    def forward(ctx, tokens, mask, query, in_w, in_b, out_w, out_b, H: int):
        L = lib()
        tokens = _req(tokens, "tokens")
        B, T, E = tokens.shape
        m = _mask_u8(mask, "mask")
        ps = [_req(t, n) for t, n in ((query, "query"), (in_w, "in_proj_weight"), (in_b, "in_proj_bias"), (out_w, "out_proj.weight"), (out_b, "out_proj.bias"))]
        out = torch.empty(B, E, dtype=torch.float32, device=tokens.device)
        saved = torch.empty(L.mvn_attn_pool_saved_bytes(B, E, H) // 4, dtype=torch.float32, device=tokens.device)
        check(L.mvn_attn_pool_fwd(_p(tokens), _p(m), *[_p(t) for t in ps], B, T, E, H, _p(out), _p(saved), saved.numel() * 4, _stream()), "attn_pool_fwd")
        _count(1)
        ctx.save_for_backward(tokens, m, *ps, saved)
        ctx.dims = (B, T, E, H)
        return out

    @staticmethod
    @nvtx_range("attn_pool.backward")
    def backward(ctx, dout):
        L = lib()
        tokens, m, query, in_w, in_b, out_w, out_b, saved = ctx.saved_tensors
        B, T, E, H = ctx.dims
        dout = _req(dout, "grad_output")
        dx = torch.empty_like(tokens)
        g = [torch.empty_like(t) for t in (query, in_w, in_b, out_w, out_b)]
        wsb = L.mvn_attn_pool_bwd_workspace_bytes(B, E, H)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dout.device)
        check(L.mvn_attn_pool_bwd(_p(tokens), _p(m), _p(query), _p(in_w), _p(in_b), _p(out_w), _p(saved), _p(dout), B, T, E, H, _p(dx),
                                  *[_p(t) for t in g], _p(ws), wsb, _stream()), "attn_pool_bwd")
        _count(5 + H)
        return (dx, None, *g, None)


def attn_pool_supported(T: int, E: int, H: int) -> bool:
    return E <= 128 and 128 % E == 0 and H <= 8 and E % H == 0 and H * T * 4 <= 96 * 1024


# ------------------------------------------------------------------------------------------------------
# CLIP loss (A10/A11)
# ------------------------------------------------------------------------------------------------------
_DP_GROUP = None


def set_data_parallel_group(group):
    """Process group whose ranks shard the global batch (None = single process).  With a group set, clip_loss
    all-gathers the embeddings (global-batch negatives) and both LSE vectors; see DESIGN.md, multi-GPU."""
    global _DP_GROUP
    _DP_GROUP = group


def get_data_parallel_group():
    return _DP_GROUP


# ---- gradient all-reduce overlapped with the backward pass -----------------------------------------------------------
# The flat gradient buffer is laid out encoder by encoder.  An encoder's backward writes its whole segment in one library
# call, on the stream its forward ran on; with a data-parallel group set, that segment's all-reduce is launched right there
# (async: NCCL's stream waits for that encoder's kernels only), so it runs while the other modality's backward is still
# executing.  finish_grad_reduce() waits for the launched pieces and reduces what is left (heads, logit scale) in one call.
_GRAD_OVERLAP = {"gbuf": None, "pending": []}


def begin_grad_overlap(gbuf: Optional[torch.Tensor]):
    """Called by the model's forward for every training step: segments of `gbuf` written by fused backward calls from now on are
    all-reduced as soon as they are complete.  None switches the mechanism off."""
    _GRAD_OVERLAP["gbuf"] = gbuf if (_DP_GROUP is not None and gbuf is not None) else None
    _GRAD_OVERLAP["pending"] = []


def _segment_done(gbuf: Optional[torch.Tensor], off: int, count: int):
    if gbuf is None or _GRAD_OVERLAP["gbuf"] is not gbuf or _DP_GROUP is None or count == 0:
        return
    import torch.distributed as dist
    work = dist.all_reduce(gbuf[off:off + count], group=_DP_GROUP, async_op=True)
    _GRAD_OVERLAP["pending"].append((off, count, work))


def finish_grad_reduce(gbuf: torch.Tensor, group=None):
    """SUM all-reduce of the flat gradient buffer over the data-parallel group: waits for the segments launched during the
    backward and reduces the remaining ranges.  Equivalent to one dist.all_reduce(gbuf)."""
    import torch.distributed as dist
    group = group if group is not None else _DP_GROUP
    if group is None:
        return
    pend = _GRAD_OVERLAP["pending"] if _GRAD_OVERLAP["gbuf"] is gbuf else []
    done = sorted((o, c) for o, c, _ in pend)
    for _, _, w in pend:
        w.wait()
    pos, rest = 0, []
    for o, c in done:
        if o > pos:
            rest.append((pos, o))
        pos = max(pos, o + c)
    if pos < gbuf.numel():
        rest.append((pos, gbuf.numel()))
    for a, b in rest:
        dist.all_reduce(gbuf[a:b], group=group)
    _GRAD_OVERLAP["pending"] = []
    _GRAD_OVERLAP["gbuf"] = None


def _all_gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    out = torch.empty((ws * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def _clip_fwd_local(e1, e2, e1_all, e2_all, n, N, D, off, ls, lb, prec):
    """This rank's rows/columns of the streamed loss: (loss share [1], lse [2, n]) -- mvn_clip_loss_fwd."""
    L = lib()
    e1 = _req(e1, "embs1"); e2 = _req(e2, "embs2"); ls = _req(ls, "logit_scale"); lb = _req(lb, "logit_bias")
    wsb = L.mvn_clip_loss_workspace_bytes(n, N, D)
    ws = torch.empty(wsb, dtype=torch.uint8, device=e1.device)
    loss = torch.empty(1, dtype=torch.float32, device=e1.device)
    lse = torch.empty(2, n, dtype=torch.float32, device=e1.device)
    check(L.mvn_clip_loss_fwd(_p(e1), _p(e2), _p(e1_all), _p(e2_all), n, N, D, off, _p(ls), _p(lb), _p(loss), _p(lse[0]), _p(lse[1]),
                              _p(ws), wsb, prec, _stream()), "clip_loss_fwd")
    _count(4)
    return loss, lse


def _clip_bwd_local(e1, e2, e1_all, e2_all, n, N, D, off, ls, lb, lse_all, g, prec):
    """(d_e1_local, d_e2_local, d_logit_scale share [1]) from the all-gathered LSE vectors -- mvn_clip_loss_bwd."""
    L = lib()
    g = _req(g, "grad_output")
    wsb = L.mvn_clip_loss_workspace_bytes(n, N, D)
    ws = torch.empty(wsb, dtype=torch.uint8, device=e1.device)
    d1 = torch.empty_like(e1); d2 = torch.empty_like(e2)
    dls = torch.empty(1, dtype=torch.float32, device=e1.device)
    check(L.mvn_clip_loss_bwd(_p(e1), _p(e2), _p(e1_all), _p(e2_all), n, N, D, off, _p(ls), _p(lb), _p(lse_all[0]), _p(lse_all[1]), _p(g),
                              _p(d1), _p(d2), _p(dls), _p(ws), wsb, prec, _stream()), "clip_loss_bwd")
    _count(5)
    return d1, d2, dls


class ClipLossFn(torch.autograd.Function):
    """Symmetric InfoNCE over the GLOBAL batch.  Host logic = the exchange protocol of DESIGN.md section 5 (all-gather
    embeddings -> rank-local LSEs -> all-gather LSEs / all-reduce the loss share); the arithmetic is the two kernels
    above.  (tests/test_dp_gloo.py drives this protocol on two CPU processes with test-side torch stand-ins for the two kernels.)"""

    @staticmethod
    @nvtx_range("clip_loss.forward")
    def forward(ctx, e1, e2, logit_scale, logit_bias, prec: int):
        if e1.shape != e2.shape or e1.dim() != 2:
            raise ValueError(f"clip_loss: embeddings must both be (N, D), got {tuple(e1.shape)} and {tuple(e2.shape)}")
        e1 = e1.contiguous(); e2 = e2.contiguous()
        ls = logit_scale.reshape(1); lb = logit_bias.reshape(1)
        n, D = e1.shape
        grp = _DP_GROUP
        if grp is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(grp), dist.get_world_size(grp)
            e1_all, e2_all = _all_gather_rows(e1, grp), _all_gather_rows(e2, grp)
        else:
            rank, world, e1_all, e2_all = 0, 1, e1, e2
        N, off = n * world, rank * n
        loss, lse = _clip_fwd_local(e1, e2, e1_all, e2_all, n, N, D, off, ls, lb, prec)
        if grp is not None:
            import torch.distributed as dist
            lse_all = _all_gather_rows(lse.t().contiguous(), grp).t().contiguous()      # (2, N), rank-major = global row order
            dist.all_reduce(loss, group=grp)
        else:
            lse_all = lse
        ctx.save_for_backward(e1, e2, e1_all, e2_all, ls, lb, lse_all)
        ctx.meta = (n, N, D, off, prec, logit_scale.shape, logit_bias.shape)
        return loss.reshape(())

    @staticmethod
    @nvtx_range("clip_loss.backward")
    def backward(ctx, g):
        e1, e2, e1_all, e2_all, ls, lb, lse_all = ctx.saved_tensors
        n, N, D, off, prec, ls_shape, lb_shape = ctx.meta
        d1, d2, dls = _clip_bwd_local(e1, e2, e1_all, e2_all, n, N, D, off, ls, lb, lse_all, g.reshape(1).contiguous(), prec)
        # d loss / d logit_bias is identically zero under the two softmaxes; autograd still reports a zero tensor.
        # d_logit_scale is this rank's share: it is summed over ranks by the flat gradient all-reduce like every parameter grad.
        dlb = torch.zeros(lb_shape, dtype=torch.float32, device=e1.device)
        return d1, d2, dls.reshape(ls_shape), dlb, None


class ClipLossMultiFn(torch.autograd.Function):
    """clip_loss_multimodal (src/loss.py:41-65) for M modalities sharing one scalar logit scale / bias: the sum over the
    M(M-1)/2 pairs with ONE all-gather of all modalities' embeddings, ONE all-gather of all pairs' LSE vectors and ONE
    all-reduce of the loss (the per-pair form gathers every embedding once per pair it takes part in).  The arithmetic is
    the same two kernels per pair as ClipLossFn."""

    @staticmethod
    @nvtx_range("clip_loss_multimodal.forward")
    def forward(ctx, logit_scale, logit_bias, prec: int, *embs):
        M = len(embs)
        n, D = embs[0].shape
        for e in embs:
            if e.shape != (n, D):
                raise ValueError(f"clip_loss_multimodal: every modality must be ({n}, {D}), got {tuple(e.shape)}")
        ls = logit_scale.reshape(1); lb = logit_bias.reshape(1)
        e_loc = torch.stack(list(embs))                                               # (M, n, D); device / dtype are checked by the kernel calls
        grp = _DP_GROUP
        if grp is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(grp), dist.get_world_size(grp)
            e_all = _all_gather_rows(e_loc.unsqueeze(0), grp).permute(1, 0, 2, 3).reshape(M, world * n, D).contiguous()
        else:
            rank, world, e_all = 0, 1, e_loc
        N, off = n * world, rank * n
        pairs = [(i, j) for i in range(M - 1) for j in range(i + 1, M)]
        total = None
        lses = []
        for i, j in pairs:
            loss, lse = _clip_fwd_local(e_loc[i], e_loc[j], e_all[i], e_all[j], n, N, D, off, ls, lb, prec)
            total = loss if total is None else total + loss
            lses.append(lse)
        lse_loc = torch.stack(lses)                                                       # (P, 2, n)
        if grp is not None:
            import torch.distributed as dist
            lse_all = _all_gather_rows(lse_loc.unsqueeze(0), grp).permute(1, 2, 0, 3).reshape(len(pairs), 2, N).contiguous()
            dist.all_reduce(total, group=grp)
        else:
            lse_all = lse_loc
        ctx.save_for_backward(e_loc, e_all, ls, lb, lse_all)
        ctx.meta = (M, n, N, D, off, prec, pairs, logit_scale.shape, logit_bias.shape)
        return total.reshape(())

    @staticmethod
    @nvtx_range("clip_loss_multimodal.backward")
    def backward(ctx, g):
        e_loc, e_all, ls, lb, lse_all = ctx.saved_tensors
        M, n, N, D, off, prec, pairs, ls_shape, lb_shape = ctx.meta
        g1 = g.reshape(1).contiguous()
        de = [None] * M
        dls_tot = None
        for p, (i, j) in enumerate(pairs):
            d1, d2, dls = _clip_bwd_local(e_loc[i], e_loc[j], e_all[i], e_all[j], n, N, D, off, ls, lb, lse_all[p], g1, prec)
            de[i] = d1 if de[i] is None else de[i] + d1
            de[j] = d2 if de[j] is None else de[j] + d2
            dls_tot = dls if dls_tot is None else dls_tot + dls
        dlb = torch.zeros(lb_shape, dtype=lb.dtype, device=e_loc.device)              # identically zero under the two softmaxes
        return (dls_tot.reshape(ls_shape), dlb, None) + tuple(de)


class WeightedCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, class_w):
        L = lib()
        logits = _req(logits, "logits")
        if labels.dtype != torch.int64:
            labels = labels.long()
        labels = labels.contiguous()
        B, C = logits.shape
        buf = torch.empty(2 + 2 * B, dtype=torch.float32, device=logits.device)
        check(L.mvn_weighted_ce_fwd(_p(logits), _p(labels), _p(class_w), B, C, _p(buf), _stream()), "weighted_ce_fwd")
        _count(2)
        ctx.save_for_backward(logits, labels, class_w, buf)
        wsum = buf[1].clone().reshape(())                      # sum of the class weights of this batch (the loss's denominator)
        ctx.mark_non_differentiable(wsum)
        return buf[0].clone().reshape(()), wsum

    @staticmethod
    def backward(ctx, g, _gw=None):
        L = lib()
        logits, labels, class_w, buf = ctx.saved_tensors
        B, C = logits.shape
        d = torch.empty_like(logits)
        g = _req(g.reshape(1), "grad_output")
        check(L.mvn_weighted_ce_bwd(_p(logits), _p(labels), _p(class_w), B, C, _p(buf), _p(g), _p(d), _stream()), "weighted_ce_bwd")
        _count(1)
        return d, None, None


class MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        L = lib()
        pred = _req(pred, "pred").reshape(-1); target = _req(target, "target").reshape(-1)
        if pred.numel() != target.numel():
            raise ValueError("mse: size mismatch")
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        check(L.mvn_mse_fwd(_p(pred), _p(target), pred.numel(), _p(loss), _stream()), "mse_fwd")
        _count(1)
        ctx.save_for_backward(pred, target)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        L = lib()
        pred, target = ctx.saved_tensors
        d = torch.empty_like(pred)
        g = _req(g.reshape(1), "grad_output")
        check(L.mvn_mse_bwd(_p(pred), _p(target), pred.numel(), _p(g), _p(d), _stream()), "mse_bwd")
        _count(1)
        return d, None


class MaskedMSEFn(torch.autograd.Function):
    """nn.MSELoss()(target[mask], pred[mask]) without the gather (src/models_pretraining.py:204-205, 221-222): mean of the squared
    differences over the positions where `mask` is set.  Returns (loss, number of selected positions)."""

    @staticmethod
    def forward(ctx, pred, target, mask):
        L = lib()
        ctx.pred_shape = pred.shape
        pred = _req(pred, "pred").reshape(-1); target = _req(target, "target").reshape(-1)
        m = _mask_u8(mask.reshape(-1), "mask_pred")
        if not (pred.numel() == target.numel() == m.numel()):
            raise ValueError("masked_mse: size mismatch")
        buf = torch.empty(2, dtype=torch.float32, device=pred.device)
        check(L.mvn_masked_mse_fwd(_p(pred), _p(target), _p(m), pred.numel(), _p(buf), _stream()), "masked_mse_fwd")
        _count(1)
        ctx.save_for_backward(pred, target, m, buf)
        loss, cnt = buf[0], buf[1]
        ctx.mark_non_differentiable(cnt)
        return loss, cnt

    @staticmethod
    def backward(ctx, g, _gcount):
        L = lib()
        pred, target, m, buf = ctx.saved_tensors
        d = torch.empty_like(pred)
        g = _req(g.reshape(1), "grad_output")
        check(L.mvn_masked_mse_bwd(_p(pred), _p(target), _p(m), pred.numel(), _p(buf), _p(g), _p(d), _stream()), "masked_mse_bwd")
        _count(1)
        return d.view(ctx.pred_shape), None, None


class _DPWeightedMeanFn(torch.autograd.Function):
    """Global-batch value of a per-rank weighted mean: sum_r(loss_r * w_r) / sum_r(w_r).  Backward scales the local
    gradient by w_local / w_global, so that the SUM all-reduce of the flat gradient buffer yields the gradient of the
    global-batch loss (a plain local mean would come out world-size times too large, and for the weighted cross-entropy
    normalised by the wrong denominator)."""

    @staticmethod
    def forward(ctx, loss_local, w_local, group):
        import torch.distributed as dist
        pair = torch.stack([loss_local.reshape(()) * w_local.reshape(()), w_local.reshape(())])
        dist.all_reduce(pair, group=group)
        ctx.save_for_backward(w_local.reshape(()) / pair[1])
        return pair[0] / pair[1]

    @staticmethod
    def backward(ctx, g):
        (ratio,) = ctx.saved_tensors
        return g * ratio, None, None


def dp_weighted_mean(loss_local: torch.Tensor, weight_local) -> torch.Tensor:
    """MSE / weighted-CE heads under data parallelism (src/models_multimodal.py:326,335-349 define them on the whole batch):
    identity on one process; with a data-parallel group the global-batch loss with the matching gradient scale."""
    grp = _DP_GROUP
    if grp is None:
        return loss_local
    if not torch.is_tensor(weight_local):
        weight_local = torch.full((), float(weight_local), dtype=loss_local.dtype, device=loss_local.device)   # a fill kernel: graph-capturable
    return _DPWeightedMeanFn.apply(loss_local, weight_local.to(loss_local.dtype), grp)


class MetaInputFn(torch.autograd.Function):
    """[class_emb[cls] | redshift repeated] (src/models_multimodal.py:295-304) as one gather kernel; backward = the embedding
    table's gradient (deterministic per-class sums)."""

    @staticmethod
    def forward(ctx, class_emb, cls, redshift):
        L = lib()
        w = _req(class_emb, "class_emb.weight")
        cls = cls.long().contiguous()
        red = _req(redshift.float(), "redshift")
        B, (n_classes, half) = cls.numel(), w.shape
        out = torch.empty(B, 2 * half, dtype=torch.float32, device=w.device)
        check(L.mvn_meta_input_fwd(_p(w), _p(cls), _p(red), B, half, n_classes, _p(out), _stream()), "meta_input_fwd")
        _count(1)
        ctx.save_for_backward(cls)
        ctx.dims = (B, half, n_classes)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = lib()
        (cls,) = ctx.saved_tensors
        B, half, n_classes = ctx.dims
        dout = _req(dout, "grad_output")
        dw = torch.empty(n_classes, half, dtype=torch.float32, device=dout.device)
        check(L.mvn_meta_input_bwd(_p(dout), _p(cls), B, half, n_classes, _p(dw), _stream()), "meta_input_bwd")
        _count(1)
        return dw, None, None


def retrieval_ranks(e1: torch.Tensor, e2: torch.Tensor) -> torch.Tensor:
    L = lib()
    e1 = _req(e1, "embs1"); e2 = _req(e2, "embs2")
    N, D = e1.shape
    r = torch.empty(N, dtype=torch.int32, device=e1.device)
    check(L.mvn_retrieval_ranks(_p(e1), _p(e2), N, D, _p(r), _stream()), "retrieval_ranks")
    _count(1)
    return r


def retrieval_curve(ranks: torch.Tensor, k_thr: torch.Tensor) -> torch.Tensor:
    """counts[t] = #{j : ranks[j] < k_thr[t]} (mvn_retrieval_curve); both int32 CUDA tensors."""
    L = lib()
    ranks = _req(ranks, "ranks", torch.int32); k_thr = _req(k_thr, "k_thr", torch.int32)
    counts = torch.empty(k_thr.numel(), dtype=torch.int32, device=ranks.device)
    check(L.mvn_retrieval_curve(_p(ranks), ranks.numel(), _p(k_thr), k_thr.numel(), _p(counts), _stream()), "retrieval_curve")
    _count(1)
    return counts
