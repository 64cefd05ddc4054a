"""Fused RAdam (A13): torch.optim.RAdam semantics (coupled L2 weight decay, variance rectification) as ONE kernel over
the model's flat parameter / gradient / moment buffers instead of ~130 per-tensor updates.
reference: src/models_multimodal.py:306-310 (torch.optim.RAdam(self.parameters(), lr, **optimizer_kwargs))."""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import torch

from . import ops
from ._lib import check, lib


class FusedRAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, model=None):
        if model is None or not hasattr(model, "flat_group"):
            raise ValueError("FusedRAdam needs model=<module with flat_group()> (the maven_b200 LightCurveImageCLIP)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedRAdam supports a single parameter group")
        self.model = model
        self._m: Optional[torch.Tensor] = None
        self._v: Optional[torch.Tensor] = None
        self._steps: Optional[list] = None          # per-parameter step counts (torch skips params whose grad is None)
        self._gstage: Optional[torch.Tensor] = None

    def _ensure_state(self, g: ops.FlatParams, flat: torch.Tensor):
        if self._m is None or self._m.numel() != g.total or self._m.device != flat.device:
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            self._steps = [0] * len(g.params)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = lib()
        grp = self.param_groups[0]
        g: ops.FlatParams = self.model.flat_group()
        flat = g.ensure()
        self._ensure_state(g, flat)
        gbuf = self.model.gather_grads()
        in_group = {id(p) for p in grp["params"]}
        # contiguous runs of parameters that (a) have a gradient, (b) share the same step count
        runs = []
        for k, p in enumerate(g.params):
            if p.grad is None or id(p) not in in_group:
                continue
            o, n = g.offsets[k], g.sizes[k]
            self._steps[k] += 1
            st = self._steps[k]
            if runs and runs[-1][1] == o and runs[-1][2] == st:
                runs[-1][1] = o + n
            else:
                runs.append([o, o + n, st])
        lr, (b1, b2), eps, wd = grp["lr"], grp["betas"], grp["eps"], grp["weight_decay"]
        stream = ops._stream()
        for o0, o1, st in runs:
            bc1 = 1.0 - b1 ** st
            bc2 = 1.0 - b2 ** st
            rho_inf = 2.0 / (1.0 - b2) - 1.0
            rho_t = rho_inf - 2.0 * st * (b2 ** st) / bc2
            rect = -1.0
            if rho_t > 5.0:
                rect = math.sqrt((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t))
            n = o1 - o0
            check(L.mvn_radam_step(ctypes.c_void_p(flat.data_ptr() + 4 * o0), ctypes.c_void_p(gbuf.data_ptr() + 4 * o0),
                                   ctypes.c_void_p(self._m.data_ptr() + 4 * o0), ctypes.c_void_p(self._v.data_ptr() + 4 * o0), n,
                                   lr, b1, b2, eps, wd, bc1, math.sqrt(bc2), rect, stream), "radam_step")
            ops._count(1)
        return loss
