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

    # ---- CUDA-graph support (maven_b200.graph.GraphedTrainStep) ------------------------------------------------
    def snapshot(self):
        return (None if self._m is None else self._m.clone(), None if self._v is None else self._v.clone(),
                None if self._steps is None else list(self._steps))

    def restore(self, snap):
        m, v, steps = snap
        if m is None:
            self._m = self._v = self._steps = None
        else:
            self._m.copy_(m); self._v.copy_(v); self._steps = list(steps)

    def enable_device_step(self):
        """From now on step() keeps the step count on the device and computes the step-dependent scalars there
        (mvn_radam_step_dev), so that it can be captured in a CUDA graph; the same counter re-seeds the dropout masks."""
        g: ops.FlatParams = self.model.flat_group()
        flat = g.ensure()
        self._ensure_state(g, flat)
        start = max(self._steps) if self._steps else 0
        if any(s_ not in (0, start) for s_ in self._steps):       # 0: parameters that never receive a gradient (e.g. logit_scale of a classifier)
            raise RuntimeError("FusedRAdam: device-side stepping needs every stepped parameter at the same step count")
        self._step_dev = torch.full((1,), start, dtype=torch.int32, device=flat.device)
        self._scal_dev = torch.zeros(4, dtype=torch.float32, device=flat.device)
        ops.set_step_counter(self._step_dev)

    def disable_device_step(self):
        """Back to host-side stepping (eager mode); the host mirror of the step count carries on from the device's."""
        if getattr(self, "_step_dev", None) is not None:
            n = int(self._step_dev.item())
            stepped = getattr(self, "_stepped", None)
            self._steps = [n if (stepped is None or stepped[k]) else s for k, s in enumerate(self._steps)]
            if ops._STEP_COUNTER is self._step_dev:
                ops.set_step_counter(None)
            self._step_dev = self._scal_dev = None

    def note_graph_replay(self):
        if self._steps is not None:
            stepped = getattr(self, "_stepped", None)
            self._steps = [s + 1 if (stepped is None or stepped[k]) else s for k, s in enumerate(self._steps)]

    # ---- checkpoint format: torch.optim.RAdam's ------------------------------------------------------------------
    # The moments live in two flat buffers, not in Optimizer.state; state_dict() / load_state_dict() translate to and from
    # torch.optim.RAdam's per-parameter {'step', 'exp_avg', 'exp_avg_sq'} entries, so Lightning checkpoints written through
    # configure_optimizers() resume exactly and a reference RAdam state loads into this optimizer (and vice versa).
    def state_dict(self):
        g: ops.FlatParams = self.model.flat_group()
        self.state.clear()
        if self._m is not None:
            if getattr(self, "_step_dev", None) is not None:          # device-side stepping: the device counter is the truth
                n = int(self._step_dev.item())
                stepped = getattr(self, "_stepped", None)
                steps = [n if (stepped is None or stepped[k]) else s_ for k, s_ in enumerate(self._steps)]
            else:
                steps = self._steps
            for k, p in enumerate(g.params):
                if steps[k] == 0:
                    continue                                          # torch creates state lazily, at a parameter's first step
                o, n_ = g.offsets[k], g.sizes[k]
                self.state[p] = {"step": torch.tensor(float(steps[k]), dtype=torch.float32),
                                 "exp_avg": self._m[o:o + n_].view(p.shape).clone(),
                                 "exp_avg_sq": self._v[o:o + n_].view(p.shape).clone()}
        try:
            return super().state_dict()
        finally:
            self.state.clear()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)                           # validates the groups, maps ids -> parameters, casts to device
        g: ops.FlatParams = self.model.flat_group()
        flat = g.ensure()
        self._m = self._v = None
        self._ensure_state(g, flat)
        index = {id(p): k for k, p in enumerate(g.params)}
        for p, st in list(self.state.items()):
            k = index.get(id(p))
            if k is None:
                continue
            o, n_ = g.offsets[k], g.sizes[k]
            self._m[o:o + n_].copy_(st["exp_avg"].reshape(-1))
            self._v[o:o + n_].copy_(st["exp_avg_sq"].reshape(-1))
            self._steps[k] = int(round(float(st["step"])))
        self.state.clear()
        if getattr(self, "_step_dev", None) is not None:
            self._step_dev.fill_(max(self._steps) if self._steps else 0)

    def _ensure_state(self, g: ops.FlatParams, flat: torch.Tensor):
        if self._m is None or self._m.numel() != g.total or self._m.device != flat.device:
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            self._steps = [0] * len(g.params)

    @torch.no_grad()
    @ops.nvtx_range("radam.step")
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
        if getattr(self, "_step_dev", None) is not None:
            if len({r[2] for r in runs}) > 1:
                raise RuntimeError("FusedRAdam: device-side stepping needs every stepped parameter at the same step count")
            self._stepped = [False] * len(self._steps)
            for k, p in enumerate(g.params):
                if p.grad is not None and id(p) in in_group:
                    self._stepped[k] = True
                    self._steps[k] -= 1              # the device counter is the truth; note_graph_replay() advances the mirror
            for i, (o0, o1, _) in enumerate(runs):   # parameters without a gradient are skipped, like torch.optim.RAdam does
                check(L.mvn_radam_step_dev(ctypes.c_void_p(flat.data_ptr() + 4 * o0), ctypes.c_void_p(gbuf.data_ptr() + 4 * o0),
                                           ctypes.c_void_p(self._m.data_ptr() + 4 * o0), ctypes.c_void_p(self._v.data_ptr() + 4 * o0), o1 - o0,
                                           lr, b1, b2, eps, wd, ctypes.c_void_p(self._step_dev.data_ptr()) if i == 0 else None,
                                           ctypes.c_void_p(self._scal_dev.data_ptr()), stream), "radam_step_dev")
                ops._count(2 if i == 0 else 1)
            if not torch.cuda.is_current_stream_capturing():
                self.note_graph_replay()             # an eager call in device-step mode really executed a step
            return loss
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
