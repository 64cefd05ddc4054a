"""Whole-step CUDA graph: forward + backward + gradient all-reduce + fused RAdam of LightCurveImageCLIP
captured once and replayed per batch.  The reference's Lightning loop (src/models_multimodal.py:312-366 + torch.optim.RAdam)
enqueues ~16.6k kernels per step from Python; the eager path of this package enqueues ~350 from ~40 library calls, which still
costs 2.5-4 ms of host time per step -- more than the device needs for the smaller configurations.  A replay costs one launch.

What changes from step to step inside a graph (whose kernel arguments are frozen at capture) lives on the device:
  * the batch      -> static input buffers, refilled by `__call__` (host or device tensors, async copies);
  * dropout masks  -> every dropout site mixes a device step counter into its hash (mvn_set_step_counter);
  * RAdam's bias corrections / rectification -> computed on the device from that counter (mvn_radam_step_dev).
No host synchronisation happens inside `__call__`; the returned loss is a device tensor (call .item() to read it).
Frozen at capture, like every kernel argument: the batch shapes, the learning rate / betas / weight decay of the optimizer
(the reference trains with a constant-lr RAdam, src/models_multimodal.py:306-310) and which parameters receive gradients.
`step.static[i]` are the device-resident inputs: a caller may fill one itself (e.g. maven_b200.augment.augment_images(...,
out=step.static[0]) from an 8-bit upload) and pass that same tensor in the batch, which skips the copy for it.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops
from ._lib import lib


class GraphedTrainStep:
    def __init__(self, model, optimizer, example_batch: Sequence[Optional[torch.Tensor]], group=None, warmup: int = 3):
        """model: LightCurveImageCLIP in train mode on a CUDA device; optimizer: its FusedRAdam; example_batch: a batch with
        the shapes/dtypes every later batch will have (the reference's 9-tuple); group: data-parallel process group whose
        ranks all-reduce the flat gradient buffer (None = single GPU).  Model/optimizer state is left exactly as it was:
        the eager warm-up steps that prime allocator and lazy initialisation are rolled back before capture."""
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("maven_b200: GraphedTrainStep needs the model on a CUDA device (no CPU fallback)")
        self.model, self.optimizer, self.group, self.device = model, optimizer, group, dev
        self.static = [None if v is None else v.to(dev, copy=True) for v in example_batch]
        L = lib()

        # ---- eager warm-up on a side stream, then roll the state back ------------------------------------------
        snap_model = {k: v.detach().clone() for k, v in model.state_dict().items()}
        snap_opt = optimizer.snapshot()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(max(warmup, 1)):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            model.load_state_dict(snap_model)
        optimizer.restore(snap_opt)
        optimizer.enable_device_step()
        optimizer.zero_grad(set_to_none=True)
        model._gbuf = None

        # ---- capture ----------------------------------------------------------------------------------------------
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.mvn_launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            loss = self._eager_step()
        self.launches_per_replay = int(L.mvn_launch_count() - n0)
        self.loss = loss.detach()
        self.replays = 0

    def _eager_step(self):
        loss = self.model.training_step(self.static, 0)
        loss.backward()
        if self.group is not None:
            self.model.reduce_gradients(self.group)             # encoder segments were launched during the backward (overlap)
        self.optimizer.step()
        return loss

    def __call__(self, batch: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
        """One training step on `batch` (None: reuse what is in the static buffers).  Returns the loss (device tensor,
        overwritten by the next call)."""
        if batch is not None:
            for dst, src in zip(self.static, batch):
                if dst is None or src is dst:
                    continue
                if src is None or src.shape != dst.shape or src.dtype != dst.dtype:
                    raise ValueError("maven_b200: GraphedTrainStep batches must keep the captured shapes and dtypes "
                                     f"(expected {tuple(dst.shape)} {dst.dtype}, got {None if src is None else (tuple(src.shape), src.dtype)})")
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        self.optimizer.note_graph_replay()
        return self.loss
