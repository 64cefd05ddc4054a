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

Input prefetch (`double_buffer=True`): a second set of static inputs with its own captured graph, so that the host->device copy
of batch i+1 (`prefetch`, on a copy stream) runs while step i computes; `step_prefetched()` then replays the graph that reads the
freshly filled set.  Both graphs update the same parameters, optimizer state and device step counter.  This is what the reference's
DataLoader workers + pinned memory do for it; here the copy of a 4 MB (C4) to 13 MB (C3, 8-bit images) batch is 0.15 - 0.5 ms that the
1.5 - 9 ms step would otherwise wait for.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops
from ._lib import lib


class GraphedTrainStep:
    def __init__(self, model, optimizer, example_batch: Sequence[Optional[torch.Tensor]], group=None, warmup: int = 3,
                 double_buffer: bool = False):
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
        # ---- optional second input set + graph (prefetch) ------------------------------------------------------------
        self._sets = [(self.static, self.graph, self.loss)]
        self._cur = 0                       # set the last replay read
        self._filled = None                 # set holding a prefetched batch
        self._copy_stream = None
        self._copy_done = None
        if double_buffer:
            static2 = [None if v is None else v.clone() for v in self.static]
            first = self.static
            self.static = static2           # _eager_step reads self.static
            optimizer.zero_grad(set_to_none=True)
            model._gbuf = None
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, capture_error_mode="thread_local"):
                loss2 = self._eager_step()
            self.static = first
            self._sets.append((static2, g2, loss2.detach()))
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_done = torch.cuda.Event()
        self._read_done = [torch.cuda.Event() for _ in self._sets]      # recorded after the replay that read set i

    def _eager_step(self):
        loss = self.model.training_step(self.static, 0)
        loss.backward()
        if self.group is not None:
            self.model.reduce_gradients(self.group)             # encoder segments were launched during the backward (overlap)
        self.optimizer.step()
        return loss

    def _fill(self, static, batch):
        for dst, src in zip(static, batch):
            if dst is None or src is dst:
                continue
            if src is None or src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError("maven_b200: GraphedTrainStep batches must keep the captured shapes and dtypes "
                                 f"(expected {tuple(dst.shape)} {dst.dtype}, got {None if src is None else (tuple(src.shape), src.dtype)})")
            dst.copy_(src, non_blocking=True)

    def prefetch_buffers(self):
        """The static inputs the NEXT `prefetch` fills (a caller that produces an input on the device, e.g. augment_images(out=...),
        writes it there inside `with torch.cuda.stream(step.copy_stream)` and passes the same tensor in the batch)."""
        if len(self._sets) < 2:
            raise RuntimeError("maven_b200: GraphedTrainStep(double_buffer=True) is needed for prefetch")
        return self._sets[self._cur ^ 1][0]

    @property
    def copy_stream(self):
        return self._copy_stream

    @ops.nvtx_range("graph.prefetch")
    def prefetch(self, batch):
        """Start the upload of the next batch into the input set the running step does not read (copy stream, asynchronous).
        `batch`: the 9-tuple, or a callable `f(static) -> batch` that is run inside the copy stream and may fill inputs on the
        device itself (e.g. an 8-bit image upload + augment_images(out=static[0])), returning those same tensors in the batch.
        Sources are host tensors (pinned, for the copy to be asynchronous) or device tensors that are already complete: the copy
        stream does NOT wait for work queued on the caller's stream -- it would queue behind the running step and lose the overlap."""
        static = self.prefetch_buffers()
        tgt = self._cur ^ 1
        self._copy_stream.wait_event(self._read_done[tgt])       # only the replay that last READ this set: the running step reads the other one
        with torch.cuda.stream(self._copy_stream):
            if callable(batch):
                batch = batch(static)
            self._fill(static, batch)
            self._copy_done.record(self._copy_stream)
        self._filled = tgt

    @ops.nvtx_range("graph.replay")
    def step_prefetched(self) -> torch.Tensor:
        """One training step on the batch handed to the last `prefetch`."""
        if self._filled is None:
            raise RuntimeError("maven_b200: step_prefetched() without a prefetch()")
        self._cur, self._filled = self._filled, None
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._copy_done)
        _, graph, loss = self._sets[self._cur]
        graph.replay()
        self._read_done[self._cur].record(main)
        self.replays += 1
        self.optimizer.note_graph_replay()
        return loss

    @ops.nvtx_range("graph.replay")
    def __call__(self, batch: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
        """One training step on `batch` (None: reuse what is in the static buffers).  Returns the loss (device tensor,
        overwritten by the next call)."""
        if batch is not None:
            self._fill(self.static, batch)
        self.graph.replay()
        self._cur = 0
        self._read_done[0].record(torch.cuda.current_stream(self.device))
        self.replays += 1
        self.optimizer.note_graph_replay()
        return self.loss
